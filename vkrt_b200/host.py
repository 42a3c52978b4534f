"""ctypes binding over libvkrt_host.so — the C host that mirrors vkrt's VKRT_* API (include/vkrt_host.h).

Python here is only a caller (tests, bench.py); scene preparation, glTF / vkrt.scene ingest, the frame protocol and image
export all live in the C library, which in turn drives libvkrt_cuda.so. Importing never falls back to anything: a missing
library raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import BuildStats, FrameStats, VkrtError, load_library

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libvkrt_host.so")

NAME_LEN = 256

# wire-format dtypes (include/vkrt_shared.h)
SHADER_VERTEX = np.dtype([("position", "<f4", 4), ("texcoord0", "<f4", 2), ("texcoord1", "<f4", 2), ("packedNormal", "<u4"),
                          ("packedTangent", "<u4"), ("packedColor", "<u4"), ("_pad", "<u4")])
VERTEX = np.dtype([("position", "<f4", 4), ("normal", "<f4", 4), ("tangent", "<f4", 4), ("color", "<f4", 4), ("texcoord0", "<f4", 2),
                   ("texcoord1", "<f4", 2)])
MESH_INFO = np.dtype([("position", "<f4", 3), ("vertexBase", "<u4"), ("rotation", "<f4", 3), ("vertexCount", "<u4"), ("scale", "<f4", 3),
                      ("indexBase", "<u4"), ("indexCount", "<u4"), ("materialIndex", "<u4"), ("renderBackfaces", "<u4"),
                      ("lightPdfArea", "<f4"), ("opacity", "<f4"), ("reserved0", "<u4"), ("reserved1", "<u4"), ("reserved2", "<u4")])
MATERIAL_BYTES = 272
EMISSIVE_MESH = np.dtype([("triOffset", "<u4"), ("triCount", "<u4"), ("pmfMesh", "<f4"), ("invTotalArea", "<f4"), ("emission", "<f4", 3),
                          ("reserved0", "<f4")])
EMISSIVE_TRIANGLE = np.dtype([("v0Area", "<f4", 4), ("e1Pad", "<f4", 4), ("e2Pad", "<f4", 4)])
SCENE_DATA_BYTES = 240


class CreateInfo(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("title", C.c_char_p), ("startMaximized", C.c_uint8),
                ("startFullscreen", C.c_uint8), ("headless", C.c_uint8), ("disableSER", C.c_uint8), ("preferredDeviceIndex", C.c_int32),
                ("preferredDeviceName", C.c_char_p), ("rank", C.c_uint32), ("worldSize", C.c_uint32), ("maxPathsInFlight", C.c_uint32),
                ("cudaFlags", C.c_uint32), ("hostOnly", C.c_uint8)]


class PreparedScene(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertexCount", C.c_uint32), ("indices", C.c_void_p), ("indexCount", C.c_uint32),
                ("meshInfos", C.c_void_p), ("world3x4", C.c_void_p), ("geometrySource", C.c_void_p), ("alphaTested", C.c_void_p),
                ("meshCount", C.c_uint32), ("materials", C.c_void_p), ("materialCount", C.c_uint32), ("emissiveMeshes", C.c_void_p),
                ("emissiveMeshCount", C.c_uint32), ("emissiveTriangles", C.c_void_p), ("emissiveTriangleCount", C.c_uint32),
                ("meshAliasQ", C.c_void_p), ("meshAliasIdx", C.c_void_p), ("triAliasQ", C.c_void_p), ("triAliasIdx", C.c_void_p),
                ("sceneData", C.c_void_p)]


class OfflineRenderResult(C.Structure):
    _fields_ = [("seconds", C.c_double), ("deviceSeconds", C.c_double), ("samples", C.c_uint64), ("frames", C.c_uint32),
                ("samplesPerFrame", C.c_uint32), ("samplesPerSecond", C.c_double), ("mpathsPerSecond", C.c_double),
                ("extensionRays", C.c_uint64), ("shadowRays", C.c_uint64)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("target", C.c_float * 3), ("up", C.c_float * 3), ("nearZ", C.c_float), ("farZ", C.c_float),
                ("vfov", C.c_float)]


class SceneSettings(C.Structure):
    _fields_ = [("camera", Camera), ("samplesPerPixel", C.c_uint32), ("rrMaxDepth", C.c_uint32), ("rrMinDepth", C.c_uint32),
                ("toneMappingMode", C.c_uint32), ("renderMode", C.c_uint32), ("spectralSamplingMode", C.c_uint32), ("exposure", C.c_float),
                ("autoExposureEnabled", C.c_uint8), ("autoSPPEnabled", C.c_uint8), ("autoSPPTargetFPS", C.c_uint32),
                ("environmentColor", C.c_float * 3), ("environmentStrength", C.c_float), ("environmentRotation", C.c_float),
                ("environmentTextureIndex", C.c_uint32), ("timeBase", C.c_float), ("timeStep", C.c_float), ("debugMode", C.c_uint32),
                ("misNeeEnabled", C.c_uint32), ("selectionEnabled", C.c_uint32), ("selectedMeshIndex", C.c_uint32)]


class RenderStatus(C.Structure):
    _fields_ = [("framesPerSecond", C.c_uint32), ("averageFrametime", C.c_float), ("frametimes", C.c_float * 128), ("displayTimeMs", C.c_float),
                ("renderTimeMs", C.c_float), ("accumulationFrame", C.c_uint32), ("totalSamples", C.c_uint64), ("renderPhase", C.c_int),
                ("renderDenoiseEnabled", C.c_uint8), ("renderTargetSamples", C.c_uint32), ("displayRenderTimeMs", C.c_float),
                ("displayFrameTimeMs", C.c_float)]


_hostlib = None


def load_host_library() -> C.CDLL:
    global _hostlib
    if _hostlib is not None:
        return _hostlib
    load_library()  # libvkrt_cuda.so first (raises if missing); the host library links against it
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError("vkrt_b200: %s not found; run `make -C %s`" % (HOST_LIB_PATH, _HERE))
    lib = C.CDLL(HOST_LIB_PATH)
    lib.VKRT_lastError.restype = C.c_char_p
    lib.VKRT_lastError.argtypes = [C.c_void_p]
    lib.VKRT_cudaContext.restype = C.c_void_p
    lib.VKRT_cudaContext.argtypes = [C.c_void_p]
    lib.VKRT_destroy.restype = None
    lib.VKRT_destroy.argtypes = [C.c_void_p]
    _hostlib = lib
    return lib


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    n = count * np.dtype(dtype).itemsize
    buf = (C.c_char * n).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


class Host:
    """One VKRT handle."""

    def __init__(self, width=1600, height=900, device=-1, rank=0, world_size=1, max_paths=0, cuda_flags=0, host_only=False):
        self.lib = load_host_library()
        self.h = C.c_void_p()
        self._check(self.lib.VKRT_create(C.byref(self.h)), "VKRT_create")
        ci = CreateInfo()
        self.lib.VKRT_defaultCreateInfo(C.byref(ci))
        ci.width, ci.height, ci.preferredDeviceIndex = width, height, device
        ci.rank, ci.worldSize, ci.maxPathsInFlight, ci.cudaFlags, ci.hostOnly = rank, world_size, max_paths, cuda_flags, 1 if host_only else 0
        rc = self.lib.VKRT_initWithCreateInfo(self.h, C.byref(ci))
        if rc != 0:
            msg = (self.lib.VKRT_lastError(self.h) or b"").decode()
            self.lib.VKRT_destroy(self.h)
            self.h = C.c_void_p()
            raise VkrtError(rc, "VKRT_initWithCreateInfo", msg)

    def close(self):
        if self.h:
            self.lib.VKRT_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise VkrtError(rc, what, (self.lib.VKRT_lastError(self.h) or b"").decode() if self.h else "")

    # ---- app layer ----
    def load_scene(self, path):
        self._check(self.lib.VKRT_appLoadScene(self.h, os.fsencode(path)), "VKRT_appLoadScene")

    def import_mesh(self, path):
        first, count = C.c_uint32(), C.c_uint32()
        self._check(self.lib.VKRT_appImportMesh(self.h, os.fsencode(path), C.byref(first), C.byref(count)), "VKRT_appImportMesh")
        return first.value, count.value

    def generate_soup(self, triangles, seed=0):
        self._check(self.lib.VKRT_appGenerateSoup(self.h, C.c_uint32(triangles), C.c_uint32(seed)), "VKRT_appGenerateSoup")

    def generate_instanced(self, glb, count, seed=0):
        self._check(self.lib.VKRT_appGenerateInstanced(self.h, os.fsencode(glb), C.c_uint32(count), C.c_uint32(seed)), "VKRT_appGenerateInstanced")

    def load_rgb2spec(self, path):
        self._check(self.lib.VKRT_loadRGB2SpecTable(self.h, os.fsencode(path)), "VKRT_loadRGB2SpecTable")

    def offline_render(self, width, height, target_samples, samples_per_frame):
        res = OfflineRenderResult()
        self._check(self.lib.VKRT_appOfflineRender(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(target_samples),
                                                   C.c_uint32(samples_per_frame), C.byref(res)), "VKRT_appOfflineRender")
        return res

    # ---- settings ----
    def set_render_mode(self, mode):
        self._check(self.lib.VKRT_setRenderMode(self.h, C.c_uint32(mode)), "VKRT_setRenderMode")

    def set_spectral_sampling_mode(self, mode):
        self._check(self.lib.VKRT_setSpectralSamplingMode(self.h, C.c_uint32(mode)), "VKRT_setSpectralSamplingMode")

    def set_samples_per_pixel(self, spp):
        self._check(self.lib.VKRT_setSamplesPerPixel(self.h, C.c_uint32(spp)), "VKRT_setSamplesPerPixel")

    def set_path_depth(self, rr_min, rr_max):
        self._check(self.lib.VKRT_setPathDepth(self.h, C.c_uint32(rr_min), C.c_uint32(rr_max)), "VKRT_setPathDepth")

    def set_debug_mode(self, mode):
        self._check(self.lib.VKRT_setDebugMode(self.h, C.c_uint32(mode)), "VKRT_setDebugMode")

    def set_environment_light(self, color, strength):
        c = (C.c_float * 3)(*color)
        self._check(self.lib.VKRT_setEnvironmentLight(self.h, c, C.c_float(strength)), "VKRT_setEnvironmentLight")

    def set_environment_texture(self, pixels_rgba32f):
        """Lat-long environment map, float32 HxWx4 (the reference loads .exr files as LINEAR RGBA32F, controller.c:1300-1312)."""
        px = np.ascontiguousarray(pixels_rgba32f, dtype=np.float32)

        class Upload(C.Structure):
            _fields_ = [("name", C.c_char_p), ("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32),
                        ("colorSpace", C.c_uint32)]
        up = Upload(b"environment", px.ctypes.data, px.shape[1], px.shape[0], 3, 1)
        self._check(self.lib.VKRT_setEnvironmentTextureFromPixels(self.h, C.byref(up)), "VKRT_setEnvironmentTextureFromPixels")

    def camera_set_pose(self, pos, target, up, vfov):
        self._check(self.lib.VKRT_cameraSetPose(self.h, (C.c_float * 3)(*pos), (C.c_float * 3)(*target), (C.c_float * 3)(*up), C.c_float(vfov)),
                    "VKRT_cameraSetPose")

    def scene_settings(self):
        s = SceneSettings()
        self._check(self.lib.VKRT_getSceneSettings(self.h, C.byref(s)), "VKRT_getSceneSettings")
        return s

    def render_status(self):
        s = RenderStatus()
        self._check(self.lib.VKRT_getRenderStatus(self.h, C.byref(s)), "VKRT_getRenderStatus")
        return s

    def mesh_count(self):
        n = C.c_uint32()
        self._check(self.lib.VKRT_getMeshCount(self.h, C.byref(n)), "VKRT_getMeshCount")
        return n.value

    # ---- frame protocol ----
    def start_render(self, width, height, target_samples):
        self._check(self.lib.VKRT_startRender(self.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(target_samples)), "VKRT_startRender")

    def update_scene(self):
        self._check(self.lib.VKRT_beginFrame(self.h), "VKRT_beginFrame")
        self._check(self.lib.VKRT_updateScene(self.h), "VKRT_updateScene")

    def draw(self):
        self._check(self.lib.VKRT_draw(self.h), "VKRT_draw")

    def save_render_image(self, path, denoise=False):
        """VKRT_saveRenderImageEx; denoise=True runs the feature AOVs through Open Image Denoise first (bound at run time, raw image if absent)."""
        settings = (C.c_uint8 * 1)(1 if denoise else 0)   # VKRT_RenderExportSettings { uint8_t denoiseEnabled; }
        self._check(self.lib.VKRT_saveRenderImageEx(self.h, os.fsencode(path), settings), "VKRT_saveRenderImageEx")

    def last_frame_stats(self):
        st = FrameStats()
        self._check(self.lib.VKRT_getLastFrameStats(self.h, C.byref(st)), "VKRT_getLastFrameStats")
        return st

    def build_stats(self):
        st = BuildStats()
        self._check(self.lib.VKRT_getBuildStats(self.h, C.byref(st)), "VKRT_getBuildStats")
        return st

    def cuda_context(self):
        return self.lib.VKRT_cudaContext(self.h)

    # ---- introspection ----
    def prepare_scene(self):
        """The device-format arrays the host hands to vkrt_cuda_set_* (copies)."""
        ps = PreparedScene()
        self._check(self.lib.VKRT_prepareScene(self.h, C.byref(ps)), "VKRT_prepareScene")
        n = ps.meshCount
        return dict(
            vertices=_view(ps.vertices, ps.vertexCount, SHADER_VERTEX), indices=_view(ps.indices, ps.indexCount, np.uint32),
            meshInfos=_view(ps.meshInfos, n, MESH_INFO), world3x4=_view(ps.world3x4, n * 12, np.float32).reshape(n, 3, 4),
            geometrySource=_view(ps.geometrySource, n, np.uint32), alphaTested=_view(ps.alphaTested, n, np.uint8),
            materials=_view(ps.materials, ps.materialCount * MATERIAL_BYTES, np.uint8).reshape(ps.materialCount, MATERIAL_BYTES),
            emissiveMeshes=_view(ps.emissiveMeshes, ps.emissiveMeshCount, EMISSIVE_MESH),
            emissiveTriangles=_view(ps.emissiveTriangles, ps.emissiveTriangleCount, EMISSIVE_TRIANGLE),
            meshAliasQ=_view(ps.meshAliasQ, ps.emissiveMeshCount, np.float32), meshAliasIdx=_view(ps.meshAliasIdx, ps.emissiveMeshCount, np.uint32),
            triAliasQ=_view(ps.triAliasQ, ps.emissiveTriangleCount, np.float32), triAliasIdx=_view(ps.triAliasIdx, ps.emissiveTriangleCount, np.uint32),
            sceneData=_view(ps.sceneData, SCENE_DATA_BYTES, np.uint8))
