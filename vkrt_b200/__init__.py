"""vkrt_b200 — B200-native path-tracing hot path for vkrt behind a C ABI.

This package is a thin ctypes binding over ``libvkrt_cuda.so`` (include/vkrt_cuda.h) and ``libvkrt_host.so``
(the C host that mirrors vkrt's VKRT_* API).  There is no Python or CPU implementation of the renderer here:
if the CUDA library has not been built, importing the binding raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(_HERE, "libvkrt_cuda.so")
HOST_LIB_PATH = os.path.join(_HERE, "libvkrt_host.so")

AOV_ACCUM, AOV_ALBEDO, AOV_NORMAL, AOV_OUTPUT, AOV_HITID_CENTER, AOV_HITID_S0, AOV_HIT_TUV = range(7)
FLAG_COUNT_RAYS = 1
FLAG_NO_MATERIAL_SORT = 2
FLAG_STAGE_TIMING = 4
FLAG_FORCE_TWO_LEVEL = 8
FLAG_FORCE_FLAT = 16
FLAG_DEEP_STACK = 32
FLAG_ENV_IMPORTANCE = 64   # extension: environment-map importance sampling (not in the reference; include/vkrt_cuda.h)

VKRT_SUCCESS = 0
_ERRORS = {0: "SUCCESS", -1: "INVALID_ARGUMENT", -2: "OPERATION_FAILED", -3: "OUT_OF_MEMORY", -4: "DEVICE_LOST",
           -5: "INITIALIZATION_FAILED"}


class VkrtError(RuntimeError):
    def __init__(self, code, what, detail=""):
        super().__init__("%s: VKRT_ERROR_%s (%d) %s" % (what, _ERRORS.get(code, "?"), code, detail))
        self.code = code


class CreateInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_uint32), ("worldSize", C.c_uint32), ("tileWidth", C.c_uint32),
                ("tileHeight", C.c_uint32), ("maxPathsInFlight", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class BuildStats(C.Structure):
    _fields_ = [("buildMs", C.c_float), ("blasMs", C.c_float), ("tlasMs", C.c_float), ("uniqueGeometries", C.c_uint32),
                ("instanceCount", C.c_uint32), ("triangleCount", C.c_uint64), ("instancedTriangleCount", C.c_uint64),
                ("bvh8NodeCount", C.c_uint64), ("accelBytes", C.c_uint64), ("flat", C.c_uint32), ("plocHierarchies", C.c_uint32)]


class FrameStats(C.Structure):
    _fields_ = [("frameMs", C.c_float), ("traceMs", C.c_float), ("shadeMs", C.c_float), ("kernelLaunches", C.c_uint32),
                ("traceLaunches", C.c_uint32), ("shadeLaunches", C.c_uint32), ("paths", C.c_uint64), ("extensionRays", C.c_uint64), ("shadowRays", C.c_uint64), ("nodesVisited", C.c_uint64),
                ("trianglesTested", C.c_uint64), ("instancesEntered", C.c_uint64), ("shadeKernelMs", C.c_float), ("reserved", C.c_uint32)]


class RGB2SpecInfo(C.Structure):
    _fields_ = [("res", C.c_uint32), ("scaleOffset", C.c_uint32), ("dataOffset", C.c_uint32)]


class TextureDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32),
                ("colorSpace", C.c_uint32)]


EXPORTS = [
    "vkrt_cuda_create", "vkrt_cuda_destroy", "vkrt_cuda_last_error", "vkrt_cuda_version", "vkrt_cuda_set_geometry",
    "vkrt_cuda_set_instances", "vkrt_cuda_set_materials", "vkrt_cuda_set_lights", "vkrt_cuda_set_textures",
    "vkrt_cuda_set_rgb2spec", "vkrt_cuda_build_accel", "vkrt_cuda_resize", "vkrt_cuda_reset_accumulation",
    "vkrt_cuda_render_frame", "vkrt_cuda_render_frame_async", "vkrt_cuda_sync", "vkrt_cuda_timer_begin", "vkrt_cuda_timer_end", "vkrt_cuda_nccl_unique_id",
    "vkrt_cuda_comm_init", "vkrt_cuda_gather", "vkrt_cuda_local_film", "vkrt_cuda_import_gathered",
    "vkrt_cuda_max_local_pixels", "vkrt_cuda_read_aov", "vkrt_cuda_read_accum_samples", "vkrt_cuda_trace_primary", "vkrt_cuda_trace_rays",
    "vkrt_cuda_eval_closures", "vkrt_cuda_invalidate_accel", "vkrt_cuda_gather_aovs",
]

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Loads libvkrt_cuda.so. Raises (never falls back) if it is missing: build it with ``make -C vkrt_b200`` or
    ``python -c 'import __graft_entry__ as g; g.build()'``."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("VKRT_CUDA_LIB") or CUDA_LIB_PATH  # VKRT_CUDA_LIB: A/B-test an alternative build of the same ABI
    if not os.path.exists(p):
        raise ImportError("vkrt_b200: %s not found. The CUDA extension is required (there is no CPU path); run `make -C %s`." % (p, _HERE))
    lib = C.CDLL(p)
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError if the ABI is incomplete
    lib.vkrt_cuda_last_error.restype = C.c_char_p
    lib.vkrt_cuda_last_error.argtypes = [C.c_void_p]
    lib.vkrt_cuda_version.restype = C.c_char_p
    lib.vkrt_cuda_max_local_pixels.restype = C.c_uint64
    lib.vkrt_cuda_max_local_pixels.argtypes = [C.c_void_p]
    lib.vkrt_cuda_destroy.argtypes = [C.c_void_p]
    lib.vkrt_cuda_destroy.restype = None
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class CudaContext:
    """One vkrt_cuda_ctx. Arrays are numpy arrays in the wire format of include/vkrt_shared.h."""

    def __init__(self, device=-1, rank=0, world_size=1, tile=(32, 32), max_paths=0, flags=0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        info = CreateInfo(device, rank, world_size, tile[0], tile[1], max_paths, flags, 0)
        rc = self.lib.vkrt_cuda_create(C.byref(info), C.byref(self.ctx))
        if rc != 0:
            raise VkrtError(rc, "vkrt_cuda_create", "(is a CUDA device visible?)")
        self.width = self.height = 0
        self._keep = []

    def close(self):
        if self.ctx:
            self.lib.vkrt_cuda_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise VkrtError(rc, what, (self.lib.vkrt_cuda_last_error(self.ctx) or b"").decode())

    def set_geometry(self, vertices, indices):
        v, i = np.ascontiguousarray(vertices), np.ascontiguousarray(indices, dtype=np.uint32)
        self._check(self.lib.vkrt_cuda_set_geometry(self.ctx, _ptr(v), C.c_uint32(len(v)), _ptr(i), C.c_uint32(len(i))), "set_geometry")

    def set_instances(self, mesh_infos, world3x4, geometry_source=None, alpha_tested=None):
        mi = np.ascontiguousarray(mesh_infos)
        w = np.ascontiguousarray(world3x4, dtype=np.float32)
        gs = np.ascontiguousarray(geometry_source, dtype=np.uint32) if geometry_source is not None else None
        al = np.ascontiguousarray(alpha_tested, dtype=np.uint8) if alpha_tested is not None else None
        self._check(self.lib.vkrt_cuda_set_instances(self.ctx, _ptr(mi), _ptr(w), _ptr(gs), _ptr(al), C.c_uint32(len(mi))), "set_instances")

    def set_materials(self, materials):
        m = np.ascontiguousarray(materials)
        self._check(self.lib.vkrt_cuda_set_materials(self.ctx, _ptr(m), C.c_uint32(len(m))), "set_materials")

    def set_lights(self, meshes, mesh_count, triangles, triangle_count, mesh_q, mesh_idx, tri_q, tri_idx):
        arrs = [np.ascontiguousarray(a) for a in (meshes, triangles, mesh_q, mesh_idx, tri_q, tri_idx)]
        self._check(self.lib.vkrt_cuda_set_lights(self.ctx, _ptr(arrs[0]), C.c_uint32(mesh_count), _ptr(arrs[1]), C.c_uint32(triangle_count),
                                                  _ptr(arrs[2]), _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5])), "set_lights")

    def set_textures(self, textures):
        arr = (TextureDesc * max(len(textures), 1))()
        for i, t in enumerate(textures):
            px = np.ascontiguousarray(t["pixels"])
            self._keep.append(px)
            arr[i] = TextureDesc(px.ctypes.data, t["width"], t["height"], t["format"], t["colorSpace"])
        self._check(self.lib.vkrt_cuda_set_textures(self.ctx, arr, C.c_uint32(len(textures))), "set_textures")

    def set_rgb2spec(self, payload, res, scale_offset=0, data_offset=None):
        p = np.ascontiguousarray(payload, dtype=np.float32)
        info = RGB2SpecInfo(res, scale_offset, res if data_offset is None else data_offset)
        self._check(self.lib.vkrt_cuda_set_rgb2spec(self.ctx, _ptr(p), C.c_uint32(len(p)), info), "set_rgb2spec")

    def build_accel(self):
        st = BuildStats()
        self._check(self.lib.vkrt_cuda_build_accel(self.ctx, C.byref(st)), "build_accel")
        return st

    def resize(self, width, height):
        self._check(self.lib.vkrt_cuda_resize(self.ctx, C.c_uint32(width), C.c_uint32(height)), "resize")
        self.width, self.height = width, height

    def reset_accumulation(self):
        self._check(self.lib.vkrt_cuda_reset_accumulation(self.ctx), "reset_accumulation")

    def render_frame(self, scene_data):
        sd = np.ascontiguousarray(scene_data)
        st = FrameStats()
        self._check(self.lib.vkrt_cuda_render_frame(self.ctx, _ptr(sd), C.byref(st)), "render_frame")
        return st

    def render_frame_async(self, scene_data):
        sd = np.ascontiguousarray(scene_data)
        self._check(self.lib.vkrt_cuda_render_frame_async(self.ctx, _ptr(sd)), "render_frame_async")

    def sync(self):
        self._check(self.lib.vkrt_cuda_sync(self.ctx), "sync")

    def trace_primary(self, scene_data):
        sd = np.ascontiguousarray(scene_data)
        self._check(self.lib.vkrt_cuda_trace_primary(self.ctx, _ptr(sd)), "trace_primary")

    def read_aov(self, which, out=None):
        shapes = {AOV_ACCUM: (np.float32, 4), AOV_ALBEDO: (np.float16, 4), AOV_NORMAL: (np.float16, 4), AOV_OUTPUT: (np.uint16, 4),
                  AOV_HITID_CENTER: (np.uint32, 2), AOV_HITID_S0: (np.uint32, 2), AOV_HIT_TUV: (np.float32, 3)}
        dt, nc = shapes[which]
        if out is None:
            out = np.zeros((self.height, self.width, nc), dtype=dt)
        self._check(self.lib.vkrt_cuda_read_aov(self.ctx, C.c_int(which), _ptr(out), C.c_size_t(out.nbytes)), "read_aov")
        return out

    def trace_rays(self, rays, any_hit=False):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros((len(r), 5), dtype=np.uint32)
        ms = C.c_float()
        self._check(self.lib.vkrt_cuda_trace_rays(self.ctx, _ptr(r), C.c_uint32(len(r)), C.c_int(1 if any_hit else 0), _ptr(hits), C.byref(ms)), "trace_rays")
        return hits, ms.value

    def eval_closures(self, queries, result_dtype):
        """Test entry (include/vkrt_closure.h): `queries` is a structured array of vkrt_closure_query records."""
        q = np.ascontiguousarray(queries)
        out = np.zeros(len(q), dtype=result_dtype)
        self._check(self.lib.vkrt_cuda_eval_closures(self.ctx, _ptr(q), C.c_uint32(len(q)), _ptr(out)), "eval_closures")
        return out

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.vkrt_cuda_comm_init(self.ctx, buf), "comm_init")

    def gather(self, aov_mask=0xF):
        ms = C.c_float()
        self._check(self.lib.vkrt_cuda_gather_aovs(self.ctx, C.c_uint32(aov_mask), C.byref(ms)), "gather")
        return ms.value

    def local_film(self, which):
        p, nbytes, npx = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.vkrt_cuda_local_film(self.ctx, C.c_int(which), C.byref(p), C.byref(nbytes), C.byref(npx)), "local_film")
        return p.value, nbytes.value, npx.value

    def import_gathered(self, which, device_ptr):
        self._check(self.lib.vkrt_cuda_import_gathered(self.ctx, C.c_int(which), C.c_void_p(device_ptr)), "import_gathered")

    def max_local_pixels(self):
        return int(self.lib.vkrt_cuda_max_local_pixels(self.ctx))


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.vkrt_cuda_nccl_unique_id(buf)
    if rc != 0:
        raise VkrtError(rc, "vkrt_cuda_nccl_unique_id")
    return buf.raw
