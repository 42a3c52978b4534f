"""Shared test harness: drives the CPU oracle (oracle/_build/liboracle.so) and the product's C ABI
(vkrt_b200/libvkrt_cuda.so) through the same Python interface, from the same prepared scene arrays."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import host_ref as hr  # noqa: E402  (oracle-side host logic; tests only)

AOV_ACCUM, AOV_ALBEDO, AOV_NORMAL, AOV_OUTPUT, AOV_HITID_CENTER, AOV_HITID_S0, AOV_HIT_TUV = range(7)


class RGB2SpecInfo(C.Structure):
    _fields_ = [("res", C.c_uint32), ("scaleOffset", C.c_uint32), ("dataOffset", C.c_uint32)]


class TextureDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32),
                ("colorSpace", C.c_uint32)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build_oracle():
    out = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return out


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.oracle_rand.restype = C.c_float
        lib.oracle_f16_to_f32.restype = C.c_float
        lib.oracle_f16_to_f32.argtypes = [C.c_uint16]
        lib.oracle_f32_to_f16.restype = C.c_uint16
        lib.oracle_f32_to_f16.argtypes = [C.c_float]
        lib.oracle_rgb2spec_eval.restype = C.c_float
        lib.oracle_rgb2spec_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        lib.oracle_hash.restype = C.c_uint32
        lib.oracle_init_pixel_seed.restype = C.c_uint32
        lib.oracle_reverse_bits.restype = C.c_uint32
        lib.oracle_srgb_lut.restype = C.POINTER(C.c_float)
        lib.oracle_srgb_lut.argtypes = [C.c_void_p]
        lib.oracle_spectral_xyz.argtypes = [C.c_float, C.c_void_p]
        _oracle_lib = lib
    return _oracle_lib


def load_rgb2spec(path=None):
    """Returns (payload float32 array, RGB2SpecInfo) from a srgb.coeff file (core/scene/rgb2spec.c:17-59)."""
    path = path or os.path.join(ROOT, "assets", "rgb2spec", "srgb.coeff")
    raw = open(path, "rb").read()
    assert raw[:4] == b"SPEC"
    res = int(np.frombuffer(raw, "<u4", 1, 4)[0])
    n = res + 9 * res ** 3
    assert len(raw) == 8 + 4 * n, "srgb.coeff size mismatch"
    payload = np.frombuffer(raw, "<f4", n, 8).copy()
    return payload, RGB2SpecInfo(res, 0, res)


class Backend:
    """Common driver over a library exporting <prefix>_set_geometry ... <prefix>_read_aov."""
    prefix = ""

    def __init__(self, lib):
        self.lib = lib
        self.ctx = C.c_void_p()
        self.width = self.height = 0
        self._keep = []

    def f(self, name):
        return getattr(self.lib, self.prefix + name)

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s%s failed: %d %s" % (self.prefix, what, rc, self.last_error()))

    def last_error(self):
        return ""

    def upload(self, prep, rgb2spec=None):
        v, idx = np.ascontiguousarray(prep["vertices"]), np.ascontiguousarray(prep["indices"])
        self.check(self.f("set_geometry")(self.ctx, _ptr(v), C.c_uint32(len(v)), _ptr(idx), C.c_uint32(len(idx))), "set_geometry")
        infos = np.ascontiguousarray(prep["meshInfos"])
        world = np.ascontiguousarray(prep["world3x4"], dtype=np.float32)
        gs = np.ascontiguousarray(prep["geometrySource"], dtype=np.uint32)
        al = np.ascontiguousarray(prep["alphaTested"], dtype=np.uint8)
        self.check(self.f("set_instances")(self.ctx, _ptr(infos), _ptr(world), _ptr(gs), _ptr(al), C.c_uint32(len(infos))), "set_instances")
        mats = np.ascontiguousarray(prep["materials"])
        self.check(self.f("set_materials")(self.ctx, _ptr(mats), C.c_uint32(len(mats))), "set_materials")
        L = prep["lights"]
        self.check(self.f("set_lights")(self.ctx, _ptr(np.ascontiguousarray(L["meshes"])), C.c_uint32(L["meshCount"]),
                                        _ptr(np.ascontiguousarray(L["triangles"])), C.c_uint32(L["triangleCount"]),
                                        _ptr(np.ascontiguousarray(L["meshAliasQ"])), _ptr(np.ascontiguousarray(L["meshAliasIdx"])),
                                        _ptr(np.ascontiguousarray(L["triAliasQ"])), _ptr(np.ascontiguousarray(L["triAliasIdx"]))), "set_lights")
        texs = prep.get("textures") or []
        if texs:
            arr = (TextureDesc * len(texs))()
            for i, t in enumerate(texs):
                px = np.ascontiguousarray(t["pixels"])
                self._keep.append(px)
                arr[i] = TextureDesc(px.ctypes.data, t["width"], t["height"], t["format"], t["colorSpace"])
            self.check(self.f("set_textures")(self.ctx, arr, C.c_uint32(len(texs))), "set_textures")
        if rgb2spec is not None:
            payload, info = rgb2spec
            self.check(self.f("set_rgb2spec")(self.ctx, _ptr(payload), C.c_uint32(len(payload)), info), "set_rgb2spec")
        self.build_accel()

    def resize(self, w, h):
        self.width, self.height = w, h
        self.check(self.f("resize")(self.ctx, C.c_uint32(w), C.c_uint32(h)), "resize")

    def reset(self):
        self.check(self.f("reset_accumulation")(self.ctx), "reset_accumulation")

    def render(self, scene_data, frames=1, first_frame=0):
        sd = scene_data.copy()
        for k in range(frames):
            sd["frameNumber"] = first_frame + k
            self.render_frame(sd)
        return sd

    def read(self, which):
        n = self.width * self.height
        shapes = {AOV_ACCUM: (np.float32, 4), AOV_ALBEDO: (np.float16, 4), AOV_NORMAL: (np.float16, 4),
                  AOV_OUTPUT: (np.uint16, 4), AOV_HITID_CENTER: (np.uint32, 2), AOV_HITID_S0: (np.uint32, 2),
                  AOV_HIT_TUV: (np.float32, 3)}
        dt, nc = shapes[which]
        out = np.zeros((self.height, self.width, nc), dtype=dt)
        self.check(self.f("read_aov")(self.ctx, C.c_int(which), _ptr(out), C.c_size_t(out.nbytes)), "read_aov")
        return out

    def trace_primary(self, scene_data):
        sd = np.ascontiguousarray(scene_data)
        self.check(self.f("trace_primary")(self.ctx, _ptr(sd)), "trace_primary")


class OracleBackend(Backend):
    prefix = "oracle_"

    def __init__(self, brute_force=False, threads=0):
        super().__init__(oracle_lib())
        self.check(self.lib.oracle_create(C.byref(self.ctx)), "create")
        self.lib.oracle_set_brute_force(self.ctx, 1 if brute_force else 0)
        self.lib.oracle_set_threads(self.ctx, threads)

    def __del__(self):
        if self.ctx:
            self.lib.oracle_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def build_accel(self):
        self.check(self.lib.oracle_build_accel(self.ctx), "build_accel")

    def render_frame(self, sd):
        sd = np.ascontiguousarray(sd)
        self.check(self.lib.oracle_render_frame(self.ctx, _ptr(sd)), "render_frame")

    def render_rows(self, sd, row_begin, row_end):
        sd = np.ascontiguousarray(sd)
        rays = (C.c_uint64 * 2)()
        self.check(self.lib.oracle_render_frame_rows(self.ctx, _ptr(sd), C.c_uint32(row_begin), C.c_uint32(row_end), rays), "render_rows")
        return int(rays[0]), int(rays[1])

    def trace_rays(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 5), dtype=np.uint32)
        self.check(self.lib.oracle_trace_rays(self.ctx, _ptr(rays), C.c_uint32(len(rays)), C.c_int(1 if any_hit else 0), _ptr(hits)), "trace_rays")
        return hits


def write_png(path, rgb8):
    """Tiny dependency-free PNG writer for debugging renders (uint8 HxWx3)."""
    import struct
    import zlib
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))
    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


class CudaBackend(Backend):
    """The product: vkrt_cuda_* through ctypes (vkrt_b200.CudaContext)."""
    prefix = "vkrt_cuda_"

    def __init__(self, **kw):
        import vkrt_b200
        self.vk = vkrt_b200
        self.cc = vkrt_b200.CudaContext(**kw)
        super().__init__(self.cc.lib)
        self.ctx = self.cc.ctx
        self.build_stats = None
        self.frame_stats = []

    def last_error(self):
        return (self.lib.vkrt_cuda_last_error(self.ctx) or b"").decode()

    def build_accel(self):
        self.build_stats = self.cc.build_accel()

    def render_frame(self, sd):
        self.frame_stats.append(self.cc.render_frame(sd))

    def trace_rays(self, rays, any_hit=False):
        return self.cc.trace_rays(rays, any_hit)[0]

    def close(self):
        self.cc.close()
        self.ctx = None


def compare_images(a, b, name=""):
    """Robust image comparison for stochastic renders that share the RNG stream but differ in transcendental rounding."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    scale = np.maximum(np.abs(a), np.abs(b)) + 1e-3
    rel = d / scale
    return dict(name=name, rmse=float(np.sqrt(np.mean(d ** 2))), mean_abs=float(d.mean()), max_abs=float(d.max()),
                frac_rel_gt_1e3=float((rel > 1e-3).mean()), frac_rel_gt_1e2=float((rel > 1e-2).mean()),
                mean_a=float(a.mean()), mean_b=float(b.mean()))
