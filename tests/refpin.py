"""Loader for oracle/_ref: the reference's OWN sources compiled for the CPU (test infrastructure, see oracle/Makefile `ref`).

  libvkrt_refshade.so  src/shaders/**/*.slang, transliterated to C++ by oracle/ref_slang/slang2cpp.py; exports oracle_* + refshade_*
  libvkrt_refhost.so   src/core/{utility/packing,scene/*,api/{mesh,settings}}.c behind a null Vulkan device; exports refhost_*

The libraries are built where /root/reference exists (this container) and travel to the GPU box prebuilt; a machine with neither
skips the tests that need them."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H

ROOT = H.ROOT
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REFERENCE = os.environ.get("VKRT_REFERENCE", "/root/reference")


def _ensure_built():
    if os.path.isdir(os.path.join(REFERENCE, "src", "shaders")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "REF=" + REFERENCE])


_cache = {}


def _load(name):
    if name not in _cache:
        if not _cache.get("_built"):
            _ensure_built()
            _cache["_built"] = True
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            pytest.skip("oracle/_ref/%s is not built and there is no reference checkout at %s" % (name, REFERENCE))
        _cache[name] = C.CDLL(path)
    return _cache[name]


def refshade_lib():
    lib = _load("libvkrt_refshade.so")
    lib.refshade_version.restype = C.c_char_p
    lib.refshade_rand.restype = C.c_float
    lib.refshade_wavelength_unit.restype = C.c_float
    lib.refshade_hash.restype = C.c_uint32
    lib.refshade_init_pixel_seed.restype = C.c_uint32
    lib.refshade_reverse_bits.restype = C.c_uint32
    lib.refshade_rgb2spec_eval.restype = C.c_float
    lib.refshade_rgb2spec_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
    lib.refshade_power_heuristic.restype = C.c_float
    lib.refshade_power_heuristic.argtypes = [C.c_float, C.c_float]
    lib.refshade_dispersive_ior.restype = C.c_float
    lib.refshade_dispersive_ior.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.refshade_spectral_xyz.argtypes = [C.c_float, C.c_void_p]
    lib.refshade_primary_ray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    lib.refshade_sample_alias.restype = C.c_uint32
    lib.refshade_sample_alias.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    return lib


def refhost_lib():
    lib = _load("libvkrt_refhost.so")
    lib.refhost_create.restype = C.c_void_p
    lib.refhost_create.argtypes = [C.c_uint32, C.c_uint32]
    lib.refhost_destroy.argtypes = [C.c_void_p]
    lib.refhost_read_light_buffer.restype = C.c_int64
    lib.refhost_read_light_buffer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
    for name in ("refhost_add_material", "refhost_set_material", "refhost_get_material", "refhost_add_mesh", "refhost_set_mesh_transform",
                 "refhost_set_mesh_transform_matrix", "refhost_get_mesh", "refhost_rebuild_lights", "refhost_set_camera", "refhost_get_scene_data",
                 "refhost_set_path_depth", "refhost_set_samples_per_pixel", "refhost_set_render_mode", "refhost_set_spectral_sampling_mode",
                 "refhost_set_tone_mapping_mode", "refhost_set_exposure", "refhost_set_environment_light", "refhost_set_mis_nee_enabled"):
        getattr(lib, name).argtypes = None
    lib.refhost_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float]
    lib.refhost_set_exposure.argtypes = [C.c_void_p, C.c_float]
    lib.refhost_set_environment_light.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
    return lib


class RefShadeBackend(H.OracleBackend):
    """The oracle's scene container, traversal and texture sampler under the REFERENCE'S shaders: same Backend interface as
    OracleBackend / CudaBackend, render_frame runs the reference's raygen entry points."""

    def __init__(self, threads=0):
        H.Backend.__init__(self, refshade_lib())
        self.check(self.lib.oracle_create(C.byref(self.ctx)), "create")
        self.lib.oracle_set_threads(self.ctx, threads)

    def render_frame(self, sd):
        sd = np.ascontiguousarray(sd)
        self.check(self.lib.refshade_render_frame(self.ctx, H._ptr(sd)), "refshade_render_frame")

    def render_rows(self, sd, row_begin, row_end):
        sd = np.ascontiguousarray(sd)
        self.check(self.lib.refshade_render_frame_rows(self.ctx, H._ptr(sd), C.c_uint32(row_begin), C.c_uint32(row_end)), "refshade_render_frame_rows")
        return 0, 0


# ---- closure records (include/vkrt_closure.h) ------------------------------------------------------------------------------------------
CLOSURE_QUERY = np.dtype([("material", H.hr.MATERIAL), ("wo", "<f4", 3), ("frontFace", "<u4"), ("wi", "<f4", 3), ("rng", "<u4"),
                          ("wavelengths", "<f4", 4), ("mode", "<u4"), ("reserved", "<u4", 3)])
CLOSURE_RESULT = np.dtype([("evalValue", "<f4", 4), ("evalPdf", "<f4", 4), ("sampleWi", "<f4", 3), ("sampleFlags", "<u4"),
                           ("sampleWeight", "<f4", 4), ("samplePdf", "<f4", 4), ("rngAfter", "<u4"), ("reserved", "<u4", 3)])
assert CLOSURE_QUERY.itemsize == 336 and CLOSURE_RESULT.itemsize == 96


def random_closure_queries(n, seed, mode, lobes="all"):
    """Randomised sanitised materials that switch every lobe of the layered closure on: sheen, clearcoat, subsurface, Oren-Nayar
    (diffuseRoughness), anisotropic GGX, specular tint, metals (eta / k), rough transmission with dispersion (abbe) and absorption."""
    rng = np.random.default_rng(seed)
    q = np.zeros(n, CLOSURE_QUERY)
    m = q["material"]
    base = H.hr.default_material()
    for name in base.dtype.names:
        m[name] = base[name]
    u = lambda *shape: rng.random(shape, dtype=np.float32)  # noqa: E731
    on = lambda p: rng.random(n) < p                          # noqa: E731
    m["baseColor"] = u(n, 3)
    m["roughness"] = np.where(on(0.15), u(n) * 0.05, u(n)).astype(np.float32)
    m["metallic"] = np.where(on(0.3), u(n), 0.0).astype(np.float32)
    m["metallic"][on(0.1)] = 1.0
    m["eta"] = np.where(on(0.3)[:, None], 0.1 + 3.0 * u(n, 3), 0.0).astype(np.float32)
    m["k"] = np.where(m["eta"].sum(axis=1, keepdims=True) > 0, 4.0 * u(n, 3), 0.0).astype(np.float32)
    m["anisotropic"] = np.where(on(0.5), u(n), 0.0).astype(np.float32)
    m["specular"] = u(n)
    m["specularTint"] = np.where(on(0.5), u(n), 0.0).astype(np.float32)
    m["sheenTintWeight"] = np.concatenate([u(n, 3), np.where(on(0.5), u(n), 0.0).astype(np.float32)[:, None]], axis=1)
    m["sheenRoughness"] = u(n)
    m["clearcoat"] = np.where(on(0.5), u(n), 0.0).astype(np.float32)
    m["clearcoatGloss"] = u(n)
    m["diffuseRoughness"] = np.where(on(0.5), u(n), 0.0).astype(np.float32)
    m["subsurface"] = np.where(on(0.4), u(n), 0.0).astype(np.float32)
    transmissive = on(0.35)
    m["transmission"] = np.where(transmissive, np.where(on(0.5), 1.0, u(n)), 0.0).astype(np.float32)
    m["ior"] = (1.0 + 1.5 * u(n)).astype(np.float32)
    m["abbeNumber"] = np.where(transmissive & on(0.5), 20.0 + 60.0 * u(n), 0.0).astype(np.float32)
    m["absorptionCoefficient"] = np.where(transmissive & on(0.5), 5.0 * u(n), 0.0).astype(np.float32)
    m["attenuationColor"] = u(n, 3)
    if lobes == "opaque":
        m["transmission"] = 0.0
    for i in range(n):
        m[i] = H.hr.sanitize_material(m[i])

    def direction(upper):
        z = rng.random(n) * 2.0 - 1.0
        if upper is not None:
            z = np.where(upper, np.abs(z), z)
        phi = rng.random(n) * 2.0 * np.pi
        r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
        return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    q["wo"] = direction(on(0.9))          # mostly above the surface, some grazing / below
    q["wi"] = direction(~transmissive)    # reflections above; transmissive materials also get directions below
    q["frontFace"] = np.where(on(0.8), 1, 0)
    q["rng"] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    unit = u(n)
    q["wavelengths"] = (360.0 + np.mod(unit[:, None] + np.array([0.0, 0.25, 0.5, 0.75], np.float32), 1.0) * 470.0).astype(np.float32)
    q["mode"] = mode
    return q


def eval_closures(lib, fn_name, ctx, queries):
    out = np.zeros(len(queries), CLOSURE_RESULT)
    q = np.ascontiguousarray(queries)
    rc = getattr(lib, fn_name)(ctx, q.ctypes.data_as(C.c_void_p), C.c_uint32(len(q)), out.ctypes.data_as(C.c_void_p))
    assert rc == 0, (fn_name, rc)
    return out
