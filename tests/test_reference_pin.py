"""The reference-held pin of the oracle and of the C host (VERDICT r01 items 2 and "parity unpinned"):

 * oracle/_ref/libvkrt_refshade.so = the reference's own src/shaders/**/*.slang compiled for the CPU. The oracle (oracle/oracle.cpp,
   shading.h), a hand restatement of those shaders, must reproduce it BIT FOR BIT: known-answer functions, per-closure evaluation and
   sampling over randomised materials with every lobe switched on, and whole frames in the three render modes (both sides use the
   oracle's stand-ins for what the Vulkan driver supplies upstream: traversal, texture sampler, image formats).
 * oracle/_ref/libvkrt_refhost.so = the reference's own host C sources behind a null device. The product's C host
   (vkrt_b200/host/*.c through libvkrt_host.so) must produce byte-identical packed vertices, struct layouts, transforms, sanitised
   materials, camera matrices and light tables.
No GPU is involved; the GPU-vs-reference tests are in tests/test_gpu_reference.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H
import refpin
import scenes

hr = H.hr


# ======================================================================================================================
# shaders: known answers
# ======================================================================================================================
def test_integer_and_sampling_kats_match_the_reference_shaders():
    ref, orc = refpin.refshade_lib(), H.oracle_lib()
    rng = np.random.default_rng(1)
    vals = np.concatenate([[0, 1, 2, 0xFFFFFFFF, 0x80000000, 0x9E3779B9], rng.integers(0, 2 ** 32, 4096, dtype=np.uint64)]).astype(np.uint32)
    for v in vals[:512]:
        v = int(v)
        assert ref.refshade_hash(C.c_uint32(v)) == orc.oracle_hash(C.c_uint32(v))
        assert ref.refshade_reverse_bits(C.c_uint32(v)) == orc.oracle_reverse_bits(C.c_uint32(v))
        a, b = C.c_uint32(v), C.c_uint32(v)
        ra, rb = ref.refshade_rand(C.byref(a)), orc.oracle_rand(C.byref(b))
        assert ra == rb and a.value == b.value and 0.0 <= ra < 1.0
    for x, y, f, s in rng.integers(0, 5000, (256, 4)):
        assert ref.refshade_init_pixel_seed(int(x), int(y), C.c_uint32(int(f)), C.c_uint32(int(s))) == \
            orc.oracle_init_pixel_seed(int(x), int(y), C.c_uint32(int(f)), C.c_uint32(int(s)))
    # the SURVEY A.5 known answers hold for the reference's own text
    assert ref.refshade_hash(C.c_uint32(0)) == 0 and ref.refshade_reverse_bits(C.c_uint32(1)) == 0x80000000


def test_unpack_and_colour_functions_match_the_reference_shaders():
    ref, orc = refpin.refshade_lib(), H.oracle_lib()
    rng = np.random.default_rng(2)
    a3, b3, a4, b4 = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32)
    for p in np.concatenate([[0, 0xFFFFFFFF, 0x7FFF7FFF, 0x80008000, 0x80000000], rng.integers(0, 2 ** 32, 2000, dtype=np.uint64)]).astype(np.uint32):
        ref.refshade_unpack_normal(C.c_uint32(int(p)), a3.ctypes.data_as(C.c_void_p))
        orc.oracle_unpack_normal(C.c_uint32(int(p)), b3.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a3.view(np.uint32), b3.view(np.uint32)), hex(int(p))
        ref.refshade_unpack_tangent(C.c_uint32(int(p)), a4.ctypes.data_as(C.c_void_p))
        orc.oracle_unpack_tangent(C.c_uint32(int(p)), b4.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a4.view(np.uint32), b4.view(np.uint32)), hex(int(p))
    for lam in np.linspace(360.0, 830.0, 471, dtype=np.float32):
        ref.refshade_spectral_xyz(C.c_float(lam), a3.ctypes.data_as(C.c_void_p))
        orc.oracle_spectral_xyz(C.c_float(lam), b3.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a3.view(np.uint32), b3.view(np.uint32)), lam
    for xyz in rng.random((256, 3), dtype=np.float32) * 2.0:
        xyz = np.ascontiguousarray(xyz)
        ref.refshade_xyz_to_srgb(xyz.ctypes.data_as(C.c_void_p), a3.ctypes.data_as(C.c_void_p))
        orc.oracle_xyz_to_srgb(xyz.ctypes.data_as(C.c_void_p), b3.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a3.view(np.uint32), b3.view(np.uint32))


def test_rgb2spec_lookup_matches_the_reference_shaders():
    ref = refpin.refshade_lib()
    o, r = H.OracleBackend(), refpin.RefShadeBackend()
    payload, info = scenes.rgb2spec()
    for b in (o, r):
        b.check(b.f("set_rgb2spec")(b.ctx, payload.ctypes.data_as(C.c_void_p), C.c_uint32(len(payload)), info), "set_rgb2spec")
    rng = np.random.default_rng(3)
    colours = np.concatenate([rng.random((400, 3), dtype=np.float32), rng.random((100, 3), dtype=np.float32) * 30.0, np.eye(3, dtype=np.float32), [[0, 0, 0], [1, 1, 1], [1e-9, 0, 0], [0.5, 0.5, 0.5]]]).astype(np.float32)
    for rgb in colours:
        rgb = np.ascontiguousarray(rgb)
        for lam in (360.0, 455.5, 550.0, 700.25, 830.0):
            a = ref.refshade_rgb2spec_eval(r.ctx, rgb.ctypes.data_as(C.c_void_p), C.c_float(lam))
            b = o.lib.oracle_rgb2spec_eval(o.ctx, rgb.ctypes.data_as(C.c_void_p), C.c_float(lam))
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (rgb, lam, a, b)


def test_primary_rays_match_the_reference_camera_shader():
    ref, orc = refpin.refshade_lib(), H.oracle_lib()
    prep = scenes.cornell(640, 360)
    sd = np.ascontiguousarray(prep["sceneData"])
    rng = np.random.default_rng(4)
    a, b = np.zeros(8, np.float32), np.zeros(8, np.float32)
    for _ in range(2000):
        px, py = int(rng.integers(0, 640)), int(rng.integers(0, 360))
        jx, jy = (rng.random(2, dtype=np.float32) - np.float32(0.5))
        ref.refshade_primary_ray(sd.ctypes.data_as(C.c_void_p), px, py, C.c_float(jx), C.c_float(jy), a.ctypes.data_as(C.c_void_p))
        orc.oracle_primary_ray(sd.ctypes.data_as(C.c_void_p), px, py, C.c_float(jx), C.c_float(jy), b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (px, py, a, b)


# ======================================================================================================================
# shaders: closures (VERDICT r01 "what's weak" 1: a check that is not common-mode with csrc/shading.cuh)
# ======================================================================================================================
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["rgb", "single", "hero"])
def test_oracle_closures_are_bit_identical_to_the_reference_shaders(mode):
    o, r = H.OracleBackend(), refpin.RefShadeBackend()
    payload, info = scenes.rgb2spec()
    for b in (o, r):
        b.check(b.f("set_rgb2spec")(b.ctx, payload.ctypes.data_as(C.c_void_p), C.c_uint32(len(payload)), info), "set_rgb2spec")
    q = refpin.random_closure_queries(20000, seed=100 + mode, mode=mode)
    want = refpin.eval_closures(r.lib, "refshade_eval_closures", r.ctx, q)
    got = refpin.eval_closures(o.lib, "oracle_eval_closures", o.ctx, q)
    # every lobe was really exercised
    m = q["material"]
    assert (m["sheenTintWeight"][:, 3] > 0).sum() > 5000 and (m["clearcoat"] > 0).sum() > 5000 and (m["subsurface"] > 0).sum() > 4000
    assert (m["diffuseRoughness"] > 0).sum() > 5000 and (m["anisotropic"] > 0).sum() > 5000 and (m["abbeNumber"] > 0).sum() > 1500
    assert (want["sampleFlags"] & 1).sum() > 10000 and (want["sampleFlags"] & 2).sum() > 500 and (want["evalPdf"][:, 0] > 0).sum() > 8000
    for f in refpin.CLOSURE_RESULT.names:
        a, b = want[f], got[f]
        same = a.view(np.uint32) == b.view(np.uint32)
        if a.dtype.kind == "f":
            same |= np.isnan(a) & np.isnan(b)
        bad = np.argwhere(~same)
        assert len(bad) == 0, "%s differs in %d of %d values; first: query %s reference %s oracle %s" % (
            f, len(bad), same.size, bad[0], a[tuple(bad[0])], b[tuple(bad[0])])


# ======================================================================================================================
# shaders: whole frames
# ======================================================================================================================
def _render_pair(prep, w, h, frames, mode, hero, debug=0, **sd_overrides):
    out = []
    for backend in (H.OracleBackend(), refpin.RefShadeBackend()):
        backend.upload(prep, rgb2spec=scenes.rgb2spec())
        backend.resize(w, h)
        sd = prep["sceneData"].copy()
        sd["packedRenderSettings"] = hr.pack_render_settings(int(sd["packedRenderSettings"]) & 0xFFFF, mode, hero)
        sd["debugMode"] = debug
        for k, v in sd_overrides.items():
            sd[k] = v
        backend.render(sd, frames=frames)
        out.append({k: backend.read(k) for k in (H.AOV_ACCUM, H.AOV_ALBEDO, H.AOV_NORMAL, H.AOV_OUTPUT)})
    return out


def _assert_frames_identical(a, b, what):
    for k in a:
        x, y = a[k].view(np.uint16 if a[k].dtype.itemsize == 2 else np.uint32), b[k].view(np.uint16 if b[k].dtype.itemsize == 2 else np.uint32)
        n = int((x != y).any(axis=-1).sum())
        assert n == 0, "%s: AOV %d differs in %d of %d pixels" % (what, k, n, x.shape[0] * x.shape[1])


@pytest.mark.parametrize("mode,hero", [(0, 0), (1, 0), (1, 1)], ids=["rgb", "single", "hero"])
@pytest.mark.parametrize("glass", [False, True], ids=["diffuse", "glass"])
def test_oracle_frames_are_bit_identical_to_the_reference_shaders_cornell(mode, hero, glass):
    w, h = 80, 56
    prep = scenes.cornell(w, h, spp=3, glass=glass)
    a, b = _render_pair(prep, w, h, 2, mode, hero)
    assert a[H.AOV_ACCUM][..., :3].mean() > 0.01
    _assert_frames_identical(a, b, "cornell")


@pytest.mark.parametrize("mode,hero", [(0, 0), (1, 1)], ids=["rgb", "hero"])
def test_oracle_frames_are_bit_identical_to_the_reference_shaders_textured_and_instanced(mode, hero):
    w, h = 72, 48
    for name, prep in (("textured", scenes.textured(w, h, spp=2, cutout=False)), ("instanced", scenes.instanced(w, h, count=27, spp=2)),
                       ("sunlit", scenes.sunlit(w, h, spp=2)), ("lobes", scenes.lobes(w, h, spp=2))):
        a, b = _render_pair(prep, w, h, 2, mode, hero)
        _assert_frames_identical(a, b, name)


@pytest.mark.parametrize("debug", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_oracle_debug_views_are_bit_identical_to_the_reference_shaders(debug):
    w, h = 48, 32
    prep = scenes.textured(w, h, spp=2, cutout=False)
    a, b = _render_pair(prep, w, h, 1, 0, 0, debug=debug)
    _assert_frames_identical(a, b, "debug view %d" % debug)


def test_reference_any_hit_alpha_agrees_with_the_oracle_in_the_mean():
    """Stochastic alpha: the reference's any-hit shader draws from the path RNG in driver traversal order (unpinnable); the oracle and the
    CUDA path hash (ray seed, instance, primitive) instead. Same acceptance probability, different random numbers:
    cut-out (MASK) coverage is deterministic and must match exactly where no BLEND surface is involved, images agree in the mean."""
    w, h = 64, 48
    for name, prep in (("alpha_blend", scenes.alpha_blend(w, h, spp=16)), ("textured+cutout", scenes.textured(w, h, spp=16))):
        a, b = _render_pair(prep, w, h, 8, 0, 0)
        ma, mb = a[H.AOV_ACCUM][..., :3].mean(axis=(0, 1)), b[H.AOV_ACCUM][..., :3].mean(axis=(0, 1))
        assert np.allclose(ma, mb, rtol=0.03), (name, ma, mb)
        blocks = lambda x: x[..., :3].reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3))  # noqa: E731
        rel = np.abs(blocks(a[H.AOV_ACCUM]) - blocks(b[H.AOV_ACCUM])) / (blocks(a[H.AOV_ACCUM]) + 0.05)
        assert float(np.sqrt((rel ** 2).mean())) < 0.08, (name, float(np.sqrt((rel ** 2).mean())))
        # the denoiser normal of the first non-specular hit does not depend on the random numbers: identical where both saw a surface
        assert np.array_equal(a[H.AOV_ACCUM][..., 3], b[H.AOV_ACCUM][..., 3])


# ======================================================================================================================
# host: the reference's C sources against vkrt_b200/host (libvkrt_host.so, no device)
# ======================================================================================================================
SHARED_STRUCTS = [
    ("Vertex", ["position", "normal", "tangent", "color", "texcoord0", "texcoord1"]),
    ("ShaderVertex", ["position", "texcoord0", "texcoord1", "packedNormal", "packedTangent", "packedColor"]),
    ("MeshInfo", ["position", "vertexBase", "rotation", "vertexCount", "scale", "indexBase", "indexCount", "materialIndex", "renderBackfaces",
                  "lightPdfArea", "opacity", "reserved0", "reserved1", "reserved2"]),
    ("Material", ["baseColor", "roughness", "emissionColor", "emissionLuminance", "eta", "metallic", "k", "anisotropic", "specular", "specularTint",
                  "abbeNumber", "reserved0", "sheenTintWeight", "clearcoat", "clearcoatGloss", "ior", "diffuseRoughness", "transmission", "subsurface",
                  "sheenRoughness", "absorptionCoefficient", "attenuationColor", "normalTextureScale", "baseColorTextureIndex",
                  "metallicRoughnessTextureIndex", "normalTextureIndex", "emissiveTextureIndex", "baseColorTextureWrap", "metallicRoughnessTextureWrap",
                  "normalTextureWrap", "emissiveTextureWrap", "opacity", "alphaCutoff", "alphaMode", "textureTexcoordSets", "baseColorTextureTransform",
                  "metallicRoughnessTextureTransform", "normalTextureTransform", "emissiveTextureTransform", "textureRotations"]),
    ("EmissiveMesh", ["triOffset", "triCount", "pmfMesh", "invTotalArea", "emission", "reserved0"]),
    ("EmissiveTriangle", ["v0Area", "e1Pad", "e2Pad"]),
    ("RGB2SpecTableInfo", ["res", "scaleOffset", "dataOffset"]),
    ("SceneData", ["viewInverse", "projInverse", "frameNumber", "samplesPerPixel", "rrMaxDepth", "rrMinDepth", "viewportRect", "packedRenderSettings",
                   "exposure", "timeBase", "timeStep", "environmentLight", "environmentTextureIndex", "environmentRotation", "debugMode", "misNeeEnabled",
                   "emissiveMeshCount", "emissiveTriangleCount", "selectionEnabled", "selectedMeshIndex", "rgb2specSRGB"]),
]


def test_shared_struct_layouts_match_the_reference_headers(tmp_path):
    """sizeof and every offsetof of include/vkrt_shared.h (compiled here with gcc) against src/shared/types.h compiled inside
    libvkrt_refhost.so: fails when the wire format drifts from the reference source."""
    lib = refpin.refhost_lib()
    want = np.zeros(512, np.uint32)
    n = lib.refhost_struct_layout(want.ctypes.data_as(C.c_void_p), C.c_uint32(512))
    assert n == sum(1 + len(f) for _, f in SHARED_STRUCTS)
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "vkrt_shared.h"', 'int main(void) {']
    for name, fields in SHARED_STRUCTS:
        src.append('printf("%%zu\\n", sizeof(%s));' % name)
        src += ['printf("%%zu\\n", offsetof(%s, %s));' % (name, f) for f in fields]
    src.append('return 0; }')
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(H.ROOT, "include"), str(c), "-o", str(exe)])
    got = np.array(subprocess.check_output([str(exe)]).split(), dtype=np.uint32)
    labels = [("sizeof(%s)" % s) for s, fs in SHARED_STRUCTS for _ in [0]] and [lab for s, fs in SHARED_STRUCTS for lab in ["sizeof(%s)" % s] + ["%s.%s" % (s, f) for f in fs]]
    bad = [(labels[i], int(want[i]), int(got[i])) for i in range(n) if want[i] != got[i]]
    assert not bad, bad
    # the numpy mirrors used by the tests follow the same layout
    assert hr.MATERIAL.itemsize == 272 and hr.SCENE_DATA.itemsize == 240 and hr.MESH_INFO.itemsize == 80


def _host():
    from vkrt_b200 import host
    return host


def _random_vertices(n, seed):
    rng = np.random.default_rng(seed)
    v = np.zeros(n, hr.VERTEX)
    v["position"] = rng.normal(size=(n, 4)).astype(np.float32)
    v["normal"][:, :3] = rng.normal(size=(n, 3)).astype(np.float32)
    v["tangent"] = rng.normal(size=(n, 4)).astype(np.float32)
    v["color"] = (rng.random((n, 4)) * 1.4 - 0.2).astype(np.float32)
    v["texcoord0"] = rng.normal(size=(n, 2)).astype(np.float32)
    v["texcoord1"] = rng.normal(size=(n, 2)).astype(np.float32)
    # corner cases of the octahedral encoders and the rounding
    k = n // 16
    v["normal"][:k, :3] = 0.0
    v["normal"][k:2 * k, :3] *= np.float32(1e-12)
    v["normal"][2 * k:3 * k, 2] = -np.abs(v["normal"][2 * k:3 * k, 2])
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 1, 0], [-1, 1, -1]], np.float32)
    v["normal"][3 * k:3 * k + 8, :3] = axes
    v["tangent"][3 * k:3 * k + 8, :3] = axes[::-1]
    v["tangent"][4 * k:5 * k, 3] = 0.0
    v["tangent"][5 * k:6 * k, 3] = -0.0
    v["color"][6 * k:7 * k] = np.float32(0.5) + (rng.integers(-3, 4, (k, 4)) / np.float32(510.0)).astype(np.float32)  # rounding ties of x * 255
    v["normal"][7 * k:7 * k + 4, 0] = [np.nan, np.inf, -np.inf, 1e38]
    return v


def test_pack_shader_vertex_is_bit_identical_to_the_reference():
    """VKRT_packShaderVertex (vkrt_b200/host/scene_prep.c) against src/core/utility/packing.c:92-156 on one million vertices."""
    ref = refpin.refhost_lib()
    hostlib = _host().load_host_library()
    n = 1_000_000
    v = _random_vertices(n, 7)
    want = np.zeros(n, _host().SHADER_VERTEX)
    ref.refhost_pack_vertices(v.ctypes.data_as(C.c_void_p), C.c_uint32(n), want.ctypes.data_as(C.c_void_p))
    got = np.zeros(n, _host().SHADER_VERTEX)
    stride_in, stride_out = v.dtype.itemsize, got.dtype.itemsize
    hostlib.VKRT_packShaderVertex.argtypes = [C.c_void_p, C.c_void_p]
    base_in, base_out = v.ctypes.data, got.ctypes.data
    for i in range(n):  # the product's entry point packs one vertex per call, like the reference's
        hostlib.VKRT_packShaderVertex(base_in + i * stride_in, base_out + i * stride_out)
    for f in ("position", "texcoord0", "texcoord1", "packedNormal", "packedTangent", "packedColor"):   # not the 4 tail-padding bytes
        a, b = want[f].reshape(n, -1).view(np.uint32), got[f].reshape(n, -1).view(np.uint32)
        bad = np.flatnonzero((a != b).any(axis=1))
        assert len(bad) == 0, (f, len(bad), v[bad[:3]], want[f][bad[:3]], got[f][bad[:3]])
    # the oracle-side numpy restatement (used to build every test scene) against the reference as well
    finite = np.isfinite(v["normal"]).all(axis=1)
    py = hr.pack_shader_vertices(v[finite])
    w = want[finite]
    for f in ("packedNormal", "packedTangent", "packedColor"):
        bad = np.flatnonzero(py[f] != w[f])
        assert len(bad) == 0, (f, len(bad), v[finite][bad[:3]], py[f][bad[:3]], w[f][bad[:3]])


def test_transforms_are_bit_identical_to_the_reference():
    ref = refpin.refhost_lib()
    hostlib = _host().load_host_library()
    rng = np.random.default_rng(8)
    m_ref, m_got = np.zeros(16, np.float32), np.zeros(16, np.float32)
    p2, r2, s2, p3, r3, s3 = (np.zeros(3, np.float32) for _ in range(6))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for i in range(4000):
        pos = rng.normal(size=3).astype(np.float32) * 5
        rot = (rng.random(3).astype(np.float32) * 720 - 360)
        scale = np.exp(rng.normal(size=3)).astype(np.float32)
        if i % 3 == 0:
            scale *= rng.choice([-1.0, 1.0], 3).astype(np.float32)
        if i % 17 == 0:
            rot[1] = rng.choice([90.0, -90.0, 270.0])      # gimbal lock
        if i % 29 == 0:
            scale[rng.integers(0, 3)] = 1e-7               # degenerate axis
        ref.refhost_build_transform(ptr(pos), ptr(rot), ptr(scale), ptr(m_ref))
        hostlib.VKRT_buildMeshTransformMatrix(ptr(pos), ptr(rot), ptr(scale), ptr(m_got))
        assert np.array_equal(m_ref.view(np.uint32), m_got.view(np.uint32)), (pos, rot, scale, m_ref, m_got)
        ref.refhost_decompose_transform(ptr(m_ref), ptr(p2), ptr(r2), ptr(s2))
        hostlib.VKRT_decomposeMeshTransform(ptr(m_ref), ptr(p3), ptr(r3), ptr(s3))
        for a, b in ((p2, p3), (r2, r3), (s2, s3)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (pos, rot, scale, p2, r2, s2, p3, r3, s3)
        # general (sheared) matrices as they arrive from glTF node hierarchies
        g = np.eye(4, dtype=np.float32)
        g[:3, :3] = rng.normal(size=(3, 3)).astype(np.float32)
        g[3, :3] = rng.normal(size=3).astype(np.float32)
        g = np.ascontiguousarray(g.reshape(16))
        ref.refhost_decompose_transform(ptr(g), ptr(p2), ptr(r2), ptr(s2))
        hostlib.VKRT_decomposeMeshTransform(ptr(g), ptr(p3), ptr(r3), ptr(s3))
        for a, b in ((p2, p3), (r2, r3), (s2, s3)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (g, p2, r2, s2, p3, r3, s3)
        ref.refhost_imported_node_transform(ptr(g), ptr(m_ref))
        hostlib.VKRT_buildImportedNodeTransform(ptr(g), ptr(m_got))
        assert np.array_equal(m_ref.view(np.uint32), m_got.view(np.uint32))


def _garbage_materials(n, seed):
    rng = np.random.default_rng(seed)
    raw = np.zeros(n, hr.MATERIAL)
    words = raw.view(np.uint32).reshape(n, -1)
    words[:] = rng.integers(0, 2 ** 32, words.shape, dtype=np.uint64).astype(np.uint32)   # arbitrary bit patterns: NaN, inf, denormals, huge
    sane = rng.random(n) < 0.5
    for name in hr.MATERIAL.names:
        if raw[name].dtype.kind == "f":
            raw[name][sane] = (rng.normal(size=raw[name][sane].shape) * 1.5).astype(np.float32)
    raw["alphaMode"] = rng.integers(0, 5, n)
    for t in ("baseColorTextureIndex", "metallicRoughnessTextureIndex", "normalTextureIndex", "emissiveTextureIndex"):
        raw[t] = np.where(rng.random(n) < 0.5, 0xFFFFFFFF, rng.integers(0, 4, n)).astype(np.uint32)
    return raw


def test_material_sanitisation_is_bit_identical_to_the_reference():
    """VKRT_addMaterial / VKRT_setMaterial of the product against src/core/api/mesh.c:107-200 on arbitrary bit patterns."""
    ref = refpin.refhost_lib()
    host = _host()
    h = ref.refhost_create(C.c_uint32(64), C.c_uint32(64))
    mine = host.Host(width=64, height=64, host_only=True)

    class Snapshot(C.Structure):   # VKRT_MaterialSnapshot (include/vkrt_host.h): 16-byte aligned like Material
        _fields_ = [("material", C.c_uint8 * 272), ("useCount", C.c_uint32), ("name", C.c_char * 256), ("_tail", C.c_uint8 * 12)]
    assert C.sizeof(Snapshot) == 544
    try:
        raw = _garbage_materials(3000, 9)
        idx_ref, idx_mine = C.c_uint32(), C.c_uint32()
        assert ref.refhost_add_material(C.c_void_p(h), None, C.byref(idx_ref)) == 0
        assert mine.lib.VKRT_addMaterial(mine.h, None, b"m", C.byref(idx_mine)) == 0
        assert idx_ref.value == idx_mine.value == 1          # index 0 is the default material on both sides
        out = np.zeros(1, hr.MATERIAL)
        snap = Snapshot()
        for i in range(len(raw)):
            m = np.ascontiguousarray(raw[i:i + 1])
            ra = ref.refhost_set_material(C.c_void_p(h), C.c_uint32(1), m.ctypes.data_as(C.c_void_p))
            rb = mine.lib.VKRT_setMaterial(mine.h, C.c_uint32(1), m.ctypes.data_as(C.c_void_p))
            assert ra == rb == 0
            assert ref.refhost_get_material(C.c_void_p(h), C.c_uint32(1), out.ctypes.data_as(C.c_void_p)) == 0
            assert mine.lib.VKRT_getMaterialSnapshot(mine.h, C.c_uint32(1), C.byref(snap)) == 0
            got = np.frombuffer(bytes(snap.material), hr.MATERIAL)
            assert out.tobytes() == got.tobytes(), (i, [(n, out[n], got[n]) for n in hr.MATERIAL.names if out[n].tobytes() != got[n].tobytes()])
        # the default material itself
        assert ref.refhost_get_material(C.c_void_p(h), C.c_uint32(0), out.ctypes.data_as(C.c_void_p)) == 0
        assert mine.lib.VKRT_getMaterialSnapshot(mine.h, C.c_uint32(0), C.byref(snap)) == 0
        assert out.tobytes() == bytes(snap.material)
    finally:
        ref.refhost_destroy(C.c_void_p(h))
        mine.close()


def test_camera_matrices_are_bit_identical_to_the_reference():
    """syncCameraMatrices (src/core/scene/camera.c:128-143: cglm lookat / perspective / Y flip / inverses) against the product's
    VKRT_cameraSetPose: these 32 floats decide every primary ray."""
    ref = refpin.refhost_lib()
    host = _host()
    rng = np.random.default_rng(10)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for w, hgt in ((1920, 1080), (512, 512), (3840, 2160), (97, 33)):
        h = ref.refhost_create(C.c_uint32(w), C.c_uint32(hgt))
        mine = host.Host(width=w, height=hgt, host_only=True)
        try:
            for _ in range(300):
                pos = rng.normal(size=3).astype(np.float32) * 4
                target = rng.normal(size=3).astype(np.float32)
                up = np.array([0, 0, 1], np.float32) if rng.random() < 0.7 else rng.normal(size=3).astype(np.float32)
                vfov = np.float32(rng.uniform(10, 120))
                assert ref.refhost_set_camera(C.c_void_p(h), ptr(pos), ptr(target), ptr(up), C.c_float(vfov), C.c_float(0.001), C.c_float(10000.0)) == 0
                sd_ref = np.zeros(1, hr.SCENE_DATA)
                ref.refhost_get_scene_data(C.c_void_p(h), ptr(sd_ref))
                mine.camera_set_pose(pos, target, up, float(vfov))
                mine.start_render(w, hgt, 1)
                mine.update_scene()
                sd = np.frombuffer(mine.prepare_scene()["sceneData"].tobytes(), hr.SCENE_DATA)
                for f in ("viewInverse", "projInverse"):
                    assert np.array_equal(sd_ref[f].view(np.uint32), sd[f].view(np.uint32)), (f, pos, target, up, vfov, sd_ref[f], sd[f])
        finally:
            ref.refhost_destroy(C.c_void_p(h))
            mine.close()


def test_light_tables_are_bit_identical_to_the_reference():
    """vkrtSceneRebuildLightBuffers (src/core/scene/lighting.c:496-544: emissive triangle list in world space, per-mesh and per-triangle
    alias tables, pmf, invTotalArea, lightPdfArea) against the product's host on a scene of transformed emissive meshes."""
    ref = refpin.refhost_lib()
    host = _host()
    rng = np.random.default_rng(11)
    w, hgt = 64, 64
    h = ref.refhost_create(C.c_uint32(w), C.c_uint32(hgt))
    mine = host.Host(width=w, height=hgt, host_only=True)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    try:
        meshes = []
        for k in range(9):
            if k % 3 == 0:
                hm = hr.uv_sphere_mesh("s%d" % k, radius=0.5 + 0.1 * k, segments=12 + k, rings=6 + k)
            elif k % 3 == 1:
                hm = hr.box_mesh("b%d" % k, half=(0.3 + 0.1 * k, 0.2, 0.5))
            else:
                hm = hr.quad_mesh("q%d" % k, size=1.0 + k)
            meshes.append(hm)
        mats = []
        for k in range(6):
            m = hr.default_material()
            m["baseColor"] = rng.random(3)
            if k in (1, 2, 4, 5):
                m["emissionColor"] = rng.random(3) + 0.05
                m["emissionLuminance"] = float(rng.uniform(0.5, 40.0))
            if k == 5:
                m["emissionLuminance"] = 0.0   # emission colour without luminance: not a light
            mats.append(np.ascontiguousarray(m))
        idx = C.c_uint32()
        for m in mats:
            assert ref.refhost_add_material(C.c_void_p(h), ptr(m), C.byref(idx)) == 0
            assert mine.lib.VKRT_addMaterial(mine.h, ptr(m), b"m", C.byref(idx)) == 0
        for k, hm in enumerate(meshes):
            v = np.ascontiguousarray(hm.vertices)
            ix = np.ascontiguousarray(hm.indices, dtype=np.uint32)
            mat_index = 1 + ((k + 1) % len(mats))
            assert ref.refhost_add_mesh(C.c_void_p(h), ptr(v), C.c_uint32(len(v)), ptr(ix), C.c_uint32(len(ix)), C.c_uint32(mat_index)) == 0
            assert mine.lib.VKRT_uploadMeshData(mine.h, ptr(v), C.c_size_t(len(v)), ptr(ix), C.c_size_t(len(ix))) == 0
            assert mine.lib.VKRT_setMeshMaterialIndex(mine.h, C.c_uint32(k), C.c_uint32(mat_index)) == 0
            pos = rng.normal(size=3).astype(np.float32) * 3
            rot = (rng.random(3) * 360 - 180).astype(np.float32)
            scale = np.exp(rng.normal(size=3) * 0.5).astype(np.float32)
            if k == 4:
                scale[0] = -scale[0]
            assert ref.refhost_set_mesh_transform(C.c_void_p(h), C.c_uint32(k), ptr(pos), ptr(rot), ptr(scale)) == 0
            assert mine.lib.VKRT_setMeshTransform(mine.h, C.c_uint32(k), ptr(pos), ptr(rot), ptr(scale)) == 0
        n_mesh, n_tri = C.c_uint32(), C.c_uint32()
        assert ref.refhost_rebuild_lights(C.c_void_p(h), C.byref(n_mesh), C.byref(n_tri)) == 0
        mine.start_render(w, hgt, 1)
        mine.update_scene()
        prep = mine.prepare_scene()
        assert n_mesh.value == len(prep["emissiveMeshes"]) > 3 and n_tri.value == len(prep["emissiveTriangles"]) > 100
        names = ["emissiveMeshes", "emissiveTriangles", "meshAliasQ", "meshAliasIdx", "triAliasQ", "triAliasIdx"]
        for which, name in enumerate(names):
            size = ref.refhost_read_light_buffer(C.c_void_p(h), which, None, C.c_uint64(0))
            buf = np.zeros(size, np.uint8)
            assert ref.refhost_read_light_buffer(C.c_void_p(h), which, ptr(buf), C.c_uint64(size)) == size
            got = prep[name].view(np.uint8).reshape(-1)
            assert size == got.size and np.array_equal(buf, got), "%s differs from the reference (%d bytes, %d different)" % (
                name, size, int((buf[:min(size, got.size)] != got[:min(size, got.size)]).sum()))
        info = np.zeros(1, hr.MESH_INFO)
        world = np.zeros(12, np.float32)
        for k in range(len(meshes)):
            assert ref.refhost_get_mesh(C.c_void_p(h), C.c_uint32(k), ptr(info), ptr(world)) == 0
            mi = prep["meshInfos"][k]
            for f in ("position", "rotation", "scale", "lightPdfArea", "vertexBase", "vertexCount", "indexBase", "indexCount", "materialIndex", "opacity"):
                assert info[f][0].tobytes() == mi[f].tobytes(), (k, f, info[f][0], mi[f])
            assert np.array_equal(world.view(np.uint32), prep["world3x4"][k].reshape(-1).view(np.uint32)), (k, world, prep["world3x4"][k])
    finally:
        ref.refhost_destroy(C.c_void_p(h))
        mine.close()


def test_render_setting_clamps_match_the_reference():
    """The setters of src/core/api/settings.c against the product's: same clamps, same SceneData words."""
    ref = refpin.refhost_lib()
    host = _host()
    h = ref.refhost_create(C.c_uint32(320), C.c_uint32(200))
    mine = host.Host(width=320, height=200, host_only=True)
    fields = ("samplesPerPixel", "rrMaxDepth", "rrMinDepth", "packedRenderSettings", "exposure", "environmentLight", "misNeeEnabled")
    try:
        cases = [(0, 0, 0), (1, 3, 2), (4096, 9, 100), (8, 0, 5), (2047, 64, 64), (2049, 65, 1000)]
        for spp, rr_min, rr_max in cases:
            for mode, spectral, tone in ((0, 0, 0), (1, 0, 1), (1, 1, 1), (7, 9, 5)):
                exposure = float(np.float32([0.0, -3.0, 1e9, 0.37, float("nan")][(spp + mode) % 5]))
                env = np.array([0.2, 5.0, -1.0], np.float32)
                strength = float([0.25, -2.0, 1e12][mode % 3])
                for lib_, hh in ((ref, C.c_void_p(h)), (None, mine)):
                    if lib_ is not None:
                        lib_.refhost_set_samples_per_pixel(hh, C.c_uint32(spp))
                        lib_.refhost_set_path_depth(hh, C.c_uint32(rr_min), C.c_uint32(rr_max))
                        lib_.refhost_set_render_mode(hh, C.c_uint32(mode))
                        lib_.refhost_set_spectral_sampling_mode(hh, C.c_uint32(spectral))
                        lib_.refhost_set_tone_mapping_mode(hh, C.c_uint32(tone))
                        lib_.refhost_set_exposure(hh, C.c_float(exposure))
                        lib_.refhost_set_environment_light(hh, env.ctypes.data_as(C.c_void_p), C.c_float(strength))
                        lib_.refhost_set_mis_nee_enabled(hh, C.c_uint32(spp % 2))
                    else:
                        L = mine.lib
                        L.VKRT_setSamplesPerPixel(mine.h, C.c_uint32(spp))
                        L.VKRT_setPathDepth(mine.h, C.c_uint32(rr_min), C.c_uint32(rr_max))
                        L.VKRT_setRenderMode(mine.h, C.c_uint32(mode))
                        L.VKRT_setSpectralSamplingMode(mine.h, C.c_uint32(spectral))
                        L.VKRT_setToneMappingMode(mine.h, C.c_uint32(tone))
                        L.VKRT_setExposure(mine.h, C.c_float(exposure))
                        L.VKRT_setEnvironmentLight(mine.h, env.ctypes.data_as(C.c_void_p), C.c_float(strength))
                        L.VKRT_setMisNeeEnabled(mine.h, C.c_uint8(spp % 2))
                sd_ref = np.zeros(1, hr.SCENE_DATA)
                ref.refhost_get_scene_data(C.c_void_p(h), sd_ref.ctypes.data_as(C.c_void_p))
                mine.start_render(320, 200, 1)
                mine.update_scene()
                sd = np.frombuffer(mine.prepare_scene()["sceneData"].tobytes(), hr.SCENE_DATA)
                for f in fields:
                    assert sd_ref[f].tobytes() == sd[f].tobytes(), (f, (spp, rr_min, rr_max, mode, spectral, tone, exposure, strength), sd_ref[f], sd[f])
    finally:
        ref.refhost_destroy(C.c_void_p(h))
        mine.close()


def test_remove_material_matches_the_reference():
    """VKRT_removeMaterial (src/core/api/mesh.c:296-342,449-459): the default material cannot be removed, meshes fall back to it, later
    indices shift down; a random sequence of add / remove / assign calls leaves both hosts with the same materials and mesh indices."""
    ref = refpin.refhost_lib()
    host = _host()
    ref.refhost_material_count.restype = C.c_uint32
    rng = np.random.default_rng(12)
    h = ref.refhost_create(C.c_uint32(64), C.c_uint32(64))
    mine = host.Host(width=64, height=64, host_only=True)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731

    class Snapshot(C.Structure):
        _fields_ = [("material", C.c_uint8 * 272), ("useCount", C.c_uint32), ("name", C.c_char * 256), ("_tail", C.c_uint8 * 12)]
    try:
        quad = hr.quad_mesh("q", 1.0)
        v, ix = np.ascontiguousarray(quad.vertices), np.ascontiguousarray(quad.indices, dtype=np.uint32)
        idx = C.c_uint32()
        for k in range(6):
            m = hr.default_material()
            m["baseColor"] = rng.random(3)
            m["roughness"] = float(rng.random())
            m = np.ascontiguousarray(m)
            assert ref.refhost_add_material(C.c_void_p(h), ptr(m), C.byref(idx)) == 0
            assert mine.lib.VKRT_addMaterial(mine.h, ptr(m), b"m", C.byref(idx)) == 0
        for k in range(8):
            assert ref.refhost_add_mesh(C.c_void_p(h), ptr(v), C.c_uint32(len(v)), ptr(ix), C.c_uint32(len(ix)), C.c_uint32(1 + k % 6)) == 0
            assert mine.lib.VKRT_uploadMeshData(mine.h, ptr(v), C.c_size_t(len(v)), ptr(ix), C.c_size_t(len(ix))) == 0
            assert mine.lib.VKRT_setMeshMaterialIndex(mine.h, C.c_uint32(k), C.c_uint32(1 + k % 6)) == 0
        for victim in (0, 3, 99, 1, 5, 1, 1, 1, 1, 1):
            ra = ref.refhost_remove_material(C.c_void_p(h), C.c_uint32(victim))
            rb = mine.lib.VKRT_removeMaterial(mine.h, C.c_uint32(victim))
            assert ra == rb, (victim, ra, rb)
            n_ref = ref.refhost_material_count(C.c_void_p(h))
            n_mine = C.c_uint32()
            mine.lib.VKRT_getMaterialCount(mine.h, C.byref(n_mine))
            assert n_ref == n_mine.value
            out = np.zeros(1, hr.MATERIAL)
            snap = Snapshot()
            for i in range(n_ref):
                assert ref.refhost_get_material(C.c_void_p(h), C.c_uint32(i), ptr(out)) == 0
                assert mine.lib.VKRT_getMaterialSnapshot(mine.h, C.c_uint32(i), C.byref(snap)) == 0
                assert out.tobytes() == bytes(snap.material), (victim, i)
            info = np.zeros(1, hr.MESH_INFO)
            world = np.zeros(12, np.float32)
            mine.start_render(64, 64, 1)
            mine.update_scene()
            prep = mine.prepare_scene()
            for k in range(8):
                assert ref.refhost_get_mesh(C.c_void_p(h), C.c_uint32(k), ptr(info), ptr(world)) == 0
                assert int(info["materialIndex"][0]) == int(prep["meshInfos"][k]["materialIndex"]), (victim, k)
    finally:
        ref.refhost_destroy(C.c_void_p(h))
        mine.close()


# ---- glTF import: the product's importer against the reference's own src/app/mesh/loader.c (cgltf) ------------------------------------------
class _GltfMesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertexCount", C.c_size_t), ("indices", C.c_void_p), ("indexCount", C.c_size_t), ("world", C.c_float * 16),
                ("materialIndex", C.c_int), ("doubleSided", C.c_int), ("name", C.c_char * 256)]


class _GltfTexture(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("colorSpace", C.c_uint32), ("name", C.c_char * 256)]


class _GltfImport(C.Structure):
    _fields_ = [("meshes", C.POINTER(_GltfMesh)), ("meshCount", C.c_uint32), ("materials", C.c_void_p), ("materialNames", C.c_void_p), ("materialCount", C.c_uint32),
                ("textures", C.POINTER(_GltfTexture)), ("textureCount", C.c_uint32)]


_VERTEX = np.dtype([("position", "<f4", 4), ("normal", "<f4", 4), ("tangent", "<f4", 4), ("color", "<f4", 4), ("texcoord0", "<f4", 2), ("texcoord1", "<f4", 2)])


def _glb_cases(tmp_path):
    import gltf_fixtures
    png = np.load(os.path.join(H.ROOT, "tests", "golden", "images.npz"))["file_png_rgba8"].tobytes()
    plain, textured = str(tmp_path / "hierarchy.glb"), str(tmp_path / "hierarchy_textured.glb")
    gltf_fixtures.hierarchy_glb(plain)
    gltf_fixtures.hierarchy_glb(textured, png)
    return [os.path.join(H.ROOT, "assets", "models", m + ".glb") for m in ("cube", "plane", "sphere", "prism", "suzanne", "bunny")] + [plain, textured]


def test_gltf_importer_matches_the_reference_loader(tmp_path):
    """vkrt_b200/host/gltf_import.c (own JSON + accessor reader) against the reference's importer compiled where it lies (src/app/mesh/loader.c
    + the vendored cgltf, oracle/ref_host/ref_loader_entry.c) on the six bundled models and on generated files that take every branch: u8 /
    u16 / u32 / absent indices, generated normals and tangents, winding alignment, float and normalised-byte colours, two UV sets, several
    primitives per mesh, TRS and matrix nodes with nesting and mirroring, alpha modes, double-sided, KHR_materials_{ior, transmission, volume,
    clearcoat, sheen, specular, emissive_strength}, texture references with samplers, texCoord sets and KHR_texture_transform.
    Entries (order, names, material index, back-face flag), vertices and indices: byte-identical; materials: byte-identical; textures: same
    list of (name, colour space) (the reference side decodes with a stand-in); node hierarchy: world matrices within 1e-5."""
    ref = refpin.refhost_lib()
    host = C.CDLL(os.path.join(H.ROOT, "vkrt_b200", "libvkrt_host.so"))
    ref.refloader_load.restype = C.c_void_p
    ref.refloader_load.argtypes = [C.c_char_p]
    for path in _glb_cases(tmp_path):
        h = ref.refloader_load(path.encode())
        assert h, path
        hp = C.c_void_p(h)
        counts = (C.c_uint32 * 4)()
        ref.refloader_counts(hp, counts)
        imp, err = _GltfImport(), C.create_string_buffer(256)
        assert host.gltfImportFile(path.encode(), C.byref(imp), err, 256) == 1, err.value
        assert (imp.meshCount, imp.materialCount, imp.textureCount) == (counts[0], counts[2], counts[3]), path
        # node world matrices of the reference: product of the local transforms up the parent chain (column-major, as cglm)
        locals_, parents = [], []
        for n in range(counts[1]):
            loc, pc, prs = (C.c_float * 16)(), (C.c_uint32 * 2)(), (C.c_float * 9)()
            ref.refloader_node(hp, n, loc, pc, prs)
            locals_.append(np.array(list(loc), np.float64).reshape(4, 4).T)
            parents.append(pc[0])

        def world_of(n):
            m = locals_[n]
            while parents[n] != 0xFFFFFFFF:
                n = parents[n]
                m = locals_[n] @ m
            return m
        for i in range(counts[0]):
            info, prs, name = (C.c_uint64 * 5)(), (C.c_float * 9)(), C.create_string_buffer(256)
            ref.refloader_entry_info(hp, i, info, prs, name, 256)
            m = imp.meshes[i]
            assert (m.vertexCount, m.indexCount) == (info[0], info[1]), (path, i)
            assert m.name == name.value, (path, i, m.name, name.value)
            assert m.materialIndex == (-1 if info[3] == 0xFFFFFFFF else int(info[3])) and m.doubleSided == info[4], (path, i)
            rv, ri = np.zeros(info[0], _VERTEX), np.zeros(info[1], np.uint32)
            ref.refloader_entry_data(hp, i, rv.ctypes.data_as(C.c_void_p), ri.ctypes.data_as(C.c_void_p))
            mv = np.frombuffer((C.c_char * (m.vertexCount * 80)).from_address(m.vertices), _VERTEX)
            mi = np.frombuffer((C.c_char * (m.indexCount * 4)).from_address(m.indices), np.uint32)
            assert np.array_equal(ri, mi), (path, i)
            for field in _VERTEX.names:
                assert np.array_equal(rv[field].view(np.uint32), mv[field].view(np.uint32)), (path, i, field)
            assert np.allclose(list(prs), [0, 0, 0, 0, 0, 0, 1, 1, 1])   # the entry itself carries no transform: its node does
            mine_world = np.array(list(m.world), np.float64).reshape(4, 4).T
            assert np.allclose(mine_world, world_of(int(info[2])), atol=1e-5), (path, i)
        for k in range(counts[2]):
            rm, name = np.zeros(1, H.hr.MATERIAL), C.create_string_buffer(256)
            ref.refloader_material(hp, k, rm.ctypes.data_as(C.c_void_p), name, 256)
            mine = (C.c_char * 272).from_address(imp.materials + 272 * k).raw
            assert rm.tobytes() == mine, (path, k, [f for f in H.hr.MATERIAL.names if rm[f].tobytes() != np.frombuffer(mine, H.hr.MATERIAL)[f].tobytes()])
            assert (C.c_char * 256).from_address(imp.materialNames + 256 * k).value == name.value
        for k in range(counts[3]):
            desc, name = (C.c_uint32 * 4)(), C.create_string_buffer(256)
            ref.refloader_texture(hp, k, desc, name, 256)
            assert imp.textures[k].name == name.value and imp.textures[k].colorSpace == desc[3], (path, k)
        host.gltfImportFree(C.byref(imp))
        ref.refloader_free(hp)


# ---- vkrt.scene files: the product's reader against the reference's own scene controller --------------------------------------------------
@pytest.mark.parametrize("scene", ["cornell", "prism", "caustics"])
def test_scene_file_reader_matches_the_reference_controller(scene):
    """VKRT_appLoadScene (vkrt_b200/host/scene_file.c: own JSON reader, object hierarchy, material table) against the reference's
    sceneControllerLoadSceneFromPath compiled where it lies (src/app/scene/controller.c + cJSON, session.c, mesh/controller.c, loader.c + cgltf,
    api/{query,geometry,texture,environment,render,mesh,settings}.c, scene/{geometry,environment,transform,...}.c behind the null device;
    oracle/ref_host/ref_scene_entry.c) on the bundled scenes: mesh order and names, geometry sharing, vertex / index ranges, material
    assignment, back-face and opacity flags identical; materials byte-identical; world matrices and the decomposed position / rotation /
    scale within a few ulp; every scene setting (camera, depths, modes, exposure, environment) identical."""
    from vkrt_b200 import host
    ref = refpin.refhost_lib()
    ref.refscene_load.argtypes = [C.c_void_p, C.c_char_p]
    ref.refscene_mesh_count.argtypes = [C.c_void_p]
    ref.refscene_mesh_count.restype = C.c_uint32
    path = os.path.join(H.ROOT, "assets", "scenes", scene + ".json")
    h = C.c_void_p(ref.refhost_create(320, 180))
    assert ref.refscene_load(h, path.encode()) == 1
    hs = host.Host(host_only=True, width=320, height=180)
    hs.load_scene(path)
    hs.start_render(320, 180, 64)
    got = hs.prepare_scene()
    n = ref.refscene_mesh_count(h)
    assert n == len(got["meshInfos"]) == hs.mesh_count()

    class MeshSnapshot(C.Structure):
        _fields_ = [("info", C.c_uint8 * 80), ("material", C.c_uint8 * 272), ("materialIndex", C.c_uint32), ("geometrySource", C.c_uint32),
                    ("hasMaterialAssignment", C.c_uint8), ("ownsGeometry", C.c_uint8), ("name", C.c_char * 256),
                    ("tail", C.c_uint8 * 6)]   # the C struct is 16-byte aligned through MeshInfo / Material: sizeof == 624
    assert C.sizeof(MeshSnapshot) == 624
    for i in range(n):
        info, world, misc, name = np.zeros(1, H.hr.MESH_INFO), (C.c_float * 16)(), (C.c_uint32 * 4)(), C.create_string_buffer(256)
        ref.refscene_mesh(h, i, info.ctypes.data_as(C.c_void_p), world, misc, name, 256)
        mine = got["meshInfos"][i]
        for key in ("vertexBase", "vertexCount", "indexBase", "indexCount", "materialIndex", "renderBackfaces"):
            assert int(mine[key]) == int(info[key][0]), (scene, i, key)
        assert float(mine["opacity"]) == float(info["opacity"][0])
        assert int(got["geometrySource"][i]) == misc[0], (scene, i)
        snap = MeshSnapshot()
        assert hs.lib.VKRT_getMeshSnapshot(hs.h, C.c_uint32(i), C.byref(snap)) == 0
        assert snap.name == name.value and snap.ownsGeometry == misc[1] and snap.hasMaterialAssignment == misc[2], (scene, i, snap.name, name.value)
        ref_world = np.array(list(world), np.float32).reshape(4, 4).T[:3]
        a, b = got["world3x4"][i].astype(np.float64), ref_world.astype(np.float64)
        assert np.all(np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * 8 * 1.2e-7 + 1e-6), (scene, i, a, b)
        assert np.allclose(mine["position"], info["position"][0], rtol=0, atol=1e-6)
        assert np.allclose(mine["scale"], info["scale"][0], rtol=1e-5, atol=1e-6)
        d = np.abs(((np.asarray(mine["rotation"]) - info["rotation"][0]) + 180.0) % 360.0 - 180.0)
        assert float(d.max()) < 2e-3, (scene, i)
    # materials (sanitised by VKRT_addMaterial / VKRT_setMaterial on both sides)
    count = C.c_uint32()
    ref.refhost_material_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)] if hasattr(ref, "refhost_material_count") else None
    mats = np.frombuffer(got["materials"].tobytes(), H.hr.MATERIAL)
    for k in range(len(mats)):
        snapshot = np.zeros(544, np.uint8)   # VKRT_MaterialSnapshot {Material, useCount, name[256]} (tests above)
        assert ref.refhost_get_material(h, C.c_uint32(k), snapshot.ctypes.data_as(C.c_void_p)) == 0, (scene, k)
        assert snapshot[:272].tobytes() == mats[k].tobytes(), (scene, k, [f for f in H.hr.MATERIAL.names
                                                                          if np.frombuffer(snapshot[:272].tobytes(), H.hr.MATERIAL)[f].tobytes() != mats[k][f].tobytes()])
    # scene settings
    rs, ms = host.SceneSettings(), hs.scene_settings()
    ref.refscene_settings(h, C.byref(rs))
    for field, _ in host.SceneSettings._fields_:
        if field == "camera":   # (near / far are engine defaults that the shim's handle creation does not set: not part of a scene file)
            for part in ("pos", "target", "up"):
                assert list(getattr(rs.camera, part)) == list(getattr(ms.camera, part)), (scene, part)
            assert rs.camera.vfov == ms.camera.vfov
        elif field == "environmentColor":
            assert list(rs.environmentColor) == list(ms.environmentColor)
        elif field in ("rrMaxDepth", "rrMinDepth", "toneMappingMode", "renderMode", "spectralSamplingMode", "exposure", "autoExposureEnabled",
                       "environmentStrength", "environmentRotation", "environmentTextureIndex", "misNeeEnabled"):   # what a scene file sets
            assert getattr(rs, field) == getattr(ms, field), (scene, field, getattr(rs, field), getattr(ms, field))
    hs.close()
    ref.refscene_close()   # (the handle itself is left to the process: refhost_destroy frees what refhost_add_mesh allocated, not what scene/geometry.c did)


# ---- feedback controllers: vkrt_b200/host/controllers.c against the reference's timing.c / exposure.c -----------------------------------------
def test_feedback_controllers_match_the_reference():
    """Auto-SPP (src/core/scene/timing.c updateAutoSPP) and auto-exposure (src/core/scene/exposure.c: the 16 x 16 probe grid recorded as 256
    one-texel copies, resolveAutoExposureReadback), compiled unmodified into libvkrt_refhost.so, driven step by step next to the product's
    vkrtAutoSPPStep / vkrtAutoExposureStep / vkrtAutoExposureProbePixels: same probe pixels for any frame size, and bit-identical controller
    state over long random sequences (non-finite and non-positive samples included)."""
    from vkrt_b200 import host
    ref = refpin.refhost_lib()
    lib = host.load_host_library()
    lib.vkrtAutoSPPStep.restype = C.c_uint32
    lib.vkrtAutoSPPStep.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_uint32]
    lib.vkrtAutoExposureStep.restype = C.c_int
    lib.vkrtAutoExposureStep.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_uint32, C.c_float, C.POINTER(C.c_float)]
    lib.vkrtAutoExposureProbePixels.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    ref.refcontrol_probe_pixels.restype = C.c_uint32
    ref.refcontrol_probe_pixels.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    ref.refcontrol_exposure_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    ref.refcontrol_autospp_step.restype = C.c_uint32
    ref.refcontrol_autospp_step.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_uint32]
    h = C.c_void_p(ref.refhost_create(64, 64))
    bits = lambda v: np.float32(v).view(np.uint32)   # noqa: E731
    # probe pixels
    for w, hh in ((1920, 1080), (3840, 2160), (512, 512), (17, 9), (16, 16), (5, 3), (1, 1)):
        a, b = np.zeros(512, np.uint32), np.zeros(512, np.uint32)
        assert ref.refcontrol_probe_pixels(h, w, hh, a.ctypes.data_as(C.c_void_p)) == 256
        lib.vkrtAutoExposureProbePixels(w, hh, b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a, b), (w, hh)
    # auto-exposure: a drifting scene brightness with hostile samples, 300 steps
    rng = np.random.default_rng(21)
    rf, re_ = C.c_float(0.0), C.c_float(1.0)
    mf, me = C.c_float(0.0), np.float32(1.0)
    for step in range(300):
        level = np.float32(10.0 ** rng.uniform(-3, 2))
        s = (rng.random((256, 4), dtype=np.float32) * level).astype(np.float32)
        if step % 7 == 0:
            s[rng.integers(0, 256, 5), rng.integers(0, 3, 5)] = [np.nan, np.inf, -np.inf, -3.0, 0.0]
        if step % 50 == 49:
            s[:] = 0.0          # black frame: the controller must hold its state
        if step % 97 == 96:
            s[:] = np.nan
        ref.refcontrol_exposure_step(h, s.ctypes.data_as(C.c_void_p), C.byref(rf), C.byref(re_))
        out = C.c_float(0.0)
        if lib.vkrtAutoExposureStep(C.byref(mf), s.ctypes.data_as(C.c_void_p), 256, C.c_float(me), C.byref(out)):
            me = np.float32(out.value)
        assert bits(rf.value) == bits(mf.value) and bits(re_.value) == bits(me), (step, rf.value, mf.value, re_.value, me)
    # auto-SPP: frame times that follow the sample count with noise, several targets, 400 steps
    for target in (1000.0 / 60.0, 1000.0 / 24.0, 5.0, 200.0):
        rc, mc = C.c_float(0.0), C.c_float(0.0)
        rspp = mspp = 1
        cost = 0.35
        for step in range(400):
            if step % 80 == 79:
                cost *= float(rng.choice([0.25, 4.0]))     # the scene got cheaper / more expensive
            measured = np.float32(max(cost * rspp * rng.uniform(0.85, 1.2) + 0.3, 0.0 if step % 37 else -1.0))
            rspp = ref.refcontrol_autospp_step(h, C.byref(rc), C.c_float(target), C.c_float(measured), rspp)
            mspp = lib.vkrtAutoSPPStep(C.byref(mc), C.c_float(target), C.c_float(measured), mspp)
            assert rspp == mspp and bits(rc.value) == bits(mc.value), (target, step, rspp, mspp, rc.value, mc.value)
        assert 1 <= rspp <= 2048


# ---- OpenEXR: host/image_decode.c + host/export.c against the reference's exr.cpp over the vendored tinyexr -----------------------------------
class _LoadedImage(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("colorSpace", C.c_uint32)]


def _pixels_of(img):
    texel = {2: (np.uint16, 4), 3: (np.float32, 4)}[img.format]     # VKRT_TEXTURE_FORMAT_RGBA16_SFLOAT = 2, RGBA32_SFLOAT = 3
    n = img.width * img.height * texel[1]
    return np.ctypeslib.as_array(C.cast(img.pixels, C.POINTER(np.ctypeslib.as_ctypes_type(texel[0]))), shape=(n,)).copy()


def test_exr_reader_and_writer_match_the_reference_codec(tmp_path):
    """The reference reads and writes OpenEXR through tinyexr (src/core/utility/exr.cpp, compiled where it lies with the vendored header into
    oracle/_ref/libvkrt_refexr.so). (1) Every EXR fixture (NONE / RLE / ZIPS / ZIP / PIZ, half and float, RGB / RGBA / Y) decodes to the same
    format and the same bits with the host's own reader; the PXR24 fixture is refused by both (the vendored tinyexr has no PXR24 either). (2) A
    file written by the host's writer reads back through tinyexr with the pixels it was given, and (3) one written by the reference reads
    back through the host's reader."""
    from vkrt_b200 import host
    ref = refpin._load("libvkrt_refexr.so")
    host.load_host_library()
    lib = C.CDLL(host.HOST_LIB_PATH)   # (a private handle: other test modules set argtypes with their own struct classes on the shared one)
    for fn in (ref.vkrtLoadEXRImageFromMemory, ref.vkrtLoadEXRImageFromFile, lib.vkrtLoadImageFromFile, lib.vkrtLoadImageFromMemory):
        fn.restype = C.c_int
    gold = np.load(os.path.join(H.ROOT, "tests", "golden", "images.npz"))
    names = [k[5:] for k in gold.files if k.startswith("file_exr_")]
    assert len(names) >= 8
    for name in names:
        data = gold["file_" + name].tobytes()
        a, b = _LoadedImage(), _LoadedImage()
        ra = ref.vkrtLoadEXRImageFromMemory(data, C.c_size_t(len(data)), name.encode(), C.byref(a))
        rb = lib.vkrtLoadImageFromMemory(data, C.c_size_t(len(data)), b"image/x-exr", C.c_uint32(1), C.byref(b))
        if "pxr24" in name:
            assert ra == 0 and rb == 0   # rejected with a message on both sides, never mis-decoded
            continue
        assert ra == 1 and rb == 1, name
        assert (a.width, a.height, a.format) == (b.width, b.height, b.format), (name, a.format, b.format)
        pa, pb = _pixels_of(a), _pixels_of(b)
        assert np.array_equal(pa.view(np.uint8), pb.view(np.uint8)), name
        lib.vkrtFreeLoadedImage(C.byref(b))
    rng = np.random.default_rng(8)
    w, h = 37, 23
    px = (rng.standard_normal((h, w, 4)) * 3.0).astype(np.float32)
    px[0, 0] = (0.0, -0.0, 65504.0, 1e-30)
    mine, theirs = str(tmp_path / "host.exr"), str(tmp_path / "reference.exr")
    lib.vkrtWriteEXRFromRGBA32F.restype = C.c_int
    assert lib.vkrtWriteEXRFromRGBA32F(mine.encode(), px.ctypes.data_as(C.c_void_p), w, h) == 1
    assert ref.vkrtWriteEXRFromRGBA32F(theirs.encode(), px.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(h)) == 1
    a, b = _LoadedImage(), _LoadedImage()
    blob = open(mine, "rb").read()    # (the reference's file reader opens paths relative to the executable: hand it the bytes instead)
    assert ref.vkrtLoadEXRImageFromMemory(blob, C.c_size_t(len(blob)), b"host.exr", C.byref(a)) == 1   # host writer -> reference reader
    assert (a.width, a.height, a.format) == (w, h, 3)
    assert np.array_equal(_pixels_of(a).view(np.uint32), px.reshape(-1).view(np.uint32))
    assert lib.vkrtLoadImageFromFile(theirs.encode(), C.c_uint32(1), C.byref(b)) == 1        # reference writer -> host reader
    assert (b.width, b.height) == (w, h)
    got = _pixels_of(b)
    want = px.reshape(-1) if b.format == 3 else px.astype(np.float16).reshape(-1).view(np.uint16)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    lib.vkrtFreeLoadedImage(C.byref(b))


# ======================================================================================================================
# data tables the device code embeds
# ======================================================================================================================
def test_sheen_ltc_table_of_the_device_code_is_the_reference_table():
    """csrc/data/sheen_ltc.inc (what k_shade's sheen lobe reads) and oracle/sheen_ltc.inc are one file, and — where the reference checkout is
    present — its 3072 values are the ones of src/shaders/bsdf/data/sheen_ltc.slang, value by value in fp32."""
    import re
    dev = open(os.path.join(H.ROOT, "vkrt_b200", "csrc", "data", "sheen_ltc.inc")).read()
    assert dev == open(os.path.join(H.ROOT, "oracle", "sheen_ltc.inc")).read()
    mine = np.array([float(v.rstrip("f")) for v in re.findall(r"-?\d+\.\d+f?", "\n".join(ln for ln in dev.splitlines() if not ln.startswith("//")))], np.float32)
    assert mine.shape == (3072,)
    slang = os.path.join(refpin.REFERENCE, "src", "shaders", "bsdf", "data", "sheen_ltc.slang")
    if not os.path.exists(slang):
        pytest.skip("reference checkout not present")
    src = open(slang).read()
    body = src[src.index("SHEEN_LTC_TABLE[3072]"):]
    body = body[body.index("{") + 1: body.index("}")]
    ref = np.array([float(v.rstrip("f")) for v in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?f?", body)], np.float32)
    assert ref.shape == (3072,) and np.array_equal(ref, mine)


def test_device_closure_code_carries_the_oracles_constants():
    """csrc/shading.cuh is the CUDA statement of the closures whose oracle restatement (oracle/shading.h) is pinned bit for bit to the reference
    shaders above. The GPU tests compare the two per call within 1e-4, which a mistyped constant of a rarely taken branch can survive; this
    compares the floating-point literals of the two sources: the same set of distinct constants, each the same number of times (the
    trivial 0 / 1 aside, which initialisers add freely)."""
    import collections
    import re

    def literals(path):
        s = open(path).read()
        s = re.sub(r"//.*", "", s)
        s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
        return collections.Counter(re.findall(r"(?<![\w.])(?:\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)f?", s))
    dev = literals(os.path.join(H.ROOT, "vkrt_b200", "csrc", "shading.cuh"))
    ora = literals(os.path.join(H.ROOT, "oracle", "shading.h"))
    assert len(ora) > 100 and set(dev) == set(ora), (sorted(set(dev) - set(ora)), sorted(set(ora) - set(dev)))
    for k in ora:
        if k not in ("0.0f", "1.0f"):
            assert dev[k] == ora[k], (k, dev[k], ora[k])
