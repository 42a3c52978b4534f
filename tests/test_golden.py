"""Golden fixtures (tests/golden/cornell_64x36.npz, written by tests/golden/make_golden.py from the oracle): the oracle must keep
reproducing them (CPU), the C host must reproduce the host-side arrays (CPU), and the CUDA path must match them on the GPU box —
there nothing reads /root/reference or regenerates expectations."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
import scenes

hr = H.hr
G = np.load(os.path.join(H.ROOT, "tests", "golden", "cornell_64x36.npz"))
W, Hh = 64, 36
SCENE = os.path.join(H.ROOT, "assets", "scenes", "cornell.json")


def test_integer_known_answers_match_survey_appendix():
    k = G["kat_pixel_seed"]
    assert [hex(int(v)) for v in k[:, 4]] == ["0x0", "0x9d5a7acd", "0x2859cc20", "0xa910c697", "0x31828193", "0x78876e17"]
    assert [hex(int(v)) for v in k[0, 5:]] == ["0x1fce552", "0xa0117be9", "0xe15a61"]
    lib = H.oracle_lib()
    for x, y, f, s, seed, a, b, c in k.tolist():
        assert lib.oracle_init_pixel_seed(C.c_int(x), C.c_int(y), C.c_uint32(f), C.c_uint32(s)) == seed
        rng = C.c_uint32(seed)
        for want in (a, b, c):
            lib.oracle_rand(C.byref(rng))
            assert rng.value == want


def test_oracle_reproduces_golden_render():
    prep = hr.load_scene_json(SCENE).prepare(W, Hh)
    prep["sceneData"]["samplesPerPixel"] = 2
    o = H.OracleBackend()
    o.upload(prep)
    o.resize(W, Hh)
    o.trace_primary(prep["sceneData"])
    assert np.array_equal(o.read(H.AOV_HITID_CENTER), G["ids_center"])
    assert np.array_equal(o.read(H.AOV_HITID_S0), G["ids_s0"])
    assert np.array_equal(o.read(H.AOV_HIT_TUV).view(np.uint32), G["tuv_center"].view(np.uint32))
    o.render(prep["sceneData"], frames=2)
    acc = o.read(H.AOV_ACCUM)
    assert np.array_equal(acc[..., 3], G["accum_rgb"][..., 3])
    assert np.allclose(acc, G["accum_rgb"], rtol=1e-5, atol=1e-6)  # same code, same libm: identical up to compiler version


def test_c_host_reproduces_golden_scene_arrays():
    from vkrt_b200 import host
    hs = host.Host(width=W, height=Hh, host_only=True)
    hs.load_scene(SCENE)
    hs.set_samples_per_pixel(2)
    hs.start_render(W, Hh, 4)
    p = hs.prepare_scene()
    v = p["vertices"]
    sums = np.array([np.bitwise_xor.reduce(v["packedNormal"]), np.bitwise_xor.reduce(v["packedTangent"]),
                     int(v["packedNormal"].astype(np.uint64).sum() & 0xFFFFFFFF), int(v["packedTangent"].astype(np.uint64).sum() & 0xFFFFFFFF),
                     len(v), len(p["indices"]), int(p["indices"].astype(np.uint64).sum() & 0xFFFFFFFF)], dtype=np.uint64)
    assert np.array_equal(sums, G["host_vertex_checksums"])                      # quantised vertex streams: checksum of checksums
    assert p["materials"].tobytes() == G["host_materials"].tobytes()
    assert np.array_equal(p["triAliasIdx"], G["host_triAliasIdx"])
    assert np.allclose(p["triAliasQ"], G["host_triAliasQ"], rtol=1e-4, atol=1e-5)
    assert np.allclose(p["world3x4"], G["host_world3x4"], rtol=1e-6, atol=1e-6)
    gi = p["meshInfos"]
    ri = np.frombuffer(G["host_meshInfos"].tobytes(), dtype=hr.MESH_INFO)
    for key in ("vertexBase", "vertexCount", "indexBase", "indexCount", "materialIndex", "renderBackfaces"):
        assert np.array_equal(gi[key], ri[key]), key
    gsd = np.frombuffer(p["sceneData"].tobytes(), dtype=hr.SCENE_DATA)[0]
    rsd = np.frombuffer(G["host_sceneData"].tobytes(), dtype=hr.SCENE_DATA)[0]
    assert np.allclose(gsd["viewInverse"], rsd["viewInverse"], rtol=2e-5, atol=2e-6)
    assert int(gsd["emissiveTriangleCount"]) == int(rsd["emissiveTriangleCount"]) == 960
    hs.close()


def _cuda_render(mode, sampling):
    """The product path end to end: C host (scene file + glb ingest + preparation) -> C ABI -> CUDA."""
    import vkrt_b200
    from vkrt_b200 import host
    hs = host.Host(width=W, height=Hh)
    hs.load_scene(SCENE)
    if mode:
        hs.set_render_mode(1)
        hs.set_spectral_sampling_mode(sampling)
        scenes.rgb2spec()
        hs.load_rgb2spec(os.path.join(H.ROOT, "assets", "rgb2spec", "srgb.coeff"))
    hs.set_samples_per_pixel(2)
    hs.start_render(W, Hh, 4)
    hs.draw()
    hs.draw()
    lib = vkrt_b200.load_library()
    ctx = C.c_void_p(hs.cuda_context())

    def read(which, dt, nc):
        out = np.zeros((Hh, W, nc), dt)
        assert lib.vkrt_cuda_read_aov(ctx, C.c_int(which), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)) == 0
        return out
    res = dict(accum=read(0, np.float32, 4), ids=read(4, np.uint32, 2), ids_s0=read(5, np.uint32, 2), tuv=read(6, np.float32, 3), output=read(3, np.uint16, 4))
    hs.close()
    return res


@pytest.mark.gpu
def test_cuda_matches_golden_primary_hits_bit_exact():
    """Same input arrays as the fixture (numpy host restatement) through the C ABI: ids and the t/u/v bits are identical."""
    prep = hr.load_scene_json(SCENE).prepare(W, Hh)
    g = H.CudaBackend()
    g.upload(prep)
    g.resize(W, Hh)
    g.trace_primary(prep["sceneData"])
    assert np.array_equal(g.read(H.AOV_HITID_CENTER), G["ids_center"])
    assert np.array_equal(g.read(H.AOV_HITID_S0), G["ids_s0"])
    assert np.array_equal(g.read(H.AOV_HIT_TUV).view(np.uint32), G["tuv_center"].view(np.uint32))
    g.close()


@pytest.mark.gpu
def test_c_host_pipeline_matches_golden_primary_hits():
    """The whole product path (C host ingest -> C ABI -> CUDA). The C host's world matrices differ from the fixture's by an ulp
    (sinf/cosf vs numpy), so a pixel exactly on a shared triangle edge may pick the neighbour: >= 99.9 % identical ids
    (the north_star bar), t within 1e-4 relative everywhere the ids agree."""
    r = _cuda_render(0, 0)
    # jittered rays are in general position: the north_star bar applies as is
    assert np.all(r["ids_s0"] == G["ids_s0"], axis=-1).mean() >= 0.999
    # un-jittered pixel centres of this 64x36 symmetric box fall EXACTLY on quad diagonals and wall/floor seams (11 of 2304 pixels,
    # all with equal t on both sides or at a silhouette); they may flip with an ulp of the matrices
    same = np.all(r["ids"] == G["ids_center"], axis=-1)
    assert same.mean() >= 0.99, int((~same).sum())
    t, tg = r["tuv"][..., 0][same], G["tuv_center"][..., 0][same]
    assert np.allclose(t, tg, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode,sampling,frac,rel", [("accum_rgb", 0, 0, 0.99, 0.03), ("accum_single", 1, 0, 0.98, 0.08), ("accum_hero", 1, 1, 0.98, 0.08)])
def test_cuda_matches_golden_images(name, mode, sampling, frac, rel):
    r = _cuda_render(mode, sampling)
    want = G[name]
    assert np.array_equal(r["accum"][..., 3], want[..., 3])
    c = H.compare_images(want[..., :3], r["accum"][..., :3])
    assert 1.0 - c["frac_rel_gt_1e3"] >= frac and c["rmse"] <= rel * max(c["mean_a"], 1e-6), c
    if name == "accum_rgb":
        d = np.abs(r["output"].astype(np.int64) - G["output_rgb"].astype(np.int64))
        assert (d > 64).mean() < 0.02
