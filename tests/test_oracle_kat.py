"""Known-answer tests that pin the oracle's integer paths and its bit-level helpers."""
import ctypes as C

import numpy as np

import harness as H


def test_hash_kat():
    lib = H.oracle_lib()
    assert lib.oracle_hash(C.c_uint32(0)) == 0x00000000
    assert lib.oracle_hash(C.c_uint32(1)) == 0x688990C0
    assert lib.oracle_hash(C.c_uint32(0xFFFFFFFF)) == 0x6768824A


def _rand3(seed):
    lib = H.oracle_lib()
    rng = C.c_uint32(seed)
    states, vals = [], []
    for _ in range(3):
        v = lib.oracle_rand(C.byref(rng))
        states.append(rng.value)
        vals.append(v)
    return states, vals


def test_pixel_seed_and_rand_kat():
    lib = H.oracle_lib()
    cases = [
        ((0, 0, 0, 0), 0x00000000, [0x01FCE552, 0xA0117BE9, 0x00E15A61]),
        ((1, 0, 0, 0), 0x9D5A7ACD, [0xA37305AE, 0x74AE8B35, 0xF95548BC]),
        ((0, 1, 0, 0), 0x2859CC20, [0x734089E9, 0x956A12AC, 0x8CF5658D]),
        ((255, 255, 0, 0), 0xA910C697, [0xFFBC6532, 0x60D3E804, 0xE9F67FD3]),
        ((255, 255, 3, 17), 0x31828193, [0xCE0E932D, 0xB1BDA95D, 0x9216D81B]),
        ((1919, 1079, 63, 1023), 0x78876E17, [0x1D631EBD, 0x2EBA6B36, 0xBB7E2BD4]),
    ]
    for (x, y, f, s), seed, states in cases:
        got = lib.oracle_init_pixel_seed(C.c_int(x), C.c_int(y), C.c_uint32(f), C.c_uint32(s))
        assert got == seed, (x, y, f, s, hex(got))
        st, vals = _rand3(seed)
        assert st == states
        for state, v in zip(st, vals):
            assert v == np.float32(state & 0x00FFFFFF) / np.float32(16777216.0)
    _, vals = _rand3(0)
    assert np.allclose(vals, [0.98787415, 0.0682969689, 0.880285323], rtol=0, atol=1e-9)


def test_reverse_bits_kat():
    lib = H.oracle_lib()
    assert lib.oracle_reverse_bits(C.c_uint32(1)) == 0x80000000
    assert lib.oracle_reverse_bits(C.c_uint32(6)) == 0x60000000
    for v in (0, 0xFFFFFFFF, 0x12345678, 0xDEADBEEF):
        assert lib.oracle_reverse_bits(C.c_uint32(v)) == int(format(v, "032b")[::-1], 2)


def test_half_conversion_matches_numpy():
    lib = H.oracle_lib()
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.normal(size=2000).astype(np.float32), (rng.normal(size=2000) * 1e-6).astype(np.float32),
                           np.array([0, -0.0, 1, -1, 65504, 65520, 1e-8, 6.1e-5, 5.96e-8, 2.98e-8, 3e-8, np.inf, -np.inf], dtype=np.float32)])
    for v in vals:
        got = lib.oracle_f32_to_f16(C.c_float(float(v)))
        want = int(np.float32(v).astype(np.float16).view(np.uint16))
        assert got == want, (v, hex(got), hex(want))
    for h in list(range(0, 0x7C00, 37)) + list(range(0x8000, 0xFC00, 41)):
        got = lib.oracle_f16_to_f32(C.c_uint16(h))
        want = float(np.array([h], dtype=np.uint16).view(np.float16)[0])
        assert got == want


def test_pack_unpack_round_trip():
    """packShaderVertex (packing.c:92-156) -> unpackOctNormal/unpackOctTangent (packing.slang:13-60)."""
    lib = H.oracle_lib()
    hr = H.hr
    rng = np.random.default_rng(7)
    n = rng.normal(size=(500, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    packed = hr.pack_oct_normal32(n)
    out = np.zeros(3, dtype=np.float32)
    for i in range(len(n)):
        lib.oracle_unpack_normal(C.c_uint32(int(packed[i])), out.ctypes.data_as(C.c_void_p))
        assert abs(np.linalg.norm(out) - 1) < 1e-5
        assert np.dot(out, n[i]) > 1 - 2e-8 * 32767  # snorm16 oct: error well below 1e-4 rad^2
        assert np.max(np.abs(out - n[i])) < 1e-4
    t = np.concatenate([n, np.where(rng.random((500, 1)) < 0.5, -1.0, 1.0).astype(np.float32)], axis=1)
    pt = hr.pack_tangent32(t)
    out4 = np.zeros(4, dtype=np.float32)
    for i in range(len(t)):
        lib.oracle_unpack_tangent(C.c_uint32(int(pt[i])), out4.ctypes.data_as(C.c_void_p))
        assert out4[3] == t[i, 3]
        assert np.max(np.abs(out4[:3] - t[i, :3])) < 3e-4
    c = rng.random((100, 4)).astype(np.float32)
    pc = hr.pack_color_rgba8(c)
    for i in range(len(c)):
        q = [(int(pc[i]) >> (8 * k)) & 0xFF for k in range(4)]
        assert np.max(np.abs(np.array(q) / 255.0 - c[i])) <= 0.5 / 255 + 1e-7


def test_alias_table_reconstructs_pmf():
    hr = H.hr
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 960):
        w = rng.random(n).astype(np.float32) + 0.01
        pmf = (w / w.sum()).astype(np.float32)
        q, idx = hr.build_alias_table(pmf)
        rec = np.zeros(n)
        for i in range(n):
            rec[i] += q[i] / n
            rec[idx[i]] += (1.0 - q[i]) / n
        assert np.allclose(rec, pmf, atol=2e-6)


def test_watertight_shared_edge_never_leaks():
    """Rays aimed exactly at the shared edge/vertices of two triangles must hit at least one of them."""
    lib = H.oracle_lib()
    rng = np.random.default_rng(11)
    tuv = np.zeros(3, dtype=np.float32)
    misses = 0
    for _ in range(2000):
        a, b, c, d = (rng.normal(size=3).astype(np.float32) for _ in range(4))
        t = np.float32(rng.random())
        target = (a * (1 - t) + b * t).astype(np.float32)  # on the shared edge a-b
        org = (target + rng.normal(size=3).astype(np.float32) * 3).astype(np.float32)
        dr = (target - org).astype(np.float32)
        hit = 0
        for tri in ((a, b, c), (b, a, d)):
            hit |= lib.oracle_watertight(org.ctypes.data_as(C.c_void_p), dr.ctypes.data_as(C.c_void_p), tri[0].ctypes.data_as(C.c_void_p),
                                         tri[1].ctypes.data_as(C.c_void_p), tri[2].ctypes.data_as(C.c_void_p), tuv.ctypes.data_as(C.c_void_p))
        # c and d must be on opposite sides of the edge as seen from the ray for the pair to tile the neighbourhood
        n = np.cross(b - a, dr)
        if np.dot(c - a, n) * np.dot(d - a, n) < 0 and not hit:
            misses += 1
    assert misses == 0


def test_bvh_matches_brute_force():
    hr = H.hr
    scene = hr.cornell_scene(sphere_segments=16)
    W = Hh = 96
    prep = scene.prepare(W, Hh)
    a, b = H.OracleBackend(brute_force=False), H.OracleBackend(brute_force=True)
    for o in (a, b):
        o.upload(prep)
        o.resize(W, Hh)
        o.trace_primary(prep["sceneData"])
    for which in (H.AOV_HITID_CENTER, H.AOV_HITID_S0):
        assert np.array_equal(a.read(which), b.read(which))
    assert np.array_equal(a.read(H.AOV_HIT_TUV).view(np.uint32), b.read(H.AOV_HIT_TUV).view(np.uint32))
    # random rays, closest + any hit
    rng = np.random.default_rng(5)
    o_ = rng.uniform(-0.9, 0.9, (4000, 3)).astype(np.float32)
    d_ = rng.normal(size=(4000, 3)).astype(np.float32)
    rays = np.concatenate([o_, np.full((4000, 1), 1e-3, np.float32), d_, np.full((4000, 1), 1e4, np.float32)], axis=1)
    assert np.array_equal(a.trace_rays(rays), b.trace_rays(rays))
    rays[:, 7] = 0.7
    assert np.array_equal(a.trace_rays(rays, any_hit=True), b.trace_rays(rays, any_hit=True))
