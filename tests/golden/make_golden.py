#!/usr/bin/env python3
"""Generates the golden fixtures in this directory from the ORACLE (oracle/oracle.cpp + oracle/host_ref.py).

The reference ships no golden vectors and cannot be run here, so these fixtures do not pin the oracle against
upstream; they pin (a) the oracle against drift, (b) the C host against the numpy restatement, and (c) the CUDA path against fixed
expected outputs that travel to the GPU box. Re-run only on purpose:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
import scenes  # noqa: E402

hr = H.hr
W, Hh = 64, 36


def main():
    out = {}
    sc = hr.load_scene_json(os.path.join(H.ROOT, "assets", "scenes", "cornell.json"))
    prep = sc.prepare(W, Hh)
    prep["sceneData"]["samplesPerPixel"] = 2
    # host-side goldens (small ones: light tables, MeshInfo, SceneData; the 1.7 MB vertex array is covered by a checksum)
    out["host_meshInfos"] = prep["meshInfos"].view(np.uint8)
    out["host_world3x4"] = prep["world3x4"]
    out["host_sceneData"] = np.frombuffer(prep["sceneData"].tobytes(), np.uint8)
    out["host_materials"] = np.frombuffer(np.ascontiguousarray(prep["materials"]).tobytes(), np.uint8)
    L = prep["lights"]
    out["host_emissiveMeshes"] = L["meshes"][:L["meshCount"]].view(np.uint8)
    out["host_triAliasIdx"] = L["triAliasIdx"][:L["triangleCount"]]
    out["host_triAliasQ"] = L["triAliasQ"][:L["triangleCount"]]
    v = prep["vertices"]
    out["host_vertex_checksums"] = np.array([np.bitwise_xor.reduce(v["packedNormal"]), np.bitwise_xor.reduce(v["packedTangent"]),
                                             int(v["packedNormal"].astype(np.uint64).sum() & 0xFFFFFFFF), int(v["packedTangent"].astype(np.uint64).sum() & 0xFFFFFFFF),
                                             len(v), len(prep["indices"]), int(prep["indices"].astype(np.uint64).sum() & 0xFFFFFFFF)], dtype=np.uint64)
    # render goldens
    o = H.OracleBackend()
    o.upload(prep)
    o.resize(W, Hh)
    o.trace_primary(prep["sceneData"])
    out["ids_center"] = o.read(H.AOV_HITID_CENTER)
    out["ids_s0"] = o.read(H.AOV_HITID_S0)
    out["tuv_center"] = o.read(H.AOV_HIT_TUV)
    o.render(prep["sceneData"], frames=2)
    out["accum_rgb"] = o.read(H.AOV_ACCUM)
    out["output_rgb"] = o.read(H.AOV_OUTPUT)
    for name, sampling in (("single", 0), ("hero", 1)):
        sc2 = hr.load_scene_json(os.path.join(H.ROOT, "assets", "scenes", "cornell.json"))
        sc2.settings.render_mode, sc2.settings.spectral_sampling = 1, sampling
        p2 = sc2.prepare(W, Hh)
        p2["sceneData"]["samplesPerPixel"] = 2
        o2 = H.OracleBackend()
        o2.upload(p2, rgb2spec=scenes.rgb2spec())
        o2.resize(W, Hh)
        o2.render(p2["sceneData"], frames=2)
        out["accum_" + name] = o2.read(H.AOV_ACCUM)
    # integer known answers so that they also exist as data
    lib = H.oracle_lib()
    import ctypes as C
    seeds = []
    for x, y, f, s in ((0, 0, 0, 0), (1, 0, 0, 0), (0, 1, 0, 0), (255, 255, 0, 0), (255, 255, 3, 17), (1919, 1079, 63, 1023)):
        seed = lib.oracle_init_pixel_seed(C.c_int(x), C.c_int(y), C.c_uint32(f), C.c_uint32(s))
        rng = C.c_uint32(seed)
        st = []
        for _ in range(3):
            lib.oracle_rand(C.byref(rng))
            st.append(rng.value)
        seeds.append([x, y, f, s, seed] + st)
    out["kat_pixel_seed"] = np.array(seeds, dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "cornell_64x36.npz"), **out)
    print("wrote", os.path.join(HERE, "cornell_64x36.npz"), {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
