"""The C host library (libvkrt_host.so: VKRT_* API, vkrt.scene + .glb ingest, scene preparation) against the oracle-side numpy
restatement of the same reference code (oracle/host_ref.py), array for array. Runs without a GPU (hostOnly handles).

Integer / byte work (packed normals, tangents, colours, indices, bases, dedup, alias indices, flags) must match bit for bit.
fp32 results that pass through libm transcendentals (sin/cos of Euler angles, tan of the field of view, atan2/asin in the
transform decomposition) are compared with a few-ulp tolerance, stated per test."""
import os

import numpy as np
import pytest

import harness as H

hr = H.hr
ASSETS = os.path.join(H.ROOT, "assets")


def _host(**kw):
    from vkrt_b200 import host
    return host.Host(host_only=True, **kw)


def _ulp_close(a, b, ulps=4, floor=1e-6):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    tol = np.maximum(np.abs(a), np.abs(b)) * np.float32(ulps * 1.2e-7) + np.float32(floor)
    return bool(np.all(np.abs(a.astype(np.float64) - b.astype(np.float64)) <= tol))


@pytest.mark.parametrize("scene", ["cornell", "prism", "caustics"])
def test_scene_file_matches_reference_restatement(scene):
    path = os.path.join(ASSETS, "scenes", scene + ".json")
    w, h = 320, 180
    hs = _host(width=w, height=h)
    hs.load_scene(path)
    hs.start_render(w, h, 64)
    got = hs.prepare_scene()
    ref = hr.load_scene_json(path).prepare(w, h)

    # geometry: byte-exact (positions/uvs are copied, normals/tangents/colours are quantised integers)
    assert got["vertices"].tobytes() == ref["vertices"].tobytes()
    assert np.array_equal(got["indices"], ref["indices"])
    assert np.array_equal(got["geometrySource"], ref["geometrySource"])
    assert np.array_equal(got["alphaTested"], ref["alphaTested"])
    gi, ri = got["meshInfos"], ref["meshInfos"]
    for key in ("vertexBase", "vertexCount", "indexBase", "indexCount", "materialIndex", "renderBackfaces"):
        assert np.array_equal(gi[key], ri[key]), key
    assert np.array_equal(gi["opacity"], ri["opacity"])
    # transforms: composed from Euler angles with sinf/cosf -> a few ulp; decomposed Euler angles in degrees -> 1e-3 degrees
    assert _ulp_close(got["world3x4"], ref["world3x4"], ulps=8)
    assert np.allclose(gi["position"], ri["position"], rtol=0, atol=1e-6)
    assert np.allclose(gi["scale"], ri["scale"], rtol=1e-5, atol=1e-6)
    d = np.abs(((gi["rotation"] - ri["rotation"]) + 180.0) % 360.0 - 180.0)
    assert float(d.max()) < 2e-3
    # materials: sanitised copies of the JSON values -> byte-exact
    assert got["materials"].tobytes() == np.ascontiguousarray(ref["materials"]).tobytes()
    # lights
    L = ref["lights"]
    assert len(got["emissiveMeshes"]) == L["meshCount"] and len(got["emissiveTriangles"]) == L["triangleCount"]
    if L["meshCount"]:
        for key in ("triOffset", "triCount"):
            assert np.array_equal(got["emissiveMeshes"][key], L["meshes"][key][:L["meshCount"]])
        assert _ulp_close(got["emissiveMeshes"]["pmfMesh"], L["meshes"]["pmfMesh"][:L["meshCount"]], ulps=64)
        assert _ulp_close(got["emissiveMeshes"]["invTotalArea"], L["meshes"]["invTotalArea"][:L["meshCount"]], ulps=64)
        assert np.array_equal(got["emissiveMeshes"]["emission"], L["meshes"]["emission"][:L["meshCount"]])
        for key in ("v0Area", "e1Pad", "e2Pad"):
            assert np.allclose(got["emissiveTriangles"][key], L["triangles"][key][:L["triangleCount"]], rtol=2e-5, atol=2e-6), key
        assert np.allclose(gi["lightPdfArea"], ri["lightPdfArea"], rtol=1e-4, atol=0)
        # alias tables: same structure whenever the pmf inputs agree to rounding; compare the distributions they encode
        for q, idx, rq, ridx, n in ((got["triAliasQ"], got["triAliasIdx"], L["triAliasQ"], L["triAliasIdx"], L["triangleCount"]),):
            em = got["emissiveMeshes"][0]
            cnt = int(em["triCount"])

            def pmf(qq, ii):
                p = np.zeros(cnt)
                np.add.at(p, np.arange(cnt), qq[:cnt].astype(np.float64) / cnt)
                np.add.at(p, ii[:cnt].astype(np.int64), (1.0 - qq[:cnt].astype(np.float64)) / cnt)
                return p
            assert np.allclose(pmf(q, idx), pmf(rq, ridx), rtol=0, atol=1e-6)
    # SceneData: integers exact, camera matrices to a few ulp (tanf, normalisation)
    gsd = np.frombuffer(got["sceneData"].tobytes(), dtype=hr.SCENE_DATA)[0]
    rsd = ref["sceneData"]
    for key in ("rrMaxDepth", "rrMinDepth", "packedRenderSettings", "environmentTextureIndex", "debugMode", "misNeeEnabled", "emissiveMeshCount",
                "emissiveTriangleCount"):
        assert int(gsd[key]) == int(rsd[key]), key
    assert np.array_equal(gsd["viewportRect"], rsd["viewportRect"])
    assert np.allclose(gsd["environmentLight"], rsd["environmentLight"], rtol=1e-6)
    assert np.allclose(gsd["viewInverse"], rsd["viewInverse"], rtol=2e-5, atol=2e-6)
    assert np.allclose(gsd["projInverse"], rsd["projInverse"], rtol=2e-5, atol=1e-4)
    hs.close()


def test_dedup_shares_geometry_across_imports():
    hs = _host()
    hs.load_scene(os.path.join(ASSETS, "scenes", "cornell.json"))
    got = hs.prepare_scene()
    gs = got["geometrySource"]
    # cornell: plane.glb x5 -> one owner, sphere.glb x2 -> one owner, bunny -> its own
    assert len(gs) == 8 and len(set(gs.tolist())) == 3
    assert len(got["vertices"]) == 4 + 559 + 34834
    hs.close()


def test_setters_clamp_like_the_reference():
    from vkrt_b200 import VkrtError
    hs = _host()
    hs.set_path_depth(9, 200)           # settings.c:41-57: max clamped to 64, min <= max
    s = hs.scene_settings()
    assert (s.rrMinDepth, s.rrMaxDepth) == (9, 64)
    hs.set_path_depth(100, 0)
    s = hs.scene_settings()
    assert (s.rrMinDepth, s.rrMaxDepth) == (1, 1)
    hs.set_samples_per_pixel(0)
    assert hs.scene_settings().samplesPerPixel == 1
    with pytest.raises(VkrtError):
        hs.set_render_mode(7)
    with pytest.raises(VkrtError):
        hs.set_debug_mode(99)
    hs.set_environment_light((float("nan"), -1.0, 2.0), float("inf"))
    s = hs.scene_settings()
    assert list(s.environmentColor) == [1.0, 0.0, 2.0] and s.environmentStrength == 0.0  # non-finite -> fallback (numeric.h:5-17)
    hs.close()


def test_render_session_protocol_without_device():
    """startRender resets the accumulation state; tracing needs a device and fails loudly on a hostOnly handle."""
    from vkrt_b200 import VkrtError
    hs = _host()
    hs.generate_soup(300)
    hs.start_render(64, 48, 8)
    st = hs.render_status()
    assert st.renderPhase == 1 and st.totalSamples == 0 and st.renderTargetSamples == 8
    hs.update_scene()
    with pytest.raises(VkrtError):
        hs.draw()
    hs.close()


def test_procedural_scenes_are_deterministic_and_well_formed():
    a, b = _host(), _host()
    for hs in (a, b):
        hs.generate_soup(1000)
    pa, pb = a.prepare_scene(), b.prepare_scene()
    assert pa["vertices"].tobytes() == pb["vertices"].tobytes()
    assert len(pa["indices"]) == 3000 + 6 and len(pa["meshInfos"]) == 17
    assert int(pa["meshInfos"]["indexCount"][:16].sum()) == 3000
    assert len(pa["emissiveTriangles"]) == 2
    a.close(); b.close()
    hs = _host()
    hs.generate_instanced(os.path.join(ASSETS, "models", "suzanne.glb"), 27)
    p = hs.prepare_scene()
    assert len(p["meshInfos"]) == 27 + 2
    assert len(set(p["geometrySource"][:27].tolist())) == 1   # one BLAS shared by every instance
    assert len(p["indices"]) == 62976 * 3 + 12
    hs.close()


def test_alias_table_reconstructs_pmf():
    """buildAliasTable exactness (SURVEY §8c iii): the table encodes exactly the input distribution."""
    import ctypes as C
    from vkrt_b200 import host
    lib = host.load_host_library()
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 960, 4099):
        pmf = rng.uniform(0.01, 1.0, n).astype(np.float32)
        pmf /= pmf.sum(dtype=np.float32)
        q, idx = hr.build_alias_table(pmf)
        cq, cidx = np.zeros(n, np.float32), np.zeros(n, np.uint32)
        assert lib.VKRT_hostBuildAliasTable(pmf.ctypes.data_as(C.c_void_p), C.c_uint32(n), cq.ctypes.data_as(C.c_void_p), cidx.ctypes.data_as(C.c_void_p)) == 1
        assert np.array_equal(cq.view(np.uint32), q.view(np.uint32)) and np.array_equal(cidx, idx)   # C host == restatement, bit for bit
        rec = np.zeros(n)
        np.add.at(rec, np.arange(n), q.astype(np.float64) / n)
        np.add.at(rec, idx.astype(np.int64), (1.0 - q.astype(np.float64)) / n)
        assert np.allclose(rec, pmf, atol=2e-6)


def test_glb_with_an_out_of_range_index_fails_to_import(tmp_path):
    """src/app/mesh/loader.c:1855-1862: an index >= vertexCount fails the primitive and with it the import (it would otherwise become an
    out-of-bounds vertex fetch on the device)."""
    import json
    import struct
    from vkrt_b200 import host
    raw = bytearray(open(os.path.join(H.ROOT, "assets", "models", "cube.glb"), "rb").read())
    json_len = struct.unpack_from("<I", raw, 12)[0]
    doc = json.loads(raw[20:20 + json_len].decode())
    acc = doc["accessors"][doc["meshes"][0]["primitives"][0]["indices"]]
    view = doc["bufferViews"][acc["bufferView"]]
    at = 20 + json_len + 8 + view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    width = {5121: 1, 5123: 2, 5125: 4}[acc["componentType"]]
    raw[at:at + width] = b"\xff" * width
    bad = tmp_path / "bad_index.glb"
    bad.write_bytes(bytes(raw))
    hs = host.Host(host_only=True, width=64, height=36)
    with pytest.raises(Exception):
        hs.import_mesh(str(bad))
    assert hs.mesh_count() == 0
    hs.import_mesh(os.path.join(H.ROOT, "assets", "models", "cube.glb"))   # the unmodified file still imports
    assert hs.mesh_count() == 1
    hs.close()


@pytest.mark.parametrize("edit", ["count_2_pow_60", "negative_offset", "stride_below_element", "nan_count", "offset_past_end"])
def test_glb_with_a_hostile_accessor_is_refused(tmp_path, edit):
    """accessor count / byteStride / byteOffset arrive as JSON doubles; a count of 2^60 with stride 16 used to wrap the
    bounds product, allocate 0 bytes and overflow the heap in the vertex loop. Every such file must fail to import, cleanly."""
    import json
    import struct
    from vkrt_b200 import host
    raw = open(os.path.join(H.ROOT, "assets", "models", "cube.glb"), "rb").read()
    json_len = struct.unpack_from("<I", raw, 12)[0]
    doc = json.loads(raw[20:20 + json_len].decode())
    prim = doc["meshes"][0]["primitives"][0]
    acc = doc["accessors"][prim["attributes"]["POSITION"]]
    view = doc["bufferViews"][acc["bufferView"]]
    if edit == "count_2_pow_60":
        acc["count"] = 2 ** 60
        view["byteStride"] = 16
    elif edit == "negative_offset":
        acc["byteOffset"] = -64
    elif edit == "stride_below_element":
        view["byteStride"] = 4
    elif edit == "nan_count":
        acc["count"] = 1e400     # json.dumps writes Infinity
    else:
        view["byteOffset"] = 2 ** 40
    text = json.dumps(doc).encode()
    text += b" " * (-len(text) % 4)
    rest = raw[20 + json_len:]
    out = bytearray(raw[:12]) + struct.pack("<I", len(text)) + b"JSON" + text + rest
    struct.pack_into("<I", out, 8, len(out))
    bad = tmp_path / (edit + ".glb")
    bad.write_bytes(bytes(out))
    hs = host.Host(host_only=True, width=64, height=36)
    with pytest.raises(Exception):
        hs.import_mesh(str(bad))
    assert hs.mesh_count() == 0
    hs.import_mesh(os.path.join(H.ROOT, "assets", "models", "cube.glb"))
    assert hs.mesh_count() == 1
    hs.close()


@pytest.mark.parametrize("edit", ["negative_mesh_material", "fractional_material_index", "huge_material_index", "missing_mesh_material"])
def test_scene_file_with_a_bad_material_reference_is_refused(tmp_path, edit):
    """controller.c:368-377,905-927: material indices in a scene file must be non-negative integers below 2^32 and every mesh must carry one;
    a negative index used to wrap to 4 billion and grow the material list until memory ran out. Indices past 2^20 are refused outright."""
    import json
    import shutil
    from vkrt_b200 import host
    doc = json.load(open(os.path.join(H.ROOT, "assets", "scenes", "prism.json")))
    if edit == "negative_mesh_material":
        doc["meshes"][0]["materialIndex"] = -2
    elif edit == "fractional_material_index":
        doc["materials"][0]["index"] = 1.5
    elif edit == "huge_material_index":
        doc["meshes"][0]["materialIndex"] = 4000000000
    else:
        del doc["meshes"][0]["materialIndex"]
    (tmp_path / "scenes").mkdir()
    shutil.copytree(os.path.join(H.ROOT, "assets", "models"), str(tmp_path / "models"))
    bad = tmp_path / "scenes" / "bad.json"
    bad.write_text(json.dumps(doc))
    hs = host.Host(host_only=True, width=64, height=36)
    with pytest.raises(Exception):
        hs.load_scene(str(bad))
    hs.close()
    good = tmp_path / "scenes" / "good.json"
    good.write_text(json.dumps(json.load(open(os.path.join(H.ROOT, "assets", "scenes", "prism.json")))))
    hs = host.Host(host_only=True, width=64, height=36)
    hs.load_scene(str(good))
    assert hs.mesh_count() > 0
    hs.close()


@pytest.mark.parametrize("bad", [float("nan"), float("inf"), 3.0e38], ids=["nan", "inf", "huge"])
def test_light_tables_stay_finite_with_a_non_finite_emissive_vertex(tmp_path, bad):
    """An emissive triangle whose area is NaN / infinite (a non-finite vertex: the triangle is inactive in the acceleration structure)
    is left out of the light tables like a zero-area one (lighting.c:302), so pmf, invTotalArea and MeshInfo.lightPdfArea stay finite
    (they used to turn NaN and with them every emitter-hit MIS weight of that mesh)."""
    import json
    import struct
    from vkrt_b200 import host
    raw = open(os.path.join(H.ROOT, "assets", "models", "cube.glb"), "rb").read()
    json_len = struct.unpack_from("<I", raw, 12)[0]
    doc = json.loads(raw[20:20 + json_len].decode())
    doc["materials"][0]["emissiveFactor"] = [1.0, 0.8, 0.6]
    js = json.dumps(doc).encode()
    js += b" " * ((4 - len(js) % 4) % 4)
    bin_chunk = bytearray(raw[20 + json_len:])
    struct.pack_into("<f", bin_chunk, 8, bad)          # x of the first POSITION (bufferView 0 starts the BIN payload)
    out = bytearray(raw[:12]) + struct.pack("<I", len(js)) + b"JSON" + js + bin_chunk
    struct.pack_into("<I", out, 8, len(out))
    glb = tmp_path / "emissive_cube.glb"
    glb.write_bytes(bytes(out))
    hs = host.Host(host_only=True, width=64, height=36)
    hs.import_mesh(str(glb))
    p = hs.prepare_scene()
    assert len(p["emissiveMeshes"]) == 1
    em = p["emissiveMeshes"][0]
    assert 0 < em["triCount"] < 12                      # the triangles of the bad vertex are left out, the rest still emit
    for k in ("pmfMesh", "invTotalArea"):
        assert np.isfinite(em[k]) and em[k] > 0
    assert np.isfinite(np.asarray(p["triAliasQ"])).all() and np.asarray(p["triAliasIdx"]).max() < em["triCount"]
    assert np.isfinite(p["meshInfos"]["lightPdfArea"]).all() and p["meshInfos"]["lightPdfArea"][0] > 0
    hs.close()
