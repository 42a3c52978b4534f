#!/usr/bin/env python3
"""Re-export the reference's bundled models and scenes as fixtures for this repo (run once in the build container; needs
/root/reference). Outputs are committed under assets/ so that tests and the bench do not need /root/reference at run time.

  assets/models/<name>.glb : minimal glTF 2.0 binary re-serialised by this script (POSITION/NORMAL/TEXCOORD_0 as float32,
                             indices in their original component type, node TRS, pbrMetallicRoughness factors, doubleSided)
  assets/scenes/<name>.json: vkrt.scene v1 documents. cornell.json swaps the missing dragon.glb (reference
                             .MISSING_LARGE_BLOBS:3) for bunny.glb and re-seats it on the floor; the others are unchanged.
"""
import json, os, struct, sys
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import host_ref as hr

COMP = {5121: "u1", 5123: "<u2", 5125: "<u4", 5126: "<f4"}
NC = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}


def reexport_glb(src, dst):
    doc, binary = hr._read_glb(src)
    out_bin = bytearray()
    views, accessors = [], []

    def add(arr, target, acc_type, comp, minmax=False):
        while len(out_bin) % 4:
            out_bin.append(0)
        off = len(out_bin)
        out_bin.extend(arr.tobytes())
        views.append({"buffer": 0, "byteOffset": off, "byteLength": arr.nbytes, "target": target})
        a = {"bufferView": len(views) - 1, "componentType": comp, "count": int(arr.shape[0]), "type": acc_type}
        if minmax:
            a["min"] = [float(v) for v in arr.min(axis=0)]
            a["max"] = [float(v) for v in arr.max(axis=0)]
        accessors.append(a)
        return len(accessors) - 1

    meshes = []
    for m in doc["meshes"]:
        prims = []
        for p in m["primitives"]:
            at = {}
            for key in ("POSITION", "NORMAL", "TANGENT", "TEXCOORD_0", "TEXCOORD_1", "COLOR_0"):
                if key in p["attributes"]:
                    src_acc = doc["accessors"][p["attributes"][key]]
                    arr = hr._accessor(doc, binary, p["attributes"][key]).astype("<f4")
                    at[key] = add(np.ascontiguousarray(arr), 34962, src_acc["type"], 5126, key == "POSITION")
            q = {"attributes": at}
            if "indices" in p:
                ia = doc["accessors"][p["indices"]]
                idx = hr._accessor(doc, binary, p["indices"], as_float=False).astype(COMP[ia["componentType"]])
                q["indices"] = add(np.ascontiguousarray(idx), 34963, "SCALAR", ia["componentType"])
            if "material" in p:
                q["material"] = p["material"]
            prims.append(q)
        meshes.append({"name": m.get("name", "mesh"), "primitives": prims})
    out = {"asset": {"version": "2.0", "generator": "vkrt-b200 tools/import_reference_assets.py"},
           "scene": doc.get("scene", 0), "scenes": doc["scenes"], "nodes": doc["nodes"], "meshes": meshes,
           "accessors": accessors, "bufferViews": views, "buffers": [{"byteLength": len(out_bin)}]}
    if "materials" in doc:
        out["materials"] = doc["materials"]
    js = json.dumps(out, separators=(",", ":")).encode()
    while len(js) % 4:
        js += b" "
    while len(out_bin) % 4:
        out_bin.append(0)
    total = 12 + 8 + len(js) + 8 + len(out_bin)
    with open(dst, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, total))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(out_bin), 0x004E4942) + bytes(out_bin))


def main():
    os.makedirs(os.path.join(ROOT, "assets/models"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "assets/scenes"), exist_ok=True)
    for name in ("plane", "cube", "prism", "sphere", "suzanne", "bunny"):
        reexport_glb(os.path.join(REF, "assets/models/%s.glb" % name), os.path.join(ROOT, "assets/models/%s.glb" % name))
        a = hr.load_glb(os.path.join(REF, "assets/models/%s.glb" % name))[0]
        b = hr.load_glb(os.path.join(ROOT, "assets/models/%s.glb" % name))[0]
        assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.indices, b.indices), name
        print(name, len(a.vertices), "vertices", len(a.indices) // 3, "triangles")
    for name in ("prism", "caustics", "cornell"):
        doc = json.load(open(os.path.join(REF, "assets/scenes/%s.json" % name)))
        if name == "cornell":
            doc["meshImports"] = [p.replace("dragon.glb", "bunny.glb") for p in doc["meshImports"]]
            bunny = hr.load_glb(os.path.join(ROOT, "assets/models/bunny.glb"))[0]
            for obj in doc["sceneObjects"]:
                if obj["name"] == "dragon":
                    obj["name"] = "bunny"
                    w = hr.build_mesh_transform(obj["localPosition"], obj["localRotation"], obj["localScale"])
                    p = hr._xform_points(hr.world3x4(w), bunny.vertices["position"][:, :3])
                    obj["localPosition"][2] = float(np.float32(obj["localPosition"][2] - (p[:, 2].min() + 1.0) + 1e-3))
                    obj["localPosition"][0] = float(np.float32(obj["localPosition"][0] - (0.5 * (p[:, 0].min() + p[:, 0].max()) + 0.25)))
            for m in doc["meshes"]:
                if m["name"] == "dragon":
                    m["name"] = "bunny"
            for m in doc["materials"]:
                if m["name"] == "dragon":
                    m["name"] = "bunny"
            doc["_note"] = "dragon.glb is a missing blob in the reference checkout; bunny.glb stands in (SURVEY.md 8c)"
        json.dump(doc, open(os.path.join(ROOT, "assets/scenes/%s.json" % name), "w"), indent=1)
        print("scene", name)


if __name__ == "__main__":
    main()
