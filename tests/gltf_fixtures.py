"""Small .glb files written in Python for the importer tests: every branch the reference's importer (src/app/mesh/loader.c) has for geometry
(indexed u8 / u16 / u32 and non-indexed primitives, missing normals / tangents, vertex colours as float and as normalised bytes, two UV
sets, several primitives per mesh), for the node hierarchy (TRS and matrix nodes, nesting, mirroring scale, meshes used twice) and for
materials (core PBR, alpha modes, double-sided, the KHR_materials_* extensions vkrt reads, texture references with samplers and transforms)."""
import json
import struct

import numpy as np


class GlbBuilder:
    def __init__(self):
        self.blob = b""
        self.views = []
        self.accessors = []

    def view(self, data):
        self.views.append({"buffer": 0, "byteOffset": len(self.blob), "byteLength": len(data)})
        self.blob += data + b"\x00" * (-len(data) % 4)
        return len(self.views) - 1

    def accessor(self, array, kind, normalized=False, with_bounds=False):
        array = np.ascontiguousarray(array)
        ctype = {np.dtype(np.float32): 5126, np.dtype(np.uint8): 5121, np.dtype(np.uint16): 5123, np.dtype(np.uint32): 5125}[array.dtype]
        acc = {"bufferView": self.view(array.tobytes()), "componentType": ctype, "count": int(array.shape[0]), "type": kind}
        if normalized:
            acc["normalized"] = True
        if with_bounds:
            acc["min"], acc["max"] = array.min(axis=0).tolist(), array.max(axis=0).tolist()
        self.accessors.append(acc)
        return len(self.accessors) - 1

    def write(self, path, doc):
        doc = dict(doc)
        doc["accessors"], doc["bufferViews"], doc["buffers"] = self.accessors, self.views, [{"byteLength": len(self.blob)}]
        js = json.dumps(doc).encode()
        js += b" " * (-len(js) % 4)
        out = struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(self.blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js + \
            struct.pack("<II", len(self.blob), 0x004E4942) + self.blob
        open(path, "wb").write(out)


def _grid(n, seed):
    """An n x n height-field patch: positions, smooth normals, tangents, uvs, triangle indices."""
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.linspace(0, 1, n, dtype=np.float32), np.linspace(0, 1, n, dtype=np.float32))
    h = (0.15 * np.sin(5 * u) * np.cos(4 * v)).astype(np.float32)
    pos = np.stack([u * 2 - 1, h, v * 2 - 1], axis=-1).reshape(-1, 3).astype(np.float32)
    nrm = np.stack([-0.75 * np.cos(5 * u) * np.cos(4 * v), np.ones_like(u), 0.6 * np.sin(5 * u) * np.sin(4 * v)], axis=-1).reshape(-1, 3)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    tan = np.concatenate([np.tile(np.array([[1, 0, 0]], np.float32), (n * n, 1)), np.where(rng.random((n * n, 1)) < 0.5, -1.0, 1.0).astype(np.float32)], axis=1)
    uv = np.stack([u, v], axis=-1).reshape(-1, 2).astype(np.float32)
    idx = []
    for j in range(n - 1):
        for i in range(n - 1):
            a = j * n + i
            idx += [a, a + n, a + 1, a + 1, a + n, a + n + 1]
    return pos, nrm, tan, uv, np.array(idx, np.uint32)


def hierarchy_glb(path, png_bytes=None):
    """Writes the file and returns nothing; `png_bytes` (any decodable PNG) becomes an embedded image, else the materials carry no textures."""
    b = GlbBuilder()
    rng = np.random.default_rng(5)
    pos, nrm, tan, uv, idx = _grid(7, 1)
    col_f = rng.random((len(pos), 4)).astype(np.float32)
    col_b = rng.integers(0, 256, (len(pos), 3), dtype=np.uint8)
    prims_a = [
        # full attribute set, u16 indices, float RGBA colours
        {"attributes": {"POSITION": b.accessor(pos, "VEC3", with_bounds=True), "NORMAL": b.accessor(nrm, "VEC3"), "TANGENT": b.accessor(tan, "VEC4"),
                        "TEXCOORD_0": b.accessor(uv, "VEC2"), "TEXCOORD_1": b.accessor(uv[:, ::-1] * 0.5, "VEC2"), "COLOR_0": b.accessor(col_f, "VEC4")},
         "indices": b.accessor(idx.astype(np.uint16), "SCALAR"), "material": 0},
        # no normals, no tangents (both generated), u32 indices, normalised byte RGB colours
        {"attributes": {"POSITION": b.accessor(pos * np.array([1, -1, 1], np.float32) + np.array([0, 1.5, 0], np.float32), "VEC3", with_bounds=True),
                        "TEXCOORD_0": b.accessor(uv, "VEC2"), "COLOR_0": b.accessor(col_b, "VEC3", normalized=True)},
         "indices": b.accessor(idx, "SCALAR"), "material": 1},
    ]
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0], [1, 0, 1], [0, 0, 1], [0, 1, 0], [0, 0, 1], [1, 0, 0]], np.float32)
    prims_b = [
        # non-indexed triangles with normals only
        {"attributes": {"POSITION": b.accessor(tri, "VEC3", with_bounds=True), "NORMAL": b.accessor(np.tile(np.array([[0, 1, 0]], np.float32), (9, 1)), "VEC3")},
         "material": 2},
        # u8 indices, no material
        {"attributes": {"POSITION": b.accessor(tri[:4] + 2.0, "VEC3", with_bounds=True)}, "indices": b.accessor(np.array([0, 1, 2, 2, 1, 3], np.uint8), "SCALAR")},
    ]
    tex = {}
    doc_images, doc_textures, doc_samplers = [], [], []
    if png_bytes is not None:
        doc_images = [{"bufferView": b.view(png_bytes), "mimeType": "image/png", "name": "checker"}]
        doc_samplers = [{"wrapS": 33648, "wrapT": 33071}]
        doc_textures = [{"source": 0, "sampler": 0, "name": "checker tex"}, {"source": 0}]
        tex = {"base": {"index": 0, "texCoord": 1, "extensions": {"KHR_texture_transform": {"offset": [0.1, 0.2], "scale": [3, 0.5], "rotation": 0.3, "texCoord": 0}}},
               "mr": {"index": 1}, "normal": {"index": 1, "scale": 0.7}, "emissive": {"index": 0}}
    materials = [
        {"name": "coated glass", "doubleSided": True, "alphaMode": "BLEND",
         "pbrMetallicRoughness": dict({"baseColorFactor": [0.9, 0.8, 0.7, 0.6], "metallicFactor": 0.25, "roughnessFactor": 0.35},
                                      **({"baseColorTexture": tex["base"], "metallicRoughnessTexture": tex["mr"]} if tex else {})),
         **({"normalTexture": tex["normal"], "emissiveTexture": tex["emissive"]} if tex else {}),
         "emissiveFactor": [0.5, 0.25, 1.0],
         "extensions": {"KHR_materials_ior": {"ior": 1.33}, "KHR_materials_transmission": {"transmissionFactor": 0.8},
                        "KHR_materials_volume": {"attenuationColor": [0.7, 0.9, 0.8], "attenuationDistance": 0.4, "thicknessFactor": 0.2},
                        "KHR_materials_clearcoat": {"clearcoatFactor": 0.6, "clearcoatRoughnessFactor": 0.15},
                        "KHR_materials_emissive_strength": {"emissiveStrength": 4.5}}},
        {"name": "velvet", "alphaMode": "MASK", "alphaCutoff": 0.35,
         "pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.05, 0.3, 1.0], "metallicFactor": 0.0, "roughnessFactor": 0.9},
         "extensions": {"KHR_materials_sheen": {"sheenColorFactor": [0.8, 0.6, 0.9], "sheenRoughnessFactor": 0.45},
                        "KHR_materials_specular": {"specularFactor": 0.7, "specularColorFactor": [1.0, 0.8, 0.6]}}},
        {"pbrMetallicRoughness": {"metallicFactor": 1.0, "roughnessFactor": 0.05}},   # unnamed metal, defaults elsewhere
    ]
    nodes = [
        {"name": "root", "translation": [0.5, 1.0, -2.0], "rotation": [0.1830127, 0.5, 0.1830127, 0.8365163], "scale": [1.5, 1.0, 0.75], "children": [1, 2]},
        {"name": "patch", "mesh": 0, "translation": [0, 0.25, 0]},
        {"name": "mirrored", "mesh": 1, "matrix": [-1, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 2, 0, 1, 2, 3, 1], "children": [3]},
        {"name": "leaf", "mesh": 0, "scale": [0.5, 0.5, 0.5]},
        {"name": "second root", "mesh": 1, "rotation": [0, 0.7071068, 0, 0.7071068]},
    ]
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 4]}], "nodes": nodes,
           "meshes": [{"name": "terrain", "primitives": prims_a}, {"name": "bits", "primitives": prims_b}], "materials": materials,
           "extensionsUsed": ["KHR_materials_ior", "KHR_materials_transmission", "KHR_materials_volume", "KHR_materials_clearcoat", "KHR_materials_sheen",
                              "KHR_materials_specular", "KHR_materials_emissive_strength", "KHR_texture_transform"]}
    if doc_images:
        doc["images"], doc["textures"], doc["samplers"] = doc_images, doc_textures, doc_samplers
    b.write(path, doc)
