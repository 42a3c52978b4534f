"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy restatement of vkrt's *host-side* scene feed (SURVEY.md §8a-2 "Host light build", "Host geometry", camera,
SceneData sync) plus the glTF / vkrt.scene ingest.  It serves two purposes:
  * it is the checker for the product's C host library (vkrt_b200/host): tests compare array-for-array;
  * it prepares device-format scene arrays for oracle-vs-CUDA parity tests.
The product never imports this module.
Parity status: pinned where it matters for the tests -- pack_shader_vertices is compared with the reference's own packing.c on a million
vertices (tests/test_reference_pin.py); the product's C host, which this module checks array for array on the bundled scenes, is itself
byte-identical to the reference's host sources (oracle/_ref/libvkrt_refhost.so) for transforms, materials, camera and light tables.

Reference files followed (paths relative to /root/reference/src):
  core/utility/packing.c:92-156          pack_shader_vertex, pack_oct_normal32, pack_tangent32, pack_color_rgba8
  core/scene/transform.c:26-35,158-223   build_mesh_transform, decompose_mesh_transform, world3x4
  core/scene/camera.c:128-143            camera_matrices (lookat, GL-clip perspective, Y flip, inverses)
  core/scene/lighting.c:55-164,267-433   build_alias_table, build_lights
  core/scene/uniform.c:93-174            default settings, scene_data
  core/api/mesh.c:107-146                sanitize_material
  core/api/vkrt_types.h:79-121           default_material
  core/render/accel/tlas.c:291-296       material_may_reject_ray_hit
  app/mesh/loader.c:1343-1907            glb import (axis swap, winding alignment, tangent generation)
  app/scene/controller.c:585-892,1528+   vkrt.scene JSON
All fp32 arithmetic is done with np.float32 scalars/arrays so that roundings match a C float implementation
(transcendentals may differ from glibc by an ulp; tests use a 2-ulp tolerance where they are involved).
"""
from __future__ import annotations

import json
import math
import os
import struct
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32
INVALID = 0xFFFFFFFF

# ---------------------------------------------------------------------------------------------------------------
# wire-format dtypes (include/vkrt_shared.h)
# ---------------------------------------------------------------------------------------------------------------
VERTEX = np.dtype([("position", "<f4", 4), ("normal", "<f4", 4), ("tangent", "<f4", 4), ("color", "<f4", 4),
                   ("texcoord0", "<f4", 2), ("texcoord1", "<f4", 2)])
SHADER_VERTEX = np.dtype([("position", "<f4", 4), ("texcoord0", "<f4", 2), ("texcoord1", "<f4", 2),
                          ("packedNormal", "<u4"), ("packedTangent", "<u4"), ("packedColor", "<u4"), ("_pad", "<u4")])
MESH_INFO = np.dtype([("position", "<f4", 3), ("vertexBase", "<u4"), ("rotation", "<f4", 3), ("vertexCount", "<u4"),
                      ("scale", "<f4", 3), ("indexBase", "<u4"), ("indexCount", "<u4"), ("materialIndex", "<u4"),
                      ("renderBackfaces", "<u4"), ("lightPdfArea", "<f4"), ("opacity", "<f4"),
                      ("reserved0", "<u4"), ("reserved1", "<u4"), ("reserved2", "<u4")])
MATERIAL = np.dtype([
    ("baseColor", "<f4", 3), ("roughness", "<f4"), ("emissionColor", "<f4", 3), ("emissionLuminance", "<f4"),
    ("eta", "<f4", 3), ("metallic", "<f4"), ("k", "<f4", 3), ("anisotropic", "<f4"),
    ("specular", "<f4"), ("specularTint", "<f4"), ("abbeNumber", "<f4"), ("reserved0", "<f4"),
    ("sheenTintWeight", "<f4", 4),
    ("clearcoat", "<f4"), ("clearcoatGloss", "<f4"), ("ior", "<f4"), ("diffuseRoughness", "<f4"),
    ("transmission", "<f4"), ("subsurface", "<f4"), ("sheenRoughness", "<f4"), ("absorptionCoefficient", "<f4"),
    ("attenuationColor", "<f4", 3), ("normalTextureScale", "<f4"),
    ("baseColorTextureIndex", "<u4"), ("metallicRoughnessTextureIndex", "<u4"), ("normalTextureIndex", "<u4"),
    ("emissiveTextureIndex", "<u4"),
    ("baseColorTextureWrap", "<u4"), ("metallicRoughnessTextureWrap", "<u4"), ("normalTextureWrap", "<u4"),
    ("emissiveTextureWrap", "<u4"),
    ("opacity", "<f4"), ("alphaCutoff", "<f4"), ("alphaMode", "<u4"), ("textureTexcoordSets", "<u4"),
    ("baseColorTextureTransform", "<f4", 4), ("metallicRoughnessTextureTransform", "<f4", 4),
    ("normalTextureTransform", "<f4", 4), ("emissiveTextureTransform", "<f4", 4), ("textureRotations", "<f4", 4)])
EMISSIVE_MESH = np.dtype([("triOffset", "<u4"), ("triCount", "<u4"), ("pmfMesh", "<f4"), ("invTotalArea", "<f4"),
                          ("emission", "<f4", 3), ("reserved0", "<f4")])
EMISSIVE_TRIANGLE = np.dtype([("v0Area", "<f4", 4), ("e1Pad", "<f4", 4), ("e2Pad", "<f4", 4)])
SCENE_DATA = np.dtype([
    ("viewInverse", "<f4", 16), ("projInverse", "<f4", 16),
    ("frameNumber", "<u4"), ("samplesPerPixel", "<u4"), ("rrMaxDepth", "<u4"), ("rrMinDepth", "<u4"),
    ("viewportRect", "<u4", 4), ("packedRenderSettings", "<u4"), ("exposure", "<f4"), ("timeBase", "<f4"),
    ("timeStep", "<f4"), ("environmentLight", "<f4", 4), ("environmentTextureIndex", "<u4"),
    ("environmentRotation", "<f4"), ("debugMode", "<u4"), ("misNeeEnabled", "<u4"), ("emissiveMeshCount", "<u4"),
    ("emissiveTriangleCount", "<u4"), ("selectionEnabled", "<u4"), ("selectedMeshIndex", "<u4"),
    ("rgb2spec_res", "<u4"), ("rgb2spec_scaleOffset", "<u4"), ("rgb2spec_dataOffset", "<u4"), ("_pad", "<u4")])
assert VERTEX.itemsize == 80 and SHADER_VERTEX.itemsize == 48 and MESH_INFO.itemsize == 80
assert MATERIAL.itemsize == 272 and EMISSIVE_MESH.itemsize == 32 and EMISSIVE_TRIANGLE.itemsize == 48
assert SCENE_DATA.itemsize == 240 and SCENE_DATA.fields["frameNumber"][1] == 128


def pack_render_settings(tone, mode, spectral):
    return (tone & 0xFFFF) | ((mode & 0xFF) << 16) | ((spectral & 0xFF) << 24)


# ---------------------------------------------------------------------------------------------------------------
# materials (api/vkrt_types.h:79-121, api/mesh.c:107-146)
# ---------------------------------------------------------------------------------------------------------------
def default_material():
    m = np.zeros((), MATERIAL)
    m["baseColor"] = (0.8, 0.8, 0.8)
    m["roughness"] = 0.5
    m["emissionColor"] = (1, 1, 1)
    m["specular"] = 0.5
    m["sheenTintWeight"] = (1, 1, 1, 0)
    m["clearcoatGloss"] = 1.0
    m["ior"] = 1.5
    m["sheenRoughness"] = 0.5
    m["attenuationColor"] = (1, 1, 1)
    m["normalTextureScale"] = 1.0
    for k in ("baseColorTextureIndex", "metallicRoughnessTextureIndex", "normalTextureIndex", "emissiveTextureIndex"):
        m[k] = INVALID
    m["opacity"] = 1.0
    m["alphaCutoff"] = 0.5
    for k in ("baseColorTextureTransform", "metallicRoughnessTextureTransform", "normalTextureTransform",
              "emissiveTextureTransform"):
        m[k] = (1, 1, 0, 0)
    return m


def _finite_clamp(v, fallback, lo, hi):
    v = float(v)
    if not math.isfinite(v):
        v = fallback
    return min(max(v, lo), hi)


def sanitize_material(m, texture_count=0):
    m = m.copy()
    for i in range(3):
        m["baseColor"][i] = _finite_clamp(m["baseColor"][i], 0, 0, 1)
        m["emissionColor"][i] = _finite_clamp(m["emissionColor"][i], 0, 0, math.inf)
        m["sheenTintWeight"][i] = _finite_clamp(m["sheenTintWeight"][i], 0, 0, 1)
        m["attenuationColor"][i] = _finite_clamp(m["attenuationColor"][i], 1, 0, 1)
        m["eta"][i] = _finite_clamp(m["eta"][i], 0, 0, math.inf)
        m["k"][i] = _finite_clamp(m["k"][i], 0, 0, math.inf)
    for k in ("metallic", "roughness", "diffuseRoughness", "specular", "specularTint", "anisotropic", "clearcoat",
              "clearcoatGloss", "transmission", "subsurface", "sheenRoughness"):
        m[k] = _finite_clamp(m[k], 0, 0, 1)
    m["sheenTintWeight"][3] = _finite_clamp(m["sheenTintWeight"][3], 0, 0, 1)
    m["ior"] = _finite_clamp(m["ior"], 1, 1, 4)
    m["abbeNumber"] = _finite_clamp(m["abbeNumber"], 0, 0, 200)
    m["absorptionCoefficient"] = _finite_clamp(m["absorptionCoefficient"], 0, 0, 1e6)
    m["emissionLuminance"] = _finite_clamp(m["emissionLuminance"], 0, 0, math.inf)
    nts = float(m["normalTextureScale"])
    m["normalTextureScale"] = max(nts if math.isfinite(nts) else 1.0, 0.0)
    m["opacity"] = _finite_clamp(m["opacity"], 1, 0, 1)
    m["alphaCutoff"] = _finite_clamp(m["alphaCutoff"], 0.5, 0, 1)
    if int(m["alphaMode"]) not in (1, 2):
        m["alphaMode"] = 0
    for slot, key in enumerate(("baseColor", "metallicRoughness", "normal", "emissive")):
        if int(m[key + "TextureIndex"]) != INVALID and int(m[key + "TextureIndex"]) >= texture_count:
            m[key + "TextureIndex"] = INVALID
    return m


def material_may_reject_ray_hit(m, mesh_opacity):
    return int(m["alphaMode"]) != 0 or float(m["opacity"]) < 0.999 or float(mesh_opacity) < 0.999


# ---------------------------------------------------------------------------------------------------------------
# vertex packing (core/utility/packing.c)
# ---------------------------------------------------------------------------------------------------------------
def _lround(x):
    """C lroundf: round half away from zero, on float32 input."""
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5)).astype(np.int64)


def _normalize3(v):
    v = np.asarray(v, dtype=np.float32).reshape(-1, 3)
    len_sq = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    ok = len_sq > f32(1e-20)
    inv = np.where(ok, f32(1.0) / np.sqrt(np.where(ok, len_sq, f32(1.0))), f32(0.0)).astype(np.float32)
    out = v * inv[:, None]
    out[~ok] = (0.0, 0.0, 1.0)
    return out


def _oct_project(v):
    n = _normalize3(v)
    inv_l1 = f32(1.0) / ((np.abs(n[:, 0]) + np.abs(n[:, 1])) + np.abs(n[:, 2]))
    px = n[:, 0] * inv_l1
    py = n[:, 1] * inv_l1
    neg = n[:, 2] < 0
    old_x = px.copy()
    px2 = (f32(1.0) - np.abs(py)) * np.where(old_x >= 0, f32(1.0), f32(-1.0))
    py2 = (f32(1.0) - np.abs(old_x)) * np.where(py >= 0, f32(1.0), f32(-1.0))
    px = np.where(neg, px2, px).astype(np.float32)
    py = np.where(neg, py2, py).astype(np.float32)
    return np.clip(px, -1, 1).astype(np.float32), np.clip(py, -1, 1).astype(np.float32)


def pack_oct_normal32(normals):
    px, py = _oct_project(normals)
    sx = _lround(px * f32(32767.0))
    sy = _lround(py * f32(32767.0))
    return ((sx & 0xFFFF) | ((sy & 0xFFFF) << 16)).astype(np.uint32)


def pack_tangent32(tangents):
    t = np.asarray(tangents, dtype=np.float32).reshape(-1, 4)
    px, py = _oct_project(t[:, :3])
    sx = _lround(px * f32(16383.0)) & 0x7FFF
    sy = _lround(py * f32(16383.0)) & 0x7FFF
    packed = (sx | (sy << 15)).astype(np.uint32)
    packed = np.where(t[:, 3] < 0, packed | np.uint32(0x80000000), packed).astype(np.uint32)
    return packed


def pack_color_rgba8(colors):
    c = np.clip(np.asarray(colors, dtype=np.float32).reshape(-1, 4), 0, 1)
    q = _lround(c * f32(255.0)).astype(np.uint32)
    return (q[:, 0] | (q[:, 1] << 8) | (q[:, 2] << 16) | (q[:, 3] << 24)).astype(np.uint32)


def pack_shader_vertices(vertices):
    out = np.zeros(len(vertices), SHADER_VERTEX)
    out["position"] = vertices["position"]
    out["texcoord0"] = vertices["texcoord0"]
    out["texcoord1"] = vertices["texcoord1"]
    out["packedNormal"] = pack_oct_normal32(vertices["normal"][:, :3])
    out["packedTangent"] = pack_tangent32(vertices["tangent"])
    out["packedColor"] = pack_color_rgba8(vertices["color"])
    return out


# ---------------------------------------------------------------------------------------------------------------
# transforms (core/scene/transform.c) — 4x4 stored as numpy [row, col]; column-major flattening via .T.ravel()
# ---------------------------------------------------------------------------------------------------------------
def _rot_axis(deg, axis):
    a = f32(math.radians(float(f32(deg))))
    c, s = f32(math.cos(float(a))), f32(math.sin(float(a)))
    m = np.eye(4, dtype=np.float32)
    if axis == 0:
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    elif axis == 1:
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    else:
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def build_mesh_transform(position, rotation_deg, scale):
    """T * Rz * Ry * Rx * S (transform.c:26-35)."""
    t = np.eye(4, dtype=np.float32)
    t[:3, 3] = np.asarray(position, dtype=np.float32)
    s = np.diag(np.array([scale[0], scale[1], scale[2], 1.0], dtype=np.float32))
    m = t @ _rot_axis(rotation_deg[2], 2) @ _rot_axis(rotation_deg[1], 1) @ _rot_axis(rotation_deg[0], 0) @ s
    return m.astype(np.float32)


def decompose_mesh_transform(world):
    """transform.c:158-210: position, Euler ZYX degrees, signed scale (best of 3 sign candidates when det < 0)."""
    world = np.asarray(world, dtype=np.float32)
    pos = world[:3, 3].copy()
    rot = np.eye(3, dtype=np.float32)
    abs_scale = np.ones(3, dtype=np.float32)
    for axis in range(3):
        col = world[:3, axis]
        n = f32(math.sqrt(float(f32(col[0] * col[0] + col[1] * col[1] + col[2] * col[2]))))
        if n < 1e-6 or not math.isfinite(float(n)):
            continue
        abs_scale[axis] = n
        rot[:, axis] = col / n
    det = float(np.dot(np.cross(rot[:, 0], rot[:, 1]), rot[:, 2]))
    candidates = [0, 1, 2] if det < 0 else [-1]
    best = None
    for flipped in candidates:
        r = rot.copy()
        s = abs_scale.copy()
        if flipped >= 0:
            s[flipped] = -s[flipped]
            r[:, flipped] = -r[:, flipped]
        sine_y = min(max(-float(r[2, 0]), -1.0), 1.0)
        ry = math.asin(sine_y)
        if abs(math.cos(ry)) > 1e-6:
            rx = math.atan2(float(r[2, 1]), float(r[2, 2]))
            rz = math.atan2(float(r[1, 0]), float(r[0, 0]))
        else:
            rx = math.atan2(-float(r[1, 2]), float(r[1, 1]))
            rz = 0.0
        deg = np.array([math.degrees(rx), math.degrees(ry), math.degrees(rz)], dtype=np.float32)
        recomposed = build_mesh_transform(pos, deg, s)
        err = float(np.max(np.abs(recomposed[:3, :] - world[:3, :])))
        if best is None or err < best[0]:
            best = (err, deg, s)
    return pos, best[1], best[2]


def world3x4(world):
    return np.asarray(world, dtype=np.float32)[:3, :].copy()


IMPORT_BASIS = _rot_axis(90.0, 0)  # kImportedMeshBasisRotationDegrees = (90, 0, 0)


def imported_node_transform(gltf_local):
    """VKRT_buildImportedNodeTransform: B * M * B^-1 with B = Rx(90deg) (transform.c:84-111)."""
    b = IMPORT_BASIS.astype(np.float64)
    return (b @ np.asarray(gltf_local, dtype=np.float64) @ np.linalg.inv(b)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# camera (core/scene/camera.c:128-143)
# ---------------------------------------------------------------------------------------------------------------
def _norm3(v):
    v = np.asarray(v, dtype=np.float32)
    n = f32(math.sqrt(float(f32(f32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))))
    if n < f32(1.1920929e-07):  # cglm glm_vec3_normalize: zero vector below FLT_EPSILON, else scale by 1/norm
        return np.zeros(3, dtype=np.float32)
    return (v * f32(f32(1.0) / n)).astype(np.float32)


def camera_matrices(pos, target, up, vfov_deg, width, height, near=0.001, far=10000.0):
    pos = np.asarray(pos, dtype=np.float32)
    target = np.asarray(target, dtype=np.float32)
    up = np.asarray(up, dtype=np.float32)
    f = _norm3(target - pos)
    s = _norm3(np.cross(f, up).astype(np.float32))
    u = np.cross(s, f).astype(np.float32)
    view = np.eye(4, dtype=np.float32)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[0, 3] = -f32(np.dot(s, pos))
    view[1, 3] = -f32(np.dot(u, pos))
    view[2, 3] = f32(np.dot(f, pos))
    ff = f32(1.0) / f32(math.tan(float(f32(math.radians(float(f32(vfov_deg)))) * f32(0.5))))
    fn = f32(1.0) / (f32(near) - f32(far))
    aspect = f32(width) / f32(height)
    proj = np.zeros((4, 4), dtype=np.float32)
    proj[0, 0] = ff / aspect
    proj[1, 1] = -ff  # proj[1][1] *= -1
    proj[2, 2] = (f32(near) + f32(far)) * fn
    proj[3, 2] = -1.0
    proj[2, 3] = f32(2.0) * f32(near) * f32(far) * fn
    view_inv = np.linalg.inv(view.astype(np.float64)).astype(np.float32)
    proj_inv = np.linalg.inv(proj.astype(np.float64)).astype(np.float32)
    return view_inv, proj_inv


# ---------------------------------------------------------------------------------------------------------------
# lights (core/scene/lighting.c)
# ---------------------------------------------------------------------------------------------------------------
def build_alias_table(pmf):
    """Vose's method with LIFO stacks in fp32, exactly lighting.c:108-164."""
    pmf = np.asarray(pmf, dtype=np.float32)
    n = len(pmf)
    q = np.zeros(n, dtype=np.float32)
    idx = np.zeros(n, dtype=np.uint32)
    scaled = (pmf * f32(n)).astype(np.float32)
    small, large = [], []
    for i in range(n):
        (small if scaled[i] < f32(1.0) else large).append(i)
    while small and large:
        s = small.pop()
        l = large.pop()
        q[s] = scaled[s]
        idx[s] = l
        scaled[l] = f32(f32(scaled[l] + scaled[s]) - f32(1.0))
        (small if scaled[l] < f32(1.0) else large).append(l)
    while large:
        l = large.pop()
        q[l] = 1.0
        idx[l] = l
    while small:
        s = small.pop()
        q[s] = 1.0
        idx[s] = s
    return q, idx


def _luminance(c):
    c = np.asarray(c, dtype=np.float32)
    return f32(f32(f32(0.2126) * c[0] + f32(0.7152) * c[1]) + f32(0.0722) * c[2])


def material_emission_weight(m):
    lum_l = float(m["emissionLuminance"])
    if not math.isfinite(lum_l) or lum_l <= 0:
        return f32(0)
    lum = _luminance(m["emissionColor"])
    if lum <= 0:
        return f32(0)
    return f32(lum * f32(m["emissionLuminance"]))


def material_eligible_for_nee(mesh_info, m):
    if float(mesh_info["opacity"]) < 0.999 or float(m["opacity"]) < 0.999:
        return False
    if int(m["emissiveTextureIndex"]) != INVALID:
        return False
    return int(m["alphaMode"]) == 0


def _xform_points(w34, p):
    w = np.asarray(w34, dtype=np.float32)
    p = np.asarray(p, dtype=np.float32)
    x = ((w[0, 0] * p[:, 0] + w[0, 1] * p[:, 1]) + w[0, 2] * p[:, 2]) + w[0, 3]
    y = ((w[1, 0] * p[:, 0] + w[1, 1] * p[:, 1]) + w[1, 2] * p[:, 2]) + w[1, 3]
    z = ((w[2, 0] * p[:, 0] + w[2, 1] * p[:, 1]) + w[2, 2] * p[:, 2]) + w[2, 3]
    return np.stack([x, y, z], axis=1).astype(np.float32)


def build_lights(meshes, mesh_infos, materials):
    """meshes: list of HostMesh (unpacked fp32 vertices/indices + world matrix). Returns the six light buffers and
    writes lightPdfArea into mesh_infos (lighting.c:267-433)."""
    mesh_infos["lightPdfArea"] = 0
    e_meshes, e_tris, tri_q, tri_idx, weights, sources = [], [], [], [], [], []
    tri_offset = 0
    for mi, mesh in enumerate(meshes):
        m = materials[int(mesh_infos[mi]["materialIndex"])]
        if not material_eligible_for_nee(mesh_infos[mi], m):
            continue
        ew = material_emission_weight(m)
        if ew <= 0 or len(mesh.indices) < 3:
            continue
        idx = mesh.indices.reshape(-1, 3)
        wp = _xform_points(world3x4(mesh.world), mesh.vertices["position"][:, :3])
        p0, p1, p2 = wp[idx[:, 0]], wp[idx[:, 1]], wp[idx[:, 2]]
        e1 = (p1 - p0).astype(np.float32)
        e2 = (p2 - p0).astype(np.float32)
        cx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
        cy = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
        cz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
        area = (f32(0.5) * np.sqrt(((cx * cx + cy * cy) + cz * cz).astype(np.float32))).astype(np.float32)
        keep = area > 0
        p0, e1, e2, area = p0[keep], e1[keep], e2[keep], area[keep]
        total = f32(0)
        for a in area:  # sequential fp32 sum like the C loop
            total = f32(total + a)
        sel = f32(total * ew)
        if sel <= 0 or len(area) == 0:
            continue
        inv_total = f32(f32(1.0) / total)
        pmf = (area * inv_total).astype(np.float32)
        q, ai = build_alias_table(pmf)
        tris = np.zeros(len(area), EMISSIVE_TRIANGLE)
        tris["v0Area"][:, :3] = p0
        tris["v0Area"][:, 3] = area
        tris["e1Pad"][:, :3] = e1
        tris["e2Pad"][:, :3] = e2
        em = np.zeros((), EMISSIVE_MESH)
        em["triOffset"] = tri_offset
        em["triCount"] = len(area)
        em["invTotalArea"] = inv_total
        em["emission"] = (m["emissionColor"] * m["emissionLuminance"]).astype(np.float32)
        e_meshes.append(em)
        e_tris.append(tris)
        tri_q.append(q)
        tri_idx.append(ai)
        weights.append(sel)
        sources.append(mi)
        tri_offset += len(area)
    n = len(e_meshes)
    out_m = np.zeros(max(n, 1), EMISSIVE_MESH)
    out_t = np.concatenate(e_tris) if e_tris else np.zeros(1, EMISSIVE_TRIANGLE)
    out_tq = np.concatenate(tri_q) if tri_q else np.zeros(1, np.float32)
    out_ti = np.concatenate(tri_idx) if tri_idx else np.zeros(1, np.uint32)
    mq, mi_ = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.uint32)
    if n:
        total_w = f32(0)
        for w in weights:
            total_w = f32(total_w + w)
        inv_w = f32(f32(1.0) / total_w)
        pmfs = np.array([f32(w * inv_w) for w in weights], dtype=np.float32)
        for k in range(n):
            e_meshes[k]["pmfMesh"] = pmfs[k]
            out_m[k] = e_meshes[k]
            mesh_infos[sources[k]]["lightPdfArea"] = f32(pmfs[k] * f32(e_meshes[k]["invTotalArea"]))
        mq, mi_ = build_alias_table(pmfs)
    return dict(meshes=out_m, triangles=out_t, meshAliasQ=mq, meshAliasIdx=mi_, triAliasQ=out_tq, triAliasIdx=out_ti,
                meshCount=n, triangleCount=tri_offset)


# ---------------------------------------------------------------------------------------------------------------
# glb import (app/mesh/loader.c) — triangles only, one mesh per primitive, DFS over scene roots
# ---------------------------------------------------------------------------------------------------------------
_COMP = {5120: ("i1", 1), 5121: ("u1", 1), 5122: ("<i2", 2), 5123: ("<u2", 2), 5125: ("<u4", 4), 5126: ("<f4", 4)}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _read_glb(path):
    b = open(path, "rb").read()
    magic, _ver, _length = struct.unpack("<III", b[:12])
    if magic != 0x46546C67:
        raise ValueError("not a glb: %s" % path)
    off = 12
    doc, binary = None, b""
    while off < len(b):
        clen, ctype = struct.unpack("<II", b[off:off + 8])
        chunk = b[off + 8:off + 8 + clen]
        if ctype == 0x4E4F534A:
            doc = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942:
            binary = chunk
        off += 8 + clen
    return doc, binary


def _accessor(doc, binary, index, as_float=True):
    a = doc["accessors"][index]
    bv = doc["bufferViews"][a["bufferView"]]
    dt, size = _COMP[a["componentType"]]
    nc = _NCOMP[a["type"]]
    stride = bv.get("byteStride", 0) or size * nc
    start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
    count = a["count"]
    raw = np.frombuffer(binary, dtype=np.uint8, count=(count - 1) * stride + size * nc, offset=start)
    if stride == size * nc:
        arr = raw.view(dt).reshape(count, nc)
    else:
        arr = np.stack([raw[i * stride:i * stride + size * nc].view(dt) for i in range(count)])
    if not as_float:
        return arr.astype(np.uint32).reshape(-1)
    arr = arr.astype(np.float32)
    if a.get("normalized"):
        scale = {5120: 127.0, 5121: 255.0, 5122: 32767.0, 5123: 65535.0}.get(a["componentType"])
        if scale:
            arr = np.maximum(arr / f32(scale), f32(-1.0))
    return arr


def _node_local_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], dtype=np.float64).reshape(4, 4).T
    t = np.array(node.get("translation", [0, 0, 0]), dtype=np.float64)
    q = np.array(node.get("rotation", [0, 0, 0, 1]), dtype=np.float64)
    s = np.array(node.get("scale", [1, 1, 1]), dtype=np.float64)
    x, y, z, w = q
    r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    m = np.eye(4)
    m[:3, :3] = r * s[None, :]
    m[:3, 3] = t
    return m


@dataclass
class HostMesh:
    name: str
    vertices: np.ndarray  # VERTEX
    indices: np.ndarray   # uint32
    world: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    material_index: int = 0
    opacity: float = 1.0
    render_backfaces: int = 0
    node_local: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    gltf_material: dict | None = None


def _dot3(a, b):
    """fp32 dot product summed left to right (np.dot may go through BLAS with a different accumulation)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def _fallback_tangent(n):
    n = np.asarray(n, dtype=np.float32)
    if float(_dot3(n, n)) <= 1e-12:
        n = np.array([0, 0, 1], dtype=np.float32)
    else:
        n = _norm3(n)
    up = np.array([1, 0, 0], dtype=np.float32) if abs(float(n[2])) > 0.999 else np.array([0, 0, 1], dtype=np.float32)
    t = np.cross(up, n).astype(np.float32)
    if float(_dot3(t, t)) <= 1e-12:
        t = np.array([1, 0, 0], dtype=np.float32)
    else:
        t = _norm3(t)
    return np.array([t[0], t[1], t[2], 1.0], dtype=np.float32)


def _orthonormalize_tangent(n, t, handed):
    n = np.asarray(n, dtype=np.float32)
    t = np.asarray(t, dtype=np.float32)
    if float(_dot3(n, n)) <= 1e-12 or float(_dot3(t, t)) <= 1e-12:
        return None
    n = _norm3(n)
    t = (t - n * _dot3(t, n)).astype(np.float32)
    if float(_dot3(t, t)) <= 1e-12:
        return None
    t = _norm3(t)
    return np.array([t[0], t[1], t[2], -1.0 if handed < 0 else 1.0], dtype=np.float32)


def _generate_tangents(verts, indices, uv):
    n = len(verts)
    t1 = np.zeros((n, 3), dtype=np.float32)
    t2 = np.zeros((n, 3), dtype=np.float32)
    pos = verts["position"][:, :3]
    for tri in indices.reshape(-1, 3):
        i0, i1, i2 = (int(v) for v in tri)
        e1 = pos[i1] - pos[i0]
        e2 = pos[i2] - pos[i0]
        du1, dv1 = uv[i1] - uv[i0]
        du2, dv2 = uv[i2] - uv[i0]
        det = f32(du1 * dv2) - f32(dv1 * du2)
        if abs(float(det)) <= 1e-12:
            continue
        inv = f32(1.0) / det
        sdir = ((dv2 * e1 - dv1 * e2) * inv).astype(np.float32)
        tdir = ((du1 * e2 - du2 * e1) * inv).astype(np.float32)
        for i in (i0, i1, i2):
            t1[i] += sdir
            t2[i] += tdir
    for i in range(n):
        nrm = verts["normal"][i, :3]
        if float(_dot3(nrm, nrm)) <= 1e-12 or float(_dot3(t1[i], t1[i])) <= 1e-12:
            verts["tangent"][i] = _fallback_tangent(nrm)
            continue
        nn = _norm3(nrm)
        handed = -1.0 if float(_dot3(np.cross(nn, t1[i]).astype(np.float32), t2[i])) < 0 else 1.0
        r = _orthonormalize_tangent(nrm, t1[i], handed)
        verts["tangent"][i] = r if r is not None else _fallback_tangent(nrm)


def load_glb(path):
    """Returns list[HostMesh] (node_local = engine-space local matrix of the owning node; one root-level node per mesh
    is assumed by the bundled models, deeper hierarchies are flattened into node_local)."""
    doc, binary = _read_glb(path)
    meshes = []
    scene = doc["scenes"][doc.get("scene", 0)]

    def visit(node_index, parent_world):
        node = doc["nodes"][node_index]
        local = imported_node_transform(_node_local_matrix(node))
        world = (parent_world.astype(np.float64) @ local.astype(np.float64)).astype(np.float32)
        if "mesh" in node:
            gm = doc["meshes"][node["mesh"]]
            for pi, prim in enumerate(gm["primitives"]):
                if prim.get("mode", 4) != 4:
                    continue
                at = prim["attributes"]
                pos = _accessor(doc, binary, at["POSITION"])
                n = len(pos)
                v = np.zeros(n, VERTEX)
                v["position"][:, 0], v["position"][:, 1], v["position"][:, 2] = pos[:, 0], -pos[:, 2], pos[:, 1]
                v["position"][:, 3] = 1.0
                has_normals = "NORMAL" in at
                if has_normals:
                    nr = _accessor(doc, binary, at["NORMAL"])
                    v["normal"][:, 0], v["normal"][:, 1], v["normal"][:, 2] = nr[:, 0], -nr[:, 2], nr[:, 1]
                use_imported_tangents = has_normals and "TANGENT" in at
                if use_imported_tangents:
                    tg = _accessor(doc, binary, at["TANGENT"])
                    v["tangent"][:, 0], v["tangent"][:, 1], v["tangent"][:, 2], v["tangent"][:, 3] = tg[:, 0], -tg[:, 2], tg[:, 1], tg[:, 3]
                v["color"] = 1.0
                if "COLOR_0" in at:
                    c = _accessor(doc, binary, at["COLOR_0"])
                    v["color"][:, :c.shape[1]] = c
                if "TEXCOORD_0" in at:
                    v["texcoord0"] = _accessor(doc, binary, at["TEXCOORD_0"])
                if "TEXCOORD_1" in at:
                    v["texcoord1"] = _accessor(doc, binary, at["TEXCOORD_1"])
                idx = _accessor(doc, binary, prim["indices"], as_float=False) if "indices" in prim else np.arange(n, dtype=np.uint32)
                idx = idx.astype(np.uint32).copy()
                gmat = doc["materials"][prim["material"]] if "material" in prim else None
                tset = 0
                if gmat and "normalTexture" in gmat:
                    tset = gmat["normalTexture"].get("texCoord", 0)
                    tset = 0 if tset > 1 else tset
                tex_key = "TEXCOORD_1" if tset == 1 else "TEXCOORD_0"
                if not has_normals:
                    _generate_normals(v, idx)
                else:
                    _align_winding(v, idx)
                    if use_imported_tangents:
                        for i in range(n):
                            r = _orthonormalize_tangent(v["normal"][i, :3], v["tangent"][i, :3], v["tangent"][i, 3])
                            v["tangent"][i] = r if r is not None else _fallback_tangent(v["normal"][i, :3])
                    elif tex_key in at:
                        _generate_tangents(v, idx, v["texcoord1"] if tset == 1 else v["texcoord0"])
                    else:
                        for i in range(n):
                            v["tangent"][i] = _fallback_tangent(v["normal"][i, :3])
                name = node.get("name") or gm.get("name") or "mesh"          # loader.c:1097-1118 buildEntryName
                if len(gm["primitives"]) > 1:
                    name = "%s_%d" % (name, pi)
                hm = HostMesh(name=name, vertices=v, indices=idx, world=world.copy(), node_local=world.copy(),
                              render_backfaces=1 if (gmat and gmat.get("doubleSided")) else 0, gltf_material=gmat)
                meshes.append(hm)
        for ch in node.get("children", []):
            visit(ch, world)

    for root in scene["nodes"]:
        visit(root, np.eye(4, dtype=np.float32))
    return meshes


def _generate_normals(v, idx):
    pos = v["position"][:, :3]
    acc = np.zeros((len(v), 3), dtype=np.float32)
    for tri in idx.reshape(-1, 3):
        i0, i1, i2 = (int(t) for t in tri)
        fn = np.cross(pos[i1] - pos[i0], pos[i2] - pos[i0]).astype(np.float32)
        if float(np.dot(fn, fn)) <= 1e-12:
            continue
        acc[i0] += fn
        acc[i1] += fn
        acc[i2] += fn
    for i in range(len(v)):
        n = acc[i]
        v["normal"][i, :3] = _norm3(n) if float(np.dot(n, n)) > 1e-12 else (0, 0, 1)
        v["normal"][i, 3] = 0


def _align_winding(v, idx):
    pos = v["position"][:, :3]
    nr = v["normal"][:, :3]
    tri = idx.reshape(-1, 3)
    p0, p1, p2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    e1 = (p1 - p0).astype(np.float32)
    e2 = (p2 - p0).astype(np.float32)
    fn = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                   e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], axis=1).astype(np.float32)
    avg = ((nr[tri[:, 0]] + nr[tri[:, 1]]) + nr[tri[:, 2]]).astype(np.float32)
    fn2 = (fn[:, 0] * fn[:, 0] + fn[:, 1] * fn[:, 1]) + fn[:, 2] * fn[:, 2]
    av2 = (avg[:, 0] * avg[:, 0] + avg[:, 1] * avg[:, 1]) + avg[:, 2] * avg[:, 2]
    d = (fn[:, 0] * avg[:, 0] + fn[:, 1] * avg[:, 1]) + fn[:, 2] * avg[:, 2]
    align = np.where((fn2 <= 1e-12) | (av2 <= 1e-12), f32(0), d)
    flip = align < 0
    t1 = tri[:, 1].copy()
    tri[flip, 1] = tri[flip, 2]
    tri[flip, 2] = t1[flip]


def gltf_material_to_vkrt(gmat):
    """Subset of app/mesh/loader.c:942-1078 used by the bundled models (metallic-roughness factors, double sided,
    KHR ior/transmission/emissive_strength)."""
    m = default_material()
    if not gmat:
        return m
    pbr = gmat.get("pbrMetallicRoughness", {})
    bc = pbr.get("baseColorFactor", [1, 1, 1, 1])
    m["baseColor"] = bc[:3]
    m["opacity"] = bc[3]
    m["metallic"] = pbr.get("metallicFactor", 1.0)
    m["roughness"] = pbr.get("roughnessFactor", 1.0)
    ext = gmat.get("extensions", {})
    if "KHR_materials_ior" in ext:
        m["ior"] = ext["KHR_materials_ior"].get("ior", 1.5)
    if "KHR_materials_transmission" in ext:
        m["transmission"] = ext["KHR_materials_transmission"].get("transmissionFactor", 0.0)
    em = np.array(gmat.get("emissiveFactor", [0, 0, 0]), dtype=np.float32)
    strength = ext.get("KHR_materials_emissive_strength", {}).get("emissiveStrength", 1.0)
    em = em * f32(strength)
    mx = float(em.max())
    if mx > 0:
        m["emissionColor"] = em / f32(mx)
        m["emissionLuminance"] = mx
    am = gmat.get("alphaMode", "OPAQUE")
    m["alphaMode"] = {"OPAQUE": 0, "MASK": 1, "BLEND": 2}[am]
    if am == "MASK":                                                             # loader.c:1546-1552
        m["alphaCutoff"] = gmat.get("alphaCutoff", 0.5)
    elif am == "BLEND":
        m["alphaCutoff"] = f32(1.0) / f32(255.0)
    return m


# ---------------------------------------------------------------------------------------------------------------
# Scene assembly
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Settings:
    camera_pos: tuple = (-0.5, 0.2, -0.2)
    camera_target: tuple = (0.0, 0.0, 0.0)
    camera_up: tuple = (0.0, 0.0, 1.0)
    vfov: float = 40.0
    samples_per_pixel: int = 8
    rr_min_depth: int = 4
    rr_max_depth: int = 8
    tone_mapping: int = 1
    render_mode: int = 0
    spectral_sampling: int = 1
    exposure: float = 1.0
    environment_color: tuple = (0.25, 0.25, 0.25)
    environment_strength: float = 1.0
    environment_rotation: float = 0.0
    environment_texture_index: int = INVALID
    debug_mode: int = 0
    mis_nee_enabled: int = 1


@dataclass
class Scene:
    meshes: list
    materials: np.ndarray
    settings: Settings
    textures: list = field(default_factory=list)

    def prepare(self, width, height, dedup=True):
        """Device-format arrays, exactly what the C host hands to vkrt_cuda_set_* (geometry.c:644-806 layout:
        owners' vertices/indices concatenated in mesh order; duplicates share their source's bases)."""
        infos = np.zeros(len(self.meshes), MESH_INFO)
        vparts, iparts, owners = [], [], {}
        geometry_source = np.zeros(len(self.meshes), np.uint32)
        vbase = ibase = 0
        for i, mesh in enumerate(self.meshes):
            key = (mesh.vertices.tobytes(), mesh.indices.tobytes()) if dedup else i
            if key in owners:
                src = owners[key]
                infos[i]["vertexBase"], infos[i]["indexBase"] = infos[src]["vertexBase"], infos[src]["indexBase"]
                geometry_source[i] = src
            else:
                owners[key] = i
                geometry_source[i] = i
                infos[i]["vertexBase"], infos[i]["indexBase"] = vbase, ibase
                vparts.append(pack_shader_vertices(mesh.vertices))
                iparts.append(mesh.indices.astype(np.uint32))
                vbase += len(mesh.vertices)
                ibase += len(mesh.indices)
            infos[i]["vertexCount"] = len(mesh.vertices)
            infos[i]["indexCount"] = len(mesh.indices)
            pos, rot, scale = decompose_mesh_transform(mesh.world)
            infos[i]["position"], infos[i]["rotation"], infos[i]["scale"] = pos, rot, scale
            infos[i]["materialIndex"] = mesh.material_index
            infos[i]["renderBackfaces"] = mesh.render_backfaces
            infos[i]["opacity"] = mesh.opacity
        lights = build_lights(self.meshes, infos, self.materials)
        world = np.stack([world3x4(m.world) for m in self.meshes]).astype(np.float32) if self.meshes else np.zeros((0, 3, 4), np.float32)
        alpha = np.array([1 if material_may_reject_ray_hit(self.materials[m.material_index], m.opacity) else 0
                          for m in self.meshes], dtype=np.uint8)
        s = self.settings
        sd = np.zeros((), SCENE_DATA)
        vi, pi = camera_matrices(s.camera_pos, s.camera_target, s.camera_up, s.vfov, width, height)
        sd["viewInverse"] = vi.T.ravel()
        sd["projInverse"] = pi.T.ravel()
        sd["samplesPerPixel"] = max(s.samples_per_pixel, 1)
        sd["rrMaxDepth"], sd["rrMinDepth"] = s.rr_max_depth, s.rr_min_depth
        sd["viewportRect"] = (0, 0, width, height)
        sd["packedRenderSettings"] = pack_render_settings(s.tone_mapping, s.render_mode, s.spectral_sampling)
        sd["exposure"] = s.exposure
        sd["timeBase"], sd["timeStep"] = -1.0, 0.5
        ec = np.asarray(s.environment_color, dtype=np.float32) * f32(s.environment_strength)
        sd["environmentLight"] = (ec[0], ec[1], ec[2], s.environment_strength)
        sd["environmentTextureIndex"] = s.environment_texture_index
        sd["environmentRotation"] = s.environment_rotation
        sd["debugMode"] = s.debug_mode
        sd["misNeeEnabled"] = 1 if s.mis_nee_enabled else 0
        sd["emissiveMeshCount"] = lights["meshCount"]
        sd["emissiveTriangleCount"] = lights["triangleCount"]
        sd["selectedMeshIndex"] = INVALID
        return dict(vertices=np.concatenate(vparts) if vparts else np.zeros(0, SHADER_VERTEX),
                    indices=np.concatenate(iparts) if iparts else np.zeros(0, np.uint32),
                    meshInfos=infos, world3x4=world, geometrySource=geometry_source, alphaTested=alpha,
                    materials=self.materials, lights=lights, sceneData=sd, textures=self.textures)


def _compose_prs(obj):
    if "localTransform" in obj:
        return np.array(obj["localTransform"], dtype=np.float32).reshape(4, 4).T
    return build_mesh_transform(obj.get("localPosition", [0, 0, 0]), obj.get("localRotation", [0, 0, 0]),
                                obj.get("localScale", [1, 1, 1]))


_MATERIAL_JSON_KEYS = {k for k in MATERIAL.names if k not in ("reserved0",)}


def load_scene_json(path, model_overrides=None):
    """vkrt.scene v1 (app/scene/controller.c:1528-1597). model_overrides maps basenames to replacement files
    (used for the missing dragon.glb -> bunny.glb substitution)."""
    doc = json.load(open(path))
    assert doc.get("format") == "vkrt.scene"
    base = os.path.dirname(os.path.abspath(path))
    imported = []  # per import: list[HostMesh]
    for rel in doc["meshImports"]:
        p = os.path.normpath(os.path.join(base, rel))
        if model_overrides and os.path.basename(p) in model_overrides:
            p = model_overrides[os.path.basename(p)]
        imported.append(load_glb(p))
    mat_count = 1 + max([m["index"] for m in doc.get("materials", [])] + [0])
    materials = np.zeros(mat_count, MATERIAL)
    materials[:] = default_material()
    for entry in doc.get("materials", []):
        m = default_material()
        for k, v in entry.get("material", {}).items():
            if k in _MATERIAL_JSON_KEYS:
                m[k] = v
        materials[entry["index"]] = sanitize_material(m)
    meshes = []
    for jm in doc["meshes"]:
        src = imported[jm["importIndex"]][jm["importLocalIndex"]]
        hm = HostMesh(name=jm.get("name", src.name), vertices=src.vertices, indices=src.indices,
                      material_index=jm.get("materialIndex", 0) if jm.get("hasMaterialAssignment", True) else 0,
                      opacity=jm.get("opacity", 1.0), render_backfaces=1 if jm.get("renderBackfaces") else 0)
        meshes.append(hm)
    worlds = []
    for obj in doc.get("sceneObjects", []):
        local = _compose_prs(obj)
        parent = obj.get("parentIndex")
        w = local if parent is None else (worlds[parent].astype(np.float64) @ local.astype(np.float64)).astype(np.float32)
        worlds.append(w)
        mi = obj.get("meshIndex")
        if mi is not None:
            meshes[mi].world = w
    st = Settings()
    ss = doc.get("sceneSettings", {})
    cam = ss.get("camera", {})
    st.camera_pos = tuple(cam.get("position", st.camera_pos))
    st.camera_target = tuple(cam.get("target", st.camera_target))
    st.camera_up = tuple(cam.get("up", st.camera_up))
    st.vfov = cam.get("vfov", st.vfov)
    st.rr_min_depth = ss.get("rrMinDepth", st.rr_min_depth)
    st.rr_max_depth = ss.get("rrMaxDepth", st.rr_max_depth)
    st.tone_mapping = ss.get("toneMappingMode", st.tone_mapping)
    st.render_mode = ss.get("renderMode", st.render_mode)
    st.spectral_sampling = ss.get("spectralSamplingMode", st.spectral_sampling)
    st.exposure = ss.get("exposure", st.exposure)
    st.environment_color = tuple(ss.get("environmentColor", st.environment_color))
    st.environment_strength = ss.get("environmentStrength", st.environment_strength)
    st.environment_rotation = ss.get("environmentRotation", st.environment_rotation)
    st.mis_nee_enabled = 1 if ss.get("misNeeEnabled", True) else 0
    return Scene(meshes=meshes, materials=materials, settings=st)


# ---------------------------------------------------------------------------------------------------------------
# Procedural geometry (no reference equivalent; test scenes)
# ---------------------------------------------------------------------------------------------------------------
def make_vertices(positions, normals=None, uvs=None):
    positions = np.asarray(positions, dtype=np.float32)
    v = np.zeros(len(positions), VERTEX)
    v["position"][:, :3] = positions
    v["position"][:, 3] = 1.0
    if normals is not None:
        v["normal"][:, :3] = np.asarray(normals, dtype=np.float32)
    v["color"] = 1.0
    if uvs is not None:
        v["texcoord0"] = np.asarray(uvs, dtype=np.float32)
    for i in range(len(v)):
        v["tangent"][i] = _fallback_tangent(v["normal"][i, :3])
    return v


def quad_mesh(name="quad", size=1.0):
    s = size
    p = [(-s, -s, 0), (s, -s, 0), (s, s, 0), (-s, s, 0)]
    n = [(0, 0, 1)] * 4
    uv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    return HostMesh(name=name, vertices=make_vertices(p, n, uv), indices=np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32))


def uv_sphere_mesh(name="sphere", radius=1.0, segments=32, rings=16):
    pos, nrm, uv, idx = [], [], [], []
    for r in range(rings + 1):
        th = math.pi * r / rings
        for s in range(segments + 1):
            ph = 2 * math.pi * s / segments
            n = (math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th))
            pos.append(tuple(radius * c for c in n))
            nrm.append(n)
            uv.append((s / segments, r / rings))
    for r in range(rings):
        for s in range(segments):
            a = r * (segments + 1) + s
            b = a + segments + 1
            if r != 0:
                idx += [a, b, a + 1]
            if r != rings - 1:
                idx += [a + 1, b, b + 1]
    return HostMesh(name=name, vertices=make_vertices(pos, nrm, uv), indices=np.array(idx, dtype=np.uint32))


def box_mesh(name="box", half=(1.0, 1.0, 1.0)):
    hx, hy, hz = half
    pos, nrm, idx = [], [], []
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    for n, u, v in faces:
        base = len(pos)
        for su, sv in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            p = tuple((n[k] + su * u[k] + sv * v[k]) * (hx, hy, hz)[k] for k in range(3))
            pos.append(p)
            nrm.append(n)
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return HostMesh(name=name, vertices=make_vertices(pos, nrm), indices=np.array(idx, dtype=np.uint32))


def pcg32_stream(seed, n):
    """n uint32 values from PCG32 (XSH-RR), state seeded like the reference implementation's pcg32_srandom(seed, 1)."""
    mask = (1 << 64) - 1
    state, inc = 0, (1 << 1) | 1
    def step():
        nonlocal state
        old = state
        state = (old * 6364136223846793005 + inc) & mask
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF
    step()
    state = (state + seed) & mask
    step()
    return np.array([step() for _ in range(n)], dtype=np.uint32)


def soup_meshes(triangle_count, mesh_count=16, seed=0x5EED0001):
    """random triangle soup, centres uniform in [-1,1]^3, edge length log-uniform in [2e-3, 2e-2]
    (scaled up for small counts so the scene stays visually dense), triangle i -> mesh i % mesh_count."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(-1, 1, (triangle_count, 3))
    scale = (1e7 / max(triangle_count, 1)) ** (1.0 / 3.0)
    edge = np.exp(rng.uniform(math.log(2e-3), math.log(2e-2), triangle_count)) * scale
    a = rng.normal(size=(triangle_count, 3))
    b = rng.normal(size=(triangle_count, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b -= a * np.sum(a * b, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    p0 = c - 0.5 * edge[:, None] * a - 0.3 * edge[:, None] * b
    p1 = c + 0.5 * edge[:, None] * a - 0.3 * edge[:, None] * b
    p2 = c + 0.6 * edge[:, None] * b
    n = np.cross(p1 - p0, p2 - p0)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    meshes = []
    for k in range(mesh_count):
        sel = np.arange(k, triangle_count, mesh_count)
        pos = np.stack([p0[sel], p1[sel], p2[sel]], axis=1).reshape(-1, 3)
        nrm = np.repeat(n[sel], 3, axis=0)
        v = np.zeros(len(pos), VERTEX)
        v["position"][:, :3] = pos
        v["position"][:, 3] = 1
        v["normal"][:, :3] = nrm
        v["color"] = 1
        v["tangent"][:, :3] = np.repeat(a[sel], 3, axis=0)
        v["tangent"][:, 3] = 1
        meshes.append(HostMesh(name="soup%d" % k, vertices=v, indices=np.arange(len(pos), dtype=np.uint32)))
    return meshes


def soup_scene(triangle_count, seed=0x5EED0001):
    meshes = soup_meshes(triangle_count, 16, seed)
    rng = np.random.Generator(np.random.PCG64(seed ^ 0xABCDEF))
    materials = np.zeros(18, MATERIAL)
    materials[:] = default_material()
    for k in range(16):
        m = default_material()
        if k < 8:
            m["roughness"] = 1.0
            m["baseColor"] = rng.uniform(0.2, 0.8, 3)
        else:
            m["metallic"] = 1.0
            m["roughness"] = (0.05, 0.2, 0.4)[k % 3]
            m["baseColor"] = rng.uniform(0.5, 0.9, 3)
        materials[1 + k] = sanitize_material(m)
        meshes[k].material_index = 1 + k
    light = quad_mesh("light", 0.5)
    light.world = build_mesh_transform((0, 0, 1.5), (180, 0, 0), (1, 1, 1))
    lm = default_material()
    lm["emissionLuminance"] = 30.0
    materials[17] = sanitize_material(lm)
    light.material_index = 17
    meshes.append(light)
    st = Settings(camera_pos=(0.0, -3.6, 0.4), camera_target=(0, 0, 0), vfov=40.0)
    return Scene(meshes=meshes, materials=materials, settings=st)


def cornell_scene(with_glass=False, sphere_segments=32):
    """Procedural Cornell-style box in vkrt's conventions (Z up): 5 planes (one geometry, 5 instances), an emissive
    sphere light, a rough-dielectric sphere and a mirror sphere. Used where the reference's .glb assets are not needed."""
    mats = np.zeros(9, MATERIAL)
    mats[:] = default_material()
    def mat(**kw):
        m = default_material()
        for k, v in kw.items():
            m[k] = v
        return sanitize_material(m)
    mats[1] = mat(roughness=1.0)
    mats[2] = mat(roughness=1.0)
    mats[3] = mat(roughness=1.0)
    mats[4] = mat(baseColor=(0.8, 0.14, 0.12), roughness=1.0)
    mats[5] = mat(baseColor=(0.13, 0.7, 0.16), roughness=1.0)
    mats[6] = mat(roughness=1.0, emissionLuminance=30.0)
    mats[7] = mat(baseColor=(0.74, 0.73, 0.72), roughness=0.0) if not with_glass else mat(baseColor=(1, 1, 1), roughness=0.05, transmission=1.0, ior=1.5)
    mats[8] = mat(baseColor=(1, 1, 1), metallic=1.0, roughness=0.0)
    meshes = []
    def plane(name, pos, rot, mi):
        q = quad_mesh(name)
        q.world = build_mesh_transform(pos, rot, (1, 1, 1))
        q.material_index = mi
        meshes.append(q)
    plane("floor", (0, 0, -1), (0, 0, 0), 1)
    plane("ceiling", (0, 0, 1), (180, 0, 0), 2)
    plane("back", (0, 1, 0), (90, 0, 0), 3)
    plane("left", (-1, 0, 0), (0, 90, 0), 4)
    plane("right", (1, 0, 0), (0, -90, 0), 5)
    light = uv_sphere_mesh("light", 1.0, sphere_segments, sphere_segments // 2)
    light.world = build_mesh_transform((0, 0, 0.75), (0, 0, 0), (0.15, 0.15, 0.15))
    light.material_index = 6
    meshes.append(light)
    big = uv_sphere_mesh("big", 1.0, sphere_segments, sphere_segments // 2)
    big.world = build_mesh_transform((-0.35, 0.3, -0.6), (0, 0, 30), (0.4, 0.4, 0.4))
    big.material_index = 7
    meshes.append(big)
    small = uv_sphere_mesh("small", 1.0, sphere_segments, sphere_segments // 2)
    small.world = build_mesh_transform((0.45, -0.2, -0.7), (0, 0, 0), (0.3, 0.3, 0.3))
    small.material_index = 8
    meshes.append(small)
    st = Settings(camera_pos=(0.0, -3.7, 0.0), camera_target=(0, 0, 0), vfov=40.0)
    return Scene(meshes=meshes, materials=mats, settings=st)
