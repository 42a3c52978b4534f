/* gltf_import.c — binary glTF (.glb) import into engine-space meshes.
 *
 * Restates the behaviour of the reference's importer (src/app/mesh/loader.c, which sits on the vendored cgltf):
 *   - one engine mesh per glTF *triangle* primitive, visited depth-first over the default scene's root nodes (:2027-2071);
 *   - positions, normals and tangents go from glTF's Y-up frame to the engine's Z-up frame as (x, y, z) -> (x, -z, y)
 *     (:1762,:1772-1774,:1816) and node matrices are conjugated by Rx(90 deg) (VKRT_buildImportedNodeTransform);
 *   - missing normals are generated from area-weighted face normals (:1599-1645); present normals make every triangle's
 *     winding agree with its averaged vertex normal (:1894-1907) — this decides which side is frontFace;
 *   - tangents: imported ones are re-orthonormalised (:1448-1462), otherwise generated from the UV set the normal map uses
 *     (:1646-1754), otherwise a fallback cross(up, n) (:1418-1446);
 *   - materials: metallic-roughness / specular-glossiness factors, KHR_materials_{specular, ior, transmission, volume,
 *     clearcoat, sheen, emissive_strength}, alpha mode / cutoff, doubleSided (:942-1078,:1472-1555).
 * Images (bufferView, data: URI or external file; PNG / JPEG / EXR through image_decode.c) are decoded once per (image, colour space)
 * pair in the reference's order (loader.c:318-357) and bound with sampler wrap modes, texture coordinate set and KHR_texture_transform
 * (loader.c:359-404, 905-940, 1486-1540). */
#include "gltf_import.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hjson.h"
#include "image_decode.h"
#include "hmath.h"

typedef struct {
    const hj_value* doc;
    const unsigned char* bin;
    size_t binSize;
    GltfImport* out;
    char* error;
    size_t errorSize;
    const char* sourcePath;   /* the .glb, for relative image URIs */
    struct { int image; uint32_t colorSpace; } refs[4096];
    uint32_t refCount;
} ImportCtx;

static int fail(ImportCtx* c, const char* what) {
    if (c->error && c->errorSize && !c->error[0]) snprintf(c->error, c->errorSize, "%s", what);
    return 0;
}

static size_t componentSize(int componentType) {
    switch (componentType) {
        case 5120: case 5121: return 1;
        case 5122: case 5123: return 2;
        case 5125: case 5126: return 4;
        default: return 0;
    }
}
static int typeComponents(const char* t) {
    if (!t) return 0;
    if (!strcmp(t, "SCALAR")) return 1;
    if (!strcmp(t, "VEC2")) return 2;
    if (!strcmp(t, "VEC3")) return 3;
    if (!strcmp(t, "VEC4")) return 4;
    if (!strcmp(t, "MAT4")) return 16;
    return 0;
}

typedef struct {
    const unsigned char* base;
    size_t stride, count;
    int componentType, components, normalized;
} Accessor;

static int openAccessor(ImportCtx* c, int index, Accessor* a) {
    const hj_value* acc = hj_at(hj_get(c->doc, "accessors"), (size_t)index);
    if (!acc) return fail(c, "glTF: accessor index out of range");
    if (hj_get(acc, "sparse")) return fail(c, "glTF: sparse accessors are not supported");
    const hj_value* bv = hj_at(hj_get(c->doc, "bufferViews"), (size_t)hj_number(hj_get(acc, "bufferView"), -1));
    if (!bv) return fail(c, "glTF: accessor without bufferView");
    if ((int)hj_number(hj_get(bv, "buffer"), 0) != 0) return fail(c, "glTF: only the embedded GLB buffer is supported");
    a->componentType = (int)hj_number(hj_get(acc, "componentType"), 0);
    a->components = typeComponents(hj_string(hj_get(acc, "type"), NULL));
    a->normalized = hj_bool(hj_get(acc, "normalized"), 0);
    size_t cs = componentSize(a->componentType);
    if (!cs || !a->components) return fail(c, "glTF: unsupported accessor type");
    size_t elem = cs * (size_t)a->components;
    /* every number below comes from the file as a double: reject anything that is not a small non-negative integer BEFORE it is cast
     * (a count of 2^60 used to wrap the bounds product and pass; the reference relied on cgltf_validate) */
    const double countD = hj_number(hj_get(acc, "count"), 0), strideD = hj_number(hj_get(bv, "byteStride"), 0);
    const double viewOffD = hj_number(hj_get(bv, "byteOffset"), 0), accOffD = hj_number(hj_get(acc, "byteOffset"), 0);
    const double limit = 4294967295.0;
    if (!(countD >= 0.0 && countD <= limit) || !(strideD >= 0.0 && strideD <= 252.0) || !(viewOffD >= 0.0 && viewOffD <= limit) ||
        !(accOffD >= 0.0 && accOffD <= limit))
        return fail(c, "glTF: accessor count / stride / offset out of range");
    a->count = (size_t)countD;
    a->stride = (size_t)strideD;
    if (!a->stride) a->stride = elem;
    if (a->stride < elem) return fail(c, "glTF: byteStride smaller than the element");
    const size_t start = (size_t)viewOffD + (size_t)accOffD;
    if (start > c->binSize || elem > c->binSize - start) {
        if (a->count) return fail(c, "glTF: accessor exceeds the binary chunk");
    } else if (a->count && a->count - 1 > (c->binSize - start - elem) / a->stride) {
        return fail(c, "glTF: accessor exceeds the binary chunk");
    }
    a->base = c->bin + (start <= c->binSize ? start : 0);
    return 1;
}
static float readComponent(const Accessor* a, size_t i, int k) {
    const unsigned char* p = a->base + i * a->stride + (size_t)k * componentSize(a->componentType);
    switch (a->componentType) {
        case 5126: { float f; memcpy(&f, p, 4); return f; }
        case 5125: { uint32_t u; memcpy(&u, p, 4); return (float)u; }
        case 5123: { uint16_t u; memcpy(&u, p, 2); return a->normalized ? (float)u / 65535.0f : (float)u; }
        case 5122: { int16_t s; memcpy(&s, p, 2); return a->normalized ? fmaxf((float)s / 32767.0f, -1.0f) : (float)s; }
        case 5121: { return a->normalized ? (float)p[0] / 255.0f : (float)p[0]; }
        case 5120: { int8_t s = (int8_t)p[0]; return a->normalized ? fmaxf((float)s / 127.0f, -1.0f) : (float)s; }
        default: return 0.0f;
    }
}
static uint32_t readIndex(const Accessor* a, size_t i) {
    const unsigned char* p = a->base + i * a->stride;
    switch (a->componentType) {
        case 5125: { uint32_t u; memcpy(&u, p, 4); return u; }
        case 5123: { uint16_t u; memcpy(&u, p, 2); return u; }
        case 5121: return p[0];
        default: return 0;
    }
}

/* ---- normals / tangents ---------------------------------------------------------------------------------------------- */
static void fallbackTangent(const float n_in[3], float out[4]) {
    float n[3] = {n_in[0], n_in[1], n_in[2]};
    if (h_dot3(n, n) <= 1e-12f) { n[0] = 0; n[1] = 0; n[2] = 1; }
    else h_normalize3(n);
    float up[3] = {0, 0, 1};
    if (fabsf(n[2]) > 0.999f) { up[0] = 1; up[2] = 0; }
    float t[3];
    h_cross3(up, n, t);
    if (h_dot3(t, t) <= 1e-12f) { t[0] = 1; t[1] = 0; t[2] = 0; }
    else h_normalize3(t);
    out[0] = t[0]; out[1] = t[1]; out[2] = t[2]; out[3] = 1.0f;
}
static int orthonormalizeTangent(const float n_in[3], const float t_in[3], float handedness, float out[4]) {
    if (h_dot3(n_in, n_in) <= 1e-12f || h_dot3(t_in, t_in) <= 1e-12f) return 0;
    float n[3] = {n_in[0], n_in[1], n_in[2]};
    h_normalize3(n);
    float d = h_dot3(t_in, n);
    float t[3] = {t_in[0] - n[0] * d, t_in[1] - n[1] * d, t_in[2] - n[2] * d};
    if (h_dot3(t, t) <= 1e-12f) return 0;
    h_normalize3(t);
    out[0] = t[0]; out[1] = t[1]; out[2] = t[2]; out[3] = handedness < 0.0f ? -1.0f : 1.0f;
    return 1;
}
static void generateNormals(Vertex* v, size_t nv, const uint32_t* idx, size_t ni) {
    float* acc = (float*)calloc(nv * 3, sizeof(float));
    if (!acc) return;
    for (size_t t = 0; t + 2 < ni; t += 3) {
        uint32_t i0 = idx[t], i1 = idx[t + 1], i2 = idx[t + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) continue;
        float e1[3], e2[3], fn[3];
        for (int k = 0; k < 3; k++) { e1[k] = v[i1].position[k] - v[i0].position[k]; e2[k] = v[i2].position[k] - v[i0].position[k]; }
        h_cross3(e1, e2, fn);
        if (h_dot3(fn, fn) <= 1e-12f) continue;
        for (int k = 0; k < 3; k++) { acc[i0 * 3 + k] += fn[k]; acc[i1 * 3 + k] += fn[k]; acc[i2 * 3 + k] += fn[k]; }
    }
    for (size_t i = 0; i < nv; i++) {
        float n[3] = {acc[i * 3], acc[i * 3 + 1], acc[i * 3 + 2]};
        if (h_dot3(n, n) > 1e-12f) h_normalize3(n);
        else { n[0] = 0; n[1] = 0; n[2] = 1; }
        v[i].normal[0] = n[0]; v[i].normal[1] = n[1]; v[i].normal[2] = n[2]; v[i].normal[3] = 0.0f;
    }
    free(acc);
}
static void alignWinding(const Vertex* v, size_t nv, uint32_t* idx, size_t ni) {
    for (size_t t = 0; t + 2 < ni; t += 3) {
        uint32_t i0 = idx[t], i1 = idx[t + 1], i2 = idx[t + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) continue;
        float e1[3], e2[3], fn[3], avg[3];
        for (int k = 0; k < 3; k++) {
            e1[k] = v[i1].position[k] - v[i0].position[k];
            e2[k] = v[i2].position[k] - v[i0].position[k];
            avg[k] = (v[i0].normal[k] + v[i1].normal[k]) + v[i2].normal[k];
        }
        h_cross3(e1, e2, fn);
        if (h_dot3(fn, fn) <= 1e-12f || h_dot3(avg, avg) <= 1e-12f) continue;
        if (h_dot3(fn, avg) < 0.0f) { idx[t + 1] = i2; idx[t + 2] = i1; }
    }
}
static void generateTangents(Vertex* v, size_t nv, const uint32_t* idx, size_t ni, int texcoordSet) {
    float* t1 = (float*)calloc(nv * 3, sizeof(float));
    float* t2 = (float*)calloc(nv * 3, sizeof(float));
    if (!t1 || !t2) { free(t1); free(t2); return; }
    for (size_t t = 0; t + 2 < ni; t += 3) {
        uint32_t i0 = idx[t], i1 = idx[t + 1], i2 = idx[t + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) continue;
        const float* uv0 = texcoordSet ? v[i0].texcoord1 : v[i0].texcoord0;
        const float* uv1 = texcoordSet ? v[i1].texcoord1 : v[i1].texcoord0;
        const float* uv2 = texcoordSet ? v[i2].texcoord1 : v[i2].texcoord0;
        float e1[3], e2[3];
        for (int k = 0; k < 3; k++) { e1[k] = v[i1].position[k] - v[i0].position[k]; e2[k] = v[i2].position[k] - v[i0].position[k]; }
        float du1 = uv1[0] - uv0[0], dv1 = uv1[1] - uv0[1], du2 = uv2[0] - uv0[0], dv2 = uv2[1] - uv0[1];
        float det = du1 * dv2 - dv1 * du2;
        if (fabsf(det) <= 1e-12f) continue;
        float inv = 1.0f / det;
        for (int k = 0; k < 3; k++) {
            float s = (dv2 * e1[k] - dv1 * e2[k]) * inv, tt = (du1 * e2[k] - du2 * e1[k]) * inv;
            t1[i0 * 3 + k] += s; t1[i1 * 3 + k] += s; t1[i2 * 3 + k] += s;
            t2[i0 * 3 + k] += tt; t2[i1 * 3 + k] += tt; t2[i2 * 3 + k] += tt;
        }
    }
    for (size_t i = 0; i < nv; i++) {
        const float* n = v[i].normal;
        if (h_dot3(n, n) <= 1e-12f || h_dot3(&t1[i * 3], &t1[i * 3]) <= 1e-12f) { fallbackTangent(n, v[i].tangent); continue; }
        float nn[3] = {n[0], n[1], n[2]}, c[3];
        h_normalize3(nn);
        h_cross3(nn, &t1[i * 3], c);
        float handed = h_dot3(c, &t2[i * 3]) < 0.0f ? -1.0f : 1.0f;
        if (!orthonormalizeTangent(n, &t1[i * 3], handed, v[i].tangent)) fallbackTangent(n, v[i].tangent);
    }
    free(t1); free(t2);
}

/* ---- materials -------------------------------------------------------------------------------------------------------- */
static float max3f(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

/* ---- textures (loader.c:262-404) ---------------------------------------------------------------------------------------------- */
/* texture coordinate set of a texture view: KHR_texture_transform.texCoord overrides textureInfo.texCoord (loader.c:378-384) */
static uint32_t textureViewTexcoordSet(const hj_value* view) {
    const hj_value* xf = hj_get(hj_get(view, "extensions"), "KHR_texture_transform");
    const hj_value* t = hj_get(xf, "texCoord");
    if (t && t->type == HJ_NUMBER && t->number >= 0.0) return (uint32_t)t->number;
    double v = hj_number(hj_get(view, "texCoord"), 0.0);
    return v >= 0.0 ? (uint32_t)v : 0u;
}
/* image index behind a texture view, -1 when there is none or the view uses a texture coordinate set above 1 (unsupported) */
static int textureViewImage(ImportCtx* c, const hj_value* view) {
    if (!view || view->type != HJ_OBJECT) return -1;
    const hj_value* tex = hj_at(hj_get(c->doc, "textures"), (size_t)hj_number(hj_get(view, "index"), -1));
    if (!tex) return -1;
    const hj_value* src = hj_get(tex, "source");
    if (!src || src->type != HJ_NUMBER || src->number < 0.0 || (size_t)src->number >= hj_count(hj_get(c->doc, "images"))) return -1;
    if (textureViewTexcoordSet(view) > 1u) return -1;
    return (int)src->number;
}
static void collectReference(ImportCtx* c, const hj_value* view, uint32_t colorSpace) {
    int image = textureViewImage(c, view);
    if (image < 0) return;
    for (uint32_t i = 0; i < c->refCount; i++)
        if (c->refs[i].image == image && c->refs[i].colorSpace == colorSpace) return;
    if (c->refCount < sizeof(c->refs) / sizeof(c->refs[0])) { c->refs[c->refCount].image = image; c->refs[c->refCount].colorSpace = colorSpace; c->refCount++; }
}
static const hj_value* baseColorView(const hj_value* gm) {  /* loader.c:311-315 */
    const hj_value* sg = hj_get(hj_get(gm, "extensions"), "KHR_materials_pbrSpecularGlossiness");
    if (sg) return hj_get(sg, "diffuseTexture");
    return hj_get(hj_get(gm, "pbrMetallicRoughness"), "baseColorTexture");
}
static int base64Value(int ch) {
    if (ch >= 'A' && ch <= 'Z') return ch - 'A';
    if (ch >= 'a' && ch <= 'z') return ch - 'a' + 26;
    if (ch >= '0' && ch <= '9') return ch - '0' + 52;
    if (ch == '+' || ch == '-') return 62;
    if (ch == '/' || ch == '_') return 63;
    return -1;
}
static unsigned char* decodeDataUri(const char* uri, size_t* outSize, char* mime, size_t mimeLen) {
    const char* comma = strchr(uri, ',');
    if (!comma || !strstr(uri, ";base64")) return NULL;
    size_t ml = (size_t)(strcspn(uri + 5, ";,"));
    snprintf(mime, mimeLen, "%.*s", (int)ml, uri + 5);
    size_t n = strlen(comma + 1);
    unsigned char* out = (unsigned char*)malloc(n * 3 / 4 + 4);
    if (!out) return NULL;
    uint32_t acc = 0;
    int bits = 0;
    size_t o = 0;
    for (const char* p = comma + 1; *p; p++) {
        int v = base64Value((unsigned char)*p);
        if (v < 0) continue;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out[o++] = (unsigned char)((acc >> bits) & 0xffu); }
    }
    *outSize = o;
    return out;
}
static void percentDecode(char* s) {
    char* w = s;
    for (const char* r = s; *r; r++) {
        if (r[0] == '%' && r[1] && r[2]) {
            char hex[3] = {r[1], r[2], 0};
            *w++ = (char)strtol(hex, NULL, 16);
            r += 2;
        } else {
            *w++ = *r;
        }
    }
    *w = 0;
}
/* loader.c:429-455 decodeTextureImage: external file, else bufferView bytes (data: URIs are decoded here as well) */
static int decodeImageReference(ImportCtx* c, int imageIndex, uint32_t colorSpace, GltfTexture* out) {
    const hj_value* image = hj_at(hj_get(c->doc, "images"), (size_t)imageIndex);
    const char* uri = hj_string(hj_get(image, "uri"), NULL);
    const char* mime = hj_string(hj_get(image, "mimeType"), NULL);
    HostImage decoded;
    char why[256] = "";
    int ok = 0;
    if (uri && uri[0] && strncmp(uri, "data:", 5) != 0) {
        char rel[4096], full[8192];
        snprintf(rel, sizeof(rel), "%s", uri);
        percentDecode(rel);
        const char* slash = strrchr(c->sourcePath, '/');
        if (rel[0] == '/' || !slash) snprintf(full, sizeof(full), "%s", rel);
        else snprintf(full, sizeof(full), "%.*s/%s", (int)(slash - c->sourcePath), c->sourcePath, rel);
        ok = hostLoadImageFile(full, colorSpace, &decoded, why, sizeof(why));
    } else if (uri && uri[0]) {
        char uriMime[64] = "";
        size_t n = 0;
        unsigned char* bytes = decodeDataUri(uri, &n, uriMime, sizeof(uriMime));
        if (bytes) ok = hostDecodeImage(bytes, n, mime ? mime : uriMime, "data URI", colorSpace, &decoded, why, sizeof(why));
        free(bytes);
    } else {
        const hj_value* bv = hj_at(hj_get(c->doc, "bufferViews"), (size_t)hj_number(hj_get(image, "bufferView"), -1));
        size_t off = (size_t)hj_number(hj_get(bv, "byteOffset"), 0), len = (size_t)hj_number(hj_get(bv, "byteLength"), 0);
        if (bv && (int)hj_number(hj_get(bv, "buffer"), 0) == 0 && c->bin && len <= c->binSize && off <= c->binSize - len)
            ok = hostDecodeImage(c->bin + off, len, mime, "glTF bufferView", colorSpace, &decoded, why, sizeof(why));
        else snprintf(why, sizeof(why), "image %d has no readable data", imageIndex);
    }
    if (!ok) {
        if (c->error && c->errorSize && !c->error[0]) snprintf(c->error, c->errorSize, "texture image %d: %s", imageIndex, why);
        return 0;
    }
    out->pixels = decoded.pixels; out->width = decoded.width; out->height = decoded.height; out->format = decoded.format; out->colorSpace = decoded.colorSpace;
    const char* name = hj_string(hj_get(image, "name"), NULL);   /* loader.c:457-465: image name, else texture name, else "Texture" */
    if (!name || !name[0]) {
        const hj_value* textures = hj_get(c->doc, "textures");
        for (size_t t = 0; t < hj_count(textures) && (!name || !name[0]); t++)
            if ((int)hj_number(hj_get(hj_at(textures, t), "source"), -1) == imageIndex) name = hj_string(hj_get(hj_at(textures, t), "name"), NULL);
    }
    snprintf(out->name, sizeof(out->name), "%s", name && name[0] ? name : "Texture");
    return 1;
}
/* Binds a texture view to one material slot: index into GltfImport.textures, wrap modes, texcoord set, transform (loader.c:905-940). */
static void bindTexture(ImportCtx* c, Material* m, uint32_t slot, const hj_value* view, uint32_t colorSpace) {
    int image = textureViewImage(c, view);
    if (image < 0) return;
    uint32_t found = VKRT_INVALID_INDEX;
    for (uint32_t i = 0; i < c->refCount; i++)
        if (c->refs[i].image == image && c->refs[i].colorSpace == colorSpace) found = i;
    if (found == VKRT_INVALID_INDEX) return;
    uint32_t* idx[4] = {&m->baseColorTextureIndex, &m->metallicRoughnessTextureIndex, &m->normalTextureIndex, &m->emissiveTextureIndex};
    uint32_t* wrap[4] = {&m->baseColorTextureWrap, &m->metallicRoughnessTextureWrap, &m->normalTextureWrap, &m->emissiveTextureWrap};
    float* xf[4] = {m->baseColorTextureTransform, m->metallicRoughnessTextureTransform, m->normalTextureTransform, m->emissiveTextureTransform};
    *idx[slot] = found;
    const hj_value* tex = hj_at(hj_get(c->doc, "textures"), (size_t)hj_number(hj_get(view, "index"), -1));
    const hj_value* sampler = hj_at(hj_get(c->doc, "samplers"), (size_t)hj_number(hj_get(tex, "sampler"), -1));
    uint32_t ws = (uint32_t)hj_number(hj_get(sampler, "wrapS"), 0), wt = (uint32_t)hj_number(hj_get(sampler, "wrapT"), 0);
    *wrap[slot] = (ws ? ws : VKRT_TEXTURE_WRAP_REPEAT) | ((wt ? wt : VKRT_TEXTURE_WRAP_REPEAT) << 16u);   /* loader.c:359-363 */
    m->textureTexcoordSets = (m->textureTexcoordSets & ~(0xffu << (8u * slot))) | ((textureViewTexcoordSet(view) & 0xffu) << (8u * slot));
    xf[slot][0] = xf[slot][1] = 1.0f; xf[slot][2] = xf[slot][3] = 0.0f;
    m->textureRotations[slot] = 0.0f;
    const hj_value* kt = hj_get(hj_get(view, "extensions"), "KHR_texture_transform");
    if (kt) {   /* loader.c:386-404: scale.xy, offset.xy, rotation */
        float sc[2] = {1.0f, 1.0f}, of[2] = {0.0f, 0.0f};
        hj_floats(hj_get(kt, "scale"), sc, 2);
        hj_floats(hj_get(kt, "offset"), of, 2);
        xf[slot][0] = sc[0]; xf[slot][1] = sc[1]; xf[slot][2] = of[0]; xf[slot][3] = of[1];
        m->textureRotations[slot] = (float)hj_number(hj_get(kt, "rotation"), 0.0);
    }
}

static Material convertMaterial(ImportCtx* c, const hj_value* gm, Material m) {
    if (!gm) return m;
    const hj_value* ext = hj_get(gm, "extensions");
    const hj_value* sg = hj_get(ext, "KHR_materials_pbrSpecularGlossiness");
    const hj_value* pbr = hj_get(gm, "pbrMetallicRoughness");
    float f4[4];
    if (sg) {
        f4[0] = f4[1] = f4[2] = f4[3] = 1.0f;
        hj_floats(hj_get(sg, "diffuseFactor"), f4, 4);
        m.baseColor[0] = f4[0]; m.baseColor[1] = f4[1]; m.baseColor[2] = f4[2]; m.opacity = f4[3];
        m.roughness = 1.0f - (float)hj_number(hj_get(sg, "glossinessFactor"), 1.0);
        bindTexture(c, &m, VKRT_MATERIAL_TEXTURE_SLOT_BASE_COLOR, hj_get(sg, "diffuseTexture"), VKRT_TEXTURE_COLOR_SPACE_SRGB);
    } else if (pbr) {
        f4[0] = f4[1] = f4[2] = f4[3] = 1.0f;
        hj_floats(hj_get(pbr, "baseColorFactor"), f4, 4);
        m.baseColor[0] = f4[0]; m.baseColor[1] = f4[1]; m.baseColor[2] = f4[2]; m.opacity = f4[3];
        m.metallic = (float)hj_number(hj_get(pbr, "metallicFactor"), 1.0);
        m.roughness = (float)hj_number(hj_get(pbr, "roughnessFactor"), 1.0);
        bindTexture(c, &m, VKRT_MATERIAL_TEXTURE_SLOT_BASE_COLOR, hj_get(pbr, "baseColorTexture"), VKRT_TEXTURE_COLOR_SPACE_SRGB);
        bindTexture(c, &m, VKRT_MATERIAL_TEXTURE_SLOT_METALLIC_ROUGHNESS, hj_get(pbr, "metallicRoughnessTexture"), VKRT_TEXTURE_COLOR_SPACE_LINEAR);
    }
    const hj_value* e;
    if ((e = hj_get(ext, "KHR_materials_specular"))) m.specular = (float)hj_number(hj_get(e, "specularFactor"), 1.0);
    if ((e = hj_get(ext, "KHR_materials_ior"))) m.ior = (float)hj_number(hj_get(e, "ior"), 1.5);
    if ((e = hj_get(ext, "KHR_materials_transmission"))) m.transmission = (float)hj_number(hj_get(e, "transmissionFactor"), 0.0);
    if ((e = hj_get(ext, "KHR_materials_volume"))) {
        float col[3] = {1, 1, 1};
        hj_floats(hj_get(e, "attenuationColor"), col, 3);
        memcpy(m.attenuationColor, col, sizeof(col));
        float dist = (float)hj_number(hj_get(e, "attenuationDistance"), INFINITY);
        m.absorptionCoefficient = (dist > 0.0f && isfinite(dist)) ? 1.0f / dist : 0.0f;
    }
    if ((e = hj_get(ext, "KHR_materials_clearcoat"))) {
        m.clearcoat = (float)hj_number(hj_get(e, "clearcoatFactor"), 0.0);
        m.clearcoatGloss = 1.0f - (float)hj_number(hj_get(e, "clearcoatRoughnessFactor"), 0.0);
    }
    if ((e = hj_get(ext, "KHR_materials_sheen"))) {
        float col[3] = {0, 0, 0};
        hj_floats(hj_get(e, "sheenColorFactor"), col, 3);
        float w = max3f(col[0], col[1], col[2]);
        if (w > 0.0f) { m.sheenTintWeight[0] = col[0] / w; m.sheenTintWeight[1] = col[1] / w; m.sheenTintWeight[2] = col[2] / w; m.sheenTintWeight[3] = w; }
        else { m.sheenTintWeight[0] = m.sheenTintWeight[1] = m.sheenTintWeight[2] = 1.0f; m.sheenTintWeight[3] = 0.0f; }
        m.sheenRoughness = (float)hj_number(hj_get(e, "sheenRoughnessFactor"), 0.0);
    }
    {
        float em[3] = {0, 0, 0};
        hj_floats(hj_get(gm, "emissiveFactor"), em, 3);
        float scale = (float)hj_number(hj_get(hj_get(ext, "KHR_materials_emissive_strength"), "emissiveStrength"), 1.0);
        for (int k = 0; k < 3; k++) em[k] *= scale;
        float mx = max3f(em[0], em[1], em[2]);
        if (mx > 0.0f) { m.emissionColor[0] = em[0] / mx; m.emissionColor[1] = em[1] / mx; m.emissionColor[2] = em[2] / mx; m.emissionLuminance = mx; }
        else { m.emissionColor[0] = m.emissionColor[1] = m.emissionColor[2] = 1.0f; m.emissionLuminance = 0.0f; }
    }
    {   /* loader.c:1486-1540: normal map (scale > 0 else 1), then the emissive map */
        const hj_value* nt = hj_get(gm, "normalTexture");
        bindTexture(c, &m, VKRT_MATERIAL_TEXTURE_SLOT_NORMAL, nt, VKRT_TEXTURE_COLOR_SPACE_LINEAR);
        if (m.normalTextureIndex != VKRT_INVALID_INDEX) {
            float scale = (float)hj_number(hj_get(nt, "scale"), 1.0);
            m.normalTextureScale = scale > 0.0f ? scale : 1.0f;
        }
        bindTexture(c, &m, VKRT_MATERIAL_TEXTURE_SLOT_EMISSIVE, hj_get(gm, "emissiveTexture"), VKRT_TEXTURE_COLOR_SPACE_SRGB);
    }
    const char* am = hj_string(hj_get(gm, "alphaMode"), "OPAQUE");
    m.alphaMode = !strcmp(am, "MASK") ? VKRT_MATERIAL_ALPHA_MODE_MASK : (!strcmp(am, "BLEND") ? VKRT_MATERIAL_ALPHA_MODE_BLEND : VKRT_MATERIAL_ALPHA_MODE_OPAQUE);
    /* loader.c:1546-1552: MASK takes the file's cutoff (glTF default 0.5), BLEND a fixed 1/255, OPAQUE keeps the material default */
    if (m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_MASK) m.alphaCutoff = (float)hj_number(hj_get(gm, "alphaCutoff"), 0.5);
    else if (m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_BLEND) m.alphaCutoff = 1.0f / 255.0f;
    return m;
}

/* ---- nodes ---------------------------------------------------------------------------------------------------------------- */
static void nodeLocalMatrix(const hj_value* node, hmat4 out) {
    const hj_value* mv = hj_get(node, "matrix");
    if (hj_count(mv) == 16) {
        float f[16];
        hj_floats(mv, f, 16);
        memcpy(out, f, sizeof(f)); /* glTF matrices are column-major, like hmat4 */
        return;
    }
    float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
    hj_floats(hj_get(node, "translation"), t, 3);
    hj_floats(hj_get(node, "rotation"), q, 4);
    hj_floats(hj_get(node, "scale"), s, 3);
    float x = q[0], y = q[1], z = q[2], w = q[3];
    h_mat4_identity(out);
    out[0][0] = (1 - 2 * (y * y + z * z)) * s[0]; out[0][1] = (2 * (x * y + z * w)) * s[0]; out[0][2] = (2 * (x * z - y * w)) * s[0];
    out[1][0] = (2 * (x * y - z * w)) * s[1]; out[1][1] = (1 - 2 * (x * x + z * z)) * s[1]; out[1][2] = (2 * (y * z + x * w)) * s[1];
    out[2][0] = (2 * (x * z + y * w)) * s[2]; out[2][1] = (2 * (y * z - x * w)) * s[2]; out[2][2] = (1 - 2 * (x * x + y * y)) * s[2];
    out[3][0] = t[0]; out[3][1] = t[1]; out[3][2] = t[2];
}

static int appendMesh(ImportCtx* c, GltfMesh* m) {
    GltfImport* o = c->out;
    GltfMesh* nm = (GltfMesh*)realloc(o->meshes, (size_t)(o->meshCount + 1u) * sizeof(GltfMesh));
    if (!nm) return fail(c, "out of memory");
    o->meshes = nm;
    o->meshes[o->meshCount++] = *m;
    return 1;
}

static int importPrimitive(ImportCtx* c, const hj_value* gmesh, const hj_value* node, const hj_value* prim, size_t primitiveIndex, hmat4 world) {
    if ((int)hj_number(hj_get(prim, "mode"), 4) != 4) return 1; /* triangles only */
    const hj_value* at = hj_get(prim, "attributes");
    const hj_value* posAcc = hj_get(at, "POSITION");
    if (!posAcc) return 1;
    Accessor pos;
    if (!openAccessor(c, (int)hj_number(posAcc, -1), &pos) || pos.components < 3) return fail(c, "glTF: bad POSITION accessor");
    size_t nv = pos.count;
    if (nv == 0) return 1;
    void* vp = NULL;
    if (posix_memalign(&vp, 16, nv * sizeof(Vertex)) != 0) return fail(c, "out of memory");
    Vertex* v = (Vertex*)vp;
    memset(v, 0, nv * sizeof(Vertex));
    for (size_t i = 0; i < nv; i++) {
        float x = readComponent(&pos, i, 0), y = readComponent(&pos, i, 1), z = readComponent(&pos, i, 2);
        v[i].position[0] = x; v[i].position[1] = -z; v[i].position[2] = y; v[i].position[3] = 1.0f;
        v[i].color[0] = v[i].color[1] = v[i].color[2] = v[i].color[3] = 1.0f;
    }
    Accessor a;
    int hasNormals = 0, hasTangents = 0, hasUv[2] = {0, 0};
    const hj_value* h;
    if ((h = hj_get(at, "NORMAL")) && openAccessor(c, (int)hj_number(h, -1), &a) && a.count == nv && a.components >= 3) {
        hasNormals = 1;
        for (size_t i = 0; i < nv; i++) {
            float x = readComponent(&a, i, 0), y = readComponent(&a, i, 1), z = readComponent(&a, i, 2);
            v[i].normal[0] = x; v[i].normal[1] = -z; v[i].normal[2] = y;
        }
    }
    if (hasNormals && (h = hj_get(at, "TANGENT")) && openAccessor(c, (int)hj_number(h, -1), &a) && a.count == nv && a.components >= 4) {
        hasTangents = 1;
        for (size_t i = 0; i < nv; i++) {
            float x = readComponent(&a, i, 0), y = readComponent(&a, i, 1), z = readComponent(&a, i, 2);
            v[i].tangent[0] = x; v[i].tangent[1] = -z; v[i].tangent[2] = y; v[i].tangent[3] = readComponent(&a, i, 3);
        }
    }
    if ((h = hj_get(at, "COLOR_0")) && openAccessor(c, (int)hj_number(h, -1), &a) && a.count == nv) {
        for (size_t i = 0; i < nv; i++)
            for (int k = 0; k < a.components && k < 4; k++) v[i].color[k] = readComponent(&a, i, k);
    }
    for (int set = 0; set < 2; set++) {
        if ((h = hj_get(at, set ? "TEXCOORD_1" : "TEXCOORD_0")) && openAccessor(c, (int)hj_number(h, -1), &a) && a.count == nv && a.components >= 2) {
            hasUv[set] = 1;
            for (size_t i = 0; i < nv; i++) {
                float* uv = set ? v[i].texcoord1 : v[i].texcoord0;
                uv[0] = readComponent(&a, i, 0); uv[1] = readComponent(&a, i, 1);
            }
        }
    }
    size_t ni;
    uint32_t* idx;
    if ((h = hj_get(prim, "indices"))) {
        if (!openAccessor(c, (int)hj_number(h, -1), &a)) { free(v); return 0; }
        ni = a.count;
        idx = (uint32_t*)malloc((ni ? ni : 1) * sizeof(uint32_t));
        if (!idx) { free(v); return fail(c, "out of memory"); }
        for (size_t i = 0; i < ni; i++) {
            idx[i] = readIndex(&a, i); /* indices are always widened to uint32 */
            if (idx[i] >= nv) { free(idx); free(v); return fail(c, "index out of range"); } /* loader.c:1859: the primitive, and with it the import, fails */
        }
    } else {
        ni = nv;
        idx = (uint32_t*)malloc(ni * sizeof(uint32_t));
        if (!idx) { free(v); return fail(c, "out of memory"); }
        for (size_t i = 0; i < ni; i++) idx[i] = (uint32_t)i;
    }
    ni -= ni % 3;
    const hj_value* gmat = hj_at(hj_get(c->doc, "materials"), (size_t)hj_number(hj_get(prim, "material"), -1));
    int normalTexSet = 0;
    if (gmat && hj_get(gmat, "normalTexture")) {
        normalTexSet = (int)hj_number(hj_get(hj_get(gmat, "normalTexture"), "texCoord"), 0);
        if (normalTexSet > 1) normalTexSet = 0;
    }
    if (!hasNormals) {
        /* loader.c:1871-1878: a primitive without normals gets generated normals and NOTHING else: no winding alignment, and its tangents stay
           zero (the shaders build a fallback frame from a zero tangent, geometry/surface.slang) — found by the pin against the reference loader */
        generateNormals(v, nv, idx, ni);
    } else {
        alignWinding(v, nv, idx, ni);
        if (hasTangents) {
            for (size_t i = 0; i < nv; i++) {
                float t[4];
                if (orthonormalizeTangent(v[i].normal, v[i].tangent, v[i].tangent[3], t)) memcpy(v[i].tangent, t, sizeof(t));
                else fallbackTangent(v[i].normal, v[i].tangent);
            }
        } else if (hasUv[normalTexSet]) {
            generateTangents(v, nv, idx, ni, normalTexSet);
        } else {
            for (size_t i = 0; i < nv; i++) fallbackTangent(v[i].normal, v[i].tangent);
        }
    }
    GltfMesh m;
    memset(&m, 0, sizeof(m));
    m.vertices = v; m.vertexCount = nv; m.indices = idx; m.indexCount = ni;
    memcpy(m.world, world, sizeof(hmat4));
    m.materialIndex = gmat ? (int)hj_number(hj_get(prim, "material"), -1) : -1;
    m.doubleSided = gmat ? hj_bool(hj_get(gmat, "doubleSided"), 0) : 0;
    /* loader.c:1097-1118 buildEntryName: the node's name, else the mesh's, else "mesh"; "_<primitive>" when the mesh has several primitives */
    const char* name = hj_string(hj_get(node, "name"), NULL);
    if (!name || !name[0]) name = hj_string(hj_get(gmesh, "name"), NULL);
    if (!name || !name[0]) name = "mesh";
    if (hj_count(hj_get(gmesh, "primitives")) > 1) snprintf(m.name, sizeof(m.name), "%s_%zu", name, primitiveIndex);
    else snprintf(m.name, sizeof(m.name), "%s", name);
    if (!appendMesh(c, &m)) { free(v); free(idx); return 0; }
    return 1;
}

static int visitNode(ImportCtx* c, int nodeIndex, hmat4 parentWorld, int depth) {
    if (depth > 256) return fail(c, "glTF: node hierarchy too deep");
    const hj_value* node = hj_at(hj_get(c->doc, "nodes"), (size_t)nodeIndex);
    if (!node) return fail(c, "glTF: node index out of range");
    hmat4 local, engineLocal, world;
    nodeLocalMatrix(node, local);
    VKRT_buildImportedNodeTransform(local, engineLocal);
    h_mat4_mul(parentWorld, engineLocal, world);
    const hj_value* meshRef = hj_get(node, "mesh");
    if (meshRef) {
        const hj_value* gmesh = hj_at(hj_get(c->doc, "meshes"), (size_t)hj_number(meshRef, -1));
        const hj_value* prims = hj_get(gmesh, "primitives");
        for (size_t p = 0; p < hj_count(prims); p++)
            if (!importPrimitive(c, gmesh, node, hj_at(prims, p), p, world)) return 0;
    }
    const hj_value* children = hj_get(node, "children");
    for (size_t k = 0; k < hj_count(children); k++)
        if (!visitNode(c, (int)hj_number(hj_at(children, k), -1), world, depth + 1)) return 0;
    return 1;
}

void gltfImportFree(GltfImport* imp) {
    if (!imp) return;
    for (uint32_t i = 0; i < imp->meshCount; i++) { free(imp->meshes[i].vertices); free(imp->meshes[i].indices); }
    free(imp->meshes);
    free(imp->materials);
    free(imp->materialNames);
    for (uint32_t i = 0; i < imp->textureCount; i++) free(imp->textures[i].pixels);
    free(imp->textures);
    memset(imp, 0, sizeof(*imp));
}

int gltfImportFile(const char* path, GltfImport* out, char* error, size_t errorSize) {
    memset(out, 0, sizeof(*out));
    if (error && errorSize) error[0] = 0;
    ImportCtx c;
    memset(&c, 0, sizeof(c));
    c.out = out; c.error = error; c.errorSize = errorSize;
    FILE* f = fopen(path, "rb");
    if (!f) { snprintf(error, errorSize, "cannot open %s", path); return 0; }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size < 20 || (unsigned long)size > (1ul << 30)) { fclose(f); snprintf(error, errorSize, "%s: unsupported size", path); return 0; } /* 1 GiB limit, loader.c:27 */
    unsigned char* data = (unsigned char*)malloc((size_t)size);
    if (!data || fread(data, 1, (size_t)size, f) != (size_t)size) { free(data); fclose(f); snprintf(error, errorSize, "%s: read failed", path); return 0; }
    fclose(f);
    uint32_t magic, version, length;
    memcpy(&magic, data, 4); memcpy(&version, data + 4, 4); memcpy(&length, data + 8, 4);
    if (magic != 0x46546C67u || version != 2u) { free(data); snprintf(error, errorSize, "%s: not a glTF 2.0 binary", path); return 0; }
    const char* json = NULL;
    size_t jsonSize = 0;
    size_t off = 12;
    while (off + 8 <= (size_t)size) {
        uint32_t clen, ctype;
        memcpy(&clen, data + off, 4); memcpy(&ctype, data + off + 4, 4);
        if (off + 8 + clen > (size_t)size) break;
        if (ctype == 0x4E4F534Au) { json = (const char*)data + off + 8; jsonSize = clen; }
        else if (ctype == 0x004E4942u && !c.bin) { c.bin = data + off + 8; c.binSize = clen; }
        off += 8 + (size_t)clen;
    }
    if (!json) { free(data); snprintf(error, errorSize, "%s: no JSON chunk", path); return 0; }
    hj_value* doc = hj_parse(json, jsonSize, error, errorSize);
    if (!doc) { free(data); return 0; }
    c.doc = doc;
    int ok = 1;
    /* materials */
    const hj_value* mats = hj_get(doc, "materials");
    uint32_t nm = (uint32_t)hj_count(mats);
    /* textures first: unique (image, colour space) pairs in material order (loader.c:318-357), decoded once each; a file whose image
     * cannot be decoded fails to import, as upstream (loader.c:876-890) */
    c.sourcePath = path;
    for (uint32_t i = 0; i < nm; i++) {
        const hj_value* gm = hj_at(mats, i);
        collectReference(&c, baseColorView(gm), VKRT_TEXTURE_COLOR_SPACE_SRGB);
        collectReference(&c, hj_get(hj_get(gm, "pbrMetallicRoughness"), "metallicRoughnessTexture"), VKRT_TEXTURE_COLOR_SPACE_LINEAR);
        collectReference(&c, hj_get(gm, "normalTexture"), VKRT_TEXTURE_COLOR_SPACE_LINEAR);
        collectReference(&c, hj_get(gm, "emissiveTexture"), VKRT_TEXTURE_COLOR_SPACE_SRGB);
    }
    if (c.refCount) {
        out->textures = (GltfTexture*)calloc(c.refCount, sizeof(GltfTexture));
        if (!out->textures) ok = fail(&c, "out of memory");
        for (uint32_t i = 0; ok && i < c.refCount; i++) {
            ok = decodeImageReference(&c, c.refs[i].image, c.refs[i].colorSpace, &out->textures[i]);
            if (ok) out->textureCount = i + 1u;
        }
    }
    if (ok && nm) {
        out->materials = (Material*)calloc(nm, sizeof(Material));
        out->materialNames = (char(*)[VKRT_NAME_LEN])calloc(nm, VKRT_NAME_LEN);
        if (!out->materials || !out->materialNames) ok = fail(&c, "out of memory");
        for (uint32_t i = 0; ok && i < nm; i++) {
            const hj_value* gm = hj_at(mats, i);
            out->materials[i] = convertMaterial(&c, gm, VKRT_materialDefault());
            const char* name = hj_string(hj_get(gm, "name"), NULL);
            if (name && name[0]) snprintf(out->materialNames[i], VKRT_NAME_LEN, "%s", name);
            else snprintf(out->materialNames[i], VKRT_NAME_LEN, "Material %u", i);
        }
        out->materialCount = nm;
    }
    /* default scene, depth first */
    const hj_value* scenes = hj_get(doc, "scenes");
    const hj_value* scene = hj_at(scenes, (size_t)hj_number(hj_get(doc, "scene"), 0));
    hmat4 identity;
    h_mat4_identity(identity);
    if (ok && scene) {
        const hj_value* roots = hj_get(scene, "nodes");
        for (size_t k = 0; ok && k < hj_count(roots); k++) ok = visitNode(&c, (int)hj_number(hj_at(roots, k), -1), identity, 0);
    } else if (ok) {
        /* no scene: every node that is nobody's child is a root (cgltf has the same fallback upstream) */
        size_t nn = hj_count(hj_get(doc, "nodes"));
        unsigned char* isChild = (unsigned char*)calloc(nn ? nn : 1, 1);
        for (size_t n = 0; isChild && n < nn; n++) {
            const hj_value* ch = hj_get(hj_at(hj_get(doc, "nodes"), n), "children");
            for (size_t k = 0; k < hj_count(ch); k++) {
                size_t ci = (size_t)hj_number(hj_at(ch, k), -1);
                if (ci < nn) isChild[ci] = 1;
            }
        }
        for (size_t n = 0; ok && isChild && n < nn; n++)
            if (!isChild[n]) ok = visitNode(&c, (int)n, identity, 0);
        free(isChild);
    }
    hj_free(doc);
    free(data);
    if (!ok) gltfImportFree(out);
    return ok;
}
