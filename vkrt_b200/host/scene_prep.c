/* scene_prep.c — host-side scene feed: everything the reference computes on the CPU before a frame can be traced.
 *
 * Restates (SURVEY.md §8a-2 "Host geometry", "Host light build", camera, SceneData):
 *   src/core/utility/packing.c:92-156        vertex quantisation (oct normal snorm16x2, oct tangent snorm15x2 + sign, RGBA8)
 *   src/core/scene/geometry.c:166-210,570+   FNV-1a fingerprint + memcmp dedup, owners packed back to back in mesh order
 *   src/core/scene/transform.c:26-35,158-223 T*Rz*Ry*Rx*S compose, lossy PRS decompose, 3x4 world transform
 *   src/core/scene/lighting.c:55-164,267-433 emissive triangle list, Vose alias tables (LIFO stacks, fp32), lightPdfArea
 *   src/core/scene/camera.c:128-143          look-at / perspective / Y flip / inverses
 *   src/core/scene/uniform.c:153-174         SceneData from the scene settings
 *   src/core/api/mesh.c:107-196              material sanitisation
 *   src/core/render/accel/tlas.c:291-296     which instances need the stochastic alpha test
 * The arrays built here are exactly what vkrt_cuda_set_* receives (include/vkrt_cuda.h). */
#include <stdarg.h>

#include "host_state.h"

VKRT_Result hostFail(VKRT* vkrt, VKRT_Result code, const char* fmt, ...) {
    if (vkrt) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(vkrt->error, sizeof(vkrt->error), fmt, ap);
        va_end(ap);
        fprintf(stderr, "[vkrt host] ERROR: %s\n", vkrt->error);
    }
    return code;
}

/* ---- vertex packing -------------------------------------------------------------------------------------------------- */
static float clampUnit(float v) { return v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v); }
static float saturatef(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

static void normalizeOrUp(const float in[3], float out[3]) {
    float lenSq = in[0] * in[0] + in[1] * in[1] + in[2] * in[2];
    if (lenSq > 1e-20f) {
        float inv = 1.0f / sqrtf(lenSq);
        out[0] = in[0] * inv; out[1] = in[1] * inv; out[2] = in[2] * inv;
    } else {
        out[0] = 0.0f; out[1] = 0.0f; out[2] = 1.0f;
    }
}

/* Octahedral projection of a direction onto [-1,1]^2 (lower hemisphere folded over the diagonals). */
static void octProject(const float dir[3], float* px, float* py) {
    float n[3];
    normalizeOrUp(dir, n);
    float invL1 = 1.0f / (fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]));
    float x = n[0] * invL1, y = n[1] * invL1;
    if (n[2] < 0.0f) {
        float ox = x;
        x = (1.0f - fabsf(y)) * (ox >= 0.0f ? 1.0f : -1.0f);
        y = (1.0f - fabsf(ox)) * (y >= 0.0f ? 1.0f : -1.0f);
    }
    *px = x;
    *py = y;
}

static uint32_t packOctNormal(const float normal[3]) {
    float x, y;
    octProject(normal, &x, &y);
    int32_t sx = (int32_t)lroundf(clampUnit(x) * 32767.0f), sy = (int32_t)lroundf(clampUnit(y) * 32767.0f);
    return ((uint32_t)sx & 0xffffu) | (((uint32_t)sy & 0xffffu) << 16);
}
static uint32_t packTangent(const float tangent[4]) {
    float x, y;
    octProject(tangent, &x, &y);
    uint32_t qx = (uint32_t)(int32_t)lroundf(clampUnit(x) * 16383.0f) & 0x7fffu;
    uint32_t qy = (uint32_t)(int32_t)lroundf(clampUnit(y) * 16383.0f) & 0x7fffu;
    uint32_t packed = qx | (qy << 15);
    if (tangent[3] < 0.0f) packed |= 0x80000000u;
    return packed;
}
static uint32_t packColor(const float c[4]) {
    uint32_t r = (uint32_t)lroundf(saturatef(c[0]) * 255.0f), g = (uint32_t)lroundf(saturatef(c[1]) * 255.0f);
    uint32_t b = (uint32_t)lroundf(saturatef(c[2]) * 255.0f), a = (uint32_t)lroundf(saturatef(c[3]) * 255.0f);
    return r | (g << 8) | (b << 16) | (a << 24);
}

void VKRT_packShaderVertex(const Vertex* v, ShaderVertex* out) {
    memset(out, 0, sizeof(*out));
    if (!v) return;
    memcpy(out->position, v->position, sizeof(out->position));
    memcpy(out->texcoord0, v->texcoord0, sizeof(out->texcoord0));
    memcpy(out->texcoord1, v->texcoord1, sizeof(out->texcoord1));
    out->packedNormal = packOctNormal(v->normal);
    out->packedTangent = packTangent(v->tangent);
    out->packedColor = packColor(v->color);
}

/* ---- geometry dedup + layout ------------------------------------------------------------------------------------------- */
static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; i++) {
        h ^= (uint64_t)p[i];
        h *= 1099511628211ull;
    }
    return h;
}
uint64_t hostGeometryFingerprint(const Vertex* vertices, size_t vertexCount, const uint32_t* indices, size_t indexCount) {
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, &vertexCount, sizeof(vertexCount));
    h = fnv1a(h, &indexCount, sizeof(indexCount));
    h = fnv1a(h, vertices, vertexCount * sizeof(Vertex));
    h = fnv1a(h, indices, indexCount * sizeof(uint32_t));
    return h;
}

#define GROW(ptr, cap, need, type)                                                     \
    do {                                                                               \
        if ((need) > (cap)) {                                                          \
            size_t ncap_ = (size_t)(need) + (size_t)(need) / 2 + 16;                   \
            void* np_ = realloc((ptr), ncap_ * sizeof(type));                          \
            if (!np_) return hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory"); \
            (ptr) = (type*)np_;                                                        \
            (cap) = (uint32_t)ncap_;                                                   \
        }                                                                              \
    } while (0)

/* Owners' packed vertices / indices back to back in mesh order; every mesh gets its owner's bases. */
VKRT_Result hostPrepareGeometry(VKRT* vkrt) {
    uint64_t nv = 0, ni = 0;
    for (uint32_t i = 0; i < vkrt->meshCount; i++)
        if (vkrt->meshes[i].ownsGeometry) { nv += vkrt->meshes[i].info.vertexCount; ni += vkrt->meshes[i].info.indexCount; }
    if (nv > 0xffffffffull || ni > 0xffffffffull) return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "geometry exceeds 32-bit buffer addressing");
    GROW(vkrt->packedVertices, vkrt->packedVertexCapacity, (uint32_t)nv + 1u, ShaderVertex);
    GROW(vkrt->packedIndices, vkrt->packedIndexCapacity, (uint32_t)ni + 1u, uint32_t);
    uint32_t vbase = 0, ibase = 0;
    for (uint32_t i = 0; i < vkrt->meshCount; i++) {
        HostMesh* m = &vkrt->meshes[i];
        if (!m->ownsGeometry) continue;
        m->info.vertexBase = vbase;
        m->info.indexBase = ibase;
        for (uint32_t k = 0; k < m->info.vertexCount; k++) VKRT_packShaderVertex(&m->vertices[k], &vkrt->packedVertices[vbase + k]);
        memcpy(vkrt->packedIndices + ibase, m->indices, (size_t)m->info.indexCount * sizeof(uint32_t));
        vbase += m->info.vertexCount;
        ibase += m->info.indexCount;
    }
    for (uint32_t i = 0; i < vkrt->meshCount; i++) {
        HostMesh* m = &vkrt->meshes[i];
        if (m->ownsGeometry) continue;
        const HostMesh* src = &vkrt->meshes[m->geometrySource];
        m->info.vertexBase = src->info.vertexBase;
        m->info.indexBase = src->info.indexBase;
    }
    vkrt->packedVertexCount = vbase;
    vkrt->packedIndexCount = ibase;
    return VKRT_SUCCESS;
}

int hostMaterialMayRejectRayHit(const Material* material, float meshOpacity) {
    if (!material) return 1;
    if (material->alphaMode != VKRT_MATERIAL_ALPHA_MODE_OPAQUE) return 1;
    return material->opacity < 0.999f || meshOpacity < 0.999f;
}

VKRT_Result hostPrepareMeshInfos(VKRT* vkrt) {
    uint32_t n = vkrt->meshCount;
    if (n + 1u > vkrt->preparedMeshCapacity) {
        uint32_t cap = n + n / 2 + 16;
        MeshInfo* a = (MeshInfo*)realloc(vkrt->meshInfos, (size_t)cap * sizeof(MeshInfo));
        float* b = (float*)realloc(vkrt->world3x4, (size_t)cap * 12 * sizeof(float));
        uint32_t* c = (uint32_t*)realloc(vkrt->geometrySource, (size_t)cap * sizeof(uint32_t));
        uint8_t* d = (uint8_t*)realloc(vkrt->alphaTested, (size_t)cap);
        if (a) vkrt->meshInfos = a;
        if (b) vkrt->world3x4 = b;
        if (c) vkrt->geometrySource = c;
        if (d) vkrt->alphaTested = d;
        if (!a || !b || !c || !d) return hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory");
        vkrt->preparedMeshCapacity = cap;
    }
    for (uint32_t i = 0; i < n; i++) {
        const HostMesh* m = &vkrt->meshes[i];
        vkrt->meshInfos[i] = m->info;
        /* getMeshWorldTransform (transform.c:212-223): row-major 3x4 = transpose of the column-major upper 3 rows */
        for (int row = 0; row < 3; row++)
            for (int col = 0; col < 4; col++) vkrt->world3x4[(size_t)i * 12 + row * 4 + col] = m->worldTransform[col][row];
        vkrt->geometrySource[i] = m->geometrySource;
        const Material* mat = m->info.materialIndex < vkrt->materialCount ? &vkrt->materials[m->info.materialIndex].material : NULL;
        vkrt->alphaTested[i] = (uint8_t)(hostMaterialMayRejectRayHit(mat, m->info.opacity) ? 1 : 0);
    }
    return VKRT_SUCCESS;
}

VKRT_Result hostPrepareMaterials(VKRT* vkrt) {
    uint32_t n = vkrt->materialCount;
    GROW(vkrt->materialArray, vkrt->materialArrayCapacity, n + 1u, Material);
    for (uint32_t i = 0; i < n; i++) vkrt->materialArray[i] = vkrt->materials[i].material;
    return VKRT_SUCCESS;
}

/* ---- transforms ------------------------------------------------------------------------------------------------------- */
void VKRT_buildMeshTransformMatrix(const vkrt_vec3 position, const vkrt_vec3 rotationDegrees, const vkrt_vec3 scale, vkrt_mat4 out) {
    if (!position || !rotationDegrees || !scale || !out) return;
    const float ax[3] = {1.0f, 0.0f, 0.0f}, ay[3] = {0.0f, 1.0f, 0.0f}, az[3] = {0.0f, 0.0f, 1.0f};
    h_mat4_identity(out);
    h_translate(out, position);
    h_rotate(out, h_rad(rotationDegrees[2]), az);
    h_rotate(out, h_rad(rotationDegrees[1]), ay);
    h_rotate(out, h_rad(rotationDegrees[0]), ax);
    h_scale(out, scale);
}

static const float kTransformEpsilon = 1e-6f;

static void eulerZYX(hmat4 r, float out[3]) {
    float sineY = -r[0][2];
    if (sineY < -1.0f) sineY = -1.0f;
    else if (sineY > 1.0f) sineY = 1.0f;
    out[1] = asinf(sineY);
    if (fabsf(cosf(out[1])) > kTransformEpsilon) {
        out[0] = atan2f(r[1][2], r[2][2]);
        out[2] = atan2f(r[0][1], r[0][0]);
    } else {
        out[0] = atan2f(-r[2][1], r[1][1]);
        out[2] = 0.0f;
    }
}

void VKRT_decomposeMeshTransform(vkrt_mat4 world, vkrt_vec3 outPosition, vkrt_vec3 outRotation, vkrt_vec3 outScale) {
    if (!outPosition || !outRotation || !outScale) return;
    outPosition[0] = world[3][0]; outPosition[1] = world[3][1]; outPosition[2] = world[3][2];
    hmat4 rot;
    float absScale[3] = {1.0f, 1.0f, 1.0f};
    h_mat4_identity(rot);
    for (int axis = 0; axis < 3; axis++) {
        float col[3] = {world[axis][0], world[axis][1], world[axis][2]};
        float s = h_norm3(col);
        if (s < kTransformEpsilon || !isfinite(s)) continue; /* keeps the identity column and scale 1 */
        absScale[axis] = s;
        rot[axis][0] = col[0] / s; rot[axis][1] = col[1] / s; rot[axis][2] = col[2] / s;
    }
    float c01[3];
    h_cross3(rot[0], rot[1], c01);
    float det = h_dot3(c01, rot[2]);
    int candidates = det < 0.0f ? 3 : 1;
    float bestError = INFINITY, bestRot[3] = {0, 0, 0}, bestScale[3] = {absScale[0], absScale[1], absScale[2]};
    for (int k = 0; k < candidates; k++) {
        int flipped = det < 0.0f ? k : -1;
        hmat4 cr;
        float cs[3] = {absScale[0], absScale[1], absScale[2]};
        h_mat4_copy(rot, cr);
        if (flipped >= 0) {
            cs[flipped] = -cs[flipped];
            cr[flipped][0] = -cr[flipped][0]; cr[flipped][1] = -cr[flipped][1]; cr[flipped][2] = -cr[flipped][2];
        }
        float rad[3], deg[3];
        eulerZYX(cr, rad);
        deg[0] = h_deg(rad[0]); deg[1] = h_deg(rad[1]); deg[2] = h_deg(rad[2]);
        hmat4 re;
        VKRT_buildMeshTransformMatrix(outPosition, deg, cs, re);
        float err = 0.0f;
        for (int col = 0; col < 4; col++)
            for (int row = 0; row < 3; row++) {
                float d = fabsf(world[col][row] - re[col][row]);
                if (d > err) err = d;
            }
        if (err < bestError) {
            bestError = err;
            memcpy(bestRot, deg, sizeof(bestRot));
            memcpy(bestScale, cs, sizeof(bestScale));
        }
    }
    for (int a = 0; a < 3; a++) { outRotation[a] = bestRot[a]; outScale[a] = bestScale[a]; }
}

/* glTF is Y-up; the engine is Z-up: M_engine = B * M * B^-1 with B = Rx(90 deg) (transform.c:84-111). */
void VKRT_buildImportedNodeTransform(vkrt_mat4 world, vkrt_mat4 outEngine) {
    if (!world || !outEngine) return;
    const float zero[3] = {0, 0, 0}, basisRot[3] = {90.0f, 0.0f, 0.0f}, one[3] = {1, 1, 1};
    hmat4 b, binv, t;
    VKRT_buildMeshTransformMatrix(zero, basisRot, one, b);
    h_mat4_inv(b, binv);
    h_mat4_mul(b, world, t);
    h_mat4_mul(t, binv, outEngine);
}
void VKRT_decomposeMeshNodeTransform(vkrt_mat4 world, vkrt_vec3 outPosition, vkrt_vec3 outRotation, vkrt_vec3 outScale) {
    if (!outPosition || !outRotation || !outScale) return;
    hmat4 e;
    VKRT_buildImportedNodeTransform(world, e);
    VKRT_decomposeMeshTransform(e, outPosition, outRotation, outScale);
}

/* ---- materials -------------------------------------------------------------------------------------------------------- */
Material VKRT_materialDefault(void) {
    Material m;
    memset(&m, 0, sizeof(m));
    m.baseColor[0] = m.baseColor[1] = m.baseColor[2] = 0.8f;
    m.roughness = 0.5f;
    m.emissionColor[0] = m.emissionColor[1] = m.emissionColor[2] = 1.0f;
    m.specular = 0.5f;
    m.sheenTintWeight[0] = m.sheenTintWeight[1] = m.sheenTintWeight[2] = 1.0f;
    m.clearcoatGloss = 1.0f;
    m.ior = 1.5f;
    m.sheenRoughness = 0.5f;
    m.attenuationColor[0] = m.attenuationColor[1] = m.attenuationColor[2] = 1.0f;
    m.normalTextureScale = 1.0f;
    m.baseColorTextureIndex = m.metallicRoughnessTextureIndex = m.normalTextureIndex = m.emissiveTextureIndex = VKRT_INVALID_INDEX;
    m.opacity = 1.0f;
    m.alphaCutoff = 0.5f;
    m.alphaMode = VKRT_MATERIAL_ALPHA_MODE_OPAQUE;
    float* tr[4] = {m.baseColorTextureTransform, m.metallicRoughnessTextureTransform, m.normalTextureTransform, m.emissiveTextureTransform};
    for (int i = 0; i < 4; i++) { tr[i][0] = 1.0f; tr[i][1] = 1.0f; tr[i][2] = 0.0f; tr[i][3] = 0.0f; }
    return m;
}

static void sanitizeTextureSlot(const VKRT* vkrt, uint32_t slot, uint32_t* index, uint32_t* wrap, uint32_t* texcoordSets, float* transform,
                                float* rotation, uint32_t expectedColorSpace) {
    int valid = 0;
    if (*index != VKRT_INVALID_INDEX && vkrt && *index < vkrt->textureCount) {
        uint32_t cs = vkrt->textures[*index].colorSpace;
        valid = expectedColorSpace == VKRT_TEXTURE_COLOR_SPACE_SRGB ? (cs == VKRT_TEXTURE_COLOR_SPACE_SRGB || cs == VKRT_TEXTURE_COLOR_SPACE_LINEAR)
                                                                    : cs == VKRT_TEXTURE_COLOR_SPACE_LINEAR;
    }
    if (!valid) { /* unbound slot: wrap 0, texcoord set 0, identity transform */
        *index = VKRT_INVALID_INDEX;
        *wrap = 0u;
        *texcoordSets &= ~(0xffu << (slot * 8u));
        transform[0] = 1.0f; transform[1] = 1.0f; transform[2] = 0.0f; transform[3] = 0.0f;
        *rotation = 0.0f;
        return;
    }
    if (*wrap == 0u) *wrap = VKRT_TEXTURE_WRAP_DEFAULT;
    if (!isfinite(*rotation)) *rotation = 0.0f;
    for (int i = 0; i < 4; i++)
        if (!isfinite(transform[i])) transform[i] = i < 2 ? 1.0f : 0.0f;
    if (((*texcoordSets >> (slot * 8u)) & 0xffu) > 1u) *texcoordSets &= ~(0xffu << (slot * 8u));
}

Material hostSanitizeMaterial(const VKRT* vkrt, Material m) {
    for (int i = 0; i < 3; i++) {
        m.baseColor[i] = hostFiniteClampf(m.baseColor[i], 0.0f, 0.0f, 1.0f);
        m.emissionColor[i] = hostFiniteClampf(m.emissionColor[i], 0.0f, 0.0f, INFINITY);
        m.sheenTintWeight[i] = hostFiniteClampf(m.sheenTintWeight[i], 0.0f, 0.0f, 1.0f);
        m.attenuationColor[i] = hostFiniteClampf(m.attenuationColor[i], 1.0f, 0.0f, 1.0f);
        m.eta[i] = hostFiniteClampf(m.eta[i], 0.0f, 0.0f, INFINITY);
        m.k[i] = hostFiniteClampf(m.k[i], 0.0f, 0.0f, INFINITY);
    }
    m.metallic = hostFiniteClampf(m.metallic, 0.0f, 0.0f, 1.0f);
    m.roughness = hostFiniteClampf(m.roughness, 0.0f, 0.0f, 1.0f);
    m.diffuseRoughness = hostFiniteClampf(m.diffuseRoughness, 0.0f, 0.0f, 1.0f);
    m.specular = hostFiniteClampf(m.specular, 0.0f, 0.0f, 1.0f);
    m.specularTint = hostFiniteClampf(m.specularTint, 0.0f, 0.0f, 1.0f);
    m.anisotropic = hostFiniteClampf(m.anisotropic, 0.0f, 0.0f, 1.0f);
    m.sheenTintWeight[3] = hostFiniteClampf(m.sheenTintWeight[3], 0.0f, 0.0f, 1.0f);
    m.clearcoat = hostFiniteClampf(m.clearcoat, 0.0f, 0.0f, 1.0f);
    m.clearcoatGloss = hostFiniteClampf(m.clearcoatGloss, 0.0f, 0.0f, 1.0f);
    m.ior = hostFiniteClampf(m.ior, 1.0f, 1.0f, 4.0f);
    m.abbeNumber = hostFiniteClampf(m.abbeNumber, 0.0f, 0.0f, 200.0f);
    m.transmission = hostFiniteClampf(m.transmission, 0.0f, 0.0f, 1.0f);
    m.subsurface = hostFiniteClampf(m.subsurface, 0.0f, 0.0f, 1.0f);
    m.sheenRoughness = hostFiniteClampf(m.sheenRoughness, 0.0f, 0.0f, 1.0f);
    m.absorptionCoefficient = hostFiniteClampf(m.absorptionCoefficient, 0.0f, 0.0f, VKRT_MAX_ABSORPTION_COEFFICIENT);
    m.emissionLuminance = hostFiniteClampf(m.emissionLuminance, 0.0f, 0.0f, INFINITY);
    m.normalTextureScale = hostFiniteOrf(m.normalTextureScale, 1.0f);
    if (m.normalTextureScale < 0.0f) m.normalTextureScale = 0.0f;
    m.opacity = hostFiniteClampf(m.opacity, 1.0f, 0.0f, 1.0f);
    m.alphaCutoff = hostFiniteClampf(m.alphaCutoff, 0.5f, 0.0f, 1.0f);
    if (m.alphaMode != VKRT_MATERIAL_ALPHA_MODE_MASK && m.alphaMode != VKRT_MATERIAL_ALPHA_MODE_BLEND) m.alphaMode = VKRT_MATERIAL_ALPHA_MODE_OPAQUE;
    sanitizeTextureSlot(vkrt, 0, &m.baseColorTextureIndex, &m.baseColorTextureWrap, &m.textureTexcoordSets, m.baseColorTextureTransform,
                        &m.textureRotations[0], VKRT_TEXTURE_COLOR_SPACE_SRGB);
    sanitizeTextureSlot(vkrt, 1, &m.metallicRoughnessTextureIndex, &m.metallicRoughnessTextureWrap, &m.textureTexcoordSets,
                        m.metallicRoughnessTextureTransform, &m.textureRotations[1], VKRT_TEXTURE_COLOR_SPACE_LINEAR);
    sanitizeTextureSlot(vkrt, 2, &m.normalTextureIndex, &m.normalTextureWrap, &m.textureTexcoordSets, m.normalTextureTransform,
                        &m.textureRotations[2], VKRT_TEXTURE_COLOR_SPACE_LINEAR);
    sanitizeTextureSlot(vkrt, 3, &m.emissiveTextureIndex, &m.emissiveTextureWrap, &m.textureTexcoordSets, m.emissiveTextureTransform,
                        &m.textureRotations[3], VKRT_TEXTURE_COLOR_SPACE_SRGB);
    return m;
}

/* ---- lights ----------------------------------------------------------------------------------------------------------- */
/* Vose's alias method with two LIFO work stacks, all fp32 (lighting.c:108-164). */
int hostBuildAliasTable(const float* pmf, uint32_t count, float* outQ, uint32_t* outIdx) {
    if (!pmf || !outQ || !outIdx || count == 0) return 0;
    float* scaled = (float*)malloc((size_t)count * sizeof(float));
    uint32_t* small = (uint32_t*)malloc((size_t)count * sizeof(uint32_t));
    uint32_t* large = (uint32_t*)malloc((size_t)count * sizeof(uint32_t));
    if (!scaled || !small || !large) { free(scaled); free(small); free(large); return 0; }
    uint32_t ns = 0, nl = 0;
    for (uint32_t i = 0; i < count; i++) {
        scaled[i] = pmf[i] * (float)count;
        if (scaled[i] < 1.0f) small[ns++] = i;
        else large[nl++] = i;
    }
    while (ns > 0 && nl > 0) {
        uint32_t s = small[--ns], l = large[--nl];
        outQ[s] = scaled[s];
        outIdx[s] = l;
        scaled[l] = (scaled[l] + scaled[s]) - 1.0f;
        if (scaled[l] < 1.0f) small[ns++] = l;
        else large[nl++] = l;
    }
    while (nl > 0) { uint32_t l = large[--nl]; outQ[l] = 1.0f; outIdx[l] = l; }
    while (ns > 0) { uint32_t s = small[--ns]; outQ[s] = 1.0f; outIdx[s] = s; }
    free(scaled); free(small); free(large);
    return 1;
}

/* test hook: the exact table builder used for the light tables */
VKRT_HOST_API int VKRT_hostBuildAliasTable(const float* pmf, uint32_t count, float* outQ, uint32_t* outIdx) { return hostBuildAliasTable(pmf, count, outQ, outIdx); }

static float emissionWeight(const Material* m) {
    if (!isfinite(m->emissionLuminance) || m->emissionLuminance <= 0.0f) return 0.0f;
    float lum = 0.2126f * m->emissionColor[0] + 0.7152f * m->emissionColor[1] + 0.0722f * m->emissionColor[2];
    if (lum <= 0.0f) return 0.0f;
    return lum * m->emissionLuminance;
}
static int eligibleForDirectLightSampling(const MeshInfo* info, const Material* m) {
    if (info->opacity < 0.999f || m->opacity < 0.999f) return 0;
    if (m->emissiveTextureIndex != VKRT_INVALID_INDEX) return 0;
    return m->alphaMode == VKRT_MATERIAL_ALPHA_MODE_OPAQUE;
}
static void toWorld(const HostMesh* mesh, const float p[4], float out[3]) {
    /* row r of the 3x4 world transform = worldTransform[col][r]; terms summed left to right */
    for (int r = 0; r < 3; r++)
        out[r] = (mesh->worldTransform[0][r] * p[0]) + (mesh->worldTransform[1][r] * p[1]) + (mesh->worldTransform[2][r] * p[2]) + mesh->worldTransform[3][r];
}

VKRT_Result hostRebuildLights(VKRT* vkrt) {
    /* capacity: every triangle of every candidate mesh */
    uint32_t meshCap = 0;
    uint64_t triCap = 0;
    for (uint32_t i = 0; i < vkrt->meshCount; i++) {
        vkrt->meshes[i].info.lightPdfArea = 0.0f;
        const HostMesh* mesh = &vkrt->meshes[i];
        if (mesh->info.materialIndex >= vkrt->materialCount) continue;
        const Material* mat = &vkrt->materials[mesh->info.materialIndex].material;
        if (!eligibleForDirectLightSampling(&mesh->info, mat) || emissionWeight(mat) <= 0.0f) continue;
        meshCap++;
        triCap += mesh->info.indexCount / 3u;
    }
    if (triCap > 0xfffffff0ull) return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "too many emissive triangles");
    if (meshCap + 1u > vkrt->emissiveMeshCapacity) {
        uint32_t cap = meshCap + 16u;
        vkrt->emissiveMeshes = (EmissiveMesh*)realloc(vkrt->emissiveMeshes, (size_t)cap * sizeof(EmissiveMesh));
        vkrt->meshAliasQ = (float*)realloc(vkrt->meshAliasQ, (size_t)cap * sizeof(float));
        vkrt->meshAliasIdx = (uint32_t*)realloc(vkrt->meshAliasIdx, (size_t)cap * sizeof(uint32_t));
        if (!vkrt->emissiveMeshes || !vkrt->meshAliasQ || !vkrt->meshAliasIdx) return hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory");
        vkrt->emissiveMeshCapacity = cap;
    }
    if ((uint32_t)triCap + 1u > vkrt->emissiveTriangleCapacity) {
        uint32_t cap = (uint32_t)triCap + 16u;
        void* a = NULL;
        if (posix_memalign(&a, 16, (size_t)cap * sizeof(EmissiveTriangle)) != 0) return hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory");
        free(vkrt->emissiveTriangles);
        vkrt->emissiveTriangles = (EmissiveTriangle*)a;
        vkrt->triAliasQ = (float*)realloc(vkrt->triAliasQ, (size_t)cap * sizeof(float));
        vkrt->triAliasIdx = (uint32_t*)realloc(vkrt->triAliasIdx, (size_t)cap * sizeof(uint32_t));
        if (!vkrt->triAliasQ || !vkrt->triAliasIdx) return hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory");
        vkrt->emissiveTriangleCapacity = cap;
    }
    float* meshWeights = (float*)malloc(((size_t)meshCap + 1u) * sizeof(float));
    uint32_t* sourceMesh = (uint32_t*)malloc(((size_t)meshCap + 1u) * sizeof(uint32_t));
    float* pmf = (float*)malloc(((size_t)(triCap > meshCap ? triCap : meshCap) + 1u) * sizeof(float));
    if (!meshWeights || !sourceMesh || !pmf) { free(meshWeights); free(sourceMesh); free(pmf); return hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory"); }

    uint32_t nMesh = 0, nTri = 0;
    float totalSelectionWeight = 0.0f;
    VKRT_Result rc = VKRT_SUCCESS;
    for (uint32_t mi = 0; mi < vkrt->meshCount && rc == VKRT_SUCCESS; mi++) {
        const HostMesh* mesh = &vkrt->meshes[mi];
        if (mesh->info.materialIndex >= vkrt->materialCount) continue;
        const Material* mat = &vkrt->materials[mesh->info.materialIndex].material;
        if (!eligibleForDirectLightSampling(&mesh->info, mat)) continue;
        float ew = emissionWeight(mat);
        if (ew <= 0.0f) continue;
        const HostMesh* geo = &vkrt->meshes[mesh->geometrySource];
        if (!geo->vertices || !geo->indices) continue;
        uint32_t triCount = mesh->info.indexCount / 3u;
        if (triCount == 0u) continue;
        uint32_t triOffset = nTri;
        float totalArea = 0.0f;
        for (uint32_t t = 0; t < triCount; t++) {
            uint32_t i0 = geo->indices[t * 3u], i1 = geo->indices[t * 3u + 1u], i2 = geo->indices[t * 3u + 2u];
            if (i0 >= mesh->info.vertexCount || i1 >= mesh->info.vertexCount || i2 >= mesh->info.vertexCount) continue;
            float p0[3], p1[3], p2[3], e1[3], e2[3], cr[3];
            toWorld(mesh, geo->vertices[i0].position, p0);
            toWorld(mesh, geo->vertices[i1].position, p1);
            toWorld(mesh, geo->vertices[i2].position, p2);
            for (int k = 0; k < 3; k++) { e1[k] = p1[k] - p0[k]; e2[k] = p2[k] - p0[k]; }
            h_cross3(e1, e2, cr);
            float area = 0.5f * h_norm3(cr);
            if (!(area > 0.0f) || !isfinite(area)) continue;   /* lighting.c:302 skips area <= 0; a NaN / infinite area (non-finite vertex: an inactive triangle) is skipped too */
            EmissiveTriangle* et = &vkrt->emissiveTriangles[nTri++];
            memset(et, 0, sizeof(*et));
            et->v0Area[0] = p0[0]; et->v0Area[1] = p0[1]; et->v0Area[2] = p0[2]; et->v0Area[3] = area;
            et->e1Pad[0] = e1[0]; et->e1Pad[1] = e1[1]; et->e1Pad[2] = e1[2];
            et->e2Pad[0] = e2[0]; et->e2Pad[1] = e2[1]; et->e2Pad[2] = e2[2];
            totalArea += area;
        }
        uint32_t valid = nTri - triOffset;
        float selectionWeight = totalArea * ew;
        if (!(selectionWeight > 0.0f) || !isfinite(selectionWeight) || valid == 0u) { nTri = triOffset; continue; }
        float invTotalArea = 1.0f / totalArea;
        for (uint32_t t = 0; t < valid; t++) pmf[t] = vkrt->emissiveTriangles[triOffset + t].v0Area[3] * invTotalArea;
        if (!hostBuildAliasTable(pmf, valid, vkrt->triAliasQ + triOffset, vkrt->triAliasIdx + triOffset)) { rc = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "triangle alias table"); break; }
        EmissiveMesh em;
        memset(&em, 0, sizeof(em));
        em.triOffset = triOffset;
        em.triCount = valid;
        em.invTotalArea = invTotalArea;
        for (int k = 0; k < 3; k++) em.emission[k] = mat->emissionColor[k] * mat->emissionLuminance;
        meshWeights[nMesh] = selectionWeight;
        totalSelectionWeight += selectionWeight;
        sourceMesh[nMesh] = mi;
        vkrt->emissiveMeshes[nMesh++] = em;
    }
    if (rc == VKRT_SUCCESS && nMesh > 0 && totalSelectionWeight > 0.0f) {
        float invTotal = 1.0f / totalSelectionWeight;
        for (uint32_t k = 0; k < nMesh; k++) {
            float p = meshWeights[k] * invTotal;
            vkrt->emissiveMeshes[k].pmfMesh = p;
            pmf[k] = p;
            vkrt->meshes[sourceMesh[k]].info.lightPdfArea = p * vkrt->emissiveMeshes[k].invTotalArea;
        }
        if (!hostBuildAliasTable(pmf, nMesh, vkrt->meshAliasQ, vkrt->meshAliasIdx)) rc = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "mesh alias table");
    }
    free(meshWeights); free(sourceMesh); free(pmf);
    if (rc != VKRT_SUCCESS) return rc;
    vkrt->emissiveMeshCount = nMesh;
    vkrt->emissiveTriangleCount = nTri;
    vkrt->sceneData.emissiveMeshCount = nMesh;
    vkrt->sceneData.emissiveTriangleCount = nTri;
    return VKRT_SUCCESS;
}

/* ---- camera + SceneData ------------------------------------------------------------------------------------------------ */
void hostSyncCameraMatrices(VKRT* vkrt) {
    hmat4 view, proj, vi, pi;
    Camera cam = vkrt->sceneSettings.camera;
    uint32_t w = vkrt->sceneData.viewportRect[2] > 0u ? vkrt->sceneData.viewportRect[2] : 1u;
    uint32_t h = vkrt->sceneData.viewportRect[3] > 0u ? vkrt->sceneData.viewportRect[3] : 1u;
    h_lookat(cam.pos, cam.target, cam.up, view);
    h_perspective(h_rad(cam.vfov), (float)w / (float)h, cam.nearZ, cam.farZ, proj);
    proj[1][1] *= -1.0f;
    h_mat4_inv(view, vi);
    h_mat4_inv(proj, pi);
    memcpy(vkrt->sceneData.viewInverse, vi, sizeof(vi));
    memcpy(vkrt->sceneData.projInverse, pi, sizeof(pi));
}

void hostWriteSceneStateUniform(VKRT* vkrt) {
    const VKRT_SceneSettingsSnapshot* s = &vkrt->sceneSettings;
    SceneData* sd = &vkrt->sceneData;
    sd->samplesPerPixel = s->samplesPerPixel > 0u ? s->samplesPerPixel : 1u;
    sd->rrMaxDepth = s->rrMaxDepth;
    sd->rrMinDepth = s->rrMinDepth;
    sd->packedRenderSettings = VKRT_PACK_RENDER_SETTINGS(s->toneMappingMode, s->renderMode, s->spectralSamplingMode);
    sd->exposure = s->exposure;
    sd->timeBase = s->timeBase;
    sd->timeStep = s->timeStep;
    sd->environmentLight[0] = s->environmentColor[0] * s->environmentStrength;
    sd->environmentLight[1] = s->environmentColor[1] * s->environmentStrength;
    sd->environmentLight[2] = s->environmentColor[2] * s->environmentStrength;
    sd->environmentLight[3] = s->environmentStrength;
    sd->environmentTextureIndex = s->environmentTextureIndex;
    sd->environmentRotation = s->environmentRotation;
    sd->debugMode = s->debugMode;
    sd->misNeeEnabled = s->misNeeEnabled ? 1u : 0u;
    sd->selectionEnabled = 0u; /* editor picking is out of scope */
    sd->selectedMeshIndex = VKRT_INVALID_INDEX;
    sd->rgb2specSRGB = vkrt->rgb2specInfo;
}

/* resetSceneData (uniform.c:227-245): restart accumulation */
void hostResetSceneData(VKRT* vkrt) {   /* scene/uniform.c:227-245 */
    hostResetAutoSPPState(vkrt, 1);
    vkrt->autoExposureFilteredLuminance = 0.0f;
    vkrt->sceneData.frameNumber = 0;
    vkrt->renderStatus.renderPhase = vkrt->renderStatus.renderPhase != VKRT_RENDER_PHASE_INACTIVE ? VKRT_RENDER_PHASE_SAMPLING : VKRT_RENDER_PHASE_INACTIVE;
    vkrt->renderStatus.accumulationFrame = 0;
    vkrt->renderStatus.totalSamples = 0;
    vkrt->renderStatus.averageFrametime = 0.0f;
    vkrt->accumulationNeedsReset = 1;
    hostWriteSceneStateUniform(vkrt);
    vkrt->sceneData.emissiveMeshCount = vkrt->emissiveMeshCount;
    vkrt->sceneData.emissiveTriangleCount = vkrt->emissiveTriangleCount;
    memset(vkrt->renderStatus.frametimes, 0, sizeof(vkrt->renderStatus.frametimes));
}
