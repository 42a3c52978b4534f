// scene_view.cuh — device view of the scene buffers (the reference's descriptor set 0, src/shaders/scene/resources.slang:7-59)
// plus geometry fetch (geometry/surface/interpolation.slang), software texture sampling (material/textures.slang) and the
// stochastic alpha test (rt/alpha_test.slang).
#pragma once
#include "accel.cuh"
#include "shading.cuh"

namespace vk {

struct TextureView {
    const uint8_t* pixels;
    uint32_t width, height, format, colorSpace;
};

// Optional extension (not in the reference, which looks the environment up on a miss only — light/environment.slang:16-27): a
// piecewise-constant distribution over the texels of the lat-long environment map, weight = luminance x sin(theta), sampled through
// an alias table by the next-event estimation and combined with BSDF sampling by the same MIS as the emissive triangles. Off unless
// the context was created with VKRT_CUDA_FLAG_ENV_IMPORTANCE and the scene has an environment texture.
struct EnvDistribution {
    const float* aliasQ;        // [width * height]
    const uint32_t* aliasIdx;
    const float* pdfUv;         // probability density over the unit (u, v) square: pmf x width x height
    uint32_t width, height;
    float pEnv;                 // probability that a vertex's light sample goes to the environment (1 when there are no emissive triangles)
    uint32_t active;
};

struct SceneView {
    const ShaderVertex* vertices;
    const uint32_t* indices;
    const MeshInfo* meshInfos;
    const MeshTrig* meshTrig;  // per mesh, written by k_mesh_trig at set_instances time
    const Material* materials;
    const EmissiveMesh* emissiveMeshes;
    const EmissiveTriangle* emissiveTriangles;
    const float* meshAliasQ;
    const uint32_t* meshAliasIdx;
    const float* triAliasQ;
    const uint32_t* triAliasIdx;
    const TextureView* textures;
    uint32_t textureCount;
    const float* srgbLut;  // 256 entries, computed once on the host (identical to the oracle's table)
    SpectralTables spectral;
    AccelView accel;
    EnvDistribution env;
};

// ---- textures: LOD 0, bilinear, Vulkan texel-centre convention, sRGB decode per texel (core/scene/textures.c:141-209) ----
__device__ __forceinline__ int wrapCoord(int i, int n, uint32_t mode) {
    if (mode == VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    if (mode == VKRT_TEXTURE_WRAP_MIRRORED_REPEAT) {
        int p = 2 * n;
        int m = i % p;
        if (m < 0) m += p;
        m -= n;
        int mir = m >= 0 ? m : -(1 + m);
        return (n - 1) - mir;
    }
    int m = i % n;
    return m < 0 ? m + n : m;
}

__device__ __forceinline__ float4 fetchTexel(const SceneView& sc, const TextureView& t, int x, int y) {
    size_t idx = (size_t)y * t.width + x;
    switch (t.format) {
        case VKRT_TEXTURE_FORMAT_RGBA8_UNORM: {
            uchar4 p = reinterpret_cast<const uchar4*>(t.pixels)[idx];
            if (t.colorSpace == VKRT_TEXTURE_COLOR_SPACE_SRGB)
                return float4(sc.srgbLut[p.x], sc.srgbLut[p.y], sc.srgbLut[p.z], float(p.w) * (1.0f / 255.0f));
            return float4(float(p.x), float(p.y), float(p.z), float(p.w)) * (1.0f / 255.0f);
        }
        case VKRT_TEXTURE_FORMAT_RGBA16_UNORM: {
            ushort4 p = reinterpret_cast<const ushort4*>(t.pixels)[idx];
            return float4(float(p.x), float(p.y), float(p.z), float(p.w)) * (1.0f / 65535.0f);
        }
        case VKRT_TEXTURE_FORMAT_RGBA16_SFLOAT: {
            ushort4 p = reinterpret_cast<const ushort4*>(t.pixels)[idx];
            return float4(f16_to_f32(p.x), f16_to_f32(p.y), f16_to_f32(p.z), f16_to_f32(p.w));
        }
        default: {
            ::float4 p = reinterpret_cast<const ::float4*>(t.pixels)[idx];
            return float4(p.x, p.y, p.z, p.w);
        }
    }
}

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) { return a + (b - a) * t; }

static __device__ __noinline__ float4 sampleTextureBilinear(const SceneView& sc, uint32_t textureIndex, float2 uv, uint32_t wrapU, uint32_t wrapV) {
    if (textureIndex >= sc.textureCount) return float4(1.0f);
    const TextureView t = sc.textures[textureIndex];
    if (t.width == 0) return float4(1.0f);
    float fx = uv.x * float(t.width) - 0.5f;
    float fy = uv.y * float(t.height) - 0.5f;
    float flx = floorf(fx), fly = floorf(fy);
    float ax = fx - flx, ay = fy - fly;
    int x0 = wrapCoord((int)flx, (int)t.width, wrapU), x1 = wrapCoord((int)flx + 1, (int)t.width, wrapU);
    int y0 = wrapCoord((int)fly, (int)t.height, wrapV), y1 = wrapCoord((int)fly + 1, (int)t.height, wrapV);
    float4 t00 = fetchTexel(sc, t, x0, y0), t10 = fetchTexel(sc, t, x1, y0);
    float4 t01 = fetchTexel(sc, t, x0, y1), t11 = fetchTexel(sc, t, x1, y1);
    return lerp4(lerp4(t00, t10, ax), lerp4(t01, t11, ax), ay);
}

__device__ __forceinline__ uint32_t wrapModeOrDefault(uint32_t m) {
    return (m == VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE || m == VKRT_TEXTURE_WRAP_MIRRORED_REPEAT) ? m : VKRT_TEXTURE_WRAP_REPEAT;
}

struct SurfaceTextureData {
    float4 color = float4(1.0f);
    float2 texcoord0, texcoord1;
};

// material/textures.slang:37-73
__device__ __forceinline__ float2 transformTextureUv(float2 uv, const float* tr, float rotation) {
    float2 scaled = uv * float2(tr[0], tr[1]);
    float s = sinf(rotation), co = cosf(rotation);
    return float2(co * scaled.x - s * scaled.y, s * scaled.x + co * scaled.y) + float2(tr[2], tr[3]);
}
static __device__ __noinline__ float4 sampleMaterialTextureBound(const SceneView& sc, uint32_t textureIndex, uint32_t packedWrap, ::float4 transform,
                                                                 float rotation, float2 uv) {
    const float tr[4] = {transform.x, transform.y, transform.z, transform.w};
    return sampleTextureBilinear(sc, textureIndex, transformTextureUv(uv, tr, rotation), wrapModeOrDefault(packedWrap & 0xffffu),
                                 wrapModeOrDefault((packedWrap >> 16) & 0xffffu));
}
__device__ __forceinline__ float4 sampleMaterialTexture(const SceneView& sc, uint32_t textureIndex, uint32_t packedWrap, const float* transform,
                                                        float rotation, uint32_t texcoordSet, const SurfaceTextureData& s, float4 fallback) {
    if (textureIndex == VKRT_INVALID_INDEX) return fallback;  // the common case: untextured slot, nothing else is executed
    float2 uv = texcoordSet == 1u ? s.texcoord1 : s.texcoord0;
    return sampleMaterialTextureBound(sc, textureIndex, packedWrap, make_float4(transform[0], transform[1], transform[2], transform[3]), rotation, uv);
}
__device__ __forceinline__ uint32_t materialTexcoordSet(const Material& m, uint32_t slot) { return (m.textureTexcoordSets >> (slot * 8u)) & 0xffu; }
__device__ __forceinline__ float4 sampleBaseColorTexture(const SceneView& sc, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(sc, m.baseColorTextureIndex, m.baseColorTextureWrap, m.baseColorTextureTransform, m.textureRotations[0],
                                 materialTexcoordSet(m, 0), s, float4(1.0f));
}
__device__ __forceinline__ float4 sampleMetallicRoughnessTexture(const SceneView& sc, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(sc, m.metallicRoughnessTextureIndex, m.metallicRoughnessTextureWrap, m.metallicRoughnessTextureTransform,
                                 m.textureRotations[1], materialTexcoordSet(m, 1), s, float4(1.0f));
}
__device__ __forceinline__ float4 sampleNormalTexture(const SceneView& sc, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(sc, m.normalTextureIndex, m.normalTextureWrap, m.normalTextureTransform, m.textureRotations[2],
                                 materialTexcoordSet(m, 2), s, float4(0.5f, 0.5f, 1.0f, 1.0f));
}
__device__ __forceinline__ float4 sampleEmissiveTexture(const SceneView& sc, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(sc, m.emissiveTextureIndex, m.emissiveTextureWrap, m.emissiveTextureTransform, m.textureRotations[3],
                                 materialTexcoordSet(m, 3), s, float4(1.0f));
}

// ---- geometry fetch (geometry/surface/interpolation.slang:8-60) ------------------------------------------------------
struct TriangleVertices {
    ShaderVertex v0, v1, v2;
};
__device__ __forceinline__ ShaderVertex loadShaderVertex(const ShaderVertex* p) {
    // 48-byte vertex = three 16-byte loads
    const ::float4* q = reinterpret_cast<const ::float4*>(p);
    ::float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    ShaderVertex v;
    v.position[0] = a.x; v.position[1] = a.y; v.position[2] = a.z; v.position[3] = a.w;
    v.texcoord0[0] = b.x; v.texcoord0[1] = b.y; v.texcoord1[0] = b.z; v.texcoord1[1] = b.w;
    v.packedNormal = __float_as_uint(c.x); v.packedTangent = __float_as_uint(c.y); v.packedColor = __float_as_uint(c.z);
    return v;
}
__device__ __forceinline__ void loadTriangleVertices(const SceneView& sc, const MeshInfo& mesh, uint32_t prim, TriangleVertices& tv) {
    uint32_t tb = mesh.indexBase + prim * 3u;
    tv.v0 = loadShaderVertex(sc.vertices + (__ldg(sc.indices + tb + 0u) + mesh.vertexBase));
    tv.v1 = loadShaderVertex(sc.vertices + (__ldg(sc.indices + tb + 1u) + mesh.vertexBase));
    tv.v2 = loadShaderVertex(sc.vertices + (__ldg(sc.indices + tb + 2u) + mesh.vertexBase));
}
__device__ __forceinline__ float2 interp2(float2 a, float2 b, float2 c, float2 bary) {
    float w = 1.0f - bary.x - bary.y;
    return a * w + b * bary.x + c * bary.y;
}
__device__ __forceinline__ float3 interp3(float3 a, float3 b, float3 c, float2 bary) {
    float w = 1.0f - bary.x - bary.y;
    return a * w + b * bary.x + c * bary.y;
}
__device__ __forceinline__ float4 interp4(float4 a, float4 b, float4 c, float2 bary) {
    float w = 1.0f - bary.x - bary.y;
    return a * w + b * bary.x + c * bary.y;
}
__device__ __forceinline__ SurfaceTextureData evaluateSurfaceTextureData(const TriangleVertices& tv, float2 bary) {
    SurfaceTextureData s;
    s.color = interp4(unpackColorRGBA8(tv.v0.packedColor), unpackColorRGBA8(tv.v1.packedColor), unpackColorRGBA8(tv.v2.packedColor), bary);
    s.texcoord0 = interp2(float2(tv.v0.texcoord0[0], tv.v0.texcoord0[1]), float2(tv.v1.texcoord0[0], tv.v1.texcoord0[1]),
                          float2(tv.v2.texcoord0[0], tv.v2.texcoord0[1]), bary);
    s.texcoord1 = interp2(float2(tv.v0.texcoord1[0], tv.v0.texcoord1[1]), float2(tv.v1.texcoord1[0], tv.v1.texcoord1[1]),
                          float2(tv.v2.texcoord1[0], tv.v2.texcoord1[1]), bary);
    return s;
}

// ---- stochastic alpha (rt/alpha_test.slang:8-75). The candidate's random number is a pure function of
// (rng at ray start, instance, primitive) so that the accepted hit does not depend on traversal order. ----
__device__ __forceinline__ float alphaCandidateRand(uint32_t raySeed, uint32_t inst, uint32_t prim) {
    uint32_t h = hash(raySeed ^ hash(inst * 0x9e3779b1u + prim + 0x7f4a7c15u));
    return float(h & 0x00ffffffu) * (1.0f / 16777216.0f);
}
static __device__ __noinline__ bool alphaHitAccepted(const SceneView& sc, uint32_t inst, uint32_t prim, float2 bary, uint32_t raySeed) {
    const MeshInfo mesh = sc.meshInfos[inst];
    const Material& m = sc.materials[mesh.materialIndex];
    const bool usesMask = m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_MASK;
    const bool usesBlend = m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_BLEND || m.opacity < 0.999f || mesh.opacity < 0.999f;
    if (!(usesMask || usesBlend)) return true;
    TriangleVertices tv;
    loadTriangleVertices(sc, mesh, prim, tv);
    SurfaceTextureData s = evaluateSurfaceTextureData(tv, bary);
    float textureAlpha = m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_OPAQUE ? s.color.w : sampleBaseColorTexture(sc, m, s).w * s.color.w;
    if (usesMask) {
        if (textureAlpha < m.alphaCutoff) return false;
        float maskOpacity = saturate(mesh.opacity * m.opacity * s.color.w);
        if (maskOpacity >= 1.0f) return true;
        if (maskOpacity <= 0.0f) return false;
        return alphaCandidateRand(raySeed, inst, prim) <= maskOpacity;
    }
    float opacity = saturate(mesh.opacity * m.opacity * textureAlpha);
    if (opacity <= 0.0f) return false;
    if (opacity >= 1.0f) return true;
    return alphaCandidateRand(raySeed, inst, prim) <= opacity;
}

} // namespace vk
