// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of vkrt's path-tracing hot path: raygen megakernel loop, surface reconstruction,
// NEE + MIS, layered principled BSDF (RGB / single-wavelength / hero-wavelength), film write-back.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// The product (vkrt_b200/) never links, includes or calls it.
//
// PARITY STATUS: PINNED against the reference's own shader sources. `make -C oracle ref` compiles /root/reference/src/shaders/**/*.slang
// (transliterated to C++ by oracle/ref_slang/) into oracle/_ref/libvkrt_refshade.so, and tests/test_reference_pin.py requires this
// restatement to reproduce it BIT FOR BIT: known-answer functions, 20 000 randomised closure evaluations / samplings per render mode
// with every lobe switched on, and whole frames (accumulation, denoiser features, display image) in RGB / single / hero. Not pinnable,
// because upstream it is the Vulkan driver's and the repository holds no source for it: BVH build and traversal order, the ray /
// triangle test, the texture filter (specified in accel.h and below; both sides of every comparison use these stand-ins).
//
// Structure follows the reference one function at a time; each block cites file:line under /root/reference/src/shaders.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include <atomic>
#include <thread>

#include "accel.h"
#include "shading.h"
#include "../include/vkrt_closure.h"

namespace orc {

// Dynamic-scheduled parallel loop over [begin, end) in chunks (std::thread; OpenMP is not reliably available here).
template <class F>
static void parallelFor(int64_t begin, int64_t end, int64_t chunk, int threads, F&& body) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 1 || end - begin <= chunk) {
        for (int64_t i = begin; i < end; i++) body(i, 0);
        return;
    }
    std::atomic<int64_t> next(begin);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t]() {
            for (;;) {
                int64_t b = next.fetch_add(chunk);
                if (b >= end) break;
                int64_t e = b + chunk < end ? b + chunk : end;
                for (int64_t i = b; i < e; i++) body(i, t);
            }
        });
    }
    for (auto& th : pool) th.join();
}

// ------------------------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------------------------
struct Texture {
    uint32_t width = 0, height = 0, format = 0, colorSpace = 0;
    std::vector<uint8_t> pixels;
};

struct Blas {
    uint32_t vertexBase = 0, indexBase = 0, indexCount = 0;
    std::vector<float3> tri; // 3 per triangle, object space, in primitive order
    Bvh2 bvh;
};

struct Instance {
    float world[12]; // row-major 3x4
    float inv[12];
    uint32_t blas = 0;
    uint8_t alphaTested = 0;
};

struct Ctx {
    std::vector<ShaderVertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<MeshInfo> meshInfos;
    std::vector<Instance> instances;
    std::vector<uint32_t> geometrySource;
    std::vector<Material> materials;
    std::vector<EmissiveMesh> emissiveMeshes;
    std::vector<EmissiveTriangle> emissiveTriangles;
    std::vector<float> meshAliasQ, triAliasQ;
    std::vector<uint32_t> meshAliasIdx, triAliasIdx;
    std::vector<Texture> textures;
    std::vector<float> rgb2spec;
    SpectralTables spectral;
    float srgbLut[256];

    std::vector<Blas> blas;
    Bvh2 tlas;
    bool accelBuilt = false;
    bool bruteForce = false;
    int threads = 0;

    uint32_t width = 0, height = 0;
    std::vector<float> accum[2];      // RGBA32F
    std::vector<uint16_t> albedo[2];  // RGBA16F
    std::vector<uint16_t> normal[2];  // RGBA16F
    std::vector<uint16_t> output;     // RGBA16 UNORM
    int readIndex = 0;                // accumulationReadImage index (frame.c:386-388 swaps per traced frame)
    SceneData lastScene;
    bool haveScene = false;
    uint64_t rayCount = 0, shadowRayCount = 0;
    std::string error;
};

// Affine inverse in fp64 (adjugate), rounded once to fp32. Mirrored by vkrt_b200/csrc/api.cu:invertAffine3x4.
static void invertAffine3x4(const float m[12], float out[12]) {
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double tx = m[3], ty = m[7], tz = m[11];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * B + c * C;
    double id = det != 0.0 ? 1.0 / det : 0.0;
    double r00 = A * id, r01 = -(b * i - c * h) * id, r02 = (b * f - c * e) * id;
    double r10 = B * id, r11 = (a * i - c * g) * id, r12 = -(a * f - c * d) * id;
    double r20 = C * id, r21 = -(a * h - b * g) * id, r22 = (a * e - b * d) * id;
    out[0] = (float)r00; out[1] = (float)r01; out[2] = (float)r02;
    out[4] = (float)r10; out[5] = (float)r11; out[6] = (float)r12;
    out[8] = (float)r20; out[9] = (float)r21; out[10] = (float)r22;
    out[3] = (float)(-(r00 * tx + r01 * ty + r02 * tz));
    out[7] = (float)(-(r10 * tx + r11 * ty + r12 * tz));
    out[11] = (float)(-(r20 * tx + r21 * ty + r22 * tz));
}

static inline float3 xformPoint(const float m[12], float3 p) {
    return float3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
                  ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
static inline float3 xformVector(const float m[12], float3 v) {
    return float3((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z,
                  (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}

static Aabb paddedBox(Aabb b) {
    float m = 0.0f;
    for (int k = 0; k < 3; k++) m = std::max(m, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
    float eps = m * 9.5367431640625e-07f; // 2^-20 relative
    b.lo = b.lo - float3(eps);
    b.hi = b.hi + float3(eps);
    return b;
}

static bool buildAccel(Ctx& c) {
    c.blas.clear();
    const uint32_t n = (uint32_t)c.meshInfos.size();
    if (c.instances.size() != n) return false;
    std::map<std::tuple<uint32_t, uint32_t, uint32_t>, uint32_t> byRange;
    std::map<uint32_t, uint32_t> bySource;
    for (uint32_t i = 0; i < n; i++) {
        const MeshInfo& mi = c.meshInfos[i];
        uint32_t blasIndex;
        bool found = false;
        if (!c.geometrySource.empty()) {
            auto it = bySource.find(c.geometrySource[i]);
            if (it != bySource.end()) { blasIndex = it->second; found = true; }
        } else {
            auto it = byRange.find(std::make_tuple(mi.vertexBase, mi.indexBase, mi.indexCount));
            if (it != byRange.end()) { blasIndex = it->second; found = true; }
        }
        if (!found) {
            blasIndex = (uint32_t)c.blas.size();
            c.blas.push_back(Blas());
            Blas& b = c.blas.back();
            b.vertexBase = mi.vertexBase;
            b.indexBase = mi.indexBase;
            b.indexCount = mi.indexCount;
            uint32_t triCount = mi.indexCount / 3u;
            b.tri.resize((size_t)triCount * 3);
            std::vector<Aabb> boxes(triCount);
            for (uint32_t t = 0; t < triCount; t++) {
                Aabb box;
                for (int k = 0; k < 3; k++) {
                    uint32_t idx = c.indices[mi.indexBase + t * 3u + k] + mi.vertexBase;
                    if (idx >= c.vertices.size()) return false;
                    const float* p = c.vertices[idx].position;
                    float3 v(p[0], p[1], p[2]);
                    b.tri[(size_t)t * 3 + k] = v;
                    box.grow(v);
                }
                boxes[t] = paddedBox(box);
            }
            b.bvh.build(boxes, 4);
            if (!c.geometrySource.empty()) bySource[c.geometrySource[i]] = blasIndex;
            else byRange[std::make_tuple(mi.vertexBase, mi.indexBase, mi.indexCount)] = blasIndex;
        }
        c.instances[i].blas = blasIndex;
    }
    std::vector<Aabb> ibox(n);
    for (uint32_t i = 0; i < n; i++) {
        const Blas& b = c.blas[c.instances[i].blas];
        Aabb w;
        if (!b.bvh.nodes.empty() && !b.tri.empty()) {
            Aabb r = b.bvh.nodes[0].box;
            for (int k = 0; k < 8; k++) {
                float3 p((k & 1) ? r.hi.x : r.lo.x, (k & 2) ? r.hi.y : r.lo.y, (k & 4) ? r.hi.z : r.lo.z);
                w.grow(xformPoint(c.instances[i].world, p));
            }
            w = paddedBox(w);
        } else {
            w.lo = w.hi = float3(0.0f);
        }
        ibox[i] = w;
    }
    c.tlas.build(ibox, 1);
    c.accelBuilt = true;
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// Textures (material/textures.slang:8-73; sampler setup core/scene/textures.c:141-209): LOD 0, bilinear, 3 wrap modes.
// The reference uses the driver's fixed-function sampler (not pinnable upstream); specified here as fp32 bilinear with
// the Vulkan texel-centre convention, sRGB decode per texel before filtering (8-bit LUT).
// ------------------------------------------------------------------------------------------------------------------
static int wrapCoord(int i, int n, uint32_t mode) {
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

static float4 fetchTexel(const Ctx& c, const Texture& t, int x, int y) {
    size_t idx = (size_t)y * t.width + x;
    switch (t.format) {
        case VKRT_TEXTURE_FORMAT_RGBA8_UNORM: {
            const uint8_t* p = &t.pixels[idx * 4];
            if (t.colorSpace == VKRT_TEXTURE_COLOR_SPACE_SRGB)
                return float4(c.srgbLut[p[0]], c.srgbLut[p[1]], c.srgbLut[p[2]], float(p[3]) * (1.0f / 255.0f));
            return float4(float(p[0]), float(p[1]), float(p[2]), float(p[3])) * (1.0f / 255.0f);
        }
        case VKRT_TEXTURE_FORMAT_RGBA16_UNORM: {
            const uint16_t* p = reinterpret_cast<const uint16_t*>(&t.pixels[idx * 8]);
            return float4(float(p[0]), float(p[1]), float(p[2]), float(p[3])) * (1.0f / 65535.0f);
        }
        case VKRT_TEXTURE_FORMAT_RGBA16_SFLOAT: {
            const uint16_t* p = reinterpret_cast<const uint16_t*>(&t.pixels[idx * 8]);
            return float4(f16_to_f32(p[0]), f16_to_f32(p[1]), f16_to_f32(p[2]), f16_to_f32(p[3]));
        }
        default: {
            const float* p = reinterpret_cast<const float*>(&t.pixels[idx * 16]);
            return float4(p[0], p[1], p[2], p[3]);
        }
    }
}

static float4 lerp4(float4 a, float4 b, float t) { return a + (b - a) * t; }

static float4 sampleTextureBilinear(const Ctx& c, uint32_t textureIndex, float2 uv, uint32_t wrapU, uint32_t wrapV) {
    if (textureIndex >= c.textures.size() || c.textures[textureIndex].width == 0) return float4(1.0f); // fallback 1x1 white (textures.c:182-201)
    const Texture& t = c.textures[textureIndex];
    float fx = uv.x * float(t.width) - 0.5f;
    float fy = uv.y * float(t.height) - 0.5f;
    float flx = std::floor(fx), fly = std::floor(fy);
    float ax = fx - flx, ay = fy - fly;
    int x0 = wrapCoord((int)flx, (int)t.width, wrapU), x1 = wrapCoord((int)flx + 1, (int)t.width, wrapU);
    int y0 = wrapCoord((int)fly, (int)t.height, wrapV), y1 = wrapCoord((int)fly + 1, (int)t.height, wrapV);
    float4 t00 = fetchTexel(c, t, x0, y0), t10 = fetchTexel(c, t, x1, y0);
    float4 t01 = fetchTexel(c, t, x0, y1), t11 = fetchTexel(c, t, x1, y1);
    return lerp4(lerp4(t00, t10, ax), lerp4(t01, t11, ax), ay);
}

static uint32_t wrapModeOrDefault(uint32_t m) {
    return (m == VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE || m == VKRT_TEXTURE_WRAP_MIRRORED_REPEAT) ? m : VKRT_TEXTURE_WRAP_REPEAT;
}

struct SurfaceTextureData {
    float4 color = float4(1.0f);
    float2 texcoord0, texcoord1;
};

// textures.slang:37-73
static float2 transformTextureUv(float2 uv, const float* tr, float rotation) {
    float2 scaled = uv * float2(tr[0], tr[1]);
    float s = std::sin(rotation), co = std::cos(rotation);
    return float2(co * scaled.x - s * scaled.y, s * scaled.x + co * scaled.y) + float2(tr[2], tr[3]);
}
static float4 sampleMaterialTexture(const Ctx& c, uint32_t textureIndex, uint32_t packedWrap, const float* transform, float rotation,
                                    uint32_t texcoordSet, const SurfaceTextureData& s, float4 fallback) {
    if (textureIndex == VKRT_INVALID_INDEX) return fallback;
    float2 uv = texcoordSet == 1u ? s.texcoord1 : s.texcoord0;
    return sampleTextureBilinear(c, textureIndex, transformTextureUv(uv, transform, rotation), wrapModeOrDefault(packedWrap & 0xffffu),
                                 wrapModeOrDefault((packedWrap >> 16) & 0xffffu));
}
static uint32_t texcoordSet(const Material& m, uint32_t slot) { return (m.textureTexcoordSets >> (slot * 8u)) & 0xffu; }
static float4 sampleBaseColorTexture(const Ctx& c, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(c, m.baseColorTextureIndex, m.baseColorTextureWrap, m.baseColorTextureTransform, m.textureRotations[0],
                                 texcoordSet(m, 0), s, float4(1.0f));
}
static float4 sampleMetallicRoughnessTexture(const Ctx& c, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(c, m.metallicRoughnessTextureIndex, m.metallicRoughnessTextureWrap, m.metallicRoughnessTextureTransform,
                                 m.textureRotations[1], texcoordSet(m, 1), s, float4(1.0f));
}
static float4 sampleNormalTexture(const Ctx& c, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(c, m.normalTextureIndex, m.normalTextureWrap, m.normalTextureTransform, m.textureRotations[2],
                                 texcoordSet(m, 2), s, float4(0.5f, 0.5f, 1.0f, 1.0f));
}
static float4 sampleEmissiveTexture(const Ctx& c, const Material& m, const SurfaceTextureData& s) {
    return sampleMaterialTexture(c, m.emissiveTextureIndex, m.emissiveTextureWrap, m.emissiveTextureTransform, m.textureRotations[3],
                                 texcoordSet(m, 3), s, float4(1.0f));
}
// textures.slang:123-155
static void applySurfaceTextures(const Ctx& c, Material& m, const SurfaceTextureData& s) {
    float4 bc = sampleBaseColorTexture(c, m, s);
    m.baseColor[0] *= bc.x * s.color.x;
    m.baseColor[1] *= bc.y * s.color.y;
    m.baseColor[2] *= bc.z * s.color.z;
    float4 mr = sampleMetallicRoughnessTexture(c, m, s);
    m.roughness = saturate(m.roughness * mr.y);
    m.metallic = saturate(m.metallic * mr.z);
    float4 et = sampleEmissiveTexture(c, m, s);
    float3 emission = float3(m.emissionColor[0], m.emissionColor[1], m.emissionColor[2]) * m.emissionLuminance * et.xyz();
    float emissionMax = maxComponent(emission);
    if (emissionMax > 0.0f) {
        float3 ec = emission / emissionMax;
        m.emissionColor[0] = ec.x; m.emissionColor[1] = ec.y; m.emissionColor[2] = ec.z;
        m.emissionLuminance = emissionMax;
    } else {
        m.emissionColor[0] = m.emissionColor[1] = m.emissionColor[2] = 1.0f;
        m.emissionLuminance = 0.0f;
    }
}
static float3 applyNormalTexture(const Ctx& c, const Material& m, const SurfaceTextureData& s, const ShadingBasis& basis) {
    if (m.normalTextureIndex == VKRT_INVALID_INDEX) return basis.normal;
    float3 ns = sampleNormalTexture(c, m, s).xyz() * 2.0f - 1.0f;
    ns.x *= m.normalTextureScale;
    ns.y *= m.normalTextureScale;
    ns = safeNormalize(ns);
    return safeNormalize(basis.tangent * ns.x + basis.bitangent * ns.y + basis.normal * ns.z);
}

// ------------------------------------------------------------------------------------------------------------------
// Geometry fetch (geometry/surface/interpolation.slang:8-60)
// ------------------------------------------------------------------------------------------------------------------
static void loadTriangleVertices(const Ctx& c, const MeshInfo& mesh, uint32_t prim, const ShaderVertex*& v0, const ShaderVertex*& v1,
                                 const ShaderVertex*& v2) {
    uint32_t tb = mesh.indexBase + prim * 3u;
    v0 = &c.vertices[c.indices[tb + 0u] + mesh.vertexBase];
    v1 = &c.vertices[c.indices[tb + 1u] + mesh.vertexBase];
    v2 = &c.vertices[c.indices[tb + 2u] + mesh.vertexBase];
}
static float2 interp2(float2 a, float2 b, float2 cc, float2 bary) {
    float w = 1.0f - bary.x - bary.y;
    return a * w + b * bary.x + cc * bary.y;
}
static float3 interp3(float3 a, float3 b, float3 cc, float2 bary) {
    float w = 1.0f - bary.x - bary.y;
    return a * w + b * bary.x + cc * bary.y;
}
static float4 interp4(float4 a, float4 b, float4 cc, float2 bary) {
    float w = 1.0f - bary.x - bary.y;
    return a * w + b * bary.x + cc * bary.y;
}
static SurfaceTextureData evaluateSurfaceTextureData(const Ctx& c, const MeshInfo& mesh, uint32_t prim, float2 bary) {
    const ShaderVertex *v0, *v1, *v2;
    loadTriangleVertices(c, mesh, prim, v0, v1, v2);
    SurfaceTextureData s;
    s.color = interp4(unpackColorRGBA8(v0->packedColor), unpackColorRGBA8(v1->packedColor), unpackColorRGBA8(v2->packedColor), bary);
    s.texcoord0 = interp2(float2(v0->texcoord0[0], v0->texcoord0[1]), float2(v1->texcoord0[0], v1->texcoord0[1]),
                          float2(v2->texcoord0[0], v2->texcoord0[1]), bary);
    s.texcoord1 = interp2(float2(v0->texcoord1[0], v0->texcoord1[1]), float2(v1->texcoord1[0], v1->texcoord1[1]),
                          float2(v2->texcoord1[0], v2->texcoord1[1]), bary);
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
// Stochastic alpha (rt/alpha_test.slang:8-75; host decision core/render/accel/tlas.c:291-322).
// DEVIATION (documented in DESIGN.md): the reference's any-hit shader consumes the path's rng in driver traversal
// order, which no software BVH can reproduce. Here the random number of a candidate is a pure function of
// (rng at ray start, instance, primitive), and the path rng is not advanced, so the accepted hit is independent of
// traversal order. Statistically equivalent; exact for opaque scenes.
// ------------------------------------------------------------------------------------------------------------------
static float alphaCandidateRand(uint32_t raySeed, uint32_t inst, uint32_t prim) {
    uint32_t h = hash(raySeed ^ hash(inst * 0x9e3779b1u + prim + 0x7f4a7c15u));
    return float(h & 0x00ffffffu) * (1.0f / 16777216.0f);
}
static bool materialUsesAlphaMask(const Material& m) { return m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_MASK; }
static bool materialUsesAlphaBlend(const Material& m, float meshOpacity) {
    return m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_BLEND || m.opacity < 0.999f || meshOpacity < 0.999f;
}
static float evaluateMaterialTextureAlpha(const Ctx& c, const Material& m, const SurfaceTextureData& s) {
    if (m.alphaMode == VKRT_MATERIAL_ALPHA_MODE_OPAQUE) return s.color.w;
    return sampleBaseColorTexture(c, m, s).w * s.color.w;
}
static bool alphaHitAccepted(const Ctx& c, uint32_t inst, uint32_t prim, float2 bary, uint32_t raySeed) {
    const MeshInfo& mesh = c.meshInfos[inst];
    const Material& m = c.materials[mesh.materialIndex];
    if (!(materialUsesAlphaMask(m) || materialUsesAlphaBlend(m, mesh.opacity))) return true;
    SurfaceTextureData s = evaluateSurfaceTextureData(c, mesh, prim, bary);
    if (materialUsesAlphaMask(m)) {
        if (evaluateMaterialTextureAlpha(c, m, s) < m.alphaCutoff) return false;
        float maskOpacity = saturate(mesh.opacity * m.opacity * s.color.w);
        if (maskOpacity >= 1.0f) return true;
        if (maskOpacity <= 0.0f) return false;
        return alphaCandidateRand(raySeed, inst, prim) <= maskOpacity;
    }
    float opacity = saturate(mesh.opacity * m.opacity * evaluateMaterialTextureAlpha(c, m, s));
    if (!materialUsesAlphaBlend(m, mesh.opacity)) return true;
    if (opacity <= 0.0f) return false;
    if (opacity >= 1.0f) return true;
    return alphaCandidateRand(raySeed, inst, prim) <= opacity;
}

// ------------------------------------------------------------------------------------------------------------------
// Ray queries (rt/queries/scene_query.slang:10-42, shadow_query.slang:8-28)
// ------------------------------------------------------------------------------------------------------------------
struct TraceCounters {
    uint64_t rays = 0, shadowRays = 0;
};

// Shadow rays (anyHit): the reference terminates on the FIRST accepted hit in driver traversal order and reports
// "unsupported transmission" if that occluder's material has transmission > 0 (entry/shadow/closest_hit.slang:5-7), which
// is order dependent when a segment crosses both glass and an opaque occluder. Pinned here (and in the CUDA kernel) as:
// an accepted hit on a non-transmissive instance ends the search (OCCLUDED); accepted hits on transmissive instances only
// raise sawTransmissive (-> UNSUPPORTED_TRANSMISSION if nothing opaque is found). Order independent.
// Optional per-thread stand-in for the driver's any-hit stage (oracle/ref_slang/refshade_entry.cpp runs the reference's own any-hit
// shaders through it); nullptr = the oracle's hashed stochastic alpha test above. `transmissiveInstance` reports which transmissive
// instance a shadow ray met (the reference's shadow closest-hit shader wants an instance index).
struct AnyHitHook {
    bool (*accept)(void* user, uint32_t inst, uint32_t prim, float t, float u, float v) = nullptr;
    void* user = nullptr;
    uint32_t transmissiveInstance = 0xFFFFFFFFu;
};
static thread_local AnyHitHook* g_anyHitHook = nullptr;

static void intersectInstance(const Ctx& c, uint32_t inst, const Ray& ray, uint32_t raySeed, bool anyHit, HitRecord& best, float& tBest,
                              bool& done, bool& sawTransmissive) {
    const Instance& I = c.instances[inst];
    const Blas& b = c.blas[I.blas];
    if (b.tri.empty()) return;
    float3 o = xformPoint(I.inv, ray.origin);
    float3 d = xformVector(I.inv, ray.direction);
    RayShear sh = makeRayShear(d);
    if (!sh.valid) return;
    const bool instanceTransmissive = anyHit && c.materials[c.meshInfos[inst].materialIndex].transmission > 0.0f;
    if (instanceTransmissive && sawTransmissive) return; // nothing new to learn from this instance
    bool skipInstance = false;
    auto testTri = [&](uint32_t prim) {
        float t, u, v;
        if (!watertightTriangle(o, sh, b.tri[(size_t)prim * 3], b.tri[(size_t)prim * 3 + 1], b.tri[(size_t)prim * 3 + 2], t, u, v)) return;
        if (!(t > ray.tMin)) return;
        if (!hitCloser(t, inst, prim, best, tBest)) return;
        if (I.alphaTested) {
            const bool accepted = g_anyHitHook && g_anyHitHook->accept ? g_anyHitHook->accept(g_anyHitHook->user, inst, prim, t, u, v)
                                                                       : alphaHitAccepted(c, inst, prim, float2(u, v), raySeed);
            if (!accepted) return;
        }
        if (anyHit) {
            if (instanceTransmissive) {
                sawTransmissive = true;
                skipInstance = true;
                if (g_anyHitHook) g_anyHitHook->transmissiveInstance = inst;
                return;
            }
            best.instance = inst;
            best.primitive = prim;
            done = true;
            return;
        }
        best.instance = inst;
        best.primitive = prim;
        best.t = t;
        best.u = u;
        best.v = v;
        tBest = t;
    };
    if (c.bruteForce) {
        uint32_t n = (uint32_t)(b.tri.size() / 3);
        for (uint32_t p = 0; p < n && !done && !skipInstance; p++) testTri(p);
        return;
    }
    float3 invD = safeInvDir(d);
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0 && !done && !skipInstance) {
        const Bvh2Node& n = b.bvh.nodes[stack[--sp]];
        if (!slabTest(n.box, o, invD, ray.tMin, tBest)) continue;
        if (n.count) {
            for (uint32_t k = 0; k < n.count && !done && !skipInstance; k++) testTri(b.bvh.prims[n.left + k]);
        } else {
            if (sp + 2 > 64) { continue; }
            stack[sp++] = n.left;
            stack[sp++] = n.left + 1;
        }
    }
}

static HitRecord traceRay(const Ctx& c, const Ray& ray, uint32_t raySeed, bool anyHit, bool* outSawTransmissive = nullptr) {
    HitRecord best;
    float tBest = ray.tMax;
    bool done = false;
    bool sawT = false;
    bool& sawTransmissive = outSawTransmissive ? *outSawTransmissive : sawT;
    if (c.instances.empty()) return best;
    if (c.bruteForce) {
        for (uint32_t i = 0; i < (uint32_t)c.instances.size() && !done; i++) intersectInstance(c, i, ray, raySeed, anyHit, best, tBest, done, sawTransmissive);
        return best;
    }
    float3 invD = safeInvDir(ray.direction);
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0 && !done) {
        const Bvh2Node& n = c.tlas.nodes[stack[--sp]];
        if (!slabTest(n.box, ray.origin, invD, ray.tMin, tBest)) continue;
        if (n.count) {
            for (uint32_t k = 0; k < n.count && !done; k++) intersectInstance(c, c.tlas.prims[n.left + k], ray, raySeed, anyHit, best, tBest, done, sawTransmissive);
        } else {
            if (sp + 2 > 64) continue;
            stack[sp++] = n.left;
            stack[sp++] = n.left + 1;
        }
    }
    return best;
}

static const uint32_t SHADOW_OCCLUDED = 0u, SHADOW_VISIBLE = 1u, SHADOW_UNSUPPORTED_TRANSMISSION = 2u;

// entry/shadow/{closest_hit,miss}.slang
static uint32_t traceShadowRay(const Ctx& c, float3 origin, float3 direction, float tMax, uint32_t raySeed) {
    Ray r;
    r.origin = origin;
    r.direction = direction;
    r.tMin = RAY_T_MIN;
    r.tMax = tMax;
    bool sawTransmissive = false;
    HitRecord h = traceRay(c, r, raySeed, true, &sawTransmissive);
    if (h.hit()) return SHADOW_OCCLUDED;
    return sawTransmissive ? SHADOW_UNSUPPORTED_TRANSMISSION : SHADOW_VISIBLE;
}

// ------------------------------------------------------------------------------------------------------------------
// Surface reconstruction (geometry/surface/reconstruct.slang:8-49, integrator/path/surface_state.slang:8-31)
// ------------------------------------------------------------------------------------------------------------------
struct SurfaceShadingData {
    uint32_t materialIndex = VKRT_INVALID_INDEX;
    uint32_t frontFace = 0u;
    float3 shadingNormal = float3(0.0f, 0.0f, 1.0f);
    float3 geometricNormal = float3(0.0f, 0.0f, 1.0f);
    float4 tangent = float4(1.0f, 0.0f, 0.0f, 1.0f);
    SurfaceTextureData textureData;
};

static SurfaceShadingData reconstructSurfaceShading(const Ctx& c, const MeshInfo& mesh, uint32_t prim, float2 bary, float3 worldRayDir) {
    const ShaderVertex *v0, *v1, *v2;
    loadTriangleVertices(c, mesh, prim, v0, v1, v2);
    float3 objectNormal = safeNormalize(
        interp3(unpackOctNormal(v0->packedNormal), unpackOctNormal(v1->packedNormal), unpackOctNormal(v2->packedNormal), bary));
    float4 objectTangent =
        interp4(unpackOctTangent(v0->packedTangent), unpackOctTangent(v1->packedTangent), unpackOctTangent(v2->packedTangent), bary);
    float handedness = objectTangent.w < 0.0f ? -1.0f : 1.0f;
    float3 shadingNormalUnoriented = meshTransformNormal(mesh, objectNormal);
    float facing = dot(shadingNormalUnoriented, worldRayDir) > 0.0f ? -1.0f : 1.0f;
    float3 p0(v0->position[0], v0->position[1], v0->position[2]);
    float3 p1(v1->position[0], v1->position[1], v1->position[2]);
    float3 p2(v2->position[0], v2->position[1], v2->position[2]);
    float3 worldEdge1 = meshTransformVector(mesh, p1 - p0);
    float3 worldEdge2 = meshTransformVector(mesh, p2 - p0);
    float3 geometricNormal = safeNormalize(cross(worldEdge1, worldEdge2));
    if (surfaceTransformSign(mesh) < 0.0f) geometricNormal = -geometricNormal;
    SurfaceShadingData hit;
    hit.materialIndex = mesh.materialIndex;
    hit.frontFace = dot(geometricNormal, worldRayDir) < 0.0f ? 1u : 0u;
    hit.shadingNormal = shadingNormalUnoriented * facing;
    hit.geometricNormal = hit.frontFace != 0u ? geometricNormal : -geometricNormal;
    hit.tangent = float4(safeNormalize(meshTransformVector(mesh, objectTangent.xyz())) * facing, handedness);
    hit.textureData = evaluateSurfaceTextureData(c, mesh, prim, bary);
    return hit;
}

struct PathSurfaceState {
    float3 hitPoint;
    SurfaceShadingData surface;
    Material material;
    ShadingBasis basis;
};

static PathSurfaceState makePathSurfaceState(const Ctx& c, const HitRecord& hit, const Ray& ray) {
    PathSurfaceState s;
    s.hitPoint = ray.origin + ray.direction * hit.t;
    s.surface = reconstructSurfaceShading(c, c.meshInfos[hit.instance], hit.primitive, float2(hit.u, hit.v), ray.direction);
    s.material = c.materials[s.surface.materialIndex];
    ShadingBasis unperturbed = makeShadingBasis(s.surface.shadingNormal, s.surface.tangent);
    s.surface.shadingNormal = applyNormalTexture(c, s.material, s.surface.textureData, unperturbed);
    s.surface.shadingNormal = sanitizeShadingNormal(s.surface.shadingNormal, s.surface.geometricNormal, -ray.direction);
    s.basis = makeShadingBasis(s.surface.shadingNormal, s.surface.tangent);
    applySurfaceTextures(c, s.material, s.surface.textureData);
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
// Environment (light/environment.slang:9-27)
// ------------------------------------------------------------------------------------------------------------------
static float3 sampleEnvironmentRadiance(const Ctx& c, const SceneData& scene, float3 worldDir) {
    if (scene.environmentTextureIndex == VKRT_INVALID_INDEX)
        return float3(scene.environmentLight[0], scene.environmentLight[1], scene.environmentLight[2]);
    float3 dir = normalize(worldDir);
    float phi = std::atan2(dir.y, dir.x) + scene.environmentRotation * (PI / 180.0f);
    float theta = std::acos(clamp(dir.z, -1.0f, 1.0f));
    float2 uv(frac(phi * (0.5f * INV_PI) + 0.5f), theta * INV_PI);
    float3 radiance = sampleTextureBilinear(c, scene.environmentTextureIndex, uv, VKRT_TEXTURE_WRAP_REPEAT, VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE).xyz();
    return radiance * scene.environmentLight[3];
}

// ------------------------------------------------------------------------------------------------------------------
// Direct lighting (light/direct/light_sampling.slang:9-63, common.slang:24-53)
// ------------------------------------------------------------------------------------------------------------------
struct LightSample {
    float3 position, normal, emission;
    float pdf = 0.0f;
    bool valid = false;
};
static LightSample sampleDirectLightPoint(const Ctx& c, const SceneData& scene, uint& rng) {
    LightSample ls;
    uint meshCount = scene.emissiveMeshCount;
    if (meshCount == 0u) return ls;
    uint meshIdx = sampleAlias(rand(rng), meshCount, 0u, c.meshAliasQ.data(), c.meshAliasIdx.data());
    const EmissiveMesh& em = c.emissiveMeshes[meshIdx];
    if (em.triCount == 0u) return ls;
    uint localTri = sampleAlias(rand(rng), em.triCount, em.triOffset, c.triAliasQ.data(), c.triAliasIdx.data());
    const EmissiveTriangle& tri = c.emissiveTriangles[em.triOffset + localTri];
    float u1 = rand(rng);
    float u2 = rand(rng);
    float sqrtU1 = sqrt(u1);
    float b1 = u2 * sqrtU1;
    float b2 = (1.0f - u2) * sqrtU1;
    float3 v0(tri.v0Area[0], tri.v0Area[1], tri.v0Area[2]);
    float3 e1(tri.e1Pad[0], tri.e1Pad[1], tri.e1Pad[2]);
    float3 e2(tri.e2Pad[0], tri.e2Pad[1], tri.e2Pad[2]);
    ls.position = v0 + b1 * e1 + b2 * e2;
    ls.normal = safeNormalize(cross(e1, e2));
    ls.emission = float3(em.emission[0], em.emission[1], em.emission[2]);
    ls.pdf = em.pmfMesh * em.invTotalArea;
    ls.valid = ls.pdf > 0.0f;
    return ls;
}
struct DirectLightSurfaceSample {
    LightSample light;
    float3 wi;
    float shadowDistance = 0.0f;
    float pdfSolidAngle = 0.0f;
    bool valid = false;
};
static DirectLightSurfaceSample sampleDirectLightSurface(const Ctx& c, const SceneData& scene, float3 hitPoint, uint& rng) {
    DirectLightSurfaceSample s;
    s.light = sampleDirectLightPoint(c, scene, rng);
    if (!s.light.valid) return s;
    float3 toLight = s.light.position - hitPoint;
    float d2 = dot(toLight, toLight);
    if (d2 <= 0.0f) return s;
    float invD = rsqrt(d2);
    float distance = d2 * invD;
    s.wi = toLight * invD;
    float cosLight = abs(dot(s.wi, s.light.normal));
    if (cosLight <= 0.0f) return s;
    s.pdfSolidAngle = s.light.pdf * d2 / cosLight;
    if (s.pdfSolidAngle <= 0.0f) return s;
    s.shadowDistance = distance - SHADOW_DISTANCE_OFFSET;
    if (s.shadowDistance <= 0.0f) return s;
    s.valid = true;
    return s;
}
struct DirectLightSample {
    DirectLightSurfaceSample surface;
    float3 wiLocal;
    bool valid = false, neeUnsupported = false;
};
static DirectLightSample sampleDirectLight(Ctx& c, const SceneData& scene, float3 hitPoint, float3 geometricNormal, const ShadingBasis& basis,
                                           const BSDFState& state, uint& rng, TraceCounters& tc) {
    DirectLightSample light;
    if (cosTheta(state.wo) <= 0.0f) return light;
    light.surface = sampleDirectLightSurface(c, scene, hitPoint, rng);
    if (!light.surface.valid) return light;
    float3 shadowOffset = dot(light.surface.wi, geometricNormal) >= 0.0f ? geometricNormal : -geometricNormal;
    float3 shadowOrigin = hitPoint + shadowOffset * SHADOW_ORIGIN_OFFSET;
    tc.shadowRays++;
    uint32_t vis = traceShadowRay(c, shadowOrigin, light.surface.wi, light.surface.shadowDistance, rng);
    if (vis == SHADOW_UNSUPPORTED_TRANSMISSION) {
        light.neeUnsupported = true;
        return light;
    }
    if (vis != SHADOW_VISIBLE) return light;
    light.wiLocal = worldToLocal(light.surface.wi, basis);
    if (materialMediumIsRefractive(state.material) && cosTheta(light.wiLocal) <= 0.0f) return light;
    light.valid = true;
    return light;
}

// light/direct/mis_weights.slang:6-42
static float lightPdfAreaToSolidAngle(float lightPdfArea, float3 geometricNormal, float3 rayDirection, float hitDistance) {
    if (lightPdfArea <= 0.0f) return 0.0f;
    float d2 = hitDistance * hitDistance;
    if (d2 <= 0.0f) return 0.0f;
    float cosLight = abs(dot(rayDirection, geometricNormal));
    if (cosLight <= 0.0f) return 0.0f;
    return lightPdfArea * d2 / cosLight;
}
static float computeBSDFEmitterMISWeight(float bsdfPdf, float lightPdfArea, float3 gn, float3 rd, float hd) {
    float lp = lightPdfAreaToSolidAngle(lightPdfArea, gn, rd, hd);
    return lp <= 0.0f ? 1.0f : powerHeuristic(bsdfPdf, lp);
}
static float computeSpectralMISWeight(float4 sampledTechniquePdf, float4 alternateTechniquePdf) {
    float denom = dot(sampledTechniquePdf, float4(1.0f)) + dot(alternateTechniquePdf, float4(1.0f));
    return denom > 0.0f ? sampledTechniquePdf.x / denom : 0.0f;
}
static float computeSpectralEmitterMISWeight(float4 prevVertexTechniquePathPdf, float4 prevBsdfTechniquePdf, float lightPdfSolidAngle) {
    float4 bsdfTp = prevVertexTechniquePathPdf * prevBsdfTechniquePdf;
    float4 lightTp = prevVertexTechniquePathPdf * lightPdfSolidAngle;
    return computeSpectralMISWeight(bsdfTp, lightTp);
}

// ------------------------------------------------------------------------------------------------------------------
// Integrator state (integrator/path/state.slang:28-204)
// ------------------------------------------------------------------------------------------------------------------
static const uint MODE_BSDF_ONLY = 1u << 0, MODE_NEE_ONLY = 1u << 1, MODE_BOUNCE_COUNT = 1u << 2, MODE_DN_ALBEDO = 1u << 3,
                  MODE_DN_NORMAL = 1u << 4, MODE_DN_VALIDITY = 1u << 5, MODE_DN_DEPTH = 1u << 6, MODE_DN_FOLLOW = 1u << 7,
                  MODE_NEE_ENABLED = 1u << 8;

struct ModeState {
    uint debugMode = 0, flags = 0;
    bool has(uint f) const { return (flags & f) != 0u; }
};
static ModeState makeModeState(const SceneData& scene) {
    ModeState m;
    m.debugMode = scene.debugMode;
    if (m.debugMode == VKRT_DEBUG_MODE_BSDF_ONLY) m.flags |= MODE_BSDF_ONLY;
    if (m.debugMode == VKRT_DEBUG_MODE_NEE_ONLY) m.flags |= MODE_NEE_ONLY;
    if (m.debugMode == VKRT_DEBUG_MODE_BOUNCE_COUNT) m.flags |= MODE_BOUNCE_COUNT;
    if (m.debugMode == VKRT_DEBUG_MODE_DENOISER_ALBEDO) m.flags |= MODE_DN_ALBEDO;
    if (m.debugMode == VKRT_DEBUG_MODE_DENOISER_NORMAL) m.flags |= MODE_DN_NORMAL;
    if (m.debugMode == VKRT_DEBUG_MODE_DENOISER_FEATURE_VALIDITY) m.flags |= MODE_DN_VALIDITY;
    if (m.debugMode == VKRT_DEBUG_MODE_DENOISER_FEATURE_DEPTH) m.flags |= MODE_DN_DEPTH;
    if (m.debugMode == VKRT_DEBUG_MODE_DENOISER_FOLLOW_SPECULAR) m.flags |= MODE_DN_FOLLOW;
    if (scene.misNeeEnabled != 0u && scene.emissiveMeshCount > 0u) m.flags |= MODE_NEE_ENABLED;
    return m;
}

struct DenoiserFeatures {
    float3 albedo = float3(0.0f), normal = float3(0.0f);
    float weight = 0.0f, depth = 0.0f, followSpecular = 0.0f;
    bool resolved = false;
};
struct FrameState {
    float3 radiance = float3(0.0f);
    DenoiserFeatures features;
    bool debugEarlyOut = false;
};
struct PathCommon {
    uint rng = 0u;
    MediumState medium;
    bool prevVertexNeeAllowed = false;
    float prevBsdfPdf = 0.0f;
    uint bounceCount = 0u;
    Ray ray;
};

static void initCommon(PathCommon& p, const SceneData& scene, int px, int py, uint previousSamples, uint sampleIndex) {
    p.rng = initPixelSeed(px, py, scene.frameNumber, previousSamples + sampleIndex);
    p.medium = MediumState();
    float jx = rand(p.rng);
    float jy = rand(p.rng);
    p.ray = makePrimaryRay(scene, px, py, float2(jx, jy) - float2(0.5f));
}

// integrator/path/debug.slang:31-64, utility/debug.slang
static bool isTextureMapDebugMode(uint m) {
    return m == VKRT_DEBUG_MODE_BASE_COLOR_MAP || m == VKRT_DEBUG_MODE_METALLIC_MAP || m == VKRT_DEBUG_MODE_ROUGHNESS_MAP ||
           m == VKRT_DEBUG_MODE_NORMAL_MAP || m == VKRT_DEBUG_MODE_EMISSIVE_MAP;
}
static bool handlePrimarySurfaceDebug(const Ctx& c, const ModeState& mode, uint sampleIndex, uint depth, const PathSurfaceState& s, float hitDistance,
                                      FrameState& frame) {
    if (sampleIndex != 0u || depth != 0u) return false;
    if (mode.debugMode == VKRT_DEBUG_MODE_NORMALS) {
        frame.radiance = s.surface.shadingNormal * 0.5f + 0.5f;
        frame.debugEarlyOut = true;
        return true;
    }
    if (mode.debugMode == VKRT_DEBUG_MODE_DEPTH) {
        frame.radiance = float3(1.0f / (1.0f + hitDistance));
        frame.debugEarlyOut = true;
        return true;
    }
    if (isTextureMapDebugMode(mode.debugMode)) {
        // NOTE: the reference evaluates these on surfaceState.material AFTER applySurfaceTextures (loop.slang:33-45).
        const Material& m = s.material;
        const SurfaceTextureData& td = s.surface.textureData;
        float3 r(0.0f);
        switch (mode.debugMode) {
            case VKRT_DEBUG_MODE_BASE_COLOR_MAP: r = sampleBaseColorTexture(c, m, td).xyz(); break;
            case VKRT_DEBUG_MODE_METALLIC_MAP: r = float3(sampleMetallicRoughnessTexture(c, m, td).z); break;
            case VKRT_DEBUG_MODE_ROUGHNESS_MAP: r = float3(sampleMetallicRoughnessTexture(c, m, td).y); break;
            case VKRT_DEBUG_MODE_NORMAL_MAP: r = sampleNormalTexture(c, m, td).xyz(); break;
            case VKRT_DEBUG_MODE_EMISSIVE_MAP: r = sampleEmissiveTexture(c, m, td).xyz(); break;
        }
        frame.radiance = r;
        frame.debugEarlyOut = true;
        return true;
    }
    return false;
}

// integrator/path/loop.slang:83-111
static void resolveDenoiserFeatures(DenoiserFeatures& f, const BSDFMaterial& m, const PathSurfaceState& s, uint depth) {
    bool follow = materialDenoiserShouldFollowSpecularHit(m, s.surface.frontFace);
    if (depth == 0u) f.followSpecular = follow ? 1.0f : 0.0f;
    if (f.resolved || follow) return;
    f.albedo = materialDenoiserAlbedo(m);
    f.normal = s.surface.shadingNormal;
    f.weight = 1.0f;
    f.depth = float(depth + 1u);
    f.resolved = true;
}
static void accumulateDenoiserFeatures(DenoiserFeatures& t, const DenoiserFeatures& s) {
    t.albedo += s.albedo;
    t.normal += s.normal;
    t.weight += s.weight;
    t.depth += s.depth;
    t.followSpecular += s.followSpecular;
}
static void advancePathRay(PathCommon& p, const PathSurfaceState& s, uint isTransmission, float3 wi) {
    float3 off = isTransmission != 0u ? -s.surface.geometricNormal : s.surface.geometricNormal;
    p.ray.origin = s.hitPoint + off * SHADOW_ORIGIN_OFFSET;
    p.ray.direction = wi;
    p.ray.tMin = RAY_T_MIN;
    p.ray.tMax = RAY_T_MAX;
}
static void applyBounceCountDebug(const SceneData& scene, const ModeState& mode, uint sampleIndex, uint bounceCount, FrameState& frame) {
    if (!mode.has(MODE_BOUNCE_COUNT) || sampleIndex != 0u) return;
    float t = clamp(float(bounceCount) / max(float(scene.rrMaxDepth), 1.0f), 0.0f, 1.0f);
    frame.radiance = float3(t, 1.0f - t, 0.0f);
    frame.debugEarlyOut = true;
}
static float bsdfEmitterMisWeight(const Ctx& c, const ModeState& mode, const PathCommon& p, const HitRecord& hit, const PathSurfaceState& s, uint depth) {
    if (!mode.has(MODE_NEE_ENABLED) || !p.prevVertexNeeAllowed || depth == 0u || p.prevBsdfPdf <= 0.0f) return 1.0f;
    return computeBSDFEmitterMISWeight(p.prevBsdfPdf, c.meshInfos[hit.instance].lightPdfArea, s.surface.geometricNormal, p.ray.direction, hit.t);
}

// ------------------------------------------------------------------------------------------------------------------
// RGB transport (integrator/path/rgb/integrator.slang:15-113, rgb/transport.slang:6-71, light/direct/rgb_lighting.slang:7-31)
// ------------------------------------------------------------------------------------------------------------------
static void traceRgbPathSample(Ctx& c, const SceneData& scene, const ModeState& mode, int px, int py, uint previousSamples, uint sampleIndex,
                               FrameState& frame, uint* firstHit, TraceCounters& tc) {
    float3 sampleRadiance(0.0f);
    DenoiserFeatures features;
    PathCommon path;
    initCommon(path, scene, px, py, previousSamples, sampleIndex);
    float3 throughput(1.0f);
    const SpectralTables& T = c.spectral;

    for (uint depth = 0u; depth < scene.rrMaxDepth; depth++) {
        tc.rays++;
        HitRecord hit = traceRay(c, path.ray, path.rng, false);
        if (sampleIndex == 0u && depth == 0u && firstHit) { firstHit[0] = hit.instance; firstHit[1] = hit.primitive; }
        if (!hit.hit()) {
            if (!mode.has(MODE_NEE_ONLY) && !mediumHasActiveBoundary(path.medium))
                sampleRadiance += throughput * sampleEnvironmentRadiance(c, scene, path.ray.direction);
            break;
        }
        if (path.medium.absorptionActive()) throughput *= mediumTransmittance(path.medium, hit.t);
        if (!anyGreater(throughput, 0.0f)) break;

        PathSurfaceState s = makePathSurfaceState(c, hit, path.ray);
        BSDFMaterial bm(s.material);
        if (handlePrimarySurfaceDebug(c, mode, sampleIndex, depth, s, hit.t, frame)) break;
        resolveDenoiserFeatures(features, bm, s, depth);

        BSDFState state(bm, worldToLocal(-path.ray.direction, s.basis), s.surface.frontFace, 0.0f, 0u);
        bool currentVertexNeeAllowed = !path.medium.refractiveActive();
        bool vertexNeeSupported = currentVertexNeeAllowed;

        float3 contribution(0.0f);
        if (!mode.has(MODE_BSDF_ONLY) && mode.has(MODE_NEE_ENABLED) && currentVertexNeeAllowed) {
            DirectLightSample light = sampleDirectLight(c, scene, s.hitPoint, s.surface.geometricNormal, s.basis, state, path.rng, tc);
            if (light.neeUnsupported) {
                vertexNeeSupported = false;
            } else if (light.valid) {
                BSDFEval e = evalBSDF(T, state, light.wiLocal);
                if (e.pdf > 0.0f) {
                    float misWeight = powerHeuristic(light.surface.pdfSolidAngle, e.pdf);
                    float3 transmittance = mediumTransmittance(path.medium, light.surface.shadowDistance);
                    float3 fCos = e.value * absCosTheta(light.wiLocal);
                    contribution += misWeight * transmittance * fCos * light.surface.light.emission / light.surface.pdfSolidAngle;
                }
            }
        }
        if (!mode.has(MODE_NEE_ONLY)) {
            float3 emission = float3(s.material.emissionColor[0], s.material.emissionColor[1], s.material.emissionColor[2]) * s.material.emissionLuminance;
            if (anyGreater(emission, 0.0f)) contribution += emission * bsdfEmitterMisWeight(c, mode, path, hit, s, depth);
        }
        sampleRadiance += throughput * contribution;
        path.bounceCount = depth + 1u;

        BSDFSample smp = sampleBSDF(T, state, s.basis, path.rng);
        if (!smp.isUsable()) break;
        throughput *= smp.weight;
        if (!anyGreater(throughput, 0.0f)) break;
        path.prevBsdfPdf = smp.pdf;

        bool sampledPathNeeAllowed = vertexNeeSupported && smp.isTransmission == 0u;
        updateMediumStateFromTransmission(T, bm, s.surface.frontFace, smp.isTransmission, 0.0f, 0u, path.medium);
        path.prevVertexNeeAllowed = sampledPathNeeAllowed;

        if (depth + 1u >= scene.rrMinDepth) {
            float cp = clamp(maxComponent(throughput), RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB);
            if (rand(path.rng) > cp) break;
            throughput /= cp;
        }
        advancePathRay(path, s, smp.isTransmission, smp.wi);
    }
    applyBounceCountDebug(scene, mode, sampleIndex, path.bounceCount, frame);
    if (frame.debugEarlyOut) return;
    frame.radiance += sampleRadiance;
    accumulateDenoiserFeatures(frame.features, features);
}

// ------------------------------------------------------------------------------------------------------------------
// Spectral single (integrator/path/spectral_single/*.slang, light/direct/spectral_lighting.slang:7-32)
// ------------------------------------------------------------------------------------------------------------------
static float sampleDirectLightSpectralSingle(Ctx& c, const SceneData& scene, const PathSurfaceState& s, const BSDFState& state, const MediumState& medium,
                                             uint& rng, bool& neeSupported, TraceCounters& tc) {
    neeSupported = true;
    DirectLightSample light = sampleDirectLight(c, scene, s.hitPoint, s.surface.geometricNormal, s.basis, state, rng, tc);
    if (light.neeUnsupported) { neeSupported = false; return 0.0f; }
    if (!light.valid) return 0.0f;
    BSDFEval e = evalSingleWavelengthBSDF(c.spectral, state, light.wiLocal);
    if (e.pdf <= 0.0f) return 0.0f;
    float spectralValue = e.value.x * absCosTheta(light.wiLocal) * spectralScalarFromLinearSrgb(c.spectral, light.surface.light.emission, state.wavelengthNm);
    float misWeight = powerHeuristic(light.surface.pdfSolidAngle, e.pdf);
    return misWeight * mediumTransmittance(medium, light.surface.shadowDistance).x * spectralValue / light.surface.pdfSolidAngle;
}

static void traceSpectralSinglePathSample(Ctx& c, const SceneData& scene, const ModeState& mode, int px, int py, uint previousSamples, uint sampleIndex,
                                          FrameState& frame, uint* firstHit, TraceCounters& tc) {
    float spectralRadiance = 0.0f;
    DenoiserFeatures features;
    PathCommon path;
    initCommon(path, scene, px, py, previousSamples, sampleIndex);
    WavelengthSample wavelength = sampleUniformWavelength(path.rng, previousSamples + sampleIndex);
    float throughput = 1.0f;
    const SpectralTables& T = c.spectral;

    for (uint depth = 0u; depth < scene.rrMaxDepth; depth++) {
        tc.rays++;
        HitRecord hit = traceRay(c, path.ray, path.rng, false);
        if (sampleIndex == 0u && depth == 0u && firstHit) { firstHit[0] = hit.instance; firstHit[1] = hit.primitive; }
        if (!hit.hit()) {
            if (!mode.has(MODE_NEE_ONLY) && !mediumHasActiveBoundary(path.medium))
                spectralRadiance += throughput * spectralScalarFromLinearSrgb(T, sampleEnvironmentRadiance(c, scene, path.ray.direction), wavelength.lambdaNm);
            break;
        }
        if (path.medium.absorptionActive()) throughput *= mediumTransmittance(path.medium, hit.t).x;
        if (!(throughput > 0.0f)) break;

        PathSurfaceState s = makePathSurfaceState(c, hit, path.ray);
        BSDFMaterial bm(s.material);
        if (handlePrimarySurfaceDebug(c, mode, sampleIndex, depth, s, hit.t, frame)) break;
        resolveDenoiserFeatures(features, bm, s, depth);

        BSDFState state(bm, worldToLocal(-path.ray.direction, s.basis), s.surface.frontFace, wavelength.lambdaNm, 1u);
        bool currentVertexNeeAllowed = !path.medium.refractiveActive();
        bool vertexNeeSupported = currentVertexNeeAllowed;
        float contribution = 0.0f;
        if (!mode.has(MODE_BSDF_ONLY) && mode.has(MODE_NEE_ENABLED) && currentVertexNeeAllowed) {
            bool sup;
            contribution += sampleDirectLightSpectralSingle(c, scene, s, state, path.medium, path.rng, sup, tc);
            vertexNeeSupported = sup;
        }
        if (!mode.has(MODE_NEE_ONLY)) {
            float3 emission = float3(s.material.emissionColor[0], s.material.emissionColor[1], s.material.emissionColor[2]) * s.material.emissionLuminance;
            if (anyGreater(emission, 0.0f))
                contribution += spectralScalarFromLinearSrgb(T, emission, wavelength.lambdaNm) * bsdfEmitterMisWeight(c, mode, path, hit, s, depth);
        }
        spectralRadiance += throughput * contribution;
        path.bounceCount = depth + 1u;

        BSDFSample smp = sampleBSDF(T, state, s.basis, path.rng);
        if (!smp.isUsable()) break;
        throughput *= smp.weight.x;
        if (throughput <= 0.0f) break;
        path.prevBsdfPdf = smp.pdf;

        bool sampledPathNeeAllowed = vertexNeeSupported && smp.isTransmission == 0u;
        updateMediumStateFromTransmission(T, bm, s.surface.frontFace, smp.isTransmission, wavelength.lambdaNm, 1u, path.medium);
        path.prevVertexNeeAllowed = sampledPathNeeAllowed;

        if (depth + 1u >= scene.rrMinDepth) {
            float cp = clamp(throughput, RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB);
            if (rand(path.rng) > cp) break;
            throughput /= cp;
        }
        advancePathRay(path, s, smp.isTransmission, smp.wi);
    }
    applyBounceCountDebug(scene, mode, sampleIndex, path.bounceCount, frame);
    if (frame.debugEarlyOut) return;
    frame.radiance += spectralSampleToXYZ(spectralRadiance, wavelength);
    accumulateDenoiserFeatures(frame.features, features);
}

// ------------------------------------------------------------------------------------------------------------------
// Spectral hero (integrator/path/spectral_hero/integrator.slang:16-172, transport.slang:6-173, spectral_lighting.slang:34-64)
// ------------------------------------------------------------------------------------------------------------------
static void traceSpectralHeroPathSample(Ctx& c, const SceneData& scene, const ModeState& mode, int px, int py, uint previousSamples, uint sampleIndex,
                                        FrameState& frame, uint* firstHit, TraceCounters& tc) {
    float3 radianceXYZ(0.0f);         // sampleState.radiance
    float spectralRadianceScalar = 0.0f;
    float4 spectralRadiance(0.0f);
    DenoiserFeatures features;
    PathCommon path;
    initCommon(path, scene, px, py, previousSamples, sampleIndex);
    WavelengthSample scalarWavelength;
    float4 wavelengthsNm = sampleHeroWavelengths4(path.rng, previousSamples + sampleIndex);
    float4 invWavelengthPdf(WAVELENGTH_RANGE_NM);
    float throughputScalar = 1.0f;
    float4 throughput(1.0f);
    float4 techniquePathPdf(1.0f), prevVertexTechniquePathPdf(0.0f), prevBsdfTechniquePdf(0.0f);
    bool heroActive = true;
    const SpectralTables& T = c.spectral;

    for (uint depth = 0u; depth < scene.rrMaxDepth; depth++) {
        tc.rays++;
        HitRecord hit = traceRay(c, path.ray, path.rng, false);
        if (sampleIndex == 0u && depth == 0u && firstHit) { firstHit[0] = hit.instance; firstHit[1] = hit.primitive; }
        if (!hit.hit()) {
            if (!mode.has(MODE_NEE_ONLY) && !mediumHasActiveBoundary(path.medium)) {
                float3 env = sampleEnvironmentRadiance(c, scene, path.ray.direction);
                if (heroActive)
                    spectralRadiance += throughput * heroWavelengthBalanceWeight(techniquePathPdf) * spectralScalarFromLinearSrgb4(T, env, wavelengthsNm);
                else
                    spectralRadianceScalar += throughputScalar * spectralScalarFromLinearSrgb(T, env, scalarWavelength.lambdaNm);
            }
            break;
        }
        // applySpectralHeroMediumTransmittance (transport.slang:26-41)
        if (path.medium.absorptionActive()) {
            if (heroActive) throughput *= mediumSpectralTransmittance(path.medium, hit.t);
            else throughputScalar *= mediumTransmittance(path.medium, hit.t).x;
        }
        if (heroActive ? !anyGreater(throughput, 0.0f) : !(throughputScalar > 0.0f)) break;

        PathSurfaceState s = makePathSurfaceState(c, hit, path.ray);
        BSDFMaterial bm(s.material);
        if (handlePrimarySurfaceDebug(c, mode, sampleIndex, depth, s, hit.t, frame)) break;
        resolveDenoiserFeatures(features, bm, s, depth);

        float stateWavelengthNm = heroActive ? wavelengthsNm.x : scalarWavelength.lambdaNm;
        BSDFState state(bm, worldToLocal(-path.ray.direction, s.basis), s.surface.frontFace, stateWavelengthNm, 1u);
        bool currentVertexNeeAllowed = !path.medium.refractiveActive();
        bool vertexNeeSupported = currentVertexNeeAllowed;

        float contributionScalar = 0.0f;
        float4 contribution(0.0f);
        if (!mode.has(MODE_BSDF_ONLY) && mode.has(MODE_NEE_ENABLED) && currentVertexNeeAllowed) {
            if (heroActive) {
                DirectLightSample light = sampleDirectLight(c, scene, s.hitPoint, s.surface.geometricNormal, s.basis, state, path.rng, tc);
                if (light.neeUnsupported) {
                    vertexNeeSupported = false;
                } else if (light.valid) {
                    float4 bsdfLocalTp(0.0f);
                    float4 fCos = evalSpectralBSDF(T, state, light.wiLocal, wavelengthsNm, bsdfLocalTp) * absCosTheta(light.wiLocal);
                    float4 lightTp = techniquePathPdf * light.surface.pdfSolidAngle;
                    float4 bsdfTp = techniquePathPdf * bsdfLocalTp;
                    float misWeight = computeSpectralMISWeight(lightTp, bsdfTp);
                    if (misWeight > 0.0f) {
                        contribution += misWeight * mediumSpectralTransmittance(path.medium, light.surface.shadowDistance) * fCos *
                                        spectralScalarFromLinearSrgb4(T, light.surface.light.emission, wavelengthsNm) / light.surface.pdfSolidAngle;
                    }
                }
            } else {
                bool sup;
                contributionScalar += sampleDirectLightSpectralSingle(c, scene, s, state, path.medium, path.rng, sup, tc);
                vertexNeeSupported = sup;
            }
        }
        if (!mode.has(MODE_NEE_ONLY)) {
            float3 emission = float3(s.material.emissionColor[0], s.material.emissionColor[1], s.material.emissionColor[2]) * s.material.emissionLuminance;
            if (anyGreater(emission, 0.0f)) {
                if (heroActive) {
                    float misWeight = heroWavelengthBalanceWeight(techniquePathPdf);
                    if (mode.has(MODE_NEE_ENABLED) && path.prevVertexNeeAllowed && depth > 0u) {
                        float lp = lightPdfAreaToSolidAngle(c.meshInfos[hit.instance].lightPdfArea, s.surface.geometricNormal, path.ray.direction, hit.t);
                        misWeight = computeSpectralEmitterMISWeight(prevVertexTechniquePathPdf, prevBsdfTechniquePdf, lp);
                    }
                    contribution += spectralScalarFromLinearSrgb4(T, emission, wavelengthsNm) * misWeight;
                } else {
                    contributionScalar += spectralScalarFromLinearSrgb(T, emission, scalarWavelength.lambdaNm) * bsdfEmitterMisWeight(c, mode, path, hit, s, depth);
                }
            }
        }
        if (heroActive) spectralRadiance += throughput * contribution;
        else spectralRadianceScalar += throughputScalar * contributionScalar;
        path.bounceCount = depth + 1u;

        // sampleSpectralHeroNextDirection (transport.slang:56-121)
        uint isTransmission = 0u;
        float3 wi(0.0f);
        if (heroActive) {
            SpectralBSDFSample smp = sampleSpectralBSDF(T, state, s.basis, wavelengthsNm, path.rng);
            if (!smp.isUsable()) break;
            prevVertexTechniquePathPdf = techniquePathPdf;
            prevBsdfTechniquePdf = smp.techniquePdf;
            techniquePathPdf *= smp.techniquePdf;
            throughput *= smp.weight;
            if (!anyGreater(throughput, 0.0f)) break;
            if (smp.isTransmission != 0u && materialMediumIsRefractive(bm) && bm.abbeNumber > 0.0f) {
                radianceXYZ += spectralSample4ToXYZ(spectralRadiance, wavelengthsNm, invWavelengthPdf);
                spectralRadiance = float4(0.0f);
                throughputScalar = throughput.x;
                throughput = float4(0.0f);
                scalarWavelength.lambdaNm = wavelengthsNm.x;
                scalarWavelength.invPdf = invWavelengthPdf.x;
                path.prevBsdfPdf = smp.techniquePdf.x;
                heroActive = false;
            }
            isTransmission = smp.isTransmission;
            wi = smp.wi;
        } else {
            BSDFSample smp = sampleBSDF(T, state, s.basis, path.rng);
            if (!smp.isUsable()) break;
            throughputScalar *= smp.weight.x;
            if (throughputScalar <= 0.0f) break;
            path.prevBsdfPdf = smp.pdf;
            isTransmission = smp.isTransmission;
            wi = smp.wi;
        }

        bool sampledPathNeeAllowed = vertexNeeSupported && isTransmission == 0u;
        if (heroActive) updateMediumStateFromTransmissionSpectral(T, bm, s.surface.frontFace, isTransmission, wavelengthsNm, path.medium);
        else updateMediumStateFromTransmission(T, bm, s.surface.frontFace, isTransmission, scalarWavelength.lambdaNm, 1u, path.medium);
        path.prevVertexNeeAllowed = sampledPathNeeAllowed;

        if (depth + 1u >= scene.rrMinDepth) {
            float cp = heroActive ? clamp(maxComponent4(throughput), RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB)
                                  : clamp(throughputScalar, RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB);
            if (rand(path.rng) > cp) break;
            if (heroActive) {
                techniquePathPdf *= float4(cp);
                throughput /= cp;
            } else {
                throughputScalar /= cp;
            }
        }
        advancePathRay(path, s, isTransmission, wi);
    }
    applyBounceCountDebug(scene, mode, sampleIndex, path.bounceCount, frame);
    if (frame.debugEarlyOut) return;
    float3 sampleXYZ = radianceXYZ;
    if (heroActive) sampleXYZ += spectralSample4ToXYZ(spectralRadiance, wavelengthsNm, invWavelengthPdf);
    else sampleXYZ += spectralSampleToXYZ(spectralRadianceScalar, scalarWavelength);
    frame.radiance += sampleXYZ;
    accumulateDenoiserFeatures(frame.features, features);
}

// ------------------------------------------------------------------------------------------------------------------
// Film (integrator/path/writeback.slang:9-123)
// ------------------------------------------------------------------------------------------------------------------
static float3 blendAccumulatedValue(float3 prev, float prevW, float3 cur, float curW) {
    float total = prevW + curW;
    if (total <= 0.0f) return float3(0.0f);
    return (prev * prevW + cur * curW) / total;
}
static void storeHalf4(uint16_t* dst, float4 v) {
    dst[0] = f32_to_f16(v.x); dst[1] = f32_to_f16(v.y); dst[2] = f32_to_f16(v.z); dst[3] = f32_to_f16(v.w);
}
static float4 loadHalf4(const uint16_t* src) { return float4(f16_to_f32(src[0]), f16_to_f32(src[1]), f16_to_f32(src[2]), f16_to_f32(src[3])); }
static uint16_t unorm16(float v) {
    v = saturate(v);
    return (uint16_t)std::lrintf(v * 65535.0f); // round-to-nearest-even, like Vulkan's UNORM conversion
}
static void storeUnorm16x4(uint16_t* dst, float4 v) {
    dst[0] = unorm16(v.x); dst[1] = unorm16(v.y); dst[2] = unorm16(v.z); dst[3] = unorm16(v.w);
}

static void renderPixel(Ctx& c, const SceneData& scene, const ModeState& mode, int px, int py, int writeIndex, TraceCounters& tc) {
    const size_t p = (size_t)py * c.width + px;
    float* accW = &c.accum[writeIndex][p * 4];
    uint16_t* albW = &c.albedo[writeIndex][p * 4];
    uint16_t* nrmW = &c.normal[writeIndex][p * 4];
    uint16_t* out = &c.output[p * 4];
    if (!insideViewport(scene, px, py)) { // raygen_rgb.slang:6-12
        accW[0] = accW[1] = accW[2] = accW[3] = 0.0f;
        for (int k = 0; k < 4; k++) albW[k] = nrmW[k] = out[k] = 0;
        return;
    }
    const float* accR = &c.accum[c.readIndex][p * 4];
    float4 previousAccumulation(accR[0], accR[1], accR[2], accR[3]);
    uint spp = max(scene.samplesPerPixel, 1u);
    uint previousSamples = (uint)(previousAccumulation.w + 0.5f);
    FrameState frame;
    const uint renderMode = VKRT_RENDER_SETTINGS_MODE(scene.packedRenderSettings);
    const uint spectralSampling = VKRT_RENDER_SETTINGS_SPECTRAL(scene.packedRenderSettings);
    uint spectralOutput = renderMode == VKRT_RENDER_MODE_SPECTRAL ? 1u : 0u;

    if (scene.debugMode == VKRT_DEBUG_MODE_SELECTION_MASK) { // selection is out of scope: mask == 0 everywhere
        frame.radiance = float3(0.05f);
        frame.debugEarlyOut = true;
    } else {
        for (uint s = 0u; s < spp && !frame.debugEarlyOut; s++) {
            if (renderMode != VKRT_RENDER_MODE_SPECTRAL) traceRgbPathSample(c, scene, mode, px, py, previousSamples, s, frame, nullptr, tc);
            else if (spectralSampling == VKRT_SPECTRAL_SAMPLING_MODE_HERO) traceSpectralHeroPathSample(c, scene, mode, px, py, previousSamples, s, frame, nullptr, tc);
            else traceSpectralSinglePathSample(c, scene, mode, px, py, previousSamples, s, frame, nullptr, tc);
        }
    }
    // finalizeFrameAccumulation
    if (!frame.debugEarlyOut) {
        frame.radiance /= float(spp);
        if (frame.features.weight > 0.0f) {
            frame.features.albedo /= frame.features.weight;
            frame.features.normal /= frame.features.weight;
        } else {
            frame.features.albedo = float3(0.0f);
            frame.features.normal = float3(0.0f);
        }
    }
    // applyFrameDebugOverrides
    if (!frame.debugEarlyOut) {
        if (mode.has(MODE_DN_ALBEDO)) {
            frame.radiance = frame.features.albedo;
            frame.debugEarlyOut = true;
        } else if (mode.has(MODE_DN_NORMAL)) {
            frame.radiance = frame.features.weight > 0.0f ? frame.features.normal * 0.5f + 0.5f : float3(0.0f);
            frame.debugEarlyOut = true;
        } else if (mode.has(MODE_DN_VALIDITY)) {
            frame.radiance = float3(saturate(frame.features.weight / float(spp)));
            frame.debugEarlyOut = true;
        } else if (mode.has(MODE_DN_DEPTH)) {
            float nd = 0.0f;
            if (frame.features.weight > 0.0f) {
                float avg = frame.features.depth / frame.features.weight;
                nd = saturate((avg - 1.0f) / max(float(scene.rrMaxDepth - 1u), 1.0f));
            }
            frame.radiance = float3(nd);
            frame.debugEarlyOut = true;
        } else if (mode.has(MODE_DN_FOLLOW)) {
            float fr = saturate(frame.features.followSpecular / float(spp));
            frame.radiance = lerp(float3(0.05f), float3(1.0f, 0.6f, 0.0f), fr);
            frame.debugEarlyOut = true;
        }
    }
    if (frame.debugEarlyOut) { // writeDebugFrameOutputs
        accW[0] = frame.radiance.x; accW[1] = frame.radiance.y; accW[2] = frame.radiance.z; accW[3] = 0.0f;
        for (int k = 0; k < 4; k++) albW[k] = nrmW[k] = 0;
        storeUnorm16x4(out, float4(encodeDisplayColor(frame.radiance), 1.0f));
        return;
    }
    // writeAccumulatedFrameOutputs
    float4 previousAlbedo = loadHalf4(&c.albedo[c.readIndex][p * 4]);
    float4 previousNormal = loadHalf4(&c.normal[c.readIndex][p * 4]);
    float previousWeight = float(previousSamples);
    float totalWeight = previousWeight + float(spp);
    float totalAlbedoWeight = previousAlbedo.w + frame.features.weight;
    float totalNormalWeight = previousNormal.w + frame.features.weight;
    float3 accumulated = blendAccumulatedValue(previousAccumulation.xyz(), previousWeight, frame.radiance, float(spp));
    float3 accAlbedo = blendAccumulatedValue(previousAlbedo.xyz(), previousAlbedo.w, frame.features.albedo, frame.features.weight);
    float3 accNormal = blendAccumulatedValue(previousNormal.xyz(), previousNormal.w, frame.features.normal, frame.features.weight);
    accW[0] = accumulated.x; accW[1] = accumulated.y; accW[2] = accumulated.z; accW[3] = totalWeight;
    storeHalf4(albW, float4(accAlbedo, totalAlbedoWeight));
    storeHalf4(nrmW, float4(accNormal, totalNormalWeight));
    float3 display = spectralOutput ? mapSceneColorToDisplay(scene, xyzToLinearSrgb(accumulated)) : mapSceneColorToDisplay(scene, accumulated);
    storeUnorm16x4(out, float4(display, 1.0f));
}

} // namespace orc

// ------------------------------------------------------------------------------------------------------------------
// C API — mirrors include/vkrt_cuda.h one-to-one so the tests can drive both back-ends with the same code.
// ------------------------------------------------------------------------------------------------------------------
using namespace orc;
extern "C" {

#define ORC_API __attribute__((visibility("default")))
typedef struct oracle_ctx oracle_ctx;
static Ctx* C(oracle_ctx* c) { return reinterpret_cast<Ctx*>(c); }

ORC_API int oracle_create(oracle_ctx** out) {
    if (!out) return -1;
    Ctx* c = new Ctx();
    for (int i = 0; i < 256; i++) {
        float v = float(i) / 255.0f;
        c->srgbLut[i] = v <= 0.04045f ? v / 12.92f : std::pow((v + 0.055f) / 1.055f, 2.4f);
    }
    *out = reinterpret_cast<oracle_ctx*>(c);
    return 0;
}
ORC_API void oracle_destroy(oracle_ctx* c) { delete C(c); }
ORC_API void oracle_set_brute_force(oracle_ctx* c, int enabled) { C(c)->bruteForce = enabled != 0; }
ORC_API void oracle_set_threads(oracle_ctx* c, int n) { C(c)->threads = n; }
ORC_API int oracle_max_threads(void) { return (int)std::thread::hardware_concurrency(); }
ORC_API const float* oracle_srgb_lut(oracle_ctx* c) { return C(c)->srgbLut; }

ORC_API int oracle_set_geometry(oracle_ctx* c, const ShaderVertex* v, uint32_t nV, const uint32_t* idx, uint32_t nI) {
    if (!c || (nV && !v) || (nI && !idx)) return -1;
    C(c)->vertices.assign(v, v + nV);
    C(c)->indices.assign(idx, idx + nI);
    C(c)->accelBuilt = false;
    return 0;
}
ORC_API int oracle_set_instances(oracle_ctx* c, const MeshInfo* infos, const float* world3x4, const uint32_t* geometrySource,
                                 const uint8_t* alphaTested, uint32_t n) {
    if (!c || (n && (!infos || !world3x4))) return -1;
    Ctx& x = *C(c);
    x.meshInfos.assign(infos, infos + n);
    x.instances.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        std::memcpy(x.instances[i].world, world3x4 + (size_t)i * 12, sizeof(float) * 12);
        invertAffine3x4(x.instances[i].world, x.instances[i].inv);
        x.instances[i].alphaTested = alphaTested ? alphaTested[i] : 0;
    }
    if (geometrySource) x.geometrySource.assign(geometrySource, geometrySource + n);
    else x.geometrySource.clear();
    x.accelBuilt = false;
    return 0;
}
ORC_API int oracle_set_materials(oracle_ctx* c, const Material* m, uint32_t n) {
    if (!c || (n && !m)) return -1;
    C(c)->materials.assign(m, m + n);
    return 0;
}
ORC_API int oracle_set_lights(oracle_ctx* c, const EmissiveMesh* meshes, uint32_t nM, const EmissiveTriangle* tris, uint32_t nT, const float* mq,
                              const uint32_t* mi, const float* tq, const uint32_t* ti) {
    if (!c) return -1;
    Ctx& x = *C(c);
    x.emissiveMeshes.assign(meshes, meshes + nM);
    x.emissiveTriangles.assign(tris, tris + nT);
    x.meshAliasQ.assign(mq, mq + nM);
    x.meshAliasIdx.assign(mi, mi + nM);
    x.triAliasQ.assign(tq, tq + nT);
    x.triAliasIdx.assign(ti, ti + nT);
    return 0;
}
struct oracle_texture {
    const void* pixels;
    uint32_t width, height, format, colorSpace;
};
ORC_API int oracle_set_textures(oracle_ctx* c, const oracle_texture* t, uint32_t n) {
    if (!c || (n && !t)) return -1;
    Ctx& x = *C(c);
    x.textures.resize(n);
    static const size_t bpp[4] = {4, 8, 8, 16};
    for (uint32_t i = 0; i < n; i++) {
        if (t[i].format >= 4) return -1;
        x.textures[i].width = t[i].width;
        x.textures[i].height = t[i].height;
        x.textures[i].format = t[i].format;
        x.textures[i].colorSpace = t[i].colorSpace;
        size_t bytes = (size_t)t[i].width * t[i].height * bpp[t[i].format];
        x.textures[i].pixels.assign((const uint8_t*)t[i].pixels, (const uint8_t*)t[i].pixels + bytes);
    }
    return 0;
}
ORC_API int oracle_set_rgb2spec(oracle_ctx* c, const float* payload, uint32_t floatCount, RGB2SpecTableInfo info) {
    if (!c || !payload) return -1;
    Ctx& x = *C(c);
    x.rgb2spec.assign(payload, payload + floatCount);
    x.spectral.info = info;
    x.spectral.table = x.rgb2spec.data();
    return 0;
}
ORC_API int oracle_build_accel(oracle_ctx* c) { return buildAccel(*C(c)) ? 0 : -2; }

ORC_API int oracle_resize(oracle_ctx* c, uint32_t w, uint32_t h) {
    Ctx& x = *C(c);
    x.width = w;
    x.height = h;
    size_t n = (size_t)w * h * 4;
    for (int k = 0; k < 2; k++) {
        x.accum[k].assign(n, 0.0f);
        x.albedo[k].assign(n, 0);
        x.normal[k].assign(n, 0);
    }
    x.output.assign(n, 0);
    x.readIndex = 0;
    return 0;
}
ORC_API int oracle_reset_accumulation(oracle_ctx* c) {
    Ctx& x = *C(c);
    for (int k = 0; k < 2; k++) {
        std::fill(x.accum[k].begin(), x.accum[k].end(), 0.0f);
        std::fill(x.albedo[k].begin(), x.albedo[k].end(), 0);
        std::fill(x.normal[k].begin(), x.normal[k].end(), 0);
    }
    return 0;
}

// Renders rows [rowBegin, rowEnd) only (bounded CPU-baseline samples); the other rows of the write image keep the
// previous accumulation so that a later full frame is still well-defined. rowEnd = 0 means the full image.
ORC_API int oracle_render_frame_rows(oracle_ctx* c, const SceneData* sd, uint32_t rowBegin, uint32_t rowEnd, uint64_t* outRays) {
    if (!c || !sd) return -1;
    Ctx& x = *C(c);
    if (!x.accelBuilt && !buildAccel(x)) return -2;
    if (x.width == 0 || x.height == 0) return -1;
    SceneData scene = *sd;
    const uint renderMode = VKRT_RENDER_SETTINGS_MODE(scene.packedRenderSettings);
    if (renderMode == VKRT_RENDER_MODE_SPECTRAL && !x.spectral.table) return -2;
    if (rowEnd == 0 || rowEnd > x.height) rowEnd = x.height;
    ModeState mode = makeModeState(scene);
    int writeIndex = 1 - x.readIndex;
    bool partial = rowBegin != 0 || rowEnd != x.height;
    if (partial) {
        x.accum[writeIndex] = x.accum[x.readIndex];
        x.albedo[writeIndex] = x.albedo[x.readIndex];
        x.normal[writeIndex] = x.normal[x.readIndex];
    }
    int nthreads = x.threads > 0 ? x.threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::vector<TraceCounters> counters((size_t)nthreads);
    parallelFor((int64_t)rowBegin, (int64_t)rowEnd, 1, nthreads, [&](int64_t py, int t) {
        for (int px = 0; px < (int)x.width; px++) renderPixel(x, scene, mode, px, (int)py, writeIndex, counters[(size_t)t]);
    });
    uint64_t rays = 0, shadow = 0;
    for (auto& tc : counters) { rays += tc.rays; shadow += tc.shadowRays; }
    x.readIndex = writeIndex; // frame.c:386-388
    x.lastScene = scene;
    x.haveScene = true;
    x.rayCount = rays;
    x.shadowRayCount = shadow;
    if (outRays) { outRays[0] = rays; outRays[1] = shadow; }
    return 0;
}
ORC_API int oracle_render_frame(oracle_ctx* c, const SceneData* sd) { return oracle_render_frame_rows(c, sd, 0, 0, nullptr); }

// AOV ids match vkrt_cuda_aov.
ORC_API int oracle_read_aov(oracle_ctx* c, int which, void* dst, size_t bytes) {
    Ctx& x = *C(c);
    size_t px = (size_t)x.width * x.height;
    switch (which) {
        case 0: if (bytes != px * 16) return -1; std::memcpy(dst, x.accum[x.readIndex].data(), bytes); return 0;
        case 1: if (bytes != px * 8) return -1; std::memcpy(dst, x.albedo[x.readIndex].data(), bytes); return 0;
        case 2: if (bytes != px * 8) return -1; std::memcpy(dst, x.normal[x.readIndex].data(), bytes); return 0;
        case 3: if (bytes != px * 8) return -1; std::memcpy(dst, x.output.data(), bytes); return 0;
        default: break;
    }
    if (which < 4 || which > 6 || !x.haveScene) return -1;
    if (!x.accelBuilt && !buildAccel(x)) return -2;
    if ((which == 6 && bytes != px * 12) || (which != 6 && bytes != px * 8)) return -1;
    const SceneData& scene = x.lastScene;
    parallelFor(0, (int64_t)x.height, 4, x.threads, [&](int64_t pyl, int) {
        const int py = (int)pyl;
        for (int pxl = 0; pxl < (int)x.width; pxl++) {
            size_t p = (size_t)py * x.width + pxl;
            Ray r;
            uint rng = 0u;
            if (which == 5) { // frame 0, sample 0, jittered (state.slang:140-146)
                rng = initPixelSeed(pxl, py, 0u, 0u);
                float jx = orc::rand(rng);
                float jy = orc::rand(rng);
                r = makePrimaryRay(scene, pxl, py, float2(jx, jy) - float2(0.5f));
            } else {
                r = makePrimaryRay(scene, pxl, py, float2(0.0f)); // captureSelectionHit, debug.slang:21-29
            }
            HitRecord h;
            if (insideViewport(scene, pxl, py)) h = traceRay(x, r, rng, false);
            if (which == 6) {
                float* o = (float*)dst + p * 3;
                o[0] = h.t; o[1] = h.u; o[2] = h.v;
            } else {
                uint32_t* o = (uint32_t*)dst + p * 2;
                o[0] = h.instance; o[1] = h.primitive;
            }
        }
    });
    return 0;
}
ORC_API int oracle_trace_primary(oracle_ctx* c, const SceneData* sd) {
    C(c)->lastScene = *sd;
    C(c)->haveScene = true;
    return 0;
}

// rays: n * 8 floats; hits: n * 5 words {instance, primitive, t, u, v}
ORC_API int oracle_trace_rays(oracle_ctx* c, const float* rays, uint32_t n, int anyHit, uint32_t* hits) {
    Ctx& x = *C(c);
    if (!x.accelBuilt && !buildAccel(x)) return -2;
    parallelFor(0, (int64_t)n, 256, x.threads, [&](int64_t i, int) {
        Ray r;
        const float* f = rays + i * 8;
        r.origin = float3(f[0], f[1], f[2]);
        r.tMin = f[3];
        r.direction = float3(f[4], f[5], f[6]);
        r.tMax = f[7];
        bool sawT = false;
        HitRecord h = traceRay(x, r, 0u, anyHit != 0, &sawT);
        uint32_t* o = hits + i * 5;
        if (anyHit) {
            o[0] = h.hit() ? 1u : (sawT ? 2u : 0u); o[1] = 0; o[2] = 0; o[3] = 0; o[4] = 0;
        } else {
            o[0] = h.instance; o[1] = h.primitive; o[2] = asuint(h.t); o[3] = asuint(h.u); o[4] = asuint(h.v);
        }
    });
    return 0;
}

// ---- unit-test hooks: expose the scalar restatements for KATs and BSDF property tests -----------------------------
ORC_API uint32_t oracle_hash(uint32_t v) { return hash(v); }
ORC_API uint32_t oracle_init_pixel_seed(int x, int y, uint32_t frame, uint32_t sample) { return initPixelSeed(x, y, frame, sample); }
ORC_API float oracle_rand(uint32_t* rng) { return orc::rand(*rng); }
ORC_API uint32_t oracle_reverse_bits(uint32_t v) { return reverseBits32(v); }
ORC_API void oracle_unpack_normal(uint32_t packed, float* out3) { float3 n = unpackOctNormal(packed); out3[0] = n.x; out3[1] = n.y; out3[2] = n.z; }
ORC_API void oracle_unpack_tangent(uint32_t packed, float* out4) { float4 t = unpackOctTangent(packed); out4[0] = t.x; out4[1] = t.y; out4[2] = t.z; out4[3] = t.w; }
ORC_API void oracle_primary_ray(const SceneData* sd, int px, int py, float jx, float jy, float* out8) {
    Ray r = makePrimaryRay(*sd, px, py, float2(jx, jy));
    out8[0] = r.origin.x; out8[1] = r.origin.y; out8[2] = r.origin.z; out8[3] = r.tMin;
    out8[4] = r.direction.x; out8[5] = r.direction.y; out8[6] = r.direction.z; out8[7] = r.tMax;
}
ORC_API uint16_t oracle_f32_to_f16(float v) { return f32_to_f16(v); }
ORC_API float oracle_f16_to_f32(uint16_t v) { return f16_to_f32(v); }
ORC_API void oracle_xyz_to_srgb(const float* xyz, float* rgb) { float3 r = xyzToLinearSrgb(float3(xyz[0], xyz[1], xyz[2])); rgb[0] = r.x; rgb[1] = r.y; rgb[2] = r.z; }
ORC_API void oracle_spectral_xyz(float lambda, float* xyz) { float3 r = spectralXYZ1931(lambda); xyz[0] = r.x; xyz[1] = r.y; xyz[2] = r.z; }
ORC_API float oracle_rgb2spec_eval(oracle_ctx* c, const float* rgb, float lambda) {
    return spectralScalarFromLinearSrgb(C(c)->spectral, float3(rgb[0], rgb[1], rgb[2]), lambda);
}
// mode 0: evalBSDF (RGB). wo/wi in the local shading frame. out = {value.rgb, pdf}
ORC_API void oracle_bsdf_eval(oracle_ctx* c, const Material* m, const float* wo, const float* wi, uint32_t frontFace, float* out4) {
    BSDFMaterial bm(*m);
    BSDFState st(bm, float3(wo[0], wo[1], wo[2]), frontFace, 0.0f, 0u);
    BSDFEval e = evalBSDF(C(c)->spectral, st, float3(wi[0], wi[1], wi[2]));
    out4[0] = e.value.x; out4[1] = e.value.y; out4[2] = e.value.z; out4[3] = e.pdf;
}
// out = {wi.xyz (local), weight.rgb, pdf, isTransmission}
ORC_API void oracle_bsdf_sample(oracle_ctx* c, const Material* m, const float* wo, uint32_t frontFace, uint32_t* rng, float* out8) {
    BSDFMaterial bm(*m);
    BSDFState st(bm, float3(wo[0], wo[1], wo[2]), frontFace, 0.0f, 0u);
    ShadingBasis b;
    b.tangent = float3(1, 0, 0); b.bitangent = float3(0, 1, 0); b.normal = float3(0, 0, 1);
    BSDFSample s = sampleBSDF(C(c)->spectral, st, b, *rng);
    out8[0] = s.wi.x; out8[1] = s.wi.y; out8[2] = s.wi.z; out8[3] = s.weight.x; out8[4] = s.weight.y; out8[5] = s.weight.z;
    out8[6] = s.pdf; out8[7] = float(s.isTransmission);
}
// hero: out = {value4, techniquePdf4}
ORC_API void oracle_bsdf_eval_hero(oracle_ctx* c, const Material* m, const float* wo, const float* wi, uint32_t frontFace, const float* wl4, float* out8) {
    BSDFMaterial bm(*m);
    float4 wl(wl4[0], wl4[1], wl4[2], wl4[3]);
    BSDFState st(bm, float3(wo[0], wo[1], wo[2]), frontFace, wl.x, 1u);
    float4 tp;
    float4 v = evalSpectralBSDF(C(c)->spectral, st, float3(wi[0], wi[1], wi[2]), wl, tp);
    out8[0] = v.x; out8[1] = v.y; out8[2] = v.z; out8[3] = v.w; out8[4] = tp.x; out8[5] = tp.y; out8[6] = tp.z; out8[7] = tp.w;
}
// Batched closure evaluation, record layout of include/vkrt_closure.h (shared with vkrt_cuda_eval_closures and refshade_eval_closures).
ORC_API int oracle_eval_closures(oracle_ctx* c, const vkrt_closure_query* q, uint32_t count, vkrt_closure_result* out) {
    if (!c || !q || !out) return -1;
    const SpectralTables& T = C(c)->spectral;
    for (uint32_t i = 0; i < count; i++) {
        const vkrt_closure_query& Q = q[i];
        vkrt_closure_result R = {};
        BSDFMaterial bm(Q.material);
        const float3 wo(Q.wo[0], Q.wo[1], Q.wo[2]), wi(Q.wi[0], Q.wi[1], Q.wi[2]);
        const float4 wl(Q.wavelengths[0], Q.wavelengths[1], Q.wavelengths[2], Q.wavelengths[3]);
        BSDFState st(bm, wo, Q.frontFace, Q.mode == 0u ? 0.0f : wl.x, Q.mode == 0u ? 0u : 1u);
        ShadingBasis b;
        b.tangent = float3(1, 0, 0); b.bitangent = float3(0, 1, 0); b.normal = float3(0, 0, 1);
        uint32_t rng = Q.rng;
        if (Q.mode == 2u) {
            float4 tp(0.0f);
            const float4 v = evalSpectralBSDF(T, st, wi, wl, tp);
            R.evalValue[0] = v.x; R.evalValue[1] = v.y; R.evalValue[2] = v.z; R.evalValue[3] = v.w;
            R.evalPdf[0] = tp.x; R.evalPdf[1] = tp.y; R.evalPdf[2] = tp.z; R.evalPdf[3] = tp.w;
            const SpectralBSDFSample s = sampleSpectralBSDF(T, st, b, wl, rng);
            R.sampleWi[0] = s.wi.x; R.sampleWi[1] = s.wi.y; R.sampleWi[2] = s.wi.z;
            R.sampleWeight[0] = s.weight.x; R.sampleWeight[1] = s.weight.y; R.sampleWeight[2] = s.weight.z; R.sampleWeight[3] = s.weight.w;
            R.samplePdf[0] = s.techniquePdf.x; R.samplePdf[1] = s.techniquePdf.y; R.samplePdf[2] = s.techniquePdf.z; R.samplePdf[3] = s.techniquePdf.w;
            R.sampleFlags = (s.isUsable() ? 1u : 0u) | (s.isTransmission != 0u ? 2u : 0u);
        } else {
            const BSDFEval e = Q.mode == 0u ? evalBSDF(T, st, wi) : evalSingleWavelengthBSDF(T, st, wi);
            R.evalValue[0] = e.value.x; R.evalValue[1] = e.value.y; R.evalValue[2] = e.value.z;
            R.evalPdf[0] = e.pdf;
            const BSDFSample s = sampleBSDF(T, st, b, rng);
            R.sampleWi[0] = s.wi.x; R.sampleWi[1] = s.wi.y; R.sampleWi[2] = s.wi.z;
            R.sampleWeight[0] = s.weight.x; R.sampleWeight[1] = s.weight.y; R.sampleWeight[2] = s.weight.z;
            R.samplePdf[0] = s.pdf;
            R.sampleFlags = (s.isUsable() ? 1u : 0u) | (s.isTransmission != 0u ? 2u : 0u);
        }
        R.rngAfter = rng;
        out[i] = R;
    }
    return 0;
}
ORC_API int oracle_watertight(const float* org, const float* dir, const float* v0, const float* v1, const float* v2, float* tuv) {
    RayShear sh = makeRayShear(float3(dir[0], dir[1], dir[2]));
    if (!sh.valid) return 0;
    return watertightTriangle(float3(org[0], org[1], org[2]), sh, float3(v0[0], v0[1], v0[2]), float3(v1[0], v1[1], v1[2]),
                              float3(v2[0], v2[1], v2[2]), tuv[0], tuv[1], tuv[2]) ? 1 : 0;
}

} // extern "C"
