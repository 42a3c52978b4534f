// wavefront.cu — raygen / shade / film kernels of the wavefront path tracer (sm_100a).
//
// Restates, stage by stage, the reference's raygen megakernels:
//   src/shaders/entry/path/raygen_{rgb,spectral_single,spectral_hero}.slang          -> k_raygen + k_film
//   src/shaders/integrator/path/{rgb,spectral_single,spectral_hero}/integrator.slang  -> k_shade<MODE> (one depth per launch)
//   src/shaders/integrator/path/writeback.slang                                       -> k_film
// RNG order per path is the reference's: the single uint rng travels in the path state and is
// consumed as jitter(2) [, wavelength(1)], then per depth NEE(4) -> BSDF selector + lobe -> RR(1).
#include <cooperative_groups.h>

#include "wavefront.cuh"
#include "../../include/vkrt_closure.h"

namespace cg = cooperative_groups;

namespace vk {

constexpr int MODE_RGB = 0, MODE_SINGLE = 1, MODE_HERO = 2;

// ---- pinned-arithmetic camera (camera/ray.slang:15-30): identical roundings to the oracle so that primary rays, and
// therefore primary-hit ids, are bit-exact ---------------------------------------------------------------------------
__device__ __forceinline__ float4 mulMat4Exact(const float* m, float4 v) {
    auto row = [&](int r) {
        return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[r], v.x), __fmul_rn(m[4 + r], v.y)), __fmul_rn(m[8 + r], v.z)), __fmul_rn(m[12 + r], v.w));
    };
    return float4(row(0), row(1), row(2), row(3));
}
__device__ __forceinline__ Ray makePrimaryRayExact(const SceneData& scene, int px, int py, float2 jitter) {
    const int ox = int(scene.viewportRect[0]), oy = int(scene.viewportRect[1]);
    const float vw = float(scene.viewportRect[2]), vh = float(scene.viewportRect[3]);
    const float vx = __fadd_rn(__fadd_rn(float(px - ox), 0.5f), jitter.x);
    const float vy = __fadd_rn(__fadd_rn(float(py - oy), 0.5f), jitter.y);
    const float u = __fdiv_rn(vx, vw), v = __fdiv_rn(vy, vh);
    const float nx = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f), ny = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
    const float4 viewDir = mulMat4Exact(scene.projInverse, float4(nx, ny, 1.0f, 1.0f));
    const float4 org = mulMat4Exact(scene.viewInverse, float4(0.0f, 0.0f, 0.0f, 1.0f));
    const float4 dir = mulMat4Exact(scene.viewInverse, float4(viewDir.x, viewDir.y, viewDir.z, 0.0f));
    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(dir.x, dir.x), __fmul_rn(dir.y, dir.y)), __fmul_rn(dir.z, dir.z));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(len2));
    Ray r;
    r.origin = float3(org.x, org.y, org.z);
    r.direction = float3(__fmul_rn(dir.x, inv), __fmul_rn(dir.y, inv), __fmul_rn(dir.z, inv));
    r.tMin = RAY_T_MIN;
    r.tMax = RAY_T_MAX;
    return r;
}

__device__ __forceinline__ ::float4 toF4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ ::float4 toF4(float4 v) { return make_float4(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ float4 fromF4(::float4 v) { return float4(v.x, v.y, v.z, v.w); }

__device__ __forceinline__ uint32_t allocSlots(uint32_t* counter) {
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, g.size());
    return g.shfl(base, 0) + g.thread_rank();
}

__device__ __forceinline__ float4 heroWavelengths(float unit) {
    return WAVELENGTH_MIN_NM + frac(unit + float4(0.0f, 0.25f, 0.5f, 0.75f)) * WAVELENGTH_RANGE_NM;
}

// ----------------------------------------------------------------------------------------------------------------------
// raygen: RaygenPixelState + PathCommonState.initCommon (integrator/path/state.slang:60-79,140-146,160-196)
// ----------------------------------------------------------------------------------------------------------------------
// Queue slots for the threads of a block that have `want` set: ONE atomic per block (a 1080p x 16 spp chunk is 1 M warps; one atomic per
// warp on the same counter serialises in L2 at ~1 ns each and was a third of k_raygen). Every thread of the block must call it.
__device__ __forceinline__ uint32_t blockAllocSlots(uint32_t* counter, bool want) {
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sBase;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = (blockDim.x + 31) >> 5;
    if (lane == 0) sWarp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        uint32_t c = lane < warps ? sWarp[lane] : 0u, x = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane < warps) sWarp[lane] = x - c;              // exclusive prefix of the warp counts
        if (lane == 31 && x) sBase = atomicAdd(counter, x);
    }
    __syncthreads();
    const uint32_t slot = sBase + sWarp[warp] + __popc(m & ((1u << lane) - 1u));
    __syncthreads();                                         // sWarp / sBase are reused by the next trip
    return slot;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_raygen(const FrameParams fp) {
    const uint32_t lpc = fp.tiles.localPixelCount;
    const uint32_t total = fp.chunkSamples * lpc;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        const uint32_t t = base + threadIdx.x;
        bool valid = t < total;
        uint32_t rng = 0u;
        float unit = 0.0f;
        Ray ray;
        if (valid) {
            // zero the sample record (coalesced)
            fp.rec.radiance[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            fp.rec.featA[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            fp.rec.featB[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            fp.rec.follow[t] = 0.0f;
            if (MODE == MODE_HERO) fp.rec.radianceScalar[t] = 0.0f;
            const uint32_t s = t / lpc, lp = t - s * lpc;
            uint32_t gx, gy;
            valid = localPixelToGlobal(fp.tiles, fp.tiles.localToGlobalTile, lp, gx, gy) && insideViewport(fp.sd, (int)gx, (int)gy);
            if (valid) {
                const float prevW = fp.film.accum[fp.readIndex][lp].w;
                const uint32_t previousSamples = (uint32_t)(prevW + 0.5f);
                const uint32_t sampleIndex = fp.chunkFirstSample + s;
                rng = initPixelSeed((int)gx, (int)gy, fp.sd.frameNumber, previousSamples + sampleIndex);
                const float jx = rand(rng);
                const float jy = rand(rng);
                ray = makePrimaryRayExact(fp.sd, (int)gx, (int)gy, float2(__fsub_rn(jx, 0.5f), __fsub_rn(jy, 0.5f)));
                if (MODE != MODE_RGB) {
                    unit = sampleUniformWavelengthUnit(rng, previousSamples + sampleIndex);
                    fp.rec.unitWavelength[t] = unit;
                }
            }
        }
        const uint32_t j = blockAllocSlots(fp.extCount, valid);
        if (!valid) continue;
        const PathState& S = fp.st[0];
        S.rayO[j] = toF4(ray.origin, ray.tMin);
        S.rayD[j] = toF4(ray.direction, ray.tMax);
        S.meta[j] = make_uint4(t, rng, MODE == MODE_HERO ? (uint32_t)PF_HERO_ACTIVE : 0u, 0u);
        // The rest of a fresh path's state is constant (throughput 1, technique pdfs 1 / 0, prevBsdfPdf 0) or a function of `unit`:
        // k_shade substitutes it at depth 0 instead of reading it back, which saves 64 B written and 64 B gathered per camera path.
        if (MODE == MODE_SINGLE) {
            const float lambda = WAVELENGTH_MIN_NM + saturate(unit) * WAVELENGTH_RANGE_NM;
            S.thr[j] = make_float4(1.f, lambda, 0.f, 0.f);
        } else if (MODE == MODE_HERO) {
            S.heroMisc[j] = make_float4(unit, 1.f, 0.f, 0.f);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// Surface reconstruction (geometry/surface/reconstruct.slang:8-49, integrator/path/surface_state.slang:8-31)
// ----------------------------------------------------------------------------------------------------------------------
struct SurfaceShadingData {
    uint32_t materialIndex;
    uint32_t frontFace;
    float3 shadingNormal, geometricNormal;
    float4 tangent;
    SurfaceTextureData textureData;
};

__device__ __forceinline__ SurfaceShadingData reconstructSurfaceShading(const SceneView& sc, const MeshInfo& mesh, const MeshTrig& trig, uint32_t prim,
                                                                        float2 bary, float3 worldRayDir) {
    TriangleVertices tv;
    loadTriangleVertices(sc, mesh, prim, tv);
    float3 objectNormal = safeNormalize(interp3(unpackOctNormal(tv.v0.packedNormal), unpackOctNormal(tv.v1.packedNormal),
                                                unpackOctNormal(tv.v2.packedNormal), bary));
    float4 objectTangent = interp4(unpackOctTangent(tv.v0.packedTangent), unpackOctTangent(tv.v1.packedTangent),
                                   unpackOctTangent(tv.v2.packedTangent), bary);
    float handedness = objectTangent.w < 0.0f ? -1.0f : 1.0f;
    float3 shadingNormalUnoriented = meshTransformNormal(mesh, trig, objectNormal);
    float facing = dot(shadingNormalUnoriented, worldRayDir) > 0.0f ? -1.0f : 1.0f;
    float3 p0(tv.v0.position[0], tv.v0.position[1], tv.v0.position[2]);
    float3 p1(tv.v1.position[0], tv.v1.position[1], tv.v1.position[2]);
    float3 p2(tv.v2.position[0], tv.v2.position[1], tv.v2.position[2]);
    float3 worldEdge1 = meshTransformVector(mesh, trig, p1 - p0);
    float3 worldEdge2 = meshTransformVector(mesh, trig, p2 - p0);
    float3 geometricNormal = safeNormalize(cross(worldEdge1, worldEdge2));
    if (surfaceTransformSign(mesh) < 0.0f) geometricNormal = -geometricNormal;
    SurfaceShadingData hit;
    hit.materialIndex = mesh.materialIndex;
    hit.frontFace = dot(geometricNormal, worldRayDir) < 0.0f ? 1u : 0u;
    hit.shadingNormal = shadingNormalUnoriented * facing;
    hit.geometricNormal = hit.frontFace != 0u ? geometricNormal : -geometricNormal;
    hit.tangent = float4(safeNormalize(meshTransformVector(mesh, trig, objectTangent.xyz())) * facing, handedness);
    hit.textureData = evaluateSurfaceTextureData(tv, bary);
    return hit;
}

// material/textures.slang:123-155
__device__ __forceinline__ void normalizeEmission(Material& m, float3 emissiveTexel) {
    float3 emission = float3(m.emissionColor[0], m.emissionColor[1], m.emissionColor[2]) * m.emissionLuminance * emissiveTexel;
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
__device__ __forceinline__ void applySurfaceTextures(const SceneView& sc, Material& m, const SurfaceTextureData& s) {
    float4 bc = sampleBaseColorTexture(sc, m, s);
    m.baseColor[0] *= bc.x * s.color.x;
    m.baseColor[1] *= bc.y * s.color.y;
    m.baseColor[2] *= bc.z * s.color.z;
    float4 mr = sampleMetallicRoughnessTexture(sc, m, s);
    m.roughness = saturate(m.roughness * mr.y);
    m.metallic = saturate(m.metallic * mr.z);
    float4 et = sampleEmissiveTexture(sc, m, s);
    normalizeEmission(m, et.xyz());
}
__device__ __forceinline__ float3 applyNormalTexture(const SceneView& sc, const Material& m, const SurfaceTextureData& s, const ShadingBasis& basis) {
    if (m.normalTextureIndex == VKRT_INVALID_INDEX) return basis.normal;
    float3 ns = sampleNormalTexture(sc, m, s).xyz() * 2.0f - 1.0f;
    ns.x *= m.normalTextureScale;
    ns.y *= m.normalTextureScale;
    ns = safeNormalize(ns);
    return safeNormalize(basis.tangent * ns.x + basis.bitangent * ns.y + basis.normal * ns.z);
}

__device__ __forceinline__ Material loadMaterial(const Material* p) {
    // 272-byte material = 17 x 16-byte read-only loads (L1/L2 resident: a scene has few materials)
    Material m;
    const ::float4* q = reinterpret_cast<const ::float4*>(p);
    ::float4* d = reinterpret_cast<::float4*>(&m);
#pragma unroll
    for (int i = 0; i < 17; i++) d[i] = __ldg(q + i);
    return m;
}
__device__ __forceinline__ MeshInfo loadMeshInfo(const MeshInfo* p) {
    MeshInfo m;
    const ::float4* q = reinterpret_cast<const ::float4*>(p);
    ::float4* d = reinterpret_cast<::float4*>(&m);
#pragma unroll
    for (int i = 0; i < 5; i++) d[i] = __ldg(q + i);
    return m;
}

// light/environment.slang:9-27
__device__ __forceinline__ float3 sampleEnvironmentRadiance(const SceneView& sc, const SceneData& scene, float3 worldDir) {
    if (scene.environmentTextureIndex == VKRT_INVALID_INDEX)
        return float3(scene.environmentLight[0], scene.environmentLight[1], scene.environmentLight[2]);
    float3 dir = normalize(worldDir);
    float phi = atan2f(dir.y, dir.x) + scene.environmentRotation * (PI / 180.0f);
    float theta = acosf(clamp(dir.z, -1.0f, 1.0f));
    float2 uv(frac(phi * (0.5f * INV_PI) + 0.5f), theta * INV_PI);
    float3 radiance = sampleTextureBilinear(sc, scene.environmentTextureIndex, uv, VKRT_TEXTURE_WRAP_REPEAT, VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE).xyz();
    return radiance * scene.environmentLight[3];
}

// light/direct/light_sampling.slang:9-63
struct DirectLightSurfaceSample {
    float3 emission, wi;
    float shadowDistance, pdfSolidAngle;
    uint emissiveMesh;   // index of the sampled emissive mesh (its emission has a spectral memo entry), 0xffffffff for the environment
    bool valid;
};

// ---- environment-map importance sampling (extension, see EnvDistribution) ----------------------------------------------------------
// Solid-angle density of sampleEnvironmentLight for a direction: the inverse of sampleEnvironmentRadiance's lat-long mapping.
static __device__ __noinline__ float environmentPdf(const SceneView& sc, const SceneData& scene, float3 worldDir) {
    const EnvDistribution& E = sc.env;
    const float3 dir = normalize(worldDir);
    const float phi = atan2f(dir.y, dir.x) + scene.environmentRotation * (PI / 180.0f);
    const float theta = acosf(clamp(dir.z, -1.0f, 1.0f));
    const float sinTheta = sinf(theta);
    if (!(sinTheta > 1e-6f)) return 0.0f;
    const float u = frac(phi * (0.5f * INV_PI) + 0.5f), v = theta * INV_PI;
    const uint32_t x = min((uint32_t)(u * float(E.width)), E.width - 1u), y = min((uint32_t)(v * float(E.height)), E.height - 1u);
    return __ldg(E.pdfUv + (size_t)y * E.width + x) / (2.0f * PI * PI * sinTheta);
}
// One direction towards the environment: texel by the alias method (48 random bits for the cell, so that millions of texels are
// addressed uniformly, one more number for the alias coin), uniform inside the texel. 5 random numbers.
static __device__ __noinline__ DirectLightSurfaceSample sampleEnvironmentLight(const SceneView& sc, const SceneData& scene, uint& rng) {
    const EnvDistribution& E = sc.env;
    DirectLightSurfaceSample s;
    s.valid = false;
    s.shadowDistance = s.pdfSolidAngle = 0.0f;
    s.emissiveMesh = 0xffffffffu;
    const uint32_t n = E.width * E.height;
    const uint64_t hi = (uint64_t)(rand(rng) * 16777216.0f), lo = (uint64_t)(rand(rng) * 16777216.0f);
    // 48 random bits x n needs a 128-bit product: in 64 bits it wraps as soon as n > 65536 texels, and every map larger than 256 x 256
    // would then start in its first 65536 texels only. High half of (bits << 16) * n = floor(bits * n / 2^48).
    uint32_t texel = (uint32_t)__umul64hi(((hi << 24) | lo) << 16, (uint64_t)n);
    texel = min(texel, n - 1u);
    if (!(rand(rng) < __ldg(E.aliasQ + texel))) texel = __ldg(E.aliasIdx + texel);
    const uint32_t ty = texel / E.width, tx = texel - ty * E.width;
    const float u = (float(tx) + rand(rng)) / float(E.width);
    const float v = (float(ty) + rand(rng)) / float(E.height);
    const float phi = (u - 0.5f) * (2.0f * PI) - scene.environmentRotation * (PI / 180.0f);
    const float theta = v * PI;
    const float sinTheta = sinf(theta), cosTheta_ = cosf(theta);
    if (!(sinTheta > 1e-6f)) return s;
    s.wi = float3(sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta_);
    s.pdfSolidAngle = __ldg(E.pdfUv + texel) / (2.0f * PI * PI * sinTheta);
    if (!(s.pdfSolidAngle > 0.0f)) return s;
    s.emission = sampleEnvironmentRadiance(sc, scene, s.wi);
    s.shadowDistance = RAY_T_MAX;
    s.valid = true;
    return s;
}
__device__ __forceinline__ DirectLightSurfaceSample sampleDirectLightSurface(const SceneView& sc, const SceneData& scene, float3 hitPoint, uint& rng) {
    DirectLightSurfaceSample s;
    s.valid = false;
    s.shadowDistance = s.pdfSolidAngle = 0.0f;
    s.emissiveMesh = 0xffffffffu;
    const uint meshCount = scene.emissiveMeshCount;
    if (meshCount == 0u) return s;
    const uint meshIdx = sampleAlias(rand(rng), meshCount, 0u, sc.meshAliasQ, sc.meshAliasIdx);
    s.emissiveMesh = meshIdx;
    const EmissiveMesh em = sc.emissiveMeshes[meshIdx];
    if (em.triCount == 0u) return s;
    const uint localTri = sampleAlias(rand(rng), em.triCount, em.triOffset, sc.triAliasQ, sc.triAliasIdx);
    const ::float4* tp = reinterpret_cast<const ::float4*>(sc.emissiveTriangles + em.triOffset + localTri);
    const ::float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
    const float u1 = rand(rng);
    const float u2 = rand(rng);
    const float sqrtU1 = sqrt(u1);
    const float b1 = u2 * sqrtU1;
    const float b2 = (1.0f - u2) * sqrtU1;
    const float3 v0(t0.x, t0.y, t0.z), e1(t1.x, t1.y, t1.z), e2(t2.x, t2.y, t2.z);
    const float3 position = v0 + b1 * e1 + b2 * e2;
    const float3 normal = safeNormalize(cross(e1, e2));
    s.emission = float3(em.emission[0], em.emission[1], em.emission[2]);
    const float pdf = em.pmfMesh * em.invTotalArea;
    if (!(pdf > 0.0f)) return s;
    const float3 toLight = position - hitPoint;
    const float d2 = dot(toLight, toLight);
    if (d2 <= 0.0f) return s;
    const float invD = rsqrt(d2);
    const float distance = d2 * invD;
    s.wi = toLight * invD;
    const float cosLight = abs(dot(s.wi, normal));
    if (cosLight <= 0.0f) return s;
    s.pdfSolidAngle = pdf * d2 / cosLight;
    if (s.pdfSolidAngle <= 0.0f) return s;
    s.shadowDistance = distance - SHADOW_DISTANCE_OFFSET;
    if (s.shadowDistance <= 0.0f) return s;
    s.valid = true;
    return s;
}

// light/direct/mis_weights.slang:6-42
__device__ __forceinline__ float lightPdfAreaToSolidAngle(float lightPdfArea, float3 gn, float3 rd, float hitDistance) {
    if (lightPdfArea <= 0.0f) return 0.0f;
    float d2 = hitDistance * hitDistance;
    if (d2 <= 0.0f) return 0.0f;
    float cosLight = abs(dot(rd, gn));
    if (cosLight <= 0.0f) return 0.0f;
    return lightPdfArea * d2 / cosLight;
}
__device__ __forceinline__ float computeSpectralMISWeight(float4 sampled, float4 alternate) {
    float denom = dot(sampled, float4(1.0f)) + dot(alternate, float4(1.0f));
    return denom > 0.0f ? sampled.x / denom : 0.0f;
}

// ----------------------------------------------------------------------------------------------------------------------
// shade: one path vertex (the body of the reference's depth loop between two TraceRay calls)
// ----------------------------------------------------------------------------------------------------------------------
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 2   // 2 x 256 threads x 128 registers: same 16 warps per SM as 1 x 512, but a phase barrier stalls 8 warps, not 16
#endif
#ifndef SHADE_BLOCK
#define SHADE_BLOCK 256      // measured with the state in shared memory (profiles/r02_notes.md): 512 x 1 29.0 ms, 384 x 1 29.4, 256 x 2 27.2,
                             // 192 x 3 (112 registers) 30.4, 128 x 4 28.3
#endif
// The body is cut into four phases separated by block barriers. k_shade is bound by instruction fetch, not by ALU or memory
// (profiles/r01_notes.md: "no instruction" is the top stall with ~200 KB of SASS live); keeping the warps of a block inside the same
// phase makes them share the instruction lines they fetch. Every thread of a block runs the same number of loop trips and reaches
// every barrier; a thread whose vertex is finished (miss, absorbed, debug view, out of range) just carries live = false through.
#ifdef SHADE_NO_BARRIER
#define SHADE_PHASE_BARRIER()
#else
#define SHADE_PHASE_BARRIER() __syncthreads()
#endif
#ifdef SHADE_FINE_BARRIERS
#define SHADE_SUBPHASE_BARRIER() __syncthreads()
#else
#define SHADE_SUBPHASE_BARRIER()
#endif
// Per-thread record in SHARED memory for what a vertex carries across the phases of k_shade. The closure state is passed by reference
// to out-of-line lobe functions, so it has to live in memory; as part of the kernel's local frame (0.5 KB x 512 threads) it overflowed
// the L1 (r02c capture: L1 hit rate 58 %, lg_throttle on the state stores, long scoreboard 4.8 per issue). In shared memory the state,
// the shading basis and the hit geometry cost no registers between phases and never leave the SM: hero shade 33.0 -> 27.2 ms per
// 33 M-path frame together with 256-thread blocks (profiles/r02_notes.md). Putting the pending shadow ray and the hero technique pdfs
// there as well was slower (30.5 ms): 2 x 100 KB of records leave the L1 only 28 KB for the spill slots.
// Records sit at a stride of an odd number of 8-byte units: 32-bit accesses conflict 2-way at worst.
struct ShadeScratch {
    BSDFState state;
    ShadingBasis basis;
    float3 geometricNormal, hitPoint, emission;
};
#ifdef SHADE_STRIDE16
constexpr size_t SHADE_SCRATCH_STRIDE = (((sizeof(ShadeScratch) + 15) / 16) | 1) * 16;
#else
static_assert(alignof(ShadeScratch) <= 8, "ShadeScratch must not need more than 8-byte alignment");
constexpr size_t SHADE_SCRATCH_STRIDE = (((sizeof(ShadeScratch) + 7) / 8) | 1) * 8;   // an odd number of 8-byte units: 2-way conflicts at worst
#endif
constexpr size_t SHADE_DYNAMIC_SMEM = SHADE_SCRATCH_STRIDE * SHADE_BLOCK;
static_assert((SHADE_DYNAMIC_SMEM + 1024 + 1280 + 64) * SHADE_MIN_BLOCKS <= 228 * 1024, "the per-vertex records of SHADE_MIN_BLOCKS blocks must fit the SM's shared memory (228 KB; 1 KB reserved and 1.25 KB of static tables per block)");
template <int MODE, bool ENVIS>
__global__ void __launch_bounds__(SHADE_BLOCK, SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ FrameParams fp, const uint32_t depth) {
    const uint32_t count = fp.extCount[depth];
    const PathState& S = fp.st[depth & 1u];
    const PathState& N = fp.st[(depth & 1u) ^ 1u];
    const SceneView& sc = fp.scene;
    const SceneData& scene = fp.sd;
    // spectral modes: the rgb2spec scale axis (res floats, searched several times per vertex) is staged in shared memory
    __shared__ float sScale[RGB2SPEC_SMEM_RES];
    SpectralTables T = sc.spectral;
    if (MODE != MODE_RGB && T.info.res <= RGB2SPEC_SMEM_RES) {
        if (threadIdx.x < T.info.res) sScale[threadIdx.x] = T.scale[threadIdx.x];
        T.scale = sScale;
        __syncthreads();
#ifndef SHADE_NO_INTERVAL_LUT
        if (T.info.res == 64u) {   // start points of the scale-axis search (SpectralTables::intervalLut), rebuilt per launch: 1 KB, ~250 compares per thread
            __shared__ unsigned char sIntervalLut[RGB2SPEC_LUT_SIZE];
            for (uint32_t q = threadIdx.x; q < RGB2SPEC_LUT_SIZE; q += blockDim.x) {
                const float xq = float(q) * (1.0f / float(RGB2SPEC_LUT_SIZE));
                uint32_t k = 0u;
                for (uint32_t e = 1u; e <= 62u; e++) k += sScale[e] <= xq ? 1u : 0u;
                sIntervalLut[q] = (unsigned char)k;
            }
            T.intervalLut = sIntervalLut;
            __syncthreads();
        }
#endif
    }
    const uint32_t lpc = fp.tiles.localPixelCount;
    const bool neeEnabled = (fp.modeFlags & MODE_NEE_ENABLED) && !(fp.modeFlags & MODE_BSDF_ONLY);
    const bool neeOnly = (fp.modeFlags & MODE_NEE_ONLY) != 0u;

    const uint32_t* __restrict__ order = fp.shadeOrder;

    for (uint32_t base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        bool live = j < count;
        const uint32_t i = (live && order) ? __ldg(order + j) : j;   // material-sorted order: neighbouring lanes shade the same closure

        // ================================ phase 1: unpack, miss, surface, material, closure state ================================
        Ray ray;
        uint32_t rec = 0u, rng = 0u, flags = 0u, hitInst = VKRT_INVALID_INDEX, hitPrim = 0u, lp = 0u, sampleIndex = 0u;
        float hitT = 0.0f, hitU = 0.0f, hitV = 0.0f;
        float3 thrRgb(0.0f);
        float thrScalar = 0.0f, prevBsdfPdf = 0.0f, lambdaScalar = 0.0f, unit = 0.0f;
        float4 thr4(0.0f), wl4(0.0f), techPdf(0.0f);
        bool heroActive = false;
        MediumState medium;
        float lightPdfArea = 0.0f;
        extern __shared__ __align__(16) unsigned char shadeDynamicSmem[];   // (per-thread records need 8-byte alignment only)
        ShadeScratch& scratch = *reinterpret_cast<ShadeScratch*>(shadeDynamicSmem + (size_t)threadIdx.x * SHADE_SCRATCH_STRIDE);
        BSDFState& state = scratch.state;        // holds the only copy of the closure parameters that outlives phase 1 (state.material)
        ShadingBasis& basis = scratch.basis;
        float3& hitPoint = scratch.hitPoint;
        float3& geometricNormal = scratch.geometricNormal;
        float3& emission = scratch.emission;     // textured, per-hit emission (integrator.slang:79-85)
        bool currentVertexNeeAllowed = false;

        if (live) {
            const ::float4 ro = SQ_LD(S.rayO + i), rd = SQ_LD(S.rayD + i);
            const ::uint4 ha = SQ_LD(fp.hitA + i);
            const ::uint4 slotMeta = SQ_LD(S.meta + i);
            rec = slotMeta.x;
            rng = slotMeta.y;
            flags = slotMeta.z;
            hitV = __uint_as_float(slotMeta.w);
            // depth 0: the constant part of a fresh path's state is not stored (k_raygen)
            const ::float4 thrRaw = (depth == 0u && MODE != MODE_SINGLE) ? make_float4(1.f, 1.f, 1.f, MODE == MODE_RGB ? 0.f : 1.f) : SQ_LD(S.thr + i);
            ray.origin = float3(ro.x, ro.y, ro.z);
            ray.direction = float3(rd.x, rd.y, rd.z);
            hitInst = ha.x;
            hitPrim = ha.y;
            hitT = __uint_as_float(ha.z);
            hitU = __uint_as_float(ha.w);
            const uint32_t sampleInChunk = rec / lpc;
            lp = rec - sampleInChunk * lpc;
            sampleIndex = fp.chunkFirstSample + sampleInChunk;

            // ---- unpack mode state ---------------------------------------------------------------------------------
            if (MODE == MODE_RGB) {
                thrRgb = float3(thrRaw.x, thrRaw.y, thrRaw.z);
                prevBsdfPdf = thrRaw.w;
            } else if (MODE == MODE_SINGLE) {
                thrScalar = thrRaw.x;
                lambdaScalar = thrRaw.y;
                prevBsdfPdf = thrRaw.z;
            } else {
                thr4 = fromF4(thrRaw);
                const ::float4 misc = SQ_LD(S.heroMisc + i);
                unit = misc.x;
                thrScalar = misc.y;
                prevBsdfPdf = misc.z;
                wl4 = heroWavelengths(unit);
                lambdaScalar = wl4.x;
                heroActive = (flags & PF_HERO_ACTIVE) != 0u;
                techPdf = depth == 0u ? float4(1.0f) : fromF4(SQ_LD(S.techPdf + i));
            }
            medium.flags = (flags >> 1) & 3u;
            if (medium.absorptionActive()) {
                const ::float4 sg = SQ_LD(S.sigma + i);
                medium.absorptionSigma = float3(sg.x, sg.y, sg.z);
                medium.spectralAbsorptionSigma = fromF4(sg);
                if (MODE == MODE_SINGLE || (MODE == MODE_HERO && !heroActive)) medium.spectralAbsorptionSigma = float4(sg.x);
            }

            // ---- miss: environment (loop.slang:4-6, */transport.slang accumulate*Environment) -------------------------
            if (hitInst == VKRT_INVALID_INDEX) {
                if (!neeOnly && !mediumHasActiveBoundary(medium)) {
                    float3 env = sampleEnvironmentRadiance(sc, scene, ray.direction);
                    float heroWeight = (MODE == MODE_HERO && heroActive) ? heroWavelengthBalanceWeight(techPdf) : 1.0f;
                    if (ENVIS) {
                        // the environment is also reached by next-event estimation: weight the BSDF-sampled arrival exactly like an
                        // emitter hit (integrator.slang:79-85 / mis_weights.slang:18-42 with the environment's solid-angle density)
                        if ((fp.modeFlags & MODE_NEE_ENABLED) && (flags & PF_PREV_VERTEX_NEE_ALLOWED) && depth > 0u) {
                            const float lp2 = environmentPdf(sc, scene, ray.direction) * sc.env.pEnv;
                            if (MODE == MODE_HERO && heroActive) {
                                const float4 pv = fromF4(SQ_LD(S.prevVertexTechPdf + i)), pb = fromF4(SQ_LD(S.prevBsdfTechPdf + i));
                                heroWeight = computeSpectralMISWeight(pv * pb, pv * lp2);
                            } else if (prevBsdfPdf > 0.0f && lp2 > 0.0f) {
                                env = env * powerHeuristic(prevBsdfPdf, lp2);
                            }
                        }
                    }
                    if (MODE == MODE_RGB) {
                        ::float4 r = fp.rec.radiance[rec];
                        const float3 c = thrRgb * env;
                        r.x += c.x; r.y += c.y; r.z += c.z;
                        fp.rec.radiance[rec] = r;
                    } else if (MODE == MODE_SINGLE) {
                        fp.rec.radiance[rec].x += thrScalar * spectralScalarFromLinearSrgb(T, env, lambdaScalar);
                    } else if (heroActive) {
                        const float4 c = thr4 * heroWeight * spectralScalarFromLinearSrgb4(T, env, wl4);
                        ::float4 r = fp.rec.radiance[rec];
                        r.x += c.x; r.y += c.y; r.z += c.z; r.w += c.w;
                        fp.rec.radiance[rec] = r;
                    } else {
                        fp.rec.radianceScalar[rec] += thrScalar * spectralScalarFromLinearSrgb(T, env, lambdaScalar);
                    }
                }
                live = false;
            }
        }

        SHADE_SUBPHASE_BARRIER();

        // ---- medium transmittance along the segment (apply*MediumTransmittance) ------------------------------------
        if (live) {
            if (medium.absorptionActive()) {
                if (MODE == MODE_RGB) thrRgb *= mediumTransmittance(medium, hitT);
                else if (MODE == MODE_HERO && heroActive) thr4 *= mediumSpectralTransmittance(medium, hitT);
                else thrScalar *= mediumTransmittance(medium, hitT).x;
            }
            const bool alive = MODE == MODE_RGB ? anyGreater(thrRgb, 0.0f)
                                                : ((MODE == MODE_HERO && heroActive) ? anyGreater(thr4, 0.0f) : thrScalar > 0.0f);
            if (!alive) live = false;
        }

        // ---- surface (PathSurfaceState.__init) -----------------------------------------------------------------------
        if (live) {
            const MeshInfo mesh = loadMeshInfo(sc.meshInfos + hitInst);
            lightPdfArea = mesh.lightPdfArea;
            hitPoint = ray.origin + ray.direction * hitT;
            MeshTrig trig;
            {
                const ::float4* tq = reinterpret_cast<const ::float4*>(sc.meshTrig + hitInst);
                const ::float4 t0 = __ldg(tq), t1 = __ldg(tq + 1);
                trig.sx = t0.x; trig.cx = t0.y; trig.sy = t0.z; trig.cy = t0.w; trig.sz = t1.x; trig.cz = t1.y; trig.pad0 = trig.pad1 = 0.0f;
            }
            SurfaceShadingData surface = reconstructSurfaceShading(sc, mesh, trig, hitPrim, float2(hitU, hitV), ray.direction);
            Material material = loadMaterial(sc.materials + surface.materialIndex);
            if (material.normalTextureIndex != VKRT_INVALID_INDEX) {
                const ShadingBasis unperturbed = makeShadingBasis(surface.shadingNormal, surface.tangent);
                surface.shadingNormal = applyNormalTexture(sc, material, surface.textureData, unperturbed);
            } else {   // without a texture applyNormalTexture returns the basis normal: the normalised shading normal; the tangent frame is not needed
                surface.shadingNormal = safeNormalize(surface.shadingNormal);
            }
            surface.shadingNormal = sanitizeShadingNormal(surface.shadingNormal, surface.geometricNormal, -ray.direction);
            basis = makeShadingBasis(surface.shadingNormal, surface.tangent);
            geometricNormal = surface.geometricNormal;
            applySurfaceTextures(sc, material, surface.textureData);
            emission = float3(material.emissionColor[0], material.emissionColor[1], material.emissionColor[2]) * material.emissionLuminance;
            const BSDFMaterial bm = makeBSDFMaterial(material);

            // ---- primary-surface debug views (integrator/path/debug.slang:31-64) -----------------------------------------
            if (sampleIndex == 0u && depth == 0u && scene.debugMode != VKRT_DEBUG_MODE_NONE) {
                const uint32_t dm = scene.debugMode;
                bool handled = true;
                float3 c(0.0f);
                if (dm == VKRT_DEBUG_MODE_NORMALS) c = surface.shadingNormal * 0.5f + 0.5f;
                else if (dm == VKRT_DEBUG_MODE_DEPTH) c = float3(1.0f / (1.0f + hitT));
                else if (dm == VKRT_DEBUG_MODE_BASE_COLOR_MAP) c = sampleBaseColorTexture(sc, material, surface.textureData).xyz();
                else if (dm == VKRT_DEBUG_MODE_METALLIC_MAP) c = float3(sampleMetallicRoughnessTexture(sc, material, surface.textureData).z);
                else if (dm == VKRT_DEBUG_MODE_ROUGHNESS_MAP) c = float3(sampleMetallicRoughnessTexture(sc, material, surface.textureData).y);
                else if (dm == VKRT_DEBUG_MODE_NORMAL_MAP) c = sampleNormalTexture(sc, material, surface.textureData).xyz();
                else if (dm == VKRT_DEBUG_MODE_EMISSIVE_MAP) c = sampleEmissiveTexture(sc, material, surface.textureData).xyz();
                else handled = false;
                if (handled) {
                    fp.film.debugColor[lp] = toF4(c, 1.0f);
                    live = false;
                }
            }
          if (live) {
            // ---- denoiser features (loop.slang:83-103) ----------------------------------------------------------------
            // (the specular-follow test is needed only while the features of this path are unresolved, and for the follow image at depth 0)
            const bool follow = (depth == 0u || !(flags & PF_FEATURES_RESOLVED)) && materialDenoiserShouldFollowSpecularHit(bm, surface.frontFace);
            if (depth == 0u) fp.rec.follow[rec] = follow ? 1.0f : 0.0f;
            if (!(flags & PF_FEATURES_RESOLVED) && !follow) {
                fp.rec.featA[rec] = toF4(materialDenoiserAlbedo(bm), 1.0f);
                fp.rec.featB[rec] = toF4(surface.shadingNormal, float(depth + 1u));
                flags |= PF_FEATURES_RESOLVED;
            }
            const float stateWavelength = MODE == MODE_RGB ? 0.0f : lambdaScalar;
            state = BSDFState(bm, worldToLocal(-ray.direction, basis), surface.frontFace, stateWavelength, MODE == MODE_RGB ? 0u : 1u);
            if (MODE != MODE_RGB) state.memoIndex = surface.materialIndex * SPECTRAL_MEMO_SLOTS;
            currentVertexNeeAllowed = !medium.refractiveActive();
          }
        }
        SHADE_PHASE_BARRIER();

        // ================================ phase 2: next-event estimation (light/direct/*.slang) ================================
        // The shadow ray is traced by the next k_trace launch.
        bool shadowPending = false;
        ::float4 shadowO = make_float4(0.f, 0.f, 0.f, 0.f), shadowD = shadowO, shadowC = shadowO;
        uint32_t shadowSeed = 0u;
        bool shadowScalarLane = false;
        if (live && neeEnabled && currentVertexNeeAllowed && cosTheta(state.wo) > 0.0f) {
            DirectLightSurfaceSample ls;
            if (ENVIS) {
                // one light sample per vertex, as in the reference: it goes to the environment with probability pEnv (a random number
                // is spent on the choice only when the scene has emissive triangles too), and both densities carry that probability
                const float pEnv = sc.env.pEnv;
                if (pEnv >= 1.0f || rand(rng) < pEnv) {
                    ls = sampleEnvironmentLight(sc, scene, rng);
                    ls.pdfSolidAngle *= pEnv;
                    if (mediumHasActiveBoundary(medium)) ls.valid = false;  // the environment is not seen from inside a medium (loop.slang:4-6)
                } else {
                    ls = sampleDirectLightSurface(sc, scene, hitPoint, rng);
                    ls.pdfSolidAngle *= 1.0f - pEnv;
                }
            } else {
                ls = sampleDirectLightSurface(sc, scene, hitPoint, rng);
            }
            if (ls.valid) {
                const float3 gn = geometricNormal;
                const float3 shadowOffset = dot(ls.wi, gn) >= 0.0f ? gn : -gn;
                const float3 shadowOrigin = hitPoint + shadowOffset * SHADOW_ORIGIN_OFFSET;
                const float3 wiLocal = worldToLocal(ls.wi, basis);
                // common.slang:47-50 rejects the sample after the visibility test; evaluating it first is equivalent for the radiance.
                float4 c(0.0f);
                const SpectralMemoEntry* lightMemo = (MODE != MODE_RGB && T.emissiveMemo && ls.emissiveMesh != 0xffffffffu) ? T.emissiveMemo + ls.emissiveMesh : nullptr;
                const bool refractiveReject = materialMediumIsRefractive(state.material) && cosTheta(wiLocal) <= 0.0f;
                // A light below the shading horizon of an opaque surface contributes exactly zero (the reflection stack needs cos > 0, the
                // transmission lobe needs transmission > 0): not evaluating it keeps the warp out of the transmission code altogether.
                const bool belowOpaqueHorizon = cosTheta(wiLocal) <= 0.0f && !(state.material.transmission > 0.0f);
                if (!refractiveReject && !belowOpaqueHorizon) {
                    if (MODE == MODE_RGB) {
                        const BSDFEval e = evalBSDF(T, state, wiLocal);
                        if (e.pdf > 0.0f) {
                            const float misWeight = powerHeuristic(ls.pdfSolidAngle, e.pdf);
                            const float3 fCos = e.value * absCosTheta(wiLocal);
                            const float3 r = misWeight * mediumTransmittance(medium, ls.shadowDistance) * fCos * ls.emission / ls.pdfSolidAngle;
                            const float3 tr = thrRgb * r;
                            c = float4(tr.x, tr.y, tr.z, 0.0f);
                        }
                    } else if (MODE == MODE_HERO && heroActive) {
                        float4 localTp(0.0f);
                        const float4 fCos = evalSpectralBSDF(T, state, wiLocal, wl4, localTp) * absCosTheta(wiLocal);
                        const float misWeight = computeSpectralMISWeight(techPdf * ls.pdfSolidAngle, techPdf * localTp);
                        if (misWeight > 0.0f) {
                            c = thr4 * (misWeight * mediumSpectralTransmittance(medium, ls.shadowDistance) * fCos *
                                        spectralScalarFromLinearSrgb4(T, lightMemo, ls.emission, wl4) / ls.pdfSolidAngle);
                        }
                    } else {
                        const BSDFEval e = evalSingleWavelengthBSDF(T, state, wiLocal);
                        if (e.pdf > 0.0f) {
                            const float sv = e.value.x * absCosTheta(wiLocal) * spectralScalarFromLinearSrgb(T, lightMemo, ls.emission, lambdaScalar);
                            const float misWeight = powerHeuristic(ls.pdfSolidAngle, e.pdf);
                            c.x = thrScalar * (misWeight * mediumTransmittance(medium, ls.shadowDistance).x * sv / ls.pdfSolidAngle);
                            shadowScalarLane = MODE == MODE_HERO;
                        }
                    }
                }
                // A shadow ray whose contribution is zero (light below the shading horizon, zero MIS weight) can only be observed through
                // "unsupported transmission" (it clears the next vertex's NEE flag): with no transmissive instance in the scene it is dropped.
                shadowPending = (fp.modeFlags & MODE_SCENE_TRANSMISSIVE) || c.x != 0.0f || c.y != 0.0f || c.z != 0.0f || c.w != 0.0f;
                shadowO = toF4(shadowOrigin, RAY_T_MIN);
                shadowD = toF4(ls.wi, ls.shadowDistance);
                shadowC = toF4(c);
                shadowSeed = rng;
            }
        }
        SHADE_PHASE_BARRIER();

        // ================================ phase 3: emission, BSDF sampling ================================
        bool pathContinues = live;
        uint32_t isTransmission = 0u;
        float3 wi(0.0f);
        float4 newPrevVertexTechPdf(0.0f), newPrevBsdfTechPdf(0.0f);
        if (live) {
            // ---- emission seen by BSDF sampling (integrator.slang:79-85; hero: integrator.slang:98-126) -------------------
            if (!neeOnly) {
                if (anyGreater(emission, 0.0f)) {
                    const bool misActive = (fp.modeFlags & MODE_NEE_ENABLED) && (flags & PF_PREV_VERTEX_NEE_ALLOWED) && depth > 0u;
                    if (MODE == MODE_HERO && heroActive) {
                        float misWeight = heroWavelengthBalanceWeight(techPdf);
                        if (misActive) {
                            const float lp2 = lightPdfAreaToSolidAngle(lightPdfArea, geometricNormal, ray.direction, hitT) * (ENVIS ? 1.0f - sc.env.pEnv : 1.0f);
                            const float4 pv = fromF4(SQ_LD(S.prevVertexTechPdf + i)), pb = fromF4(SQ_LD(S.prevBsdfTechPdf + i));
                            misWeight = computeSpectralMISWeight(pv * pb, pv * lp2);
                        }
                        const float4 c = thr4 * (spectralScalarFromLinearSrgb4(T, spectralMemoOf(T, state, SPECTRAL_MEMO_EMISSION), emission, wl4) * misWeight);
                        ::float4 r = fp.rec.radiance[rec];
                        r.x += c.x; r.y += c.y; r.z += c.z; r.w += c.w;
                        fp.rec.radiance[rec] = r;
                    } else {
                        float misWeight = 1.0f;
                        if (misActive && prevBsdfPdf > 0.0f) {
                            const float lp2 = lightPdfAreaToSolidAngle(lightPdfArea, geometricNormal, ray.direction, hitT) * (ENVIS ? 1.0f - sc.env.pEnv : 1.0f);
                            misWeight = lp2 <= 0.0f ? 1.0f : powerHeuristic(prevBsdfPdf, lp2);
                        }
                        if (MODE == MODE_RGB) {
                            const float3 c = thrRgb * (emission * misWeight);
                            ::float4 r = fp.rec.radiance[rec];
                            r.x += c.x; r.y += c.y; r.z += c.z;
                            fp.rec.radiance[rec] = r;
                        } else {
                            const float c = thrScalar * (spectralScalarFromLinearSrgb(T, spectralMemoOf(T, state, SPECTRAL_MEMO_EMISSION), emission, lambdaScalar) * misWeight);
                            if (MODE == MODE_SINGLE) fp.rec.radiance[rec].x += c;
                            else fp.rec.radianceScalar[rec] += c;
                        }
                    }
                }
            }
            if ((fp.modeFlags & MODE_BOUNCE_COUNT) && sampleIndex == 0u) fp.film.bounceCount[lp] = depth + 1u;
        }
        SHADE_SUBPHASE_BARRIER();
        if (live) {
            // ---- sample the next direction (sample*NextDirection) -------------------------------------------------------
            if (MODE == MODE_HERO && heroActive) {
                const SpectralBSDFSample smp = sampleSpectralBSDF(T, state, basis, wl4, rng);
                if (!smp.isUsable()) pathContinues = false;
                if (pathContinues) {
                    newPrevVertexTechPdf = techPdf;
                    newPrevBsdfTechPdf = smp.techniquePdf;
                    techPdf *= smp.techniquePdf;
                    thr4 *= smp.weight;
                    if (!anyGreater(thr4, 0.0f)) pathContinues = false;
                }
                if (pathContinues) {
                    if (smp.isTransmission != 0u && materialMediumIsRefractive(state.material) && state.material.abbeNumber > 0.0f) {
                        // dispersive collapse (spectral_hero/transport.slang:77-87). The 4-lane radiance stays in the record and
                        // is converted to XYZ by the film kernel together with the scalar lane.
                        thrScalar = thr4.x;
                        thr4 = float4(0.0f);
                        prevBsdfPdf = smp.techniquePdf.x;
                        heroActive = false;
                        flags &= ~PF_HERO_ACTIVE;
                    }
                    isTransmission = smp.isTransmission;
                    wi = smp.wi;
                }
            } else {
                const BSDFSample smp = sampleBSDF(T, state, basis, rng);
                if (!smp.isUsable()) pathContinues = false;
                if (pathContinues) {
                    if (MODE == MODE_RGB) {
                        thrRgb *= smp.weight;
                        if (!anyGreater(thrRgb, 0.0f)) pathContinues = false;
                    } else {
                        thrScalar *= smp.weight.x;
                        if (thrScalar <= 0.0f) pathContinues = false;
                    }
                    prevBsdfPdf = smp.pdf;
                    isTransmission = smp.isTransmission;
                    wi = smp.wi;
                }
            }
        }
        SHADE_PHASE_BARRIER();

        // ================================ phase 4: medium update, Russian roulette, queue writes ================================
        if (live) {
            if (pathContinues) {
                // medium update + NEE bookkeeping (integrator.slang:96-98)
                if (MODE == MODE_RGB) updateMediumStateFromTransmission(T, state.material, state.frontFace, isTransmission, 0.0f, 0u, medium);
                else if (MODE == MODE_HERO && heroActive) updateMediumStateFromTransmissionSpectral(T, state.material, state.frontFace, isTransmission, wl4, medium);
                else updateMediumStateFromTransmission(T, state.material, state.frontFace, isTransmission, lambdaScalar, 1u, medium);
                const bool sampledPathNeeAllowed = currentVertexNeeAllowed && isTransmission == 0u;  // shadow kernel clears it on "unsupported"
                flags = (flags & ~(PF_PREV_VERTEX_NEE_ALLOWED | PF_MEDIUM_REFRACTIVE | PF_MEDIUM_ABSORPTION)) |
                        (sampledPathNeeAllowed ? PF_PREV_VERTEX_NEE_ALLOWED : 0u) | (medium.flags << 1);
                // Russian roulette (integrator.slang:100-106)
                if (depth + 1u >= scene.rrMinDepth) {
                    float cp;
                    if (MODE == MODE_RGB) cp = clamp(maxComponent(thrRgb), RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB);
                    else if (MODE == MODE_HERO && heroActive) cp = clamp(maxComponent4(thr4), RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB);
                    else cp = clamp(thrScalar, RR_MIN_CONTINUE_PROB, RR_MAX_CONTINUE_PROB);
                    if (rand(rng) > cp) {
                        pathContinues = false;
                    } else if (MODE == MODE_RGB) {
                        thrRgb /= cp;
                    } else if (MODE == MODE_HERO && heroActive) {
                        techPdf *= float4(cp);
                        thr4 /= cp;
                    } else {
                        thrScalar /= cp;
                    }
                }
            }
            if (depth + 1u >= scene.rrMaxDepth) pathContinues = false;

            // ---- write the continuing path at its new (compacted) position ------------------------------------------------
            uint32_t newPos = 0x7fffffffu;
            if (pathContinues) {
                newPos = allocSlots(fp.extCount + depth + 1u);
                const float3 gn = geometricNormal;
                const float3 off = isTransmission != 0u ? -gn : gn;
                SQ_ST(N.rayO + newPos, toF4(hitPoint + off * SHADOW_ORIGIN_OFFSET, RAY_T_MIN));
                SQ_ST(N.rayD + newPos, toF4(wi, RAY_T_MAX));
                SQ_ST(N.meta + newPos, make_uint4(rec, rng, flags, 0u));
                if (MODE == MODE_RGB) {
                    SQ_ST(N.thr + newPos, toF4(thrRgb, prevBsdfPdf));
                    if (medium.absorptionActive()) SQ_ST(N.sigma + newPos, toF4(medium.absorptionSigma, 0.0f));
                } else if (MODE == MODE_SINGLE) {
                    SQ_ST(N.thr + newPos, make_float4(thrScalar, lambdaScalar, prevBsdfPdf, 0.0f));
                    if (medium.absorptionActive()) SQ_ST(N.sigma + newPos, toF4(medium.absorptionSigma, 0.0f));
                } else {
                    SQ_ST(N.thr + newPos, toF4(thr4));
                    SQ_ST(N.heroMisc + newPos, make_float4(unit, thrScalar, prevBsdfPdf, 0.0f));
                    SQ_ST(N.techPdf + newPos, toF4(techPdf));
                    SQ_ST(N.prevVertexTechPdf + newPos, toF4(newPrevVertexTechPdf));
                    SQ_ST(N.prevBsdfTechPdf + newPos, toF4(newPrevBsdfTechPdf));
                    if (medium.absorptionActive())
                        SQ_ST(N.sigma + newPos, heroActive ? toF4(medium.spectralAbsorptionSigma) : toF4(medium.absorptionSigma, 0.0f));
                }
            }
            if (shadowPending) {
                const uint32_t k = allocSlots(fp.shCount + depth);
                SQ_ST(fp.shO + k, shadowO);
                SQ_ST(fp.shD + k, shadowD);
                SQ_ST(fp.shContribution + k, shadowC);
                SQ_ST(fp.shTarget + k, make_uint2(rec, newPos | (shadowScalarLane ? SHADOW_KIND_SCALAR : 0u)));
                SQ_ST(fp.shSeed + k, shadowSeed);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// Material sort of one depth's queue: a 256-bin counting sort on key = 0 (miss) | 1 + materialIndex. Three small kernels
// (count, scan, scatter); block-local histograms in shared memory keep the global atomics to one per bin per block.
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t shadeSortKey(const FrameParams& fp, uint32_t i) {
    const uint32_t inst = fp.hitA[i].x;
    if (inst == VKRT_INVALID_INDEX) return 0u;
    const uint32_t m = __ldg(&fp.scene.meshInfos[inst].materialIndex) + 1u;
    return m > 255u ? 255u : m;
}
__global__ void __launch_bounds__(256) k_sort_count(const FrameParams fp, const uint32_t depth) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t count = fp.extCount[depth];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) atomicAdd(&hist[shadeSortKey(fp, i)], 1u);
    __syncthreads();
    if (hist[threadIdx.x]) atomicAdd(&fp.sortBins[threadIdx.x], hist[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_sort_scan(const FrameParams fp) {
    __shared__ uint32_t s[256];
    const uint32_t c = fp.sortBins[threadIdx.x];
    s[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        uint32_t v = threadIdx.x >= (uint32_t)o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += v;
        __syncthreads();
    }
    fp.sortBins[threadIdx.x] = s[threadIdx.x] - c;  // exclusive start
    fp.sortBins[256 + threadIdx.x] = 0u;            // cursor
}
__global__ void __launch_bounds__(256) k_sort_scatter(const FrameParams fp, const uint32_t depth) {
    __shared__ uint32_t hist[256], base[256];
    const uint32_t count = fp.extCount[depth];
    // every block handles contiguous chunks of 256 * ITEMS entries so that one reservation per bin covers a whole chunk
    constexpr uint32_t ITEMS = 8;
    for (uint32_t chunk = blockIdx.x * 256u * ITEMS; chunk < count; chunk += gridDim.x * 256u * ITEMS) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t key[ITEMS], rank[ITEMS];
#pragma unroll
        for (uint32_t k = 0; k < ITEMS; k++) {
            const uint32_t i = chunk + k * 256u + threadIdx.x;
            key[k] = i < count ? shadeSortKey(fp, i) : 0xffffffffu;
            if (i < count) rank[k] = atomicAdd(&hist[key[k]], 1u);
        }
        __syncthreads();
        if (hist[threadIdx.x]) base[threadIdx.x] = fp.sortBins[threadIdx.x] + atomicAdd(&fp.sortBins[256 + threadIdx.x], hist[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < ITEMS; k++) {
            const uint32_t i = chunk + k * 256u + threadIdx.x;
            if (i < count) fp.shadeOrder[base[key[k]] + rank[k]] = i;
        }
        __syncthreads();
    }
}
void launchShadeSort(const FrameParams& fp, uint32_t depth, int smCount, cudaStream_t st) {
    cudaMemsetAsync(fp.sortBins, 0, sizeof(uint32_t) * 512, st);
    k_sort_count<<<smCount * 8, 256, 0, st>>>(fp, depth);
    k_sort_scan<<<1, 256, 0, st>>>(fp);
    k_sort_scatter<<<smCount * 8, 256, 0, st>>>(fp, depth);
}

// ----------------------------------------------------------------------------------------------------------------------
// film (integrator/path/writeback.slang:9-123, film/tonemap.slang)
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 blendAccumulatedValue(float3 prev, float prevW, float3 cur, float curW) {
    const float total = prevW + curW;
    if (total <= 0.0f) return float3(0.0f);
    return (prev * prevW + cur * curW) / total;
}
__device__ __forceinline__ ::uint2 packHalf4(float4 v) {
    return make_uint2((uint32_t)f32_to_f16(v.x) | ((uint32_t)f32_to_f16(v.y) << 16), (uint32_t)f32_to_f16(v.z) | ((uint32_t)f32_to_f16(v.w) << 16));
}
__device__ __forceinline__ float4 unpackHalf4(::uint2 p) {
    return float4(f16_to_f32((uint16_t)(p.x & 0xffffu)), f16_to_f32((uint16_t)(p.x >> 16)), f16_to_f32((uint16_t)(p.y & 0xffffu)),
                  f16_to_f32((uint16_t)(p.y >> 16)));
}
__device__ __forceinline__ uint32_t unorm16(float v) { return (uint32_t)__float2int_rn(saturate(v) * 65535.0f); }
__device__ __forceinline__ ::uint2 packUnorm16x4(float4 v) {
    return make_uint2(unorm16(v.x) | (unorm16(v.y) << 16), unorm16(v.z) | (unorm16(v.w) << 16));
}

template <int MODE>
__global__ void __launch_bounds__(256) k_film(const FrameParams fp, const int firstChunk, const int lastChunk) {
    const uint32_t lpc = fp.tiles.localPixelCount;
    const SceneData& scene = fp.sd;
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < lpc; lp += gridDim.x * blockDim.x) {
        uint32_t gx, gy;
        const bool valid = localPixelToGlobal(fp.tiles, fp.tiles.localToGlobalTile, lp, gx, gy);
        const int w = 1 - fp.readIndex;
        if (!valid || !insideViewport(scene, (int)gx, (int)gy)) {  // raygen_rgb.slang:6-12
            if (lastChunk) {
                fp.film.accum[w][lp] = make_float4(0.f, 0.f, 0.f, 0.f);
                fp.film.albedo[w][lp] = make_uint2(0u, 0u);
                fp.film.normal[w][lp] = make_uint2(0u, 0u);
                fp.film.output[lp] = make_uint2(0u, 0u);
            }
            continue;
        }
        // sum this chunk's samples in sample order
        float3 radiance(0.0f), albedo(0.0f), normal(0.0f);
        float weight = 0.0f, fdepth = 0.0f, follow = 0.0f;
        if (!firstChunk) {
            const ::float4 a = fp.film.frameRadiance[lp], b = fp.film.frameFeatA[lp], c = fp.film.frameFeatB[lp];
            radiance = float3(a.x, a.y, a.z);
            albedo = float3(b.x, b.y, b.z); weight = b.w;
            normal = float3(c.x, c.y, c.z); fdepth = c.w;
            follow = fp.film.frameFollow[lp];
        }
        for (uint32_t s = 0; s < fp.chunkSamples; s++) {
            const uint32_t rec = s * lpc + lp;
            const ::float4 r = fp.rec.radiance[rec];
            if (MODE == MODE_RGB) {
                radiance += float3(r.x, r.y, r.z);
            } else if (MODE == MODE_SINGLE) {
                WavelengthSample ws;
                ws.lambdaNm = WAVELENGTH_MIN_NM + saturate(fp.rec.unitWavelength[rec]) * WAVELENGTH_RANGE_NM;
                ws.invPdf = WAVELENGTH_RANGE_NM;
                radiance += spectralSampleToXYZ(r.x, ws);
            } else {
                const float4 wl = heroWavelengths(fp.rec.unitWavelength[rec]);
                float3 xyz = spectralSample4ToXYZ(float4(r.x, r.y, r.z, r.w), wl, float4(WAVELENGTH_RANGE_NM));
                WavelengthSample ws;
                ws.lambdaNm = wl.x;
                ws.invPdf = WAVELENGTH_RANGE_NM;
                xyz += spectralSampleToXYZ(fp.rec.radianceScalar[rec], ws);
                radiance += xyz;
            }
            const ::float4 fa = fp.rec.featA[rec], fb = fp.rec.featB[rec];
            albedo += float3(fa.x, fa.y, fa.z);
            normal += float3(fb.x, fb.y, fb.z);
            weight += fa.w;
            fdepth += fb.w;
            follow += fp.rec.follow[rec];
        }
        if (!lastChunk) {
            fp.film.frameRadiance[lp] = toF4(radiance, 0.0f);
            fp.film.frameFeatA[lp] = toF4(albedo, weight);
            fp.film.frameFeatB[lp] = toF4(normal, fdepth);
            fp.film.frameFollow[lp] = follow;
            continue;
        }
        const uint32_t spp = max(scene.samplesPerPixel, 1u);
        const ::float4 prevAcc = fp.film.accum[fp.readIndex][lp];
        const uint32_t previousSamples = (uint32_t)(prevAcc.w + 0.5f);
        bool debugOut = false;
        float3 debugRadiance(0.0f);
        if (scene.debugMode != VKRT_DEBUG_MODE_NONE) {
            const ::float4 dc = fp.film.debugColor[lp];
            if (scene.debugMode == VKRT_DEBUG_MODE_SELECTION_MASK) {
                debugOut = true;
                debugRadiance = float3(0.05f);
            } else if (dc.w > 0.0f) {
                debugOut = true;
                debugRadiance = float3(dc.x, dc.y, dc.z);
            } else if (fp.modeFlags & MODE_BOUNCE_COUNT) {
                const float t = clamp(float(fp.film.bounceCount[lp]) / max(float(scene.rrMaxDepth), 1.0f), 0.0f, 1.0f);
                debugOut = true;
                debugRadiance = float3(t, 1.0f - t, 0.0f);
            }
        }
        if (!debugOut) {  // finalizeFrameAccumulation + applyFrameDebugOverrides
            radiance /= float(spp);
            if (weight > 0.0f) { albedo /= weight; normal /= weight; }
            else { albedo = float3(0.0f); normal = float3(0.0f); }
            if (fp.modeFlags & MODE_DN_ALBEDO) { debugOut = true; debugRadiance = albedo; }
            else if (fp.modeFlags & MODE_DN_NORMAL) { debugOut = true; debugRadiance = weight > 0.0f ? normal * 0.5f + 0.5f : float3(0.0f); }
            else if (fp.modeFlags & MODE_DN_VALIDITY) { debugOut = true; debugRadiance = float3(saturate(weight / float(spp))); }
            else if (fp.modeFlags & MODE_DN_DEPTH) {
                float nd = 0.0f;
                if (weight > 0.0f) nd = saturate((fdepth / weight - 1.0f) / max(float(scene.rrMaxDepth - 1u), 1.0f));
                debugOut = true;
                debugRadiance = float3(nd);
            } else if (fp.modeFlags & MODE_DN_FOLLOW) {
                debugOut = true;
                debugRadiance = lerp(float3(0.05f), float3(1.0f, 0.6f, 0.0f), saturate(follow / float(spp)));
            }
        }
        if (debugOut) {  // writeDebugFrameOutputs
            fp.film.accum[w][lp] = toF4(debugRadiance, 0.0f);
            fp.film.albedo[w][lp] = make_uint2(0u, 0u);
            fp.film.normal[w][lp] = make_uint2(0u, 0u);
            fp.film.output[lp] = packUnorm16x4(float4(encodeDisplayColor(debugRadiance), 1.0f));
            continue;
        }
        // writeAccumulatedFrameOutputs
        const float4 prevAlbedo = unpackHalf4(fp.film.albedo[fp.readIndex][lp]);
        const float4 prevNormal = unpackHalf4(fp.film.normal[fp.readIndex][lp]);
        const float previousWeight = float(previousSamples);
        const float totalWeight = previousWeight + float(spp);
        const float3 accumulated = blendAccumulatedValue(float3(prevAcc.x, prevAcc.y, prevAcc.z), previousWeight, radiance, float(spp));
        const float3 accAlbedo = blendAccumulatedValue(prevAlbedo.xyz(), prevAlbedo.w, albedo, weight);
        const float3 accNormal = blendAccumulatedValue(prevNormal.xyz(), prevNormal.w, normal, weight);
        fp.film.accum[w][lp] = toF4(accumulated, totalWeight);
        fp.film.albedo[w][lp] = packHalf4(float4(accAlbedo, prevAlbedo.w + weight));
        fp.film.normal[w][lp] = packHalf4(float4(accNormal, prevNormal.w + weight));
        const float3 display = MODE == MODE_RGB ? mapSceneColorToDisplay(scene, accumulated) : mapSceneColorToDisplay(scene, xyzToLinearSrgb(accumulated));
        fp.film.output[lp] = packUnorm16x4(float4(display, 1.0f));
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// primary visibility AOVs (HITID_CENTER: captureSelectionHit, debug.slang:21-29; HITID_S0: frame 0, sample 0)
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_primary_raygen(const FrameParams fp, const int jittered) {
    const uint32_t lpc = fp.tiles.localPixelCount;
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < lpc; lp += gridDim.x * blockDim.x) {
        fp.film.hitId[lp] = make_uint2(VKRT_INVALID_INDEX, VKRT_INVALID_INDEX);
        fp.film.hitTuv[lp] = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t gx, gy;
        if (!localPixelToGlobal(fp.tiles, fp.tiles.localToGlobalTile, lp, gx, gy)) continue;
        if (!insideViewport(fp.sd, (int)gx, (int)gy)) continue;
        float2 jitter(0.0f);
        uint32_t rng = 0u;
        if (jittered) {
            rng = initPixelSeed((int)gx, (int)gy, 0u, 0u);
            const float jx = rand(rng);
            const float jy = rand(rng);
            jitter = float2(__fsub_rn(jx, 0.5f), __fsub_rn(jy, 0.5f));
        }
        const Ray ray = makePrimaryRayExact(fp.sd, (int)gx, (int)gy, jitter);
        const uint32_t j = allocSlots(fp.extCount);
        fp.st[0].rayO[j] = toF4(ray.origin, ray.tMin);
        fp.st[0].rayD[j] = toF4(ray.direction, ray.tMax);
        fp.st[0].meta[j] = make_uint4(lp, rng, 0u, 0u);
    }
}
__global__ void __launch_bounds__(256) k_primary_store(const FrameParams fp) {
    const uint32_t count = fp.extCount[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const ::uint4 h = fp.hitA[i];
        const ::uint4 mt = fp.st[0].meta[i];
        const uint32_t lp = mt.x;
        fp.film.hitId[lp] = make_uint2(h.x, h.y);
        fp.film.hitTuv[lp] = make_float4(__uint_as_float(h.z), __uint_as_float(h.w), __uint_as_float(mt.w), 0.0f);
    }
}

// Tile-compact -> full-frame row-major (read_aov, import_gathered). elemSize in 4-byte words per pixel.
__global__ void __launch_bounds__(256) k_untile(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, TileMap tm, const uint32_t* __restrict__ l2g,
                                                uint32_t words, uint32_t srcStridePixels) {
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < tm.localPixelCount; lp += gridDim.x * blockDim.x) {
        uint32_t gx, gy;
        if (!localPixelToGlobal(tm, l2g, lp, gx, gy)) continue;
        const size_t d = ((size_t)gy * tm.width + gx) * words;
        const size_t s = ((size_t)lp) * words;
        (void)srcStridePixels;
        for (uint32_t k = 0; k < words; k++) dst[d + k] = src[s + k];
    }
}

// Rank 0 after the NCCL gather: every rank's tile-compact buffer (rank r at r * strideWords) -> the full-frame row-major image, in ONE
// launch. One thread per full-frame pixel: its tile's owner is (ty * tilesX + tx + ty) mod G (tiles.h) and tileLocalIndex[tile] is the
// position of that tile among the owner's tiles (uploaded once per resize), so writes are coalesced and reads are 32-pixel runs.
__global__ void __launch_bounds__(256) k_untile_all(const uint32_t* __restrict__ gathered, uint32_t* __restrict__ dst, TileMap tm, uint32_t worldSize,
                                                    const uint32_t* __restrict__ tileLocalIndex, uint64_t rankStrideWords, uint32_t words) {
    const uint64_t pixels = (uint64_t)tm.width * tm.height;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t gy = (uint32_t)(p / tm.width), gx = (uint32_t)(p - (uint64_t)gy * tm.width);
        const uint32_t tx = gx / tm.tileW, ty = gy / tm.tileH;
        const uint32_t tile = ty * tm.tilesX + tx;
        const uint32_t owner = (uint32_t)(((uint64_t)ty * tm.tilesX + tx + ty) % worldSize);
        const uint64_t lp = (uint64_t)tileLocalIndex[tile] * tm.tileW * tm.tileH + (gy % tm.tileH) * tm.tileW + gx % tm.tileW;
        const uint32_t* src = gathered + owner * rankStrideWords + lp * words;
        uint32_t* d = dst + p * words;
        if (words == 4u) *reinterpret_cast<::uint4*>(d) = *reinterpret_cast<const ::uint4*>(src);
        else *reinterpret_cast<::uint2*>(d) = *reinterpret_cast<const ::uint2*>(src);
    }
}
void launchUntileAll(const void* gathered, void* dst, const TileMap& tm, uint32_t worldSize, const uint32_t* tileLocalIndex, uint64_t rankStrideWords,
                     uint32_t words, int grid, cudaStream_t st) {
    k_untile_all<<<grid, 256, 0, st>>>((const uint32_t*)gathered, (uint32_t*)dst, tm, worldSize, tileLocalIndex, rankStrideWords, words);
}

// Accumulation texels at a few global pixel positions (auto-exposure probe): global pixel -> owning local tile by binary search in the
// ascending local-to-global tile list.
__global__ void __launch_bounds__(256) k_probe_accum(const ::float4* __restrict__ accum, TileMap tm, const uint32_t* __restrict__ l2g, const ::uint2* __restrict__ xy,
                                                     uint32_t count, ::float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const ::uint2 p = xy[i];
    ::float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.x < tm.width && p.y < tm.height && tm.localTileCount) {
        const uint32_t gt = (p.y / tm.tileH) * tm.tilesX + p.x / tm.tileW;
        uint32_t lo = 0, hi = tm.localTileCount;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (l2g[mid] <= gt) lo = mid; else hi = mid;
        }
        if (l2g[lo] == gt) v = accum[(size_t)lo * tm.tileW * tm.tileH + (p.y % tm.tileH) * tm.tileW + p.x % tm.tileW];
    }
    out[i] = v;
}
void launchProbeAccum(const ::float4* accum, const TileMap& tm, const uint32_t* l2g, const ::uint2* xy, uint32_t count, ::float4* out, cudaStream_t st) {
    k_probe_accum<<<(count + 255) / 256, 256, 0, st>>>(accum, tm, l2g, xy, count, out);
}

// ----------------------------------------------------------------------------------------------------------------------
// Per-call closure evaluation (TEST ENTRY, include/vkrt_closure.h): one thread per query through the SAME device functions, compiled
// in this translation unit with the same flags, that k_shade calls: BSDFMaterial / BSDFState construction, evalBSDF /
// evalSingleWavelengthBSDF / evalSpectralBSDF (the NEE path) and sampleBSDF / sampleSpectralBSDF (the continuation), with the rgb2spec
// scale axis staged in shared memory as in k_shade. Lets every lobe be compared call by call with the reference's own shaders.
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_eval_closures(const SceneView sc, const vkrt_closure_query* __restrict__ queries, uint32_t count,
                                                       vkrt_closure_result* __restrict__ results) {
    __shared__ float sScale[RGB2SPEC_SMEM_RES];
    SpectralTables T = sc.spectral;
    if (T.info.res <= RGB2SPEC_SMEM_RES && T.info.res > 0u) {
        if (threadIdx.x < T.info.res) sScale[threadIdx.x] = T.scale[threadIdx.x];
        T.scale = sScale;
        __syncthreads();
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const vkrt_closure_query& Q = queries[i];
    vkrt_closure_result R = {};
    const Material material = loadMaterial(&Q.material);
    const BSDFMaterial bm = makeBSDFMaterial(material);
    const float3 wo(Q.wo[0], Q.wo[1], Q.wo[2]), wi(Q.wi[0], Q.wi[1], Q.wi[2]);
    const float4 wl(Q.wavelengths[0], Q.wavelengths[1], Q.wavelengths[2], Q.wavelengths[3]);
    const BSDFState st(bm, wo, Q.frontFace, Q.mode == 0u ? 0.0f : wl.x, Q.mode == 0u ? 0u : 1u);
    ShadingBasis basis;
    basis.tangent = float3(1.0f, 0.0f, 0.0f);
    basis.bitangent = float3(0.0f, 1.0f, 0.0f);
    basis.normal = float3(0.0f, 0.0f, 1.0f);
    uint rng = Q.rng;
    if (Q.mode == 2u) {
        float4 tp(0.0f);
        const float4 v = evalSpectralBSDF(T, st, wi, wl, tp);
        R.evalValue[0] = v.x; R.evalValue[1] = v.y; R.evalValue[2] = v.z; R.evalValue[3] = v.w;
        R.evalPdf[0] = tp.x; R.evalPdf[1] = tp.y; R.evalPdf[2] = tp.z; R.evalPdf[3] = tp.w;
        const SpectralBSDFSample s = sampleSpectralBSDF(T, st, basis, wl, rng);
        R.sampleWi[0] = s.wi.x; R.sampleWi[1] = s.wi.y; R.sampleWi[2] = s.wi.z;
        R.sampleWeight[0] = s.weight.x; R.sampleWeight[1] = s.weight.y; R.sampleWeight[2] = s.weight.z; R.sampleWeight[3] = s.weight.w;
        R.samplePdf[0] = s.techniquePdf.x; R.samplePdf[1] = s.techniquePdf.y; R.samplePdf[2] = s.techniquePdf.z; R.samplePdf[3] = s.techniquePdf.w;
        R.sampleFlags = (s.isUsable() ? 1u : 0u) | (s.isTransmission != 0u ? 2u : 0u);
    } else {
        const BSDFEval e = Q.mode == 0u ? evalBSDF(T, st, wi) : evalSingleWavelengthBSDF(T, st, wi);
        R.evalValue[0] = e.value.x; R.evalValue[1] = e.value.y; R.evalValue[2] = e.value.z;
        R.evalPdf[0] = e.pdf;
        const BSDFSample s = sampleBSDF(T, st, basis, rng);
        R.sampleWi[0] = s.wi.x; R.sampleWi[1] = s.wi.y; R.sampleWi[2] = s.wi.z;
        R.sampleWeight[0] = s.weight.x; R.sampleWeight[1] = s.weight.y; R.sampleWeight[2] = s.weight.z;
        R.samplePdf[0] = s.pdf;
        R.sampleFlags = (s.isUsable() ? 1u : 0u) | (s.isTransmission != 0u ? 2u : 0u);
    }
    R.rngAfter = rng;
    results[i] = R;
}
void launchEvalClosures(const SceneView& sc, const vkrt_closure_query* queries, uint32_t count, vkrt_closure_result* results, cudaStream_t st) {
    if (count) k_eval_closures<<<(count + 127u) / 128u, 128, 0, st>>>(sc, queries, count, results);
}

// One thread per mesh: the rotation trig the shade kernel would otherwise re-evaluate per hit (see MeshTrig).
__global__ void __launch_bounds__(128) k_mesh_trig(const MeshInfo* __restrict__ infos, MeshTrig* __restrict__ out, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = makeMeshTrig(float3(infos[i].rotation[0], infos[i].rotation[1], infos[i].rotation[2]));
}
// rgb2spec payload -> float4 cells (shading.cuh SpectralTables)
// Spectral memo (shading.cuh, SpectralMemoEntry): the table lookups of the colours that are constants of a material (diffuse colour, emission;
// both as k_shade derives them when no texture or vertex colour intervenes) and of an emissive mesh. Same device functions, same arithmetic
// as the per-vertex path; an entry whose key does not match at run time is simply not used.
__global__ void __launch_bounds__(128) k_spectral_memo(const SceneView sc, uint32_t materialCount, uint32_t emissiveMeshCount, SpectralMemoEntry* __restrict__ materialMemo,
                                                       SpectralMemoEntry* __restrict__ emissiveMemo) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const SpectralTables& T = sc.spectral;
    if (i < materialCount) {
        Material m = loadMaterial(sc.materials + i);
        m.roughness = saturate(m.roughness * 1.0f);   // applySurfaceTextures with unit texels and unit vertex colour
        m.metallic = saturate(m.metallic * 1.0f);
        normalizeEmission(m, float3(1.0f));
        const float3 emission = float3(m.emissionColor[0], m.emissionColor[1], m.emissionColor[2]) * m.emissionLuminance;
        const BSDFMaterial bm = makeBSDFMaterial(m);
        const float3 diffuse = saturate(bsdfDiffuseColor(bm));
        materialMemo[i * SPECTRAL_MEMO_SLOTS + SPECTRAL_MEMO_DIFFUSE] = {make_float4(diffuse.x, diffuse.y, diffuse.z, 0.0f), spectralCoefficientsFromLinearSrgb(T, diffuse)};
        materialMemo[i * SPECTRAL_MEMO_SLOTS + SPECTRAL_MEMO_EMISSION] = {make_float4(emission.x, emission.y, emission.z, 0.0f), spectralCoefficientsFromLinearSrgb(T, emission)};
    } else if (i - materialCount < emissiveMeshCount) {
        const uint32_t k = i - materialCount;
        const EmissiveMesh em = sc.emissiveMeshes[k];
        const float3 emission(em.emission[0], em.emission[1], em.emission[2]);
        emissiveMemo[k] = {make_float4(emission.x, emission.y, emission.z, 0.0f), spectralCoefficientsFromLinearSrgb(T, emission)};
    }
}
void launchSpectralMemo(const SceneView& sc, uint32_t materialCount, uint32_t emissiveMeshCount, SpectralMemoEntry* materialMemo, SpectralMemoEntry* emissiveMemo, cudaStream_t st) {
    const uint32_t n = materialCount + emissiveMeshCount;
    if (n) k_spectral_memo<<<(n + 127) / 128, 128, 0, st>>>(sc, materialCount, emissiveMeshCount, materialMemo, emissiveMemo);
}

__global__ void __launch_bounds__(256) k_pack_rgb2spec(const float* __restrict__ table, uint32_t dataOffset, size_t cellCount, ::float4* __restrict__ cells) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cellCount; i += (size_t)gridDim.x * blockDim.x) {
        const float* c = table + dataOffset + 3u * i;
        cells[i] = make_float4(c[0], c[1], c[2], 0.0f);
    }
}
void launchPackRgb2spec(const float* table, uint32_t dataOffset, size_t cellCount, ::float4* cells, int grid, cudaStream_t st) {
    k_pack_rgb2spec<<<grid, 256, 0, st>>>(table, dataOffset, cellCount, cells);
}
void launchMeshTrig(const MeshInfo* infos, MeshTrig* out, uint32_t count, cudaStream_t st) {
    if (count) k_mesh_trig<<<(count + 127) / 128, 128, 0, st>>>(infos, out, count);
}

int traceBlocksPerSm(bool count) {
    int n = 0;
    // the two-level variant has the larger register footprint; one grid size serves both
    if (count) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<true, false, TRACE_STACK>, TRACE_BLOCK, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<false, false, TRACE_STACK>, TRACE_BLOCK, 0);
    return n > 0 ? n : 1;
}
int shadeBlocksPerSm(int mode) {
    int n = 0;
    if (SHADE_DYNAMIC_SMEM > 0) {   // more than the default 48 KB of dynamic shared memory needs an opt-in per kernel
        const int bytes = (int)SHADE_DYNAMIC_SMEM;
        cudaFuncSetAttribute(k_shade<MODE_RGB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaFuncSetAttribute(k_shade<MODE_SINGLE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaFuncSetAttribute(k_shade<MODE_HERO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaFuncSetAttribute(k_shade<MODE_RGB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaFuncSetAttribute(k_shade<MODE_SINGLE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaFuncSetAttribute(k_shade<MODE_HERO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    }
    if (mode == MODE_RGB) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_shade<MODE_RGB, false>, SHADE_BLOCK, SHADE_DYNAMIC_SMEM);
    else if (mode == MODE_SINGLE) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_shade<MODE_SINGLE, false>, SHADE_BLOCK, SHADE_DYNAMIC_SMEM);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_shade<MODE_HERO, false>, SHADE_BLOCK, SHADE_DYNAMIC_SMEM);
    return n > 0 ? n : 1;
}

// ---- host-side launch helpers (called from api.cu) ---------------------------------------------------------------------
void launchRaygen(int mode, const FrameParams& fp, int grid, cudaStream_t st) {
    if (mode == MODE_RGB) k_raygen<MODE_RGB><<<grid, 256, 0, st>>>(fp);
    else if (mode == MODE_SINGLE) k_raygen<MODE_SINGLE><<<grid, 256, 0, st>>>(fp);
    else k_raygen<MODE_HERO><<<grid, 256, 0, st>>>(fp);
}
void launchShade(int mode, const FrameParams& fp, uint32_t depth, int grid, cudaStream_t st) {
    if (fp.scene.env.active) {  // environment-map importance sampling: separate instantiations, the default kernels are untouched
        if (mode == MODE_RGB) k_shade<MODE_RGB, true><<<grid, SHADE_BLOCK, SHADE_DYNAMIC_SMEM, st>>>(fp, depth);
        else if (mode == MODE_SINGLE) k_shade<MODE_SINGLE, true><<<grid, SHADE_BLOCK, SHADE_DYNAMIC_SMEM, st>>>(fp, depth);
        else k_shade<MODE_HERO, true><<<grid, SHADE_BLOCK, SHADE_DYNAMIC_SMEM, st>>>(fp, depth);
        return;
    }
    if (mode == MODE_RGB) k_shade<MODE_RGB, false><<<grid, SHADE_BLOCK, SHADE_DYNAMIC_SMEM, st>>>(fp, depth);
    else if (mode == MODE_SINGLE) k_shade<MODE_SINGLE, false><<<grid, SHADE_BLOCK, SHADE_DYNAMIC_SMEM, st>>>(fp, depth);
    else k_shade<MODE_HERO, false><<<grid, SHADE_BLOCK, SHADE_DYNAMIC_SMEM, st>>>(fp, depth);
}
// weight of every environment texel for the importance-sampling table: luminance x sin(theta of the texel row)
__global__ void k_env_weights(const SceneView sc, uint32_t textureIndex, float* __restrict__ out) {
    const TextureView t = sc.textures[textureIndex];
    const uint32_t n = t.width * t.height;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t y = i / t.width, x = i - y * t.width;
        const float4 c = fetchTexel(sc, t, (int)x, (int)y);
        float w = linearSrgbLuminance(float3(c.x, c.y, c.z)) * sinf(PI * (float(y) + 0.5f) / float(t.height));
        if (!(w > 0.0f) || !(w < 3.0e38f)) w = 0.0f;
        out[i] = w;
    }
}
void launchEnvWeights(const SceneView& sc, uint32_t textureIndex, float* out, int grid, cudaStream_t st) { k_env_weights<<<grid, 256, 0, st>>>(sc, textureIndex, out); }
void launchFilm(int mode, const FrameParams& fp, int firstChunk, int lastChunk, int grid, cudaStream_t st) {
    if (mode == MODE_RGB) k_film<MODE_RGB><<<grid, 256, 0, st>>>(fp, firstChunk, lastChunk);
    else if (mode == MODE_SINGLE) k_film<MODE_SINGLE><<<grid, 256, 0, st>>>(fp, firstChunk, lastChunk);
    else k_film<MODE_HERO><<<grid, 256, 0, st>>>(fp, firstChunk, lastChunk);
}
template <int STACK>
static void launchTraceStack(const TraceParams& tp, bool count, bool flat, int grid, cudaStream_t st) {
    if (count) { if (flat) k_trace<true, true, STACK><<<grid, TRACE_BLOCK, 0, st>>>(tp); else k_trace<true, false, STACK><<<grid, TRACE_BLOCK, 0, st>>>(tp); }
    else { if (flat) k_trace<false, true, STACK><<<grid, TRACE_BLOCK, 0, st>>>(tp); else k_trace<false, false, STACK><<<grid, TRACE_BLOCK, 0, st>>>(tp); }
}
void launchTrace(const TraceParams& tp, bool count, int grid, cudaStream_t st) {
    const bool flat = tp.scene.accel.flat != 0u;
    // the stack size follows the depth of the built trees (api.cu refuses trees deeper than TRACE_STACK_DEEP)
    if (tp.scene.accel.stackNeed > (uint32_t)TRACE_STACK) launchTraceStack<TRACE_STACK_DEEP>(tp, count, flat, grid, st);
    else launchTraceStack<TRACE_STACK>(tp, count, flat, grid, st);
}
void launchPrimaryRaygen(const FrameParams& fp, int jittered, int grid, cudaStream_t st) { k_primary_raygen<<<grid, 256, 0, st>>>(fp, jittered); }
void launchPrimaryStore(const FrameParams& fp, int grid, cudaStream_t st) { k_primary_store<<<grid, 256, 0, st>>>(fp); }
void launchUntile(const void* src, void* dst, const TileMap& tm, const uint32_t* l2g, uint32_t words, int grid, cudaStream_t st) {
    k_untile<<<grid, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, tm, l2g, words, 0u);
}

} // namespace vk
