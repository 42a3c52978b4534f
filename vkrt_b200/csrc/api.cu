// api.cu — implementation of the vkrt_cuda_* C ABI (include/vkrt_cuda.h): device context, scene uploads, acceleration
// structure build orchestration, the per-frame wavefront launch sequence and read-back.
//
// Replaces, on the reference side: src/core/runtime/command/record.c:109-179,448-486,577-599 (per-frame command recording),
// src/core/render/accel/{blas,tlas}.c (acceleration structures), src/core/runtime/{buffer,images}.c (device memory) and
// src/core/utility/export/api.c:170-242 (read-back).  There is NO CPU fallback: every entry point needs a CUDA device.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/vkrt_cuda.h"
#include "build.h"
#include "tiles.h"
#include "wavefront.cuh"
#include "../../include/vkrt_closure.h"

namespace vk {
void launchRaygen(int mode, const FrameParams& fp, int grid, cudaStream_t st);
void launchShade(int mode, const FrameParams& fp, uint32_t depth, int grid, cudaStream_t st);
void launchShadeSort(const FrameParams& fp, uint32_t depth, int smCount, cudaStream_t st);
void launchFilm(int mode, const FrameParams& fp, int firstChunk, int lastChunk, int grid, cudaStream_t st);
void launchEnvWeights(const SceneView& sc, uint32_t textureIndex, float* out, int grid, cudaStream_t st);
void launchSpectralMemo(const SceneView& sc, uint32_t materialCount, uint32_t emissiveMeshCount, SpectralMemoEntry* materialMemo, SpectralMemoEntry* emissiveMemo, cudaStream_t st);
void launchTrace(const TraceParams& tp, bool count, int grid, cudaStream_t st);  // picks the flat / two-level kernel from tp.scene.accel.flat
void launchPrimaryRaygen(const FrameParams& fp, int jittered, int grid, cudaStream_t st);
void launchProbeAccum(const ::float4* accum, const TileMap& tm, const uint32_t* l2g, const ::uint2* xy, uint32_t count, ::float4* out, cudaStream_t st);
void launchPackRgb2spec(const float* table, uint32_t dataOffset, size_t cellCount, ::float4* cells, int grid, cudaStream_t st);
void launchPrimaryStore(const FrameParams& fp, int grid, cudaStream_t st);
void launchUntile(const void* src, void* dst, const TileMap& tm, const uint32_t* l2g, uint32_t words, int grid, cudaStream_t st);
void launchMeshTrig(const MeshInfo* infos, MeshTrig* out, uint32_t count, cudaStream_t st);
void launchUntileAll(const void* gathered, void* dst, const TileMap& tm, uint32_t worldSize, const uint32_t* tileLocalIndex, uint64_t rankStrideWords,
                     uint32_t words, int grid, cudaStream_t st);
void launchEvalClosures(const SceneView& sc, const vkrt_closure_query* queries, uint32_t count, vkrt_closure_result* results, cudaStream_t st);
int traceBlocksPerSm(bool count);
int shadeBlocksPerSm(int mode);
}  // namespace vk

using namespace vk;

namespace {

constexpr uint32_t MAX_DEPTH_SLOTS = 72;  // rrMaxDepth is clamped to 64 by the reference (api/settings.c:45-46)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* src, size_t count, cudaStream_t st) {
        cudaError_t e = alloc(count ? count : 1);
        if (e != cudaSuccess) return e;
        if (count) e = cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st);
        return e;
    }
};

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, char[128], int) = nullptr;  // ncclUniqueId passed by value (128 bytes)
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
};

}  // namespace

struct vkrt_cuda_ctx {
    int device = 0;
    int smCount = 148;
    uint32_t rank = 0, worldSize = 1, tileW = 32, tileH = 32;
    uint32_t requestedCapacity = 0, flags = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr, evT0 = nullptr, evT1 = nullptr;
    std::string error;

    // scene
    DevBuf<ShaderVertex> vertices;
    DevBuf<uint32_t> indices;
    uint32_t vertexCount = 0, indexCount = 0;
    std::vector<MeshInfo> hostMeshInfos;
    std::vector<float> hostWorld;
    std::vector<uint32_t> hostGeometrySource;
    std::vector<uint8_t> hostAlpha;
    std::vector<Material> hostMaterials;
    DevBuf<MeshInfo> meshInfos;
    DevBuf<MeshTrig> meshTrig;
    DevBuf<Material> materials;
    DevBuf<EmissiveMesh> emissiveMeshes;
    DevBuf<SpectralMemoEntry> materialMemo, emissiveMemo;   // memoised rgb2spec lookups of material / light constants (k_spectral_memo)
    bool memoDisabled = false;                                // VKRT_NO_SPECTRAL_MEMO=1: every lookup goes to the table (A/B and the bit-identity test)
    bool memoDirty = true;                                    // materials, lights or the rgb2spec table changed since the memo was filled
    DevBuf<EmissiveTriangle> emissiveTriangles;
    DevBuf<float> meshAliasQ, triAliasQ, rgb2spec, srgbLut, world3x4;
    DevBuf<::float4> rgb2specCells;   // the coefficient cells of rgb2spec re-packed as float4 (shading.cuh SpectralTables)
    DevBuf<uint32_t> meshAliasIdx, triAliasIdx;
    RGB2SpecTableInfo rgb2specInfo = {0, 0, 0};
    bool haveRgb2spec = false;
    std::vector<DevBuf<uint8_t>*> texturePixels;
    DevBuf<TextureView> textures;
    uint32_t textureCount = 0;
    uint32_t lightMeshCount = 0;  // emissive meshes uploaded by vkrt_cuda_set_lights (SceneData.emissiveMeshCount may not exceed it)
    // environment-map importance sampling (extension, VKRT_CUDA_FLAG_ENV_IMPORTANCE): table of the texture it was built from
    DevBuf<float> envAliasQ, envPdfUv;
    DevBuf<uint32_t> envAliasIdx;
    uint32_t envTableTexture = VKRT_INVALID_INDEX, envTableW = 0, envTableH = 0;
    bool envTableUsable = false;

    // accel
    AccelBuilder builder;
    DevBuf<Bvh8Node> nodes;
    DevBuf<::float4> triangles;
    DevBuf<InstanceRecord> instanceRecords, instancesLeafOrder;
    DevBuf<::uint2> flatPrims;
    DevBuf<::float4> flatTriangles;   // flat variant: leaf-ordered triangle records, b.w = instance (what k_trace<.., true> reads)
    bool accelFlat = false;
    uint32_t stackNeed = 0;  // traversal stack entries the built trees can require (AccelView::stackNeed)
    DevBuf<::float4> blasBounds;
    DevBuf<uint32_t> instanceBlas;
    uint32_t tlasRoot = 0;
    bool accelValid = false;
    uint32_t traceRefill = 0;       // VKRT_TRACE_REFILL (tuning override of AccelView::refillLanes; 0 = automatic)
    bool anyTransmissive = false;   // some instance record carries INSTANCE_FLAG_TRANSMISSIVE (set by build_accel)
    vkrt_cuda_build_stats buildStats = {};

    // film + wavefront
    uint32_t width = 0, height = 0;
    TileMap tiles = {};
    std::vector<uint32_t> hostL2G;
    DevBuf<uint32_t> l2g;
    uint32_t capacity = 0;
    DevBuf<::float4> f4pool[40];
    DevBuf<uint32_t> u32pool[16];
    DevBuf<float> f32pool[8];
    DevBuf<::uint4> hitA;
    DevBuf<::uint2> u2pool[8];
    DevBuf<uint32_t> counters;  // extCount | shCount | traceWork, MAX_DEPTH_SLOTS each
    DevBuf<uint32_t> shadeOrder, sortBins;
    DevBuf<unsigned long long> stats;
    FrameParams fp = {};
    int readIndex = 0;
    SceneData lastScene = {};
    bool haveScene = false;
    int traceGrid = 0, shadeGrid[3] = {0, 0, 0};
    DevBuf<uint8_t> staging;  // full-frame un-tiled image for read_aov
    DevBuf<uint8_t> gathered; // rank 0: concatenated tile-compact buffers of all ranks
    bool filmIsFullFrame[8] = {false};
    DevBuf<uint8_t> fullFrame[4];  // rank 0 after gather: accum, albedo, normal, output (row-major)
    DevBuf<uint32_t> tileLocalIndex;  // per global tile: its position among its owner's tiles (built by resize, used by the gather)

    // per-launch stage timing (VKRT_CUDA_FLAG_STAGE_TIMING): event after every launch, kind 0 = raygen/shade/film, 1 = trace
    std::vector<cudaEvent_t> stageEvents;
    std::vector<int> stageKinds;
    size_t stageUsed = 0;

    // nccl
    NcclApi nccl;
    void* comm = nullptr;
};

namespace {

VKRT_Result fail(vkrt_cuda_ctx* c, VKRT_Result code, const char* fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->error = buf;
    return code;
}
VKRT_Result cudaFail(vkrt_cuda_ctx* c, cudaError_t e, const char* what) {
    VKRT_Result code = VKRT_ERROR_OPERATION_FAILED;
    if (e == cudaErrorMemoryAllocation) code = VKRT_ERROR_OUT_OF_MEMORY;
    else if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorDevicesUnavailable || e == cudaErrorIllegalAddress ||
             e == cudaErrorLaunchFailure || e == cudaErrorECCUncorrectable || e == cudaErrorHardwareStackError)
        code = VKRT_ERROR_DEVICE_LOST;
    return fail(c, code, "%s: %s", what, cudaGetErrorString(e));
}
#define CU(call)                                                  \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return cudaFail(ctx, e_, #call);   \
    } while (0)

// Affine inverse in fp64 (adjugate), rounded once to fp32; pinned formula shared with the oracle's restatement.
void invertAffine3x4(const float m[12], float out[12]) {
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

uint32_t modeFlagsFor(const SceneData& sd, bool envImportance) {
    uint32_t f = 0;
    if (sd.debugMode == VKRT_DEBUG_MODE_BSDF_ONLY) f |= MODE_BSDF_ONLY;
    if (sd.debugMode == VKRT_DEBUG_MODE_NEE_ONLY) f |= MODE_NEE_ONLY;
    if (sd.debugMode == VKRT_DEBUG_MODE_BOUNCE_COUNT) f |= MODE_BOUNCE_COUNT;
    if (sd.debugMode == VKRT_DEBUG_MODE_DENOISER_ALBEDO) f |= MODE_DN_ALBEDO;
    if (sd.debugMode == VKRT_DEBUG_MODE_DENOISER_NORMAL) f |= MODE_DN_NORMAL;
    if (sd.debugMode == VKRT_DEBUG_MODE_DENOISER_FEATURE_VALIDITY) f |= MODE_DN_VALIDITY;
    if (sd.debugMode == VKRT_DEBUG_MODE_DENOISER_FEATURE_DEPTH) f |= MODE_DN_DEPTH;
    if (sd.debugMode == VKRT_DEBUG_MODE_DENOISER_FOLLOW_SPECULAR) f |= MODE_DN_FOLLOW;
    if (sd.misNeeEnabled != 0u && (sd.emissiveMeshCount > 0u || envImportance)) f |= MODE_NEE_ENABLED;
    return f;
}

SceneView makeSceneView(vkrt_cuda_ctx* c) {
    SceneView v = {};
    v.vertices = c->vertices.p;
    v.indices = c->indices.p;
    v.meshInfos = c->meshInfos.p;
    v.meshTrig = c->meshTrig.p;
    v.materials = c->materials.p;
    v.emissiveMeshes = c->emissiveMeshes.p;
    v.emissiveTriangles = c->emissiveTriangles.p;
    v.meshAliasQ = c->meshAliasQ.p;
    v.meshAliasIdx = c->meshAliasIdx.p;
    v.triAliasQ = c->triAliasQ.p;
    v.triAliasIdx = c->triAliasIdx.p;
    v.textures = c->textures.p;
    v.textureCount = c->textureCount;
    v.srgbLut = c->srgbLut.p;
    v.spectral.info = c->rgb2specInfo;
    v.spectral.table = c->rgb2spec.p;
    v.spectral.cells = c->rgb2specCells.p;
    v.spectral.scale = c->rgb2spec.p + c->rgb2specInfo.scaleOffset;
    if (!c->memoDirty && !c->memoDisabled) {
        v.spectral.materialMemo = c->materialMemo.p;
        v.spectral.emissiveMemo = c->emissiveMemo.p;
    }
    v.accel.nodes = c->nodes.p;
    v.accel.triangles = c->accelFlat ? c->flatTriangles.p : c->triangles.p;
    v.accel.instances = c->instancesLeafOrder.p;
    v.accel.tlasRoot = c->tlasRoot;
    v.accel.instanceCount = (uint32_t)c->hostMeshInfos.size();
    v.accel.flatPrims = c->flatPrims.p;
    v.accel.flat = c->accelFlat ? 1u : 0u;
    v.env = {};
    v.accel.stackNeed = (c->flags & VKRT_CUDA_FLAG_DEEP_STACK) ? (uint32_t)TRACE_STACK_DEEP : c->stackNeed;
    // Refill threshold (measured, profiles/r02_notes.md): long traversals of a big single-level tree amortise a refill over many trips and
    // want the warp full (24); short ones (a few nodes per ray) pay for every refill and do better at 20, like the two-level walk.
    v.accel.refillLanes = c->traceRefill ? c->traceRefill
                                         : ((c->accelFlat && c->buildStats.instancedTriangleCount >= TRACE_REFILL_BIG_SCENE) ? TRACE_REFILL_FLAT : TRACE_REFILL);
    return v;
}

VKRT_Result allocateWavefront(vkrt_cuda_ctx* ctx) {
    const uint32_t lpc = ctx->tiles.localPixelCount;
    uint32_t cap = ctx->requestedCapacity ? ctx->requestedCapacity : (1u << 24);  // 16 Mi paths in flight by default
    if (cap < lpc) cap = lpc;
    cap = (cap / lpc) * lpc;
    ctx->capacity = cap;
    int f4 = 0, u32 = 0, f32 = 0, u2 = 0;
    FrameParams& fp = ctx->fp;
    auto F4 = [&](size_t n) -> ::float4* { return ctx->f4pool[f4].alloc(n) == cudaSuccess ? ctx->f4pool[f4++].p : nullptr; };
    auto U32 = [&](size_t n) -> uint32_t* { return ctx->u32pool[u32].alloc(n) == cudaSuccess ? ctx->u32pool[u32++].p : nullptr; };
    auto F32 = [&](size_t n) -> float* { return ctx->f32pool[f32].alloc(n) == cudaSuccess ? ctx->f32pool[f32++].p : nullptr; };
    auto U2 = [&](size_t n) -> ::uint2* { return ctx->u2pool[u2].alloc(n) == cudaSuccess ? ctx->u2pool[u2++].p : nullptr; };
    bool ok = true;
    for (int s = 0; s < 2; s++) {
        PathState& P = fp.st[s];
        ok &= (P.rayO = F4(cap)) && (P.rayD = F4(cap)) && (P.thr = F4(cap)) && (P.sigma = F4(cap)) && (P.techPdf = F4(cap)) &&
              (P.prevVertexTechPdf = F4(cap)) && (P.prevBsdfTechPdf = F4(cap)) && (P.heroMisc = F4(cap));
        ok &= (P.meta = reinterpret_cast<::uint4*>(F4(cap))) != nullptr;
    }
    ok &= (fp.shO = F4(cap)) && (fp.shD = F4(cap)) && (fp.shContribution = F4(cap)) && (fp.shSeed = U32(cap)) && (fp.shTarget = U2(cap));
    ok &= ctx->hitA.alloc(cap) == cudaSuccess;
    fp.hitA = ctx->hitA.p;
    ok &= (fp.rec.radiance = F4(cap)) && (fp.rec.featA = F4(cap)) && (fp.rec.featB = F4(cap));
    ok &= (fp.rec.radianceScalar = F32(cap)) && (fp.rec.unitWavelength = F32(cap)) && (fp.rec.follow = F32(cap));
    Film& film = fp.film;
    for (int s = 0; s < 2; s++) ok &= (film.accum[s] = F4(lpc)) && (film.albedo[s] = U2(lpc)) && (film.normal[s] = U2(lpc));
    ok &= (film.output = U2(lpc)) && (film.frameRadiance = F4(lpc)) && (film.frameFeatA = F4(lpc)) && (film.frameFeatB = F4(lpc)) &&
          (film.frameFollow = F32(lpc)) && (film.debugColor = F4(lpc)) && (film.bounceCount = U32(lpc)) && (film.hitId = U2(lpc)) &&
          (film.hitTuv = F4(lpc));
    if (!ok) return fail(ctx, VKRT_ERROR_OUT_OF_MEMORY, "wavefront allocation failed (capacity %u paths, %u local pixels)", cap, lpc);
    if (ctx->counters.alloc(MAX_DEPTH_SLOTS * 4) != cudaSuccess || ctx->stats.alloc(4) != cudaSuccess)
        return fail(ctx, VKRT_ERROR_OUT_OF_MEMORY, "counter allocation failed");
    if (ctx->shadeOrder.alloc(cap) != cudaSuccess || ctx->sortBins.alloc(512) != cudaSuccess)
        return fail(ctx, VKRT_ERROR_OUT_OF_MEMORY, "sort buffer allocation failed");
    fp.shadeOrder = (ctx->flags & VKRT_CUDA_FLAG_NO_MATERIAL_SORT) ? nullptr : ctx->shadeOrder.p;
    fp.sortBins = ctx->sortBins.p;
    fp.extCount = ctx->counters.p;
    fp.shCount = ctx->counters.p + MAX_DEPTH_SLOTS;
    fp.traceWork = ctx->counters.p + 2 * MAX_DEPTH_SLOTS;
    fp.capacity = cap;
    return VKRT_SUCCESS;
}

VKRT_Result resetAccumulation(vkrt_cuda_ctx* ctx) {
    const size_t lpc = ctx->tiles.localPixelCount;
    if (!lpc) return VKRT_SUCCESS;
    Film& film = ctx->fp.film;
    for (int s = 0; s < 2; s++) {
        CU(cudaMemsetAsync(film.accum[s], 0, lpc * sizeof(::float4), ctx->stream));
        CU(cudaMemsetAsync(film.albedo[s], 0, lpc * sizeof(::uint2), ctx->stream));
        CU(cudaMemsetAsync(film.normal[s], 0, lpc * sizeof(::uint2), ctx->stream));
    }
    CU(cudaMemsetAsync(film.output, 0, lpc * sizeof(::uint2), ctx->stream));
    return VKRT_SUCCESS;
}

// Builds (once per environment texture) the alias table the extension samples: texel weights on the device, Vose's method in
// double precision on the host. Returns false when the texture carries no energy (the extension then stays off for the frame).
VKRT_Result ensureEnvTable(vkrt_cuda_ctx* ctx, uint32_t textureIndex) {
    if (ctx->envTableTexture == textureIndex) return VKRT_SUCCESS;
    ctx->envTableTexture = textureIndex;
    ctx->envTableUsable = false;
    TextureView tv;
    CU(cudaMemcpy(&tv, ctx->textures.p + textureIndex, sizeof(tv), cudaMemcpyDeviceToHost));
    const uint64_t n64 = (uint64_t)tv.width * tv.height;
    if (n64 == 0 || n64 > (1ull << 26)) return VKRT_SUCCESS;
    const uint32_t n = (uint32_t)n64;
    CU(ctx->envPdfUv.alloc(n));
    launchEnvWeights(makeSceneView(ctx), textureIndex, ctx->envPdfUv.p, ctx->smCount * 8, ctx->stream);
    std::vector<float> w(n);
    CU(cudaMemcpyAsync(w.data(), ctx->envPdfUv.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (float x : w) sum += x;
    if (!(sum > 0.0)) return VKRT_SUCCESS;
    std::vector<double> scaled(n);
    std::vector<float> q(n), pdfUv(n);
    std::vector<uint32_t> alias(n), small, large;
    for (uint32_t i = 0; i < n; i++) {
        const double pmf = w[i] / sum;
        pdfUv[i] = (float)(pmf * n);
        scaled[i] = pmf * n;
        alias[i] = i;
        (scaled[i] < 1.0 ? small : large).push_back(i);
    }
    while (!small.empty() && !large.empty()) {
        const uint32_t a = small.back(), b = large.back();
        small.pop_back();
        q[a] = (float)scaled[a];
        alias[a] = b;
        scaled[b] -= 1.0 - scaled[a];
        if (scaled[b] < 1.0) { large.pop_back(); small.push_back(b); }
    }
    for (uint32_t i : large) q[i] = 1.0f;
    for (uint32_t i : small) q[i] = 1.0f;
    CU(ctx->envPdfUv.upload(pdfUv.data(), n, ctx->stream));
    CU(ctx->envAliasQ.upload(q.data(), n, ctx->stream));
    CU(ctx->envAliasIdx.upload(alias.data(), n, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->envTableW = tv.width;
    ctx->envTableH = tv.height;
    ctx->envTableUsable = true;
    return VKRT_SUCCESS;
}

void fillFrameParams(vkrt_cuda_ctx* ctx, const SceneData& sd) {
    FrameParams& fp = ctx->fp;
    fp.scene = makeSceneView(ctx);
    const bool envImportance = (ctx->flags & VKRT_CUDA_FLAG_ENV_IMPORTANCE) && sd.environmentTextureIndex < ctx->textureCount &&
                               ctx->envTableTexture == sd.environmentTextureIndex && ctx->envTableUsable;
    if (envImportance) {
        EnvDistribution& e = fp.scene.env;
        e.aliasQ = ctx->envAliasQ.p;
        e.aliasIdx = ctx->envAliasIdx.p;
        e.pdfUv = ctx->envPdfUv.p;
        e.width = ctx->envTableW;
        e.height = ctx->envTableH;
        e.pEnv = sd.emissiveMeshCount > 0u ? 0.5f : 1.0f;
        e.active = 1u;
    }
    fp.sd = sd;
    fp.tiles = ctx->tiles;
    fp.tiles.localToGlobalTile = ctx->l2g.p;
    fp.readIndex = ctx->readIndex;
    fp.modeFlags = modeFlagsFor(sd, envImportance) | (ctx->anyTransmissive ? MODE_SCENE_TRANSMISSIVE : 0u);
}

TraceParams makeTraceParams(vkrt_cuda_ctx* ctx, uint32_t depth, bool haveExt, bool haveShadow) {
    const FrameParams& fp = ctx->fp;
    TraceParams tp = {};
    tp.scene = fp.scene;
    const PathState& S = fp.st[depth & 1u];
    if (haveExt) {
        tp.rayO = S.rayO;
        tp.rayD = S.rayD;
        tp.raySeed = reinterpret_cast<const uint32_t*>(S.meta) + 1;
        tp.hitA = fp.hitA;
        tp.hitB = reinterpret_cast<float*>(S.meta) + 3;
        tp.extCount = fp.extCount + depth;
    }
    if (haveShadow) {
        tp.shO = fp.shO;
        tp.shD = fp.shD;
        tp.shContribution = fp.shContribution;
        tp.shTarget = fp.shTarget;
        tp.shSeed = fp.shSeed;
        tp.shCount = fp.shCount + (depth - 1u);
        tp.radiance = fp.rec.radiance;
        tp.radianceScalar = fp.rec.radianceScalar;
        tp.pathFlags = reinterpret_cast<uint32_t*>(S.meta) + 2;
    }
    tp.slotStride = 4u;
    tp.workCounter = fp.traceWork + depth;
    tp.stats = (ctx->flags & VKRT_CUDA_FLAG_COUNT_RAYS) ? ctx->stats.p : nullptr;
    return tp;
}

int renderModeOf(const SceneData& sd) {
    if (VKRT_RENDER_SETTINGS_MODE(sd.packedRenderSettings) != VKRT_RENDER_MODE_SPECTRAL) return 0;
    return VKRT_RENDER_SETTINGS_SPECTRAL(sd.packedRenderSettings) == VKRT_SPECTRAL_SAMPLING_MODE_HERO ? 2 : 1;
}

VKRT_Result enqueueFrame(vkrt_cuda_ctx* ctx, const SceneData* sceneData, uint32_t* launches) {
    if (!ctx->accelValid) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "render_frame: acceleration structure not built");
    if (!ctx->width) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "render_frame: resize not called");
    SceneData sd = *sceneData;
    const int mode = renderModeOf(sd);
    if (mode != 0 && !ctx->haveRgb2spec) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "spectral rendering needs vkrt_cuda_set_rgb2spec");
    if (mode != 0 && ctx->memoDirty) {   // spectral memo of the material / light constants (k_spectral_memo), refreshed after a scene edit
        const uint32_t nm = (uint32_t)ctx->hostMaterials.size(), ne = ctx->lightMeshCount;
        CU(ctx->materialMemo.alloc((size_t)std::max(nm, 1u) * SPECTRAL_MEMO_SLOTS));
        CU(ctx->emissiveMemo.alloc(std::max(ne, 1u)));
        launchSpectralMemo(makeSceneView(ctx), nm, ne, ctx->materialMemo.p, ctx->emissiveMemo.p, ctx->stream);
        CU(cudaGetLastError());
        ctx->memoDirty = false;
    }
    if (sd.rrMaxDepth > 64u) sd.rrMaxDepth = 64u;
    if (sd.emissiveMeshCount > ctx->lightMeshCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "render_frame: SceneData names %u emissive meshes, %u uploaded", sd.emissiveMeshCount, ctx->lightMeshCount);
    const uint32_t spp = std::max(sd.samplesPerPixel, 1u);
    sd.samplesPerPixel = spp;
    if ((ctx->flags & VKRT_CUDA_FLAG_ENV_IMPORTANCE) && sd.environmentTextureIndex < ctx->textureCount) {
        const VKRT_Result r = ensureEnvTable(ctx, sd.environmentTextureIndex);
        if (r != VKRT_SUCCESS) return r;
    }
    fillFrameParams(ctx, sd);
    FrameParams& fp = ctx->fp;
    const uint32_t lpc = ctx->tiles.localPixelCount;
    const uint32_t samplesPerChunk = std::max(1u, std::min(spp, ctx->capacity / lpc));
    const bool count = (ctx->flags & VKRT_CUDA_FLAG_COUNT_RAYS) != 0;
    cudaStream_t st = ctx->stream;
    uint32_t nl = 0;
    const bool timing = launches != nullptr && (ctx->flags & VKRT_CUDA_FLAG_STAGE_TIMING) != 0;
    ctx->stageUsed = 0;
    ctx->stageKinds.clear();
    auto mark = [&](int kind) {
        if (!timing) return;
        if (ctx->stageUsed == ctx->stageEvents.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            ctx->stageEvents.push_back(e);
        }
        cudaEventRecord(ctx->stageEvents[ctx->stageUsed++], st);
        ctx->stageKinds.push_back(kind);
    };
    mark(-1);
    if (sd.debugMode != VKRT_DEBUG_MODE_NONE) {
        CU(cudaMemsetAsync(fp.film.debugColor, 0, lpc * sizeof(::float4), st));
        CU(cudaMemsetAsync(fp.film.bounceCount, 0, lpc * sizeof(uint32_t), st));
    }
    for (uint32_t s0 = 0; s0 < spp; s0 += samplesPerChunk) {
        fp.chunkFirstSample = s0;
        fp.chunkSamples = std::min(samplesPerChunk, spp - s0);
        CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(uint32_t) * MAX_DEPTH_SLOTS * 4, st));
        launchRaygen(mode, fp, ctx->shadeGrid[mode] * 2, st);
        nl++;
        mark(0);
        for (uint32_t d = 0; d < sd.rrMaxDepth; d++) {
            launchTrace(makeTraceParams(ctx, d, true, d > 0), count, ctx->traceGrid, st);
            mark(1);
            if (fp.shadeOrder) { launchShadeSort(fp, d, ctx->smCount, st); nl += 3; mark(0); }
            launchShade(mode, fp, d, ctx->shadeGrid[mode], st);
            mark(2);
            nl += 2;
        }
        if (sd.rrMaxDepth > 0) {
            launchTrace(makeTraceParams(ctx, sd.rrMaxDepth, false, true), count, ctx->traceGrid, st);
            nl++;
            mark(1);
        }
        launchFilm(mode, fp, s0 == 0, s0 + fp.chunkSamples >= spp, ctx->smCount * 4, st);
        nl++;
        mark(0);
    }
    CU(cudaGetLastError());
    ctx->readIndex = 1 - ctx->readIndex;  // frame.c:386-388
    ctx->lastScene = sd;
    ctx->haveScene = true;
    for (bool& b : ctx->filmIsFullFrame) b = false;
    if (launches) *launches = nl;
    return VKRT_SUCCESS;
}

}  // namespace

// ======================================================================================================================
// C ABI
// ======================================================================================================================
extern "C" {

VKRT_CUDA_API const char* vkrt_cuda_version(void) { return "vkrt-b200 0.1 (sm_100a)"; }

VKRT_CUDA_API const char* vkrt_cuda_last_error(const vkrt_cuda_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

VKRT_CUDA_API VKRT_Result vkrt_cuda_create(const vkrt_cuda_create_info* info, vkrt_cuda_ctx** outCtx) {
    if (!outCtx) return VKRT_ERROR_INVALID_ARGUMENT;
    *outCtx = nullptr;
    int deviceCount = 0;
    cudaError_t e = cudaGetDeviceCount(&deviceCount);
    if (e != cudaSuccess || deviceCount == 0) {
        fprintf(stderr, "vkrt_cuda_create: no CUDA device (%s). This library has no CPU path.\n", cudaGetErrorString(e));
        return VKRT_ERROR_INITIALIZATION_FAILED;
    }
    vkrt_cuda_ctx* ctx = new vkrt_cuda_ctx();
    int dev = info ? info->device : -1;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= deviceCount) {
        delete ctx;
        return VKRT_ERROR_INVALID_ARGUMENT;
    }
    ctx->device = dev;
    if (info) {
        ctx->rank = info->rank;
        ctx->worldSize = info->worldSize ? info->worldSize : 1;
        ctx->tileW = info->tileWidth ? info->tileWidth : 32;
        ctx->tileH = info->tileHeight ? info->tileHeight : 32;
        ctx->requestedCapacity = info->maxPathsInFlight;
        ctx->flags = info->flags;
    }
    if (ctx->rank >= ctx->worldSize) {
        delete ctx;
        return VKRT_ERROR_INVALID_ARGUMENT;
    }
    ctx->builder.mode = (ctx->flags & VKRT_CUDA_FLAG_LBVH) ? AccelBuilder::BUILD_LBVH : ((ctx->flags & VKRT_CUDA_FLAG_PLOC) ? AccelBuilder::BUILD_PLOC : AccelBuilder::BUILD_BEST);
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->evA) != cudaSuccess || cudaEventCreate(&ctx->evB) != cudaSuccess) {
        delete ctx;
        return VKRT_ERROR_INITIALIZATION_FAILED;
    }
    cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, dev);
    ctx->traceGrid = ctx->smCount * traceBlocksPerSm((ctx->flags & VKRT_CUDA_FLAG_COUNT_RAYS) != 0);
    if (const char* m = getenv("VKRT_NO_SPECTRAL_MEMO")) ctx->memoDisabled = atoi(m) != 0;
    if (const char* r = getenv("VKRT_TRACE_REFILL")) ctx->traceRefill = (uint32_t)std::min(std::max(atoi(r), 0), 32);   // tuning knob (tests/perf_probe.py sweeps)
    for (int m = 0; m < 3; m++) ctx->shadeGrid[m] = ctx->smCount * shadeBlocksPerSm(m);
    // sRGB decode table (256 entries), same formula as the oracle
    float lut[256];
    for (int i = 0; i < 256; i++) {
        float v = float(i) / 255.0f;
        lut[i] = v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
    }
    if (ctx->srgbLut.upload(lut, 256, ctx->stream) != cudaSuccess) {
        delete ctx;
        return VKRT_ERROR_OUT_OF_MEMORY;
    }
    // placeholders so that empty scenes still have valid pointers
    ctx->emissiveMeshes.alloc(1); ctx->emissiveTriangles.alloc(1); ctx->meshAliasQ.alloc(1); ctx->meshAliasIdx.alloc(1);
    ctx->triAliasQ.alloc(1); ctx->triAliasIdx.alloc(1); ctx->textures.alloc(1); ctx->rgb2spec.alloc(1); ctx->rgb2specCells.alloc(1);
    cudaStreamSynchronize(ctx->stream);
    *outCtx = ctx;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API void vkrt_cuda_destroy(vkrt_cuda_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
    for (auto* t : ctx->texturePixels) delete t;
    for (cudaEvent_t e : ctx->stageEvents) cudaEventDestroy(e);
    if (ctx->evA) cudaEventDestroy(ctx->evA);
    if (ctx->evT0) cudaEventDestroy(ctx->evT0);
    if (ctx->evT1) cudaEventDestroy(ctx->evT1);
    if (ctx->evB) cudaEventDestroy(ctx->evB);
    cudaStream_t st = ctx->stream;
    delete ctx;
    if (st) cudaStreamDestroy(st);
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_set_geometry(vkrt_cuda_ctx* ctx, const ShaderVertex* vertices, uint32_t vertexCount, const uint32_t* indices,
                                                 uint32_t indexCount) {
    if (!ctx || (vertexCount && !vertices) || (indexCount && !indices)) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    CU(ctx->vertices.upload(vertices, vertexCount, ctx->stream));
    CU(ctx->indices.upload(indices, indexCount, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->vertexCount = vertexCount;
    ctx->indexCount = indexCount;
    ctx->accelValid = false;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_set_instances(vkrt_cuda_ctx* ctx, const MeshInfo* infos, const float* world3x4, const uint32_t* geometrySource,
                                                  const uint8_t* alphaTested, uint32_t instanceCount) {
    if (!ctx || (instanceCount && (!infos || !world3x4))) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    // The acceleration structure depends on: which geometry each instance shares, its world matrix, its any-hit flag and whether its
    // material transmits. An update that changes none of these (material index between two opaque materials, lightPdfArea after a
    // light edit, opacity above the any-hit threshold ...) keeps the built BVH: interactive edits must not pay a rebuild.
    auto transmissive = [&](uint32_t materialIndex) { return materialIndex < ctx->hostMaterials.size() && ctx->hostMaterials[materialIndex].transmission > 0.0f; };
    bool same = ctx->accelValid && ctx->hostMeshInfos.size() == instanceCount && ctx->hostWorld.size() == (size_t)instanceCount * 12 &&
                (instanceCount == 0 || memcmp(ctx->hostWorld.data(), world3x4, sizeof(float) * 12 * instanceCount) == 0) &&
                (geometrySource ? (ctx->hostGeometrySource.size() == instanceCount && memcmp(ctx->hostGeometrySource.data(), geometrySource, sizeof(uint32_t) * instanceCount) == 0)
                                : ctx->hostGeometrySource.empty());
    for (uint32_t i = 0; same && i < instanceCount; i++) {
        const MeshInfo &a = ctx->hostMeshInfos[i], &b = infos[i];
        same = a.vertexBase == b.vertexBase && a.vertexCount == b.vertexCount && a.indexBase == b.indexBase && a.indexCount == b.indexCount &&
               transmissive(a.materialIndex) == transmissive(b.materialIndex) && b.materialIndex < std::max<size_t>(ctx->hostMaterials.size(), 1) &&
               ctx->hostAlpha[i] == (alphaTested ? alphaTested[i] : 0);
    }
    ctx->hostMeshInfos.assign(infos, infos + instanceCount);
    ctx->hostWorld.assign(world3x4, world3x4 + (size_t)instanceCount * 12);
    if (geometrySource) ctx->hostGeometrySource.assign(geometrySource, geometrySource + instanceCount);
    else ctx->hostGeometrySource.clear();
    if (alphaTested) ctx->hostAlpha.assign(alphaTested, alphaTested + instanceCount);
    else ctx->hostAlpha.assign(instanceCount, 0);
    CU(ctx->meshInfos.upload(infos, instanceCount, ctx->stream));
    CU(ctx->meshTrig.alloc(instanceCount));
    launchMeshTrig(ctx->meshInfos.p, ctx->meshTrig.p, instanceCount, ctx->stream);
    CU(ctx->world3x4.upload(world3x4, (size_t)instanceCount * 12, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (!same) ctx->accelValid = false;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_set_materials(vkrt_cuda_ctx* ctx, const Material* materials, uint32_t materialCount) {
    if (!ctx || (materialCount && !materials)) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    // the instance records carry one bit per material: "transmits" (shadow rays: unsupported transmission). Only a change of that bit,
    // or of the material count the instances were validated against, invalidates the built BVH.
    bool same = ctx->accelValid && ctx->hostMaterials.size() == materialCount;
    for (uint32_t i = 0; same && i < materialCount; i++) same = (ctx->hostMaterials[i].transmission > 0.0f) == (materials[i].transmission > 0.0f);
    ctx->hostMaterials.assign(materials, materials + materialCount);
    ctx->memoDirty = true;
    CU(ctx->materials.upload(materials, materialCount, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (!same) ctx->accelValid = false;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_set_lights(vkrt_cuda_ctx* ctx, const EmissiveMesh* meshes, uint32_t meshCount, const EmissiveTriangle* triangles,
                                               uint32_t triangleCount, const float* meshAliasQ, const uint32_t* meshAliasIdx, const float* triAliasQ,
                                               const uint32_t* triAliasIdx) {
    if (!ctx) return VKRT_ERROR_INVALID_ARGUMENT;
    if (meshCount && (!meshes || !meshAliasQ || !meshAliasIdx)) return VKRT_ERROR_INVALID_ARGUMENT;
    if (triangleCount && (!triangles || !triAliasQ || !triAliasIdx)) return VKRT_ERROR_INVALID_ARGUMENT;
    // the tables are followed blindly by the light sampler (light_sampling.slang:9-37): every link must stay inside them
    for (uint32_t m = 0; m < meshCount; m++) {
        if (meshAliasIdx[m] >= meshCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "light tables: mesh alias %u out of range", m);
        const EmissiveMesh& em = meshes[m];
        if ((uint64_t)em.triOffset + em.triCount > triangleCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "light tables: emissive mesh %u triangle range out of bounds", m);
        for (uint32_t k = 0; k < em.triCount; k++)
            if (triAliasIdx[em.triOffset + k] >= em.triCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "light tables: triangle alias %u of mesh %u out of range", k, m);
    }
    cudaSetDevice(ctx->device);
    ctx->lightMeshCount = meshCount;
    ctx->memoDirty = true;
    CU(ctx->emissiveMeshes.upload(meshes, meshCount, ctx->stream));
    CU(ctx->emissiveTriangles.upload(triangles, triangleCount, ctx->stream));
    CU(ctx->meshAliasQ.upload(meshAliasQ, meshCount, ctx->stream));
    CU(ctx->meshAliasIdx.upload(meshAliasIdx, meshCount, ctx->stream));
    CU(ctx->triAliasQ.upload(triAliasQ, triangleCount, ctx->stream));
    CU(ctx->triAliasIdx.upload(triAliasIdx, triangleCount, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_set_textures(vkrt_cuda_ctx* ctx, const vkrt_cuda_texture* textures, uint32_t textureCount) {
    if (!ctx || (textureCount && !textures) || textureCount > VKRT_MAX_BINDLESS_TEXTURES) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    static const size_t bpp[4] = {4, 8, 8, 16};
    for (auto* t : ctx->texturePixels) delete t;
    ctx->texturePixels.clear();
    std::vector<TextureView> views(textureCount);
    for (uint32_t i = 0; i < textureCount; i++) {
        if (textures[i].format >= VKRT_TEXTURE_FORMAT_COUNT || !textures[i].pixels) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "texture %u invalid", i);
        auto* buf = new DevBuf<uint8_t>();
        ctx->texturePixels.push_back(buf);
        size_t bytes = (size_t)textures[i].width * textures[i].height * bpp[textures[i].format];
        CU(buf->upload((const uint8_t*)textures[i].pixels, bytes, ctx->stream));
        views[i].pixels = buf->p;
        views[i].width = textures[i].width;
        views[i].height = textures[i].height;
        views[i].format = textures[i].format;
        views[i].colorSpace = textures[i].colorSpace;
    }
    CU(ctx->textures.upload(views.data(), textureCount, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->textureCount = textureCount;
    ctx->envTableTexture = VKRT_INVALID_INDEX;  // the importance-sampling table follows the texture set
    ctx->envTableUsable = false;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_set_rgb2spec(vkrt_cuda_ctx* ctx, const float* payload, uint32_t floatCount, RGB2SpecTableInfo info) {
    if (!ctx || !payload || info.res < 2) return VKRT_ERROR_INVALID_ARGUMENT;
    const uint64_t need = (uint64_t)info.dataOffset + 9ull * info.res * info.res * info.res;
    if (need > floatCount || (uint64_t)info.scaleOffset + info.res > floatCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "rgb2spec payload too small");
    cudaSetDevice(ctx->device);
    CU(ctx->rgb2spec.upload(payload, floatCount, ctx->stream));
    const size_t cellCount = 3ull * info.res * info.res * info.res;
    CU(ctx->rgb2specCells.alloc(cellCount));
    launchPackRgb2spec(ctx->rgb2spec.p, info.dataOffset, cellCount, ctx->rgb2specCells.p, ctx->smCount * 8, ctx->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    if (const char* e = getenv("VKRT_L2_PERSIST")) {   // tuning knob (profiles/r02_notes.md): pin the coefficient cells in the persisting part of L2
        if (atoi(e) != 0) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
                const size_t bytes = std::min(std::min(cellCount * sizeof(::float4), (size_t)prop.persistingL2CacheMaxSize), (size_t)prop.accessPolicyMaxWindowSize);
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes);
                cudaStreamAttrValue av = {};
                av.accessPolicyWindow.base_ptr = ctx->rgb2specCells.p;
                av.accessPolicyWindow.num_bytes = bytes;
                av.accessPolicyWindow.hitRatio = 1.0f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
                cudaGetLastError();
            }
        }
    }
    ctx->rgb2specInfo = info;
    ctx->haveRgb2spec = true;
    ctx->memoDirty = true;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_build_accel(vkrt_cuda_ctx* ctx, vkrt_cuda_build_stats* outStats) {
    if (!ctx) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (ctx->accelValid) {   // nothing the BVH depends on changed since the last build (see vkrt_cuda_set_instances / set_materials)
        if (outStats) *outStats = ctx->buildStats;
        return VKRT_SUCCESS;
    }
    const uint32_t n = (uint32_t)ctx->hostMeshInfos.size();
    uint32_t plocKept = 0;
    // --- unique geometries (geometry.c:166-210 decides sharing on the host; here it arrives as geometrySource) ---
    struct BlasDesc { uint32_t vertexBase, indexBase, triCount, nodeBase, primBase, nodeCount, vertexLimit; };
    std::vector<BlasDesc> blas;
    std::vector<uint32_t> instanceBlas(n);
    std::map<std::tuple<uint32_t, uint32_t, uint32_t>, uint32_t> byRange;
    std::map<uint32_t, uint32_t> bySource;
    uint64_t totalTris = 0, instancedTris = 0;
    for (uint32_t i = 0; i < n; i++) {
        const MeshInfo& mi = ctx->hostMeshInfos[i];
        if ((uint64_t)mi.indexBase + mi.indexCount > ctx->indexCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "instance %u: index range out of bounds", i);
        if (mi.materialIndex >= ctx->hostMaterials.size()) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "instance %u: material index out of range", i);
        if ((uint64_t)mi.vertexBase + mi.vertexCount > ctx->vertexCount) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "instance %u: vertex range out of bounds", i);
        auto key = std::make_tuple(mi.vertexBase, mi.indexBase, mi.indexCount);
        uint32_t b;
        bool found = false;
        if (!ctx->hostGeometrySource.empty()) {
            auto it = bySource.find(ctx->hostGeometrySource[i]);
            if (it != bySource.end()) { b = it->second; found = true; }
        } else {
            auto it = byRange.find(key);
            if (it != byRange.end()) { b = it->second; found = true; }
        }
        if (!found) {
            b = (uint32_t)blas.size();
            blas.push_back({mi.vertexBase, mi.indexBase, mi.indexCount / 3u, 0u, 0u, 0u, mi.vertexCount ? mi.vertexCount : ctx->vertexCount - mi.vertexBase});
            totalTris += mi.indexCount / 3u;
            if (!ctx->hostGeometrySource.empty()) bySource[ctx->hostGeometrySource[i]] = b;
            else byRange[key] = b;
        }
        instanceBlas[i] = b;
        instancedTris += mi.indexCount / 3u;
    }
    if (totalTris > 0xFFFFFFF0ull) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "too many triangles");
    {   // every index must address a vertex of its geometry before any kernel dereferences it (one pass over the index buffer)
        DevBuf<uint32_t> badIndex;
        CU(badIndex.alloc(1));
        CU(cudaMemsetAsync(badIndex.p, 0, sizeof(uint32_t), ctx->stream));
        for (const BlasDesc& d : blas) launchValidateIndices(ctx->indices.p, d.indexBase, d.triCount * 3u, d.vertexLimit, badIndex.p, ctx->stream);
        uint32_t bad = 0;
        CU(cudaMemcpyAsync(&bad, badIndex.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (bad) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "index buffer addresses a vertex outside its geometry");
    }
    // ---- single-level variant: one BVH over all instanced triangles in world space ------------------------------------------------
    // Instancing (shared BLASes under a TLAS) pays when geometry is reused many times; when it is not — a handful of instances, or many
    // overlapping meshes like the 16-mesh triangle soup, where a ray enters 9 BLASes — a flat BVH traverses far fewer nodes.
    // Decision (radix-tree builder): average overlap depth of the instances' world boxes = sum of box volumes / volume of their union box. Separated
    // instances (cornell: depth < 1) keep their own well-fitting BLASes — mixing ten wall-sized triangles into one Morton-ordered
    // tree with 70 k small ones costs 1.7x the node visits; stacked instances (soup: depth 16) are flattened. Flattening is also
    // limited to scenes where it does not multiply memory (instanced <= 2 x unique triangles, or <= 4 Mi triangles in total).
    bool flat = false;
    if (instancedTris > 0 && instancedTris < 0x7FFFFFF0ull && instancedTris <= std::max<uint64_t>(4ull << 20, 2 * totalTris) &&
        !(ctx->flags & (VKRT_CUDA_FLAG_FORCE_TWO_LEVEL | VKRT_CUDA_FLAG_FORCE_FLAT))) {
        DevBuf<int> gb;
        CU(gb.alloc(8 * std::max<size_t>(blas.size(), 1)));
        for (size_t b = 0; b < blas.size(); b++)
            launchGeometryBounds(ctx->vertices.p, ctx->indices.p, blas[b].vertexBase, blas[b].indexBase, blas[b].triCount, gb.p + 8 * b, ctx->stream);
        std::vector<int> hb(8 * blas.size());
        CU(cudaMemcpyAsync(hb.data(), gb.p, sizeof(int) * hb.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        double ulo[3] = {1e300, 1e300, 1e300}, uhi[3] = {-1e300, -1e300, -1e300};
        std::vector<double> boxes((size_t)n * 6);
        for (uint32_t i = 0; i < n; i++) {
            const int* g = &hb[8 * instanceBlas[i]];
            float lo[3] = {orderedIntToFloatHost(g[0]), orderedIntToFloatHost(g[1]), orderedIntToFloatHost(g[2])};
            float hi[3] = {orderedIntToFloatHost(g[3]), orderedIntToFloatHost(g[4]), orderedIntToFloatHost(g[5])};
            double* bx = &boxes[(size_t)i * 6];
            for (int a = 0; a < 3; a++) { bx[a] = 1e300; bx[3 + a] = -1e300; }
            if (lo[0] > hi[0]) continue;
            const float* m = &ctx->hostWorld[(size_t)i * 12];
            for (int k = 0; k < 8; k++) {
                double p[3] = {(k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]};
                for (int a = 0; a < 3; a++) {
                    double w = m[a * 4] * p[0] + m[a * 4 + 1] * p[1] + m[a * 4 + 2] * p[2] + m[a * 4 + 3];
                    bx[a] = std::min(bx[a], w);
                    bx[3 + a] = std::max(bx[3 + a], w);
                }
            }
            for (int a = 0; a < 3; a++) { ulo[a] = std::min(ulo[a], bx[a]); uhi[a] = std::max(uhi[a], bx[3 + a]); }
        }
        double diag = 0.0;
        for (int a = 0; a < 3; a++) diag = std::max(diag, uhi[a] - ulo[a]);
        const double pad = 1e-3 * diag;  // planes get a sliver of volume, not zero
        double unionVol = 1.0, sumVol = 0.0;
        for (int a = 0; a < 3; a++) unionVol *= (uhi[a] - ulo[a]) + 2 * pad;
        for (uint32_t i = 0; i < n; i++) {
            const double* bx = &boxes[(size_t)i * 6];
            if (bx[0] > bx[3]) continue;
            double v = 1.0;
            for (int a = 0; a < 3; a++) v *= (bx[3 + a] - bx[a]) + 2 * pad;
            sumVol += v;
        }
        flat = unionVol > 0.0 && sumVol / unionVol > 2.0;
        // With the surface-area driven hierarchy (PLOC, picked per BVH by cost) large triangles stay near the root, so one BVH over all
        // instanced triangles beats BLAS + TLAS even for separated instances (cornell: 32.7 against 33.7 ms of traversal per 16-spp
        // 1080p frame, 0.95 against 1.61 instance transforms per ray): flatten whenever it does not multiply memory. A radix tree alone
        // needs the overlap test above (flat cornell: 9.3 node visits per ray against 5.5).
        if (ctx->builder.mode != AccelBuilder::BUILD_LBVH) flat = true;
    }
    if (ctx->flags & VKRT_CUDA_FLAG_FORCE_TWO_LEVEL) flat = false;
    if ((ctx->flags & VKRT_CUDA_FLAG_FORCE_FLAT) && instancedTris > 0 && instancedTris < 0x7FFFFFF0ull) flat = true;
    ctx->accelFlat = flat;
    if (flat) {
        cudaStream_t st = ctx->stream;
        uint32_t pb = 0;
        for (auto& b : blas) { b.primBase = pb; pb += b.triCount; }
        std::vector<::uint4> instTable(n);
        std::vector<uint32_t> triOffsets(n + 1, 0u);
        std::vector<InstanceRecord> records(n);
        bool anyTransmissive = false;
        for (uint32_t i = 0; i < n; i++) {
            const BlasDesc& d = blas[instanceBlas[i]];
            instTable[i] = make_uint4(d.vertexBase, d.indexBase, d.primBase, d.triCount);
            triOffsets[i + 1] = triOffsets[i] + d.triCount;
            float inv[12];
            invertAffine3x4(&ctx->hostWorld[(size_t)i * 12], inv);
            InstanceRecord& r = records[i];
            r.inv0 = make_float4(inv[0], inv[1], inv[2], inv[3]);
            r.inv1 = make_float4(inv[4], inv[5], inv[6], inv[7]);
            r.inv2 = make_float4(inv[8], inv[9], inv[10], inv[11]);
            r.blasRoot = 0;
            r.flags = 0;
            if (ctx->hostAlpha[i]) r.flags |= INSTANCE_FLAG_ALPHA_TESTED;
            if (ctx->hostMaterials[ctx->hostMeshInfos[i].materialIndex].transmission > 0.0f) { r.flags |= INSTANCE_FLAG_TRANSMISSIVE; anyTransmissive = true; }
            if (d.triCount == 0) r.flags |= INSTANCE_FLAG_EMPTY;
            r.instanceIndex = i;
            r.pad = 0;
        }
        const uint32_t total = (uint32_t)instancedTris;
        DevBuf<::uint4> dInstTable;
        DevBuf<uint32_t> dTriOffsets;
        DevBuf<::uint2> flatScratch;
        DevBuf<Bvh8Node> scratchNodes;
        CU(dInstTable.upload(instTable.data(), n, st));
        CU(dTriOffsets.upload(triOffsets.data(), n + 1, st));
        CU(flatScratch.alloc(total));
        CU(scratchNodes.alloc(total));
        CU(ctx->triangles.alloc((size_t)std::max<uint64_t>(totalTris, 1) * 3));
        CU(ctx->flatPrims.alloc(total));
        CU(ctx->instanceRecords.upload(records.data(), records.size(), st));
        CU(ctx->instancesLeafOrder.upload(records.data(), records.size(), st));  // instance order in the flat variant
        CU(cudaEventRecord(ctx->evA, st));
        for (const BlasDesc& d : blas) launchPackTriangles(ctx->vertices.p, ctx->indices.p, d.vertexBase, d.indexBase, d.triCount, ctx->triangles.p, d.primBase, st);
        uint32_t nodeCount = 0, primCount = 0;
        if (!ctx->builder.buildFlat(st, ctx->vertices.p, ctx->indices.p, dInstTable.p, dTriOffsets.p, n, ctx->world3x4.p, total, flatScratch.p, scratchNodes.p, 0u,
                                    ctx->flatPrims.p, &nodeCount, &primCount))
            return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "flat BVH build failed: %s", ctx->builder.err);
        if (primCount != total) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "flat BVH emitted %u of %u triangles", primCount, total);
        // one parked node group per level (+ margin): the traversal kernel is chosen by this bound, deeper trees are refused
        plocKept += ctx->builder.lastBuilder;
        ctx->stackNeed = ctx->builder.lastLevels + 2u;
        if (ctx->stackNeed > (uint32_t)TRACE_STACK_DEEP)
            return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "flat BVH has %u levels: deeper than the traversal stack (%d)", ctx->builder.lastLevels, TRACE_STACK_DEEP);
        CU(ctx->nodes.alloc(std::max(nodeCount, 1u)));
        launchRelocateNodes(scratchNodes.p, ctx->nodes.p, nodeCount, 0u, 0u, st);
        CU(ctx->flatTriangles.alloc((size_t)std::max(total, 1u) * 3));
        launchFlatGatherTriangles(ctx->flatPrims.p, ctx->triangles.p, total, ctx->flatTriangles.p, st);
        CU(cudaEventRecord(ctx->evB, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        ctx->tlasRoot = 0;
        ctx->accelValid = true;
        ctx->anyTransmissive = anyTransmissive;
        vkrt_cuda_build_stats& s = ctx->buildStats;
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->evA, ctx->evB);
        s.blasMs = ms;
        s.tlasMs = 0.0f;
        s.buildMs = ms;
        s.uniqueGeometries = (uint32_t)blas.size();
        s.instanceCount = n;
        s.triangleCount = totalTris;
        s.instancedTriangleCount = instancedTris;
        s.bvh8NodeCount = nodeCount;
        s.accelBytes = (uint64_t)nodeCount * sizeof(Bvh8Node) + totalTris * 48ull + instancedTris * 8ull + (uint64_t)n * sizeof(InstanceRecord);
        s.flat = 1;
        s.plocHierarchies = plocKept;
        if (outStats) *outStats = s;
        return VKRT_SUCCESS;
    }
    // worst-case node budget: one 8-wide node per binary internal node; trimmed after the build
    uint64_t nodeBudget = 0;
    for (auto& b : blas) { b.nodeBase = (uint32_t)nodeBudget; nodeBudget += std::max(b.triCount, 1u); }
    const uint32_t tlasNodeBase = (uint32_t)nodeBudget;
    nodeBudget += std::max(n, 1u);
    DevBuf<Bvh8Node> scratchNodes;
    CU(scratchNodes.alloc(nodeBudget));
    CU(ctx->triangles.alloc((size_t)std::max<uint64_t>(totalTris, 1) * 3));
    CU(ctx->blasBounds.alloc(std::max<size_t>(blas.size(), 1) * 2));
    cudaStream_t st = ctx->stream;
    CU(cudaEventRecord(ctx->evA, st));
    uint32_t primBase = 0, maxBlasLevels = 0;
    std::vector<::float4> emptyBounds = {make_float4(1e30f, 1e30f, 1e30f, 0.f), make_float4(-1e30f, -1e30f, -1e30f, 0.f)};
    for (size_t b = 0; b < blas.size(); b++) {
        BlasDesc& d = blas[b];
        d.primBase = primBase;
        if (d.triCount == 0) {
            CU(cudaMemcpyAsync(ctx->blasBounds.p + 2 * b, emptyBounds.data(), sizeof(::float4) * 2, cudaMemcpyHostToDevice, st));
            continue;
        }
        uint32_t nodeCount = 0, primCount = 0;
        if (!ctx->builder.buildBlas(st, ctx->vertices.p, ctx->indices.p, d.vertexBase, d.indexBase, d.triCount, scratchNodes.p, d.nodeBase,
                                    ctx->triangles.p, d.primBase, ctx->blasBounds.p + 2 * b, &nodeCount, &primCount))
            return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "BLAS build failed: %s", ctx->builder.err);
        if (primCount != d.triCount) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "BLAS %zu emitted %u of %u triangles", b, primCount, d.triCount);
        d.nodeCount = nodeCount;
        maxBlasLevels = std::max(maxBlasLevels, ctx->builder.lastLevels);
        plocKept += ctx->builder.lastBuilder;
        primBase += d.triCount;
    }
    CU(cudaEventRecord(ctx->evB, st));
    // --- instance records ---
    std::vector<InstanceRecord> records(std::max(n, 1u));
    bool anyTransmissive = false;
    for (uint32_t i = 0; i < n; i++) {
        float inv[12];
        invertAffine3x4(&ctx->hostWorld[(size_t)i * 12], inv);
        InstanceRecord& r = records[i];
        r.inv0 = make_float4(inv[0], inv[1], inv[2], inv[3]);
        r.inv1 = make_float4(inv[4], inv[5], inv[6], inv[7]);
        r.inv2 = make_float4(inv[8], inv[9], inv[10], inv[11]);
        const BlasDesc& d = blas[instanceBlas[i]];
        r.blasRoot = d.nodeBase;  // rebased below once nodes are compacted
        r.flags = 0;
        if (ctx->hostAlpha[i]) r.flags |= INSTANCE_FLAG_ALPHA_TESTED;
        if (ctx->hostMaterials[ctx->hostMeshInfos[i].materialIndex].transmission > 0.0f) { r.flags |= INSTANCE_FLAG_TRANSMISSIVE; anyTransmissive = true; }
        if (d.triCount == 0) r.flags |= INSTANCE_FLAG_EMPTY;
        r.instanceIndex = i;
        r.pad = 0;
    }
    // compact node ranges: final layout = BLAS nodes back to back, then TLAS nodes
    std::vector<uint32_t> finalBase(blas.size());
    uint64_t finalNodes = 0;
    for (size_t b = 0; b < blas.size(); b++) { finalBase[b] = (uint32_t)finalNodes; finalNodes += blas[b].nodeCount; }
    // childBase fields inside nodes are absolute, so builds must target their final base: rebuild offsets by giving each BLAS
    // its final base up front is impossible before counts are known -> instead relocate: node indices are (base + local), and
    // every BLAS was built with nodeBase = d.nodeBase; relocation subtracts (d.nodeBase - finalBase) from childBase.
    CU(ctx->nodes.alloc(std::max<uint64_t>(finalNodes + std::max(n, 1u), 1)));
    for (uint32_t i = 0; i < n; i++) records[i].blasRoot = finalBase[instanceBlas[i]];
    CU(ctx->instanceRecords.upload(records.data(), records.size(), st));
    CU(ctx->instancesLeafOrder.alloc(std::max(n, 1u)));
    CU(ctx->instanceBlas.upload(instanceBlas.data(), instanceBlas.size(), st));
    for (size_t b = 0; b < blas.size(); b++) {
        if (!blas[b].nodeCount) continue;
        launchRelocateNodes(scratchNodes.p + blas[b].nodeBase, ctx->nodes.p + finalBase[b], blas[b].nodeCount, blas[b].nodeBase, finalBase[b], st);
    }
    cudaEvent_t evC;
    CU(cudaEventCreate(&evC));
    uint32_t tlasNodes = 0, tlasPrims = 0, tlasLevels = 0;
    ctx->tlasRoot = (uint32_t)finalNodes;
    if (n > 0) {
        if (!ctx->builder.buildTlas(st, ctx->blasBounds.p, ctx->instanceBlas.p, ctx->world3x4.p, ctx->instanceRecords.p, n, scratchNodes.p, tlasNodeBase,
                                    ctx->instancesLeafOrder.p, &tlasNodes, &tlasPrims)) {
            cudaEventDestroy(evC);
            return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "TLAS build failed: %s", ctx->builder.err);
        }
        launchRelocateNodes(scratchNodes.p + tlasNodeBase, ctx->nodes.p + ctx->tlasRoot, tlasNodes, tlasNodeBase, ctx->tlasRoot, st);
        tlasLevels = ctx->builder.lastLevels;
        plocKept += ctx->builder.lastBuilder;
    }
    // one parked node group per level, two entries parked when a ray enters an instance (+ margin)
    ctx->stackNeed = tlasLevels + 2u + maxBlasLevels + 2u;
    if (ctx->stackNeed > (uint32_t)TRACE_STACK_DEEP) {
        cudaEventDestroy(evC);
        return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "BVH has %u + %u levels: deeper than the traversal stack (%d)", tlasLevels, maxBlasLevels, TRACE_STACK_DEEP);
    }
    CU(cudaEventRecord(evC, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    float blasMs = 0, tlasMs = 0;
    cudaEventElapsedTime(&blasMs, ctx->evA, ctx->evB);
    cudaEventElapsedTime(&tlasMs, ctx->evB, evC);
    cudaEventDestroy(evC);
    ctx->accelValid = true;
    ctx->anyTransmissive = anyTransmissive;
    vkrt_cuda_build_stats& s = ctx->buildStats;
    s.blasMs = blasMs;
    s.tlasMs = tlasMs;
    s.buildMs = blasMs + tlasMs;
    s.uniqueGeometries = (uint32_t)blas.size();
    s.instanceCount = n;
    s.triangleCount = totalTris;
    s.instancedTriangleCount = instancedTris;
    s.bvh8NodeCount = finalNodes + tlasNodes;
    s.accelBytes = s.bvh8NodeCount * sizeof(Bvh8Node) + totalTris * 48ull + (uint64_t)n * sizeof(InstanceRecord);
    s.flat = 0;
    s.plocHierarchies = plocKept;
    if (outStats) *outStats = s;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_invalidate_accel(vkrt_cuda_ctx* ctx) {
    if (!ctx) return VKRT_ERROR_INVALID_ARGUMENT;
    ctx->accelValid = false;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_resize(vkrt_cuda_ctx* ctx, uint32_t width, uint32_t height) {
    if (!ctx || width == 0 || height == 0 || width > 16384 || height > 16384) return VKRT_ERROR_INVALID_ARGUMENT;  // render.c:252-253
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->width = width;
    ctx->height = height;
    vkrt_tile_layout lay;
    vkrt_tile_layout_init(&lay, width, height, ctx->tileW, ctx->tileH, ctx->rank, ctx->worldSize);
    ctx->hostL2G.resize(std::max(lay.localTileCount, 1u));
    vkrt_tile_layout_local_tiles(&lay, ctx->hostL2G.data());
    CU(ctx->l2g.upload(ctx->hostL2G.data(), ctx->hostL2G.size(), ctx->stream));
    TileMap& tm = ctx->tiles;
    tm.width = width; tm.height = height; tm.tileW = lay.tileW; tm.tileH = lay.tileH; tm.tilesX = lay.tilesX; tm.tilesY = lay.tilesY;
    tm.localTileCount = lay.localTileCount;
    tm.localPixelCount = lay.localTileCount * lay.tileW * lay.tileH;
    tm.localToGlobalTile = ctx->l2g.p;
    if (tm.localPixelCount == 0) {
        // leave the context "not resized": a later render_frame must fail on its width guard, not divide by a zero pixel count
        ctx->width = ctx->height = 0;
        ctx->tiles = {};
        return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "rank %u owns no tiles of a %ux%u image (%u ranks, %ux%u tiles): use fewer ranks or smaller tiles",
                    ctx->rank, width, height, ctx->worldSize, lay.tileW, lay.tileH);
    }
    VKRT_Result r = allocateWavefront(ctx);
    if (r != VKRT_SUCCESS) {
        ctx->width = ctx->height = 0;
        return r;
    }
    if (ctx->worldSize > 1) {   // the un-tiling table of the gather: once per resize, not once per rank, AOV and call
        std::vector<uint32_t> local((size_t)lay.tilesX * lay.tilesY), next(ctx->worldSize, 0u);
        for (uint32_t ty = 0; ty < lay.tilesY; ty++)
            for (uint32_t tx = 0; tx < lay.tilesX; tx++) local[(size_t)ty * lay.tilesX + tx] = next[vkrt_tile_owner(&lay, tx, ty)]++;
        CU(ctx->tileLocalIndex.upload(local.data(), local.size(), ctx->stream));
    }
    ctx->readIndex = 0;
    r = resetAccumulation(ctx);
    if (r != VKRT_SUCCESS) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_reset_accumulation(vkrt_cuda_ctx* ctx) {
    if (!ctx) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    VKRT_Result r = resetAccumulation(ctx);
    if (r != VKRT_SUCCESS) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_render_frame_async(vkrt_cuda_ctx* ctx, const SceneData* sceneData) {
    if (!ctx || !sceneData) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    return enqueueFrame(ctx, sceneData, nullptr);
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_sync(vkrt_cuda_ctx* ctx) {
    if (!ctx) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_timer_begin(vkrt_cuda_ctx* ctx) {
    if (!ctx) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (!ctx->evT0) { CU(cudaEventCreate(&ctx->evT0)); CU(cudaEventCreate(&ctx->evT1)); }
    CU(cudaEventRecord(ctx->evT0, ctx->stream));
    return VKRT_SUCCESS;
}
VKRT_CUDA_API VKRT_Result vkrt_cuda_timer_end(vkrt_cuda_ctx* ctx, float* outMs) {
    if (!ctx || !outMs) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!ctx->evT0) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "timer_end without timer_begin");
    cudaSetDevice(ctx->device);
    CU(cudaEventRecord(ctx->evT1, ctx->stream));
    CU(cudaEventSynchronize(ctx->evT1));
    CU(cudaEventElapsedTime(outMs, ctx->evT0, ctx->evT1));
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_render_frame(vkrt_cuda_ctx* ctx, const SceneData* sceneData, vkrt_cuda_frame_stats* outStats) {
    if (!ctx || !sceneData) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (ctx->flags & VKRT_CUDA_FLAG_COUNT_RAYS) CU(cudaMemsetAsync(ctx->stats.p, 0, sizeof(unsigned long long) * 4, ctx->stream));
    CU(cudaEventRecord(ctx->evA, ctx->stream));
    uint32_t launches = 0;
    VKRT_Result r = enqueueFrame(ctx, sceneData, &launches);
    if (r != VKRT_SUCCESS) return r;
    CU(cudaEventRecord(ctx->evB, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (outStats) {
        memset(outStats, 0, sizeof(*outStats));
        cudaEventElapsedTime(&outStats->frameMs, ctx->evA, ctx->evB);
        outStats->kernelLaunches = launches;
        for (size_t k = 1; k < ctx->stageUsed; k++) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ctx->stageEvents[k - 1], ctx->stageEvents[k]) != cudaSuccess) continue;
            if (ctx->stageKinds[k] == 1) { outStats->traceMs += ms; outStats->traceLaunches++; }
            else outStats->shadeMs += ms;
            if (ctx->stageKinds[k] == 2) { outStats->shadeKernelMs += ms; outStats->shadeLaunches++; }
        }
        const uint32_t spp = std::max(sceneData->samplesPerPixel, 1u);
        // counters hold the LAST chunk only; ray totals are exact when the frame fits one chunk (the common case)
        uint32_t host[MAX_DEPTH_SLOTS * 2];
        CU(cudaMemcpy(host, ctx->counters.p, sizeof(host), cudaMemcpyDeviceToHost));
        uint64_t ext = 0, sh = 0;
        for (uint32_t d = 0; d < MAX_DEPTH_SLOTS; d++) { ext += host[d]; sh += host[MAX_DEPTH_SLOTS + d]; }
        const uint32_t lpc = ctx->tiles.localPixelCount;
        const uint32_t samplesPerChunk = std::max(1u, std::min(spp, ctx->capacity / lpc));
        const uint32_t chunks = (spp + samplesPerChunk - 1) / samplesPerChunk;
        const uint32_t lastChunkSamples = spp - (chunks - 1) * samplesPerChunk;
        const double scale = (double)spp / (double)lastChunkSamples;
        outStats->extensionRays = (uint64_t)((double)ext * scale);
        outStats->shadowRays = (uint64_t)((double)sh * scale);
        outStats->paths = (uint64_t)((double)host[0] * scale);
        if (ctx->flags & VKRT_CUDA_FLAG_COUNT_RAYS) {
            unsigned long long st[4];
            CU(cudaMemcpy(st, ctx->stats.p, sizeof(st), cudaMemcpyDeviceToHost));
            outStats->nodesVisited = st[0];
            outStats->trianglesTested = st[1];
            outStats->instancesEntered = st[2];
        }
    }
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_trace_primary(vkrt_cuda_ctx* ctx, const SceneData* sceneData) {
    if (!ctx || !sceneData) return VKRT_ERROR_INVALID_ARGUMENT;
    ctx->lastScene = *sceneData;
    ctx->haveScene = true;
    return VKRT_SUCCESS;
}

static VKRT_Result tracePrimary(vkrt_cuda_ctx* ctx, int jittered) {
    if (!ctx->accelValid || !ctx->haveScene || !ctx->width) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "hit-id AOV needs a built scene, a size and a camera");
    fillFrameParams(ctx, ctx->lastScene);
    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(uint32_t) * MAX_DEPTH_SLOTS * 3, st));
    launchPrimaryRaygen(ctx->fp, jittered, ctx->smCount * 4, st);
    launchTrace(makeTraceParams(ctx, 0, true, false), false, ctx->traceGrid, st);
    launchPrimaryStore(ctx->fp, ctx->smCount * 4, st);
    CU(cudaGetLastError());
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_read_accum_samples(vkrt_cuda_ctx* ctx, const uint32_t* xy, uint32_t count, float* outRgba) {
    if (!ctx || !xy || !outRgba) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!ctx->width) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "read_accum_samples before resize");
    if (count == 0) return VKRT_SUCCESS;
    cudaSetDevice(ctx->device);
    CU(ctx->staging.alloc((size_t)count * 24 + 16));   // xy pairs, then RGBA
    ::uint2* dxy = reinterpret_cast<::uint2*>(ctx->staging.p);
    ::float4* dout = reinterpret_cast<::float4*>(reinterpret_cast<char*>(ctx->staging.p) + (((size_t)count * 8 + 15) & ~(size_t)15));
    CU(cudaMemcpyAsync(dxy, xy, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    launchProbeAccum(ctx->fp.film.accum[ctx->readIndex], ctx->tiles, ctx->l2g.p, dxy, count, dout, ctx->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(outRgba, dout, (size_t)count * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_read_aov(vkrt_cuda_ctx* ctx, vkrt_cuda_aov which, void* dst, size_t bytes) {
    if (!ctx || !dst) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!ctx->width) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "read_aov before resize");
    cudaSetDevice(ctx->device);
    const size_t px = (size_t)ctx->width * ctx->height;
    const void* src = nullptr;
    uint32_t words = 0;
    const Film& film = ctx->fp.film;
    switch (which) {
        case VKRT_CUDA_AOV_ACCUM_RGBA32F: src = film.accum[ctx->readIndex]; words = 4; break;
        case VKRT_CUDA_AOV_ALBEDO_RGBA16F: src = film.albedo[ctx->readIndex]; words = 2; break;
        case VKRT_CUDA_AOV_NORMAL_RGBA16F: src = film.normal[ctx->readIndex]; words = 2; break;
        case VKRT_CUDA_AOV_OUTPUT_RGBA16: src = film.output; words = 2; break;
        case VKRT_CUDA_AOV_HITID_CENTER:
        case VKRT_CUDA_AOV_HITID_S0: {
            VKRT_Result r = tracePrimary(ctx, which == VKRT_CUDA_AOV_HITID_S0);
            if (r != VKRT_SUCCESS) return r;
            src = film.hitId; words = 2; break;
        }
        case VKRT_CUDA_AOV_HIT_T_UV_CENTER: {
            VKRT_Result r = tracePrimary(ctx, 0);
            if (r != VKRT_SUCCESS) return r;
            src = film.hitTuv; words = 4; break;
        }
        default: return VKRT_ERROR_INVALID_ARGUMENT;
    }
    const uint32_t outWords = which == VKRT_CUDA_AOV_HIT_T_UV_CENTER ? 3u : words;
    if (bytes != px * outWords * 4) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "read_aov: expected %zu bytes, got %zu", px * outWords * 4, bytes);
    const int slot = which <= VKRT_CUDA_AOV_OUTPUT_RGBA16 ? (int)which : -1;
    if (slot >= 0 && ctx->filmIsFullFrame[slot]) {  // rank 0 after a gather
        CU(cudaMemcpyAsync(dst, ctx->fullFrame[slot].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return VKRT_SUCCESS;
    }
    CU(ctx->staging.alloc(px * words * 4));
    CU(cudaMemsetAsync(ctx->staging.p, 0, px * words * 4, ctx->stream));
    launchUntile(src, ctx->staging.p, ctx->tiles, ctx->l2g.p, words, ctx->smCount * 4, ctx->stream);
    if (outWords == words) {
        CU(cudaMemcpyAsync(dst, ctx->staging.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    } else {  // drop the 4th word of each pixel on the host
        std::vector<float> tmp(px * 4);
        CU(cudaMemcpyAsync(tmp.data(), ctx->staging.p, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float* o = (float*)dst;
        for (size_t i = 0; i < px; i++) { o[i * 3] = tmp[i * 4]; o[i * 3 + 1] = tmp[i * 4 + 1]; o[i * 3 + 2] = tmp[i * 4 + 2]; }
    }
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_trace_rays(vkrt_cuda_ctx* ctx, const float* rays, uint32_t rayCount, int anyHit, uint32_t* hits, float* outKernelMs) {
    if (!ctx || !rays || !hits) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!ctx->accelValid) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "trace_rays: acceleration structure not built");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    DevBuf<::float4> o, d;
    DevBuf<::uint4> hA;
    DevBuf<float> hB;
    DevBuf<uint32_t> res, ctr;
    CU(o.alloc(rayCount)); CU(d.alloc(rayCount)); CU(hA.alloc(rayCount)); CU(hB.alloc(rayCount)); CU(res.alloc(rayCount)); CU(ctr.alloc(4));
    std::vector<::float4> ho(rayCount), hd(rayCount);
    for (uint32_t i = 0; i < rayCount; i++) {
        ho[i] = make_float4(rays[i * 8 + 0], rays[i * 8 + 1], rays[i * 8 + 2], rays[i * 8 + 3]);
        hd[i] = make_float4(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6], rays[i * 8 + 7]);
    }
    CU(cudaMemcpyAsync(o.p, ho.data(), sizeof(::float4) * rayCount, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d.p, hd.data(), sizeof(::float4) * rayCount, cudaMemcpyHostToDevice, st));
    uint32_t hostCtr[4] = {rayCount, 0u, 0u, 0u};  // [0] count, [1] work counter
    CU(cudaMemcpyAsync(ctr.p, hostCtr, sizeof(hostCtr), cudaMemcpyHostToDevice, st));
    TraceParams tp = {};
    tp.scene = makeSceneView(ctx);
    if (anyHit) {
        tp.shO = o.p; tp.shD = d.p; tp.shCount = ctr.p; tp.shadowResult = res.p; tp.shSeed = res.p /* unused unless alpha */;
    } else {
        tp.rayO = o.p; tp.rayD = d.p; tp.hitA = hA.p; tp.hitB = hB.p; tp.extCount = ctr.p;
    }
    tp.workCounter = ctr.p + 1;
    tp.slotStride = 1u;
    CU(cudaEventRecord(ctx->evA, st));
    launchTrace(tp, false, ctx->traceGrid, st);
    CU(cudaEventRecord(ctx->evB, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (outKernelMs) cudaEventElapsedTime(outKernelMs, ctx->evA, ctx->evB);
    if (anyHit) {
        std::vector<uint32_t> r(rayCount);
        CU(cudaMemcpy(r.data(), res.p, sizeof(uint32_t) * rayCount, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < rayCount; i++) { hits[i * 5] = r[i]; hits[i * 5 + 1] = hits[i * 5 + 2] = hits[i * 5 + 3] = hits[i * 5 + 4] = 0; }
    } else {
        std::vector<::uint4> a(rayCount);
        std::vector<float> b(rayCount);
        CU(cudaMemcpy(a.data(), hA.p, sizeof(::uint4) * rayCount, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(b.data(), hB.p, sizeof(float) * rayCount, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < rayCount; i++) {
            hits[i * 5] = a[i].x; hits[i * 5 + 1] = a[i].y; hits[i * 5 + 2] = a[i].z; hits[i * 5 + 3] = a[i].w;
            memcpy(&hits[i * 5 + 4], &b[i], 4);
        }
    }
    return VKRT_SUCCESS;
}

// Test entry (include/vkrt_closure.h): closures evaluated on the device by the functions k_shade uses.
VKRT_CUDA_API VKRT_Result vkrt_cuda_eval_closures(vkrt_cuda_ctx* ctx, const vkrt_closure_query* queries, uint32_t count, vkrt_closure_result* results) {
    if (!ctx || (count && (!queries || !results))) return VKRT_ERROR_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < count; i++)
        if (queries[i].mode > 2u) return fail(ctx, VKRT_ERROR_INVALID_ARGUMENT, "eval_closures: query %u has mode %u", i, queries[i].mode);
    bool spectral = false;
    for (uint32_t i = 0; i < count; i++) spectral |= queries[i].mode != 0u;
    if (spectral && !ctx->haveRgb2spec) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "eval_closures: spectral queries need vkrt_cuda_set_rgb2spec");
    cudaSetDevice(ctx->device);
    DevBuf<vkrt_closure_query> q;
    DevBuf<vkrt_closure_result> r;
    CU(q.upload(queries, count, ctx->stream));
    CU(r.alloc(count));
    launchEvalClosures(makeSceneView(ctx), q.p, count, r.p, ctx->stream);
    CU(cudaGetLastError());
    if (count) CU(cudaMemcpyAsync(results, r.p, sizeof(vkrt_closure_result) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VKRT_SUCCESS;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------------
static bool loadNccl(NcclApi& n) {
    if (n.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) return false;
    n.GetUniqueId = (int (*)(void*))dlsym(n.lib, "ncclGetUniqueId");
    n.CommInitRank = (int (*)(void**, int, char[128], int))dlsym(n.lib, "ncclCommInitRank");
    n.CommDestroy = (int (*)(void*))dlsym(n.lib, "ncclCommDestroy");
    n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
    n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
    n.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclSend");
    n.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclRecv");
    return n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv;
}

struct NcclIdByValue { char bytes[128]; };

VKRT_CUDA_API VKRT_Result vkrt_cuda_nccl_unique_id(void* outId128) {
    if (!outId128) return VKRT_ERROR_INVALID_ARGUMENT;
    NcclApi n;
    if (!loadNccl(n)) return VKRT_ERROR_INITIALIZATION_FAILED;
    return n.GetUniqueId(outId128) == 0 ? VKRT_SUCCESS : VKRT_ERROR_OPERATION_FAILED;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_comm_init(vkrt_cuda_ctx* ctx, const void* uniqueId128) {
    if (!ctx || !uniqueId128) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!loadNccl(ctx->nccl)) return fail(ctx, VKRT_ERROR_INITIALIZATION_FAILED, "libnccl.so.2 not found: %s", dlerror());
    cudaSetDevice(ctx->device);
    NcclIdByValue id;
    memcpy(id.bytes, uniqueId128, 128);
    typedef int (*InitFn)(void**, int, NcclIdByValue, int);
    InitFn init = (InitFn)dlsym(ctx->nccl.lib, "ncclCommInitRank");
    int rc = init(&ctx->comm, (int)ctx->worldSize, id, (int)ctx->rank);
    if (rc != 0) return fail(ctx, VKRT_ERROR_INITIALIZATION_FAILED, "ncclCommInitRank failed (%d)", rc);
    return VKRT_SUCCESS;
}

VKRT_CUDA_API uint64_t vkrt_cuda_max_local_pixels(const vkrt_cuda_ctx* ctx) {
    if (!ctx || !ctx->width) return 0;
    vkrt_tile_layout lay;
    uint64_t m = 0;
    for (uint32_t r = 0; r < ctx->worldSize; r++) {
        vkrt_tile_layout_init(&lay, ctx->width, ctx->height, ctx->tileW, ctx->tileH, r, ctx->worldSize);
        m = std::max<uint64_t>(m, (uint64_t)lay.localTileCount * lay.tileW * lay.tileH);
    }
    return m;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_local_film(vkrt_cuda_ctx* ctx, vkrt_cuda_aov which, void** outDevicePtr, uint64_t* outBytes, uint64_t* outLocalPixelCount) {
    if (!ctx || !outDevicePtr || !ctx->width) return VKRT_ERROR_INVALID_ARGUMENT;
    const Film& film = ctx->fp.film;
    uint32_t words = 2;
    switch (which) {
        case VKRT_CUDA_AOV_ACCUM_RGBA32F: *outDevicePtr = film.accum[ctx->readIndex]; words = 4; break;
        case VKRT_CUDA_AOV_ALBEDO_RGBA16F: *outDevicePtr = film.albedo[ctx->readIndex]; break;
        case VKRT_CUDA_AOV_NORMAL_RGBA16F: *outDevicePtr = film.normal[ctx->readIndex]; break;
        case VKRT_CUDA_AOV_OUTPUT_RGBA16: *outDevicePtr = film.output; break;
        default: return VKRT_ERROR_INVALID_ARGUMENT;
    }
    if (outBytes) *outBytes = (uint64_t)ctx->tiles.localPixelCount * words * 4;
    if (outLocalPixelCount) *outLocalPixelCount = ctx->tiles.localPixelCount;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_import_gathered(vkrt_cuda_ctx* ctx, vkrt_cuda_aov which, const void* deviceGathered) {
    if (!ctx || !deviceGathered || !ctx->width || which > VKRT_CUDA_AOV_OUTPUT_RGBA16) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    const uint32_t words = which == VKRT_CUDA_AOV_ACCUM_RGBA32F ? 4u : 2u;
    const size_t px = (size_t)ctx->width * ctx->height;
    const uint64_t stride = vkrt_cuda_max_local_pixels(ctx);
    CU(ctx->fullFrame[which].alloc(px * words * 4));
    if (ctx->worldSize > 1) {
        launchUntileAll(deviceGathered, ctx->fullFrame[which].p, ctx->tiles, ctx->worldSize, ctx->tileLocalIndex.p, stride * words, words, ctx->smCount * 8, ctx->stream);
    } else {
        launchUntile(deviceGathered, ctx->fullFrame[which].p, ctx->tiles, ctx->l2g.p, words, ctx->smCount * 4, ctx->stream);
    }
    CU(cudaGetLastError());
    ctx->filmIsFullFrame[which] = true;
    return VKRT_SUCCESS;
}

VKRT_CUDA_API VKRT_Result vkrt_cuda_gather(vkrt_cuda_ctx* ctx, float* outGatherMs) { return vkrt_cuda_gather_aovs(ctx, 0xFu, outGatherMs); }

VKRT_CUDA_API VKRT_Result vkrt_cuda_gather_aovs(vkrt_cuda_ctx* ctx, uint32_t aovMask, float* outGatherMs) {
    if (!ctx || !ctx->width || (aovMask & ~0xFu)) return VKRT_ERROR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (outGatherMs) *outGatherMs = 0.0f;
    if (ctx->worldSize == 1) return VKRT_SUCCESS;
    if (!ctx->comm) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "gather: vkrt_cuda_comm_init not called");
    const uint64_t stride = vkrt_cuda_max_local_pixels(ctx);
    const vkrt_cuda_aov aovs[4] = {VKRT_CUDA_AOV_ACCUM_RGBA32F, VKRT_CUDA_AOV_ALBEDO_RGBA16F, VKRT_CUDA_AOV_NORMAL_RGBA16F, VKRT_CUDA_AOV_OUTPUT_RGBA16};
    CU(cudaEventRecord(ctx->evA, ctx->stream));
    for (vkrt_cuda_aov which : aovs) {
        if (!(aovMask & (1u << (uint32_t)which))) continue;
        const uint32_t words = which == VKRT_CUDA_AOV_ACCUM_RGBA32F ? 4u : 2u;
        void* local = nullptr;
        uint64_t bytes = 0;
        vkrt_cuda_local_film(ctx, which, &local, &bytes, nullptr);
        if (ctx->rank == 0) CU(ctx->gathered.alloc((size_t)stride * words * 4 * ctx->worldSize));
        ctx->nccl.GroupStart();
        if (ctx->rank == 0) {
            CU(cudaMemcpyAsync(ctx->gathered.p, local, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            for (uint32_t r = 1; r < ctx->worldSize; r++) {
                vkrt_tile_layout lay;
                vkrt_tile_layout_init(&lay, ctx->width, ctx->height, ctx->tileW, ctx->tileH, r, ctx->worldSize);
                size_t rb = (size_t)lay.localTileCount * lay.tileW * lay.tileH * words * 4;
                ctx->nccl.Recv(ctx->gathered.p + (size_t)r * stride * words * 4, rb, /*ncclUint8*/ 1, (int)r, ctx->comm, ctx->stream);
            }
        } else {
            ctx->nccl.Send(local, bytes, /*ncclUint8*/ 1, 0, ctx->comm, ctx->stream);
        }
        int rc = ctx->nccl.GroupEnd();
        if (rc != 0) return fail(ctx, VKRT_ERROR_OPERATION_FAILED, "nccl gather failed (%d)", rc);
        if (ctx->rank == 0) {
            VKRT_Result r = vkrt_cuda_import_gathered(ctx, which, ctx->gathered.p);
            if (r != VKRT_SUCCESS) return r;
        }
    }
    CU(cudaEventRecord(ctx->evB, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (outGatherMs) cudaEventElapsedTime(outGatherMs, ctx->evA, ctx->evB);
    return VKRT_SUCCESS;
}

}  // extern "C"
