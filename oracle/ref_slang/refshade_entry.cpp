// refshade_entry.cpp -- TEST INFRASTRUCTURE (oracle/). Builds oracle/_ref/libvkrt_refshade.so: the REFERENCE'S OWN shader sources
// (/root/reference/src/shaders/**/*.slang, transliterated to C++ at build time by slang2cpp.py into oracle/_ref/shaders_gen.inc, never
// committed) compiled with g++ and run on the CPU. This is the reference-held pin of the oracle: every closure, sampler, MIS weight,
// integrator loop and film write-back below `refslang::` is the upstream text, not a restatement.
//
// What is NOT reference code here, because upstream gets it from the Vulkan driver and the repository holds no source for it
// (SURVEY §0.2): TraceRay (BVH traversal + ray/triangle test), the bilinear texture sampler, and the storage-image format
// conversions. Those three are served by the oracle's stand-ins (oracle.cpp: traceRay, sampleTextureBilinear, f32<->f16 / unorm16),
// so a difference between this library and the oracle isolates the oracle's restatement of the SHADERS.
//
// The library exports the whole oracle_* API (scene upload, accel build, AOV read-back: it includes oracle.cpp) plus refshade_*:
// the reference's raygen entry points run per pixel over the oracle context's film, and per-call wrappers of the reference's BSDF /
// sampling / colour functions for the known-answer and closure-parity tests.
#include "../oracle.cpp"
#include "hlsl_prelude.h"

#include "../../include/vkrt_closure.h"

namespace refslang {

// ---- driver-side objects of the binding table (scene/resources.slang) --------------------------------------------------------------
struct RaytracingAccelerationStructure {};
struct BuiltInTriangleIntersectionAttributes {
    float2 barycentrics;
};
static const uint RAY_FLAG_NONE = 0u;
static const uint RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH = 4u;

enum ImageFormat { IMG_RGBA32F, IMG_RGBA16F, IMG_RGBA16_UNORM, IMG_R32UI };
template <class T>
struct RWTexture2D;
template <>
struct RWTexture2D<float4> {
    void* data = nullptr;
    uint width = 0;
    ImageFormat format = IMG_RGBA32F;
    struct Texel {
        const RWTexture2D* img;
        size_t index;
        operator float4() const {
            if (img->format == IMG_RGBA32F) {
                const float* p = static_cast<const float*>(img->data) + index * 4;
                return float4(p[0], p[1], p[2], p[3]);
            }
            const uint16_t* p = static_cast<const uint16_t*>(img->data) + index * 4;
            if (img->format == IMG_RGBA16F) return float4(orc::f16_to_f32(p[0]), orc::f16_to_f32(p[1]), orc::f16_to_f32(p[2]), orc::f16_to_f32(p[3]));
            return float4(float(p[0]), float(p[1]), float(p[2]), float(p[3])) * (1.0f / 65535.0f);
        }
        void operator=(float4 v) const {
            if (img->format == IMG_RGBA32F) {
                float* p = static_cast<float*>(img->data) + index * 4;
                p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
                return;
            }
            uint16_t* p = static_cast<uint16_t*>(img->data) + index * 4;
            if (img->format == IMG_RGBA16F) orc::storeHalf4(p, orc::float4(v.x, v.y, v.z, v.w));
            else orc::storeUnorm16x4(p, orc::float4(v.x, v.y, v.z, v.w));
        }
    };
    Texel operator[](int2 p) const { return Texel{this, size_t(p.y) * width + size_t(p.x)}; }
};
template <>
struct RWTexture2D<uint> {
    uint dummy = 0;
    uint& operator[](int2) { return dummy; }
};
struct SamplerState {
    uint variant = 0;  // axisVariantU * 3 + axisVariantV (material/textures.slang:16-31): 0 repeat, 1 clamp, 2 mirror
};
static const orc::Ctx* g_ctx = nullptr;  // the oracle context whose scene / textures / film the shaders are bound to
template <class T>
struct Texture2D;
template <>
struct Texture2D<float4> {
    uint index = 0;
    float4 SampleLevel(SamplerState s, float2 uv, float) const {
        static const uint32_t wrapOf[3] = {VKRT_TEXTURE_WRAP_REPEAT, VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE, VKRT_TEXTURE_WRAP_MIRRORED_REPEAT};
        const orc::float4 t = orc::sampleTextureBilinear(*g_ctx, index, orc::float2(uv.x, uv.y), wrapOf[(s.variant / 3u) % 3u], wrapOf[s.variant % 3u]);
        return float4(t.x, t.y, t.z, t.w);
    }
};
inline uint NonUniformResourceIndex(uint i) { return i; }

// ---- per-invocation built-ins ------------------------------------------------------------------------------------------------------
struct Invocation {
    uint3 launchIndex;
    uint instanceIndex = 0, primitiveIndex = 0;
    float rayT = 0.0f;
    bool ignoreHit = false;
};
static thread_local Invocation tl;
inline uint3 DispatchRaysIndex() { return tl.launchIndex; }
inline uint InstanceIndex() { return tl.instanceIndex; }
inline uint PrimitiveIndex() { return tl.primitiveIndex; }
inline float RayTCurrent() { return tl.rayT; }
inline void IgnoreHit() { tl.ignoreHit = true; }

struct SceneRayPayload;
struct ShadowPayload;
void TraceRay(RaytracingAccelerationStructure, uint flags, uint mask, uint sbtOffset, uint sbtStride, uint missIndex, RayDesc ray, SceneRayPayload& payload);
void TraceRay(RaytracingAccelerationStructure, uint flags, uint mask, uint sbtOffset, uint sbtStride, uint missIndex, RayDesc ray, ShadowPayload& payload);

#include "../_ref/shaders_gen.inc"

// ---- TraceRay: the oracle's traversal + the reference's hit / miss / any-hit shaders -------------------------------------------------
static bool sceneAnyHit(void* user, uint32_t inst, uint32_t prim, float t, float u, float v) {
    SceneRayPayload& payload = *static_cast<SceneRayPayload*>(user);
    tl.instanceIndex = inst; tl.primitiveIndex = prim; tl.rayT = t; tl.ignoreHit = false;
    BuiltInTriangleIntersectionAttributes attr;
    attr.barycentrics = float2(u, v);
    pathAnyHitMain(payload, attr);
    return !tl.ignoreHit;
}
static bool shadowAnyHit(void* user, uint32_t inst, uint32_t prim, float t, float u, float v) {
    ShadowPayload& payload = *static_cast<ShadowPayload*>(user);
    tl.instanceIndex = inst; tl.primitiveIndex = prim; tl.rayT = t; tl.ignoreHit = false;
    BuiltInTriangleIntersectionAttributes attr;
    attr.barycentrics = float2(u, v);
    shadowAnyHitMain(payload, attr);
    return !tl.ignoreHit;
}
static orc::Ray toOrcRay(const RayDesc& r) {
    orc::Ray o;
    o.origin = orc::float3(r.Origin.x, r.Origin.y, r.Origin.z);
    o.direction = orc::float3(r.Direction.x, r.Direction.y, r.Direction.z);
    o.tMin = r.TMin;
    o.tMax = r.TMax;
    return o;
}
void TraceRay(RaytracingAccelerationStructure, uint, uint, uint, uint, uint, RayDesc ray, SceneRayPayload& payload) {
    orc::AnyHitHook hook;
    hook.accept = sceneAnyHit;
    hook.user = &payload;
    orc::g_anyHitHook = &hook;
    const orc::HitRecord h = orc::traceRay(*g_ctx, toOrcRay(ray), payload.rng, false);
    orc::g_anyHitHook = nullptr;
    if (h.hit()) {
        tl.instanceIndex = h.instance; tl.primitiveIndex = h.primitive; tl.rayT = h.t;
        BuiltInTriangleIntersectionAttributes attr;
        attr.barycentrics = float2(h.u, h.v);
        pathClosestHitMain(payload, attr);
    } else {
        pathMissMain(payload);
    }
}
// Shadow rays: ACCEPT_FIRST_HIT_AND_END_SEARCH makes the reference's verdict depend on the driver's traversal order when a segment
// crosses both a transmissive and an opaque occluder; the oracle pins the order-independent reading (oracle.cpp, "Shadow rays"): an
// opaque occluder wins, else a transmissive one. The reference's shadow closest-hit shader then classifies the instance it is given.
void TraceRay(RaytracingAccelerationStructure, uint, uint, uint, uint, uint, RayDesc ray, ShadowPayload& payload) {
    orc::AnyHitHook hook;
    hook.accept = shadowAnyHit;
    hook.user = &payload;
    orc::g_anyHitHook = &hook;
    bool sawTransmissive = false;
    const orc::HitRecord h = orc::traceRay(*g_ctx, toOrcRay(ray), payload.rng, true, &sawTransmissive);
    orc::g_anyHitHook = nullptr;
    BuiltInTriangleIntersectionAttributes attr;
    if (h.hit()) {
        tl.instanceIndex = h.instance; tl.primitiveIndex = h.primitive;
        shadowClosestHitMain(payload, attr);
    } else if (sawTransmissive && hook.transmissiveInstance != 0xFFFFFFFFu) {
        tl.instanceIndex = hook.transmissiveInstance; tl.primitiveIndex = 0u;
        shadowClosestHitMain(payload, attr);
    } else {
        shadowMissMain(payload);
    }
}

static Selection g_selectionSlot;

// Binds the reference's descriptor set (scene/resources.slang:7-59) to the oracle context's buffers.
static void bindResources(orc::Ctx& c, const ::SceneData& sd, int writeIndex) {
    g_ctx = &c;
    static_assert(sizeof(SceneData) == sizeof(::SceneData) && sizeof(Material) == sizeof(::Material) && sizeof(MeshInfo) == sizeof(::MeshInfo) &&
                      sizeof(ShaderVertex) == sizeof(::ShaderVertex) && sizeof(EmissiveMesh) == sizeof(::EmissiveMesh) &&
                      sizeof(EmissiveTriangle) == sizeof(::EmissiveTriangle) && sizeof(Vertex) == sizeof(::Vertex),
                  "the reference's src/shared/types.h compiled as C++ must have the wire sizes of include/vkrt_shared.h");
    std::memcpy(&scene, &sd, sizeof(sd));
    // the table and its layout words are one upload (oracle_set_rgb2spec); the host mirrors the latter into SceneData (scene/rgb2spec.c:61-89)
    std::memcpy(&scene.rgb2specSRGB, &c.spectral.info, sizeof(scene.rgb2specSRGB));
    vertices.data = reinterpret_cast<const ShaderVertex*>(c.vertices.data());
    indices.data = c.indices.data();
    meshInfos.data = reinterpret_cast<const MeshInfo*>(c.meshInfos.data());
    materials.data = reinterpret_cast<const Material*>(c.materials.data());
    emissiveMeshes.data = reinterpret_cast<const EmissiveMesh*>(c.emissiveMeshes.data());
    emissiveTriangles.data = reinterpret_cast<const EmissiveTriangle*>(c.emissiveTriangles.data());
    meshAliasQ.data = c.meshAliasQ.data();
    meshAliasIdx.data = c.meshAliasIdx.data();
    triAliasQ.data = c.triAliasQ.data();
    triAliasIdx.data = c.triAliasIdx.data();
    rgb2specSRGBTable.data = c.rgb2spec.data();
    selection.data = &g_selectionSlot;
    static bool tablesBound = false;   // slot i of the bindless arrays is texture i / sampler variant i, once and for all
    if (!tablesBound) {
        for (uint i = 0; i < VKRT_MAX_BINDLESS_TEXTURES; i++) sceneTextures[i].index = i;
        for (uint i = 0; i < uint(sizeof(textureSamplers) / sizeof(textureSamplers[0])); i++) textureSamplers[i].variant = i;
        tablesBound = true;
    }
    const int r = c.readIndex, w = writeIndex;
    accumulationReadImage = {c.accum[r].data(), c.width, IMG_RGBA32F};
    accumulationWriteImage = {c.accum[w].data(), c.width, IMG_RGBA32F};
    albedoReadImage = {c.albedo[r].data(), c.width, IMG_RGBA16F};
    albedoWriteImage = {c.albedo[w].data(), c.width, IMG_RGBA16F};
    normalReadImage = {c.normal[r].data(), c.width, IMG_RGBA16F};
    normalWriteImage = {c.normal[w].data(), c.width, IMG_RGBA16F};
    outputImage = {c.output.data(), c.width, IMG_RGBA16_UNORM};
}

static void evalClosures(orc::Ctx& x, const vkrt_closure_query* q, uint32_t count, vkrt_closure_result* out) {
    ::SceneData sd = {};
    sd.rgb2specSRGB = x.spectral.info;
    bindResources(x, sd, 1 - x.readIndex);
    for (uint32_t i = 0; i < count; i++) {
        const vkrt_closure_query& Q = q[i];
        vkrt_closure_result R = {};
        Material m;
        std::memcpy(&m, &Q.material, sizeof(m));
        const BSDFMaterial bm = BSDFMaterial(m);
        const float3 wo(Q.wo[0], Q.wo[1], Q.wo[2]), wi(Q.wi[0], Q.wi[1], Q.wi[2]);
        const float4 wl(Q.wavelengths[0], Q.wavelengths[1], Q.wavelengths[2], Q.wavelengths[3]);
        // the spectral flag of the scene decides how colours are turned into scalars (utility/spectral.slang:20-27)
        scene.packedRenderSettings = VKRT_PACK_RENDER_SETTINGS(0u, Q.mode == 0u ? 0u : 1u, Q.mode == 2u ? 1u : 0u);
        BSDFState st = BSDFState(bm, wo, Q.frontFace, Q.mode == 0u ? 0.0f : wl.x, Q.mode == 0u ? 0u : 1u);
        ShadingBasis basis;
        basis.tangent = float3(1.0f, 0.0f, 0.0f);
        basis.bitangent = float3(0.0f, 1.0f, 0.0f);
        basis.normal = float3(0.0f, 0.0f, 1.0f);
        uint rng = Q.rng;
        if (Q.mode == 2u) {
            float4 tp(0.0f);
            const float4 v = evalSpectralBSDF(st, wi, wl, tp);
            R.evalValue[0] = v.x; R.evalValue[1] = v.y; R.evalValue[2] = v.z; R.evalValue[3] = v.w;
            R.evalPdf[0] = tp.x; R.evalPdf[1] = tp.y; R.evalPdf[2] = tp.z; R.evalPdf[3] = tp.w;
            SpectralBSDFSample s = sampleSpectralBSDF(st, basis, wl, rng);
            R.sampleWi[0] = s.wi.x; R.sampleWi[1] = s.wi.y; R.sampleWi[2] = s.wi.z;
            R.sampleWeight[0] = s.weight.x; R.sampleWeight[1] = s.weight.y; R.sampleWeight[2] = s.weight.z; R.sampleWeight[3] = s.weight.w;
            R.samplePdf[0] = s.techniquePdf.x; R.samplePdf[1] = s.techniquePdf.y; R.samplePdf[2] = s.techniquePdf.z; R.samplePdf[3] = s.techniquePdf.w;
            R.sampleFlags = (s.isUsable() ? 1u : 0u) | (s.isTransmission != 0u ? 2u : 0u);
        } else {
            const BSDFEval e = Q.mode == 0u ? evalBSDF(st, wi) : evalSingleWavelengthBSDF(st, wi);
            R.evalValue[0] = e.value.x; R.evalValue[1] = e.value.y; R.evalValue[2] = e.value.z;
            R.evalPdf[0] = e.pdf;
            BSDFSample s = sampleBSDF(st, basis, rng);
            R.sampleWi[0] = s.wi.x; R.sampleWi[1] = s.wi.y; R.sampleWi[2] = s.wi.z;
            R.sampleWeight[0] = s.weight.x; R.sampleWeight[1] = s.weight.y; R.sampleWeight[2] = s.weight.z;
            R.samplePdf[0] = s.pdf;
            R.sampleFlags = (s.isUsable() ? 1u : 0u) | (s.isTransmission != 0u ? 2u : 0u);
        }
        R.rngAfter = rng;
        out[i] = R;
    }
}


}  // namespace refslang

// ======================================================================================================================
// C API
// ======================================================================================================================
extern "C" {

ORC_API const char* refshade_version(void) { return "vkrt reference shaders (src/shaders/**/*.slang) transliterated by oracle/ref_slang/slang2cpp.py, g++ fp32"; }

// One frame = the reference's raygen entry point for the scene's render mode (entry/path/raygen_{rgb,spectral_single,spectral_hero}.slang)
// invoked for every pixel of rows [rowBegin, rowEnd) (0, 0 = the whole image), followed by the host's per-frame accumulation swap
// (src/core/api/frame.c:386-388), exactly like oracle_render_frame_rows.
ORC_API int refshade_render_frame_rows(oracle_ctx* ctx, const SceneData* sd, uint32_t rowBegin, uint32_t rowEnd) {
    if (!ctx || !sd) return -1;
    orc::Ctx& x = *reinterpret_cast<orc::Ctx*>(ctx);
    if (!x.accelBuilt && !orc::buildAccel(x)) return -2;
    if (x.width == 0 || x.height == 0) return -1;
    const uint32_t renderMode = VKRT_RENDER_SETTINGS_MODE(sd->packedRenderSettings);
    const uint32_t spectralSampling = VKRT_RENDER_SETTINGS_SPECTRAL(sd->packedRenderSettings);
    if (renderMode == VKRT_RENDER_MODE_SPECTRAL && !x.spectral.table) return -2;
    if (rowEnd == 0 || rowEnd > x.height) rowEnd = x.height;
    const int writeIndex = 1 - x.readIndex;
    if (rowBegin != 0 || rowEnd != x.height) {
        x.accum[writeIndex] = x.accum[x.readIndex];
        x.albedo[writeIndex] = x.albedo[x.readIndex];
        x.normal[writeIndex] = x.normal[x.readIndex];
    }
    refslang::bindResources(x, *sd, writeIndex);
    int nthreads = x.threads > 0 ? x.threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    orc::parallelFor((int64_t)rowBegin, (int64_t)rowEnd, 1, nthreads, [&](int64_t py, int) {
        for (uint32_t px = 0; px < x.width; px++) {
            refslang::tl.launchIndex = refslang::uint3(px, (uint32_t)py, 0u);
            if (renderMode != VKRT_RENDER_MODE_SPECTRAL) refslang::raygenRgbMain();
            else if (spectralSampling == VKRT_SPECTRAL_SAMPLING_MODE_HERO) refslang::raygenSpectralHeroMain();
            else refslang::raygenSpectralSingleMain();
        }
    });
    x.readIndex = writeIndex;
    x.lastScene = *sd;
    x.haveScene = true;
    return 0;
}
ORC_API int refshade_render_frame(oracle_ctx* ctx, const SceneData* sd) { return refshade_render_frame_rows(ctx, sd, 0, 0); }

// ---- closures: the reference's evalBSDF / evalSingleWavelengthBSDF / evalSpectralBSDF / sampleBSDF / sampleSpectralBSDF --------------------
// (bsdf/principled/eval_rgb.slang:104-140, eval_spectral.slang:96-115, bsdf/sample_rgb.slang:6, sample_spectral.slang:6) with the
// identity shading basis. ctx supplies the rgb2spec table (oracle_set_rgb2spec).
ORC_API int refshade_eval_closures(oracle_ctx* ctx, const vkrt_closure_query* q, uint32_t count, vkrt_closure_result* out) {
    if (!ctx || !q || !out) return -1;
    refslang::evalClosures(*reinterpret_cast<orc::Ctx*>(ctx), q, count, out);
    return 0;
}

// ---- known-answer wrappers (same signatures as the oracle_* ones in oracle.cpp) --------------------------------------------------------
ORC_API uint32_t refshade_hash(uint32_t v) { return refslang::hash(v); }
ORC_API uint32_t refshade_init_pixel_seed(int x, int y, uint32_t frame, uint32_t sample) { return refslang::initPixelSeed(refslang::int2(x, y), frame, sample); }
ORC_API float refshade_rand(uint32_t* rng) { return refslang::rand(*rng); }
ORC_API uint32_t refshade_reverse_bits(uint32_t v) { return refslang::reverseBits32(v); }
ORC_API float refshade_wavelength_unit(uint32_t* rng, uint32_t sampleIndex) { return refslang::sampleUniformWavelengthUnit(*rng, sampleIndex); }
ORC_API void refshade_unpack_normal(uint32_t packed, float* out3) { const refslang::float3 n = refslang::unpackOctNormal(packed); out3[0] = n.x; out3[1] = n.y; out3[2] = n.z; }
ORC_API void refshade_unpack_tangent(uint32_t packed, float* out4) { const refslang::float4 t = refslang::unpackOctTangent(packed); out4[0] = t.x; out4[1] = t.y; out4[2] = t.z; out4[3] = t.w; }
ORC_API void refshade_unpack_color(uint32_t packed, float* out4) { const refslang::float4 t = refslang::unpackColorRGBA8(packed); out4[0] = t.x; out4[1] = t.y; out4[2] = t.z; out4[3] = t.w; }
ORC_API void refshade_primary_ray(const SceneData* sd, int px, int py, float jx, float jy, float* out8) {
    std::memcpy(&refslang::scene, sd, sizeof(*sd));
    const refslang::RayDesc r = refslang::makePrimaryRay(refslang::int2(px, py), refslang::float2(jx, jy));
    out8[0] = r.Origin.x; out8[1] = r.Origin.y; out8[2] = r.Origin.z; out8[3] = r.TMin;
    out8[4] = r.Direction.x; out8[5] = r.Direction.y; out8[6] = r.Direction.z; out8[7] = r.TMax;
}
ORC_API void refshade_xyz_to_srgb(const float* xyz, float* rgb) { const refslang::float3 r = refslang::xyzToLinearSrgb(refslang::float3(xyz[0], xyz[1], xyz[2])); rgb[0] = r.x; rgb[1] = r.y; rgb[2] = r.z; }
ORC_API void refshade_spectral_xyz(float lambda, float* xyz) { const refslang::float3 r = refslang::spectralXYZ1931(lambda); xyz[0] = r.x; xyz[1] = r.y; xyz[2] = r.z; }
// spectralScalarFromLinearSrgb (utility/spectral.slang:37-48): rgb2spec lookup with the black and > 1 cases, as oracle_rgb2spec_eval
ORC_API float refshade_rgb2spec_eval(oracle_ctx* ctx, const float* rgb, float lambda) {
    orc::Ctx& x = *reinterpret_cast<orc::Ctx*>(ctx);
    ::SceneData sd = {};
    refslang::bindResources(x, sd, 1 - x.readIndex);
    return refslang::spectralScalarFromLinearSrgb(refslang::float3(rgb[0], rgb[1], rgb[2]), lambda);
}
ORC_API float refshade_power_heuristic(float a, float b) { return refslang::powerHeuristic(a, b); }
ORC_API float refshade_dispersive_ior(float ior, float abbe, float lambda) { return refslang::dispersiveIor(ior, abbe, lambda); }
ORC_API void refshade_map_to_display(const SceneData* sd, const float* rgb, float* out3) {
    std::memcpy(&refslang::scene, sd, sizeof(*sd));
    const refslang::float3 r = refslang::mapSceneColorToDisplay(refslang::float3(rgb[0], rgb[1], rgb[2]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
ORC_API uint32_t refshade_sample_alias(float u, uint32_t count, uint32_t offset, const float* q, const uint32_t* idx) {
    refslang::StructuredBuffer<float> aq; aq.data = q;
    refslang::StructuredBuffer<refslang::uint> ai; ai.data = idx;
    return refslang::sampleAlias(u, count, offset, aq, ai);
}

}  // extern "C"
