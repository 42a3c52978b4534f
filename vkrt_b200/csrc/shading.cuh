// shading.cuh — device restatement of vkrt's Slang shading library (src/shaders/{sampling,camera,utility,bsdf,film,geometry}).
// Every block cites the reference file:line it follows. Device-only code; the wavefront kernels in wavefront.cu call it.
#define VK_D __device__ __forceinline__
// Out-of-line on purpose: k_shade is instruction-fetch bound when the closure code is inlined at every call site
// (717 KB of SASS in hero mode, profiles/r01_shade_hero_baseline.txt); one copy per function keeps it I-cache resident.
#define VK_NOINLINE static __device__ __noinline__
#pragma once
#include "vmath.cuh"
#include "../../include/vkrt_shared.h"

namespace vk {

// ---- scene/constants.slang:6-15 ---------------------------------------------------------------------------------
static constexpr float RAY_T_MIN = 0.001f;
static constexpr float RAY_T_MAX = 10000.0f;
static constexpr float SHADOW_ORIGIN_OFFSET = 0.001f;
static constexpr float SHADOW_DISTANCE_OFFSET = 0.002f;
static constexpr float RR_MIN_CONTINUE_PROB = 0.05f;
static constexpr float RR_MAX_CONTINUE_PROB = 0.95f;
static constexpr float SAFE_NORMALIZE_EPS = 1e-12f;
static constexpr float TANGENT_PARALLEL_THRESHOLD = 0.999f;
static constexpr float PI = 3.14159265358979323846f;     // bsdf/base/math.slang:6
static constexpr float INV_PI = 0.31830988618379067154f; // bsdf/base/math.slang:7

// ---- sampling/random.slang:7-27 -----------------------------------------------------------------------------------
VK_D uint hash(uint value) {
    value ^= value >> 16;
    value *= 0x7feb352du;
    value ^= value >> 15;
    value *= 0x846ca68bu;
    value ^= value >> 16;
    return value;
}
VK_D float rand(uint& rng) {
    rng = hash(rng + 0x9e3779b9u);
    return float(rng & 0x00ffffffu) * (1.0f / 16777216.0f);
}
VK_D uint initPixelSeed(int px, int py, uint frameNumber, uint sampleIndex) {
    uint seed = uint(px) * 73856093u;
    seed ^= uint(py) * 19349663u;
    seed ^= frameNumber * 83492791u;
    seed ^= sampleIndex * 2654435761u;
    return hash(seed);
}

// ---- sampling/discrete.slang:8-18 ---------------------------------------------------------------------------------
VK_D float powerHeuristic(float pdfA, float pdfB) {
    float a2 = pdfA * pdfA;
    return a2 / (a2 + pdfB * pdfB);
}
VK_D uint sampleAlias(float u, uint count, uint offset, const float* aQ, const uint* aIdx) {
    float scaled = u * float(count);
    uint i = min(uint(scaled), count - 1u);
    float remainder = scaled - float(i);
    return (remainder < aQ[offset + i]) ? i : aIdx[offset + i];
}

// ---- sampling/wavelength.slang:10-47 ------------------------------------------------------------------------------
static constexpr float WAVELENGTH_MIN_NM = 360.0f;
static constexpr float WAVELENGTH_MAX_NM = 830.0f;
static constexpr float WAVELENGTH_RANGE_NM = WAVELENGTH_MAX_NM - WAVELENGTH_MIN_NM;
static constexpr float INV_UINT32 = 1.0f / 4294967296.0f;
struct WavelengthSample {
    float lambdaNm = 0.0f;
    float invPdf = 0.0f;
};
VK_D uint reverseBits32(uint value) {
    value = ((value & 0x55555555u) << 1u) | ((value >> 1u) & 0x55555555u);
    value = ((value & 0x33333333u) << 2u) | ((value >> 2u) & 0x33333333u);
    value = ((value & 0x0f0f0f0fu) << 4u) | ((value >> 4u) & 0x0f0f0f0fu);
    value = ((value & 0x00ff00ffu) << 8u) | ((value >> 8u) & 0x00ff00ffu);
    return (value << 16u) | (value >> 16u);
}
VK_D float sampleUniformWavelengthUnit(uint& rng, uint sampleIndex) {
    float radInvB2 = float(reverseBits32(sampleIndex)) * INV_UINT32;
    return frac(radInvB2 + rand(rng));
}
VK_D WavelengthSample sampleUniformWavelength(uint& rng, uint sampleIndex) {
    WavelengthSample s;
    float unitSample = sampleUniformWavelengthUnit(rng, sampleIndex);
    s.lambdaNm = WAVELENGTH_MIN_NM + saturate(unitSample) * WAVELENGTH_RANGE_NM;
    s.invPdf = WAVELENGTH_RANGE_NM;
    return s;
}
VK_D float4 sampleHeroWavelengths4(uint& rng, uint sampleIndex) {
    float unitSample = sampleUniformWavelengthUnit(rng, sampleIndex);
    return WAVELENGTH_MIN_NM + frac(unitSample + float4(0.0f, 0.25f, 0.5f, 0.75f)) * WAVELENGTH_RANGE_NM;
}

// ---- camera/ray.slang:32-39 ---------------------------------------------------------------------------------------
VK_D float3 safeNormalize(float3 value) {
    float lenSq = dot(value, value);
    if (lenSq <= SAFE_NORMALIZE_EPS) return float3(0.0f, 0.0f, 1.0f);
    return value * rsqrt(lenSq);
}

struct Ray {
    float3 origin;
    float3 direction;
    float tMin, tMax;
};

// Column-major 4x4 (cglm memory image) times column vector, terms summed left to right.
VK_D float4 mulMat4(const float* m, float4 v) {
    return float4(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w,
                  ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w,
                  ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w,
                  ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w);
}

// camera/ray.slang:15-30
VK_D Ray makePrimaryRay(const SceneData& scene, int px, int py, float2 jitter) {
    int ox = int(scene.viewportRect[0]), oy = int(scene.viewportRect[1]);
    float2 viewportSize = float2(float(scene.viewportRect[2]), float(scene.viewportRect[3]));
    float2 viewportPixel = float2(float(px - ox), float(py - oy)) + 0.5f + jitter;
    float2 uv = viewportPixel / viewportSize;
    float2 ndc = uv * 2.0f - 1.0f;
    float4 viewDir = mulMat4(scene.projInverse, float4(ndc.x, ndc.y, 1.0f, 1.0f));
    Ray r;
    r.origin = mulMat4(scene.viewInverse, float4(0.0f, 0.0f, 0.0f, 1.0f)).xyz();
    r.direction = normalize(mulMat4(scene.viewInverse, float4(viewDir.xyz(), 0.0f)).xyz());
    r.tMin = RAY_T_MIN;
    r.tMax = RAY_T_MAX;
    return r;
}

// camera/viewport.slang:6-16
VK_D bool insideViewport(const SceneData& scene, int px, int py) {
    int ox = int(scene.viewportRect[0]), oy = int(scene.viewportRect[1]);
    int sx = int(scene.viewportRect[2]), sy = int(scene.viewportRect[3]);
    if (sx <= 0 || sy <= 0) return false;
    return px >= ox && py >= oy && px < ox + sx && py < oy + sy;
}

// ---- utility/color.slang:6 ----------------------------------------------------------------------------------------
VK_D float linearSrgbLuminance(float3 rgb) { return dot(rgb, float3(0.2126f, 0.7152f, 0.0722f)); }

// ---- utility/rgb2spec.slang:14-90 ---------------------------------------------------------------------------------
// Device layout of the coefficient table (written once by vkrt_cuda_set_rgb2spec, k_pack_rgb2spec): `cells` holds one float4 per
// grid cell {c0, c1, c2, 0} in the payload's cell order, so the 8 corners of a lookup are 8 aligned 128-bit loads (two neighbours along
// x share a 32-byte sector) instead of 24 scalar loads; `scale` is the res-entry scale axis, which k_shade stages in shared memory.
struct SpectralMemoEntry {
    ::float4 key;     // xyz = the linear sRGB colour the value was computed for
    ::float4 value;   // {c0, c1, c2} of rgb / scale, w = scale (spectralScalarFromLinearSrgb4), 0 when the colour is black
};
enum : uint { SPECTRAL_MEMO_DIFFUSE = 0u, SPECTRAL_MEMO_EMISSION = 1u, SPECTRAL_MEMO_SLOTS = 2u };
struct SpectralTables {
    RGB2SpecTableInfo info = {0, 0, 0};
    const float* table = nullptr;      // the payload as uploaded (rgb2spec.c:17-59)
    const ::float4* cells = nullptr;   // 3 * res^3 packed coefficient cells
    const float* scale = nullptr;      // table + scaleOffset, or its shared-memory copy
    // optional (k_shade, res = 64): for q in [0, 1024), the interval index of x = q / 1024. A search starts there and walks up; it ends where
    // the full search ends (rgb2specFindInterval), after 0 - 1 steps for all but near-black inputs, instead of 14 loads and compares.
    const unsigned char* intervalLut = nullptr;
    // Memoised upsamplings (k_spectral_memo): colours that are constants of a material or of a light are looked up in the table once per
    // scene edit, not once per path vertex. An entry is used only when its key equals the colour at hand bit for bit, so a texture or a
    // vertex colour that changes the colour simply misses; nullptr = no memo (closure test entry, RGB frames).
    const SpectralMemoEntry* materialMemo = nullptr;   // SPECTRAL_MEMO_SLOTS entries per material
    const SpectralMemoEntry* emissiveMemo = nullptr;   // one entry per emissive mesh (EmissiveMesh::emission)
};
static constexpr float RGB2SPEC_EPSILON = 1e-8f;
static constexpr uint RGB2SPEC_SMEM_RES = 64u;

// rgb2spec.slang:14-30: the largest index in [0, res-2] whose scale entry is <= x. The scale axis is non-decreasing, so that index is
// the NUMBER of entries k in [1, res-2] with scale[k] <= x: for the standard res = 64 two rounds of independent loads (7 block pivots,
// then 7 entries of the block) replace six dependent ones.
constexpr uint RGB2SPEC_LUT_SIZE = 1024u;
VK_D uint rgb2specFindInterval(const float* __restrict__ scale, uint res, float x, const unsigned char* __restrict__ lut = nullptr) {
    if (lut) {   // (res == 64) same result as below: the table entry is the answer for a lower bound of x, and the scale axis is non-decreasing
        uint k = lut[min(uint(x * float(RGB2SPEC_LUT_SIZE)), RGB2SPEC_LUT_SIZE - 1u)];
        while (k < 62u && scale[k + 1u] <= x) k++;
        return k;
    }
    if (res == 64u) {
        // all loads of a round are issued before the first comparison: two load latencies per search instead of fourteen
        const float p0 = scale[8], p1 = scale[16], p2 = scale[24], p3 = scale[32], p4 = scale[40], p5 = scale[48], p6 = scale[56];
        const uint block = (p0 <= x ? 1u : 0u) + (p1 <= x ? 1u : 0u) + (p2 <= x ? 1u : 0u) + (p3 <= x ? 1u : 0u) + (p4 <= x ? 1u : 0u) +
                           (p5 <= x ? 1u : 0u) + (p6 <= x ? 1u : 0u);
        const float* s = scale + 8u * block;
        const float q1 = s[1], q2 = s[2], q3 = s[3], q4 = s[4], q5 = s[5], q6 = s[6], q7 = s[7];
        const uint left = 8u * block + (q1 <= x ? 1u : 0u) + (q2 <= x ? 1u : 0u) + (q3 <= x ? 1u : 0u) + (q4 <= x ? 1u : 0u) +
                          (q5 <= x ? 1u : 0u) + (q6 <= x ? 1u : 0u) + (q7 <= x ? 1u : 0u);
        return min(left, 62u);
    }
    int left = 0;
    int size = int(res) - 2;
    while (size > 0) {
        int half = size >> 1;
        int middle = left + half + 1;
        if (scale[uint(middle)] <= x) {
            left = middle;
            size -= half + 1;
        } else {
            size = half;
        }
    }
    return min(uint(left), res - 2u);
}

VK_NOINLINE float3 rgb2specFetchTable(const ::float4* __restrict__ cells, const float* __restrict__ scale, uint res, float3 rgb,
                                      const unsigned char* __restrict__ lut = nullptr) {
    float z = max(rgb.x, max(rgb.y, rgb.z));
    if (z <= RGB2SPEC_EPSILON) return float3(0.0f);
    // rgb2spec.slang:39-44: the LAST channel that reaches the maximum dominates; the other two follow in cyclic order. Written as
    // selects: indexing a float3 with a runtime channel compiles to branches.
    if (rgb.x == rgb.y && rgb.y == rgb.z) {
        // Grey (a Schlick colour or a directional attenuation of an untinted dielectric): the last channel dominates and both ratios are 1,
        // i.e. x = y = res - 1 up to the rounding of the division below: cell (res - 2, res - 2) with weights (0, 1). The trilinear form
        // then reduces to the z interpolation of two cells: 2 loads instead of 8, 3 lerps instead of 21, same value.
        const uint zi = rgb2specFindInterval(scale, res, z, lut);
        const ::float4* c = cells + ((((size_t)2u * res + zi) * res + (res - 1u)) * res + (res - 1u));
        const ::float4 lo = __ldg(c), hi = __ldg(c + res * res);
        const float scale0 = scale[zi], scale1 = scale[zi + 1u];
        const float z1 = (z - scale0) / max(scale1 - scale0, RGB2SPEC_EPSILON);
        const float z0 = 1.0f - z1;
        return float3(lo.x * z0 + hi.x * z1, lo.y * z0 + hi.y * z1, lo.z * z0 + hi.z * z1);
    }
    uint dominantChannel = 0u;
    float top = rgb.x;
    if (rgb.y >= top) { dominantChannel = 1u; top = rgb.y; }
    if (rgb.z >= top) dominantChannel = 2u;
    const float cx = dominantChannel == 0u ? rgb.y : (dominantChannel == 1u ? rgb.z : rgb.x);
    const float cy = dominantChannel == 0u ? rgb.z : (dominantChannel == 1u ? rgb.x : rgb.y);
    float xyScale = float(res - 1u) / z;
    float x = cx * xyScale;
    float y = cy * xyScale;
    uint xi = min(uint(x), res - 2u);
    uint yi = min(uint(y), res - 2u);
    uint zi = rgb2specFindInterval(scale, res, z, lut);
    const ::float4* c = cells + ((((size_t)dominantChannel * res + zi) * res + yi) * res + xi);
    const uint dy = res, dz = res * res;
    const ::float4 c000 = __ldg(c), c100 = __ldg(c + 1), c010 = __ldg(c + dy), c110 = __ldg(c + dy + 1);
    const ::float4 c001 = __ldg(c + dz), c101 = __ldg(c + dz + 1), c011 = __ldg(c + dz + dy), c111 = __ldg(c + dz + dy + 1);
    float x1 = x - float(xi), y1 = y - float(yi);
    float x0 = 1.0f - x1, y0 = 1.0f - y1;
    float scale0 = scale[zi];
    float scale1 = scale[zi + 1u];
    float z1 = (z - scale0) / max(scale1 - scale0, RGB2SPEC_EPSILON);
    float z0 = 1.0f - z1;
    float3 coeff;
    coeff.x = ((c000.x * x0 + c100.x * x1) * y0 + (c010.x * x0 + c110.x * x1) * y1) * z0 +
              ((c001.x * x0 + c101.x * x1) * y0 + (c011.x * x0 + c111.x * x1) * y1) * z1;
    coeff.y = ((c000.y * x0 + c100.y * x1) * y0 + (c010.y * x0 + c110.y * x1) * y1) * z0 +
              ((c001.y * x0 + c101.y * x1) * y0 + (c011.y * x0 + c111.y * x1) * y1) * z1;
    coeff.z = ((c000.z * x0 + c100.z * x1) * y0 + (c010.z * x0 + c110.z * x1) * y1) * z0 +
              ((c001.z * x0 + c101.z * x1) * y0 + (c011.z * x0 + c111.z * x1) * y1) * z1;
    return coeff;
}
VK_D float3 rgb2specFetch(const SpectralTables& t, float3 rgb) {
    return rgb2specFetchTable(t.cells, t.scale, t.info.res, rgb, t.intervalLut);
}
VK_D float rgb2specEvalCoeffs(float3 coeff, float lambdaNm) {
    float x = (coeff.x * lambdaNm + coeff.y) * lambdaNm + coeff.z;
    return 0.5f * x * rsqrt(x * x + 1.0f) + 0.5f;
}
VK_D float rgb2specEval(const SpectralTables& t, float3 rgb, float lambdaNm) {
    return rgb2specEvalCoeffs(rgb2specFetch(t, rgb), lambdaNm);
}

// ---- utility/spectral.slang:37-136 --------------------------------------------------------------------------------
VK_D float spectralScalarFromLinearSrgb(const SpectralTables& t, float3 rgb, float lambdaNm) {
    float maxValue = max(rgb.x, max(rgb.y, rgb.z));
    if (maxValue <= 0.0f) return 0.0f;
    if (maxValue <= 1.0f) return rgb2specEval(t, rgb, lambdaNm);
    return maxValue * rgb2specEval(t, rgb / maxValue, lambdaNm);
}
VK_D float4 spectralScalarFromLinearSrgb4(const SpectralTables& t, float3 rgb, float4 lambdaNm) {
    float maxValue = max(rgb.x, max(rgb.y, rgb.z));
    if (maxValue <= 0.0f) return float4(0.0f);
    float scale = maxValue <= 1.0f ? 1.0f : maxValue;
    float3 coeff = rgb2specFetch(t, rgb / scale);
    return scale * float4(rgb2specEvalCoeffs(coeff, lambdaNm.x), rgb2specEvalCoeffs(coeff, lambdaNm.y),
                          rgb2specEvalCoeffs(coeff, lambdaNm.z), rgb2specEvalCoeffs(coeff, lambdaNm.w));
}
// The coefficient fetch of spectralScalarFromLinearSrgb4 on its own (memo producer), and the two consumers: same arithmetic, same order.
VK_D ::float4 spectralCoefficientsFromLinearSrgb(const SpectralTables& t, float3 rgb) {
    float maxValue = max(rgb.x, max(rgb.y, rgb.z));
    if (maxValue <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float scale = maxValue <= 1.0f ? 1.0f : maxValue;
    float3 coeff = rgb2specFetch(t, rgb / scale);
    return make_float4(coeff.x, coeff.y, coeff.z, scale);
}
VK_D bool spectralMemoHit(const SpectralMemoEntry* e, float3 rgb, ::float4& value) {
    if (!e) return false;
    const ::float4 k = __ldg(&e->key);
    if (__float_as_uint(k.x) != __float_as_uint(rgb.x) || __float_as_uint(k.y) != __float_as_uint(rgb.y) || __float_as_uint(k.z) != __float_as_uint(rgb.z)) return false;
    value = __ldg(&e->value);
    return true;
}
VK_D float4 spectralScalarFromLinearSrgb4(const SpectralTables& t, const SpectralMemoEntry* memo, float3 rgb, float4 lambdaNm) {
    ::float4 v;
    if (!spectralMemoHit(memo, rgb, v)) return spectralScalarFromLinearSrgb4(t, rgb, lambdaNm);
    if (v.w <= 0.0f) return float4(0.0f);
    const float3 coeff(v.x, v.y, v.z);
    return v.w * float4(rgb2specEvalCoeffs(coeff, lambdaNm.x), rgb2specEvalCoeffs(coeff, lambdaNm.y),
                        rgb2specEvalCoeffs(coeff, lambdaNm.z), rgb2specEvalCoeffs(coeff, lambdaNm.w));
}
VK_D float spectralScalarFromLinearSrgb(const SpectralTables& t, const SpectralMemoEntry* memo, float3 rgb, float lambdaNm) {
    ::float4 v;
    if (!spectralMemoHit(memo, rgb, v)) return spectralScalarFromLinearSrgb(t, rgb, lambdaNm);
    if (v.w <= 0.0f) return 0.0f;
    const float e = rgb2specEvalCoeffs(float3(v.x, v.y, v.z), lambdaNm);
    return v.w <= 1.0f ? e : v.w * e;
}
VK_D float spectralXFit1931(float l) {
    float t1 = (l - 442.0f) * (l < 442.0f ? 0.0624f : 0.0374f);
    float t2 = (l - 599.8f) * (l < 599.8f ? 0.0264f : 0.0323f);
    float t3 = (l - 501.1f) * (l < 501.1f ? 0.0490f : 0.0382f);
    return 0.362f * expf(-0.5f * t1 * t1) + 1.056f * expf(-0.5f * t2 * t2) - 0.065f * expf(-0.5f * t3 * t3);
}
VK_D float spectralYFit1931(float l) {
    float t1 = (l - 568.8f) * (l < 568.8f ? 0.0213f : 0.0247f);
    float t2 = (l - 530.9f) * (l < 530.9f ? 0.0613f : 0.0322f);
    return 0.821f * expf(-0.5f * t1 * t1) + 0.286f * expf(-0.5f * t2 * t2);
}
VK_D float spectralZFit1931(float l) {
    float t1 = (l - 437.0f) * (l < 437.0f ? 0.0845f : 0.0278f);
    float t2 = (l - 459.0f) * (l < 459.0f ? 0.0385f : 0.0725f);
    return 1.217f * expf(-0.5f * t1 * t1) + 0.681f * expf(-0.5f * t2 * t2);
}
VK_D float3 spectralXYZ1931(float l) {
    return max(float3(spectralXFit1931(l), spectralYFit1931(l), spectralZFit1931(l)), float3(0.0f));
}
VK_D float3 mul3x3(const float* m, float3 v) { // row-major float3x3 (spectral.slang:14-19)
    return float3((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[3] * v.x + m[4] * v.y) + m[5] * v.z,
                  (m[6] * v.x + m[7] * v.y) + m[8] * v.z);
}
__device__ const float XYZ_TO_LINEAR_SRGB[9] = {3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f,
                                            0.0415560f, 0.0556434f,  -0.2040259f, 1.0572252f};
__device__ const float BRADFORD[9] = {0.8951f, 0.2664f, -0.1614f, -0.7502f, 1.7135f, 0.0367f, 0.0389f, -0.0685f, 1.0296f};
__device__ const float BRADFORD_INVERSE[9] = {0.9869929f, -0.1470543f, 0.1599627f, 0.4323053f, 0.5183603f,
                                          0.0492912f, -0.0085287f, 0.0400428f, 0.9684867f};
static constexpr float CIE_Y_INTEGRAL_1931_FIT = 106.9461715f;
VK_D float3 adaptEqualEnergyXYZToD65(float3 xyz) {
    float3 lms = mul3x3(BRADFORD, xyz);
    float3 adapted = lms * float3(0.9413344f, 1.0404175f, 1.0895327f);
    return mul3x3(BRADFORD_INVERSE, adapted);
}
VK_D float3 xyzToLinearSrgb(float3 xyz) { return mul3x3(XYZ_TO_LINEAR_SRGB, adaptEqualEnergyXYZToD65(xyz)); }
VK_D float3 spectralSampleToXYZ(float scalarValue, WavelengthSample w) {
    float xyzScale = (scalarValue * w.invPdf) / CIE_Y_INTEGRAL_1931_FIT;
    return spectralXYZ1931(w.lambdaNm) * xyzScale;
}
VK_D float3 spectralSample4ToXYZ(float4 v, float4 l, float4 invPdf) {
    float3 xyz(0.0f);
    xyz += spectralXYZ1931(l.x) * ((v.x * invPdf.x) / CIE_Y_INTEGRAL_1931_FIT);
    xyz += spectralXYZ1931(l.y) * ((v.y * invPdf.y) / CIE_Y_INTEGRAL_1931_FIT);
    xyz += spectralXYZ1931(l.z) * ((v.z * invPdf.z) / CIE_Y_INTEGRAL_1931_FIT);
    xyz += spectralXYZ1931(l.w) * ((v.w * invPdf.w) / CIE_Y_INTEGRAL_1931_FIT);
    return xyz;
}
VK_D float heroWavelengthBalanceWeight(float4 techniquePathPdf) {
    float combined = dot(techniquePathPdf, float4(1.0f));
    return combined > 0.0f ? techniquePathPdf.x / combined : 0.0f;
}
VK_D float dispersiveIor(float ior, float abbeNumber, float lambdaNm) {
    if (ior <= 1.0f + 1e-4f || abbeNumber <= 0.0f) return max(ior, 1.0f);
    const float C_UM = 0.6562725f, D_UM = 0.5875618f, F_UM = 0.4861327f;
    float lambdaUm = max(lambdaNm * 0.001f, 1e-4f);
    float nFMinusNC = (ior - 1.0f) / max(abbeNumber, 1e-6f);
    float invF2 = 1.0f / (F_UM * F_UM);
    float invC2 = 1.0f / (C_UM * C_UM);
    float B = nFMinusNC / (invF2 - invC2);
    float A = ior - B / (D_UM * D_UM);
    return max(A + B / (lambdaUm * lambdaUm), 1.0f);
}

// ---- bsdf/base/material.slang:4-79 --------------------------------------------------------------------------------
struct BSDFMaterial {
    float3 baseColor = float3(1.0f);
    float roughness = 1.0f;
    float3 eta = float3(1.0f);
    float metallic = 0.0f;
    float3 k = float3(0.0f);
    float anisotropic = 0.0f;
    float specular = 0.5f;
    float specularTint = 0.0f;
    float abbeNumber = 0.0f;
    float3 sheenTint = float3(0.0f);   // sheenTintWeight.xyz / .w (two members: a float4 would force 16-byte alignment on the closure
    float sheenWeight = 0.0f;          // state, which k_shade keeps in shared memory at a bank-conflict-avoiding odd stride)
    float clearcoat = 0.0f;
    float clearcoatGloss = 0.0f;
    float ior = 1.0f;
    float diffuseRoughness = 0.0f;
    float transmission = 0.0f;
    float subsurface = 0.0f;
    float sheenRoughness = 0.0f;
    float absorptionCoefficient = 0.0f;
    float3 attenuationColor = float3(1.0f);
    // Derived once per vertex by makeBSDFMaterial (the only way to convert a Material): the Schlick F0 of the dielectric interface, which the
    // branch weights, the directional attenuation and every evaluation of the dielectric lobe would each recompute (ior, specular, tint).
    float3 specularF0 = float3(0.0f);
    __device__ BSDFMaterial() {}
    friend __device__ BSDFMaterial makeBSDFMaterial(const Material& m);

private:
    __device__ explicit BSDFMaterial(const Material& m) {
        baseColor = float3(m.baseColor[0], m.baseColor[1], m.baseColor[2]);
        roughness = m.roughness;
        eta = float3(m.eta[0], m.eta[1], m.eta[2]);
        metallic = m.metallic;
        k = float3(m.k[0], m.k[1], m.k[2]);
        anisotropic = m.anisotropic;
        specular = m.specular;
        specularTint = m.specularTint;
        abbeNumber = m.abbeNumber;
        sheenTint = float3(m.sheenTintWeight[0], m.sheenTintWeight[1], m.sheenTintWeight[2]);
        sheenWeight = m.sheenTintWeight[3];
        clearcoat = m.clearcoat;
        clearcoatGloss = m.clearcoatGloss;
        ior = m.ior;
        diffuseRoughness = m.diffuseRoughness;
        transmission = m.transmission;
        subsurface = m.subsurface;
        sheenRoughness = m.sheenRoughness;
        absorptionCoefficient = m.absorptionCoefficient;
        attenuationColor = float3(m.attenuationColor[0], m.attenuationColor[1], m.attenuationColor[2]);
    }

public:
};
VK_D float3 bsdfDiffuseColor(const BSDFMaterial& m) { return m.baseColor * (1.0f - m.metallic); }
VK_D float3 bsdfTintColor(float3 baseColor) {
    float lum = linearSrgbLuminance(baseColor);
    return lum <= 0.0f ? float3(1.0f) : baseColor / lum;
}
VK_D float bsdfSheenWeight(const BSDFMaterial& m) { return saturate(m.sheenWeight); }
VK_D float3 bsdfSheenTint(const BSDFMaterial& m) { return saturate(m.sheenTint); }
VK_D float3 bsdfSheenColor(const BSDFMaterial& m) { return bsdfSheenTint(m) * bsdfSheenWeight(m); }
VK_D float3 bsdfTransmissionColor(const BSDFMaterial& m) {
    return m.transmission > 0.0f && m.metallic <= 0.0f ? float3(1.0f) : m.baseColor;
}

// ---- bsdf/base/math.slang:9-67 ------------------------------------------------------------------------------------
VK_D float pow5(float v) { float v2 = v * v; return v2 * v2 * v; }
VK_D float schlickWeight(float cosTheta) { return pow5(1.0f - saturate(cosTheta)); }
VK_D float cosTheta(float3 w) { return w.z; }
VK_D float absCosTheta(float3 w) { return abs(w.z); }
VK_D float cos2Theta(float3 w) { return w.z * w.z; }
VK_D float sin2Theta(float3 w) { return saturate(1.0f - cos2Theta(w)); }
VK_D float cosineHemispherePdf(float3 wi) { return cosTheta(wi) > 0.0f ? cosTheta(wi) * INV_PI : 0.0f; }
VK_D float3 sampleCosineHemisphere(uint& rng) {
    float u1 = rand(rng);
    float u2 = rand(rng);
    float r = sqrt(u1);
    float phi = 2.0f * PI * u2;
    return float3(r * cosf(phi), r * sinf(phi), sqrt(max(0.0f, 1.0f - u1)));
}
VK_D float cosPhiDifference(float3 wi, float3 wo) {
    float s2i = sin2Theta(wi), s2o = sin2Theta(wo);
    if (s2i <= 0.0f || s2o <= 0.0f) return 0.0f;
    float inv = rsqrt(s2i * s2o);
    return clamp((wi.x * wo.x + wi.y * wo.y) * inv, -1.0f, 1.0f);
}

// ---- bsdf/base/basis.slang:4-47 -----------------------------------------------------------------------------------
struct ShadingBasis {
    float3 tangent, bitangent, normal;
};
VK_D float3 makeFallbackTangent(float3 normal) {
    float3 up = abs(normal.z) < TANGENT_PARALLEL_THRESHOLD ? float3(0.0f, 0.0f, 1.0f) : float3(1.0f, 0.0f, 0.0f);
    return safeNormalize(cross(up, normal));
}
VK_D ShadingBasis makeShadingBasis(float3 normal, float4 tangentData) {
    ShadingBasis b;
    b.normal = safeNormalize(normal);
    float3 tangent = tangentData.xyz() - b.normal * dot(tangentData.xyz(), b.normal);
    tangent = dot(tangent, tangent) <= SAFE_NORMALIZE_EPS ? makeFallbackTangent(b.normal) : safeNormalize(tangent);
    float handedness = tangentData.w < 0.0f ? -1.0f : 1.0f;
    b.tangent = tangent;
    b.bitangent = cross(b.normal, b.tangent) * handedness;
    return b;
}
VK_D float3 sanitizeShadingNormal(float3 shadingNormal, float3 geometricNormal, float3 outgoing) {
    float3 n = safeNormalize(shadingNormal);
    float3 g = safeNormalize(geometricNormal);
    float3 o = safeNormalize(outgoing);
    if (dot(n, g) <= 1e-4f || dot(n, o) <= 1e-4f) return g;
    return n;
}
VK_D float3 localToWorld(float3 l, const ShadingBasis& b) { return b.tangent * l.x + b.bitangent * l.y + b.normal * l.z; }
VK_D float3 worldToLocal(float3 w, const ShadingBasis& b) {
    return float3(dot(w, b.tangent), dot(w, b.bitangent), dot(w, b.normal));
}

// ---- bsdf/base/fresnel.slang:4-68 ---------------------------------------------------------------------------------
VK_D float dielectricF0(float eta) {
    float f0 = (eta - 1.0f) / max(eta + 1.0f, 1e-6f);
    return f0 * f0;
}
VK_D bool materialHasConductor(const BSDFMaterial& m) { return anyGreater(m.k, 0.0f); }
VK_D float3 computeDielectricSpecularF0(const BSDFMaterial& m) {
    float dielectric = dielectricF0(m.ior);
    float specularScale = m.specular / 0.5f;
    float3 tint = lerp(float3(1.0f), bsdfTintColor(m.baseColor), saturate(m.specularTint));
    return saturate(dielectric * specularScale * tint);
}
__device__ __forceinline__ BSDFMaterial makeBSDFMaterial(const Material& src) {
    BSDFMaterial m(src);
    m.specularF0 = computeDielectricSpecularF0(m);
    return m;
}
VK_D float3 bsdfDielectricSpecularF0(const BSDFMaterial& m) { return m.specularF0; }
VK_D float bsdfDielectricSpecularF0Luminance(const BSDFMaterial& m) {
    return saturate(linearSrgbLuminance(bsdfDielectricSpecularF0(m)));
}
VK_D float3 fresnelSchlick(float cosT, float3 f0) {
    float w = schlickWeight(cosT);
    return f0 + (1.0f - f0) * w;
}
VK_D float fresnelDielectric(float cosT, float eta) {
    float cosI = clamp(cosT, -1.0f, 1.0f);
    if (cosI < 0.0f) {
        eta = 1.0f / max(eta, 1e-6f);
        cosI = -cosI;
    }
    float sin2I = max(1.0f - cosI * cosI, 0.0f);
    float sin2T = sin2I / max(eta * eta, 1e-6f);
    if (sin2T >= 1.0f) return 1.0f;
    float cosTt = sqrt(max(1.0f - sin2T, 0.0f));
    float rPar = (eta * cosI - cosTt) / max(eta * cosI + cosTt, 1e-6f);
    float rPerp = (cosI - eta * cosTt) / max(cosI + eta * cosTt, 1e-6f);
    return 0.5f * (rPar * rPar + rPerp * rPerp);
}
VK_NOINLINE float3 fresnelConductor(float cosT, float3 eta, float3 k) {
    float cosI = saturate(cosT);
    float cos2I = cosI * cosI;
    float sin2I = max(1.0f - cos2I, 0.0f);
    float3 eta2 = eta * eta;
    float3 k2 = k * k;
    float3 t0 = eta2 - k2 - sin2I;
    float3 a2PlusB2 = sqrt(max(t0 * t0 + 4.0f * eta2 * k2, float3(0.0f)));
    float3 t1 = a2PlusB2 + cos2I;
    float3 a = sqrt(max(0.5f * (a2PlusB2 + t0), float3(0.0f)));
    float3 t2 = 2.0f * cosI * a;
    float3 rs = (t1 - t2) / max(t1 + t2, float3(1e-6f));
    float3 t3 = cos2I * a2PlusB2 + sin2I * sin2I;
    float3 t4 = t2 * sin2I;
    float3 rp = rs * ((t3 - t4) / max(t3 + t4, float3(1e-6f)));
    return 0.5f * (rp + rs);
}

// ---- bsdf/base/interface.slang:4-31 -------------------------------------------------------------------------------
VK_D float interfaceIor(const BSDFMaterial& m) { return max(m.ior, 1.0f); }
VK_D float interfaceIor(const BSDFMaterial& m, float lambdaNm, uint spectralMode) {
    float ior = interfaceIor(m);
    return spectralMode != 0u ? dispersiveIor(ior, m.abbeNumber, lambdaNm) : ior;
}
VK_D float interfaceEta(const BSDFMaterial& m, uint frontFace, float lambdaNm, uint spectralMode) {
    float ior = interfaceIor(m, lambdaNm, spectralMode);
    return frontFace != 0u ? ior : (1.0f / ior);
}
VK_D float interfaceRefractionEta(const BSDFMaterial& m, uint frontFace, float lambdaNm, uint spectralMode) {
    float ior = interfaceIor(m, lambdaNm, spectralMode);
    return frontFace != 0u ? (1.0f / ior) : ior;
}

// ---- bsdf/base/medium.slang:7-113 ---------------------------------------------------------------------------------
static constexpr uint MEDIUM_FLAG_REFRACTIVE_ACTIVE = 1u << 0;
static constexpr uint MEDIUM_FLAG_ABSORPTION_ACTIVE = 1u << 1;
struct MediumState {
    uint flags = 0u;
    float3 absorptionSigma = float3(0.0f);
    float4 spectralAbsorptionSigma = float4(0.0f);
    __device__ bool refractiveActive() const { return (flags & MEDIUM_FLAG_REFRACTIVE_ACTIVE) != 0u; }
    __device__ bool absorptionActive() const { return (flags & MEDIUM_FLAG_ABSORPTION_ACTIVE) != 0u; }
    __device__ void setRefractiveActive(bool e) { flags = e ? (flags | MEDIUM_FLAG_REFRACTIVE_ACTIVE) : (flags & ~MEDIUM_FLAG_REFRACTIVE_ACTIVE); }
    __device__ void setAbsorptionActive(bool e) { flags = e ? (flags | MEDIUM_FLAG_ABSORPTION_ACTIVE) : (flags & ~MEDIUM_FLAG_ABSORPTION_ACTIVE); }
};
VK_D bool materialHasAbsorption(const BSDFMaterial& m) {
    return m.absorptionCoefficient > 0.0f && anyLess(m.attenuationColor, 0.9999f);
}
VK_D bool materialMediumIsRefractive(const BSDFMaterial& m) {
    return m.transmission > 0.0f && interfaceIor(m) > 1.0f + 1e-4f;
}
VK_D float3 materialAbsorptionSigma(const BSDFMaterial& m) {
    float3 tint = max(saturate(m.attenuationColor), float3(1e-6f));
    return -log(tint) * m.absorptionCoefficient;
}
VK_D float3 materialAbsorptionSigma(const SpectralTables& t, const BSDFMaterial& m, float lambdaNm, uint spectralMode) {
    if (spectralMode == 0u) return materialAbsorptionSigma(m);
    float tint = max(spectralScalarFromLinearSrgb(t, saturate(m.attenuationColor), lambdaNm), 1e-6f);
    return float3(-logf(tint) * m.absorptionCoefficient);
}
VK_D float4 materialAbsorptionSigma4(const SpectralTables& t, const BSDFMaterial& m, float4 wl) {
    float4 tint = max(spectralScalarFromLinearSrgb4(t, saturate(m.attenuationColor), wl), float4(1e-6f));
    return -log(tint) * m.absorptionCoefficient;
}
VK_D bool mediumHasActiveBoundary(const MediumState& m) { return m.flags != 0u; }
VK_D float3 mediumTransmittance(const MediumState& m, float distance) {
    return m.absorptionActive() ? exp(-m.absorptionSigma * distance) : float3(1.0f);
}
VK_D float4 mediumSpectralTransmittance(const MediumState& m, float distance) {
    return m.absorptionActive() ? exp(-m.spectralAbsorptionSigma * distance) : float4(1.0f);
}
VK_D void updateMediumStateFromTransmission(const SpectralTables& t, const BSDFMaterial& m, uint frontFace,
                                              uint isTransmission, float lambdaNm, uint spectralMode, MediumState& medium) {
    if (isTransmission == 0u) return;
    bool entering = frontFace != 0u;
    medium.setRefractiveActive(entering && materialMediumIsRefractive(m));
    medium.setAbsorptionActive(entering && materialHasAbsorption(m));
    medium.absorptionSigma = medium.absorptionActive() ? materialAbsorptionSigma(t, m, lambdaNm, spectralMode) : float3(0.0f);
    medium.spectralAbsorptionSigma = float4(medium.absorptionSigma.x);
}
VK_D void updateMediumStateFromTransmissionSpectral(const SpectralTables& t, const BSDFMaterial& m, uint frontFace,
                                                      uint isTransmission, float4 wl, MediumState& medium) {
    if (isTransmission == 0u) return;
    bool entering = frontFace != 0u;
    medium.setRefractiveActive(entering && materialMediumIsRefractive(m));
    medium.setAbsorptionActive(entering && materialHasAbsorption(m));
    if (!medium.absorptionActive()) {
        medium.absorptionSigma = float3(0.0f);
        medium.spectralAbsorptionSigma = float4(0.0f);
        return;
    }
    medium.spectralAbsorptionSigma = materialAbsorptionSigma4(t, m, wl);
    medium.absorptionSigma = float3(medium.spectralAbsorptionSigma.x);
}

// ---- bsdf/types.slang ---------------------------------------------------------------------------------------------
struct BSDFSample {
    float3 wi = float3(0.0f);
    float3 weight = float3(0.0f);
    float pdf = 0.0f;
    uint isTransmission = 0u;
    __device__ bool isUsable() const { return pdf > 0.0f && anyGreater(weight, 0.0f); }
};
struct BSDFEval {
    float3 value = float3(0.0f);
    float pdf = 0.0f;
};
struct SpectralBSDFSample {
    float3 wi = float3(0.0f);
    float4 weight = float4(0.0f);
    float4 techniquePdf = float4(0.0f);
    uint isTransmission = 0u;
    __device__ bool isUsable() const { return techniquePdf.x > 0.0f && anyGreater(weight, 0.0f); }
};

// ---- bsdf/lobes/ggx.slang:16-227 ----------------------------------------------------------------------------------
static constexpr float GGX_MIN_ALPHA = 1e-3f;
static constexpr float GGX_EPSILON = 1e-6f;
struct GGXParams {
    float2 alpha;
    float vndfK = 0.0f;
    float projectedWoLength = 0.0f;
};
VK_D float2 ggxAlpha(const BSDFMaterial& m) {
    float roughness = max(m.roughness, GGX_MIN_ALPHA);
    float aspect = sqrt(max(1.0f - 0.9f * m.anisotropic, GGX_MIN_ALPHA));
    float a = roughness * roughness;
    return max(float2(a / aspect, a * aspect), float2(GGX_MIN_ALPHA));
}
VK_D float ggxVNDFK(float3 wo, float2 alpha) {
    float a = saturate(min(alpha.x, alpha.y));
    float s = 1.0f + length(float2(wo.x, wo.y));
    float a2 = a * a, s2 = s * s;
    return (1.0f - a2) * s2 / max(s2 + a2 * wo.z * wo.z, GGX_EPSILON);
}
VK_D float ggxLambda(float3 w, float2 alpha) {
    float2 alpha2 = alpha * alpha;
    float z2 = max(w.z * w.z, GGX_EPSILON);
    float2 wxy(w.x, w.y);
    float slope2 = dot(alpha2 * wxy, wxy) / z2;
    return 0.5f * (sqrt(1.0f + slope2) - 1.0f);
}
VK_D float ggxDistribution(float3 m, float2 alpha) {
    if (m.z <= 0.0f) return 0.0f;
    float c = (m.x * m.x) / (alpha.x * alpha.x) + (m.y * m.y) / (alpha.y * alpha.y) + m.z * m.z;
    return 1.0f / (PI * alpha.x * alpha.y * c * c);
}
VK_D float ggxMasking(float3 wo, float3 wi, float2 alpha) { return 1.0f / (1.0f + ggxLambda(wo, alpha) + ggxLambda(wi, alpha)); }
VK_D float ggxMasking1(float3 w, float2 alpha) { return 1.0f / (1.0f + ggxLambda(w, alpha)); }
VK_D float ggxProjectedLength(float3 wo, float2 alpha) {
    float2 s = alpha * float2(wo.x, wo.y);
    return sqrt(dot(s, s) + wo.z * wo.z);
}
VK_D GGXParams makeGGXParams(const BSDFMaterial& m, float3 wo) {
    GGXParams p;
    p.alpha = ggxAlpha(m);
    p.vndfK = ggxVNDFK(wo, p.alpha);
    p.projectedWoLength = ggxProjectedLength(wo, p.alpha);
    return p;
}
VK_D bool ggxHalfVector(float3 wo, float3 wi, float3& m, float& woDotM) {
    float3 h = wo + wi;
    float h2 = dot(h, h);
    if (h2 <= GGX_EPSILON) {
        m = float3(0.0f);
        woDotM = 0.0f;
        return false;
    }
    m = h * rsqrt(h2);
    woDotM = dot(wo, m);
    return woDotM > 0.0f;
}
VK_D float ggxVisibleNormalPdf(float3 wo, float3 m, const GGXParams& p) {
    float woDotM = dot(wo, m);
    if (cosTheta(wo) <= 0.0f || m.z <= 0.0f || woDotM <= 0.0f) return 0.0f;
    float Dm = ggxDistribution(m, p.alpha);
    float G1 = ggxMasking1(wo, p.alpha);
    return Dm * G1 * woDotM / max(cosTheta(wo), GGX_EPSILON);
}
VK_D float ggxReflectionPdf(float3 wo, float3 m, const GGXParams& p) {
    float visiblePdf = ggxVisibleNormalPdf(wo, m, p);
    return visiblePdf / max(4.0f * abs(dot(wo, m)), GGX_EPSILON);
}
VK_D float3 ggxConductorFresnelColor(const BSDFMaterial& mat, float cosT) {
    return materialHasConductor(mat) ? fresnelConductor(cosT, max(mat.eta, float3(1e-3f)), max(mat.k, float3(0.0f)))
                                     : fresnelSchlick(cosT, mat.baseColor);
}
VK_D float3 ggxConductorFresnel(const SpectralTables& t, const BSDFMaterial& mat, float cosT, float lambdaNm, uint spectralMode) {
    float3 c = ggxConductorFresnelColor(mat, cosT);
    return spectralMode != 0u ? float3(spectralScalarFromLinearSrgb(t, saturate(c), lambdaNm)) : c;
}
VK_D float4 ggxConductorFresnel4(const SpectralTables& t, const BSDFMaterial& mat, float cosT, float4 wl) {
    return spectralScalarFromLinearSrgb4(t, saturate(ggxConductorFresnelColor(mat, cosT)), wl);
}
VK_D BSDFEval evalGGX(const SpectralTables& t, const BSDFMaterial& mat, float3 wo, float3 wi, const GGXParams& p,
                        float lambdaNm, uint spectralMode) {
    BSDFEval e;
    if (cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return e;
    float3 m;
    float woDotM;
    if (!ggxHalfVector(wo, wi, m, woDotM)) return e;
    float Dm = ggxDistribution(m, p.alpha);
    float G = ggxMasking(wo, wi, p.alpha);
    float3 F = ggxConductorFresnel(t, mat, woDotM, lambdaNm, spectralMode);
    e.value = F * (Dm * G / max(4.0f * cosTheta(wo) * cosTheta(wi), GGX_EPSILON));
    e.pdf = ggxReflectionPdf(wo, m, p);
    return e;
}
VK_D float4 evalSpectralGGX(const SpectralTables& t, const BSDFMaterial& mat, float3 wo, float3 wi, const GGXParams& p,
                              float4 wl, float& pdf) {
    pdf = 0.0f;
    if (cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return float4(0.0f);
    float3 m;
    float woDotM;
    if (!ggxHalfVector(wo, wi, m, woDotM)) return float4(0.0f);
    float Dm = ggxDistribution(m, p.alpha);
    float G = ggxMasking(wo, wi, p.alpha);
    pdf = ggxReflectionPdf(wo, m, p);
    return ggxConductorFresnel4(t, mat, woDotM, wl) * (Dm * G / max(4.0f * cosTheta(wo) * cosTheta(wi), GGX_EPSILON));
}
VK_D bool sampleGGXVNDF(float3 wo, const GGXParams& p, float2 u, float3& m) {
    if (p.projectedWoLength <= GGX_EPSILON) {
        m = float3(0.0f);
        return false;
    }
    float3 woStd = float3(wo.x * p.alpha.x, wo.y * p.alpha.y, wo.z) / p.projectedWoLength;
    float phi = 2.0f * PI * u.x;
    float b = p.vndfK * woStd.z;
    float z = (1.0f - u.y) * (1.0f + b) - b;
    float sinT = sqrt(saturate(1.0f - z * z));
    float3 wiStd = float3(sinT * cosf(phi), sinT * sinf(phi), z);
    float3 mStd = woStd + wiStd;
    float mStd2 = dot(mStd, mStd);
    if (mStd2 <= GGX_EPSILON) {
        m = float3(0.0f);
        return false;
    }
    float3 un = float3(mStd.x * p.alpha.x, mStd.y * p.alpha.y, mStd.z);
    float un2 = dot(un, un);
    if (un2 <= GGX_EPSILON) {
        m = float3(0.0f);
        return false;
    }
    m = un * rsqrt(un2);
    return m.z > 0.0f;
}
VK_D bool sampleGGX(float3 wo, const GGXParams& p, uint& rng, float3& wi) {
    float ux = rand(rng);
    float uy = rand(rng);
    float3 m;
    if (!sampleGGXVNDF(wo, p, float2(ux, uy), m)) {
        wi = float3(0.0f);
        return false;
    }
    float woDotM = dot(wo, m);
    if (woDotM <= 0.0f) {
        wi = float3(0.0f);
        return false;
    }
    wi = 2.0f * woDotM * m - wo;
    return cosTheta(wi) > 0.0f;
}

// ---- bsdf/lobes/diffuse.slang:13-53, subsurface.slang:11-31 -------------------------------------------------------
VK_D BSDFEval evalLambertian(float3 diffuseColor, float3 wi) {
    BSDFEval e;
    e.value = cosTheta(wi) > 0.0f ? diffuseColor * INV_PI : float3(0.0f);
    e.pdf = cosineHemispherePdf(wi);
    return e;
}
VK_NOINLINE BSDFEval evalOrenNayar(float3 diffuseColor, float roughness, float3 wo, float3 wi) {
    BSDFEval e;
    if (cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return e;
    float sigma = saturate(roughness) * (0.5f * PI);
    float sigma2 = sigma * sigma;
    float A = 1.0f - sigma2 / (2.0f * (sigma2 + 0.33f));
    float B = 0.45f * sigma2 / (sigma2 + 0.09f);
    float sinI = sqrt(sin2Theta(wi));
    float sinO = sqrt(sin2Theta(wo));
    float maxCos = max(0.0f, cosPhiDifference(wi, wo));
    float sinAlpha, tanBeta;
    if (absCosTheta(wi) > absCosTheta(wo)) {
        sinAlpha = sinO;
        tanBeta = sinI / max(absCosTheta(wi), 1e-6f);
    } else {
        sinAlpha = sinI;
        tanBeta = sinO / max(absCosTheta(wo), 1e-6f);
    }
    e.value = diffuseColor * (INV_PI * (A + B * maxCos * sinAlpha * tanBeta));
    e.pdf = cosineHemispherePdf(wi);
    return e;
}
VK_NOINLINE BSDFEval evalFakeSubsurface(float3 diffuseColor, float roughness, float3 wo, float3 wi) {
    BSDFEval e;
    if (cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return e;
    float3 h = safeNormalize(wo + wi);
    float wiDotH = saturate(dot(wi, h));
    float fss90 = wiDotH * wiDotH * roughness;
    float fssIn = lerp(1.0f, fss90, schlickWeight(cosTheta(wi)));
    float fssOut = lerp(1.0f, fss90, schlickWeight(cosTheta(wo)));
    float fss = fssIn * fssOut;
    float ss = 1.25f * (fss * (1.0f / max(cosTheta(wi) + cosTheta(wo), 1e-6f) - 0.5f) + 0.5f);
    e.value = diffuseColor * (INV_PI * ss);
    e.pdf = cosineHemispherePdf(wi);
    return e;
}

// ---- bsdf/lobes/clearcoat.slang:14-97 -----------------------------------------------------------------------------
static constexpr float CLEARCOAT_EPSILON = 1e-6f;
struct ClearcoatParams { float alpha = 0.0f; };
VK_D ClearcoatParams makeClearcoatParams(const BSDFMaterial& m) {
    ClearcoatParams p;
    p.alpha = max(lerp(0.1f, 0.001f, saturate(m.clearcoatGloss)), 1e-3f);
    return p;
}
VK_D float clearcoatSmithG1(float cosT, float alpha) {
    float alpha2 = alpha * alpha;
    float cos2 = max(cosT * cosT, CLEARCOAT_EPSILON);
    float tan2 = max(1.0f - cos2, 0.0f) / cos2;
    return 2.0f / (1.0f + sqrt(1.0f + alpha2 * tan2));
}
VK_D float clearcoatDistribution(float cosThetaM, float alpha) {
    float alpha2 = alpha * alpha;
    if (alpha2 >= 1.0f - CLEARCOAT_EPSILON) return INV_PI;
    float denom = PI * logf(alpha2) * (1.0f + (alpha2 - 1.0f) * cosThetaM * cosThetaM);
    return (alpha2 - 1.0f) / min(denom, -CLEARCOAT_EPSILON);
}
VK_NOINLINE BSDFEval evalClearcoat(float clearcoatWeight, float3 wo, float3 wi, const ClearcoatParams& p) {
    BSDFEval e;
    if (clearcoatWeight <= 0.0f || cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return e;
    float3 h = safeNormalize(wo + wi);
    float woDotH = saturate(dot(wo, h));
    if (woDotH <= 0.0f) return e;
    float D = clearcoatDistribution(saturate(cosTheta(h)), p.alpha);
    float F = lerp(0.04f, 1.0f, schlickWeight(woDotH));
    float G = clearcoatSmithG1(cosTheta(wo), 0.25f) * clearcoatSmithG1(cosTheta(wi), 0.25f);
    float value = 0.25f * clearcoatWeight * D * F * G / max(4.0f * cosTheta(wo) * cosTheta(wi), CLEARCOAT_EPSILON);
    float pdfM = D * saturate(cosTheta(h));
    e.value = float3(value);
    e.pdf = pdfM / max(4.0f * woDotH, CLEARCOAT_EPSILON);
    return e;
}
VK_NOINLINE bool sampleClearcoat(float3 wo, const ClearcoatParams& p, uint& rng, float3& wi) {
    float u1 = rand(rng);
    float u2 = rand(rng);
    float alpha2 = p.alpha * p.alpha;
    float cosM;
    if (alpha2 >= 1.0f - CLEARCOAT_EPSILON) {
        cosM = sqrt(max(0.0f, 1.0f - u1));
    } else {
        float exponent = powf(alpha2, 1.0f - u1);
        cosM = sqrt(saturate((1.0f - exponent) / max(1.0f - alpha2, CLEARCOAT_EPSILON)));
    }
    float sinM = sqrt(saturate(1.0f - cosM * cosM));
    float phi = 2.0f * PI * u2;
    float3 m = float3(sinM * cosf(phi), sinM * sinf(phi), cosM);
    float woDotM = dot(wo, m);
    if (woDotM <= 0.0f) {
        wi = float3(0.0f);
        return false;
    }
    wi = 2.0f * woDotM * m - wo;
    return cosTheta(wi) > 0.0f;
}

// ---- bsdf/lobes/sheen.slang:14-113 (+ data/sheen_ltc.slang) -------------------------------------------------------
static constexpr uint SHEEN_LTC_SIZE = 32u;
static constexpr uint SHEEN_LTC_LAYER_SIZE = SHEEN_LTC_SIZE * SHEEN_LTC_SIZE;
__device__ const float SHEEN_LTC_TABLE[3072] = {
#include "data/sheen_ltc.inc"
};
static constexpr float SHEEN_EPSILON = 1e-6f;
struct SheenParams {
    ShadingBasis basis;
    float transformA = 0.0f, transformB = 0.0f, albedo = 0.0f;
};
VK_D float sheenLtcLookup(float cosT, float roughness, uint layer) {
    float2 uv = saturate(float2(cosT, roughness)) * float(SHEEN_LTC_SIZE - 1u);
    uint x0 = min(uint(uv.x), SHEEN_LTC_SIZE - 1u);
    uint y0 = min(uint(uv.y), SHEEN_LTC_SIZE - 1u);
    uint x1 = min(x0 + 1u, SHEEN_LTC_SIZE - 1u);
    uint y1 = min(y0 + 1u, SHEEN_LTC_SIZE - 1u);
    float tx = frac(uv.x), ty = frac(uv.y);
    uint off = layer * SHEEN_LTC_LAYER_SIZE;
    float e00 = SHEEN_LTC_TABLE[off + y0 * SHEEN_LTC_SIZE + x0];
    float e10 = SHEEN_LTC_TABLE[off + y0 * SHEEN_LTC_SIZE + x1];
    float e01 = SHEEN_LTC_TABLE[off + y1 * SHEEN_LTC_SIZE + x0];
    float e11 = SHEEN_LTC_TABLE[off + y1 * SHEEN_LTC_SIZE + x1];
    return lerp(lerp(e00, e10, tx), lerp(e01, e11, tx), ty);
}
VK_D SheenParams makeSheenParams(float3 wo, float sheenRoughness) {
    SheenParams p;
    p.basis = makeShadingBasis(float3(0.0f, 0.0f, 1.0f), float4(wo, 1.0f));
    float cosO = cosTheta(wo);
    if (cosO <= 0.0f) return p;
    float roughness = clamp(sheenRoughness, 1e-3f, 1.0f);
    p.transformA = sheenLtcLookup(cosO, roughness, 0u);
    p.transformB = sheenLtcLookup(cosO, roughness, 1u);
    p.albedo = sheenLtcLookup(cosO, roughness, 2u);
    return p;
}
VK_D float sheenDistributionValue(float3 localWi, const SheenParams& p) {
    float z = max(localWi.z, 0.0f);
    if (z <= 0.0f || abs(p.transformA) < 1e-5f || p.albedo < 1e-5f) return 0.0f;
    float axbz = p.transformA * localWi.x + p.transformB * localWi.z;
    float lenSqr = axbz * axbz + (p.transformA * localWi.y) * (p.transformA * localWi.y) + localWi.z * localWi.z;
    if (lenSqr <= SHEEN_EPSILON) return 0.0f;
    float scale = p.transformA / lenSqr;
    return INV_PI * z * scale * scale;
}
VK_NOINLINE float sheenDirectionalAlbedo(float cosT, float sheenRoughness) {
    return sheenLtcLookup(cosT, clamp(sheenRoughness, 1e-3f, 1.0f), 2u);
}
VK_D float sheenLayerAttenuation(float sheenWeight, float cosT, float sheenRoughness) {
    if (sheenWeight <= 0.0f) return 1.0f;   // saturate(1 - 0 * albedo) without the LTC lookup (three per vertex for a material without sheen)
    return saturate(1.0f - sheenWeight * sheenDirectionalAlbedo(cosT, sheenRoughness));
}
VK_NOINLINE BSDFEval evalSheen(float3 sheenColor, float sheenRoughness, float3 wo, float3 wi) {
    BSDFEval e;
    if (cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return e;
    SheenParams p = makeSheenParams(wo, sheenRoughness);
    float3 localWi = worldToLocal(wi, p.basis);
    float value = sheenDistributionValue(localWi, p);
    e.value = sheenColor * (p.albedo * value);
    e.pdf = value;
    return e;
}
VK_NOINLINE bool sampleSheen(float3 wo, float sheenRoughness, uint& rng, float3& wi) {
    SheenParams p = makeSheenParams(wo, sheenRoughness);
    if (abs(p.transformA) < 1e-5f || p.albedo < 1e-5f) {
        wi = float3(0.0f);
        return false;
    }
    float r = sqrt(rand(rng));
    float phi = 2.0f * PI * rand(rng);
    float2 disk = r * float2(cosf(phi), sinf(phi));
    float diskZ = sqrt(max(1.0f - dot(disk, disk), 0.0f));
    float3 localWi = normalize(float3(disk.x - diskZ * p.transformB, disk.y, diskZ * p.transformA));
    wi = localToWorld(localWi, p.basis);
    return cosTheta(wi) > 0.0f;
}

// ---- bsdf/lobes/dielectric/interface.slang:13-100 -----------------------------------------------------------------
VK_D bool dielectricIsIdentity(const BSDFMaterial& m) { return interfaceIor(m) <= 1.0f + 1e-4f; }
VK_D float dielectricFresnel(const BSDFMaterial& m, uint frontFace, float cosI, float lambdaNm, uint spectralMode) {
    if (dielectricIsIdentity(m)) return 0.0f;
    return fresnelDielectric(cosI, interfaceEta(m, frontFace, lambdaNm, spectralMode));
}
VK_D float4 dielectricFresnel4(const BSDFMaterial& m, uint frontFace, float cosI, float4 wl) {
    if (dielectricIsIdentity(m)) return float4(0.0f);
    return float4(fresnelDielectric(cosI, interfaceEta(m, frontFace, wl.x, 1u)),
                  fresnelDielectric(cosI, interfaceEta(m, frontFace, wl.y, 1u)),
                  fresnelDielectric(cosI, interfaceEta(m, frontFace, wl.z, 1u)),
                  fresnelDielectric(cosI, interfaceEta(m, frontFace, wl.w, 1u)));
}
VK_D float3 dielectricReflectionColor(const SpectralTables& t, const BSDFMaterial& m, float cosI, float fresnel,
                                        float lambdaNm, uint spectralMode) {
    if (m.transmission > 0.0f && m.metallic <= 0.0f) return float3(fresnel);
    float3 c = fresnelSchlick(cosI, bsdfDielectricSpecularF0(m));
    return spectralMode != 0u ? float3(spectralScalarFromLinearSrgb(t, saturate(c), lambdaNm)) : c;
}
VK_D float4 dielectricReflectionColor4(const SpectralTables& t, const BSDFMaterial& m, float cosI, float4 fresnel, float4 wl) {
    if (m.transmission > 0.0f && m.metallic <= 0.0f) return fresnel;
    float3 c = fresnelSchlick(cosI, bsdfDielectricSpecularF0(m));
    return spectralScalarFromLinearSrgb4(t, saturate(c), wl);
}
VK_D float dielectricTransmissionProbability(const BSDFMaterial& m, float fresnel) {
    float tw = m.transmission * (1.0f - fresnel);
    float total = fresnel + tw;
    return total <= 0.0f ? 0.0f : tw / total;
}
VK_D bool dielectricTransmissionHalfVector(float3 wo, float3 wi, float etap, float3& wm, float& woDotWm, float& wiDotWm,
                                             float& denom) {
    wm = float3(0.0f);
    woDotWm = wiDotWm = denom = 0.0f;
    if (cosTheta(wo) <= 0.0f || cosTheta(wi) >= 0.0f) return false;
    float3 h = wo + wi * etap;
    if (dot(h, h) <= GGX_EPSILON) return false;
    wm = safeNormalize(h);
    if (wm.z < 0.0f) wm = -wm;
    woDotWm = dot(wo, wm);
    wiDotWm = dot(wi, wm);
    denom = wiDotWm + woDotWm / etap;
    return woDotWm > 0.0f && wiDotWm < 0.0f && abs(denom) > GGX_EPSILON;
}

// ---- bsdf/lobes/dielectric/eval_rgb.slang:6-95 --------------------------------------------------------------------
VK_D BSDFEval evalDielectricReflection(const SpectralTables& t, const BSDFMaterial& mat, float3 wo, float3 wi, uint frontFace,
                                         const GGXParams& p, float lambdaNm, uint spectralMode) {
    BSDFEval e;
    if (dielectricIsIdentity(mat) || cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return e;
    float3 wm;
    float woDotWm;
    if (!ggxHalfVector(wo, wi, wm, woDotWm)) return e;
    // An opaque dielectric reflects with probability 1 and takes its colour from Schlick's F0 (dielectricReflectionColor): the exact Fresnel
    // term is needed only when the material transmits.
    const bool transmits = mat.transmission > 0.0f;
    float fresnel = transmits ? dielectricFresnel(mat, frontFace, woDotWm, lambdaNm, spectralMode) : 0.0f;
    float reflectionProbability = transmits ? 1.0f - dielectricTransmissionProbability(mat, fresnel) : 1.0f;
    if (reflectionProbability <= 0.0f) return e;
    float Dm = ggxDistribution(wm, p.alpha);
    float G = ggxMasking(wo, wi, p.alpha);
    e.value = dielectricReflectionColor(t, mat, woDotWm, fresnel, lambdaNm, spectralMode) *
              (Dm * G / max(4.0f * cosTheta(wo) * cosTheta(wi), GGX_EPSILON));
    e.pdf = reflectionProbability * ggxReflectionPdf(wo, wm, p);
    return e;
}
VK_NOINLINE BSDFEval evalDielectricTransmission(const SpectralTables& t, const BSDFMaterial& mat, float3 wo, float3 wi, uint frontFace,
                                           const GGXParams& p, float coatAttenuation, float lambdaNm, uint spectralMode) {
    BSDFEval e;
    if (mat.transmission <= 0.0f || cosTheta(wo) <= 0.0f || cosTheta(wi) >= 0.0f) return e;
    float3 transmissionColor = bsdfTransmissionColor(mat) * ((1.0f - mat.metallic) * mat.transmission);
    if (spectralMode != 0u) transmissionColor = float3(spectralScalarFromLinearSrgb(t, saturate(transmissionColor), lambdaNm));
    if (dielectricIsIdentity(mat)) {
        if (dot(wi, -wo) < 0.9999f) return e;
        e.value = coatAttenuation * transmissionColor / max(absCosTheta(wi), GGX_EPSILON);
        e.pdf = 1.0f;
        return e;
    }
    float etap = interfaceEta(mat, frontFace, lambdaNm, spectralMode);
    float3 wm;
    float woDotWm, wiDotWm, denom;
    if (!dielectricTransmissionHalfVector(wo, wi, etap, wm, woDotWm, wiDotWm, denom)) return e;
    float fresnel = dielectricFresnel(mat, frontFace, woDotWm, lambdaNm, spectralMode);
    float tp = dielectricTransmissionProbability(mat, fresnel);
    if (tp <= 0.0f) return e;
    float Dm = ggxDistribution(wm, p.alpha);
    float G = ggxMasking(wo, wi, p.alpha);
    float denom2 = denom * denom;
    float transportScale = 1.0f / max(etap * etap, GGX_EPSILON);
    float transmissionTerm = abs((wiDotWm * woDotWm) / max(absCosTheta(wi) * cosTheta(wo) * denom2, GGX_EPSILON));
    float dwmDwi = abs(wiDotWm) / max(denom2, GGX_EPSILON);
    e.value = coatAttenuation * transmissionColor * ((1.0f - fresnel) * Dm * G * transmissionTerm * transportScale);
    e.pdf = tp * ggxVisibleNormalPdf(wo, wm, p) * dwmDwi;
    return e;
}

// ---- bsdf/lobes/dielectric/eval_spectral.slang:6-156 --------------------------------------------------------------
VK_D float4 evalSpectralDielectricReflection(const SpectralTables& t, const BSDFMaterial& mat, float3 wo, float3 wi,
                                               uint frontFace, const GGXParams& p, float4 wl, float4& techniquePdf) {
    techniquePdf = float4(0.0f);
    if (dielectricIsIdentity(mat) || cosTheta(wo) <= 0.0f || cosTheta(wi) <= 0.0f) return float4(0.0f);
    float3 wm;
    float woDotWm;
    if (!ggxHalfVector(wo, wi, wm, woDotWm)) return float4(0.0f);
    float Dm = ggxDistribution(wm, p.alpha);
    float G = ggxMasking(wo, wi, p.alpha);
    float basePdf = ggxReflectionPdf(wo, wm, p);
    float4 fresnel(0.0f);
    if (mat.transmission > 0.0f) {
        fresnel = dielectricFresnel4(mat, frontFace, woDotWm, wl);
        techniquePdf = float4(1.0f - dielectricTransmissionProbability(mat, fresnel.x),
                              1.0f - dielectricTransmissionProbability(mat, fresnel.y),
                              1.0f - dielectricTransmissionProbability(mat, fresnel.z),
                              1.0f - dielectricTransmissionProbability(mat, fresnel.w)) *
                       basePdf;
    } else {   // opaque: reflection probability 1 at every wavelength, colour from Schlick's F0: the four exact Fresnel terms are not needed
        techniquePdf = float4(basePdf);
    }
    if (!anyGreater(techniquePdf, 0.0f)) return float4(0.0f);
    return dielectricReflectionColor4(t, mat, woDotWm, fresnel, wl) * (Dm * G / max(4.0f * cosTheta(wo) * cosTheta(wi), GGX_EPSILON));
}
VK_D float evalSpectralDielectricTransmissionLane(const BSDFMaterial& mat, float3 wo, float3 wi, uint frontFace,
                                                    const GGXParams& p, float coatAttenuation, float transmissionColor,
                                                    float lambdaNm, float& pdf) {
    pdf = 0.0f;
    if (mat.transmission <= 0.0f || cosTheta(wo) <= 0.0f || cosTheta(wi) >= 0.0f) return 0.0f;
    if (dielectricIsIdentity(mat)) {
        if (dot(wi, -wo) < 0.9999f) return 0.0f;
        pdf = 1.0f;
        return coatAttenuation * transmissionColor / max(absCosTheta(wi), GGX_EPSILON);
    }
    float etap = interfaceEta(mat, frontFace, lambdaNm, 1u);
    float3 wm;
    float woDotWm, wiDotWm, denom;
    if (!dielectricTransmissionHalfVector(wo, wi, etap, wm, woDotWm, wiDotWm, denom)) return 0.0f;
    float fresnel = fresnelDielectric(woDotWm, etap);
    float tp = dielectricTransmissionProbability(mat, fresnel);
    if (tp <= 0.0f) return 0.0f;
    float Dm = ggxDistribution(wm, p.alpha);
    float G = ggxMasking(wo, wi, p.alpha);
    float denom2 = denom * denom;
    float transportScale = 1.0f / max(etap * etap, GGX_EPSILON);
    float transmissionTerm = abs((wiDotWm * woDotWm) / max(absCosTheta(wi) * cosTheta(wo) * denom2, GGX_EPSILON));
    float dwmDwi = abs(wiDotWm) / max(denom2, GGX_EPSILON);
    pdf = tp * ggxVisibleNormalPdf(wo, wm, p) * dwmDwi;
    return coatAttenuation * transmissionColor * ((1.0f - fresnel) * Dm * G * transmissionTerm * transportScale);
}
VK_NOINLINE float4 evalSpectralDielectricTransmission(const SpectralTables& t, const BSDFMaterial& mat, float3 wo, float3 wi,
                                                 uint frontFace, const GGXParams& p, float coatAttenuation, float4 wl,
                                                 float4& techniquePdf) {
    float4 value(0.0f);
    techniquePdf = float4(0.0f);
    float4 tc = spectralScalarFromLinearSrgb4(t, saturate(bsdfTransmissionColor(mat)), wl) * ((1.0f - mat.metallic) * mat.transmission);
    value.x = evalSpectralDielectricTransmissionLane(mat, wo, wi, frontFace, p, coatAttenuation, tc.x, wl.x, techniquePdf.x);
    value.y = evalSpectralDielectricTransmissionLane(mat, wo, wi, frontFace, p, coatAttenuation, tc.y, wl.y, techniquePdf.y);
    value.z = evalSpectralDielectricTransmissionLane(mat, wo, wi, frontFace, p, coatAttenuation, tc.z, wl.z, techniquePdf.z);
    value.w = evalSpectralDielectricTransmissionLane(mat, wo, wi, frontFace, p, coatAttenuation, tc.w, wl.w, techniquePdf.w);
    return value;
}

// ---- bsdf/lobes/dielectric/sample.slang:7-53 ----------------------------------------------------------------------
struct DielectricSample {
    float3 wi = float3(0.0f);
    uint isTransmission = 0u;
};
VK_D bool sampleDielectric(const BSDFMaterial& mat, float3 wo, uint frontFace, const GGXParams& p, float lambdaNm,
                             uint spectralMode, uint& rng, DielectricSample& s) {
    s = DielectricSample();
    if (cosTheta(wo) <= 0.0f) return false;
    if (dielectricIsIdentity(mat)) {
        s.wi = -wo;
        s.isTransmission = mat.transmission > 0.0f ? 1u : 0u;
        return mat.transmission > 0.0f;
    }
    float3 wm;
    float ux = rand(rng);
    float uy = rand(rng);
    if (!sampleGGXVNDF(wo, p, float2(ux, uy), wm)) return false;
    float woDotWm = dot(wo, wm);
    if (woDotWm <= 0.0f) return false;
    float tp = mat.transmission > 0.0f ? dielectricTransmissionProbability(mat, dielectricFresnel(mat, frontFace, woDotWm, lambdaNm, spectralMode)) : 0.0f;
    if (tp > 0.0f && rand(rng) < tp) {
        float eta = interfaceRefractionEta(mat, frontFace, lambdaNm, spectralMode);
        float3 wi = refract(-wo, wm, eta);
        if (dot(wi, wi) > 0.0f && cosTheta(wi) < 0.0f) {
            s.wi = wi;
            s.isTransmission = 1u;
            return true;
        }
    }
    s.wi = 2.0f * woDotWm * wm - wo;
    s.isTransmission = 0u;
    return cosTheta(s.wi) > 0.0f;
}

// ---- bsdf/principled/attenuation.slang:4-66 -----------------------------------------------------------------------
VK_D float scalarSchlickFresnel(float cosT, float f0) { return lerp(f0, 1.0f, schlickWeight(abs(cosT))); }
VK_D float coatDirectionalAttenuation(const BSDFMaterial& m, float3 w) {
    if (m.clearcoat <= 0.0f) return 1.0f;
    if (absCosTheta(w) <= 0.0f) return 0.0f;
    float coatFresnel = 0.25f * m.clearcoat * scalarSchlickFresnel(absCosTheta(w), 0.04f);
    return 1.0f - saturate(coatFresnel);
}
VK_D float sheenDirectionalAttenuation(float3 sheenColor, float3 w, float sheenRoughness) {
    return sheenLayerAttenuation(maxComponent(sheenColor), cosTheta(w), sheenRoughness);
}
VK_D float2 dielectricDirectionalAlbedoAB(float roughness, float cosT) {
    float cr = saturate(roughness);
    float cc = saturate(cosT);
    float4 c0(-1.0f, -0.0275f, -0.572f, 0.022f);
    float4 c1(1.0f, 0.0425f, 1.04f, -0.04f);
    float4 r = cr * c0 + c1;
    float a004 = min(r.x * r.x, exp2f(-9.28f * cc)) * r.x + r.y;
    return saturate(float2(-1.04f, 1.04f) * a004 + float2(r.z, r.w));
}
VK_D float3 dielectricDirectionalAlbedo(const BSDFMaterial& m, float3 w) {
    if (absCosTheta(w) <= 0.0f) return float3(1.0f);
    float2 ab = dielectricDirectionalAlbedoAB(m.roughness, absCosTheta(w));
    return saturate(bsdfDielectricSpecularF0(m) * ab.x + ab.y);
}
VK_D float3 dielectricDirectionalAttenuation(const BSDFMaterial& m, float3 w) {
    return saturate(1.0f - dielectricDirectionalAlbedo(m, w));
}
VK_D float reflectionStackAttenuation(const BSDFMaterial& m, float viewCoatAttenuation, float3 wi) {
    return viewCoatAttenuation * coatDirectionalAttenuation(m, wi);
}
VK_D float materialTransmissionStackAttenuation(const BSDFMaterial& m, float3 wo) {
    return sheenDirectionalAttenuation(bsdfSheenColor(m), wo, m.sheenRoughness) * coatDirectionalAttenuation(m, wo);
}

// ---- bsdf/principled/weights.slang:4-62, types.slang:4-38 ---------------------------------------------------------
struct BSDFBranchWeights {
    float sheen = 0, coat = 0, metal = 0, dielectric = 0, diffuse = 0, subsurface = 0;
};
VK_D BSDFBranchWeights normalizeBranchWeights(BSDFBranchWeights w) {
    float total = w.sheen + w.coat + w.metal + w.dielectric + w.diffuse + w.subsurface;
    if (total <= 0.0f) {
        w = BSDFBranchWeights();
        w.diffuse = 1.0f;
        return w;
    }
    w.sheen /= total;
    w.coat /= total;
    w.metal /= total;
    w.dielectric /= total;
    w.diffuse /= total;
    w.subsurface /= total;
    return w;
}
VK_D BSDFBranchWeights makeBSDFBranchWeights(const BSDFMaterial& m, float3 wo, uint frontFace) {
    BSDFBranchWeights w;
    float3 diffuseColor = bsdfDiffuseColor(m);
    float3 sheenColor = bsdfSheenColor(m);
    float nonMetal = 1.0f - m.metallic;
    float baseColorWeight = max(linearSrgbLuminance(m.baseColor), 1e-3f);
    float diffuseColorWeight = max(linearSrgbLuminance(diffuseColor), 1e-3f);
    float baseScatter = nonMetal * (1.0f - m.transmission);
    float transmissionWeight = m.transmission * max(linearSrgbLuminance(bsdfTransmissionColor(m)), 1e-3f);
    float dielectricSpecularWeight = nonMetal * max(bsdfDielectricSpecularF0Luminance(m) + transmissionWeight, 0.0f);
    w.sheen = linearSrgbLuminance(sheenColor);
    if (m.transmission > 0.0f && frontFace == 0u) {
        w.sheen = 0.0f;
        w.coat = 0.0f;
        w.metal = m.metallic * baseColorWeight;
        w.dielectric = dielectricSpecularWeight;
        w.diffuse = 0.0f;
        w.subsurface = 0.0f;
        return normalizeBranchWeights(w);
    }
    float vs = sheenDirectionalAttenuation(sheenColor, wo, m.sheenRoughness);
    float vc = coatDirectionalAttenuation(m, wo);
    float da = linearSrgbLuminance(dielectricDirectionalAttenuation(m, wo));
    w.coat = vs * m.clearcoat * 0.25f;
    w.metal = vs * vc * m.metallic * baseColorWeight;
    w.dielectric = vs * vc * dielectricSpecularWeight;
    w.diffuse = vs * vc * da * baseScatter * (1.0f - m.subsurface) * diffuseColorWeight;
    w.subsurface = vs * vc * da * baseScatter * m.subsurface * diffuseColorWeight;
    return normalizeBranchWeights(w);
}
struct BSDFState {
    BSDFMaterial material;
    float3 wo;
    GGXParams ggx;
    float wavelengthNm = 0.0f;
    BSDFBranchWeights sampleWeights;
    // Hero mode evaluates the closure twice per vertex with the same (material, wo): once for the light sample, once for the sampled
    // direction. The two spectral upsamplings that do not depend on wi are kept from the first evaluation (the state lives in the
    // kernel's local frame; each lookup is ~100 instructions and 8 gathered 128-bit loads).
    mutable float cachedDiffuse4[4], cachedVd4[4];   // plain floats (no alignment demands on the state record, which k_shade keeps in shared memory)
    // One word for the small fields: k_shade keeps this record in shared memory and every 8 bytes of it cost 4 KB of L1 per SM (the hero
    // kernel's spill slots live there). The clearcoat constant is recomputed where the lobe runs (makeClearcoatParams) for the same reason.
    uint frontFace : 1;
    uint spectralMode : 1;
    mutable uint cachedMask : 2;   // bit 0: cachedDiffuse4 valid, bit 1: cachedVd4 valid
    uint memoIndex : 28;           // first SpectralTables::materialMemo entry of this vertex's material (k_shade), or SPECTRAL_MEMO_NONE
    static constexpr uint SPECTRAL_MEMO_NONE = 0x0fffffffu;
    __device__ BSDFState() {}
    __device__ BSDFState(const BSDFMaterial& m, float3 wo_, uint ff, float wl, uint sm)
        : material(m), wo(wo_), wavelengthNm(wl), frontFace(ff), spectralMode(sm), cachedMask(0u), memoIndex(SPECTRAL_MEMO_NONE) {
        ggx = makeGGXParams(m, wo_);
        sampleWeights = makeBSDFBranchWeights(m, wo_, ff);
    }
};
VK_D const SpectralMemoEntry* spectralMemoOf(const SpectralTables& t, const BSDFState& s, uint slot) {
    return (t.materialMemo && s.memoIndex != BSDFState::SPECTRAL_MEMO_NONE) ? t.materialMemo + (s.memoIndex + slot) : nullptr;
}
VK_D bool useInteriorDielectricInterface(const BSDFState& s) { return s.material.transmission > 0.0f && s.frontFace == 0u; }

// ---- bsdf/principled/eval_rgb.slang:4-140 -------------------------------------------------------------------------
VK_D BSDFEval evalScalarReflectionStack(const SpectralTables& t, const BSDFState& s, float3 wi, uint spectralMode) {
    BSDFEval e;
    if (cosTheta(wi) <= 0.0f) return e;
    const BSDFMaterial& m = s.material;
    bool interior = useInteriorDielectricInterface(s);
    float3 sheenValue(0.0f), coatValue(0.0f), metalValue(0.0f), dielectricValue(0.0f), substrateValue(0.0f), sheenColor(0.0f);
    if (s.sampleWeights.sheen > 0.0f || !interior) {
        sheenColor = bsdfSheenColor(m);
        if (spectralMode != 0u) sheenColor = float3(spectralScalarFromLinearSrgb(t, saturate(sheenColor), s.wavelengthNm));
    }
    if (s.sampleWeights.sheen > 0.0f) {
        BSDFEval sh = evalSheen(sheenColor, m.sheenRoughness, s.wo, wi);
        sheenValue = sh.value;
        e.pdf += s.sampleWeights.sheen * sh.pdf;
    }
    if (s.sampleWeights.coat > 0.0f) {
        BSDFEval c = evalClearcoat(m.clearcoat, s.wo, wi, makeClearcoatParams(m));
        coatValue = c.value;
        e.pdf += s.sampleWeights.coat * c.pdf;
    }
    if (s.sampleWeights.metal > 0.0f) {
        BSDFEval mt = evalGGX(t, m, s.wo, wi, s.ggx, s.wavelengthNm, spectralMode);
        metalValue = mt.value;
        e.pdf += s.sampleWeights.metal * mt.pdf;
    }
    if (s.sampleWeights.dielectric > 0.0f) {
        BSDFEval d = evalDielectricReflection(t, m, s.wo, wi, s.frontFace, s.ggx, s.wavelengthNm, spectralMode);
        dielectricValue = d.value;
        e.pdf += s.sampleWeights.dielectric * d.pdf;
    }
    if (s.sampleWeights.diffuse > 0.0f || s.sampleWeights.subsurface > 0.0f) {
        float3 diffuseColor = bsdfDiffuseColor(m);
        if (spectralMode != 0u) diffuseColor = float3(spectralScalarFromLinearSrgb(t, spectralMemoOf(t, s, SPECTRAL_MEMO_DIFFUSE), saturate(diffuseColor), s.wavelengthNm));
        if (s.sampleWeights.diffuse > 0.0f) {
            float3 dw = diffuseColor * ((1.0f - m.transmission) * (1.0f - m.subsurface));
            BSDFEval d = m.diffuseRoughness <= 0.0f ? evalLambertian(dw, wi) : evalOrenNayar(dw, m.diffuseRoughness, s.wo, wi);
            substrateValue += d.value;
            e.pdf += s.sampleWeights.diffuse * d.pdf;
        }
        if (s.sampleWeights.subsurface > 0.0f) {
            float3 sc = diffuseColor * (1.0f - m.transmission);
            BSDFEval ss = evalFakeSubsurface(sc * m.subsurface, m.diffuseRoughness, s.wo, wi);
            substrateValue += ss.value;
            e.pdf += s.sampleWeights.subsurface * ss.pdf;
        }
    }
    if (interior) {
        float nonMetal = 1.0f - m.metallic;
        e.value = m.metallic * metalValue + nonMetal * dielectricValue;
        return e;
    }
    float vs = sheenDirectionalAttenuation(sheenColor, s.wo, m.sheenRoughness);
    float vc = coatDirectionalAttenuation(m, s.wo);
    float3 vd = dielectricDirectionalAttenuation(m, s.wo);
    float nonMetal = 1.0f - m.metallic;
    float3 baseValue = coatValue + reflectionStackAttenuation(m, vc, wi) *
                                       (m.metallic * metalValue + nonMetal * (dielectricValue + vd * substrateValue));
    e.value = sheenValue + vs * baseValue;
    return e;
}
VK_NOINLINE BSDFEval evalBSDFMode(const SpectralTables& t, const BSDFState& s, float3 wi, uint spectralMode) {
    if (cosTheta(wi) > 0.0f) return evalScalarReflectionStack(t, s, wi, spectralMode);
    if (!(s.material.transmission > 0.0f)) return BSDFEval();   // opaque: the transmission lobe is zero with zero density (checked again inside)
    BSDFEval tr = evalDielectricTransmission(
        t, s.material, s.wo, wi, s.frontFace, s.ggx,
        useInteriorDielectricInterface(s) ? 1.0f : materialTransmissionStackAttenuation(s.material, s.wo), s.wavelengthNm,
        spectralMode);
    tr.pdf *= s.sampleWeights.dielectric;
    return tr;
}
VK_D BSDFEval evalBSDF(const SpectralTables& t, const BSDFState& s, float3 wi) { return evalBSDFMode(t, s, wi, 0u); }
VK_D BSDFEval evalSingleWavelengthBSDF(const SpectralTables& t, const BSDFState& s, float3 wi) { return evalBSDFMode(t, s, wi, 1u); }

// ---- bsdf/principled/eval_spectral.slang:4-115 --------------------------------------------------------------------
VK_D float4 evalSpectralReflectionStack(const SpectralTables& t, const BSDFState& s, float3 wi, float4 wl, float4& techniquePdf) {
    techniquePdf = float4(0.0f);
    if (cosTheta(wi) <= 0.0f) return float4(0.0f);
    const BSDFMaterial& m = s.material;
    bool interior = useInteriorDielectricInterface(s);
    float4 sheenValue(0.0f), metalValue(0.0f), dielectricValue(0.0f), substrateValue(0.0f), sheenColor(0.0f);
    float coatValue = 0.0f;
    if (s.sampleWeights.sheen > 0.0f || !interior) sheenColor = spectralScalarFromLinearSrgb4(t, saturate(bsdfSheenColor(m)), wl);
    if (s.sampleWeights.sheen > 0.0f) {
        BSDFEval sh = evalSheen(float3(1.0f), m.sheenRoughness, s.wo, wi);
        sheenValue = sheenColor * sh.value.x;
        techniquePdf += float4(s.sampleWeights.sheen * sh.pdf);
    }
    if (s.sampleWeights.coat > 0.0f) {
        BSDFEval c = evalClearcoat(m.clearcoat, s.wo, wi, makeClearcoatParams(m));
        coatValue = c.value.x;
        techniquePdf += float4(s.sampleWeights.coat * c.pdf);
    }
    if (s.sampleWeights.metal > 0.0f) {
        float metalPdf = 0.0f;
        metalValue = evalSpectralGGX(t, m, s.wo, wi, s.ggx, wl, metalPdf);
        techniquePdf += float4(s.sampleWeights.metal * metalPdf);
    }
    if (s.sampleWeights.dielectric > 0.0f) {
        float4 dp(0.0f);
        dielectricValue = evalSpectralDielectricReflection(t, m, s.wo, wi, s.frontFace, s.ggx, wl, dp);
        techniquePdf += s.sampleWeights.dielectric * dp;
    }
    if (s.sampleWeights.diffuse > 0.0f || s.sampleWeights.subsurface > 0.0f) {
        if (!(s.cachedMask & 1u)) {
            const float4 c = spectralScalarFromLinearSrgb4(t, spectralMemoOf(t, s, SPECTRAL_MEMO_DIFFUSE), saturate(bsdfDiffuseColor(m)), wl);
            s.cachedDiffuse4[0] = c.x; s.cachedDiffuse4[1] = c.y; s.cachedDiffuse4[2] = c.z; s.cachedDiffuse4[3] = c.w;
            s.cachedMask |= 1u;
        }
        const float4 diffuseColor(s.cachedDiffuse4[0], s.cachedDiffuse4[1], s.cachedDiffuse4[2], s.cachedDiffuse4[3]);
        if (s.sampleWeights.diffuse > 0.0f) {
            float ds = (1.0f - m.transmission) * (1.0f - m.subsurface);
            BSDFEval d = m.diffuseRoughness <= 0.0f ? evalLambertian(float3(ds), wi) : evalOrenNayar(float3(ds), m.diffuseRoughness, s.wo, wi);
            substrateValue += diffuseColor * d.value.x;
            techniquePdf += float4(s.sampleWeights.diffuse * d.pdf);
        }
        if (s.sampleWeights.subsurface > 0.0f) {
            float ss = (1.0f - m.transmission) * m.subsurface;
            BSDFEval sub = evalFakeSubsurface(float3(ss), m.diffuseRoughness, s.wo, wi);
            substrateValue += diffuseColor * sub.value.x;
            techniquePdf += float4(s.sampleWeights.subsurface * sub.pdf);
        }
    }
    if (interior) {
        float nonMetal = 1.0f - m.metallic;
        return m.metallic * metalValue + nonMetal * dielectricValue;
    }
    float vc = coatDirectionalAttenuation(m, s.wo);
    const float4 vs = anyGreater(sheenColor, 0.0f) ? saturate(1.0f - sheenColor * sheenDirectionalAlbedo(cosTheta(s.wo), m.sheenRoughness)) : float4(1.0f);
    if (!(s.cachedMask & 2u)) {
        const float4 c = spectralScalarFromLinearSrgb4(t, saturate(dielectricDirectionalAttenuation(m, s.wo)), wl);
        s.cachedVd4[0] = c.x; s.cachedVd4[1] = c.y; s.cachedVd4[2] = c.z; s.cachedVd4[3] = c.w;
        s.cachedMask |= 2u;
    }
    const float4 vd(s.cachedVd4[0], s.cachedVd4[1], s.cachedVd4[2], s.cachedVd4[3]);
    float nonMetal = 1.0f - m.metallic;
    float ra = reflectionStackAttenuation(m, vc, wi);
    float4 baseValue = float4(coatValue) + ra * (m.metallic * metalValue + nonMetal * (dielectricValue + vd * substrateValue));
    return sheenValue + vs * baseValue;
}
VK_NOINLINE float4 evalSpectralBSDF(const SpectralTables& t, const BSDFState& s, float3 wi, float4 wl, float4& techniquePdf) {
    if (cosTheta(wi) > 0.0f) return evalSpectralReflectionStack(t, s, wi, wl, techniquePdf);
    if (!(s.material.transmission > 0.0f)) {   // opaque: every lane of the transmission lobe is zero with zero density; skip its colour lookup
        techniquePdf = float4(0.0f);
        return float4(0.0f);
    }
    float coatAtt = useInteriorDielectricInterface(s) ? 1.0f : materialTransmissionStackAttenuation(s.material, s.wo);
    float4 v = evalSpectralDielectricTransmission(t, s.material, s.wo, wi, s.frontFace, s.ggx, coatAtt, wl, techniquePdf);
    techniquePdf *= s.sampleWeights.dielectric;
    return v;
}

// ---- bsdf/principled/sample.slang:4-51 ----------------------------------------------------------------------------
struct BSDFDirectionSample {
    float3 wi = float3(0.0f);
    uint isTransmission = 0u;
};
VK_NOINLINE bool sampleBSDFDirection(const BSDFState& s, uint& rng, BSDFDirectionSample& out) {
    out = BSDFDirectionSample();
    float selector = rand(rng);
    if (selector < s.sampleWeights.sheen) return sampleSheen(s.wo, s.material.sheenRoughness, rng, out.wi);
    selector -= s.sampleWeights.sheen;
    if (selector < s.sampleWeights.coat) return sampleClearcoat(s.wo, makeClearcoatParams(s.material), rng, out.wi);
    selector -= s.sampleWeights.coat;
    if (selector < s.sampleWeights.metal) return sampleGGX(s.wo, s.ggx, rng, out.wi);
    selector -= s.sampleWeights.metal;
    if (selector < s.sampleWeights.dielectric) {
        DielectricSample ds;
        if (!sampleDielectric(s.material, s.wo, s.frontFace, s.ggx, s.wavelengthNm, s.spectralMode, rng, ds)) return false;
        out.wi = ds.wi;
        out.isTransmission = ds.isTransmission;
        return true;
    }
    out.wi = sampleCosineHemisphere(rng); // diffuse and subsurface branches are identical (sample.slang:44-50)
    return true;
}
// bsdf/sample_rgb.slang:6-23
VK_D BSDFSample sampleBSDF(const SpectralTables& t, const BSDFState& s, const ShadingBasis& basis, uint& rng) {
    BSDFSample out;
    if (cosTheta(s.wo) <= 0.0f) return out;
    BSDFDirectionSample d;
    if (!sampleBSDFDirection(s, rng, d)) return out;
    out.isTransmission = d.isTransmission;
    out.wi = localToWorld(d.wi, basis);
    BSDFEval e = s.spectralMode != 0u ? evalSingleWavelengthBSDF(t, s, d.wi) : evalBSDF(t, s, d.wi);
    if (e.pdf <= 0.0f) return out;
    out.pdf = e.pdf;
    out.weight = e.value * absCosTheta(d.wi) / out.pdf;
    return out;
}
// bsdf/sample_spectral.slang:6-24
VK_D SpectralBSDFSample sampleSpectralBSDF(const SpectralTables& t, const BSDFState& s, const ShadingBasis& basis, float4 wl, uint& rng) {
    SpectralBSDFSample out;
    if (cosTheta(s.wo) <= 0.0f) return out;
    BSDFDirectionSample d;
    if (!sampleBSDFDirection(s, rng, d)) return out;
    out.isTransmission = d.isTransmission;
    out.wi = localToWorld(d.wi, basis);
    float4 tp(0.0f);
    float4 v = evalSpectralBSDF(t, s, d.wi, wl, tp);
    float sampledPdf = tp.x;
    if (!(sampledPdf > 0.0f)) return out;
    out.weight = v * absCosTheta(d.wi) / sampledPdf;
    out.techniquePdf = tp;
    return out;
}

// ---- bsdf/principled/denoiser.slang:4-46 --------------------------------------------------------------------------
VK_D bool materialDenoiserShouldFollowSpecularHit(const BSDFMaterial& m, uint frontFace) {
    const float EPS = 1e-3f;
    float diffuseWeight = maxComponent(bsdfDiffuseColor(m)) * (1.0f - m.transmission);
    float subsurfaceWeight = diffuseWeight * m.subsurface;
    float sheenLayerWeight = maxComponent(bsdfSheenColor(m)) * (1.0f - m.transmission);
    float clearcoatLayerWeight = frontFace != 0u ? m.clearcoat : 0.0f;
    bool hasVisibleBaseLayer = diffuseWeight > EPS || subsurfaceWeight > EPS || sheenLayerWeight > EPS || clearcoatLayerWeight > EPS;
    if (hasVisibleBaseLayer) return false;
    if (m.transmission > EPS && dielectricIsIdentity(m)) return true;
    if (m.roughness > 0.02f) return false;
    return m.transmission > EPS || m.metallic >= 1.0f - EPS || maxComponent(bsdfDielectricSpecularF0(m)) > EPS;
}
VK_D float3 materialDenoiserAlbedo(const BSDFMaterial& m) {
    float nonMetal = 1.0f - m.metallic;
    float3 diffuseColor = bsdfDiffuseColor(m);
    float3 sheenColor = bsdfSheenColor(m);
    float3 dF0 = bsdfDielectricSpecularF0(m);
    float3 mF0 = materialHasConductor(m) ? fresnelConductor(1.0f, max(m.eta, float3(1e-3f)), max(m.k, float3(0.0f)))
                                         : saturate(m.baseColor);
    float diffuseWeight = nonMetal * (1.0f - m.transmission) * (1.0f - m.subsurface);
    float subsurfaceWeight = nonMetal * (1.0f - m.transmission) * m.subsurface;
    float specularWeight = nonMetal * (1.0f - m.transmission) * bsdfDielectricSpecularF0Luminance(m);
    float transmissionWeight = nonMetal * m.transmission;
    float metallicWeight = m.metallic;
    float sheenLayerWeight = maxComponent(sheenColor) * nonMetal * (1.0f - m.transmission);
    float total = diffuseWeight + subsurfaceWeight + specularWeight + transmissionWeight + metallicWeight + sheenLayerWeight;
    if (total <= 0.0f) return saturate(diffuseColor);
    float3 albedo = diffuseColor * diffuseWeight + diffuseColor * subsurfaceWeight + dF0 * specularWeight +
                    float3(1.0f) * transmissionWeight + mF0 * metallicWeight + sheenColor * sheenLayerWeight;
    return saturate(albedo / total);
}

// ---- film/tonemap.slang:9-32 --------------------------------------------------------------------------------------
VK_D float srgbEncodeScalar(float v) {
    if (v <= 0.0031308f) return 12.92f * v;
    return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
VK_D float3 encodeDisplayColor(float3 c) {
    c = saturate(c);
    return float3(srgbEncodeScalar(c.x), srgbEncodeScalar(c.y), srgbEncodeScalar(c.z));
}
VK_D float3 toneMapACES(float3 c) { return (c * (2.51f * c + 0.03f)) / (c * (2.43f * c + 0.59f) + 0.14f); }
VK_D float3 mapSceneColorToDisplay(const SceneData& scene, float3 color) {
    float3 c = max(color * scene.exposure, float3(0.0f));
    if (VKRT_RENDER_SETTINGS_TONE(scene.packedRenderSettings) == VKRT_TONE_MAPPING_MODE_ACES) c = toneMapACES(c);
    return encodeDisplayColor(c);
}

// ---- geometry/packing.slang:13-60 ---------------------------------------------------------------------------------
VK_D float3 decodeOct(float2 p) {
    float z = 1.0f - abs(p.x) - abs(p.y);
    if (z < 0.0f) {
        float oldX = p.x;
        p.x = (1.0f - abs(p.y)) * (oldX >= 0.0f ? 1.0f : -1.0f);
        p.y = (1.0f - abs(oldX)) * (p.y >= 0.0f ? 1.0f : -1.0f);
    }
    return normalize(float3(p.x, p.y, z));
}
VK_D float3 unpackOctNormal(uint packed) {
    float2 p = float2(float(int32_t(packed << 16) >> 16) / 32767.0f, float(int32_t(packed) >> 16) / 32767.0f);
    return decodeOct(p);
}
VK_D float4 unpackColorRGBA8(uint packed) {
    const float s = 1.0f / 255.0f;
    return float4(float(packed & 0xffu) * s, float((packed >> 8) & 0xffu) * s, float((packed >> 16) & 0xffu) * s,
                  float((packed >> 24) & 0xffu) * s);
}
VK_D float unpackSnorm15(uint value) {
    int32_t sv = int32_t(value & 0x7fffu);
    if ((sv & 0x4000) != 0) sv |= ~int32_t(0x7fff);
    return clamp(float(sv) / 16383.0f, -1.0f, 1.0f);
}
VK_D float4 unpackOctTangent(uint packed) {
    float2 p = float2(unpackSnorm15(packed), unpackSnorm15(packed >> 15u));
    float handedness = (packed & 0x80000000u) != 0u ? -1.0f : 1.0f;
    return float4(decodeOct(p), handedness);
}

// ---- geometry/surface/transform.slang:6-55 ------------------------------------------------------------------------
static constexpr float SURFACE_DEG_TO_RAD = 0.01745329251994329577f;
static constexpr float SURFACE_SCALE_EPSILON = 1e-6f;
// sin/cos of a mesh's Euler rotation. The reference's shader re-evaluates these six transcendentals in every
// rotateMeshVector call (three calls per hit, surface/transform.slang:9-29); they depend on MeshInfo only, so they are
// evaluated once per mesh by k_mesh_trig (same device sinf/cosf, same bits) and fetched per hit.
struct MeshTrig {
    float sx, cx, sy, cy, sz, cz, pad0, pad1;
};
VK_D MeshTrig makeMeshTrig(float3 rotationDegrees) {
    float3 r = rotationDegrees * SURFACE_DEG_TO_RAD;
    MeshTrig t;
    t.sx = sinf(r.x); t.cx = cosf(r.x);
    t.sy = sinf(r.y); t.cy = cosf(r.y);
    t.sz = sinf(r.z); t.cz = cosf(r.z);
    t.pad0 = t.pad1 = 0.0f;
    return t;
}
VK_D float3 rotateMeshVector(float3 v, const MeshTrig& t) {
    float3 a(v.x, t.cx * v.y - t.sx * v.z, t.sx * v.y + t.cx * v.z);          // rotateX
    float3 b(t.cy * a.x + t.sy * a.z, a.y, -t.sy * a.x + t.cy * a.z);         // rotateY
    return float3(t.cz * b.x - t.sz * b.y, t.sz * b.x + t.cz * b.y, b.z);     // rotateZ
}
VK_D float safeSignedReciprocal(float v) {
    if (abs(v) > SURFACE_SCALE_EPSILON) return 1.0f / v;
    return v < 0.0f ? -1.0f / SURFACE_SCALE_EPSILON : 1.0f / SURFACE_SCALE_EPSILON;
}
VK_D float3 meshScale(const MeshInfo& m) { return float3(m.scale[0], m.scale[1], m.scale[2]); }
VK_D float3 meshRotation(const MeshInfo& m) { return float3(m.rotation[0], m.rotation[1], m.rotation[2]); }
VK_D float3 meshTransformVector(const MeshInfo& m, const MeshTrig& t, float3 v) { return rotateMeshVector(v * meshScale(m), t); }
VK_D float3 meshTransformNormal(const MeshInfo& m, const MeshTrig& t, float3 n) {
    float3 inv(safeSignedReciprocal(m.scale[0]), safeSignedReciprocal(m.scale[1]), safeSignedReciprocal(m.scale[2]));
    return safeNormalize(rotateMeshVector(n * inv, t));
}
VK_D float surfaceTransformSign(const MeshInfo& m) { return m.scale[0] * m.scale[1] * m.scale[2] < 0.0f ? -1.0f : 1.0f; }

} // namespace vk
