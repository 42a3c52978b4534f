// vmath.cuh — HLSL/Slang-flavoured fp32 vector math for the device restatement of vkrt's shaders (src/shaders/**).
// Conventions (the Slang intrinsics are implementation-defined on a Vulkan driver; pinned here):
//   rsqrt(x) = 1.0f / sqrtf(x) (IEEE, not the approximate MUFU.RSQ), normalize(v) = v * rsqrt(dot(v,v)),
//   lerp(a,b,t) = a + (b - a) * t, saturate(x) = min(max(x,0),1), frac(x) = x - floorf(x).
// Compiled without --use_fast_math: division and sqrt are IEEE-rounded; FMA contraction is left to the compiler
// except in intersect.cuh / camera code, which pin every rounding with __fmul_rn/__fadd_rn.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define VK_HD __host__ __device__ __forceinline__
#define VK_HDM __host__ __device__

namespace vk {

typedef uint32_t uint;

struct alignas(8) float2 {
    float x, y;
    VK_HDM float2() : x(0), y(0) {}
    VK_HDM float2(float a) : x(a), y(a) {}
    VK_HDM float2(float a, float b) : x(a), y(b) {}
};
struct float3 {
    float x, y, z;
    VK_HDM float3() : x(0), y(0), z(0) {}
    VK_HDM float3(float a) : x(a), y(a), z(a) {}
    VK_HDM float3(float a, float b, float c) : x(a), y(b), z(c) {}
    VK_HDM float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    VK_HDM float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct alignas(16) float4 {
    float x, y, z, w;
    VK_HDM float4() : x(0), y(0), z(0), w(0) {}
    VK_HDM float4(float a) : x(a), y(a), z(a), w(a) {}
    VK_HDM float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    VK_HDM float4(float3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    VK_HDM float3 xyz() const { return float3(x, y, z); }
    VK_HDM float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    VK_HDM float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

#define ORC_OP2(T, op)                                                                   \
    VK_HD T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y); }                 \
    VK_HD T operator op(T a, float b) { return T(a.x op b, a.y op b); }                 \
    VK_HD T operator op(float a, T b) { return T(a op b.x, a op b.y); }
#define ORC_OP3(T, op)                                                                   \
    VK_HD T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y, a.z op b.z); }     \
    VK_HD T operator op(T a, float b) { return T(a.x op b, a.y op b, a.z op b); }       \
    VK_HD T operator op(float a, T b) { return T(a op b.x, a op b.y, a op b.z); }
#define ORC_OP4(T, op)                                                                            \
    VK_HD T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); }  \
    VK_HD T operator op(T a, float b) { return T(a.x op b, a.y op b, a.z op b, a.w op b); }      \
    VK_HD T operator op(float a, T b) { return T(a op b.x, a op b.y, a op b.z, a op b.w); }
ORC_OP2(float2, +) ORC_OP2(float2, -) ORC_OP2(float2, *) ORC_OP2(float2, /)
ORC_OP3(float3, +) ORC_OP3(float3, -) ORC_OP3(float3, *) ORC_OP3(float3, /)
ORC_OP4(float4, +) ORC_OP4(float4, -) ORC_OP4(float4, *) ORC_OP4(float4, /)
VK_HD float2 operator-(float2 a) { return float2(-a.x, -a.y); }
VK_HD float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
VK_HD float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
VK_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
VK_HD float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
VK_HD float3& operator*=(float3& a, float b) { a = a * b; return a; }
VK_HD float3& operator/=(float3& a, float b) { a = a / b; return a; }
VK_HD float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
VK_HD float4& operator*=(float4& a, float4 b) { a = a * b; return a; }
VK_HD float4& operator*=(float4& a, float b) { a = a * b; return a; }
VK_HD float4& operator/=(float4& a, float b) { a = a / b; return a; }

VK_HD float fmin2(float a, float b) { return a < b ? a : b; }
VK_HD float fmax2(float a, float b) { return a > b ? a : b; }
// HLSL min/max: NaN handling irrelevant on this path (inputs sanitised); written as compare+select.
VK_HD float min(float a, float b) { return fmin2(a, b); }
VK_HD float max(float a, float b) { return fmax2(a, b); }
VK_HD uint min(uint a, uint b) { return a < b ? a : b; }
VK_HD uint max(uint a, uint b) { return a > b ? a : b; }
VK_HD float2 max(float2 a, float2 b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
VK_HD float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
VK_HD float4 max(float4 a, float4 b) { return float4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
VK_HD float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
VK_HD float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
VK_HD float saturate(float v) { return clamp(v, 0.0f, 1.0f); }
VK_HD float2 saturate(float2 v) { return float2(saturate(v.x), saturate(v.y)); }
VK_HD float3 saturate(float3 v) { return float3(saturate(v.x), saturate(v.y), saturate(v.z)); }
VK_HD float4 saturate(float4 v) { return float4(saturate(v.x), saturate(v.y), saturate(v.z), saturate(v.w)); }
VK_HD float abs(float v) { return fabsf(v); }
VK_HD float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
VK_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VK_HD float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
VK_HD float3 cross(float3 a, float3 b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
VK_HD float rsqrt(float v) { return 1.0f / sqrtf(v); }
VK_HD float sqrt(float v) { return sqrtf(v); }
VK_HD float3 sqrt(float3 v) { return float3(sqrtf(v.x), sqrtf(v.y), sqrtf(v.z)); }
VK_HD float length(float2 v) { return sqrtf(dot(v, v)); }
VK_HD float length(float3 v) { return sqrtf(dot(v, v)); }
VK_HD float3 normalize(float3 v) { return v * rsqrt(dot(v, v)); }
VK_HD float lerp(float a, float b, float t) { return a + (b - a) * t; }
VK_HD float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
VK_HD float frac(float v) { return v - floorf(v); }
VK_HD float4 frac(float4 v) { return float4(frac(v.x), frac(v.y), frac(v.z), frac(v.w)); }
VK_HD float3 exp(float3 v) { return float3(expf(v.x), expf(v.y), expf(v.z)); }
VK_HD float4 exp(float4 v) { return float4(expf(v.x), expf(v.y), expf(v.z), expf(v.w)); }
VK_HD float3 log(float3 v) { return float3(logf(v.x), logf(v.y), logf(v.z)); }
VK_HD float4 log(float4 v) { return float4(logf(v.x), logf(v.y), logf(v.z), logf(v.w)); }
VK_HD bool anyGreater(float3 v, float t) { return v.x > t || v.y > t || v.z > t; }
VK_HD bool anyGreater(float4 v, float t) { return v.x > t || v.y > t || v.z > t || v.w > t; }
VK_HD bool anyLess(float3 v, float t) { return v.x < t || v.y < t || v.z < t; }
VK_HD float maxComponent(float3 v) { return max(v.x, max(v.y, v.z)); }
VK_HD float maxComponent4(float4 v) { return max(max(v.x, v.y), max(v.z, v.w)); }
// GLSL/HLSL refract
VK_HD float3 refract(float3 I, float3 N, float eta) {
    float NdotI = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - NdotI * NdotI);
    if (k < 0.0f) return float3(0.0f);
    return eta * I - (eta * NdotI + sqrtf(k)) * N;
}

__device__ __forceinline__ uint asuint(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float asfloat(uint u) { return __uint_as_float(u); }

// binary16 conversions: hardware round-to-nearest-even (bit-identical to the oracle's software routine).
__device__ __forceinline__ uint16_t f32_to_f16(float v) { return __half_as_ushort(__float2half_rn(v)); }
__device__ __forceinline__ float f16_to_f32(uint16_t h) { return __half2float(__ushort_as_half(h)); }

} // namespace vk
