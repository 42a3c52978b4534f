// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; nothing under vkrt_b200/ may include this.
//
// Minimal HLSL/Slang-flavoured scalar vector math for the CPU restatement of vkrt's shaders.
// All arithmetic is plain IEEE fp32, one rounding per operation: compile with -ffp-contract=off.
// Conventions pinned here (the Slang intrinsics are implementation-defined on a Vulkan driver):
//   rsqrt(x)      = 1.0f / sqrtf(x)
//   normalize(v)  = v * rsqrt(dot(v,v))
//   lerp(a,b,t)   = a + (b - a) * t
//   saturate(x)   = min(max(x, 0), 1)
//   frac(x)       = x - floorf(x)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

typedef uint32_t uint;

struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float a) : x(a), y(a) {}
    float2(float a, float b) : x(a), y(b) {}
};
struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float a) : x(a), y(a), z(a) {}
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a) : x(a), y(a), z(a), w(a) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(float3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float3 xyz() const { return float3(x, y, z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

#define ORC_OP2(T, op)                                                                   \
    inline T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y); }                 \
    inline T operator op(T a, float b) { return T(a.x op b, a.y op b); }                 \
    inline T operator op(float a, T b) { return T(a op b.x, a op b.y); }
#define ORC_OP3(T, op)                                                                   \
    inline T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y, a.z op b.z); }     \
    inline T operator op(T a, float b) { return T(a.x op b, a.y op b, a.z op b); }       \
    inline T operator op(float a, T b) { return T(a op b.x, a op b.y, a op b.z); }
#define ORC_OP4(T, op)                                                                            \
    inline T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); }  \
    inline T operator op(T a, float b) { return T(a.x op b, a.y op b, a.z op b, a.w op b); }      \
    inline T operator op(float a, T b) { return T(a op b.x, a op b.y, a op b.z, a op b.w); }
ORC_OP2(float2, +) ORC_OP2(float2, -) ORC_OP2(float2, *) ORC_OP2(float2, /)
ORC_OP3(float3, +) ORC_OP3(float3, -) ORC_OP3(float3, *) ORC_OP3(float3, /)
ORC_OP4(float4, +) ORC_OP4(float4, -) ORC_OP4(float4, *) ORC_OP4(float4, /)
inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float b) { a = a * b; return a; }
inline float3& operator/=(float3& a, float b) { a = a / b; return a; }
inline float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
inline float4& operator*=(float4& a, float4 b) { a = a * b; return a; }
inline float4& operator*=(float4& a, float b) { a = a * b; return a; }
inline float4& operator/=(float4& a, float b) { a = a / b; return a; }

inline float fmin2(float a, float b) { return a < b ? a : b; }
inline float fmax2(float a, float b) { return a > b ? a : b; }
// HLSL min/max: NaN handling irrelevant on this path (inputs sanitised); written as compare+select.
inline float min(float a, float b) { return fmin2(a, b); }
inline float max(float a, float b) { return fmax2(a, b); }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float2 max(float2 a, float2 b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float4 max(float4 a, float4 b) { return float4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float saturate(float v) { return clamp(v, 0.0f, 1.0f); }
inline float2 saturate(float2 v) { return float2(saturate(v.x), saturate(v.y)); }
inline float3 saturate(float3 v) { return float3(saturate(v.x), saturate(v.y), saturate(v.z)); }
inline float4 saturate(float4 v) { return float4(saturate(v.x), saturate(v.y), saturate(v.z), saturate(v.w)); }
inline float abs(float v) { return std::fabs(v); }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float3 cross(float3 a, float3 b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float rsqrt(float v) { return 1.0f / std::sqrt(v); }
inline float sqrt(float v) { return std::sqrt(v); }
inline float3 sqrt(float3 v) { return float3(std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)); }
inline float length(float2 v) { return std::sqrt(dot(v, v)); }
inline float length(float3 v) { return std::sqrt(dot(v, v)); }
inline float3 normalize(float3 v) { return v * rsqrt(dot(v, v)); }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float frac(float v) { return v - std::floor(v); }
inline float4 frac(float4 v) { return float4(frac(v.x), frac(v.y), frac(v.z), frac(v.w)); }
inline float3 exp(float3 v) { return float3(std::exp(v.x), std::exp(v.y), std::exp(v.z)); }
inline float4 exp(float4 v) { return float4(std::exp(v.x), std::exp(v.y), std::exp(v.z), std::exp(v.w)); }
inline float3 log(float3 v) { return float3(std::log(v.x), std::log(v.y), std::log(v.z)); }
inline float4 log(float4 v) { return float4(std::log(v.x), std::log(v.y), std::log(v.z), std::log(v.w)); }
inline bool anyGreater(float3 v, float t) { return v.x > t || v.y > t || v.z > t; }
inline bool anyGreater(float4 v, float t) { return v.x > t || v.y > t || v.z > t || v.w > t; }
inline bool anyLess(float3 v, float t) { return v.x < t || v.y < t || v.z < t; }
inline float maxComponent(float3 v) { return max(v.x, max(v.y, v.z)); }
inline float maxComponent4(float4 v) { return max(max(v.x, v.y), max(v.z, v.w)); }
// GLSL/HLSL refract
inline float3 refract(float3 I, float3 N, float eta) {
    float NdotI = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - NdotI * NdotI);
    if (k < 0.0f) return float3(0.0f);
    return eta * I - (eta * NdotI + std::sqrt(k)) * N;
}

inline uint asuint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline float asfloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }

// IEEE binary16 conversions (round-to-nearest-even), used for the RGBA16F feature images.
inline uint16_t f32_to_f16(float value) {
    uint32_t bits = asuint(value);
    uint32_t sign = (bits >> 16) & 0x8000u;
    uint32_t mant = bits & 0x007fffffu;
    int32_t exp = (int32_t)((bits >> 23) & 0xffu) - 127;
    if (exp == 128) return (uint16_t)(sign | 0x7c00u | (mant ? (0x200u | (mant >> 13)) : 0u));
    if (exp > 15) return (uint16_t)(sign | 0x7c00u);
    if (exp >= -14) {
        uint32_t half = sign | ((uint32_t)(exp + 15) << 10) | (mant >> 13);
        uint32_t round = (mant >> 12) & 1u, sticky = mant & 0xfffu;
        if (round && (sticky || (half & 1u))) half++;
        return (uint16_t)half;
    }
    if (exp < -25) return (uint16_t)sign;
    mant |= 0x00800000u;
    // subnormal half: value = mant * 2^(exp-23); half ulp = 2^-24 -> q = mant >> (-(exp) - 1)
    uint32_t s = (uint32_t)(-exp - 1);
    uint32_t q = mant >> s;
    uint32_t r = (mant >> (s - 1)) & 1u;
    uint32_t st = mant & ((1u << (s - 1)) - 1u);
    uint32_t h = sign | q;
    if (r && (st || (h & 1u))) h++;
    return (uint16_t)h;
}
inline float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t mant = h & 0x3ffu;
    if (exp == 0) {
        if (mant == 0) return asfloat(sign);
        float f = (float)mant * 5.9604644775390625e-08f; // 2^-24
        return sign ? -f : f;
    }
    if (exp == 31) return asfloat(sign | 0x7f800000u | (mant << 13));
    return asfloat(sign | ((exp + 112u) << 23) | (mant << 13));
}

} // namespace orc
