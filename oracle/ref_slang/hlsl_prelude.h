// hlsl_prelude.h -- TEST INFRASTRUCTURE (oracle/): the part of the Slang/HLSL core library that the reference's shaders use,
// written here as plain scalar C++ so that the reference's own .slang sources (transliterated token by token by slang2cpp.py into
// oracle/_ref/, never committed) compile with g++ and can be called from the parity tests.
//
// Nothing in this file is reference code: it is the language runtime a Slang compiler would supply. Semantics follow the HLSL
// intrinsics (saturate = clamp to [0,1] with NaN -> 0, frac = x - floor(x), lerp = a + (b - a) t, rsqrt = 1 / sqrt, mul(M, v) with
// `-matrix-layout-column-major` = cglm's column-major mat4 times vector). One rounding per operation: compile with
// -ffp-contract=off -fno-fast-math (oracle/Makefile does).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace refslang {

typedef uint32_t uint;

struct bool2 { bool x, y; };
struct bool3 { bool x, y, z; };
struct bool4 { bool x, y, z, w; };
inline bool any(bool v) { return v; }
inline bool any(bool2 v) { return v.x || v.y; }
inline bool any(bool3 v) { return v.x || v.y || v.z; }
inline bool any(bool4 v) { return v.x || v.y || v.z || v.w; }
inline bool all(bool v) { return v; }
inline bool all(bool2 v) { return v.x && v.y; }
inline bool all(bool3 v) { return v.x && v.y && v.z; }
inline bool all(bool4 v) { return v.x && v.y && v.z && v.w; }

struct int2;
struct uint2;
struct float2 {
    float x, y;
    float2() : x(0.0f), y(0.0f) {}
    float2(float s) : x(s), y(s) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
    explicit float2(const int2& v);
    explicit float2(const uint2& v);
    float& operator[](uint i) { return (&x)[i]; }
    float operator[](uint i) const { return (&x)[i]; }
    float2 xy() const { return *this; }
    float2 yx() const { return float2(y, x); }
};
struct float3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; };   // anonymous structs: GNU extension, fine for g++
    float3() : x(0.0f), y(0.0f), z(0.0f) {}
    float3(float s) : x(s), y(s), z(s) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(float2 v, float z_) : x(v.x), y(v.y), z(z_) {}
    float3(const float v[3]) : x(v[0]), y(v[1]), z(v[2]) {}
    float& operator[](uint i) { return (&x)[i]; }
    float operator[](uint i) const { return (&x)[i]; }
    float2 xy() const { return float2(x, y); }
    float3 xyz() const { return *this; }
    float3 rgb() const { return *this; }
};
struct alignas(16) float4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; };
    float4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(float3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float4(float2 a, float2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    float4(float2 a, float z_, float w_) : x(a.x), y(a.y), z(z_), w(w_) {}
    float& operator[](uint i) { return (&x)[i]; }
    float operator[](uint i) const { return (&x)[i]; }
    float2 xy() const { return float2(x, y); }
    float2 zw() const { return float2(z, w); }
    float3 xyz() const { return float3(x, y, z); }
    float3 rgb() const { return float3(x, y, z); }
    float4 xyzw() const { return *this; }
    float4 rgba() const { return *this; }
};
struct int2 {
    int x, y;
    int2() : x(0), y(0) {}
    int2(int s) : x(s), y(s) {}
    int2(int x_, int y_) : x(x_), y(y_) {}
    explicit int2(const float2& v) : x(int(v.x)), y(int(v.y)) {}
    explicit int2(const uint2& v);
    int2 xy() const { return *this; }
};
inline float2::float2(const int2& v) : x(float(v.x)), y(float(v.y)) {}
struct uint2 {
    uint x, y;
    uint2() : x(0), y(0) {}
    uint2(uint x_, uint y_) : x(x_), y(y_) {}
    explicit uint2(const int2& v) : x(uint(v.x)), y(uint(v.y)) {}
    uint2 xy() const { return *this; }
};
inline float2::float2(const uint2& v) : x(float(v.x)), y(float(v.y)) {}
inline int2::int2(const uint2& v) : x(int(v.x)), y(int(v.y)) {}
struct uint3 {
    uint x, y, z;
    uint3() : x(0), y(0), z(0) {}
    uint3(uint x_, uint y_, uint z_) : x(x_), y(y_), z(z_) {}
    uint2 xy() const { return uint2(x, y); }
};
struct alignas(16) uint4 {
    uint x, y, z, w;
    uint4() : x(0), y(0), z(0), w(0) {}
    uint4(uint x_, uint y_, uint z_, uint w_) : x(x_), y(y_), z(z_), w(w_) {}
    uint& operator[](uint i) { return (&x)[i]; }
    uint operator[](uint i) const { return (&x)[i]; }
    uint2 xy() const { return uint2(x, y); }
    uint2 zw() const { return uint2(z, w); }
};
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator-(int2 a, int2 b) { return int2(a.x - b.x, a.y - b.y); }
inline bool2 operator==(int2 a, int2 b) { return {a.x == b.x, a.y == b.y}; }
inline bool2 operator>=(int2 a, int2 b) { return {a.x >= b.x, a.y >= b.y}; }
inline bool2 operator<(int2 a, int2 b) { return {a.x < b.x, a.y < b.y}; }

// ---- scalar intrinsics (float versions only: no silent promotion to double) -----------------------------------------------------
inline float abs(float v) { return ::fabsf(v); }
inline int abs(int v) { return v < 0 ? -v : v; }
inline float sqrt(float v) { return ::sqrtf(v); }
inline float rsqrt(float v) { return 1.0f / ::sqrtf(v); }
inline float exp(float v) { return ::expf(v); }
inline float exp2(float v) { return ::exp2f(v); }
inline float log(float v) { return ::logf(v); }
inline float sin(float v) { return ::sinf(v); }
inline float cos(float v) { return ::cosf(v); }
inline float acos(float v) { return ::acosf(v); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float pow(float a, float b) { return ::powf(a, b); }
inline float floor(float v) { return ::floorf(v); }
inline float round(float v) { return ::nearbyintf(v); }  // SPIR-V Round: implementation-defined ties, RoundEven in practice
inline float frac(float v) { return v - ::floorf(v); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float clamp(float v, float lo, float hi) { return ::fminf(::fmaxf(v, lo), hi); }
inline int clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }
inline uint clamp(uint v, uint lo, uint hi) { return min(max(v, lo), hi); }
inline float saturate(float v) { return ::fminf(::fmaxf(v, 0.0f), 1.0f); }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline uint asuint(float v) { uint u; std::memcpy(&u, &v, 4); return u; }
inline float asfloat(uint v) { float f; std::memcpy(&f, &v, 4); return f; }
inline bool isnan(float v) { return v != v; }
inline bool isinf(float v) { return std::isinf(v); }
inline bool isfinite(float v) { return std::isfinite(v); }
inline uint reversebits(uint v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
inline float f16tof32(uint h) {
    const uint s = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
    if (e == 0u) return (s ? -1.0f : 1.0f) * ::ldexpf(float(m), -24);
    if (e == 31u) return m ? NAN : (s ? -INFINITY : INFINITY);
    return (s ? -1.0f : 1.0f) * ::ldexpf(float(m | 1024u), int(e) - 25);
}

// ---- component-wise vector operators ------------------------------------------------------------------------------------------
#define REFSLANG_VEC_OPS(V, B, ...)                                                                                     \
    REFSLANG_BIN(V, +) REFSLANG_BIN(V, -) REFSLANG_BIN(V, *) REFSLANG_BIN(V, /)                                         \
    REFSLANG_CMP(V, B, >) REFSLANG_CMP(V, B, <) REFSLANG_CMP(V, B, >=) REFSLANG_CMP(V, B, <=) REFSLANG_CMP(V, B, ==)    \
    REFSLANG_CMP(V, B, !=)                                                                                              \
    inline V operator-(V a) { return REFSLANG_NEG(V, a); }

#define REFSLANG_BIN(V, OP)                                                  \
    inline V operator OP(V a, V b) { return REFSLANG_APPLY2(V, a, b, OP); }  \
    inline V operator OP(V a, float b) { return a OP V(b); }                 \
    inline V operator OP(float a, V b) { return V(a) OP b; }                 \
    inline V& operator OP##=(V& a, V b) { a = a OP b; return a; }            \
    inline V& operator OP##=(V& a, float b) { a = a OP V(b); return a; }
#define REFSLANG_CMP(V, B, OP)                                               \
    inline B operator OP(V a, V b) { return REFSLANG_APPLYB(B, a, b, OP); }  \
    inline B operator OP(V a, float b) { return a OP V(b); }                 \
    inline B operator OP(float a, V b) { return V(a) OP b; }

#define REFSLANG_APPLY2(V, a, b, OP) V(a.x OP b.x, a.y OP b.y)
#define REFSLANG_APPLYB(B, a, b, OP) B{a.x OP b.x, a.y OP b.y}
#define REFSLANG_NEG(V, a) V(-a.x, -a.y)
REFSLANG_VEC_OPS(float2, bool2)
#undef REFSLANG_APPLY2
#undef REFSLANG_APPLYB
#undef REFSLANG_NEG
#define REFSLANG_APPLY2(V, a, b, OP) V(a.x OP b.x, a.y OP b.y, a.z OP b.z)
#define REFSLANG_APPLYB(B, a, b, OP) B{a.x OP b.x, a.y OP b.y, a.z OP b.z}
#define REFSLANG_NEG(V, a) V(-a.x, -a.y, -a.z)
REFSLANG_VEC_OPS(float3, bool3)
#undef REFSLANG_APPLY2
#undef REFSLANG_APPLYB
#undef REFSLANG_NEG
#define REFSLANG_APPLY2(V, a, b, OP) V(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w)
#define REFSLANG_APPLYB(B, a, b, OP) B{a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w}
#define REFSLANG_NEG(V, a) V(-a.x, -a.y, -a.z, -a.w)
REFSLANG_VEC_OPS(float4, bool4)
#undef REFSLANG_APPLY2
#undef REFSLANG_APPLYB
#undef REFSLANG_NEG

#define REFSLANG_MAP1(F)                                                              \
    inline float2 F(float2 v) { return float2(F(v.x), F(v.y)); }                      \
    inline float3 F(float3 v) { return float3(F(v.x), F(v.y), F(v.z)); }              \
    inline float4 F(float4 v) { return float4(F(v.x), F(v.y), F(v.z), F(v.w)); }
REFSLANG_MAP1(abs) REFSLANG_MAP1(sqrt) REFSLANG_MAP1(rsqrt) REFSLANG_MAP1(exp) REFSLANG_MAP1(exp2) REFSLANG_MAP1(log) REFSLANG_MAP1(sin)
REFSLANG_MAP1(cos) REFSLANG_MAP1(floor) REFSLANG_MAP1(frac) REFSLANG_MAP1(saturate) REFSLANG_MAP1(round)
#define REFSLANG_MAP2(F)                                                                                       \
    inline float2 F(float2 a, float2 b) { return float2(F(a.x, b.x), F(a.y, b.y)); }                           \
    inline float3 F(float3 a, float3 b) { return float3(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z)); }              \
    inline float4 F(float4 a, float4 b) { return float4(F(a.x, b.x), F(a.y, b.y), F(a.z, b.z), F(a.w, b.w)); } \
    inline float2 F(float2 a, float b) { return F(a, float2(b)); }                                             \
    inline float3 F(float3 a, float b) { return F(a, float3(b)); }                                             \
    inline float4 F(float4 a, float b) { return F(a, float4(b)); }                                             \
    inline float2 F(float a, float2 b) { return F(float2(a), b); }                                             \
    inline float3 F(float a, float3 b) { return F(float3(a), b); }                                             \
    inline float4 F(float a, float4 b) { return F(float4(a), b); }
REFSLANG_MAP2(min) REFSLANG_MAP2(max) REFSLANG_MAP2(pow)
#define REFSLANG_MAP3(F)                                                                                                        \
    inline float2 F(float2 a, float2 b, float2 c) { return float2(F(a.x, b.x, c.x), F(a.y, b.y, c.y)); }                        \
    inline float3 F(float3 a, float3 b, float3 c) { return float3(F(a.x, b.x, c.x), F(a.y, b.y, c.y), F(a.z, b.z, c.z)); }      \
    inline float4 F(float4 a, float4 b, float4 c) { return float4(F(a.x, b.x, c.x), F(a.y, b.y, c.y), F(a.z, b.z, c.z), F(a.w, b.w, c.w)); } \
    inline float2 F(float2 a, float b, float c) { return F(a, float2(b), float2(c)); }                                          \
    inline float3 F(float3 a, float b, float c) { return F(a, float3(b), float3(c)); }                                          \
    inline float4 F(float4 a, float b, float c) { return F(a, float4(b), float4(c)); }
REFSLANG_MAP3(clamp)
inline float2 lerp(float2 a, float2 b, float t) { return a + (b - a) * t; }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float4 lerp(float4 a, float4 b, float t) { return a + (b - a) * t; }
inline float2 lerp(float2 a, float2 b, float2 t) { return a + (b - a) * t; }
inline float3 lerp(float3 a, float3 b, float3 t) { return a + (b - a) * t; }
inline float4 lerp(float4 a, float4 b, float4 t) { return a + (b - a) * t; }

inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(float2 v) { return sqrt(dot(v, v)); }
inline float length(float3 v) { return sqrt(dot(v, v)); }
inline float3 normalize(float3 v) { return v * rsqrt(dot(v, v)); }
inline float2 normalize(float2 v) { return v * rsqrt(dot(v, v)); }
inline float3 refract(float3 i, float3 n, float eta) {
    const float d = dot(n, i), k = 1.0f - eta * eta * (1.0f - d * d);
    return k < 0.0f ? float3(0.0f) : eta * i - (eta * d + sqrt(k)) * n;
}

inline float2 unpackHalf2x16ToFloat(uint v) { return float2(f16tof32(v & 0xffffu), f16tof32(v >> 16)); }

// broadcast swizzles (`s.xxx` on a scalar or a vector): slang2cpp.py rewrites them into these calls
inline float3 swz_xxx(float v) { return float3(v); }
inline float3 swz_xxx(float3 v) { return float3(v.x); }
inline float3 swz_xxx(float4 v) { return float3(v.x); }
inline float4 swz_xxxx(float v) { return float4(v); }
inline float4 swz_xxxx(float3 v) { return float4(v.x); }
inline float4 swz_xxxx(float4 v) { return float4(v.x); }
inline float3 swz_ggg(float3 v) { return float3(v.y); }
inline float3 swz_ggg(float4 v) { return float3(v.y); }
inline float3 swz_bbb(float3 v) { return float3(v.z); }
inline float3 swz_bbb(float4 v) { return float3(v.z); }

// ---- matrices: memory layout as uploaded by the host (cglm, column-major: m[c][r]); mul(M, v)[r] = sum_c m[c][r] v[c] ------------
struct alignas(16) float4x4 {
    float m[4][4];
};
inline float4 mul(const float4x4& M, float4 v) {
    float4 r;
    for (int i = 0; i < 4; i++) r[uint(i)] = M.m[0][i] * v.x + M.m[1][i] * v.y + M.m[2][i] * v.z + M.m[3][i] * v.w;
    return r;
}
struct float3x3 {
    float3 r[3];   // rows, float3x3(a, b, c) takes rows
    float3x3() {}
    float3x3(float3 a, float3 b, float3 c) { r[0] = a; r[1] = b; r[2] = c; }
    float3x3(float a, float b, float c, float d, float e, float f, float g, float h, float i) { r[0] = float3(a, b, c); r[1] = float3(d, e, f); r[2] = float3(g, h, i); }
};
inline float3 mul(const float3x3& M, float3 v) { return float3(dot(M.r[0], v), dot(M.r[1], v), dot(M.r[2], v)); }
inline float3 mul(float3 v, const float3x3& M) { return M.r[0] * v.x + M.r[1] * v.y + M.r[2] * v.z; }

// ---- resources: what the descriptor set binds (scene/resources.slang) becomes plain pointers set by the test harness --------------
template <class T>
struct StructuredBuffer {
    const T* data = nullptr;
    const T& operator[](uint i) const { return data[i]; }
    const T& operator[](int i) const { return data[i]; }
};
template <class T>
struct RWStructuredBuffer {
    T* data = nullptr;
    T& operator[](uint i) const { return data[i]; }
};
template <class T>
using ConstantBuffer = T;

struct RayDesc {
    float3 Origin;
    float TMin;
    float3 Direction;
    float TMax;
};

}  // namespace refslang
