// intersect.cuh — the ray/triangle and instance-transform semantics of the traversal kernels.
//
// The reference delegates this to the Vulkan driver (TraceRay, src/shaders/rt/queries/scene_query.slang:30-39;
// acceleration structures built by src/core/render/accel/{blas,tlas}.c), so nothing pins the arithmetic upstream.
// The specification implemented here (and, independently, by the CPU oracle) is:
//   * object-space ray = inverse(3x4 world) applied to origin/direction, direction not renormalised (t preserved);
//   * Woop/Benthin/Wald 2013 watertight test, fp32, every product and sum rounded once (no FMA: contraction would
//     break the shared-edge symmetry the test relies on), fp64 fallback when an edge function is exactly zero;
//   * no culling; barycentrics (u,v) weight vertices 1 and 2; hit accepted iff tMin < t and
//     (t < tBest or (t == tBest and (instance, primitive) lexicographically smaller)) -> traversal-order independent.
#pragma once
#include "vmath.cuh"

namespace vk {

// Shear constants of a ray (Woop et al. 2013). The permutation (kx, ky, kz) is stored as PRMT selectors: picking component k of a
// vector is two byte-permutes, no compare / select chains or branches inside the per-triangle code (profiles/r01_notes.md).
struct RayShear {
    uint32_t x1, x2, y1, y2, z1, z2;
    float Sx, Sy, Sz;
};

__device__ __forceinline__ float comp(const float3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ float pick(const float3& v, uint32_t s1, uint32_t s2) {
    return __uint_as_float(__byte_perm(__byte_perm(__float_as_uint(v.x), __float_as_uint(v.y), s1), __float_as_uint(v.z), s2));
}

__device__ __forceinline__ bool makeRayShear(const float3& d, RayShear& s) {
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0;
    float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; m = az; }
    int kx = kz == 2 ? 0 : kz + 1;
    int ky = kx == 2 ? 0 : kx + 1;
    float dz = comp(d, kz);
    if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }
    s.x1 = kx == 1 ? 0x7654u : 0x3210u; s.x2 = kx == 2 ? 0x7654u : 0x3210u;
    s.y1 = ky == 1 ? 0x7654u : 0x3210u; s.y2 = ky == 2 ? 0x7654u : 0x3210u;
    s.z1 = kz == 1 ? 0x7654u : 0x3210u; s.z2 = kz == 2 ? 0x7654u : 0x3210u;
    s.Sx = __fdiv_rn(pick(d, s.x1, s.x2), dz);
    s.Sy = __fdiv_rn(pick(d, s.y1, s.y2), dz);
    s.Sz = __fdiv_rn(1.0f, dz);
    return m > 0.0f;
}

__device__ __forceinline__ bool watertightTriangle(const float3& org, const RayShear& s, const float3& v0, const float3& v1,
                                                    const float3& v2, float& t, float& u, float& v) {
    const float3 A = float3(__fsub_rn(v0.x, org.x), __fsub_rn(v0.y, org.y), __fsub_rn(v0.z, org.z));
    const float3 B = float3(__fsub_rn(v1.x, org.x), __fsub_rn(v1.y, org.y), __fsub_rn(v1.z, org.z));
    const float3 C = float3(__fsub_rn(v2.x, org.x), __fsub_rn(v2.y, org.y), __fsub_rn(v2.z, org.z));
    const float Akz = pick(A, s.z1, s.z2), Bkz = pick(B, s.z1, s.z2), Ckz = pick(C, s.z1, s.z2);
    const float Ax = __fsub_rn(pick(A, s.x1, s.x2), __fmul_rn(s.Sx, Akz)), Ay = __fsub_rn(pick(A, s.y1, s.y2), __fmul_rn(s.Sy, Akz));
    const float Bx = __fsub_rn(pick(B, s.x1, s.x2), __fmul_rn(s.Sx, Bkz)), By = __fsub_rn(pick(B, s.y1, s.y2), __fmul_rn(s.Sy, Bkz));
    const float Cx = __fsub_rn(pick(C, s.x1, s.x2), __fmul_rn(s.Sx, Ckz)), Cy = __fsub_rn(pick(C, s.y1, s.y2), __fmul_rn(s.Sy, Ckz));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;
    const float Az = __fmul_rn(s.Sz, Akz), Bz = __fmul_rn(s.Sz, Bkz), Cz = __fmul_rn(s.Sz, Ckz);
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
    const float invDet = __fdiv_rn(1.0f, det);
    t = __fmul_rn(T, invDet);
    u = __fmul_rn(V, invDet);
    v = __fmul_rn(W, invDet);
    return true;
}

// row-major 3x4 affine applied to a point / vector, terms summed left to right, one rounding per op
__device__ __forceinline__ float3 xformPoint(const float4& r0, const float4& r1, const float4& r2, const float3& p) {
    return float3(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, p.x), __fmul_rn(r0.y, p.y)), __fmul_rn(r0.z, p.z)), r0.w),
                  __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, p.x), __fmul_rn(r1.y, p.y)), __fmul_rn(r1.z, p.z)), r1.w),
                  __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, p.x), __fmul_rn(r2.y, p.y)), __fmul_rn(r2.z, p.z)), r2.w));
}
__device__ __forceinline__ float3 xformVector(const float4& r0, const float4& r1, const float4& r2, const float3& v) {
    return float3(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, v.x), __fmul_rn(r0.y, v.y)), __fmul_rn(r0.z, v.z)),
                  __fadd_rn(__fadd_rn(__fmul_rn(r1.x, v.x), __fmul_rn(r1.y, v.y)), __fmul_rn(r1.z, v.z)),
                  __fadd_rn(__fadd_rn(__fmul_rn(r2.x, v.x), __fmul_rn(r2.y, v.y)), __fmul_rn(r2.z, v.z)));
}

// Reciprocal direction for the slab tests only (never for a hit): components below 1e-20 are replaced by +-1e-20, a displacement of
// t * 1e-20 along that axis (far below fp32 resolution of any coordinate), so that 2^(e+15) * idir stays finite (trace.cuh).
__device__ __forceinline__ float3 safeInvDir(const float3& d) {
    const float eps = 1e-20f;
    float x = fabsf(d.x) < eps ? (d.x < 0.0f ? -eps : eps) : d.x;
    float y = fabsf(d.y) < eps ? (d.y < 0.0f ? -eps : eps) : d.y;
    float z = fabsf(d.z) < eps ? (d.z < 0.0f ? -eps : eps) : d.z;
    return float3(1.0f / x, 1.0f / y, 1.0f / z);
}

} // namespace vk
