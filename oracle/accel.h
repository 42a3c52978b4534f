// ORACLE — TEST INFRASTRUCTURE ONLY. CPU acceleration structure + the pinned ray/triangle semantics.
//
// The reference has NO intersection code: BVH build, traversal order and the ray/triangle test are the Vulkan
// driver's (rt/queries/scene_query.slang:30-39 -> TraceRay; core/render/accel/{blas,tlas}.c). Parity is therefore
// UNPINNED at this seam. What is pinned here, and mirrored operation-for-operation by the CUDA
// traversal kernels (vkrt_b200/csrc/intersect.cuh), is the *specification* both sides implement:
//   * rays are transformed into instance object space with the inverse of the 3x4 world transform
//     (direction NOT renormalised, so t is preserved) — same as Vulkan's instance semantics;
//   * watertight ray/triangle test of Woop, Benthin, Wald 2013 in fp32 with one rounding per op (no FMA),
//     fp64 fallback when an edge function is exactly 0; no back-face culling (RAY_FLAG_NONE);
//   * barycentrics (u, v) weight vertices 1 and 2 (Vulkan convention, interpolation.slang:27-34);
//   * a hit is accepted iff tMin < t and (t < tBest or (t == tBest and (instance, primitive) < best ids)),
//     which makes the closest hit independent of traversal order; initial tBest = tMax (exclusive).
// Traversal here is a plain two-level BVH2 (binned SAH); oracle_set_brute_force() switches to an exhaustive loop
// used by the tests to validate the BVH itself.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "vecmath.h"

namespace orc {

struct HitRecord {
    uint instance = 0xFFFFFFFFu;
    uint primitive = 0xFFFFFFFFu;
    float t = 0.0f;
    float u = 0.0f, v = 0.0f;
    bool hit() const { return instance != 0xFFFFFFFFu; }
};

struct RayShear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
    bool valid;
};

inline RayShear makeRayShear(float3 d) {
    RayShear s;
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int kz = 0;
    float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; m = az; }
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    if (d[kz] < 0.0f) std::swap(kx, ky);
    s.kx = kx; s.ky = ky; s.kz = kz;
    s.valid = m > 0.0f;
    float dz = d[kz];
    s.Sx = d[kx] / dz;
    s.Sy = d[ky] / dz;
    s.Sz = 1.0f / dz;
    return s;
}

// Returns true and fills t,u,v when the ray's supporting line crosses the triangle (no t-range test).
inline bool watertightTriangle(float3 org, const RayShear& s, float3 v0, float3 v1, float3 v2, float& t, float& u, float& v) {
    const float3 A = v0 - org, B = v1 - org, C = v2 - org;
    const float Ax = A[s.kx] - s.Sx * A[s.kz], Ay = A[s.ky] - s.Sy * A[s.kz];
    const float Bx = B[s.kx] - s.Sx * B[s.kz], By = B[s.ky] - s.Sy * B[s.kz];
    const float Cx = C[s.kx] - s.Sx * C[s.kz], Cy = C[s.ky] - s.Sy * C[s.kz];
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = s.Sz * A[s.kz], Bz = s.Sz * B[s.kz], Cz = s.Sz * C[s.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float invDet = 1.0f / det;
    t = T * invDet;
    u = V * invDet;
    v = W * invDet;
    return true;
}

inline bool hitCloser(float t, uint inst, uint prim, const HitRecord& best, float tBest) {
    if (t < tBest) return true;
    if (t == tBest && best.hit()) return inst < best.instance || (inst == best.instance && prim < best.primitive);
    return false;
}

struct Aabb {
    float3 lo = float3(1e30f), hi = float3(-1e30f);
    void grow(float3 p) { lo = min(lo, p); hi = max(hi, p); }
    void grow(const Aabb& b) { lo = min(lo, b.lo); hi = max(hi, b.hi); }
    float halfArea() const {
        float3 e = hi - lo;
        return e.x * e.y + e.y * e.z + e.z * e.x;
    }
    float3 center() const { return (lo + hi) * 0.5f; }
};

struct Bvh2Node {
    Aabb box;
    uint left = 0;   // inner: left child index (right = left + 1); leaf: first primitive slot
    uint count = 0;  // 0 = inner, else number of primitives
};

// Generic binned-SAH BVH2 over boxes. prims = permutation of [0, n).
struct Bvh2 {
    std::vector<Bvh2Node> nodes;
    std::vector<uint> prims;

    void build(const std::vector<Aabb>& boxes, uint maxLeaf) {
        const uint n = (uint)boxes.size();
        prims.resize(n);
        for (uint i = 0; i < n; i++) prims[i] = i;
        nodes.clear();
        nodes.reserve(n ? 2 * n : 1);
        nodes.push_back(Bvh2Node());
        if (n == 0) return;
        std::vector<float3> centers(n);
        for (uint i = 0; i < n; i++) centers[i] = boxes[i].center();
        struct Work { uint node, begin, end; };
        std::vector<Work> stack;
        stack.push_back({0, 0, n});
        while (!stack.empty()) {
            Work w = stack.back();
            stack.pop_back();
            Aabb box, cbox;
            for (uint i = w.begin; i < w.end; i++) {
                box.grow(boxes[prims[i]]);
                cbox.grow(centers[prims[i]]);
            }
            nodes[w.node].box = box;
            uint cnt = w.end - w.begin;
            if (cnt <= maxLeaf) {
                nodes[w.node].left = w.begin;
                nodes[w.node].count = cnt;
                continue;
            }
            float3 ce = cbox.hi - cbox.lo;
            int axis = 0;
            if (ce.y > ce.x) axis = 1;
            if (ce.z > ce[axis]) axis = 2;
            uint mid = w.begin;
            if (ce[axis] > 0.0f) {
                const int NB = 16;
                Aabb bb[NB];
                uint bc[NB] = {0};
                float k = NB * (1.0f - 1e-6f) / ce[axis];
                for (uint i = w.begin; i < w.end; i++) {
                    int b = (int)((centers[prims[i]][axis] - cbox.lo[axis]) * k);
                    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    bb[b].grow(boxes[prims[i]]);
                    bc[b]++;
                }
                float rightArea[NB];
                Aabb acc;
                for (int b = NB - 1; b > 0; b--) {
                    acc.grow(bb[b]);
                    rightArea[b] = acc.halfArea();
                }
                uint rc[NB];
                uint r = 0;
                for (int b = NB - 1; b > 0; b--) { r += bc[b]; rc[b] = r; }
                Aabb lacc;
                uint lc = 0;
                float best = 1e30f;
                int bestSplit = -1;
                for (int b = 0; b < NB - 1; b++) {
                    lacc.grow(bb[b]);
                    lc += bc[b];
                    if (lc == 0 || rc[b + 1] == 0) continue;
                    float cost = lacc.halfArea() * lc + rightArea[b + 1] * rc[b + 1];
                    if (cost < best) { best = cost; bestSplit = b; }
                }
                if (bestSplit >= 0) {
                    auto it = std::partition(prims.begin() + w.begin, prims.begin() + w.end, [&](uint p) {
                        int b = (int)((centers[p][axis] - cbox.lo[axis]) * k);
                        b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                        return b <= bestSplit;
                    });
                    mid = (uint)(it - prims.begin());
                }
            }
            if (mid == w.begin || mid == w.end) mid = w.begin + cnt / 2; // degenerate: split in the middle
            uint left = (uint)nodes.size();
            nodes.push_back(Bvh2Node());
            nodes.push_back(Bvh2Node());
            nodes[w.node].left = left;
            nodes[w.node].count = 0;
            stack.push_back({left, w.begin, mid});
            stack.push_back({left + 1, mid, w.end});
        }
    }
};

// Conservative slab test: boxes are padded by a few ulps so that the BVH can never cull a triangle the
// watertight test would accept (validated against brute force in tests/test_oracle_accel.py).
inline bool slabTest(const Aabb& b, float3 o, float3 invD, float tMin, float tMax) {
    float tx0 = (b.lo.x - o.x) * invD.x, tx1 = (b.hi.x - o.x) * invD.x;
    float ty0 = (b.lo.y - o.y) * invD.y, ty1 = (b.hi.y - o.y) * invD.y;
    float tz0 = (b.lo.z - o.z) * invD.z, tz1 = (b.hi.z - o.z) * invD.z;
    float tn = std::max(std::max(std::min(tx0, tx1), std::min(ty0, ty1)), std::max(std::min(tz0, tz1), tMin));
    float tf = std::min(std::min(std::max(tx0, tx1), std::max(ty0, ty1)), std::min(std::max(tz0, tz1), tMax));
    // NaN (0 * inf) compares false on both min/max the way std::min/max are written -> treated as "inside".
    return tn * 0.999999f <= tf * 1.000001f;
}

inline float3 safeInvDir(float3 d) {
    auto inv = [](float v) {
        const float eps = 1e-30f;
        if (std::fabs(v) < eps) v = v < 0.0f ? -eps : eps;
        return 1.0f / v;
    };
    return float3(inv(d.x), inv(d.y), inv(d.z));
}

} // namespace orc
