// trace.cuh — persistent-thread traversal of the two-level compressed 8-wide BVH.
//
// This kernel is what the reference obtains from the driver with TraceRay (src/shaders/rt/queries/scene_query.slang:10-42
// closest hit, RAY_FLAG_NONE, mask 0xff; src/shaders/rt/queries/shadow_query.slang:8-28 first-hit-terminate) plus the tiny hit
// shaders (entry/path/closest_hit.slang, miss.slang, any_hit.slang; entry/shadow/*.slang).
//
// Design (B200, no RT cores):
//   * one launch processes one bounce's extension-ray queue AND the previous shade's shadow-ray queue as a single work
//     list; rays are read/written as 16-byte SoA vectors (coalesced: lane i <-> ray base+i);
//   * persistent warps (grid = k * 148 SMs) pull rays with a warp-aggregated atomic; a warp refills its idle lanes as soon
//     as fewer than AccelView::refillLanes lanes (20 or 24, chosen per scene) are still traversing (Aila & Laine 2009 dynamic fetch), so a long ray does
//     not hold 31 idle lanes hostage;
//   * traversal stack: uint2 entries (node group / primitive group, Ylitie et al. 2017) in per-thread local memory (L1);
//   * 80-byte nodes are fetched as five 128-bit loads, triangles as three; everything read-only goes through LDG.
#pragma once
#include "intersect.cuh"
#include "scene_view.cuh"

namespace vk {

constexpr int TRACE_STACK = 40;        // stack entries of the default kernel: one per BVH8 level + 2 per instance entry (AccelView::stackNeed)
constexpr int TRACE_STACK_DEEP = 192;  // second instantiation for degenerate (chain-like) trees; deeper trees are refused at build time
constexpr int TRACE_BLOCK = 128;
#ifndef TRACE_REFILL
#define TRACE_REFILL 20        // two-level kernel; measured 16 / 20 / 24 / 28 lanes: cornell 33.1 / 33.2 / 34.8 / 39.0 ms
#endif
#ifndef TRACE_REFILL_FLAT
#define TRACE_REFILL_FLAT 24   // single-level BVH over TRACE_REFILL_BIG_SCENE triangles or more. Measured at 20 / 24 lanes (profiles/r02k_sweeps.txt): soup of 1 M / 3 M /
                               // 10 M triangles 30.2 / 29.6, 18.5 / 18.2, 20.2 / 19.9 ms; 1000 spheres (0.6 M) 34.3 / 34.0 ms; but the flat cornell box (70 k
                               // triangles, 6 node visits per ray: a refill is a large share of such a short walk) 28.3 / 29.4 ms -> 20 for small scenes
#endif
constexpr unsigned long long TRACE_REFILL_BIG_SCENE = 262144ull;
#ifndef TRACE_PRIM_VOTE
#define TRACE_PRIM_VOTE 8   // lanes with a pending primitive wait until 8 of them can run the primitive section together (0 = every trip)
#endif
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 7   // 72 registers, 28 warps per SM. Measured again once the parameter block stayed in constant memory (profiles/r02z_trace_occupancy.txt):
                             // 5 / 6 / 7 / 8 blocks per SM: cornell 27.3 / 26.2 / 25.7 / 26.7 ms, soup 18.6 / 18.0 / 17.8 / 18.7, 1000 spheres 32.5 / 30.9 / 30.2 / 31.6
#endif
// Cache hints for the wavefront queues (path state, rays, hit records, shadow queue): every entry is written once by one kernel and read
// once or twice by the next, and a launch streams 1 - 5 GB of them through the 126 MB L2 beside the ~30 MB the kernels re-read (BVH,
// geometry, materials, rgb2spec cells). ld.global.cs / st.global.cs (SASS: LDG.E.EF / STG.E.EF) mark queue lines evict-first; results
// cannot change. Measured on the same box (profiles/r03i_hints_ab.txt): cornell hero frame 48.0 -> 47.7 ms (trace 25.42 -> 25.25, shade
// 22.57 -> 22.45), RGB 39.05 -> 38.98 ms. -DNO_TRACE_STREAM_HINTS / -DNO_SHADE_STREAM_HINTS build the plain accesses for an A/B.
#ifndef NO_TRACE_STREAM_HINTS
#define TQ_LD(p) __ldcs(p)
#define TQ_ST(p, v) __stcs((p), (v))
#else
#define TQ_LD(p) __ldg(p)
#define TQ_ST(p, v) (*(p) = (v))
#endif
#ifndef NO_SHADE_STREAM_HINTS
#define SQ_LD(p) __ldcs(p)
#define SQ_ST(p, v) __stcs((p), (v))
#else
#define SQ_LD(p) (*(p))
#define SQ_ST(p, v) (*(p) = (v))
#endif
constexpr uint32_t SHADOW_KIND_SCALAR = 1u << 31;  // in ShadowTarget.statePos: contribution goes to the scalar lane (hero fallback)

struct TraceParams {
    SceneView scene;
    // extension rays (closest hit)
    const ::float4* rayO;        // origin.xyz, tMin
    const ::float4* rayD;        // direction.xyz, tMax
    const uint32_t* raySeed;     // rng at ray start; read only when an alpha-tested instance is met
    uint32_t slotStride;         // words between consecutive slots of raySeed / hitB / pathFlags (1 = plain arrays, 4 = PathState::meta)
    ::uint4* hitA;               // instance, primitive, t bits, u bits
    float* hitB;                 // v
    const uint32_t* extCount;    // device-resident queue length
    // shadow rays (any hit); results are applied in place
    const ::float4* shO;
    const ::float4* shD;
    const ::float4* shContribution;  // rgb / 4 wavelengths, already multiplied by throughput
    const ::uint2* shTarget;         // x = sample-record index, y = new path-state position (0x7fffffff = path ended) | SHADOW_KIND_SCALAR
    const uint32_t* shSeed;
    const uint32_t* shCount;
    ::float4* radiance;          // per sample-record accumulators
    float* radianceScalar;       // hero mode: single-wavelength lane after dispersive collapse
    uint32_t* pathFlags;         // new path-state flags (bit 0 = prevVertexNeeAllowed)
    uint32_t* shadowResult;      // optional (trace_rays API): 0 visible, 1 occluded, 2 unsupported transmission
    // bookkeeping
    uint32_t* workCounter;       // zero-initialised per launch
    unsigned long long* stats;   // optional: [0] nodes visited, [1] triangles tested, [2] instances entered
};

// Byte j of v dropped into the mantissa of 1.0f: PRMT builds 0x3F80bb00 = 1 + b * 2^-15 (no conversion pipe, no bias subtraction:
// the "-1" is folded into the per-node plane constants below). `one` must live in a register so that the selector is the
// immediate operand of PRMT (with a literal constant ptxas keeps the four selectors in uniform registers and copies one into a
// vector register before every PRMT: 45 extra instructions per node, profiles/r01_notes.md).
__device__ __forceinline__ float byteToUnitFloat(uint32_t v, uint32_t one, uint32_t j) {
    return __uint_as_float(__byte_perm(one, v, 0x3240u + (j << 4)));
}

// Intersects the 8 children of one compressed node; returns the hit mask: bits 24..31 = internal children ordered by
// (slot ^ octinv) (nearest = highest bit), bits 0..23 = primitives of the leaf children. Branch-free: the meta byte of a child
// encodes {count bits << 5 | bit position} for both kinds (internal children: 1 << 5 | 24 + slot), so one shift places the bits;
// internal positions are re-ordered by XOR with octinv; empty slots (meta 0) contribute nothing.
//
// Slab arithmetic: plane = p + q * 2^e, t = (plane - o) * idir = f * A + B with f = 1 + q * 2^-15 (byteToUnitFloat),
// A = 2^(e+15) * idir, B = (p - o) * idir - A: ONE FMA per plane. B carries a rounding error of up to 2^-24 |A| = 2^-9 of a
// quantisation step of ITS axis, so the near planes use B - 2^-22 |A| and the far planes B + 2^-22 |A| (1/128 step; a common margin
// for the three axes would be a disaster: an axis the ray is almost parallel to has a huge |A|). The relative slack covers the
// rounding of t itself. Culling never changes a result: the closest hit is independent of the traversal order (intersect.cuh).
__device__ __forceinline__ uint32_t intersectNode8(const Bvh8Node* __restrict__ node, const float3& o, const float3& idir, uint32_t octinv,
                                                   uint32_t one, float tMin, float tMax, uint32_t& childBase, uint32_t& primBase, uint32_t& imask) {
    const ::float4* q = reinterpret_cast<const ::float4*>(node);
    const ::float4 n0 = __ldg(q), n1 = __ldg(q + 1), n2 = __ldg(q + 2), n3 = __ldg(q + 3), n4 = __ldg(q + 4);
    const uint32_t ebits = __float_as_uint(n0.w);
    imask = ebits >> 24;
    childBase = __float_as_uint(n1.x);
    primBase = __float_as_uint(n1.y);
    const float Ax = __uint_as_float(((ebits & 0xffu) + 15u) << 23) * idir.x;
    const float Ay = __uint_as_float((((ebits >> 8) & 0xffu) + 15u) << 23) * idir.y;
    const float Az = __uint_as_float((((ebits >> 16) & 0xffu) + 15u) << 23) * idir.z;
    const float Bx = fmaf(n0.x - o.x, idir.x, -Ax), By = fmaf(n0.y - o.y, idir.y, -Ay), Bz = fmaf(n0.z - o.z, idir.z, -Az);
    constexpr float M = 2.384185791015625e-07f;  // 2^-22
    const float Bnx = fmaf(fabsf(Ax), -M, Bx), Bny = fmaf(fabsf(Ay), -M, By), Bnz = fmaf(fabsf(Az), -M, Bz);
    const float Bfx = fmaf(fabsf(Ax), M, Bx), Bfy = fmaf(fabsf(Ay), M, By), Bfz = fmaf(fabsf(Az), M, Bz);
    const bool negx = idir.x < 0.0f, negy = idir.y < 0.0f, negz = idir.z < 0.0f;
    const uint32_t octinv4 = octinv * 0x01010101u;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
        const uint32_t lox4 = __float_as_uint(half ? n2.y : n2.x), loy4 = __float_as_uint(half ? n2.w : n2.z);
        const uint32_t loz4 = __float_as_uint(half ? n3.y : n3.x), hix4 = __float_as_uint(half ? n3.w : n3.z);
        const uint32_t hiy4 = __float_as_uint(half ? n4.y : n4.x), hiz4 = __float_as_uint(half ? n4.w : n4.z);
        // four children at a time (Ylitie et al. 2017, listing 1): internal children have bits 3 and 4 of their position set
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;  // 0x10 -> 0xff per byte (no carries between bytes)
        const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t nx4 = negx ? hix4 : lox4, fx4 = negx ? lox4 : hix4;
        const uint32_t ny4 = negy ? hiy4 : loy4, fy4 = negy ? loy4 : hiy4;
        const uint32_t nz4 = negz ? hiz4 : loz4, fz4 = negz ? loz4 : hiz4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tnx = fmaf(byteToUnitFloat(nx4, one, j), Ax, Bnx), tfx = fmaf(byteToUnitFloat(fx4, one, j), Ax, Bfx);
            const float tny = fmaf(byteToUnitFloat(ny4, one, j), Ay, Bny), tfy = fmaf(byteToUnitFloat(fy4, one, j), Ay, Bfy);
            const float tnz = fmaf(byteToUnitFloat(nz4, one, j), Az, Bnz), tfz = fmaf(byteToUnitFloat(fz4, one, j), Az, Bfz);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tMin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tMax));
            const bool hit = tn <= tf * 1.000002f;
            const uint32_t bits = ((childBits4 >> (j * 8)) & 0xffu) << ((bitIndex4 >> (j * 8)) & 0xffu);
            hitmask |= hit ? bits : 0u;
        }
    }
    return hitmask;
}

struct LaneRay {
    float3 o, d;        // world ray
    float tMin, tBest;
    uint32_t index;     // position in the combined work list
    uint32_t hitInst, hitPrim;
    float hitU, hitV;
    bool anyHit, sawTransmissive, active;
};

template <bool COUNT, bool FLAT, int STACK>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS) k_trace(const __grid_constant__ TraceParams P) {
    const int lane = threadIdx.x & 31;
    const uint32_t extCount = P.extCount ? *P.extCount : 0u;
    const uint32_t shCount = P.shCount ? *P.shCount : 0u;
    const uint32_t total = extCount + shCount;
    const AccelView& A = P.scene.accel;

    uint2 stack[STACK];
    int sp = 0;
    LaneRay R;
    R.active = false;
    // traversal registers
    float3 o, d, idir;
    uint32_t octinv = 0;
    bool inBlas = false;
    int blasBase = 0;
    uint32_t curInst = 0, curFlags = 0;
    RayShear shear;
    float3 objO(0.0f);  // flat variant: ray origin in the cached instance's object space
    uint2 G = make_uint2(0u, 0u), Gt = make_uint2(0u, 0u);  // pending node group / pending primitive group of this lane
    bool exhausted = false;
    uint32_t one;  // 1.0f as an opaque register value (byteToUnitFloat)
    asm volatile("mov.b32 %0, 0x3F800000;" : "=r"(one));
    unsigned long long nNodes = 0, nTris = 0, nInst = 0;

    while (true) {
        // ---- refill idle lanes ----------------------------------------------------------------------------------
        if (!exhausted) {
            const unsigned idle = __ballot_sync(0xffffffffu, !R.active);
            if (idle) {
                uint32_t base = 0;
                const int leader = __ffs(idle) - 1;
                if (lane == leader) base = atomicAdd(P.workCounter, (uint32_t)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!R.active) {
                    const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < total) {
                        R.anyHit = idx >= extCount;
                        R.index = idx;
                        ::float4 ro, rd;
                        if (!R.anyHit) {
                            ro = TQ_LD(P.rayO + idx); rd = TQ_LD(P.rayD + idx);
                        } else {
                            ro = TQ_LD(P.shO + (idx - extCount)); rd = TQ_LD(P.shD + (idx - extCount));
                        }
                        R.o = float3(ro.x, ro.y, ro.z);
                        R.d = float3(rd.x, rd.y, rd.z);
                        R.tMin = ro.w;
                        R.tBest = rd.w;
                        R.hitInst = VKRT_INVALID_INDEX;
                        R.hitPrim = VKRT_INVALID_INDEX;
                        R.hitU = R.hitV = 0.0f;
                        R.sawTransmissive = false;
                        R.active = true;
                        o = R.o; d = R.d;
                        idir = safeInvDir(d);
                        octinv = 7u ^ ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
                        inBlas = false;
                        if (FLAT) curInst = VKRT_INVALID_INDEX;  // no object-space ray cached yet
                        sp = 0;
                        G = make_uint2(A.tlasRoot, 0x80000000u);
                        Gt = make_uint2(0u, 0u);
                        if (A.instanceCount == 0u) G = make_uint2(0u, 0u);
                    }
                }
                if (base + (uint32_t)__popc(idle) >= total) exhausted = true;
            }
        }
        if (__ballot_sync(0xffffffffu, R.active) == 0u) break;

        // ---- traverse until this lane finishes or the warp wants a refill ------------------------------------------
        // Each trip of this loop is three short, warp-convergent sections (profiles/r01_notes.md: the previous
        // "test a node, then loop over all of its primitives" shape ran at 10.5 of 32 lanes):
        //   1. lanes without a pending primitive test ONE node (or unpack a popped primitive group),
        //   2. lanes with a pending primitive process ONE primitive (triangle test, or instance entry in the TLAS),
        //   3. lanes with nothing pending pop the stack / leave the BLAS / retire the ray.
        while (R.active) {
            if (Gt.y == 0u) {
                if (G.y & 0xff000000u) {
                    const int bit = 31 - __clz(G.y & 0xff000000u);
                    const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
                    G.y &= ~(1u << bit);
                    if (G.y & 0xff000000u) { if (sp < STACK) stack[sp++] = G; }
                    const uint32_t nodeIndex = G.x + __popc(G.y & 0xffu & ((1u << slot) - 1u));
                    uint32_t childBase, primBase, imask;
                    const uint32_t hits = intersectNode8(A.nodes + nodeIndex, o, idir, octinv, one, R.tMin, R.tBest, childBase, primBase, imask);
                    if (COUNT) nNodes++;
                    G = make_uint2(childBase, (hits & 0xff000000u) | imask);
                    Gt = make_uint2(primBase, hits & 0x00ffffffu);
                } else if (G.y) {  // a primitive group that was parked on the stack
                    Gt = G;
                    G = make_uint2(0u, 0u);
                }
            }

#if TRACE_PRIM_VOTE > 0
            // The primitive section runs when at least TRACE_PRIM_VOTE lanes have a pending primitive, or when no lane of the warp has
            // node work left to do meanwhile: pending lanes sit out a few node trips and then test their primitives together.
            const unsigned inLoop = __activemask();
            const unsigned pending = __ballot_sync(inLoop, Gt.y != 0u);
            const bool runPrims = __popc(pending) >= TRACE_PRIM_VOTE || pending == inLoop;
            if (Gt.y && runPrims) {
#else
            if (Gt.y) {
#endif
                const int i = __ffs(Gt.y) - 1;
                Gt.y &= Gt.y - 1u;
                const uint32_t primIndex = Gt.x + (uint32_t)i;
                if (FLAT) {
                    // ---- single-level BVH: the leaf-ordered triangle record carries its instance in b.w (ONE load stage per test); the
                    // triangle stays in object space and the ray is taken into the instance's space (cached per lane until the instance
                    // changes) -> same arithmetic, same bits as the two-level structure.
                    const ::float4* tri = A.triangles + (size_t)primIndex * 3;
                    const ::float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
                    const uint32_t triInst = __float_as_uint(b.w);
                    bool testable = true;
                    if (triInst != curInst) {
                        const InstanceRecord* rec = A.instances + triInst;
                        const ::uint4 meta = __ldg(reinterpret_cast<const ::uint4*>(rec) + 3);
                        const ::float4 r0 = __ldg(reinterpret_cast<const ::float4*>(rec));
                        const ::float4 r1 = __ldg(reinterpret_cast<const ::float4*>(rec) + 1);
                        const ::float4 r2 = __ldg(reinterpret_cast<const ::float4*>(rec) + 2);
                        const float4 i0(r0.x, r0.y, r0.z, r0.w), i1(r1.x, r1.y, r1.z, r1.w), i2(r2.x, r2.y, r2.z, r2.w);
                        objO = xformPoint(i0, i1, i2, R.o);
                        const float3 od = xformVector(i0, i1, i2, R.d);
                        curFlags = meta.y;
                        if (makeRayShear(od, shear)) {
                            curInst = triInst;
                            if (COUNT) nInst++;
                        } else {
                            curInst = VKRT_INVALID_INDEX;
                            testable = false;
                        }
                    }
                    if (testable && R.anyHit && R.sawTransmissive && (curFlags & INSTANCE_FLAG_TRANSMISSIVE)) testable = false;
                    if (testable) {
                        if (COUNT) nTris++;
                        float t, u, v;
                        bool accept = watertightTriangle(objO, shear, float3(a.x, a.y, a.z), float3(b.x, b.y, b.z), float3(c.x, c.y, c.z), t, u, v) && t > R.tMin;
                        const uint32_t prim = __float_as_uint(a.w);
                        if (accept) {
                            bool closer = t < R.tBest;
                            if (!closer && t == R.tBest && R.hitInst != VKRT_INVALID_INDEX)
                                closer = curInst < R.hitInst || (curInst == R.hitInst && prim < R.hitPrim);
                            accept = closer;
                        }
                        if (accept && (curFlags & INSTANCE_FLAG_ALPHA_TESTED)) {
                            const uint32_t seed = R.anyHit ? __ldg(P.shSeed + (R.index - extCount)) : (P.raySeed ? __ldg(P.raySeed + (size_t)R.index * P.slotStride) : 0u);
                            accept = alphaHitAccepted(P.scene, curInst, prim, float2(u, v), seed);
                        }
                        if (accept) {
                            if (!R.anyHit) {
                                R.hitInst = curInst;
                                R.hitPrim = prim;
                                R.tBest = t;
                                R.hitU = u;
                                R.hitV = v;
                            } else if (curFlags & INSTANCE_FLAG_TRANSMISSIVE) {
                                R.sawTransmissive = true;   // keep looking for an opaque occluder; triangles of transmissive instances are skipped from now on
                            } else {
                                R.hitInst = curInst;        // occluded: done
                                R.hitPrim = prim;
                                sp = 0;
                                G = make_uint2(0u, 0u);
                                Gt.y = 0u;
                            }
                        }
                    }
                } else if (!inBlas) {
                    // TLAS leaf = instance: park the remaining TLAS work and descend into the BLAS
                    if (Gt.y) { if (sp < STACK) stack[sp++] = Gt; }
                    if (G.y & 0xff000000u) { if (sp < STACK) stack[sp++] = G; }
                    G = make_uint2(0u, 0u);
                    Gt.y = 0u;
                    const InstanceRecord* rec = A.instances + primIndex;
                    // the whole 64-byte record in one load stage (skipped instances are rare)
                    const ::uint4 meta = __ldg(reinterpret_cast<const ::uint4*>(rec) + 3);
                    const ::float4 r0 = __ldg(reinterpret_cast<const ::float4*>(rec));
                    const ::float4 r1 = __ldg(reinterpret_cast<const ::float4*>(rec) + 1);
                    const ::float4 r2 = __ldg(reinterpret_cast<const ::float4*>(rec) + 2);
                    const bool skip = (meta.y & INSTANCE_FLAG_EMPTY) || (R.anyHit && R.sawTransmissive && (meta.y & INSTANCE_FLAG_TRANSMISSIVE));
                    if (!skip) {
                        const float4 i0(r0.x, r0.y, r0.z, r0.w), i1(r1.x, r1.y, r1.z, r1.w), i2(r2.x, r2.y, r2.z, r2.w);
                        const float3 oo = xformPoint(i0, i1, i2, R.o);
                        const float3 od = xformVector(i0, i1, i2, R.d);
                        if (makeRayShear(od, shear)) {
                            o = oo; d = od;
                            idir = safeInvDir(d);
                            octinv = 7u ^ ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
                            inBlas = true;
                            blasBase = sp;
                            curInst = meta.z;
                            curFlags = meta.y;
                            G = make_uint2(meta.x, 0x80000000u);
                            if (COUNT) nInst++;
                        }
                    }
                } else {
                    const ::float4* tri = A.triangles + (size_t)primIndex * 3;
                    const ::float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
                    if (COUNT) nTris++;
                    float t, u, v;
                    bool accept = watertightTriangle(o, shear, float3(a.x, a.y, a.z), float3(b.x, b.y, b.z), float3(c.x, c.y, c.z), t, u, v) && t > R.tMin;
                    const uint32_t prim = __float_as_uint(a.w);
                    if (accept) {
                        bool closer = t < R.tBest;
                        if (!closer && t == R.tBest && R.hitInst != VKRT_INVALID_INDEX)
                            closer = curInst < R.hitInst || (curInst == R.hitInst && prim < R.hitPrim);
                        accept = closer;
                    }
                    if (accept && (curFlags & INSTANCE_FLAG_ALPHA_TESTED)) {
                        const uint32_t seed = R.anyHit ? __ldg(P.shSeed + (R.index - extCount)) : (P.raySeed ? __ldg(P.raySeed + (size_t)R.index * P.slotStride) : 0u);
                        accept = alphaHitAccepted(P.scene, curInst, prim, float2(u, v), seed);
                    }
                    if (accept) {
                        if (!R.anyHit) {
                            R.hitInst = curInst;
                            R.hitPrim = prim;
                            R.tBest = t;
                            R.hitU = u;
                            R.hitV = v;
                        } else if (curFlags & INSTANCE_FLAG_TRANSMISSIVE) {
                            R.sawTransmissive = true;   // keep looking for an opaque occluder, but not in this instance
                            sp = blasBase;
                            G = make_uint2(0u, 0u);
                            Gt.y = 0u;
                        } else {
                            R.hitInst = curInst;        // occluded: done
                            R.hitPrim = prim;
                            sp = 0;
                            inBlas = false;
                            G = make_uint2(0u, 0u);
                            Gt.y = 0u;
                        }
                    }
                }
            }

            if (Gt.y == 0u && (G.y & 0xff000000u) == 0u) {
                if (inBlas && sp == blasBase) {  // BLAS exhausted: back to the world-space ray
                    inBlas = false;
                    o = R.o; d = R.d;
                    idir = safeInvDir(d);
                    octinv = 7u ^ ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
                }
                if (sp == 0) {
                    // ---- ray finished: write back ----------------------------------------------------------------
                    if (!R.anyHit) {
                        TQ_ST(P.hitA + R.index, make_uint4(R.hitInst, R.hitPrim, __float_as_uint(R.hitInst != VKRT_INVALID_INDEX ? R.tBest : 0.0f),
                                                           __float_as_uint(R.hitU)));
                        P.hitB[(size_t)R.index * P.slotStride] = R.hitV;
                    } else {
                        const uint32_t k = R.index - extCount;
                        const bool occluded = R.hitInst != VKRT_INVALID_INDEX;
                        if (P.shadowResult) P.shadowResult[k] = occluded ? 1u : (R.sawTransmissive ? 2u : 0u);
                        if (P.shTarget) {
                            const ::uint2 target = TQ_LD(P.shTarget + k);
                            if (!occluded && !R.sawTransmissive) {
                                const ::float4 c = TQ_LD(P.shContribution + k);
                                if (target.y & SHADOW_KIND_SCALAR) {
                                    atomicAdd(P.radianceScalar + target.x, c.x);
                                } else {
                                    // one shadow ray per record per launch: the reduction is race-free and fire-and-forget
                                    atomicAdd(P.radiance + target.x, c);
                                }
                            } else if (!occluded) {  // unsupported transmission: NEE is not trusted at this vertex
                                const uint32_t pos = target.y & 0x7fffffffu;
                                if (pos != 0x7fffffffu) P.pathFlags[(size_t)pos * P.slotStride] &= ~1u;
                            }
                        }
                    }
                    R.active = false;
                    break;
                }
                G = stack[--sp];
            }
            if (!exhausted && __popc(__activemask()) < (int)A.refillLanes) break;
        }
    }
    if (COUNT && P.stats) {
        atomicAdd(&P.stats[0], nNodes);
        atomicAdd(&P.stats[1], nTris);
        atomicAdd(&P.stats[2], nInst);
    }
}

} // namespace vk
