// accel.cuh — device acceleration-structure layout shared by the builder (bvh_build.cu) and the traversal kernels.
//
// Replaces the driver-side VkAccelerationStructureKHR objects the reference creates in
// src/core/render/accel/blas.c:222-262 (one BLAS per unique geometry, PREFER_FAST_TRACE) and
// src/core/render/accel/tlas.c:324-346,535-559 (one TLAS instance per mesh, instanceCustomIndex = mesh index).
//
// Layout in HBM (all arrays 16-byte aligned, read with 128-bit loads):
//   bvh8Nodes   : compressed 8-wide nodes, 80 B each (Ylitie, Karras, Laine 2017), BLASes first, TLAS last
//   triangles   : 3 x ::float4 per triangle in leaf order: v0.xyz|prim, v1.xyz|-, v2.xyz|-   (48 B, object space)
//   instances   : 64 B per TLAS leaf: inverse 3x4 (row-major, 3 x ::float4) + {blasRoot, triBase, flags, instanceIndex}
// Flat variant (vkrt_cuda_build_accel picks it when instancing does not pay): ONE BVH over the world-space boxes of all
// instanced triangles; leaves index leaf-ordered 48-byte triangle records {v0 | primitive, v1 | instance, v2} (gathered once after the
// build from the builder's `flatPrims` {triangle record, instance} list); triangles stay in object space and the ray is taken into
// the instance's space at the triangle test, so hits are bit-identical to the two-level structure.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vk {

struct alignas(16) Bvh8Node {
    float px, py, pz;         // quantisation origin
    uint8_t ex, ey, ez;       // per-axis exponent (biased, value = 2^(e-127))
    uint8_t imask;            // bit i set = child slot i is an internal node
    uint32_t childBase;       // index of the first internal child
    uint32_t primBase;        // index of the first primitive referenced by this node's leaf slots
    uint8_t meta[8];          // internal: 0b001xxxxx (24 + slot); leaf: unary count << 5 | offset; empty: 0
    uint8_t qlox[8], qloy[8], qloz[8];
    uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

struct alignas(16) InstanceRecord {
    ::float4 inv0, inv1, inv2;  // rows of the inverse world transform
    uint32_t blasRoot;        // node index of the BLAS root in bvh8Nodes
    uint32_t flags;           // INSTANCE_FLAG_*
    uint32_t instanceIndex;   // == mesh index (InstanceIndex(), closest_hit.slang:4)
    uint32_t pad;
};
static_assert(sizeof(InstanceRecord) == 64, "InstanceRecord must be 64 bytes");

enum : uint32_t {
    INSTANCE_FLAG_ALPHA_TESTED = 1u << 0,  // FORCE_NO_OPAQUE: candidates go through the stochastic alpha test
    INSTANCE_FLAG_TRANSMISSIVE = 1u << 1,  // material.transmission > 0 (shadow rays: "unsupported transmission")
    INSTANCE_FLAG_EMPTY = 1u << 2
};

// Binary LBVH as produced by the builder, kept for the collapse step (and for the debug BVH2 traversal).
// Node i < n-1 is internal, node (n-1)+k is the leaf holding sorted primitive k.
struct Lbvh {
    ::float4* lo;        // xyz = box min, w = left child index (int bits) for internal nodes
    ::float4* hi;        // xyz = box max, w = right child index
    uint32_t* parent;
    uint32_t primCount;
};

struct AccelView {
    const Bvh8Node* nodes;
    const ::float4* triangles;          // two-level: BLAS leaf order; flat: leaf order of the single BVH, b.w = instance
    const InstanceRecord* instances;    // two-level: TLAS leaf order; flat: instance order
    const ::uint2* flatPrims;           // flat only, leaf order: x = triangle record, y = instance (build product; traversal reads `triangles`)
    uint32_t tlasRoot;      // node index of the TLAS root (flat: of the single BVH)
    uint32_t instanceCount; // 0 = nothing to hit
    uint32_t flat;          // 1 = single-level BVH over instanced triangles
    uint32_t stackNeed;     // upper bound of the traversal stack entries a ray can need (from the built tree depths); picks the k_trace stack size
    uint32_t refillLanes;   // a warp of k_trace fetches new rays when fewer lanes than this are still traversing (trace.cuh)
};

} // namespace vk
