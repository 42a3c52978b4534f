// bvh_build.cu — GPU acceleration-structure builder for sm_100a.
//
// Replaces the driver builds recorded by the reference in src/core/render/accel/blas.c:222-262
// (vkCmdBuildAccelerationStructuresKHR, one BLAS per unique geometry) and src/core/render/accel/tlas.c:236-289,535-559
// (TLAS over one instance per mesh).  Pipeline per BVH (BLAS over triangles, TLAS over instance boxes):
//   1. primitive AABBs + scene bounds                         (k_triangle_bounds / k_instance_bounds)
//   2. Morton codes of the box centres, 3*b bits              (k_morton)
//   3. hand-written LSD radix sort, 8-bit digits, no CUB      (k_rs_histogram / k_rs_scan_* / k_rs_scatter)
//   4. Karras 2012 binary radix tree                          (k_hierarchy)
//   5. bottom-up AABB refit with per-node arrival counters    (k_refit)
//   6. greedy surface-area collapse to compressed 8-wide nodes (k_collapse), level by level
// All stages stream their inputs with coalesced 8/16-byte accesses; the sort is the HBM-bound part
// (DESIGN.md "Build" gives the algorithmic bytes per triangle).
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "accel.cuh"
#include "build.h"
#include "vmath.cuh"

namespace vk {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(err, sizeof(err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return false;                                                                          \
        }                                                                                          \
    } while (0)

// ---- ordered-int float atomics for the bounds reduction --------------------------------------------------------------
__device__ __forceinline__ int floatToOrdered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float orderedToFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_init_bounds(int* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = floatToOrdered(FLT_MAX);
    else if (threadIdx.x < 6) bounds[threadIdx.x] = floatToOrdered(-FLT_MAX);
    else if (threadIdx.x < 8) bounds[threadIdx.x] = 0;  // padding of the 8-int slot (copied to the host with the rest)
}

// Block-level reduction of per-thread boxes, then ONE set of six ordered-int atomics per block. Every thread of the block must call
// it (blockDim a multiple of 32, at most 1024). The kernels below accumulate over a grid-stride loop first, so a 10 M-triangle
// geometry issues ~7 k atomics instead of 1.9 M on the same six words (profiles/r01_notes.md: 2.1 ms -> bandwidth-bound).
__device__ __forceinline__ void reduceBounds(float3 lo, float3 hi, int* bounds) {
    __shared__ float sBox[6][32];
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = (blockDim.x + 31) >> 5;
    if (lane == 0) {
        sBox[0][warp] = lo.x; sBox[1][warp] = lo.y; sBox[2][warp] = lo.z;
        sBox[3][warp] = hi.x; sBox[4][warp] = hi.y; sBox[5][warp] = hi.z;
    }
    __syncthreads();
    if (warp == 0) {
        for (int c = 0; c < 6; c++) {
            float v = lane < warps ? sBox[c][lane] : (c < 3 ? FLT_MAX : -FLT_MAX);
            for (int o = 16; o > 0; o >>= 1) {
                const float w = __shfl_xor_sync(0xffffffffu, v, o);
                v = c < 3 ? fminf(v, w) : fmaxf(v, w);
            }
            if (lane == 0 && (c < 3 ? v < FLT_MAX : v > -FLT_MAX)) {
                if (c < 3) atomicMin(&bounds[c], floatToOrdered(v)); else atomicMax(&bounds[c], floatToOrdered(v));
            }
        }
    }
}
constexpr int BOUNDS_MAX_BLOCKS = 148 * 8;
static inline int boundsGrid(uint32_t n) { return (int)std::min<uint32_t>((n + 255u) / 256u, (uint32_t)BOUNDS_MAX_BLOCKS); }

// Pads a box by 2^-20 of its largest absolute coordinate so that the (rounded) watertight test can never accept a
// hit outside the boxes that cull for it.
__device__ __forceinline__ void padBox(float3& lo, float3& hi) {
    float m = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
    float eps = m * 9.5367431640625e-07f;
    lo = lo - float3(eps);
    hi = hi + float3(eps);
}

// A triangle with a non-finite (or absurdly large: extents and areas must not overflow) coordinate is INACTIVE, as in the Vulkan acceleration
// structure the reference builds ("if the X component of a vertex is NaN the triangle is inactive"): it keeps its slot and its primitive
// index but collapses to one finite point, so it is never hit and the builder only ever sees finite boxes.
__device__ __forceinline__ bool finiteCoord(float x) { return fabsf(x) <= 1.0e18f; }  // false for NaN
__device__ __forceinline__ bool finitePoint(float3 v) { return finiteCoord(v.x) && finiteCoord(v.y) && finiteCoord(v.z); }
__device__ __forceinline__ void deactivateIfNonFinite(float3& a, float3& b, float3& c) {
    if (finitePoint(a) && finitePoint(b) && finitePoint(c)) return;
    const float3 p = finitePoint(a) ? a : (finitePoint(b) ? b : (finitePoint(c) ? c : float3(0.0f)));
    a = b = c = p;
}

__global__ void k_triangle_bounds(const ShaderVertex* __restrict__ vertices, const uint32_t* __restrict__ indices, uint32_t vertexBase,
                                  uint32_t indexBase, uint32_t triCount, ::float4* __restrict__ primLo, ::float4* __restrict__ primHi,
                                  int* bounds) {
    float3 accLo(FLT_MAX), accHi(-FLT_MAX);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < triCount; t += gridDim.x * blockDim.x) {
        float3 v[3];
        for (int k = 0; k < 3; k++) {
            uint32_t vi = indices[indexBase + t * 3u + k] + vertexBase;
            ::float4 p = *reinterpret_cast<const ::float4*>(vertices[vi].position);
            v[k] = float3(p.x, p.y, p.z);
        }
        deactivateIfNonFinite(v[0], v[1], v[2]);
        float3 lo = min(min(v[0], v[1]), v[2]), hi = max(max(v[0], v[1]), v[2]);
        padBox(lo, hi);
        primLo[t] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        primHi[t] = make_float4(hi.x, hi.y, hi.z, 0.0f);
        accLo = min(accLo, lo);
        accHi = max(accHi, hi);
    }
    reduceBounds(accLo, accHi, bounds);
}

// World-space AABB of every instance = transformed corners of its BLAS root box (conservative), padded.
__global__ void k_instance_bounds(const ::float4* __restrict__ blasBounds, const uint32_t* __restrict__ instanceBlas,
                                  const float* __restrict__ world3x4, uint32_t instanceCount, ::float4* __restrict__ primLo,
                                  ::float4* __restrict__ primHi, int* bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < instanceCount;
    float3 lo(FLT_MAX), hi(-FLT_MAX);
    if (valid) {
        ::float4 blo = blasBounds[instanceBlas[i] * 2u], bhi = blasBounds[instanceBlas[i] * 2u + 1u];
        const float* m = world3x4 + (size_t)i * 12;
        if (blo.x <= bhi.x) {
            for (int k = 0; k < 8; k++) {
                float3 p((k & 1) ? bhi.x : blo.x, (k & 2) ? bhi.y : blo.y, (k & 4) ? bhi.z : blo.z);
                float3 w(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                         m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
                lo = min(lo, w);
                hi = max(hi, w);
            }
            padBox(lo, hi);
            // a second, slightly larger pad absorbs the rounding of the 8 corner transforms themselves
            padBox(lo, hi);
        } else {  // empty BLAS: degenerate box at the instance origin
            lo = hi = float3(m[3], m[7], m[11]);
        }
        primLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        primHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
    reduceBounds(lo, hi, bounds);
}

// Object-space AABB of one geometry (ordered-int atomics into bounds6), for the flat / two-level decision.
__global__ void k_geometry_bounds(const ShaderVertex* __restrict__ vertices, const uint32_t* __restrict__ indices, uint32_t vertexBase, uint32_t indexBase,
                                  uint32_t triCount, int* bounds6) {
    float3 lo(FLT_MAX), hi(-FLT_MAX);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < triCount; t += gridDim.x * blockDim.x) {
        float3 v[3];
        for (int k = 0; k < 3; k++) {
            uint32_t vi = indices[indexBase + t * 3u + k] + vertexBase;
            ::float4 p = *reinterpret_cast<const ::float4*>(vertices[vi].position);
            v[k] = float3(p.x, p.y, p.z);
        }
        deactivateIfNonFinite(v[0], v[1], v[2]);
        lo = min(lo, min(min(v[0], v[1]), v[2]));
        hi = max(hi, max(max(v[0], v[1]), v[2]));
    }
    reduceBounds(lo, hi, bounds6);
}
void launchGeometryBounds(const ShaderVertex* vertices, const uint32_t* indices, uint32_t vertexBase, uint32_t indexBase, uint32_t triCount, int* bounds6,
                          cudaStream_t st) {
    k_init_bounds<<<1, 32, 0, st>>>(bounds6);
    if (triCount) k_geometry_bounds<<<boundsGrid(triCount), 256, 0, st>>>(vertices, indices, vertexBase, indexBase, triCount, bounds6);
}
// Raises *flag when an index of the range does not address a vertex of its geometry (the ABI copies caller data in; an index past the
// vertex buffer would otherwise become an out-of-bounds fetch in every later kernel).
__global__ void k_validate_indices(const uint32_t* __restrict__ indices, uint32_t indexBase, uint32_t indexCount, uint32_t vertexLimit, uint32_t* flag) {
    bool bad = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < indexCount; i += gridDim.x * blockDim.x) bad |= indices[indexBase + i] >= vertexLimit;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
void launchValidateIndices(const uint32_t* indices, uint32_t indexBase, uint32_t indexCount, uint32_t vertexLimit, uint32_t* flag, cudaStream_t st) {
    if (indexCount) k_validate_indices<<<boundsGrid(indexCount), 256, 0, st>>>(indices, indexBase, indexCount, vertexLimit, flag);
}
float orderedIntToFloatHost(int i) {
    int b = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

// ---- flat (single-level) build inputs ----------------------------------------------------------------------------------
// Enumerates every (instance, local triangle) pair: triOffsets is the exclusive prefix sum of the instances' triangle counts.
__global__ void k_flat_enumerate(const uint32_t* __restrict__ triOffsets, uint32_t instanceCount, uint32_t total, uint2* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    uint32_t lo = 0, hi = instanceCount;  // last instance whose offset <= p
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (triOffsets[mid] <= p) lo = mid; else hi = mid;
    }
    out[p] = make_uint2(lo, p - triOffsets[lo]);
}
// World-space AABB of an instanced triangle: the three vertices through the instance's 3x4 world transform, padded twice (once for
// the watertight test's own rounding, once for the rounding of these transforms; the hit itself is computed in object space).
__global__ void k_flat_bounds(const ShaderVertex* __restrict__ vertices, const uint32_t* __restrict__ indices, const uint2* __restrict__ flatIn,
                              const uint4* __restrict__ flatInstances, const float* __restrict__ world3x4, uint32_t total,
                              ::float4* __restrict__ primLo, ::float4* __restrict__ primHi, int* bounds) {
    float3 accLo(FLT_MAX), accHi(-FLT_MAX);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        float3 lo(FLT_MAX), hi(-FLT_MAX);
        const uint2 fp = flatIn[p];
        const uint4 inst = flatInstances[fp.x];
        const float* m = world3x4 + (size_t)fp.x * 12;
        float3 o[3], w[3];
        for (int k = 0; k < 3; k++) {
            uint32_t vi = indices[inst.y + fp.y * 3u + k] + inst.x;
            ::float4 q = *reinterpret_cast<const ::float4*>(vertices[vi].position);
            o[k] = float3(q.x, q.y, q.z);
        }
        deactivateIfNonFinite(o[0], o[1], o[2]);  // the same object-space decision as the packed triangle record (k_pack_triangles)
        for (int k = 0; k < 3; k++)
            w[k] = float3(m[0] * o[k].x + m[1] * o[k].y + m[2] * o[k].z + m[3], m[4] * o[k].x + m[5] * o[k].y + m[6] * o[k].z + m[7],
                          m[8] * o[k].x + m[9] * o[k].y + m[10] * o[k].z + m[11]);
        if (!(finitePoint(w[0]) && finitePoint(w[1]) && finitePoint(w[2]))) w[0] = w[1] = w[2] = float3(0.0f);  // a transform that overflows
        lo = min(min(w[0], w[1]), w[2]);
        hi = max(max(w[0], w[1]), w[2]);
        padBox(lo, hi);
        padBox(lo, hi);
        primLo[p] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        primHi[p] = make_float4(hi.x, hi.y, hi.z, 0.0f);
        accLo = min(accLo, lo);
        accHi = max(accHi, hi);
    }
    reduceBounds(accLo, accHi, bounds);
}
// Object-space triangle records of one unique geometry in primitive order (a.w = primitive index), shared by all of its instances.
__global__ void k_pack_triangles(const ShaderVertex* __restrict__ vertices, const uint32_t* __restrict__ indices, uint32_t vertexBase, uint32_t indexBase,
                                 uint32_t triCount, ::float4* __restrict__ trianglesOut, uint32_t primBase) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    ::float4 v[3];
    for (int k = 0; k < 3; k++) {
        uint32_t vi = indices[indexBase + t * 3u + k] + vertexBase;
        v[k] = *reinterpret_cast<const ::float4*>(vertices[vi].position);
    }
    {
        float3 a(v[0].x, v[0].y, v[0].z), b(v[1].x, v[1].y, v[1].z), c(v[2].x, v[2].y, v[2].z);
        deactivateIfNonFinite(a, b, c);   // an inactive triangle becomes a zero-area record: the watertight test rejects it (det == 0)
        v[0] = make_float4(a.x, a.y, a.z, 0.0f); v[1] = make_float4(b.x, b.y, b.z, 0.0f); v[2] = make_float4(c.x, c.y, c.z, 0.0f);
    }
    v[0].w = __uint_as_float(t);
    v[1].w = 0.0f;
    v[2].w = 0.0f;
    const size_t o = (size_t)(primBase + t) * 3;
    trianglesOut[o] = v[0]; trianglesOut[o + 1] = v[1]; trianglesOut[o + 2] = v[2];
}
void launchPackTriangles(const ShaderVertex* vertices, const uint32_t* indices, uint32_t vertexBase, uint32_t indexBase, uint32_t triCount,
                         ::float4* trianglesOut, uint32_t primBase, cudaStream_t st) {
    if (triCount) k_pack_triangles<<<(triCount + 255) / 256, 256, 0, st>>>(vertices, indices, vertexBase, indexBase, triCount, trianglesOut, primBase);
}

// ---- Morton codes ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const ::float4* __restrict__ primLo, const ::float4* __restrict__ primHi, uint32_t n, const int* __restrict__ bounds,
                         uint32_t bitsPerAxis, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 blo(orderedToFloat(bounds[0]), orderedToFloat(bounds[1]), orderedToFloat(bounds[2]));
    float3 bhi(orderedToFloat(bounds[3]), orderedToFloat(bounds[4]), orderedToFloat(bounds[5]));
    ::float4 lo = primLo[i], hi = primHi[i];
    float3 c((lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f);
    float3 ext = bhi - blo;
    float scale = float(1u << bitsPerAxis);
    float3 nrm((ext.x > 0.0f ? (c.x - blo.x) / ext.x : 0.0f), (ext.y > 0.0f ? (c.y - blo.y) / ext.y : 0.0f),
               (ext.z > 0.0f ? (c.z - blo.z) / ext.z : 0.0f));
    uint32_t maxq = (1u << bitsPerAxis) - 1u;
    uint32_t qx = min((uint32_t)fmaxf(nrm.x * scale, 0.0f), maxq);
    uint32_t qy = min((uint32_t)fmaxf(nrm.y * scale, 0.0f), maxq);
    uint32_t qz = min((uint32_t)fmaxf(nrm.z * scale, 0.0f), maxq);
    keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    vals[i] = i;
}

// ---- LSD radix sort (64-bit keys, 32-bit values), 8-bit digits -------------------------------------------------------
// Tile = 8 warps x 16 rounds x 32 keys = 4096 keys; warp w owns the contiguous chunk [w*512, (w+1)*512) of the tile so
// that stability only needs (a) per-warp digit counts, (b) an exclusive prefix over warps, (c) in-round match_any ranks.
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = 8;
constexpr int RS_ROUNDS = 8;   // 2048-key tiles: 24 KB of keys + values in shared memory for the ordered scatter
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;

__device__ __forceinline__ void warpDigitCount(uint32_t (*warpHist)[256], int warp, uint32_t digit, bool valid) {
    unsigned active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        unsigned mask = __match_any_sync(active, digit);
        int leader = __ffs(mask) - 1;
        if ((int)(threadIdx.x & 31) == leader) warpHist[warp][digit] += __popc(mask);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_histogram(const uint64_t* __restrict__ keys, uint32_t n, uint32_t shift,
                                                             uint32_t numTiles, uint32_t* __restrict__ tileHist) {
    __shared__ uint32_t warpHist[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warpHist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE + warp * (RS_ROUNDS * 32);
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t idx = base + r * 32 + lane;
        bool valid = idx < n;
        uint32_t digit = valid ? (uint32_t)((keys[idx] >> shift) & 0xffu) : 0u;
        warpDigitCount(warpHist, warp, digit, valid);
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) total += warpHist[w][threadIdx.x];
    tileHist[threadIdx.x * numTiles + blockIdx.x] = total;  // digit-major so that one block can scan one digit
}

// One block per digit: exclusive scan of that digit's per-tile counts (in place) + digit total.
__global__ void __launch_bounds__(1024) k_rs_scan_tiles(uint32_t* __restrict__ tileHist, uint32_t numTiles, uint32_t* __restrict__ digitTotals) {
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    uint32_t* row = tileHist + (size_t)blockIdx.x * numTiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < numTiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < numTiles ? row[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)(threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t s = warpSums[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if ((int)threadIdx.x >= o) s += y;
            }
            warpSums[threadIdx.x] = s;
        }
        __syncthreads();
        uint32_t warpOffset = (threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u;
        uint32_t c = carry;
        if (i < numTiles) row[i] = c + warpOffset + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warpOffset + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) digitTotals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_rs_scan_digits(uint32_t* __restrict__ digitTotals) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = digitTotals[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < 256; i++) {
            uint32_t v = s[i];
            s[i] = acc;
            acc += v;
        }
    }
    __syncthreads();
    digitTotals[256 + threadIdx.x] = s[threadIdx.x];  // exclusive digit bases
}

// Scatter of one 8-bit pass. The tile is first ordered by digit in shared memory (stable: warp, round, lane order is the input
// order), then written out: element i of the ordered tile goes to globalBase[digit] + i, so that the keys of one digit leave as one
// contiguous run (a 2048-key tile has 8-key = 64-byte runs on average) instead of 32 scattered 8-byte stores per warp instruction.
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                           uint64_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t n,
                                                           uint32_t shift, uint32_t numTiles, const uint32_t* __restrict__ tileHist,
                                                           const uint32_t* __restrict__ digitTotals) {
    __shared__ uint32_t warpHist[RS_WARPS][256];   // per-warp digit counts, then running local offsets
    __shared__ uint32_t globalBase[256];           // output position of the tile's first key of a digit, minus its local position
    __shared__ uint32_t scanTmp[RS_WARPS];
    __shared__ uint64_t sKeys[RS_TILE];
    __shared__ uint32_t sVals[RS_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warpHist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tileBase = blockIdx.x * RS_TILE;
    const uint32_t base = tileBase + warp * (RS_ROUNDS * 32);
    const uint32_t tileCount = min((uint32_t)RS_TILE, n - tileBase);
    uint64_t key[RS_ROUNDS];
    uint32_t val[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t idx = base + r * 32 + lane;
        bool valid = idx < n;
        key[r] = valid ? keysIn[idx] : 0ull;
        val[r] = valid ? valsIn[idx] : 0u;
        warpDigitCount(warpHist, warp, (uint32_t)((key[r] >> shift) & 0xffu), valid);
    }
    __syncthreads();
    {   // thread d: block-exclusive start of digit d (scan over the 256 digit totals), then per-warp running offsets
        const uint32_t d = threadIdx.x;
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) total += warpHist[w][d];
        uint32_t x = total;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) scanTmp[warp] = x;
        __syncthreads();
        uint32_t warpOffset = 0;
        for (int w = 0; w < warp; w++) warpOffset += scanTmp[w];
        const uint32_t localStart = warpOffset + x - total;
        globalBase[d] = digitTotals[256 + d] + tileHist[d * numTiles + blockIdx.x] - localStart;
        uint32_t running = localStart;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = warpHist[w][d];
            warpHist[w][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t idx = base + r * 32 + lane;
        bool valid = idx < n;
        uint32_t digit = (uint32_t)((key[r] >> shift) & 0xffu);
        unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned mask = __match_any_sync(active, digit);
            uint32_t rank = __popc(mask & ((1u << lane) - 1u));
            uint32_t off = warpHist[warp][digit];
            __syncwarp(mask);
            if (rank == 0) warpHist[warp][digit] = off + __popc(mask);
            sKeys[off + rank] = key[r];
            sVals[off + rank] = val[r];
        }
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tileCount; i += RS_THREADS) {
        const uint64_t k = sKeys[i];
        const uint32_t pos = globalBase[(uint32_t)((k >> shift) & 0xffu)] + i;
        keysOut[pos] = k;
        valsOut[pos] = sVals[i];
    }
}

// Flat variant: leaf-ordered triangle records with the instance in b.w, gathered from the per-geometry records (traversal then needs
// one load stage per triangle test instead of {leaf entry -> record}).
__global__ void __launch_bounds__(256) k_flat_gather_triangles(const uint2* __restrict__ flatPrims, const ::float4* __restrict__ src, uint32_t total,
                                                               ::float4* __restrict__ dst) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint2 fp = flatPrims[i];
        const ::float4* s = src + (size_t)fp.x * 3;
        ::float4 a = s[0], b = s[1], c = s[2];
        b.w = __uint_as_float(fp.y);
        ::float4* d = dst + (size_t)i * 3;
        d[0] = a; d[1] = b; d[2] = c;
    }
}
void launchFlatGatherTriangles(const uint2* flatPrims, const ::float4* src, uint32_t total, ::float4* dst, cudaStream_t st) {
    if (total) k_flat_gather_triangles<<<boundsGrid(total), 256, 0, st>>>(flatPrims, src, total, dst);
}

// ---- Karras 2012 hierarchy -------------------------------------------------------------------------------------------
__device__ __forceinline__ int deltaKeys(const uint64_t* __restrict__ keys, int n, int i, uint64_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t kj = keys[j];
    if (ki == kj) return 64 + __clz((uint32_t)(i ^ j));
    return __clzll((long long)(ki ^ kj));
}

// Node numbering: internal i in [0, n-1); leaf k stored at (n-1)+k. lo.w / hi.w of internal nodes hold the child indices.
__global__ void k_hierarchy(const uint64_t* __restrict__ keys, uint32_t n, ::float4* __restrict__ nodeLo, ::float4* __restrict__ nodeHi,
                            uint32_t* __restrict__ parent, uint32_t* __restrict__ subFirst, uint32_t* __restrict__ subCount) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int N = (int)n;
    if (i >= N - 1) return;
    uint64_t ki = keys[i];
    int d = (deltaKeys(keys, N, i, ki, i + 1) - deltaKeys(keys, N, i, ki, i - 1)) >= 0 ? 1 : -1;
    int dmin = deltaKeys(keys, N, i, ki, i - d);
    int lmax = 2;
    while (deltaKeys(keys, N, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (deltaKeys(keys, N, i, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = deltaKeys(keys, N, i, ki, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (deltaKeys(keys, N, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? d : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    uint32_t left = (lo == gamma) ? (uint32_t)(N - 1 + gamma) : (uint32_t)gamma;
    uint32_t right = (hi == gamma + 1) ? (uint32_t)(N - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    nodeLo[i].w = __uint_as_float(left);
    nodeHi[i].w = __uint_as_float(right);
    parent[left] = (uint32_t)i;
    parent[right] = (uint32_t)i;
    subFirst[i] = (uint32_t)lo;
    subCount[i] = (uint32_t)(hi - lo + 1);
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

__global__ void k_leaves(const ::float4* __restrict__ primLo, const ::float4* __restrict__ primHi, const uint32_t* __restrict__ sortedVals,
                         uint32_t n, ::float4* __restrict__ nodeLo, ::float4* __restrict__ nodeHi, uint32_t* __restrict__ subFirst,
                         uint32_t* __restrict__ subCount, uint32_t* __restrict__ parent) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t p = sortedVals[k];
    ::float4 lo = primLo[p], hi = primHi[p];
    lo.w = __uint_as_float(p);  // leaves carry the original primitive index
    hi.w = __uint_as_float(0xFFFFFFFFu);
    nodeLo[n - 1 + k] = lo;
    nodeHi[n - 1 + k] = hi;
    subFirst[n - 1 + k] = k;
    subCount[n - 1 + k] = 1;
    if (n == 1) parent[0] = 0xFFFFFFFFu;
}

__global__ void k_refit(uint32_t n, ::float4* nodeLo, ::float4* nodeHi, const uint32_t* __restrict__ parent, uint32_t* arrival) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || n < 2) return;
    uint32_t cur = parent[n - 1 + k];
    while (cur != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(&arrival[cur], 1u) == 0u) return;  // first child to arrive: the sibling will finish this node
        volatile ::float4* vlo = nodeLo;
        volatile ::float4* vhi = nodeHi;
        uint32_t l = __float_as_uint(vlo[cur].w), r = __float_as_uint(vhi[cur].w);
        float lx0 = vlo[l].x, ly0 = vlo[l].y, lz0 = vlo[l].z, hx0 = vhi[l].x, hy0 = vhi[l].y, hz0 = vhi[l].z;
        float lx1 = vlo[r].x, ly1 = vlo[r].y, lz1 = vlo[r].z, hx1 = vhi[r].x, hy1 = vhi[r].y, hz1 = vhi[r].z;
        vlo[cur].x = fminf(lx0, lx1); vlo[cur].y = fminf(ly0, ly1); vlo[cur].z = fminf(lz0, lz1);
        vhi[cur].x = fmaxf(hx0, hx1); vhi[cur].y = fmaxf(hy0, hy1); vhi[cur].z = fmaxf(hz0, hz1);
        cur = parent[cur];
    }
}


// ---- PLOC hierarchy (Meister & Bittner 2018, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction") ----
// The reference asks its driver for PREFER_FAST_TRACE acceleration structures (src/core/render/accel/blas.c:39, tlas.c:188). A Karras
// radix tree follows the Morton code alone: ten wall-sized triangles among 70 k small ones end up deep inside it (the flat BVH over the
// cornell box cost 1.7x the node visits of the two-level one, profiles/r01_notes.md). PLOC keeps the Morton ORDER but builds the tree
// bottom-up by surface area: every cluster looks `PLOC_RADIUS` neighbours to each side for the partner whose merged box is smallest, and
// mutually-nearest pairs merge; surviving clusters are compacted in order and the search repeats. Large boxes find no cheap partner and
// stay near the root. Each iteration = neighbour search + flags (k_ploc_search), one block-level scan (k_ploc_scan), merge + compaction
// (k_ploc_apply); the last <= PLOC_TAIL clusters are finished by one block in shared memory (k_ploc_tail).
// Node numbering as k_hierarchy's: leaves at (n-1)+k, internal nodes 0..n-2 with the ROOT AT 0 (ids are handed out from n-2 downwards,
// and a binary tree over n leaves has exactly n-1 merges).
#ifndef VKRT_PLOC_RADIUS
#define VKRT_PLOC_RADIUS 10
#endif
constexpr int PLOC_RADIUS = VKRT_PLOC_RADIUS;
constexpr int PLOC_THREADS = 256;
constexpr int PLOC_TAIL = 512;

struct PlocState {          // device-resident
    uint32_t count, merged, iterations;   // clusters alive, merges so far (= internal nodes created), iterations run
    uint32_t prevCount, prevMerged;       // the same two before the iteration in flight (k_ploc_apply works on those)
    uint32_t pad[3];
};

__device__ __forceinline__ float mergedHalfArea(float3 alo, float3 ahi, float3 blo, float3 bhi) {
    const float ex = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), ey = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), ez = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return ex * ey + ey * ez + ez * ex;
}
// Candidate j beats the current best for cluster i. Ties (identical boxes are common: instanced or degenerate triangles) are broken by
// distance in the order and then towards the partner i ^ 1, so that runs of equal boxes pair up (0,1)(2,3)... instead of forming a chain
// in which only one pair per iteration is mutual.
__device__ __forceinline__ bool plocBetter(float cost, int i, int j, float bestCost, int bestJ) {
    if (cost != bestCost) return cost < bestCost;
    const int dj = j > i ? j - i : i - j, db = bestJ > i ? bestJ - i : i - bestJ;
    if (dj != db) return dj < db;
    return j == (i ^ 1);
}

// Nearest neighbour of every cluster of this block's tile (and of the halo, so that "mutual" can be decided without another pass),
// the survival / merge flags, their block sums, and nn[] for k_ploc_apply.
__global__ void __launch_bounds__(PLOC_THREADS) k_ploc_search(const PlocState* __restrict__ state, const uint32_t* __restrict__ clusters,
                                                                const ::float4* __restrict__ nodeLo, const ::float4* __restrict__ nodeHi,
                                                                uint32_t* __restrict__ nn, uint2* __restrict__ blockCounts, uint32_t forcePairs) {
    constexpr int R = PLOC_RADIUS, T = PLOC_THREADS, W = T + 4 * R;
    __shared__ float sLx[W], sLy[W], sLz[W], sHx[W], sHy[W], sHz[W];
    __shared__ int sNN[T + 2 * R];
    __shared__ uint32_t sKeep[T / 32], sMerge[T / 32];
    const int c = (int)state->count;
    const int tile0 = blockIdx.x * T;
    if (tile0 >= c) return;
    const int base = tile0 - 2 * R;   // global index of sLo[0]
    for (int k = threadIdx.x; k < W; k += T) {
        const int g = base + k;
        if (g >= 0 && g < c) {
            const uint32_t node = clusters[g];
            const ::float4 lo = __ldg(nodeLo + node), hi = __ldg(nodeHi + node);
            sLx[k] = lo.x; sLy[k] = lo.y; sLz[k] = lo.z;
            sHx[k] = hi.x; sHy[k] = hi.y; sHz[k] = hi.z;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < T + 2 * R; k += T) {   // element g = tile0 - R + k
        const int g = tile0 - R + k;
        int best = -1;
        if (g >= 0 && g < c) {
            if (forcePairs) {
                best = (g ^ 1) < c ? (g ^ 1) : -1;
            } else {
                float bestCost = FLT_MAX;
                const int s = k + R;   // index of g in sLo
                const float3 alo(sLx[s], sLy[s], sLz[s]), ahi(sHx[s], sHy[s], sHz[s]);
                for (int d = -R; d <= R; d++) {
                    const int j = g + d;
                    if (d == 0 || j < 0 || j >= c) continue;
                    const float cost = mergedHalfArea(alo, ahi, float3(sLx[s + d], sLy[s + d], sLz[s + d]), float3(sHx[s + d], sHy[s + d], sHz[s + d]));
                    if (best < 0 || plocBetter(cost, g, j, bestCost, best)) { bestCost = cost; best = j; }
                }
            }
        }
        sNN[k] = best;
    }
    __syncthreads();
    const int g = tile0 + threadIdx.x;
    bool keep = false, merge = false;
    if (g < c) {
        const int j = sNN[threadIdx.x + R];
        nn[g] = (uint32_t)j;
        const bool mutual = j >= 0 && sNN[j - (tile0 - R)] == g;
        merge = mutual && g < j;
        keep = !(mutual && g > j);
    }
    const unsigned km = __ballot_sync(0xffffffffu, keep), mm = __ballot_sync(0xffffffffu, merge);
    if ((threadIdx.x & 31) == 0) { sKeep[threadIdx.x >> 5] = __popc(km); sMerge[threadIdx.x >> 5] = __popc(mm); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0, b = 0;
        for (int w = 0; w < T / 32; w++) { a += sKeep[w]; b += sMerge[w]; }
        blockCounts[blockIdx.x] = make_uint2(a, b);
    }
}

// Exclusive scan of the per-block {survivors, merges} (one block; <= n / PLOC_THREADS entries), new cluster count and merge total.
__global__ void __launch_bounds__(1024) k_ploc_scan(PlocState* __restrict__ state, uint2* __restrict__ blockCounts, uint2* __restrict__ blockOffsets) {
    __shared__ uint32_t sA[1024], sB[1024];
    __shared__ uint32_t carryA, carryB;
    const uint32_t c = state->count;
    const uint32_t blocks = (c + PLOC_THREADS - 1) / PLOC_THREADS;
    if (threadIdx.x == 0) { carryA = 0; carryB = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint2 v = i < blocks ? blockCounts[i] : make_uint2(0u, 0u);
        sA[threadIdx.x] = v.x; sB[threadIdx.x] = v.y;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const uint32_t a = threadIdx.x >= (uint32_t)o ? sA[threadIdx.x - o] : 0u, b = threadIdx.x >= (uint32_t)o ? sB[threadIdx.x - o] : 0u;
            __syncthreads();
            sA[threadIdx.x] += a; sB[threadIdx.x] += b;
            __syncthreads();
        }
        if (i < blocks) blockOffsets[i] = make_uint2(carryA + sA[threadIdx.x] - v.x, carryB + sB[threadIdx.x] - v.y);
        __syncthreads();
        if (threadIdx.x == 1023) { carryA += sA[1023]; carryB += sB[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        state->prevCount = c;
        state->prevMerged = state->merged;
        state->count = carryA;
        state->merged += carryB;
        state->iterations++;
    }
}

// Creates the merged nodes and writes the surviving clusters, in order, to clustersOut.
__global__ void __launch_bounds__(PLOC_THREADS) k_ploc_apply(const PlocState* __restrict__ state, const uint32_t* __restrict__ clusters,
                                                               uint32_t* __restrict__ clustersOut, const uint32_t* __restrict__ nn,
                                                               const uint2* __restrict__ blockOffsets, ::float4* __restrict__ nodeLo, ::float4* __restrict__ nodeHi, uint32_t* __restrict__ subCount, uint32_t n) {
    __shared__ uint32_t sKeep[PLOC_THREADS / 32], sMerge[PLOC_THREADS / 32];
    const uint32_t T = PLOC_THREADS;
    const uint32_t oldCount = state->prevCount, mergedBefore = state->prevMerged;
    const uint32_t g = blockIdx.x * T + threadIdx.x;
    if (blockIdx.x * T >= oldCount) return;
    bool keep = false, merge = false;
    uint32_t j = 0xFFFFFFFFu;
    if (g < oldCount) {
        j = nn[g];
        const bool mutual = j != 0xFFFFFFFFu && nn[j] == g;
        merge = mutual && g < j;
        keep = !(mutual && g > j);
    }
    const unsigned km = __ballot_sync(0xffffffffu, keep), mm = __ballot_sync(0xffffffffu, merge);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sKeep[warp] = __popc(km); sMerge[warp] = __popc(mm); }
    __syncthreads();
    uint32_t keepBase = 0, mergeBase = 0;
    for (int w = 0; w < warp; w++) { keepBase += sKeep[w]; mergeBase += sMerge[w]; }
    const uint2 off = blockOffsets[blockIdx.x];
    const uint32_t pos = off.x + keepBase + __popc(km & ((1u << lane) - 1u));
    const uint32_t mrank = mergedBefore + off.y + mergeBase + __popc(mm & ((1u << lane) - 1u));
    if (merge) {
        const uint32_t a = clusters[g], b = clusters[j];
        const uint32_t id = (n - 2u) - mrank;
        const ::float4 alo = nodeLo[a], ahi = nodeHi[a], blo = nodeLo[b], bhi = nodeHi[b];
        nodeLo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), __uint_as_float(a));
        nodeHi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), __uint_as_float(b));
        subCount[id] = 2u;   // "not leaf-like" for the collapse (one primitive per leaf slot)
        clustersOut[pos] = id;
    } else if (keep) {
        clustersOut[pos] = clusters[g];
    }
}

// The last <= PLOC_TAIL clusters: every iteration in shared memory, one block, no host round trips.
__global__ void __launch_bounds__(PLOC_TAIL) k_ploc_tail(PlocState* __restrict__ state, const uint32_t* __restrict__ clusters, ::float4* __restrict__ nodeLo,
                                                           ::float4* __restrict__ nodeHi, uint32_t* __restrict__ subCount, uint32_t n) {
    constexpr int R = PLOC_RADIUS;
    __shared__ float sB[2][6][PLOC_TAIL];   // boxes, ping-pong: lo.xyz, hi.xyz
    __shared__ uint32_t sNode[2][PLOC_TAIL];
    __shared__ int sNN[PLOC_TAIL];
    __shared__ uint32_t sScanA[PLOC_TAIL], sScanB[PLOC_TAIL];
    __shared__ uint32_t sCount, sMerged;
    const int t = threadIdx.x;
    int c = (int)state->count;
    if (c > PLOC_TAIL) return;   // not yet: the host keeps iterating globally
    if (t < c) {
        const uint32_t node = clusters[t];
        const ::float4 lo = nodeLo[node], hi = nodeHi[node];
        sB[0][0][t] = lo.x; sB[0][1][t] = lo.y; sB[0][2][t] = lo.z;
        sB[0][3][t] = hi.x; sB[0][4][t] = hi.y; sB[0][5][t] = hi.z;
        sNode[0][t] = node;
    }
    if (t == 0) { sCount = (uint32_t)c; sMerged = state->merged; }
    __syncthreads();
    int cur = 0;
    uint32_t iterations = 0;
    while (c > 1) {
        int best = -1;
        if (t < c) {
            float bestCost = FLT_MAX;
            const float3 alo(sB[cur][0][t], sB[cur][1][t], sB[cur][2][t]), ahi(sB[cur][3][t], sB[cur][4][t], sB[cur][5][t]);
            for (int d = -R; d <= R; d++) {
                const int j = t + d;
                if (d == 0 || j < 0 || j >= c) continue;
                const float cost = mergedHalfArea(alo, ahi, float3(sB[cur][0][j], sB[cur][1][j], sB[cur][2][j]), float3(sB[cur][3][j], sB[cur][4][j], sB[cur][5][j]));
                if (best < 0 || plocBetter(cost, t, j, bestCost, best)) { bestCost = cost; best = j; }
            }
        }
        sNN[t] = best;
        __syncthreads();
        bool keep = false, merge = false;
        if (t < c) {
            const bool mutual = best >= 0 && sNN[best] == t;
            merge = mutual && t < best;
            keep = !(mutual && t > best);
        }
        sScanA[t] = keep ? 1u : 0u;
        sScanB[t] = merge ? 1u : 0u;
        __syncthreads();
        for (int o = 1; o < PLOC_TAIL; o <<= 1) {
            const uint32_t a = t >= o ? sScanA[t - o] : 0u, b = t >= o ? sScanB[t - o] : 0u;
            __syncthreads();
            sScanA[t] += a; sScanB[t] += b;
            __syncthreads();
        }
        const uint32_t mergedBefore = sMerged;
        const int nxt = 1 - cur;
        if (merge) {
            const uint32_t pos = sScanA[t] - 1u, id = (n - 2u) - (mergedBefore + sScanB[t] - 1u);
            const uint32_t a = sNode[cur][t], b = sNode[cur][best];
            float box[6];
            for (int k = 0; k < 3; k++) box[k] = fminf(sB[cur][k][t], sB[cur][k][best]);
            for (int k = 3; k < 6; k++) box[k] = fmaxf(sB[cur][k][t], sB[cur][k][best]);
            nodeLo[id] = make_float4(box[0], box[1], box[2], __uint_as_float(a));
            nodeHi[id] = make_float4(box[3], box[4], box[5], __uint_as_float(b));
            subCount[id] = 2u;
            for (int k = 0; k < 6; k++) sB[nxt][k][pos] = box[k];
            sNode[nxt][pos] = id;
        } else if (keep) {
            const uint32_t pos = sScanA[t] - 1u;
            for (int k = 0; k < 6; k++) sB[nxt][k][pos] = sB[cur][k][t];
            sNode[nxt][pos] = sNode[cur][t];
        }
        __syncthreads();
        if (t == 0) { sCount = sScanA[PLOC_TAIL - 1]; sMerged = mergedBefore + sScanB[PLOC_TAIL - 1]; }
        __syncthreads();
        c = (int)sCount;
        cur = nxt;
        iterations++;
    }
    if (t == 0) { state->count = (uint32_t)c; state->merged = sMerged; state->iterations += iterations; }
}

// ---- collapse to compressed 8-wide nodes -----------------------------------------------------------------------------
struct CollapseParams {
    const ::float4* nodeLo;
    const ::float4* nodeHi;
    const uint32_t* subFirst;
    const uint32_t* subCount;
    const uint32_t* sortedVals;   // sorted position -> original primitive index
    uint32_t primCount;
    uint32_t maxLeaf;             // primitives per leaf slot (1; the node format allows up to 3)
    Bvh8Node* nodesOut;           // global node array
    uint32_t nodeBase;            // global index of this BVH's root node
    uint32_t* counters;           // [0] work-out count, [1] nodes allocated (local), [2] primitives emitted (local)
    // triangle emission
    const ShaderVertex* vertices;
    const uint32_t* indices;
    uint32_t vertexBase, indexBase;
    ::float4* trianglesOut;         // global triangle array (3 ::float4 per triangle)
    uint32_t primBase;            // global primitive index of this BVH's first emitted primitive
    // instance emission
    const InstanceRecord* instanceRecords;  // by instance index
    InstanceRecord* instancesOut;           // in TLAS leaf order
    // flat (single-level) emission: primitive p = (instance, local triangle) -> {object-space triangle record, instance}
    const uint2* flatIn;
    const uint4* flatInstances;             // per instance: vertexBase, indexBase, first triangle record, triangle count
    uint2* flatOut;                         // in leaf order
};

__device__ __forceinline__ float boxHalfArea(::float4 lo, ::float4 hi) {
    float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return ex * ey + ey * ez + ez * ex;
}

__global__ void __launch_bounds__(64) k_collapse(CollapseParams P, const uint2* __restrict__ workIn, uint32_t workCount, uint2* __restrict__ workOut) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= workCount) return;
    const uint2 w = workIn[t];
    const uint32_t n = P.primCount;
    const uint32_t b2 = w.x;
    auto isLeafLike = [&](uint32_t node) { return node >= n - 1 || P.subCount[node] <= P.maxLeaf; };

    uint32_t child[8];
    float area[8];
    int nc = 0;
    if (isLeafLike(b2)) {
        child[nc++] = b2;
    } else {
        child[0] = __float_as_uint(P.nodeLo[b2].w);
        child[1] = __float_as_uint(P.nodeHi[b2].w);
        nc = 2;
        for (int k = 0; k < 2; k++) area[k] = isLeafLike(child[k]) ? -1.0f : boxHalfArea(P.nodeLo[child[k]], P.nodeHi[child[k]]);
        while (nc < 8) {
            int best = -1;
            float bestArea = -1.0f;
            for (int k = 0; k < nc; k++)
                if (area[k] > bestArea) { bestArea = area[k]; best = k; }
            if (best < 0) break;
            uint32_t c = child[best];
            uint32_t l = __float_as_uint(P.nodeLo[c].w), r = __float_as_uint(P.nodeHi[c].w);
            child[best] = l;
            area[best] = isLeafLike(l) ? -1.0f : boxHalfArea(P.nodeLo[l], P.nodeHi[l]);
            child[nc] = r;
            area[nc] = isLeafLike(r) ? -1.0f : boxHalfArea(P.nodeLo[r], P.nodeHi[r]);
            nc++;
        }
    }

    // node frame
    ::float4 nlo = P.nodeLo[b2], nhi = P.nodeHi[b2];
    float3 center((nlo.x + nhi.x) * 0.5f, (nlo.y + nhi.y) * 0.5f, (nlo.z + nhi.z) * 0.5f);

    // greedy slot assignment: slot bit set = child on the positive side of that axis (x: bit0, y: bit1, z: bit2)
    ::float4 clo[8], chi[8];
    float cost[8][8];
    for (int k = 0; k < nc; k++) {
        clo[k] = P.nodeLo[child[k]];
        chi[k] = P.nodeHi[child[k]];
        float3 c((clo[k].x + chi[k].x) * 0.5f - center.x, (clo[k].y + chi[k].y) * 0.5f - center.y, (clo[k].z + chi[k].z) * 0.5f - center.z);
        for (int s = 0; s < 8; s++) cost[k][s] = ((s & 1) ? c.x : -c.x) + ((s & 2) ? c.y : -c.y) + ((s & 4) ? c.z : -c.z);
    }
    int slotOf[8];
    int childAt[8];
    for (int k = 0; k < 8; k++) { slotOf[k] = -1; childAt[k] = -1; }
    for (int it = 0; it < nc; it++) {
        float bestCost = -FLT_MAX;
        int bk = -1, bs = -1;
        for (int k = 0; k < nc; k++) {
            if (slotOf[k] >= 0) continue;
            for (int s = 0; s < 8; s++) {
                if (childAt[s] >= 0) continue;
                if (cost[k][s] > bestCost) { bestCost = cost[k][s]; bk = k; bs = s; }
            }
        }
        slotOf[bk] = bs;
        childAt[bs] = bk;
    }

    // quantisation exponents: smallest e with extent <= 255 * 2^e
    Bvh8Node node;
    node.px = nlo.x; node.py = nlo.y; node.pz = nlo.z;
    float ext[3] = {nhi.x - nlo.x, nhi.y - nlo.y, nhi.z - nlo.z};
    uint8_t eb[3];
    float inv2e[3], pow2e[3];
    for (int a = 0; a < 3; a++) {
        int e;
        if (!(ext[a] > 0.0f)) {
            e = -126;
        } else {
            e = (int)ceilf(log2f(ext[a] * (1.0f / 255.0f)));
            if (e < -126) e = -126;
            while (ext[a] > 255.0f * __uint_as_float((uint32_t)(e + 127) << 23)) e++;
        }
        eb[a] = (uint8_t)(e + 127);
        pow2e[a] = __uint_as_float((uint32_t)eb[a] << 23);
        inv2e[a] = 1.0f / pow2e[a];
    }
    node.ex = eb[0]; node.ey = eb[1]; node.ez = eb[2];

    // allocation
    uint32_t nInternal = 0, nPrims = 0;
    for (int k = 0; k < nc; k++) {
        if (isLeafLike(child[k])) nPrims += (child[k] >= n - 1) ? 1u : P.subCount[child[k]];
        else nInternal++;
    }
    uint32_t childBaseLocal = nInternal ? atomicAdd(&P.counters[1], nInternal) : 0u;
    uint32_t primBaseLocal = nPrims ? atomicAdd(&P.counters[2], nPrims) : 0u;
    uint32_t workBase = nInternal ? atomicAdd(&P.counters[0], nInternal) : 0u;
    node.childBase = P.nodeBase + childBaseLocal;
    node.primBase = P.primBase + primBaseLocal;
    node.imask = 0;

    uint32_t internalRank = 0, primOffset = 0;
    const float p3[3] = {nlo.x, nlo.y, nlo.z};
    for (int s = 0; s < 8; s++) {
        int k = childAt[s];
        if (k < 0) {
            node.meta[s] = 0;
            node.qlox[s] = node.qloy[s] = node.qloz[s] = 255;  // inverted box: never hit
            node.qhix[s] = node.qhiy[s] = node.qhiz[s] = 0;
            continue;
        }
        const float lo3[3] = {clo[k].x, clo[k].y, clo[k].z}, hi3[3] = {chi[k].x, chi[k].y, chi[k].z};
        uint8_t ql[3], qh[3];
        for (int a = 0; a < 3; a++) {
            float fl = floorf((lo3[a] - p3[a]) * inv2e[a]);
            float fh = ceilf((hi3[a] - p3[a]) * inv2e[a]);
            fl = fminf(fmaxf(fl, 0.0f), 255.0f);
            fh = fminf(fmaxf(fh, 0.0f), 255.0f);
            // make the decoded box provably conservative under the decode arithmetic p + q * 2^e
            while (fl > 0.0f && p3[a] + fl * pow2e[a] > lo3[a]) fl -= 1.0f;
            while (fh < 255.0f && p3[a] + fh * pow2e[a] < hi3[a]) fh += 1.0f;
            ql[a] = (uint8_t)fl;
            qh[a] = (uint8_t)fh;
        }
        node.qlox[s] = ql[0]; node.qloy[s] = ql[1]; node.qloz[s] = ql[2];
        node.qhix[s] = qh[0]; node.qhiy[s] = qh[1]; node.qhiz[s] = qh[2];
        uint32_t c = child[k];
        if (!isLeafLike(c)) {
            node.imask |= (uint8_t)(1u << s);
            node.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            workOut[workBase + internalRank] = make_uint2(c, childBaseLocal + internalRank);
            internalRank++;
        } else {
            uint32_t cnt = (c >= n - 1) ? 1u : P.subCount[c];
            uint32_t first = P.subFirst[c];
            uint32_t unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
            node.meta[s] = (uint8_t)((unary << 5) | primOffset);
            for (uint32_t q = 0; q < cnt; q++) {
                uint32_t prim = P.sortedVals[first + q];
                uint32_t outIndex = P.primBase + primBaseLocal + primOffset + q;
                if (P.flatOut) {
                    const uint2 fp = P.flatIn[prim];  // x = instance, y = local triangle
                    P.flatOut[outIndex] = make_uint2(P.flatInstances[fp.x].z + fp.y, fp.x);
                } else if (P.trianglesOut) {
                    ::float4 v[3];
                    for (int kk = 0; kk < 3; kk++) {
                        uint32_t vi = P.indices[P.indexBase + prim * 3u + kk] + P.vertexBase;
                        v[kk] = *reinterpret_cast<const ::float4*>(P.vertices[vi].position);
                    }
                    v[0].w = __uint_as_float(prim);
                    v[1].w = 0.0f;
                    v[2].w = 0.0f;
                    P.trianglesOut[(size_t)outIndex * 3 + 0] = v[0];
                    P.trianglesOut[(size_t)outIndex * 3 + 1] = v[1];
                    P.trianglesOut[(size_t)outIndex * 3 + 2] = v[2];
                } else {
                    P.instancesOut[outIndex] = P.instanceRecords[prim];
                }
            }
            primOffset += cnt;
        }
    }
    P.nodesOut[P.nodeBase + w.y] = node;
}

// Copies a finished BVH from the worst-case-sized scratch array to its final, compacted position; internal-child links
// are absolute node indices and move by (newBase - oldBase).
__global__ void k_relocate_nodes(const Bvh8Node* __restrict__ src, Bvh8Node* __restrict__ dst, uint32_t count, uint32_t oldBase, uint32_t newBase) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Bvh8Node n = src[i];
    n.childBase = n.childBase - oldBase + newBase;
    dst[i] = n;
}
void launchRelocateNodes(const Bvh8Node* src, Bvh8Node* dst, uint32_t count, uint32_t oldBase, uint32_t newBase, cudaStream_t st) {
    if (count) k_relocate_nodes<<<(count + 255) / 256, 256, 0, st>>>(src, dst, count, oldBase, newBase);
}

// ----------------------------------------------------------------------------------------------------------------------
// Host driver
// ----------------------------------------------------------------------------------------------------------------------
// Sum of the half-areas of the internal nodes of a binary hierarchy: with the leaves fixed, the surface-area-heuristic cost of two trees
// over the same primitives differs only in this sum.
__global__ void __launch_bounds__(256) k_tree_cost(const ::float4* __restrict__ nodeLo, const ::float4* __restrict__ nodeHi, uint32_t internalCount, double* __restrict__ out) {
    __shared__ double sSum[8];
    double acc = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < internalCount; i += gridDim.x * blockDim.x) {
        const ::float4 lo = nodeLo[i], hi = nodeHi[i];
        const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
        const float a = ex * ey + ey * ez + ez * ex;
        if (a == a && a < 3.0e38f) acc += (double)a;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sSum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sSum[w];
        atomicAdd(out, t);
    }
}

__global__ void k_ploc_init(uint32_t* __restrict__ clusters, uint32_t n) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) clusters[k] = n - 1u + k;   // leaf k of the Morton order
}

static uint32_t ceilLog2(uint32_t v) {
    uint32_t r = 0;
    while ((1ull << r) < v) r++;
    return r;
}

void AccelBuilder::release() {
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(primLo); F(primHi); F(keys[0]); F(keys[1]); F(vals[0]); F(vals[1]); F(tileHist); F(digitTotals); F(nodeLo); F(nodeHi);
    F(parent); F(arrival); F(subFirst); F(subCount); F(work[0]); F(work[1]); F(counters); F(bounds); F(plocState);
    F(nodeLoB); F(nodeHiB); F(plocClusters); F(plocNN); F(treeCost);
    plocState = nullptr;
    nodeLoB = nodeHiB = nullptr;
    plocClusters = plocNN = nullptr;
    treeCost = nullptr;
    primLo = primHi = nodeLo = nodeHi = nullptr;
    keys[0] = keys[1] = nullptr;
    vals[0] = vals[1] = nullptr;
    tileHist = digitTotals = parent = arrival = subFirst = subCount = counters = nullptr;
    work[0] = work[1] = nullptr;
    bounds = nullptr;
    capacity = 0;
}

bool AccelBuilder::reserve(uint32_t n) {
    if (n <= capacity) return true;
    release();
    uint32_t cap = n < 1024 ? 1024 : n;
    uint32_t tiles = (cap + RS_TILE - 1) / RS_TILE;
    CK(cudaMalloc(&primLo, sizeof(::float4) * cap));
    CK(cudaMalloc(&primHi, sizeof(::float4) * cap));
    for (int k = 0; k < 2; k++) {
        CK(cudaMalloc(&keys[k], sizeof(uint64_t) * cap));
        CK(cudaMalloc(&vals[k], sizeof(uint32_t) * cap));
        CK(cudaMalloc(&work[k], sizeof(uint2) * cap));
    }
    CK(cudaMalloc(&tileHist, sizeof(uint32_t) * 256 * tiles));
    CK(cudaMalloc(&digitTotals, sizeof(uint32_t) * 512));
    CK(cudaMalloc(&nodeLo, sizeof(::float4) * 2 * cap));
    CK(cudaMalloc(&nodeHi, sizeof(::float4) * 2 * cap));
    CK(cudaMalloc(&parent, sizeof(uint32_t) * 2 * cap));
    CK(cudaMalloc(&arrival, sizeof(uint32_t) * cap));
    CK(cudaMalloc(&subFirst, sizeof(uint32_t) * 2 * cap));
    CK(cudaMalloc(&subCount, sizeof(uint32_t) * 2 * cap));
    CK(cudaMalloc(&counters, sizeof(uint32_t) * 4));
    CK(cudaMalloc(&bounds, sizeof(int) * 8));
    CK(cudaMalloc(&plocState, 64));
    CK(cudaMalloc(&nodeLoB, sizeof(::float4) * 2 * cap));
    CK(cudaMalloc(&nodeHiB, sizeof(::float4) * 2 * cap));
    CK(cudaMalloc(&plocClusters, sizeof(uint32_t) * 2 * cap));
    CK(cudaMalloc(&plocNN, sizeof(uint32_t) * cap));
    CK(cudaMalloc(&treeCost, sizeof(double) * 2));
    capacity = cap;
    return true;
}

// Steps 2-6 for primitives whose boxes are already in primLo/primHi (bounds reduced into `bounds`).
bool AccelBuilder::buildFromBoxes(cudaStream_t st, uint32_t n, const BuildTarget& tgt, uint32_t* outNodeCount, uint32_t* outPrimCount) {
    const int TB = 256;
    const uint32_t grid = (n + TB - 1) / TB;
    uint32_t bitsPerAxis = ceilLog2(n < 2 ? 2 : n) / 3 + 7;
    if (bitsPerAxis > 21) bitsPerAxis = 21;
    const uint32_t keyBits = bitsPerAxis * 3;
    k_morton<<<grid, TB, 0, st>>>(primLo, primHi, n, bounds, bitsPerAxis, keys[0], vals[0]);
    const uint32_t tiles = (n + RS_TILE - 1) / RS_TILE;
    int cur = 0;
    for (uint32_t shift = 0; shift < keyBits; shift += 8) {
        k_rs_histogram<<<tiles, RS_THREADS, 0, st>>>(keys[cur], n, shift, tiles, tileHist);
        k_rs_scan_tiles<<<256, 1024, 0, st>>>(tileHist, tiles, digitTotals);
        k_rs_scan_digits<<<1, 256, 0, st>>>(digitTotals);
        k_rs_scatter<<<tiles, RS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[1 - cur], vals[1 - cur], n, shift, tiles, tileHist, digitTotals);
        cur = 1 - cur;
    }
    sortedVals = vals[cur];
    k_leaves<<<grid, TB, 0, st>>>(primLo, primHi, vals[cur], n, nodeLo, nodeHi, subFirst, subCount, parent);
    treeLo = nodeLo;
    treeHi = nodeHi;
    lastBuilder = 0;
    if (n > 1) {
        const bool wantLbvh = mode != BUILD_PLOC, wantPloc = mode != BUILD_LBVH;
        if (wantPloc) {
            // PLOC works on its own copy of the leaves (nodeLoB / nodeHiB) so that both hierarchies exist side by side
            CK(cudaMemcpyAsync(nodeLoB + (n - 1), nodeLo + (n - 1), sizeof(::float4) * n, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(nodeHiB + (n - 1), nodeHi + (n - 1), sizeof(::float4) * n, cudaMemcpyDeviceToDevice, st));
        }
        if (wantLbvh) {
            k_hierarchy<<<grid, TB, 0, st>>>(keys[cur], n, nodeLo, nodeHi, parent, subFirst, subCount);
            CK(cudaMemsetAsync(arrival, 0, sizeof(uint32_t) * n, st));
            k_refit<<<grid, TB, 0, st>>>(n, nodeLo, nodeHi, parent, arrival);
        }
        if (wantPloc && !buildPloc(st, n)) return false;
        if (wantLbvh && wantPloc) {
            // surface-area cost of the two BINARY trees (sum of internal half-areas), kept for the build statistics
            CK(cudaMemsetAsync(treeCost, 0, sizeof(double) * 2, st));
            const int cg = (int)std::min<uint32_t>((n + 255u) / 256u, 2048u);
            k_tree_cost<<<cg, 256, 0, st>>>(nodeLo, nodeHi, n - 1, treeCost);
            k_tree_cost<<<cg, 256, 0, st>>>(nodeLoB, nodeHiB, n - 1, treeCost + 1);
            double host[2];
            CK(cudaMemcpyAsync(host, treeCost, sizeof(host), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            lastCost[0] = host[0];
            lastCost[1] = host[1];
            // PREFER_FAST_TRACE: PLOC is kept when it lowers the surface-area cost of the binary tree by more than 20 %. Measured on B200
            // (profiles/r02_notes.md): flat cornell 0.47 of the radix tree's cost -> 9.3 to 6.2 node visits per ray; 1000 instances
            // 0.65; the C4 TLAS 0.53; caustics 0.13. Near parity the radix tree is the better one although its cost is a few percent
            // higher (bunny 0.93, suzanne 0.90, 10 M-triangle soup 0.94 -> 20.9 against PLOC's 24.0 visits): its children sit in the
            // octants of their parent, which is what the octant-ordered front-to-back traversal of the 8-wide nodes assumes, and the
            // surface-area cost -- of the binary tree or of the collapsed wide tree, both were tried -- does not see traversal order.
            if (host[1] < 0.8 * host[0]) { treeLo = nodeLoB; treeHi = nodeHiB; lastBuilder = 1; }
            if (getenv("VKRT_BUILD_VERBOSE"))
                fprintf(stderr, "[vkrt build] n=%u surface-area cost: radix tree %.6g, PLOC %.6g (ratio %.3f) -> %s\n", n, host[0], host[1],
                        host[0] > 0.0 ? host[1] / host[0] : 0.0, lastBuilder ? "PLOC" : "radix tree");
        } else if (wantPloc) {
            treeLo = nodeLoB; treeHi = nodeHiB; lastBuilder = 1;
        }
    }
    // collapse
    CollapseParams P = {};
    P.nodeLo = treeLo; P.nodeHi = treeHi; P.subFirst = subFirst; P.subCount = subCount; P.sortedVals = vals[cur];
    P.primCount = n;
    // One primitive per leaf slot. A watertight triangle test costs ~4 child-box tests and runs at half their lane efficiency, so a
    // slot box that culls a single triangle pays for itself: 3 -> 1 triangles per slot took the 10 M-triangle soup from 13.4 to 4.5
    // triangle tests per ray for 1.3 more node visits (762 -> 1020 Mrays/s) and the cornell box from 3.1 to 2.3 (profiles/r01_notes.md).
    P.maxLeaf = 1u;
    P.nodesOut = tgt.nodesOut; P.nodeBase = tgt.nodeBase; P.counters = counters;
    P.vertices = tgt.vertices; P.indices = tgt.indices; P.vertexBase = tgt.vertexBase; P.indexBase = tgt.indexBase;
    P.trianglesOut = tgt.trianglesOut; P.primBase = tgt.primBase;
    P.instanceRecords = tgt.instanceRecords; P.instancesOut = tgt.instancesOut;
    P.flatIn = tgt.flatIn; P.flatInstances = tgt.flatInstances; P.flatOut = tgt.flatOut;
    uint32_t hostCounters[4] = {0u, 0u, 0u, 0u};
    auto runCollapse = [&](CollapseParams cp, uint32_t* levels) -> bool {
        uint32_t initCounters[4] = {0u, 1u, 0u, 0u};  // node 0 (root) is pre-allocated
        CK(cudaMemcpyAsync(counters, initCounters, sizeof(initCounters), cudaMemcpyHostToDevice, st));
        uint2 rootWork = make_uint2(0u, 0u);  // bvh2 node 0 (for n == 1 that is the single leaf) -> local bvh8 node 0
        CK(cudaMemcpyAsync(work[0], &rootWork, sizeof(uint2), cudaMemcpyHostToDevice, st));
        uint32_t workCount = 1;
        int wq = 0;
        *levels = 0;
        while (workCount > 0) {
            (*levels)++;
            k_collapse<<<(workCount + 63) / 64, 64, 0, st>>>(cp, work[wq], workCount, work[1 - wq]);
            CK(cudaMemcpyAsync(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            workCount = hostCounters[0];
            uint32_t zero = 0;
            CK(cudaMemcpyAsync(counters, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            wq = 1 - wq;
        }
        return true;
    };
    if (!runCollapse(P, &lastLevels)) return false;
    *outNodeCount = hostCounters[1];
    *outPrimCount = hostCounters[2];
    CK(cudaGetLastError());
    return true;
}

// PLOC over the Morton-sorted leaves in nodeLoB / nodeHiB[(n-1)+k]. Scratch: plocClusters (ping-pong cluster lists), plocNN (nearest
// neighbours), `tileHist` (per-block counts and offsets; free once the sort is done).
bool AccelBuilder::buildPloc(cudaStream_t st, uint32_t n) {
    uint32_t* cl[2] = {plocClusters, plocClusters + capacity};
    uint2* blockCounts = reinterpret_cast<uint2*>(tileHist);
    const uint32_t maxBlocks = (n + PLOC_THREADS - 1) / PLOC_THREADS;
    uint2* blockOffsets = blockCounts + maxBlocks + 1;
    PlocState* state = reinterpret_cast<PlocState*>(plocState);
    PlocState init = {};
    init.count = n;
    CK(cudaMemcpyAsync(state, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_ploc_init<<<(n + 255) / 256, 256, 0, st>>>(cl[0], n);
    uint32_t upper = n;       // host-side upper bound of the cluster count (refreshed every few iterations)
    int cur = 0;
    uint32_t iterations = 0;
    const uint32_t maxIterations = 64u + 8u * ceilLog2(n);
    while (upper > (uint32_t)PLOC_TAIL) {
        const uint32_t force = iterations >= maxIterations ? 1u : 0u;   // pathological input: finish by pairing neighbours (halves the count)
        for (int k = 0; k < 4 && upper > (uint32_t)PLOC_TAIL; k++) {
            const uint32_t blocks = (upper + PLOC_THREADS - 1) / PLOC_THREADS;
            k_ploc_search<<<blocks, PLOC_THREADS, 0, st>>>(state, cl[cur], nodeLoB, nodeHiB, plocNN, blockCounts, force);
            k_ploc_scan<<<1, 1024, 0, st>>>(state, blockCounts, blockOffsets);
            k_ploc_apply<<<blocks, PLOC_THREADS, 0, st>>>(state, cl[cur], cl[1 - cur], plocNN, blockOffsets, nodeLoB, nodeHiB, subCount, n);
            cur = 1 - cur;
            iterations++;
        }
        PlocState host;
        CK(cudaMemcpyAsync(&host, state, sizeof(host), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        upper = host.count;
    }
    k_ploc_tail<<<1, PLOC_TAIL, 0, st>>>(state, cl[cur], nodeLoB, nodeHiB, subCount, n);
    CK(cudaGetLastError());
    return true;
}

bool AccelBuilder::buildBlas(cudaStream_t st, const ShaderVertex* vertices, const uint32_t* indices, uint32_t vertexBase, uint32_t indexBase,
                             uint32_t triCount, Bvh8Node* nodesOut, uint32_t nodeBase, ::float4* trianglesOut, uint32_t primBase,
                             ::float4* blasBoundsOut, uint32_t* outNodeCount, uint32_t* outPrimCount) {
    if (!reserve(triCount)) return false;
    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_triangle_bounds<<<boundsGrid(triCount), 256, 0, st>>>(vertices, indices, vertexBase, indexBase, triCount, primLo, primHi, bounds);
    BuildTarget tgt = {};
    tgt.nodesOut = nodesOut; tgt.nodeBase = nodeBase; tgt.vertices = vertices; tgt.indices = indices;
    tgt.vertexBase = vertexBase; tgt.indexBase = indexBase; tgt.trianglesOut = trianglesOut; tgt.primBase = primBase;
    if (!buildFromBoxes(st, triCount, tgt, outNodeCount, outPrimCount)) return false;
    // BLAS root box (LBVH node 0 after refit; for a single triangle node 0 is that leaf)
    CK(cudaMemcpyAsync(blasBoundsOut, treeLo, sizeof(::float4), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(blasBoundsOut + 1, treeHi, sizeof(::float4), cudaMemcpyDeviceToDevice, st));
    return true;
}

bool AccelBuilder::buildTlas(cudaStream_t st, const ::float4* blasBounds, const uint32_t* instanceBlas, const float* world3x4,
                             const InstanceRecord* records, uint32_t instanceCount, Bvh8Node* nodesOut, uint32_t nodeBase,
                             InstanceRecord* instancesOut, uint32_t* outNodeCount, uint32_t* outPrimCount) {
    if (!reserve(instanceCount)) return false;
    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_instance_bounds<<<(instanceCount + 255) / 256, 256, 0, st>>>(blasBounds, instanceBlas, world3x4, instanceCount, primLo, primHi, bounds);
    BuildTarget tgt = {};
    tgt.nodesOut = nodesOut; tgt.nodeBase = nodeBase; tgt.instanceRecords = records; tgt.instancesOut = instancesOut;
    return buildFromBoxes(st, instanceCount, tgt, outNodeCount, outPrimCount);
}

// Single-level BVH over every instanced triangle in world space (used when instancing does not pay: few or heavily overlapping
// instances). flatScratch must hold totalPrims uint2 entries.
bool AccelBuilder::buildFlat(cudaStream_t st, const ShaderVertex* vertices, const uint32_t* indices, const uint4* flatInstances,
                             const uint32_t* triOffsets, uint32_t instanceCount, const float* world3x4, uint32_t totalPrims, uint2* flatScratch,
                             Bvh8Node* nodesOut, uint32_t nodeBase, uint2* flatOut, uint32_t* outNodeCount, uint32_t* outPrimCount) {
    if (!reserve(totalPrims)) return false;
    const uint32_t grid = (totalPrims + 255) / 256;
    k_flat_enumerate<<<grid, 256, 0, st>>>(triOffsets, instanceCount, totalPrims, flatScratch);
    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_flat_bounds<<<boundsGrid(totalPrims), 256, 0, st>>>(vertices, indices, flatScratch, flatInstances, world3x4, totalPrims, primLo, primHi, bounds);
    BuildTarget tgt = {};
    tgt.nodesOut = nodesOut; tgt.nodeBase = nodeBase;
    tgt.flatIn = flatScratch; tgt.flatInstances = flatInstances; tgt.flatOut = flatOut;
    return buildFromBoxes(st, totalPrims, tgt, outNodeCount, outPrimCount);
}

} // namespace vk
