// build.h — host-side interface of the GPU BVH builder (bvh_build.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/vkrt_shared.h"
#include "accel.cuh"

namespace vk {

struct BuildTarget {
    Bvh8Node* nodesOut = nullptr;
    uint32_t nodeBase = 0;
    const ShaderVertex* vertices = nullptr;
    const uint32_t* indices = nullptr;
    uint32_t vertexBase = 0, indexBase = 0;
    ::float4* trianglesOut = nullptr;
    uint32_t primBase = 0;
    const InstanceRecord* instanceRecords = nullptr;
    InstanceRecord* instancesOut = nullptr;
    const uint2* flatIn = nullptr;
    const uint4* flatInstances = nullptr;
    uint2* flatOut = nullptr;
};

void launchPackTriangles(const ShaderVertex* vertices, const uint32_t* indices, uint32_t vertexBase, uint32_t indexBase, uint32_t triCount,
                         ::float4* trianglesOut, uint32_t primBase, cudaStream_t st);
void launchGeometryBounds(const ShaderVertex* vertices, const uint32_t* indices, uint32_t vertexBase, uint32_t indexBase, uint32_t triCount, int* bounds6,
                          cudaStream_t st);
void launchValidateIndices(const uint32_t* indices, uint32_t indexBase, uint32_t indexCount, uint32_t vertexLimit, uint32_t* flag, cudaStream_t st);
float orderedIntToFloatHost(int i);
void launchFlatGatherTriangles(const uint2* flatPrims, const ::float4* src, uint32_t total, ::float4* dst, cudaStream_t st);
void launchRelocateNodes(const Bvh8Node* src, Bvh8Node* dst, uint32_t count, uint32_t oldBase, uint32_t newBase, cudaStream_t st);

// Scratch memory + launch sequence for one BVH build at a time (reused across BLASes and the TLAS).
class AccelBuilder {
public:
    ~AccelBuilder() { release(); }
    bool reserve(uint32_t primCount);
    void release();
    // nodesOut must have room for max(triCount, 1) nodes starting at nodeBase; trianglesOut for triCount triangles at primBase.
    bool buildBlas(cudaStream_t st, const ShaderVertex* vertices, const uint32_t* indices, uint32_t vertexBase, uint32_t indexBase,
                   uint32_t triCount, Bvh8Node* nodesOut, uint32_t nodeBase, ::float4* trianglesOut, uint32_t primBase,
                   ::float4* blasBoundsOut, uint32_t* outNodeCount, uint32_t* outPrimCount);
    bool buildTlas(cudaStream_t st, const ::float4* blasBounds, const uint32_t* instanceBlas, const float* world3x4,
                   const InstanceRecord* records, uint32_t instanceCount, Bvh8Node* nodesOut, uint32_t nodeBase,
                   InstanceRecord* instancesOut, uint32_t* outNodeCount, uint32_t* outPrimCount);
    bool buildFlat(cudaStream_t st, const ShaderVertex* vertices, const uint32_t* indices, const uint4* flatInstances, const uint32_t* triOffsets,
                   uint32_t instanceCount, const float* world3x4, uint32_t totalPrims, uint2* flatScratch, Bvh8Node* nodesOut, uint32_t nodeBase,
                   uint2* flatOut, uint32_t* outNodeCount, uint32_t* outPrimCount);
    char err[512] = {0};
    enum Mode { BUILD_BEST = 0, BUILD_LBVH = 1, BUILD_PLOC = 2 };
    // BEST: build the Karras radix tree AND the PLOC tree; PLOC is kept when its surface-area cost is below 0.8 x the radix tree's
    // (PREFER_FAST_TRACE; see buildFromBoxes);
    // LBVH: radix tree only (fastest build); PLOC: PLOC only (A/B measurements)
    Mode mode = BUILD_BEST;
    uint32_t lastBuilder = 0;           // hierarchy the most recent build kept: 0 = radix tree, 1 = PLOC
    double lastCost[2] = {0.0, 0.0};       // sum of internal-node half-areas of the two binary hierarchies (BEST mode)
    uint32_t lastLevels = 0;  // BVH8 levels of the most recent build (number of collapse rounds)

private:
    bool buildPloc(cudaStream_t st, uint32_t n);
    bool buildFromBoxes(cudaStream_t st, uint32_t n, const BuildTarget& tgt, uint32_t* outNodeCount, uint32_t* outPrimCount);
    uint32_t capacity = 0;
    ::float4 *primLo = nullptr, *primHi = nullptr;
    uint64_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t *tileHist = nullptr, *digitTotals = nullptr;
    ::float4 *nodeLo = nullptr, *nodeHi = nullptr;
    uint32_t *parent = nullptr, *arrival = nullptr, *subFirst = nullptr, *subCount = nullptr;
    uint2* work[2] = {nullptr, nullptr};
    uint32_t* counters = nullptr;
    int* bounds = nullptr;
    void* plocState = nullptr;
    ::float4 *nodeLoB = nullptr, *nodeHiB = nullptr;      // PLOC's hierarchy (the radix tree lives in nodeLo / nodeHi)
    const ::float4 *treeLo = nullptr, *treeHi = nullptr;  // the hierarchy the collapse reads
    uint32_t *plocClusters = nullptr, *plocNN = nullptr;
    double* treeCost = nullptr;
    const uint32_t* sortedVals = nullptr;
};

} // namespace vk
