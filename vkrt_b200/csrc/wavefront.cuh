// wavefront.cuh — data layout of the wavefront path tracer (all SoA, 16-byte vectors, resident in HBM).
//
// The reference runs one megakernel invocation per pixel (src/shaders/entry/path/raygen_*.slang: for spp { for depth {...} }).
// Here the same per-sample program is cut at every TraceRay into stages that exchange paths through queues:
//   raygen  -> [trace: extension rays of depth d + shadow rays of depth d-1] -> shade(d) -> ... -> film
// A path is identified by its sample-record index  rec = sampleInChunk * localPixelCount + localPixel ; radiance and the
// denoiser features of that sample accumulate in the record, and the film kernel reduces the records of a pixel in sample
// order, which reproduces the reference's summation order  frameState.radiance += sampleState.radiance  exactly.
#pragma once
#include "scene_view.cuh"
#include "trace.cuh"

namespace vk {

// path flags (state.slang:22-26 + medium.slang:4-5 folded into one word)
enum : uint32_t {
    PF_PREV_VERTEX_NEE_ALLOWED = 1u << 0,
    PF_MEDIUM_REFRACTIVE = 1u << 1,
    PF_MEDIUM_ABSORPTION = 1u << 2,
    PF_FEATURES_RESOLVED = 1u << 3,
    PF_HERO_ACTIVE = 1u << 4
};

struct TileMap {
    uint32_t width, height;          // full image
    uint32_t tileW, tileH, tilesX, tilesY;
    uint32_t localTileCount;         // tiles owned by this rank
    uint32_t localPixelCount;        // localTileCount * tileW * tileH (edge tiles are padded, padded pixels are invalid)
    const uint32_t* localToGlobalTile;  // [localTileCount]
};

__host__ __device__ inline bool localPixelToGlobal(const TileMap& tm, const uint32_t* localToGlobalTile, uint32_t lp, uint32_t& gx, uint32_t& gy) {
    const uint32_t tilePixels = tm.tileW * tm.tileH;
    const uint32_t lt = lp / tilePixels, in = lp % tilePixels;
    const uint32_t gt = localToGlobalTile[lt];
    gx = (gt % tm.tilesX) * tm.tileW + in % tm.tileW;
    gy = (gt / tm.tilesX) * tm.tileH + in / tm.tileW;
    return gx < tm.width && gy < tm.height;
}

struct PathState {  // one set per ping-pong side
    ::float4* rayO;
    ::float4* rayD;
    ::uint4* meta;        // {sample-record index, rng, path flags, hit v (float bits, written by k_trace for this slot)}: the four
                          // words a vertex needs from its slot in ONE 16-byte gather (shading runs in material-sorted order, so every
                          // separate array costs a 32-byte sector per vertex)
    ::float4* thr;        // RGB: throughput.xyz, prevBsdfPdf | single: throughput, lambda, prevBsdfPdf, - | hero: throughput4
    ::float4* sigma;      // medium absorption sigma (rgb or 4 wavelengths); valid iff PF_MEDIUM_ABSORPTION
    ::float4* techPdf;    // hero: techniquePathPdf
    ::float4* prevVertexTechPdf;
    ::float4* prevBsdfTechPdf;
    ::float4* heroMisc;   // hero: unit wavelength sample, scalar throughput, prevBsdfPdf, -
};

struct SampleRecords {
    ::float4* radiance;       // RGB radiance | single: x | hero: 4 wavelength lanes
    float* radianceScalar;    // hero: single-wavelength lane after dispersive collapse
    float* unitWavelength;    // spectral: the sample's wavelength rotation in [0,1)
    ::float4* featA;          // denoiser albedo.xyz, weight
    ::float4* featB;          // denoiser normal.xyz, depth
    float* follow;            // followSpecular (depth 0)
};

struct Film {  // tile-compact, indexed by local pixel
    ::float4* accum[2];       // RGBA32F; w = sample count (scene/resources.slang:12-15)
    ::uint2* albedo[2];       // RGBA16F
    ::uint2* normal[2];       // RGBA16F
    ::uint2* output;          // RGBA16 UNORM
    ::float4* frameRadiance;  // per-frame partial sums across sample chunks
    ::float4* frameFeatA;
    ::float4* frameFeatB;
    float* frameFollow;
    ::float4* debugColor;     // primary-surface debug override (w = 1 when set)
    uint32_t* bounceCount;    // sample 0's bounce count (bounce-count debug view)
    ::uint2* hitId;           // primary visibility AOV
    ::float4* hitTuv;
};

struct FrameParams {
    SceneView scene;
    SceneData sd;
    TileMap tiles;
    PathState st[2];
    ::uint4* hitA;           // instance, primitive, t, u of the slot's extension ray (v lives in PathState::meta.w)
    // shadow queue
    ::float4* shO;
    ::float4* shD;
    ::float4* shContribution;
    ::uint2* shTarget;
    uint32_t* shSeed;
    // per-depth counters, zeroed at chunk start: extCount[d], shCount[d], traceWork[d]
    uint32_t* extCount;
    uint32_t* shCount;
    uint32_t* traceWork;
    // material-sorted shading order (north_star: "sorted-by-material shading kernels"): order[j] = queue slot, grouped by
    // key = 0 for a miss, 1 + materialIndex (clamped to 255) for a hit; nullptr = shade in queue order
    uint32_t* shadeOrder;
    uint32_t* sortBins;       // [0..255] counts -> starts, [256..511] cursors
    SampleRecords rec;
    Film film;
    int readIndex;            // accumulation read image
    uint32_t chunkFirstSample;  // first sample index of this chunk within the frame
    uint32_t chunkSamples;
    uint32_t capacity;        // path slots
    uint32_t modeFlags;       // RaygenModeState.flags (state.slang:28-58)
};

enum : uint32_t {
    MODE_BSDF_ONLY = 1u << 0, MODE_NEE_ONLY = 1u << 1, MODE_BOUNCE_COUNT = 1u << 2, MODE_DN_ALBEDO = 1u << 3,
    MODE_DN_NORMAL = 1u << 4, MODE_DN_VALIDITY = 1u << 5, MODE_DN_DEPTH = 1u << 6, MODE_DN_FOLLOW = 1u << 7,
    MODE_NEE_ENABLED = 1u << 8,
    MODE_SCENE_TRANSMISSIVE = 1u << 9   // some instance has material.transmission > 0: shadow rays can report "unsupported transmission"
};

} // namespace vk
