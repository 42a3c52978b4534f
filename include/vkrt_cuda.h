/*
 * vkrt_cuda.h — the drop-in C ABI for vkrt's path-tracing hot path on B200 (sm_100a).
 *
 * This is the boundary SURVEY.md §8(b) names: everything the reference does between
 * "host scene data is ready" and "accumulation image can be read back" — i.e. the Vulkan
 * acceleration-structure code (src/core/render/accel/{blas,tlas}.c), the command recording
 * (src/core/runtime/command/record.c:448-486,577-599) and the Slang shaders
 * (everything under src/shaders) — is replaced by the functions below.  Plain pointers and sizes only;
 * every upload copies, every read-back fills a caller buffer; the API is single-threaded
 * and non-reentrant, exactly like the reference's VKRT_* API (src/core/api/vkrt.h).
 *
 * Error convention = the reference's VKRT_Result (src/core/api/vkrt_types.h:16-26).
 *
 * Each entry point cites the reference interface it replaces.  The binding a vkrt
 * maintainer would add is shown in INTEGRATION.md.
 */
#ifndef VKRT_CUDA_H
#define VKRT_CUDA_H

#include "vkrt_shared.h"
#include "vkrt_closure.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKRT_CUDA_API __declspec(dllexport)
#else
#define VKRT_CUDA_API __attribute__((visibility("default")))
#endif

/* reference: src/core/api/vkrt_types.h:16-26 */
typedef int32_t VKRT_Result;
enum {
    VKRT_SUCCESS = 0,
    VKRT_ERROR_INVALID_ARGUMENT = -1,
    VKRT_ERROR_OPERATION_FAILED = -2,
    VKRT_ERROR_OUT_OF_MEMORY = -3,
    VKRT_ERROR_DEVICE_LOST = -4,
    VKRT_ERROR_INITIALIZATION_FAILED = -5
};

typedef struct vkrt_cuda_ctx vkrt_cuda_ctx;

/* Replaces createInstanceAndDevice (src/core/api/lifecycle.c:416): picks the CUDA device and,
 * for multi-GPU runs, this process's share of the image (interleaved tiles). */
typedef struct vkrt_cuda_create_info {
    int32_t device;        /* CUDA ordinal; -1 = current device */
    uint32_t rank;         /* this process's rank in [0, worldSize) */
    uint32_t worldSize;    /* 1 = whole image on this GPU */
    uint32_t tileWidth;    /* interleaved tile size in pixels; 0 = default 32 */
    uint32_t tileHeight;   /* 0 = default 32 */
    uint32_t maxPathsInFlight; /* wavefront pool capacity; 0 = default */
    uint32_t flags;        /* VKRT_CUDA_FLAG_* */
    uint32_t reserved;
} vkrt_cuda_create_info;

enum {
    VKRT_CUDA_FLAG_NONE = 0u,
    VKRT_CUDA_FLAG_COUNT_RAYS = 1u << 0,    /* per-frame ray/node/triangle counters (instrumented build of the same kernels) */
    VKRT_CUDA_FLAG_NO_MATERIAL_SORT = 1u << 1, /* shade in queue order instead of material-sorted order */
    VKRT_CUDA_FLAG_STAGE_TIMING = 1u << 2,    /* vkrt_cuda_render_frame records a CUDA event after every launch and fills traceMs / shadeMs */
    VKRT_CUDA_FLAG_FORCE_TWO_LEVEL = 1u << 3, /* always build BLAS per unique geometry + TLAS (default: chosen from the instancing ratio) */
    VKRT_CUDA_FLAG_FORCE_FLAT = 1u << 4,      /* always build one BVH over all instanced triangles */
    VKRT_CUDA_FLAG_DEEP_STACK = 1u << 5,      /* always traverse with the deep-tree kernel (default: chosen from the depth of the built trees) */
    VKRT_CUDA_FLAG_LBVH = 1u << 7,            /* build every hierarchy as a Karras radix tree over the Morton codes only (fastest build). Default: also
                                                 build the PLOC tree (bottom-up clustering by surface area over the same Morton order) and keep it, per
                                                 BVH, when it lowers the surface-area cost by more than 20 %: the counterpart of the
                                                 PREFER_FAST_TRACE the reference asks its driver for (src/core/render/accel/blas.c:39, tlas.c:188) */
    VKRT_CUDA_FLAG_PLOC = 1u << 8,            /* PLOC only (A/B measurements) */
    VKRT_CUDA_FLAG_ENV_IMPORTANCE = 1u << 6   /* EXTENSION (not in the reference, which reads the environment map on a miss only:
                                                 src/shaders/light/environment.slang:16-27): next-event estimation also samples the
                                                 lat-long environment texture by luminance x sin(theta), MIS-combined with BSDF sampling.
                                                 Same expectation, different samples: off by default so that results match the reference's. */
};

typedef struct vkrt_cuda_build_stats {
    float buildMs;            /* device time of BLAS + TLAS build (CUDA events) */
    float blasMs;
    float tlasMs;
    uint32_t uniqueGeometries; /* BLAS count */
    uint32_t instanceCount;
    uint64_t triangleCount;    /* unique (BLAS) triangles */
    uint64_t instancedTriangleCount;
    uint64_t bvh8NodeCount;
    uint64_t accelBytes;       /* nodes + repacked triangles + instance records */
    uint32_t flat;             /* 1 = single-level BVH over instanced triangles was built, 0 = BLAS per geometry + TLAS */
    uint32_t plocHierarchies;  /* how many of the built BVHs kept the PLOC hierarchy (the others the radix tree; VKRT_CUDA_FLAG_LBVH) */
} vkrt_cuda_build_stats;

typedef struct vkrt_cuda_frame_stats {
    float frameMs;       /* device time of the whole frame (CUDA events on the render stream) */
    float traceMs;       /* device time inside traversal kernels (sum over bounces); needs VKRT_CUDA_FLAG_STAGE_TIMING */
    float shadeMs;       /* device time inside raygen + shading + film kernels; needs VKRT_CUDA_FLAG_STAGE_TIMING */
    uint32_t kernelLaunches;
    uint32_t traceLaunches; /* traversal launches among kernelLaunches */
    uint32_t shadeLaunches; /* shading-kernel launches proper among kernelLaunches (one per depth and sample chunk) */
    uint64_t paths;      /* camera paths started = local pixels * spp */
    uint64_t extensionRays; /* closest-hit rays traced */
    uint64_t shadowRays;    /* any-hit rays traced */
    uint64_t nodesVisited;  /* only with VKRT_CUDA_FLAG_COUNT_RAYS */
    uint64_t trianglesTested;
    uint64_t instancesEntered;
    float shadeKernelMs;    /* device time inside the shading kernels alone (part of shadeMs); needs VKRT_CUDA_FLAG_STAGE_TIMING */
    uint32_t reserved;
} vkrt_cuda_frame_stats;

/* reference: VKRT_TextureUpload, src/core/api/vkrt_types.h:70-77 */
typedef struct vkrt_cuda_texture {
    const void* pixels;
    uint32_t width;
    uint32_t height;
    uint32_t format;     /* VKRT_TEXTURE_FORMAT_* */
    uint32_t colorSpace; /* VKRT_TEXTURE_COLOR_SPACE_* (sRGB decode only for RGBA8) */
} vkrt_cuda_texture;

typedef enum vkrt_cuda_aov {
    VKRT_CUDA_AOV_ACCUM_RGBA32F = 0, /* accumulationWriteImage: xyz = mean radiance (RGB, or XYZ in spectral), w = sample count */
    VKRT_CUDA_AOV_ALBEDO_RGBA16F = 1,
    VKRT_CUDA_AOV_NORMAL_RGBA16F = 2,
    VKRT_CUDA_AOV_OUTPUT_RGBA16 = 3, /* tone-mapped sRGB, 16-bit UNORM */
    VKRT_CUDA_AOV_HITID_CENTER = 4,  /* uint32 x2 per pixel {instance, primitive}; jitter = 0 (debug.slang:21-29) */
    VKRT_CUDA_AOV_HITID_S0 = 5,      /* uint32 x2 per pixel; frame 0 sample 0, jittered (integrator.slang:25-27) */
    VKRT_CUDA_AOV_HIT_T_UV_CENTER = 6 /* float x3 per pixel {t, u, v} of the HITID_CENTER ray */
} vkrt_cuda_aov;

/* lifecycle — replaces VKRT_create/VKRT_destroy's device half (src/core/api/lifecycle.c:295-309,416-545) */
VKRT_CUDA_API VKRT_Result vkrt_cuda_create(const vkrt_cuda_create_info* info, vkrt_cuda_ctx** outCtx);
VKRT_CUDA_API void vkrt_cuda_destroy(vkrt_cuda_ctx* ctx);
VKRT_CUDA_API const char* vkrt_cuda_last_error(const vkrt_cuda_ctx* ctx);
VKRT_CUDA_API const char* vkrt_cuda_version(void);

/* Global vertex/index buffers with per-mesh bases (bindings 10/11, src/shaders/scene/resources.slang:29-32;
 * host side src/core/scene/geometry.c:644-806). */
VKRT_CUDA_API VKRT_Result vkrt_cuda_set_geometry(vkrt_cuda_ctx* ctx, const ShaderVertex* vertices, uint32_t vertexCount,
                                                 const uint32_t* indices, uint32_t indexCount);

/* One instance per mesh (src/core/render/accel/tlas.c:324-346): MeshInfo (binding 14), row-major 3x4
 * world transform (getMeshWorldTransform, src/core/scene/transform.c:212-223), the dedup id that decides
 * BLAS sharing (NULL = derive from vertexBase/indexBase/indexCount), and the per-instance
 * FORCE_NO_OPAQUE decision (materialMayRejectRayHit, tlas.c:291-296; NULL = all opaque). */
VKRT_CUDA_API VKRT_Result vkrt_cuda_set_instances(vkrt_cuda_ctx* ctx, const MeshInfo* infos, const float* world3x4,
                                                  const uint32_t* geometrySource, const uint8_t* alphaTested,
                                                  uint32_t instanceCount);

/* binding 15 (materials); rebuilt by src/core/scene/rebuild.c:88-155 */
VKRT_CUDA_API VKRT_Result vkrt_cuda_set_materials(vkrt_cuda_ctx* ctx, const Material* materials, uint32_t materialCount);

/* bindings 16-21; built by vkrtSceneRebuildLightBuffers (src/core/scene/lighting.c:496-544) */
VKRT_CUDA_API VKRT_Result vkrt_cuda_set_lights(vkrt_cuda_ctx* ctx, const EmissiveMesh* meshes, uint32_t meshCount,
                                               const EmissiveTriangle* triangles, uint32_t triangleCount,
                                               const float* meshAliasQ, const uint32_t* meshAliasIdx,
                                               const float* triAliasQ, const uint32_t* triAliasIdx);

/* bindings 22/23 (bindless textures + 9 sampler variants; src/core/scene/textures.c:141-209) */
VKRT_CUDA_API VKRT_Result vkrt_cuda_set_textures(vkrt_cuda_ctx* ctx, const vkrt_cuda_texture* textures, uint32_t textureCount);

/* binding 24; payload = scale[res] ++ coeff[3*res^3*3] (src/core/scene/rgb2spec.c:17-59) */
VKRT_CUDA_API VKRT_Result vkrt_cuda_set_rgb2spec(vkrt_cuda_ctx* ctx, const float* payload, uint32_t floatCount,
                                                 RGB2SpecTableInfo info);

/* Replaces recordBottomLevelAccelerationStructureBuilds (blas.c:222-262) + recordTopLevelAccelerationStructureBuilds
 * (tlas.c:535-559): Morton sort, two binary hierarchies over the sorted primitives (Karras radix tree and PLOC, the one with the lower surface-area
 * cost is kept: the counterpart of the PREFER_FAST_TRACE flag the reference passes, blas.c:39 / tlas.c:188), collapse to compressed 8-wide nodes;
 * ONE flat BVH over all instanced triangles when flattening does not multiply memory, otherwise one BLAS per unique geometry + a TLAS over instances. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_build_accel(vkrt_cuda_ctx* ctx, vkrt_cuda_build_stats* outStats);
/* vkrt_cuda_build_accel returns the previous build when nothing it depends on changed (geometry, instance matrices / sharing / any-hit
 * flags, the "transmits" bit of a material): like the reference, which rebuilds a BLAS only when blasBuildPending is set
 * (src/core/render/accel/blas.c:222-262). This forces the next call to build (build-time measurements). */
VKRT_CUDA_API VKRT_Result vkrt_cuda_invalidate_accel(vkrt_cuda_ctx* ctx);

/* Replaces createGPUImages (src/core/runtime/images.c:260-320): (re)allocates film images for the FULL image size;
 * this rank stores only its own tiles. Resets accumulation. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_resize(vkrt_cuda_ctx* ctx, uint32_t width, uint32_t height);

/* record.c:580-585 (accumulationNeedsReset) */
VKRT_CUDA_API VKRT_Result vkrt_cuda_reset_accumulation(vkrt_cuda_ctx* ctx);

/* Replaces recordMainTracePass (record.c:448-486): one vkCmdTraceRaysKHR(W,H,1) worth of work =
 * sceneData->samplesPerPixel samples for every pixel, accumulated into the film exactly as
 * writeback.slang:87-114, followed by the accumulation read/write swap of VKRT_endFrame (frame.c:386-388).
 * The caller owns frameNumber (frame.c:380-389 increments it per traced frame). Blocks until the frame is done. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_render_frame(vkrt_cuda_ctx* ctx, const SceneData* sceneData,
                                                 vkrt_cuda_frame_stats* outStats);

/* Same work, but only enqueued on the context's stream (no host sync, no stats); pair with vkrt_cuda_sync. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_render_frame_async(vkrt_cuda_ctx* ctx, const SceneData* sceneData);
VKRT_CUDA_API VKRT_Result vkrt_cuda_sync(vkrt_cuda_ctx* ctx);
/* Device-side stopwatch on the context's own stream (the stream every kernel of this library is launched on): begin records a CUDA
 * event, end records a second one, waits for it and returns the elapsed device time between the two. Replaces the reference's GPU
 * timestamp pair around the frame (record.c:274-280,795-800) for multi-frame measurements. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_timer_begin(vkrt_cuda_ctx* ctx);
VKRT_CUDA_API VKRT_Result vkrt_cuda_timer_end(vkrt_cuda_ctx* ctx, float* outMs);

/* Multi-GPU (no reference equivalent). The library talks to NCCL through dlopen("libnccl.so.2"),
 * so a single-GPU host needs no NCCL. uniqueId is the 128-byte ncclUniqueId obtained from
 * vkrt_cuda_nccl_unique_id on rank 0 and distributed by the caller. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_nccl_unique_id(void* outId128);
VKRT_CUDA_API VKRT_Result vkrt_cuda_comm_init(vkrt_cuda_ctx* ctx, const void* uniqueId128);
/* Gathers every rank's tile-compact film images to rank 0 and un-permutes them into full-frame images there. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_gather(vkrt_cuda_ctx* ctx, float* outGatherMs);
/* The same for a subset: bit k of aovMask = AOV k (bit 0 accumulation, 1 albedo, 2 normal, 3 display image). A progressive read-back
 * of the accumulation alone moves a third of the bytes; one NCCL group and one un-tiling launch per AOV, one host synchronisation per call. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_gather_aovs(vkrt_cuda_ctx* ctx, uint32_t aovMask, float* outGatherMs);
/* Device pointers of this rank's tile-compact images, for callers that run the collective themselves
 * (e.g. torch.distributed): bytes = localPixelCount * 16 (accum) / 8 (albedo, normal, output). */
VKRT_CUDA_API VKRT_Result vkrt_cuda_local_film(vkrt_cuda_ctx* ctx, vkrt_cuda_aov which, void** outDevicePtr,
                                               uint64_t* outBytes, uint64_t* outLocalPixelCount);
/* Rank 0: scatter `count` ranks' tile-compact buffers (concatenated, rank-major, each padded to
 * vkrt_cuda_max_local_pixels() pixels) into the full-frame image of `which`. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_import_gathered(vkrt_cuda_ctx* ctx, vkrt_cuda_aov which, const void* deviceGathered);
VKRT_CUDA_API uint64_t vkrt_cuda_max_local_pixels(const vkrt_cuda_ctx* ctx);

/* Replaces readbackImagePixels (src/core/utility/export/api.c:170-242). For worldSize == 1, or after
 * vkrt_cuda_gather / vkrt_cuda_import_gathered on rank 0, the result is the full W*H image in row-major pixel order.
 * HITID_* AOVs trace a primary-visibility pass with the camera of the last render_frame call (or of `sceneData`
 * passed to vkrt_cuda_trace_primary). bytes must equal W*H*pixelSize. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_read_aov(vkrt_cuda_ctx* ctx, vkrt_cuda_aov which, void* dst, size_t bytes);
VKRT_CUDA_API VKRT_Result vkrt_cuda_trace_primary(vkrt_cuda_ctx* ctx, const SceneData* sceneData);
/* Replaces the 256 one-texel vkCmdCopyImageToBuffer regions of the auto-exposure probe (src/core/scene/exposure.c:126-185):
 * out[4 i .. 4 i + 3] = accumulation RGBA (w = sample count) of pixel (xy[2 i], xy[2 i + 1]); pixels this rank does not own read 0. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_read_accum_samples(vkrt_cuda_ctx* ctx, const uint32_t* xy, uint32_t count, float* outRgba);

/* Standalone traversal entry (no reference equivalent; used by the traversal benchmark and parity tests):
 * rays = n * {ox,oy,oz,tmin, dx,dy,dz,tmax} floats on the HOST, hits = n * {instance, primitive, t, u, v} (uint32/float bits).
 * anyHit != 0 traces shadow-style (first accepted hit terminates; hits[i].instance = 1 occluded / 0 visible). */
VKRT_CUDA_API VKRT_Result vkrt_cuda_trace_rays(vkrt_cuda_ctx* ctx, const float* rays, uint32_t rayCount, int anyHit,
                                               uint32_t* hits, float* outKernelMs);

/* Test entry (no reference equivalent; vkrt_closure.h): evaluates and samples the layered closure for `count` independent queries on
 * the device, through the same device functions the shading kernel calls, so that every lobe of src/shaders/bsdf can be compared
 * call by call with the CPU oracle and with the reference's own shaders compiled for the CPU (tests/test_gpu_reference.py).
 * queries / results are HOST arrays. Spectral modes need vkrt_cuda_set_rgb2spec. */
VKRT_CUDA_API VKRT_Result vkrt_cuda_eval_closures(vkrt_cuda_ctx* ctx, const vkrt_closure_query* queries, uint32_t count,
                                                  vkrt_closure_result* results);

#ifdef __cplusplus
}
#endif

#endif /* VKRT_CUDA_H */
