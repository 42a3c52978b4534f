/*
 * vkrt_host.h — the C host API above the vkrt_cuda_* boundary (libvkrt_host.so).
 *
 * This mirrors the part of the reference's public API (src/core/api/vkrt.h, vkrt_types.h) that feeds and drives the
 * path-tracing hot path: same function names, argument meaning, clamping and error behaviour, so that vkrt's app layer
 * (src/app/{cli,render,scene,mesh,session}) links against it unchanged for headless rendering. Functions of the reference
 * API that only serve the window/editor (swapchain, selection outline, overlay, camera mouse input, render-view pan/zoom,
 * the viewport variant of the denoiser) are not part of the path and are not provided. The save-time denoise stage of
 * VKRT_saveRenderImageEx (feature AOVs -> Open Image Denoise, src/core/utility/export/image.c:840-905) is provided; the OIDN
 * library itself is bound at run time (VKRT_OIDN_LIBRARY or the usual sonames) and a missing library saves the raw image.
 *
 * What the implementation does differently underneath: VKRT_updateScene keeps the reference's dirty-revision logic
 * (src/core/api/frame.c:227-263) but calls vkrt_cuda_set_* / vkrt_cuda_build_accel instead of rebuilding Vulkan buffers and
 * acceleration structures; VKRT_trace calls vkrt_cuda_render_frame instead of recording vkCmdTraceRaysKHR
 * (src/core/runtime/command/record.c:448-486); VKRT_saveRenderImageEx reads the film back with vkrt_cuda_read_aov.
 */
#ifndef VKRT_HOST_H
#define VKRT_HOST_H

#include "vkrt_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKRT_HOST_API __declspec(dllexport)
#else
#define VKRT_HOST_API __attribute__((visibility("default")))
#endif

enum { VKRT_DEVICE_NAME_LEN = 256, VKRT_NAME_LEN = 256, VKRT_FRAMETIME_HISTORY_SIZE = 128 };

typedef float vkrt_vec3[3];
typedef float vkrt_mat4[4][4]; /* column-major m[col][row], like cglm's mat4 */

typedef uint32_t VKRT_ToneMappingMode;
typedef uint32_t VKRT_RenderMode;
typedef uint32_t VKRT_SpectralSamplingMode;
typedef uint32_t VKRT_DebugMode;

/* reference: vkrt_types.h:36-39 */
typedef struct Camera {
    vkrt_vec3 pos, target, up;
    float nearZ, farZ, vfov;
} Camera;

/* reference: vkrt_types.h:52-62. The window fields are accepted and ignored (B200 has no display stack); the device is chosen
 * with preferredDeviceIndex (CUDA ordinal, -1 = current). The last three fields are this implementation's multi-GPU knobs. */
typedef struct VKRT_CreateInfo {
    uint32_t width;
    uint32_t height;
    const char* title;
    uint8_t startMaximized;
    uint8_t startFullscreen;
    uint8_t headless;
    uint8_t disableSER;
    int32_t preferredDeviceIndex;
    const char* preferredDeviceName;
    uint32_t rank;              /* tile partition: this process's rank */
    uint32_t worldSize;         /* 0 or 1 = single GPU */
    uint32_t maxPathsInFlight;  /* 0 = default */
    uint32_t cudaFlags;         /* VKRT_CUDA_FLAG_* */
    uint8_t hostOnly;           /* 1 = no device: scene preparation only (CPU tests of the host logic); VKRT_trace fails */
} VKRT_CreateInfo;

typedef struct VKRT_MeshUpload {
    const Vertex* vertices;
    size_t vertexCount;
    const uint32_t* indices;
    size_t indexCount;
} VKRT_MeshUpload;

typedef struct VKRT_TextureUpload {
    const char* name;
    const void* pixels;
    uint32_t width;
    uint32_t height;
    uint32_t format;
    uint32_t colorSpace;
} VKRT_TextureUpload;

typedef struct VKRT VKRT;

typedef struct VKRT_TextureSnapshot { /* vkrt_types.h:266-273 */
    uint32_t width;
    uint32_t height;
    uint32_t format;
    uint32_t colorSpace;
    uint32_t useCount;
    char name[VKRT_NAME_LEN];
} VKRT_TextureSnapshot;

typedef struct VKRT_SceneSettingsSnapshot {
    Camera camera;
    uint32_t samplesPerPixel;
    uint32_t rrMaxDepth;
    uint32_t rrMinDepth;
    VKRT_ToneMappingMode toneMappingMode;
    VKRT_RenderMode renderMode;
    uint32_t spectralSamplingMode;
    float exposure;
    uint8_t autoExposureEnabled;
    uint8_t autoSPPEnabled;
    uint32_t autoSPPTargetFPS;
    vkrt_vec3 environmentColor;
    float environmentStrength;
    float environmentRotation;
    uint32_t environmentTextureIndex;
    float timeBase;
    float timeStep;
    uint32_t debugMode;
    uint32_t misNeeEnabled;
    uint32_t selectionEnabled;
    uint32_t selectedMeshIndex;
} VKRT_SceneSettingsSnapshot;

typedef enum VKRT_RenderPhase {
    VKRT_RENDER_PHASE_INACTIVE = 0,
    VKRT_RENDER_PHASE_SAMPLING,
    VKRT_RENDER_PHASE_DENOISING,
    VKRT_RENDER_PHASE_COMPLETE_RAW,
    VKRT_RENDER_PHASE_COMPLETE_DENOISED
} VKRT_RenderPhase;

typedef struct VKRT_RenderStatusSnapshot {
    uint32_t framesPerSecond;
    float averageFrametime;
    float frametimes[VKRT_FRAMETIME_HISTORY_SIZE];
    float displayTimeMs;
    float renderTimeMs; /* device time of the last traced frame (vkrt_cuda_frame_stats.frameMs) */
    uint32_t accumulationFrame;
    uint64_t totalSamples;
    VKRT_RenderPhase renderPhase;
    uint8_t renderDenoiseEnabled;
    uint32_t renderTargetSamples;
    float displayRenderTimeMs;
    float displayFrameTimeMs;
} VKRT_RenderStatusSnapshot;

typedef struct VKRT_RenderExportSettings {
    uint8_t denoiseEnabled; /* 1: denoise with the albedo / normal AOVs before saving (needs libOpenImageDenoise at run time; raw image otherwise) */
} VKRT_RenderExportSettings;

typedef struct VKRT_SystemInfo {
    char deviceName[VKRT_DEVICE_NAME_LEN];
    uint32_t vendorID;
    uint32_t driverVersion;
} VKRT_SystemInfo;

typedef struct VKRT_MeshSnapshot {
    MeshInfo info;
    Material material;
    uint32_t materialIndex;
    uint32_t geometrySource;
    uint8_t hasMaterialAssignment;
    uint8_t ownsGeometry;
    char name[VKRT_NAME_LEN];
} VKRT_MeshSnapshot;

typedef struct VKRT_MaterialSnapshot {
    Material material;
    uint32_t useCount;
    char name[VKRT_NAME_LEN];
} VKRT_MaterialSnapshot;

/* ---- lifecycle (src/core/api/lifecycle.c:295-309,547-609) ---- */
VKRT_HOST_API void VKRT_defaultCreateInfo(VKRT_CreateInfo* createInfo);
VKRT_HOST_API VKRT_Result VKRT_create(VKRT** outVkrt);
VKRT_HOST_API void VKRT_destroy(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_initWithCreateInfo(VKRT* vkrt, const VKRT_CreateInfo* createInfo);
VKRT_HOST_API VKRT_Result VKRT_init(VKRT* vkrt);
VKRT_HOST_API void VKRT_deinit(VKRT* vkrt);

/* ---- frame protocol (src/core/api/frame.c:74-420) ---- */
VKRT_HOST_API VKRT_Result VKRT_beginFrame(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_updateScene(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_trace(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_present(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_endFrame(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_draw(VKRT* vkrt);

/* ---- geometry (src/core/api/geometry.c:31, src/core/scene/geometry.c:166-210,729-806) ---- */
VKRT_HOST_API VKRT_Result VKRT_uploadMeshData(VKRT* vkrt, const Vertex* vertices, size_t vertexCount, const uint32_t* indices, size_t indexCount);
VKRT_HOST_API VKRT_Result VKRT_uploadMeshDataBatch(VKRT* vkrt, const VKRT_MeshUpload* uploads, size_t uploadCount);
VKRT_HOST_API VKRT_Result VKRT_removeMesh(VKRT* vkrt, uint32_t meshIndex);

/* ---- settings (src/core/api/settings.c:30-293; each setter clamps like the reference and restarts accumulation) ---- */
VKRT_HOST_API VKRT_Result VKRT_invalidateAccumulation(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_setSamplesPerPixel(VKRT* vkrt, uint32_t samplesPerPixel);
VKRT_HOST_API VKRT_Result VKRT_setPathDepth(VKRT* vkrt, uint32_t rrMinDepth, uint32_t rrMaxDepth);
VKRT_HOST_API VKRT_Result VKRT_setAutoSPPEnabled(VKRT* vkrt, uint8_t enabled);
VKRT_HOST_API VKRT_Result VKRT_setAutoSPPTargetFPS(VKRT* vkrt, uint32_t targetFPS);   /* vkrt.h:41; clamped to 30..360, 0 = 60 */
VKRT_HOST_API VKRT_Result VKRT_setAutoExposureEnabled(VKRT* vkrt, uint8_t enabled);  /* vkrt.h:46; single rank only */
VKRT_HOST_API VKRT_Result VKRT_setToneMappingMode(VKRT* vkrt, VKRT_ToneMappingMode toneMappingMode);
VKRT_HOST_API VKRT_Result VKRT_setRenderMode(VKRT* vkrt, VKRT_RenderMode renderMode);
VKRT_HOST_API VKRT_Result VKRT_setSpectralSamplingMode(VKRT* vkrt, VKRT_SpectralSamplingMode spectralSamplingMode);
VKRT_HOST_API VKRT_Result VKRT_setExposure(VKRT* vkrt, float exposure);
VKRT_HOST_API VKRT_Result VKRT_setEnvironmentLight(VKRT* vkrt, vkrt_vec3 color, float strength);
VKRT_HOST_API VKRT_Result VKRT_setEnvironmentRotation(VKRT* vkrt, float rotationDegrees);
VKRT_HOST_API VKRT_Result VKRT_setEnvironmentTextureFromPixels(VKRT* vkrt, const VKRT_TextureUpload* upload);
VKRT_HOST_API VKRT_Result VKRT_setEnvironmentTextureFromFile(VKRT* vkrt, const char* path); /* vkrt.h:49; PNG / JPEG / EXR, stored LINEAR */
VKRT_HOST_API VKRT_Result VKRT_clearEnvironmentTexture(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_setDebugMode(VKRT* vkrt, VKRT_DebugMode mode);
VKRT_HOST_API VKRT_Result VKRT_setMisNeeEnabled(VKRT* vkrt, uint8_t enabled);
VKRT_HOST_API VKRT_Result VKRT_setTimeRange(VKRT* vkrt, float timeBase, float timeStep);
VKRT_HOST_API VKRT_Result VKRT_setRenderViewport(VKRT* vkrt, uint32_t x, uint32_t y, uint32_t width, uint32_t height);
VKRT_HOST_API VKRT_Result VKRT_cameraSetPose(VKRT* vkrt, vkrt_vec3 position, vkrt_vec3 target, vkrt_vec3 upVector, float vfov);
VKRT_HOST_API VKRT_Result VKRT_cameraGetPose(const VKRT* vkrt, vkrt_vec3 position, vkrt_vec3 target, vkrt_vec3 upVector, float* vfov);
/* rgb2spec table: the reference embeds srgb.coeff at build time (src/core/scene/rgb2spec.c:17-89); here it is loaded at run time. */
VKRT_HOST_API VKRT_Result VKRT_loadRGB2SpecTable(VKRT* vkrt, const char* path);

/* ---- render session (src/core/api/render.c:82-310) ---- */
VKRT_HOST_API void VKRT_defaultRenderExportSettings(VKRT_RenderExportSettings* settings);
VKRT_HOST_API VKRT_Result VKRT_saveRenderImageEx(VKRT* vkrt, const char* path, const VKRT_RenderExportSettings* settings);
VKRT_HOST_API VKRT_Result VKRT_saveRenderImage(VKRT* vkrt, const char* path);
VKRT_HOST_API VKRT_Result VKRT_startRender(VKRT* vkrt, uint32_t width, uint32_t height, uint32_t targetSamples);
VKRT_HOST_API VKRT_Result VKRT_continueRender(VKRT* vkrt, uint32_t targetSamples);
VKRT_HOST_API VKRT_Result VKRT_stopRenderSampling(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_stopRender(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_getSceneSettings(const VKRT* vkrt, VKRT_SceneSettingsSnapshot* outSettings);
VKRT_HOST_API VKRT_Result VKRT_getRenderStatus(const VKRT* vkrt, VKRT_RenderStatusSnapshot* outStatus);
VKRT_HOST_API VKRT_Result VKRT_getSystemInfo(const VKRT* vkrt, VKRT_SystemInfo* outSystemInfo);

/* ---- meshes / materials / textures (src/core/api/mesh.c:414-664, texture.c) ---- */
VKRT_HOST_API VKRT_Result VKRT_getMeshCount(const VKRT* vkrt, uint32_t* outMeshCount);
VKRT_HOST_API VKRT_Result VKRT_getMeshSnapshot(const VKRT* vkrt, uint32_t meshIndex, VKRT_MeshSnapshot* outMesh);
VKRT_HOST_API VKRT_Result VKRT_getMaterialCount(const VKRT* vkrt, uint32_t* outMaterialCount);
VKRT_HOST_API VKRT_Result VKRT_getMaterialSnapshot(const VKRT* vkrt, uint32_t materialIndex, VKRT_MaterialSnapshot* outMaterial);
VKRT_HOST_API VKRT_Result VKRT_getTextureCount(const VKRT* vkrt, uint32_t* outTextureCount);
VKRT_HOST_API VKRT_Result VKRT_getTextureSnapshot(const VKRT* vkrt, uint32_t textureIndex, VKRT_TextureSnapshot* outTexture); /* vkrt.h:78 */
VKRT_HOST_API VKRT_Result VKRT_addTextureFromPixels(VKRT* vkrt, const VKRT_TextureUpload* upload, uint32_t* outTextureIndex);
VKRT_HOST_API VKRT_Result VKRT_addTextureFromFile(VKRT* vkrt, const char* path, const char* name, uint32_t colorSpace, uint32_t* outTextureIndex); /* vkrt.h:80-86 */
VKRT_HOST_API VKRT_Result VKRT_removeTexture(VKRT* vkrt, uint32_t textureIndex); /* vkrt.h:87 */
VKRT_HOST_API VKRT_Result VKRT_addTexturesBatch(VKRT* vkrt, const VKRT_TextureUpload* uploads, size_t uploadCount, uint32_t* outTextureIndices); /* vkrt.h:88-93 */
VKRT_HOST_API VKRT_Result VKRT_setMaterialTexture(VKRT* vkrt, uint32_t materialIndex, uint32_t textureSlot, uint32_t textureIndex);
VKRT_HOST_API VKRT_Result VKRT_addMaterial(VKRT* vkrt, const Material* material, const char* name, uint32_t* outMaterialIndex);
VKRT_HOST_API VKRT_Result VKRT_removeMaterial(VKRT* vkrt, uint32_t materialIndex); /* vkrt.h:96 */
VKRT_HOST_API VKRT_Result VKRT_setMaterialName(VKRT* vkrt, uint32_t materialIndex, const char* name);
VKRT_HOST_API VKRT_Result VKRT_setMaterial(VKRT* vkrt, uint32_t materialIndex, const Material* material);
VKRT_HOST_API VKRT_Result VKRT_setMeshMaterialIndex(VKRT* vkrt, uint32_t meshIndex, uint32_t materialIndex);
VKRT_HOST_API VKRT_Result VKRT_clearMeshMaterialAssignment(VKRT* vkrt, uint32_t meshIndex);
VKRT_HOST_API VKRT_Result VKRT_setMeshOpacity(VKRT* vkrt, uint32_t meshIndex, float opacity);
VKRT_HOST_API VKRT_Result VKRT_setMeshName(VKRT* vkrt, uint32_t meshIndex, const char* name);
VKRT_HOST_API VKRT_Result VKRT_setMeshTransform(VKRT* vkrt, uint32_t meshIndex, vkrt_vec3 position, vkrt_vec3 rotation, vkrt_vec3 scale);
VKRT_HOST_API VKRT_Result VKRT_setMeshTransformMatrix(VKRT* vkrt, uint32_t meshIndex, vkrt_mat4 worldTransform);
VKRT_HOST_API VKRT_Result VKRT_setMeshRenderBackfaces(VKRT* vkrt, uint32_t meshIndex, uint32_t enabled);
VKRT_HOST_API Material VKRT_materialDefault(void); /* vkrt_types.h:79-121 (a static inline there) */

/* ---- transforms (src/core/scene/transform.c:26-35,84-223) ---- */
VKRT_HOST_API void VKRT_buildMeshTransformMatrix(const vkrt_vec3 position, const vkrt_vec3 rotationDegrees, const vkrt_vec3 scale, vkrt_mat4 outMatrix);
VKRT_HOST_API void VKRT_buildImportedNodeTransform(vkrt_mat4 worldTransform, vkrt_mat4 outEngineTransform);
VKRT_HOST_API void VKRT_decomposeMeshTransform(vkrt_mat4 worldTransform, vkrt_vec3 outPosition, vkrt_vec3 outRotation, vkrt_vec3 outScale);
VKRT_HOST_API void VKRT_decomposeMeshNodeTransform(vkrt_mat4 worldTransform, vkrt_vec3 outPosition, vkrt_vec3 outRotation, vkrt_vec3 outScale);
/* src/core/utility/packing.c:144-156 */
VKRT_HOST_API void VKRT_packShaderVertex(const Vertex* vertex, ShaderVertex* outVertex);

/* feedback controllers as pure step functions (src/core/scene/timing.c:121-178, exposure.c:14-67,139-147), for front ends and tests */
VKRT_HOST_API uint32_t vkrtAutoSPPStep(float* ioControlMsPerSpp, float targetFrameMs, float measuredFrameMs, uint32_t samplesPerPixel);
VKRT_HOST_API int vkrtAutoExposureStep(float* ioFilteredLuminance, const float* samplesRgba, uint32_t sampleCount, float currentExposure, float* outExposure);
VKRT_HOST_API void vkrtAutoExposureProbePixels(uint32_t width, uint32_t height, uint32_t* outXY);

/* image decoding (src/core/utility/image.h:9-27): PNG / JPEG / EXR -> RGBA8 / RGBA16 UNORM / RGBA16F / RGBA32F; return 1 on success */
typedef struct VKRT_LoadedImage {
    void* pixels;
    uint32_t width;
    uint32_t height;
    uint32_t format;
    uint32_t colorSpace;
} VKRT_LoadedImage;
VKRT_HOST_API int vkrtLoadImageFromFile(const char* path, uint32_t preferredColorSpace, VKRT_LoadedImage* outImage);
VKRT_HOST_API int vkrtLoadImageFromMemory(const void* data, size_t size, const char* mimeType, uint32_t preferredColorSpace, VKRT_LoadedImage* outImage);
VKRT_HOST_API void vkrtFreeLoadedImage(VKRT_LoadedImage* image);
/* baseline 4:4:4 JPEG writer behind VKRT_saveRenderImage("*.jpg") (src/core/utility/export/image.c:220-263 uses quality 95) */
/* scanline OpenEXR writer behind VKRT_saveRenderImage("*.exr") (src/core/utility/exr.h:14 vkrtWriteEXRFromRGBA32F): RGBA, 32-bit float, uncompressed */
VKRT_HOST_API int vkrtWriteEXRFromRGBA32F(const char* path, const float* rgba32f, uint32_t width, uint32_t height);
VKRT_HOST_API int vkrtWriteJPEGFromRGBA8(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height, int quality);

/* The host-side stages of a denoised save, exported for the parity tests (tests/test_denoise.py compares them with the reference's own
 * source over a stand-in OIDN): image.c:840-905 denoiseLinearRenderOutput on a linear RGBA32F image with the RGBA16F feature AOVs
 * (returns 0 when the save must fail, else 1; `note` says why the raw image was kept), and image.c:641-699 convertLinearToDisplayRGBA16. */
VKRT_HOST_API int vkrtHostDenoiseLinear(float* linear, const uint16_t* albedoHalf, const uint16_t* normalHalf, uint32_t width, uint32_t height, int allowRawFallback,
                                        char* note, size_t noteLength);
/* image.c:907-960 prepareLinearRenderOutput on read-back buffers (what a denoised save runs between the read-backs and the file writer):
 * accumulation (RGBA32F, XYZ when `spectral`) -> linear sRGB with alpha 1, vkrtHostDenoiseLinear when `denoise`, non-finite channels -> 0 */
VKRT_HOST_API int vkrtHostPrepareLinearOutput(float* accumulation, const uint16_t* albedoHalf, const uint16_t* normalHalf, uint32_t width, uint32_t height, int spectral,
                                              int denoise, int allowRawFallback, char* note, size_t noteLength);
VKRT_HOST_API void vkrtHostLinearToDisplay16(const float* linear, uint32_t width, uint32_t height, uint32_t toneMappingMode, float exposure, uint32_t debugMode,
                                             uint16_t* outRgba16);
VKRT_HOST_API void vkrtHostResetDenoiser(void); /* forget the bound OIDN library (it is looked up again on the next use) */

/* ---- app layer: scene files, model import, procedural benchmark scenes, offline render loop ----
 * (src/app/scene/controller.c:1528-1597, src/app/mesh/loader.c:2133-2179, src/app/render/benchmark.c:13-293) */
VKRT_HOST_API VKRT_Result VKRT_appLoadScene(VKRT* vkrt, const char* scenePath);            /* vkrt.scene JSON v1 */
VKRT_HOST_API VKRT_Result VKRT_appImportMesh(VKRT* vkrt, const char* glbPath, uint32_t* outFirstMesh, uint32_t* outMeshCount);
VKRT_HOST_API VKRT_Result VKRT_appGenerateSoup(VKRT* vkrt, uint32_t triangleCount, uint32_t seed);       /* SURVEY §8d config C3 */
VKRT_HOST_API VKRT_Result VKRT_appGenerateInstanced(VKRT* vkrt, const char* glbPath, uint32_t instanceCount, uint32_t seed); /* C4 */
typedef struct VKRT_OfflineRenderResult {
    double seconds;        /* host wall clock over the timed frames */
    double deviceSeconds;  /* sum of device frame times */
    uint64_t samples;      /* spp accumulated during the timed frames */
    uint32_t frames;
    uint32_t samplesPerFrame;
    double samplesPerSecond; /* the reference's benchmark line (benchmark.c:109-137): 1 sample = 1 spp over the whole frame */
    double mpathsPerSecond;  /* samplesPerSecond * width * height / 1e6 */
    uint64_t extensionRays, shadowRays;
} VKRT_OfflineRenderResult;
VKRT_HOST_API VKRT_Result VKRT_appOfflineRender(VKRT* vkrt, uint32_t width, uint32_t height, uint32_t targetSamples, uint32_t samplesPerFrame,
                                                VKRT_OfflineRenderResult* outResult);

/* ---- introspection used by the tests and the bench (no reference equivalent) ---- */
typedef struct VKRT_PreparedScene {
    const ShaderVertex* vertices; uint32_t vertexCount;
    const uint32_t* indices; uint32_t indexCount;
    const MeshInfo* meshInfos; const float* world3x4; const uint32_t* geometrySource; const uint8_t* alphaTested; uint32_t meshCount;
    const Material* materials; uint32_t materialCount;
    const EmissiveMesh* emissiveMeshes; uint32_t emissiveMeshCount;
    const EmissiveTriangle* emissiveTriangles; uint32_t emissiveTriangleCount;
    const float* meshAliasQ; const uint32_t* meshAliasIdx; const float* triAliasQ; const uint32_t* triAliasIdx;
    const SceneData* sceneData;
} VKRT_PreparedScene;
/* Runs the host half of VKRT_updateScene (packing, dedup layout, MeshInfo, light tables, camera, SceneData) and exposes the arrays
 * that are handed to vkrt_cuda_set_*. Pointers stay valid until the next scene mutation. */
VKRT_HOST_API VKRT_Result VKRT_prepareScene(VKRT* vkrt, VKRT_PreparedScene* outScene);
VKRT_HOST_API vkrt_cuda_ctx* VKRT_cudaContext(VKRT* vkrt);
VKRT_HOST_API VKRT_Result VKRT_getLastFrameStats(const VKRT* vkrt, vkrt_cuda_frame_stats* outStats);
VKRT_HOST_API VKRT_Result VKRT_getBuildStats(const VKRT* vkrt, vkrt_cuda_build_stats* outStats);
VKRT_HOST_API const char* VKRT_lastError(const VKRT* vkrt);
/* Interleaved-tile partition of a width x height image over worldSize ranks (vkrt_b200/csrc/tiles.h; tile size 0 = 32): the ascending
 * global tile ids owned by `rank`. A rank's tile-compact film stores its tiles back to back, tileWidth*tileHeight pixels each. */
VKRT_HOST_API VKRT_Result VKRT_tilePartition(uint32_t width, uint32_t height, uint32_t tileWidth, uint32_t tileHeight, uint32_t rank, uint32_t worldSize,
                                             uint32_t* outLocalTileCount, uint32_t* outLocalToGlobalTile, uint32_t capacity, uint32_t* outTilesX,
                                             uint32_t* outTilesY);

#ifdef __cplusplus
}
#endif
#endif /* VKRT_HOST_H */
