/* host_state.h — internal state of libvkrt_host (the VKRT handle). Mirrors what the reference keeps in VKRT.core /
 * VKRT.sceneSettings / VKRT.renderStatus (src/core/internal/vkrt_internal.h), minus every Vulkan object. */
#ifndef VKRT_HOST_STATE_H
#define VKRT_HOST_STATE_H

#include <stdio.h>
#include <stdlib.h>

#include "../../include/vkrt_host.h"
#include "hmath.h"

typedef struct HostMesh {
    Vertex* vertices;      /* owned iff ownsGeometry */
    uint32_t* indices;
    MeshInfo info;
    hmat4 worldTransform;
    uint64_t fingerprint;
    uint32_t geometrySource; /* index of the mesh that owns the geometry (self when ownsGeometry) */
    uint8_t ownsGeometry;
    uint8_t hasMaterialAssignment;
    int8_t renderBackfacesOverride; /* -1 = follow material default */
    char name[VKRT_NAME_LEN];
} HostMesh;

typedef struct HostMaterial {
    Material material;
    char name[VKRT_NAME_LEN];
} HostMaterial;

typedef struct HostTexture {
    void* pixels;
    uint32_t width, height, format, colorSpace;
    char name[VKRT_NAME_LEN];
} HostTexture;

struct VKRT {
    int initialized;
    int hostOnly;
    VKRT_CreateInfo createInfo;
    vkrt_cuda_ctx* cuda;
    char error[512];

    HostMesh* meshes;
    uint32_t meshCount, meshCapacity;
    HostMaterial* materials;
    uint32_t materialCount;
    HostTexture* textures;
    uint32_t textureCount;

    VKRT_SceneSettingsSnapshot sceneSettings;
    SceneData sceneData;
    VKRT_RenderStatusSnapshot renderStatus;
    uint32_t renderWidth, renderHeight;

    /* dirty tracking (reference: revision counters in src/core/internal/state.c:105-128) */
    int geometryDirty, sceneResourcesDirty, materialsDirty, lightsDirty, texturesDirty, accelDirty, filmDirty;
    int accumulationNeedsReset;
    /* feedback controllers (controllers.c): smoothed ms per spp, explicit frame budget, filtered probe luminance */
    float autoSPPControlMs, autoSPPTargetFrameMs, autoExposureFilteredLuminance;
    int frameTraced, framePresented;

    /* prepared device-format arrays (what vkrt_cuda_set_* receives) */
    ShaderVertex* packedVertices; uint32_t packedVertexCount, packedVertexCapacity;
    uint32_t* packedIndices; uint32_t packedIndexCount, packedIndexCapacity;
    MeshInfo* meshInfos; float* world3x4; uint32_t* geometrySource; uint8_t* alphaTested; uint32_t preparedMeshCapacity;
    Material* materialArray; uint32_t materialArrayCapacity;
    EmissiveMesh* emissiveMeshes; EmissiveTriangle* emissiveTriangles;
    float* meshAliasQ; uint32_t* meshAliasIdx; float* triAliasQ; uint32_t* triAliasIdx;
    uint32_t emissiveMeshCount, emissiveTriangleCount, emissiveMeshCapacity, emissiveTriangleCapacity;

    float* rgb2spec; uint32_t rgb2specFloats; RGB2SpecTableInfo rgb2specInfo; int rgb2specDirty;

    vkrt_cuda_frame_stats lastFrameStats;
    vkrt_cuda_build_stats buildStats;
    uint64_t totalExtensionRays, totalShadowRays;
    double totalDeviceMs;
};

/* scene_prep.c */
VKRT_Result hostPrepareGeometry(VKRT* vkrt);
VKRT_Result hostPrepareMeshInfos(VKRT* vkrt);
VKRT_Result hostPrepareMaterials(VKRT* vkrt);
VKRT_Result hostRebuildLights(VKRT* vkrt);
void hostSyncCameraMatrices(VKRT* vkrt);
void hostWriteSceneStateUniform(VKRT* vkrt);
int hostBuildAliasTable(const float* pmf, uint32_t count, float* outQ, uint32_t* outIdx);
uint64_t hostGeometryFingerprint(const Vertex* vertices, size_t vertexCount, const uint32_t* indices, size_t indexCount);
Material hostSanitizeMaterial(const VKRT* vkrt, Material material);
int hostMaterialMayRejectRayHit(const Material* material, float meshOpacity);
VKRT_Result hostFail(VKRT* vkrt, VKRT_Result code, const char* fmt, ...);
void hostResetSceneData(VKRT* vkrt);
void hostResetAutoSPPState(VKRT* vkrt, int resetSamplesPerPixel);
void hostUpdateAutoSPP(VKRT* vkrt);
VKRT_Result hostUpdateAutoExposure(VKRT* vkrt);

static inline float hostFiniteClampf(float v, float fallback, float lo, float hi) {
    if (!isfinite(v)) v = fallback;
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}
static inline float hostFiniteOrf(float v, float fallback) { return isfinite(v) ? v : fallback; }

#endif
