/* ref_scene_entry.c — the reference's OWN vkrt.scene reader and mesh upload path behind a flat C interface (TEST INFRASTRUCTURE, part of
 * libvkrt_refhost.so).
 *
 * Compiled where they lie by oracle/Makefile `ref`: src/app/scene/controller.c (the JSON reader, cJSON vendored), src/app/session/session.c
 * (scene-object hierarchy -> world transforms), src/app/mesh/controller.c (import glue), src/core/api/{query,geometry,texture,environment,
 * render}.c and src/core/scene/{geometry,environment}.c (mesh upload with geometry de-duplication). What stays stubbed is the GPU side of
 * those paths (zero-initialised device buffers, buffer addresses, BLAS builds, descriptor updates, swapchain / viewport / image state) and the
 * texture store (scene/textures.c is Vulkan image management: replaced by a list of {name, colour space}; the bundled scenes have no
 * textures, and scene files with textures are NOT covered by this pin). tests/test_reference_pin.py loads the bundled scenes through
 * sceneControllerLoadSceneFromPath and compares meshes, transforms, materials and settings with the product's VKRT_appLoadScene. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vkrt_internal.h"
#include "scene/controller.h"
#include "session.h"
#include "buffer.h"
#include "images.h"
#include "textures.h"
#include "view.h"

#define REFHOST_API __attribute__((visibility("default")))

/* ---- GPU-side stubs ---- */
VKRT_Result createDeviceBufferFromData(VKRT* vkrt, const void* hostData, VkDeviceSize size, VkBufferUsageFlags usage, VkBuffer* outBuffer,
                                       VkDeviceMemory* outMemory, VkDeviceAddress* outDeviceAddress);
VKRT_Result createZeroInitializedDeviceBuffer(VKRT* vkrt, VkDeviceSize size, VkBufferUsageFlags usage, Buffer* outBuffer) {
    void* zeros = calloc(1, (size_t)size ? (size_t)size : 1u);
    VKRT_Result r = createDeviceBufferFromData(vkrt, zeros, size, usage, &outBuffer->buffer, &outBuffer->memory, &outBuffer->deviceAddress);
    free(zeros);
    return r;
}
VkDeviceAddress queryBufferDeviceAddress(VKRT* vkrt, VkBuffer buffer) { (void)vkrt; return (VkDeviceAddress)(uintptr_t)buffer; }
VKRT_Result createBottomLevelAccelerationStructureForGeometry(VKRT* vkrt, const MeshInfo* meshInfo, VkDeviceAddress vertexDataAddress,
                                                              VkDeviceAddress indexDataAddress, AccelerationStructure* out) {
    (void)vkrt; (void)meshInfo; (void)vertexDataAddress; (void)indexDataAddress;
    if (out) memset(out, 0, sizeof(*out));
    return VKRT_SUCCESS;
}
VKRT_Result updateAllDescriptorSets(VKRT* vkrt) { (void)vkrt; return VKRT_SUCCESS; }
void vkrtCleanupPendingGeometryUploads(VKRT* vkrt, FrameSceneUpdate* update) { (void)vkrt; (void)update; }
VkBool32 vkrtUsesRenderPresentProfile(const VKRT* vkrt) { (void)vkrt; return VK_FALSE; }
void vkrtRefreshPresentModeIfNeeded(VKRT* vkrt, VkBool32 previous) { (void)vkrt; (void)previous; }
void vkrtClampViewportRect(VkExtent2D extent, uint32_t* x, uint32_t* y, uint32_t* width, uint32_t* height) { (void)extent; (void)x; (void)y; (void)width; (void)height; }
void vkrtQueryRenderViewCropExtent(VkExtent2D r, VkExtent2D v, float zoom, uint32_t* w, uint32_t* h, VkBool32* fill) {
    (void)v; (void)zoom;
    if (w) *w = r.width;
    if (h) *h = r.height;
    if (fill) *fill = VK_TRUE;
}
void vkrtClampRenderViewPanOffset(VkExtent2D r, VkExtent2D v, float zoom, float* panX, float* panY) { (void)r; (void)v; (void)zoom; (void)panX; (void)panY; }
int saveCurrentRenderImageEx(VKRT* vkrt, const char* path, const VKRT_RenderExportSettings* settings) { (void)vkrt; (void)path; (void)settings; return -1; }
VKRT_Result createGPUImageState(VKRT* vkrt, VkExtent2D extent, GPUImageState* outState) { (void)vkrt; (void)extent; if (outState) memset(outState, 0, sizeof(*outState)); return VKRT_SUCCESS; }
void destroyGPUImageState(VKRT* vkrt, GPUImageState* state) { (void)vkrt; (void)state; }
void captureGPUImageState(const VKRT* vkrt, GPUImageState* outState) { (void)vkrt; if (outState) memset(outState, 0, sizeof(*outState)); }
void applyGPUImageState(VKRT* vkrt, const GPUImageState* state) { (void)vkrt; (void)state; }

/* ---- texture store stand-in (see header) ---- */
static VKRT_Result appendTexture(VKRT* vkrt, const char* name, uint32_t colorSpace, uint32_t* outIndex) {
    SceneTexture* grown = (SceneTexture*)realloc(vkrt->core.textures, (size_t)(vkrt->core.textureCount + 1u) * sizeof(SceneTexture));
    if (!grown) return VKRT_ERROR_OUT_OF_MEMORY;
    vkrt->core.textures = grown;
    SceneTexture* t = &grown[vkrt->core.textureCount];
    memset(t, 0, sizeof(*t));
    t->width = t->height = 1u;
    t->colorSpace = colorSpace;
    snprintf(t->name, sizeof(t->name), "%s", name ? name : "Texture");
    if (outIndex) *outIndex = vkrt->core.textureCount;
    vkrt->core.textureCount++;
    return VKRT_SUCCESS;
}
VKRT_Result vkrtSceneAddTextureFromFile(VKRT* vkrt, const char* path, const char* name, uint32_t colorSpace, uint32_t* outTextureIndex) {
    return appendTexture(vkrt, name && name[0] ? name : path, colorSpace, outTextureIndex);
}
VKRT_Result vkrtSceneAddTextureFromPixels(VKRT* vkrt, const VKRT_TextureUpload* upload, uint32_t* outTextureIndex) {
    return appendTexture(vkrt, upload ? upload->name : NULL, upload ? upload->colorSpace : 0u, outTextureIndex);
}
VKRT_Result vkrtSceneAddTexturesBatch(VKRT* vkrt, const VKRT_TextureUpload* uploads, size_t uploadCount, uint32_t* outTextureIndices) {
    for (size_t i = 0; i < uploadCount; i++) {
        VKRT_Result r = appendTexture(vkrt, uploads[i].name, uploads[i].colorSpace, outTextureIndices ? &outTextureIndices[i] : NULL);
        if (r != VKRT_SUCCESS) return r;
    }
    return VKRT_SUCCESS;
}
VKRT_Result vkrtSceneSetMaterialTexture(VKRT* vkrt, uint32_t materialIndex, uint32_t textureSlot, uint32_t textureIndex) {
    if (!vkrt || materialIndex >= vkrt->core.materialCount) return VKRT_ERROR_INVALID_ARGUMENT;
    Material* m = &vkrt->core.materials[materialIndex].material;
    uint32_t* slots[4] = {&m->baseColorTextureIndex, &m->metallicRoughnessTextureIndex, &m->normalTextureIndex, &m->emissiveTextureIndex};
    if (textureSlot >= 4u) return VKRT_ERROR_INVALID_ARGUMENT;
    *slots[textureSlot] = textureIndex;
    return VKRT_SUCCESS;
}

/* ---- entry points ---- */
static Session g_session;
static int g_sessionLive = 0;

REFHOST_API int refscene_load(void* h, const char* path) {
    if (g_sessionLive) sessionDeinit(&g_session);
    sessionInit(&g_session);
    g_sessionLive = 1;
    return sceneControllerLoadSceneFromPath((VKRT*)h, &g_session, path);
}
REFHOST_API void refscene_close(void) {
    if (g_sessionLive) sessionDeinit(&g_session);
    g_sessionLive = 0;
}
REFHOST_API uint32_t refscene_mesh_count(void* h) { return ((VKRT*)h)->core.meshCount; }
/* info: the MeshInfo as uploaded; world: 4x4 column-major; misc: [0] geometrySource, [1] ownsGeometry, [2] hasMaterialAssignment, [3] renderBackfaces (resolved) */
REFHOST_API void refscene_mesh(void* h, uint32_t i, MeshInfo* info, float* world16, uint32_t* misc, char* name, size_t nameSize) {
    const Mesh* m = &((VKRT*)h)->core.meshes[i];
    *info = m->info;
    memcpy(world16, m->worldTransform, 64);
    misc[0] = m->geometrySource; misc[1] = m->ownsGeometry; misc[2] = m->hasMaterialAssignment; misc[3] = m->info.renderBackfaces;
    if (name && nameSize) { snprintf(name, nameSize, "%s", m->name); }
}
REFHOST_API void refscene_settings(void* h, VKRT_SceneSettingsSnapshot* out) { VKRT_getSceneSettings((VKRT*)h, out); }
