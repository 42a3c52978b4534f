/* ref_host_entry.c -- TEST INFRASTRUCTURE (oracle/). Builds oracle/_ref/libvkrt_refhost.so from the REFERENCE'S OWN host sources,
 * compiled unmodified where they lie under /root/reference (oracle/Makefile passes -I<reference>/...):
 *
 *     src/core/utility/packing.c     packShaderVertex: the vertex wire format the shaders decode
 *     src/core/scene/transform.c     mesh transform compose / decompose (the lossy Euler PRS the shaders rebuild normals from)
 *     src/core/scene/camera.c        syncCameraMatrices: viewInverse / projInverse
 *     src/core/scene/lighting.c      emissive triangle list, two-level alias tables, lightPdfArea
 *     src/core/api/mesh.c            VKRT_addMaterial / VKRT_setMaterial (sanitizeMaterial), VKRT_setMeshTransform
 *     src/core/api/settings.c        render-setting setters and their clamps
 *     src/core/scene/uniform.c       resetSceneData / SceneData defaults
 *     src/core/internal/state.c      material / dirty-flag helpers
 *
 * The reference keeps its scene in one `VKRT` struct whose leaves are Vulkan buffers. This file supplies a "null device" behind that
 * struct: createDeviceBufferFromData keeps a host copy of what would have been uploaded (so the tests can read the light tables the
 * reference builds), every other device call is a stub. The Vulkan / GLFW type names come from oracle/ref_host/shim/. Nothing from the
 * reference is copied into the repository; a test that links this library fails if vkrt_b200/host/ or include/vkrt_shared.h drift from
 * the upstream sources. */
#include "vkrt_internal.h"
#include "buffer.h"
#include <stdarg.h>
#include "debug.h"
#include "packing.h"
#include "scene.h"
#include "state.h"
#include "textures.h"
#include "vkrt.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define REFHOST_API __attribute__((visibility("default")))

/* ---- null device -------------------------------------------------------------------------------------------------------------- */
typedef struct HostCopy {
    size_t size;
    unsigned char bytes[];
} HostCopy;

VKRT_Result createDeviceBufferFromData(VKRT* vkrt, const void* hostData, VkDeviceSize size, VkBufferUsageFlags usage, VkBuffer* outBuffer,
                                       VkDeviceMemory* outMemory, VkDeviceAddress* outDeviceAddress) {
    (void)vkrt; (void)usage;
    HostCopy* copy = (HostCopy*)malloc(sizeof(HostCopy) + (size_t)size);
    if (!copy) return VKRT_ERROR_OUT_OF_MEMORY;
    copy->size = (size_t)size;
    if (hostData && size) memcpy(copy->bytes, hostData, (size_t)size);
    *outBuffer = (VkBuffer)copy;
    if (outMemory) *outMemory = (VkDeviceMemory)copy;
    if (outDeviceAddress) *outDeviceAddress = (VkDeviceAddress)(uintptr_t)copy->bytes;
    return VKRT_SUCCESS;
}
VKRT_Result createBuffer(VKRT* vkrt, VkDeviceSize size, VkBufferUsageFlags usage, VkMemoryPropertyFlags properties, VkBuffer* buffer,
                         VkDeviceMemory* bufferMemory) {
    (void)properties;
    return createDeviceBufferFromData(vkrt, NULL, size, usage, buffer, bufferMemory, NULL);
}
void destroyBufferResources(VKRT* vkrt, Buffer* buffer) {
    (void)vkrt;
    if (!buffer) return;
    if (buffer->buffer) free((void*)buffer->buffer);
    buffer->buffer = VK_NULL_HANDLE;
    buffer->memory = VK_NULL_HANDLE;
    buffer->deviceAddress = 0;
    buffer->count = 0;
}
void vkUnmapMemory(VkDevice device, VkDeviceMemory memory) { (void)device; (void)memory; }
void vkDestroyBuffer(VkDevice device, VkBuffer buffer, const VkAllocationCallbacks* allocator) { (void)device; (void)buffer; (void)allocator; }
void vkFreeMemory(VkDevice device, VkDeviceMemory memory, const VkAllocationCallbacks* allocator) { (void)device; (void)allocator; free((void*)memory); }
VkResult vkMapMemory(VkDevice device, VkDeviceMemory memory, VkDeviceSize offset, VkDeviceSize size, VkFlags flags, void** data) {
    (void)device; (void)size; (void)flags;
    *data = ((HostCopy*)memory)->bytes + offset;
    return VK_SUCCESS;
}
VkResult vkGetQueryPoolResults(VkDevice d, VkQueryPool p, uint32_t first, uint32_t count, size_t size, void* data, VkDeviceSize stride, VkFlags flags) {
    (void)d; (void)p; (void)first; (void)count; (void)size; (void)data; (void)stride; (void)flags;
    return -1;
}
VKRT_Result createAutoExposureReadbacks(VKRT* vkrt) { (void)vkrt; return VKRT_SUCCESS; }
uint64_t getMicroseconds(void) { return 0; }   /* (wins over utility/platform.c's clock: first definition on the link line, --allow-multiple-definition) */
void vkrtLogLine(FILE* stream, const char* level, const char* format, ...) {   /* silent unless REFHOST_LOG is set (debugging the tests) */
    if (!getenv("REFHOST_LOG")) return;
    va_list args;
    va_start(args, format);
    fprintf(stream ? stream : stderr, "%s ", level ? level : "");
    vfprintf(stream ? stream : stderr, format, args);
    fputc('\n', stream ? stream : stderr);
    va_end(args);
}
/* texture registry (scene/textures.c needs the image loaders): no textures in the pinned host scenes */
void vkrtAdjustMaterialTextureUseCounts(VKRT* vkrt, const Material* material, int delta) { (void)vkrt; (void)material; (void)delta; }
uint32_t vkrtCountTextureUsers(const VKRT* vkrt, uint32_t textureIndex) { (void)vkrt; (void)textureIndex; return 0u; }
const SceneTexture* vkrtGetSceneTexture(const VKRT* vkrt, uint32_t textureIndex) {
    if (!vkrt || textureIndex >= vkrt->core.textureCount || !vkrt->core.textures) return NULL;
    return &vkrt->core.textures[textureIndex];
}
VKRT_Result vkrtSceneRemoveTexture(VKRT* vkrt, uint32_t textureIndex) { (void)vkrt; (void)textureIndex; return VKRT_SUCCESS; }
VkResult vkWaitForFences(VkDevice d, uint32_t n, const VkFence* f, VkBool32 all, uint64_t timeout) { (void)d; (void)n; (void)f; (void)all; (void)timeout; return VK_SUCCESS; }

/* ---- handle ---------------------------------------------------------------------------------------------------------------------- */
REFHOST_API void* refhost_create(uint32_t width, uint32_t height) {
    VKRT* vkrt = (VKRT*)calloc(1, sizeof(VKRT));
    if (!vkrt) return NULL;
    vkrt->core.sceneData = &vkrt->core.sceneDataHost;
    vkrt->runtime.renderExtent.width = width;
    vkrt->runtime.renderExtent.height = height;
    vkrt->core.sceneData->viewportRect[2] = width;
    vkrt->core.sceneData->viewportRect[3] = height;
    return vkrt;
}
REFHOST_API void refhost_destroy(void* h) {
    VKRT* vkrt = (VKRT*)h;
    if (!vkrt) return;
    for (uint32_t i = 0; i < vkrt->core.meshCount; i++) {
        free(vkrt->core.meshes[i].vertices);
        free(vkrt->core.meshes[i].indices);
    }
    free(vkrt->core.meshes);
    free(vkrt->core.materials);
    Buffer* lights[6] = {&vkrt->core.sceneEmissiveMeshData, &vkrt->core.sceneEmissiveTriangleData, &vkrt->core.sceneMeshAliasQ,
                         &vkrt->core.sceneMeshAliasIdx, &vkrt->core.sceneTriAliasQ, &vkrt->core.sceneTriAliasIdx};
    for (int i = 0; i < 6; i++) destroyBufferResources(vkrt, lights[i]);
    free(vkrt);
}

/* ---- pure functions ------------------------------------------------------------------------------------------------------------------ */
REFHOST_API void refhost_pack_vertices(const Vertex* in, uint32_t count, ShaderVertex* out) {
    for (uint32_t i = 0; i < count; i++) out[i] = packShaderVertex(&in[i]);
}
/* {sizeof, then offsetof of every field} of the seven shared structs, in declaration order (src/shared/types.h:25-150) */
REFHOST_API uint32_t refhost_struct_layout(uint32_t* out, uint32_t capacity) {
    uint32_t n = 0;
#define PUT(v) do { if (n < capacity) out[n] = (uint32_t)(v); n++; } while (0)
#define OFF(T, f) PUT(offsetof(T, f))
    PUT(sizeof(Vertex)); OFF(Vertex, position); OFF(Vertex, normal); OFF(Vertex, tangent); OFF(Vertex, color); OFF(Vertex, texcoord0); OFF(Vertex, texcoord1);
    PUT(sizeof(ShaderVertex)); OFF(ShaderVertex, position); OFF(ShaderVertex, texcoord0); OFF(ShaderVertex, texcoord1); OFF(ShaderVertex, packedNormal);
    OFF(ShaderVertex, packedTangent); OFF(ShaderVertex, packedColor);
    PUT(sizeof(MeshInfo)); OFF(MeshInfo, position); OFF(MeshInfo, vertexBase); OFF(MeshInfo, rotation); OFF(MeshInfo, vertexCount); OFF(MeshInfo, scale);
    OFF(MeshInfo, indexBase); OFF(MeshInfo, indexCount); OFF(MeshInfo, materialIndex); OFF(MeshInfo, renderBackfaces); OFF(MeshInfo, lightPdfArea);
    OFF(MeshInfo, opacity); OFF(MeshInfo, reserved0); OFF(MeshInfo, reserved1); OFF(MeshInfo, reserved2);
    PUT(sizeof(Material)); OFF(Material, baseColor); OFF(Material, roughness); OFF(Material, emissionColor); OFF(Material, emissionLuminance); OFF(Material, eta);
    OFF(Material, metallic); OFF(Material, k); OFF(Material, anisotropic); OFF(Material, specular); OFF(Material, specularTint); OFF(Material, abbeNumber);
    OFF(Material, reserved0); OFF(Material, sheenTintWeight); OFF(Material, clearcoat); OFF(Material, clearcoatGloss); OFF(Material, ior);
    OFF(Material, diffuseRoughness); OFF(Material, transmission); OFF(Material, subsurface); OFF(Material, sheenRoughness); OFF(Material, absorptionCoefficient);
    OFF(Material, attenuationColor); OFF(Material, normalTextureScale); OFF(Material, baseColorTextureIndex); OFF(Material, metallicRoughnessTextureIndex);
    OFF(Material, normalTextureIndex); OFF(Material, emissiveTextureIndex); OFF(Material, baseColorTextureWrap); OFF(Material, metallicRoughnessTextureWrap);
    OFF(Material, normalTextureWrap); OFF(Material, emissiveTextureWrap); OFF(Material, opacity); OFF(Material, alphaCutoff); OFF(Material, alphaMode);
    OFF(Material, textureTexcoordSets); OFF(Material, baseColorTextureTransform); OFF(Material, metallicRoughnessTextureTransform);
    OFF(Material, normalTextureTransform); OFF(Material, emissiveTextureTransform); OFF(Material, textureRotations);
    PUT(sizeof(EmissiveMesh)); OFF(EmissiveMesh, triOffset); OFF(EmissiveMesh, triCount); OFF(EmissiveMesh, pmfMesh); OFF(EmissiveMesh, invTotalArea);
    OFF(EmissiveMesh, emission); OFF(EmissiveMesh, reserved0);
    PUT(sizeof(EmissiveTriangle)); OFF(EmissiveTriangle, v0Area); OFF(EmissiveTriangle, e1Pad); OFF(EmissiveTriangle, e2Pad);
    PUT(sizeof(RGB2SpecTableInfo)); OFF(RGB2SpecTableInfo, res); OFF(RGB2SpecTableInfo, scaleOffset); OFF(RGB2SpecTableInfo, dataOffset);
    PUT(sizeof(SceneData)); OFF(SceneData, viewInverse); OFF(SceneData, projInverse); OFF(SceneData, frameNumber); OFF(SceneData, samplesPerPixel);
    OFF(SceneData, rrMaxDepth); OFF(SceneData, rrMinDepth); OFF(SceneData, viewportRect); OFF(SceneData, packedRenderSettings); OFF(SceneData, exposure);
    OFF(SceneData, timeBase); OFF(SceneData, timeStep); OFF(SceneData, environmentLight); OFF(SceneData, environmentTextureIndex);
    OFF(SceneData, environmentRotation); OFF(SceneData, debugMode); OFF(SceneData, misNeeEnabled); OFF(SceneData, emissiveMeshCount);
    OFF(SceneData, emissiveTriangleCount); OFF(SceneData, selectionEnabled); OFF(SceneData, selectedMeshIndex); OFF(SceneData, rgb2specSRGB);
#undef OFF
#undef PUT
    return n;
}
REFHOST_API void refhost_build_transform(const float* position, const float* rotationDegrees, const float* scale, float* outMat16) {
    mat4 m;
    VKRT_buildMeshTransformMatrix((float*)position, (float*)rotationDegrees, (float*)scale, m);
    memcpy(outMat16, m, sizeof(m));
}
REFHOST_API void refhost_decompose_transform(const float* mat16, float* position, float* rotationDegrees, float* scale) {
    mat4 m;
    memcpy(m, mat16, sizeof(m));
    VKRT_decomposeMeshTransform(m, position, rotationDegrees, scale);
}
REFHOST_API void refhost_imported_node_transform(const float* world16, float* outEngine16) {
    mat4 in, out;
    memcpy(in, world16, sizeof(in));
    VKRT_buildImportedNodeTransform(in, out);
    memcpy(outEngine16, out, sizeof(out));
}
REFHOST_API void refhost_material_default(Material* out) { *out = VKRT_materialDefault(); }

/* ---- scene through the reference's own API ------------------------------------------------------------------------------------------ */
REFHOST_API int refhost_add_material(void* h, const Material* material, uint32_t* outIndex) { return VKRT_addMaterial((VKRT*)h, material, "m", outIndex); }
REFHOST_API int refhost_set_material(void* h, uint32_t index, const Material* material) { return VKRT_setMaterial((VKRT*)h, index, material); }
REFHOST_API int refhost_remove_material(void* h, uint32_t index) { return VKRT_removeMaterial((VKRT*)h, index); }
REFHOST_API uint32_t refhost_material_count(void* h) { return ((VKRT*)h)->core.materialCount; }
REFHOST_API int refhost_get_material(void* h, uint32_t index, Material* out) {
    const Material* m = vkrtGetSceneMaterialData((VKRT*)h, index);
    if (!m) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = *m;
    return VKRT_SUCCESS;
}
/* Stands in for vkrtSceneUploadMeshDataBatch (scene/geometry.c:729-806), whose remainder is Vulkan staging: appends one mesh that owns
 * its geometry, identity transform, opacity 1, running vertex / index bases. */
REFHOST_API int refhost_add_mesh(void* h, const Vertex* vertices, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount, uint32_t materialIndex) {
    VKRT* vkrt = (VKRT*)h;
    Mesh* resized = (Mesh*)realloc(vkrt->core.meshes, (size_t)(vkrt->core.meshCount + 1u) * sizeof(Mesh));
    if (!resized) return VKRT_ERROR_OUT_OF_MEMORY;
    vkrt->core.meshes = resized;
    Mesh* mesh = &vkrt->core.meshes[vkrt->core.meshCount];
    memset(mesh, 0, sizeof(*mesh));
    uint32_t vertexBase = 0, indexBase = 0;
    for (uint32_t i = 0; i < vkrt->core.meshCount; i++) {
        vertexBase += vkrt->core.meshes[i].info.vertexCount;
        indexBase += vkrt->core.meshes[i].info.indexCount;
    }
    mesh->vertices = (Vertex*)aligned_alloc(16, ((size_t)vertexCount * sizeof(Vertex) + 15u) & ~(size_t)15u);
    mesh->indices = (uint32_t*)malloc((size_t)indexCount * sizeof(uint32_t));
    if (!mesh->vertices || !mesh->indices) return VKRT_ERROR_OUT_OF_MEMORY;
    memcpy(mesh->vertices, vertices, (size_t)vertexCount * sizeof(Vertex));
    memcpy(mesh->indices, indices, (size_t)indexCount * sizeof(uint32_t));
    mesh->info.vertexBase = vertexBase;
    mesh->info.vertexCount = vertexCount;
    mesh->info.indexBase = indexBase;
    mesh->info.indexCount = indexCount;
    mesh->info.materialIndex = materialIndex;
    mesh->info.opacity = 1.0f;
    mesh->info.scale[0] = mesh->info.scale[1] = mesh->info.scale[2] = 1.0f;
    glm_mat4_identity(mesh->worldTransform);
    mesh->ownsGeometry = 1;
    mesh->geometrySource = vkrt->core.meshCount;
    mesh->renderBackfacesOverride = -1;
    mesh->hasMaterialAssignment = 1;
    vkrt->core.meshCount++;
    return VKRT_SUCCESS;
}
REFHOST_API int refhost_set_mesh_transform(void* h, uint32_t meshIndex, float* position, float* rotationDegrees, float* scale) {
    return VKRT_setMeshTransform((VKRT*)h, meshIndex, position, rotationDegrees, scale);
}
REFHOST_API int refhost_set_mesh_transform_matrix(void* h, uint32_t meshIndex, const float* mat16) {
    mat4 m;
    memcpy(m, mat16, sizeof(m));
    return VKRT_setMeshTransformMatrix((VKRT*)h, meshIndex, m);
}
REFHOST_API int refhost_get_mesh(void* h, uint32_t meshIndex, MeshInfo* outInfo, float* outWorld3x4) {
    VKRT* vkrt = (VKRT*)h;
    if (meshIndex >= vkrt->core.meshCount) return VKRT_ERROR_INVALID_ARGUMENT;
    *outInfo = vkrt->core.meshes[meshIndex].info;
    VkTransformMatrixKHR t = getMeshWorldTransform(&vkrt->core.meshes[meshIndex]);
    memcpy(outWorld3x4, t.matrix, sizeof(t.matrix));
    return VKRT_SUCCESS;
}
REFHOST_API int refhost_rebuild_lights(void* h, uint32_t* outMeshCount, uint32_t* outTriangleCount) {
    VKRT* vkrt = (VKRT*)h;
    VKRT_Result r = vkrtSceneRebuildLightBuffers(vkrt);
    if (outMeshCount) *outMeshCount = vkrt->core.emissiveMeshCount;
    if (outTriangleCount) *outTriangleCount = vkrt->core.emissiveTriangleCount;
    return r;
}
/* which: 0 emissive meshes, 1 emissive triangles, 2 mesh alias q, 3 mesh alias idx, 4 triangle alias q, 5 triangle alias idx */
REFHOST_API int64_t refhost_read_light_buffer(void* h, int which, void* dst, uint64_t capacity) {
    VKRT* vkrt = (VKRT*)h;
    Buffer* lights[6] = {&vkrt->core.sceneEmissiveMeshData, &vkrt->core.sceneEmissiveTriangleData, &vkrt->core.sceneMeshAliasQ,
                         &vkrt->core.sceneMeshAliasIdx, &vkrt->core.sceneTriAliasQ, &vkrt->core.sceneTriAliasIdx};
    if (which < 0 || which > 5 || !lights[which]->buffer) return -1;
    HostCopy* copy = (HostCopy*)lights[which]->buffer;
    if (dst && capacity >= copy->size) memcpy(dst, copy->bytes, copy->size);
    return (int64_t)copy->size;
}
REFHOST_API int refhost_set_camera(void* h, const float* pos, const float* target, const float* up, float vfov, float nearZ, float farZ) {
    VKRT* vkrt = (VKRT*)h;
    Camera* cam = &vkrt->sceneSettings.camera;
    memcpy(cam->pos, pos, sizeof(vec3));
    memcpy(cam->target, target, sizeof(vec3));
    memcpy(cam->up, up, sizeof(vec3));
    cam->vfov = vfov;
    cam->nearZ = nearZ;
    cam->farZ = farZ;
    syncCameraMatrices(vkrt);
    return VKRT_SUCCESS;
}
REFHOST_API void refhost_get_scene_data(void* h, SceneData* out) { *out = *((VKRT*)h)->core.sceneData; }
REFHOST_API int refhost_set_path_depth(void* h, uint32_t rrMin, uint32_t rrMax) { return VKRT_setPathDepth((VKRT*)h, rrMin, rrMax); }
REFHOST_API int refhost_set_samples_per_pixel(void* h, uint32_t spp) { return VKRT_setSamplesPerPixel((VKRT*)h, spp); }
REFHOST_API int refhost_set_render_mode(void* h, uint32_t mode) { return VKRT_setRenderMode((VKRT*)h, (VKRT_RenderMode)mode); }
REFHOST_API int refhost_set_spectral_sampling_mode(void* h, uint32_t mode) { return VKRT_setSpectralSamplingMode((VKRT*)h, (VKRT_SpectralSamplingMode)mode); }
REFHOST_API int refhost_set_tone_mapping_mode(void* h, uint32_t mode) { return VKRT_setToneMappingMode((VKRT*)h, (VKRT_ToneMappingMode)mode); }
REFHOST_API int refhost_set_exposure(void* h, float exposure) { return VKRT_setExposure((VKRT*)h, exposure); }
REFHOST_API int refhost_set_environment_light(void* h, float* color, float strength) { return VKRT_setEnvironmentLight((VKRT*)h, color, strength); }
REFHOST_API int refhost_set_mis_nee_enabled(void* h, uint32_t enabled) { return VKRT_setMisNeeEnabled((VKRT*)h, enabled); }
