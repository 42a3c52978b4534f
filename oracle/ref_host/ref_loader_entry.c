/* ref_loader_entry.c — the reference's OWN glTF importer behind a flat C interface (TEST INFRASTRUCTURE, part of libvkrt_refhost.so).
 *
 * oracle/Makefile `ref` compiles /root/reference/src/app/mesh/{loader.c, cgltf_impl.c} (with the vendored cgltf.h) and
 * src/core/utility/{io.c, platform.c} where they lie; this file adds the two things the importer links against that are not built here —
 * the image decoders (libspng / libjpeg-turbo / tinyexr behind vkrtLoadImage*: replaced by a decoder that reports a 1 x 1 white RGBA8
 * texel, so texture ENTRIES and material texture indices can still be compared) — and accessors that copy the import result out.
 * tests/test_reference_pin.py compares the product's importer (vkrt_b200/host/gltf_import.c) with it on the bundled .glb models. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "image.h"
#include "loader.h"

#define REFHOST_API __attribute__((visibility("default")))

static void whiteTexel(uint32_t colorSpace, VKRT_LoadedImage* out) {
    uint8_t* px = (uint8_t*)malloc(4);
    px[0] = px[1] = px[2] = px[3] = 255u;
    memset(out, 0, sizeof(*out));
    out->pixels = px;
    out->width = out->height = 1u;
    out->format = 0u; /* VKRT_TEXTURE_FORMAT_RGBA8_UNORM */
    out->colorSpace = colorSpace;
}
int vkrtLoadImageFromFile(const char* path, uint32_t preferredColorSpace, VKRT_LoadedImage* outImage) {
    (void)path;
    whiteTexel(preferredColorSpace, outImage);
    return 1;
}
int vkrtLoadImageFromMemory(const void* data, size_t size, const char* mimeType, uint32_t preferredColorSpace, VKRT_LoadedImage* outImage) {
    (void)data; (void)size; (void)mimeType;
    whiteTexel(preferredColorSpace, outImage);
    return 1;
}
void vkrtFreeLoadedImage(VKRT_LoadedImage* image) {
    if (image) { free(image->pixels); memset(image, 0, sizeof(*image)); }
}

REFHOST_API void* refloader_load(const char* path) {
    MeshImportData* d = (MeshImportData*)calloc(1, sizeof(MeshImportData));
    if (meshLoadFromFile(path, d) != 0) { free(d); return NULL; }
    return d;
}
REFHOST_API void refloader_free(void* h) {
    if (!h) return;
    meshReleaseImportData((MeshImportData*)h);
    free(h);
}
/* counts: [0] mesh entries, [1] nodes, [2] materials, [3] textures */
REFHOST_API void refloader_counts(void* h, uint32_t* out) {
    MeshImportData* d = (MeshImportData*)h;
    out[0] = d->count; out[1] = d->nodeCount; out[2] = d->materialCount; out[3] = d->textureCount;
}
/* info: [0] vertexCount, [1] indexCount, [2] nodeIndex, [3] materialIndex, [4] renderBackfaces; prs = position, rotation, scale */
REFHOST_API void refloader_entry_info(void* h, uint32_t i, uint64_t* info, float* prs, char* name, size_t nameSize) {
    const MeshImportEntry* e = &((MeshImportData*)h)->entries[i];
    info[0] = e->vertexCount; info[1] = e->indexCount; info[2] = e->nodeIndex; info[3] = e->materialIndex; info[4] = e->renderBackfaces;
    memcpy(prs, e->position, 12); memcpy(prs + 3, e->rotation, 12); memcpy(prs + 6, e->scale, 12);
    if (name && nameSize) { strncpy(name, e->name ? e->name : "", nameSize - 1); name[nameSize - 1] = 0; }
}
REFHOST_API void refloader_entry_data(void* h, uint32_t i, Vertex* vertices, uint32_t* indices) {
    const MeshImportEntry* e = &((MeshImportData*)h)->entries[i];
    if (vertices) memcpy(vertices, e->vertices, e->vertexCount * sizeof(Vertex));
    if (indices) memcpy(indices, e->indices, e->indexCount * sizeof(uint32_t));
}
REFHOST_API void refloader_material(void* h, uint32_t i, Material* out, char* name, size_t nameSize) {
    const MaterialImportEntry* m = &((MeshImportData*)h)->materials[i];
    *out = m->material;
    if (name && nameSize) { strncpy(name, m->name ? m->name : "", nameSize - 1); name[nameSize - 1] = 0; }
}
/* node: local 4x4 (column-major, cglm), parent index, mesh entry count, position / rotation / scale */
REFHOST_API void refloader_node(void* h, uint32_t i, float* local16, uint32_t* parentAndCount, float* prs) {
    const NodeImportEntry* n = &((MeshImportData*)h)->nodes[i];
    memcpy(local16, n->localTransform, 64);
    parentAndCount[0] = n->parentIndex; parentAndCount[1] = n->meshEntryCount;
    memcpy(prs, n->position, 12); memcpy(prs + 3, n->rotation, 12); memcpy(prs + 6, n->scale, 12);
}
REFHOST_API void refloader_texture(void* h, uint32_t i, uint32_t* whFormatSpace, char* name, size_t nameSize) {
    const TextureImportEntry* t = &((MeshImportData*)h)->textures[i];
    whFormatSpace[0] = t->width; whFormatSpace[1] = t->height; whFormatSpace[2] = t->format; whFormatSpace[3] = t->colorSpace;
    if (name && nameSize) { strncpy(name, t->name ? t->name : "", nameSize - 1); name[nameSize - 1] = 0; }
}
