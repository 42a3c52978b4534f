/* gltf_import.h — result of importing one .glb (see gltf_import.c). */
#ifndef VKRT_HOST_GLTF_IMPORT_H
#define VKRT_HOST_GLTF_IMPORT_H

#include "../../include/vkrt_host.h"

typedef struct GltfMesh {
    Vertex* vertices;     /* engine space (Z-up), 16-byte aligned */
    size_t vertexCount;
    uint32_t* indices;
    size_t indexCount;
    float world[4][4];    /* engine-space world matrix of the owning node (column-major) */
    int materialIndex;    /* into GltfImport.materials, -1 = none */
    int doubleSided;
    char name[VKRT_NAME_LEN];
} GltfMesh;

typedef struct GltfTexture {   /* one decoded (image, colour space) pair; Material.*TextureIndex of this import index this list */
    void* pixels;
    uint32_t width, height, format, colorSpace;
    char name[VKRT_NAME_LEN];
} GltfTexture;

typedef struct GltfImport {
    GltfMesh* meshes;
    uint32_t meshCount;
    Material* materials;
    char (*materialNames)[VKRT_NAME_LEN];
    uint32_t materialCount;
    GltfTexture* textures;
    uint32_t textureCount;
} GltfImport;

/* (exported so that tests/test_reference_pin.py can compare the importer with the reference's own src/app/mesh/loader.c) */
VKRT_HOST_API int gltfImportFile(const char* path, GltfImport* out, char* error, size_t errorSize); /* 1 = ok */
VKRT_HOST_API void gltfImportFree(GltfImport* imp);

#endif
