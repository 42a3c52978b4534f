/* Mutation fuzzer for the .glb importer (vkrt_b200/host/gltf_import.c + hjson.c): GLB container, JSON chunk, accessors / bufferViews,
 * materials, embedded images. Built with -fsanitize=address,undefined by tests/test_fuzz_decoders.py.
 *   fuzz_gltf_import <seed> <iterations> <scratch path> file.glb...                                                                 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../vkrt_b200/host/gltf_import.h"

static uint64_t s;
static uint32_t rnd(void) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    s = strtoull(argv[1], NULL, 10) * 0x9E3779B97F4A7C15ull + 1;
    const long iterations = atol(argv[2]);
    const char* scratch = argv[3];
    long imported = 0, rejected = 0;
    for (int f = 4; f < argc; f++) {
        FILE* fp = fopen(argv[f], "rb");
        if (!fp) { fprintf(stderr, "cannot open %s\n", argv[f]); return 2; }
        fseek(fp, 0, SEEK_END);
        const long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        uint8_t* orig = (uint8_t*)malloc((size_t)n);
        if (fread(orig, 1, (size_t)n, fp) != (size_t)n) return 2;
        fclose(fp);
        /* the JSON chunk starts at byte 20 and its length is at byte 12 */
        uint32_t jsonLen = 0;
        if (n > 20) memcpy(&jsonLen, orig + 12, 4);
        if (jsonLen > (uint32_t)(n - 20)) jsonLen = (uint32_t)(n - 20);
        for (long it = 0; it < iterations; it++) {
            size_t len = (size_t)n;
            const uint32_t kind = rnd() % 8u;
            if (kind == 0u) len = rnd() % (uint32_t)(n + 1);
            uint8_t* buf = (uint8_t*)malloc(len ? len : 1);
            memcpy(buf, orig, len);
            const uint32_t edits = kind == 0u ? 0u : 1u + rnd() % 4u;
            for (uint32_t e = 0; e < edits && len; e++) {
                size_t at = rnd() % len;
                if (kind <= 5u && jsonLen) at = (20u + rnd() % jsonLen) % len;   /* most edits land in the JSON chunk */
                if (kind == 6u) at = rnd() % (len < 28 ? len : 28);               /* container header */
                switch (rnd() % 6u) {
                    case 0: buf[at] ^= (uint8_t)(1u << (rnd() % 8u)); break;
                    case 1: buf[at] = (uint8_t)rnd(); break;
                    case 2: buf[at] = "0123456789-.e"[rnd() % 13u]; break;       /* turn a digit into another: counts, offsets, indices */
                    case 3: buf[at] = "{}[],:\"\\"[rnd() % 8u]; break;
                    case 4: if (at + 1 < len) { buf[at] = '9'; buf[at + 1] = '9'; } break;
                    default: for (size_t k = 0; k < 4 && at + k < len; k++) buf[at + k] = 0xFF;
                }
            }
            fp = fopen(scratch, "wb");
            if (!fp) return 2;
            fwrite(buf, 1, len, fp);
            fclose(fp);
            GltfImport imp;
            char err[512];
            memset(&imp, 0, sizeof(imp));
            if (gltfImportFile(scratch, &imp, err, sizeof(err))) {
                volatile float sink = 0;
                for (uint32_t m = 0; m < imp.meshCount; m++) {
                    const GltfMesh* g = &imp.meshes[m];
                    for (size_t k = 0; k < g->indexCount; k++) {
                        if (g->indices[k] >= g->vertexCount) { fprintf(stderr, "index %u out of %zu vertices accepted\n", g->indices[k], g->vertexCount); abort(); }
                        sink += g->vertices[g->indices[k]].position[0];
                    }
                    if (g->materialIndex >= (int)imp.materialCount) { fprintf(stderr, "material index out of range accepted\n"); abort(); }
                }
                (void)sink;
                gltfImportFree(&imp);
                imported++;
                /* and through the host: import into a host-only handle, then the host half of the scene update (packing, dedup,
                 * instance transforms, emissive / alias tables) */
                VKRT* vkrt = NULL;
                if (VKRT_create(&vkrt) == VKRT_SUCCESS) {
                    VKRT_CreateInfo ci;
                    VKRT_defaultCreateInfo(&ci);
                    ci.hostOnly = 1;
                    ci.width = 64; ci.height = 64;
                    if (VKRT_initWithCreateInfo(vkrt, &ci) == VKRT_SUCCESS && VKRT_appImportMesh(vkrt, scratch, NULL, NULL) == VKRT_SUCCESS) {
                        VKRT_PreparedScene prepared;
                        memset(&prepared, 0, sizeof(prepared));
                        if (VKRT_updateScene(vkrt) == VKRT_SUCCESS && VKRT_prepareScene(vkrt, &prepared) == VKRT_SUCCESS) {
                            for (uint32_t k = 0; k < prepared.emissiveMeshCount; k++)
                                if (prepared.meshAliasIdx[k] >= prepared.emissiveMeshCount) { fprintf(stderr, "mesh alias out of range\n"); abort(); }
                        }
                    }
                    VKRT_destroy(vkrt);
                }
            } else {
                rejected++;
            }
            free(buf);
        }
        free(orig);
    }
    remove(scratch);
    printf("imported %ld rejected %ld\n", imported, rejected);
    return 0;
}
