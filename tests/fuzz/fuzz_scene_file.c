/* Mutation fuzzer for the vkrt.scene reader (vkrt_b200/host/scene_file.c + hjson.c) on a host-only handle (no device): every mutated
 * scene file must load or be rejected with a message. Built with -fsanitize=address,undefined by tests/test_fuzz_decoders.py.
 *   fuzz_scene_file <seed> <iterations> <scratch path inside a scenes/ directory> scene.json...                                     */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vkrt_host.h"

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
    long loaded = 0, rejected = 0;
    for (int f = 4; f < argc; f++) {
        FILE* fp = fopen(argv[f], "rb");
        if (!fp) { fprintf(stderr, "cannot open %s\n", argv[f]); return 2; }
        fseek(fp, 0, SEEK_END);
        const long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        char* orig = (char*)malloc((size_t)n);
        if (fread(orig, 1, (size_t)n, fp) != (size_t)n) return 2;
        fclose(fp);
        for (long it = 0; it < iterations; it++) {
            size_t len = (size_t)n;
            const uint32_t kind = rnd() % 8u;
            if (kind == 0u) len = rnd() % (uint32_t)(n + 1);
            char* buf = (char*)malloc(len ? len : 1);
            memcpy(buf, orig, len);
            const uint32_t edits = kind == 0u ? 0u : 1u + rnd() % 4u;
            for (uint32_t e = 0; e < edits && len; e++) {
                const size_t at = rnd() % len;
                switch (rnd() % 6u) {
                    case 0: buf[at] ^= (char)(1u << (rnd() % 7u)); break;
                    case 1: buf[at] = "0123456789-.eE"[rnd() % 14u]; break;
                    case 2: buf[at] = "{}[],:\"\\"[rnd() % 8u]; break;
                    case 3: if (at + 3 < len) memcpy(buf + at, "1e99", 4); break;
                    case 4: if (at + 3 < len) memcpy(buf + at, "-999", 4); break;
                    default: buf[at] = (char)rnd();
                }
            }
            fp = fopen(scratch, "wb");
            if (!fp) return 2;
            fwrite(buf, 1, len, fp);
            fclose(fp);
            VKRT* vkrt = NULL;
            if (VKRT_create(&vkrt) != VKRT_SUCCESS) return 2;
            VKRT_CreateInfo ci;
            VKRT_defaultCreateInfo(&ci);
            ci.hostOnly = 1;
            ci.width = 64; ci.height = 64;
            if (VKRT_initWithCreateInfo(vkrt, &ci) != VKRT_SUCCESS) return 2;
            if (VKRT_appLoadScene(vkrt, scratch) == VKRT_SUCCESS) {
                /* the host half of the scene update: geometry packing, dedup, instance transforms, light tables, uniform */
                VKRT_PreparedScene prepared;
                memset(&prepared, 0, sizeof(prepared));
                if (VKRT_updateScene(vkrt) == VKRT_SUCCESS) (void)VKRT_prepareScene(vkrt, &prepared);
                loaded++;
            } else {
                rejected++;
            }
            VKRT_destroy(vkrt);
            free(buf);
        }
        free(orig);
    }
    remove(scratch);
    printf("loaded %ld rejected %ld\n", loaded, rejected);
    return 0;
}
