/* Mutation fuzzer for the texture decoders (vkrt_b200/host/image_decode.c: PNG, baseline JPEG, scanline EXR incl. PIZ).
 * Built with -fsanitize=address,undefined by tests/test_fuzz_decoders.py and fed the committed fixture files: every mutated input must
 * either decode or be rejected with a message; any out-of-bounds access, overflow or leak aborts the process.
 *   fuzz_image_decode <seed> <iterations> file...                                                                                     */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../vkrt_b200/host/image_decode.h"

static uint64_t s;
static uint32_t rnd(void) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    s = strtoull(argv[1], NULL, 10) * 0x9E3779B97F4A7C15ull + 1;
    const long iterations = atol(argv[2]);
    long decoded = 0, rejected = 0;
    for (int f = 3; f < argc; f++) {
        FILE* fp = fopen(argv[f], "rb");
        if (!fp) { fprintf(stderr, "cannot open %s\n", argv[f]); return 2; }
        fseek(fp, 0, SEEK_END);
        const long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        uint8_t* orig = (uint8_t*)malloc((size_t)n);
        if (fread(orig, 1, (size_t)n, fp) != (size_t)n) return 2;
        fclose(fp);
        for (long it = 0; it < iterations; it++) {
            size_t len = (size_t)n;
            /* exact-size heap copy so that a read one byte past the input trips ASan */
            const uint32_t kind = rnd() % 8u;
            if (kind == 0u) len = rnd() % (uint32_t)(n + 1);                      /* truncation */
            uint8_t* buf = (uint8_t*)malloc(len ? len : 1);
            memcpy(buf, orig, len);
            if (len) {
                const uint32_t edits = kind == 0u ? 0u : 1u + rnd() % 6u;
                for (uint32_t e = 0; e < edits; e++) {
                    const size_t at = (kind == 1u ? rnd() % 64u : rnd()) % len;   /* kind 1: concentrate on the header */
                    switch (rnd() % 5u) {
                        case 0: buf[at] ^= (uint8_t)(1u << (rnd() % 8u)); break;
                        case 1: buf[at] = (uint8_t)rnd(); break;
                        case 2: buf[at] = 0xFF; break;
                        case 3: buf[at] = 0x00; break;
                        default: {                                                 /* clobber a 4- or 8-byte length / offset field */
                            const size_t width = (rnd() & 1u) ? 8 : 4;
                            const uint32_t mode = rnd() % 3u;
                            for (size_t k = 0; k < width && at + k < len; k++) buf[at + k] = mode == 0u ? 0xFF : (mode == 1u ? (uint8_t)rnd() : ((rnd() & 1u) ? 0xFF : 0x7F));
                        }
                    }
                }
            }
            HostImage img;
            char err[256];
            memset(&img, 0, sizeof(img));
            if (hostDecodeImage(buf, len, NULL, "fuzz", (it & 1) ? 1u : 0u, &img, err, sizeof(err))) {
                /* touch every output byte */
                static const size_t bpp[4] = {4, 8, 8, 16};
                const size_t bytes = (size_t)img.width * img.height * bpp[img.format & 3u];
                volatile uint8_t sink = 0;
                for (size_t k = 0; k < bytes; k += 61) sink ^= ((uint8_t*)img.pixels)[k];
                if (bytes) sink ^= ((uint8_t*)img.pixels)[bytes - 1];
                (void)sink;
                hostFreeImage(&img);
                decoded++;
            } else {
                rejected++;
            }
            free(buf);
        }
        free(orig);
    }
    printf("decoded %ld rejected %ld\n", decoded, rejected);
    return 0;
}
