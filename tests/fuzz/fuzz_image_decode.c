/* Mutation fuzzer for the texture decoders (vkrt_b200/host/image_decode.c: PNG, baseline JPEG, scanline EXR incl. PIZ).
 * Built with -fsanitize=address,undefined by tests/test_fuzz_decoders.py and fed the committed fixture files: every mutated input must
 * either decode or be rejected with a message; any out-of-bounds access, overflow or leak aborts the process.
 *   fuzz_image_decode <seed> <iterations> file...                                                                                     */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <zlib.h>

#include "../../vkrt_b200/host/image_decode.h"

static uint64_t s;
static uint32_t rnd(void) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void putBe32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static size_t putChunk(uint8_t* out, const char* type, const uint8_t* data, uint32_t len) {
    putBe32(out, len);
    memcpy(out + 4, type, 4);
    if (len) memcpy(out + 8, data, len);
    uLong crc = crc32(0L, (const Bytef*)type, 4);
    if (len) crc = crc32(crc, data, len);   /* crc32(x, NULL, 0) would reset the value */
    putBe32(out + 8 + len, (uint32_t)crc);
    return 12u + len;
}
/* Structure-aware PNG mutation: most byte flips die at the chunk CRC or the zlib checksum, so this one edits the IHDR fields and the
 * FILTERED scanline bytes (filter types, pixel data) and then re-compresses and re-signs the file. Returns a malloc'ed file or NULL. */
static uint8_t* mutatePng(const uint8_t* in, size_t n, size_t* outLen) {
    if (n < 33 || memcmp(in, "\x89PNG\r\n\x1a\n", 8) != 0) return NULL;
    uint8_t ihdr[13];
    uint8_t* idat = (uint8_t*)malloc(n);
    uint8_t* other = (uint8_t*)malloc(n);   /* ancillary chunks before IDAT, kept verbatim (already chunk-framed) */
    size_t idatLen = 0, otherLen = 0, pos = 8;
    int haveIhdr = 0;
    while (pos + 12 <= n) {
        const uint32_t len = be32(in + pos);
        if (len > n - pos - 12) break;
        const uint8_t* type = in + pos + 4;
        if (!memcmp(type, "IHDR", 4) && len == 13) { memcpy(ihdr, in + pos + 8, 13); haveIhdr = 1; }
        else if (!memcmp(type, "IDAT", 4)) { memcpy(idat + idatLen, in + pos + 8, len); idatLen += len; }
        else if (memcmp(type, "IEND", 4)) { memcpy(other + otherLen, in + pos, 12u + len); otherLen += 12u + len; }
        pos += 12u + len;
    }
    uint8_t* result = NULL;
    uLongf rawLen = 1u << 22;
    uint8_t* raw = (uint8_t*)malloc(rawLen);
    if (haveIhdr && uncompress(raw, &rawLen, idat, (uLong)idatLen) == Z_OK && rawLen) {
        const uint32_t what = rnd() % 4u;
        if (what != 0u) {   /* scanline bytes: a filter byte now and then, pixel bytes otherwise */
            const uint32_t edits = 1u + rnd() % 8u;
            for (uint32_t e = 0; e < edits; e++) raw[rnd() % rawLen] = (rnd() & 3u) ? (uint8_t)rnd() : (uint8_t)(rnd() % 6u);
            if ((rnd() & 7u) == 0u) rawLen = 1u + rnd() % rawLen;   /* short pixel data */
        }
        if (what != 1u) {   /* header fields */
            switch (rnd() % 6u) {
                case 0: putBe32(ihdr, be32(ihdr) + (rnd() % 5u) - 2u); break;                 /* width +-2 */
                case 1: putBe32(ihdr + 4, be32(ihdr + 4) + (rnd() % 5u) - 2u); break;         /* height +-2 */
                case 2: ihdr[8] = (uint8_t)(1u << (rnd() % 5u)); break;                        /* bit depth 1..16 */
                case 3: ihdr[9] = (uint8_t)("\0\2\3\4\6\5\7"[rnd() % 7u]); break;            /* colour type */
                case 4: ihdr[12] = (uint8_t)(rnd() % 3u); break;                               /* interlace */
                default: putBe32(ihdr + (rnd() & 4u), (rnd() & 1u) ? 0x7FFFFFFFu : (rnd() % 70000u)); /* huge / odd dimension */
            }
        }
        uLongf zLen = compressBound(rawLen);
        uint8_t* z = (uint8_t*)malloc(zLen);
        if (compress2(z, &zLen, raw, rawLen, 1) == Z_OK) {
            result = (uint8_t*)malloc(8 + 25 + otherLen + 12 + zLen + 12);
            size_t o = 8;
            memcpy(result, in, 8);
            o += putChunk(result + o, "IHDR", ihdr, 13);
            memcpy(result + o, other, otherLen); o += otherLen;
            o += putChunk(result + o, "IDAT", z, (uint32_t)zLen);
            o += putChunk(result + o, "IEND", NULL, 0);
            *outLen = o;
        }
        free(z);
    }
    free(raw); free(idat); free(other);
    return result;
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
            uint8_t* buf = NULL;
            if (kind >= 4u) {   /* PNG inputs: half of the iterations are structure-aware */
                size_t mlen = 0;
                uint8_t* m = mutatePng(orig, (size_t)n, &mlen);
                if (m) { buf = (uint8_t*)malloc(mlen ? mlen : 1); memcpy(buf, m, mlen); free(m); len = mlen; }
            }
            const int structured = buf != NULL;
            if (!buf) { buf = (uint8_t*)malloc(len ? len : 1); memcpy(buf, orig, len); }
            if (len && !structured) {
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
