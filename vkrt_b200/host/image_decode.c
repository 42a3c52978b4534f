/* image_decode.c — PNG / JPEG / OpenEXR decoding for textures and environment maps.
 *
 * Mirrors src/core/utility/image.c:86-330 and src/core/utility/exr.cpp:40-302 of the reference, which delegate to libspng,
 * libturbojpeg and tinyexr (none of them vendored here; zlib is the only dependency):
 *   PNG  -> RGBA8 UNORM, or RGBA16 UNORM (host-endian) when the file is 16 bits per sample; tRNS is not applied (spng flags = 0);
 *           16-bit files requested as sRGB are converted to linear and stored LINEAR (image.c:70-84,185-189);
 *   JPEG -> RGBA8 UNORM, alpha 255; baseline and extended sequential Huffman, 8-bit, 1 or 3 components, the IJG "islow" IDCT
 *           (TJFLAG_ACCURATEDCT), triangle-filter ("fancy") chroma upsampling for h2v1 / h2v2 and the IJG YCbCr tables, i.e. the same
 *           integers libjpeg-turbo produces; progressive and arithmetic-coded files are rejected;
 *   EXR  -> single-part scanline files, NONE / RLE / ZIPS / ZIP / PIZ compression, HALF / FLOAT / UINT channels; RGBA16F when every
 *           channel is HALF, else RGBA32F; R,G,B(,A) by name or by ".R" suffix, Y fallback, missing A = 1 (exr.cpp:40-172);
 *           tiled, multi-part, deep, PXR24 / B44 / DWA files are rejected with a message.
 */
#include "image_decode.h"

#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../../include/vkrt_shared.h"

static int imgFail(char* err, size_t errLen, const char* fmt, ...) {
    if (err && errLen) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errLen, fmt, ap);
        va_end(ap);
    }
    return 0;
}

void hostFreeImage(HostImage* image) {
    if (!image) return;
    free(image->pixels);
    memset(image, 0, sizeof(*image));
}

/* ============================================================ PNG ============================================================ */
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* Reverses the per-scanline filters of one (sub)image in place; rows are [filter byte][stride bytes]. */
static int pngUnfilter(uint8_t* data, size_t stride, uint32_t rows, uint32_t bpp) {
    const uint8_t* prev = NULL;
    for (uint32_t y = 0; y < rows; y++) {
        uint8_t* row = data + (size_t)y * (stride + 1);
        const uint8_t filter = row[0];
        uint8_t* cur = row + 1;
        switch (filter) {
            case 0: break;
            case 1: for (size_t i = bpp; i < stride; i++) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
            case 2: if (prev) for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(cur[i] + prev[i]); break;
            case 3:
                for (size_t i = 0; i < stride; i++) {
                    int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
                    cur[i] = (uint8_t)(cur[i] + ((a + b) >> 1));
                }
                break;
            case 4:
                for (size_t i = 0; i < stride; i++) {
                    int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
                    cur[i] = (uint8_t)(cur[i] + paeth(a, b, c));
                }
                break;
            default: return 0;
        }
        prev = cur;
    }
    return 1;
}

typedef struct PngInfo {
    uint32_t width, height;
    uint8_t depth, colorType, interlace;
    uint8_t palette[256][3];
    uint32_t paletteCount;
} PngInfo;

static uint32_t pngChannels(uint8_t colorType) {
    switch (colorType) { case 0: return 1; case 2: return 3; case 3: return 1; case 4: return 2; case 6: return 4; default: return 0; }
}

/* Sample k of a row at `depth` bits (big-endian bit order inside bytes, big-endian 16-bit samples). */
static uint32_t pngSample(const uint8_t* row, uint32_t k, uint8_t depth) {
    switch (depth) {
        case 16: return ((uint32_t)row[2 * k] << 8) | row[2 * k + 1];
        case 8: return row[k];
        case 4: return (row[k >> 1] >> ((1u - (k & 1u)) * 4u)) & 0xfu;
        case 2: return (row[k >> 2] >> ((3u - (k & 3u)) * 2u)) & 0x3u;
        default: return (row[k >> 3] >> (7u - (k & 7u))) & 0x1u;
    }
}

/* Writes pixel x of an unfiltered row into the RGBA8 / RGBA16 output at (ox, oy). */
static void pngStorePixel(const PngInfo* in, const uint8_t* row, uint32_t x, void* out, uint32_t ox, uint32_t oy) {
    const uint32_t ch = pngChannels(in->colorType);
    if (in->depth == 16) {
        uint16_t* px = (uint16_t*)out + ((size_t)oy * in->width + ox) * 4u;
        uint32_t s[4];
        for (uint32_t c = 0; c < ch; c++) s[c] = pngSample(row, x * ch + c, 16);
        switch (in->colorType) {
            case 0: px[0] = px[1] = px[2] = (uint16_t)s[0]; px[3] = 65535u; break;
            case 2: px[0] = (uint16_t)s[0]; px[1] = (uint16_t)s[1]; px[2] = (uint16_t)s[2]; px[3] = 65535u; break;
            case 4: px[0] = px[1] = px[2] = (uint16_t)s[0]; px[3] = (uint16_t)s[1]; break;
            default: px[0] = (uint16_t)s[0]; px[1] = (uint16_t)s[1]; px[2] = (uint16_t)s[2]; px[3] = (uint16_t)s[3]; break;
        }
        return;
    }
    uint8_t* px = (uint8_t*)out + ((size_t)oy * in->width + ox) * 4u;
    if (in->colorType == 3) {
        uint32_t idx = pngSample(row, x, in->depth);
        if (idx >= in->paletteCount) { px[0] = px[1] = px[2] = 0; px[3] = 255; return; }
        px[0] = in->palette[idx][0]; px[1] = in->palette[idx][1]; px[2] = in->palette[idx][2]; px[3] = 255;
        return;
    }
    uint32_t s[4] = {0, 0, 0, 0};
    for (uint32_t c = 0; c < ch; c++) s[c] = pngSample(row, x * ch + c, in->depth);
    if (in->depth < 8) {  /* grey 1/2/4 bits: scale to the full 8-bit range */
        const uint32_t maxv = (1u << in->depth) - 1u;
        s[0] = s[0] * 255u / maxv;
    }
    switch (in->colorType) {
        case 0: px[0] = px[1] = px[2] = (uint8_t)s[0]; px[3] = 255; break;
        case 2: px[0] = (uint8_t)s[0]; px[1] = (uint8_t)s[1]; px[2] = (uint8_t)s[2]; px[3] = 255; break;
        case 4: px[0] = px[1] = px[2] = (uint8_t)s[0]; px[3] = (uint8_t)s[1]; break;
        default: px[0] = (uint8_t)s[0]; px[1] = (uint8_t)s[1]; px[2] = (uint8_t)s[2]; px[3] = (uint8_t)s[3]; break;
    }
}

/* image.c:57-68 srgbDecodeScalar */
static float srgbDecodeScalar(float v) {
    if (v <= 0.04045f) return v / 12.92f;
    return powf((v + 0.055f) / 1.055f, 2.4f);
}

static int decodePng(const uint8_t* data, size_t size, const char* label, uint32_t preferredColorSpace, HostImage* out, char* err, size_t errLen) {
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (size < 8 + 25 || memcmp(data, sig, 8) != 0) return imgFail(err, errLen, "PNG decode from %s failed (bad signature)", label);
    PngInfo in;
    memset(&in, 0, sizeof(in));
    uint8_t* idat = NULL;
    size_t idatLen = 0, idatCap = 0;
    int haveHeader = 0, sawEnd = 0;
    size_t pos = 8;
    while (pos + 12 <= size && !sawEnd) {
        const uint32_t len = be32(data + pos);
        const uint8_t* tag = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if ((size_t)len > size - pos - 12) { free(idat); return imgFail(err, errLen, "PNG decode from %s failed (truncated chunk)", label); }
        if (crc32(crc32(0L, Z_NULL, 0), tag, (uInt)(len + 4)) != be32(body + len)) {
            free(idat);
            return imgFail(err, errLen, "PNG decode from %s failed (chunk CRC mismatch)", label);
        }
        if (!memcmp(tag, "IHDR", 4)) {
            if (len != 13) { free(idat); return imgFail(err, errLen, "PNG header decode from %s failed", label); }
            in.width = be32(body); in.height = be32(body + 4);
            in.depth = body[8]; in.colorType = body[9]; in.interlace = body[12];
            const uint32_t ch = pngChannels(in.colorType);
            int depthOk = in.depth == 8 || in.depth == 16 || (in.colorType == 0 && (in.depth == 1 || in.depth == 2 || in.depth == 4)) ||
                          (in.colorType == 3 && (in.depth == 1 || in.depth == 2 || in.depth == 4));
            if (in.colorType == 3 && in.depth == 16) depthOk = 0;
            if (!ch || !depthOk || body[10] != 0 || body[11] != 0 || in.interlace > 1 || in.width == 0 || in.height == 0 || in.width > (1u << 24) ||
                in.height > (1u << 24)) {
                free(idat);
                return imgFail(err, errLen, "PNG header decode from %s failed (unsupported header)", label);
            }
            haveHeader = 1;
        } else if (!memcmp(tag, "PLTE", 4)) {
            if (len % 3 != 0 || len > 768) { free(idat); return imgFail(err, errLen, "PNG decode from %s failed (bad palette)", label); }
            in.paletteCount = len / 3;
            memcpy(in.palette, body, len);
        } else if (!memcmp(tag, "IDAT", 4)) {
            if (idatLen + len > idatCap) {
                idatCap = (idatLen + len) * 2 + 4096;
                uint8_t* grown = (uint8_t*)realloc(idat, idatCap);
                if (!grown) { free(idat); return imgFail(err, errLen, "Failed to allocate PNG decode buffer for %s", label); }
                idat = grown;
            }
            memcpy(idat + idatLen, body, len);
            idatLen += len;
        } else if (!memcmp(tag, "IEND", 4)) {
            sawEnd = 1;
        }
        pos += 12 + (size_t)len;
    }
    if (!haveHeader || !idatLen) { free(idat); return imgFail(err, errLen, "PNG decode from %s failed (missing IHDR or IDAT)", label); }
    if (in.colorType == 3 && in.paletteCount == 0) { free(idat); return imgFail(err, errLen, "PNG decode from %s failed (missing palette)", label); }

    const uint32_t bitsPerPixel = pngChannels(in.colorType) * in.depth;
    const uint32_t bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;
    /* Adam7 pass geometry (a non-interlaced image is one "pass" covering everything) */
    static const uint8_t px0[7] = {0, 4, 0, 2, 0, 1, 0}, py0[7] = {0, 0, 4, 0, 2, 0, 1}, pdx[7] = {8, 8, 4, 4, 2, 2, 1}, pdy[7] = {8, 8, 8, 4, 4, 2, 2};
    const int passes = in.interlace ? 7 : 1;
    size_t rawSize = 0;
    uint32_t passW[7], passH[7];
    for (int p = 0; p < passes; p++) {
        passW[p] = in.interlace ? (in.width + pdx[p] - 1 - px0[p]) / pdx[p] : in.width;
        passH[p] = in.interlace ? (in.height + pdy[p] - 1 - py0[p]) / pdy[p] : in.height;
        if (in.interlace && (in.width <= px0[p] || in.height <= py0[p])) passW[p] = passH[p] = 0;
        if (passW[p] && passH[p]) rawSize += ((size_t)((passW[p] * (size_t)bitsPerPixel + 7) / 8) + 1) * passH[p];
    }
    uint8_t* raw = (uint8_t*)malloc(rawSize ? rawSize : 1);
    const size_t texel = in.depth == 16 ? 8u : 4u;
    void* pixels = malloc((size_t)in.width * in.height * texel);
    if (!raw || !pixels) { free(raw); free(pixels); free(idat); return imgFail(err, errLen, "Failed to allocate PNG decode buffer for %s", label); }
    uLongf got = (uLongf)rawSize;
    int zr = uncompress(raw, &got, idat, (uLong)idatLen);
    free(idat);
    if ((zr != Z_OK && zr != Z_BUF_ERROR) || got != rawSize) {
        free(raw); free(pixels);
        return imgFail(err, errLen, "PNG decode from %s failed (inflate: %d, %lu of %zu bytes)", label, zr, (unsigned long)got, rawSize);
    }
    uint8_t* cursor = raw;
    for (int p = 0; p < passes; p++) {
        if (!passW[p] || !passH[p]) continue;
        const size_t stride = (passW[p] * (size_t)bitsPerPixel + 7) / 8;
        if (!pngUnfilter(cursor, stride, passH[p], bpp)) { free(raw); free(pixels); return imgFail(err, errLen, "PNG decode from %s failed (bad filter)", label); }
        for (uint32_t y = 0; y < passH[p]; y++) {
            const uint8_t* row = cursor + (size_t)y * (stride + 1) + 1;
            for (uint32_t x = 0; x < passW[p]; x++) {
                const uint32_t ox = in.interlace ? px0[p] + x * pdx[p] : x, oy = in.interlace ? py0[p] + y * pdy[p] : y;
                pngStorePixel(&in, row, x, pixels, ox, oy);
            }
        }
        cursor += (stride + 1) * passH[p];
    }
    free(raw);
    uint32_t storage = preferredColorSpace;
    if (in.depth == 16 && preferredColorSpace == VKRT_TEXTURE_COLOR_SPACE_SRGB) {  /* image.c:70-84 */
        uint16_t* p16 = (uint16_t*)pixels;
        const size_t count = (size_t)in.width * in.height;
        for (size_t i = 0; i < count; i++)
            for (int c = 0; c < 3; c++) {
                float linear = srgbDecodeScalar((float)p16[i * 4 + c] / 65535.0f);
                long q = lrintf(linear * 65535.0f);
                p16[i * 4 + c] = (uint16_t)(q < 0 ? 0 : (q > 65535 ? 65535 : q));
            }
        storage = VKRT_TEXTURE_COLOR_SPACE_LINEAR;
    }
    out->pixels = pixels; out->width = in.width; out->height = in.height;
    out->format = in.depth == 16 ? VKRT_TEXTURE_FORMAT_RGBA16_UNORM : VKRT_TEXTURE_FORMAT_RGBA8_UNORM;
    out->colorSpace = storage;
    return 1;
}

/* ============================================================ JPEG =========================================================== */
/* ITU T.81 baseline / extended sequential Huffman decoder. The arithmetic that decides the output integers follows the published IJG
 * algorithms that libjpeg-turbo implements (and the reference therefore gets): jidctint "islow" 8x8 inverse DCT (CONST_BITS 13,
 * PASS1_BITS 2), h2v1 / h2v2 "fancy" triangle upsampling, and the fixed-point YCbCr -> RGB conversion of jdcolor (SCALEBITS 16). */
typedef struct JHuff {
    uint16_t code[256];
    uint8_t size[256], value[256];
    int count;
    int16_t lookup[512];  /* 9-bit fast table: (length << 8) | value, -1 = longer code */
    int32_t maxcode[18], valptr[17], mincode[17];
} JHuff;

typedef struct JComp {
    int id, h, v, tq, td, ta;
    int dcPred;
    int blocksW, blocksH;   /* allocated plane size in blocks (MCU padded) */
    uint8_t* plane;         /* blocksW*8 x blocksH*8 samples */
} JComp;

typedef struct JDec {
    const uint8_t* data;
    size_t size, pos;
    uint32_t bitBuf;
    int bitCount;
    int hitMarker;
    uint16_t quant[4][64];
    JHuff dc[4], ac[4];
    int haveDc[4], haveAc[4], haveQuant[4];
    JComp comp[3];
    int ncomp, width, height, hmax, vmax, restartInterval;
    int adobeTransform;  /* -1 absent */
} JDec;

static const uint8_t jZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static int jBuildHuff(JHuff* h, const uint8_t* bits, const uint8_t* vals, int total) {
    int k = 0;
    uint32_t code = 0;
    h->count = total;
    for (int len = 1; len <= 16; len++) {
        h->valptr[len] = k;
        h->mincode[len] = (int32_t)code;
        for (int i = 0; i < bits[len - 1]; i++) {
            if (k >= 256) return 0;
            h->size[k] = (uint8_t)len; h->code[k] = (uint16_t)code; h->value[k] = vals[k];
            k++; code++;
        }
        h->maxcode[len] = bits[len - 1] ? (int32_t)code - 1 : -1;
        if (code > (1u << len)) return 0;
        code <<= 1;
    }
    h->maxcode[17] = 0x7fffffff;
    for (int i = 0; i < 512; i++) h->lookup[i] = -1;
    for (int i = 0; i < k; i++) {
        if (h->size[i] <= 9) {
            int shift = 9 - h->size[i];
            int first = h->code[i] << shift;
            for (int j = 0; j < (1 << shift); j++) h->lookup[first + j] = (int16_t)((h->size[i] << 8) | h->value[i]);
        }
    }
    return 1;
}

static void jFill(JDec* d) {
    while (d->bitCount <= 24) {
        uint32_t byte = 0;
        if (!d->hitMarker && d->pos < d->size) {
            byte = d->data[d->pos];
            if (byte == 0xff) {
                uint8_t next = d->pos + 1 < d->size ? d->data[d->pos + 1] : 0xd9;
                if (next == 0x00) d->pos += 2;
                else { d->hitMarker = 1; byte = 0; }
            } else {
                d->pos++;
            }
        }
        d->bitBuf |= byte << (24 - d->bitCount);
        d->bitCount += 8;
    }
}
static inline int jGetBits(JDec* d, int n) {
    if (n == 0) return 0;
    if (d->bitCount < n) jFill(d);
    int v = (int)(d->bitBuf >> (32 - n));
    d->bitBuf <<= n;
    d->bitCount -= n;
    return v;
}
static int jDecodeHuff(JDec* d, const JHuff* h) {
    if (d->bitCount < 16) jFill(d);
    int look = h->lookup[d->bitBuf >> 23];
    if (look >= 0) {
        int len = look >> 8;
        d->bitBuf <<= len;
        d->bitCount -= len;
        return look & 0xff;
    }
    int32_t code = (int32_t)(d->bitBuf >> 22);  /* 10 bits */
    int len = 10;
    while (len <= 16 && (h->maxcode[len] < 0 || code > h->maxcode[len])) {
        len++;
        code = (int32_t)(d->bitBuf >> (32 - len));
    }
    if (len > 16) return -1;
    d->bitBuf <<= len;
    d->bitCount -= len;
    int idx = h->valptr[len] + (code - h->mincode[len]);
    return (idx >= 0 && idx < h->count) ? h->value[idx] : -1;
}
static inline int jExtend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }

/* jidctint.c "islow": 13-bit constants, two passes, intermediate scaled by 2^PASS1_BITS */
#define J_FIX_0_298631336 2446
#define J_FIX_0_390180644 3196
#define J_FIX_0_541196100 4433
#define J_FIX_0_765366865 6270
#define J_FIX_0_899976223 7373
#define J_FIX_1_175875602 9633
#define J_FIX_1_501321110 12299
#define J_FIX_1_847759065 15137
#define J_FIX_1_961570560 16069
#define J_FIX_2_053119869 16819
#define J_FIX_2_562915447 20995
#define J_FIX_3_072711026 25172
#define J_DESCALE(x, n) (((x) + ((int64_t)1 << ((n) - 1))) >> (n))
static inline uint8_t jClampSample(int64_t v) {
    v += 128;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
static void jIdctIslow(const int16_t* coef, const uint16_t* quant, uint8_t* out, int stride) {
    int ws[64];
    for (int c = 0; c < 8; c++) {
        const int16_t* in = coef + c;
        const uint16_t* q = quant + c;
        int* w = ws + c;
        if (!in[8] && !in[16] && !in[24] && !in[32] && !in[40] && !in[48] && !in[56]) {
            int64_t dc = ((int64_t)in[0] * q[0]) * 4;
            for (int r = 0; r < 8; r++) w[r * 8] = (int)dc;
            continue;
        }
        int64_t z2 = (int64_t)in[16] * q[16], z3 = (int64_t)in[48] * q[48];
        int64_t z1 = (z2 + z3) * J_FIX_0_541196100;
        int64_t tmp2 = z1 + z3 * (-J_FIX_1_847759065);
        int64_t tmp3 = z1 + z2 * J_FIX_0_765366865;
        z2 = (int64_t)in[0] * q[0]; z3 = (int64_t)in[32] * q[32];
        int64_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
        int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = (int64_t)in[56] * q[56]; tmp1 = (int64_t)in[40] * q[40]; tmp2 = (int64_t)in[24] * q[24]; tmp3 = (int64_t)in[8] * q[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int64_t z4 = tmp1 + tmp3;
        int64_t z5 = (z3 + z4) * J_FIX_1_175875602;
        tmp0 *= J_FIX_0_298631336; tmp1 *= J_FIX_2_053119869; tmp2 *= J_FIX_3_072711026; tmp3 *= J_FIX_1_501321110;
        z1 *= -J_FIX_0_899976223; z2 *= -J_FIX_2_562915447; z3 *= -J_FIX_1_961570560; z4 *= -J_FIX_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        w[0] = J_DESCALE(tmp10 + tmp3, 11); w[56] = J_DESCALE(tmp10 - tmp3, 11);
        w[8] = J_DESCALE(tmp11 + tmp2, 11); w[48] = J_DESCALE(tmp11 - tmp2, 11);
        w[16] = J_DESCALE(tmp12 + tmp1, 11); w[40] = J_DESCALE(tmp12 - tmp1, 11);
        w[24] = J_DESCALE(tmp13 + tmp0, 11); w[32] = J_DESCALE(tmp13 - tmp0, 11);
    }
    for (int r = 0; r < 8; r++) {
        const int* w = ws + r * 8;
        uint8_t* o = out + r * stride;
        if (!w[1] && !w[2] && !w[3] && !w[4] && !w[5] && !w[6] && !w[7]) {
            uint8_t dcs = jClampSample(J_DESCALE(w[0], 5));
            for (int c = 0; c < 8; c++) o[c] = dcs;
            continue;
        }
        int64_t z2 = w[2], z3 = w[6];
        int64_t z1 = (z2 + z3) * J_FIX_0_541196100;
        int64_t tmp2 = z1 + z3 * (-J_FIX_1_847759065);
        int64_t tmp3 = z1 + z2 * J_FIX_0_765366865;
        int64_t tmp0 = ((int64_t)w[0] + w[4]) * 8192, tmp1 = ((int64_t)w[0] - w[4]) * 8192;
        int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int64_t z4 = tmp1 + tmp3;
        int64_t z5 = (z3 + z4) * J_FIX_1_175875602;
        tmp0 *= J_FIX_0_298631336; tmp1 *= J_FIX_2_053119869; tmp2 *= J_FIX_3_072711026; tmp3 *= J_FIX_1_501321110;
        z1 *= -J_FIX_0_899976223; z2 *= -J_FIX_2_562915447; z3 *= -J_FIX_1_961570560; z4 *= -J_FIX_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = jClampSample(J_DESCALE(tmp10 + tmp3, 18)); o[7] = jClampSample(J_DESCALE(tmp10 - tmp3, 18));
        o[1] = jClampSample(J_DESCALE(tmp11 + tmp2, 18)); o[6] = jClampSample(J_DESCALE(tmp11 - tmp2, 18));
        o[2] = jClampSample(J_DESCALE(tmp12 + tmp1, 18)); o[5] = jClampSample(J_DESCALE(tmp12 - tmp1, 18));
        o[3] = jClampSample(J_DESCALE(tmp13 + tmp0, 18)); o[4] = jClampSample(J_DESCALE(tmp13 - tmp0, 18));
    }
}

static int jDecodeBlock(JDec* d, JComp* c, int16_t* coef) {
    memset(coef, 0, 64 * sizeof(int16_t));
    int t = jDecodeHuff(d, &d->dc[c->td]);
    if (t < 0 || t > 16) return 0;
    int diff = t ? jExtend(jGetBits(d, t), t) : 0;
    c->dcPred = (int)((unsigned)c->dcPred + (unsigned)diff);  /* wraps instead of overflowing on corrupt streams */
    coef[0] = (int16_t)c->dcPred;
    for (int k = 1; k < 64;) {
        int rs = jDecodeHuff(d, &d->ac[c->ta]);
        if (rs < 0) return 0;
        int r = rs >> 4, s = rs & 15;
        if (s == 0) {
            if (r != 15) break;
            k += 16;
            continue;
        }
        k += r;
        if (k > 63) return 0;
        coef[jZigzag[k]] = (int16_t)jExtend(jGetBits(d, s), s);
        k++;
    }
    return 1;
}

/* jdsample.c h2v1_fancy_upsample: one output row from one input row */
static void jUpsampleH2Fancy(const uint8_t* in, int inWidth, uint8_t* out) {
    if (inWidth == 1) { out[0] = out[1] = in[0]; return; }
    out[0] = in[0];
    out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
    for (int i = 1; i < inWidth - 1; i++) {
        int v = in[i] * 3;
        out[2 * i] = (uint8_t)((v + in[i - 1] + 1) >> 2);
        out[2 * i + 1] = (uint8_t)((v + in[i + 1] + 2) >> 2);
    }
    int v = in[inWidth - 1] * 3;
    out[2 * (inWidth - 1)] = (uint8_t)((v + in[inWidth - 2] + 1) >> 2);
    out[2 * (inWidth - 1) + 1] = in[inWidth - 1];
}
/* jdsample.c h2v2_fancy_upsample: output row from the nearer (in0) and the farther (in1) input row */
static void jUpsampleH2V2Fancy(const uint8_t* in0, const uint8_t* in1, int inWidth, uint8_t* out) {
    int thisSum = in0[0] * 3 + in1[0];
    if (inWidth == 1) { out[0] = (uint8_t)((thisSum * 4 + 8) >> 4); out[1] = (uint8_t)((thisSum * 4 + 7) >> 4); return; }
    int nextSum = in0[1] * 3 + in1[1];
    out[0] = (uint8_t)((thisSum * 4 + 8) >> 4);
    out[1] = (uint8_t)((thisSum * 3 + nextSum + 7) >> 4);
    int lastSum = thisSum;
    thisSum = nextSum;
    for (int i = 1; i < inWidth - 1; i++) {
        nextSum = in0[i + 1] * 3 + in1[i + 1];
        out[2 * i] = (uint8_t)((thisSum * 3 + lastSum + 8) >> 4);
        out[2 * i + 1] = (uint8_t)((thisSum * 3 + nextSum + 7) >> 4);
        lastSum = thisSum;
        thisSum = nextSum;
    }
    out[2 * (inWidth - 1)] = (uint8_t)((thisSum * 3 + lastSum + 8) >> 4);
    out[2 * (inWidth - 1) + 1] = (uint8_t)((thisSum * 4 + 7) >> 4);
}

static int decodeJpeg(const uint8_t* data, size_t size, const char* label, uint32_t preferredColorSpace, HostImage* out, char* err, size_t errLen) {
    JDec* d = (JDec*)calloc(1, sizeof(JDec));
    if (!d) return imgFail(err, errLen, "JPEG decoder initialization failed for %s", label);
    d->data = data; d->size = size; d->adobeTransform = -1;
    int ok = 0, sawFrame = 0, sawScan = 0;
    size_t pos = 2;
    const char* why = "no image data";
    if (size < 4 || data[0] != 0xff || data[1] != 0xd8) { why = "not a JPEG stream"; goto done; }
    while (pos + 4 <= size && !sawScan) {
        if (data[pos] != 0xff) { pos++; continue; }
        const uint8_t marker = data[pos + 1];
        if (marker == 0xff) { pos++; continue; }
        if (marker == 0xd8 || (marker >= 0xd0 && marker <= 0xd7) || marker == 0x01) { pos += 2; continue; }
        if (marker == 0xd9) break;
        const size_t len = ((size_t)data[pos + 2] << 8) | data[pos + 3];
        if (len < 2 || pos + 2 + len > size) { why = "truncated segment"; goto done; }
        const uint8_t* seg = data + pos + 4;
        const size_t segLen = len - 2;
        if (marker == 0xdb) {  /* DQT */
            size_t p = 0;
            while (p < segLen) {
                const int pq = seg[p] >> 4, tq = seg[p] & 15;
                p++;
                if (tq > 3 || p + (pq ? 128u : 64u) > segLen) { why = "bad quantisation table"; goto done; }
                for (int i = 0; i < 64; i++) {
                    d->quant[tq][jZigzag[i]] = pq ? (uint16_t)((seg[p] << 8) | seg[p + 1]) : seg[p];
                    p += pq ? 2 : 1;
                }
                d->haveQuant[tq] = 1;
            }
        } else if (marker == 0xc4) {  /* DHT */
            size_t p = 0;
            while (p + 17 <= segLen) {
                const int tc = seg[p] >> 4, th = seg[p] & 15;
                int total = 0;
                for (int i = 0; i < 16; i++) total += seg[p + 1 + i];
                if (tc > 1 || th > 3 || total > 256 || p + 17 + (size_t)total > segLen) { why = "bad Huffman table"; goto done; }
                if (!jBuildHuff(tc ? &d->ac[th] : &d->dc[th], seg + p + 1, seg + p + 17, total)) { why = "bad Huffman table"; goto done; }
                if (tc) d->haveAc[th] = 1; else d->haveDc[th] = 1;
                p += 17 + (size_t)total;
            }
        } else if (marker == 0xc0 || marker == 0xc1) {  /* SOF0 / SOF1 */
            if (segLen < 6 || seg[0] != 8) { why = "only 8-bit samples are supported"; goto done; }
            d->height = (seg[1] << 8) | seg[2];
            d->width = (seg[3] << 8) | seg[4];
            d->ncomp = seg[5];
            if ((d->ncomp != 1 && d->ncomp != 3) || segLen < 6u + 3u * (size_t)d->ncomp || d->width <= 0 || d->height <= 0) { why = "unsupported component count or size"; goto done; }
            for (int i = 0; i < d->ncomp; i++) {
                d->comp[i].id = seg[6 + 3 * i];
                d->comp[i].h = seg[7 + 3 * i] >> 4;
                d->comp[i].v = seg[7 + 3 * i] & 15;
                d->comp[i].tq = seg[8 + 3 * i];
                if (d->comp[i].h < 1 || d->comp[i].h > 2 || d->comp[i].v < 1 || d->comp[i].v > 2 || d->comp[i].tq > 3) { why = "unsupported sampling factors"; goto done; }
                if (d->comp[i].h > d->hmax) d->hmax = d->comp[i].h;
                if (d->comp[i].v > d->vmax) d->vmax = d->comp[i].v;
            }
            if (d->ncomp == 1) { d->comp[0].h = d->comp[0].v = 1; d->hmax = d->vmax = 1; }
            sawFrame = 1;
        } else if (marker == 0xc2 || (marker >= 0xc5 && marker <= 0xcf && marker != 0xc8 && marker != 0xcc) || marker == 0xc3) {
            why = marker == 0xc2 ? "progressive JPEG is not supported" : "unsupported JPEG process (lossless / arithmetic / hierarchical)";
            goto done;
        } else if (marker == 0xdd) {
            if (segLen >= 2) d->restartInterval = (seg[0] << 8) | seg[1];
        } else if (marker == 0xee) {
            if (segLen >= 12 && !memcmp(seg, "Adobe", 5)) d->adobeTransform = seg[11];
        } else if (marker == 0xda) {  /* SOS */
            if (!sawFrame) { why = "scan before frame header"; goto done; }
            const int ns = seg[0];
            if (ns != d->ncomp || segLen < 1u + 2u * (size_t)ns + 3u) { why = "non-interleaved scans are not supported"; goto done; }
            for (int i = 0; i < ns; i++) {
                int found = -1;
                for (int k = 0; k < d->ncomp; k++) if (d->comp[k].id == seg[1 + 2 * i]) found = k;
                if (found < 0) { why = "scan names an unknown component"; goto done; }
                d->comp[found].td = seg[2 + 2 * i] >> 4;
                d->comp[found].ta = seg[2 + 2 * i] & 15;
                if (d->comp[found].td > 3 || d->comp[found].ta > 3 || !d->haveDc[d->comp[found].td] || !d->haveAc[d->comp[found].ta] || !d->haveQuant[d->comp[found].tq]) {
                    why = "scan refers to a missing table";
                    goto done;
                }
            }
            sawScan = 1;
        }
        pos += 2 + len;
    }
    if (!sawScan) goto done;
    {
        const int mcuW = 8 * d->hmax, mcuH = 8 * d->vmax;
        const int mcusX = (d->width + mcuW - 1) / mcuW, mcusY = (d->height + mcuH - 1) / mcuH;
        for (int i = 0; i < d->ncomp; i++) {
            d->comp[i].blocksW = mcusX * d->comp[i].h;
            d->comp[i].blocksH = mcusY * d->comp[i].v;
            d->comp[i].plane = (uint8_t*)malloc((size_t)d->comp[i].blocksW * 8 * d->comp[i].blocksH * 8);
            if (!d->comp[i].plane) { why = "out of memory"; goto done; }
        }
        d->pos = pos;
        int16_t coef[64];
        int restartCountdown = d->restartInterval;
        for (int my = 0; my < mcusY; my++) {
            for (int mx = 0; mx < mcusX; mx++) {
                if (d->restartInterval && restartCountdown == 0) {
                    /* align to the RSTn marker and reset the predictors */
                    d->bitBuf = 0; d->bitCount = 0; d->hitMarker = 0;
                    while (d->pos + 1 < d->size && !(d->data[d->pos] == 0xff && d->data[d->pos + 1] >= 0xd0 && d->data[d->pos + 1] <= 0xd7)) d->pos++;
                    if (d->pos + 1 < d->size) d->pos += 2;
                    for (int i = 0; i < d->ncomp; i++) d->comp[i].dcPred = 0;
                    restartCountdown = d->restartInterval;
                }
                for (int i = 0; i < d->ncomp; i++) {
                    JComp* c = &d->comp[i];
                    const int stride = c->blocksW * 8;
                    for (int by = 0; by < c->v; by++)
                        for (int bx = 0; bx < c->h; bx++) {
                            if (!jDecodeBlock(d, c, coef)) { why = "corrupt entropy-coded data"; goto done; }
                            uint8_t* dst = c->plane + (size_t)((my * c->v + by) * 8) * stride + (size_t)(mx * c->h + bx) * 8;
                            jIdctIslow(coef, d->quant[c->tq], dst, stride);
                        }
                }
                if (d->restartInterval) restartCountdown--;
            }
        }
        /* colour: upsample chroma to full resolution, then YCbCr -> RGB (jdcolor.c tables) */
        const int W = d->width, H = d->height;
        uint8_t* rgba = (uint8_t*)malloc((size_t)W * H * 4);
        uint8_t* rows[3] = {NULL, NULL, NULL};
        if (!rgba) { why = "out of memory"; goto done; }
        int colorTransform = d->ncomp == 3;
        if (d->ncomp == 3 && d->adobeTransform == 0) colorTransform = 0;
        if (d->ncomp == 3 && d->adobeTransform < 0 && d->comp[0].id == 'R' && d->comp[1].id == 'G' && d->comp[2].id == 'B') colorTransform = 0;
        for (int i = 0; i < d->ncomp; i++) {
            rows[i] = (uint8_t*)malloc((size_t)d->comp[i].blocksW * 16 + 16);
            if (!rows[i]) { free(rgba); free(rows[0]); free(rows[1]); free(rows[2]); why = "out of memory"; goto done; }
        }
        for (int y = 0; y < H; y++) {
            const uint8_t* line[3] = {NULL, NULL, NULL};
            for (int i = 0; i < d->ncomp; i++) {
                JComp* c = &d->comp[i];
                const int stride = c->blocksW * 8;
                /* downsampled_width as libjpeg computes it: ceil(width * h / hmax) */
                const int cw = (W * c->h + d->hmax - 1) / d->hmax, chh = (H * c->v + d->vmax - 1) / d->vmax;
                const int hs = d->hmax / c->h, vs = d->vmax / c->v;
                if (hs == 1 && vs == 1) {
                    line[i] = c->plane + (size_t)y * stride;
                } else if (hs == 2 && vs == 1) {
                    jUpsampleH2Fancy(c->plane + (size_t)y * stride, cw, rows[i]);
                    line[i] = rows[i];
                } else if (hs == 2 && vs == 2) {
                    const int cy = y >> 1;
                    int other = (y & 1) ? cy + 1 : cy - 1;   /* the farther row: below for odd output rows, above for even ones */
                    if (other < 0) other = 0;
                    if (other > chh - 1) other = chh - 1;
                    jUpsampleH2V2Fancy(c->plane + (size_t)cy * stride, c->plane + (size_t)other * stride, cw, rows[i]);
                    line[i] = rows[i];
                } else {  /* h1v2: libjpeg-turbo uses its own fancy vertical filter; replicate rows (rare layout) */
                    const int cy = y / vs;
                    const uint8_t* src = c->plane + (size_t)cy * stride;
                    for (int x = 0; x < W; x++) rows[i][x] = src[x / hs];
                    line[i] = rows[i];
                }
            }
            uint8_t* o = rgba + (size_t)y * W * 4;
            if (d->ncomp == 1) {
                for (int x = 0; x < W; x++) { o[4 * x] = o[4 * x + 1] = o[4 * x + 2] = line[0][x]; o[4 * x + 3] = 255; }
            } else if (!colorTransform) {
                for (int x = 0; x < W; x++) { o[4 * x] = line[0][x]; o[4 * x + 1] = line[1][x]; o[4 * x + 2] = line[2][x]; o[4 * x + 3] = 255; }
            } else {
                for (int x = 0; x < W; x++) {
                    const int Y = line[0][x], cb = line[1][x] - 128, cr = line[2][x] - 128;
                    /* jdcolor.c: Cr_r = (FIX(1.40200)*x + ONE_HALF) >> 16, Cb_b likewise, Cr_g = -FIX(0.71414)*x, Cb_g = -FIX(0.34414)*x + ONE_HALF */
                    const int r = Y + ((91881 * cr + 32768) >> 16);
                    const int g = Y + ((-22554 * cb - 46802 * cr + 32768) >> 16);
                    const int b = Y + ((116130 * cb + 32768) >> 16);
                    o[4 * x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
                    o[4 * x + 1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
                    o[4 * x + 2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
                    o[4 * x + 3] = 255;
                }
            }
        }
        free(rows[0]); free(rows[1]); free(rows[2]);
        out->pixels = rgba; out->width = (uint32_t)W; out->height = (uint32_t)H;
        out->format = VKRT_TEXTURE_FORMAT_RGBA8_UNORM; out->colorSpace = preferredColorSpace;
        ok = 1;
    }
done:
    for (int i = 0; i < 3; i++) free(d->comp[i].plane);
    free(d);
    if (!ok) return imgFail(err, errLen, "JPEG decode from %s failed (%s)", label, why);
    return 1;
}

/* ============================================================ EXR ============================================================ */
typedef struct ExrChannel { char name[64]; int type; int xs, ys; size_t offsetInLine; } ExrChannel;

static int exrChannelMatches(const char* channel, const char* component) {  /* exr.cpp:40-46 */
    if (!strcmp(channel, component)) return 1;
    const char* suffix = strrchr(channel, '.');
    return suffix && !strcmp(suffix + 1, component);
}
static float exrHalfToFloat(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { e++; man <<= 1; } while ((man & 0x400u) == 0);
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static uint16_t exrFloatToHalf(float f) {  /* round to nearest even, as tinyexr's float_to_half_full does for representable ranges */
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xffu) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xffu) == 0xffu) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0u));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t half = man >> shift, rem = man & ((1u << shift) - 1u), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1u))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13), rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
    return (uint16_t)(sign | half);
}

/* Undoes OpenEXR's zip/rle pre-processing: delta predictor, then de-interleave of the two byte halves. */
static void exrUnpredict(uint8_t* tmp, size_t n, uint8_t* outBuf) {
    for (size_t i = 1; i < n; i++) tmp[i] = (uint8_t)(tmp[i - 1] + tmp[i] - 128);
    const uint8_t* t1 = tmp;
    const uint8_t* t2 = tmp + (n + 1) / 2;
    size_t o = 0;
    while (o < n) {
        outBuf[o++] = *t1++;
        if (o < n) outBuf[o++] = *t2++;
    }
}
static int exrRleDecode(const uint8_t* in, size_t inLen, uint8_t* outBuf, size_t outLen) {
    size_t i = 0, o = 0;
    while (i < inLen) {
        int8_t count = (int8_t)in[i++];
        if (count < 0) {
            size_t n = (size_t)(-count);
            if (i + n > inLen || o + n > outLen) return 0;
            memcpy(outBuf + o, in + i, n);
            i += n; o += n;
        } else {
            size_t n = (size_t)count + 1;
            if (i >= inLen || o + n > outLen) return 0;
            memset(outBuf + o, in[i++], n);
            o += n;
        }
    }
    return o == outLen;
}


/* ---- PIZ (OpenEXR's default for photographic data): a 16-bit value LUT from a presence bitmap, a 2-D Haar-like wavelet per channel
 * plane and a canonical Huffman code with run-length escapes, as published in the OpenEXR technical documentation (ImfPizCompressor,
 * ImfHuf, ImfWav). ---------------------------------------------------------------------------------------------------------------- */
#define PIZ_HUF_ENCSIZE 65537          /* 65536 values + the run-length symbol */
typedef struct PizBits { const uint8_t* p; const uint8_t* end; uint64_t acc; int count; } PizBits;
static inline uint32_t pizGetBits(PizBits* b, int n) {
    while (b->count < n) {
        b->acc = (b->acc << 8) | (b->p < b->end ? *b->p : 0u);
        b->p++;
        b->count += 8;
    }
    b->count -= n;
    return (uint32_t)((b->acc >> b->count) & ((1ull << n) - 1ull));
}
static int pizHufDecode(const uint8_t* data, size_t size, uint16_t* out, size_t outCount) {
    if (size < 20) return 0;
    uint32_t im, iM, nBits;
    memcpy(&im, data, 4); memcpy(&iM, data + 4, 4); memcpy(&nBits, data + 12, 4);
    if (im >= PIZ_HUF_ENCSIZE || iM >= PIZ_HUF_ENCSIZE || im > iM) return 0;
    uint8_t* len = (uint8_t*)calloc(PIZ_HUF_ENCSIZE, 1);
    uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * PIZ_HUF_ENCSIZE);
    if (!len || !sorted) { free(len); free(sorted); return 0; }
    PizBits b = {data + 20, data + size, 0, 0};
    /* packed code lengths: 6 bits each, 59..62 = short zero runs (2..5), 63 = long zero run (8 more bits + 6) */
    for (uint32_t s = im; s <= iM; s++) {
        uint32_t l = pizGetBits(&b, 6);
        if (l == 63u) {
            uint32_t run = pizGetBits(&b, 8) + 6u;
            if (s + run > iM + 1u) { free(len); free(sorted); return 0; }
            s += run - 1u;
        } else if (l >= 59u) {
            uint32_t run = l - 59u + 2u;
            if (s + run > iM + 1u) { free(len); free(sorted); return 0; }
            s += run - 1u;
        } else {
            len[s] = (uint8_t)l;
        }
    }
    /* canonical codes: lengths 58..1, the first code of a length follows from the longer ones; symbols of one length in index order */
    uint64_t count[59] = {0}, first[59] = {0};
    uint32_t start[60] = {0};
    for (uint32_t s = im; s <= iM; s++) count[len[s]]++;
    uint64_t c = 0;
    for (int l = 58; l > 0; l--) { uint64_t nc = (c + count[l]) >> 1; first[l] = c; c = nc; }
    for (int l = 1; l <= 58; l++) start[l + 1] = start[l] + (uint32_t)count[l];
    { uint32_t fill[60]; memcpy(fill, start, sizeof(fill)); for (uint32_t s = im; s <= iM; s++) if (len[s]) sorted[fill[len[s]]++] = s; }
    /* the bit stream starts at the next byte boundary after the table */
    const uint8_t* bitsStart = b.p - (b.count / 8);
    PizBits d = {bitsStart, data + size, 0, 0};
    uint64_t bitsLeft = nBits;
    size_t o = 0;
    int ok = 1;
    while (bitsLeft > 0 && ok) {
        uint64_t code = 0;
        int l = 0;
        uint32_t sym = 0xffffffffu;
        while (l < 58 && bitsLeft > 0) {
            code = (code << 1) | pizGetBits(&d, 1);
            bitsLeft--;
            l++;
            if (count[l] && code >= first[l] && code - first[l] < count[l]) { sym = sorted[start[l] + (uint32_t)(code - first[l])]; break; }
        }
        if (sym == 0xffffffffu) {   /* trailing padding bits of the last byte are not a symbol */
            break;
        }
        if (sym == iM) {            /* run-length symbol: repeat the previous value */
            if (bitsLeft < 8 || o == 0) { ok = 0; break; }
            uint32_t run = pizGetBits(&d, 8);
            bitsLeft -= 8;
            if (o + run > outCount) { ok = 0; break; }
            for (uint32_t k = 0; k < run; k++) out[o + k] = out[o - 1];
            o += run;
        } else {
            if (o >= outCount) { ok = 0; break; }
            out[o++] = (uint16_t)sym;
        }
    }
    free(len); free(sorted);
    return ok && o == outCount;
}
static inline void pizWdec14(uint16_t l, uint16_t h, uint16_t* a, uint16_t* b) {
    int ls = (int16_t)l, hs = (int16_t)h;
    int ai = ls + (hs & 1) + (hs >> 1);
    *a = (uint16_t)(int16_t)ai;
    *b = (uint16_t)(int16_t)(ai - hs);
}
static inline void pizWdec16(uint16_t l, uint16_t h, uint16_t* a, uint16_t* b) {
    int m = l, d = h;
    int bb = (m - (d >> 1)) & 0xffff;
    int aa = (d + bb - 0x8000) & 0xffff;
    *b = (uint16_t)bb;
    *a = (uint16_t)aa;
}
static void pizWav2Decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const int w14 = mx < (1 << 14);
    int n = nx > ny ? ny : nx, p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1;
    p2 = p;
    p >>= 1;
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* ey = in + (ptrdiff_t)oy * (ny - p2);
        const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                uint16_t* p10 = px + oy1;
                uint16_t* p11 = p10 + ox1;
                if (w14) { pizWdec14(*px, *p10, &i00, &i10); pizWdec14(*p01, *p11, &i01, &i11); pizWdec14(i00, i01, px, p01); pizWdec14(i10, i11, p10, p11); }
                else { pizWdec16(*px, *p10, &i00, &i10); pizWdec16(*p01, *p11, &i01, &i11); pizWdec16(i00, i01, px, p01); pizWdec16(i10, i11, p10, p11); }
            }
            if (nx & p) {
                uint16_t* p10 = px + oy1;
                if (w14) pizWdec14(*px, *p10, &i00, p10); else pizWdec16(*px, *p10, &i00, p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t* px = py;
            uint16_t* ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                if (w14) pizWdec14(*px, *p01, &i00, p01); else pizWdec16(*px, *p01, &i00, p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}
/* One PIZ block -> the same scanline-interleaved bytes an uncompressed block holds. wordsPerPixel[c] = 1 (HALF) or 2 (FLOAT / UINT). */
static int exrPizDecode(const uint8_t* src, size_t srcLen, uint8_t* outBuf, size_t rawLen, uint32_t width, uint32_t lines, const int* wordsPerPixel, int nch) {
    const size_t totalWords = rawLen / 2;
    if (srcLen < 4) return 0;
    uint16_t minNonZero, maxNonZero;
    memcpy(&minNonZero, src, 2); memcpy(&maxNonZero, src + 2, 2);
    uint8_t* bitmap = (uint8_t*)calloc(8192, 1);
    uint16_t* lut = (uint16_t*)calloc(65536, sizeof(uint16_t));
    uint16_t* tmp = (uint16_t*)malloc(totalWords * sizeof(uint16_t) + 2);
    int ok = bitmap && lut && tmp;
    size_t pos = 4;
    if (ok && minNonZero <= maxNonZero) {
        size_t nb = (size_t)maxNonZero - minNonZero + 1;
        if (maxNonZero >= 8192 || pos + nb > srcLen) ok = 0;
        else { memcpy(bitmap + minNonZero, src + pos, nb); pos += nb; }
    }
    uint16_t maxValue = 0;
    if (ok) {
        uint32_t k = 0;
        for (uint32_t i = 0; i < 65536u; i++)
            if (i == 0 || (bitmap[i >> 3] & (1u << (i & 7u)))) lut[k++] = (uint16_t)i;
        maxValue = (uint16_t)(k - 1u);
        int32_t hufLen = 0;
        if (pos + 4 > srcLen) ok = 0;
        else { memcpy(&hufLen, src + pos, 4); pos += 4; }
        if (ok && (hufLen < 0 || pos + (size_t)hufLen > srcLen)) ok = 0;
        if (ok) ok = pizHufDecode(src + pos, (size_t)hufLen, tmp, totalWords);
    }
    if (ok) {
        uint16_t* plane = tmp;
        for (int c = 0; c < nch; c++) {
            for (int j = 0; j < wordsPerPixel[c]; j++) pizWav2Decode(plane + j, (int)width, wordsPerPixel[c], (int)lines, (int)width * wordsPerPixel[c], maxValue);
            plane += (size_t)width * lines * wordsPerPixel[c];
        }
        for (size_t i = 0; i < totalWords; i++) tmp[i] = lut[tmp[i]];
        /* channel-planar -> per scanline, channels in order */
        uint16_t* o = (uint16_t*)outBuf;
        for (uint32_t y = 0; y < lines; y++) {
            const uint16_t* chanBase = tmp;
            for (int c = 0; c < nch; c++) {
                const size_t rowWords = (size_t)width * wordsPerPixel[c];
                memcpy(o, chanBase + (size_t)y * rowWords, rowWords * 2);
                o += rowWords;
                chanBase += rowWords * lines;
            }
        }
    }
    free(bitmap); free(lut); free(tmp);
    return ok;
}

static int decodeExr(const uint8_t* data, size_t size, const char* label, HostImage* out, char* err, size_t errLen) {
    if (size < 8 || data[0] != 0x76 || data[1] != 0x2f || data[2] != 0x31 || data[3] != 0x01) return imgFail(err, errLen, "Invalid EXR file: %s", label);
    const uint32_t versionField = (uint32_t)data[4] | ((uint32_t)data[5] << 8) | ((uint32_t)data[6] << 16) | ((uint32_t)data[7] << 24);
    if ((versionField & 0xffu) != 2) return imgFail(err, errLen, "Invalid EXR file: %s", label);
    if (versionField & 0x200u) return imgFail(err, errLen, "EXR header decode from %s failed (tiled images are not supported)", label);
    if (versionField & 0x1800u) return imgFail(err, errLen, "EXR header decode from %s failed (deep / multi-part files are not supported)", label);
    ExrChannel ch[16];
    int nch = 0, compression = -1, haveWindow = 0;
    int32_t xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    size_t pos = 8;
    for (;;) {  /* attributes: name\0 type\0 int32 size, data; an empty name ends the header */
        if (pos >= size) return imgFail(err, errLen, "EXR header decode from %s failed (truncated header)", label);
        if (data[pos] == 0) { pos++; break; }
        const char* name = (const char*)data + pos;
        size_t nl = strnlen(name, size - pos);
        if (pos + nl + 1 >= size) return imgFail(err, errLen, "EXR header decode from %s failed (truncated header)", label);
        const char* type = name + nl + 1;
        size_t tl = strnlen(type, size - (pos + nl + 1));
        size_t p = pos + nl + 1 + tl + 1;
        if (p + 4 > size) return imgFail(err, errLen, "EXR header decode from %s failed (truncated header)", label);
        uint32_t asz;
        memcpy(&asz, data + p, 4);
        p += 4;
        if ((size_t)asz > size - p) return imgFail(err, errLen, "EXR header decode from %s failed (truncated attribute)", label);
        const uint8_t* a = data + p;
        if (!strcmp(name, "channels") && !strcmp(type, "chlist")) {
            size_t q = 0;
            while (q < asz && a[q] != 0) {
                size_t cl = strnlen((const char*)a + q, asz - q);
                if (q + cl + 1 + 16 > asz || nch >= 16 || cl >= sizeof(ch[0].name)) return imgFail(err, errLen, "EXR header decode from %s failed (bad channel list)", label);
                memcpy(ch[nch].name, a + q, cl + 1);
                q += cl + 1;
                int32_t vals[4];
                memcpy(&vals[0], a + q, 4); memcpy(&vals[2], a + q + 8, 4); memcpy(&vals[3], a + q + 12, 4);
                ch[nch].type = vals[0]; ch[nch].xs = vals[2]; ch[nch].ys = vals[3];
                q += 16;
                nch++;
            }
        } else if (!strcmp(name, "compression") && asz >= 1) {
            compression = a[0];
        } else if (!strcmp(name, "dataWindow") && asz >= 16) {
            memcpy(&xmin, a, 4); memcpy(&ymin, a + 4, 4); memcpy(&xmax, a + 8, 4); memcpy(&ymax, a + 12, 4);
            haveWindow = 1;
        }
        pos = p + asz;
    }
    if (!nch || compression < 0 || !haveWindow || xmax < xmin || ymax < ymin) return imgFail(err, errLen, "EXR header decode from %s failed (missing attributes)", label);
    if (compression > 4) return imgFail(err, errLen, "EXR decode from %s failed (compression %d: only NONE, RLE, ZIPS, ZIP and PIZ are supported)", label, compression);
    const int64_t W64 = (int64_t)xmax - xmin + 1, H64 = (int64_t)ymax - ymin + 1;  /* 64-bit: corrupt windows span the whole int range */
    if (W64 > (1 << 20) || H64 > (1 << 20)) return imgFail(err, errLen, "EXR image dimensions overflow for %s", label);
    const uint32_t W = (uint32_t)W64, H = (uint32_t)H64;
    size_t lineBytes = 0;
    int allHalf = 1;
    for (int i = 0; i < nch; i++) {
        if (ch[i].xs != 1 || ch[i].ys != 1) return imgFail(err, errLen, "EXR decode from %s failed (subsampled channels are not supported)", label);
        if (ch[i].type < 0 || ch[i].type > 2) return imgFail(err, errLen, "EXR decode from %s failed (bad pixel type)", label);
        ch[i].offsetInLine = lineBytes;
        lineBytes += (size_t)W * (ch[i].type == 1 ? 2u : 4u);
        if (ch[i].type != 1) allHalf = 0;
    }
    int iR = -1, iG = -1, iB = -1, iA = -1, iY = -1;
    for (int i = nch - 1; i >= 0; i--) {  /* first match wins, as queryChannelIndex scans upwards */
        if (exrChannelMatches(ch[i].name, "R")) iR = i;
        if (exrChannelMatches(ch[i].name, "G")) iG = i;
        if (exrChannelMatches(ch[i].name, "B")) iB = i;
        if (exrChannelMatches(ch[i].name, "A")) iA = i;
        if (exrChannelMatches(ch[i].name, "Y")) iY = i;
    }
    if ((iR < 0 || iG < 0 || iB < 0) && iY < 0) return imgFail(err, errLen, "EXR image from %s did not contain RGB(A) or Y channels", label);
    const uint32_t linesPerBlock = compression == 3 ? 16u : (compression == 4 ? 32u : 1u);
    int wordsPerPixel[16];
    for (int i = 0; i < nch; i++) wordsPerPixel[i] = ch[i].type == 1 ? 1 : 2;
    const uint32_t blocks = (H + linesPerBlock - 1) / linesPerBlock;
    if (pos + (size_t)blocks * 8 > size) return imgFail(err, errLen, "EXR decode from %s failed (truncated offset table)", label);
    const size_t texel = allHalf ? 8u : 16u;
    uint8_t* pixels = (uint8_t*)malloc((size_t)W * H * texel);
    uint8_t* blockBuf = (uint8_t*)malloc(lineBytes * linesPerBlock);
    uint8_t* tmpBuf = (uint8_t*)malloc(lineBytes * linesPerBlock);
    if (!pixels || !blockBuf || !tmpBuf) { free(pixels); free(blockBuf); free(tmpBuf); return imgFail(err, errLen, "Failed to allocate EXR decode buffer for %s", label); }
    const int sel[4] = {iR >= 0 ? iR : iY, iG >= 0 ? iG : iY, iB >= 0 ? iB : iY, iA};
    int good = 1;
    const char* why = "";
    for (uint32_t b = 0; b < blocks && good; b++) {
        uint64_t off;
        memcpy(&off, data + pos + (size_t)b * 8, 8);
        if (off > size || size - off < 8) { good = 0; why = "bad block offset"; break; }  /* no wrap for offsets near 2^64 */
        int32_t y0;
        uint32_t dataSize;
        memcpy(&y0, data + off, 4);
        memcpy(&dataSize, data + off + 4, 4);
        if (dataSize > size - off - 8 || y0 < ymin || y0 > ymax) { good = 0; why = "bad block header"; break; }
        const uint32_t firstLine = (uint32_t)(y0 - ymin);
        const uint32_t lines = firstLine + linesPerBlock <= H ? linesPerBlock : H - firstLine;
        const size_t rawLen = lineBytes * lines;
        const uint8_t* src = data + off + 8;
        if (compression == 0 || dataSize == rawLen) {  /* a block that did not shrink is stored raw */
            if (dataSize != rawLen) { good = 0; why = "bad block size"; break; }
            memcpy(blockBuf, src, rawLen);
        } else if (compression == 4) {
            if (!exrPizDecode(src, dataSize, blockBuf, rawLen, W, lines, wordsPerPixel, nch)) { good = 0; why = "bad PIZ data"; break; }
        } else if (compression == 1) {
            if (!exrRleDecode(src, dataSize, tmpBuf, rawLen)) { good = 0; why = "bad RLE data"; break; }
            exrUnpredict(tmpBuf, rawLen, blockBuf);
        } else {
            uLongf got = (uLongf)rawLen;
            if (uncompress(tmpBuf, &got, src, dataSize) != Z_OK || got != rawLen) { good = 0; why = "bad zlib data"; break; }
            exrUnpredict(tmpBuf, rawLen, blockBuf);
        }
        for (uint32_t l = 0; l < lines; l++) {
            const uint8_t* line = blockBuf + (size_t)l * lineBytes;
            const uint32_t y = firstLine + l;
            for (int c = 0; c < 4; c++) {
                const int ci = sel[c];
                for (uint32_t x = 0; x < W; x++) {
                    float fv = 1.0f;
                    uint16_t hv = 0x3c00u;
                    if (ci >= 0) {
                        const uint8_t* s = line + ch[ci].offsetInLine;
                        if (ch[ci].type == 1) { memcpy(&hv, s + 2 * (size_t)x, 2); if (!allHalf) fv = exrHalfToFloat(hv); }
                        else if (ch[ci].type == 2) { memcpy(&fv, s + 4 * (size_t)x, 4); }
                        else { uint32_t u; memcpy(&u, s + 4 * (size_t)x, 4); fv = (float)u; }
                    }
                    if (allHalf) ((uint16_t*)pixels)[((size_t)y * W + x) * 4 + c] = hv;
                    else ((float*)pixels)[((size_t)y * W + x) * 4 + c] = fv;
                }
            }
        }
    }
    free(blockBuf); free(tmpBuf);
    (void)exrFloatToHalf;
    if (!good) { free(pixels); return imgFail(err, errLen, "EXR decode from %s failed (%s)", label, why); }
    out->pixels = pixels; out->width = W; out->height = H;
    out->format = allHalf ? VKRT_TEXTURE_FORMAT_RGBA16_SFLOAT : VKRT_TEXTURE_FORMAT_RGBA32_SFLOAT;
    out->colorSpace = VKRT_TEXTURE_COLOR_SPACE_LINEAR;
    return 1;
}

/* ============================================================ front end ====================================================== */
enum { CODEC_UNKNOWN = 0, CODEC_PNG, CODEC_JPEG, CODEC_EXR };

static int codecFromMime(const char* mime) {  /* image.c:86-94 */
    if (!mime || !mime[0]) return CODEC_UNKNOWN;
    if (!strncmp(mime, "image/png", 9)) return CODEC_PNG;
    if (!strncmp(mime, "image/jpeg", 10)) return CODEC_JPEG;
    if (!strncmp(mime, "image/exr", 9) || !strncmp(mime, "image/x-exr", 11)) return CODEC_EXR;
    return CODEC_UNKNOWN;
}
static int codecFromBytes(const uint8_t* d, size_t n) {  /* image.c:96-110 */
    static const uint8_t png[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (n >= 8 && !memcmp(d, png, 8)) return CODEC_PNG;
    if (n >= 3 && d[0] == 0xff && d[1] == 0xd8 && d[2] == 0xff) return CODEC_JPEG;
    if (n >= 4 && d[0] == 0x76 && d[1] == 0x2f && d[2] == 0x31 && d[3] == 0x01) return CODEC_EXR;
    return CODEC_UNKNOWN;
}

int hostDecodeImage(const void* data, size_t size, const char* mimeType, const char* label, uint32_t preferredColorSpace, HostImage* out, char* err,
                    size_t errLen) {
    if (!out) return 0;
    memset(out, 0, sizeof(*out));
    if (!data || !size) return imgFail(err, errLen, "empty image data for %s", label ? label : "?");
    if (!label) label = "<memory>";
    int codec = codecFromMime(mimeType);
    if (codec == CODEC_UNKNOWN) codec = codecFromBytes((const uint8_t*)data, size);
    switch (codec) {
        case CODEC_PNG: return decodePng((const uint8_t*)data, size, label, preferredColorSpace, out, err, errLen);
        case CODEC_JPEG: return decodeJpeg((const uint8_t*)data, size, label, preferredColorSpace, out, err, errLen);
        case CODEC_EXR: return decodeExr((const uint8_t*)data, size, label, out, err, errLen);
        default: return imgFail(err, errLen, "Unsupported image format for %s (%s); supported: PNG, JPEG, EXR", label, mimeType && mimeType[0] ? mimeType : "unknown type");
    }
}

int hostLoadImageFile(const char* path, uint32_t preferredColorSpace, HostImage* out, char* err, size_t errLen) {
    if (out) memset(out, 0, sizeof(*out));
    if (!path || !path[0] || !out) return 0;
    FILE* f = fopen(path, "rb");
    if (!f) return imgFail(err, errLen, "Failed to open image file: %s", path);
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n <= 0) { fclose(f); return imgFail(err, errLen, "Image file is empty: %s", path); }
    uint8_t* bytes = (uint8_t*)malloc((size_t)n);
    if (!bytes || fread(bytes, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(bytes); return imgFail(err, errLen, "Failed to read image file: %s", path); }
    fclose(f);
    int ok = hostDecodeImage(bytes, (size_t)n, NULL, path, preferredColorSpace, out, err, errLen);
    free(bytes);
    return ok;
}

/* ---- exported utility entry points (reference: src/core/utility/image.h:16-27); errors go to stderr like LOG_ERROR ------------- */
#include "../../include/vkrt_host.h"
int vkrtLoadImageFromFile(const char* path, uint32_t preferredColorSpace, VKRT_LoadedImage* outImage) {
    char why[256] = "";
    HostImage img;
    if (!outImage) return 0;
    memset(outImage, 0, sizeof(*outImage));
    if (!hostLoadImageFile(path, preferredColorSpace, &img, why, sizeof(why))) { if (why[0]) fprintf(stderr, "[vkrt host] %s\n", why); return 0; }
    outImage->pixels = img.pixels; outImage->width = img.width; outImage->height = img.height; outImage->format = img.format; outImage->colorSpace = img.colorSpace;
    return 1;
}
int vkrtLoadImageFromMemory(const void* data, size_t size, const char* mimeType, uint32_t preferredColorSpace, VKRT_LoadedImage* outImage) {
    char why[256] = "";
    HostImage img;
    if (!outImage) return 0;
    memset(outImage, 0, sizeof(*outImage));
    if (!hostDecodeImage(data, size, mimeType, "<memory>", preferredColorSpace, &img, why, sizeof(why))) { if (why[0]) fprintf(stderr, "[vkrt host] %s\n", why); return 0; }
    outImage->pixels = img.pixels; outImage->width = img.width; outImage->height = img.height; outImage->format = img.format; outImage->colorSpace = img.colorSpace;
    return 1;
}
void vkrtFreeLoadedImage(VKRT_LoadedImage* image) {
    if (!image) return;
    free(image->pixels);
    memset(image, 0, sizeof(*image));
}
