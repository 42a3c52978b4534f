/* jpeg_encode.c — baseline JPEG writer for VKRT_saveRenderImage("*.jpg").
 *
 * Mirrors src/core/utility/export/image.c:220-263 (tjCompress2, TJPF_RGBA, TJSAMP_444, quality 95, TJFLAG_ACCURATEDCT): 4:4:4 YCbCr,
 * the IJG quality-scaled Annex K quantisation tables, the "islow" forward DCT (jfdctint, CONST_BITS 13 / PASS1_BITS 2), the fixed-point
 * RGB -> YCbCr conversion of jccolor (SCALEBITS 16) and the standard Annex K Huffman tables (libjpeg's default when optimize_coding is
 * off). Alpha is dropped. libjpeg-turbo itself is not vendored here. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "image_decode.h"

static const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
/* ITU T.81 Annex K.1 (natural order) */
static const uint8_t kLumaQuant[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                       18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t kChromaQuant[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                         99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
/* Annex K.3 Huffman tables: BITS[1..16] and HUFFVAL */
static const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
    0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
    0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1,
    0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
    0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

typedef struct { uint16_t code[256]; uint8_t len[256]; } EncTable;
typedef struct { uint8_t* p; size_t n, cap; uint32_t acc; int bits; int failed; } Out;

static void outByte(Out* o, uint8_t b) {
    if (o->n == o->cap) {
        size_t cap = o->cap ? o->cap * 2 : 1 << 16;
        uint8_t* g = (uint8_t*)realloc(o->p, cap);
        if (!g) { o->failed = 1; return; }
        o->p = g; o->cap = cap;
    }
    o->p[o->n++] = b;
}
static void outBytes(Out* o, const void* src, size_t n) { for (size_t i = 0; i < n; i++) outByte(o, ((const uint8_t*)src)[i]); }
static void outU16(Out* o, unsigned v) { outByte(o, (uint8_t)(v >> 8)); outByte(o, (uint8_t)v); }
static void outBits(Out* o, uint32_t code, int len) {
    o->acc = (o->acc << len) | (code & ((1u << len) - 1u));
    o->bits += len;
    while (o->bits >= 8) {
        uint8_t b = (uint8_t)(o->acc >> (o->bits - 8));
        outByte(o, b);
        if (b == 0xff) outByte(o, 0x00);
        o->bits -= 8;
    }
}
static void buildEncTable(EncTable* t, const uint8_t* bits, const uint8_t* vals) {
    memset(t, 0, sizeof(*t));
    uint32_t code = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1]; i++) { t->code[vals[k]] = (uint16_t)code; t->len[vals[k]] = (uint8_t)l; k++; code++; }
        code <<= 1;
    }
}
static void writeDht(Out* o, int cls, int id, const uint8_t* bits, const uint8_t* vals, int count) {
    outU16(o, 0xffc4); outU16(o, (unsigned)(2 + 1 + 16 + count));
    outByte(o, (uint8_t)((cls << 4) | id));
    outBytes(o, bits, 16);
    outBytes(o, vals, (size_t)count);
}

/* jfdctint.c jpeg_fdct_islow on samples already level-shifted by -128; output scaled by 8 as in libjpeg */
static void fdctIslow(int* data) {
    enum { CB = 13, P1 = 2 };
#define FX(x) ((int)((x) * (1 << CB) + 0.5))
#define DS(x, n) (((x) + (1 << ((n) - 1))) >> (n))
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 8; i++) {
            int* d = pass == 0 ? data + i * 8 : data + i;
            const int s = pass == 0 ? 1 : 8;
            int tmp0 = d[0] + d[7 * s], tmp7 = d[0] - d[7 * s], tmp1 = d[s] + d[6 * s], tmp6 = d[s] - d[6 * s];
            int tmp2 = d[2 * s] + d[5 * s], tmp5 = d[2 * s] - d[5 * s], tmp3 = d[3 * s] + d[4 * s], tmp4 = d[3 * s] - d[4 * s];
            int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            if (pass == 0) {
                d[0] = (tmp10 + tmp11) << P1;
                d[4 * s] = (tmp10 - tmp11) << P1;
            } else {
                d[0] = DS(tmp10 + tmp11, P1);
                d[4 * s] = DS(tmp10 - tmp11, P1);
            }
            const int sh = pass == 0 ? CB - P1 : CB + P1;
            int z1 = (tmp12 + tmp13) * FX(0.541196100);
            d[2 * s] = DS(z1 + tmp13 * FX(0.765366865), sh);
            d[6 * s] = DS(z1 + tmp12 * (-FX(1.847759065)), sh);
            z1 = tmp4 + tmp7;
            int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
            int z5 = (z3 + z4) * FX(1.175875602);
            tmp4 *= FX(0.298631336); tmp5 *= FX(2.053119869); tmp6 *= FX(3.072711026); tmp7 *= FX(1.501321110);
            z1 *= -FX(0.899976223); z2 *= -FX(2.562915447); z3 *= -FX(1.961570560); z4 *= -FX(0.390180644);
            z3 += z5; z4 += z5;
            d[7 * s] = DS(tmp4 + z1 + z3, sh);
            d[5 * s] = DS(tmp5 + z2 + z4, sh);
            d[3 * s] = DS(tmp6 + z2 + z3, sh);
            d[s] = DS(tmp7 + z1 + z4, sh);
        }
    }
#undef FX
#undef DS
}

static int bitCount(int v) {
    int n = 0;
    if (v < 0) v = -v;
    while (v) { n++; v >>= 1; }
    return n;
}
static void encodeBlock(Out* o, const int* coef, const uint16_t* quant, int* dcPred, const EncTable* dc, const EncTable* ac) {
    int q[64];
    for (int i = 0; i < 64; i++) {   /* jcdctmgr.c: divide by 8 * quant with rounding to nearest, symmetric about zero */
        const int qv = quant[i] << 3;
        int t = coef[i];
        if (t < 0) { t = -t; t += qv >> 1; t = t >= qv ? t / qv : 0; t = -t; }
        else { t += qv >> 1; t = t >= qv ? t / qv : 0; }
        q[i] = t;
    }
    int diff = q[0] - *dcPred;
    *dcPred = q[0];
    int nb = bitCount(diff);
    outBits(o, dc->code[nb], dc->len[nb]);
    if (nb) outBits(o, (uint32_t)(diff < 0 ? diff - 1 : diff), nb);
    int run = 0;
    for (int k = 1; k < 64; k++) {
        const int v = q[kZigzag[k]];
        if (v == 0) { run++; continue; }
        while (run > 15) { outBits(o, ac->code[0xf0], ac->len[0xf0]); run -= 16; }
        nb = bitCount(v);
        const int sym = (run << 4) | nb;
        outBits(o, ac->code[sym], ac->len[sym]);
        outBits(o, (uint32_t)(v < 0 ? v - 1 : v), nb);
        run = 0;
    }
    if (run > 0) outBits(o, ac->code[0x00], ac->len[0x00]);
}

int hostWriteJpegRgba8(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height, int quality) {
    if (!path || !rgba8 || !width || !height || width > 65535u || height > 65535u) return 0;
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;   /* jcparam.c jpeg_quality_scaling */
    uint16_t quant[2][64];
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < 64; i++) {
            long v = ((long)(t ? kChromaQuant[i] : kLumaQuant[i]) * scale + 50L) / 100L;
            quant[t][i] = (uint16_t)(v < 1 ? 1 : (v > 255 ? 255 : v));     /* baseline: 8-bit tables */
        }
    Out o;
    memset(&o, 0, sizeof(o));
    outU16(&o, 0xffd8);
    static const uint8_t jfif[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    outU16(&o, 0xffe0); outU16(&o, 16); outBytes(&o, jfif, 14);
    for (int t = 0; t < 2; t++) {
        outU16(&o, 0xffdb); outU16(&o, 67); outByte(&o, (uint8_t)t);
        for (int i = 0; i < 64; i++) outByte(&o, (uint8_t)quant[t][kZigzag[i]]);
    }
    outU16(&o, 0xffc0); outU16(&o, 17); outByte(&o, 8); outU16(&o, height); outU16(&o, width); outByte(&o, 3);
    outByte(&o, 1); outByte(&o, 0x11); outByte(&o, 0);
    outByte(&o, 2); outByte(&o, 0x11); outByte(&o, 1);
    outByte(&o, 3); outByte(&o, 0x11); outByte(&o, 1);
    writeDht(&o, 0, 0, kDcLumaBits, kDcVals, 12);
    writeDht(&o, 1, 0, kAcLumaBits, kAcLumaVals, 162);
    writeDht(&o, 0, 1, kDcChromaBits, kDcVals, 12);
    writeDht(&o, 1, 1, kAcChromaBits, kAcChromaVals, 162);
    outU16(&o, 0xffda); outU16(&o, 12); outByte(&o, 3);
    outByte(&o, 1); outByte(&o, 0x00); outByte(&o, 2); outByte(&o, 0x11); outByte(&o, 3); outByte(&o, 0x11);
    outByte(&o, 0); outByte(&o, 63); outByte(&o, 0);
    EncTable dcT[2], acT[2];
    buildEncTable(&dcT[0], kDcLumaBits, kDcVals); buildEncTable(&dcT[1], kDcChromaBits, kDcVals);
    buildEncTable(&acT[0], kAcLumaBits, kAcLumaVals); buildEncTable(&acT[1], kAcChromaBits, kAcChromaVals);
    int dcPred[3] = {0, 0, 0};
    for (uint32_t by = 0; by < height; by += 8)
        for (uint32_t bx = 0; bx < width; bx += 8) {
            int blk[3][64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    const uint32_t sx = bx + x < width ? bx + x : width - 1, sy = by + y < height ? by + y : height - 1;   /* edge replication */
                    const uint8_t* p = rgba8 + ((size_t)sy * width + sx) * 4;
                    const int r = p[0], g = p[1], b = p[2];
                    /* jccolor.c rgb_ycc_convert, SCALEBITS 16 */
                    blk[0][y * 8 + x] = ((19595 * r + 38470 * g + 7471 * b + 32768) >> 16) - 128;
                    blk[1][y * 8 + x] = ((-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16) - 128;
                    blk[2][y * 8 + x] = ((32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16) - 128;
                }
            for (int c = 0; c < 3; c++) {
                fdctIslow(blk[c]);
                encodeBlock(&o, blk[c], quant[c ? 1 : 0], &dcPred[c], &dcT[c ? 1 : 0], &acT[c ? 1 : 0]);
            }
        }
    if (o.bits) outBits(&o, 0x7f, 8 - o.bits);   /* pad the last byte with ones */
    outU16(&o, 0xffd9);
    int ok = !o.failed;
    if (ok) {
        FILE* f = fopen(path, "wb");
        ok = f && fwrite(o.p, 1, o.n, f) == o.n;
        if (f) fclose(f);
    }
    free(o.p);
    return ok;
}

/* exported for tests and front ends (include/vkrt_host.h) */
#include "../../include/vkrt_host.h"
int vkrtWriteJPEGFromRGBA8(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height, int quality) {
    return hostWriteJpegRgba8(path, rgba8, width, height, quality);
}
