/* export.c — VKRT_saveRenderImageEx: film read-back and image files.
 *
 * Restates src/core/utility/export/api.c:170-242 (which buffer is read for which format) and export/image.c:388-447,907-1016
 * (spectral accumulation holds XYZ under an equal-energy white: Bradford-adapt to D65, XYZ -> linear sRGB, sanitise, alpha 1).
 * Writers are written for this repo (the reference links tinyexr / libspng / libjpeg-turbo):
 *   .exr : OpenEXR 2 single-part scanline image, RGBA 32-bit float, uncompressed — the accumulation buffer in linear sRGB;
 *   .png : 16-bit RGBA, zlib deflate — the already tone-mapped RGBA16-UNORM output image (offline mode saves undenoised). */
#include <zlib.h>

#include "host_state.h"
#include "image_decode.h"

static void mul3(const float m[9], const float v[3], float out[3]) {
    for (int r = 0; r < 3; r++) out[r] = (m[r * 3 + 0] * v[0]) + (m[r * 3 + 1] * v[1]) + (m[r * 3 + 2] * v[2]);
}
static void adaptEqualEnergyXYZToD65(const float xyz[3], float adapted[3]) {
    static const float bradford[9] = {0.8951f, 0.2664f, -0.1614f, -0.7502f, 1.7135f, 0.0367f, 0.0389f, -0.0685f, 1.0296f};
    static const float bradfordInverse[9] = {0.9869929f, -0.1470543f, 0.1599627f, 0.4323053f, 0.5183603f, 0.0492912f, -0.0085287f, 0.0400428f, 0.9684867f};
    static const float scale[3] = {0.9413344f, 1.0404175f, 1.0895327f};
    float lms[3], a[3];
    mul3(bradford, xyz, lms);
    for (int c = 0; c < 3; c++) a[c] = lms[c] * scale[c];
    mul3(bradfordInverse, a, adapted);
}

/* accumulation (RGBA32F, w = sample count) -> linear sRGB, alpha 1; non-finite channels become 0 when `sanitise` is set (the denoiser
 * is handed the image before that step, as in image.c:907-952) */
static void linearizeAccumulationEx(float* px, size_t count, int spectral, int sanitise) {
    for (size_t i = 0; i < count; i++) {
        float* p = px + i * 4;
        if (spectral) {
            float a[3];
            adaptEqualEnergyXYZToD65(p, a);
            p[0] = (3.2404542f * a[0]) + (-1.5371385f * a[1]) + (-0.4985314f * a[2]);
            p[1] = (-0.9692660f * a[0]) + (1.8760108f * a[1]) + (0.0415560f * a[2]);
            p[2] = (0.0556434f * a[0]) + (-0.2040259f * a[1]) + (1.0572252f * a[2]);
        }
        for (int c = 0; sanitise && c < 3; c++)
            if (!isfinite(p[c])) p[c] = 0.0f;
        p[3] = 1.0f;
    }
}
static void linearizeAccumulation(float* px, size_t count, int spectral) { linearizeAccumulationEx(px, count, spectral, 1); }

/* image.c:907-960 prepareLinearRenderOutput on the read-back buffers: accumulation -> linear sRGB (alpha 1), the denoise stage when asked
 * for, then non-finite channels -> 0. Returns 0 when the denoiser failed and no raw fallback is allowed. */
int vkrtHostPrepareLinearOutput(float* accum, const uint16_t* albedoHalf, const uint16_t* normalHalf, uint32_t w, uint32_t h, int spectral, int denoise,
                                int allowRawFallback, char* note, size_t noteLen) {
    const size_t px = (size_t)w * h;
    if (note && noteLen) note[0] = 0;
    linearizeAccumulationEx(accum, px, spectral, 0);
    if (denoise && !vkrtHostDenoiseLinear(accum, albedoHalf, normalHalf, w, h, allowRawFallback, note, noteLen)) return 0;
    for (size_t i = 0; i < px * 4; i++)   /* sanitizeLinearRGBA32FInPlace(.., 1.0f, 0): alpha is 1 already */
        if ((i & 3u) != 3u && !isfinite(accum[i])) accum[i] = 0.0f;
    return 1;
}

/* ---- OpenEXR ------------------------------------------------------------------------------------------------------------- */
static void putStr(FILE* f, const char* s) { fwrite(s, 1, strlen(s) + 1, f); }
static void putI32(FILE* f, int32_t v) { fwrite(&v, 4, 1, f); }
static void putAttr(FILE* f, const char* name, const char* type, int32_t size) { putStr(f, name); putStr(f, type); putI32(f, size); }

static int writeExr(const char* path, const float* rgba, uint32_t w, uint32_t h) {
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    putI32(f, 20000630);
    putI32(f, 2); /* version 2, single-part scanline */
    const char* names[4] = {"A", "B", "G", "R"}; /* channels are stored in alphabetical order */
    putAttr(f, "channels", "chlist", 4 * 18 + 1);
    for (int c = 0; c < 4; c++) {
        putStr(f, names[c]);
        putI32(f, 2); /* FLOAT */
        unsigned char pl[4] = {0, 0, 0, 0};
        fwrite(pl, 1, 4, f);
        putI32(f, 1); putI32(f, 1);
    }
    fputc(0, f);
    putAttr(f, "compression", "compression", 1); fputc(0, f); /* NO_COMPRESSION */
    putAttr(f, "dataWindow", "box2i", 16); putI32(f, 0); putI32(f, 0); putI32(f, (int32_t)w - 1); putI32(f, (int32_t)h - 1);
    putAttr(f, "displayWindow", "box2i", 16); putI32(f, 0); putI32(f, 0); putI32(f, (int32_t)w - 1); putI32(f, (int32_t)h - 1);
    putAttr(f, "lineOrder", "lineOrder", 1); fputc(0, f);
    float one = 1.0f, zero2[2] = {0.0f, 0.0f};
    putAttr(f, "pixelAspectRatio", "float", 4); fwrite(&one, 4, 1, f);
    putAttr(f, "screenWindowCenter", "v2f", 8); fwrite(zero2, 4, 2, f);
    putAttr(f, "screenWindowWidth", "float", 4); fwrite(&one, 4, 1, f);
    fputc(0, f);
    uint64_t tableStart = (uint64_t)ftell(f);
    uint64_t rowBytes = (uint64_t)w * 16u;
    uint64_t first = tableStart + (uint64_t)h * 8u;
    for (uint32_t y = 0; y < h; y++) {
        uint64_t off = first + (uint64_t)y * (8u + rowBytes);
        fwrite(&off, 8, 1, f);
    }
    float* row = (float*)malloc((size_t)rowBytes);
    if (!row) { fclose(f); return 0; }
    const int src[4] = {3, 2, 1, 0}; /* A, B, G, R <- rgba */
    for (uint32_t y = 0; y < h; y++) {
        putI32(f, (int32_t)y);
        putI32(f, (int32_t)rowBytes);
        for (int c = 0; c < 4; c++)
            for (uint32_t x = 0; x < w; x++) row[(size_t)c * w + x] = rgba[((size_t)y * w + x) * 4 + src[c]];
        fwrite(row, 1, (size_t)rowBytes, f);
    }
    free(row);
    int ok = ferror(f) == 0;
    fclose(f);
    return ok;
}

/* ---- PNG (16-bit RGBA) ---------------------------------------------------------------------------------------------------- */
static void pngChunk(FILE* f, const char* tag, const unsigned char* data, uint32_t len) {
    unsigned char hdr[8] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len,
                            (unsigned char)tag[0], (unsigned char)tag[1], (unsigned char)tag[2], (unsigned char)tag[3]};
    fwrite(hdr, 1, 8, f);
    if (len) fwrite(data, 1, len, f);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (len) crc = crc32(crc, data, len);
    unsigned char c[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    fwrite(c, 1, 4, f);
}
static int writePng16(const char* path, const uint16_t* rgba, uint32_t w, uint32_t h) {
    size_t stride = (size_t)w * 8 + 1;
    unsigned char* raw = (unsigned char*)malloc(stride * h);
    if (!raw) return 0;
    for (uint32_t y = 0; y < h; y++) {
        unsigned char* row = raw + (size_t)y * stride;
        row[0] = 0; /* filter: none */
        for (size_t k = 0; k < (size_t)w * 4; k++) {
            uint16_t v = rgba[(size_t)y * w * 4 + k];
            row[1 + k * 2] = (unsigned char)(v >> 8);
            row[2 + k * 2] = (unsigned char)(v & 0xff);
        }
    }
    uLongf clen = compressBound((uLong)(stride * h));
    unsigned char* comp = (unsigned char*)malloc(clen);
    if (!comp || compress2(comp, &clen, raw, (uLong)(stride * h), 6) != Z_OK) { free(raw); free(comp); return 0; }
    free(raw);
    FILE* f = fopen(path, "wb");
    if (!f) { free(comp); return 0; }
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    fwrite(sig, 1, 8, f);
    unsigned char ihdr[13] = {(unsigned char)(w >> 24), (unsigned char)(w >> 16), (unsigned char)(w >> 8), (unsigned char)w,
                              (unsigned char)(h >> 24), (unsigned char)(h >> 16), (unsigned char)(h >> 8), (unsigned char)h, 16, 6, 0, 0, 0};
    pngChunk(f, "IHDR", ihdr, 13);
    pngChunk(f, "IDAT", comp, (uint32_t)clen);
    pngChunk(f, "IEND", NULL, 0);
    free(comp);
    int ok = ferror(f) == 0;
    fclose(f);
    return ok;
}

/* src/core/utility/exr.h: vkrtWriteEXRFromRGBA32F (the reference writes through tinyexr; tests/test_reference_pin.py reads this writer's
 * files back with the reference's tinyexr loader, and the reference's files with this host's reader) */
int vkrtWriteEXRFromRGBA32F(const char* path, const float* rgba32f, uint32_t width, uint32_t height) {
    if (!path || !rgba32f || !width || !height) return 0;
    return writeExr(path, rgba32f, width, height);
}

static int hasSuffix(const char* s, const char* suffix) {
    size_t n = strlen(s), m = strlen(suffix);
    if (m > n) return 0;
    for (size_t i = 0; i < m; i++) {
        char a = s[n - m + i], b = suffix[i];
        if (a >= 'A' && a <= 'Z') a = (char)(a - 'A' + 'a');
        if (a != b) return 0;
    }
    return 1;
}

void VKRT_defaultRenderExportSettings(VKRT_RenderExportSettings* s) {
    if (s) s->denoiseEnabled = 0;
}

VKRT_Result VKRT_saveRenderImageEx(VKRT* v, const char* path, const VKRT_RenderExportSettings* settings) {
    if (!v || !path || !path[0]) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!v->initialized || v->hostOnly || !v->cuda) return VKRT_ERROR_OPERATION_FAILED;
    char withExtension[4096];
    {   /* export/image.c:134-149 resolveRenderImagePath: a path without an extension is saved as PNG */
        const char* base = strrchr(path, '/');
        base = base ? base + 1 : path;
        if (!strchr(base, '.')) {
            if (snprintf(withExtension, sizeof(withExtension), "%s.png", path) >= (int)sizeof(withExtension)) return VKRT_ERROR_INVALID_ARGUMENT;
            path = withExtension;
        }
    }
    const uint32_t w = v->renderWidth, h = v->renderHeight;
    const size_t px = (size_t)w * h;
    VKRT_Result r;
    if (v->createInfo.worldSize > 1u) { /* rank 0 assembles the full frame (no-op when the caller already gathered) */
        float ms = 0.0f;
        r = vkrt_cuda_gather(v->cuda, &ms);
        if (r != VKRT_SUCCESS) return hostFail(v, r, "gather: %s", vkrt_cuda_last_error(v->cuda));
        if (v->createInfo.rank != 0u) return VKRT_SUCCESS;
    }
    /* export/api.c:213-236 + image.c:907-1016: with the denoiser on (and no debug view) every format starts from the accumulation buffer:
       linear sRGB -> OIDN with the feature AOVs -> sanitise; EXR stores that, PNG / JPEG tone-map it on the CPU. A failed or missing
       denoiser saves the raw image (allowRawFallback = 1 for a save) and leaves the reason in the last-error string. */
    const int denoise = settings && settings->denoiseEnabled && v->sceneSettings.debugMode == VKRT_DEBUG_MODE_NONE;
    const int isExr = hasSuffix(path, ".exr"), isPng = hasSuffix(path, ".png"), isJpeg = hasSuffix(path, ".jpg") || hasSuffix(path, ".jpeg");
    if (denoise && (isExr || isPng || isJpeg)) {
        float* acc = (float*)malloc(px * 16);
        uint16_t* albedo = (uint16_t*)malloc(px * 8);
        uint16_t* normal = (uint16_t*)malloc(px * 8);
        uint16_t* display = (uint16_t*)malloc(px * 8);
        r = (acc && albedo && normal && display) ? VKRT_SUCCESS : VKRT_ERROR_OUT_OF_MEMORY;
        if (r == VKRT_SUCCESS) r = vkrt_cuda_read_aov(v->cuda, VKRT_CUDA_AOV_ACCUM_RGBA32F, acc, px * 16);
        if (r == VKRT_SUCCESS) r = vkrt_cuda_read_aov(v->cuda, VKRT_CUDA_AOV_ALBEDO_RGBA16F, albedo, px * 8);
        if (r == VKRT_SUCCESS) r = vkrt_cuda_read_aov(v->cuda, VKRT_CUDA_AOV_NORMAL_RGBA16F, normal, px * 8);
        if (r != VKRT_SUCCESS) {
            hostFail(v, r, "read_aov: %s", v->cuda ? vkrt_cuda_last_error(v->cuda) : "");
        } else {
            char note[256];
            if (!vkrtHostPrepareLinearOutput(acc, albedo, normal, w, h, v->sceneSettings.renderMode == VKRT_RENDER_MODE_SPECTRAL, 1, 1, note, sizeof(note))) {
                r = hostFail(v, VKRT_ERROR_OPERATION_FAILED, "OIDN denoising failed for '%s': %s", path, note);
            } else {
                if (note[0]) hostFail(v, VKRT_SUCCESS, "OIDN denoising failed for '%s'; using raw render (%s)", path, note);
                int ok;
                if (isExr) {
                    ok = writeExr(path, acc, w, h);
                } else {
                    vkrtHostLinearToDisplay16(acc, w, h, v->sceneSettings.toneMappingMode, v->sceneSettings.exposure, v->sceneSettings.debugMode, display);
                    if (isPng) {
                        ok = writePng16(path, display, w, h);
                    } else {
                        uint8_t* out8 = (uint8_t*)albedo; /* (px * 8 bytes: room for px * 4) */
                        for (size_t i = 0; i < px * 4; i++) out8[i] = (uint8_t)((((uint32_t)display[i] * 255u) + 32767u) / 65535u);
                        ok = hostWriteJpegRgba8(path, out8, w, h, 95);
                        if (!ok) remove(path);
                    }
                }
                if (!ok) r = hostFail(v, VKRT_ERROR_OPERATION_FAILED, "cannot write %s", path);
            }
        }
        free(acc); free(albedo); free(normal); free(display);
        return r;
    }
    if (hasSuffix(path, ".exr")) {
        float* acc = (float*)malloc(px * 16);
        if (!acc) return VKRT_ERROR_OUT_OF_MEMORY;
        r = vkrt_cuda_read_aov(v->cuda, VKRT_CUDA_AOV_ACCUM_RGBA32F, acc, px * 16);
        if (r == VKRT_SUCCESS) {
            linearizeAccumulation(acc, px, v->sceneSettings.renderMode == VKRT_RENDER_MODE_SPECTRAL && v->sceneSettings.debugMode == VKRT_DEBUG_MODE_NONE);
            if (!writeExr(path, acc, w, h)) r = hostFail(v, VKRT_ERROR_OPERATION_FAILED, "cannot write %s", path);
        } else hostFail(v, r, "read_aov: %s", vkrt_cuda_last_error(v->cuda));
        free(acc);
        return r;
    }
    if (hasSuffix(path, ".png")) {
        uint16_t* out = (uint16_t*)malloc(px * 8);
        if (!out) return VKRT_ERROR_OUT_OF_MEMORY;
        r = vkrt_cuda_read_aov(v->cuda, VKRT_CUDA_AOV_OUTPUT_RGBA16, out, px * 8);
        if (r == VKRT_SUCCESS) {
            if (!writePng16(path, out, w, h)) r = hostFail(v, VKRT_ERROR_OPERATION_FAILED, "cannot write %s", path);
        } else hostFail(v, r, "read_aov: %s", vkrt_cuda_last_error(v->cuda));
        free(out);
        return r;
    }
    if (hasSuffix(path, ".jpg") || hasSuffix(path, ".jpeg")) {   /* export/image.c:298-327: RGBA16 UNORM output -> 8 bits, quality 95, 4:4:4 */
        uint16_t* out = (uint16_t*)malloc(px * 8);
        uint8_t* out8 = (uint8_t*)malloc(px * 4);
        if (!out || !out8) { free(out); free(out8); return VKRT_ERROR_OUT_OF_MEMORY; }
        r = vkrt_cuda_read_aov(v->cuda, VKRT_CUDA_AOV_OUTPUT_RGBA16, out, px * 8);
        if (r == VKRT_SUCCESS) {
            for (size_t i = 0; i < px * 4; i++) out8[i] = (uint8_t)((((uint32_t)out[i] * 255u) + 32767u) / 65535u);
            if (!hostWriteJpegRgba8(path, out8, w, h, 95)) { remove(path); r = hostFail(v, VKRT_ERROR_OPERATION_FAILED, "cannot write %s", path); }
        } else hostFail(v, r, "read_aov: %s", vkrt_cuda_last_error(v->cuda));
        free(out); free(out8);
        return r;
    }
    return hostFail(v, VKRT_ERROR_INVALID_ARGUMENT, "%s: unsupported render export format (use .png, .jpg or .exr)", path);
}
VKRT_Result VKRT_saveRenderImage(VKRT* v, const char* path) {
    VKRT_RenderExportSettings s;
    VKRT_defaultRenderExportSettings(&s);
    return VKRT_saveRenderImageEx(v, path, &s);
}
