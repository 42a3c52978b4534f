/* denoise.c — the save-time denoise stage of VKRT_saveRenderImageEx: feature AOVs -> Open Image Denoise.
 *
 * Restates src/core/utility/export/image.c:488-640,701-905 (feature preparation, prefilter, "RT" filter, raw fallback) and
 * src/core/utility/denoise.c:99-340 (the OIDN call sequence and filter parameters). The reference links libOpenImageDenoise at build
 * time; here the library is bound at run time with dlopen (VKRT_OIDN_LIBRARY, else the usual sonames), because it is a third-party
 * CPU library that is not part of the path: when it is absent the stage behaves like a filter that failed — the raw image is saved
 * and the reason is kept in the host's last-error string. Everything around the library call is deterministic host arithmetic and is
 * checked byte for byte against the reference's own source with a stand-in OIDN (tests/fake_oidn.c, tests/test_denoise.py). */
#include <dlfcn.h>

#include "host_state.h"

/* ---- the part of the OIDN 2 C API the stage uses (OpenImageDenoise/oidn.h) ------------------------------------------------------------ */
typedef struct OIDNDeviceImpl* OIDNDevice;
typedef struct OIDNFilterImpl* OIDNFilter;
typedef struct OIDNBufferImpl* OIDNBuffer;
enum { OIDN_DEVICE_TYPE_CPU = 1, OIDN_FORMAT_FLOAT3 = 3, OIDN_ERROR_NONE = 0, OIDN_QUALITY_HIGH = 6 };

typedef struct OidnApi {
    void* handle;
    int tried;
    char why[192];
    OIDNDevice (*newDevice)(int);
    void (*commitDevice)(OIDNDevice);
    void (*syncDevice)(OIDNDevice);
    void (*releaseDevice)(OIDNDevice);
    int (*getDeviceError)(OIDNDevice, const char**);
    OIDNFilter (*newFilter)(OIDNDevice, const char*);
    void (*releaseFilter)(OIDNFilter);
    void (*setFilterImage)(OIDNFilter, const char*, OIDNBuffer, int, size_t, size_t, size_t, size_t, size_t);
    void (*setFilterBool)(OIDNFilter, const char*, _Bool);
    void (*setFilterInt)(OIDNFilter, const char*, int);
    void (*commitFilter)(OIDNFilter);
    void (*executeFilter)(OIDNFilter);
    OIDNBuffer (*newBuffer)(OIDNDevice, size_t);
    void (*releaseBuffer)(OIDNBuffer);
    void (*writeBuffer)(OIDNBuffer, size_t, size_t, const void*);
    void (*readBuffer)(OIDNBuffer, size_t, size_t, void*);
} OidnApi;

static OidnApi g_oidn;

static int oidnBind(void) {
    OidnApi* a = &g_oidn;
    if (a->tried) return a->handle != NULL;
    a->tried = 1;
    const char* names[4] = {getenv("VKRT_OIDN_LIBRARY"), "libOpenImageDenoise.so.2", "libOpenImageDenoise.so.1", "libOpenImageDenoise.so"};
    void* h = NULL;
    for (int i = 0; i < 4 && !h; i++)
        if (names[i] && names[i][0]) h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        snprintf(a->why, sizeof(a->why), "Open Image Denoise is not installed (%s)", dlerror());
        return 0;
    }
    struct { void** slot; const char* name; } syms[] = {
        {(void**)&a->newDevice, "oidnNewDevice"},           {(void**)&a->commitDevice, "oidnCommitDevice"},   {(void**)&a->syncDevice, "oidnSyncDevice"},
        {(void**)&a->releaseDevice, "oidnReleaseDevice"},   {(void**)&a->getDeviceError, "oidnGetDeviceError"}, {(void**)&a->newFilter, "oidnNewFilter"},
        {(void**)&a->releaseFilter, "oidnReleaseFilter"},   {(void**)&a->setFilterImage, "oidnSetFilterImage"}, {(void**)&a->setFilterBool, "oidnSetFilterBool"},
        {(void**)&a->setFilterInt, "oidnSetFilterInt"},     {(void**)&a->commitFilter, "oidnCommitFilter"},   {(void**)&a->executeFilter, "oidnExecuteFilter"},
        {(void**)&a->newBuffer, "oidnNewBuffer"},           {(void**)&a->releaseBuffer, "oidnReleaseBuffer"}, {(void**)&a->writeBuffer, "oidnWriteBuffer"},
        {(void**)&a->readBuffer, "oidnReadBuffer"},
    };
    for (size_t i = 0; i < sizeof(syms) / sizeof(syms[0]); i++) {
        *syms[i].slot = dlsym(h, syms[i].name);
        if (!*syms[i].slot) {
            snprintf(a->why, sizeof(a->why), "Open Image Denoise library lacks %s", syms[i].name);
            dlclose(h);
            return 0;
        }
    }
    a->handle = h;
    return 1;
}

/* test hook: forget the bound library so that the next call binds again (VKRT_OIDN_LIBRARY may have changed) */
VKRT_HOST_API void vkrtHostResetDenoiser(void) {
    if (g_oidn.handle) dlclose(g_oidn.handle);
    memset(&g_oidn, 0, sizeof(g_oidn));
}

/* One "RT" filter run (denoise.c:245-280): `mainName` is "color" for the beauty pass (hdr, auxiliary images attached) and "albedo" /
 * "normal" for the prefilter of a feature image (ldr, nothing attached). The output buffer starts as a copy of the main image. */
static int runRtFilter(const char* mainName, const float* mainImage, const float* albedo, const float* normal, uint32_t w, uint32_t h, int hdr, int cleanAux,
                       float* out, char* err, size_t errLen) {
    if (!oidnBind()) {
        snprintf(err, errLen, "%s", g_oidn.why);
        return 0;
    }
    const OidnApi* a = &g_oidn;
    const size_t pixelStride = sizeof(float) * 4u, rowStride = pixelStride * w, bytes = rowStride * h;
    OIDNDevice dev = a->newDevice(OIDN_DEVICE_TYPE_CPU);
    if (!dev) {
        snprintf(err, errLen, "failed to create OIDN device");
        return 0;
    }
    a->commitDevice(dev);
    OIDNFilter filter = a->newFilter(dev, "RT");
    OIDNBuffer bufs[4] = {NULL, NULL, NULL, NULL}; /* main, output, albedo, normal */
    const float* src[4] = {mainImage, mainImage, albedo, normal};
    const char* label[4] = {"color", "output", "albedo", "normal"};
    int ok = filter != NULL;
    if (!ok) snprintf(err, errLen, "failed to create OIDN RT filter");
    for (int i = 0; ok && i < 4; i++) {
        if (!src[i]) continue;
        bufs[i] = a->newBuffer(dev, bytes);
        if (!bufs[i]) {
            snprintf(err, errLen, "failed to allocate OIDN %s buffer", label[i]);
            ok = 0;
        } else {
            a->writeBuffer(bufs[i], 0u, bytes, src[i]);
        }
    }
    if (ok) {
        a->setFilterImage(filter, mainName, bufs[0], OIDN_FORMAT_FLOAT3, w, h, 0u, pixelStride, rowStride);
        if (albedo && strcmp(mainName, "albedo") != 0) a->setFilterImage(filter, "albedo", bufs[2], OIDN_FORMAT_FLOAT3, w, h, 0u, pixelStride, rowStride);
        if (normal && strcmp(mainName, "normal") != 0) a->setFilterImage(filter, "normal", bufs[3], OIDN_FORMAT_FLOAT3, w, h, 0u, pixelStride, rowStride);
        a->setFilterImage(filter, "output", bufs[1], OIDN_FORMAT_FLOAT3, w, h, 0u, pixelStride, rowStride);
        a->setFilterBool(filter, "hdr", hdr != 0);
        a->setFilterBool(filter, "srgb", 0);
        a->setFilterBool(filter, "cleanAux", cleanAux != 0);
        a->setFilterInt(filter, "quality", OIDN_QUALITY_HIGH);
        a->commitFilter(filter);
        a->executeFilter(filter);
        a->syncDevice(dev);
        const char* msg = NULL;
        int e = a->getDeviceError(dev, &msg);
        if (e == OIDN_ERROR_NONE) {
            a->readBuffer(bufs[1], 0u, bytes, out);
            e = a->getDeviceError(dev, &msg);
        }
        if (e != OIDN_ERROR_NONE) {
            snprintf(err, errLen, "%s", (msg && msg[0]) ? msg : "OIDN filtering failed");
            ok = 0;
        }
    }
    static const int releaseOrder[4] = {1, 3, 2, 0}; /* output, normal, albedo, colour (denoise.c:77-97) */
    for (int k = 0; k < 4; k++)
        if (bufs[releaseOrder[k]]) a->releaseBuffer(bufs[releaseOrder[k]]);
    if (filter) a->releaseFilter(filter);
    a->releaseDevice(dev);
    return ok;
}

/* ---- feature images (image.c:346-373,488-640) ------------------------------------------------------------------------------------------- */
static float halfToFloat(uint16_t v) {
    uint32_t sign = ((uint32_t)v & 0x8000u) << 16, e = ((uint32_t)v >> 10) & 0x1fu, m = (uint32_t)v & 0x3ffu, bits;
    if (e == 0u) {
        if (m == 0u) {
            bits = sign;
        } else { /* subnormal half: renormalise */
            e = 1u;
            while (!(m & 0x400u)) { m <<= 1; e++; }
            bits = sign | ((127u - 15u - e + 1u) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (e == 0x1fu) {
        bits = sign | 0x7f800000u | (m << 13);
    } else {
        bits = sign | ((e + 112u) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* RGBA16F feature AOV -> RGBA32F with the accumulated weight kept in .w; a pixel without weight is cleared (no first non-specular hit
 * was ever recorded there). Albedo is clamped to >= 0, normals are re-normalised (or cleared when shorter than 1e-10). */
static float* prepareFeature(const uint16_t* half, size_t pixels, int isNormal) {
    float* out = (float*)malloc(pixels * 16u);
    if (!out) return NULL;
    for (size_t i = 0; i < pixels; i++) {
        float p[4];
        for (int c = 0; c < 4; c++) p[c] = halfToFloat(half[i * 4u + (size_t)c]);
        p[3] = fmaxf(p[3], 0.0f);
        if (p[3] <= 0.0f) {
            p[0] = p[1] = p[2] = 0.0f;
        } else if (!isNormal) {
            p[0] = fmaxf(p[0], 0.0f); p[1] = fmaxf(p[1], 0.0f); p[2] = fmaxf(p[2], 0.0f);
        } else {
            float l2 = (p[0] * p[0]) + (p[1] * p[1]) + (p[2] * p[2]);
            if (l2 > 1e-20f) {
                float inv = 1.0f / sqrtf(l2);
                p[0] *= inv; p[1] *= inv; p[2] *= inv;
            } else {
                p[0] = p[1] = p[2] = 0.0f;
            }
        }
        memcpy(out + i * 4u, p, 16);
    }
    return out;
}
static int hasCoverage(const float* px, size_t pixels) {
    for (size_t i = 0; i < pixels; i++)
        if (px[i * 4u + 3u] > 0.0f) return 1;
    return 0;
}
static void prefilterInPlace(const char* name, float** image, uint32_t w, uint32_t h, int* done, char* err, size_t errLen) {
    *done = 0;
    float* filtered = (float*)malloc((size_t)w * h * 16u);
    if (!filtered) return;
    memcpy(filtered, *image, (size_t)w * h * 16u);
    if (runRtFilter(name, *image, NULL, NULL, w, h, 0, 0, filtered, err, errLen)) {
        free(*image);
        *image = filtered;
        *done = 1;
    } else {
        free(filtered); /* the raw feature image is used (image.c:736-765) */
    }
}

/* image.c:840-905 denoiseLinearRenderOutput. `linear` (RGBA32F, alpha 1) is replaced by the denoised image on success. Returns 1 when
 * the caller may go on (denoised, or failed with raw fallback allowed), 0 when the save must fail. `note` receives why a fallback happened. */
VKRT_HOST_API int vkrtHostDenoiseLinear(float* linear, const uint16_t* albedoHalf, const uint16_t* normalHalf, uint32_t w, uint32_t h, int allowRawFallback, char* note,
                                        size_t noteLen) {
    if (note && noteLen) note[0] = 0;
    if (!linear || !w || !h) return 0;
    const size_t pixels = (size_t)w * h;
    char err[256] = "";
    float *albedo = NULL, *normal = NULL;
    int features = 0;
    if (albedoHalf && normalHalf) {
        albedo = prepareFeature(albedoHalf, pixels, 0);
        normal = albedo ? prepareFeature(normalHalf, pixels, 1) : NULL;
        features = albedo && normal && (hasCoverage(albedo, pixels) || hasCoverage(normal, pixels));
    }
    int cleanAlbedo = 0, cleanNormal = 0;
    if (features) {
        prefilterInPlace("albedo", &albedo, w, h, &cleanAlbedo, err, sizeof(err));
        prefilterInPlace("normal", &normal, w, h, &cleanNormal, err, sizeof(err));
    }
    float* denoised = (float*)malloc(pixels * 16u);
    int ok = 0;
    if (denoised && runRtFilter("color", linear, features ? albedo : NULL, features ? normal : NULL, w, h, 1, cleanAlbedo && cleanNormal, denoised, err, sizeof(err))) {
        for (size_t i = 0; i < pixels; i++) { /* sanitizeLinearRGBA32FInPlace(.., 1.0f, forceOpaqueAlpha) */
            float* p = denoised + i * 4u;
            for (int c = 0; c < 3; c++)
                if (!isfinite(p[c])) p[c] = 0.0f;
            p[3] = 1.0f;
        }
        memcpy(linear, denoised, pixels * 16u);
        ok = 1;
    } else {
        if (note && noteLen) snprintf(note, noteLen, "%s", err[0] ? err : "out of memory");
        ok = allowRawFallback ? 1 : 0;
    }
    free(denoised);
    free(normal);
    free(albedo);
    return ok;
}

/* image.c:449-487,641-699 convertLinearToDisplayRGBA16: exposure, ACES (or none), sRGB transfer, 16-bit UNORM; a debug view is encoded only. */
VKRT_HOST_API void vkrtHostLinearToDisplay16(const float* linear, uint32_t w, uint32_t h, uint32_t toneMappingMode, float exposure, uint32_t debugMode, uint16_t* out) {
    if (!isfinite(exposure) || exposure < 0.0f) exposure = 1.0f;
    const size_t pixels = (size_t)w * h;
    for (size_t i = 0; i < pixels; i++) {
        float c[3] = {linear[i * 4u], linear[i * 4u + 1u], linear[i * 4u + 2u]};
        if (debugMode == VKRT_DEBUG_MODE_NONE) {
            for (int k = 0; k < 3; k++) c[k] *= exposure;
            if (toneMappingMode == VKRT_TONE_MAPPING_MODE_ACES)
                for (int k = 0; k < 3; k++) c[k] = (c[k] * ((2.51f * c[k]) + 0.03f)) / ((c[k] * ((2.43f * c[k]) + 0.59f)) + 0.14f);
        }
        for (int k = 0; k < 3; k++) {
            float v = c[k] < 0.0f ? 0.0f : (c[k] > 1.0f ? 1.0f : c[k]); /* (a NaN passes both comparisons, as in the reference's clampf) */
            v = v <= 0.0031308f ? 12.92f * v : (1.055f * powf(v, 1.0f / 2.4f)) - 0.055f;
            v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            out[i * 4u + (size_t)k] = (uint16_t)((v * 65535.0f) + 0.5f);
        }
        out[i * 4u + 3u] = 65535u;
    }
}
