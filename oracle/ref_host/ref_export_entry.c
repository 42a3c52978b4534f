/* ref_export_entry.c — the reference's OWN save-time denoise / export stage compiled for the parity tests (TEST INFRASTRUCTURE).
 *
 * oracle/Makefile `ref` compiles this file together with /root/reference/src/core/utility/denoise.c into oracle/_ref/libvkrt_refexport.so,
 * linked against the stand-in OIDN (fake_oidn.c). The reference's export/image.c is #included where it lies, so that its static
 * functions — prepareLinearRenderOutput (image.c:907-960: spectral XYZ -> sRGB, denoiseLinearRenderOutput with feature preparation and
 * prefilter, sanitise) and convertLinearToDisplayRGBA16 (image.c:641-699) — can be called; its file writers (libspng, libjpeg-turbo,
 * tinyexr) are stubbed and never reached. Nothing of the reference is copied into this repository. */
#include "utility/export/image.c"

/* ---- stubs for what image.c links against but these entry points never call ---- */
spng_ctx* spng_ctx_new(int flags) { (void)flags; return NULL; }
void spng_ctx_free(spng_ctx* ctx) { (void)ctx; }
int spng_set_png_file(spng_ctx* ctx, FILE* file) { (void)ctx; (void)file; return 1; }
int spng_set_ihdr(spng_ctx* ctx, struct spng_ihdr* ihdr) { (void)ctx; (void)ihdr; return 1; }
int spng_encode_image(spng_ctx* ctx, const void* img, size_t len, int fmt, int flags) { (void)ctx; (void)img; (void)len; (void)fmt; (void)flags; return 1; }
const char* spng_strerror(int err) { (void)err; return "stub"; }
tjhandle tjInitCompress(void) { return NULL; }
int tjCompress2(tjhandle h, const unsigned char* s, int w, int p, int hh, int pf, unsigned char** jb, unsigned long* js, int ss, int q, int f) {
    (void)h; (void)s; (void)w; (void)p; (void)hh; (void)pf; (void)jb; (void)js; (void)ss; (void)q; (void)f; return -1;
}
int tjDestroy(tjhandle h) { (void)h; return 0; }
void tjFree(unsigned char* b) { (void)b; }
char* tjGetErrorStr(void) { return (char*)"stub"; }
char* tjGetErrorStr2(tjhandle h) { (void)h; return (char*)"stub"; }
int vkrtWriteEXRFromRGBA32F(const char* path, const float* rgba32f, uint32_t width, uint32_t height) { (void)path; (void)rgba32f; (void)width; (void)height; return 0; }
int vkrtTryComputeImageByteCount(uint32_t width, uint32_t height, uint32_t channels, size_t* outByteCount) { (void)width; (void)height; (void)channels; (void)outByteCount; return 0; }
char* stringDuplicate(const char* value) { return value ? strdup(value) : NULL; }
const char* pathBasename(const char* path) { const char* s = path ? strrchr(path, '/') : NULL; return s ? s + 1 : path; }
void vkrtLogLine(FILE* stream, const char* level, const char* format, ...) { (void)stream; (void)level; (void)format; }
int vkrtInfoLoggingEnabled(void) { return 0; }

#define REFEXPORT __attribute__((visibility("default")))

/* prepareLinearRenderOutput on caller-provided buffers: beauty = the accumulation read-back (RGBA32F), albedo / normal = the RGBA16F
 * feature AOVs (NULL = not read back). Returns the function's result (1 = linear image produced) and copies the image out. */
REFEXPORT int refexport_prepare_linear(const float* beautyRgba32f, const uint16_t* albedoRgba16f, const uint16_t* normalRgba16f, uint32_t width, uint32_t height,
                                       uint32_t renderMode, uint32_t debugMode, int denoiseEnabled, int allowRawFallback, float* outLinear) {
    RenderImageBuffer beauty = {(void*)beautyRgba32f, RENDER_IMAGE_BUFFER_FORMAT_RGBA32F};
    RenderImageBuffer albedo = {(void*)albedoRgba16f, RENDER_IMAGE_BUFFER_FORMAT_RGBA16F};
    RenderImageBuffer normal = {(void*)normalRgba16f, RENDER_IMAGE_BUFFER_FORMAT_RGBA16F};
    VKRT_RenderExportSettings settings;
    memset(&settings, 0, sizeof(settings));
    settings.denoiseEnabled = denoiseEnabled ? 1u : 0u;
    VKRT_SceneSettingsSnapshot scene;
    memset(&scene, 0, sizeof(scene));
    scene.renderMode = renderMode;
    scene.debugMode = debugMode;
    LinearRenderOutputRequest request = {
        .label = "parity test", .beautyBuffer = &beauty, .albedoBuffer = &albedo, .normalBuffer = &normal, .width = width, .height = height,
        .settings = &settings, .sceneSettings = &scene, .allowRawFallback = allowRawFallback,
    };
    float* linear = NULL;
    int ok = prepareLinearRenderOutput(&request, &linear);
    if (ok && linear) memcpy(outLinear, linear, (size_t)width * height * 16u);
    free(linear);
    return ok;
}

REFEXPORT int refexport_linear_to_display(const float* linear, uint32_t width, uint32_t height, uint32_t toneMappingMode, float exposure, uint32_t debugMode, uint16_t* outRgba16) {
    VKRT_SceneSettingsSnapshot scene;
    memset(&scene, 0, sizeof(scene));
    scene.toneMappingMode = toneMappingMode;
    scene.exposure = exposure;
    scene.debugMode = debugMode;
    uint16_t* pixels = NULL;
    int ok = convertLinearToDisplayRGBA16(linear, (size_t)width * height * 16u, width, height, &scene, &pixels);
    if (ok && pixels) memcpy(outRgba16, pixels, (size_t)width * height * 8u);
    free(pixels);
    return ok;
}
