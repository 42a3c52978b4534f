/* fake_oidn.c — a stand-in for libOpenImageDenoise (TEST INFRASTRUCTURE; the product never links or ships it).
 *
 * Open Image Denoise is a third-party neural denoiser that is not installed here, and its output could not be compared anyway. What CAN
 * be compared is everything around it: which images the caller attaches under which names, strides and formats, which filter
 * parameters it sets, that the output buffer is seeded, and what it does with the result or with a failure. This library exports the
 * part of the OIDN 2 C API that src/core/utility/denoise.c calls and "filters" with a small deterministic function in which every input
 * and every parameter is visible, so that the reference's own denoise stage (oracle/_ref/libvkrt_refexport.so, linked against this
 * file's library) and the product's (vkrt_b200/host/denoise.c, which dlopens it through VKRT_OIDN_LIBRARY) must produce the same bytes:
 *
 *   out.rgb = box3x3(main).rgb * (hdr ? 1 : 0.5) + 0.125 * albedo.rgb + 0.0625 * normal.rgb + (cleanAux ? 1/32 : 0)
 *             + (srgb ? 500 : 0) + (quality == HIGH ? 0 : 1000);   out.a is not written (FLOAT3): the seed stays
 *
 * FAKE_OIDN_FAIL = "device" | "execute" | "prefilter" | "read" makes the corresponding step fail the way OIDN reports failures;
 * FAKE_OIDN_LOG = file appends one line per executed filter with everything the caller set. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

typedef struct Device { int error; char message[128]; int committed; } Device;
typedef struct Buffer { float* data; size_t bytes; Device* device; } Buffer;
typedef struct Image { Buffer* buffer; int format; size_t w, h, offset, pixelStride, rowStride; int set; } Image;
typedef struct Filter {
    Device* device;
    Image color, albedo, normal, output;
    int hdr, srgb, cleanAux, quality, committed;
} Filter;

static int failMode(const char* what) {
    const char* f = getenv("FAKE_OIDN_FAIL");
    return f && strcmp(f, what) == 0;
}
static void setError(Device* d, const char* msg) {
    if (!d || d->error) return; /* the first error sticks until it is read */
    d->error = 1;               /* OIDN_ERROR_UNKNOWN */
    snprintf(d->message, sizeof(d->message), "%s", msg);
}

API void* oidnNewDevice(int type) {
    if (type != 1 /* OIDN_DEVICE_TYPE_CPU */ || failMode("device")) return NULL;
    return calloc(1, sizeof(Device));
}
API void oidnCommitDevice(void* d) { ((Device*)d)->committed = 1; }
API void oidnSyncDevice(void* d) { (void)d; }
API void oidnReleaseDevice(void* d) { free(d); }
API int oidnGetDeviceError(void* dv, const char** outMessage) {
    Device* d = (Device*)dv;
    static __thread char last[128];
    int e = d->error;
    snprintf(last, sizeof(last), "%s", d->message);
    if (outMessage) *outMessage = e ? last : NULL;
    d->error = 0;
    d->message[0] = 0;
    return e;
}
API void* oidnNewFilter(void* d, const char* type) {
    if (!d || !type || strcmp(type, "RT") != 0) return NULL;
    Filter* f = (Filter*)calloc(1, sizeof(Filter));
    f->device = (Device*)d;
    f->hdr = f->srgb = f->cleanAux = 0;
    f->quality = 0;
    return f;
}
API void oidnReleaseFilter(void* f) { free(f); }
API void* oidnNewBuffer(void* d, size_t bytes) {
    Buffer* b = (Buffer*)calloc(1, sizeof(Buffer));
    b->device = (Device*)d;
    b->data = (float*)malloc(bytes ? bytes : 1);
    memset(b->data, 0xA5, bytes); /* an unseeded output buffer shows */
    b->bytes = bytes;
    return b;
}
API void oidnReleaseBuffer(void* bv) {
    Buffer* b = (Buffer*)bv;
    if (b) free(b->data);
    free(b);
}
API void oidnWriteBuffer(void* bv, size_t offset, size_t bytes, const void* src) {
    Buffer* b = (Buffer*)bv;
    if (offset + bytes <= b->bytes) memcpy((char*)b->data + offset, src, bytes);
}
API void oidnReadBuffer(void* bv, size_t offset, size_t bytes, void* dst) {
    Buffer* b = (Buffer*)bv;
    if (failMode("read")) { setError(b->device, "simulated read-back failure"); return; }
    if (offset + bytes <= b->bytes) memcpy(dst, (char*)b->data + offset, bytes);
}
API void oidnSetFilterImage(void* fv, const char* name, void* buffer, int format, size_t w, size_t h, size_t offset, size_t pixelStride, size_t rowStride) {
    Filter* f = (Filter*)fv;
    Image* im = !strcmp(name, "color") ? &f->color : !strcmp(name, "albedo") ? &f->albedo : !strcmp(name, "normal") ? &f->normal : !strcmp(name, "output") ? &f->output : NULL;
    if (!im) { setError(f->device, "unknown filter image"); return; }
    im->buffer = (Buffer*)buffer; im->format = format; im->w = w; im->h = h; im->offset = offset; im->pixelStride = pixelStride; im->rowStride = rowStride; im->set = 1;
    f->committed = 0;
}
API void oidnSetFilterBool(void* fv, const char* name, _Bool value) {
    Filter* f = (Filter*)fv;
    if (!strcmp(name, "hdr")) f->hdr = value; else if (!strcmp(name, "srgb")) f->srgb = value; else if (!strcmp(name, "cleanAux")) f->cleanAux = value;
    else setError(f->device, "unknown filter parameter");
    f->committed = 0;
}
API void oidnSetFilterInt(void* fv, const char* name, int value) {
    Filter* f = (Filter*)fv;
    if (!strcmp(name, "quality")) f->quality = value; else setError(f->device, "unknown filter parameter");
    f->committed = 0;
}
API void oidnCommitFilter(void* fv) { ((Filter*)fv)->committed = 1; }

static const float* texel(const Image* im, size_t x, size_t y) { return (const float*)((const char*)im->buffer->data + im->offset + y * im->rowStride + x * im->pixelStride); }

API void oidnExecuteFilter(void* fv) {
    Filter* f = (Filter*)fv;
    Device* d = f->device;
    const Image* mainImage = f->color.set ? &f->color : (f->albedo.set ? &f->albedo : (f->normal.set ? &f->normal : NULL));
    const char* mainName = f->color.set ? "color" : (f->albedo.set ? "albedo" : "normal");
    if (!f->committed || !d->committed) { setError(d, "filter or device not committed"); return; }
    if (!mainImage || !f->output.set) { setError(d, "main or output image missing"); return; }
    const Image* aux[2] = {(f->color.set && f->albedo.set) ? &f->albedo : NULL, (f->color.set && f->normal.set) ? &f->normal : NULL};
    const Image* all[4] = {mainImage, &f->output, aux[0], aux[1]};
    for (int i = 0; i < 4; i++) {
        const Image* im = all[i];
        if (!im) continue;
        if (im->format != 3 || im->w != mainImage->w || im->h != mainImage->h || im->pixelStride < 12 || im->rowStride < im->pixelStride * im->w ||
            !im->buffer || im->offset + im->rowStride * im->h > im->buffer->bytes) { setError(d, "inconsistent image description"); return; }
    }
    const char* logPath = getenv("FAKE_OIDN_LOG");
    if (logPath && logPath[0]) {
        FILE* lf = fopen(logPath, "a");
        if (lf) {
            fprintf(lf, "RT main=%s albedo=%d normal=%d %zux%zu pixelStride=%zu rowStride=%zu hdr=%d srgb=%d cleanAux=%d quality=%d\n", mainName, aux[0] != NULL, aux[1] != NULL,
                    mainImage->w, mainImage->h, mainImage->pixelStride, mainImage->rowStride, f->hdr, f->srgb, f->cleanAux, f->quality);
            fclose(lf);
        }
    }
    if (failMode("execute") || (failMode("prefilter") && !f->color.set)) { setError(d, "simulated filter failure"); return; }
    const size_t w = mainImage->w, h = mainImage->h;
    const float gain = f->hdr ? 1.0f : 0.5f;
    const float bias = (f->cleanAux ? 0.03125f : 0.0f) + (f->srgb ? 500.0f : 0.0f) + (f->quality == 6 ? 0.0f : 1000.0f);
    float* result = (float*)malloc(w * h * 3 * sizeof(float)); /* (the output may alias nothing here, but a filter reads all inputs first) */
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) {
                float sum = 0.0f;
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        size_t xx = (size_t)((long)x + dx < 0 ? 0 : ((size_t)((long)x + dx) >= w ? w - 1 : (size_t)((long)x + dx)));
                        size_t yy = (size_t)((long)y + dy < 0 ? 0 : ((size_t)((long)y + dy) >= h ? h - 1 : (size_t)((long)y + dy)));
                        sum = sum + texel(mainImage, xx, yy)[c];
                    }
                float v = (sum / 9.0f) * gain;
                if (aux[0]) v = v + 0.125f * texel(aux[0], x, y)[c];
                if (aux[1]) v = v + 0.0625f * texel(aux[1], x, y)[c];
                result[(y * w + x) * 3 + (size_t)c] = v + bias;
            }
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) memcpy((void*)texel(&f->output, x, y), result + (y * w + x) * 3, 12);
    free(result);
}
