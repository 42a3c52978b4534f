/* image_decode.h — PNG / JPEG / EXR decoding for textures and environment maps (reference: src/core/utility/image.h). */
#ifndef VKRT_HOST_IMAGE_DECODE_H
#define VKRT_HOST_IMAGE_DECODE_H
#include <stddef.h>
#include <stdint.h>

typedef struct HostImage {   /* VKRT_LoadedImage, image.h */
    void* pixels;            /* malloc'ed; RGBA8 / RGBA16 UNORM / RGBA16F / RGBA32F, row-major, top row first */
    uint32_t width, height;
    uint32_t format;         /* VKRT_TEXTURE_FORMAT_* */
    uint32_t colorSpace;     /* VKRT_TEXTURE_COLOR_SPACE_* actually stored */
} HostImage;

/* mimeType may be NULL (the codec is then taken from the signature bytes). Return 1 on success; on failure `err` holds the message. */
int hostDecodeImage(const void* data, size_t size, const char* mimeType, const char* label, uint32_t preferredColorSpace, HostImage* out, char* err,
                    size_t errLen);
int hostLoadImageFile(const char* path, uint32_t preferredColorSpace, HostImage* out, char* err, size_t errLen);
void hostFreeImage(HostImage* image);
/* baseline JPEG, 4:4:4, IJG quality scale (jpeg_encode.c; reference: export/image.c:220-263 writes quality 95). 1 on success. */
int hostWriteJpegRgba8(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height, int quality);
#endif
