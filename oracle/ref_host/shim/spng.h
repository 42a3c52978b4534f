/* Type and prototype names of libspng that src/core/utility/export/image.c mentions (TEST INFRASTRUCTURE, oracle/Makefile `ref`): enough
 * to compile the reference's file where it lies; the writers are never called through libvkrt_refexport.so and the functions are stubs. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
typedef struct spng_ctx spng_ctx;
struct spng_ihdr { uint32_t width, height; uint8_t bit_depth, color_type, compression_method, filter_method, interlace_method; };
enum { SPNG_CTX_ENCODER = 2, SPNG_FMT_PNG = 256, SPNG_ENCODE_FINALIZE = 2, SPNG_COLOR_TYPE_TRUECOLOR_ALPHA = 6, SPNG_INTERLACE_NONE = 0 };
spng_ctx* spng_ctx_new(int flags);
void spng_ctx_free(spng_ctx* ctx);
int spng_set_png_file(spng_ctx* ctx, FILE* file);
int spng_set_ihdr(spng_ctx* ctx, struct spng_ihdr* ihdr);
int spng_encode_image(spng_ctx* ctx, const void* img, size_t len, int fmt, int flags);
const char* spng_strerror(int err);
