/* ref_exr_entry.c — what the reference's EXR codec (src/core/utility/exr.cpp over the vendored tinyexr) links against besides io.c / platform.c:
 * the logger (TEST INFRASTRUCTURE, libvkrt_refexr.so; the codec's own entry points are already extern "C"). Silent unless REFHOST_LOG is set. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
__attribute__((visibility("default"))) const char* refexr_version(void) { return "reference exr.cpp + tinyexr"; }
void vkrtLogLine(FILE* stream, const char* level, const char* format, ...) {
    if (!getenv("REFHOST_LOG")) return;
    va_list args;
    va_start(args, format);
    fprintf(stderr, "%s ", level ? level : "");
    vfprintf(stderr, format, args);
    fputc('\n', stderr);
    va_end(args);
}
