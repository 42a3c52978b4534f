/* hjson.h — a small JSON DOM (RFC 8259) for vkrt.scene documents and glTF JSON chunks.
 * The reference uses the vendored cJSON / cgltf for this (external/cjson, external/cgltf); neither is copied here. */
#ifndef VKRT_HOST_HJSON_H
#define VKRT_HOST_HJSON_H

#include <stddef.h>

typedef enum { HJ_NULL, HJ_BOOL, HJ_NUMBER, HJ_STRING, HJ_ARRAY, HJ_OBJECT } hj_type;

typedef struct hj_value {
    hj_type type;
    double number;      /* HJ_NUMBER; HJ_BOOL: 0/1 */
    char* string;       /* HJ_STRING (unescaped, NUL-terminated) */
    struct hj_value** items; /* HJ_ARRAY elements / HJ_OBJECT values */
    char** keys;        /* HJ_OBJECT keys */
    size_t count;
} hj_value;

hj_value* hj_parse(const char* text, size_t length, char* error, size_t errorSize);
void hj_free(hj_value* v);
const hj_value* hj_get(const hj_value* object, const char* key);              /* NULL if absent or not an object */
const hj_value* hj_at(const hj_value* array, size_t index);                   /* NULL if out of range */
size_t hj_count(const hj_value* v);                                           /* array/object size, else 0 */
double hj_number(const hj_value* v, double fallback);
int hj_bool(const hj_value* v, int fallback);
const char* hj_string(const hj_value* v, const char* fallback);
/* reads up to n numbers of an array into out; returns how many were read */
size_t hj_floats(const hj_value* array, float* out, size_t n);

#endif
