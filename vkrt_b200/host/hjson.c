/* hjson.c — recursive-descent JSON parser (see hjson.h). */
#include "hjson.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const char* p;
    const char* end;
    char* error;
    size_t errorSize;
    int depth;
} hj_parser;

static void hj_error(hj_parser* ps, const char* what) {
    if (ps->error && ps->errorSize && !ps->error[0]) snprintf(ps->error, ps->errorSize, "JSON: %s near offset %ld", what, (long)(ps->end - ps->p));
}
static void skipWs(hj_parser* ps) {
    while (ps->p < ps->end && (*ps->p == ' ' || *ps->p == '\t' || *ps->p == '\n' || *ps->p == '\r')) ps->p++;
}
static hj_value* newValue(hj_type t) {
    hj_value* v = (hj_value*)calloc(1, sizeof(hj_value));
    if (v) v->type = t;
    return v;
}
static hj_value* parseValue(hj_parser* ps);

static void appendUtf8(char** out, unsigned cp) {
    char* o = *out;
    if (cp < 0x80) *o++ = (char)cp;
    else if (cp < 0x800) { *o++ = (char)(0xC0 | (cp >> 6)); *o++ = (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { *o++ = (char)(0xE0 | (cp >> 12)); *o++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *o++ = (char)(0x80 | (cp & 0x3F)); }
    else { *o++ = (char)(0xF0 | (cp >> 18)); *o++ = (char)(0x80 | ((cp >> 12) & 0x3F)); *o++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *o++ = (char)(0x80 | (cp & 0x3F)); }
    *out = o;
}
static int hex4(const char* p, unsigned* out) {
    unsigned v = 0;
    for (int i = 0; i < 4; i++) {
        char c = p[i];
        v <<= 4;
        if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
        else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
        else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
        else return 0;
    }
    *out = v;
    return 1;
}
static char* parseStringRaw(hj_parser* ps) {
    if (ps->p >= ps->end || *ps->p != '"') { hj_error(ps, "expected string"); return NULL; }
    ps->p++;
    const char* start = ps->p;
    while (ps->p < ps->end && *ps->p != '"') {
        if (*ps->p == '\\') ps->p++;
        ps->p++;
    }
    if (ps->p >= ps->end) { hj_error(ps, "unterminated string"); return NULL; }
    size_t rawLen = (size_t)(ps->p - start);
    char* out = (char*)malloc(rawLen + 1);
    if (!out) return NULL;
    char* o = out;
    for (const char* s = start; s < ps->p; s++) {
        if (*s != '\\') { *o++ = *s; continue; }
        s++;
        switch (*s) {
            case 'n': *o++ = '\n'; break;
            case 't': *o++ = '\t'; break;
            case 'r': *o++ = '\r'; break;
            case 'b': *o++ = '\b'; break;
            case 'f': *o++ = '\f'; break;
            case 'u': {
                unsigned cp = 0;
                if (s + 4 < ps->p + 1 && hex4(s + 1, &cp)) {
                    s += 4;
                    if (cp >= 0xD800 && cp <= 0xDBFF && s + 6 < ps->p + 1 && s[1] == '\\' && s[2] == 'u') {
                        unsigned lo = 0;
                        if (hex4(s + 3, &lo) && lo >= 0xDC00 && lo <= 0xDFFF) { cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); s += 6; }
                    }
                    appendUtf8(&o, cp);
                }
                break;
            }
            default: *o++ = *s; break; /* \" \\ \/ */
        }
    }
    *o = 0;
    ps->p++; /* closing quote */
    return out;
}

static int pushItem(hj_value* v, char* key, hj_value* item) {
    hj_value** ni = (hj_value**)realloc(v->items, (v->count + 1) * sizeof(hj_value*));
    if (!ni) return 0;
    v->items = ni;
    if (v->type == HJ_OBJECT) {
        char** nk = (char**)realloc(v->keys, (v->count + 1) * sizeof(char*));
        if (!nk) return 0;
        v->keys = nk;
        v->keys[v->count] = key;
    }
    v->items[v->count++] = item;
    return 1;
}

static hj_value* parseValue(hj_parser* ps) {
    skipWs(ps);
    if (ps->p >= ps->end) { hj_error(ps, "unexpected end"); return NULL; }
    if (++ps->depth > 256) { hj_error(ps, "nesting too deep"); return NULL; }
    hj_value* v = NULL;
    char c = *ps->p;
    if (c == '{' || c == '[') {
        const char close = c == '{' ? '}' : ']';
        v = newValue(c == '{' ? HJ_OBJECT : HJ_ARRAY);
        ps->p++;
        skipWs(ps);
        if (ps->p < ps->end && *ps->p == close) { ps->p++; ps->depth--; return v; }
        while (v) {
            char* key = NULL;
            if (v->type == HJ_OBJECT) {
                skipWs(ps);
                key = parseStringRaw(ps);
                skipWs(ps);
                if (!key || ps->p >= ps->end || *ps->p != ':') { free(key); hj_error(ps, "expected ':'"); hj_free(v); return NULL; }
                ps->p++;
            }
            hj_value* item = parseValue(ps);
            if (!item || !pushItem(v, key, item)) { free(key); hj_free(item); hj_free(v); return NULL; }
            skipWs(ps);
            if (ps->p < ps->end && *ps->p == ',') { ps->p++; continue; }
            if (ps->p < ps->end && *ps->p == close) { ps->p++; break; }
            hj_error(ps, "expected ',' or closing bracket");
            hj_free(v);
            return NULL;
        }
    } else if (c == '"') {
        char* s = parseStringRaw(ps);
        if (!s) return NULL;
        v = newValue(HJ_STRING);
        if (v) v->string = s; else free(s);
    } else if (c == 't' && ps->end - ps->p >= 4 && !memcmp(ps->p, "true", 4)) { v = newValue(HJ_BOOL); if (v) v->number = 1; ps->p += 4; }
    else if (c == 'f' && ps->end - ps->p >= 5 && !memcmp(ps->p, "false", 5)) { v = newValue(HJ_BOOL); ps->p += 5; }
    else if (c == 'n' && ps->end - ps->p >= 4 && !memcmp(ps->p, "null", 4)) { v = newValue(HJ_NULL); ps->p += 4; }
    else {
        char buf[64];
        size_t n = 0;
        while (ps->p + n < ps->end && n < sizeof(buf) - 1 && strchr("+-0123456789.eE", ps->p[n])) n++;
        if (n == 0) { hj_error(ps, "unexpected character"); return NULL; }
        memcpy(buf, ps->p, n);
        buf[n] = 0;
        char* endp = NULL;
        double d = strtod(buf, &endp);
        if (endp == buf) { hj_error(ps, "bad number"); return NULL; }
        ps->p += (size_t)(endp - buf);
        v = newValue(HJ_NUMBER);
        if (v) v->number = d;
    }
    ps->depth--;
    return v;
}

hj_value* hj_parse(const char* text, size_t length, char* error, size_t errorSize) {
    if (error && errorSize) error[0] = 0;
    hj_parser ps = {text, text + length, error, errorSize, 0};
    hj_value* v = parseValue(&ps);
    if (v) {
        skipWs(&ps);
        while (ps.p < ps.end && *ps.p == 0) ps.p++; /* glTF chunks may be NUL/space padded */
        skipWs(&ps);
        if (ps.p != ps.end) { hj_error(&ps, "trailing characters"); hj_free(v); return NULL; }
    }
    return v;
}

void hj_free(hj_value* v) {
    if (!v) return;
    for (size_t i = 0; i < v->count; i++) {
        hj_free(v->items[i]);
        if (v->keys) free(v->keys[i]);
    }
    free(v->items);
    free(v->keys);
    free(v->string);
    free(v);
}

const hj_value* hj_get(const hj_value* o, const char* key) {
    if (!o || o->type != HJ_OBJECT) return NULL;
    for (size_t i = 0; i < o->count; i++)
        if (strcmp(o->keys[i], key) == 0) return o->items[i];
    return NULL;
}
const hj_value* hj_at(const hj_value* a, size_t i) { return (a && a->type == HJ_ARRAY && i < a->count) ? a->items[i] : NULL; }
size_t hj_count(const hj_value* v) { return (v && (v->type == HJ_ARRAY || v->type == HJ_OBJECT)) ? v->count : 0; }
double hj_number(const hj_value* v, double fallback) { return (v && (v->type == HJ_NUMBER || v->type == HJ_BOOL)) ? v->number : fallback; }
int hj_bool(const hj_value* v, int fallback) {
    if (!v) return fallback;
    if (v->type == HJ_BOOL || v->type == HJ_NUMBER) return v->number != 0.0;
    return fallback;
}
const char* hj_string(const hj_value* v, const char* fallback) { return (v && v->type == HJ_STRING) ? v->string : fallback; }
size_t hj_floats(const hj_value* a, float* out, size_t n) {
    size_t k = 0;
    if (!a || a->type != HJ_ARRAY) return 0;
    for (; k < n && k < a->count; k++) out[k] = (float)hj_number(a->items[k], 0.0);
    return k;
}
