/* hmath.h — the handful of fp32 vector/matrix routines the C host needs.
 *
 * The reference does this arithmetic with cglm (external/cglm, scalar paths): glm_translate / glm_rotate / glm_scale in
 * VKRT_buildMeshTransformMatrix (src/core/scene/transform.c:26-35), glm_lookat / glm_perspective / glm_mat4_inv in
 * syncCameraMatrices (src/core/scene/camera.c:128-143). These are restatements of the same formulas (operation order follows
 * the published cglm scalar code so that results agree to the last bit where no SIMD path is involved), written for this repo.
 * mat4 is column-major: m[col][row], the memory image of cglm's mat4 and of SceneData.viewInverse / projInverse. */
#ifndef VKRT_HOST_HMATH_H
#define VKRT_HOST_HMATH_H

#include <math.h>
#include <string.h>

typedef float hvec3[3];
typedef float hmat4[4][4];

#define H_PI 3.14159265358979323846264338327950288f

static inline float h_rad(float deg) { return deg * H_PI / 180.0f; }
static inline float h_deg(float rad) { return rad * 180.0f / H_PI; }

static inline float h_dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline float h_norm3(const float* v) { return sqrtf(h_dot3(v, v)); }
static inline void h_cross3(const float* a, const float* b, float* d) {
    float c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
    d[0] = c0; d[1] = c1; d[2] = c2;
}
static inline void h_normalize3(float* v) {
    float n = h_norm3(v);
    if (n < 1.1920928955078125e-7f) { v[0] = v[1] = v[2] = 0.0f; return; }
    float inv = 1.0f / n;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

static inline void h_mat4_identity(hmat4 m) {
    memset(m, 0, sizeof(hmat4));
    m[0][0] = m[1][1] = m[2][2] = m[3][3] = 1.0f;
}
static inline void h_mat4_copy(hmat4 src, hmat4 dst) { memcpy(dst, src, sizeof(hmat4)); }

/* dest = a * b (column vectors) */
static inline void h_mat4_mul(hmat4 a, hmat4 b, hmat4 dest) {
    hmat4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++)
            r[c][row] = a[0][row] * b[c][0] + a[1][row] * b[c][1] + a[2][row] * b[c][2] + a[3][row] * b[c][3];
    memcpy(dest, r, sizeof(hmat4));
}

/* m = m * T(v) */
static inline void h_translate(hmat4 m, const float* v) {
    for (int row = 0; row < 4; row++) m[3][row] = m[0][row] * v[0] + m[1][row] * v[1] + m[2][row] * v[2] + m[3][row];
}
/* m = m * S(v) */
static inline void h_scale(hmat4 m, const float* v) {
    for (int row = 0; row < 4; row++) { m[0][row] *= v[0]; m[1][row] *= v[1]; m[2][row] *= v[2]; }
}
/* Axis-angle rotation matrix (Rodrigues), then m = m * R on the upper 3 columns. */
static inline void h_rotate(hmat4 m, float angle, const float* axis) {
    float c = cosf(angle), s = sinf(angle);
    float an[3] = {axis[0], axis[1], axis[2]};
    h_normalize3(an);
    float v[3] = {an[0] * (1.0f - c), an[1] * (1.0f - c), an[2] * (1.0f - c)};
    float vs[3] = {an[0] * s, an[1] * s, an[2] * s};
    float r[3][3];
    for (int k = 0; k < 3; k++) { r[0][k] = an[k] * v[0]; r[1][k] = an[k] * v[1]; r[2][k] = an[k] * v[2]; }
    r[0][0] += c;     r[1][0] -= vs[2]; r[2][0] += vs[1];
    r[0][1] += vs[2]; r[1][1] += c;     r[2][1] -= vs[0];
    r[0][2] -= vs[1]; r[1][2] += vs[0]; r[2][2] += c;
    float a[3][4];
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 4; row++) a[col][row] = m[col][row];
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 4; row++) m[col][row] = a[0][row] * r[col][0] + a[1][row] * r[col][1] + a[2][row] * r[col][2];
}

/* General 4x4 inverse (fp32). The reference calls cglm's glm_mat4_inv (src/core/scene/camera.c:141-142), which on x86-64 is the SSE2
 * routine glm_mat4_inv_sse2 (cglm simd/sse2/mat4.h, built without FMA): twelve 2x2 sub-determinants c1..c12 (product - product), the
 * determinant as ((c2 c7 + c3 c10) + (c1 c8 + c4 c9)) - (c12 c5 + c11 c6), the sub-determinants scaled by 1 / det BEFORE the cofactor
 * sums, each cofactor as (x cA - y cB) + z cC, signs applied by flipping the sign bit. This scalar code performs the same IEEE
 * operations in the same order, lane by lane, so that viewInverse / projInverse -- and with them every primary ray -- are bit-identical
 * to the reference's (tests/test_reference_pin.py::test_camera_matrices_are_bit_identical_to_the_reference). */
static inline void h_mat4_inv(hmat4 mat, hmat4 dest) {
    const float a = mat[0][0], b = mat[0][1], c = mat[0][2], d = mat[0][3], e = mat[1][0], f = mat[1][1], g = mat[1][2], h = mat[1][3],
                i = mat[2][0], j = mat[2][1], k = mat[2][2], l = mat[2][3], m = mat[3][0], n = mat[3][1], o = mat[3][2], p = mat[3][3];
    float c1 = k * p - l * o, c2 = c * h - d * g, c3 = i * p - l * m, c4 = a * h - d * e;
    float c5 = j * p - l * n, c6 = b * h - d * f, c11 = i * o - k * m, c12 = a * g - c * e;
    float c7 = i * n - j * m, c8 = a * f - b * e, c9 = j * o - k * n, c10 = b * g - c * f;
    const float det = ((c2 * c7 + c3 * c10) + (c1 * c8 + c4 * c9)) - (c12 * c5 + c11 * c6);
    const float idt = 1.0f / det;
    c1 *= idt; c2 *= idt; c3 *= idt; c4 *= idt; c5 *= idt; c6 *= idt;
    c7 *= idt; c8 *= idt; c9 *= idt; c10 *= idt; c11 *= idt; c12 *= idt;
    hmat4 r;
    r[0][0] = (f * c1 - g * c5) + h * c9;
    r[0][1] = -((b * c1 - c * c5) + d * c9);
    r[0][2] = (n * c2 - o * c6) + p * c10;
    r[0][3] = -((j * c2 - k * c6) + l * c10);
    r[1][0] = -((e * c1 - g * c3) + h * c11);
    r[1][1] = (a * c1 - c * c3) + d * c11;
    r[1][2] = -((m * c2 - o * c4) + p * c12);
    r[1][3] = (i * c2 - k * c4) + l * c12;
    r[2][0] = (e * c5 - f * c3) + h * c7;
    r[2][1] = -((a * c5 - b * c3) + d * c7);
    r[2][2] = (m * c6 - n * c4) + p * c8;
    r[2][3] = -((i * c6 - j * c4) + l * c8);
    r[3][0] = -((e * c9 - f * c11) + g * c7);
    r[3][1] = (a * c9 - b * c11) + c * c7;
    r[3][2] = -((m * c10 - n * c12) + o * c8);
    r[3][3] = (i * c10 - j * c12) + k * c8;
    memcpy(dest, r, sizeof(hmat4));
}

/* Right-handed look-at. */
static inline void h_lookat(const float* eye, const float* center, const float* up, hmat4 dest) {
    float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
    h_normalize3(f);
    float s[3], u[3];
    h_cross3(f, up, s);
    h_normalize3(s);
    h_cross3(s, f, u);
    dest[0][0] = s[0]; dest[0][1] = u[0]; dest[0][2] = -f[0]; dest[0][3] = 0.0f;
    dest[1][0] = s[1]; dest[1][1] = u[1]; dest[1][2] = -f[1]; dest[1][3] = 0.0f;
    dest[2][0] = s[2]; dest[2][1] = u[2]; dest[2][2] = -f[2]; dest[2][3] = 0.0f;
    dest[3][0] = -h_dot3(s, eye); dest[3][1] = -h_dot3(u, eye); dest[3][2] = h_dot3(f, eye); dest[3][3] = 1.0f;
}

/* Right-handed perspective, clip z in [-1, 1] (GL convention; the shader only ever un-projects ndc z = 1). */
static inline void h_perspective(float fovy, float aspect, float nearZ, float farZ, hmat4 dest) {
    memset(dest, 0, sizeof(hmat4));
    float f = 1.0f / tanf(fovy * 0.5f);
    float fn = 1.0f / (nearZ - farZ);
    dest[0][0] = f / aspect;
    dest[1][1] = f;
    dest[2][2] = (nearZ + farZ) * fn;
    dest[2][3] = -1.0f;
    dest[3][2] = 2.0f * nearZ * farZ * fn;
}

#endif
