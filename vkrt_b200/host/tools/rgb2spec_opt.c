/*
 * rgb2spec_opt — build-time generator for the sRGB -> sigmoid-polynomial spectrum table ("srgb.coeff").
 *
 * The reference embeds external/rgb2spec/srgb.coeff (src/core/meson.build:55-62), produced by mitsuba-renderer/rgb2spec's
 * rgb2spec_opt; both the blob and the optimiser source are absent from the reference checkout (.MISSING_LARGE_BLOBS:4,
 * external/rgb2spec/README.md).  This tool re-implements the published algorithm (Jakob & Hanika 2019, "A Low-Dimensional
 * Function Space for Efficient Spectral Upsampling": Gauss-Newton on 3 polynomial coefficients, CIELAB residual, warm
 * started along the brightness axis, smoothstep^2 scale) and writes the exact file format parsed by
 * src/core/scene/rgb2spec.c:17-59:  "SPEC", uint32 res, float scale[res], float coeff[3][res][res][res][3].
 *
 * Colour pipeline: the table is optimised against the renderer's OWN spectral-to-RGB path (src/shaders/utility/spectral.slang:
 * Wyman-Sloan-Shirley CIE 1931 fits, equal-energy white, division by 106.9461715, Bradford E->D65, XYZ->linear sRGB), so that
 * RGB -> spectrum -> XYZ -> RGB round-trips through the shaders' own arithmetic.  (The upstream blob targets tabulated CIE data
 * under D65; parity with it is unpinned, see DESIGN.md.)
 *
 * usage: rgb2spec_opt <res> <out.coeff> [threads]
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LAMBDA_MIN 360.0
#define LAMBDA_MAX 830.0
#define NSAMPLES 189 /* 2.5 nm trapezoid */

static double rgb_tbl[3][NSAMPLES];
static double lambda_tbl[NSAMPLES];
static const double XYZ_TO_RGB[3][3] = {{3.2404542, -1.5371385, -0.4985314}, {-0.9692660, 1.8760108, 0.0415560}, {0.0556434, -0.2040259, 1.0572252}};
static const double BRADFORD[3][3] = {{0.8951, 0.2664, -0.1614}, {-0.7502, 1.7135, 0.0367}, {0.0389, -0.0685, 1.0296}};
static const double BRADFORD_INV[3][3] = {{0.9869929, -0.1470543, 0.1599627}, {0.4323053, 0.5183603, 0.0492912}, {-0.0085287, 0.0400428, 0.9684867}};
static const double BRADFORD_SCALE[3] = {0.9413344, 1.0404175, 1.0895327};
static const double CIE_Y_INTEGRAL = 106.9461715;
static double RGB_TO_XYZ[3][3];
static const double WHITE_D65[3] = {0.95047, 1.0, 1.08883};

static double xfit(double l) {
    double t1 = (l - 442.0) * (l < 442.0 ? 0.0624 : 0.0374), t2 = (l - 599.8) * (l < 599.8 ? 0.0264 : 0.0323), t3 = (l - 501.1) * (l < 501.1 ? 0.0490 : 0.0382);
    return 0.362 * exp(-0.5 * t1 * t1) + 1.056 * exp(-0.5 * t2 * t2) - 0.065 * exp(-0.5 * t3 * t3);
}
static double yfit(double l) {
    double t1 = (l - 568.8) * (l < 568.8 ? 0.0213 : 0.0247), t2 = (l - 530.9) * (l < 530.9 ? 0.0613 : 0.0322);
    return 0.821 * exp(-0.5 * t1 * t1) + 0.286 * exp(-0.5 * t2 * t2);
}
static double zfit(double l) {
    double t1 = (l - 437.0) * (l < 437.0 ? 0.0845 : 0.0278), t2 = (l - 459.0) * (l < 459.0 ? 0.0385 : 0.0725);
    return 1.217 * exp(-0.5 * t1 * t1) + 0.681 * exp(-0.5 * t2 * t2);
}
static void mul3(const double m[3][3], const double v[3], double out[3]) {
    for (int r = 0; r < 3; r++) out[r] = m[r][0] * v[0] + m[r][1] * v[1] + m[r][2] * v[2];
}
static void invert3(const double m[3][3], double inv[3][3]) {
    double a = m[0][0], b = m[0][1], c = m[0][2], d = m[1][0], e = m[1][1], f = m[1][2], g = m[2][0], h = m[2][1], i = m[2][2];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g, det = a * A + b * B + c * C;
    inv[0][0] = A / det; inv[0][1] = -(b * i - c * h) / det; inv[0][2] = (b * f - c * e) / det;
    inv[1][0] = B / det; inv[1][1] = (a * i - c * g) / det; inv[1][2] = -(a * f - c * d) / det;
    inv[2][0] = C / det; inv[2][1] = -(a * h - b * g) / det; inv[2][2] = (a * e - b * d) / det;
}

static void init_tables(void) {
    invert3(XYZ_TO_RGB, RGB_TO_XYZ);
    double h = (LAMBDA_MAX - LAMBDA_MIN) / (NSAMPLES - 1);
    for (int i = 0; i < NSAMPLES; i++) {
        double l = LAMBDA_MIN + h * i;
        lambda_tbl[i] = l;
        double w = (i == 0 || i == NSAMPLES - 1) ? 0.5 * h : h;
        double xyz[3] = {fmax(xfit(l), 0.0), fmax(yfit(l), 0.0), fmax(zfit(l), 0.0)};
        double lms[3], ad[3], rgb[3];
        mul3(BRADFORD, xyz, lms);
        for (int k = 0; k < 3; k++) lms[k] *= BRADFORD_SCALE[k];
        mul3(BRADFORD_INV, lms, ad);
        mul3(XYZ_TO_RGB, ad, rgb);
        for (int k = 0; k < 3; k++) rgb_tbl[k][i] = rgb[k] * w / CIE_Y_INTEGRAL;
    }
}

static double smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }

static void cie_lab(double p[3]) {
    double xyz[3];
    mul3(RGB_TO_XYZ, p, xyz);
    double f[3];
    for (int i = 0; i < 3; i++) {
        double t = xyz[i] / WHITE_D65[i];
        const double delta = 6.0 / 29.0;
        f[i] = t > delta * delta * delta ? cbrt(t) : t / (3.0 * delta * delta) + 4.0 / 29.0;
    }
    p[0] = 116.0 * f[1] - 16.0;
    p[1] = 500.0 * (f[0] - f[1]);
    p[2] = 200.0 * (f[1] - f[2]);
}

static void eval_residual(const double c[3], const double rgb[3], double res[3]) {
    double out[3] = {0, 0, 0};
    for (int i = 0; i < NSAMPLES; i++) {
        double l = (lambda_tbl[i] - LAMBDA_MIN) / (LAMBDA_MAX - LAMBDA_MIN);
        double x = (c[0] * l + c[1]) * l + c[2];
        double s = 0.5 * x / sqrt(1.0 + x * x) + 0.5;
        out[0] += rgb_tbl[0][i] * s;
        out[1] += rgb_tbl[1][i] * s;
        out[2] += rgb_tbl[2][i] * s;
    }
    double t[3] = {rgb[0], rgb[1], rgb[2]};
    cie_lab(out);
    cie_lab(t);
    for (int j = 0; j < 3; j++) res[j] = t[j] - out[j];
}

static int solve3(double J[3][3], const double r[3], double x[3]) {
    double inv[3][3];
    double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (fabs(det) < 1e-15) return 0;
    invert3((const double(*)[3])J, inv);
    for (int i = 0; i < 3; i++) x[i] = inv[i][0] * r[0] + inv[i][1] * r[1] + inv[i][2] * r[2];
    return 1;
}

static void gauss_newton(const double rgb[3], double c[3]) {
    for (int it = 0; it < 15; it++) {
        double r0[3];
        eval_residual(c, rgb, r0);
        double J[3][3];
        const double eps = 1e-4;
        for (int i = 0; i < 3; i++) {
            double t[3] = {c[0], c[1], c[2]}, ra[3], rb[3];
            t[i] -= eps;
            eval_residual(t, rgb, ra);
            t[i] += 2 * eps;
            eval_residual(t, rgb, rb);
            for (int j = 0; j < 3; j++) J[j][i] = (rb[j] - ra[j]) / (2 * eps);
        }
        double x[3];
        if (!solve3(J, r0, x)) break;
        double r = 0;
        for (int j = 0; j < 3; j++) {
            c[j] -= x[j];
            r += r0[j] * r0[j];
        }
        double mx = fmax(fmax(c[0], c[1]), c[2]);
        if (mx > 200.0) for (int j = 0; j < 3; j++) c[j] *= 200.0 / mx;
        if (r < 1e-6) break;
    }
}

typedef struct { int res, l, jBegin, jEnd; const float* scale; float* out; } Job;

static void* worker(void* arg) {
    Job* job = (Job*)arg;
    const int res = job->res, l = job->l;
    const double c0 = LAMBDA_MIN, c1 = 1.0 / (LAMBDA_MAX - LAMBDA_MIN);
    for (int j = job->jBegin; j < job->jEnd; j++) {
        const double y = (double)j / (res - 1);
        for (int i = 0; i < res; i++) {
            const double x = (double)i / (res - 1);
            const int start = res / 5;
            for (int dir = 0; dir < 2; dir++) {
                double coeffs[3] = {0, 0, 0};
                for (int k = dir == 0 ? start : start; dir == 0 ? k < res : k >= 0; k += dir == 0 ? 1 : -1) {
                    double b = job->scale[k];
                    double rgb[3];
                    rgb[l] = b;
                    rgb[(l + 1) % 3] = x * b;
                    rgb[(l + 2) % 3] = y * b;
                    gauss_newton(rgb, coeffs);
                    double A = coeffs[0], B = coeffs[1], C = coeffs[2];
                    size_t idx = ((((size_t)l * res + k) * res + j) * res + i) * 3;
                    job->out[idx + 0] = (float)(A * c1 * c1);
                    job->out[idx + 1] = (float)(B * c1 - 2 * A * c0 * c1 * c1);
                    job->out[idx + 2] = (float)(C - B * c0 * c1 + A * (c0 * c1) * (c0 * c1));
                }
            }
        }
    }
    return NULL;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <res> <out.coeff> [threads]\n", argv[0]);
        return 2;
    }
    int res = atoi(argv[1]);
    int threads = argc > 3 ? atoi(argv[3]) : 8;
    if (res < 4 || res > 256 || threads < 1) return 2;
    init_tables();
    float* scale = (float*)malloc(sizeof(float) * res);
    for (int k = 0; k < res; k++) scale[k] = (float)smoothstep(smoothstep((double)k / (res - 1)));
    size_t n = (size_t)3 * res * res * res * 3;
    float* out = (float*)calloc(n, sizeof(float));
    for (int l = 0; l < 3; l++) {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
        Job* jobs = (Job*)malloc(sizeof(Job) * threads);
        for (int t = 0; t < threads; t++) {
            jobs[t] = (Job){res, l, (int)((long)res * t / threads), (int)((long)res * (t + 1) / threads), scale, out};
            pthread_create(&th[t], NULL, worker, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        free(th);
        free(jobs);
    }
    FILE* f = fopen(argv[2], "wb");
    if (!f) { perror("fopen"); return 1; }
    uint32_t r32 = (uint32_t)res;
    fwrite("SPEC", 4, 1, f);
    fwrite(&r32, 4, 1, f);
    fwrite(scale, sizeof(float), res, f);
    fwrite(out, sizeof(float), n, f);
    fclose(f);
    return 0;
}
