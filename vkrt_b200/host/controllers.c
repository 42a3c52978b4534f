/* controllers.c — the two host-side feedback loops around the hot path:
 *   auto-SPP      src/core/scene/timing.c:67-178   samples per pixel per frame from the measured render time and a frame-time budget
 *   auto-exposure src/core/scene/exposure.c:14-67,187-222   exposure from a 16 x 16 luminance probe of the accumulation image
 * Both are pure arithmetic on a few floats; the device contributes the frame time (CUDA events) and the probe read
 * (vkrt_cuda_read_accum_samples). The step functions are exported so that front ends that own the clock, and the tests, can drive them. */
#include <math.h>
#include <string.h>

#include "host_state.h"

/* timing.c:121-178 updateAutoSPP. *ioControlMs is the smoothed ms-per-spp estimate (0 = no history). Returns the next spp. */
uint32_t vkrtAutoSPPStep(float* ioControlMs, float targetMs, float measuredFrameMs, uint32_t spp) {
    const float measurementSmoothing = 0.35f, budgetScale = 0.90f, upwardDeadbandScale = 0.18f, downwardDeadbandScale = 0.08f;
    const float maxUpwardScale = 1.25f, maxDownwardScale = 0.60f;
    if (!ioControlMs || targetMs <= 0.0f || measuredFrameMs <= 0.0f) return spp;
    if (spp == 0) spp = 1;
    const float sppf = (float)spp;
    const float measuredMsPerSPP = measuredFrameMs / sppf;
    if (measuredMsPerSPP <= 0.0f) return spp;
    if (*ioControlMs <= 0.0f) *ioControlMs = measuredMsPerSPP;
    else *ioControlMs = (*ioControlMs * (1.0f - measurementSmoothing)) + (measuredMsPerSPP * measurementSmoothing);
    float desired = (targetMs * budgetScale) / *ioControlMs;
    if (desired < 1.0f) desired = 1.0f;
    if (desired > 2048.0f) desired = 2048.0f;
    const float delta = desired - sppf;
    const float deadband = fmaxf(sppf * (delta > 0.0f ? upwardDeadbandScale : downwardDeadbandScale), 1.0f);
    if (fabsf(delta) <= deadband) return spp;
    uint32_t next;
    if (delta > 0.0f) {
        const float limited = fminf(desired, ceilf(sppf * maxUpwardScale));
        next = (uint32_t)floorf(limited);
        if (next <= spp && spp < 2048u) next = spp + 1u;
    } else {
        const float limited = fmaxf(desired, floorf(sppf * maxDownwardScale));
        next = (uint32_t)ceilf(limited);
        if (next >= spp && spp > 1u) next = spp - 1u;
    }
    if (next < 1u) next = 1u;
    if (next > 2048u) next = 2048u;
    return next;
}

/* exposure.c:41-58,187-222: average luminance of the finite probe samples -> exponential filter -> key / L^0.65.
 * Returns 1 and writes *outExposure when the exposure should change (by at least 1e-4). */
int vkrtAutoExposureStep(float* ioFilteredLuminance, const float* samplesRgba, uint32_t sampleCount, float currentExposure, float* outExposure) {
    const float key = 0.18f, smoothing = 0.18f, adaptationStrength = 0.65f;
    if (!ioFilteredLuminance || !samplesRgba || !outExposure) return 0;
    float sum = 0.0f;
    uint32_t count = 0;
    for (uint32_t i = 0; i < sampleCount; i++) {
        const float* s = samplesRgba + (size_t)i * 4u;
        const float luminance = 0.2126f * s[0] + 0.7152f * s[1] + 0.0722f * s[2];   /* scene/color.h linearSRGBLuminance */
        if (!isfinite(luminance)) continue;
        sum += luminance;
        count++;
    }
    if (count == 0) return 0;
    const float average = sum / (float)count;
    if (!isfinite(average) || average <= 0.0f) return 0;
    *ioFilteredLuminance = *ioFilteredLuminance <= 0.0f ? average : (*ioFilteredLuminance * (1.0f - smoothing)) + (average * smoothing);
    if (!isfinite(*ioFilteredLuminance) || *ioFilteredLuminance <= 0.0f) return 0;
    float adapted = powf(*ioFilteredLuminance, adaptationStrength);
    adapted = fmaxf(adapted, 1e-4f);
    const float next = key / adapted;
    if (!isfinite(next) || fabsf(currentExposure - next) < 1e-4f) return 0;
    *outExposure = next;
    return 1;
}

/* the probe positions of exposure.c:139-147: centres of a 16 x 16 grid over the render extent */
void vkrtAutoExposureProbePixels(uint32_t width, uint32_t height, uint32_t* outXY /* 256 pairs */) {
    uint32_t k = 0;
    for (uint32_t y = 0; y < 16u; y++)
        for (uint32_t x = 0; x < 16u; x++) {
            uint32_t sx = (((2u * x) + 1u) * width) / 32u, sy = (((2u * y) + 1u) * height) / 32u;
            if (sx >= width) sx = width - 1u;
            if (sy >= height) sy = height - 1u;
            outXY[k++] = sx;
            outXY[k++] = sy;
        }
}

/* ---- wiring into the frame protocol --------------------------------------------------------------------------------------- */
void hostResetAutoSPPState(VKRT* v, int resetSamplesPerPixel) {   /* timing.c:67-78 */
    if (!v) return;
    if (resetSamplesPerPixel && v->sceneSettings.autoSPPEnabled) v->sceneSettings.samplesPerPixel = 1u;
    v->autoSPPControlMs = 0.0f;
    v->renderStatus.renderTimeMs = 0.0f;
}
void hostUpdateAutoSPP(VKRT* v) {   /* frame.c:381 */
    if (!v || !v->sceneSettings.autoSPPEnabled) return;
    float targetMs = v->autoSPPTargetFrameMs > 0.0f ? v->autoSPPTargetFrameMs : 1000.0f / (float)(v->sceneSettings.autoSPPTargetFPS ? v->sceneSettings.autoSPPTargetFPS : 60u);
    uint32_t next = vkrtAutoSPPStep(&v->autoSPPControlMs, targetMs, v->renderStatus.renderTimeMs, v->sceneSettings.samplesPerPixel);
    if (next != v->sceneSettings.samplesPerPixel) {
        v->sceneSettings.samplesPerPixel = next;
        hostWriteSceneStateUniform(v);
    }
}
/* record.c:592 + frame.c:102: probe after the trace, resolve before the next frame. The device call is synchronous here, so both
 * halves run back to back after VKRT_trace. Single rank only: a tile-partitioned film holds a fraction of the probe pixels. */
VKRT_Result hostUpdateAutoExposure(VKRT* v) {
    if (!v || v->hostOnly || !v->sceneSettings.autoExposureEnabled || v->sceneSettings.debugMode != VKRT_DEBUG_MODE_NONE) return VKRT_SUCCESS;
    if (v->createInfo.worldSize > 1u || v->renderWidth == 0 || v->renderHeight == 0) return VKRT_SUCCESS;
    uint32_t xy[512];
    float samples[1024];
    vkrtAutoExposureProbePixels(v->renderWidth, v->renderHeight, xy);
    VKRT_Result r = vkrt_cuda_read_accum_samples(v->cuda, xy, 256u, samples);
    if (r != VKRT_SUCCESS) return hostFail(v, r, "read_accum_samples: %s", vkrt_cuda_last_error(v->cuda));
    float next = 0.0f;
    if (vkrtAutoExposureStep(&v->autoExposureFilteredLuminance, samples, 256u, v->sceneSettings.exposure, &next)) {
        v->sceneSettings.exposure = next;
        hostWriteSceneStateUniform(v);
    }
    return VKRT_SUCCESS;
}
