/* ref_control_entry.c — the reference's OWN feedback controllers behind a flat C interface (TEST INFRASTRUCTURE, part of libvkrt_refhost.so).
 *
 * src/core/scene/timing.c (updateAutoSPP) and src/core/scene/exposure.c (the 16 x 16 probe grid recorded as 256 one-texel image copies, and
 * resolveAutoExposureReadback) are compiled unmodified; the two Vulkan commands exposure.c records are captured here instead of executed.
 * tests/test_reference_pin.py drives them step by step next to the product's vkrtAutoSPPStep / vkrtAutoExposureStep /
 * vkrtAutoExposureProbePixels (vkrt_b200/host/controllers.c). */
#include <string.h>

#include "vkrt_internal.h"
#include "scene.h"
#include "command/record.h"

#define REFHOST_API __attribute__((visibility("default")))

static VkBufferImageCopy g_regions[256];
static uint32_t g_regionCount = 0;

void transitionImageLayout(VkCommandBuffer commandBuffer, VkImage image, VkImageLayout oldLayout, VkImageLayout newLayout) {
    (void)commandBuffer; (void)image; (void)oldLayout; (void)newLayout;
}
void vkCmdCopyImageToBuffer(VkCommandBuffer commandBuffer, VkImage srcImage, VkImageLayout srcImageLayout, VkBuffer dstBuffer, uint32_t regionCount,
                            const VkBufferImageCopy* regions) {
    (void)commandBuffer; (void)srcImage; (void)srcImageLayout; (void)dstBuffer;
    g_regionCount = regionCount > 256u ? 256u : regionCount;
    memcpy(g_regions, regions, sizeof(VkBufferImageCopy) * g_regionCount);
}

/* the probe pixels of recordAutoExposureReadback for a width x height frame: outXY = 256 (x, y) pairs in buffer order; returns the count */
REFHOST_API uint32_t refcontrol_probe_pixels(void* h, uint32_t width, uint32_t height, uint32_t* outXY) {
    VKRT* vkrt = (VKRT*)h;
    vkrt->sceneSettings.autoExposureEnabled = 1u;
    vkrt->sceneSettings.debugMode = VKRT_DEBUG_MODE_NONE;
    vkrt->runtime.currentFrame = 0u;
    vkrt->renderControl.autoExposure.readbacks[0].buffer.buffer = (VkBuffer)(uintptr_t)1u;   /* "allocated" */
    g_regionCount = 0u;
    recordAutoExposureReadback(vkrt, (VkCommandBuffer)(uintptr_t)1u, (VkImage)(uintptr_t)1u, (VkExtent2D){width, height});
    for (uint32_t i = 0; i < g_regionCount; i++) {
        outXY[g_regions[i].bufferOffset / 16u * 2u + 0u] = (uint32_t)g_regions[i].imageOffset.x;
        outXY[g_regions[i].bufferOffset / 16u * 2u + 1u] = (uint32_t)g_regions[i].imageOffset.y;
    }
    vkrt->renderControl.autoExposure.readbacks[0].buffer.buffer = VK_NULL_HANDLE;
    vkrt->renderControl.autoExposure.readbacks[0].pending = 0u;
    return g_regionCount;
}
/* one resolveAutoExposureReadback over 256 RGBA samples, starting from (filtered luminance, exposure) and returning the new pair */
REFHOST_API void refcontrol_exposure_step(void* h, const float* samplesRgba, float* ioFilteredLuminance, float* ioExposure) {
    VKRT* vkrt = (VKRT*)h;
    vkrt->sceneSettings.autoExposureEnabled = 1u;
    vkrt->sceneSettings.debugMode = VKRT_DEBUG_MODE_NONE;
    vkrt->sceneSettings.exposure = *ioExposure;
    vkrt->renderControl.autoExposure.filteredLuminance = *ioFilteredLuminance;
    vkrt->renderControl.autoExposure.readbacks[0].mappedSamples = (float*)samplesRgba;
    vkrt->renderControl.autoExposure.readbacks[0].pending = 1u;
    resolveAutoExposureReadback(vkrt, 0u);
    vkrt->renderControl.autoExposure.readbacks[0].mappedSamples = NULL;
    *ioFilteredLuminance = vkrt->renderControl.autoExposure.filteredLuminance;
    *ioExposure = vkrt->sceneSettings.exposure;
}
/* one updateAutoSPP: smoothed per-sample cost in / out, target and measured frame time in ms, current spp -> next spp */
REFHOST_API uint32_t refcontrol_autospp_step(void* h, float* ioControlMsPerSpp, float targetFrameMs, float measuredFrameMs, uint32_t samplesPerPixel) {
    VKRT* vkrt = (VKRT*)h;
    vkrt->sceneSettings.autoSPPEnabled = 1u;
    vkrt->sceneSettings.samplesPerPixel = samplesPerPixel;
    vkrt->renderControl.autoSPP.targetFrameMs = targetFrameMs;
    vkrt->renderControl.autoSPP.controlMs = *ioControlMsPerSpp;
    vkrt->renderStatus.renderTimeMs = measuredFrameMs;
    updateAutoSPP(vkrt);
    *ioControlMsPerSpp = vkrt->renderControl.autoSPP.controlMs;
    return vkrt->sceneSettings.samplesPerPixel;
}
