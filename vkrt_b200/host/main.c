/* main.c — headless command-line driver: the offline half of the reference's app (src/app/main.c:15-83, src/app/cli/cli.c:144-392,
 * src/app/render/benchmark.c:13-293). Window, editor and interactive flags (--width, --fullscreen, --no-ser, --render with
 * presentation) do not exist on a display-less B200; everything else keeps its name and meaning. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vkrt_host.h"

static void printHelp(void) {
    printf("\nB200-native path tracer (vkrt's hot path on CUDA sm_100a).\n");
    printf("\nUsage: vkrt [options]\n\nOptions:\n");
    printf("  --help                    Show this help message and exit\n");
    printf("  --version                 Show version and exit\n");
    printf("  --device-index <index>    CUDA device ordinal\n");
    printf("  --empty-scene             Skip loading the default starter scene\n");
    printf("  --scene <path>            Load a vkrt scene (.json) on startup\n");
    printf("  --import <path>           Import a mesh (.glb) on startup\n");
    printf("  --soup <triangles>        Procedural triangle-soup scene (benchmark config C3)\n");
    printf("  --instanced <n> <glb>     n instances of a model on a jittered grid (benchmark config C4)\n");
    printf("  --render-headless         Run an offline render offscreen (implied; there is no window)\n");
    printf("  --benchmark               Alias for --render-headless\n");
    printf("  --render-width <px>       Offline render width (default: 3840)\n");
    printf("  --render-height <px>      Offline render height (default: 2160)\n");
    printf("  --render-samples <n>      Offline render target samples (default: 16384)\n");
    printf("  --render-spp <n>          Samples per pixel per frame (default: as many as fit the wavefront pool, at most 64: the reference\n"
           "                            auto-tunes this by wall clock, an offline render here fills the GPU instead)\n");
    printf("  --builder <best|lbvh|ploc> BVH hierarchy: lower surface-area cost of radix tree and PLOC (default), radix tree only, PLOC only\n");
    printf("  --render-output <path>    Save the image after completion (.exr linear, .png tone-mapped 16-bit)\n");
    printf("  --spectral <0|1|2>        Override render mode: 0 RGB, 1 spectral single wavelength, 2 spectral hero\n");
    printf("  --rgb2spec <path>         rgb2spec coefficient table (default: assets/rgb2spec/srgb.coeff)\n");
    printf("  --env-importance          Extension: next-event estimation also samples the environment texture (not in the reference)\n");
    printf("\nFile Formats:\n  Import: binary glTF (.glb; geometry + scalar material factors)\n  Export: Image (.png, .exr)\n");
    printf("\nRequirements:\n  An NVIDIA B200 (sm_100a) and libvkrt_cuda.so\n");
}

int main(int argc, char** argv) {
    const char* scene = NULL; const char* import = NULL; const char* output = NULL; const char* instancedGlb = NULL; const char* rgb2spec = "assets/rgb2spec/srgb.coeff";
    uint32_t width = 3840, height = 2160, samples = 16384, spp = 0, soup = 0, instanced = 0, builderFlags = 0;
    int device = -1, emptyScene = 0, spectral = -1, envImportance = 0;
    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
#define NEXT() (i + 1 < argc ? argv[++i] : (fprintf(stderr, "missing value for %s\n", a), exit(2), ""))
        if (!strcmp(a, "--help")) { printHelp(); return 0; }
        else if (!strcmp(a, "--version")) { printf("%s\n", vkrt_cuda_version()); return 0; }
        else if (!strcmp(a, "--device-index")) device = atoi(NEXT());
        else if (!strcmp(a, "--empty-scene")) emptyScene = 1;
        else if (!strcmp(a, "--scene")) scene = NEXT();
        else if (!strcmp(a, "--import")) import = NEXT();
        else if (!strcmp(a, "--soup")) soup = (uint32_t)strtoul(NEXT(), NULL, 10);
        else if (!strcmp(a, "--instanced")) { instanced = (uint32_t)strtoul(NEXT(), NULL, 10); instancedGlb = NEXT(); }
        else if (!strcmp(a, "--render-headless") || !strcmp(a, "--benchmark") || !strcmp(a, "--render")) {}
        else if (!strcmp(a, "--render-width")) width = (uint32_t)strtoul(NEXT(), NULL, 10);
        else if (!strcmp(a, "--render-height")) height = (uint32_t)strtoul(NEXT(), NULL, 10);
        else if (!strcmp(a, "--render-samples")) samples = (uint32_t)strtoul(NEXT(), NULL, 10);
        else if (!strcmp(a, "--render-spp")) spp = (uint32_t)strtoul(NEXT(), NULL, 10);
        else if (!strcmp(a, "--render-output")) output = NEXT();
        else if (!strcmp(a, "--spectral")) spectral = atoi(NEXT());
        else if (!strcmp(a, "--rgb2spec")) rgb2spec = NEXT();
        else if (!strcmp(a, "--env-importance")) envImportance = 1;
        else if (!strcmp(a, "--builder")) {
            const char* b = NEXT();
            builderFlags = !strcmp(b, "lbvh") ? VKRT_CUDA_FLAG_LBVH : (!strcmp(b, "ploc") ? VKRT_CUDA_FLAG_PLOC : 0u);
        }
        else { fprintf(stderr, "unknown option %s (see --help)\n", a); return 2; }
    }
    VKRT* vkrt = NULL;
    if (VKRT_create(&vkrt) != VKRT_SUCCESS) return 1;
    VKRT_CreateInfo ci;
    VKRT_defaultCreateInfo(&ci);
    ci.width = width; ci.height = height; ci.preferredDeviceIndex = device;
    if (envImportance) ci.cudaFlags |= VKRT_CUDA_FLAG_ENV_IMPORTANCE;
    ci.cudaFlags |= builderFlags;
    {   /* an offline render fills the GPU: one frame = as many samples per pixel as a 64 Mi-path wavefront pool holds (~28 GB of the 180 GB
           of HBM; a 512 x 512 frame at 8 spp is launch-bound), at most 64. The reference auto-tunes spp per frame by wall clock instead. */
        const uint64_t pixels = (width && height) ? (uint64_t)width * height : 1;
        if (spp == 0) {
            spp = (uint32_t)((64ull << 20) / pixels);
            if (spp < 1) spp = 1;
            if (spp > 64) spp = 64;
            if (spp > samples) spp = samples ? samples : 1;
        }
        uint64_t pool = pixels * spp + 65536;
        if (pool > (64ull << 20) + 65536) pool = (64ull << 20) + 65536;
        if (pool < (16ull << 20)) pool = 16ull << 20;
        ci.maxPathsInFlight = (uint32_t)pool;
    }
    if (VKRT_initWithCreateInfo(vkrt, &ci) != VKRT_SUCCESS) { fprintf(stderr, "init failed: %s\n", VKRT_lastError(vkrt)); VKRT_destroy(vkrt); return 1; }
    VKRT_Result r = VKRT_SUCCESS;
    if (soup) r = VKRT_appGenerateSoup(vkrt, soup, 0);
    else if (instanced) r = VKRT_appGenerateInstanced(vkrt, instancedGlb, instanced, 0);
    else if (scene) r = VKRT_appLoadScene(vkrt, scene);
    else if (!emptyScene && !import) r = VKRT_appLoadScene(vkrt, "assets/scenes/cornell.json"); /* controller.c:28 default scene */
    if (r == VKRT_SUCCESS && import) r = VKRT_appImportMesh(vkrt, import, NULL, NULL);
    if (r == VKRT_SUCCESS && spectral >= 0) {
        r = VKRT_setRenderMode(vkrt, spectral ? VKRT_RENDER_MODE_SPECTRAL : VKRT_RENDER_MODE_RGB);
        if (r == VKRT_SUCCESS && spectral) r = VKRT_setSpectralSamplingMode(vkrt, spectral == 2 ? VKRT_SPECTRAL_SAMPLING_MODE_HERO : VKRT_SPECTRAL_SAMPLING_MODE_SINGLE);
    }
    VKRT_SceneSettingsSnapshot st;
    VKRT_getSceneSettings(vkrt, &st);
    if (r == VKRT_SUCCESS && st.renderMode == VKRT_RENDER_MODE_SPECTRAL) r = VKRT_loadRGB2SpecTable(vkrt, rgb2spec);
    if (r != VKRT_SUCCESS) { fprintf(stderr, "scene setup failed: %s\n", VKRT_lastError(vkrt)); VKRT_destroy(vkrt); return 1; }
    VKRT_OfflineRenderResult res;
    r = VKRT_appOfflineRender(vkrt, width, height, samples, spp, &res);
    if (r != VKRT_SUCCESS) { fprintf(stderr, "render failed: %s\n", VKRT_lastError(vkrt)); VKRT_destroy(vkrt); return 1; }
    vkrt_cuda_build_stats bs;
    VKRT_getBuildStats(vkrt, &bs);
    printf("Acceleration structure: %.3f ms, %llu triangles in %u BLAS, %u instances, %llu BVH8 nodes\n", bs.buildMs, (unsigned long long)bs.triangleCount,
           bs.uniqueGeometries, bs.instanceCount, (unsigned long long)bs.bvh8NodeCount);
    {   /* the first build of a process also pays for scratch allocation and module loading: time a rebuild of the same scene */
        vkrt_cuda_build_stats again;
        vkrt_cuda_invalidate_accel(VKRT_cudaContext(vkrt));
        if (vkrt_cuda_build_accel(VKRT_cudaContext(vkrt), &again) == VKRT_SUCCESS)
            printf("Acceleration structure rebuild: %.3f ms (%.1f M triangles/s), %u of %u hierarchies kept PLOC\n", again.buildMs,
                   again.buildMs > 0 ? (double)again.triangleCount / again.buildMs / 1e3 : 0.0, again.plocHierarchies, again.flat ? 1u : again.uniqueGeometries + 1u);
    }
    printf("Offline render complete: %.3f s, %.2f samples/s, %.3f ms/sample, %u spp/frame, actual %llu samples\n", res.seconds, res.samplesPerSecond,
           res.samples ? res.seconds * 1000.0 / (double)res.samples : 0.0, res.samplesPerFrame, (unsigned long long)res.samples);
    printf("  %ux%u: %.1f Mpaths/s, %.1f Mrays/s (%llu extension + %llu shadow rays)\n", width, height, res.mpathsPerSecond,
           res.seconds > 0 ? (double)(res.extensionRays + res.shadowRays) / res.seconds / 1e6 : 0.0, (unsigned long long)res.extensionRays, (unsigned long long)res.shadowRays);
    if (output) {
        r = VKRT_saveRenderImage(vkrt, output);
        if (r != VKRT_SUCCESS) { fprintf(stderr, "save failed: %s\n", VKRT_lastError(vkrt)); VKRT_destroy(vkrt); return 1; }
        printf("Saved %s\n", output);
    }
    VKRT_destroy(vkrt);
    return 0;
}
