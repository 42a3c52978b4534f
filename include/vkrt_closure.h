/*
 * vkrt_closure.h — per-call closure evaluation records (TEST ENTRY of the C ABI, no reference equivalent).
 *
 * vkrt_cuda_eval_closures (vkrt_cuda.h) runs the SAME device functions k_shade calls — BSDFState construction, evalBSDF /
 * evalSingleWavelengthBSDF / evalSpectralBSDF and sampleBSDF / sampleSpectralBSDF (reference: src/shaders/bsdf/principled/
 * eval_rgb.slang:104-140, eval_spectral.slang:96-115, bsdf/sample_rgb.slang:6, bsdf/sample_spectral.slang:6) — on an array of
 * independent queries, so that every lobe can be compared call by call with the CPU oracle (oracle_eval_closures) and with the
 * reference's own Slang sources compiled for the CPU (refshade_eval_closures, oracle/ref_slang/). Directions are in the local
 * shading frame (z = shading normal); the sampler runs with the identity shading basis, so sampleWi is local as well.
 */
#ifndef VKRT_CLOSURE_H
#define VKRT_CLOSURE_H

#include "vkrt_shared.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct VKRT_ALIGN16 vkrt_closure_query {
    Material material;      /* sanitised material (as VKRT_setMaterial would store it) */
    float wo[3];            /* outgoing direction, local frame, unit length */
    uint32_t frontFace;     /* 1 = the ray arrived on the geometric front side */
    float wi[3];            /* incident direction for the eval part */
    uint32_t rng;           /* RNG state handed to the sampler */
    float wavelengths[4];   /* nm; single-wavelength mode uses [0] */
    uint32_t mode;          /* 0 = RGB, 1 = spectral single, 2 = spectral hero (4 wavelengths) */
    uint32_t reserved[3];
} vkrt_closure_query;

typedef struct VKRT_ALIGN16 vkrt_closure_result {
    float evalValue[4];     /* RGB: f.rgb | single: f in [0] | hero: f at the 4 wavelengths */
    float evalPdf[4];       /* RGB / single: pdf in [0] | hero: technique pdf per wavelength */
    float sampleWi[3];
    uint32_t sampleFlags;   /* bit 0 = isUsable(), bit 1 = isTransmission */
    float sampleWeight[4];  /* f |cos| / pdf */
    float samplePdf[4];     /* RGB / single: pdf in [0] | hero: technique pdf per wavelength */
    uint32_t rngAfter;      /* RNG state after sampling: pins the number of random numbers consumed */
    uint32_t reserved[3];
} vkrt_closure_result;

#ifdef __cplusplus
}
#endif
#endif /* VKRT_CLOSURE_H */
