/*
 * vkrt_shared.h — wire-format PODs shared by the C host, the C-ABI (vkrt_cuda.h),
 * the CUDA kernels and the CPU oracle.
 *
 * These are byte-for-byte layout restatements of the reference's C/Slang-shared structs
 * (reference: src/shared/types.h:25-150, src/shared/constants.h:4-76, src/shared/formats.h:6-14).
 * They are the data contract between host scene preparation and the device path, so the
 * field order, sizes and offsets are pinned by the static asserts at the bottom of this file
 * (sizes measured on the reference headers with gcc: Vertex 80, ShaderVertex 48, MeshInfo 80,
 * Material 272, EmissiveMesh 32, EmissiveTriangle 48, SceneData 240).
 *
 * Plain C99/C++11/CUDA; no dependency on cglm. Matrices are column-major float[16]
 * (m[col*4+row]), exactly the memory image of the reference's cglm mat4.
 */
#ifndef VKRT_B200_SHARED_H
#define VKRT_B200_SHARED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (reference: src/shared/constants.h) ---------------------------------- */
#define VKRT_MAX_ABSORPTION_COEFFICIENT 1000000.0f

enum { VKRT_TONE_MAPPING_MODE_NONE = 0u, VKRT_TONE_MAPPING_MODE_ACES = 1u, VKRT_TONE_MAPPING_MODE_COUNT = 2u };
enum { VKRT_RENDER_MODE_RGB = 0u, VKRT_RENDER_MODE_SPECTRAL = 1u, VKRT_RENDER_MODE_COUNT = 2u };
enum {
    VKRT_SPECTRAL_SAMPLING_MODE_SINGLE = 0u,
    VKRT_SPECTRAL_SAMPLING_MODE_HERO = 1u,
    VKRT_SPECTRAL_SAMPLING_MODE_COUNT = 2u
};

/* packedRenderSettings = toneMap[15:0] | renderMode[23:16] | spectralSamplingMode[31:24] */
#define VKRT_PACK_RENDER_SETTINGS(tone, mode, spectral) \
    (((uint32_t)(tone) & 0xFFFFu) | (((uint32_t)(mode) & 0xFFu) << 16) | (((uint32_t)(spectral) & 0xFFu) << 24))
#define VKRT_RENDER_SETTINGS_TONE(v) ((uint32_t)(v) & 0xFFFFu)
#define VKRT_RENDER_SETTINGS_MODE(v) (((uint32_t)(v) >> 16) & 0xFFu)
#define VKRT_RENDER_SETTINGS_SPECTRAL(v) (((uint32_t)(v) >> 24) & 0xFFu)

enum {
    VKRT_DEBUG_MODE_NONE = 0u,
    VKRT_DEBUG_MODE_NORMALS = 1u,
    VKRT_DEBUG_MODE_DEPTH = 2u,
    VKRT_DEBUG_MODE_BOUNCE_COUNT = 3u,
    VKRT_DEBUG_MODE_NEE_ONLY = 4u,
    VKRT_DEBUG_MODE_BSDF_ONLY = 5u,
    VKRT_DEBUG_MODE_SELECTION_MASK = 6u,
    VKRT_DEBUG_MODE_BASE_COLOR_MAP = 7u,
    VKRT_DEBUG_MODE_METALLIC_MAP = 8u,
    VKRT_DEBUG_MODE_ROUGHNESS_MAP = 9u,
    VKRT_DEBUG_MODE_NORMAL_MAP = 10u,
    VKRT_DEBUG_MODE_EMISSIVE_MAP = 11u,
    VKRT_DEBUG_MODE_DENOISER_ALBEDO = 12u,
    VKRT_DEBUG_MODE_DENOISER_NORMAL = 13u,
    VKRT_DEBUG_MODE_DENOISER_FEATURE_VALIDITY = 14u,
    VKRT_DEBUG_MODE_DENOISER_FEATURE_DEPTH = 15u,
    VKRT_DEBUG_MODE_DENOISER_FOLLOW_SPECULAR = 16u,
    VKRT_DEBUG_MODE_COUNT = 17u
};

#define VKRT_INVALID_INDEX 0xFFFFFFFFu

enum { VKRT_MATERIAL_ALPHA_MODE_OPAQUE = 0u, VKRT_MATERIAL_ALPHA_MODE_MASK = 1u, VKRT_MATERIAL_ALPHA_MODE_BLEND = 2u };
enum { VKRT_TEXTURE_COLOR_SPACE_SRGB = 0u, VKRT_TEXTURE_COLOR_SPACE_LINEAR = 1u };
#define VKRT_MAX_BINDLESS_TEXTURES 1024u
enum {
    VKRT_MATERIAL_TEXTURE_SLOT_BASE_COLOR = 0u,
    VKRT_MATERIAL_TEXTURE_SLOT_METALLIC_ROUGHNESS = 1u,
    VKRT_MATERIAL_TEXTURE_SLOT_NORMAL = 2u,
    VKRT_MATERIAL_TEXTURE_SLOT_EMISSIVE = 3u,
    VKRT_MATERIAL_TEXTURE_SLOT_COUNT = 4u
};
/* glTF sampler wrap enums; packed as u | v << 16 in Material.*TextureWrap */
#define VKRT_TEXTURE_WRAP_REPEAT 10497u
#define VKRT_TEXTURE_WRAP_CLAMP_TO_EDGE 33071u
#define VKRT_TEXTURE_WRAP_MIRRORED_REPEAT 33648u
#define VKRT_TEXTURE_WRAP_DEFAULT (VKRT_TEXTURE_WRAP_REPEAT | (VKRT_TEXTURE_WRAP_REPEAT << 16u))

/* texture upload formats (reference: src/shared/formats.h:6-14) */
enum {
    VKRT_TEXTURE_FORMAT_RGBA8_UNORM = 0u,
    VKRT_TEXTURE_FORMAT_RGBA16_UNORM = 1u,
    VKRT_TEXTURE_FORMAT_RGBA16_SFLOAT = 2u,
    VKRT_TEXTURE_FORMAT_RGBA32_SFLOAT = 3u,
    VKRT_TEXTURE_FORMAT_COUNT = 4u
};

/* ---- structs (reference: src/shared/types.h) ------------------------------------------ */

/* The reference's float4 is cglm's 16-byte-aligned vec4, so every struct holding one is 16-byte aligned
 * (this is what pads ShaderVertex from 44 to 48 bytes). */
#if defined(_MSC_VER)
#define VKRT_ALIGN16 __declspec(align(16))
#else
#define VKRT_ALIGN16 __attribute__((aligned(16)))
#endif

/* Host-side fat vertex (types.h:25-32). */
typedef struct VKRT_ALIGN16 Vertex {
    float position[4];
    float normal[4];
    float tangent[4]; /* w = handedness */
    float color[4];
    float texcoord0[2];
    float texcoord1[2];
} Vertex;

/* Device vertex, also the BLAS position stream, stride 48 (types.h:34-41). */
typedef struct VKRT_ALIGN16 ShaderVertex {
    float position[4]; /* w unused */
    float texcoord0[2];
    float texcoord1[2];
    uint32_t packedNormal;  /* oct, snorm16 x | snorm16 y << 16 */
    uint32_t packedTangent; /* oct, snorm15 x | snorm15 y << 15 | sign << 31 */
    uint32_t packedColor;   /* RGBA8 */
} ShaderVertex;

/* One per mesh == one per TLAS instance (types.h:43-58). */
typedef struct MeshInfo {
    float position[3];
    uint32_t vertexBase;
    float rotation[3]; /* Euler degrees, applied Rz*Ry*Rx */
    uint32_t vertexCount;
    float scale[3];
    uint32_t indexBase;
    uint32_t indexCount;
    uint32_t materialIndex;
    uint32_t renderBackfaces;
    float lightPdfArea;
    float opacity;
    uint32_t reserved0;
    uint32_t reserved1;
    uint32_t reserved2;
} MeshInfo;

/* types.h:60-101 */
typedef struct VKRT_ALIGN16 Material {
    float baseColor[3];
    float roughness;
    float emissionColor[3];
    float emissionLuminance;
    float eta[3];
    float metallic;
    float k[3];
    float anisotropic;
    float specular;
    float specularTint;
    float abbeNumber;
    float reserved0;
    float sheenTintWeight[4];
    float clearcoat;
    float clearcoatGloss;
    float ior;
    float diffuseRoughness;
    float transmission;
    float subsurface;
    float sheenRoughness;
    float absorptionCoefficient;
    float attenuationColor[3];
    float normalTextureScale;
    uint32_t baseColorTextureIndex;
    uint32_t metallicRoughnessTextureIndex;
    uint32_t normalTextureIndex;
    uint32_t emissiveTextureIndex;
    uint32_t baseColorTextureWrap;
    uint32_t metallicRoughnessTextureWrap;
    uint32_t normalTextureWrap;
    uint32_t emissiveTextureWrap;
    float opacity;
    float alphaCutoff;
    uint32_t alphaMode;
    uint32_t textureTexcoordSets; /* 8 bits per slot */
    float baseColorTextureTransform[4]; /* scale.xy, offset.xy */
    float metallicRoughnessTextureTransform[4];
    float normalTextureTransform[4];
    float emissiveTextureTransform[4];
    float textureRotations[4];
} Material;

/* types.h:103-116 */
typedef struct EmissiveMesh {
    uint32_t triOffset;
    uint32_t triCount;
    float pmfMesh;
    float invTotalArea;
    float emission[3];
    float reserved0;
} EmissiveMesh;

typedef struct VKRT_ALIGN16 EmissiveTriangle {
    float v0Area[4]; /* world v0.xyz, area */
    float e1Pad[4];
    float e2Pad[4];
} EmissiveTriangle;

typedef struct RGB2SpecTableInfo {
    uint32_t res;
    uint32_t scaleOffset;
    uint32_t dataOffset;
} RGB2SpecTableInfo;

/* The per-frame uniform block (types.h:128-150). */
typedef struct VKRT_ALIGN16 SceneData {
    float viewInverse[16]; /* column-major */
    float projInverse[16];
    uint32_t frameNumber;
    uint32_t samplesPerPixel;
    uint32_t rrMaxDepth;
    uint32_t rrMinDepth;
    uint32_t viewportRect[4]; /* x, y, w, h */
    uint32_t packedRenderSettings;
    float exposure;
    float timeBase;
    float timeStep;
    float environmentLight[4]; /* rgb*strength, strength */
    uint32_t environmentTextureIndex;
    float environmentRotation; /* degrees */
    uint32_t debugMode;
    uint32_t misNeeEnabled;
    uint32_t emissiveMeshCount;
    uint32_t emissiveTriangleCount;
    uint32_t selectionEnabled;
    uint32_t selectedMeshIndex;
    RGB2SpecTableInfo rgb2specSRGB;
} SceneData;

#ifdef __cplusplus
}
#define VKRT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define VKRT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

VKRT_STATIC_ASSERT(sizeof(Vertex) == 80, "Vertex layout");
VKRT_STATIC_ASSERT(sizeof(ShaderVertex) == 48, "ShaderVertex layout");
VKRT_STATIC_ASSERT(sizeof(MeshInfo) == 80, "MeshInfo layout");
VKRT_STATIC_ASSERT(sizeof(Material) == 272, "Material layout");
VKRT_STATIC_ASSERT(offsetof(Material, sheenTintWeight) == 80, "Material.sheenTintWeight");
VKRT_STATIC_ASSERT(offsetof(Material, baseColorTextureIndex) == 144, "Material.baseColorTextureIndex");
VKRT_STATIC_ASSERT(offsetof(Material, baseColorTextureTransform) == 192, "Material.baseColorTextureTransform");
VKRT_STATIC_ASSERT(sizeof(EmissiveMesh) == 32, "EmissiveMesh layout");
VKRT_STATIC_ASSERT(sizeof(EmissiveTriangle) == 48, "EmissiveTriangle layout");
VKRT_STATIC_ASSERT(sizeof(SceneData) == 240, "SceneData layout");
VKRT_STATIC_ASSERT(offsetof(SceneData, frameNumber) == 128, "SceneData.frameNumber");
VKRT_STATIC_ASSERT(offsetof(SceneData, rgb2specSRGB) == 224, "SceneData.rgb2specSRGB");

#endif /* VKRT_B200_SHARED_H */
