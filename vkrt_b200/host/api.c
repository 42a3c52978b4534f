/* api.c — the VKRT_* host API over vkrt_cuda_* (see include/vkrt_host.h for the mapping to the reference's src/core/api). */
#include <time.h>

#include "host_state.h"
#include "image_decode.h"
#include "../csrc/tiles.h"

static double nowSeconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + (double)ts.tv_nsec * 1e-9;
}

static VKRT_Result requireReady(const VKRT* vkrt) {
    if (!vkrt) return VKRT_ERROR_INVALID_ARGUMENT;
    return vkrt->initialized ? VKRT_SUCCESS : VKRT_ERROR_OPERATION_FAILED;
}
static VKRT_Result cudaCheck(VKRT* vkrt, VKRT_Result r, const char* what) {
    if (r != VKRT_SUCCESS) hostFail(vkrt, r, "%s: %s", what, vkrt->cuda ? vkrt_cuda_last_error(vkrt->cuda) : "no device context");
    return r;
}
const char* VKRT_lastError(const VKRT* vkrt) { return vkrt ? vkrt->error : "null handle"; }

/* ---- lifecycle ---------------------------------------------------------------------------------------------------------- */
void VKRT_defaultCreateInfo(VKRT_CreateInfo* ci) {
    if (!ci) return;
    memset(ci, 0, sizeof(*ci));
    ci->width = 1600u;  /* VKRT_DEFAULT_WIDTH / HEIGHT, api/config.h:4-5 */
    ci->height = 900u;
    ci->title = "vkrt";
    ci->headless = 1;
    ci->preferredDeviceIndex = -1;
    ci->worldSize = 1;
}

VKRT_Result VKRT_create(VKRT** outVkrt) {
    if (!outVkrt) return VKRT_ERROR_INVALID_ARGUMENT;
    VKRT* v = (VKRT*)calloc(1, sizeof(VKRT));
    if (!v) return VKRT_ERROR_OUT_OF_MEMORY;
    *outVkrt = v;
    return VKRT_SUCCESS;
}

static void freeMeshes(VKRT* v) {
    for (uint32_t i = 0; i < v->meshCount; i++)
        if (v->meshes[i].ownsGeometry) { free(v->meshes[i].vertices); free(v->meshes[i].indices); }
    free(v->meshes);
    v->meshes = NULL;
    v->meshCount = v->meshCapacity = 0;
}

void VKRT_deinit(VKRT* v) {
    if (!v) return;
    if (v->cuda) vkrt_cuda_destroy(v->cuda);
    v->cuda = NULL;
    freeMeshes(v);
    free(v->materials); v->materials = NULL; v->materialCount = 0;
    for (uint32_t i = 0; i < v->textureCount; i++) free(v->textures[i].pixels);
    free(v->textures); v->textures = NULL; v->textureCount = 0;
    free(v->packedVertices); free(v->packedIndices); free(v->meshInfos); free(v->world3x4); free(v->geometrySource); free(v->alphaTested);
    free(v->materialArray); free(v->emissiveMeshes); free(v->emissiveTriangles); free(v->meshAliasQ); free(v->meshAliasIdx);
    free(v->triAliasQ); free(v->triAliasIdx); free(v->rgb2spec);
    VKRT_CreateInfo ci = v->createInfo;
    memset(v, 0, sizeof(*v));
    v->createInfo = ci;
}
void VKRT_destroy(VKRT* v) {
    if (!v) return;
    VKRT_deinit(v);
    free(v);
}

static VKRT_Result ensureDefaultMaterial(VKRT* v) { /* material 0 is always "Default Material" (internal/state.c:42-54) */
    if (v->materialCount > 0) return VKRT_SUCCESS;
    v->materials = (HostMaterial*)calloc(1, sizeof(HostMaterial));
    if (!v->materials) return VKRT_ERROR_OUT_OF_MEMORY;
    v->materials[0].material = hostSanitizeMaterial(v, VKRT_materialDefault());
    snprintf(v->materials[0].name, sizeof(v->materials[0].name), "Default Material");
    v->materialCount = 1;
    v->materialsDirty = 1;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_initWithCreateInfo(VKRT* v, const VKRT_CreateInfo* createInfo) {
    if (!v || !createInfo) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->initialized) return VKRT_ERROR_OPERATION_FAILED;
    v->createInfo = *createInfo;
    v->hostOnly = createInfo->hostOnly ? 1 : 0;
    if (!v->hostOnly) {
        vkrt_cuda_create_info ci;
        memset(&ci, 0, sizeof(ci));
        ci.device = createInfo->preferredDeviceIndex;
        ci.rank = createInfo->rank;
        ci.worldSize = createInfo->worldSize ? createInfo->worldSize : 1u;
        ci.maxPathsInFlight = createInfo->maxPathsInFlight;
        ci.flags = createInfo->cudaFlags;
        VKRT_Result r = vkrt_cuda_create(&ci, &v->cuda);
        if (r != VKRT_SUCCESS) return hostFail(v, VKRT_ERROR_INITIALIZATION_FAILED, "vkrt_cuda_create failed (%d): a CUDA device is required, there is no CPU renderer", (int)r);
    }
    /* initializeDefaultSceneSettings (scene/uniform.c:93-151) */
    uint32_t w = createInfo->width ? createInfo->width : 1600u, h = createInfo->height ? createInfo->height : 900u;
    VKRT_SceneSettingsSnapshot* s = &v->sceneSettings;
    memset(s, 0, sizeof(*s));
    s->samplesPerPixel = 8; s->rrMaxDepth = 8; s->rrMinDepth = 4;
    s->toneMappingMode = VKRT_TONE_MAPPING_MODE_ACES;
    s->renderMode = VKRT_RENDER_MODE_RGB;
    s->spectralSamplingMode = VKRT_SPECTRAL_SAMPLING_MODE_HERO;
    s->exposure = 1.0f;
    s->environmentColor[0] = s->environmentColor[1] = s->environmentColor[2] = 0.25f;
    s->environmentStrength = 1.0f;
    s->environmentTextureIndex = VKRT_INVALID_INDEX;
    s->autoSPPEnabled = 0; /* the reference defaults to 1 and locks spp after a wall-clock warm-up (benchmark.c:87-102); fixed spp keeps runs reproducible */
    s->autoSPPTargetFPS = 60;
    s->camera.nearZ = 0.001f; s->camera.farZ = 10000.0f; s->camera.vfov = 40.0f;
    s->camera.pos[0] = -0.5f; s->camera.pos[1] = 0.2f; s->camera.pos[2] = -0.2f;
    s->camera.up[2] = 1.0f;
    s->timeBase = -1.0f; s->timeStep = 0.5f;
    s->debugMode = VKRT_DEBUG_MODE_NONE;
    s->misNeeEnabled = 1u;
    s->selectedMeshIndex = VKRT_INVALID_INDEX;
    memset(&v->sceneData, 0, sizeof(v->sceneData));
    v->sceneData.viewportRect[2] = w;
    v->sceneData.viewportRect[3] = h;
    v->renderWidth = w;
    v->renderHeight = h;
    v->renderStatus.renderPhase = VKRT_RENDER_PHASE_INACTIVE;
    v->initialized = 1;
    VKRT_Result r = ensureDefaultMaterial(v);
    if (r != VKRT_SUCCESS) return r;
    hostSyncCameraMatrices(v);
    hostWriteSceneStateUniform(v);
    v->filmDirty = 1;
    v->geometryDirty = v->sceneResourcesDirty = v->materialsDirty = v->lightsDirty = v->accelDirty = 1;
    v->accumulationNeedsReset = 1;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_init(VKRT* v) {
    VKRT_CreateInfo ci;
    VKRT_defaultCreateInfo(&ci);
    return VKRT_initWithCreateInfo(v, &ci);
}

/* ---- geometry ------------------------------------------------------------------------------------------------------------ */
VKRT_Result VKRT_uploadMeshDataBatch(VKRT* v, const VKRT_MeshUpload* uploads, size_t uploadCount) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (!uploads || uploadCount == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    for (size_t i = 0; i < uploadCount; i++) {
        const VKRT_MeshUpload* u = &uploads[i];
        if (!u->vertices || !u->indices || u->vertexCount == 0 || u->indexCount == 0 || u->vertexCount > 0xffffffffu || u->indexCount > 0xffffffffu)
            return VKRT_ERROR_INVALID_ARGUMENT;
    }
    uint32_t firstNew = v->meshCount;
    if ((uint64_t)v->meshCount + uploadCount > v->meshCapacity) {
        uint32_t cap = (uint32_t)(v->meshCount + uploadCount) * 2u + 8u;
        HostMesh* nm = (HostMesh*)realloc(v->meshes, (size_t)cap * sizeof(HostMesh));
        if (!nm) return VKRT_ERROR_OUT_OF_MEMORY;
        v->meshes = nm;
        v->meshCapacity = cap;
    }
    for (size_t i = 0; i < uploadCount; i++) {
        const VKRT_MeshUpload* u = &uploads[i];
        HostMesh* m = &v->meshes[v->meshCount];
        memset(m, 0, sizeof(*m));
        h_mat4_identity(m->worldTransform);
        m->renderBackfacesOverride = -1;
        m->info.vertexCount = (uint32_t)u->vertexCount;
        m->info.indexCount = (uint32_t)u->indexCount;
        m->info.opacity = 1.0f;
        m->info.scale[0] = m->info.scale[1] = m->info.scale[2] = 1.0f;
        m->fingerprint = hostGeometryFingerprint(u->vertices, u->vertexCount, u->indices, u->indexCount);
        /* instancing is decided here: identical geometry shares the first uploader's buffers and BLAS (geometry.c:180-210) */
        uint32_t source = VKRT_INVALID_INDEX;
        for (uint32_t k = 0; k < v->meshCount; k++) {
            const HostMesh* e = &v->meshes[k];
            if (!e->ownsGeometry || e->info.vertexCount != m->info.vertexCount || e->info.indexCount != m->info.indexCount) continue;
            if (e->fingerprint != m->fingerprint) continue;
            if (memcmp(e->vertices, u->vertices, u->vertexCount * sizeof(Vertex)) != 0) continue;
            if (memcmp(e->indices, u->indices, u->indexCount * sizeof(uint32_t)) != 0) continue;
            source = e->geometrySource;
            break;
        }
        if (source != VKRT_INVALID_INDEX) {
            m->ownsGeometry = 0;
            m->geometrySource = source;
            m->vertices = v->meshes[source].vertices;
            m->indices = v->meshes[source].indices;
        } else {
            void* vp = NULL;
            if (posix_memalign(&vp, 16, u->vertexCount * sizeof(Vertex)) != 0) vp = NULL;
            uint32_t* ip = (uint32_t*)malloc(u->indexCount * sizeof(uint32_t));
            if (!vp || !ip) {
                free(vp); free(ip);
                for (uint32_t k = firstNew; k < v->meshCount; k++)
                    if (v->meshes[k].ownsGeometry) { free(v->meshes[k].vertices); free(v->meshes[k].indices); }
                v->meshCount = firstNew;
                return VKRT_ERROR_OUT_OF_MEMORY;
            }
            memcpy(vp, u->vertices, u->vertexCount * sizeof(Vertex));
            memcpy(ip, u->indices, u->indexCount * sizeof(uint32_t));
            m->vertices = (Vertex*)vp;
            m->indices = ip;
            m->ownsGeometry = 1;
            m->geometrySource = v->meshCount;
        }
        snprintf(m->name, sizeof(m->name), "Mesh %u", v->meshCount);
        v->meshCount++;
    }
    v->geometryDirty = v->sceneResourcesDirty = v->lightsDirty = v->accelDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_uploadMeshData(VKRT* v, const Vertex* vertices, size_t vertexCount, const uint32_t* indices, size_t indexCount) {
    VKRT_MeshUpload u = {vertices, vertexCount, indices, indexCount};
    return VKRT_uploadMeshDataBatch(v, &u, 1);
}

VKRT_Result VKRT_removeMesh(VKRT* v, uint32_t meshIndex) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (meshIndex >= v->meshCount) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh removed = v->meshes[meshIndex];
    if (removed.ownsGeometry) {
        /* hand the geometry to the first duplicate, if any (scene/geometry.c removal path) */
        uint32_t heir = VKRT_INVALID_INDEX;
        for (uint32_t k = 0; k < v->meshCount; k++)
            if (k != meshIndex && !v->meshes[k].ownsGeometry && v->meshes[k].geometrySource == meshIndex) { heir = k; break; }
        if (heir != VKRT_INVALID_INDEX) {
            v->meshes[heir].ownsGeometry = 1;
            for (uint32_t k = 0; k < v->meshCount; k++)
                if (v->meshes[k].geometrySource == meshIndex) v->meshes[k].geometrySource = heir;
        } else {
            free(removed.vertices);
            free(removed.indices);
        }
    }
    memmove(&v->meshes[meshIndex], &v->meshes[meshIndex + 1], (size_t)(v->meshCount - meshIndex - 1) * sizeof(HostMesh));
    v->meshCount--;
    for (uint32_t k = 0; k < v->meshCount; k++)
        if (v->meshes[k].geometrySource > meshIndex) v->meshes[k].geometrySource--;
    v->geometryDirty = v->sceneResourcesDirty = v->lightsDirty = v->accelDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}

/* ---- settings ------------------------------------------------------------------------------------------------------------- */
VKRT_Result VKRT_invalidateAccumulation(VKRT* v) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setSamplesPerPixel(VKRT* v, uint32_t spp) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (spp == 0) spp = 1;
    if (v->sceneSettings.samplesPerPixel == spp) return VKRT_SUCCESS;
    v->sceneSettings.samplesPerPixel = spp;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setPathDepth(VKRT* v, uint32_t rrMin, uint32_t rrMax) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (rrMax < 1u) rrMax = 1u;
    if (rrMax > 64u) rrMax = 64u;
    if (rrMin > rrMax) rrMin = rrMax;
    if (v->sceneSettings.rrMinDepth == rrMin && v->sceneSettings.rrMaxDepth == rrMax) return VKRT_SUCCESS;
    v->sceneSettings.rrMinDepth = rrMin;
    v->sceneSettings.rrMaxDepth = rrMax;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setAutoSPPEnabled(VKRT* v, uint8_t enabled) { /* api/settings.c:59-64 */
    if (!v) return VKRT_ERROR_INVALID_ARGUMENT;
    v->sceneSettings.autoSPPEnabled = enabled ? 1 : 0;
    hostResetAutoSPPState(v, 0);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setAutoSPPTargetFPS(VKRT* v, uint32_t targetFPS) { /* api/settings.c:66-79; no display: refresh rate 60 */
    if (!v) return VKRT_ERROR_INVALID_ARGUMENT;
    if (targetFPS == 0) targetFPS = 60u;
    if (targetFPS < 30u) targetFPS = 30u;
    if (targetFPS > 360u) targetFPS = 360u;
    v->sceneSettings.autoSPPTargetFPS = targetFPS;
    v->autoSPPTargetFrameMs = 1000.0f / (float)targetFPS;
    hostResetAutoSPPState(v, 0);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setAutoExposureEnabled(VKRT* v, uint8_t enabled) { /* api/settings.c:135-149 */
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    enabled = enabled ? 1u : 0u;
    if (v->sceneSettings.autoExposureEnabled == enabled) return VKRT_SUCCESS;
    v->sceneSettings.autoExposureEnabled = enabled;
    v->autoExposureFilteredLuminance = 0.0f;
    hostWriteSceneStateUniform(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setToneMappingMode(VKRT* v, VKRT_ToneMappingMode mode) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (mode >= VKRT_TONE_MAPPING_MODE_COUNT) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->sceneSettings.toneMappingMode == mode) return VKRT_SUCCESS;
    v->sceneSettings.toneMappingMode = mode;
    hostWriteSceneStateUniform(v); /* display-only: no accumulation restart (settings.c:95-107) */
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setRenderMode(VKRT* v, VKRT_RenderMode mode) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (mode >= VKRT_RENDER_MODE_COUNT) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->sceneSettings.renderMode == mode) return VKRT_SUCCESS;
    v->sceneSettings.renderMode = mode;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setSpectralSamplingMode(VKRT* v, VKRT_SpectralSamplingMode mode) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (mode >= VKRT_SPECTRAL_SAMPLING_MODE_COUNT) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->sceneSettings.spectralSamplingMode == mode) return VKRT_SUCCESS;
    v->sceneSettings.spectralSamplingMode = mode;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setExposure(VKRT* v, float exposure) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    exposure = hostFiniteClampf(exposure, 1.0f, 0.0f, INFINITY);
    if (v->sceneSettings.exposure == exposure) return VKRT_SUCCESS;
    v->sceneSettings.exposure = exposure;
    hostWriteSceneStateUniform(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setEnvironmentLight(VKRT* v, vkrt_vec3 color, float strength) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (!color) return VKRT_ERROR_INVALID_ARGUMENT;
    float c[3];
    for (int i = 0; i < 3; i++) c[i] = hostFiniteClampf(color[i], 1.0f, 0.0f, INFINITY);
    strength = hostFiniteClampf(strength, 0.0f, 0.0f, INFINITY);
    VKRT_SceneSettingsSnapshot* s = &v->sceneSettings;
    if (s->environmentColor[0] == c[0] && s->environmentColor[1] == c[1] && s->environmentColor[2] == c[2] && s->environmentStrength == strength)
        return VKRT_SUCCESS;
    memcpy(s->environmentColor, c, sizeof(c));
    s->environmentStrength = strength;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setEnvironmentRotation(VKRT* v, float deg) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    deg = fmodf(hostFiniteOrf(deg, 0.0f), 360.0f);
    if (deg < -180.0f) deg += 360.0f;
    if (deg >= 180.0f) deg -= 360.0f;
    if (v->sceneSettings.environmentRotation == deg) return VKRT_SUCCESS;
    v->sceneSettings.environmentRotation = deg;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setDebugMode(VKRT* v, VKRT_DebugMode mode) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (mode >= VKRT_DEBUG_MODE_COUNT) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->sceneSettings.debugMode == mode) return VKRT_SUCCESS;
    v->sceneSettings.debugMode = mode;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMisNeeEnabled(VKRT* v, uint8_t enabled) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    enabled = enabled ? 1u : 0u;
    if (v->sceneSettings.misNeeEnabled == enabled) return VKRT_SUCCESS;
    v->sceneSettings.misNeeEnabled = enabled;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setTimeRange(VKRT* v, float timeBase, float timeStep) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    timeBase = hostFiniteOrf(timeBase, -1.0f);
    timeStep = hostFiniteOrf(timeStep, timeBase);
    if (timeBase < 0.0f) { timeBase = -1.0f; timeStep = -1.0f; }
    else if (timeStep < timeBase) timeStep = timeBase;
    if (v->sceneSettings.timeBase == timeBase && v->sceneSettings.timeStep == timeStep) return VKRT_SUCCESS;
    v->sceneSettings.timeBase = timeBase;
    v->sceneSettings.timeStep = timeStep;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setRenderViewport(VKRT* v, uint32_t x, uint32_t y, uint32_t width, uint32_t height) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (width == 0 || height == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    if (x >= v->renderWidth) x = v->renderWidth - 1u;
    if (y >= v->renderHeight) y = v->renderHeight - 1u;
    if (width > v->renderWidth - x) width = v->renderWidth - x;
    if (height > v->renderHeight - y) height = v->renderHeight - y;
    uint32_t* r = v->sceneData.viewportRect;
    if (r[0] == x && r[1] == y && r[2] == width && r[3] == height) return VKRT_SUCCESS;
    r[0] = x; r[1] = y; r[2] = width; r[3] = height;
    hostSyncCameraMatrices(v);
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_cameraSetPose(VKRT* v, vkrt_vec3 position, vkrt_vec3 target, vkrt_vec3 up, float vfov) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    Camera* c = &v->sceneSettings.camera;
    if (position) memcpy(c->pos, position, sizeof(vkrt_vec3));
    if (target) memcpy(c->target, target, sizeof(vkrt_vec3));
    if (up) memcpy(c->up, up, sizeof(vkrt_vec3));
    if (vfov > 0.0f) c->vfov = vfov;
    hostSyncCameraMatrices(v);
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_cameraGetPose(const VKRT* v, vkrt_vec3 position, vkrt_vec3 target, vkrt_vec3 up, float* vfov) {
    if (!v) return VKRT_ERROR_INVALID_ARGUMENT;
    const Camera* c = &v->sceneSettings.camera;
    if (position) memcpy(position, c->pos, sizeof(vkrt_vec3));
    if (target) memcpy(target, c->target, sizeof(vkrt_vec3));
    if (up) memcpy(up, c->up, sizeof(vkrt_vec3));
    if (vfov) *vfov = c->vfov;
    return VKRT_SUCCESS;
}

/* srgb.coeff: "SPEC", u32 res, float scale[res], float coeff[3][res][res][res][3] (src/core/scene/rgb2spec.c:17-59) */
VKRT_Result VKRT_loadRGB2SpecTable(VKRT* v, const char* path) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (!path) return VKRT_ERROR_INVALID_ARGUMENT;
    FILE* f = fopen(path, "rb");
    if (!f) return hostFail(v, VKRT_ERROR_OPERATION_FAILED, "cannot open rgb2spec table %s", path);
    char magic[4];
    uint32_t res = 0;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "SPEC", 4) != 0 || fread(&res, 4, 1, f) != 1 || res < 2 || res > 256) {
        fclose(f);
        return hostFail(v, VKRT_ERROR_OPERATION_FAILED, "%s is not an rgb2spec coefficient file", path);
    }
    size_t count = (size_t)res + (size_t)9 * res * res * res;
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    if ((size_t)size != 8 + 4 * count) {
        fclose(f);
        return hostFail(v, VKRT_ERROR_OPERATION_FAILED, "%s: size %ld does not match res %u", path, size, res);
    }
    fseek(f, 8, SEEK_SET);
    float* data = (float*)malloc(count * sizeof(float));
    if (!data || fread(data, sizeof(float), count, f) != count) {
        free(data);
        fclose(f);
        return hostFail(v, VKRT_ERROR_OPERATION_FAILED, "%s: short read", path);
    }
    fclose(f);
    free(v->rgb2spec);
    v->rgb2spec = data;
    v->rgb2specFloats = (uint32_t)count;
    v->rgb2specInfo.res = res;
    v->rgb2specInfo.scaleOffset = 0;
    v->rgb2specInfo.dataOffset = res;
    v->rgb2specDirty = 1;
    hostWriteSceneStateUniform(v);
    return VKRT_SUCCESS;
}

/* ---- textures ------------------------------------------------------------------------------------------------------------- */
static size_t texelSize(uint32_t format) {
    switch (format) {
        case VKRT_TEXTURE_FORMAT_RGBA8_UNORM: return 4;
        case VKRT_TEXTURE_FORMAT_RGBA16_UNORM: case VKRT_TEXTURE_FORMAT_RGBA16_SFLOAT: return 8;
        case VKRT_TEXTURE_FORMAT_RGBA32_SFLOAT: return 16;
        default: return 0;
    }
}
VKRT_Result VKRT_addTextureFromPixels(VKRT* v, const VKRT_TextureUpload* up, uint32_t* outIndex) {
    if (outIndex) *outIndex = VKRT_INVALID_INDEX;
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (!up || !up->pixels || up->width == 0 || up->height == 0 || texelSize(up->format) == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->textureCount >= VKRT_MAX_BINDLESS_TEXTURES) return VKRT_ERROR_OPERATION_FAILED;
    if (up->colorSpace == VKRT_TEXTURE_COLOR_SPACE_SRGB && up->format != VKRT_TEXTURE_FORMAT_RGBA8_UNORM) return VKRT_ERROR_INVALID_ARGUMENT;
    HostTexture* nt = (HostTexture*)realloc(v->textures, (size_t)(v->textureCount + 1u) * sizeof(HostTexture));
    if (!nt) return VKRT_ERROR_OUT_OF_MEMORY;
    v->textures = nt;
    HostTexture* t = &v->textures[v->textureCount];
    memset(t, 0, sizeof(*t));
    size_t bytes = (size_t)up->width * up->height * texelSize(up->format);
    t->pixels = malloc(bytes);
    if (!t->pixels) return VKRT_ERROR_OUT_OF_MEMORY;
    memcpy(t->pixels, up->pixels, bytes);
    t->width = up->width; t->height = up->height; t->format = up->format; t->colorSpace = up->colorSpace;
    snprintf(t->name, sizeof(t->name), "%s", up->name && up->name[0] ? up->name : "Texture");
    if (outIndex) *outIndex = v->textureCount;
    v->textureCount++;
    v->texturesDirty = 1;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getTextureCount(const VKRT* v, uint32_t* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->textureCount;
    return VKRT_SUCCESS;
}
/* users of a texture: material slots that name it, plus the environment (scene/textures.c:374-381) */
static uint32_t textureMaterialUsers(const VKRT* v, uint32_t textureIndex) {
    uint32_t users = 0;
    for (uint32_t i = 0; i < v->materialCount; i++) {
        const Material* m = &v->materials[i].material;
        users += (m->baseColorTextureIndex == textureIndex) + (m->metallicRoughnessTextureIndex == textureIndex) + (m->normalTextureIndex == textureIndex) +
                 (m->emissiveTextureIndex == textureIndex);
    }
    return users;
}
static uint32_t textureUsers(const VKRT* v, uint32_t textureIndex) {
    return textureMaterialUsers(v, textureIndex) + (v->sceneSettings.environmentTextureIndex == textureIndex ? 1u : 0u);
}
VKRT_Result VKRT_getTextureSnapshot(const VKRT* v, uint32_t textureIndex, VKRT_TextureSnapshot* out) {
    if (!v || !out || textureIndex >= v->textureCount) return VKRT_ERROR_INVALID_ARGUMENT;
    const HostTexture* t = &v->textures[textureIndex];
    memset(out, 0, sizeof(*out));
    out->width = t->width; out->height = t->height; out->format = t->format; out->colorSpace = t->colorSpace;
    out->useCount = textureMaterialUsers(v, textureIndex);
    memcpy(out->name, t->name, VKRT_NAME_LEN);
    return VKRT_SUCCESS;
}
/* scene/textures.c:438-482: all or nothing */
VKRT_Result VKRT_addTexturesBatch(VKRT* v, const VKRT_TextureUpload* uploads, size_t uploadCount, uint32_t* outIndices) {
    if (!v || !uploads || uploadCount == 0u) return VKRT_ERROR_INVALID_ARGUMENT;
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    for (size_t i = 0; i < uploadCount; i++) {
        const VKRT_TextureUpload* up = &uploads[i];
        if (!up->pixels || up->width == 0 || up->height == 0 || texelSize(up->format) == 0) return VKRT_ERROR_INVALID_ARGUMENT;
        if (up->colorSpace == VKRT_TEXTURE_COLOR_SPACE_SRGB && up->format != VKRT_TEXTURE_FORMAT_RGBA8_UNORM) return VKRT_ERROR_INVALID_ARGUMENT;
    }
    if ((size_t)v->textureCount + uploadCount > VKRT_MAX_BINDLESS_TEXTURES) return VKRT_ERROR_OPERATION_FAILED;
    const uint32_t before = v->textureCount;
    for (size_t i = 0; i < uploadCount; i++) {
        VKRT_Result r = VKRT_addTextureFromPixels(v, &uploads[i], outIndices ? &outIndices[i] : NULL);
        if (r != VKRT_SUCCESS) {
            while (v->textureCount > before) { free(v->textures[v->textureCount - 1u].pixels); v->textureCount--; }
            return r;
        }
    }
    return VKRT_SUCCESS;
}
/* scene/textures.c:484-516 + utility/image.c: decode a PNG / JPEG / EXR file and add it */
VKRT_Result VKRT_addTextureFromFile(VKRT* v, const char* path, const char* name, uint32_t colorSpace, uint32_t* outIndex) {
    if (outIndex) *outIndex = VKRT_INVALID_INDEX;
    if (!v || !path || !path[0] || (colorSpace != VKRT_TEXTURE_COLOR_SPACE_SRGB && colorSpace != VKRT_TEXTURE_COLOR_SPACE_LINEAR)) return VKRT_ERROR_INVALID_ARGUMENT;
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    FILE* probe = fopen(path, "rb");
    if (!probe) return VKRT_ERROR_INVALID_ARGUMENT;   /* resolveExistingPath failure */
    fclose(probe);
    HostImage image;
    char why[256] = "";
    if (!hostLoadImageFile(path, colorSpace, &image, why, sizeof(why))) return hostFail(v, VKRT_ERROR_OPERATION_FAILED, "%s", why);
    const char* base = strrchr(path, '/');
    VKRT_TextureUpload up = {name && name[0] ? name : (base ? base + 1 : path), image.pixels, image.width, image.height, image.format, image.colorSpace};
    VKRT_Result r = VKRT_addTextureFromPixels(v, &up, outIndex);
    hostFreeImage(&image);
    return r;
}
/* scene/textures.c:518-578: refuses while a material still uses it; later indices shift down by one */
VKRT_Result VKRT_removeTexture(VKRT* v, uint32_t textureIndex) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (textureIndex >= v->textureCount) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->sceneSettings.environmentTextureIndex == textureIndex) v->sceneSettings.environmentTextureIndex = VKRT_INVALID_INDEX;
    if (textureMaterialUsers(v, textureIndex) != 0u) return VKRT_ERROR_OPERATION_FAILED;
    free(v->textures[textureIndex].pixels);
    memmove(&v->textures[textureIndex], &v->textures[textureIndex + 1u], (size_t)(v->textureCount - textureIndex - 1u) * sizeof(HostTexture));
    v->textureCount--;
    for (uint32_t i = 0; i < v->materialCount; i++) {
        Material* m = &v->materials[i].material;
        uint32_t* idx[4] = {&m->baseColorTextureIndex, &m->metallicRoughnessTextureIndex, &m->normalTextureIndex, &m->emissiveTextureIndex};
        for (int k = 0; k < 4; k++)
            if (*idx[k] != VKRT_INVALID_INDEX && *idx[k] > textureIndex) (*idx[k])--;
    }
    uint32_t env = v->sceneSettings.environmentTextureIndex;   /* scene/environment.c:69-78 */
    if (env != VKRT_INVALID_INDEX && env > textureIndex) v->sceneSettings.environmentTextureIndex = env - 1u;
    v->texturesDirty = v->materialsDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
/* scene/environment.c:9-37: the environment must be LINEAR; the texture it replaces is dropped when nothing else uses it */
static VKRT_Result replaceEnvironmentTexture(VKRT* v, uint32_t next) {
    uint32_t previous = v->sceneSettings.environmentTextureIndex;
    if (next != VKRT_INVALID_INDEX && (next >= v->textureCount || v->textures[next].colorSpace != VKRT_TEXTURE_COLOR_SPACE_LINEAR)) return VKRT_ERROR_INVALID_ARGUMENT;
    if (previous == next) return VKRT_SUCCESS;
    v->sceneSettings.environmentTextureIndex = VKRT_INVALID_INDEX;
    if (previous != VKRT_INVALID_INDEX && previous < v->textureCount && textureUsers(v, previous) == 0u) {
        VKRT_Result r = VKRT_removeTexture(v, previous);
        if (r != VKRT_SUCCESS) { v->sceneSettings.environmentTextureIndex = previous; return r; }
        if (next != VKRT_INVALID_INDEX && previous < next) next--;
    }
    v->sceneSettings.environmentTextureIndex = next;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setEnvironmentTextureFromPixels(VKRT* v, const VKRT_TextureUpload* up) {
    uint32_t idx = VKRT_INVALID_INDEX;
    VKRT_Result r = VKRT_addTextureFromPixels(v, up, &idx);
    if (r != VKRT_SUCCESS) return r;
    r = replaceEnvironmentTexture(v, idx);
    if (r != VKRT_SUCCESS && idx < v->textureCount && textureUsers(v, idx) == 0u) VKRT_removeTexture(v, idx);
    return r;
}
VKRT_Result VKRT_setEnvironmentTextureFromFile(VKRT* v, const char* path) {
    if (!v || !path || !path[0]) return VKRT_ERROR_INVALID_ARGUMENT;
    uint32_t idx = VKRT_INVALID_INDEX;
    VKRT_Result r = VKRT_addTextureFromFile(v, path, NULL, VKRT_TEXTURE_COLOR_SPACE_LINEAR, &idx);
    if (r != VKRT_SUCCESS) return r;
    r = replaceEnvironmentTexture(v, idx);
    if (r != VKRT_SUCCESS && idx < v->textureCount && textureUsers(v, idx) == 0u) VKRT_removeTexture(v, idx);
    return r;
}
VKRT_Result VKRT_clearEnvironmentTexture(VKRT* v) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    return replaceEnvironmentTexture(v, VKRT_INVALID_INDEX);
}

/* ---- materials ------------------------------------------------------------------------------------------------------------ */
VKRT_Result VKRT_addMaterial(VKRT* v, const Material* material, const char* name, uint32_t* outIndex) {
    if (outIndex) *outIndex = VKRT_INVALID_INDEX;
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    VKRT_Result d = ensureDefaultMaterial(v);
    if (d != VKRT_SUCCESS) return d;
    uint32_t idx = v->materialCount;
    HostMaterial* nm = (HostMaterial*)realloc(v->materials, (size_t)(idx + 1u) * sizeof(HostMaterial));
    if (!nm) return VKRT_ERROR_OUT_OF_MEMORY;
    v->materials = nm;
    memset(&v->materials[idx], 0, sizeof(HostMaterial));
    v->materials[idx].material = hostSanitizeMaterial(v, material ? *material : VKRT_materialDefault());
    if (name && name[0]) snprintf(v->materials[idx].name, VKRT_NAME_LEN, "%s", name);
    else snprintf(v->materials[idx].name, VKRT_NAME_LEN, "Material %u", idx);
    v->materialCount = idx + 1u;
    v->materialsDirty = 1;
    hostResetSceneData(v);
    if (outIndex) *outIndex = idx;
    return VKRT_SUCCESS;
}
/* api/mesh.c:296-342,449-459: the default material (index 0) cannot be removed; meshes that used the removed material fall back to it
 * and lose their assignment, later material indices shift down by one, and textures only the removed material referenced are dropped. */
VKRT_Result VKRT_removeMaterial(VKRT* v, uint32_t materialIndex) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (materialIndex >= v->materialCount) return VKRT_ERROR_INVALID_ARGUMENT;
    if (materialIndex == 0u) return VKRT_ERROR_OPERATION_FAILED;
    const Material removed = v->materials[materialIndex].material;
    for (uint32_t i = 0; i < v->meshCount; i++) {
        if (v->meshes[i].info.materialIndex == materialIndex) {
            v->meshes[i].info.materialIndex = 0u;
            v->meshes[i].hasMaterialAssignment = 0;
        }
    }
    memmove(&v->materials[materialIndex], &v->materials[materialIndex + 1u], (size_t)(v->materialCount - materialIndex - 1u) * sizeof(HostMaterial));
    v->materialCount--;
    for (uint32_t i = 0; i < v->meshCount; i++)
        if (v->meshes[i].info.materialIndex > materialIndex) v->meshes[i].info.materialIndex--;
    v->materialsDirty = v->sceneResourcesDirty = v->lightsDirty = v->accelDirty = 1;
    hostResetSceneData(v);
    /* releaseTexturesReferencedByMaterialIfUnused: distinct indices, highest first, so that a removal does not shift the ones still to do */
    uint32_t tex[4] = {removed.baseColorTextureIndex, removed.metallicRoughnessTextureIndex, removed.normalTextureIndex, removed.emissiveTextureIndex};
    for (int a = 0; a < 4; a++)
        for (int b = a + 1; b < 4; b++)
            if (tex[b] != VKRT_INVALID_INDEX && (tex[a] == VKRT_INVALID_INDEX || tex[b] > tex[a])) { uint32_t t = tex[a]; tex[a] = tex[b]; tex[b] = t; }
    for (int a = 0; a < 4; a++) {
        if (tex[a] == VKRT_INVALID_INDEX || tex[a] >= v->textureCount || (a > 0 && tex[a] == tex[a - 1])) continue;
        if (textureUsers(v, tex[a]) != 0u) continue;
        VKRT_Result r = VKRT_removeTexture(v, tex[a]);
        if (r != VKRT_SUCCESS) return r;
    }
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMaterialName(VKRT* v, uint32_t idx, const char* name) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->materialCount || !name) return VKRT_ERROR_INVALID_ARGUMENT;
    snprintf(v->materials[idx].name, VKRT_NAME_LEN, "%s", name[0] ? name : "(unknown)");
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMaterial(VKRT* v, uint32_t idx, const Material* material) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->materialCount || !material) return VKRT_ERROR_INVALID_ARGUMENT;
    Material s = hostSanitizeMaterial(v, *material);
    if (memcmp(&s, &v->materials[idx].material, sizeof(Material)) == 0) return VKRT_SUCCESS;
    v->materials[idx].material = s;
    v->materialsDirty = v->lightsDirty = v->sceneResourcesDirty = v->accelDirty = 1; /* alpha / transmission flags live in the instance records */
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMaterialTexture(VKRT* v, uint32_t materialIndex, uint32_t slot, uint32_t textureIndex) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (materialIndex >= v->materialCount || slot >= VKRT_MATERIAL_TEXTURE_SLOT_COUNT) return VKRT_ERROR_INVALID_ARGUMENT;
    if (textureIndex != VKRT_INVALID_INDEX) {
        if (textureIndex >= v->textureCount) return VKRT_ERROR_INVALID_ARGUMENT;
        /* colour slots take sRGB or linear data, the data slots only linear (scene/textures.c:94-105) */
        const int colourSlot = slot == VKRT_MATERIAL_TEXTURE_SLOT_BASE_COLOR || slot == VKRT_MATERIAL_TEXTURE_SLOT_EMISSIVE;
        const uint32_t cs = v->textures[textureIndex].colorSpace;
        if (!(cs == VKRT_TEXTURE_COLOR_SPACE_LINEAR || (colourSlot && cs == VKRT_TEXTURE_COLOR_SPACE_SRGB))) return VKRT_ERROR_INVALID_ARGUMENT;
    }
    Material m = v->materials[materialIndex].material;
    uint32_t* idx[4] = {&m.baseColorTextureIndex, &m.metallicRoughnessTextureIndex, &m.normalTextureIndex, &m.emissiveTextureIndex};
    uint32_t* wrap[4] = {&m.baseColorTextureWrap, &m.metallicRoughnessTextureWrap, &m.normalTextureWrap, &m.emissiveTextureWrap};
    float* xf[4] = {m.baseColorTextureTransform, m.metallicRoughnessTextureTransform, m.normalTextureTransform, m.emissiveTextureTransform};
    if (*idx[slot] == textureIndex) return VKRT_SUCCESS;
    const uint32_t previous = *idx[slot];
    *idx[slot] = textureIndex;
    *wrap[slot] = VKRT_TEXTURE_WRAP_DEFAULT;
    m.textureTexcoordSets &= ~(0xffu << (8u * slot));
    xf[slot][0] = xf[slot][1] = 1.0f; xf[slot][2] = xf[slot][3] = 0.0f;
    m.textureRotations[slot] = 0.0f;
    VKRT_Result r = VKRT_setMaterial(v, materialIndex, &m);
    if (r != VKRT_SUCCESS) return r;
    if (previous != VKRT_INVALID_INDEX && previous < v->textureCount && textureUsers(v, previous) == 0u) return VKRT_removeTexture(v, previous); /* api/mesh.c:397-402 */
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getMaterialCount(const VKRT* v, uint32_t* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->materialCount;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getMaterialSnapshot(const VKRT* v, uint32_t idx, VKRT_MaterialSnapshot* out) {
    if (!v || !out || idx >= v->materialCount) return VKRT_ERROR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    out->material = v->materials[idx].material;
    for (uint32_t k = 0; k < v->meshCount; k++)
        if (v->meshes[k].info.materialIndex == idx) out->useCount++;
    memcpy(out->name, v->materials[idx].name, VKRT_NAME_LEN);
    return VKRT_SUCCESS;
}

/* ---- meshes --------------------------------------------------------------------------------------------------------------- */
VKRT_Result VKRT_getMeshCount(const VKRT* v, uint32_t* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->meshCount;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getMeshSnapshot(const VKRT* v, uint32_t idx, VKRT_MeshSnapshot* out) {
    if (!v || !out || idx >= v->meshCount) return VKRT_ERROR_INVALID_ARGUMENT;
    const HostMesh* m = &v->meshes[idx];
    memset(out, 0, sizeof(*out));
    out->info = m->info;
    out->materialIndex = m->info.materialIndex;
    if (m->info.materialIndex < v->materialCount) out->material = v->materials[m->info.materialIndex].material;
    out->geometrySource = m->geometrySource;
    out->hasMaterialAssignment = m->hasMaterialAssignment;
    out->ownsGeometry = m->ownsGeometry;
    memcpy(out->name, m->name, VKRT_NAME_LEN);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMeshName(VKRT* v, uint32_t idx, const char* name) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount || !name) return VKRT_ERROR_INVALID_ARGUMENT;
    snprintf(v->meshes[idx].name, VKRT_NAME_LEN, "%s", name[0] ? name : "(unknown)");
    return VKRT_SUCCESS;
}
static int vec3Finite(const float* p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }
static int emissive(const VKRT* v, uint32_t materialIndex) {
    return materialIndex < v->materialCount && v->materials[materialIndex].material.emissionLuminance > 0.0f;
}
VKRT_Result VKRT_setMeshTransformMatrix(VKRT* v, uint32_t idx, vkrt_mat4 world) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount || !world) return VKRT_ERROR_INVALID_ARGUMENT;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            if (!isfinite(world[c][r])) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh* m = &v->meshes[idx];
    if (memcmp(m->worldTransform, world, sizeof(hmat4)) == 0) return VKRT_SUCCESS;
    memcpy(m->worldTransform, world, sizeof(hmat4));
    VKRT_decomposeMeshTransform(m->worldTransform, m->info.position, m->info.rotation, m->info.scale);
    v->sceneResourcesDirty = v->accelDirty = 1;
    if (emissive(v, m->info.materialIndex)) v->lightsDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMeshTransform(VKRT* v, uint32_t idx, vkrt_vec3 position, vkrt_vec3 rotation, vkrt_vec3 scale) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount) return VKRT_ERROR_INVALID_ARGUMENT;
    if ((position && !vec3Finite(position)) || (rotation && !vec3Finite(rotation))) return VKRT_ERROR_INVALID_ARGUMENT;
    if (scale && (!vec3Finite(scale) || scale[0] == 0.0f || scale[1] == 0.0f || scale[2] == 0.0f)) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh* m = &v->meshes[idx];
    float p[3], r[3], s[3];
    memcpy(p, position ? position : m->info.position, sizeof(p));
    memcpy(r, rotation ? rotation : m->info.rotation, sizeof(r));
    memcpy(s, scale ? scale : m->info.scale, sizeof(s));
    hmat4 w;
    VKRT_buildMeshTransformMatrix(p, r, s, w);
    return VKRT_setMeshTransformMatrix(v, idx, w);
}
VKRT_Result VKRT_setMeshMaterialIndex(VKRT* v, uint32_t idx, uint32_t materialIndex) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount || materialIndex >= v->materialCount) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh* m = &v->meshes[idx];
    if (m->info.materialIndex == materialIndex && m->hasMaterialAssignment) return VKRT_SUCCESS;
    int lighting = emissive(v, m->info.materialIndex) || emissive(v, materialIndex);
    m->info.materialIndex = materialIndex;
    m->hasMaterialAssignment = 1;
    v->sceneResourcesDirty = v->accelDirty = 1;
    if (lighting) v->lightsDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_clearMeshMaterialAssignment(VKRT* v, uint32_t idx) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh* m = &v->meshes[idx];
    if (!m->hasMaterialAssignment && m->info.materialIndex == 0u) return VKRT_SUCCESS;
    int lighting = emissive(v, m->info.materialIndex) || emissive(v, 0u);
    m->info.materialIndex = 0u;
    m->hasMaterialAssignment = 0;
    v->sceneResourcesDirty = v->accelDirty = 1;
    if (lighting) v->lightsDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMeshOpacity(VKRT* v, uint32_t idx, float opacity) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount || !isfinite(opacity) || opacity < 0.0f || opacity > 1.0f) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh* m = &v->meshes[idx];
    if (m->info.opacity == opacity) return VKRT_SUCCESS;
    m->info.opacity = opacity;
    v->sceneResourcesDirty = v->accelDirty = 1;
    if (emissive(v, m->info.materialIndex)) v->lightsDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_setMeshRenderBackfaces(VKRT* v, uint32_t idx, uint32_t enabled) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (idx >= v->meshCount) return VKRT_ERROR_INVALID_ARGUMENT;
    HostMesh* m = &v->meshes[idx];
    uint32_t n = enabled ? 1u : 0u;
    if (m->renderBackfacesOverride == (int8_t)n && m->info.renderBackfaces == n) return VKRT_SUCCESS;
    m->renderBackfacesOverride = (int8_t)n;
    m->info.renderBackfaces = n;
    v->sceneResourcesDirty = 1;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}

/* ---- scene update: the host half (always) and the device half (unless hostOnly) ------------------------------------------- */
static VKRT_Result prepareHostScene(VKRT* v) {
    VKRT_Result r;
    if (v->geometryDirty) {
        if ((r = hostPrepareGeometry(v)) != VKRT_SUCCESS) return r;
        v->sceneResourcesDirty = 1;
    }
    if (v->lightsDirty) {
        if ((r = hostRebuildLights(v)) != VKRT_SUCCESS) return r; /* writes MeshInfo.lightPdfArea */
        v->sceneResourcesDirty = 1;
    }
    if (v->sceneResourcesDirty && (r = hostPrepareMeshInfos(v)) != VKRT_SUCCESS) return r;
    if (v->materialsDirty && (r = hostPrepareMaterials(v)) != VKRT_SUCCESS) return r;
    v->sceneData.emissiveMeshCount = v->emissiveMeshCount;
    v->sceneData.emissiveTriangleCount = v->emissiveTriangleCount;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_prepareScene(VKRT* v, VKRT_PreparedScene* out) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (!out) return VKRT_ERROR_INVALID_ARGUMENT;
    /* force a full host rebuild but leave the device-dirty flags set for the next VKRT_updateScene */
    int g = v->geometryDirty, s = v->sceneResourcesDirty, m = v->materialsDirty, l = v->lightsDirty;
    v->geometryDirty = v->sceneResourcesDirty = v->materialsDirty = v->lightsDirty = 1;
    VKRT_Result r = prepareHostScene(v);
    v->geometryDirty |= g; v->sceneResourcesDirty |= s; v->materialsDirty |= m; v->lightsDirty |= l;
    if (r != VKRT_SUCCESS) return r;
    memset(out, 0, sizeof(*out));
    out->vertices = v->packedVertices; out->vertexCount = v->packedVertexCount;
    out->indices = v->packedIndices; out->indexCount = v->packedIndexCount;
    out->meshInfos = v->meshInfos; out->world3x4 = v->world3x4; out->geometrySource = v->geometrySource; out->alphaTested = v->alphaTested;
    out->meshCount = v->meshCount;
    out->materials = v->materialArray; out->materialCount = v->materialCount;
    out->emissiveMeshes = v->emissiveMeshes; out->emissiveMeshCount = v->emissiveMeshCount;
    out->emissiveTriangles = v->emissiveTriangles; out->emissiveTriangleCount = v->emissiveTriangleCount;
    out->meshAliasQ = v->meshAliasQ; out->meshAliasIdx = v->meshAliasIdx; out->triAliasQ = v->triAliasQ; out->triAliasIdx = v->triAliasIdx;
    out->sceneData = &v->sceneData;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_beginFrame(VKRT* v) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    v->frameTraced = v->framePresented = 0;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_updateScene(VKRT* v) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    VKRT_Result r = prepareHostScene(v);
    if (r != VKRT_SUCCESS) return r;
    if (v->hostOnly) {
        v->geometryDirty = v->sceneResourcesDirty = v->materialsDirty = v->lightsDirty = v->texturesDirty = 0;
        return VKRT_SUCCESS;
    }
    vkrt_cuda_ctx* c = v->cuda;
    if (v->geometryDirty) {
        if ((r = cudaCheck(v, vkrt_cuda_set_geometry(c, v->packedVertices, v->packedVertexCount, v->packedIndices, v->packedIndexCount), "set_geometry")) != VKRT_SUCCESS) return r;
        v->accelDirty = 1;
    }
    if (v->texturesDirty) {
        vkrt_cuda_texture* tx = (vkrt_cuda_texture*)calloc(v->textureCount ? v->textureCount : 1u, sizeof(vkrt_cuda_texture));
        if (!tx) return VKRT_ERROR_OUT_OF_MEMORY;
        for (uint32_t i = 0; i < v->textureCount; i++) {
            tx[i].pixels = v->textures[i].pixels; tx[i].width = v->textures[i].width; tx[i].height = v->textures[i].height;
            tx[i].format = v->textures[i].format; tx[i].colorSpace = v->textures[i].colorSpace;
        }
        r = cudaCheck(v, vkrt_cuda_set_textures(c, tx, v->textureCount), "set_textures");
        free(tx);
        if (r != VKRT_SUCCESS) return r;
    }
    if (v->materialsDirty && (r = cudaCheck(v, vkrt_cuda_set_materials(c, v->materialArray, v->materialCount), "set_materials")) != VKRT_SUCCESS) return r;
    if (v->sceneResourcesDirty || v->materialsDirty || v->lightsDirty) {   /* MeshInfo carries materialIndex, opacity and lightPdfArea */
        if ((r = cudaCheck(v, vkrt_cuda_set_instances(c, v->meshInfos, v->world3x4, v->geometrySource, v->alphaTested, v->meshCount), "set_instances")) != VKRT_SUCCESS) return r;
        v->accelDirty = 1;   /* vkrt_cuda_build_accel returns at once when nothing the BVH depends on changed (a colour / roughness edit) */
    }
    if (v->lightsDirty &&
        (r = cudaCheck(v, vkrt_cuda_set_lights(c, v->emissiveMeshes, v->emissiveMeshCount, v->emissiveTriangles, v->emissiveTriangleCount, v->meshAliasQ,
                                               v->meshAliasIdx, v->triAliasQ, v->triAliasIdx), "set_lights")) != VKRT_SUCCESS) return r;
    if (v->rgb2specDirty && v->rgb2spec) {
        if ((r = cudaCheck(v, vkrt_cuda_set_rgb2spec(c, v->rgb2spec, v->rgb2specFloats, v->rgb2specInfo), "set_rgb2spec")) != VKRT_SUCCESS) return r;
        v->rgb2specDirty = 0;
    }
    if (v->accelDirty) {
        if ((r = cudaCheck(v, vkrt_cuda_build_accel(c, &v->buildStats), "build_accel")) != VKRT_SUCCESS) return r;
        v->accelDirty = 0;
    }
    if (v->filmDirty) {
        if ((r = cudaCheck(v, vkrt_cuda_resize(c, v->renderWidth, v->renderHeight), "resize")) != VKRT_SUCCESS) return r;
        v->filmDirty = 0;
        v->accumulationNeedsReset = 0; /* resize clears the film */
    }
    v->geometryDirty = v->sceneResourcesDirty = v->materialsDirty = v->lightsDirty = v->texturesDirty = 0;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_trace(VKRT* v) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (v->hostOnly) return hostFail(v, VKRT_ERROR_OPERATION_FAILED, "VKRT_trace: created hostOnly, no device");
    if (v->renderStatus.renderPhase != VKRT_RENDER_PHASE_SAMPLING && v->renderStatus.renderPhase != VKRT_RENDER_PHASE_INACTIVE) return VKRT_SUCCESS;
    VKRT_Result r;
    if (v->accumulationNeedsReset) { /* record.c:580-585 */
        if ((r = cudaCheck(v, vkrt_cuda_reset_accumulation(v->cuda), "reset_accumulation")) != VKRT_SUCCESS) return r;
        v->accumulationNeedsReset = 0;
    }
    if ((r = cudaCheck(v, vkrt_cuda_render_frame(v->cuda, &v->sceneData, &v->lastFrameStats), "render_frame")) != VKRT_SUCCESS) return r;
    v->renderStatus.renderTimeMs = v->lastFrameStats.frameMs;
    if ((r = hostUpdateAutoExposure(v)) != VKRT_SUCCESS) return r;
    v->totalDeviceMs += v->lastFrameStats.frameMs;
    v->totalExtensionRays += v->lastFrameStats.extensionRays;
    v->totalShadowRays += v->lastFrameStats.shadowRays;
    v->frameTraced = 1;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_present(VKRT* v) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    v->framePresented = 1; /* headless: nothing to show (frame.c:339-368) */
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_stopRenderSampling(VKRT* v) {
    if (!v) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->renderStatus.renderPhase != VKRT_RENDER_PHASE_SAMPLING) return VKRT_SUCCESS;
    v->renderStatus.renderPhase = VKRT_RENDER_PHASE_COMPLETE_RAW;
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_endFrame(VKRT* v) { /* frame.c:370-402 */
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (v->framePresented) {
        int contributed = v->frameTraced && !v->accumulationNeedsReset &&
                          !(v->renderStatus.renderPhase == VKRT_RENDER_PHASE_DENOISING || v->renderStatus.renderPhase >= VKRT_RENDER_PHASE_COMPLETE_RAW);
        if (contributed) {
            const uint32_t renderedSPP = v->sceneData.samplesPerPixel;   /* frame.c:373: what this frame traced, before the controller moves it */
            hostUpdateAutoSPP(v);
            v->renderStatus.accumulationFrame++;
            v->renderStatus.totalSamples += renderedSPP;
            v->sceneData.frameNumber++;
            /* the accumulation read/write swap happens inside vkrt_cuda_render_frame */
        }
        if (v->renderStatus.renderPhase == VKRT_RENDER_PHASE_SAMPLING && v->renderStatus.renderTargetSamples > 0 &&
            v->renderStatus.totalSamples >= v->renderStatus.renderTargetSamples)
            VKRT_stopRenderSampling(v);
    }
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_draw(VKRT* v) {
    VKRT_Result r;
    if ((r = VKRT_beginFrame(v)) != VKRT_SUCCESS) return r;
    if ((r = VKRT_updateScene(v)) != VKRT_SUCCESS) return r;
    if ((r = VKRT_trace(v)) != VKRT_SUCCESS) return r;
    if ((r = VKRT_present(v)) != VKRT_SUCCESS) return r;
    return VKRT_endFrame(v);
}

/* ---- render session ------------------------------------------------------------------------------------------------------- */
VKRT_Result VKRT_startRender(VKRT* v, uint32_t width, uint32_t height, uint32_t targetSamples) {
    if (!v || width == 0 || height == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (width > 16384) width = 16384;
    if (height > 16384) height = 16384;
    if (v->renderWidth != width || v->renderHeight != height) {
        v->renderWidth = width;
        v->renderHeight = height;
        v->filmDirty = 1;
    }
    v->renderStatus.renderPhase = VKRT_RENDER_PHASE_SAMPLING;
    v->renderStatus.renderDenoiseEnabled = 0;
    v->renderStatus.renderTargetSamples = targetSamples;
    v->sceneData.viewportRect[0] = 0; v->sceneData.viewportRect[1] = 0;
    v->sceneData.viewportRect[2] = width; v->sceneData.viewportRect[3] = height;
    hostSyncCameraMatrices(v);
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_continueRender(VKRT* v, uint32_t targetSamples) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (v->renderStatus.renderPhase != VKRT_RENDER_PHASE_COMPLETE_RAW && v->renderStatus.renderPhase != VKRT_RENDER_PHASE_COMPLETE_DENOISED)
        return VKRT_ERROR_OPERATION_FAILED;
    v->renderStatus.renderPhase = VKRT_RENDER_PHASE_SAMPLING;
    v->renderStatus.renderTargetSamples = targetSamples;
    hostWriteSceneStateUniform(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_stopRender(VKRT* v) {
    if (!v) return VKRT_ERROR_INVALID_ARGUMENT;
    if (v->renderStatus.renderPhase == VKRT_RENDER_PHASE_INACTIVE) return VKRT_SUCCESS;
    v->renderStatus.renderPhase = VKRT_RENDER_PHASE_INACTIVE;
    v->renderStatus.renderTargetSamples = 0;
    hostResetSceneData(v);
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getSceneSettings(const VKRT* v, VKRT_SceneSettingsSnapshot* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->sceneSettings;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getRenderStatus(const VKRT* v, VKRT_RenderStatusSnapshot* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->renderStatus;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getSystemInfo(const VKRT* v, VKRT_SystemInfo* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    snprintf(out->deviceName, sizeof(out->deviceName), "%s", v->hostOnly ? "host only (no device)" : vkrt_cuda_version());
    out->vendorID = 0x10DE;
    return VKRT_SUCCESS;
}
vkrt_cuda_ctx* VKRT_cudaContext(VKRT* v) { return v ? v->cuda : NULL; }
VKRT_Result VKRT_getLastFrameStats(const VKRT* v, vkrt_cuda_frame_stats* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->lastFrameStats;
    return VKRT_SUCCESS;
}
VKRT_Result VKRT_getBuildStats(const VKRT* v, vkrt_cuda_build_stats* out) {
    if (!v || !out) return VKRT_ERROR_INVALID_ARGUMENT;
    *out = v->buildStats;
    return VKRT_SUCCESS;
}

/* The reference's offline driver (src/app/render/benchmark.c:13-293) without the wall-clock auto-SPP warm-up: a fixed number of
 * samples per frame so that a run is reproducible; the summary uses the reference's definition 1 sample = 1 spp over the frame. */
VKRT_Result VKRT_appOfflineRender(VKRT* v, uint32_t width, uint32_t height, uint32_t targetSamples, uint32_t samplesPerFrame,
                                  VKRT_OfflineRenderResult* out) {
    VKRT_Result ready = requireReady(v);
    if (ready != VKRT_SUCCESS) return ready;
    if (targetSamples == 0 || samplesPerFrame == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    VKRT_Result r;
    if ((r = VKRT_setSamplesPerPixel(v, samplesPerFrame)) != VKRT_SUCCESS) return r;
    if ((r = VKRT_startRender(v, width, height, targetSamples)) != VKRT_SUCCESS) return r;
    /* scene upload + BVH build happen in the first VKRT_updateScene; keep them out of the timed frames like the reference's setup frames */
    if ((r = VKRT_beginFrame(v)) != VKRT_SUCCESS || (r = VKRT_updateScene(v)) != VKRT_SUCCESS) return r;
    v->totalDeviceMs = 0.0;
    v->totalExtensionRays = v->totalShadowRays = 0;
    double t0 = nowSeconds();
    uint32_t frames = 0;
    while (v->renderStatus.renderPhase == VKRT_RENDER_PHASE_SAMPLING) {
        if ((r = VKRT_draw(v)) != VKRT_SUCCESS) return r;
        frames++;
    }
    double t1 = nowSeconds();
    if (out) {
        memset(out, 0, sizeof(*out));
        out->seconds = t1 - t0;
        out->deviceSeconds = v->totalDeviceMs * 1e-3;
        out->samples = v->renderStatus.totalSamples;
        out->frames = frames;
        out->samplesPerFrame = samplesPerFrame;
        out->samplesPerSecond = out->seconds > 0.0 ? (double)out->samples / out->seconds : 0.0;
        out->mpathsPerSecond = out->samplesPerSecond * (double)width * (double)height / 1e6;
        out->extensionRays = v->totalExtensionRays;
        out->shadowRays = v->totalShadowRays;
    }
    return VKRT_SUCCESS;
}

/* ---- tile partition (shared definition with the CUDA library: csrc/tiles.h) ------------------------------------------------ */
VKRT_Result VKRT_tilePartition(uint32_t width, uint32_t height, uint32_t tileWidth, uint32_t tileHeight, uint32_t rank, uint32_t worldSize,
                               uint32_t* outLocalTileCount, uint32_t* outLocalToGlobalTile, uint32_t capacity, uint32_t* outTilesX, uint32_t* outTilesY) {
    if (width == 0 || height == 0 || worldSize == 0 || rank >= worldSize) return VKRT_ERROR_INVALID_ARGUMENT;
    vkrt_tile_layout lay;
    vkrt_tile_layout_init(&lay, width, height, tileWidth, tileHeight, rank, worldSize);
    if (outLocalTileCount) *outLocalTileCount = lay.localTileCount;
    if (outTilesX) *outTilesX = lay.tilesX;
    if (outTilesY) *outTilesY = lay.tilesY;
    if (outLocalToGlobalTile) {
        if (capacity < lay.localTileCount) return VKRT_ERROR_INVALID_ARGUMENT;
        vkrt_tile_layout_local_tiles(&lay, outLocalToGlobalTile);
    }
    return VKRT_SUCCESS;
}
