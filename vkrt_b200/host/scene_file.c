/* scene_file.c — the app layer above the VKRT_* API: vkrt.scene documents, model import, procedural benchmark scenes.
 *
 * Restates src/app/scene/controller.c:585-638 (material JSON), :737-769 (object transforms), :823-892 (scene settings),
 * :1339-1373 (unreferenced imports are dropped), :1374-1526 (materials / meshes / scene objects), :1528-1597 (load order);
 * src/app/session/session.c:257-300 (object hierarchy -> world matrices); src/app/mesh/controller.c:539+ (standalone import).
 * The procedural generators implement SURVEY.md §8(d) configs C3 (triangle soup) and C4 (instanced model). */
#include <libgen.h>

#include "gltf_import.h"
#include "hjson.h"
#include "host_state.h"

static char* readTextFile(const char* path, size_t* outSize) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* s = (char*)malloc((size_t)n + 1);
    if (!s || fread(s, 1, (size_t)n, f) != (size_t)n) { free(s); fclose(f); return NULL; }
    fclose(f);
    s[n] = 0;
    *outSize = (size_t)n;
    return s;
}

static void optFloat(const hj_value* o, const char* key, float* out) {
    const hj_value* v = hj_get(o, key);
    if (v && v->type == HJ_NUMBER) *out = (float)v->number;
}
static void optFloats(const hj_value* o, const char* key, float* out, size_t n) {
    const hj_value* v = hj_get(o, key);
    if (v && v->type == HJ_ARRAY && v->count == n) hj_floats(v, out, n);
}
static void optUInt(const hj_value* o, const char* key, uint32_t* out) {
    const hj_value* v = hj_get(o, key);
    if (v && v->type == HJ_NUMBER && v->number >= 0.0) *out = (uint32_t)v->number;
}
static void optIndex(const hj_value* o, const char* key, uint32_t* out) { /* number, or null = VKRT_INVALID_INDEX */
    const hj_value* v = hj_get(o, key);
    if (!v) return;
    if (v->type == HJ_NULL) *out = VKRT_INVALID_INDEX;
    else if (v->type == HJ_NUMBER && v->number >= 0.0) *out = (uint32_t)v->number;
}

/* controller.c:368-377 jsonToUInt32: a JSON number that is a non-negative integer below 2^32, nothing else */
static int jsonToUInt32(const hj_value* v, uint32_t* out) {
    *out = 0u;
    if (!v || v->type != HJ_NUMBER) return 0;
    const double d = hj_number(v, -1.0);
    if (!(d >= 0.0) || d > 4294967295.0) return 0;
    const uint32_t c = (uint32_t)d;
    if ((double)c != d) return 0;
    *out = c;
    return 1;
}
/* Slots a scene file may name. The reference grows its material list one VKRT_addMaterial at a time up to whatever index the file
 * states (controller.c:933-945), i.e. until memory runs out for a corrupt index; here such a file is refused instead. */
#define SCENE_MAX_MATERIAL_INDEX (1u << 20)

static Material parseMaterialJson(const hj_value* o) {
    Material m = VKRT_materialDefault();
    optFloats(o, "baseColor", m.baseColor, 3);
    optFloat(o, "roughness", &m.roughness);
    optFloats(o, "emissionColor", m.emissionColor, 3);
    optFloat(o, "emissionLuminance", &m.emissionLuminance);
    optFloats(o, "eta", m.eta, 3);
    optFloat(o, "metallic", &m.metallic);
    optFloats(o, "k", m.k, 3);
    optFloat(o, "anisotropic", &m.anisotropic);
    optFloat(o, "specular", &m.specular);
    optFloat(o, "specularTint", &m.specularTint);
    optFloats(o, "sheenTintWeight", m.sheenTintWeight, 4);
    optFloat(o, "clearcoat", &m.clearcoat);
    optFloat(o, "clearcoatGloss", &m.clearcoatGloss);
    optFloat(o, "ior", &m.ior);
    optFloat(o, "abbeNumber", &m.abbeNumber);
    optFloat(o, "diffuseRoughness", &m.diffuseRoughness);
    optFloat(o, "transmission", &m.transmission);
    optFloat(o, "subsurface", &m.subsurface);
    optFloat(o, "sheenRoughness", &m.sheenRoughness);
    optFloat(o, "absorptionCoefficient", &m.absorptionCoefficient);
    optFloats(o, "attenuationColor", m.attenuationColor, 3);
    optFloat(o, "normalTextureScale", &m.normalTextureScale);
    optIndex(o, "baseColorTextureIndex", &m.baseColorTextureIndex);
    optIndex(o, "metallicRoughnessTextureIndex", &m.metallicRoughnessTextureIndex);
    optIndex(o, "normalTextureIndex", &m.normalTextureIndex);
    optIndex(o, "emissiveTextureIndex", &m.emissiveTextureIndex);
    optUInt(o, "baseColorTextureWrap", &m.baseColorTextureWrap);
    optUInt(o, "metallicRoughnessTextureWrap", &m.metallicRoughnessTextureWrap);
    optUInt(o, "normalTextureWrap", &m.normalTextureWrap);
    optUInt(o, "emissiveTextureWrap", &m.emissiveTextureWrap);
    optFloat(o, "opacity", &m.opacity);
    optFloat(o, "alphaCutoff", &m.alphaCutoff);
    optUInt(o, "alphaMode", &m.alphaMode);
    optUInt(o, "textureTexcoordSets", &m.textureTexcoordSets);
    optFloats(o, "baseColorTextureTransform", m.baseColorTextureTransform, 4);
    optFloats(o, "metallicRoughnessTextureTransform", m.metallicRoughnessTextureTransform, 4);
    optFloats(o, "normalTextureTransform", m.normalTextureTransform, 4);
    optFloats(o, "emissiveTextureTransform", m.emissiveTextureTransform, 4);
    optFloats(o, "textureRotations", m.textureRotations, 4);
    return m;
}

/* Uploads every mesh of one .glb as a batch; returns the index of its first mesh. */
/* The import's decoded textures, appended in import order; *outTextureBase is the scene index of the import's texture 0. */
static VKRT_Result uploadImportTextures(VKRT* vkrt, const GltfImport* imp, uint32_t* outTextureBase) {
    *outTextureBase = vkrt->textureCount;
    for (uint32_t i = 0; i < imp->textureCount; i++) {
        const GltfTexture* t = &imp->textures[i];
        VKRT_TextureUpload up = {t->name, t->pixels, t->width, t->height, t->format, t->colorSpace};
        VKRT_Result r = VKRT_addTextureFromPixels(vkrt, &up, NULL);
        if (r != VKRT_SUCCESS) return r;
    }
    return VKRT_SUCCESS;
}
static Material rebaseMaterialTextures(Material m, uint32_t textureBase) {
    uint32_t* idx[4] = {&m.baseColorTextureIndex, &m.metallicRoughnessTextureIndex, &m.normalTextureIndex, &m.emissiveTextureIndex};
    for (int k = 0; k < 4; k++)
        if (*idx[k] != VKRT_INVALID_INDEX) *idx[k] += textureBase;
    return m;
}

static VKRT_Result uploadImport(VKRT* vkrt, const GltfImport* imp, uint32_t* outFirst) {
    *outFirst = vkrt->meshCount;
    if (imp->meshCount == 0) return VKRT_SUCCESS;
    VKRT_MeshUpload* ups = (VKRT_MeshUpload*)calloc(imp->meshCount, sizeof(VKRT_MeshUpload));
    if (!ups) return VKRT_ERROR_OUT_OF_MEMORY;
    for (uint32_t i = 0; i < imp->meshCount; i++) {
        ups[i].vertices = imp->meshes[i].vertices; ups[i].vertexCount = imp->meshes[i].vertexCount;
        ups[i].indices = imp->meshes[i].indices; ups[i].indexCount = imp->meshes[i].indexCount;
    }
    VKRT_Result r = VKRT_uploadMeshDataBatch(vkrt, ups, imp->meshCount);
    free(ups);
    if (r != VKRT_SUCCESS) return r;
    for (uint32_t i = 0; i < imp->meshCount; i++) {
        uint32_t mi = *outFirst + i;
        VKRT_setMeshName(vkrt, mi, imp->meshes[i].name);
        vkrt->meshes[mi].info.renderBackfaces = imp->meshes[i].doubleSided ? 1u : 0u; /* material default; an explicit override wins */
    }
    return VKRT_SUCCESS;
}

VKRT_Result VKRT_appImportMesh(VKRT* vkrt, const char* glbPath, uint32_t* outFirstMesh, uint32_t* outMeshCount) {
    if (!vkrt || !glbPath) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!vkrt->initialized) return VKRT_ERROR_OPERATION_FAILED;
    GltfImport imp;
    char err[256];
    if (!gltfImportFile(glbPath, &imp, err, sizeof(err))) return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "import %s: %s", glbPath, err);
    uint32_t first = 0;
    VKRT_Result r = uploadImport(vkrt, &imp, &first);
    if (r == VKRT_SUCCESS) {
        /* standalone import: the file's materials are appended and assigned; node transforms are applied */
        uint32_t materialBase = vkrt->materialCount, textureBase = 0;
        r = uploadImportTextures(vkrt, &imp, &textureBase);
        for (uint32_t i = 0; i < imp.materialCount && r == VKRT_SUCCESS; i++) {
            Material m = rebaseMaterialTextures(imp.materials[i], textureBase);
            r = VKRT_addMaterial(vkrt, &m, imp.materialNames[i], NULL);
        }
        for (uint32_t i = 0; i < imp.meshCount && r == VKRT_SUCCESS; i++) {
            if (imp.meshes[i].materialIndex >= 0) r = VKRT_setMeshMaterialIndex(vkrt, first + i, materialBase + (uint32_t)imp.meshes[i].materialIndex);
            if (r == VKRT_SUCCESS) r = VKRT_setMeshTransformMatrix(vkrt, first + i, imp.meshes[i].world);
        }
    }
    if (outFirstMesh) *outFirstMesh = first;
    if (outMeshCount) *outMeshCount = imp.meshCount;
    gltfImportFree(&imp);
    return r;
}

VKRT_Result VKRT_appLoadScene(VKRT* vkrt, const char* scenePath) {
    if (!vkrt || !scenePath) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!vkrt->initialized) return VKRT_ERROR_OPERATION_FAILED;
    size_t size = 0;
    char* text = readTextFile(scenePath, &size);
    if (!text) return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "cannot read scene %s", scenePath);
    char err[256];
    hj_value* root = hj_parse(text, size, err, sizeof(err));
    free(text);
    if (!root) return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "%s: %s", scenePath, err);
    VKRT_Result r = VKRT_SUCCESS;
    uint32_t* importFirst = NULL;
    uint32_t* importCount = NULL;
    uint32_t* savedToLoaded = NULL;
    uint32_t* textureMap = NULL;
    uint32_t textureMapCount = 0;
    hmat4* worlds = NULL;
    if (strcmp(hj_string(hj_get(root, "format"), ""), "vkrt.scene") != 0 || (int)hj_number(hj_get(root, "version"), 0) != 1) {
        r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "%s: not a vkrt.scene version 1 document", scenePath);
        goto done;
    }
    if (vkrt->meshCount != 0) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "VKRT_appLoadScene needs an empty scene"); goto done; }

    /* 1. import every listed model (a file may repeat: its geometry dedups into instances) */
    char* pathCopy = strdup(scenePath);
    const char* baseDir = dirname(pathCopy);
    const hj_value* imports = hj_get(root, "meshImports");
    size_t nImports = hj_count(imports);
    importFirst = (uint32_t*)calloc(nImports ? nImports : 1, sizeof(uint32_t));
    importCount = (uint32_t*)calloc(nImports ? nImports : 1, sizeof(uint32_t));
    if (!pathCopy || !importFirst || !importCount) { free(pathCopy); r = hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory"); goto done; }
    for (size_t i = 0; i < nImports && r == VKRT_SUCCESS; i++) {
        const char* rel = hj_string(hj_at(imports, i), NULL);
        if (!rel) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "meshImports[%zu] is not a path", i); break; }
        char full[4096];
        if (rel[0] == '/') snprintf(full, sizeof(full), "%s", rel);
        else snprintf(full, sizeof(full), "%s/%s", baseDir, rel);
        GltfImport imp;
        if (!gltfImportFile(full, &imp, err, sizeof(err))) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "import %s: %s", full, err); break; }
        r = uploadImport(vkrt, &imp, &importFirst[i]);
        importCount[i] = imp.meshCount;
        /* the file's textures keep their import-order indices, which is what the saved materials name (controller.c:1586-1593) */
        uint32_t textureBase = 0;
        if (r == VKRT_SUCCESS) r = uploadImportTextures(vkrt, &imp, &textureBase);
        gltfImportFree(&imp);
    }
    /* 1b. standalone textures at their saved indices, then the environment map (controller.c:690-716,1274-1312) */
    const hj_value* texImports = hj_get(root, "textureImports");
    for (size_t k = 0; k < hj_count(texImports); k++) {
        uint32_t saved = VKRT_INVALID_INDEX;
        if (!jsonToUInt32(hj_get(hj_at(texImports, k), "index"), &saved)) saved = VKRT_INVALID_INDEX;
        if (saved != VKRT_INVALID_INDEX && saved >= VKRT_MAX_BINDLESS_TEXTURES) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "textureImports[%zu]: index %u out of range", k, saved); goto done; }
        if (saved != VKRT_INVALID_INDEX && saved + 1u > textureMapCount) textureMapCount = saved + 1u;
    }
    textureMap = (uint32_t*)malloc((textureMapCount ? textureMapCount : 1u) * sizeof(uint32_t));
    for (uint32_t k = 0; k < textureMapCount; k++) textureMap[k] = VKRT_INVALID_INDEX;
    for (size_t k = 0; k < hj_count(texImports) && r == VKRT_SUCCESS; k++) {
        const hj_value* jt = hj_at(texImports, k);
        const char* rel = hj_string(hj_get(jt, "path"), NULL);
        const hj_value* jIndex = hj_get(jt, "index");
        const hj_value* jSpace = hj_get(jt, "colorSpace");
        if (!rel || !jIndex || jIndex->type != HJ_NUMBER || !jSpace || jSpace->type != HJ_NUMBER) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "textureImports[%zu] malformed", k); break; }
        char full[4096];
        if (rel[0] == '/') snprintf(full, sizeof(full), "%s", rel);
        else snprintf(full, sizeof(full), "%s/%s", baseDir, rel);
        uint32_t loaded = VKRT_INVALID_INDEX;
        r = VKRT_addTextureFromFile(vkrt, full, NULL, (uint32_t)jSpace->number, &loaded);
        if (r != VKRT_SUCCESS) { hostFail(vkrt, r, "textureImports[%zu]: %s: %s", k, full, VKRT_lastError(vkrt)); break; }
        textureMap[(uint32_t)jIndex->number] = loaded;
    }
    const char* envRel = hj_string(hj_get(root, "environmentTexturePath"), NULL);
    if (r == VKRT_SUCCESS && envRel && envRel[0]) {
        char full[4096];
        if (envRel[0] == '/') snprintf(full, sizeof(full), "%s", envRel);
        else snprintf(full, sizeof(full), "%s/%s", baseDir, envRel);
        r = VKRT_setEnvironmentTextureFromFile(vkrt, full);
        if (r != VKRT_SUCCESS) hostFail(vkrt, r, "environmentTexturePath %s: %s", full, VKRT_lastError(vkrt));
    }
    free(pathCopy);
    if (r != VKRT_SUCCESS) goto done;

    /* 2. saved mesh index -> loaded mesh index; imported meshes nobody references are removed (controller.c:1339-1373) */
    const hj_value* meshes = hj_get(root, "meshes");
    size_t nSaved = hj_count(meshes);
    savedToLoaded = (uint32_t*)calloc(nSaved ? nSaved : 1, sizeof(uint32_t));
    unsigned char* keep = (unsigned char*)calloc(vkrt->meshCount ? vkrt->meshCount : 1, 1);
    if (!savedToLoaded || !keep) { free(keep); r = hostFail(vkrt, VKRT_ERROR_OUT_OF_MEMORY, "out of memory"); goto done; }
    for (size_t k = 0; k < nSaved; k++) {
        const hj_value* jm = hj_at(meshes, k);
        uint32_t ii, li;
        if (!jsonToUInt32(hj_get(jm, "importIndex"), &ii)) ii = VKRT_INVALID_INDEX;
        if (!jsonToUInt32(hj_get(jm, "importLocalIndex"), &li)) li = VKRT_INVALID_INDEX;
        if (ii >= nImports || li >= importCount[ii]) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "meshes[%zu]: bad import reference", k); break; }
        savedToLoaded[k] = importFirst[ii] + li;
        keep[savedToLoaded[k]] = 1;
    }
    if (r == VKRT_SUCCESS) {
        for (uint32_t mi = vkrt->meshCount; mi-- > 0;) {
            if (keep[mi]) continue;
            VKRT_removeMesh(vkrt, mi);
            for (size_t k = 0; k < nSaved; k++)
                if (savedToLoaded[k] > mi) savedToLoaded[k]--;
        }
    }
    free(keep);
    if (r != VKRT_SUCCESS) goto done;

    /* 3. materials at their saved slots (slot 0 stays the default material) */
    const hj_value* mats = hj_get(root, "materials");
    uint32_t highest = 0;
    /* controller.c:905-927: an explicit material index and every mesh's materialIndex must be valid unsigned integers */
    for (size_t k = 0; k < hj_count(mats); k++) {
        uint32_t idx = (uint32_t)k;
        const hj_value* explicitIndex = hj_get(hj_at(mats, k), "index");
        if (explicitIndex && !jsonToUInt32(explicitIndex, &idx)) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "materials[%zu]: bad index", k); goto done; }
        if (idx > highest) highest = idx;
    }
    for (size_t k = 0; k < nSaved; k++) {
        uint32_t idx;
        if (!jsonToUInt32(hj_get(hj_at(meshes, k), "materialIndex"), &idx)) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "meshes[%zu]: bad materialIndex", k); goto done; }
        if (idx > highest) highest = idx;
    }
    if (highest > SCENE_MAX_MATERIAL_INDEX) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "material index %u out of range", highest); goto done; }
    while (vkrt->materialCount < highest + 1u && r == VKRT_SUCCESS) r = VKRT_addMaterial(vkrt, NULL, NULL, NULL);
    for (size_t k = 0; k < hj_count(mats) && r == VKRT_SUCCESS; k++) {
        const hj_value* jm = hj_at(mats, k);
        uint32_t idx = (uint32_t)k;
        if (hj_get(jm, "index")) (void)jsonToUInt32(hj_get(jm, "index"), &idx);   /* validated above */
        const hj_value* body = hj_get(jm, "material");
        if (!body || body->type != HJ_OBJECT || !hj_string(hj_get(jm, "name"), NULL)) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "materials[%zu] malformed", k); break; }
        Material m = parseMaterialJson(body);
        {   /* saved texture indices -> loaded ones (controller.c:640-662 remapStandaloneTextureIndices) */
            uint32_t* ti[4] = {&m.baseColorTextureIndex, &m.metallicRoughnessTextureIndex, &m.normalTextureIndex, &m.emissiveTextureIndex};
            for (int q = 0; q < 4; q++)
                if (*ti[q] != VKRT_INVALID_INDEX && *ti[q] < textureMapCount && textureMap[*ti[q]] != VKRT_INVALID_INDEX) *ti[q] = textureMap[*ti[q]];
        }
        if ((r = VKRT_setMaterial(vkrt, idx, &m)) == VKRT_SUCCESS) r = VKRT_setMaterialName(vkrt, idx, hj_string(hj_get(jm, "name"), ""));
    }
    /* 4. per-mesh state */
    for (size_t k = 0; k < nSaved && r == VKRT_SUCCESS; k++) {
        const hj_value* jm = hj_at(meshes, k);
        uint32_t mi = savedToLoaded[k];
        VKRT_setMeshName(vkrt, mi, hj_string(hj_get(jm, "name"), "mesh"));
        int assigned = hj_bool(hj_get(jm, "hasMaterialAssignment"), 0);
        uint32_t meshMaterial = 0u;
        (void)jsonToUInt32(hj_get(jm, "materialIndex"), &meshMaterial);   /* validated above */
        r = assigned ? VKRT_setMeshMaterialIndex(vkrt, mi, meshMaterial) : VKRT_clearMeshMaterialAssignment(vkrt, mi);
        if (r == VKRT_SUCCESS) r = VKRT_setMeshOpacity(vkrt, mi, (float)hj_number(hj_get(jm, "opacity"), 1.0));
        if (r == VKRT_SUCCESS) r = VKRT_setMeshRenderBackfaces(vkrt, mi, hj_bool(hj_get(jm, "renderBackfaces"), 0) ? 1u : 0u);
    }
    /* 5. scene objects: world = parent world * local; parents precede children (controller.c:1492) */
    const hj_value* objects = hj_get(root, "sceneObjects");
    size_t nObj = hj_count(objects);
    worlds = (hmat4*)calloc(nObj ? nObj : 1, sizeof(hmat4));
    for (size_t k = 0; k < nObj && r == VKRT_SUCCESS; k++) {
        const hj_value* jo = hj_at(objects, k);
        hmat4 local;
        if (hj_get(jo, "localPosition") || hj_get(jo, "localRotation") || hj_get(jo, "localScale")) {
            float p[3] = {0, 0, 0}, rot[3] = {0, 0, 0}, s[3] = {1, 1, 1};
            if (hj_floats(hj_get(jo, "localPosition"), p, 3) != 3 || hj_floats(hj_get(jo, "localRotation"), rot, 3) != 3 || hj_floats(hj_get(jo, "localScale"), s, 3) != 3) {
                r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "sceneObjects[%zu]: incomplete transform", k);
                break;
            }
            VKRT_buildMeshTransformMatrix(p, rot, s, local);
        } else {
            float f[16];
            if (hj_floats(hj_get(jo, "localTransform"), f, 16) != 16) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "sceneObjects[%zu]: no transform", k); break; }
            memcpy(local, f, sizeof(f));
        }
        const hj_value* parent = hj_get(jo, "parentIndex");
        if (parent && parent->type == HJ_NUMBER) {
            size_t pi = (size_t)parent->number;
            if (pi >= k) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "sceneObjects[%zu]: parent must precede child", k); break; }
            h_mat4_mul(worlds[pi], local, worlds[k]);
        } else {
            h_mat4_copy(local, worlds[k]);
        }
        const hj_value* meshRef = hj_get(jo, "meshIndex");
        if (meshRef && meshRef->type == HJ_NUMBER) {
            size_t saved = (size_t)meshRef->number;
            if (saved >= nSaved) { r = hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "sceneObjects[%zu]: bad meshIndex", k); break; }
            r = VKRT_setMeshTransformMatrix(vkrt, savedToLoaded[saved], worlds[k]);
        }
    }
    /* 6. settings, in the reference's order (controller.c:878-891) */
    const hj_value* ss = hj_get(root, "sceneSettings");
    if (r == VKRT_SUCCESS && ss && ss->type == HJ_OBJECT) {
        VKRT_SceneSettingsSnapshot cur = vkrt->sceneSettings;
        const hj_value* cam = hj_get(ss, "camera");
        float pos[3], tgt[3], up[3], vfov = cur.camera.vfov, env[3], envStrength = cur.environmentStrength, envRot = cur.environmentRotation, exposure = cur.exposure;
        memcpy(pos, cur.camera.pos, sizeof(pos)); memcpy(tgt, cur.camera.target, sizeof(tgt)); memcpy(up, cur.camera.up, sizeof(up));
        memcpy(env, cur.environmentColor, sizeof(env));
        uint32_t rrMin = cur.rrMinDepth, rrMax = cur.rrMaxDepth, tone = cur.toneMappingMode, mode = cur.renderMode, spectral = cur.spectralSamplingMode;
        optFloats(cam, "position", pos, 3); optFloats(cam, "target", tgt, 3); optFloats(cam, "up", up, 3); optFloat(cam, "vfov", &vfov);
        optUInt(ss, "rrMinDepth", &rrMin); optUInt(ss, "rrMaxDepth", &rrMax); optUInt(ss, "toneMappingMode", &tone);
        optUInt(ss, "renderMode", &mode); optUInt(ss, "spectralSamplingMode", &spectral);
        optFloat(ss, "exposure", &exposure); optFloats(ss, "environmentColor", env, 3);
        optFloat(ss, "environmentStrength", &envStrength); optFloat(ss, "environmentRotation", &envRot);
        int nee = hj_bool(hj_get(ss, "misNeeEnabled"), (int)cur.misNeeEnabled);
        if ((r = VKRT_setPathDepth(vkrt, rrMin, rrMax)) == VKRT_SUCCESS && (r = VKRT_setToneMappingMode(vkrt, tone)) == VKRT_SUCCESS &&
            (r = VKRT_setRenderMode(vkrt, mode)) == VKRT_SUCCESS && (r = VKRT_setSpectralSamplingMode(vkrt, spectral)) == VKRT_SUCCESS &&
            (r = VKRT_setExposure(vkrt, exposure)) == VKRT_SUCCESS && (r = VKRT_setEnvironmentLight(vkrt, env, envStrength)) == VKRT_SUCCESS &&
            (r = VKRT_setEnvironmentRotation(vkrt, envRot)) == VKRT_SUCCESS && (r = VKRT_setMisNeeEnabled(vkrt, nee ? 1u : 0u)) == VKRT_SUCCESS)
            r = VKRT_cameraSetPose(vkrt, pos, tgt, up, vfov);
    }
done:
    free(importFirst); free(importCount); free(savedToLoaded); free(worlds); free(textureMap);
    hj_free(root);
    return r;
}

/* ---- procedural scenes ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t state, inc; } pcg32;
static uint32_t pcgNext(pcg32* g) {
    uint64_t old = g->state;
    g->state = old * 6364136223846793005ull + g->inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
}
static void pcgSeed(pcg32* g, uint64_t seed) { g->state = 0; g->inc = 3; pcgNext(g); g->state += seed; pcgNext(g); }
static float pcgFloat(pcg32* g) { return (float)(pcgNext(g) >> 8) * (1.0f / 16777216.0f); }
static void pcgUnitVector(pcg32* g, float out[3]) {
    float z = 2.0f * pcgFloat(g) - 1.0f, phi = 2.0f * H_PI * pcgFloat(g), r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    out[0] = r * cosf(phi); out[1] = r * sinf(phi); out[2] = z;
}

static VKRT_Result addQuadLight(VKRT* vkrt, float halfSize, float z, float luminance) {
    Vertex* v = NULL;
    if (posix_memalign((void**)&v, 16, 4 * sizeof(Vertex)) != 0) return VKRT_ERROR_OUT_OF_MEMORY;
    memset(v, 0, 4 * sizeof(Vertex));
    const float p[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
    for (int i = 0; i < 4; i++) {
        v[i].position[0] = p[i][0] * halfSize; v[i].position[1] = p[i][1] * halfSize; v[i].position[3] = 1.0f;
        v[i].normal[2] = 1.0f;
        v[i].tangent[0] = 1.0f; v[i].tangent[3] = 1.0f;
        v[i].color[0] = v[i].color[1] = v[i].color[2] = v[i].color[3] = 1.0f;
        v[i].texcoord0[0] = 0.5f * (p[i][0] + 1.0f); v[i].texcoord0[1] = 0.5f * (p[i][1] + 1.0f);
    }
    const uint32_t idx[6] = {0, 1, 2, 0, 2, 3};
    uint32_t mi = vkrt->meshCount;
    VKRT_Result r = VKRT_uploadMeshData(vkrt, v, 4, idx, 6);
    free(v);
    if (r != VKRT_SUCCESS) return r;
    Material lm = VKRT_materialDefault();
    lm.emissionLuminance = luminance;
    uint32_t li = 0;
    if ((r = VKRT_addMaterial(vkrt, &lm, "light", &li)) != VKRT_SUCCESS) return r;
    if ((r = VKRT_setMeshMaterialIndex(vkrt, mi, li)) != VKRT_SUCCESS) return r;
    float pos[3] = {0, 0, z}, rot[3] = {180, 0, 0}, one[3] = {1, 1, 1}; /* facing down */
    VKRT_setMeshName(vkrt, mi, "light");
    return VKRT_setMeshTransform(vkrt, mi, pos, rot, one);
}

/* C3: N triangles, centres uniform in [-1,1]^3, edge length log-uniform in [2e-3, 2e-2] (x (1e7/N)^(1/3) so that smaller soups stay
 * equally dense), random orientation, unindexed; triangle i belongs to mesh i % 16; meshes 0-7 diffuse, 8-15 glossy metal. */
VKRT_Result VKRT_appGenerateSoup(VKRT* vkrt, uint32_t triangleCount, uint32_t seed) {
    if (!vkrt || triangleCount == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!vkrt->initialized) return VKRT_ERROR_OPERATION_FAILED;
    const uint32_t meshCount = 16;
    pcg32 g;
    pcgSeed(&g, seed ? seed : 0x5EED0001u);
    Vertex* verts[16];
    uint32_t counts[16], fill[16];
    for (uint32_t k = 0; k < meshCount; k++) {
        counts[k] = triangleCount / meshCount + (k < triangleCount % meshCount ? 1u : 0u);
        fill[k] = 0;
        verts[k] = NULL;
        if (counts[k] && posix_memalign((void**)&verts[k], 16, (size_t)counts[k] * 3 * sizeof(Vertex)) != 0) {
            for (uint32_t j = 0; j < k; j++) free(verts[j]);
            return VKRT_ERROR_OUT_OF_MEMORY;
        }
    }
    const float density = cbrtf(1e7f / (float)triangleCount);
    const float logLo = logf(2e-3f), logHi = logf(2e-2f);
    for (uint32_t t = 0; t < triangleCount; t++) {
        float c[3] = {2.0f * pcgFloat(&g) - 1.0f, 2.0f * pcgFloat(&g) - 1.0f, 2.0f * pcgFloat(&g) - 1.0f};
        float edge = expf(logLo + (logHi - logLo) * pcgFloat(&g)) * density;
        float a[3], b[3], n[3];
        pcgUnitVector(&g, a);
        pcgUnitVector(&g, b);
        float d = h_dot3(a, b);
        for (int k = 0; k < 3; k++) b[k] -= a[k] * d;
        if (h_dot3(b, b) < 1e-8f) { b[0] = -a[1]; b[1] = a[0]; b[2] = 0.0f; if (h_dot3(b, b) < 1e-8f) { b[0] = 0; b[1] = -a[2]; b[2] = a[1]; } }
        h_normalize3(b);
        h_cross3(a, b, n);
        float p[3][3];
        for (int k = 0; k < 3; k++) {
            p[0][k] = c[k] - 0.5f * edge * a[k] - 0.3f * edge * b[k];
            p[1][k] = c[k] + 0.5f * edge * a[k] - 0.3f * edge * b[k];
            p[2][k] = c[k] + 0.6f * edge * b[k];
        }
        uint32_t mk = t % meshCount;
        Vertex* v = verts[mk] + (size_t)fill[mk] * 3;
        fill[mk]++;
        for (int j = 0; j < 3; j++) {
            memset(&v[j], 0, sizeof(Vertex));
            v[j].position[0] = p[j][0]; v[j].position[1] = p[j][1]; v[j].position[2] = p[j][2]; v[j].position[3] = 1.0f;
            v[j].normal[0] = n[0]; v[j].normal[1] = n[1]; v[j].normal[2] = n[2];
            v[j].tangent[0] = a[0]; v[j].tangent[1] = a[1]; v[j].tangent[2] = a[2]; v[j].tangent[3] = 1.0f;
            v[j].color[0] = v[j].color[1] = v[j].color[2] = v[j].color[3] = 1.0f;
        }
    }
    VKRT_Result r = VKRT_SUCCESS;
    uint32_t maxCount = 0;
    for (uint32_t k = 0; k < meshCount; k++) if (counts[k] > maxCount) maxCount = counts[k];
    uint32_t* iota = (uint32_t*)malloc((size_t)maxCount * 3 * sizeof(uint32_t));
    if (!iota) r = VKRT_ERROR_OUT_OF_MEMORY;
    for (uint32_t i = 0; iota && i < maxCount * 3; i++) iota[i] = i;
    const float gloss[3] = {0.05f, 0.2f, 0.4f};
    for (uint32_t k = 0; k < meshCount && r == VKRT_SUCCESS; k++) {
        if (!counts[k]) continue;
        uint32_t mi = vkrt->meshCount;
        if ((r = VKRT_uploadMeshData(vkrt, verts[k], (size_t)counts[k] * 3, iota, (size_t)counts[k] * 3)) != VKRT_SUCCESS) break;
        Material m = VKRT_materialDefault();
        if (k < 8) {
            m.roughness = 1.0f;
            for (int ch = 0; ch < 3; ch++) m.baseColor[ch] = 0.2f + 0.6f * pcgFloat(&g);
        } else {
            m.metallic = 1.0f;
            m.roughness = gloss[k % 3];
            for (int ch = 0; ch < 3; ch++) m.baseColor[ch] = 0.5f + 0.4f * pcgFloat(&g);
        }
        uint32_t matIndex = 0;
        char name[32];
        snprintf(name, sizeof(name), "soup%u", k);
        if ((r = VKRT_addMaterial(vkrt, &m, name, &matIndex)) != VKRT_SUCCESS) break;
        VKRT_setMeshName(vkrt, mi, name);
        r = VKRT_setMeshMaterialIndex(vkrt, mi, matIndex);
    }
    free(iota);
    for (uint32_t k = 0; k < meshCount; k++) free(verts[k]);
    if (r != VKRT_SUCCESS) return r;
    if ((r = addQuadLight(vkrt, 0.5f, 1.5f, 30.0f)) != VKRT_SUCCESS) return r;
    float pos[3] = {0.0f, -3.6f, 0.4f}, tgt[3] = {0, 0, 0}, up[3] = {0, 0, 1};
    return VKRT_cameraSetPose(vkrt, pos, tgt, up, 40.0f);
}

/* C4: `instanceCount` instances of one model on a jittered cubic grid in [-1,1]^3, random Euler rotation, uniform scale in [0.3, 0.6] /
 * gridSide, four materials cycled, a floor and an area light; one BLAS, a TLAS over all instances. */
VKRT_Result VKRT_appGenerateInstanced(VKRT* vkrt, const char* glbPath, uint32_t instanceCount, uint32_t seed) {
    if (!vkrt || !glbPath || instanceCount == 0) return VKRT_ERROR_INVALID_ARGUMENT;
    if (!vkrt->initialized) return VKRT_ERROR_OPERATION_FAILED;
    GltfImport imp;
    char err[256];
    if (!gltfImportFile(glbPath, &imp, err, sizeof(err))) return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "import %s: %s", glbPath, err);
    if (imp.meshCount == 0) { gltfImportFree(&imp); return hostFail(vkrt, VKRT_ERROR_OPERATION_FAILED, "%s has no triangle mesh", glbPath); }
    const GltfMesh* src = &imp.meshes[0];
    /* normalise the model into a unit box so that "scale" means the same for every input */
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < src->vertexCount; i++)
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], src->vertices[i].position[k]); hi[k] = fmaxf(hi[k], src->vertices[i].position[k]); }
    float extent = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
    float norm = extent > 0.0f ? 2.0f / extent : 1.0f;
    for (size_t i = 0; i < src->vertexCount; i++)
        for (int k = 0; k < 3; k++) src->vertices[i].position[k] = (src->vertices[i].position[k] - 0.5f * (lo[k] + hi[k])) * norm;
    Material mats[4];
    const float cols[4][3] = {{0.8f, 0.3f, 0.2f}, {0.2f, 0.6f, 0.8f}, {0.9f, 0.9f, 0.9f}, {0.3f, 0.8f, 0.3f}};
    const float rough[4] = {0.6f, 0.2f, 0.05f, 1.0f}, metal[4] = {0.0f, 1.0f, 0.0f, 0.0f};
    uint32_t matIndex[4];
    VKRT_Result r = VKRT_SUCCESS;
    for (int k = 0; k < 4 && r == VKRT_SUCCESS; k++) {
        mats[k] = VKRT_materialDefault();
        memcpy(mats[k].baseColor, cols[k], sizeof(cols[k]));
        mats[k].roughness = rough[k];
        mats[k].metallic = metal[k];
        r = VKRT_addMaterial(vkrt, &mats[k], NULL, &matIndex[k]);
    }
    pcg32 g;
    pcgSeed(&g, seed ? seed : 0x5EED0002u);
    uint32_t side = (uint32_t)ceilf(cbrtf((float)instanceCount));
    uint32_t first = vkrt->meshCount;
    /* upload in batches: every copy dedups onto the first one */
    const uint32_t batch = 256;
    VKRT_MeshUpload ups[256];
    for (uint32_t k = 0; k < batch; k++) { ups[k].vertices = src->vertices; ups[k].vertexCount = src->vertexCount; ups[k].indices = src->indices; ups[k].indexCount = src->indexCount; }
    for (uint32_t done = 0; done < instanceCount && r == VKRT_SUCCESS; done += batch) r = VKRT_uploadMeshDataBatch(vkrt, ups, instanceCount - done < batch ? instanceCount - done : batch);
    for (uint32_t n = 0; n < instanceCount && r == VKRT_SUCCESS; n++) {
        uint32_t ix = n % side, iy = (n / side) % side, iz = n / (side * side);
        float pos[3] = {((float)ix + 0.5f) / (float)side * 2.0f - 1.0f + (pcgFloat(&g) - 0.5f) * 0.4f / (float)side,
                        ((float)iy + 0.5f) / (float)side * 2.0f - 1.0f + (pcgFloat(&g) - 0.5f) * 0.4f / (float)side,
                        ((float)iz + 0.5f) / (float)side * 2.0f - 1.0f + (pcgFloat(&g) - 0.5f) * 0.4f / (float)side};
        float rot[3] = {360.0f * pcgFloat(&g) - 180.0f, 360.0f * pcgFloat(&g) - 180.0f, 360.0f * pcgFloat(&g) - 180.0f};
        float s = (0.3f + 0.3f * pcgFloat(&g)) / (float)side;
        float scale[3] = {s, s, s};
        if ((r = VKRT_setMeshTransform(vkrt, first + n, pos, rot, scale)) != VKRT_SUCCESS) break;
        r = VKRT_setMeshMaterialIndex(vkrt, first + n, matIndex[n % 4]);
    }
    gltfImportFree(&imp);
    if (r != VKRT_SUCCESS) return r;
    /* floor */
    {
        Vertex* v = NULL;
        if (posix_memalign((void**)&v, 16, 4 * sizeof(Vertex)) != 0) return VKRT_ERROR_OUT_OF_MEMORY;
        memset(v, 0, 4 * sizeof(Vertex));
        const float p[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (int i = 0; i < 4; i++) {
            v[i].position[0] = p[i][0] * 3.0f; v[i].position[1] = p[i][1] * 3.0f; v[i].position[2] = -1.2f; v[i].position[3] = 1.0f;
            v[i].normal[2] = 1.0f; v[i].tangent[0] = 1.0f; v[i].tangent[3] = 1.0f;
            v[i].color[0] = v[i].color[1] = v[i].color[2] = v[i].color[3] = 1.0f;
        }
        const uint32_t idx[6] = {0, 1, 2, 0, 2, 3};
        r = VKRT_uploadMeshData(vkrt, v, 4, idx, 6);
        free(v);
        if (r != VKRT_SUCCESS) return r;
    }
    if ((r = addQuadLight(vkrt, 0.6f, 1.6f, 20.0f)) != VKRT_SUCCESS) return r;
    float pos[3] = {2.2f, -3.0f, 1.4f}, tgt[3] = {0.0f, 0.0f, -0.1f}, up[3] = {0, 0, 1};
    return VKRT_cameraSetPose(vkrt, pos, tgt, up, 40.0f);
}
