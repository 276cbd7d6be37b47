/*
 * Host-side scene model: implementation. See vengine.hpp for the reference files each part mirrors.
 */
#include "vengine.hpp"

#include <dlfcn.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace vengine {

uint64_t nextResourceUid() {
    static std::atomic<uint64_t> counter{0};
    return ++counter;
}

uint32_t Entity::s_nextId = 1;

/* ====================================================================== textures */
static ImageU8 solidImage(uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
    ImageU8 im;
    im.width = im.height = 1;
    im.channels = 4;
    im.data = {r, g, b, a};
    return im;
}

Textures::Textures() {
    /* VulkanTextures::createBaseTextures (VulkanTextures.cpp:71-76): slots 0, 1, 2; core/Image.hpp:56-85 */
    createTexture("white", solidImage(255, 255, 255, 255), ColorSpace::LINEAR);
    createTexture("whiteColor", solidImage(255, 255, 255, 255), ColorSpace::sRGB);
    createTexture("normalmapdefault", solidImage(0x80, 0x80, 0xFF, 0xFF), ColorSpace::LINEAR);
}

Texture *Textures::createTexture(const std::string &name, const ImageU8 &image, ColorSpace colorSpace) {
    if (m_map.has(name)) return m_map.get(name);
    auto t = std::make_unique<Texture>();
    t->name = name;
    t->image = image;
    t->colorSpace = colorSpace;
    t->bindlessResourceIndex = (uint32_t)m_textures.size();
    Texture *raw = t.get();
    m_textures.push_back(std::move(t));
    return m_map.add(name, raw);
}

Texture *Textures::createTexture(const AssetInfo &info, ColorSpace colorSpace) {
    if (m_map.has(info.name)) return m_map.get(info.name);
    ImageU8 im;
    /* every 8-bit texture is loaded vertically flipped because the HDR loader leaves
     * stbi_set_flip_vertically_on_load(true) set globally (core/Image.cpp:39; SURVEY trap T11) */
    if (!loadImageU8(info.filepath, im, true)) {
        std::fprintf(stderr, "Textures::createTexture(): unable to load %s\n", info.filepath.c_str());
        return nullptr;
    }
    return createTexture(info.name, im, colorSpace);
}

/* ====================================================================== materials */
ptc_material &Material::block() { return m_materials.blocks()[m_index]; }
const ptc_material &Material::block() const { return m_materials.blocks()[m_index]; }

static bool isBlack3(vec3 c, float eps) { return std::fabs(c.x) <= eps && std::fabs(c.y) <= eps && std::fabs(c.z) <= eps; }

MaterialPBRStandard::MaterialPBRStandard(const AssetInfo &info, Materials &materials, MaterialIndex index) : Material(info, materials, index) {
    ptc_material &b = block();
    std::memset(&b, 0, sizeof(b));
    b.uv_tiling[2] = (float)MaterialType::MATERIAL_PBR_STANDARD;
    albedo() = vec4(1, 1, 1, 1);
    metallic() = 0;
    roughness() = 1;
    ao() = 1;
    emissive() = vec4(0, 0, 0, 1);
    uTiling() = 1;
    vTiling() = 1;
    Textures &tx = materials.textures();
    setAlbedoTexture(tx.get("whiteColor"));
    setMetallicTexture(tx.get("white"));
    setRoughnessTexture(tx.get("white"));
    setAOTexture(tx.get("white"));
    setEmissiveTexture(tx.get("whiteColor"));
    setNormalTexture(tx.get("normalmapdefault"));
    setAlphaTexture(tx.get("white"));
}
bool MaterialPBRStandard::isEmissive() const {
    return (block().emissive[3] > std::numeric_limits<float>::epsilon()) && !isBlack3(emissiveColor(), 0.01f);
}

MaterialLambert::MaterialLambert(const AssetInfo &info, Materials &materials, MaterialIndex index) : Material(info, materials, index) {
    ptc_material &b = block();
    std::memset(&b, 0, sizeof(b));
    b.uv_tiling[2] = (float)MaterialType::MATERIAL_LAMBERT;
    albedo() = vec4(1, 1, 1, 1);
    ao() = 1;
    emissive() = vec4(0, 0, 0, 1);
    uTiling() = 1;
    vTiling() = 1;
    Textures &tx = materials.textures();
    setAlbedoTexture(tx.get("whiteColor"));
    setAOTexture(tx.get("white"));
    setEmissiveTexture(tx.get("whiteColor"));
    setNormalTexture(tx.get("normalmapdefault"));
    setAlphaTexture(tx.get("white"));
}
bool MaterialLambert::isEmissive() const {
    return (block().emissive[3] > std::numeric_limits<float>::epsilon()) && !isBlack3(emissiveColor(), 0.01f);
}

MaterialVolume::MaterialVolume(const AssetInfo &info, Materials &materials, MaterialIndex index) : Material(info, materials, index) {
    ptc_material &b = block();
    std::memset(&b, 0, sizeof(b));
    b.uv_tiling[2] = (float)MaterialType::MATERIAL_VOLUME;
    sigmaS() = vec4(0.2f);
    sigmaA() = vec4(0.01f);
    g() = 0.0f;
}

Material *Materials::createMaterial(const AssetInfo &info, MaterialType type) {
    switch (type) {
        case MaterialType::MATERIAL_PBR_STANDARD: return createMaterial<MaterialPBRStandard>(info);
        case MaterialType::MATERIAL_LAMBERT: return createMaterial<MaterialLambert>(info);
        case MaterialType::MATERIAL_VOLUME: return createMaterial<MaterialVolume>(info);
        default: return nullptr;
    }
}

/* vulkan/resources/VulkanMaterials.cpp:154-233 */
std::vector<Material *> Materials::createImportedMaterials(const std::vector<ImportedMaterial> &imported) {
    auto getTexture = [&](const ImportedTexture &tex) -> Texture * {
        if (tex.image) {
            Texture *t = m_textures.createTexture(tex.name, *tex.image, tex.colorSpace);
            if (t) {
                t->embedded = tex.embedded;
                t->filepath = tex.filepath;
            }
            return t;
        }
        return m_textures.get(tex.name); /* a texture some model embedded earlier, referenced by name */
    };
    std::vector<Material *> out;
    for (const ImportedMaterial &mat : imported) {
        if (mat.type == ImportedMaterialType::EMBEDDED) {
            out.push_back(nullptr);
        } else if (mat.type == ImportedMaterialType::LAMBERT) {
            auto *m = createMaterial<MaterialLambert>(mat.info);
            out.push_back(m);
            if (!m) continue;
            m->albedo() = mat.albedo;
            m->ao() = mat.ao;
            m->emissive() = vec4(mat.emissiveColor, mat.emissiveStrength);
            Texture *t;
            if (mat.albedoTexture && (t = getTexture(*mat.albedoTexture))) m->setAlbedoTexture(t);
            if (mat.aoTexture && (t = getTexture(*mat.aoTexture))) m->setAOTexture(t);
            if (mat.emissiveTexture && (t = getTexture(*mat.emissiveTexture))) m->setEmissiveTexture(t);
            if (mat.normalTexture && (t = getTexture(*mat.normalTexture))) m->setNormalTexture(t);
            if (mat.alphaTexture && (t = getTexture(*mat.alphaTexture))) m->setAlphaTexture(t);
            m->setTransparent(mat.transparent);
            m->uTiling() = mat.scale.x;
            m->vTiling() = mat.scale.y;
        } else if (mat.type == ImportedMaterialType::PBR_STANDARD) {
            auto *m = createMaterial<MaterialPBRStandard>(mat.info);
            out.push_back(m);
            if (!m) continue;
            m->albedo() = mat.albedo;
            m->roughness() = mat.roughness;
            m->metallic() = mat.metallic;
            m->ao() = mat.ao;
            m->emissive() = vec4(mat.emissiveColor, mat.emissiveStrength);
            Texture *t;
            if (mat.albedoTexture && (t = getTexture(*mat.albedoTexture))) m->setAlbedoTexture(t);
            if (mat.roughnessTexture && (t = getTexture(*mat.roughnessTexture))) m->setRoughnessTexture(t);
            if (mat.metallicTexture && (t = getTexture(*mat.metallicTexture))) m->setMetallicTexture(t);
            if (mat.aoTexture && (t = getTexture(*mat.aoTexture))) m->setAOTexture(t);
            if (mat.emissiveTexture && (t = getTexture(*mat.emissiveTexture))) m->setEmissiveTexture(t);
            if (mat.normalTexture && (t = getTexture(*mat.normalTexture))) m->setNormalTexture(t);
            if (mat.alphaTexture && (t = getTexture(*mat.alphaTexture))) m->setAlphaTexture(t);
            m->setTransparent(mat.transparent);
            m->uTiling() = mat.scale.x;
            m->vTiling() = mat.scale.y;
        } else {
            auto *m = createMaterial<MaterialVolume>(mat.info);
            out.push_back(m);
            if (!m) continue;
            m->sigmaS() = vec4(mat.sigmaS, 1.0f);
            m->sigmaA() = vec4(mat.sigmaA, 1.0f);
            m->g() = mat.g;
        }
    }
    return out;
}

/* ====================================================================== transform */
void Transform::setRotation(vec3 forward, vec3 up) {
    vec3 newZ = vm::normalize(-forward);
    vec3 newY = vm::normalize(up);
    vec3 newX = vm::normalize(vm::cross(newY, newZ));
    newY = vm::normalize(vm::cross(newZ, newX));
    mat4 r(1.0f);
    r[0][0] = newX.x; r[0][1] = newX.y; r[0][2] = newX.z;
    r[1][0] = newY.x; r[1][1] = newY.y; r[1][2] = newY.z;
    r[2][0] = newZ.x; r[2][1] = newZ.y; r[2][2] = newZ.z;
    setRotation(vm::quat_cast(r));
}

/* ====================================================================== scene */
SceneObject *Scene::addSceneObject(const std::string &name, SceneObject *parent, Transform transform) {
    m_objects.push_back(std::make_unique<SceneObject>(name));
    SceneObject *o = m_objects.back().get();
    o->setLocalTransform(transform);
    if (parent == nullptr)
        m_sceneGraph.push_back(o);
    else
        parent->addChild(o);
    return o;
}

void Scene::clear() {
    m_sceneGraph.clear();
    m_objects.clear();
    m_instances.invalidate();
}

void Scene::update() {
    for (SceneObject *root : m_sceneGraph) root->update(nullptr);
    m_instances.invalidate();
    m_instances.build();
}

static void flatRec(SceneObject *o, SceneObjectVector &out) {
    out.push_back(o);
    for (SceneObject *c : o->children()) flatRec(c, out);
}
SceneObjectVector Scene::getSceneObjectsFlat() const {
    SceneObjectVector v;
    for (SceneObject *r : m_sceneGraph) flatRec(r, v);
    return v;
}

/* VulkanScene.cpp:115-142: lights live in the asset manager's lights map; an existing name returns the existing light */
Light *Scene::createLight(const AssetInfo &info, LightType type, vec4 color) {
    if (type != LightType::POINT_LIGHT && type != LightType::DIRECTIONAL_LIGHT) return nullptr;
    AssetMap<Light> &lights = m_engine.lightsMap();
    if (lights.has(info.name)) return lights.get(info.name);
    ptc_light_data ld{};
    ld.color[0] = color.x; ld.color[1] = color.y; ld.color[2] = color.z; ld.color[3] = color.w;
    ld.type[0] = (uint32_t)type;
    if (m_lightData.size() >= 1024) return nullptr;
    m_lightData.reserve(1024); /* Light holds a reference into the table */
    m_lightData.push_back(ld);
    m_lights.push_back(std::make_unique<Light>(info, type, m_lightData, (LightIndex)(m_lightData.size() - 1)));
    return lights.add(info.name, m_lights.back().get());
}

ptc_scene_data Scene::getSceneData() const {
    ptc_scene_data sd{};
    mat4 view = m_camera->viewMatrix();
    mat4 viewInv = m_camera->viewMatrixInverse();
    std::memcpy(sd.view, view.data(), 64);
    std::memcpy(sd.view_inverse, viewInv.data(), 64);
    sd.exposure[0] = m_exposure;
    sd.exposure[1] = m_environmentIntensity;
    sd.exposure[2] = m_camera->lensRadius();
    sd.exposure[3] = m_camera->focalDistance();
    sd.background[0] = m_backgroundColor.x;
    sd.background[1] = m_backgroundColor.y;
    sd.background[2] = m_backgroundColor.z;
    sd.background[3] = (float)m_environmentType;
    sd.volumes[0] = -1;
    sd.volumes[1] = m_camera->znear();
    sd.volumes[2] = m_camera->zfar();
    sd.volumes[3] = 0;
    return sd;
}

/* ====================================================================== instances (core/Instances.cpp) */
void InstancesManager::invalidate() {
    m_instancesOpaque.clear();
    m_transparent.clear();
    m_lights.clear();
    m_meshLights.clear();
    m_volumes.clear();
    m_order.clear();
}

void InstancesManager::build() {
    /* fillSceneObjectVectors, Instances.cpp:100-143 */
    for (SceneObject *so : m_scene->getSceneObjectsFlat()) {
        if (!so->isActive()) continue;
        if (so->has<ComponentMesh>()) {
            Mesh *mesh = so->get<ComponentMesh>().mesh();
            if (so->has<ComponentMaterial>() && mesh != nullptr) {
                Material *material = so->get<ComponentMaterial>().material();
                if (material == nullptr) continue;
                if (material->isEmissive()) m_meshLights.push_back(so);
                if (!material->isTransparent()) {
                    bool found = false;
                    for (auto &g : m_instancesOpaque)
                        if (g.first == mesh) {
                            g.second.push_back(so);
                            found = true;
                            break;
                        }
                    if (!found) m_instancesOpaque.push_back({mesh, SceneObjectVector{so}});
                } else {
                    m_transparent.push_back(so);
                }
            }
        }
        if (so->has<ComponentLight>()) {
            if (so->get<ComponentLight>().light() != nullptr) m_lights.push_back(so);
        }
        if (so->has<ComponentVolume>()) {
            ComponentVolume &v = so->get<ComponentVolume>();
            if (v.frontFacing() != nullptr || v.backFacing() != nullptr) m_volumes.push_back(so);
        }
    }
    /* buildInstanceDataFromScratch, Instances.cpp:145-181: opaque groups, transparent, lights */
    for (auto &g : m_instancesOpaque)
        for (SceneObject *so : g.second) m_order.push_back(so);
    for (SceneObject *so : m_transparent) m_order.push_back(so);
    for (SceneObject *so : m_lights) m_order.push_back(so);
}

/* ====================================================================== engine */
static std::string dirOf(const std::string &p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
}
static std::string selfDir() {
    Dl_info info;
    if (dladdr((void *)&selfDir, &info) && info.dli_fname) return dirOf(info.dli_fname);
    return ".";
}

Engine::Engine(const std::string &name, const std::string &assetRoot) : m_name(name), m_assetRoot(assetRoot) {
    if (m_assetRoot.empty()) {
        const char *env = std::getenv("VVIEWER_ASSETS");
        m_assetRoot = env ? env : (selfDir() + "/../..");
    }
    m_textures = std::make_unique<Textures>();
    m_materials = std::make_unique<Materials>(*m_textures);
    m_scene = std::make_unique<Scene>(*this);
    m_renderer = std::make_unique<Renderer>(std::make_unique<CudaRendererPathTracing>(*this));
}
Engine::~Engine() {}

std::string Engine::assetPath(const std::string &rel) const {
    if (!rel.empty() && rel[0] == '/') return rel;
    return m_assetRoot + "/" + rel;
}

void Engine::initResources() {
    /* initDefaultData, VulkanEngine.cpp:331-391 */
    auto *defaultMaterial = m_materials->createMaterial<MaterialPBRStandard>(AssetInfo("defaultMaterial", AssetSource::ENGINE));
    defaultMaterial->albedo() = vec4(0.8f, 0.8f, 0.8f, 1);
    defaultMaterial->metallic() = 0.5f;
    defaultMaterial->roughness() = 0.5f;
    defaultMaterial->ao() = 1.0f;
    defaultMaterial->emissive() = vec4(0, 0, 0, 1);
    auto *defaultEmissive = m_materials->createMaterial<MaterialPBRStandard>(AssetInfo("defaultEmissive", AssetSource::ENGINE));
    defaultEmissive->albedo() = vec4(1, 1, 1, 1);
    defaultEmissive->emissive() = vec4(1, 1, 1, 1);
    auto *defaultVolume = m_materials->createMaterial<MaterialVolume>(AssetInfo("defaultVolume", AssetSource::ENGINE));
    defaultVolume->sigmaS() = vec4(0.01f, 0.01f, 0.01f, 1);
    defaultVolume->sigmaA() = vec4(0.01f, 0.01f, 0.01f, 1);
    defaultVolume->g() = 0.0f;

    m_lightsMap.add("defaultPointLight", m_scene->createLight(AssetInfo("defaultPointLight", AssetSource::ENGINE), LightType::POINT_LIGHT, vec4(1, 1, 1, 1)));
    m_lightsMap.add("defaultDirectionalLightSun",
                    m_scene->createLight(AssetInfo("defaultDirectionalLightSun", AssetSource::ENGINE), LightType::DIRECTIONAL_LIGHT, vec4(1, 0.9f, 0.8f, 1)));
    m_lightsMap.add("defaultDirectionalLightMoon",
                    m_scene->createLight(AssetInfo("defaultDirectionalLightMoon", AssetSource::ENGINE), LightType::DIRECTIONAL_LIGHT, vec4(0.31f, 0.4f, 0.52f, 1)));

    importModel(AssetInfo("assets/models/uvsphere.obj", AssetSource::ENGINE), false);
    importModel(AssetInfo("assets/models/plane.obj", AssetSource::ENGINE), false);
    importModel(AssetInfo("assets/models/cube.obj", AssetSource::ENGINE), false);

    m_scene->skyboxMaterial() = importEnvironmentMap(AssetInfo("assets/HDR/harbor.hdr", AssetSource::ENGINE));
}

Model3D *Engine::addModel(std::unique_ptr<Model3D> model) {
    Model3D *raw = model.get();
    for (auto &m : raw->meshes) {
        m->poolIndex = (uint32_t)m_meshPool.size();
        m->model = raw;
        m_meshPool.push_back(m.get());
    }
    m_ownedModels.push_back(std::move(model));
    m_models.add(raw->name, raw);
    return raw;
}

/* VulkanEngine::importModel (VulkanEngine.cpp:157-192) + VulkanModel3D::importNode (VulkanModel3D.cpp:47-76) */
Model3D *Engine::importModel(const AssetInfo &info, bool importMaterials) {
    if (m_models.has(info.name)) return m_models.get(info.name);
    const std::string path = assetPath(info.filepath);
    std::string ext;
    size_t dot = path.find_last_of('.');
    if (dot != std::string::npos) ext = path.substr(dot + 1);
    for (char &c : ext) c = (char)std::tolower((unsigned char)c);

    ImportedModelNode root;
    std::vector<ImportedMaterial> importedMaterials;
    std::string err;
    bool ok;
    if (ext == "obj")
        ok = loadOBJ(path, root, importMaterials ? &importedMaterials : nullptr, &err);
    else if (ext == "gltf" || ext == "glb")
        ok = loadGLTF(path, root, importMaterials ? &importedMaterials : nullptr, &err);
    else {
        ok = false;
        err = "unsupported model format ." + ext + " (" + path + ")"; /* the reference also lists .fbx, served by assimp */
    }
    if (!ok) {
        std::fprintf(stderr, "Engine::importModel(): Failed to import a model: %s\n", err.c_str());
        return nullptr;
    }
    std::vector<Material *> materials;
    if (importMaterials) materials = m_materials->createImportedMaterials(importedMaterials);

    auto model = std::make_unique<Model3D>();
    model->name = info.name;
    model->filepath = path;
    model->internal = info.source == AssetSource::ENGINE;
    /* importNode() appends the children of EVERY node to the model's root (`m_nodeTree.add()`, VulkanModel3D.cpp:73):
     * the tree is flattened to root + all descendants, each keeping its local transform (SURVEY trap T13) */
    std::vector<ImportedModelNode *> descendants; /* pre-order, the order importNode() appends them in */
    std::function<void(ImportedModelNode &)> collect = [&](ImportedModelNode &n) {
        for (ImportedModelNode &c : n.children) {
            descendants.push_back(&c);
            collect(c);
        }
    };
    collect(root);
    {
        Model3D::Model3DNode &dst = model->nodeTree;
        dst.name = root.name;
        dst.transform = root.transform;
        auto take = [&](ImportedModelNode &src, Model3D::Model3DNode &node) {
            node.name = src.name;
            node.transform = src.transform;
            for (size_t i = 0; i < src.meshes.size(); i++) {
                int32_t mi = src.materialIndices[i];
                node.meshes.push_back(src.meshes[i].get());
                node.materials.push_back(mi >= 0 && (size_t)mi < materials.size() ? materials[(size_t)mi] : nullptr);
                model->meshes.push_back(std::move(src.meshes[i]));
            }
        };
        take(root, dst);
        dst.children.resize(descendants.size());
        for (size_t i = 0; i < descendants.size(); i++) take(*descendants[i], dst.children[i]);
    }
    return addModel(std::move(model));
}

EnvironmentMap *Engine::importEnvironmentMap(const AssetInfo &info) {
    for (auto &e : m_envMaps)
        if (e->name == info.name) return e.get();
    auto env = std::make_unique<EnvironmentMap>();
    env->name = info.name;
    env->filepath = assetPath(info.filepath);
    if (!loadImageHDR(assetPath(info.filepath), env->equirect, true)) {
        std::fprintf(stderr, "Engine::importEnvironmentMap(): unable to load %s\n", info.filepath.c_str());
        return nullptr;
    }
    m_envMaps.push_back(std::move(env));
    return m_envMaps.back().get();
}

void Engine::flatten(FlatScene &out) {
    out = FlatScene();
    /* geometry pools: only meshes that are instanced, in the order the instances first name them */
    std::unordered_map<Mesh *, uint32_t> meshSlot;
    std::vector<Mesh *> slotMesh;
    InstancesManager &im = m_scene->instancesManager();
    auto slotOf = [&](Mesh *mesh) -> uint32_t {
        auto it = meshSlot.find(mesh);
        if (it != meshSlot.end()) return it->second;
        const uint32_t s = (uint32_t)slotMesh.size();
        slotMesh.push_back(mesh);
        meshSlot[mesh] = s;
        return s;
    };
    std::unordered_map<SceneObject *, uint32_t> instanceSlot;
    /* initInstanceData, Instances.cpp:76-98 + VulkanInstances.cpp:117-130. Light-only objects own an
     * InstanceData slot in the reference as well but carry no geometry; they are skipped here because
     * nothing on the path-tracing path reads them. */
    for (SceneObject *so : im.instanceOrder()) {
        if (!so->has<ComponentMesh>() || !so->has<ComponentMaterial>()) continue;
        Mesh *mesh = so->get<ComponentMesh>().mesh();
        Material *mat = so->get<ComponentMaterial>().material();
        if (!mesh || !mat) continue;
        ptc_instance inst{};
        std::memcpy(inst.model, so->modelMatrix().data(), 64);
        inst.id[0] = (float)so->getID();
        inst.id[1] = -1;
        inst.id[2] = -1;
        inst.id[3] = 0;
        inst.material_index = mat->materialIndex();
        if (so->has<ComponentVolume>()) {
            ComponentVolume &v = so->get<ComponentVolume>();
            if (v.frontFacing()) inst.id[1] = (float)v.frontFacing()->materialIndex();
            if (v.backFacing()) inst.id[2] = (float)v.backFacing()->materialIndex();
        }
        inst.mesh_index = slotOf(mesh);
        inst.num_triangles = mesh->nTriangles();
        instanceSlot[so] = (uint32_t)out.instances.size();
        out.instances.push_back(inst);
    }
    {
        std::vector<uint64_t> key;
        key.reserve(slotMesh.size() * 3);
        for (Mesh *mesh : slotMesh) {
            key.push_back(mesh->uid);
            key.push_back((uint64_t)mesh->vertices.size());
            key.push_back((uint64_t)mesh->indices.size());
        }
        if (!m_geometry || m_geometry->key != key) {
            auto pools = std::make_shared<GeometryPools>();
            size_t nv = 0, ni = 0;
            for (Mesh *mesh : slotMesh) {
                nv += mesh->vertices.size();
                ni += (size_t)mesh->nTriangles() * 3;
            }
            pools->vertices.reserve(nv);
            pools->indices.reserve(ni);
            for (Mesh *mesh : slotMesh) {
                ptc_mesh pm{};
                pm.first_index = (uint32_t)pools->indices.size();
                pm.tri_count = mesh->nTriangles();
                pm.first_vertex = (uint32_t)pools->vertices.size();
                pm.vertex_count = (uint32_t)mesh->vertices.size();
                pools->vertices.insert(pools->vertices.end(), mesh->vertices.begin(), mesh->vertices.end());
                pools->indices.insert(pools->indices.end(), mesh->indices.begin(), mesh->indices.begin() + (size_t)pm.tri_count * 3);
                pools->meshes.push_back(pm);
            }
            pools->key = std::move(key);
            m_geometry = std::move(pools);
        }
        out.geometry = m_geometry;
    }
    out.materials = m_materials->blocks();
    out.lightData = m_scene->lightData();
    /* VulkanInstancesManager::build, VulkanInstances.cpp:66-109 */
    for (SceneObject *so : im.lights()) {
        Light *light = so->get<ComponentLight>().light();
        ptc_light_instance li{};
        li.info[0] = light->lightIndex();
        li.info[1] = 0;
        li.info[3] = (uint32_t)light->type();
        if (light->type() == LightType::POINT_LIGHT) {
            vec3 p = so->worldPosition();
            li.position[0] = p.x; li.position[1] = p.y; li.position[2] = p.z;
        } else if (light->type() == LightType::DIRECTIONAL_LIGHT) {
            vec4 d = so->modelMatrix() * vec4(0, 0, 1, 0);
            li.position[0] = d.x; li.position[1] = d.y; li.position[2] = d.z;
        }
        li.position[3] = so->get<ComponentLight>().castShadows() ? 1.0f : 0.0f;
        out.lightInstances.push_back(li);
    }
    for (SceneObject *so : im.meshLights()) {
        auto it = instanceSlot.find(so);
        if (it == instanceSlot.end()) continue;
        ptc_light_instance li{};
        li.info[1] = it->second;
        li.info[3] = (uint32_t)LightType::MESH_LIGHT;
        const mat4 &m = so->modelMatrix();
        for (int c = 0; c < 4; c++) {
            li.position[c] = m[c][0];
            li.position1[c] = m[c][1];
            li.position2[c] = m[c][2];
        }
        out.lightInstances.push_back(li);
    }
    for (auto &t : m_textures->all()) {
        ptc_texture pt{};
        pt.width = (uint32_t)t->image.width;
        pt.height = (uint32_t)t->image.height;
        pt.channels = (uint32_t)t->image.channels;
        pt.srgb = t->colorSpace == ColorSpace::sRGB ? 1u : 0u;
        pt.data = t->image.data.data();
        pt.uid = t->uid;
        out.textures.push_back(pt);
    }
    ptc_scene_desc &d = out.desc;
    d.vertices = out.geometry->vertices.data();
    d.n_vertices = out.geometry->vertices.size();
    d.indices = out.geometry->indices.data();
    d.n_indices = out.geometry->indices.size();
    d.meshes = out.geometry->meshes.data();
    d.n_meshes = (uint32_t)out.geometry->meshes.size();
    d.instances = out.instances.data();
    d.n_instances = (uint32_t)out.instances.size();
    d.materials = out.materials.data();
    d.n_materials = (uint32_t)out.materials.size();
    d.light_data = out.lightData.data();
    d.n_light_data = (uint32_t)out.lightData.size();
    d.light_instances = out.lightInstances.data();
    d.n_light_instances = (uint32_t)out.lightInstances.size();
    d.textures = out.textures.data();
    d.n_textures = (uint32_t)out.textures.size();
    EnvironmentMap *env = m_scene->skyboxMaterial();
    if (env && !env->equirect.data.empty()) {
        d.env.equirect_rgba = env->equirect.data.data();
        d.env.width = (uint32_t)env->equirect.width;
        d.env.height = (uint32_t)env->equirect.height;
        d.env.uid = env->uid;
    }
}

/* ====================================================================== renderer */
bool PtcBackend::load(const std::string &libPath, std::string *err) {
    path = libPath;
    handle = dlopen(libPath.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!handle) {
        if (err) *err = std::string("cannot load path-tracing backend '") + libPath + "': " + dlerror();
        return false;
    }
#define PTC_SYM(field, sym)                                                      \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, #sym));              \
    if (!field) {                                                                \
        if (err) *err = std::string("backend is missing symbol ") + #sym;        \
        return false;                                                            \
    }
    PTC_SYM(create, ptc_create)
    PTC_SYM(destroy, ptc_destroy)
    PTC_SYM(last_error, ptc_last_error)
    PTC_SYM(backend_name, ptc_backend_name)
    PTC_SYM(upload_scene, ptc_upload_scene)
    PTC_SYM(build_accel, ptc_build_accel)
    PTC_SYM(render, ptc_render)
    PTC_SYM(progress, ptc_progress)
    PTC_SYM(get_stats, ptc_get_stats)
    PTC_SYM(device_count, ptc_device_count)
    PTC_SYM(comm_unique_id, ptc_comm_unique_id)
    PTC_SYM(comm_init_rank, ptc_comm_init_rank)
    PTC_SYM(set_sampler_tables, ptc_set_sampler_tables)
#undef PTC_SYM
    return true;
}

CudaRendererPathTracing::CudaRendererPathTracing(Engine &engine) : m_engine(engine) {
    /* like VulkanRendererPathTracing::initResources (…PathTracing.cpp:45-91) a failure leaves
     * isRayTracingEnabled() == false; render() then reports the error and returns */
    if (!m_backend.load(selfDir() + "/libptc_cuda.so", &m_error)) {
        std::fprintf(stderr, "CudaRendererPathTracing: %s\n", m_error.c_str());
        return;
    }
    if (m_backend.create(&m_ctx, nullptr, 0) != 0 || !m_ctx) {
        m_error = std::string("ptc_create failed: ") + (m_ctx ? m_backend.last_error(m_ctx) : "no context");
        std::fprintf(stderr, "CudaRendererPathTracing: %s\n", m_error.c_str());
        return;
    }
    m_isInitialized = true;
}

CudaRendererPathTracing::~CudaRendererPathTracing() {
    if (m_ctx) m_backend.destroy(m_ctx);
}

bool CudaRendererPathTracing::setDevices(const std::vector<int> &deviceIds) {
    if (!m_backend.create || m_renderInProgress) return false;
    ptc_ctx *fresh = nullptr;
    if (m_backend.create(&fresh, deviceIds.empty() ? nullptr : deviceIds.data(), (int)deviceIds.size()) != 0 || !fresh) {
        m_error = std::string("ptc_create failed: ") + (fresh ? m_backend.last_error(fresh) : "no context");
        if (fresh) m_backend.destroy(fresh);
        std::fprintf(stderr, "CudaRendererPathTracing::setDevices(): %s\n", m_error.c_str());
        return false;
    }
    if (m_ctx) m_backend.destroy(m_ctx);
    m_ctx = fresh;
    m_samplerTablesSet = false;
    m_commRank = 0;
    m_commWorld = 1;
    m_isInitialized = true;
    return true;
}

bool CudaRendererPathTracing::commUniqueId(uint8_t out128[128]) { return m_backend.comm_unique_id && m_backend.comm_unique_id(out128) == 0; }

bool CudaRendererPathTracing::commInitRank(const uint8_t id128[128], int rank, int world) {
    if (!m_isInitialized) return false;
    if (m_backend.comm_init_rank(m_ctx, id128, rank, world) != 0) {
        m_error = m_backend.last_error(m_ctx);
        std::fprintf(stderr, "CudaRendererPathTracing::commInitRank(): %s\n", m_error.c_str());
        return false;
    }
    m_commRank = rank;
    m_commWorld = world;
    return true;
}

/* VulkanRandom::initResources / createBuffers (vulkan/resources/VulkanRandom.cpp:15-72): the PMJ02BN sequences and the blue-noise
 * textures as device tables.  The numbers are the reference's (math/PMJSequences.cpp, math/BlueNoise.cpp), kept as binary files under
 * assets/tables (tools/extract_sampler_tables.py); they are only loaded when the PMJ sampler is asked for. */
bool CudaRendererPathTracing::ensureSamplerTables() {
    if (m_samplerTablesSet) return true;
    auto readAll = [](const std::string &path, std::vector<uint8_t> &out) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) return false;
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (n < 0) {
            std::fclose(f);
            return false;
        }
        out.resize((size_t)n);
        const bool ok = std::fread(out.data(), 1, out.size(), f) == out.size();
        std::fclose(f);
        return ok;
    };
    std::vector<uint8_t> pmjBytes, blueBytes;
    const size_t nPmj = (size_t)16 * 16384 * 2, nBlue = (size_t)48 * 128 * 128;
    if (!readAll(m_engine.assetPath("assets/tables/pmj02bn.f32"), pmjBytes) || pmjBytes.size() != nPmj * 4 ||
        !readAll(m_engine.assetPath("assets/tables/bluenoise.u16"), blueBytes) || blueBytes.size() != nBlue * 2) {
        m_error = "PMJ02BN sampler tables not found under assets/tables (tools/extract_sampler_tables.py writes them)";
        return false;
    }
    std::vector<float> pmj(nPmj), blue(nBlue);
    std::memcpy(pmj.data(), pmjBytes.data(), nPmj * 4);
    for (size_t i = 0; i < nBlue; i++) blue[i] = (float)((uint32_t)blueBytes[2 * i] | ((uint32_t)blueBytes[2 * i + 1] << 8)) / 65536.0f;
    if (m_backend.set_sampler_tables(m_ctx, pmj.data(), 16, 16384, blue.data(), 48, 128) != 0) {
        m_error = m_backend.last_error(m_ctx);
        return false;
    }
    m_samplerTablesSet = true;
    return true;
}

float CudaRendererPathTracing::renderProgress() { return (m_ctx && m_isInitialized) ? m_backend.progress(m_ctx) : 0.0f; }

ptc_render_params CudaRendererPathTracing::makeRenderParams() {
    /* VulkanRendererPathTracing::render, …PathTracing.cpp:143-199 */
    Scene &scene = m_engine.scene();
    auto camera = scene.camera();
    ptc_render_params rp{};
    rp.scene = scene.getSceneData();
    uint32_t width = renderInfo().width, height = renderInfo().height;
    mat4 proj(1.0f);
    if (camera->type() == CameraType::PERSPECTIVE) {
        auto pc = std::static_pointer_cast<PerspectiveCamera>(camera);
        proj = vm::perspective(vm::radians(pc->fov()), static_cast<float>(width) / height, camera->znear(), camera->zfar());
        rp.camera_type = PTC_CAMERA_PERSPECTIVE;
    } else {
        /* trap T10: the reference builds glm::ortho in pixel units and still shoots rays from the eye
         * point; we render a true orthographic view of orthoWidth x orthoWidth / aspect */
        auto oc = std::static_pointer_cast<OrthographicCamera>(camera);
        float ow = oc->orthoWidth(), oh = ow / (static_cast<float>(width) / height);
        proj = vm::ortho(-ow / 2, ow / 2, -oh / 2, oh / 2, camera->znear(), camera->zfar());
        rp.camera_type = PTC_CAMERA_ORTHOGRAPHIC;
        rp.ortho_width = ow;
        rp.ortho_height = oh;
    }
    proj[1][1] *= -1;
    mat4 projInv = vm::inverse(proj);
    std::memcpy(rp.scene.projection, proj.data(), 64);
    std::memcpy(rp.scene.projection_inverse, projInv.data(), 64);
    rp.scene.volumes[0] = -1;
    if (camera->volume() != nullptr) {
        if (camera->volume()->type() != MaterialType::MATERIAL_VOLUME)
            std::fprintf(stderr, "The material set for the render camera's volume is not a volume material\n");
        else
            rp.scene.volumes[0] = static_cast<float>(camera->volume()->materialIndex());
    }
    rp.samples = renderInfo().samples;
    rp.batch_size = renderInfo().batchSize;
    rp.depth = renderInfo().depth;
    rp.width = width;
    rp.height = height;
    /* several GPUs: the core fills in rank / world itself; the caller only chooses how the image is cut */
    rp.split_mode = renderInfo().multiGpuSplit == 1u ? PTC_SPLIT_TILE : (renderInfo().multiGpuSplit == 2u ? PTC_SPLIT_SAMPLE : PTC_SPLIT_NONE);
    rp.rank = 0;
    rp.world = 1;
    rp.flags = renderInfo().lowDiscrepancySampler ? PTC_FLAG_SAMPLER_SOBOL : 0u;
    if (renderInfo().pmjSampler) rp.flags = (rp.flags & ~PTC_FLAG_SAMPLER_SOBOL) | PTC_FLAG_SAMPLER_PMJ;
    if (renderInfo().environmentImportanceSampling) rp.flags |= PTC_FLAG_ENV_IMPORTANCE;
    return rp;
}

bool CudaRendererPathTracing::renderToMemory(std::vector<float> &radiance, std::vector<float> &albedo, std::vector<float> &normal) {
    const RenderInfo &ri = renderInfo();
    const size_t n = (size_t)ri.width * ri.height * 4;
    radiance.resize(n);
    albedo.resize(n);
    normal.resize(n);
    return renderToBuffers(radiance.data(), albedo.data(), normal.data());
}

bool CudaRendererPathTracing::renderToBuffers(float *radiance, float *albedo, float *normal) {
    if (!m_isInitialized) {
        std::fprintf(stderr, "CudaRendererPathTracing::render(): backend not initialised: %s\n", m_error.c_str());
        return false;
    }
    if (m_renderInProgress) return false;
    Scene &scene = m_engine.scene();
    if (scene.getSceneObjectsFlat().empty()) {
        std::fprintf(stderr, "Trying to render an empty scene\n");
        return false;
    }
    m_renderInProgress = true;
    bool ok = false, tablesFailed = false;
    const bool verbose = std::getenv("PTC_VERBOSE") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        auto now = std::chrono::steady_clock::now();
        if (verbose) std::fprintf(stderr, "[render] %-10s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    do {
        FlatScene flat;
        m_engine.flatten(flat);
        lap("flatten");
        ptc_render_params rp = makeRenderParams();
        if ((rp.flags & PTC_FLAG_SAMPLER_PMJ) && !ensureSamplerTables()) {
            tablesFailed = true;
            break;
        }
        if (m_backend.upload_scene(m_ctx, &flat.desc) != 0) break;
        lap("upload");
        if (m_backend.build_accel(m_ctx) != 0) break;
        lap("build");
        if (m_backend.render(m_ctx, &rp, radiance, albedo, normal) != 0) break;
        lap("render+d2h");
        m_backend.get_stats(m_ctx, &m_stats);
        ok = true;
    } while (false);
    if (!ok) {
        if (!tablesFailed) m_error = m_backend.last_error(m_ctx);
        std::fprintf(stderr, "CudaRendererPathTracing::render(): %s\n", m_error.c_str());
    }
    m_renderInProgress = false;
    return ok;
}

void CudaRendererPathTracing::render() {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<float> radiance, albedo, normal;
    if (!renderToMemory(radiance, albedo, normal)) return;
    if (m_commRank != 0) return; /* one process per GPU: the image lives on rank 0 */
    /* storeToDisk, …PathTracing.cpp:958-1027 (OIDN is out of scope: denoise == true writes the
     * un-denoised radiance and, with writeAllFiles, the _radiance AOV) */
    const RenderInfo &ri = renderInfo();
    const uint32_t channels = 4;
    if (ri.writeAllFiles) {
        writeToDisk(albedo, ri.filename + "_albedo", FileType::HDR, ri.width, ri.height, channels);
        writeToDisk(normal, ri.filename + "_normal", FileType::HDR, ri.width, ri.height, channels);
        if (ri.denoise) writeToDisk(radiance, ri.filename + "_radiance", FileType::HDR, ri.width, ri.height, channels);
    }
    if (ri.fileType == FileType::PNG && ri.exposure != 0.0f) applyExposure(radiance, ri.exposure, channels);
    writeToDisk(radiance, ri.filename, ri.fileType, ri.width, ri.height, channels);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::string ext = (ri.fileType == FileType::PNG) ? "png" : "hdr";
    std::printf("Scene rendered: %s.%s in: %dms\n", ri.filename.c_str(), ext.c_str(), (int)ms);
}

}  // namespace vengine
