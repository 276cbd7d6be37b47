/*
 * vengine host-side scene model for the B200 path tracer — the feeder of the hot path.
 *
 * This mirrors, with no Vulkan types, the subset of the reference's abstract API that
 * src/bin/offlinerender and src/bin/unittests/RenderTests.cpp use to describe a scene and to call
 * RendererPathTracing::render():
 *   Engine / Renderer / RendererPathTracing   src/lib/vengine/core/Engine.hpp:15-54, Renderer.hpp:11-60
 *   Scene, SceneObject, SceneNode             core/Scene.hpp:39-119, core/SceneObject.hpp, core/SceneNode.hpp:57-78
 *   Entity / Component*                       utils/ECS.hpp:43-109, 289-430
 *   Camera                                    core/Camera.hpp, Camera.cpp
 *   Material*, Texture, Light, Mesh, Model3D  core/Material.hpp, Light.hpp, Mesh.hpp, Model3D.hpp
 *   InstancesManager                          core/Instances.cpp:23-181, vulkan/VulkanInstances.cpp:66-130
 * Same class and method names, same defaults, same flattening rules.  Everything below the
 * RendererPathTracing boundary goes through the C-ABI in include/ptc.h.
 */
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <cstdlib>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <vector>

#include "vmath.hpp"
#include "../../include/ptc.h"

namespace glm = vengine::vm; /* scene recipes read like the reference's */

namespace vengine {

using vm::mat4;
using vm::quat;
using vm::vec2;
using vm::vec3;
using vm::vec4;

/* ------------------------------------------------------------------ assets */
enum class AssetSource { DISK = 0, ENGINE = 1 };
struct AssetInfo { /* core/Asset.hpp */
    std::string name, filepath;
    AssetSource source = AssetSource::DISK;
    bool embedded = false; /* AssetLocation::DISK_EMBEDDED: lives inside a model file */
    AssetInfo() {}
    AssetInfo(const char *n) : name(n), filepath(n) {}
    AssetInfo(const std::string &n) : name(n), filepath(n) {}
    AssetInfo(const std::string &n, const std::string &f) : name(n), filepath(f) {}
    AssetInfo(const std::string &n, AssetSource s) : name(n), filepath(n), source(s) {}
};

template <typename T>
class AssetMap { /* core/AssetManager.hpp:84-114 */
public:
    bool has(const std::string &n) const { return m_map.count(n) != 0; }
    T *get(const std::string &n) const {
        auto it = m_map.find(n);
        return it == m_map.end() ? nullptr : it->second;
    }
    T *add(const std::string &n, T *t) {
        m_map[n] = t;
        return t;
    }
    void remove(const std::string &n) { m_map.erase(n); }
    const std::map<std::string, T *> &all() const { return m_map; }
    void reset() { m_map.clear(); }

private:
    std::map<std::string, T *> m_map;
};

enum class ColorSpace { sRGB = 0, LINEAR = 1 };
enum class FileType { PNG = 0, HDR = 1 };

/* ------------------------------------------------------------------ images / textures */
struct ImageU8 {
    int width = 0, height = 0, channels = 0; /* rows in memory order; loaders flip vertically (trap T11) */
    std::vector<uint8_t> data;
};
struct ImageF32 {
    int width = 0, height = 0, channels = 0;
    std::vector<float> data;
};
/* io_image.cpp */
/* PNG (8/16-bit, non-interlaced) or JPEG (baseline / progressive). 1-channel files stay 1 channel, everything else becomes
 * RGBA (Image<stbi_uc>::loadDiskImage); srcChannels = the channel count of the file as stbi reports it */
bool loadImageU8(const std::string &path, ImageU8 &out, bool flipVertically, int *srcChannels = nullptr);
bool decodeImageU8(const uint8_t *bytes, size_t nBytes, ImageU8 &out, int *srcChannels, bool flipVertically);
bool decodeJPEG(const uint8_t *bytes, size_t nBytes, ImageU8 &outRGBA, int *srcChannels); /* io_jpeg.cpp; rows top first */
bool loadImageHDR(const std::string &path, ImageF32 &out, bool flipVertically);   /* Radiance RGBE -> RGBA32F */
bool writeImageHDR(const std::string &path, int w, int h, int channels, const float *data); /* like stbi_write_hdr */
bool writeImagePNG(const std::string &path, int w, int h, int channels, const uint8_t *data);
/* core/ImageUtils.cpp:34-76 */
float linearToSRGB(float v);
void applyExposure(std::vector<float> &in, float exposure, uint32_t channels);
void writeToDisk(const std::vector<float> &in, const std::string &filename, FileType type, uint32_t w, uint32_t h, uint32_t channels);

uint64_t nextResourceUid(); /* process-wide, never 0 */

class Texture {
public:
    std::string name;
    std::string filepath;  /* disk textures; empty for engine-made and embedded ones */
    bool embedded = false; /* decoded out of a model file (AssetLocation::DISK_EMBEDDED) */
    ImageU8 image;
    ColorSpace colorSpace = ColorSpace::sRGB;
    uint32_t bindlessResourceIndex = 0; /* slot in the texture table (VulkanTextures.cpp:145) */
    uint64_t uid = nextResourceUid();   /* textures are immutable once created: identity for ptc_texture.uid */
};

class Textures { /* vulkan/resources/VulkanTextures.cpp:71-145 */
public:
    Textures();
    Texture *createTexture(const AssetInfo &info, ColorSpace colorSpace = ColorSpace::sRGB);
    Texture *createTexture(const std::string &name, const ImageU8 &image, ColorSpace colorSpace);
    Texture *get(const std::string &name) const { return m_map.get(name); }
    Texture *bySlot(uint32_t slot) const { return slot < m_textures.size() ? m_textures[slot].get() : nullptr; }
    const std::vector<std::unique_ptr<Texture>> &all() const { return m_textures; }

private:
    std::vector<std::unique_ptr<Texture>> m_textures;
    AssetMap<Texture> m_map;
};

class EnvironmentMap {
public:
    std::string name, filepath;
    ImageF32 equirect; /* RGBA32F, flipped like Image<float>::loadDiskImage (core/Image.cpp:27-43) */
    uint64_t uid = nextResourceUid();
};

/* ------------------------------------------------------------------ materials */
enum class MaterialType { MATERIAL_NOT_SET = -1, MATERIAL_PBR_STANDARD = 0, MATERIAL_SKYBOX = 1, MATERIAL_LAMBERT = 2, MATERIAL_VOLUME = 3 };
typedef uint32_t MaterialIndex;
class Materials;
struct ImportedMaterial;

class Material { /* core/Material.hpp */
public:
    Material(const AssetInfo &info, Materials &materials, MaterialIndex index) : m_info(info), m_materials(materials), m_index(index) {}
    virtual ~Material() {}
    virtual MaterialType type() const = 0;
    const std::string &name() const { return m_info.name; }
    bool isEmbedded() const { return m_info.embedded; }
    MaterialIndex materialIndex() const { return m_index; }
    virtual bool isEmissive() const { return false; }
    virtual bool isTransparent() const { return false; }
    virtual void setTransparent(bool) {}
    ptc_material &block();
    const ptc_material &block() const;

protected:
    AssetInfo m_info;
    Materials &m_materials;
    MaterialIndex m_index;
};

class MaterialPBRStandard : public Material { /* vulkan/resources/VulkanMaterial.cpp:30-61 defaults */
public:
    MaterialPBRStandard(const AssetInfo &info, Materials &materials, MaterialIndex index);
    MaterialType type() const override { return MaterialType::MATERIAL_PBR_STANDARD; }
    vec4 &albedo() { return *reinterpret_cast<vec4 *>(block().albedo); }
    float &metallic() { return block().metallic_roughness_ao[0]; }
    float &roughness() { return block().metallic_roughness_ao[1]; }
    float &ao() { return block().metallic_roughness_ao[2]; }
    vec4 &emissive() { return *reinterpret_cast<vec4 *>(block().emissive); }
    float &emissiveIntensity() { return block().emissive[3]; }
    vec3 emissiveColor() const { return vec3(block().emissive[0], block().emissive[1], block().emissive[2]); }
    float &uTiling() { return block().uv_tiling[0]; }
    float &vTiling() { return block().uv_tiling[1]; }
    void setAlbedoTexture(Texture *t) { block().tex1[0] = t->bindlessResourceIndex; }
    void setMetallicTexture(Texture *t) { block().tex1[1] = t->bindlessResourceIndex; }
    void setRoughnessTexture(Texture *t) { block().tex1[2] = t->bindlessResourceIndex; }
    void setAOTexture(Texture *t) { block().tex1[3] = t->bindlessResourceIndex; }
    void setEmissiveTexture(Texture *t) { block().tex2[0] = t->bindlessResourceIndex; }
    void setNormalTexture(Texture *t) { block().tex2[1] = t->bindlessResourceIndex; }
    void setAlphaTexture(Texture *t) { block().tex2[3] = t->bindlessResourceIndex; }
    bool isEmissive() const override; /* core/Material.cpp:85-88 */
    bool isTransparent() const override { return block().metallic_roughness_ao[3] > 0; }
    void setTransparent(bool t) override { block().metallic_roughness_ao[3] = t ? 1.0f : 0.0f; }
};

class MaterialLambert : public Material { /* VulkanMaterial.cpp:229-248 defaults */
public:
    MaterialLambert(const AssetInfo &info, Materials &materials, MaterialIndex index);
    MaterialType type() const override { return MaterialType::MATERIAL_LAMBERT; }
    vec4 &albedo() { return *reinterpret_cast<vec4 *>(block().albedo); }
    float &ao() { return block().metallic_roughness_ao[2]; }
    vec4 &emissive() { return *reinterpret_cast<vec4 *>(block().emissive); }
    float &emissiveIntensity() { return block().emissive[3]; }
    vec3 emissiveColor() const { return vec3(block().emissive[0], block().emissive[1], block().emissive[2]); }
    float &uTiling() { return block().uv_tiling[0]; }
    float &vTiling() { return block().uv_tiling[1]; }
    void setAlbedoTexture(Texture *t) { block().tex1[0] = t->bindlessResourceIndex; }
    void setAOTexture(Texture *t) { block().tex1[3] = t->bindlessResourceIndex; }
    void setEmissiveTexture(Texture *t) { block().tex2[0] = t->bindlessResourceIndex; }
    void setNormalTexture(Texture *t) { block().tex2[1] = t->bindlessResourceIndex; }
    void setAlphaTexture(Texture *t) { block().tex2[3] = t->bindlessResourceIndex; }
    bool isEmissive() const override;
    bool isTransparent() const override { return block().metallic_roughness_ao[3] > 0; }
    void setTransparent(bool t) override { block().metallic_roughness_ao[3] = t ? 1.0f : 0.0f; }
};

class MaterialVolume : public Material { /* VulkanMaterial.cpp:436-482: aliases of the same 128 B block */
public:
    MaterialVolume(const AssetInfo &info, Materials &materials, MaterialIndex index);
    MaterialType type() const override { return MaterialType::MATERIAL_VOLUME; }
    vec4 &sigmaA() { return *reinterpret_cast<vec4 *>(block().albedo); }
    vec4 &sigmaS() { return *reinterpret_cast<vec4 *>(block().metallic_roughness_ao); }
    float &g() { return block().emissive[0]; }
};

class Materials {
public:
    explicit Materials(Textures &textures) : m_textures(textures) { m_blocks.reserve(512); }
    template <typename T>
    T *createMaterial(const AssetInfo &info) {
        if (m_map.has(info.name)) return dynamic_cast<T *>(m_map.get(info.name));
        if (m_blocks.size() >= 512) return nullptr; /* 512-slot UBO, VulkanLimits.hpp */
        m_blocks.push_back(ptc_material{});
        auto *m = new T(info, *this, (MaterialIndex)(m_blocks.size() - 1));
        m_owned.emplace_back(m);
        m_map.add(info.name, m);
        return m;
    }
    Material *createMaterial(const AssetInfo &info, MaterialType type);
    /* vulkan/resources/VulkanMaterials.cpp:154-233; EMBEDDED entries give nullptr */
    std::vector<Material *> createImportedMaterials(const std::vector<ImportedMaterial> &imported);
    Material *get(const std::string &n) const { return m_map.get(n); }
    const std::map<std::string, Material *> &all() const { return m_map.all(); }
    std::vector<ptc_material> &blocks() { return m_blocks; }
    const std::vector<ptc_material> &blocks() const { return m_blocks; }
    Textures &textures() { return m_textures; }

private:
    Textures &m_textures;
    std::vector<ptc_material> m_blocks;
    std::vector<std::unique_ptr<Material>> m_owned;
    AssetMap<Material> m_map;
};

/* ------------------------------------------------------------------ lights */
enum class LightType { POINT_LIGHT = 0, DIRECTIONAL_LIGHT = 1, MESH_LIGHT = 2, ENVIRONMENT_MAP = 3 };
typedef uint32_t LightIndex;
class Light { /* core/Light.hpp; LightData block vulkan/resources/VulkanLight.cpp:5-48 */
public:
    Light(const AssetInfo &info, LightType type, std::vector<ptc_light_data> &table, LightIndex idx)
        : m_info(info), m_type(type), m_table(table), m_index(idx) {}
    LightType type() const { return m_type; }
    LightIndex lightIndex() const { return m_index; }
    const std::string &name() const { return m_info.name; }
    /* RGB = colour, A = intensity */
    vec4 &color() { return *reinterpret_cast<vec4 *>(m_table[m_index].color); }

private:
    AssetInfo m_info;
    LightType m_type;
    std::vector<ptc_light_data> &m_table;
    LightIndex m_index;
};
typedef Light PointLight;
typedef Light DirectionalLight;

/* ------------------------------------------------------------------ math/Transform.cpp */
class Transform {
public:
    Transform() { computeBasisVectors(); }
    Transform(vec3 pos) : m_position(pos) { computeBasisVectors(); }
    Transform(vec3 pos, vec3 scale) : m_position(pos), m_scale(scale) { computeBasisVectors(); }
    Transform(vec3 pos, quat rot) : m_position(pos), m_rotation(rot) { computeBasisVectors(); }
    Transform(vec3 pos, vec3 scale, quat rot) : m_position(pos), m_scale(scale), m_rotation(rot) { computeBasisVectors(); }
    Transform(vec3 pos, vec3 scale, vec3 euler) : m_position(pos), m_scale(scale), m_rotation(quat(euler)) { computeBasisVectors(); }
    vec3 &position() { return m_position; }
    const vec3 &position() const { return m_position; }
    vec3 &scale() { return m_scale; }
    const vec3 &scale() const { return m_scale; }
    const quat &rotation() const { return m_rotation; }
    void setRotation(const quat &q) {
        m_rotation = q;
        computeBasisVectors();
    }
    void setRotationEuler(float x, float y, float z) { setRotation(quat(vec3(x, y, z))); }
    void setRotation(vec3 forward, vec3 up);
    mat4 getModelMatrix() const { return vm::translate(m_position) * vm::toMat4(m_rotation) * vm::scale(m_scale); }
    vec3 forward() const { return -m_z; }
    vec3 up() const { return m_y; }
    vec3 right() const { return m_x; }
    static constexpr float WORLD_Z[3] = {0, 0, 1};

private:
    void computeBasisVectors() {
        m_x = vm::rotate(m_rotation, vec3(1, 0, 0));
        m_y = vm::rotate(m_rotation, vec3(0, 1, 0));
        m_z = vm::rotate(m_rotation, vec3(0, 0, 1));
    }
    vec3 m_position{0, 0, 0}, m_scale{1, 1, 1};
    quat m_rotation;
    vec3 m_x, m_y, m_z;
};

/* ------------------------------------------------------------------ geometry */
typedef ptc_vertex Vertex; /* core/Mesh.hpp:15-40 */
class Model3D;
class Mesh {
public:
    std::string name;
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    uint32_t nTriangles() const { return (uint32_t)(indices.size() / 3); }
    uint32_t poolIndex = 0; /* index into the engine's mesh list */
    uint64_t uid = nextResourceUid(); /* identity of this mesh object for the flattened geometry pools (addresses can be reused) */
    Model3D *model = nullptr; /* Mesh::m_model */
};
class Model3D { /* core/Model3D.hpp */
public:
    struct Model3DNode {
        std::string name;
        std::vector<Mesh *> meshes;
        std::vector<Material *> materials; /* per mesh; nullptr = none imported */
        Transform transform;
        std::vector<Model3DNode> children;
    };
    std::string name, filepath;
    bool internal = false; /* engine-provided (AssetSource::ENGINE): not exported */
    std::vector<std::unique_ptr<Mesh>> meshes;
    Model3DNode nodeTree;
    Mesh *mesh(const std::string &n) const {
        for (auto &m : meshes)
            if (m->name == n) return m.get();
        return nullptr;
    }
};
void computeTangents(Mesh &mesh);
void computeNormals(Mesh &mesh); /* area-weighted vertex normals for files that carry none */

/* ------------------------------------------------------------------ core/io/ImportTypes.hpp */
enum class ImportedMaterialType { LAMBERT = 0, PBR_STANDARD = 1, EMBEDDED = 2, VOLUME = 3 };
struct ImportedTexture {
    std::string name;                /* texture asset name */
    std::string filepath;            /* STANDALONE textures of a scene file */
    bool embedded = true;            /* AssetLocation::DISK_EMBEDDED vs DISK_STANDALONE */
    std::shared_ptr<ImageU8> image;  /* decoded texels (already flipped like every stbi load of the reference, trap T11) */
    ColorSpace colorSpace = ColorSpace::sRGB;
};
struct ImportedMaterial {
    AssetInfo info;
    ImportedMaterialType type = ImportedMaterialType::PBR_STANDARD;
    vec4 albedo{1, 1, 1, 1};
    std::optional<ImportedTexture> albedoTexture;
    float roughness = 0.5f;
    std::optional<ImportedTexture> roughnessTexture;
    float metallic = 0.5f;
    std::optional<ImportedTexture> metallicTexture;
    float ao = 1.0f;
    std::optional<ImportedTexture> aoTexture;
    vec3 emissiveColor{0, 0, 0};
    std::optional<ImportedTexture> emissiveTexture;
    float emissiveStrength = 1.0f;
    std::optional<ImportedTexture> normalTexture;
    std::optional<ImportedTexture> alphaTexture;
    bool transparent = false;
    vec2 scale{1, 1};
    vec3 sigmaS{0.2f, 0.2f, 0.2f};
    vec3 sigmaA{0, 0, 0};
    float g = 0.0f;
};
struct ImportedModelNode {
    std::string name;
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<int32_t> materialIndices; /* per mesh, -1 = none */
    Transform transform;
    std::vector<ImportedModelNode> children;
};
struct ImportedCamera {
    vec3 position{0, 0, 0}, target{0, 0, -1}, up{0, 1, 0};
    float znear = 0.01f, zfar = 200.0f, lensRadius = 0.0f, focalDistance = 10.0f, fov = 60.0f;
    std::string volumeMaterial;
};
struct ImportedModel {
    std::string name, filepath;
};
struct ImportedLight {
    std::string name;
    LightType type = LightType::POINT_LIGHT;
    vec3 color{1, 1, 1};
    float intensity = 1.0f;
};
struct ImportedEnvironment {
    std::string path;
    vec3 backgroundColor{0, 0, 0};
    int environmentType = 0; /* EnvironmentType::SOLID_COLOR */
};
struct ImportedSceneObject {
    std::string name;
    bool active = true;
    Transform transform;
    bool hasMesh = false, hasMaterial = false, hasLight = false, hasVolume = false;
    std::string modelName, submesh; /* mesh component */
    std::string materialName;
    std::string lightName;
    bool lightShadows = true;
    std::string volumeFront, volumeBack;
    std::vector<ImportedSceneObject> children;
};
struct ImportedScene {
    ImportedCamera camera;
    std::vector<ImportedSceneObject> objects; /* roots */
    std::vector<ImportedModel> models;
    std::vector<ImportedMaterial> materials;
    std::vector<ImportedLight> lights;
    ImportedEnvironment environment;
    bool hasEnvironment = false;
    std::string sceneFolder; /* with trailing '/' */
};
/* io_obj.cpp / io_gltf.cpp: model import with the conventions of core/io/AssimpLoadModel.cpp (aiProcess_Triangulate |
 * aiProcess_FlipUVs | aiProcess_CalcTangentSpace followed by uv.y = 1 - uv.y); materials == nullptr skips them */
bool loadOBJ(const std::string &path, ImportedModelNode &root, std::vector<ImportedMaterial> *materials, std::string *err);
bool loadGLTF(const std::string &path, ImportedModelNode &root, std::vector<ImportedMaterial> *materials, std::string *err);
/* glm::decompose restricted to affine matrices without shear (AssimpLoadModel.cpp:154-160) */
Transform decomposeTransform(const mat4 &m);
/* io_scene.cpp: the scene file format of core/io/Import.cpp:504-565 / Export.cpp:628-784; throws std::runtime_error */
void importSceneFile(const std::string &filename, ImportedScene &out);

/* ------------------------------------------------------------------ core/Camera.cpp */
enum class CameraType { PERSPECTIVE = 0, ORTHOGRAPHIC = 1 };
class Camera {
public:
    virtual ~Camera() {}
    virtual CameraType type() const = 0;
    Transform &transform() { return m_transform; }
    const Transform &transform() const { return m_transform; }
    Material *&volume() { return m_volume; }
    mat4 viewMatrix() const { return vm::lookAt(m_transform.position(), m_transform.position() + m_transform.forward(), m_transform.up()); }
    mat4 viewMatrixInverse() const { return vm::inverse(viewMatrix()); }
    float &znear() { return m_znear; }
    float &zfar() { return m_zfar; }
    float &lensRadius() { return m_lensRadius; }
    float &focalDistance() { return m_focalDistance; }

private:
    Transform m_transform;
    Material *m_volume = nullptr;
    float m_znear = 0.5f, m_zfar = 50.0f, m_lensRadius = 0.0f, m_focalDistance = 10.0f;
};
class PerspectiveCamera : public Camera {
public:
    CameraType type() const override { return CameraType::PERSPECTIVE; }
    float &fov() { return m_fov; } /* degrees */
private:
    float m_fov = 60.0f;
};
class OrthographicCamera : public Camera {
public:
    CameraType type() const override { return CameraType::ORTHOGRAPHIC; }
    void setOrthoWidth(float w) { m_orthoWidth = w; }
    float orthoWidth() const { return m_orthoWidth; }
private:
    float m_orthoWidth = 10.0f;
};

/* ------------------------------------------------------------------ utils/ECS.hpp */
class Entity;
class Component {
public:
    virtual ~Component() {}
    bool shared = false;
    std::vector<Entity *> owners;
};
class ComponentMesh : public Component {
public:
    Mesh *mesh() const { return m_mesh; }
    void setMesh(Mesh *m) { m_mesh = m; }
private:
    Mesh *m_mesh = nullptr;
};
class ComponentMaterial : public Component {
public:
    Material *material() const { return m_material; }
    void setMaterial(Material *m) { m_material = m; }
private:
    Material *m_material = nullptr;
};
class ComponentLight : public Component {
public:
    Light *light() const { return m_light; }
    void setLight(Light *l) { m_light = l; }
    bool castShadows() const { return m_castShadows; }
    void setCastShadows(bool c) { m_castShadows = c; }
private:
    Light *m_light = nullptr;
    bool m_castShadows = true;
};
class ComponentVolume : public Component {
public:
    MaterialVolume *frontFacing() const { return m_front; }
    MaterialVolume *backFacing() const { return m_back; }
    void setFrontFacingVolume(MaterialVolume *m) { m_front = m; }
    void setBackFacingVolume(MaterialVolume *m) { m_back = m; }
private:
    MaterialVolume *m_front = nullptr, *m_back = nullptr;
};
struct ComponentOwnerUnique {};
struct ComponentOwnerShared {};
class ComponentManager { /* ECS.hpp:180-286: components live in per-type pools, 16 384 per type (ECS.hpp:22) */
public:
    static ComponentManager &getInstance() {
        static ComponentManager cm;
        return cm;
    }
    template <typename T, typename Owner>
    T *create() {
        T *c = new T();
        c->shared = std::is_same<Owner, ComponentOwnerShared>::value;
        m_live.emplace_back(c);
        return c;
    }
    template <typename T>
    void remove(T *c) {
        for (auto it = m_live.begin(); it != m_live.end(); ++it)
            if (it->get() == c) {
                m_live.erase(it);
                return;
            }
    }
private:
    std::vector<std::unique_ptr<Component>> m_live;
};
class Entity {
public:
    Entity() : m_id(s_nextId++) {}
    virtual ~Entity() {}
    template <typename T>
    T &add() {
        if (has<T>()) return get<T>();
        T *c = ComponentManager::getInstance().create<T, ComponentOwnerUnique>();
        c->owners.push_back(this);
        m_components[typeid(T).name()] = {c, false};
        return *c;
    }
    template <typename T>
    void add_shared(Component *sharedComponent) {
        if (has<T>()) return;
        sharedComponent->owners.push_back(this);
        m_components[typeid(T).name()] = {sharedComponent, true};
    }
    template <typename T>
    T &get() const {
        auto it = m_components.find(typeid(T).name());
        if (it == m_components.end()) throw std::runtime_error("Entity::get(): Component doesn't exist");
        return *static_cast<T *>(it->second.first);
    }
    template <typename T>
    bool has() const {
        return m_components.find(typeid(T).name()) != m_components.end();
    }
    template <typename T>
    void remove() {
        auto it = m_components.find(typeid(T).name());
        if (it == m_components.end()) return;
        auto c = it->second;
        m_components.erase(it);
        if (!c.second) ComponentManager::getInstance().remove<T>(static_cast<T *>(c.first));
    }
    uint32_t getID() const { return m_id; }
private:
    static uint32_t s_nextId;
    uint32_t m_id;
    std::map<std::string, std::pair<Component *, bool>> m_components;
};

/* ------------------------------------------------------------------ core/SceneNode.hpp + SceneObject */
class SceneObject : public Entity {
public:
    explicit SceneObject(const std::string &name) : m_name(name) {}
    const std::string &name() const { return m_name; }
    bool isActive() const { return m_active; }
    void setActive(bool a) { m_active = a; }
    void setLocalTransform(const Transform &t) { m_localTransform = t; }
    Transform &localTransform() { return m_localTransform; }
    const Transform &localTransform() const { return m_localTransform; }
    SceneObject *parent() const { return m_parent; }
    const std::vector<SceneObject *> &children() const { return m_children; }
    SceneObject *addChild(SceneObject *c) {
        m_children.push_back(c);
        c->m_parent = this;
        return c;
    }
    const mat4 &modelMatrix() const { return m_modelMatrix; }
    vec3 worldPosition() const { return vec3(m_modelMatrix[3][0], m_modelMatrix[3][1], m_modelMatrix[3][2]); }
    /* SceneNode::update (SceneNode.hpp:57-78): world = parent world * local TRS */
    void update(const mat4 *parentWorld) {
        m_modelMatrix = parentWorld ? (*parentWorld) * m_localTransform.getModelMatrix() : m_localTransform.getModelMatrix();
        for (auto *c : m_children) c->update(&m_modelMatrix);
    }
private:
    std::string m_name;
    bool m_active = true;
    Transform m_localTransform;
    SceneObject *m_parent = nullptr;
    std::vector<SceneObject *> m_children;
    mat4 m_modelMatrix{1.0f};
};
typedef std::vector<SceneObject *> SceneObjectVector;

/* ------------------------------------------------------------------ core/Scene.hpp */
enum class EnvironmentType { SOLID_COLOR = 0, HDRI = 1, SOLID_COLOR_WITH_HDRI_LIGHTING = 2 };
class Engine;
class Scene;

/* flattened, POD view of the scene = what crosses the C-ABI (core/Instances.cpp:145-181,
 * vulkan/VulkanInstances.cpp:66-130) */
/* The geometry half of a flattened scene: vertex / index pools of the instanced meshes.  Meshes do not change once imported (the
 * reference uploads them and builds their BLAS at import time, vulkan/resources/VulkanMesh.cpp), so the pools are rebuilt only when
 * the LIST of instanced meshes changes (key), not on every render() */
struct GeometryPools {
    std::vector<ptc_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<ptc_mesh> meshes;
    std::vector<uint64_t> key; /* per mesh slot: Mesh::uid, vertex count, index count */
};

struct FlatScene {
    std::shared_ptr<const GeometryPools> geometry;
    std::vector<ptc_instance> instances;
    std::vector<ptc_material> materials;
    std::vector<ptc_light_data> lightData;
    std::vector<ptc_light_instance> lightInstances;
    std::vector<ptc_texture> textures;
    ptc_scene_desc desc{};
};

class InstancesManager {
public:
    explicit InstancesManager(Scene *scene) : m_scene(scene) {}
    void build();
    void invalidate();
    const SceneObjectVector &lights() const { return m_lights; }
    const SceneObjectVector &meshLights() const { return m_meshLights; }
    const SceneObjectVector &volumes() const { return m_volumes; }
    const SceneObjectVector &instanceOrder() const { return m_order; } /* InstanceData slot -> object */
private:
    Scene *m_scene;
    std::vector<std::pair<Mesh *, SceneObjectVector>> m_instancesOpaque; /* grouped by mesh, insertion order */
    SceneObjectVector m_transparent, m_lights, m_meshLights, m_volumes, m_order;
};

class Scene {
public:
    explicit Scene(Engine &engine) : m_engine(engine), m_instances(this) {}
    std::shared_ptr<Camera> &camera() { return m_camera; }
    float &exposure() { return m_exposure; }
    float &environmentIntensity() { return m_environmentIntensity; }
    EnvironmentType &environmentType() { return m_environmentType; }
    vec3 &backgroundColor() { return m_backgroundColor; }
    EnvironmentMap *&skyboxMaterial() { return m_skybox; }

    SceneObject *addSceneObject(const std::string &name, Transform transform) { return addSceneObject(name, nullptr, transform); }
    SceneObject *addSceneObject(const std::string &name, SceneObject *parent, Transform transform);
    void clear();
    void update(); /* Scene.cpp:105-145 */
    SceneObjectVector getSceneObjectsFlat() const;
    const SceneObjectVector &sceneGraph() const { return m_sceneGraph; }
    Light *createLight(const AssetInfo &info, LightType type, vec4 color = vec4(1, 1, 1, 1));
    InstancesManager &instancesManager() { return m_instances; }
    ptc_scene_data getSceneData() const; /* Scene.cpp:31-44 + VulkanScene.cpp:85-92 */
    std::vector<ptc_light_data> &lightData() { return m_lightData; }
    Engine &engine() { return m_engine; }

private:
    Engine &m_engine;
    std::shared_ptr<Camera> m_camera;
    float m_exposure = 0.0f, m_environmentIntensity = 1.0f;
    EnvironmentType m_environmentType = EnvironmentType::HDRI;
    vec3 m_backgroundColor{0, 0, 0};
    EnvironmentMap *m_skybox = nullptr;
    std::vector<std::unique_ptr<SceneObject>> m_objects;
    SceneObjectVector m_sceneGraph; /* roots */
    std::vector<ptc_light_data> m_lightData;
    std::vector<std::unique_ptr<Light>> m_lights;
    InstancesManager m_instances;
};

/* core/SceneUtils.cpp:76-130: instantiates an imported model's node tree under `parent` (root transform and material can be overridden) */
void addModel3D(Scene &scene, SceneObject *parent, const std::string &modelName, std::optional<Transform> overrideRootTransform = std::nullopt,
                std::optional<std::string> overrideMaterial = std::nullopt);

/* ------------------------------------------------------------------ core/Renderer.hpp:11-39 */
class RendererPathTracing {
public:
    struct RenderInfo {
        uint32_t samples = 256u;
        uint32_t depth = 9u;
        uint32_t batchSize = 4u;
        std::string filename = "test";
        FileType fileType = FileType::HDR;
        float exposure = 0.0F;
        uint32_t width = 1080u;
        uint32_t height = 1080u;
        bool denoise = false;
        bool writeAllFiles = false;
        /* extension: the reference selects its sampler at shader-compile time (SAMPLING_RTGEMS / SAMPLING_PMJ in
         * defines_pt.glsl); here it is a render setting.  true = low-discrepancy (PTC_FLAG_SAMPLER_SOBOL) */
        bool lowDiscrepancySampler = false;
        /* extension, off for parity (the reference never light-samples the environment, lightSampling.glsl:101-106):
         * luminance importance sampling of the HDRI environment + MIS (PTC_FLAG_ENV_IMPORTANCE) */
        bool environmentImportanceSampling = false;
        /* extension: how a renderer that drives several GPUs partitions the image: 0 = backend default (sample
         * batches), 1 = image tiles (ptc_split_mode PTC_SPLIT_TILE), 2 = sample batches.  Ignored on one GPU. */
        uint32_t multiGpuSplit = 0u;
        /* extension: the reference's optional PMJ02BN sampler (rng_pmj.glsl; tables of math/PMJSequences.cpp, BlueNoise.cpp) */
        bool pmjSampler = false;
    };
    virtual ~RendererPathTracing() {}
    RenderInfo &renderInfo() { return m_renderInfo; }
    const RenderInfo &renderInfo() const { return m_renderInfo; }
    virtual bool isRayTracingEnabled() const = 0;
    virtual void render() = 0;
    virtual float renderProgress() = 0;
private:
    RenderInfo m_renderInfo;
};

/* the function table of the CUDA path-tracing core (libptc_cuda.so next to this library; include/ptc.h).  The host library
 * resolves it at run time so that it links no CUDA itself; there is no other backend and no CPU fallback. */
struct PtcBackend {
    void *handle = nullptr;
    std::string path;
    decltype(&ptc_create) create = nullptr;
    decltype(&ptc_destroy) destroy = nullptr;
    decltype(&ptc_last_error) last_error = nullptr;
    decltype(&ptc_backend_name) backend_name = nullptr;
    decltype(&ptc_upload_scene) upload_scene = nullptr;
    decltype(&ptc_build_accel) build_accel = nullptr;
    decltype(&ptc_render) render = nullptr;
    decltype(&ptc_progress) progress = nullptr;
    decltype(&ptc_get_stats) get_stats = nullptr;
    decltype(&ptc_device_count) device_count = nullptr;
    decltype(&ptc_comm_unique_id) comm_unique_id = nullptr;
    decltype(&ptc_comm_init_rank) comm_init_rank = nullptr;
    decltype(&ptc_set_sampler_tables) set_sampler_tables = nullptr;
    bool load(const std::string &libPath, std::string *err);
};

/* B200 implementation of the boundary: replaces VulkanRendererPathTracing (…PathTracing.cpp:121-226, 791-1027) */
class CudaRendererPathTracing : public RendererPathTracing {
public:
    explicit CudaRendererPathTracing(Engine &engine);
    ~CudaRendererPathTracing() override;
    bool isRayTracingEnabled() const override { return m_isInitialized; }
    void render() override;
    float renderProgress() override;
    /* extras used by tools / tests */
    bool renderToMemory(std::vector<float> &radiance, std::vector<float> &albedo, std::vector<float> &normal);
    /* same, into caller-owned width * height * 4 floats each (any may be null: that target is not read back) */
    bool renderToBuffers(float *radiance, float *albedo, float *normal);
    ptc_render_params makeRenderParams();
    const ptc_stats &lastStats() const { return m_stats; }
    const std::string &lastError() const { return m_error; }
    const char *backendName() const { return m_backend.backend_name ? m_backend.backend_name() : "none"; }
    /* multi-GPU.  setDevices: this renderer drives the listed GPUs of the box from one process (NCCL communicator
     * inside the core); the scene is re-uploaded by the next render().  commUniqueId / commInitRank: one process per GPU
     * (torchrun, MPI) - every rank's renderer joins one communicator; render() is then collective and only rank 0 gets /
     * writes the image. */
    bool setDevices(const std::vector<int> &deviceIds);
    int deviceCount() const { return (m_ctx && m_backend.device_count) ? m_backend.device_count(m_ctx) : 0; }
    bool commUniqueId(uint8_t out128[128]);
    bool commInitRank(const uint8_t id128[128], int rank, int world);
    int commRank() const { return m_commRank; }
    int commWorld() const { return m_commWorld; }

private:
    Engine &m_engine;
    PtcBackend m_backend;
    ptc_ctx *m_ctx = nullptr;
    bool m_isInitialized = false;
    bool m_renderInProgress = false;
    int m_commRank = 0, m_commWorld = 1;
    bool m_samplerTablesSet = false; /* VulkanRandom's two tables were handed to the current context */
    bool ensureSamplerTables();
    ptc_stats m_stats{};
    std::string m_error;
};

class Renderer { /* core/Renderer.hpp:41-60 */
public:
    explicit Renderer(std::unique_ptr<CudaRendererPathTracing> pt) : m_pt(std::move(pt)) {}
    CudaRendererPathTracing &rendererPathTracing() { return *m_pt; }
private:
    std::unique_ptr<CudaRendererPathTracing> m_pt;
};

/* ------------------------------------------------------------------ core/Engine.hpp + vulkan/VulkanEngine.cpp */
class Engine {
public:
    /* The path tracer is the CUDA core next to this library (vviewer_b200/_lib/libptc_cuda.so): no other backend, no CPU
     * fallback.  assetRoot: directory that contains assets/ (empty = $VVIEWER_ASSETS or the repository root). */
    explicit Engine(const std::string &name, const std::string &assetRoot = "");
    ~Engine();
    void initResources(); /* VulkanEngine.cpp:32-48 + initDefaultData :331-391 */
    void releaseResources() {}
    Scene &scene() { return *m_scene; }
    Textures &textures() { return *m_textures; }
    Materials &materials() { return *m_materials; }
    Renderer &renderer() { return *m_renderer; }
    AssetMap<Model3D> &modelsMap() { return m_models; }
    AssetMap<Light> &lightsMap() { return m_lightsMap; }
    Model3D *importModel(const AssetInfo &info, bool importMaterials = true); /* VulkanEngine.cpp:157-192 */
    EnvironmentMap *importEnvironmentMap(const AssetInfo &info);              /* VulkanEngine.cpp:194-232 */
    Model3D *addModel(std::unique_ptr<Model3D> model);                       /* procedural meshes */
    /* the vviewer "Import scene" action (src/bin/vviewer/UI/MainWindow.cpp:580-680) without the UI: parse the scene file, import its
     * models / materials / lights, rebuild the scene graph, camera and environment.  false + message on failure */
    bool importScene(const std::string &filename, std::string *err = nullptr);
    /* Scene::exportScene (core/Scene.cpp:174, core/io/Export.cpp:628-784): writes <name>/scene.json + <name>/assets/ */
    bool exportScene(const std::string &name, std::string *err = nullptr);
    std::string assetPath(const std::string &rel) const;
    const std::vector<Mesh *> &meshPool() const { return m_meshPool; }
    /* flatten the current scene (after Scene::update) into the POD arrays of include/ptc.h */
    void flatten(FlatScene &out);

private:
    std::string m_name, m_assetRoot;
    std::unique_ptr<Textures> m_textures;
    std::unique_ptr<Materials> m_materials;
    std::unique_ptr<Scene> m_scene;
    std::unique_ptr<Renderer> m_renderer;
    AssetMap<Model3D> m_models;
    AssetMap<Light> m_lightsMap;
    std::vector<std::unique_ptr<Model3D>> m_ownedModels;
    std::vector<std::unique_ptr<EnvironmentMap>> m_envMaps;
    std::vector<Mesh *> m_meshPool;
    std::shared_ptr<const GeometryPools> m_geometry; /* of the last flatten() */
};

}  // namespace vengine
