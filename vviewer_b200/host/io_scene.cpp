/*
 * Scene files: the reference's on-disk scene format (`<scene>/scene.json` + `<scene>/assets/{models,textures}/`).
 *   importSceneFile()      src/lib/vengine/core/io/Import.cpp:41-565 (field names, defaults, accepted vector / rotation
 *                          spellings, error cases)
 *   Engine::importScene()  what vviewer does with the parsed data, src/bin/vviewer/UI/MainWindow.cpp:580-680, :320-379
 *   Engine::exportScene()  src/lib/vengine/core/io/Export.cpp:53-784 (member order, default-transform elision, asset copies)
 *   addModel3D()           src/lib/vengine/core/SceneUtils.cpp:76-130
 * This gives offlinerender a data-driven `--scene <file.json>` entry next to the compiled-in recipes.
 */
#include "json.hpp"
#include "vengine.hpp"

#include <sys/stat.h>
#include <cstdio>
#include <fstream>
#include <sstream>

namespace vengine {
namespace {

using json::Value;

float parseFloat(const Value &o, const char *name, float def) {
    const Value *v = o.find(name);
    return v ? v->getFloat() : def;
}

/* vectors are accepted as {x,y,..} / {r,g,..} objects, as arrays of 1 or N numbers, or as one number (Import.cpp:52-190) */
template <int N>
bool parseVecN(const Value &o, const char *name, float *out, bool *malformed = nullptr) {
    const Value *v = o.find(name);
    if (!v) return false;
    static const char *xyzw[4] = {"x", "y", "z", "w"}, *rgba[4] = {"r", "g", "b", "a"};
    auto bad = [&]() -> bool {
        if (malformed) {
            *malformed = true;
            return false;
        }
        throw std::runtime_error("parseVec" + std::to_string(N) + "(): Field " + name + " is malformed");
    };
    if (v->isObject()) {
        for (int i = 0; i < N; i++) {
            const Value *c = v->find(xyzw[i]);
            if (!c) c = v->find(rgba[i]);
            if (!c || !c->isNumber()) return bad();
            out[i] = c->getFloat();
        }
    } else if (v->isArray()) {
        if (v->size() == 1) {
            for (int i = 0; i < N; i++) out[i] = (*v)[0].getFloat();
        } else if (v->size() == (size_t)N) {
            for (int i = 0; i < N; i++) out[i] = (*v)[(size_t)i].getFloat();
        } else {
            return bad();
        }
    } else if (v->isNumber()) {
        for (int i = 0; i < N; i++) out[i] = v->getFloat();
    } else {
        return bad();
    }
    return true;
}
vec2 parseVec2(const Value &o, const char *name, vec2 def) {
    float f[2];
    return parseVecN<2>(o, name, f) ? vec2(f[0], f[1]) : def;
}
vec3 parseVec3(const Value &o, const char *name, vec3 def) {
    float f[3];
    return parseVecN<3>(o, name, f) ? vec3(f[0], f[1], f[2]) : def;
}
vec4 parseVec4(const Value &o, const char *name, vec4 def) {
    float f[4];
    return parseVecN<4>(o, name, f) ? vec4(f[0], f[1], f[2], f[3]) : def;
}
/* three numbers = Euler angles in degrees, four = quaternion (x, y, z, w) (Import.cpp:192-214) */
quat parseRotation(const Value &o, const char *name, quat def) {
    if (!o.has(name)) return def;
    float f[4];
    bool malformed = false;
    if (parseVecN<3>(o, name, f, &malformed)) return quat(vec3(vm::radians(f[0]), vm::radians(f[1]), vm::radians(f[2])));
    if (!parseVecN<4>(o, name, f)) return def;
    return quat(f[3], f[0], f[1], f[2]);
}

ImportedCamera parseCamera(const Value &o) {
    ImportedCamera c;
    c.position = parseVec3(o, "position", c.position);
    if (o.has("target")) {
        c.target = parseVec3(o, "target", c.target);
        c.up = parseVec3(o, "up", c.up);
    } else if (o.has("rotation")) {
        quat r = parseRotation(o, "rotation", quat());
        c.target = c.position + vm::rotate(r, vec3(0, 0, -1));
        c.up = vm::rotate(r, vec3(0, 1, 0));
    } else {
        throw std::runtime_error("parseCamera(): Missing target or rotation information");
    }
    c.fov = parseFloat(o, "fov", c.fov);
    c.znear = parseFloat(o, "znear", c.znear);
    c.zfar = parseFloat(o, "zfar", c.zfar);
    c.lensRadius = parseFloat(o, "lensRadius", c.lensRadius);
    c.focalDistance = parseFloat(o, "focalDistance", c.focalDistance);
    c.volumeMaterial = o.has("volume") ? o["volume"].getString() : std::string();
    return c;
}

ImportedSceneObject parseSceneObject(const Value &o) {
    ImportedSceneObject obj;
    obj.name = o["name"].getString();
    if (o.has("active")) obj.active = o["active"].getBool();
    if (o.has("transform")) {
        const Value &t = o["transform"];
        vec3 position = parseVec3(t, "position", vec3(0, 0, 0));
        vec3 scale = parseVec3(t, "scale", vec3(1, 1, 1));
        quat rotation = parseRotation(t, "rotation", quat());
        obj.transform = Transform(position, scale, rotation);
    }
    if (o.has("mesh")) {
        const Value &m = o["mesh"];
        if (!m.has("modelName")) throw std::runtime_error("parseMeshComponent(): Imported mesh doesn't have a modelName set");
        if (!m.has("submesh")) throw std::runtime_error("parseMeshComponent(): Imported mesh doesn't have a submesh set");
        obj.hasMesh = true;
        obj.modelName = m["modelName"].getString();
        obj.submesh = m["submesh"].getString();
    }
    if (o.has("material")) {
        const Value &m = o["material"];
        if (!m.has("name")) throw std::runtime_error("parseMaterialComponent(): Imported material doesn't have a name");
        obj.hasMaterial = true;
        obj.materialName = m["name"].getString();
    }
    if (o.has("light")) {
        const Value &l = o["light"];
        if (!l.has("name")) throw std::runtime_error("parseLightComponent(): Imported light doesn't have a name");
        obj.hasLight = true;
        obj.lightName = l["name"].getString();
        if (l.has("shadows")) obj.lightShadows = l["shadows"].getBool();
    }
    if (o.has("volume")) {
        const Value &v = o["volume"];
        obj.hasVolume = true;
        if (v.has("frontFacing")) obj.volumeFront = v["frontFacing"].getString();
        if (v.has("backFacing")) obj.volumeBack = v["backFacing"].getString();
    }
    if (o.has("children"))
        for (const Value &c : o["children"].items()) obj.children.push_back(parseSceneObject(c));
    return obj;
}

std::optional<ImportedTexture> parseTexture(const Value &o, const std::string &folder, ColorSpace cs) {
    if (!o.has("texture")) return std::nullopt;
    const Value &t = o["texture"];
    ImportedTexture tex;
    std::string type = t["type"].getString();
    if (type != "STANDALONE" && type != "EMBEDDED") throw std::runtime_error("parseTexture(): unknown texture type " + type);
    tex.embedded = type == "EMBEDDED";
    tex.name = t["name"].getString();
    tex.colorSpace = cs;
    if (!tex.embedded && t.has("filepath")) {
        tex.filepath = folder + t["filepath"].getString();
        auto img = std::make_shared<ImageU8>();
        /* Image<uint8_t>(AssetInfo, colorSpace) -> stbi_load with the engine-wide vertical flip (trap T11) */
        if (!loadImageU8(tex.filepath, *img, true)) throw std::runtime_error("parseTexture(): unable to load " + tex.filepath);
        tex.image = img;
    }
    return tex;
}

void parseMaterial(const Value &o, const std::string &folder, ImportedMaterial &m) {
    std::string name = o["name"].getString();
    std::string type = o["type"].getString();
    if (type == "LAMBERT") m.type = ImportedMaterialType::LAMBERT;
    else if (type == "PBR_STANDARD") m.type = ImportedMaterialType::PBR_STANDARD;
    else if (type == "EMBEDDED") m.type = ImportedMaterialType::EMBEDDED;
    else if (type == "VOLUME") m.type = ImportedMaterialType::VOLUME;
    else throw std::runtime_error("parseMaterial(): " + type + " material not supported");
    m.info = AssetInfo(name);
    if (m.type == ImportedMaterialType::EMBEDDED) return;
    if (m.type == ImportedMaterialType::PBR_STANDARD) {
        if (o.has("roughness")) {
            m.roughnessTexture = parseTexture(o["roughness"], folder, ColorSpace::LINEAR);
            m.roughness = parseFloat(o["roughness"], "value", m.roughness);
        }
        if (o.has("metallic")) {
            m.metallicTexture = parseTexture(o["metallic"], folder, ColorSpace::LINEAR);
            m.metallic = parseFloat(o["metallic"], "value", m.metallic);
        }
    }
    if (o.has("albedo")) {
        m.albedoTexture = parseTexture(o["albedo"], folder, ColorSpace::sRGB);
        m.albedo = parseVec4(o["albedo"], "value", m.albedo);
    }
    if (o.has("ao")) {
        m.aoTexture = parseTexture(o["ao"], folder, ColorSpace::LINEAR);
        m.ao = parseFloat(o["ao"], "value", m.ao);
    }
    if (o.has("emissive")) {
        m.emissiveTexture = parseTexture(o["emissive"], folder, ColorSpace::sRGB);
        vec4 e = parseVec4(o["emissive"], "value", vec4(m.emissiveColor, m.emissiveStrength));
        m.emissiveColor = vec3(e.x, e.y, e.z);
        m.emissiveStrength = e.w;
    }
    if (o.has("normal")) m.normalTexture = parseTexture(o["normal"], folder, ColorSpace::LINEAR);
    if (o.has("alpha")) m.alphaTexture = parseTexture(o["alpha"], folder, ColorSpace::LINEAR);
    if (o.has("transparent")) m.transparent = o["transparent"].getBool();
    if (o.has("scale")) m.scale = parseVec2(o, "scale", m.scale);
    if (o.has("scattering")) m.sigmaS = parseVec3(o["scattering"], "value", m.sigmaS);
    if (o.has("absorption")) m.sigmaA = parseVec3(o["absorption"], "value", m.sigmaA);
    if (o.has("g")) m.g = o["g"].getFloat();
}

ImportedLight parseLight(const Value &o) {
    ImportedLight l;
    l.name = o["name"].getString();
    if (o.has("type")) {
        std::string type = o["type"].getString();
        if (type == "POINT") l.type = LightType::POINT_LIGHT;
        else if (type == "DIRECTIONAL") l.type = LightType::DIRECTIONAL_LIGHT;
        else throw std::runtime_error("parseLight(): " + type + " light not supported");
    }
    l.color = parseVec3(o, "color", l.color);
    if (o.has("intensity")) l.intensity = o["intensity"].getFloat();
    return l;
}

bool makeDir(const std::string &p) {
    if (mkdir(p.c_str(), 0777) == 0) return true;
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
std::string fileNameOf(const std::string &p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? p : p.substr(k + 1);
}
/* copyFileToDirectoryAndGetFileName (Export.cpp:33-51); an existing destination is kept */
std::string copyInto(const std::string &file, const std::string &directory, const std::string &newName = "") {
    std::string name = newName.empty() ? fileNameOf(file) : newName;
    std::string dst = directory + name;
    struct stat st;
    if (stat(dst.c_str(), &st) != 0) {
        std::ifstream in(file, std::ios::binary);
        if (in) {
            std::ofstream out(dst, std::ios::binary);
            out << in.rdbuf();
        } else {
            std::fprintf(stderr, "Export copy error: cannot read %s\n", file.c_str());
        }
    }
    return name;
}

Value vecValue(std::initializer_list<float> v) {
    Value a = Value::array();
    for (float f : v) a.push(Value(f));
    return a;
}

/* glm::eulerAngles(quat): pitch, yaw, roll */
vec3 eulerAngles(const quat &q) {
    float py = 2.0f * (q.y * q.z + q.w * q.x), px = q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
    float pitch = (std::fabs(px) < 1e-7f && std::fabs(py) < 1e-7f) ? 2.0f * std::atan2(q.x, q.w) : std::atan2(py, px);
    float yaw = std::asin(std::min(std::max(-2.0f * (q.x * q.z - q.w * q.y), -1.0f), 1.0f));
    float ry = 2.0f * (q.x * q.y + q.w * q.z), rx = q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z;
    float roll = (std::fabs(rx) < 1e-7f && std::fabs(ry) < 1e-7f) ? 0.0f : std::atan2(ry, rx);
    return vec3(pitch, yaw, roll);
}

}  // namespace

void importSceneFile(const std::string &filename, ImportedScene &out) {
    out = ImportedScene();
    size_t slash = filename.find_last_of('/');
    out.sceneFolder = slash == std::string::npos ? std::string("") : filename.substr(0, slash + 1);
    std::ifstream in(filename, std::ios::binary);
    if (!in) throw std::runtime_error("Can't open file: " + filename);
    std::stringstream ss;
    ss << in.rdbuf();
    Value doc;
    try {
        doc = json::parse(ss.str());
    } catch (std::exception &) {
        throw std::runtime_error("Malformed scene file: " + filename);
    }
    if (!doc.isObject()) throw std::runtime_error("Malformed scene file: " + filename);
    out.camera = parseCamera(doc["camera"]);
    for (const Value &o : doc["scene"].items()) out.objects.push_back(parseSceneObject(o));
    for (const Value &o : doc["models"].items()) {
        ImportedModel m;
        if (!o.has("name")) throw std::runtime_error("parseModel(): Imported model doesn't have a name set");
        if (!o.has("filepath")) throw std::runtime_error("parseModel(): Imported mesh doesn't have a filepath set");
        m.name = o["name"].getString();
        m.filepath = out.sceneFolder + o["filepath"].getString();
        out.models.push_back(m);
    }
    const Value &mats = doc["materials"];
    out.materials.resize(mats.size());
    for (size_t i = 0; i < mats.size(); i++) parseMaterial(mats[i], out.sceneFolder, out.materials[i]);
    for (const Value &o : doc["lights"].items()) out.lights.push_back(parseLight(o));
    if (doc.has("environment")) {
        const Value &e = doc["environment"];
        out.hasEnvironment = true;
        if (e.has("path")) out.environment.path = e["path"].getString();
        out.environment.environmentType = (int)e["environmentType"].getInt();
        out.environment.backgroundColor = parseVec3(e, "backgroundColor", out.environment.backgroundColor);
    }
}

void addModel3D(Scene &scene, SceneObject *parent, const std::string &modelName, std::optional<Transform> overrideRootTransform,
                std::optional<std::string> overrideMaterial) {
    Engine &engine = scene.engine();
    Model3D *model = engine.modelsMap().get(modelName);
    if (!model) throw std::runtime_error("addModel3D(): model " + modelName + " is not imported");
    Material *defaultMat = engine.materials().get("defaultMaterial");
    Material *overrideMat = overrideMaterial ? engine.materials().get(*overrideMaterial) : nullptr;
    std::function<void(const Model3D::Model3DNode &, SceneObject *, bool)> add = [&](const Model3D::Model3DNode &node, SceneObject *under, bool isRoot) {
        Transform t = (isRoot && overrideRootTransform) ? *overrideRootTransform : node.transform;
        /* one object per mesh of the node, all named after the model */
        for (size_t i = 0; i < node.meshes.size(); i++) {
            SceneObject *so = scene.addSceneObject(modelName, under, t);
            so->add<ComponentMesh>().setMesh(node.meshes[i]);
            Material *m = overrideMat ? overrideMat : (node.materials[i] ? node.materials[i] : defaultMat);
            so->add<ComponentMaterial>().setMaterial(m);
        }
        if (!node.children.empty()) {
            SceneObject *group = scene.addSceneObject(node.name, under, t);
            for (const Model3D::Model3DNode &c : node.children) add(c, group, false);
        }
    };
    add(model->nodeTree, parent, true);
}

bool Engine::importScene(const std::string &filename, std::string *err) {
    ImportedScene in;
    try {
        importSceneFile(filename, in);
    } catch (std::exception &e) {
        if (err) *err = std::string("Unable to open scene file: ") + e.what();
        return false;
    }
    try {
        /* camera (MainWindow.cpp:609-625) */
        auto cam = std::make_shared<PerspectiveCamera>();
        cam->transform().position() = in.camera.position;
        cam->transform().setRotation(vm::normalize(in.camera.target - in.camera.position), in.camera.up);
        cam->znear() = in.camera.znear;
        cam->zfar() = in.camera.zfar;
        cam->lensRadius() = in.camera.lensRadius;
        cam->focalDistance() = in.camera.focalDistance;
        cam->fov() = in.camera.fov;
        m_scene->camera() = cam;
        m_scene->clear();

        for (const ImportedModel &m : in.models)
            if (!importModel(AssetInfo(m.name, m.filepath), true)) throw std::runtime_error("failed to import model " + m.filepath);
        m_materials->createImportedMaterials(in.materials);
        for (const ImportedLight &l : in.lights)
            if (!m_lightsMap.has(l.name)) m_lightsMap.add(l.name, m_scene->createLight(AssetInfo(l.name), l.type, vec4(l.color, l.intensity)));

        /* scene graph (MainWindow.cpp:320-379) */
        std::function<void(const ImportedSceneObject &, SceneObject *)> addObject = [&](const ImportedSceneObject &o, SceneObject *parent) {
            SceneObject *so = m_scene->addSceneObject(o.name, parent, o.transform);
            if (!o.active) so->setActive(false);
            if (o.hasMesh) {
                Model3D *model = m_models.get(o.modelName);
                if (!model) throw std::runtime_error("scene object " + o.name + " uses the unknown model " + o.modelName);
                Mesh *mesh = model->mesh(o.submesh);
                if (!mesh) throw std::runtime_error("model " + o.modelName + " has no submesh " + o.submesh);
                so->add<ComponentMesh>().setMesh(mesh);
            }
            if (o.hasMaterial) {
                Material *mat = m_materials->get(o.materialName);
                if (!mat) throw std::runtime_error("scene object " + o.name + " uses the unknown material " + o.materialName);
                so->add<ComponentMaterial>().setMaterial(mat);
            }
            if (o.hasLight) {
                Light *light = m_lightsMap.get(o.lightName);
                if (!light) throw std::runtime_error("scene object " + o.name + " uses the unknown light " + o.lightName);
                ComponentLight &lc = so->add<ComponentLight>();
                lc.setLight(light);
                lc.setCastShadows(o.lightShadows);
            }
            if (o.hasVolume) {
                ComponentVolume &vc = so->add<ComponentVolume>();
                if (!o.volumeFront.empty()) {
                    Material *v = m_materials->get(o.volumeFront);
                    if (v && v->type() == MaterialType::MATERIAL_VOLUME) vc.setFrontFacingVolume(static_cast<MaterialVolume *>(v));
                }
                if (!o.volumeBack.empty()) {
                    Material *v = m_materials->get(o.volumeBack);
                    if (v && v->type() == MaterialType::MATERIAL_VOLUME) vc.setBackFacingVolume(static_cast<MaterialVolume *>(v));
                }
            }
            for (const ImportedSceneObject &c : o.children) addObject(c, so);
        };
        for (const ImportedSceneObject &o : in.objects) addObject(o, nullptr);

        /* environment (MainWindow.cpp:655-672) */
        if (!in.environment.path.empty()) {
            EnvironmentMap *env = importEnvironmentMap(AssetInfo(in.sceneFolder + in.environment.path));
            if (env) m_scene->skyboxMaterial() = env;
        }
        if (in.hasEnvironment) {
            m_scene->backgroundColor() = in.environment.backgroundColor;
            m_scene->environmentType() = (EnvironmentType)in.environment.environmentType;
        }
        if (!in.camera.volumeMaterial.empty()) {
            Material *v = m_materials->get(in.camera.volumeMaterial);
            if (v && v->type() == MaterialType::MATERIAL_VOLUME) cam->volume() = v;
        }
        m_scene->update();
    } catch (std::exception &e) {
        if (err) *err = e.what();
        return false;
    }
    return true;
}

bool Engine::exportScene(const std::string &name, std::string *err) {
    auto failWith = [&](const std::string &m) {
        if (err) *err = m;
        return false;
    };
    const std::string sceneFolder = name + "/", assetsFolder = sceneFolder + "assets/", modelsFolder = assetsFolder + "models/",
                      texturesFolder = assetsFolder + "textures/";
    if (!makeDir(sceneFolder) || !makeDir(assetsFolder) || !makeDir(modelsFolder) || !makeDir(texturesFolder))
        return failWith("cannot create " + sceneFolder);
    Camera *camera = m_scene->camera().get();
    if (!camera || camera->type() != CameraType::PERSPECTIVE) return failWith("SceneExport: Only perspective camera supported");

    Value d = Value::object();
    d.set("version", "1.0");
    d.set("name", "scene.json");

    Value cam = Value::object();
    {
        const Transform &t = camera->transform();
        vec3 p = t.position(), target = p + t.forward(), up = t.up();
        cam.set("position", vecValue({p.x, p.y, p.z}));
        cam.set("target", vecValue({target.x, target.y, target.z}));
        cam.set("up", vecValue({up.x, up.y, up.z}));
        cam.set("znear", Value(camera->znear()));
        cam.set("zfar", Value(camera->zfar()));
        cam.set("lensRadius", Value(camera->lensRadius()));
        cam.set("focalDistance", Value(camera->focalDistance()));
        cam.set("fov", Value(static_cast<PerspectiveCamera *>(camera)->fov()));
        cam.set("volume", camera->volume() ? camera->volume()->name() : std::string(""));
    }

    /* which materials are in use decides nothing here: like the reference, every material of the map is written */
    std::function<void(Value &, SceneObject *)> writeObject = [&](Value &array, SceneObject *so) {
        Value o = Value::object();
        o.set("name", so->name());
        if (!so->isActive()) o.set("active", false);
        {
            const Transform &t = so->localTransform();
            vec3 e = eulerAngles(t.rotation());
            vec3 deg(e.x * 57.295779513082320876798154814105f, e.y * 57.295779513082320876798154814105f, e.z * 57.295779513082320876798154814105f);
            bool isDefault = t.position().x == 0 && t.position().y == 0 && t.position().z == 0 && t.scale().x == 1 && t.scale().y == 1 &&
                             t.scale().z == 1 && deg.x == 0 && deg.y == 0 && deg.z == 0;
            if (!isDefault) {
                Value tr = Value::object();
                tr.set("position", vecValue({t.position().x, t.position().y, t.position().z}));
                tr.set("scale", vecValue({t.scale().x, t.scale().y, t.scale().z}));
                tr.set("rotation", vecValue({deg.x, deg.y, deg.z}));
                o.set("transform", tr);
            }
        }
        if (so->has<ComponentMesh>() && so->get<ComponentMesh>().mesh()) {
            Mesh *mesh = so->get<ComponentMesh>().mesh();
            Value m = Value::object();
            m.set("modelName", mesh->model ? mesh->model->name : std::string(""));
            m.set("submesh", mesh->name);
            o.set("mesh", m);
        }
        if (so->has<ComponentMaterial>() && so->get<ComponentMaterial>().material()) {
            Value m = Value::object();
            m.set("name", so->get<ComponentMaterial>().material()->name());
            o.set("material", m);
        }
        if (so->has<ComponentLight>() && so->get<ComponentLight>().light()) {
            Value l = Value::object();
            l.set("name", so->get<ComponentLight>().light()->name());
            l.set("shadows", so->get<ComponentLight>().castShadows());
            o.set("light", l);
        }
        if (so->has<ComponentVolume>()) {
            ComponentVolume &vc = so->get<ComponentVolume>();
            Value v = Value::object();
            if (vc.frontFacing()) v.set("frontFacing", vc.frontFacing()->name());
            if (vc.backFacing()) v.set("backFacing", vc.backFacing()->name());
            o.set("volume", v);
        }
        if (!so->children().empty()) {
            Value children = Value::array();
            for (SceneObject *c : so->children()) writeObject(children, c);
            o.set("children", children);
        }
        array.push(o);
    };
    Value scene = Value::array();
    for (SceneObject *root : m_scene->sceneGraph()) writeObject(scene, root);

    Value models = Value::array();
    for (auto &it : m_models.all()) {
        Model3D *model = it.second;
        if (model->internal || model->filepath.empty()) continue; /* engine assets and procedural meshes have no file */
        Value m = Value::object();
        m.set("name", model->name);
        m.set("filepath", "assets/models/" + copyInto(model->filepath, modelsFolder));
        models.push(m);
    }

    Value materials = Value::array();
    for (auto &it : m_materials->all()) {
        Material *material = it.second;
        const ptc_material &b = material->block();
        Value m = Value::object();
        m.set("name", material->name());
        const std::string matDir = texturesFolder + material->name() + "/", matPrefix = "assets/textures/" + material->name() + "/";
        bool dirMade = false;
        /* {"texture": {...}, "value": ...}; engine default textures (slots 0-2) are not written */
        auto slotObject = [&](uint32_t slot, const std::string &finalName) {
            Value o = Value::object();
            Texture *tex = slot > 2 ? m_textures->bySlot(slot) : nullptr;
            if (tex) {
                Value t = Value::object();
                if (tex->embedded || tex->filepath.empty()) {
                    t.set("type", "EMBEDDED");
                    t.set("name", tex->name);
                } else {
                    t.set("type", "STANDALONE");
                    t.set("name", tex->name);
                    if (!dirMade) dirMade = makeDir(matDir);
                    size_t dot = tex->filepath.find_last_of('.');
                    std::string ext = dot == std::string::npos ? std::string("png") : tex->filepath.substr(dot + 1);
                    t.set("filepath", matPrefix + copyInto(tex->filepath, matDir, finalName + "." + ext));
                }
                o.set("texture", t);
            }
            return o;
        };
        switch (material->type()) {
            case MaterialType::MATERIAL_LAMBERT:
            case MaterialType::MATERIAL_PBR_STANDARD: {
                bool pbr = material->type() == MaterialType::MATERIAL_PBR_STANDARD;
                if (pbr && material->isEmbedded()) { /* comes back with its model file (Export.cpp:347-352) */
                    m.set("type", "EMBEDDED");
                    break;
                }
                m.set("type", pbr ? "PBR_STANDARD" : "LAMBERT");
                Value albedo = slotObject(b.tex1[0], "albedo");
                albedo.set("value", vecValue({b.albedo[0], b.albedo[1], b.albedo[2], b.albedo[3]}));
                m.set("albedo", albedo);
                if (pbr) {
                    Value roughness = slotObject(b.tex1[2], "roughness");
                    roughness.set("value", Value(b.metallic_roughness_ao[1]));
                    m.set("roughness", roughness);
                    Value metallic = slotObject(b.tex1[1], "metallic");
                    metallic.set("value", Value(b.metallic_roughness_ao[0]));
                    m.set("metallic", metallic);
                    Value ao = slotObject(b.tex1[3], "ao");
                    ao.set("value", Value(b.metallic_roughness_ao[2]));
                    m.set("ao", ao);
                }
                Value emissive = slotObject(b.tex2[0], "emissive");
                emissive.set("value", vecValue({b.emissive[0], b.emissive[1], b.emissive[2], b.emissive[3]}));
                m.set("emissive", emissive);
                m.set("normal", slotObject(b.tex2[1], "normal"));
                m.set("alpha", slotObject(b.tex2[3], "alpha"));
                m.set("transparent", material->isTransparent());
                m.set("scale", vecValue({b.uv_tiling[0], b.uv_tiling[1]}));
                break;
            }
            case MaterialType::MATERIAL_VOLUME: {
                m.set("type", "VOLUME");
                Value s = Value::object(), a = Value::object();
                s.set("value", vecValue({b.metallic_roughness_ao[0], b.metallic_roughness_ao[1], b.metallic_roughness_ao[2]}));
                a.set("value", vecValue({b.albedo[0], b.albedo[1], b.albedo[2]}));
                m.set("scattering", s);
                m.set("absorption", a);
                m.set("g", Value(b.emissive[0]));
                break;
            }
            default: continue;
        }
        materials.push(m);
    }

    Value lights = Value::array();
    for (auto &it : m_lightsMap.all()) {
        Light *light = it.second;
        if (light->type() != LightType::POINT_LIGHT && light->type() != LightType::DIRECTIONAL_LIGHT) continue;
        Value l = Value::object();
        l.set("name", light->name());
        l.set("type", light->type() == LightType::POINT_LIGHT ? "POINT" : "DIRECTIONAL");
        vec4 c = light->color();
        l.set("color", vecValue({c.x, c.y, c.z}));
        l.set("intensity", Value(c.w));
        lights.push(l);
    }

    Value environment = Value::object();
    if (m_scene->skyboxMaterial() && !m_scene->skyboxMaterial()->filepath.empty())
        environment.set("path", "assets/" + copyInto(m_scene->skyboxMaterial()->filepath, assetsFolder));
    environment.set("environmentType", (int)m_scene->environmentType());
    vec3 bg = m_scene->backgroundColor();
    environment.set("backgroundColor", vecValue({bg.x, bg.y, bg.z}));

    d.set("camera", cam);
    d.set("scene", scene);
    d.set("models", models);
    d.set("materials", materials);
    d.set("lights", lights);
    d.set("environment", environment);

    std::ofstream of(sceneFolder + "scene.json");
    if (!of) return failWith("cannot write " + sceneFolder + "scene.json");
    of << json::dump(d, 4) << "\n";
    return true;
}

}  // namespace vengine
