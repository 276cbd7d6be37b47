/*
 * Scene recipes — see scenes.hpp.
 */
#include "scenes.hpp"

#include <cmath>
#include <cstdio>
#include <functional>
#include <random>

namespace vengine {
namespace scenes {

using RI = RendererPathTracing::RenderInfo;

static RI &info(Engine &e) { return e.renderer().rendererPathTracing().renderInfo(); }

static void testRenderInfo(Engine &e, const std::string &name, uint32_t samples, uint32_t batch) {
    RI &ri = info(e);
    ri = RI();
    ri.filename = name + "_test";
    ri.width = 256;
    ri.height = 256;
    ri.samples = samples;
    ri.batchSize = batch;
    ri.fileType = FileType::HDR;
    ri.denoise = false;
    ri.writeAllFiles = false;
}

struct BaseMeshes {
    Mesh *sphere, *plane, *cube;
};
static BaseMeshes baseMeshes(Engine &e) {
    BaseMeshes b;
    b.sphere = e.modelsMap().get("assets/models/uvsphere.obj")->mesh("defaultobject");
    b.plane = e.modelsMap().get("assets/models/plane.obj")->mesh("Plane");
    b.cube = e.modelsMap().get("assets/models/cube.obj")->mesh("Cube");
    return b;
}

static SceneObject *addMeshObject(Scene &scene, const std::string &name, Transform t, Mesh *mesh, Material *mat, SceneObject *parent = nullptr) {
    SceneObject *so = scene.addSceneObject(name, parent, t);
    so->add<ComponentMesh>().setMesh(mesh);
    so->add<ComponentMaterial>().setMaterial(mat);
    return so;
}

static std::shared_ptr<PerspectiveCamera> makeCamera(Scene &scene, vec3 pos, vec3 eulerRad) {
    auto camera = std::make_shared<PerspectiveCamera>();
    camera->transform().position() = pos;
    camera->transform().setRotation(quat(eulerRad));
    scene.camera() = camera;
    return camera;
}

/* ====================================================================== RenderTests.cpp recipes */

/* RenderTests.cpp:74-153 (FurnacePBR / FurnaceLambert) */
static void furnace(Engine &e, bool lambert) {
    Scene &scene = e.scene();
    auto camera = makeCamera(scene, vec3(0, 0, 2), vec3(0, 0, 0));
    camera->fov() = 60.0f;
    camera->lensRadius() = 0.0f;
    camera->focalDistance() = 0.0f;
    BaseMeshes bm = baseMeshes(e);
    Material *mat;
    if (lambert) {
        auto m = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial1"));
        m->albedo() = vec4(0.6f, 0.6f, 0.6f, 1);
        mat = m;
    } else {
        auto m = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial1"));
        m->albedo() = vec4(0.6f, 0.6f, 0.6f, 1);
        mat = m;
    }
    addMeshObject(scene, "sphere", Transform({0, 0, 0}, {1, 1, 1}), bm.sphere, mat);
    scene.environmentIntensity() = 1.0f;
    scene.environmentType() = EnvironmentType::SOLID_COLOR;
    scene.backgroundColor() = vec3(1, 1, 1);
    scene.update();
    testRenderInfo(e, lambert ? "FurnaceLambert" : "FurnacePBR", 2048, 64);
}

/* RenderTests.cpp:156-205 */
static void environmentMap(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 1, 4), vec3(0, 0, 0));
    BaseMeshes bm = baseMeshes(e);
    auto pbr = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial2"));
    auto lam = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial2"));
    addMeshObject(scene, "sphere", Transform({0, 1, 0}, {1, 1, 1}), bm.sphere, pbr);
    addMeshObject(scene, "plane", Transform({0, 0, 0}, {10, 10, 10}), bm.plane, lam);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    testRenderInfo(e, "EnvironmentMap", 1024, 64);
}

/* RenderTests.cpp:207-267; variant = roughness*2 + metallic encoded as in the golden names RM */
static void environmentMapPBR(Engine &e, int roughness, int metallic) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 0, 3), vec3(0, 0, 0));
    BaseMeshes bm = baseMeshes(e);
    auto pbr = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial3"));
    addMeshObject(scene, "sphere", Transform({0, 0, 0}, {1, 1, 1}), bm.sphere, pbr);
    scene.environmentType() = EnvironmentType::HDRI; /* inherited from the previous test in the reference suite */
    scene.environmentIntensity() = 1.0f;
    scene.update();
    pbr->roughness() = (float)roughness;
    pbr->metallic() = (float)metallic;
    testRenderInfo(e, std::string("EnvironmentMapPBR") + char('0' + roughness) + char('0' + metallic), 1024, 64);
}

/* RenderTests.cpp:269-307 */
static void environmentMapLambert(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 0, 3), vec3(0, 0, 0));
    BaseMeshes bm = baseMeshes(e);
    auto lam = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial3"));
    addMeshObject(scene, "sphere", Transform({0, 0, 0}, {1, 1, 1}), bm.sphere, lam);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    lam->albedo() = vec4(0.1f, 0.3f, 0.5f, 1.0f);
    testRenderInfo(e, "EnvironmentMapLambert", 1024, 64);
}

/* RenderTests.cpp:309-439 (Volume1 -> goldens Volume0..5) and :441-565 (Volume2 -> goldens Volume6..9) */
static void volume(Engine &e, int k) {
    Scene &scene = e.scene();
    BaseMeshes bm = baseMeshes(e);
    const bool second = k >= 6;
    auto camera = makeCamera(scene, second ? vec3(0, 1.0f, 7) : vec3(0, 0.5f, 4), vec3(0, 0, 0));
    if (second) camera->zfar() = 10.0f;

    auto pbr = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo(second ? "pbrmaterial5" : "pbrmaterial4"));
    pbr->setTransparent(true);
    pbr->albedo() = vec4(1, 1, 1, 0.2f);
    pbr->metallic() = 0;
    pbr->roughness() = 0;
    auto vol1 = e.materials().createMaterial<MaterialVolume>(AssetInfo(second ? "volumematerial2" : "volumematerial1"));
    MaterialVolume *vol2 = nullptr;
    if (second) vol2 = e.materials().createMaterial<MaterialVolume>(AssetInfo("volumematerial3"));

    SceneObject *cube = addMeshObject(scene, "cube1", Transform({0, 0.5f, 0}, {1, 1, 1}), bm.cube, pbr);
    cube->add<ComponentVolume>().setBackFacingVolume(vol1);
    if (second) cube->add<ComponentVolume>().setFrontFacingVolume(vol2);
    addMeshObject(scene, "plane", Transform({0, -1, 0}, {10, 10, 10}), bm.plane, e.materials().get("defaultMaterial"));

    if (!second) {
        vol1->sigmaA() = vec4(0.2f, 0.4f, 0.8f, 0.0f);
        vol1->sigmaS() = vec4(0.8f, 0.4f, 0.2f, 0.0f);
        vol1->g() = 0;
    } else {
        vol1->sigmaA() = vec4(0.2f, 0.4f, 0.0f, 0.0f);
        vol1->sigmaS() = vec4(0.8f, 0.4f, 0.2f, 0.0f);
        vol2->sigmaA() = vec4(0.04f, 0.04f, 0.04f, 0.0f);
        vol2->sigmaS() = vec4(0.04f, 0.04f, 0.04f, 0.0f);
        vol1->g() = 0;
        vol2->g() = 0;
        camera->volume() = vol2;
    }
    scene.environmentIntensity() = 1.0f;
    scene.environmentType() = EnvironmentType::HDRI;
    const int stage = second ? (k - 6) : (k <= 2 ? 0 : k - 2); /* 0 env, 1 directional, 2 point, 3 mesh light */
    if (!second) {
        if (k == 1) vol1->g() = -0.99f;
        if (k == 2) vol1->g() = 0.99f;
    }
    if (stage >= 1) {
        scene.environmentType() = EnvironmentType::SOLID_COLOR;
        scene.backgroundColor() = vec3(0, 0, 0);
        SceneObject *soLight = scene.addSceneObject("light", Transform({0, 4, 0}, {1, 1, 1}, vec3(vm::radians(45.0f), vm::radians(90.0f), 0)));
        if (stage == 1) {
            Light *dl = scene.createLight(AssetInfo("Directional light 1"), LightType::DIRECTIONAL_LIGHT);
            dl->color().w = 1;
            soLight->add<ComponentLight>().setLight(dl);
        } else if (stage == 2) {
            Light *pl = scene.createLight(AssetInfo("Point light 1"), LightType::POINT_LIGHT);
            pl->color().w = 15;
            soLight->add<ComponentLight>().setLight(pl);
        } else {
            auto em = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo(second ? "emissivepbrmaterial2" : "emissivepbrmaterial1"));
            em->emissive() = vec4(0.3f, 0.3f, 0.7f, 15.0f);
            soLight->setLocalTransform(Transform({0, 1.5f, -3}, {1, 1, 1}, vec3(vm::radians(90.0f), 0, 0)));
            soLight->add<ComponentMesh>().setMesh(bm.plane);
            soLight->add<ComponentMaterial>().setMaterial(em);
        }
    }
    scene.update();
    testRenderInfo(e, std::string("Volume") + char('0' + k), 2048, 64);
}

/* RenderTests.cpp:35-72 */
static void prepareTwoBallsOnPlaneScene(Engine &e, Material *floorMaterial, Material *ball1, Material *ball2) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 1, 4), vec3(0, 0, 0));
    BaseMeshes bm = baseMeshes(e);
    addMeshObject(scene, "sphere", Transform({-1, 1, 0}, {1, 1, 1}), bm.sphere, ball1);
    addMeshObject(scene, "sphere", Transform({1, 1, 0}, {1, 1, 1}), bm.sphere, ball2);
    addMeshObject(scene, "plane", Transform({0, 0, 0}, {10, 10, 10}), bm.plane, floorMaterial);
    scene.environmentIntensity() = 1.0f;
}

/* RenderTests.cpp:567-689: kind 0 PointLight, 1 DirectionalLight, 2 MeshLight */
static void lights(Engine &e, int kind) {
    Scene &scene = e.scene();
    auto pbr = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial6"));
    auto lam = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial6"));
    auto floorM = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial7"));
    pbr->albedo() = vec4(0.7f, 0.1f, 0.1f, 1.0f);
    pbr->roughness() = 0.3f;
    lam->albedo() = vec4(0.1f, 0.7f, 0.1f, 1.0f);
    prepareTwoBallsOnPlaneScene(e, floorM, pbr, lam);
    scene.environmentType() = EnvironmentType::SOLID_COLOR;
    scene.backgroundColor() = vec3(0, 0, 0);
    if (kind == 0) {
        Light *pl = scene.createLight(AssetInfo("Point light"), LightType::POINT_LIGHT);
        pl->color().w = 10;
        scene.addSceneObject("Point light", Transform({0, 4, 0}))->add<ComponentLight>().setLight(pl);
    } else if (kind == 1) {
        Light *dl = scene.createLight(AssetInfo("Directional light"), LightType::DIRECTIONAL_LIGHT);
        dl->color().w = 1;
        scene.addSceneObject("Directional light", Transform({0, 0, 0}, {1, 1, 1}, vec3(vm::radians(45.0f), vm::radians(90.0f), 0)))
            ->add<ComponentLight>()
            .setLight(dl);
    } else {
        BaseMeshes bm = baseMeshes(e);
        auto em = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("emissivepbrmaterial3"));
        em->emissive() = vec4(0.3f, 0.3f, 0.7f, 15.0f);
        addMeshObject(scene, "Mesh light", Transform({0, 2, 0}, {0.4f, 0.4f, 0.4f}, vec3(vm::radians(90.0f), 0, 0)), bm.plane, em);
        scene.camera()->transform().position() = vec3(0, 3, 5);
        scene.camera()->transform().setRotation(quat(vec3(vm::radians(-20.0f), 0, 0)));
    }
    scene.update();
    static const char *names[3] = {"PointLight", "DirectionalLight", "MeshLight"};
    testRenderInfo(e, names[kind], 2048, 64);
}

/* RenderTests.cpp:691-742 */
static void transparency(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 2, 4), vec3(vm::radians(-20.0f), 0, 0));
    BaseMeshes bm = baseMeshes(e);
    auto m1 = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial7"));
    m1->setTransparent(true);
    m1->albedo() = vec4(0.7f, 0.7f, 0.7f, 0.5f);
    addMeshObject(scene, "sphere", Transform({0, 0, 0}, {1, 1, 1}), bm.sphere, m1);
    auto m2 = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial8"));
    Texture *tex = e.textures().createTexture(AssetInfo("checkerboard", e.assetPath("assets/textures/checkerboard.png")));
    if (tex) m2->setAlphaTexture(tex);
    m2->setTransparent(true);
    addMeshObject(scene, "plane", Transform({0, 0, 0}, {5, 5, 5}), bm.plane, m2);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    testRenderInfo(e, "Transparency", 2048, 64);
}

/* RenderTests.cpp:744-786 */
static void normalMap(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 3, 8), vec3(vm::radians(-20.0f), 0, 0));
    BaseMeshes bm = baseMeshes(e);
    auto m = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial9"));
    Texture *tex = e.textures().createTexture(AssetInfo("normal", e.assetPath("assets/textures/normal.png")), ColorSpace::LINEAR);
    if (tex) m->setNormalTexture(tex);
    addMeshObject(scene, "plane", Transform({0, 0, 0}, {5, 5, 5}), bm.plane, m);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    testRenderInfo(e, "NormalMap", 2048, 64);
}

/* RenderTests.cpp:833-903 */
static void hierarchy(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 3, 12), vec3(vm::radians(-20.0f), 0, 0));
    BaseMeshes bm = baseMeshes(e);
    Material *def = e.materials().get("defaultMaterial");
    addMeshObject(scene, "root1", Transform({0, -8, 0}, {100, 100, 100}), bm.plane, def);
    SceneObject *root2 = addMeshObject(scene, "root2", Transform({1, 0, 0}, {1.1f, 1.1f, 1.1f}, vec3(0, vm::radians(15.0f), 0)), bm.cube, def);
    SceneObject *l1_1 = addMeshObject(scene, "l1_1", Transform({0, 3, 0}, {0.5f, 0.5f, 0.5f}), bm.cube, def, root2);
    SceneObject *l1_2 = addMeshObject(scene, "l1_2", Transform({0, -4, 0}, {1, 1, 1}, vec3(vm::radians(20.0f), 0, 0)), bm.cube, def, root2);
    SceneObject *l2_1 = addMeshObject(scene, "l2_1", Transform({3, 0, 4}, {1, 1, 1}, vec3(vm::radians(45.0f), vm::radians(90.0f), 0)), bm.cube, def, l1_1);
    addMeshObject(scene, "l2_2", Transform({-3, 0, -4}), bm.cube, def, l1_2);
    Light *dl = scene.createLight(AssetInfo("Directional light"), LightType::DIRECTIONAL_LIGHT);
    dl->color() = vec4(1, 1, 1, 1);
    scene.addSceneObject("Directional light", l2_1, Transform())->add<ComponentLight>().setLight(dl);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.backgroundColor() = vec3(0, 0, 0);
    scene.environmentIntensity() = 0.5f;
    scene.update();
    testRenderInfo(e, "Hierarchy", 2048, 64);
}

/* RenderTests.cpp:788-831: the glTF import path (DamagedHelmet: embedded buffers, 5 JPEG textures, emissive + normal + AO maps) */
static void gltf(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 0.5f, 2), vec3(vm::radians(-20.0f), 0, 0));
    BaseMeshes bm = baseMeshes(e);
    addMeshObject(scene, "plane", Transform({0, -1, 0}, {5, 5, 5}), bm.plane, e.materials().get("defaultMaterial"));
    const std::string assetName = "assets/models/DamagedHelmet.gltf";
    if (!e.importModel(AssetInfo(assetName), true)) throw std::runtime_error("GLTF scene: cannot import " + assetName);
    addModel3D(scene, nullptr, assetName, std::nullopt, std::nullopt);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    testRenderInfo(e, "GLTF", 2048, 64);
}

/* RenderTests.cpp:905-967 */
static void depthOfField(Engine &e) {
    Scene &scene = e.scene();
    auto camera = makeCamera(scene, vec3(0, 1, 8), vec3(0, 0, 0));
    camera->focalDistance() = 8.0f;
    camera->lensRadius() = 0.4f;
    camera->fov() = 60.0f;
    BaseMeshes bm = baseMeshes(e);
    auto pbr = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial10"));
    auto lam = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial8"));
    addMeshObject(scene, "sphere1", Transform({0, 1, 0}, {1, 1, 1}), bm.sphere, pbr);
    addMeshObject(scene, "sphere2", Transform({-2, 1, -2}, {1, 1, 1}), bm.sphere, pbr);
    addMeshObject(scene, "sphere3", Transform({2, 1, 2}, {1, 1, 1}), bm.sphere, pbr);
    addMeshObject(scene, "plane", Transform({0, 0, 0}, {10, 10, 10}), bm.plane, lam);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    testRenderInfo(e, "DepthOfField", 2048, 4);
}

/* src/bin/offlinerender/PtSceneBallOnPlane.cpp:8-32 (create) */
static void ballOnPlane(Engine &e) {
    Scene &scene = e.scene();
    auto camera = makeCamera(scene, vec3(0, 3, 10), vec3(vm::radians(-15.0f), 0, 0));
    camera->fov() = 60.0f;
    camera->lensRadius() = 0.03f;
    camera->focalDistance() = 7.0f;
    BaseMeshes bm = baseMeshes(e);
    Material *def = e.materials().get("defaultMaterial");
    addMeshObject(scene, "plane", Transform({0, -3, 0}, {10, 10, 10}), bm.plane, def);
    addMeshObject(scene, "uvsphere", Transform({0, 0, 0}, {3, 3, 3}), bm.sphere, def);
    scene.addSceneObject("light", Transform({4, 1.5f, 0}))->add<ComponentLight>().setLight(e.lightsMap().get("defaultPointLight"));
    scene.environmentIntensity() = 1.0f;
    scene.update();
    /* PtSceneBallOnPlane.cpp:38-42 (the reference also denoises, which is out of scope: OIDN) */
    RI &ri = info(e);
    ri = RI();
    ri.filename = "0";
    ri.samples = 64;
    ri.batchSize = 64;
    ri.fileType = FileType::PNG;
    ri.denoise = false;
    ri.writeAllFiles = false;
}

/* PtSceneBallOnPlane.cpp:44-55: frame i of the render sequence - the camera orbits the ball in 45 degree steps */
void ballOnPlaneFrame(Engine &e, int frame) {
    const float height = 1.0f, radius = 10.0f, a = vm::radians(45.0f) * (float)frame;
    e.renderer().rendererPathTracing().renderInfo().filename = std::to_string(frame);
    e.scene().camera()->transform().position() = vec3(radius * std::sin(a), height, radius * std::cos(a));
    e.scene().camera()->transform().setRotationEuler(0, a, 0);
}

/* RenderTests.cpp:969-1031 */
static void sharedComponents(Engine &e) {
    Scene &scene = e.scene();
    auto camera = makeCamera(scene, vec3(0, 20, 50), vec3(vm::radians(-15.0f), vm::radians(90.0f), 0));
    camera->fov() = 60.0f;
    BaseMeshes bm = baseMeshes(e);
    Material *def = e.materials().get("defaultMaterial");
    scene.addSceneObject("root", nullptr, Transform());
    ComponentManager &cm = ComponentManager::getInstance();
    ComponentMesh *meshComponent = cm.create<ComponentMesh, ComponentOwnerShared>();
    ComponentMaterial *materialComponent = cm.create<ComponentMaterial, ComponentOwnerShared>();
    int32_t size = 150;
    for (int32_t i = -size; i < size; i += 3)
        for (int32_t j = -size; j < size; j += 3) {
            SceneObject *so = scene.addSceneObject("cube" + std::to_string(i) + std::to_string(j), nullptr, Transform({(float)i, 0, (float)j}, {1, 1, 1}));
            so->add_shared<ComponentMesh>(meshComponent);
            so->add_shared<ComponentMaterial>(materialComponent);
        }
    meshComponent->setMesh(bm.cube);
    materialComponent->setMaterial(def);
    scene.addSceneObject("directionalLight", nullptr, Transform({0, 2, 0}, {1, 1, 1}, vec3(vm::radians(45.0f), vm::radians(90.0f), 0)))
        ->add<ComponentLight>()
        .setLight(e.lightsMap().get("defaultDirectionalLightSun"));
    scene.environmentType() = EnvironmentType::SOLID_COLOR;
    scene.backgroundColor() = vec3(0, 0, 0); /* left over from the previous test in the reference suite */
    scene.environmentIntensity() = 0.0f;
    scene.update();
    testRenderInfo(e, "SharedComponents", 2048, 4);
}

/* RenderTests.cpp:1033-1087 */
static void denoise(Engine &e) {
    Scene &scene = e.scene();
    makeCamera(scene, vec3(0, 1, 4), vec3(0, 0, 0));
    BaseMeshes bm = baseMeshes(e);
    auto pbr = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("pbrmaterial11"));
    auto lam = e.materials().createMaterial<MaterialLambert>(AssetInfo("lambertmaterial9"));
    addMeshObject(scene, "sphere1", Transform({0, 1, 0}, {1, 1, 1}), bm.sphere, pbr);
    addMeshObject(scene, "plane", Transform({0, 0, 0}, {10, 10, 10}), bm.plane, lam);
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    scene.update();
    testRenderInfo(e, "Denoise", 1024, 4);
    info(e).denoise = true;
    info(e).writeAllFiles = true;
}

/* ====================================================================== procedural geometry */
namespace {

struct MeshBuilder {
    std::unique_ptr<Mesh> mesh;
    explicit MeshBuilder(const std::string &name) : mesh(std::make_unique<Mesh>()) { mesh->name = name; }
    uint32_t vertex(vec3 p, vec3 n, vec2 uv) {
        Vertex v{};
        v.position[0] = p.x; v.position[1] = p.y; v.position[2] = p.z;
        v.normal[0] = n.x; v.normal[1] = n.y; v.normal[2] = n.z;
        v.uv[0] = uv.x; v.uv[1] = uv.y;
        v.color[0] = v.color[1] = v.color[2] = 1.0f;
        mesh->vertices.push_back(v);
        return (uint32_t)mesh->vertices.size() - 1;
    }
    void tri(uint32_t a, uint32_t b, uint32_t c) {
        mesh->indices.push_back(a);
        mesh->indices.push_back(b);
        mesh->indices.push_back(c);
    }
    /* parametric surface p(u,v) on an nu x nv grid, smooth normals by finite differences */
    /* dropDegenerate: leave out triangles with coincident corners (the poles of a sphere).  An EMISSIVE mesh must not contain them: the
     * light sampler picks triangles uniformly and divides by their area (lightSampling.glsl:44-100), so a zero-area triangle gives an
     * infinite pdf and the power heuristic returns NaN - in the reference exactly as here. */
    void surface(int nu, int nv, const std::function<vec3(float, float)> &p, bool flip = false, float uvScaleU = 1, float uvScaleV = 1,
                 bool dropDegenerate = false) {
        uint32_t base = (uint32_t)mesh->vertices.size();
        const float eps = 1e-3f;
        for (int j = 0; j <= nv; j++)
            for (int i = 0; i <= nu; i++) {
                float u = (float)i / nu, v = (float)j / nv;
                vec3 P = p(u, v);
                vec3 du = p(std::min(u + eps, 1.0f), v) - p(std::max(u - eps, 0.0f), v);
                vec3 dv = p(u, std::min(v + eps, 1.0f)) - p(u, std::max(v - eps, 0.0f));
                vec3 n = vm::cross(du, dv);
                float l = vm::length(n);
                n = l > 0 ? n / l : vec3(0, 1, 0);
                if (flip) n = -n;
                vertex(P, n, vec2(u * uvScaleU, v * uvScaleV));
            }
        for (int j = 0; j < nv; j++)
            for (int i = 0; i < nu; i++) {
                uint32_t a = base + j * (nu + 1) + i, b = a + 1, c = a + (nu + 1), d = c + 1;
                auto put = [&](uint32_t x, uint32_t y, uint32_t z) {
                    if (dropDegenerate) {
                        auto pos = [&](uint32_t k) { const float *q = mesh->vertices[k].position; return vec3(q[0], q[1], q[2]); };
                        const vec3 px = pos(x), py = pos(y), pz = pos(z);
                        const float e2 = std::max(vm::dot(py - px, py - px), std::max(vm::dot(pz - px, pz - px), vm::dot(pz - py, pz - py)));
                        if (vm::length(vm::cross(py - px, pz - px)) <= 1e-5f * e2) return; /* incl. sin(pi) != 0 in float at the far pole */
                    }
                    tri(x, y, z);
                };
                if (!flip) {
                    put(a, b, d);
                    put(a, d, c);
                } else {
                    put(a, d, b);
                    put(a, c, d);
                }
            }
    }
    /* axis-aligned box, flat normals, 12 triangles */
    void box(vec3 lo, vec3 hi) {
        vec3 c[8];
        for (int i = 0; i < 8; i++) c[i] = vec3((i & 1) ? hi.x : lo.x, (i & 2) ? hi.y : lo.y, (i & 4) ? hi.z : lo.z);
        static const int faces[6][4] = {{0, 2, 3, 1}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 6, 7, 3}, {0, 4, 6, 2}, {1, 3, 7, 5}};
        static const float normals[6][3] = {{0, 0, -1}, {0, 0, 1}, {0, -1, 0}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0}};
        for (int f = 0; f < 6; f++) {
            vec3 n(normals[f][0], normals[f][1], normals[f][2]);
            uint32_t a = vertex(c[faces[f][0]], n, vec2(0, 0)), b = vertex(c[faces[f][1]], n, vec2(1, 0));
            uint32_t cc = vertex(c[faces[f][2]], n, vec2(1, 1)), d = vertex(c[faces[f][3]], n, vec2(0, 1));
            tri(a, b, cc);
            tri(a, cc, d);
        }
    }
    std::unique_ptr<Mesh> finish() {
        computeTangents(*mesh);
        return std::move(mesh);
    }
};

/* hash-based value noise, deterministic */
inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
inline float hashf(int x, int y, uint32_t seed) { return (hash32((uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ seed * 83492791u) & 0xffffff) / 16777215.0f; }
inline float vnoise(float x, float y, uint32_t seed, int period) {
    int xi = (int)std::floor(x), yi = (int)std::floor(y);
    float fx = x - xi, fy = y - yi;
    fx = fx * fx * (3 - 2 * fx);
    fy = fy * fy * (3 - 2 * fy);
    auto w = [&](int a) { return ((a % period) + period) % period; };
    float a = hashf(w(xi), w(yi), seed), b = hashf(w(xi + 1), w(yi), seed);
    float c = hashf(w(xi), w(yi + 1), seed), d = hashf(w(xi + 1), w(yi + 1), seed);
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy;
}
inline float fbm(float u, float v, uint32_t seed, int baseFreq) {
    float s = 0, amp = 0.5f;
    int f = baseFreq;
    for (int o = 0; o < 4; o++) {
        s += amp * vnoise(u * f, v * f, seed + o * 131u, f);
        amp *= 0.5f;
        f *= 2;
    }
    return s / 0.9375f;
}

struct TexSet {
    Texture *albedo, *normal, *roughness;
};
/* tileable procedural PBR texture set: albedo (sRGB), normal (linear, from a height field), roughness (R8) */
TexSet makeTextures(Engine &e, const std::string &name, int size, uint32_t seed, vec3 colA, vec3 colB, int pattern) {
    ImageU8 alb, nrm, rgh;
    alb.width = alb.height = nrm.width = nrm.height = rgh.width = rgh.height = size;
    alb.channels = nrm.channels = 4;
    rgh.channels = 1;
    alb.data.resize((size_t)size * size * 4);
    nrm.data.resize((size_t)size * size * 4);
    rgh.data.resize((size_t)size * size);
    std::vector<float> height((size_t)size * size);
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            float u = (x + 0.5f) / size, v = (y + 0.5f) / size;
            float n = fbm(u, v, seed, 4);
            float h = n;
            float mixv = n;
            if (pattern == 1) { /* tiles / bricks with mortar */
                float bu = u * 8, bv = v * 16;
                int row = (int)std::floor(bv);
                bu += (row & 1) ? 0.5f : 0.0f;
                float fu = bu - std::floor(bu), fv = bv - std::floor(bv);
                float edge = std::min(std::min(fu, 1 - fu) * 8.0f / 16.0f, std::min(fv, 1 - fv));
                float m = std::min(edge * 12.0f, 1.0f);
                h = 0.7f * m + 0.3f * n;
                mixv = 0.5f * n + 0.5f * hashf((int)std::floor(bu) % 8, row % 16, seed + 7u);
                if (m < 0.5f) mixv *= 0.4f;
            } else if (pattern == 2) { /* woven fabric */
                float w = 0.5f + 0.25f * std::sin(u * 6.2831853f * 64) + 0.25f * std::sin(v * 6.2831853f * 64);
                h = 0.6f * w + 0.4f * n;
                mixv = 0.7f * n + 0.3f * w;
            } else if (pattern == 3) { /* checker marble */
                int cx = (int)std::floor(u * 8), cy = (int)std::floor(v * 8);
                float ch = ((cx + cy) & 1) ? 1.0f : 0.0f;
                float vein = std::fabs(std::sin((u + v) * 20.0f + 6.0f * n));
                mixv = 0.75f * ch + 0.25f * vein;
                h = 0.9f + 0.1f * n;
            }
            height[(size_t)y * size + x] = h;
            vec3 c = colA * (1.0f - mixv) + colB * mixv;
            size_t o = ((size_t)y * size + x) * 4;
            for (int k = 0; k < 3; k++) alb.data[o + k] = (uint8_t)std::min(255.0f, std::max(0.0f, 255.0f * linearToSRGB(std::min(std::max(c[k], 0.0f), 1.0f))));
            alb.data[o + 3] = 255;
            rgh.data[(size_t)y * size + x] = (uint8_t)(255.0f * std::min(1.0f, std::max(0.08f, 0.35f + 0.6f * fbm(u, v, seed + 977u, 8))));
        }
    const float strength = 2.0f;
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            auto H = [&](int a, int b) { return height[(size_t)((b + size) % size) * size + ((a + size) % size)]; };
            float dx = (H(x + 1, y) - H(x - 1, y)) * 0.5f * size / 64.0f * strength;
            float dy = (H(x, y + 1) - H(x, y - 1)) * 0.5f * size / 64.0f * strength;
            vec3 n = vm::normalize(vec3(-dx, -dy, 1.0f));
            size_t o = ((size_t)y * size + x) * 4;
            nrm.data[o + 0] = (uint8_t)(255.0f * (n.x * 0.5f + 0.5f));
            nrm.data[o + 1] = (uint8_t)(255.0f * (n.y * 0.5f + 0.5f));
            nrm.data[o + 2] = (uint8_t)(255.0f * (n.z * 0.5f + 0.5f));
            nrm.data[o + 3] = 255;
        }
    TexSet t;
    t.albedo = e.textures().createTexture(name + "_albedo", alb, ColorSpace::sRGB);
    t.normal = e.textures().createTexture(name + "_normal", nrm, ColorSpace::LINEAR);
    t.roughness = e.textures().createTexture(name + "_roughness", rgh, ColorSpace::LINEAR);
    return t;
}

MaterialPBRStandard *texturedMaterial(Engine &e, const std::string &name, int texSize, uint32_t seed, vec3 a, vec3 b, int pattern, float metallic, float tiling) {
    auto m = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo(name));
    TexSet t = makeTextures(e, name, texSize, seed, a, b, pattern);
    m->albedo() = vec4(1, 1, 1, 1);
    m->metallic() = metallic;
    m->roughness() = 1.0f;
    m->setAlbedoTexture(t.albedo);
    m->setNormalTexture(t.normal);
    m->setRoughnessTexture(t.roughness);
    m->uTiling() = tiling;
    m->vTiling() = tiling;
    return m;
}

Mesh *registerMesh(Engine &e, const std::string &modelName, std::unique_ptr<Mesh> mesh) {
    auto model = std::make_unique<Model3D>();
    model->name = modelName;
    Mesh *raw = mesh.get();
    model->meshes.push_back(std::move(mesh));
    e.addModel(std::move(model));
    return raw;
}

const float TWO_PI = 6.28318530718f;

}  // namespace

/* ---------------------------------------------------------------------- C1 Cornell */
static void cornell(Engine &e) {
    Scene &scene = e.scene();
    auto camera = makeCamera(scene, vec3(0, 1, 3.75f), vec3(0, 0, 0));
    camera->fov() = 40.0f;
    BaseMeshes bm = baseMeshes(e);
    auto white = e.materials().createMaterial<MaterialLambert>(AssetInfo("cornellWhite"));
    white->albedo() = vec4(0.73f, 0.73f, 0.73f, 1);
    auto red = e.materials().createMaterial<MaterialLambert>(AssetInfo("cornellRed"));
    red->albedo() = vec4(0.65f, 0.05f, 0.05f, 1);
    auto green = e.materials().createMaterial<MaterialLambert>(AssetInfo("cornellGreen"));
    green->albedo() = vec4(0.12f, 0.45f, 0.15f, 1);
    auto metal = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("cornellMetal"));
    metal->albedo() = vec4(0.9f, 0.9f, 0.9f, 1);
    metal->metallic() = 0.8f;
    metal->roughness() = 0.2f;
    auto lightM = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("cornellLight"));
    lightM->emissive() = vec4(1.0f, 0.9f, 0.7f, 15.0f);
    const float h = vm::radians(90.0f);
    addMeshObject(scene, "floor", Transform({0, 0, 0}, {1, 1, 1}), bm.plane, white);
    addMeshObject(scene, "ceiling", Transform({0, 2, 0}, {1, 1, 1}, vec3(2 * h, 0, 0)), bm.plane, white);
    addMeshObject(scene, "back", Transform({0, 1, -1}, {1, 1, 1}, vec3(h, 0, 0)), bm.plane, white);
    addMeshObject(scene, "left", Transform({-1, 1, 0}, {1, 1, 1}, vec3(0, 0, -h)), bm.plane, red);
    addMeshObject(scene, "right", Transform({1, 1, 0}, {1, 1, 1}, vec3(0, 0, h)), bm.plane, green);
    addMeshObject(scene, "tall", Transform({-0.35f, 0.6f, -0.3f}, {0.3f, 0.6f, 0.3f}, vec3(0, vm::radians(17.0f), 0)), bm.cube, metal);
    addMeshObject(scene, "short", Transform({0.35f, 0.3f, 0.35f}, {0.3f, 0.3f, 0.3f}, vec3(0, vm::radians(-17.0f), 0)), bm.cube, white);
    addMeshObject(scene, "light", Transform({0, 1.98f, 0}, {0.25f, 0.25f, 0.25f}, vec3(2 * h, 0, 0)), bm.plane, lightM);
    scene.environmentType() = EnvironmentType::SOLID_COLOR;
    scene.backgroundColor() = vec3(0, 0, 0);
    scene.environmentIntensity() = 1.0f;
    scene.update();
    RI &ri = info(e);
    ri = RI();
    ri.filename = "Cornell";
    ri.width = 512;
    ri.height = 512;
    ri.samples = 64;
    ri.batchSize = 8;
    ri.depth = 8;
}

/* ---------------------------------------------------------------------- C2 Sponza-class atrium / C3 fog */
static void atrium(Engine &e, const Options &opt, bool fog) {
    Scene &scene = e.scene();
    const float s = std::max(0.05f, opt.scale * (fog ? 0.58f : 1.0f));
    auto N = [&](int n) { return std::max(2, (int)std::lround(n * std::sqrt(s))); };
    const int ts = std::max(4, opt.textureSize);

    /* 24 textured PBR materials, seed 2 */
    std::vector<MaterialPBRStandard *> mats;
    std::mt19937 rng(2);
    std::uniform_real_distribution<float> U(0.0f, 1.0f);
    auto rc = [&](float lo, float hi) { return vec3(lo + (hi - lo) * U(rng), lo + (hi - lo) * U(rng), lo + (hi - lo) * U(rng)); };
    mats.push_back(texturedMaterial(e, "atriumFloor", ts, 200, vec3(0.75f, 0.72f, 0.66f), vec3(0.25f, 0.22f, 0.2f), 3, 0.0f, 24.0f));
    for (int i = 0; i < 5; i++) mats.push_back(texturedMaterial(e, "atriumStone" + std::to_string(i), ts, 210 + i, rc(0.55f, 0.8f), rc(0.3f, 0.5f), 0, 0.0f, 2.0f));
    for (int i = 0; i < 5; i++) mats.push_back(texturedMaterial(e, "atriumBrick" + std::to_string(i), ts, 220 + i, rc(0.2f, 0.35f), vec3(0.6f, 0.3f, 0.2f) * (0.7f + 0.6f * U(rng)), 1, 0.0f, 4.0f));
    for (int i = 0; i < 8; i++) {
        vec3 c = rc(0.05f, 0.9f);
        mats.push_back(texturedMaterial(e, "atriumFabric" + std::to_string(i), ts, 230 + i, c, c * 0.5f, 2, 0.0f, 6.0f));
    }
    for (int i = 0; i < 5; i++) mats.push_back(texturedMaterial(e, "atriumMetal" + std::to_string(i), ts, 240 + i, rc(0.6f, 0.95f), rc(0.4f, 0.7f), 0, 0.9f, 3.0f));
    /* = 1 + 5 + 5 + 8 + 5 = 24 */

    const float L = 20.0f, Wd = 8.0f, Hs = 6.0f; /* half length, half width, storey height */
    /* floor: bumpy tile grid */
    {
        MeshBuilder b("atriumFloorMesh");
        b.surface(N(224), N(112), [&](float u, float v) {
            float x = -L + 2 * L * u, z = -Wd + 2 * Wd * v;
            float y = 0.02f * fbm(u * 4, v * 2, 11, 8) + 0.01f * std::sin(x * 6.0f) * std::sin(z * 6.0f);
            return vec3(x, y, z);
        }, true);
        addMeshObject(scene, "floor", Transform(), registerMesh(e, "atrium/floor", b.finish()), mats[0]);
    }
    /* column: fluted cylinder with base and capital, instanced */
    Mesh *column;
    {
        MeshBuilder b("atriumColumnMesh");
        b.surface(N(48), N(24), [&](float u, float v) {
            float a = TWO_PI * u;
            float r = 0.38f - 0.05f * v + 0.02f * std::cos(a * 16.0f);
            return vec3(r * std::cos(a), 0.4f + (Hs - 1.0f) * v, r * std::sin(a));
        }, true);
        b.box(vec3(-0.55f, 0, -0.55f), vec3(0.55f, 0.4f, 0.55f));
        b.box(vec3(-0.5f, Hs - 0.6f, -0.5f), vec3(0.5f, Hs - 0.3f, 0.5f));
        column = registerMesh(e, "atrium/column", b.finish());
    }
    /* arch: swept rectangular section along a semicircle, instanced */
    Mesh *arch;
    {
        MeshBuilder b("atriumArchMesh");
        const float R = 1.6f, t = 0.25f, d = 0.45f;
        for (int side = 0; side < 4; side++)
            b.surface(N(48), 2, [&](float u, float v) {
                float a = 3.14159265f * u;
                float r, z;
                switch (side) {
                    case 0: r = R - t; z = -d + 2 * d * v; break;
                    case 1: r = R + t; z = d - 2 * d * v; break;
                    case 2: r = R - t + 2 * t * v; z = d; break;
                    default: r = R + t - 2 * t * v; z = -d; break;
                }
                return vec3(r * std::cos(a), r * std::sin(a), z);
            }, side == 1 || side == 2 ? false : true);
        arch = registerMesh(e, "atrium/arch", b.finish());
    }
    const int nCols = 10;
    for (int storey = 0; storey < 2; storey++)
        for (int row = 0; row < 2; row++) {
            float z = row == 0 ? -4.0f : 4.0f;
            for (int i = 0; i < nCols; i++) {
                float x = -L + 2.0f + i * (2 * L - 4.0f) / (nCols - 1);
                int mi = 1 + (i + storey * 3 + row) % 5;
                addMeshObject(scene, "column", Transform({x, storey * Hs, z}, {1, 1, 1}, vec3(0, 0.37f * i, 0)), column, mats[mi]);
                if (i + 1 < nCols) {
                    float xn = -L + 2.0f + (i + 1) * (2 * L - 4.0f) / (nCols - 1);
                    float span = (xn - x) * 0.5f;
                    addMeshObject(scene, "arch", Transform({(x + xn) * 0.5f, storey * Hs + Hs - 2.0f, z}, {span / 1.85f, 1.0f, 1.0f}), arch, mats[1 + (i + row) % 5]);
                }
            }
        }
    /* side walls with brick relief; end walls */
    for (int side = 0; side < 2; side++) {
        MeshBuilder b("atriumWallMesh" + std::to_string(side));
        float zs = side == 0 ? -Wd : Wd;
        b.surface(N(160), N(56), [&](float u, float v) {
            float x = -L + 2 * L * u, y = 2 * Hs * v;
            float relief = 0.05f * fbm(u * 8, v * 4, 31 + side, 8);
            float win = (std::fabs(std::fmod(x + L, 4.0f) - 2.0f) < 0.6f && std::fabs(std::fmod(y, Hs) - 3.2f) < 1.2f) ? 0.35f : 0.0f;
            return vec3(x, y, zs + (side == 0 ? 1 : -1) * (relief - win));
        }, side == 0);
        addMeshObject(scene, "wall", Transform(), registerMesh(e, "atrium/wall" + std::to_string(side), b.finish()), mats[6 + side]);
    }
    for (int end = 0; end < 2; end++) {
        MeshBuilder b("atriumEndMesh" + std::to_string(end));
        float xs = end == 0 ? -L : L;
        b.surface(N(64), N(56), [&](float u, float v) {
            float z = -Wd + 2 * Wd * u, y = 2 * Hs * v;
            return vec3(xs + (end == 0 ? 1 : -1) * 0.05f * fbm(u * 4, v * 4, 41 + end, 8), y, z);
        }, end == 1);
        addMeshObject(scene, "end", Transform(), registerMesh(e, "atrium/end" + std::to_string(end), b.finish()), mats[8 + end]);
    }
    /* upper gallery slabs and aisle ceilings (the nave |z| < 4 stays open to the sky) */
    {
        MeshBuilder b("atriumSlabMesh");
        b.box(vec3(-L, -0.2f, 0), vec3(L, 0.0f, 4.0f));
        Mesh *slab = registerMesh(e, "atrium/slab", b.finish());
        addMeshObject(scene, "slab", Transform({0, Hs, 4.0f}), slab, mats[10]);
        addMeshObject(scene, "slab", Transform({0, Hs, -8.0f}), slab, mats[10]);
        addMeshObject(scene, "slab", Transform({0, 2 * Hs, 4.0f}), slab, mats[10]);
        addMeshObject(scene, "slab", Transform({0, 2 * Hs, -8.0f}), slab, mats[10]);
    }
    /* drapes: wavy cloth sheets hanging across the nave */
    {
        const int nDrapes = 5;
        for (int k = 0; k < nDrapes; k++) {
            MeshBuilder b("atriumDrapeMesh" + std::to_string(k));
            float x0 = -L + 6.0f + k * 7.0f;
            b.surface(N(80), N(80), [&](float u, float v) {
                float z = -3.5f + 7.0f * u;
                float sag = 1.5f * (1.0f - std::pow(2 * u - 1, 2.0f));
                float y = 2 * Hs - 0.5f - sag - 3.0f * v;
                float x = x0 + 0.25f * std::sin(z * 3.0f + k) * (0.3f + v) + 0.1f * std::sin(v * 25.0f + u * 7.0f);
                return vec3(x, y, z);
            });
            addMeshObject(scene, "drape", Transform(), registerMesh(e, "atrium/drape" + std::to_string(k), b.finish()), mats[11 + k]);
        }
    }
    /* metal railings on the gallery */
    {
        MeshBuilder b("atriumRailMesh");
        b.surface(N(24), N(64), [&](float u, float v) {
            float a = TWO_PI * u;
            return vec3(-L + 2 * L * v, 0.06f * std::sin(a), 0.06f * std::cos(a));
        }, true);
        Mesh *rail = registerMesh(e, "atrium/rail", b.finish());
        for (int row = 0; row < 2; row++)
            for (int k = 0; k < 3; k++) addMeshObject(scene, "rail", Transform({0, Hs + 0.4f + 0.3f * k, row == 0 ? -3.9f : 3.9f}), rail, mats[19 + (k + row) % 5]);
    }

    auto camera = makeCamera(scene, vec3(-17.0f, 3.0f, 0.5f), vec3(vm::radians(8.0f), vm::radians(-90.0f), 0));
    camera->fov() = 60.0f;
    camera->zfar() = 100.0f;
    scene.environmentType() = EnvironmentType::HDRI;
    scene.environmentIntensity() = 1.0f;
    RI &ri = info(e);
    ri = RI();
    ri.width = 1920;
    ri.height = 1080;
    if (!fog) {
        ri.filename = "Atrium";
        ri.samples = 1024;
        ri.batchSize = 16;
        ri.depth = 9;
    } else {
        /* C3: camera volume, point light + 4 emissive quads, depth 32, depth of field */
        ri.filename = "Fog";
        ri.samples = 256;
        ri.batchSize = 16;
        ri.depth = 32;
        auto vol = e.materials().createMaterial<MaterialVolume>(AssetInfo("fogVolume"));
        vol->sigmaA() = vec4(0.01f, 0.01f, 0.01f, 0);
        vol->sigmaS() = vec4(0.05f, 0.05f, 0.05f, 0);
        vol->g() = 0.3f;
        camera->volume() = vol;
        camera->lensRadius() = 0.05f;
        camera->focalDistance() = 8.0f;
        Light *pl = scene.createLight(AssetInfo("fogPoint"), LightType::POINT_LIGHT);
        pl->color() = vec4(1.0f, 0.85f, 0.7f, 50.0f);
        scene.addSceneObject("fogPoint", Transform({0, 8, 0}))->add<ComponentLight>().setLight(pl);
        BaseMeshes bm = baseMeshes(e);
        for (int k = 0; k < 4; k++) {
            auto em = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("fogEmissive" + std::to_string(k)));
            em->emissive() = vec4(0.4f + 0.2f * k, 0.7f, 1.0f - 0.2f * k, 12.0f);
            addMeshObject(scene, "fogQuad", Transform({-12.0f + 8.0f * k, 5.5f, (k & 1) ? 6.0f : -6.0f}, {0.8f, 0.8f, 0.8f}, vec3(vm::radians(180.0f), 0, 0)), bm.plane, em);
        }
        scene.environmentIntensity() = 0.3f;
    }
    scene.update();
}

/* ---------------------------------------------------------------------- C5 progressive emissive spheres */
static void progressive(Engine &e, const Options &opt) {
    Scene &scene = e.scene();
    BaseMeshes bm = baseMeshes(e);
    Mesh *sphere;
    {
        MeshBuilder b("lowSphereMesh");
        b.surface(16, 10, [&](float u, float v) {
            float th = 3.14159265f * v, ph = TWO_PI * u;
            return vec3(std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph));
        }, true, 1, 1, true);
        sphere = registerMesh(e, "progressive/sphere", b.finish());
    }
    auto floorM = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("progressiveFloor"));
    floorM->albedo() = vec4(0.6f, 0.6f, 0.6f, 1);
    floorM->metallic() = 0.2f;
    floorM->roughness() = 0.4f;
    addMeshObject(scene, "floor", Transform({0, 0, 0}, {30, 30, 30}), bm.plane, floorM);
    std::mt19937 rng(4);
    std::uniform_real_distribution<float> U(0.0f, 1.0f);
    int count = std::max(4, (int)std::lround(400 * opt.scale));
    for (int i = 0; i < count; i++) {
        float hue = U(rng) * 6.0f, sat = 0.6f + 0.4f * U(rng);
        float c = sat, x = c * (1 - std::fabs(std::fmod(hue, 2.0f) - 1));
        vec3 rgb = hue < 1 ? vec3(c, x, 0) : hue < 2 ? vec3(x, c, 0) : hue < 3 ? vec3(0, c, x) : hue < 4 ? vec3(0, x, c) : hue < 5 ? vec3(x, 0, c) : vec3(c, 0, x);
        rgb = rgb + vec3(1 - sat);
        auto em = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("progressiveEmissive" + std::to_string(i % 64)));
        if (i < 64) em->emissive() = vec4(rgb.x, rgb.y, rgb.z, 4.0f);
        float r = 0.25f + 0.35f * U(rng);
        addMeshObject(scene, "emitter", Transform({-18.0f + 36.0f * U(rng), r + 3.0f * U(rng), -18.0f + 36.0f * U(rng)}, {r, r, r}), sphere, em);
    }
    if (opt.camera == 1) {
        auto cam = std::make_shared<OrthographicCamera>();
        cam->transform().position() = vec3(0, 25, 30);
        cam->transform().setRotation(quat(vec3(vm::radians(-40.0f), 0, 0)));
        cam->setOrthoWidth(40.0f);
        cam->zfar() = 100.0f;
        scene.camera() = cam;
    } else {
        auto cam = makeCamera(scene, vec3(0, 8, 30), vec3(vm::radians(-15.0f), 0, 0));
        cam->fov() = 50.0f;
        cam->zfar() = 100.0f;
    }
    scene.environmentType() = EnvironmentType::SOLID_COLOR;
    scene.backgroundColor() = vec3(0, 0, 0);
    scene.update();
    RI &ri = info(e);
    ri = RI();
    ri.filename = "Progressive";
    ri.width = 1920;
    ri.height = 1080;
    ri.samples = 4096;
    ri.batchSize = 16;
    ri.depth = 9;
}

/* ---------------------------------------------------------------------- C4 instanced forest */
static void instanced(Engine &e, const Options &opt) {
    Scene &scene = e.scene();
    const float s = std::max(0.001f, opt.scale);
    auto N = [&](int n) { return std::max(2, (int)std::lround(n * std::sqrt(s))); };
    std::mt19937 rng(3);
    std::uniform_real_distribution<float> U(0.0f, 1.0f);
    /* alpha-tested leaf texture: 512^2 noise alpha */
    const int ts = std::max(8, std::min(512, opt.textureSize));
    ImageU8 alpha;
    alpha.width = alpha.height = ts;
    alpha.channels = 1;
    alpha.data.resize((size_t)ts * ts);
    for (int y = 0; y < ts; y++)
        for (int x = 0; x < ts; x++) {
            float u = (x + 0.5f) / ts, v = (y + 0.5f) / ts;
            float r = std::hypot(u - 0.5f, v - 0.5f);
            float n = fbm(u, v, 3, 8);
            alpha.data[(size_t)y * ts + x] = (r + 0.25f * (n - 0.5f) < 0.42f) ? 255 : 0;
        }
    Texture *alphaTex = e.textures().createTexture("leafAlpha", alpha, ColorSpace::LINEAR);
    auto leafM = e.materials().createMaterial<MaterialLambert>(AssetInfo("leaf"));
    leafM->albedo() = vec4(0.15f, 0.45f, 0.1f, 1);
    leafM->setAlphaTexture(alphaTex);
    leafM->setTransparent(true);
    auto barkM = e.materials().createMaterial<MaterialLambert>(AssetInfo("bark"));
    barkM->albedo() = vec4(0.35f, 0.25f, 0.15f, 1);
    auto rockM = e.materials().createMaterial<MaterialPBRStandard>(AssetInfo("rock"));
    rockM->albedo() = vec4(0.5f, 0.5f, 0.48f, 1);
    rockM->roughness() = 0.8f;
    auto groundM = e.materials().createMaterial<MaterialLambert>(AssetInfo("ground"));
    groundM->albedo() = vec4(0.3f, 0.35f, 0.2f, 1);

    const int nTreeKinds = 16, nRockKinds = 8;
    std::vector<Mesh *> trunks, crowns, rocks;
    for (int k = 0; k < nTreeKinds; k++) {
        MeshBuilder tb("trunk" + std::to_string(k));
        float hgt = 6.0f + 4.0f * U(rng);
        tb.surface(N(32), N(60), [&](float u, float v) {
            float a = TWO_PI * u, r = 0.35f * (1.0f - 0.8f * v) + 0.03f * std::sin(a * 5 + v * 9);
            return vec3(r * std::cos(a) + 0.3f * std::sin(v * 3 + k), hgt * v, r * std::sin(a));
        }, true);
        trunks.push_back(registerMesh(e, "forest/trunk" + std::to_string(k), tb.finish()));
        MeshBuilder cb("crown" + std::to_string(k));
        int cards = std::max(8, (int)std::lround(18000 * s)); /* 2 triangles per leaf card */
        for (int c = 0; c < cards; c++) {
            float th = std::acos(1 - 2 * U(rng)), ph = TWO_PI * U(rng), rr = 2.5f * std::cbrt(U(rng));
            vec3 ctr(rr * std::sin(th) * std::cos(ph), hgt * 0.8f + rr * std::cos(th) * 0.8f, rr * std::sin(th) * std::sin(ph));
            vec3 n = vm::normalize(vec3(U(rng) - 0.5f, U(rng) - 0.2f, U(rng) - 0.5f));
            vec3 t = vm::normalize(vm::cross(n, vec3(0.3f, 1, 0.2f)));
            vec3 bt = vm::cross(n, t);
            float sz = 0.15f + 0.1f * U(rng);
            uint32_t a0 = cb.vertex(ctr - t * sz - bt * sz, n, vec2(0, 0)), a1 = cb.vertex(ctr + t * sz - bt * sz, n, vec2(1, 0));
            uint32_t a2 = cb.vertex(ctr + t * sz + bt * sz, n, vec2(1, 1)), a3 = cb.vertex(ctr - t * sz + bt * sz, n, vec2(0, 1));
            cb.tri(a0, a1, a2);
            cb.tri(a0, a2, a3);
        }
        crowns.push_back(registerMesh(e, "forest/crown" + std::to_string(k), cb.finish()));
    }
    for (int k = 0; k < nRockKinds; k++) {
        MeshBuilder rb("rock" + std::to_string(k));
        rb.surface(N(100), N(50), [&](float u, float v) {
            float th = 3.14159265f * v, ph = TWO_PI * u;
            float r = 1.0f + 0.35f * (fbm(u * 2, v, 50 + k, 4) - 0.5f);
            return vec3(r * std::sin(th) * std::cos(ph), 0.6f * r * std::cos(th), r * std::sin(th) * std::sin(ph));
        }, true);
        rocks.push_back(registerMesh(e, "forest/rock" + std::to_string(k), rb.finish()));
    }
    const float half = 1000.0f * std::sqrt(std::min(1.0f, s * 4.0f));
    auto terrainH = [&](float x, float z) { return 12.0f * (fbm(x / (2 * half) + 0.5f, z / (2 * half) + 0.5f, 77, 4) - 0.5f); };
    {
        MeshBuilder gb("terrain");
        gb.surface(N(316), N(316), [&](float u, float v) {
            float x = -half + 2 * half * u, z = -half + 2 * half * v;
            return vec3(x, terrainH(x, z), z);
        }, true);
        addMeshObject(scene, "terrain", Transform(), registerMesh(e, "forest/terrain", gb.finish()), groundM);
    }
    int nInst = std::max(8, (int)std::lround(1250 * std::min(1.0f, s * 4.0f)));
    for (int i = 0; i < nInst; i++) {
        float x = (U(rng) * 2 - 1) * half * 0.3f, z = (U(rng) * 2 - 1) * half * 0.3f;
        float y = terrainH(x, z);
        float sc = 0.8f + 0.6f * U(rng);
        Transform t({x, y, z}, {sc, sc, sc}, vec3(0, TWO_PI * U(rng), 0));
        if (i % 5 == 4) {
            addMeshObject(scene, "rock", t, rocks[i % nRockKinds], rockM);
        } else {
            int k = i % nTreeKinds;
            addMeshObject(scene, "trunk", t, trunks[k], barkM);
            addMeshObject(scene, "crown", t, crowns[k], leafM);
        }
    }
    Light *sun = e.lightsMap().get("defaultDirectionalLightSun");
    scene.addSceneObject("sun", Transform({0, 0, 0}, {1, 1, 1}, vec3(vm::radians(50.0f), vm::radians(30.0f), 0)))->add<ComponentLight>().setLight(sun);
    auto cam = makeCamera(scene, vec3(0, terrainH(0, half * 0.32f) + 12.0f, half * 0.32f), vec3(vm::radians(-8.0f), 0, 0));
    cam->fov() = 55.0f;
    cam->zfar() = 3000.0f;
    scene.environmentType() = EnvironmentType::SOLID_COLOR;
    scene.backgroundColor() = vec3(0.55f, 0.7f, 0.95f);
    scene.update();
    RI &ri = info(e);
    ri = RI();
    ri.filename = "Instanced";
    ri.width = 3840;
    ri.height = 2160;
    ri.samples = 256;
    ri.batchSize = 8;
    ri.depth = 6;
}

/* ====================================================================== registry */
std::vector<std::string> list() {
    return {"FurnacePBR", "FurnaceLambert", "EnvironmentMap", "EnvironmentMapPBR00", "EnvironmentMapPBR01", "EnvironmentMapPBR10", "EnvironmentMapPBR11",
            "EnvironmentMapLambert", "Volume0", "Volume1", "Volume2", "Volume3", "Volume4", "Volume5", "Volume6", "Volume7", "Volume8", "Volume9",
            "PointLight", "DirectionalLight", "MeshLight", "Transparency", "NormalMap", "GLTF", "Hierarchy", "DepthOfField", "SharedComponents", "Denoise",
            "Cornell", "Atrium", "Fog", "Instanced", "Progressive", "BallOnPlane"};
}

bool build(Engine &e, const std::string &name, const Options &opt) {
    e.scene().clear();
    if (name == "FurnacePBR") furnace(e, false);
    else if (name == "FurnaceLambert") furnace(e, true);
    else if (name == "EnvironmentMap") environmentMap(e);
    else if (name.rfind("EnvironmentMapPBR", 0) == 0 && name.size() == 19) environmentMapPBR(e, name[17] - '0', name[18] - '0');
    else if (name == "EnvironmentMapLambert") environmentMapLambert(e);
    else if (name.rfind("Volume", 0) == 0 && name.size() == 7) volume(e, name[6] - '0');
    else if (name == "PointLight") lights(e, 0);
    else if (name == "DirectionalLight") lights(e, 1);
    else if (name == "MeshLight") lights(e, 2);
    else if (name == "Transparency") transparency(e);
    else if (name == "NormalMap") normalMap(e);
    else if (name == "GLTF") gltf(e);
    else if (name == "Hierarchy") hierarchy(e);
    else if (name == "DepthOfField") depthOfField(e);
    else if (name == "SharedComponents") sharedComponents(e);
    else if (name == "Denoise") denoise(e);
    else if (name == "Cornell") cornell(e);
    else if (name == "Atrium") atrium(e, opt, false);
    else if (name == "Fog") atrium(e, opt, true);
    else if (name == "Progressive") progressive(e, opt);
    else if (name == "Instanced") instanced(e, opt);
    else if (name == "BallOnPlane") ballOnPlane(e);
    else return false;
    return true;
}

}  // namespace scenes
}  // namespace vengine
