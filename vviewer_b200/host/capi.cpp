/*
 * Small C API over the C++ host library so that Python (tests, bench.py) can build the recipe scenes,
 * obtain the flattened POD scene that crosses include/ptc.h, and drive RendererPathTracing::render().
 * Declared in include/vengine_host.h.
 */
#include "../../include/vengine_host.h"
#include "scenes.hpp"

#include <cstring>
#include <sstream>
#include <string>

#include "json.hpp"

using namespace vengine;

struct vh_engine {
    std::unique_ptr<Engine> engine;
    FlatScene flat;
    bool flatValid = false;
    std::string lastError, apiError, description;
};

static uint64_t fnv1a(const void *data, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t *p = static_cast<const uint8_t *>(data);
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}
static std::string hex64(uint64_t v) {
    char buf[20];
    std::snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)v);
    return buf;
}

extern "C" {

PTC_API vh_engine *vh_engine_create(const char *asset_root) {
    auto *h = new vh_engine();
    h->engine = std::make_unique<Engine>("vh", asset_root ? asset_root : "");
    h->engine->initResources();
    return h;
}

PTC_API void vh_engine_destroy(vh_engine *h) { delete h; }

PTC_API int vh_backend_ok(vh_engine *h) { return h && h->engine->renderer().rendererPathTracing().isRayTracingEnabled() ? 1 : 0; }

PTC_API const char *vh_last_error(vh_engine *h) {
    if (!h) return "null engine";
    h->lastError = !h->apiError.empty() ? h->apiError : h->engine->renderer().rendererPathTracing().lastError();
    return h->lastError.c_str();
}

PTC_API const char *vh_scene_list(void) {
    static std::string s;
    s.clear();
    for (auto &n : scenes::list()) {
        if (!s.empty()) s += ",";
        s += n;
    }
    return s.c_str();
}

PTC_API int vh_build_scene(vh_engine *h, const char *name, int texture_size, float scale, int camera) {
    if (!h || !name) return 1;
    scenes::Options opt;
    if (texture_size > 0) opt.textureSize = texture_size;
    if (scale > 0) opt.scale = scale;
    opt.camera = camera;
    h->flatValid = false;
    h->apiError.clear();
    try {
        return scenes::build(*h->engine, name, opt) ? 0 : 2;
    } catch (std::exception &e) {
        h->apiError = e.what();
        return 3;
    }
}

PTC_API int vh_import_model(vh_engine *h, const char *path, int import_materials) {
    if (!h || !path) return 1;
    h->apiError.clear();
    if (!h->engine->importModel(AssetInfo(path), import_materials != 0)) {
        h->apiError = std::string("failed to import model ") + path;
        return 2;
    }
    return 0;
}

PTC_API int vh_add_model(vh_engine *h, const char *model_name) {
    if (!h || !model_name) return 1;
    h->apiError.clear();
    h->flatValid = false;
    try {
        addModel3D(h->engine->scene(), nullptr, model_name, std::nullopt, std::nullopt);
        h->engine->scene().update();
    } catch (std::exception &e) {
        h->apiError = e.what();
        return 2;
    }
    return 0;
}

PTC_API int vh_import_scene(vh_engine *h, const char *scene_json) {
    if (!h || !scene_json) return 1;
    h->apiError.clear();
    h->flatValid = false;
    std::string err;
    if (!h->engine->importScene(scene_json, &err)) {
        h->apiError = err;
        return 2;
    }
    return 0;
}

PTC_API int vh_export_scene(vh_engine *h, const char *name) {
    if (!h || !name) return 1;
    h->apiError.clear();
    std::string err;
    if (!h->engine->exportScene(name, &err)) {
        h->apiError = err;
        return 2;
    }
    return 0;
}

/* JSON description of the engine state for tests: models (node tree, meshes with counts / bounds / content hashes), materials
 * (the 128-byte block + texture names), textures (size, channels, colour space, content hash), scene objects */
PTC_API const char *vh_describe(vh_engine *h) {
    if (!h) return "{}";
    using json::Value;
    Engine &e = *h->engine;
    Value d = Value::object();
    Value models = Value::array();
    for (auto &it : e.modelsMap().all()) {
        Model3D *m = it.second;
        Value mo = Value::object();
        mo.set("name", m->name);
        mo.set("internal", m->internal);
        std::function<Value(const Model3D::Model3DNode &)> nodeValue = [&](const Model3D::Model3DNode &n) {
            Value v = Value::object();
            v.set("name", n.name);
            const Transform &t = n.transform;
            Value tr = Value::array();
            for (float f : {t.position().x, t.position().y, t.position().z, t.scale().x, t.scale().y, t.scale().z, t.rotation().w, t.rotation().x,
                            t.rotation().y, t.rotation().z})
                tr.push(Value(f));
            v.set("transform", tr);
            Value meshes = Value::array();
            for (size_t i = 0; i < n.meshes.size(); i++) {
                Value me = Value::object();
                me.set("name", n.meshes[i]->name);
                me.set("material", n.materials[i] ? n.materials[i]->name() : std::string(""));
                meshes.push(me);
            }
            v.set("meshes", meshes);
            Value children = Value::array();
            for (auto &c : n.children) children.push(nodeValue(c));
            v.set("children", children);
            return v;
        };
        mo.set("nodeTree", nodeValue(m->nodeTree));
        Value meshes = Value::array();
        for (auto &mesh : m->meshes) {
            Value me = Value::object();
            me.set("name", mesh->name);
            me.set("vertices", (double)mesh->vertices.size());
            me.set("triangles", (double)mesh->nTriangles());
            float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, uvlo[2] = {1e30f, 1e30f}, uvhi[2] = {-1e30f, -1e30f};
            double tdotn = 0, tlen = 0, nlen = 0;
            for (const Vertex &v : mesh->vertices) {
                for (int c = 0; c < 3; c++) {
                    lo[c] = std::min(lo[c], v.position[c]);
                    hi[c] = std::max(hi[c], v.position[c]);
                }
                for (int c = 0; c < 2; c++) {
                    uvlo[c] = std::min(uvlo[c], v.uv[c]);
                    uvhi[c] = std::max(uvhi[c], v.uv[c]);
                }
                tdotn += std::fabs(v.tangent[0] * v.normal[0] + v.tangent[1] * v.normal[1] + v.tangent[2] * v.normal[2]);
                tlen += std::sqrt(v.tangent[0] * v.tangent[0] + v.tangent[1] * v.tangent[1] + v.tangent[2] * v.tangent[2]);
                nlen += std::sqrt(v.normal[0] * v.normal[0] + v.normal[1] * v.normal[1] + v.normal[2] * v.normal[2]);
            }
            Value b = Value::array();
            for (int c = 0; c < 3; c++) b.push(Value(lo[c]));
            for (int c = 0; c < 3; c++) b.push(Value(hi[c]));
            me.set("bounds", b);
            Value ub = Value::array();
            for (float f : {uvlo[0], uvlo[1], uvhi[0], uvhi[1]}) ub.push(Value(f));
            me.set("uvBounds", ub);
            double nv = (double)std::max<size_t>(mesh->vertices.size(), 1);
            me.set("meanAbsTangentDotNormal", tdotn / nv);
            me.set("meanTangentLength", tlen / nv);
            me.set("meanNormalLength", nlen / nv);
            me.set("indexHash", hex64(fnv1a(mesh->indices.data(), mesh->indices.size() * sizeof(uint32_t))));
            /* positions + uv + normals only: tangents are derived data */
            uint64_t hv = 1469598103934665603ull;
            for (const Vertex &v : mesh->vertices) {
                hv = fnv1a(v.position, sizeof(v.position), hv);
                hv = fnv1a(v.uv, sizeof(v.uv), hv);
                hv = fnv1a(v.normal, sizeof(v.normal), hv);
            }
            me.set("vertexHash", hex64(hv));
            meshes.push(me);
        }
        mo.set("meshes", meshes);
        models.push(mo);
    }
    d.set("models", models);
    Value materials = Value::array();
    for (auto &it : e.materials().all()) {
        Material *m = it.second;
        const ptc_material &b = m->block();
        Value mo = Value::object();
        mo.set("name", m->name());
        mo.set("type", (int)m->type());
        mo.set("embedded", m->isEmbedded());
        auto vec = [](const float *f, int n) {
            Value a = Value::array();
            for (int i = 0; i < n; i++) a.push(Value(f[i]));
            return a;
        };
        mo.set("albedo", vec(b.albedo, 4));
        mo.set("metallicRoughnessAO", vec(b.metallic_roughness_ao, 4));
        mo.set("emissive", vec(b.emissive, 4));
        mo.set("uvTiling", vec(b.uv_tiling, 4));
        static const char *slots[8] = {"albedo", "metallic", "roughness", "ao", "emissive", "normal", "brdfLUT", "alpha"};
        Value tex = Value::object();
        for (int i = 0; i < 8; i++) {
            uint32_t slot = i < 4 ? b.tex1[i] : b.tex2[i - 4];
            Texture *t = e.textures().bySlot(slot);
            tex.set(slots[i], t ? t->name : std::string("?"));
        }
        mo.set("textures", tex);
        mo.set("emissiveFlag", m->isEmissive());
        mo.set("transparentFlag", m->isTransparent());
        materials.push(mo);
    }
    d.set("materials", materials);
    Value textures = Value::array();
    for (auto &t : e.textures().all()) {
        Value to = Value::object();
        to.set("name", t->name);
        to.set("width", t->image.width);
        to.set("height", t->image.height);
        to.set("channels", t->image.channels);
        to.set("srgb", t->colorSpace == ColorSpace::sRGB);
        to.set("embedded", t->embedded);
        to.set("hash", hex64(fnv1a(t->image.data.data(), t->image.data.size())));
        textures.push(to);
    }
    d.set("textures", textures);
    Value objects = Value::array();
    for (SceneObject *so : e.scene().getSceneObjectsFlat()) {
        Value o = Value::object();
        o.set("name", so->name());
        o.set("parent", so->parent() ? so->parent()->name() : std::string(""));
        o.set("active", so->isActive());
        o.set("mesh", so->has<ComponentMesh>() && so->get<ComponentMesh>().mesh() ? so->get<ComponentMesh>().mesh()->name : std::string(""));
        o.set("material", so->has<ComponentMaterial>() && so->get<ComponentMaterial>().material() ? so->get<ComponentMaterial>().material()->name() : std::string(""));
        o.set("light", so->has<ComponentLight>() && so->get<ComponentLight>().light() ? so->get<ComponentLight>().light()->name() : std::string(""));
        Value mm = Value::array();
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) mm.push(Value(so->modelMatrix()[c][r]));
        o.set("modelMatrix", mm);
        objects.push(o);
    }
    d.set("objects", objects);
    Scene &sc = e.scene();
    Value env = Value::object();
    env.set("environmentType", (int)sc.environmentType());
    env.set("map", sc.skyboxMaterial() ? sc.skyboxMaterial()->name : std::string(""));
    d.set("environment", env);
    if (sc.camera()) {
        Value cam = Value::object();
        Camera &c = *sc.camera();
        Value p = Value::array();
        for (float f : {c.transform().position().x, c.transform().position().y, c.transform().position().z}) p.push(Value(f));
        cam.set("position", p);
        vec3 fw = c.transform().forward();
        Value fv = Value::array();
        for (float f : {fw.x, fw.y, fw.z}) fv.push(Value(f));
        cam.set("forward", fv);
        cam.set("znear", Value(c.znear()));
        cam.set("zfar", Value(c.zfar()));
        cam.set("lensRadius", Value(c.lensRadius()));
        cam.set("focalDistance", Value(c.focalDistance()));
        if (c.type() == CameraType::PERSPECTIVE) cam.set("fov", Value(static_cast<PerspectiveCamera &>(c).fov()));
        cam.set("volume", c.volume() ? c.volume()->name() : std::string(""));
        d.set("camera", cam);
    }
    h->description = json::dump(d, 0);
    return h->description.c_str();
}

PTC_API void vh_set_render_info(vh_engine *h, int width, int height, int samples, int batch_size, int depth) {
    auto &ri = h->engine->renderer().rendererPathTracing().renderInfo();
    if (width > 0) ri.width = (uint32_t)width;
    if (height > 0) ri.height = (uint32_t)height;
    if (samples > 0) ri.samples = (uint32_t)samples;
    if (batch_size > 0) ri.batchSize = (uint32_t)batch_size;
    if (depth > 0) ri.depth = (uint32_t)depth;
}

PTC_API int vh_set_devices(vh_engine *h, const int *device_ids, int n_devices) {
    if (!h) return 1;
    std::vector<int> ids;
    if (device_ids && n_devices > 0) ids.assign(device_ids, device_ids + n_devices);
    return h->engine->renderer().rendererPathTracing().setDevices(ids) ? 0 : 2;
}
PTC_API int vh_device_count(vh_engine *h) { return h ? h->engine->renderer().rendererPathTracing().deviceCount() : 0; }
PTC_API int vh_comm_unique_id(vh_engine *h, uint8_t *out128) {
    if (!h || !out128) return 1;
    return h->engine->renderer().rendererPathTracing().commUniqueId(out128) ? 0 : 2;
}
PTC_API int vh_comm_init_rank(vh_engine *h, const uint8_t *id128, int rank, int world) {
    if (!h || !id128) return 1;
    return h->engine->renderer().rendererPathTracing().commInitRank(id128, rank, world) ? 0 : 2;
}
PTC_API void vh_set_render_options(vh_engine *h, int multi_gpu_split, int sampler, int env_importance) {
    auto &ri = h->engine->renderer().rendererPathTracing().renderInfo();
    if (multi_gpu_split >= 0) ri.multiGpuSplit = (uint32_t)multi_gpu_split;
    if (sampler >= 0) {
        ri.lowDiscrepancySampler = sampler == 1;
        ri.pmjSampler = sampler == 2;
    }
    if (env_importance >= 0) ri.environmentImportanceSampling = env_importance != 0;
}

PTC_API void vh_get_render_info(vh_engine *h, int *width, int *height, int *samples, int *batch_size, int *depth) {
    auto &ri = h->engine->renderer().rendererPathTracing().renderInfo();
    if (width) *width = (int)ri.width;
    if (height) *height = (int)ri.height;
    if (samples) *samples = (int)ri.samples;
    if (batch_size) *batch_size = (int)ri.batchSize;
    if (depth) *depth = (int)ri.depth;
}

PTC_API const ptc_scene_desc *vh_scene_desc(vh_engine *h) {
    if (!h) return nullptr;
    if (!h->flatValid) {
        h->engine->flatten(h->flat);
        h->flatValid = true;
    }
    return &h->flat.desc;
}

PTC_API int vh_render_params(vh_engine *h, ptc_render_params *out) {
    if (!h || !out) return 1;
    *out = h->engine->renderer().rendererPathTracing().makeRenderParams();
    return 0;
}

PTC_API int vh_render_to_memory(vh_engine *h, float *radiance, float *albedo, float *normal) {
    if (!h) return 1;
    return h->engine->renderer().rendererPathTracing().renderToBuffers(radiance, albedo, normal) ? 0 : 2;
}

PTC_API int vh_render(vh_engine *h, const char *filename) {
    if (!h) return 1;
    auto &pt = h->engine->renderer().rendererPathTracing();
    if (filename) pt.renderInfo().filename = filename;
    if (!pt.isRayTracingEnabled()) return 2;
    pt.render();
    return 0;
}

/* frame `frame` of the BallOnPlane render sequence (PtSceneBallOnPlane.cpp:44-55): only the camera moves */
PTC_API int vh_set_sequence_frame(vh_engine *h, int frame) {
    if (!h || !h->engine->scene().camera()) return 1;
    scenes::ballOnPlaneFrame(*h->engine, frame);
    h->flatValid = false;
    return 0;
}

/* the output half of RenderInfo (core/Renderer.hpp:14-28): file type, PNG exposure, AOV files, denoise flag; negative keeps */
PTC_API void vh_set_output(vh_engine *h, int file_type, float exposure, int write_all_files, int denoise) {
    auto &ri = h->engine->renderer().rendererPathTracing().renderInfo();
    if (file_type >= 0) ri.fileType = file_type == 1 ? FileType::PNG : FileType::HDR;
    ri.exposure = exposure;
    if (write_all_files >= 0) ri.writeAllFiles = write_all_files != 0;
    if (denoise >= 0) ri.denoise = denoise != 0;
}

/* storeToDisk's last step on a caller's image (VulkanRendererPathTracing.cpp:958-975 + core/ImageUtils.cpp:34-76): PNG gets
 * v * 2^exposure when exposure != 0, clamp, linear -> sRGB, uchar(255 x) truncation; HDR is written as RGBE */
PTC_API int vh_write_image(const char *filename_no_ext, int w, int h, int channels, const float *data, int file_type, float exposure) {
    if (!filename_no_ext || !data || w <= 0 || h <= 0 || channels <= 0) return 1;
    std::vector<float> img(data, data + (size_t)w * h * channels);
    const FileType ft = file_type == 1 ? FileType::PNG : FileType::HDR;
    if (ft == FileType::PNG && exposure != 0.0f) applyExposure(img, exposure, (uint32_t)channels);
    writeToDisk(img, filename_no_ext, ft, (uint32_t)w, (uint32_t)h, (uint32_t)channels);
    return 0;
}

PTC_API float vh_render_progress(vh_engine *h) { return h ? h->engine->renderer().rendererPathTracing().renderProgress() : 0.0f; }

PTC_API int vh_get_stats(vh_engine *h, ptc_stats *out) {
    if (!h || !out) return 1;
    *out = h->engine->renderer().rendererPathTracing().lastStats();
    return 0;
}

PTC_API int vh_read_hdr(const char *path, int *w, int *h, float *rgba_out) {
    ImageF32 im;
    if (!loadImageHDR(path, im, false)) return 1;
    if (w) *w = im.width;
    if (h) *h = im.height;
    if (rgba_out) std::memcpy(rgba_out, im.data.data(), im.data.size() * sizeof(float));
    return 0;
}

PTC_API int vh_write_hdr(const char *path, int w, int h, int channels, const float *data) { return writeImageHDR(path, w, h, channels, data) ? 0 : 1; }

/* decode a PNG / JPEG held in memory to RGBA8 (or 1 channel for grey PNGs); query with out == NULL first */
PTC_API int vh_decode_image(const uint8_t *bytes, uint64_t n_bytes, int flip, int *w, int *h, int *channels, int *src_channels, uint8_t *out,
                            uint64_t out_capacity) {
    static thread_local ImageU8 cached;
    static thread_local const uint8_t *cachedKey = nullptr;
    static thread_local int cachedSrc = 0, cachedFlip = -1;
    if (!bytes) return 1;
    if (cachedKey != bytes || cachedFlip != flip || out == nullptr) {
        cached = ImageU8();
        if (!decodeImageU8(bytes, (size_t)n_bytes, cached, &cachedSrc, flip != 0)) {
            cachedKey = nullptr;
            return 2;
        }
        cachedKey = bytes;
        cachedFlip = flip;
    }
    if (w) *w = cached.width;
    if (h) *h = cached.height;
    if (channels) *channels = cached.channels;
    if (src_channels) *src_channels = cachedSrc;
    if (out) {
        if (out_capacity < cached.data.size()) return 3;
        std::memcpy(out, cached.data.data(), cached.data.size());
        cachedKey = nullptr;
        cached = ImageU8();
    }
    return 0;
}

} /* extern "C" */
