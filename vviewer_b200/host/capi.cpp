/*
 * Small C API over the C++ host library so that Python (tests, bench.py) can build the recipe scenes,
 * obtain the flattened POD scene that crosses include/ptc.h, and drive RendererPathTracing::render().
 * Declared in include/vengine_host.h.
 */
#include "../../include/vengine_host.h"
#include "scenes.hpp"

#include <cstring>
#include <string>

using namespace vengine;

struct vh_engine {
    std::unique_ptr<Engine> engine;
    FlatScene flat;
    bool flatValid = false;
    std::string lastError;
};

extern "C" {

PTC_API vh_engine *vh_engine_create(const char *backend_lib, const char *asset_root) {
    auto *h = new vh_engine();
    h->engine = std::make_unique<Engine>("vh", backend_lib ? backend_lib : "", asset_root ? asset_root : "");
    h->engine->initResources();
    return h;
}

PTC_API void vh_engine_destroy(vh_engine *h) { delete h; }

PTC_API int vh_backend_ok(vh_engine *h) { return h && h->engine->renderer().rendererPathTracing().isRayTracingEnabled() ? 1 : 0; }

PTC_API const char *vh_last_error(vh_engine *h) {
    if (!h) return "null engine";
    h->lastError = h->engine->renderer().rendererPathTracing().lastError();
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
    return scenes::build(*h->engine, name, opt) ? 0 : 2;
}

PTC_API void vh_set_render_info(vh_engine *h, int width, int height, int samples, int batch_size, int depth) {
    auto &ri = h->engine->renderer().rendererPathTracing().renderInfo();
    if (width > 0) ri.width = (uint32_t)width;
    if (height > 0) ri.height = (uint32_t)height;
    if (samples > 0) ri.samples = (uint32_t)samples;
    if (batch_size > 0) ri.batchSize = (uint32_t)batch_size;
    if (depth > 0) ri.depth = (uint32_t)depth;
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
    std::vector<float> r, a, n;
    if (!h->engine->renderer().rendererPathTracing().renderToMemory(r, a, n)) return 2;
    if (radiance) std::memcpy(radiance, r.data(), r.size() * sizeof(float));
    if (albedo) std::memcpy(albedo, a.data(), a.size() * sizeof(float));
    if (normal) std::memcpy(normal, n.data(), n.size() * sizeof(float));
    return 0;
}

PTC_API int vh_render(vh_engine *h, const char *filename) {
    if (!h) return 1;
    auto &pt = h->engine->renderer().rendererPathTracing();
    if (filename) pt.renderInfo().filename = filename;
    if (!pt.isRayTracingEnabled()) return 2;
    pt.render();
    return 0;
}

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

} /* extern "C" */
