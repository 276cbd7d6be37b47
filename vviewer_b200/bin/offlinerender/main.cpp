/*
 * offlinerender — headless driver of the path tracer, same shape as the reference's
 * src/bin/offlinerender/main.cpp:13-25 (create engine, initResources, create a scene, render) with the
 * command line the reference lacks (SURVEY.md §5 "Config / flags").
 *
 *   offlinerender --scene Atrium | --scene path/to/scene.json   (a recipe name, or a scene file of core/io/Import.cpp's format)
 *                 [--export-scene DIR]  (writes DIR/scene.json + DIR/assets like Scene::exportScene, then renders)
 *                 [--width W --height H --spp N --batch B --depth D] [--out name]
 *                 [--png] [--exposure E] [--all-files] [--texsize T] [--scale S] [--camera 0|1] [--sampler default|sobol|pmj] [--env-importance] [--list]
 *                 [--frames N]   (scene BallOnPlane: the reference demo's render sequence, PtSceneBallOnPlane.cpp:44-55; files 0, 1, ...)
 *                 [--gpus N | --devices 0,2,5] [--split tile|sample]   (several GPUs of the box: scene replicated, image partitioned,
 *                                                                        buffers reduced over NCCL onto the first device)
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../host/scenes.hpp"

using namespace vengine;

int main(int argc, char **argv) {
    std::string scene = "Cornell", out, exportDir, sampler = "default", split;
    std::vector<int> devices;
    scenes::Options opt;
    int width = 0, height = 0, spp = 0, batch = 0, depth = 0, frames = 0;
    bool png = false, allFiles = false, envImportance = false;
    float exposure = 0.0f;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--scene") scene = next();
        else if (a == "--width") width = std::atoi(next());
        else if (a == "--height") height = std::atoi(next());
        else if (a == "--spp") spp = std::atoi(next());
        else if (a == "--batch") batch = std::atoi(next());
        else if (a == "--depth") depth = std::atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--frames") frames = std::atoi(next());
        else if (a == "--png") png = true;
        else if (a == "--exposure") exposure = (float)std::atof(next());
        else if (a == "--all-files") allFiles = true;
        else if (a == "--texsize") opt.textureSize = std::atoi(next());
        else if (a == "--scale") opt.scale = (float)std::atof(next());
        else if (a == "--camera") opt.camera = std::atoi(next());
        else if (a == "--sampler") sampler = next();
        else if (a == "--gpus") {
            const int n = std::atoi(next());
            devices.clear();
            for (int k = 0; k < n; k++) devices.push_back(k);
        } else if (a == "--devices") {
            devices.clear();
            std::string list = next();
            for (size_t p = 0; p < list.size();) {
                size_t q = list.find(',', p);
                if (q == std::string::npos) q = list.size();
                devices.push_back(std::atoi(list.substr(p, q - p).c_str()));
                p = q + 1;
            }
        } else if (a == "--split") split = next();
        else if (a == "--env-importance") envImportance = true;
        else if (a == "--export-scene") exportDir = next();
        else if (a == "--list") {
            for (auto &n : scenes::list()) std::printf("%s\n", n.c_str());
            return 0;
        } else {
            std::fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (sampler != "default" && sampler != "sobol" && sampler != "pmj") {
        std::fprintf(stderr, "unknown sampler %s\n", sampler.c_str());
        return 2;
    }
    if (!split.empty() && split != "tile" && split != "sample") {
        std::fprintf(stderr, "unknown split %s\n", split.c_str());
        return 2;
    }
    Engine engine("offlinerender");
    engine.initResources();
    auto &pt = engine.renderer().rendererPathTracing();
    if (!pt.isRayTracingEnabled()) {
        std::fprintf(stderr, "path tracing backend unavailable: %s\n", pt.lastError().c_str());
        return 1;
    }
    if (devices.size() > 1 && !pt.setDevices(devices)) {
        std::fprintf(stderr, "cannot use the requested GPUs: %s\n", pt.lastError().c_str());
        return 1;
    }
    const bool isFile = scene.size() > 5 && scene.compare(scene.size() - 5, 5, ".json") == 0;
    if (isFile) {
        std::string err;
        if (!engine.importScene(scene, &err)) {
            std::fprintf(stderr, "cannot import %s: %s\n", scene.c_str(), err.c_str());
            return 2;
        }
    } else if (!scenes::build(engine, scene, opt)) {
        std::fprintf(stderr, "unknown scene '%s' (use --list, or pass a scene.json)\n", scene.c_str());
        return 2;
    }
    if (!exportDir.empty()) {
        std::string err;
        if (!engine.exportScene(exportDir, &err)) {
            std::fprintf(stderr, "cannot export to %s: %s\n", exportDir.c_str(), err.c_str());
            return 2;
        }
    }
    auto &ri = pt.renderInfo();
    if (width > 0) ri.width = (uint32_t)width;
    if (height > 0) ri.height = (uint32_t)height;
    if (spp > 0) ri.samples = (uint32_t)spp;
    if (batch > 0) ri.batchSize = (uint32_t)batch;
    if (depth > 0) ri.depth = (uint32_t)depth;
    if (!out.empty()) ri.filename = out;
    if (png) ri.fileType = FileType::PNG;
    ri.exposure = exposure;
    if (allFiles) ri.writeAllFiles = true;
    ri.lowDiscrepancySampler = sampler == "sobol";
    ri.pmjSampler = sampler == "pmj";
    ri.multiGpuSplit = split == "tile" ? 1u : (split == "sample" ? 2u : 0u);
    ri.environmentImportanceSampling = envImportance;
    if (frames > 0 && scene == "BallOnPlane") {
        /* every frame is one render() call: geometry is re-flattened and re-uploaded, textures and the environment stay on the device */
        for (int f = 0; f < frames; f++) {
            scenes::ballOnPlaneFrame(engine, f);
            if (!out.empty()) ri.filename = out + std::to_string(f);
            pt.render();
        }
    } else {
        pt.render();
    }
    const ptc_stats &st = pt.lastStats();
    std::printf("{\"backend\": \"%s\", \"gpus\": %d, \"scene\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %u, \"segments\": %llu, \"shadow_rays\": %llu, "
                "\"probe_rays\": %llu, \"render_ms\": %.3f, \"build_ms\": %.3f, \"triangles\": %llu, \"Mseg_per_s\": %.2f}\n",
                pt.backendName(), pt.deviceCount(), scene.c_str(), ri.width, ri.height, ri.samples, (unsigned long long)st.segments,
                (unsigned long long)st.shadow_rays, (unsigned long long)st.probe_rays, st.render_ms, st.build_ms,
                (unsigned long long)st.n_triangles, st.render_ms > 0 ? st.segments / st.render_ms / 1e3 : 0.0);
    engine.releaseResources();
    return 0;
}
