/*
 * vengine_host.h — C view of the C++ host library (vviewer_b200/_lib/libvengine_host.so).
 *
 * The host library is the vengine-shaped scene model (Engine / Scene / ECS / Camera / Materials / Lights /
 * RendererPathTracing, vviewer_b200/host/vengine.hpp) that feeds the path tracer.  C++ programs use the
 * classes directly (vviewer_b200/bin/offlinerender); Python (tests/, bench.py) uses this C view.
 * It mirrors what a caller of the reference does: create the engine (VulkanEngine + initResources,
 * src/bin/offlinerender/main.cpp:13-17), create a scene (PtScene*::create), set renderInfo() and call
 * rendererPathTracing().render() (PtSceneDragonsOnPlane.cpp:69-78).
 */
#ifndef VENGINE_HOST_H
#define VENGINE_HOST_H

#include "ptc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vh_engine vh_engine;

/* backend_lib: shared library exporting ptc.h (NULL/"" = the CUDA product library next to the host
 * library). asset_root: directory that contains assets/ (NULL/"" = $VVIEWER_ASSETS or the repo root). */
PTC_API vh_engine *vh_engine_create(const char *backend_lib, const char *asset_root);
PTC_API void vh_engine_destroy(vh_engine *e);
PTC_API int vh_backend_ok(vh_engine *e); /* RendererPathTracing::isRayTracingEnabled */
PTC_API const char *vh_last_error(vh_engine *e);

PTC_API const char *vh_scene_list(void); /* comma-separated recipe names */
/* build a recipe scene (vviewer_b200/host/scenes.hpp); texture_size/scale <= 0 keep the defaults */
PTC_API int vh_build_scene(vh_engine *e, const char *name, int texture_size, float scale, int camera);
PTC_API void vh_set_render_info(vh_engine *e, int width, int height, int samples, int batch_size, int depth); /* <= 0 keeps */
PTC_API void vh_get_render_info(vh_engine *e, int *width, int *height, int *samples, int *batch_size, int *depth);

/* the flattened POD scene and render parameters that cross ptc.h (valid until the next vh_build_scene) */
PTC_API const ptc_scene_desc *vh_scene_desc(vh_engine *e);
PTC_API int vh_render_params(vh_engine *e, ptc_render_params *out);

/* RendererPathTracing::render(): upload, build, render, read back; to memory or to <filename>.hdr/.png */
PTC_API int vh_render_to_memory(vh_engine *e, float *radiance_rgba, float *albedo_rgba, float *normal_rgba);
PTC_API int vh_render(vh_engine *e, const char *filename);
PTC_API int vh_get_stats(vh_engine *e, ptc_stats *out);

/* Radiance HDR helpers (RGBA32F in memory, top row first) */
PTC_API int vh_read_hdr(const char *path, int *w, int *h, float *rgba_out);
PTC_API int vh_write_hdr(const char *path, int w, int h, int channels, const float *data);

#ifdef __cplusplus
}
#endif
#endif
