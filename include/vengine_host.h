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

/* The renderer of the engine is the CUDA core next to the host library (libptc_cuda.so); there is no other backend.
 * asset_root: directory that contains assets/ (NULL/"" = $VVIEWER_ASSETS or the repo root). */
PTC_API vh_engine *vh_engine_create(const char *asset_root);
PTC_API void vh_engine_destroy(vh_engine *e);
PTC_API int vh_backend_ok(vh_engine *e); /* RendererPathTracing::isRayTracingEnabled */
PTC_API const char *vh_last_error(vh_engine *e);

PTC_API const char *vh_scene_list(void); /* comma-separated recipe names */
/* build a recipe scene (vviewer_b200/host/scenes.hpp); texture_size/scale <= 0 keep the defaults */
PTC_API int vh_build_scene(vh_engine *e, const char *name, int texture_size, float scale, int camera);
/* Engine::importModel (VulkanEngine.cpp:157-192): .obj (+ .mtl) and .gltf / .glb; path relative to the asset root or absolute */
PTC_API int vh_import_model(vh_engine *e, const char *path, int import_materials);
/* addModel3D (core/SceneUtils.cpp:76-130) under the scene root, then Scene::update */
PTC_API int vh_add_model(vh_engine *e, const char *model_name);
/* scene files of core/io/Import.cpp / Export.cpp: <dir>/scene.json + <dir>/assets/ */
PTC_API int vh_import_scene(vh_engine *e, const char *scene_json);
PTC_API int vh_export_scene(vh_engine *e, const char *directory);
/* JSON text describing models / materials / textures / scene objects / camera (valid until the next call); for tests */
PTC_API const char *vh_describe(vh_engine *e);
PTC_API void vh_set_render_info(vh_engine *e, int width, int height, int samples, int batch_size, int depth); /* <= 0 keeps */
/* multi-GPU: drive several GPUs of the box from this process (the next render re-uploads the scene) ... */
PTC_API int vh_set_devices(vh_engine *e, const int *device_ids, int n_devices);
PTC_API int vh_device_count(vh_engine *e);
/* ... or one process per GPU: 128 bytes from ONE rank's vh_comm_unique_id go to every rank's vh_comm_init_rank (collective);
 * vh_render* are then collective and rank 0 alone receives / writes the image */
PTC_API int vh_comm_unique_id(vh_engine *e, uint8_t *out128);
PTC_API int vh_comm_init_rank(vh_engine *e, const uint8_t *id128, int rank, int world);
/* RenderInfo extensions; a negative value keeps the setting.  multi_gpu_split: 0 default (sample batches), 1 tiles, 2 sample
 * batches.  sampler: 0 default xorshift stream (rng_def.glsl), 1 Owen-scrambled Sobol, 2 PMJ02BN (rng_pmj.glsl). */
PTC_API void vh_set_render_options(vh_engine *e, int multi_gpu_split, int sampler, int env_importance);
PTC_API void vh_get_render_info(vh_engine *e, int *width, int *height, int *samples, int *batch_size, int *depth);

/* the flattened POD scene and render parameters that cross ptc.h (valid until the next vh_build_scene) */
PTC_API const ptc_scene_desc *vh_scene_desc(vh_engine *e);
PTC_API int vh_render_params(vh_engine *e, ptc_render_params *out);

/* RendererPathTracing::render(): upload, build, render, read back; to memory or to <filename>.hdr/.png */
PTC_API int vh_render_to_memory(vh_engine *e, float *radiance_rgba, float *albedo_rgba, float *normal_rgba);
PTC_API int vh_render(vh_engine *e, const char *filename);
/* moves the camera of a built "BallOnPlane" scene to frame 0..7 of the reference demo's orbit (PtSceneBallOnPlane.cpp:44-55) */
PTC_API int vh_set_sequence_frame(vh_engine *e, int frame);
/* output settings of RenderInfo: file_type 0 = HDR, 1 = PNG (negative keeps); exposure applies to PNG only
 * (VulkanRendererPathTracing.cpp:972-974); write_all_files adds <filename>_albedo / _normal (/ _radiance with denoise) */
PTC_API void vh_set_output(vh_engine *e, int file_type, float exposure, int write_all_files, int denoise);
/* the writer behind vh_render on a caller's image: <filename>.png (exposure, clamp, sRGB, truncation: core/ImageUtils.cpp:34-76)
 * or <filename>.hdr */
PTC_API int vh_write_image(const char *filename_no_ext, int w, int h, int channels, const float *data, int file_type, float exposure);
PTC_API int vh_get_stats(vh_engine *e, ptc_stats *out);
/* RendererPathTracing::renderProgress() (core/Renderer.hpp:37): 0..1, may be polled from another thread while vh_render* runs
 * (the reference's UI does exactly that, MainWindow.cpp:874-896) */
PTC_API float vh_render_progress(vh_engine *e);

/* Radiance HDR helpers (RGBA32F in memory, top row first) */
PTC_API int vh_read_hdr(const char *path, int *w, int *h, float *rgba_out);
PTC_API int vh_write_hdr(const char *path, int w, int h, int channels, const float *data);

/* PNG / JPEG decode from memory with the reference's conventions (stbi_load_from_memory(..., STBI_rgb_alpha),
 * src/lib/vengine/core/io/AssimpLoadModel.cpp:190; flip = stbi_set_flip_vertically_on_load, core/Image.cpp:39).
 * Call with out == NULL to get the size (w * h * channels bytes), then again with a buffer. */
PTC_API int vh_decode_image(const uint8_t *bytes, uint64_t n_bytes, int flip, int *w, int *h, int *channels, int *src_channels, uint8_t *out,
                            uint64_t out_capacity);

#ifdef __cplusplus
}
#endif
#endif
