/*
 * ptc.h — C-ABI of the B200 path-tracing core ("ptc").
 *
 * This is the drop-in boundary for vengine's offline GPU path-tracing render path.
 * It replaces, for that path only, what the reference reaches through
 *   class RendererPathTracing            (src/lib/vengine/core/Renderer.hpp:11-39)
 *   VulkanRendererPathTracing::render()  (src/lib/vengine/vulkan/renderers/VulkanRendererPathTracing.cpp:121-226, 791-901)
 *   vkCmdTraceRaysKHR + pt shaders          (VulkanRendererPathTracing.cpp:848-855, src/lib/vengine/shaders/pt/)
 *   driver BLAS/TLAS build               (src/lib/vengine/vulkan/resources/VulkanAccelerationStructure.cpp:137,258)
 *
 * Rules of the boundary: extern "C", plain pointers and sizes, caller-owned host memory that is
 * copied during the call, 0 = success / non-zero = error (text through ptc_last_error), no
 * exceptions, no CUDA / Vulkan / torch types.  Two libraries export exactly this surface:
 *   vviewer_b200/_lib/libptc_cuda.so   the product (hand-written sm_100a CUDA, no CPU fallback)
 *   oracle/_build/liboracle.so         the CPU restatement used ONLY by tests / smoke / bench cpu_baseline
 *
 * All matrices are column-major float[16] (glm layout), like the reference's UBOs.
 */
#ifndef PTC_H
#define PTC_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PTC_API __attribute__((visibility("default")))
#else
#define PTC_API
#endif

/* ---------------------------------------------------------------- POD records */

/* Vertex, 17 floats = 68 B, scalar layout.  Reference: src/lib/vengine/core/Mesh.hpp:15-40,
 * src/lib/vengine/shaders/include/structs.glsl:4-12 */
typedef struct ptc_vertex {
    float position[3];
    float uv[2];
    float normal[3];
    float color[3];
    float tangent[3];
    float bitangent[3];
} ptc_vertex;

/* One mesh = a range of the shared vertex / index pools (replaces the per-mesh device addresses of
 * InstanceData, src/lib/vengine/core/Instances.hpp:17-35). Indices are relative to first_vertex. */
typedef struct ptc_mesh {
    uint32_t first_index;  /* offset into indices[] (in uint32 units, multiple of 3) */
    uint32_t tri_count;
    uint32_t first_vertex; /* offset into vertices[] */
    uint32_t vertex_count;
} ptc_mesh;

/* InstanceData, 128 B. Reference: src/lib/vengine/core/Instances.hpp:17-35, structs.glsl:26-38.
 * id = (object id, front-facing volume material | -1, back-facing volume material | -1, 0) as floats. */
typedef struct ptc_instance {
    float model[16];
    float id[4];
    uint32_t material_index;
    uint32_t mesh_index; /* replaces vertexAddress/indexAddress */
    uint32_t num_triangles;
    uint32_t pad[9];
} ptc_instance;

/* MaterialData, 128 B. Reference: src/lib/vengine/vulkan/common/VulkanStructs.hpp:51-63, structs.glsl:41-52.
 * Volume materials alias: albedo.rgb = sigma_a, metallic_roughness_ao.rgb = sigma_s, emissive[0] = g. */
typedef struct ptc_material {
    float albedo[4];                /* rgb albedo, a alpha */
    float metallic_roughness_ao[4]; /* r metallic, g roughness, b ao, a transparent flag */
    float emissive[4];              /* rgb colour, a intensity */
    uint32_t tex1[4];               /* albedo, metallic, roughness, ao texture index */
    uint32_t tex2[4];               /* emissive, normal, brdf-lut (unused), alpha texture index */
    float uv_tiling[4];             /* u, v, material type (ptc_material_type), unused */
    uint32_t pad1[4];
    uint32_t pad2[4];
} ptc_material;

enum ptc_material_type { PTC_MATERIAL_PBR_STANDARD = 0, PTC_MATERIAL_SKYBOX = 1, PTC_MATERIAL_LAMBERT = 2, PTC_MATERIAL_VOLUME = 3 };

/* LightData 64 B / LightInstance 64 B. Reference: VulkanStructs.hpp:65-71, Instances.hpp:39-44,
 * packing src/lib/vengine/vulkan/VulkanInstances.cpp:73-108. */
typedef struct ptc_light_data {
    float color[4]; /* rgb, a intensity */
    uint32_t type[4];
    uint32_t pad1[4];
    uint32_t pad2[4];
} ptc_light_data;

typedef struct ptc_light_instance {
    uint32_t info[4];   /* r LightData index, g InstanceData index, b unused, a type 0 point / 1 directional / 2 mesh */
    float position[4];  /* point: world pos; directional: model * (0,0,1,0) (unnormalised); mesh: row 0 of model */
    float position1[4]; /* mesh: row 1 */
    float position2[4]; /* mesh: row 2 */
} ptc_light_instance;

/* 8-bit texture, rows stored in memory order (row 0 is sampled at v = 0). Sampling is bilinear, REPEAT,
 * LOD 0 (ray-tracing stages have no derivatives), sRGB decode before filtering when srgb != 0.
 * Reference: src/lib/vengine/vulkan/resources/VulkanTexture.cpp:219-232, VulkanUtils.cpp:535-554. */
typedef struct ptc_texture {
    uint32_t width, height;
    uint32_t channels; /* 1 or 4 */
    uint32_t srgb;
    const uint8_t *data;
    /* identity of IMMUTABLE content, chosen by the caller; 0 = none.  The reference uploads a texture once, when it is
     * created (VulkanTextures.cpp:71-145), and render() only references it; a backend may likewise keep the device copy
     * of a texture whose non-zero uid, size and format it has seen in this context's previous upload. */
    uint64_t uid;
} ptc_texture;

/* Equirectangular RGBA32F environment, rows in memory order (row 0 sampled at v = 0, i.e. the image
 * as loaded with the reference's vertical flip, src/lib/vengine/core/Image.cpp:39-40).  The backend
 * resamples it into a cubemap with face size min(width/4, 1080) like
 * VulkanRendererSkybox::createCubemap (VulkanRendererSkybox.cpp:98-135). May be NULL. */
typedef struct ptc_env {
    const float *equirect_rgba;
    uint32_t width, height;
    uint64_t uid; /* as ptc_texture.uid (the reference builds the cubemap at import time, VulkanEngine.cpp:194-232) */
} ptc_env;

typedef struct ptc_scene_desc {
    const ptc_vertex *vertices;
    uint64_t n_vertices;
    const uint32_t *indices;
    uint64_t n_indices;
    const ptc_mesh *meshes;
    uint32_t n_meshes;
    const ptc_instance *instances;
    uint32_t n_instances;
    const ptc_material *materials;
    uint32_t n_materials;
    const ptc_light_data *light_data;
    uint32_t n_light_data;
    const ptc_light_instance *light_instances; /* ComponentLight objects first, then mesh lights */
    uint32_t n_light_instances;
    const ptc_texture *textures;
    uint32_t n_textures;
    ptc_env env;
} ptc_scene_desc;

/* SceneData, 304 B. Reference: src/lib/vengine/core/Scene.hpp:27-35, structs.glsl:15-23. */
typedef struct ptc_scene_data {
    float view[16];
    float view_inverse[16];
    float projection[16];
    float projection_inverse[16];
    float exposure[4];   /* r exposure, g environment intensity, b lens radius, a focal distance */
    float background[4]; /* rgb colour, a environment type 0 solid / 1 HDRI / 2 solid + HDRI lighting */
    float volumes[4];    /* r camera volume material | -1, g znear, b zfar */
} ptc_scene_data;

enum ptc_camera_type { PTC_CAMERA_PERSPECTIVE = 0, PTC_CAMERA_ORTHOGRAPHIC = 1 };
enum ptc_split_mode { PTC_SPLIT_NONE = 0, PTC_SPLIT_TILE = 1, PTC_SPLIT_SAMPLE = 2 };

/* Render settings = RendererPathTracing::RenderInfo (Renderer.hpp:14-28) + PathTracingData
 * (VulkanRendererPathTracing.hpp:57-61) + the multi-GPU partition of this rank. */
typedef struct ptc_render_params {
    ptc_scene_data scene;
    uint32_t samples;    /* total samples per pixel; batches = samples / batch_size (remainder dropped, T7) */
    uint32_t batch_size;
    uint32_t depth;
    uint32_t width, height;
    uint32_t camera_type; /* ptc_camera_type; orthographic = true parallel rays (documented deviation T10) */
    float ortho_width, ortho_height;
    /* partition: this context renders only its share; the sum over ranks is the full image (alpha = 1 is written by rank 0
     * only, so the sum keeps it).  A multi-device context / a context with a communicator sets rank and world itself. */
    uint32_t split_mode; /* ptc_split_mode */
    uint32_t rank, world;
    uint32_t tile_size;  /* tile edge in pixels for PTC_SPLIT_TILE (0 = 32); tile (tx, ty) belongs to rank (tx + ty) mod world */
    uint32_t flags;      /* PTC_FLAG_* */
    uint32_t reserved[7];
} ptc_render_params;

#define PTC_FLAG_WORLD_ORIGIN_PROBE_PDF 1u /* use the world-space ray origin in the probe-ray pdf instead of reproducing rayNEE.rahit.glsl:122 */
#define PTC_FLAG_SAMPLER_SOBOL 4u         /* low-discrepancy sampler (shuffled, Owen-scrambled Sobol; plays the role of the reference's optional PMJ02BN sampler, rng_pmj.glsl) instead of the default xorshift stream */
/* The reference's optional sampler (SAMPLING_PMJ in defines_pt.glsl:1-5): pbrt-v4 style PMJ02BN, rng_pmj.glsl:20-107, with the
 * reference's own tables (math/PMJSequences.cpp: 16 x 16384 x 2 floats; math/BlueNoise.cpp: 48 x 128 x 128 floats).  Start
 * dimension pixel.y * width + pixel.y like raygen.rgen.glsl:59.  Integer-exact between the product and the oracle. */
#define PTC_FLAG_SAMPLER_PMJ 16u
/* Extension, off for parity: the reference never light-samples the environment (lightSampling.glsl:101-106 is a TODO, trap T3).
 * With this flag an HDRI environment (type 1 or 2) becomes one more light of the uniform light pick: directions are drawn from a
 * 512 x 256 luminance x cos(latitude) table over the equirectangular domain, the shadow chain decides visibility, and both this
 * sample and the environment radiance found by a sampled direction (miss) are weighted with the power heuristic.  Same
 * expectation as without the flag, far lower variance under small bright sources (sun). */
#define PTC_FLAG_ENV_IMPORTANCE 8u
#define PTC_FLAG_TIME_KERNELS 2u          /* bracket every kernel class with CUDA events (fills ptc_stats.*_ms; serialises launches) */

typedef struct ptc_stats {
    uint64_t segments;    /* iterations of the raygen depth loop (raygen.rgen.glsl:100-124) */
    uint64_t path_rays;   /* closest-hit traces */
    uint64_t shadow_rays; /* shadow chains started (lightSampling.glsl:108-144) */
    uint64_t shadow_hops;
    uint64_t probe_rays;  /* probe chains started (next_event_estimation.glsl) */
    uint64_t probe_hops;
    double render_ms;     /* wall time of the last ptc_render* call (device time for the CUDA backend) */
    double trace_ms;      /* time inside the closest-hit traversal kernel (CUDA events), 0 for the oracle */
    double shade_ms;
    double shadow_ms;
    double build_ms;      /* last ptc_build_accel */
    uint64_t trace_launches;
    uint64_t kernel_launches; /* all kernels launched by the last render */
    uint64_t n_triangles;     /* world-space triangles in the acceleration structure */
    uint64_t n_bvh_nodes;
    uint64_t scene_bytes;     /* device bytes held by scene + accel */
    uint64_t reserved[4];
    uint64_t upload_bytes;    /* host -> device bytes of the last ptc_upload_scene, per device (textures / the environment that stayed
                                 resident by uid are not copied and not counted) */
    double reduce_ms;         /* multi-GPU: the NCCL reduce of the accumulation buffers inside the last render */
    double reserved_ms;       /* (was: time inside the ray-binning kernel of a rejected experiment) */
    uint64_t accel_levels;    /* 1 = one tree over world triangles, 2 = per-mesh trees under an instance tree */
    uint64_t traversal_bytes; /* nodes + triangles the traversal kernels read */
} ptc_stats;

typedef struct ptc_ctx ptc_ctx;

/* ---------------------------------------------------------------- lifecycle */
/* device_ids == NULL / n_devices == 0: the calling thread's current device.
 * n_devices > 1: ONE context drives several GPUs of the box: one worker thread + streams per GPU and an NCCL
 * communicator over them (ncclCommInitAll).  ptc_upload_scene / ptc_build_accel then replicate the scene and its acceleration
 * structure on every GPU (the build is deterministic), and ptc_render* partition the render over the GPUs - by tiles
 * (PTC_SPLIT_TILE) or by sample batches (PTC_SPLIT_SAMPLE; also what PTC_SPLIT_NONE selects) - and sum the accumulation buffers
 * onto the first device with one ncclReduce per GPU over NVLink; rank / world of ptc_render_params are filled in by the context. */
PTC_API int ptc_create(ptc_ctx **out, const int *device_ids, int n_devices);
PTC_API void ptc_destroy(ptc_ctx *ctx);
PTC_API int ptc_device_count(const ptc_ctx *ctx);
/* One process per GPU (torchrun, MPI): every rank creates a single-device context; ONE rank obtains 128 opaque bytes from
 * ptc_comm_unique_id (ncclGetUniqueId) and the launcher hands them to all ranks, which each call ptc_comm_init_rank
 * (ncclCommInitRank; collective).  From then on ptc_render* of every rank take part in one partitioned render: the context fills
 * in rank / world, the buffers are reduced onto rank 0, and only rank 0's output pointers are written (others may pass NULL).
 * The oracle has no communicator: both return non-zero there. */
PTC_API int ptc_comm_unique_id(uint8_t *out128);
PTC_API int ptc_comm_init_rank(ptc_ctx *ctx, const uint8_t *id128, int rank, int world);
PTC_API const char *ptc_last_error(const ptc_ctx *ctx);
PTC_API const char *ptc_backend_name(void); /* "cuda-sm_100a" or "cpu-oracle" */

/* ---------------------------------------------------------------- scene */
/* replaces VulkanScene::updateFrame + VulkanMaterials::updateBuffers + VulkanTextures::updateTextures
 * (VulkanRendererPathTracing.cpp:202-205) */
PTC_API int ptc_upload_scene(ptc_ctx *ctx, const ptc_scene_desc *scene);
/* replaces the driver BLAS/TLAS builds: on-device LBVH over world-space triangles */
PTC_API int ptc_build_accel(ptc_ctx *ctx);

/* The tables of the reference's optional PMJ02BN sampler (PTC_FLAG_SAMPLER_PMJ), copied during the call: pmj =
 * [n_sequences = 16][n_samples = 16384][2] floats (math/PMJSequences.cpp), blue = [n_textures = 48][resolution = 128][128] floats
 * (math/BlueNoise.cpp).  Replaces VulkanRandom::createBuffers (vulkan/resources/VulkanRandom.cpp:40-72). */
PTC_API int ptc_set_sampler_tables(ptc_ctx *ctx, const float *pmj, uint32_t n_sequences, uint32_t n_samples, const float *blue, uint32_t n_textures,
                                   uint32_t resolution);

/* Hierarchy over the Morton-sorted triangles, before the collapse to the 8-wide compressed BVH:
 *   PTC_HIERARCHY_LBVH  Karras 2012 radix tree (splits at the highest differing Morton bit)
 *   PTC_HIERARCHY_PLOC  parallel locally-ordered clustering (Meister & Bittner 2018) with the given search radius
 *                       (1..32, 0 = default 16): bottom-up merging of area-nearest neighbours along the Morton order.
 * Both are built on the device and restated on the CPU by the oracle, bit for bit.  Default: PLOC (fewer node visits per
 * ray); the environment variable PTC_HIERARCHY=lbvh|ploc overrides the default at ptc_create.  Call before ptc_build_accel. */
enum ptc_hierarchy { PTC_HIERARCHY_LBVH = 0, PTC_HIERARCHY_PLOC = 1 };
PTC_API int ptc_set_build_options(ptc_ctx *ctx, uint32_t hierarchy, uint32_t ploc_radius);

/* One level or two?  The reference's driver builds a BLAS per mesh and one TLAS over the instances (VulkanScene.cpp:306-381).
 *   PTC_ACCEL_FLAT       one tree over world-space triangles (instances flattened): no per-ray transform, no overlapping instance
 *                        boxes - fastest while the traversal data fits the caches
 *   PTC_ACCEL_TWO_LEVEL  one tree per mesh in object space + one tree over the instances' world boxes; rays are transformed with the
 *                        instance's world->object matrix on the way in - for heavily instanced scenes (BASELINE C4: 42.5 M world
 *                        triangles, 0.6 M unique ones)
 *   PTC_ACCEL_AUTO       (default) flat, unless the flattened data would take more than a quarter of the device memory and the scene
 *                        instances its meshes at least twice on average (measured on C4: flat 1025 Mseg/s, two levels 648 - a forest's
 *                        instance boxes overlap; what two levels buy is memory, 8.6 GB -> 0.13 GB)
 * Call before ptc_build_accel.  Hits are the same up to the rounding of the object-space intersection. */
enum ptc_accel_mode { PTC_ACCEL_AUTO = 0, PTC_ACCEL_FLAT = 1, PTC_ACCEL_TWO_LEVEL = 2 };
PTC_API int ptc_set_accel_mode(ptc_ctx *ctx, uint32_t mode);
/* Two-level dump for the bit-exact build check: level -1 = top-level tree (primitives = instances), level m >= 0 = tree of mesh m
 * (primitives = the mesh's triangles).  node_words[n_nodes * 20] with child / primitive indices relative to the tree, prim_order[n_prims]
 * (tree position -> primitive), box6 = the tree's bounds (lo xyz, hi xyz).  NULL arrays query the counts.  Error on a single-level
 * structure. */
PTC_API int ptc_get_accel_level(ptc_ctx *ctx, int32_t level, uint64_t *n_nodes_out, uint64_t *n_prims_out, uint32_t *node_words, uint32_t *prim_order,
                                float *box6);

/* ---------------------------------------------------------------- render */
/* replaces render(VkDescriptorSet) batch loop + readback (VulkanRendererPathTracing.cpp:791-956).
 * Outputs: width*height*4 floats each, row-major, top-left origin, alpha = 1; any may be NULL. Blocking. */
PTC_API int ptc_render(ptc_ctx *ctx, const ptc_render_params *params, float *radiance_rgba, float *albedo_rgba,
                       float *normal_rgba);
/* same, but the three outputs are DEVICE pointers on the context's device (for the NCCL reduce). The
 * oracle returns an error. */
PTC_API int ptc_render_device(ptc_ctx *ctx, const ptc_render_params *params, void *d_radiance_rgba, void *d_albedo_rgba,
                              void *d_normal_rgba);
PTC_API float ptc_progress(const ptc_ctx *ctx); /* RendererPathTracing::renderProgress */
PTC_API int ptc_get_stats(ptc_ctx *ctx, ptc_stats *out);

/* ---------------------------------------------------------------- parity hooks */
/* rays: n x 8 floats (origin xyz, tmin, direction xyz, tmax). Outputs per ray: instance (-1 on miss),
 * primitive id within the instance's mesh, t, barycentric u, v. */
PTC_API int ptc_trace_closest(ptc_ctx *ctx, const float *rays, int n, int *inst, int *prim, float *t, float *u, float *v);

/* LBVH structure dump for the bit-exact build check. Arrays are caller-allocated:
 * morton[n] (sorted 64-bit keys), order[n] (sorted -> world triangle id), parent/left/right for the 2n-1 nodes
 * (internal nodes 0..n-2, leaves n-1..2n-2; children encoded the same way), aabb[(2n-1)*6].
 * Pass NULL to skip an array. Returns the triangle count through n_out. */
PTC_API int ptc_get_lbvh(ptc_ctx *ctx, uint64_t *n_out, uint64_t *morton, uint32_t *order, int32_t *parent, int32_t *left,
                         int32_t *right, float *aabb);

/* Wide (8-ary, compressed) BVH dump for the bit-exact collapse check: node_words[n_nodes * 20] (80 B per node, layout
 * in vviewer_b200/csrc/lbvh.cuh), tri_order[n_tris] (position in the traversal triangle array -> world triangle id).
 * Pass NULL arrays to query the counts. */
PTC_API int ptc_get_wide_bvh(ptc_ctx *ctx, uint64_t *n_nodes_out, uint64_t *n_tris_out, uint32_t *node_words, uint32_t *tri_order);

/* BSDF parity hooks (src/lib/vengine/shaders/include/brdfs/pbrStandard.glsl:92-165).  Per item:
 * params = albedo rgb, metallic, roughness (5 floats); wi, wo local frame (y = normal).
 * eval: out_f[3], out_pdf[1].   sample: u[3] = (u0, u1, lobe pick) -> out_wi[3], out_f[3], out_pdf[1]. */
PTC_API int ptc_bsdf_eval(ptc_ctx *ctx, int n, const float *params, const float *wi, const float *wo, float *out_f,
                          float *out_pdf);
PTC_API int ptc_bsdf_sample(ptc_ctx *ctx, int n, const float *params, const float *wo, const float *u, float *out_wi,
                            float *out_f, float *out_pdf);

/* Sampler parity hook: the rand2D() points (x, y interleaved in out_xy[2 * count]) that the sampler selected by `flags`
 * (PTC_FLAG_SAMPLER_SOBOL, PTC_FLAG_SAMPLER_PMJ or 0) hands to pixel (px, py) of an image `width` wide at dimension `dimension`, for
 * the global sample indices first_index .. first_index + count - 1.  PMJ02BN: `dimension` counts from the sample's start dimension
 * (pixel.y * width + pixel.y, raygen.rgen.glsl:59) and samplesPerPixel = first_index + count. */
#define PTC_SAMPLER_HOOK_1D 0x80000000u /* in `flags` of ptc_sampler_points: out_xy holds two consecutive rand1D() draws instead of one rand2D() */
PTC_API int ptc_sampler_points(ptc_ctx *ctx, uint32_t px, uint32_t py, uint32_t width, uint32_t first_index, uint32_t count,
                               uint32_t dimension, uint32_t flags, float *out_xy);

/* sRGB decode parity hook: out256[i] = what a texture fetch returns for the 8-bit sRGB code i (VK_FORMAT_R8G8B8A8_SRGB texels are
 * decoded before filtering, VulkanUtils.cpp:535-554).  Texture units do this with a fixed table whose entries differ from the
 * analytic curve by up to 1.2e-3 relative; the oracle carries a copy of the table (oracle/srgb_table.h), this hook is how it was
 * read and how tests/test_gpu_parity.py keeps the copy honest. */
PTC_API int ptc_srgb_table(ptc_ctx *ctx, float *out256);

/* Environment lookup parity hook: n directions (xyz) -> rgb of the backend's cubemap at LOD 0. */
PTC_API int ptc_env_lookup(ptc_ctx *ctx, int n, const float *dirs, float *out_rgb);

/* Environment importance-sampling parity hooks (PTC_FLAG_ENV_IMPORTANCE): n pairs of uniform numbers -> the sampled unit
 * directions (xyz) and their solid-angle densities; n directions -> the density the sampler has for them.  Error without an
 * environment. */
PTC_API int ptc_env_sample(ptc_ctx *ctx, int n, const float *u01, float *out_dirs, float *out_pdf);
PTC_API int ptc_env_pdf(ptc_ctx *ctx, int n, const float *dirs, float *out_pdf);

#ifdef __cplusplus
}
#endif
#endif /* PTC_H */
