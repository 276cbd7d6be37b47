/*
 * Shared device helpers of the B200 path-tracing core: float3 algebra, error handling, device-side
 * scene records.  sm_100a only; no other architecture is targeted.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include "../../include/ptc.h"

#define PTC_HD __host__ __device__ __forceinline__
#define PTC_D __device__ __forceinline__

/* ------------------------------------------------------------------ float3 algebra */
PTC_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PTC_HD float3 f3(float s) { return make_float3(s, s, s); }
PTC_HD float3 f3(const float4 &v) { return make_float3(v.x, v.y, v.z); }
PTC_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PTC_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PTC_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PTC_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PTC_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PTC_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
PTC_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
PTC_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
PTC_HD void operator+=(float3 &a, float3 b) { a = a + b; }
PTC_HD void operator*=(float3 &a, float3 b) { a = a * b; }
PTC_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PTC_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PTC_HD float length(float3 a) { return sqrtf(dot(a, a)); }
PTC_HD float3 normalize(float3 a) { return a / length(a); }
PTC_HD float3 fmin3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PTC_HD float3 fmax3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PTC_HD float max3(float3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
PTC_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
PTC_HD float3 clamp3(float3 a, float lo, float hi) { return f3(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)); }
PTC_HD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
PTC_HD float3 mix3(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
PTC_HD float3 exp3(float3 a) { return f3(expf(a.x), expf(a.y), expf(a.z)); }
PTC_HD float comp(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
PTC_HD bool isBlack(float3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; }
PTC_HD bool isBlackEps(float3 c, float e) { return fabsf(c.x) <= e && fabsf(c.y) <= e && fabsf(c.z) <= e; }

/* ------------------------------------------------------------------ device scene records */
/* per-instance record derived from ptc_instance at upload (112 B + ids) */
struct DInstance {
    float m[12];       /* object->world, row-major 3x4 */
    float nrm[9];      /* inverse of the upper 3x3, row-major: n_world = nrm^T * n (gl_WorldToObjectEXT trick) */
    float w2o[12];     /* world->object, row-major 3x4 */
    float volFront;    /* id.g */
    float volBack;     /* id.b */
    uint32_t material;
    uint32_t firstIndex;  /* into the index pool */
    uint32_t firstVertex; /* into the vertex pool */
    uint32_t numTriangles;
    uint32_t firstWorldTri; /* prefix of world triangles */
    uint32_t mesh;          /* index of the instance's mesh (two-level acceleration structure: which bottom-level tree) */
};

#define PTC_MAX_EMISSIVE_BOXES 16
#define TEX_WHITE 0xffffffffu /* texture whose every texel is (255, 255, 255, 255): reads as exactly 1 without a fetch */
/* texture INDEX a material's roughness slot carries on the device when its roughness map was packed into the alpha channel of a copy of
 * its normal map at upload (ptc_cuda.cu::createTextures): the normal-map tap already returned the roughness sample */
#define TEX_IN_NORMAL_ALPHA 0xfffffffeu

struct DScene {
    const ptc_vertex *vertices;
    const uint32_t *indices;
    const DInstance *instances;
    const ptc_material *materials;
    const ptc_light_data *lightData;
    const ptc_light_instance *lightInstances;
    const cudaTextureObject_t *texClasses; /* one LAYERED texture object per (width, height, sRGB) class */
    const uint32_t *texRef;                /* per texture: class << 16 | layer, or TEX_WHITE */
    cudaTextureObject_t cubemap;
    const float *envCdfV, *envCdfU; /* PTC_FLAG_ENV_IMPORTANCE tables (envdist.cuh), nullptr without an environment */
    uint32_t nInstances, nMaterials, nLightInstances, nTextures;
    uint32_t hasCubemap;
    /* acceleration structure (see lbvh.cuh) */
    const float4 *bvhNodes; /* 5 x float4 per 8-wide compressed node, breadth first, node 0 = root */
    const float4 *tris;     /* 3 x float4 per world triangle, wide-node order */
    const float4 *shading;  /* 9 x float4 per world triangle, same order (lbvh.cuh::k_gather_shading) */
    uint32_t nTris;
    uint32_t nWideNodes;
    uint32_t prmtMagic; /* 0x47000000, see traverse.cuh::byteToFloat */
    /* two-level structure (lbvh.cuh::TwoLevel; twoLevel != 0): bvhNodes starts with the top-level tree over the instances' world boxes
     * (node 0 = its root), followed by one bottom-level tree per mesh over object-space triangles; tris / shading hold the meshes'
     * triangles in bottom-level order; a top-level leaf entry p is instance tlasInst[p]; meshRoot[m] = root node of mesh m's tree */
    uint32_t twoLevel;
    const uint32_t *tlasInst;
    const uint32_t *meshRoot;
    /* scene-level switches that let whole ray types be skipped without changing any result */
    uint32_t anyEmissive;    /* some instanced material can pass the probe's emissive test */
    /* world boxes (lo, hi pairs) of the instances whose material can pass that test, when there are at most PTC_MAX_EMISSIVE_BOXES of
     * them (else 0): a probe ray that misses every box can only return black and is not traced */
    const float4 *emissiveBoxes;
    uint32_t nEmissiveBoxes;
    uint32_t anyTransparent; /* some instanced material has the transparent flag */
    uint32_t anyVolume;      /* some instance changes the volume, or the camera starts inside one */
};

/* ------------------------------------------------------------------ errors */
struct CudaError {
    std::string msg;
};
#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            throw CudaError{std::string(#expr) + " -> " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +     \
                            std::to_string(__LINE__) + ")"};                                                   \
        }                                                                                                      \
    } while (0)

template <typename T>
struct DBuf { /* owning device buffer */
    T *p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        if (count == 0) return;
        CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void upload(const T *h, size_t count, cudaStream_t s) {
        alloc(count);
        if (count) CUDA_TRY(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    size_t bytes() const { return n * sizeof(T); }
};
