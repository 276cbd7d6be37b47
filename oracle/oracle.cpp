/*
 * ORACLE — "parity pinned by the reference's 31 golden renders" (tests/golden/reference_images, scene
 * recipes from src/bin/unittests/RenderTests.cpp; see tests/test_oracle_goldens.py).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's offline path tracer
 * (the GLSL ray-tracing pipeline under /root/reference/src/lib/vengine/shaders/pt/) behind the same
 * C-ABI as the CUDA product (include/ptc.h).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * The reference's own implementation cannot be compiled here (GLSL for a Vulkan RT pipeline; no
 * Vulkan loader, glslang, assimp or OIDN in the image — DESIGN.md §oracle), so this file follows the
 * shader sources function by function; every function cites the file:line it restates.
 *
 * Deliberate, documented differences (SURVEY.md §8a traps):
 *   T1  any-hit candidates are processed nearest-first (hardware traversal order is unspecified).
 *   T8  accumulation buffers are zero-initialised explicitly.
 *   T9  one RNG stream per (pixel, global sample index) instead of per (pixel, batch).
 *   T10 orthographic camera shoots true parallel rays.
 *   AOV albedo/normal start at zero for every sample (the reference leaves a stale payload value when a
 *       path never reaches a surface or a miss at surfaceDepth 0).
 */
#include "../include/ptc.h"
#include "omath.hpp"
#include "bsdf.hpp"
#include "accel.hpp"
#include "envdist.hpp"

#include <atomic>
#include <chrono>
#include <string>
#include <vector>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

struct Texture {
    uint32_t w = 1, h = 1, ch = 4;
    std::vector<float> px; /* w*h*4 linear floats */
};

struct Stats {
    uint64_t segments = 0, path_rays = 0, shadow_rays = 0, shadow_hops = 0, probe_rays = 0, probe_hops = 0;
    void add(const Stats &o) {
        segments += o.segments;
        path_rays += o.path_rays;
        shadow_rays += o.shadow_rays;
        shadow_hops += o.shadow_hops;
        probe_rays += o.probe_rays;
        probe_hops += o.probe_hops;
    }
};

struct InstanceX {
    ptc_instance d;
    mat4 model;
    mat3 invT; /* inverse of the upper 3x3 (rows), used as n * worldToObject */
    mat4 worldToObject;
};

mat4 inverseAffine(const mat4 &M) {
    mat3 A = upper3(M);
    mat3 Ai = inverse3(A);
    vec3 t(M.at(0, 3), M.at(1, 3), M.at(2, 3));
    vec3 ti = mul3(Ai, t);
    mat4 R;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) R.m[c * 4 + r] = (r == c) ? 1.0f : 0.0f;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) R.m[c * 4 + r] = Ai.at(r, c);
        R.m[3 * 4 + r] = -ti[r];
    }
    return R;
}

#include "srgb_table.h" /* kSrgbToLinear: the texture unit's decode table (hardware-defined) */

}  // namespace

struct ptc_ctx {
    std::string err;
    uint32_t hierarchy = 1, plocRadius = 16; /* ptc_set_build_options / PTC_HIERARCHY, like the device build */
    /* scene */
    std::vector<ptc_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<ptc_mesh> meshes;
    std::vector<InstanceX> instances;
    std::vector<ptc_material> materials;
    std::vector<ptc_light_data> lightData;
    std::vector<ptc_light_instance> lightInstances;
    std::vector<Texture> textures;
    /* environment cubemap, 6 faces of N*N rgba */
    uint32_t cubeN = 0;
    std::vector<float> cube;
    envdist::Tables envTables; /* PTC_FLAG_ENV_IMPORTANCE */
    std::vector<float> pmjTable, blueTable; /* ptc_set_sampler_tables (PTC_FLAG_SAMPLER_PMJ) */
    /* accel */
    std::vector<WorldTri> tris;
    std::vector<uint64_t> instFirstTri;
    QueryBVH bvh;
    LBVH lbvh;
    bool accelBuilt = false;
    bool anyEmissive = false; /* some instanced material can pass the probe's emissive test (|e| > 0.05) */
    /* render */
    std::atomic<float> progress{0.0f};
    ptc_stats stats{};
};

namespace {

/* ------------------------------------------------------------------ textures */
/* texture(sampler2D, uv): bilinear, REPEAT, LOD 0 (VulkanTexture.cpp:219-232; ray-tracing stages
 * have no derivatives) */
struct vec4f {
    float r, g, b, a;
};
vec4f sampleTexture(const Texture &T, float u, float v) {
    float x = u * (float)T.w - 0.5f, y = v * (float)T.h - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float ax = x - fx, ay = y - fy;
    /* NVIDIA texture units (the reference ran on an RTX GPU, the product on a B200) hold the bilinear weights in
     * 1.8 fixed point (CUDA programming guide, "Linear Filtering") */
    ax = std::floor(ax * 256.0f + 0.5f) * (1.0f / 256.0f);
    ay = std::floor(ay * 256.0f + 0.5f) * (1.0f / 256.0f);
    long ix = (long)fx, iy = (long)fy;
    auto wrap = [](long i, long n) {
        long m = i % n;
        return m < 0 ? m + n : m;
    };
    long x0 = wrap(ix, T.w), x1 = wrap(ix + 1, T.w), y0 = wrap(iy, T.h), y1 = wrap(iy + 1, T.h);
    const float *p00 = &T.px[(y0 * T.w + x0) * 4], *p10 = &T.px[(y0 * T.w + x1) * 4];
    const float *p01 = &T.px[(y1 * T.w + x0) * 4], *p11 = &T.px[(y1 * T.w + x1) * 4];
    float o[4];
    for (int c = 0; c < 4; c++) {
        float a = p00[c] * (1 - ax) + p10[c] * ax;
        float b = p01[c] * (1 - ax) + p11[c] * ax;
        o[c] = a * (1 - ay) + b * ay;
    }
    return {o[0], o[1], o[2], o[3]};
}

/* ------------------------------------------------------------------ environment */
/* include/environmentMap.glsl:1-10 */
vec2 sampleEquirectangularMap(vec3 v) {
    float ux = std::atan2(v.z, v.x) * 0.1591f + 0.5f;
    float uy = std::asin(clampf(v.y, -1.0f, 1.0f)) * 0.3183f + 0.5f;
    ux = ux + 0.25f;
    ux = ux - std::floor(ux); /* mod(x, 1.0) */
    return {ux, uy};
}
/* cube face (f, s, t in [-1,1]) -> direction; inverse of the Vulkan/CUDA cubemap face selection */
vec3 cubeDir(int f, float sc, float tc) {
    switch (f) {
        case 0: return vec3(1, -tc, -sc);
        case 1: return vec3(-1, -tc, sc);
        case 2: return vec3(sc, 1, tc);
        case 3: return vec3(sc, -1, -tc);
        case 4: return vec3(sc, -tc, 1);
        default: return vec3(-sc, -tc, -1);
    }
}
void cubeCoords(vec3 d, int &f, float &sc, float &tc, float &ma) {
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    if (ax >= ay && ax >= az) {
        ma = ax;
        if (d.x >= 0) { f = 0; sc = -d.z; tc = -d.y; } else { f = 1; sc = d.z; tc = -d.y; }
    } else if (ay >= az) {
        ma = ay;
        if (d.y >= 0) { f = 2; sc = d.x; tc = d.z; } else { f = 3; sc = d.x; tc = -d.z; }
    } else {
        ma = az;
        if (d.z >= 0) { f = 4; sc = d.x; tc = -d.y; } else { f = 5; sc = -d.x; tc = -d.y; }
    }
}
/* VulkanRendererSkybox::createCubemap (VulkanRendererSkybox.cpp:98-135, 313-420) +
 * skybox/skyboxCubemapWrite.frag.glsl:12-15: every face texel = bilinear equirect lookup along the
 * texel-centre direction */
void buildCubemap(ptc_ctx *c, const ptc_env &env) {
    c->cubeN = 0;
    c->cube.clear();
    if (!env.equirect_rgba || env.width == 0 || env.height == 0) return;
    uint32_t N = std::min(env.width / 4u, 1080u);
    if (N == 0) N = 1;
    Texture eq;
    eq.w = env.width;
    eq.h = env.height;
    eq.px.assign(env.equirect_rgba, env.equirect_rgba + (size_t)env.width * env.height * 4);
    c->cubeN = N;
    c->cube.resize((size_t)6 * N * N * 4);
#pragma omp parallel for schedule(static)
    for (int fy = 0; fy < (int)(6 * N); fy++) {
        int f = fy / N, j = fy % N;
        for (uint32_t i = 0; i < N; i++) {
            float sc = 2.0f * ((float)i + 0.5f) / (float)N - 1.0f;
            float tc = 2.0f * ((float)j + 0.5f) / (float)N - 1.0f;
            vec3 d = normalize(cubeDir(f, sc, tc));
            vec2 uv = sampleEquirectangularMap(d);
            /* REPEAT in u, and (reference sampler is REPEAT in v as well) */
            vec4f v = sampleTexture(eq, uv.x, uv.y);
            float *o = &c->cube[(((size_t)f * N + j) * N + i) * 4];
            o[0] = v.r; o[1] = v.g; o[2] = v.b; o[3] = v.a;
        }
    }
}
/* textureLod(samplerCube, dir, 0): bilinear within the face; taps that fall outside the face are
 * re-projected onto the neighbouring face (seamless filtering) */
vec3 sampleCubemap(const ptc_ctx *c, vec3 d) {
    if (c->cubeN == 0) return vec3(0.0f);
    int N = (int)c->cubeN;
    int f;
    float sc, tc, ma;
    cubeCoords(d, f, sc, tc, ma);
    float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    float x = s * N - 0.5f, y = t * N - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float ax = x - fx, ay = y - fy;
    int ix = (int)fx, iy = (int)fy;
    auto fetch = [&](int i, int j) -> vec3 {
        int ff = f;
        if (i < 0 || i >= N || j < 0 || j >= N) {
            float s2 = 2.0f * ((float)i + 0.5f) / (float)N - 1.0f;
            float t2 = 2.0f * ((float)j + 0.5f) / (float)N - 1.0f;
            vec3 dd = cubeDir(f, s2, t2);
            float sc2, tc2, ma2;
            cubeCoords(dd, ff, sc2, tc2, ma2);
            i = std::min(std::max((int)std::floor(0.5f * (sc2 / ma2 + 1.0f) * N), 0), N - 1);
            j = std::min(std::max((int)std::floor(0.5f * (tc2 / ma2 + 1.0f) * N), 0), N - 1);
        }
        const float *p = &c->cube[(((size_t)ff * N + j) * N + i) * 4];
        return vec3(p[0], p[1], p[2]);
    };
    vec3 a = fetch(ix, iy) * (1 - ax) + fetch(ix + 1, iy) * ax;
    vec3 b = fetch(ix, iy + 1) * (1 - ax) + fetch(ix + 1, iy + 1) * ax;
    return a * (1 - ay) + b * ay;
}

/* ------------------------------------------------------------------ the estimator */
struct Payload { /* pt/structs_pt.glsl:4-30 */
    vec3 radiance, beta;
    uint32_t recursionDepth = 0, surfaceDepth = 0;
    vec3 albedo, normal;
    vec3 origin, direction;
    bool stop = false, insideVolume = false;
    float vtmin = 0.001f;
    uint32_t volumeMaterialIndex = 0;
    float lastPdf = 0.0f; /* PTC_FLAG_ENV_IMPORTANCE: density of the last sampled direction (0 = camera ray) */
    Rng rng;
};

struct LightSamplingRecord { /* structs_pt.glsl:56-61 */
    vec3 direction, radiance;
    float pdf;
    bool isDeltaLight;
};

struct HitInfo { /* pt/process_hit.glsl + pt/construct_frame.glsl */
    const InstanceX *inst;
    const ptc_material *mat;
    ptc_vertex v0, v1, v2;
    vec3 bary;
    float uvx, uvy;
    vec3 worldPosition, worldNormal, worldTangent, worldBitangent;
};

struct Tracer {
    const ptc_ctx *c;
    const ptc_render_params *rp;
    uint32_t depth, totalLights; /* totalLights counts the environment when it is light-sampled */
    bool envLight = false;       /* PTC_FLAG_ENV_IMPORTANCE and an HDRI environment (type 1 or 2) */
    float zfar;
    Stats st;
    std::vector<Hit> scratch;

    static vec3 v3(const float *p) { return vec3(p[0], p[1], p[2]); }

    vec4f tex(uint32_t idx, float u, float v) const {
        if (idx >= c->textures.size()) return {1, 1, 1, 1};
        return sampleTexture(c->textures[idx], u, v);
    }

    /* pt/process_hit.glsl:1-17 */
    void processHit(const Hit &h, HitInfo &hi) const {
        const WorldTri &T = c->tris[h.tri];
        const InstanceX &I = c->instances[T.inst];
        hi.inst = &I;
        hi.mat = &c->materials[I.d.material_index];
        const ptc_mesh &M = c->meshes[I.d.mesh_index];
        const uint32_t *ind = &c->indices[M.first_index + 3 * (size_t)T.prim];
        hi.v0 = c->vertices[M.first_vertex + ind[0]];
        hi.v1 = c->vertices[M.first_vertex + ind[1]];
        hi.v2 = c->vertices[M.first_vertex + ind[2]];
        hi.bary = vec3(1.0f - h.u - h.v, h.u, h.v);
        hi.uvx = hi.v0.uv[0] * hi.bary.x + hi.v1.uv[0] * hi.bary.y + hi.v2.uv[0] * hi.bary.z;
        hi.uvy = hi.v0.uv[1] * hi.bary.x + hi.v1.uv[1] * hi.bary.y + hi.v2.uv[1] * hi.bary.z;
    }
    static vec3 interp(const float *a, const float *b, const float *cc, vec3 w) { return v3(a) * w.x + v3(b) * w.y + v3(cc) * w.z; }
    /* pt/construct_frame.glsl:1-15 (without fixFrame) */
    void constructFrame(HitInfo &hi) const {
        vec3 lp = interp(hi.v0.position, hi.v1.position, hi.v2.position, hi.bary);
        hi.worldPosition = xform_point(hi.inst->model, lp);
        vec3 ln = interp(hi.v0.normal, hi.v1.normal, hi.v2.normal, hi.bary);
        vec3 lt = interp(hi.v0.tangent, hi.v1.tangent, hi.v2.tangent, hi.bary);
        vec3 lb = interp(hi.v0.bitangent, hi.v1.bitangent, hi.v2.bitangent, hi.bary);
        /* vec3(n * gl_WorldToObjectEXT) = transpose(inverse(M3)) * n */
        hi.worldNormal = normalize(mul3_transposed(hi.inst->invT, ln));
        hi.worldTangent = normalize(mul3_transposed(hi.inst->invT, lt));
        hi.worldBitangent = normalize(mul3_transposed(hi.inst->invT, lb));
    }

    struct Sigma {
        vec3 sigma_s, sigma_t;
        float g;
    };
    /* pt/process_volume_hit.glsl:3-8 and pt/process_volume_transmittance.glsl:1-7 */
    Sigma volumeCoeffs(uint32_t volumeMaterialIndex) const {
        const ptc_material &vm = c->materials[volumeMaterialIndex];
        vec3 sigma_a = v3(vm.albedo);
        vec3 sigma_s = vmax(v3(vm.metallic_roughness_ao), vec3(EPSILON));
        return {sigma_s, sigma_a + sigma_s, vm.emissive[0]};
    }
    vec3 volumeTransmittance(uint32_t volIdx, float vtstart, float vtend) const {
        Sigma s = volumeCoeffs(volIdx);
        float dist = std::max(vtend - vtstart, EPSILON);
        return vexp(-(s.sigma_t * dist));
    }
    /* volume change at a transparent boundary: rayPrimaryPBRStandard.rchit.glsl:74-90,
     * raySecondary.rchit.glsl:63-74, rayNEE.rchit.glsl:67-78 */
    static void volumeChange(const InstanceX &I, bool flipped, bool &insideVolume, uint32_t &volIdx) {
        insideVolume = false;
        float nv = flipped ? I.d.id[1] : I.d.id[2];
        if (nv != -1.0f) {
            insideVolume = true;
            volIdx = (uint32_t)nv;
        }
    }

    /* shadow chain: pt/lightSampling.glsl:108-144 + raySecondary.rahit/.rchit/.rmiss */
    vec3 shadowChain(vec3 origin, vec3 dir, float tmax, bool insideVolume, uint32_t volIdx) {
        st.shadow_rays++;
        const float tmin = 0.0001f;
        vec3 throughput(1.0f);
        bool shadowed = false, stop = false;
        float distanceT = tmax - tmin;
        std::vector<Hit> &hits = scratch;
        for (uint32_t d = 0; d < depth; d++) {
            float vtmin = tmin;
            st.shadow_hops++;
            c->bvh.all(origin, dir, tmin, distanceT, hits);
            bool ended = false;
            for (const Hit &h : hits) {
                HitInfo hi;
                processHit(h, hi);
                float tu = hi.uvx * hi.mat->uv_tiling[0], tv = hi.uvy * hi.mat->uv_tiling[1];
                float alpha = hi.mat->albedo[3] * tex(hi.mat->tex2[3], tu, tv).r;
                bool isTransparent = hi.mat->metallic_roughness_ao[3] >= 0.99f;
                if (!isTransparent) { /* raySecondary.rahit.glsl:42-47 */
                    stop = true;
                    shadowed = true;
                    ended = true;
                    break;
                }
                throughput = throughput * (1.0f - alpha); /* :52 */
                if (hi.inst->d.id[1] != hi.inst->d.id[2]) { /* :55-58 accept -> raySecondary.rchit.glsl:32-78 */
                    constructFrame(hi);
                    bool flipped = dot(hi.worldNormal, dir) > 0;
                    if (insideVolume) {
                        throughput = throughput * volumeTransmittance(volIdx, vtmin, h.t);
                        if (max3(throughput) < EPSILON) {
                            stop = true;
                            shadowed = true;
                            ended = true;
                            break;
                        }
                    }
                    vtmin = h.t;
                    volumeChange(*hi.inst, flipped, insideVolume, volIdx);
                    origin = hi.worldPosition;
                    ended = true;
                    break;
                }
                if (max3(throughput) > EPSILON) { /* :61-65 ignore */
                    shadowed = false;
                    continue;
                }
                stop = true; /* :66-71 */
                shadowed = true;
                ended = true;
                break;
            }
            if (!ended) { /* raySecondary.rmiss.glsl:17-44 */
                stop = true;
                shadowed = false;
                if (insideVolume) {
                    float vtend = std::min((float)(uint32_t)zfar, distanceT);
                    throughput = throughput * volumeTransmittance(volIdx, vtmin, vtend);
                    shadowed = !(max3(throughput) > EPSILON);
                }
            }
            distanceT -= vtmin; /* lightSampling.glsl:131 */
            if (stop) break;
        }
        if (shadowed) throughput = vec3(0.0f);
        return throughput;
    }

    /* pt/lightSampling.glsl:1-147 */
    LightSamplingRecord sampleLight(Payload &P, vec3 originPosition) {
        LightSamplingRecord lsr;
        lsr.isDeltaLight = true;
        lsr.radiance = vec3(0.0f);
        lsr.pdf = 1.0f;
        lsr.direction = vec3(0, 1, 0);
        if (totalLights == 0) return lsr;
        float pdf = 1.0f / (float)totalLights;
        uint32_t randomLight = (uint32_t)(P.rng.rand1D() * (float)totalLights);
        if (randomLight >= totalLights) randomLight = totalLights - 1;
        float tmax = 10000.0f;
        if (envLight && randomLight == totalLights - 1) {
            /* extension (no reference counterpart, trap T3): the environment is the last light; direction by luminance
             * importance sampling (oracle/envdist.hpp), radiance as the miss shader would return it after a bounce */
            vec2 u2 = P.rng.rand2D();
            float eu, ev, puv;
            envdist::sampleUV(c->envTables, u2.x, u2.y, eu, ev, puv);
            vec3 d;
            envdist::direction(eu, ev, d.x, d.y, d.z);
            lsr.direction = d;
            lsr.radiance = sampleCubemap(c, d) * (rp->scene.background[3] == 1.0f ? rp->scene.exposure[1] : 1.0f);
            lsr.pdf = pdf * envdist::solidAnglePdf(puv, ev);
            lsr.isDeltaLight = false;
            if (!(lsr.pdf > 0.0f)) lsr.radiance = vec3(0.0f);
            if (!isBlack(lsr.radiance)) {
                vec3 thr = shadowChain(originPosition, lsr.direction, tmax, P.insideVolume, P.volumeMaterialIndex);
                lsr.radiance = lsr.radiance * thr;
            }
            return lsr;
        }
        const ptc_light_instance &light = c->lightInstances[randomLight];
        if (light.info[3] == 0) {
            const ptc_light_data &ld = c->lightData[light.info[0]];
            vec3 lp = v3(light.position);
            vec3 direction = lp - originPosition;
            tmax = length(direction);
            lsr.direction = direction / tmax;
            float dist = length(originPosition - lp); /* include/lighting.glsl:1-5 */
            lsr.radiance = v3(ld.color) * (1.0f / (dist * dist)) * ld.color[3];
            lsr.pdf = pdf * 1.0f;
            lsr.isDeltaLight = true;
        } else if (light.info[3] == 1) {
            const ptc_light_data &ld = c->lightData[light.info[0]];
            lsr.direction = -v3(light.position); /* unnormalised, trap T4 */
            tmax = zfar;
            lsr.radiance = v3(ld.color) * ld.color[3];
            lsr.pdf = pdf * 1.0f;
            lsr.isDeltaLight = true;
        } else if (light.info[3] == 2) {
            const InstanceX &I = c->instances[light.info[1]];
            const ptc_mesh &M = c->meshes[I.d.mesh_index];
            const ptc_material &material = c->materials[I.d.material_index];
            /* transform rebuilt from the three rows, lightSampling.glsl:53-54 */
            mat4 transform;
            for (int col = 0; col < 4; col++) {
                transform.m[col * 4 + 0] = light.position[col];
                transform.m[col * 4 + 1] = light.position1[col];
                transform.m[col * 4 + 2] = light.position2[col];
                transform.m[col * 4 + 3] = col == 3 ? 1.0f : 0.0f;
            }
            uint32_t numTriangles = I.d.num_triangles;
            uint32_t randomTriangle = (uint32_t)(P.rng.rand1D() * (float)numTriangles);
            if (randomTriangle >= numTriangles) randomTriangle = numTriangles - 1;
            const uint32_t *ind = &c->indices[M.first_index + 3 * (size_t)randomTriangle];
            float trianglePdf = 1.0f / (float)numTriangles;
            vec2 bc = uniformSampleTriangle(P.rng.rand2D());
            vec3 sb(bc.x, bc.y, 1.0f - bc.x - bc.y);
            const ptc_vertex &v0 = c->vertices[M.first_vertex + ind[0]];
            const ptc_vertex &v1 = c->vertices[M.first_vertex + ind[1]];
            const ptc_vertex &v2 = c->vertices[M.first_vertex + ind[2]];
            vec3 p0 = xform_point(transform, v3(v0.position));
            vec3 p1 = xform_point(transform, v3(v1.position));
            vec3 p2 = xform_point(transform, v3(v2.position));
            float triangleArea = 0.5f * length(cross(p1 - p0, p2 - p0));
            float sampledPointPdf = 1.0f / triangleArea;
            vec3 sampledPoint = v3(v0.position) * sb.x + v3(v1.position) * sb.y + v3(v2.position) * sb.z;
            vec3 sampledNormal = v3(v0.normal) * sb.x + v3(v1.normal) * sb.y + v3(v2.normal) * sb.z;
            float su = sb.x * v0.uv[0] + sb.y * v1.uv[0] + sb.z * v2.uv[0];
            float sv = sb.x * v0.uv[1] + sb.y * v1.uv[1] + sb.z * v2.uv[1];
            sampledPoint = xform_point(transform, sampledPoint);
            sampledNormal = mul3_transposed(inverse3(upper3(transform)), sampledNormal); /* not normalised, trap T5 */
            vec3 direction = sampledPoint - originPosition;
            tmax = length(direction);
            lsr.direction = direction / tmax;
            float dotProduct = dot(-lsr.direction, sampledNormal);
            if (dotProduct > 0) {
                vec4f et = tex(material.tex2[0], su * material.uv_tiling[0], sv * material.uv_tiling[1]);
                vec3 emissive = v3(material.emissive) * material.emissive[3] * vec3(et.r, et.g, et.b);
                lsr.radiance = emissive;
                float dd = length(originPosition - sampledPoint);
                lsr.pdf = pdf * trianglePdf * sampledPointPdf * (dd * dd) / dotProduct;
            } else {
                lsr.radiance = vec3(0.0f);
                lsr.pdf = 0.0f;
            }
            lsr.isDeltaLight = false;
        }
        if (!isBlack(lsr.radiance)) {
            vec3 thr = shadowChain(originPosition, lsr.direction, tmax, P.insideVolume, P.volumeMaterialIndex);
            lsr.radiance = lsr.radiance * thr;
        }
        return lsr;
    }

    /* pt/next_event_estimation.glsl:1-33 + rayNEE.rahit/.rchit/.rmiss */
    void nextEventEstimation(Payload &P, float sampleDirectionPDF) {
        /* result-identical shortcut: without an emissive instance the probe can only return black */
        if (!c->anyEmissive) return;
        st.probe_rays++;
        const float tmin = 0.0001f;
        const float tmax = zfar;
        bool stop = false;
        vec3 origin = P.origin;
        vec3 dir = P.direction;
        vec3 throughput(1.0f), emissive(0.0f);
        float pdf = 0.0f;
        bool insideVolume = P.insideVolume;
        uint32_t volIdx = P.volumeMaterialIndex;
        std::vector<Hit> &hits = scratch;
        for (uint32_t d = 0; d < depth; d++) {
            float vtmin = tmin;
            st.probe_hops++;
            c->bvh.all(origin, dir, tmin, tmax, hits);
            bool ended = false;
            for (const Hit &h : hits) {
                HitInfo hi;
                processHit(h, hi);
                float tu = hi.uvx * hi.mat->uv_tiling[0], tv = hi.uvy * hi.mat->uv_tiling[1];
                float alpha = hi.mat->albedo[3] * tex(hi.mat->tex2[3], tu, tv).r;
                vec4f et = tex(hi.mat->tex2[0], tu, tv);
                vec3 em = v3(hi.mat->emissive) * hi.mat->emissive[3] * vec3(et.r, et.g, et.b);
                bool isTransparent = hi.mat->metallic_roughness_ao[3] >= 0.99f;
                if (isBlack(em, 0.05f)) { /* rayNEE.rahit.glsl:44-71 */
                    if (!isTransparent) {
                        stop = true;
                        throughput = vec3(0.0f);
                        emissive = vec3(0.0f);
                        ended = true;
                        break;
                    }
                    throughput = throughput * (1.0f - alpha);
                    if (hi.inst->d.id[1] != hi.inst->d.id[2]) { /* accept -> rayNEE.rchit.glsl:33-82 */
                        constructFrame(hi);
                        bool flipped = dot(hi.worldNormal, dir) > 0;
                        if (insideVolume) {
                            throughput = throughput * volumeTransmittance(volIdx, vtmin, h.t);
                            if (max3(throughput) < EPSILON) {
                                stop = true;
                                throughput = vec3(0.0f);
                                emissive = vec3(0.0f);
                                ended = true;
                                break;
                            }
                        }
                        vtmin = h.t;
                        volumeChange(*hi.inst, flipped, insideVolume, volIdx);
                        origin = hi.worldPosition;
                        ended = true;
                        break;
                    }
                    continue; /* ignoreIntersectionEXT */
                }
                /* emissive surface, rayNEE.rahit.glsl:73-131 */
                stop = true;
                ended = true;
                if (insideVolume) {
                    throughput = throughput * volumeTransmittance(volIdx, vtmin, h.t);
                    if (max3(throughput) < EPSILON) {
                        throughput = vec3(0.0f);
                        emissive = vec3(0.0f);
                        break;
                    }
                }
                constructFrame(hi);
                bool flipped = dot(hi.worldNormal, dir) > 0;
                if (flipped) {
                    emissive = vec3(0.0f);
                    throughput = vec3(0.0f);
                    break;
                }
                emissive = em;
                vec3 w0 = xform_point(hi.inst->model, v3(hi.v0.position));
                vec3 w1 = xform_point(hi.inst->model, v3(hi.v1.position));
                vec3 w2 = xform_point(hi.inst->model, v3(hi.v2.position));
                float triangleArea = 0.5f * length(cross(w1 - w0, w2 - w0));
                float sampledPointPdf = (1.0f / (float)hi.inst->d.num_triangles) * (1.0f / triangleArea);
                float dotProduct = dot(-dir, hi.worldNormal);
                if (dotProduct > 0) {
                    /* rayNEE.rahit.glsl:122 measures from gl_ObjectRayOriginEXT (trap T6) */
                    vec3 ro = (rp->flags & PTC_FLAG_WORLD_ORIGIN_PROBE_PDF) ? origin : xform_point(hi.inst->worldToObject, origin);
                    float dd = length(ro - hi.worldPosition);
                    float lightDirectPdf = sampledPointPdf * (dd * dd) / dotProduct;
                    pdf = lightDirectPdf * (1.0f / (float)totalLights);
                } else {
                    pdf = 0;
                    emissive = vec3(0.0f);
                }
                break;
            }
            if (!ended) { /* rayNEE.rmiss.glsl:12-19 */
                stop = true;
                emissive = vec3(0.0f);
                pdf = 0;
            }
            if (stop) break;
        }
        if (!isBlack(emissive)) {
            float w = PowerHeuristic(1, sampleDirectionPDF, 1, pdf);
            P.radiance += throughput * emissive * P.beta * w;
        }
    }

    /* pt/russian_roulette.glsl:1-11; returns true if the path was terminated */
    bool russianRoulette(Payload &P) {
        float rrsample = P.rng.rand1D();
        if (P.recursionDepth > 3) {
            float maxBeta = max3(P.beta);
            if (rrsample >= maxBeta) {
                P.stop = true;
                return true;
            }
            P.beta = P.beta * (1.0f / maxBeta);
        }
        return false;
    }

    /* pt/process_volume_hit.glsl:1-80; returns sampledMedium */
    bool processVolumeHit(Payload &P, vec3 rayDir, float vtstart, float vtend) {
        Sigma s = volumeCoeffs(P.volumeMaterialIndex);
        float g = std::max(std::min(s.g, 0.99f), -0.99f);
        vec3 worldRayDirection = normalize(rayDir);
        vec3 wo = -worldRayDirection;
        float distance_inside_volume = std::max(vtend - vtstart, EPSILON);
        uint32_t channel = std::min((uint32_t)(P.rng.rand1D() * 3.0f), 2u);
        float hit_distance = -std::log(1.0f - P.rng.rand1D()) / s.sigma_t[channel];
        bool sampledMedium = hit_distance < distance_inside_volume;
        float transmittanceDistance = std::min(hit_distance, distance_inside_volume);
        vec3 transmittance = vexp(-(s.sigma_t * transmittanceDistance));
        vec3 density = sampledMedium ? (s.sigma_t * transmittance) : transmittance;
        float pdf = 0;
        for (int i = 0; i < 3; i++) pdf += density[i];
        pdf *= 0.3333333f;
        if (pdf == 0) pdf = 1.0f;
        vec3 F = sampledMedium ? (transmittance * s.sigma_s / pdf) : (transmittance / pdf);
        P.beta *= F;
        if (sampledMedium) {
            vec3 scatteringPosition = P.origin + worldRayDirection * (vtstart + hit_distance);
            LightSamplingRecord lsr = sampleLight(P, scatteringPosition);
            if (!isBlack(lsr.radiance)) {
                float p = HG_p(wo, lsr.direction, g);
                vec3 Fp(p);
                if (!isBlack(Fp)) {
                    if (lsr.isDeltaLight) {
                        P.radiance += lsr.radiance * Fp * P.beta / lsr.pdf;
                    } else {
                        float w = PowerHeuristic(1, lsr.pdf, 1, p);
                        P.radiance += lsr.radiance * Fp * P.beta * w / lsr.pdf;
                    }
                }
            }
            vec3 sampleDirectionWorld;
            float sampleDirectionPDF = HG_Sample(wo, sampleDirectionWorld, P.rng.rand2D(), g);
            P.origin = scatteringPosition;
            P.direction = sampleDirectionWorld;
            P.lastPdf = sampleDirectionPDF;
            nextEventEstimation(P, sampleDirectionPDF);
        }
        return sampledMedium;
    }

    /* rayPrimaryLambert.rchit.glsl:45-163 and rayPrimaryPBRStandard.rchit.glsl:45-175 */
    void closestHit(Payload &P, const Hit &h, vec3 rayDir) {
        HitInfo hi;
        processHit(h, hi);
        constructFrame(hi);
        Frame frame{hi.worldNormal, hi.worldTangent, hi.worldBitangent};
        bool flipped = fixFrame(frame.normal, frame.tangent, frame.bitangent, rayDir);
        const ptc_material &material = *hi.mat;
        const bool isLambert = (int)material.uv_tiling[2] == PTC_MATERIAL_LAMBERT;

        bool sampledMedium = false;
        if (P.insideVolume) sampledMedium = processVolumeHit(P, rayDir, P.vtmin, h.t);

        if (!sampledMedium) {
            float tu = hi.uvx * material.uv_tiling[0], tv = hi.uvy * material.uv_tiling[1];
            float alpha = material.albedo[3] * tex(material.tex2[3], tu, tv).r;
            float transparent = material.metallic_roughness_ao[3];
            if (transparent > 0) {
                float r = P.rng.rand1D();
                if (alpha < EPSILON || r > alpha) {
                    if (hi.inst->d.id[1] != hi.inst->d.id[2]) volumeChange(*hi.inst, flipped, P.insideVolume, P.volumeMaterialIndex);
                    P.origin = hi.worldPosition;
                    P.direction = normalize(rayDir);
                    return;
                }
            }
            vec4f nt = tex(material.tex2[1], tu, tv);
            applyNormalToFrame(frame, processNormalFromNormalMap(vec3(nt.r, nt.g, nt.b)));

            vec4f at = tex(material.tex1[0], tu, tv);
            vec3 albedo = v3(material.albedo) * vec3(at.r, at.g, at.b);
            vec4f et = tex(material.tex2[0], tu, tv);
            vec3 emissive = v3(material.emissive) * material.emissive[3] * vec3(et.r, et.g, et.b);
            PBRStandard pbr;
            pbr.albedo = albedo;
            pbr.metallic = 0;
            pbr.roughness = 1;
            if (!isLambert) {
                pbr.metallic = material.metallic_roughness_ao[0] * tex(material.tex1[1], tu, tv).r;
                pbr.roughness = material.metallic_roughness_ao[1] * tex(material.tex1[2], tu, tv).r;
                pbr.roughness = std::max(pbr.roughness, 0.035f);
            }
            if (P.surfaceDepth == 0) {
                P.albedo = albedo;
                P.normal = frame.normal * 0.5f + vec3(0.5f);
            }
            if (P.surfaceDepth == 0 && !isBlack(emissive, isLambert ? 0.05f : 0.1f) && !flipped) {
                P.radiance += emissive * P.beta;
                P.stop = true;
                return;
            }
            P.surfaceDepth += 1;

            vec3 wo = worldToLocal(frame, -rayDir);

            LightSamplingRecord lsr = sampleLight(P, hi.worldPosition);
            if (!isBlack(lsr.radiance)) {
                vec3 wi = worldToLocal(frame, lsr.direction);
                vec3 F;
                float bsdfPdf;
                if (isLambert) {
                    float cosTheta = clampf(wi.y, 0.0f, 1.0f);
                    F = albedo * INV_PI * cosTheta;
                    bsdfPdf = cosineSampleHemispherePdf(cosTheta);
                } else {
                    F = evalPBRStandard(pbr, wi, wo, normalize(wo + wi));
                    bsdfPdf = 0;
                }
                if (!isBlack(F)) {
                    if (lsr.isDeltaLight) {
                        P.radiance += lsr.radiance * F * P.beta / lsr.pdf;
                    } else {
                        if (!isLambert) bsdfPdf = pdfPBRStandard(wi, wo, pbr);
                        float w = PowerHeuristic(1, lsr.pdf, 1, bsdfPdf);
                        P.radiance += lsr.radiance * F * P.beta * w / lsr.pdf;
                    }
                }
            }

            float sampleDirectionPDF;
            vec3 sampleDirectionLocal;
            if (isLambert) {
                sampleDirectionLocal = cosineSampleHemisphere(P.rng.rand2D(), sampleDirectionPDF);
                P.origin = hi.worldPosition;
                P.direction = localToWorld(frame, sampleDirectionLocal);
                P.beta *= albedo;
            } else {
                vec2 u2 = P.rng.rand2D();
                float u1 = P.rng.rand1D();
                vec3 F = samplePBRStandard(sampleDirectionLocal, wo, sampleDirectionPDF, pbr, u2, u1);
                P.origin = hi.worldPosition;
                P.direction = localToWorld(frame, sampleDirectionLocal);
                if (isBlack(F)) {
                    P.stop = true;
                    return;
                }
                P.beta *= vclamp(F / sampleDirectionPDF, 0.0f, 1.0f);
            }
            P.lastPdf = sampleDirectionPDF;
            nextEventEstimation(P, sampleDirectionPDF);
        }
        russianRoulette(P);
    }

    /* PTC_FLAG_ENV_IMPORTANCE: power-heuristic weight of an environment hit by a sampled direction against the light
     * sampler's density for the same direction (1 for camera rays and when the option is off) */
    float envMisWeight(const Payload &P, vec3 rayDir) const {
        if (!envLight || !(P.lastPdf > 0.0f)) return 1.0f;
        vec2 uv = sampleEquirectangularMap(normalize(rayDir));
        float pl = (1.0f / (float)totalLights) * envdist::solidAnglePdf(envdist::pdfUV(c->envTables, uv.x, uv.y), uv.y);
        return PowerHeuristic(1, P.lastPdf, 1, pl);
    }

    /* rayPrimary.rmiss.glsl:40-106 */
    void miss(Payload &P, vec3 rayDir, float rayTmax) {
        bool sampledMedium = false;
        if (P.insideVolume) {
            float vtend = std::min((float)(uint32_t)zfar, rayTmax);
            sampledMedium = processVolumeHit(P, rayDir, P.vtmin, vtend);
        }
        if (!sampledMedium) {
            P.stop = true;
            const float *bg = rp->scene.background;
            if (bg[3] == 0.0f) {
                vec3 col = v3(bg);
                if (P.surfaceDepth == 0) {
                    P.albedo = col;
                    P.normal = vec3(0.0f);
                }
                P.radiance += col * P.beta;
            } else if (bg[3] == 1.0f) {
                vec3 col = sampleCubemap(c, rayDir) * rp->scene.exposure[1];
                if (P.surfaceDepth == 0) {
                    P.albedo = col;
                    P.normal = vec3(0.0f);
                }
                P.radiance += col * P.beta * envMisWeight(P, rayDir);
            } else if (bg[3] == 2.0f) {
                vec3 env = sampleCubemap(c, rayDir);
                vec3 solid = v3(bg);
                if (P.surfaceDepth == 0) {
                    P.albedo = solid;
                    P.normal = vec3(0.0f);
                    P.radiance += solid * P.beta;
                } else {
                    P.radiance += env * P.beta * envMisWeight(P, rayDir);
                }
            }
            return;
        }
        russianRoulette(P);
    }

    /* raygen.rgen.glsl:55-129 for one sample of one pixel */
    void samplePixel(uint32_t px, uint32_t py, uint32_t sampleIndex, vec3 &radiance, vec3 &albedo, vec3 &normal) {
        Payload P;
        if (rp->flags & PTC_FLAG_SAMPLER_PMJ)
            P.rng.initPmj(px, py, rp->width, sampleIndex, (rp->samples / rp->batch_size) * rp->batch_size, PmjTables{c->pmjTable.data(), c->blueTable.data()});
        else
            P.rng.init(px, py, rp->width, sampleIndex, (rp->flags & PTC_FLAG_SAMPLER_SOBOL) != 0u);
        const mat4 projInv = mat4_from(rp->scene.projection_inverse);
        const mat4 viewInv = mat4_from(rp->scene.view_inverse);
        float lensRadius = rp->scene.exposure[2];
        float focalDistance = rp->scene.exposure[3];
        bool insideVolume = rp->scene.volumes[0] != -1.0f;

        vec2 off = P.rng.rand2D();
        float inU = ((float)px + off.x) / (float)rp->width;
        float inV = ((float)py + off.y) / (float)rp->height;
        float dx = inU * 2.0f - 1.0f, dy = inV * 2.0f - 1.0f;
        vec3 originCam(0, 0, 0), dirCam;
        if (rp->camera_type == PTC_CAMERA_ORTHOGRAPHIC) {
            /* documented deviation T10: true orthographic camera (Camera.cpp:129-153); y grows downwards
             * in the image exactly like the flipped perspective projection */
            originCam = vec3(dx * 0.5f * rp->ortho_width, -dy * 0.5f * rp->ortho_height, 0);
            dirCam = vec3(0, 0, -1);
        } else {
            /* target = projectionInverse * (d.x, d.y, 1, 1); no w-divide (raygen.rgen.glsl:67-71) */
            vec3 target(projInv.at(0, 0) * dx + projInv.at(0, 1) * dy + projInv.at(0, 2) + projInv.at(0, 3),
                        projInv.at(1, 0) * dx + projInv.at(1, 1) * dy + projInv.at(1, 2) + projInv.at(1, 3),
                        projInv.at(2, 0) * dx + projInv.at(2, 1) * dy + projInv.at(2, 2) + projInv.at(2, 3));
            dirCam = normalize(target);
        }
        if (lensRadius > 0) { /* raygen.rgen.glsl:74-85 */
            vec2 lo = P.rng.rand2D();
            vec2 disk = concentricSampleDisk(lo);
            vec3 lensOrigin(lensRadius * disk.x, lensRadius * disk.y, 0);
            float ft = focalDistance / (-dirCam.z);
            vec3 focusPoint = originCam + dirCam * ft;
            originCam = originCam + lensOrigin;
            dirCam = normalize(focusPoint - originCam);
        }
        vec3 origin = xform_point(viewInv, originCam);
        vec3 direction = xform_dir(viewInv, dirCam);

        P.origin = origin;
        P.direction = direction;
        P.surfaceDepth = 0;
        P.insideVolume = insideVolume;
        P.volumeMaterialIndex = insideVolume ? (uint32_t)(int)rp->scene.volumes[0] : 0u;
        P.albedo = vec3(0.0f);
        P.normal = vec3(0.0f);
        P.beta = vec3(1.0f);
        P.radiance = vec3(0.0f);
        for (uint32_t d = 0; d < depth; d++) {
            P.stop = false;
            P.recursionDepth = d;
            P.vtmin = 0.001f;
            st.segments++;
            st.path_rays++;
            Hit h = c->bvh.closest(origin, direction, 0.001f, 10000.0f);
            if (h.tri >= 0)
                closestHit(P, h, direction);
            else
                miss(P, direction, 10000.0f);
            if (P.stop) break;
            origin = P.origin;
            direction = P.direction;
        }
        radiance = P.radiance;
        albedo = P.albedo;
        normal = P.normal;
    }
};

int fail(ptc_ctx *c, const char *msg) {
    if (c) c->err = msg;
    return 1;
}

}  // namespace

/* ====================================================================== C-ABI */
extern "C" {

PTC_API const char *ptc_backend_name(void) { return "cpu-oracle"; }

PTC_API int ptc_create(ptc_ctx **out, const int *, int) {
    if (!out) return 1;
    *out = new ptc_ctx();
    if (const char *h = getenv("PTC_HIERARCHY")) {
        if (!strcmp(h, "lbvh")) (*out)->hierarchy = 0;
        if (!strcmp(h, "ploc")) (*out)->hierarchy = 1;
    }
    if (const char *r = getenv("PTC_PLOC_RADIUS")) {
        const int v = atoi(r);
        if (v >= 1 && v <= 32) (*out)->plocRadius = (uint32_t)v;
    }
    return 0;
}
PTC_API void ptc_destroy(ptc_ctx *ctx) { delete ctx; }
PTC_API int ptc_device_count(const ptc_ctx *ctx) { return ctx ? 1 : 0; }
PTC_API int ptc_comm_unique_id(uint8_t *) { return 1; } /* the oracle has no communicator */
PTC_API int ptc_comm_init_rank(ptc_ctx *c, const uint8_t *, int, int) { return fail(c, "the CPU oracle has no communicator"); }
PTC_API const char *ptc_last_error(const ptc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

PTC_API int ptc_upload_scene(ptc_ctx *c, const ptc_scene_desc *s) {
    if (!c || !s) return fail(c, "null argument");
    c->vertices.assign(s->vertices, s->vertices + s->n_vertices);
    c->indices.assign(s->indices, s->indices + s->n_indices);
    c->meshes.assign(s->meshes, s->meshes + s->n_meshes);
    c->materials.assign(s->materials, s->materials + s->n_materials);
    c->lightData.assign(s->light_data, s->light_data + s->n_light_data);
    c->lightInstances.assign(s->light_instances, s->light_instances + s->n_light_instances);
    c->instances.resize(s->n_instances);
    for (uint32_t i = 0; i < s->n_instances; i++) {
        InstanceX &I = c->instances[i];
        I.d = s->instances[i];
        I.model = mat4_from(I.d.model);
        I.invT = inverse3(upper3(I.model));
        I.worldToObject = inverseAffine(I.model);
        if (I.d.mesh_index >= s->n_meshes) return fail(c, "instance mesh index out of range");
        if (I.d.material_index >= s->n_materials) return fail(c, "instance material index out of range");
    }
    c->anyEmissive = false;
    for (uint32_t i = 0; i < s->n_instances; i++) {
        const ptc_material &m = c->materials[c->instances[i].d.material_index];
        for (int k = 0; k < 3; k++)
            if (std::fabs(m.emissive[3] * m.emissive[k]) > 0.05f) c->anyEmissive = true;
    }
    c->textures.resize(s->n_textures);
    for (uint32_t t = 0; t < s->n_textures; t++) {
        const ptc_texture &in = s->textures[t];
        Texture &T = c->textures[t];
        T.w = in.width;
        T.h = in.height;
        T.ch = in.channels;
        if (in.channels != 1 && in.channels != 4) return fail(c, "texture channels must be 1 or 4");
        T.px.resize((size_t)T.w * T.h * 4);
        for (size_t p = 0; p < (size_t)T.w * T.h; p++) {
            float v[4] = {0, 0, 0, 1};
            for (uint32_t k = 0; k < in.channels; k++) {
                float f = in.data[p * in.channels + k] / 255.0f;
                if (in.srgb && k < 3) f = kSrgbToLinear[in.data[p * in.channels + k]]; /* VK_FORMAT_R8G8B8A8_SRGB: decoded before filtering */
                v[k] = f;
            }
            for (int k = 0; k < 4; k++) T.px[p * 4 + k] = v[k];
        }
    }
    buildCubemap(c, s->env);
    envdist::build(s->env.equirect_rgba, s->env.width, s->env.height, c->envTables);

    /* flatten to world space: instance order, then primitive order */
    c->tris.clear();
    c->instFirstTri.assign(s->n_instances + 1, 0);
    for (uint32_t i = 0; i < s->n_instances; i++) {
        const InstanceX &I = c->instances[i];
        const ptc_mesh &M = c->meshes[I.d.mesh_index];
        c->instFirstTri[i] = c->tris.size();
        for (uint32_t p = 0; p < M.tri_count; p++) {
            const uint32_t *ind = &c->indices[M.first_index + 3 * (size_t)p];
            vec3 a = xform_point(I.model, vec3(c->vertices[M.first_vertex + ind[0]].position[0], c->vertices[M.first_vertex + ind[0]].position[1],
                                               c->vertices[M.first_vertex + ind[0]].position[2]));
            vec3 b = xform_point(I.model, vec3(c->vertices[M.first_vertex + ind[1]].position[0], c->vertices[M.first_vertex + ind[1]].position[1],
                                               c->vertices[M.first_vertex + ind[1]].position[2]));
            vec3 cc = xform_point(I.model, vec3(c->vertices[M.first_vertex + ind[2]].position[0], c->vertices[M.first_vertex + ind[2]].position[1],
                                                c->vertices[M.first_vertex + ind[2]].position[2]));
            c->tris.push_back(WorldTri{a, b - a, cc - a, i, p});
        }
    }
    c->instFirstTri[s->n_instances] = c->tris.size();
    c->accelBuilt = false;
    return 0;
}

PTC_API int ptc_build_accel(ptc_ctx *c) {
    if (!c) return 1;
    auto t0 = std::chrono::steady_clock::now();
    c->bvh.build(c->tris);
    c->lbvh = LBVH();
    c->accelBuilt = true;
    c->stats.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->stats.n_triangles = c->tris.size();
    c->stats.n_bvh_nodes = c->bvh.nodes.size();
    return 0;
}

PTC_API int ptc_render(ptc_ctx *c, const ptc_render_params *rp, float *radiance, float *albedo, float *normal) {
    if (!c || !rp) return fail(c, "null argument");
    if (!c->accelBuilt) return fail(c, "ptc_build_accel has not been called");
    if (rp->batch_size == 0 || rp->width == 0 || rp->height == 0) return fail(c, "bad render params");
    if ((rp->flags & PTC_FLAG_SAMPLER_PMJ) && c->pmjTable.empty()) return fail(c, "PTC_FLAG_SAMPLER_PMJ needs ptc_set_sampler_tables");
    const uint32_t W = rp->width, H = rp->height;
    const uint32_t batches = rp->samples / rp->batch_size; /* VulkanRendererPathTracing.cpp:798-799 (T7) */
    const uint32_t totalSamples = batches * rp->batch_size;
    const uint32_t tile = rp->tile_size ? rp->tile_size : 32;
    const uint32_t world = rp->world ? rp->world : 1;
    c->progress = 0.0f;
    auto t0 = std::chrono::steady_clock::now();
    const size_t npx = (size_t)W * H;
    /* alpha = 1 is written by ONE rank of a partitioned render (include/ptc.h: the parts are summed) */
    const float alpha = (rp->split_mode == PTC_SPLIT_NONE || world <= 1 || rp->rank == 0) ? 1.0f : 0.0f;
    for (float *buf : {radiance, albedo, normal})
        if (buf)
            for (size_t i = 0; i < npx; i++) {
                buf[i * 4 + 0] = buf[i * 4 + 1] = buf[i * 4 + 2] = 0.0f;
                buf[i * 4 + 3] = alpha;
            }
    Stats total;
    std::atomic<uint32_t> rowsDone{0};
#pragma omp parallel
    {
        Tracer T;
        T.c = c;
        T.rp = rp;
        T.depth = rp->depth;
        T.totalLights = (uint32_t)c->lightInstances.size();
        T.envLight = (rp->flags & PTC_FLAG_ENV_IMPORTANCE) != 0u && c->cubeN > 0 && c->envTables.valid &&
                     (rp->scene.background[3] == 1.0f || rp->scene.background[3] == 2.0f);
        if (T.envLight) T.totalLights += 1u;
        T.zfar = rp->scene.volumes[2];
#pragma omp for schedule(dynamic, 1)
        for (int y = 0; y < (int)H; y++) {
            for (uint32_t x = 0; x < W; x++) {
                if (rp->split_mode == PTC_SPLIT_TILE) {
                    if (((uint32_t)y / tile + x / tile) % world != rp->rank) continue; /* diagonal stripes of tiles, like the product */
                }
                vec3 sumR(0.0f), sumA(0.0f), sumN(0.0f);
                for (uint32_t b = 0; b < batches; b++) {
                    if (rp->split_mode == PTC_SPLIT_SAMPLE && (b % world) != rp->rank) continue;
                    vec3 cumR(0.0f), cumA(0.0f), cumN(0.0f);
                    for (uint32_t s = 0; s < rp->batch_size; s++) {
                        vec3 r, a, n;
                        T.samplePixel(x, (uint32_t)y, b * rp->batch_size + s, r, a, n);
                        cumR += r / (float)totalSamples; /* raygen.rgen.glsl:126-128 */
                        cumA += a / (float)totalSamples;
                        cumN += n / (float)totalSamples;
                    }
                    sumR += cumR; /* raygen.rgen.glsl:134-144 */
                    sumA += cumA;
                    sumN += cumN;
                }
                size_t o = ((size_t)y * W + x) * 4;
                if (radiance) { radiance[o] = sumR.x; radiance[o + 1] = sumR.y; radiance[o + 2] = sumR.z; }
                if (albedo) { albedo[o] = sumA.x; albedo[o + 1] = sumA.y; albedo[o + 2] = sumA.z; }
                if (normal) { normal[o] = sumN.x; normal[o + 1] = sumN.y; normal[o + 2] = sumN.z; }
            }
            c->progress = (float)(++rowsDone) / (float)H;
        }
#pragma omp critical
        total.add(T.st);
    }
    c->stats.segments = total.segments;
    c->stats.path_rays = total.path_rays;
    c->stats.shadow_rays = total.shadow_rays;
    c->stats.shadow_hops = total.shadow_hops;
    c->stats.probe_rays = total.probe_rays;
    c->stats.probe_hops = total.probe_hops;
    c->stats.render_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->progress = 1.0f;
    return 0;
}

PTC_API int ptc_render_device(ptc_ctx *c, const ptc_render_params *, void *, void *, void *) {
    return fail(c, "the CPU oracle has no device buffers");
}
PTC_API float ptc_progress(const ptc_ctx *c) { return c ? c->progress.load() : 0.0f; }
PTC_API int ptc_get_stats(ptc_ctx *c, ptc_stats *out) {
    if (!c || !out) return 1;
    *out = c->stats;
    return 0;
}

PTC_API int ptc_trace_closest(ptc_ctx *c, const float *rays, int n, int *inst, int *prim, float *t, float *u, float *v) {
    if (!c || !rays) return fail(c, "null argument");
    if (!c->accelBuilt) return fail(c, "ptc_build_accel has not been called");
    /* brute force on purpose: this is the ground truth the device traversal is compared with */
    const bool brute = c->tris.size() <= 200000;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        const float *r = rays + (size_t)i * 8;
        vec3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
        Hit h = brute ? bruteClosest(c->tris, o, d, r[3], r[7]) : c->bvh.closest(o, d, r[3], r[7]);
        if (h.tri >= 0) {
            if (inst) inst[i] = (int)c->tris[h.tri].inst;
            if (prim) prim[i] = (int)c->tris[h.tri].prim;
            if (t) t[i] = h.t;
            if (u) u[i] = h.u;
            if (v) v[i] = h.v;
        } else {
            if (inst) inst[i] = -1;
            if (prim) prim[i] = -1;
            if (t) t[i] = r[7];
            if (u) u[i] = 0;
            if (v) v[i] = 0;
        }
    }
    return 0;
}

PTC_API int ptc_set_build_options(ptc_ctx *c, uint32_t hierarchy, uint32_t ploc_radius) {
    if (!c) return 1;
    if (hierarchy > 1 || ploc_radius > 32) return fail(c, "bad build options");
    c->hierarchy = hierarchy;
    c->plocRadius = ploc_radius ? ploc_radius : 16u;
    c->lbvh = LBVH();
    return 0;
}

PTC_API int ptc_get_lbvh(ptc_ctx *c, uint64_t *n_out, uint64_t *morton, uint32_t *order, int32_t *parent, int32_t *left,
                         int32_t *right, float *aabb) {
    if (!c) return 1;
    if (c->lbvh.n != c->tris.size() || c->lbvh.morton.empty()) {
        c->lbvh.hierarchy = c->hierarchy;
        c->lbvh.plocRadius = c->plocRadius;
        c->lbvh.build(c->tris);
    }
    const LBVH &L = c->lbvh;
    if (n_out) *n_out = L.n;
    if (L.n == 0) return 0;
    if (morton) std::copy(L.morton.begin(), L.morton.end(), morton);
    if (order) std::copy(L.order.begin(), L.order.end(), order);
    size_t nn = 2 * L.n - 1;
    if (parent) std::copy(L.parent.begin(), L.parent.end(), parent);
    if (left) std::copy(L.left.begin(), L.left.end(), left);
    if (right) std::copy(L.right.begin(), L.right.end(), right);
    if (aabb)
        for (size_t i = 0; i < nn; i++) {
            aabb[i * 6 + 0] = L.box[i].lo.x; aabb[i * 6 + 1] = L.box[i].lo.y; aabb[i * 6 + 2] = L.box[i].lo.z;
            aabb[i * 6 + 3] = L.box[i].hi.x; aabb[i * 6 + 4] = L.box[i].hi.y; aabb[i * 6 + 5] = L.box[i].hi.z;
        }
    return 0;
}

PTC_API int ptc_set_accel_mode(ptc_ctx *c, uint32_t mode) {
    if (!c) return 1;
    if (mode > PTC_ACCEL_TWO_LEVEL) return fail(c, "unknown acceleration-structure mode");
    return 0; /* the oracle's own queries are structure independent (brute force / its median-split tree over world triangles) */
}

/* CPU restatement of the two-level build (vviewer_b200/csrc/lbvh.cuh::TwoLevel; the reference's driver builds one BLAS per mesh and a
 * TLAS over the instances, VulkanScene.cpp:306-381): level m >= 0 = the tree over mesh m's OBJECT-space triangles (v0, v1 - v0, v2 - v0),
 * empty when no instance uses the mesh; level -1 = the tree over the instances' world boxes, each the bounds of the 8 corners of its
 * mesh's box through the instance matrix (same fixed operation order as the flatten). */
PTC_API int ptc_get_accel_level(ptc_ctx *c, int32_t level, uint64_t *n_nodes_out, uint64_t *n_prims_out, uint32_t *node_words, uint32_t *prim_order,
                                float *box6) {
    if (!c) return 1;
    if (level < -1 || level >= (int32_t)c->meshes.size()) return fail(c, "no such level");
    std::vector<uint8_t> used(c->meshes.size(), 0);
    for (const InstanceX &I : c->instances) used[I.d.mesh_index] = 1;
    auto v3 = [](const float *p) { return vec3(p[0], p[1], p[2]); };
    auto meshBounds = [&](uint32_t m, std::vector<AABB> &tb) {
        const ptc_mesh &M = c->meshes[m];
        tb.clear();
        if (!used[m]) return;
        tb.resize(M.tri_count);
        for (uint32_t p = 0; p < M.tri_count; p++) {
            const uint32_t *ind = &c->indices[M.first_index + 3 * (size_t)p];
            vec3 a = v3(c->vertices[M.first_vertex + ind[0]].position), b = v3(c->vertices[M.first_vertex + ind[1]].position),
                 cc = v3(c->vertices[M.first_vertex + ind[2]].position);
            tb[p] = triBounds(WorldTri{a, b - a, cc - a, 0u, p});
        }
    };
    std::vector<AABB> boxes;
    if (level >= 0) {
        meshBounds((uint32_t)level, boxes);
    } else {
        std::vector<AABB> meshBox(c->meshes.size());
        std::vector<AABB> tb;
        for (uint32_t m = 0; m < c->meshes.size(); m++) {
            meshBounds(m, tb);
            AABB b;
            for (const AABB &t : tb) b.grow(t);
            meshBox[m] = b;
        }
        boxes.resize(c->instances.size());
        for (size_t i = 0; i < c->instances.size(); i++) {
            const InstanceX &I = c->instances[i];
            const AABB &mb = meshBox[I.d.mesh_index];
            AABB w;
            for (int corner = 0; corner < 8; corner++)
                w.grow(xform_point(I.model, vec3((corner & 1) ? mb.hi.x : mb.lo.x, (corner & 2) ? mb.hi.y : mb.lo.y, (corner & 4) ? mb.hi.z : mb.lo.z)));
            boxes[i] = w;
        }
    }
    LBVH L;
    L.hierarchy = c->hierarchy;
    L.plocRadius = c->plocRadius;
    L.buildBounds(boxes);
    WideBVH W;
    W.build(L);
    if (n_nodes_out) *n_nodes_out = L.n ? W.nNodes : 0;
    if (n_prims_out) *n_prims_out = L.n;
    if (box6 && L.n) {
        box6[0] = L.scene.lo.x; box6[1] = L.scene.lo.y; box6[2] = L.scene.lo.z;
        box6[3] = L.scene.hi.x; box6[4] = L.scene.hi.y; box6[5] = L.scene.hi.z;
    }
    if (node_words && L.n) memcpy(node_words, W.words.data(), W.words.size() * 4);
    if (prim_order && L.n) memcpy(prim_order, W.triOrder.data(), W.triOrder.size() * 4);
    return 0;
}

PTC_API int ptc_get_wide_bvh(ptc_ctx *c, uint64_t *n_nodes_out, uint64_t *n_tris_out, uint32_t *node_words, uint32_t *tri_order) {
    if (!c) return 1;
    if (c->lbvh.n != c->tris.size() || c->lbvh.morton.empty()) {
        c->lbvh.hierarchy = c->hierarchy;
        c->lbvh.plocRadius = c->plocRadius;
        c->lbvh.build(c->tris);
    }
    WideBVH W;
    W.build(c->lbvh);
    if (n_nodes_out) *n_nodes_out = W.nNodes;
    if (n_tris_out) *n_tris_out = c->lbvh.n;
    if (node_words) std::copy(W.words.begin(), W.words.end(), node_words);
    if (tri_order) std::copy(W.triOrder.begin(), W.triOrder.end(), tri_order);
    return 0;
}

PTC_API int ptc_bsdf_eval(ptc_ctx *, int n, const float *params, const float *wi, const float *wo, float *out_f, float *out_pdf) {
    for (int i = 0; i < n; i++) {
        PBRStandard pbr{vec3(params[i * 5], params[i * 5 + 1], params[i * 5 + 2]), params[i * 5 + 3], params[i * 5 + 4]};
        vec3 a(wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]), b(wo[i * 3], wo[i * 3 + 1], wo[i * 3 + 2]);
        vec3 F = evalPBRStandard(pbr, a, b, normalize(b + a));
        float p = pdfPBRStandard(a, b, pbr);
        out_f[i * 3] = F.x; out_f[i * 3 + 1] = F.y; out_f[i * 3 + 2] = F.z;
        out_pdf[i] = p;
    }
    return 0;
}
PTC_API int ptc_bsdf_sample(ptc_ctx *, int n, const float *params, const float *wo, const float *u, float *out_wi, float *out_f,
                            float *out_pdf) {
    for (int i = 0; i < n; i++) {
        PBRStandard pbr{vec3(params[i * 5], params[i * 5 + 1], params[i * 5 + 2]), params[i * 5 + 3], params[i * 5 + 4]};
        vec3 b(wo[i * 3], wo[i * 3 + 1], wo[i * 3 + 2]);
        vec3 w;
        float p;
        vec3 F = samplePBRStandard(w, b, p, pbr, vec2{u[i * 3], u[i * 3 + 1]}, u[i * 3 + 2]);
        out_wi[i * 3] = w.x; out_wi[i * 3 + 1] = w.y; out_wi[i * 3 + 2] = w.z;
        out_f[i * 3] = F.x; out_f[i * 3 + 1] = F.y; out_f[i * 3 + 2] = F.z;
        out_pdf[i] = p;
    }
    return 0;
}
PTC_API int ptc_set_sampler_tables(ptc_ctx *c, const float *pmj, uint32_t n_sequences, uint32_t n_samples, const float *blue, uint32_t n_textures,
                                   uint32_t resolution) {
    if (!c || !pmj || !blue) return fail(c, "null argument");
    if (n_sequences != PMJ_N_SEQUENCES || n_samples != PMJ_N_SAMPLES || n_textures != BLUE_NOISE_TEXTURES || resolution != BLUE_NOISE_RESOLUTION)
        return fail(c, "sampler tables must be 16 x 16384 x 2 and 48 x 128 x 128 (rng_pmj_defines.glsl, bluenoise_defines.glsl)");
    c->pmjTable.assign(pmj, pmj + (size_t)n_sequences * n_samples * 2);
    c->blueTable.assign(blue, blue + (size_t)n_textures * resolution * resolution);
    return 0;
}

PTC_API int ptc_sampler_points(ptc_ctx *c, uint32_t px, uint32_t py, uint32_t width, uint32_t first_index, uint32_t count, uint32_t dimension,
                               uint32_t flags, float *out_xy) {
    if ((flags & PTC_FLAG_SAMPLER_PMJ) && (!c || c->pmjTable.empty())) return fail(c, "ptc_set_sampler_tables has not been called");
    for (uint32_t i = 0; i < count; i++) {
        Rng r;
        if (flags & PTC_FLAG_SAMPLER_PMJ) {
            r.initPmj(px, py, width, first_index + i, first_index + count, PmjTables{c->pmjTable.data(), c->blueTable.data()});
            r.state += dimension;
            vec2 p;
            if (flags & PTC_SAMPLER_HOOK_1D) {
                p.x = r.rand1D();
                p.y = r.rand1D();
            } else {
                p = r.rand2D();
            }
            out_xy[2 * i] = p.x;
            out_xy[2 * i + 1] = p.y;
            continue;
        }
        r.init(px, py, width, first_index + i, (flags & PTC_FLAG_SAMPLER_SOBOL) != 0u);
        if (r.ld)
            r.state = dimension;
        else
            for (uint32_t k = 0; k < dimension; k++) r.rand1D();
        vec2 p;
        if (flags & PTC_SAMPLER_HOOK_1D) {
            p.x = r.rand1D();
            p.y = r.rand1D();
        } else {
            p = r.rand2D();
        }
        out_xy[2 * i] = p.x;
        out_xy[2 * i + 1] = p.y;
    }
    return 0;
}

PTC_API int ptc_srgb_table(ptc_ctx *, float *out256) {
    if (!out256) return 1;
    for (int i = 0; i < 256; i++) out256[i] = kSrgbToLinear[i];
    return 0;
}

PTC_API int ptc_env_sample(ptc_ctx *c, int n, const float *u01, float *out_dirs, float *out_pdf) {
    if (!c) return 1;
    if (!c->envTables.valid) return fail(c, "no environment");
    for (int i = 0; i < n; i++) {
        float u, v, puv;
        envdist::sampleUV(c->envTables, u01[2 * i], u01[2 * i + 1], u, v, puv);
        envdist::direction(u, v, out_dirs[3 * i], out_dirs[3 * i + 1], out_dirs[3 * i + 2]);
        out_pdf[i] = envdist::solidAnglePdf(puv, v);
    }
    return 0;
}
PTC_API int ptc_env_pdf(ptc_ctx *c, int n, const float *dirs, float *out_pdf) {
    if (!c) return 1;
    if (!c->envTables.valid) return fail(c, "no environment");
    for (int i = 0; i < n; i++) {
        vec2 uv = sampleEquirectangularMap(normalize(vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2])));
        out_pdf[i] = envdist::solidAnglePdf(envdist::pdfUV(c->envTables, uv.x, uv.y), uv.y);
    }
    return 0;
}

PTC_API int ptc_env_lookup(ptc_ctx *c, int n, const float *dirs, float *out_rgb) {
    if (!c) return 1;
    for (int i = 0; i < n; i++) {
        vec3 v = sampleCubemap(c, vec3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]));
        out_rgb[i * 3] = v.x; out_rgb[i * 3 + 1] = v.y; out_rgb[i * 3 + 2] = v.z;
    }
    return 0;
}

} /* extern "C" */
