/*
 * Wavefront path-tracing pipeline: raygen -> [extend -> shade -> shadow -> probe] x depth -> accumulate.
 * Replaces the Vulkan ray-tracing pipeline of the reference (paths under src/lib/vengine/shaders/pt/):
 *   k_raygen      raygen.rgen.glsl:24-98        camera ray, thin lens, path state init
 *   k_extend      traceRayEXT at raygen.rgen.glsl:110 (closest hit, sbt offset 0)
 *   k_shade       rayPrimaryLambert.rchit / rayPrimaryPBRStandard.rchit / rayPrimary.rmiss incl.
 *                 process_volume_hit.glsl, lightSampling.glsl:1-106 (sampling part), russian_roulette.glsl
 *   k_shadow      lightSampling.glsl:108-144 + raySecondary.rahit/.rchit/.rmiss (shadow chain)
 *   k_probe       next_event_estimation.glsl + rayNEE.rahit/.rchit/.rmiss (BSDF-sample emissive probe)
 *   k_accumulate  raygen.rgen.glsl:126-144
 * Path state is SoA in HBM, one slot per (sample, pixel) of the batch; queues hold slot ids and are compacted
 * with ballot/popc warp-aggregated atomics.  Kernels are persistent grid-stride loops that read their work
 * count from device memory, so a whole batch runs without a single host synchronisation.
 */
#pragma once
#include "common.cuh"
#include "bsdf.cuh"
#include "traverse.cuh"
#include "envdist.cuh"

namespace wf {

/* flags word of a path */
#define PF_DEPTH_MASK 0xffu
#define PF_SURFACE 0x100u /* surfaceDepth > 0 */
#define PF_INVOL 0x200u   /* insideVolume */
#define PF_PROBE_DEFERRED 0x400u /* the BSDF probe of the previous event is answered by this segment's closest hit (RenderConst::fuseProbe) */
#define PF_VOL_SHIFT 16

enum { CNT_ACTIVE = 0, CNT_SHADOW = 1, CNT_PROBE = 2, CNT_FETCH = 3, CNT_FETCH_SHADOW = 4, CNT_FETCH_PROBE = 5, CNT_STRIDE = 8 };
enum { ST_SEGMENTS = 0, ST_SHADOW_RAYS, ST_SHADOW_HOPS, ST_PROBE_RAYS, ST_PROBE_HOPS, ST_NODE_VISITS, ST_TRI_TESTS, ST_NODE_ITERS, ST_TRI_ITERS, ST_COUNT };

struct Wave {
    float4 *orgRng;    /* origin.xyz, rng state */
    float4 *dirFlags;  /* direction.xyz, flags */
    float4 *beta;      /* throughput.xyz */
    float4 *radiance;  /* rgb */
    float4 *hit;       /* t, u, v, triangle position (int bits, -1 = miss) */
    uint32_t *hitInst; /* two-level structure: the instance of the hit (the triangle position alone names a mesh triangle) */
    float4 *aovAlbedo; /* first-hit albedo */
    float4 *aovNormal; /* first-hit normal * 0.5 + 0.5 */
    float4 *shOrgTmax; /* shadow request: origin.xyz, tmax */
    float4 *shDirVol;  /* direction.xyz, volume state (flags bits) */
    float4 *shContrib; /* unshadowed contribution rgb */
    float4 *prBetaPdf; /* probe request: beta before roulette .xyz, sampling pdf */
    uint32_t *queue[2];
    uint32_t *qShadow, *qProbe;
    uint32_t *counters; /* [depth + 1][CNT_STRIDE] */
    unsigned long long *stats;
};

struct RenderConst {
    ptc_scene_data sd;
    uint32_t width, height, depth, totalSamples;
    uint32_t cameraType;
    float orthoW, orthoH;
    uint32_t totalLights;    /* light instances, + 1 when the environment is light-sampled */
    uint32_t fuseProbe;      /* opaque, media-free scene: the probe ray of a path that goes on is its next path ray (see k_shade) */
    uint32_t envLight;       /* PTC_FLAG_ENV_IMPORTANCE and an HDRI environment: light index totalLights - 1 is the environment */
    uint32_t flags;
    uint32_t nPixLocal;      /* pixels rendered by this rank */
    const uint32_t *pixmap;  /* local pixel -> global pixel, or nullptr = identity */
};

PTC_D uint32_t laneId() { return threadIdx.x & 31u; }

/* path state is written once and read once per kernel, and a batch's state (6 GB at 1080p x 16) is far larger than L2:
 * stream it (evict-first) so that it does not displace the BVH, the materials and the textures */
PTC_D float4 ldS(const float4 *p) { return __ldcs(p); }
PTC_D void stS(float4 *p, float4 v) { __stcs(p, v); }

/* ballot/popc compaction: every lane of the warp must call this */
PTC_D void queuePush(uint32_t *__restrict__ queue, uint32_t *__restrict__ counter, bool pred, uint32_t value) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)laneId() == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(m & ((1u << laneId()) - 1u))] = value;
}
PTC_D void statAdd(unsigned long long *stat, uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (laneId() == 0 && v) atomicAdd(stat, (unsigned long long)v);
}

PTC_D float3 mulPoint(const float *m /*3x4 row-major*/, float3 p) {
    return f3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
PTC_D float3 mulNormal(const float *a /*3x3 row-major inverse*/, float3 n) { /* a^T * n */
    return f3(a[0] * n.x + a[3] * n.y + a[6] * n.z, a[1] * n.x + a[4] * n.y + a[7] * n.z, a[2] * n.x + a[5] * n.y + a[8] * n.z);
}
PTC_D float3 ld3(const float *p) { return f3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

/* Bindless material maps.  A warp's 32 hits touch many different textures, and a texture instruction needs a warp-uniform
 * handle (per-lane handles are serialised by a compiler-generated loop over the unique handles: 12 % of k_shade's
 * instructions at 6 of 32 lanes, profiles/r1_v2).  So textures of one (width, height, sRGB) class share ONE layered texture
 * object and the lane picks its layer; all-white textures (the engine's defaults) are not fetched at all. */
PTC_D float4 texFetch(const DScene &sc, uint32_t idx, float u, float v) {
    if (idx >= sc.nTextures) return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    const uint32_t r = __ldg(&sc.texRef[idx]);
    if (r == TEX_WHITE) return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    return tex2DLayered<float4>(__ldg(&sc.texClasses[r >> 16]), u, v, (int)(r & 0xffffu));
}
PTC_D float3 envFetch(const DScene &sc, float3 d) {
    if (!sc.hasCubemap) return f3(0.0f);
    float4 c = texCubemap<float4>(sc.cubemap, d.x, d.y, d.z);
    return f3(c);
}

/* ------------------------------------------------------------------ geometry of a hit (process_hit.glsl + construct_frame.glsl) */
struct Surf {
    const DInstance *inst;
    const ptc_material *mat;
    float3 p0, p1, p2; /* object-space positions */
    float2 uv;
    float3 pos, n, t; /* world position, unnormalised->normalised world normal / tangent */
};
/* hitInst: the instance the hit reported (two-level structure); single level: the record names its instance itself */
PTC_D void loadSurf(const DScene &sc, int32_t triPos, uint32_t hitInst, float u, float v, bool wantFrame, Surf &s) {
    const float4 *__restrict__ r = sc.shading + 9 * (size_t)triPos;
    /* (plain cached loads: streaming the records past L1 with ld.global.cs was measured 1 % slower, 2054 vs 2074 Mseg/s) */
    const float4 r0 = __ldg(r + 0), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4), r5 = __ldg(r + 5), r6 = __ldg(r + 6);
    const DInstance *I = &sc.instances[sc.twoLevel ? hitInst : __float_as_uint(r6.w)];
    s.inst = I;
    s.mat = &sc.materials[I->material];
    const float w0 = 1.0f - u - v, w1 = u, w2 = v;
    s.p0 = f3(r0);
    s.p1 = f3(r1);
    s.p2 = f3(r2);
    s.uv = make_float2(r0.w * w0 + r2.w * w1 + r4.w * w2, r1.w * w0 + r3.w * w1 + r5.w * w2);
    float3 lp = s.p0 * w0 + s.p1 * w1 + s.p2 * w2;
    s.pos = mulPoint(I->m, lp);
    float3 ln = f3(r3) * w0 + f3(r4) * w1 + f3(r5) * w2;
    s.n = normalize(mulNormal(I->nrm, ln));
    if (wantFrame) {
        const float4 r7 = __ldg(r + 7), r8 = __ldg(r + 8);
        float3 lt = f3(r6) * w0 + f3(r7) * w1 + f3(r8) * w2;
        s.t = normalize(mulNormal(I->nrm, lt));
    }
}

/* material of the triangle at a traversal position without touching the rest of its record: the any-hit shaders of the shadow and
 * probe rays decide most candidates from the material alone (opaque and not emissive: the ray ends there) */
PTC_D const ptc_material *hitMaterial(const DScene &sc, int32_t triPos, uint32_t hitInst) {
    if (sc.twoLevel) return &sc.materials[sc.instances[hitInst].material];
    const float4 r6 = __ldg(sc.shading + 9 * (size_t)triPos + 6);
    return &sc.materials[sc.instances[__float_as_uint(r6.w)].material];
}

/* ------------------------------------------------------------------ volumes */
struct Medium {
    float3 sigma_s, sigma_t;
    float g;
};
PTC_D Medium loadMedium(const DScene &sc, uint32_t idx) { /* process_volume_hit.glsl:3-8 */
    const ptc_material *m = &sc.materials[idx];
    float3 sa = ld3(m->albedo);
    float3 ss = fmax3(ld3(m->metallic_roughness_ao), f3(PT_EPSILON));
    Medium md;
    md.sigma_s = ss;
    md.sigma_t = sa + ss;
    md.g = __ldg(&m->emissive[0]);
    return md;
}
PTC_D float3 transmittance(const DScene &sc, uint32_t volIdx, float vtstart, float vtend) { /* process_volume_transmittance.glsl */
    Medium md = loadMedium(sc, volIdx);
    float dist = fmaxf(vtend - vtstart, PT_EPSILON);
    return exp3(-(md.sigma_t * dist));
}
PTC_D void volumeChange(const DInstance *I, bool flipped, uint32_t &flags) { /* rchit :74-90 */
    flags &= ~(PF_INVOL | 0xffff0000u);
    float nv = flipped ? I->volFront : I->volBack;
    if (nv != -1.0f) flags |= PF_INVOL | ((uint32_t)nv << PF_VOL_SHIFT);
}

/* ------------------------------------------------------------------ light sampling (lightSampling.glsl:1-106) */
struct LightSample {
    float3 dir, radiance;
    float pdf, tmax;
    bool delta;
};
template <bool ENV>
PTC_D LightSample sampleLight(const DScene &sc, const RenderConst &rc, Rng &rng, float3 origin) {
    LightSample ls;
    ls.delta = true;
    ls.radiance = f3(0.0f);
    ls.pdf = 1.0f;
    ls.dir = f3(0.0f, 1.0f, 0.0f);
    ls.tmax = 10000.0f;
    const uint32_t totalLights = rc.totalLights;
    if (totalLights == 0u) return ls;
    const float pick = 1.0f / (float)totalLights;
    uint32_t li = min((uint32_t)(rnd(rng) * (float)totalLights), totalLights - 1u);
    if (ENV && li == totalLights - 1u) {
        /* extension (trap T3, include/ptc.h PTC_FLAG_ENV_IMPORTANCE): the environment as the last light */
        const float2 u01 = rnd2(rng);
        const envd::DeviceTables et{sc.envCdfV, sc.envCdfU};
        float eu, ev, puv;
        envd::sampleUV(et, u01.x, u01.y, eu, ev, puv);
        ls.dir = envd::direction(eu, ev);
        ls.radiance = envFetch(sc, ls.dir) * (rc.sd.background[3] == 1.0f ? rc.sd.exposure[1] : 1.0f);
        ls.pdf = pick * envd::solidAnglePdf(puv, ev);
        ls.delta = false;
        if (!(ls.pdf > 0.0f)) ls.radiance = f3(0.0f);
        return ls;
    }
    const ptc_light_instance *L = &sc.lightInstances[li];
    const uint32_t type = __ldg(&L->info[3]);
    if (type == 0u) { /* point */
        const ptc_light_data *ld = &sc.lightData[__ldg(&L->info[0])];
        float3 lp = ld3(L->position);
        float3 d = lp - origin;
        ls.tmax = length(d);
        ls.dir = d / ls.tmax;
        float dist = length(origin - lp);
        ls.radiance = ld3(ld->color) * (1.0f / (dist * dist)) * __ldg(&ld->color[3]);
        ls.pdf = pick;
    } else if (type == 1u) { /* directional; direction left unnormalised (trap T4) */
        const ptc_light_data *ld = &sc.lightData[__ldg(&L->info[0])];
        ls.dir = -ld3(L->position);
        ls.tmax = rc.sd.volumes[2];
        ls.radiance = ld3(ld->color) * __ldg(&ld->color[3]);
        ls.pdf = pick;
    } else if (type == 2u) { /* mesh light */
        const DInstance *I = &sc.instances[__ldg(&L->info[1])];
        const ptc_material *mat = &sc.materials[I->material];
        const uint32_t nt = I->numTriangles;
        uint32_t tri = min((uint32_t)(rnd(rng) * (float)nt), nt - 1u);
        const float2 u01 = rnd2(rng);
        float2 bc = sampleTriangle(u01.x, u01.y);
        float3 sb = f3(bc.x, bc.y, 1.0f - bc.x - bc.y);
        const uint32_t *ind = sc.indices + I->firstIndex + 3 * (size_t)tri;
        const ptc_vertex *V0 = &sc.vertices[I->firstVertex + __ldg(ind + 0)];
        const ptc_vertex *V1 = &sc.vertices[I->firstVertex + __ldg(ind + 1)];
        const ptc_vertex *V2 = &sc.vertices[I->firstVertex + __ldg(ind + 2)];
        float3 q0 = ld3(V0->position), q1 = ld3(V1->position), q2 = ld3(V2->position);
        float3 w0 = mulPoint(I->m, q0), w1 = mulPoint(I->m, q1), w2 = mulPoint(I->m, q2);
        float area = 0.5f * length(cross(w1 - w0, w2 - w0));
        float3 sp = mulPoint(I->m, q0 * sb.x + q1 * sb.y + q2 * sb.z);
        /* transpose(inverse(M)) * n, NOT normalised (trap T5) */
        float3 sn = mulNormal(I->nrm, ld3(V0->normal) * sb.x + ld3(V1->normal) * sb.y + ld3(V2->normal) * sb.z);
        float su = sb.x * __ldg(&V0->uv[0]) + sb.y * __ldg(&V1->uv[0]) + sb.z * __ldg(&V2->uv[0]);
        float sv = sb.x * __ldg(&V0->uv[1]) + sb.y * __ldg(&V1->uv[1]) + sb.z * __ldg(&V2->uv[1]);
        float3 d = sp - origin;
        ls.tmax = length(d);
        ls.dir = d / ls.tmax;
        float dp = dot(-ls.dir, sn);
        if (dp > 0.0f) {
            float4 et = texFetch(sc, __ldg(&mat->tex2[0]), su * __ldg(&mat->uv_tiling[0]), sv * __ldg(&mat->uv_tiling[1]));
            ls.radiance = ld3(mat->emissive) * __ldg(&mat->emissive[3]) * f3(et);
            float dd = length(origin - sp);
            ls.pdf = pick * (1.0f / (float)nt) * (1.0f / area) * (dd * dd) / dp;
        } else {
            ls.radiance = f3(0.0f);
            ls.pdf = 0.0f;
        }
        ls.delta = false;
    }
    return ls;
}

/* Does the line origin + t dir, t >= 0, meet the (padded) world box of any emissive instance?  The probe chain of
 * next_event_estimation.glsl only ever adds radiance when it ends on an emissive triangle, and all its hops stay on this line, so a
 * ray that misses every box is not traced (result-identical). */
PTC_D bool rayMeetsEmitters(const DScene &sc, float3 o, float3 d) {
    const float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); /* +-inf for axis-parallel rays: the slab test handles it, NaN = inside */
    for (uint32_t k = 0; k < sc.nEmissiveBoxes; k++) {
        const float4 lo = __ldg(sc.emissiveBoxes + 2 * k), hi = __ldg(sc.emissiveBoxes + 2 * k + 1);
        const float tx0 = (lo.x - o.x) * inv.x, tx1 = (hi.x - o.x) * inv.x;
        const float ty0 = (lo.y - o.y) * inv.y, ty1 = (hi.y - o.y) * inv.y;
        const float tz0 = (lo.z - o.z) * inv.z, tz1 = (hi.z - o.z) * inv.z;
        /* fminf / fmaxf drop NaNs (0 * inf when the origin lies on a slab plane of a parallel ray), which keeps the test conservative */
        const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
        const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
        if (tn <= tf * 1.00001f + 1e-6f) return true;
    }
    return false;
}

#ifdef PTC_NAN_TRAP
#define NAN_TRAP(cond, ...) do { if (cond) printf(__VA_ARGS__); } while (0)
#else
#define NAN_TRAP(cond, ...) do { } while (0)
#endif
#define BAD3(v) (!(isfinite((v).x) && isfinite((v).y) && isfinite((v).z)))

/* requests produced by one shading event */
struct Requests {
    bool shadow, probe;
    float3 shOrigin, shDir, shContrib;
    float shTmax;
    float3 prBeta;
    float prPdf;
};

/* ------------------------------------------------------------------ k_raygen */
__global__ void __launch_bounds__(256) k_raygen(Wave w, const __grid_constant__ RenderConst rc, uint32_t nSlots, uint32_t firstSample) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot == 0) w.counters[CNT_ACTIVE] = nSlots;
    if (slot >= nSlots) return;
    const uint32_t p = slot % rc.nPixLocal, s = slot / rc.nPixLocal;
    const uint32_t pixel = rc.pixmap ? rc.pixmap[p] : p;
    const uint32_t px = pixel % rc.width, py = pixel / rc.width;
    Rng rng = rngInit(px, py, rc.width, firstSample + s, samplerOfFlags(rc.flags));
    const float2 u01 = rnd2(rng);
    const float u0 = u01.x, u1 = u01.y;
    const float dx = (((float)px + u0) / (float)rc.width) * 2.0f - 1.0f;
    const float dy = (((float)py + u1) / (float)rc.height) * 2.0f - 1.0f;
    float3 oc = f3(0.0f), dc;
    if (rc.cameraType == PTC_CAMERA_ORTHOGRAPHIC) { /* documented deviation T10 */
        oc = f3(dx * 0.5f * rc.orthoW, -dy * 0.5f * rc.orthoH, 0.0f);
        dc = f3(0.0f, 0.0f, -1.0f);
    } else { /* projectionInverse * (d.x, d.y, 1, 1), no w divide (raygen.rgen.glsl:67-71) */
        const float *pi = rc.sd.projection_inverse;
        float3 target = f3(pi[0] * dx + pi[4] * dy + pi[8] + pi[12], pi[1] * dx + pi[5] * dy + pi[9] + pi[13], pi[2] * dx + pi[6] * dy + pi[10] + pi[14]);
        dc = normalize(target);
    }
    const float lensRadius = rc.sd.exposure[2];
    if (lensRadius > 0.0f) { /* raygen.rgen.glsl:74-85 */
        const float2 l01 = rnd2(rng);
        float2 disk = concentricDisk(l01.x, l01.y);
        float ft = rc.sd.exposure[3] / (-dc.z);
        float3 focus = oc + dc * ft;
        oc = oc + f3(lensRadius * disk.x, lensRadius * disk.y, 0.0f);
        dc = normalize(focus - oc);
    }
    const float *vi = rc.sd.view_inverse;
    /* same evaluation order as the oracle's xform_point / xform_dir */
    float3 o = f3(((vi[0] * oc.x + vi[4] * oc.y) + vi[8] * oc.z) + vi[12], ((vi[1] * oc.x + vi[5] * oc.y) + vi[9] * oc.z) + vi[13],
                  ((vi[2] * oc.x + vi[6] * oc.y) + vi[10] * oc.z) + vi[14]);
    float3 d = f3((vi[0] * dc.x + vi[4] * dc.y) + vi[8] * dc.z, (vi[1] * dc.x + vi[5] * dc.y) + vi[9] * dc.z, (vi[2] * dc.x + vi[6] * dc.y) + vi[10] * dc.z);
    uint32_t flags = 0;
    if (rc.sd.volumes[0] != -1.0f) flags |= PF_INVOL | ((uint32_t)(int)rc.sd.volumes[0] << PF_VOL_SHIFT);
    stS(&w.orgRng[slot], make_float4(o.x, o.y, o.z, __uint_as_float(rng.s)));
    stS(&w.dirFlags[slot], make_float4(d.x, d.y, d.z, __uint_as_float(flags)));
    stS(&w.beta[slot], make_float4(1.0f, 1.0f, 1.0f, 0.0f));
    stS(&w.radiance[slot], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    stS(&w.aovAlbedo[slot], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    stS(&w.aovNormal[slot], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
}

/* ------------------------------------------------------------------ warp-persistent traversal loop */
/* Shared by k_extend, k_shadow and k_probe.  Every warp pulls work items from a queue (one global atomic per 64 items) and
 * keeps one traversal per lane.  All 32 lanes run the same loop body:
 *   node phase      every lane with node work visits ONE wide node (8 boxes); triangle groups the visit exposes are NOT tested
 *                   right away (that ran at 2.8 of 32 lanes, profiles/README.md) but stashed in a per-lane shared-memory ring;
 *   triangle phase  entered by warp vote once enough lanes hold stashed triangles (or lanes are blocked on them): every lane
 *                   with a stash tests one triangle per iteration;
 *   completion      a lane whose query is finished hands the hit to its Policy, which either starts the item's next query
 *                   (shadow / probe chains: the next hop, or the next-nearest candidate) or retires the item;
 *   refill          retired lanes fetch new items as soon as fewer than tune.minActive lanes are traversing.
 * Policy (one object per lane):
 *   bool begin(uint32_t index, trv::Trav &tr)   load item `index` and describe its first query in tr (o, d, tmin, tmax, t0, id0);
 *                                               false = the item retired at once
 *   bool next(trv::Trav &tr)                    the query is done (tr.best); true = tr describes another query to run
 *   bool anyHit()                                                          the running query may stop at the first accepted triangle */
#define EXTEND_MIN_ACTIVE 28
#define EXTEND_STASH 4      /* stashed triangle groups per lane */
#define EXTEND_TRI_ENTER 12 /* lanes with stashed triangles that trigger a triangle phase */
#define EXTEND_TRI_LEAVE 6  /* the phase ends when fewer lanes than this still hold triangles */
#define EXTEND_BLOCKED 8    /* lanes that have nothing but stashed triangles left */
struct ExtendTune { /* warp-vote thresholds; defaults above, overridable through PTC_EXTEND_TUNE for tuning runs */
    uint32_t minActive, triEnter, triLeave, blocked;
};
struct TraceCounters { /* -DPTC_TRAV_STATS builds only (tools/trav_stats.py) */
    uint32_t node = 0, tri = 0, nodeIt = 0, triIt = 0;
};
#ifdef PTC_TRAV_STATS
#define TRV_COUNT(x) (x)++
#else
#define TRV_COUNT(x)
#endif

/* One stashed triangle.  Policies with ALL_HITS (the shadow ray in scenes without media) look at EVERY triangle the ray crosses in
 * traversal order instead of asking for the next-nearest one again and again: Policy::candidate() returns true when the ray is
 * decided (it is then marked as hit, which ends the occlusion query). */
template <class Policy>
PTC_D void testTriangle(const DScene &sc, Policy &pol, trv::Trav &tr, int32_t pos) {
    if constexpr (Policy::ALL_HITS) {
        if (pol.allHits()) {
            const float4 *__restrict__ tris = sc.tris;
            const float4 v0 = __ldg(&tris[3 * (size_t)pos + 0]), e1 = __ldg(&tris[3 * (size_t)pos + 1]), e2 = __ldg(&tris[3 * (size_t)pos + 2]);
            float t, u, v;
            if (tr.best.pos < 0 && trv::intersectTri(v0, e1, e2, tr.o, tr.d, t, u, v) && t > tr.tmin && t < tr.tmax &&
                pol.candidate(pos, Policy::TWO_LEVEL ? tr.inst : 0xffffffffu, u, v))
                tr.best.pos = pos;
            return;
        }
    }
    tr.template triTest<Policy::TWO_LEVEL>(sc, pos);
}

template <class Policy>
PTC_D void traceLoop(const DScene &sc, Policy &pol, uint32_t count, uint32_t *fetchCounter, const ExtendTune tune, const trv::Stack &stack, uint2 *stash,
                     TraceCounters &cnt) {
    trv::WarpFeeder feeder;
    feeder.sizeFor(count);
    trv::Trav tr;
    tr.sp = -1;
    tr.ng = tr.tg = make_uint2(0u, 0u);
    tr.best.pos = -1;
    uint32_t nStash = 0;
    bool active = false;  /* the lane owns an item */
    bool pending = false; /* tr describes a query that has not been started yet */
    while (true) {
        const uint32_t i = feeder.fetch(!active, fetchCounter, count);
        if (i != 0xffffffffu) active = pending = pol.begin(i, tr);
        if (!__any_sync(0xffffffffu, active)) {
            if (!__any_sync(0xffffffffu, i != 0xffffffffu)) break; /* nothing left to fetch */
            continue;
        }
        while (true) { /* warp-convergent: no lane leaves this loop alone */
            if (pending) {
                tr.template startLevel<Policy::TWO_LEVEL>(sc);
                nStash = 0;
                pending = false;
            }
            if constexpr (Policy::TWO_LEVEL) {
                /* out of node work: the next stack entry (more nodes, the rest of a top-level leaf, or the way out of the instance) */
                if (active && tr.ng.y <= 0x00ffffffu && (!tr.top || tr.ig.y == 0u) && tr.sp > 0) tr.popNext(stack, tr.tg.y == 0u && nStash == 0u);
                /* top level: step into the next instance of the current leaf */
                if (active && tr.top && tr.ig.y != 0u && tr.ng.y <= 0x00ffffffu) {
                    const uint32_t k = 31u - (uint32_t)__clz(tr.ig.y);
                    tr.ig.y &= ~(1u << k);
                    if (tr.ig.y != 0u) tr.push(stack, make_uint2(tr.ig.x | TRV_INSTANCE_FLAG, tr.ig.y));
                    tr.push(stack, make_uint2(TRV_MARKER_X, 0u));
                    tr.enterInstance(sc, tr.ig.x + k);
                    tr.ig.y = 0u;
                }
            }
            TRV_COUNT(cnt.nodeIt);
            if (Policy::TWO_LEVEL && active && tr.top && tr.ng.y > 0x00ffffffu) {
                /* a top-level node: what it exposes are instances, visited before the node's remaining children */
                TRV_COUNT(cnt.node);
                const uint2 g = tr.nodeStep(sc, stack);
                if (g.y != 0u) {
                    if (tr.ng.y > 0x00ffffffu) tr.push(stack, tr.ng);
                    tr.ng.y = 0u;
                    tr.ig = g;
                }
            } else if (active && tr.ng.y > 0x00ffffffu) {
                TRV_COUNT(cnt.node);
                const uint2 g = tr.nodeStep(sc, stack);
                if (g.y != 0u) {
                    if (nStash == EXTEND_STASH) { /* ring full: make room by finishing the current group now (rare) */
                        while (tr.tg.y != 0u) {
                            const uint32_t k = 31u - (uint32_t)__clz(tr.tg.y);
                            tr.tg.y &= ~(1u << k);
                            testTriangle(sc, pol, tr, (int32_t)(tr.tg.x + k));
                        }
                        tr.tg = stash[(--nStash) * TRV_BLOCK];
                    }
                    if (tr.tg.y == 0u)
                        tr.tg = g;
                    else
                        stash[(nStash++) * TRV_BLOCK] = g;
                }
                if (!Policy::TWO_LEVEL && tr.ng.y <= 0x00ffffffu && tr.sp > 0) tr.ng = tr.pop(stack);
            }
            /* (two levels: a lane with stack entries or pending instances still has work, it just takes it at the top of the loop) */
            const bool hasNode = active && (tr.ng.y > 0x00ffffffu || (Policy::TWO_LEVEL && (tr.ig.y != 0u || (tr.sp > 0 && tr.tg.y == 0u))));
            bool hasTri = active && tr.tg.y != 0u;
            const unsigned mN = __ballot_sync(0xffffffffu, hasNode);
            unsigned mT = __ballot_sync(0xffffffffu, hasTri);
            if (__popc(mT) >= tune.triEnter || __popc(mT & ~mN) >= tune.blocked || (mN == 0u && mT != 0u)) {
                do {
                    TRV_COUNT(cnt.triIt);
                    if (hasTri) {
                        TRV_COUNT(cnt.tri);
                        const uint32_t k = 31u - (uint32_t)__clz(tr.tg.y);
                        tr.tg.y &= ~(1u << k);
                        testTriangle(sc, pol, tr, (int32_t)(tr.tg.x + k));
                        if (tr.tg.y == 0u && nStash > 0u) tr.tg = stash[(--nStash) * TRV_BLOCK];
                        hasTri = tr.tg.y != 0u;
                    }
                    mT = __ballot_sync(0xffffffffu, hasTri);
                } while (__popc(mT) >= tune.triLeave || (mT & ~mN) != 0u);
            }
            /* query complete: nothing left to visit, or an occlusion query that already has its answer */
            const bool moreWork = Policy::TWO_LEVEL ? (tr.ng.y > 0x00ffffffu || hasTri || tr.sp > 0 || tr.ig.y != 0u) : (hasNode || hasTri);
            if (active && (!moreWork || (pol.anyHit() && tr.best.pos >= 0))) {
                if constexpr (Policy::TWO_LEVEL) {
                    if (!tr.top) tr.leaveInstance(); /* an occlusion query may end inside an instance: the policy sees the world-space ray */
                }
                active = pending = pol.next(tr);
                tr.ng.y = 0u;
                tr.tg.y = 0u;
            }
            const unsigned mA = __ballot_sync(0xffffffffu, active);
            if (mA == 0u || (!feeder.exhausted && __popc(mA) < tune.minActive)) break;
        }
    }
}

PTC_D void traceCountersFlush(const Wave &w, const TraceCounters &cnt) {
#ifdef PTC_TRAV_STATS
    statAdd(&w.stats[ST_NODE_VISITS], cnt.node);
    statAdd(&w.stats[ST_TRI_TESTS], cnt.tri);
    if (laneId() == 0) {
        atomicAdd(&w.stats[ST_NODE_ITERS], (unsigned long long)cnt.nodeIt);
        atomicAdd(&w.stats[ST_TRI_ITERS], (unsigned long long)cnt.triIt);
    }
#endif
}

/* ------------------------------------------------------------------ k_extend */
/* closest hit of the path ray: traceRayEXT at raygen.rgen.glsl:110 (tmin 1e-3, tmax 1e4) */
/* VOLUMES: the scene has media.  A path inside a medium already knows how far it will fly (k_shade draws the free-flight distance of
 * the NEXT segment from a copy of the random stream and leaves 1e-3 + distance in the slot's hit record): the ray only needs to be traced
 * that far - a surface beyond the scattering point changes nothing (process_volume_hit.glsl:13-25). */
template <bool VOLUMES, bool TL>
struct ExtendPolicy {
    static constexpr bool ALL_HITS = false;
    static constexpr bool TWO_LEVEL = TL;
    const Wave &w;
    const uint32_t *__restrict__ q;
    uint32_t bounce, slot;
    PTC_D ExtendPolicy(const Wave &w_, uint32_t bounce_) : w(w_), q(w_.queue[bounce_ & 1u]), bounce(bounce_), slot(0) {}
    PTC_D bool anyHit() const { return false; }
    PTC_D bool begin(uint32_t i, trv::Trav &tr) {
        slot = bounce == 0u ? i : q[i];
        const float4 o = ldS(&w.orgRng[slot]), d = ldS(&w.dirFlags[slot]);
        tr.o = f3(o);
        tr.d = f3(d);
        tr.tmin = tr.t0 = 0.001f;
        tr.tmax = 10000.0f;
        if (VOLUMES && bounce > 0u) tr.tmax = ldS(&w.hit[slot]).x;
        tr.id0 = 0xffffffffu;
        return true;
    }
    PTC_D bool next(trv::Trav &tr) {
        const trv::HitRec &h = tr.best;
        stS(&w.hit[slot], make_float4(h.t, h.u, h.v, __int_as_float(h.pos)));
        if constexpr (TL) w.hitInst[slot] = h.inst;
        return false;
    }
};
#ifndef TRV_EXTEND_MINBLOCKS
#define TRV_EXTEND_MINBLOCKS 8
#endif
#ifndef TRV_TWO_LEVEL_MINBLOCKS
#define TRV_TWO_LEVEL_MINBLOCKS 6 /* the two-level instantiations carry the world-space ray and the instance state */
#endif
template <bool VOLUMES, bool TL>
__global__ void __launch_bounds__(TRV_BLOCK, TL ? TRV_TWO_LEVEL_MINBLOCKS : TRV_EXTEND_MINBLOCKS) k_extend(Wave w, const __grid_constant__ DScene sc, uint32_t bounce, ExtendTune tune) {
    TRV_DECLARE_STACK(stack);
    __shared__ uint2 stashMem[EXTEND_STASH * TRV_BLOCK];
    ExtendPolicy<VOLUMES, TL> pol(w, bounce);
    TraceCounters cnt;
    traceLoop(sc, pol, w.counters[bounce * CNT_STRIDE + CNT_ACTIVE], &w.counters[bounce * CNT_STRIDE + CNT_FETCH], tune, stack, stashMem + threadIdx.x, cnt);
    traceCountersFlush(w, cnt);
}

/* ------------------------------------------------------------------ k_shade */
/* process_volume_hit.glsl:1-40: free flight through the medium the path is in.  Returns true when the path scatters before vtend
 * (then sp is the scattering point); beta takes the transmittance / collision weights either way.  The light and phase-function
 * sampling of a scattering event (:47-78) happen in k_shade next to the surface's, so that a warp samples its lights once. */
struct MediumHit {
    float3 sp, wo;
    float g;
};
PTC_D bool freeFlight(const DScene &sc, Rng &rng, uint32_t flags, float3 origin, float3 dir, float3 &beta, float vtstart, float vtend, MediumHit &mh) {
    Medium md = loadMedium(sc, flags >> PF_VOL_SHIFT);
    mh.g = fmaxf(fminf(md.g, 0.99f), -0.99f);
    const float3 wdir = normalize(dir);
    mh.wo = -wdir;
    const float distInside = fmaxf(vtend - vtstart, PT_EPSILON);
    const uint32_t channel = min((uint32_t)(rnd(rng) * 3.0f), 2u);
    const float hitDistance = -logf(1.0f - rnd(rng)) / comp(md.sigma_t, (int)channel);
    const bool sampled = hitDistance < distInside;
    const float3 T = exp3(-(md.sigma_t * fminf(hitDistance, distInside)));
    const float3 density = sampled ? (md.sigma_t * T) : T;
    float pdf = (density.x + density.y + density.z) * 0.3333333f;
    if (pdf == 0.0f) pdf = 1.0f;
    beta *= sampled ? (T * md.sigma_s / pdf) : (T / pdf);
    mh.sp = origin + wdir * (vtstart + hitDistance);
    return sampled;
}

/* russian_roulette.glsl:1-11; returns true when the path dies */
PTC_D bool roulette(Rng &rng, uint32_t depth, float3 &beta) {
    const float r = rnd(rng);
    if (depth > 3u) {
        const float mb = max3(beta);
        if (r >= mb) return true;
        beta = beta * (1.0f / mb);
    }
    return false;
}

/* 8 blocks per SM (64 registers, 240 B of spills) beats 6 (80 registers, 80 B) and 5 (96 registers, none): the kernel is DRAM bound and
 * wants warps in flight more than registers - measured 4: 2027, 5: 2123, 6: 2081, 7: 2126, 8: 2139, 9: 2115, 10: 2097, 12: 2011 Mseg/s */
#ifndef SHADE_MINBLOCKS
#define SHADE_MINBLOCKS 8
#endif
/* LIGHTS: the light pick is not empty (rc.totalLights > 0); VOLUMES: a path can be inside a medium (camera volume or an instance that
 * changes it).  The common "environment only, no media" scene gets a kernel without the light-sampling and free-flight code (fewer
 * registers under the same launch bound); the instantiations are result-identical where their preconditions hold. */
#ifndef SHADE_MINBLOCKS_LV
#define SHADE_MINBLOCKS_LV 6 /* the instantiation with lights AND media (the largest): fog 5: 1291, 6: 1298, 7: 1273, 8: 1266 Mseg/s */
#endif
/* SAMPLER (0 default stream, 1 Sobol, 2 PMJ02BN) is a compile-time switch too: the generators are inlined at every rnd() of this
 * kernel, and carrying the two table / hash based ones as run-time branches cost the default path 3-7 % (A/B against the round-1
 * library on one box, profiles/r2_ab_vs_r1.log) */
/* ENV: the environment is light-sampled (PTC_FLAG_ENV_IMPORTANCE, rc.envLight): the table inversion and the MIS of the miss are only
 * compiled into the instantiations that use them */
template <bool LIGHTS, bool VOLUMES, int SAMPLER, bool ENV>
__global__ void __launch_bounds__(128, (LIGHTS && VOLUMES) ? SHADE_MINBLOCKS_LV : SHADE_MINBLOCKS) k_shade(Wave w, const __grid_constant__ DScene sc, const __grid_constant__ RenderConst rc, uint32_t bounce,
                                                         uint32_t firstSample) {
    const uint32_t count = w.counters[bounce * CNT_STRIDE + CNT_ACTIVE];
    const uint32_t *__restrict__ q = w.queue[bounce & 1u];
    uint32_t *__restrict__ qNext = w.queue[(bounce + 1u) & 1u];
    uint32_t *cntNext = &w.counters[(bounce + 1u) * CNT_STRIDE + CNT_ACTIVE];
    uint32_t *cntShadow = &w.counters[bounce * CNT_STRIDE + CNT_SHADOW];
    uint32_t *cntProbe = &w.counters[bounce * CNT_STRIDE + CNT_PROBE];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounded = (count + 31u) & ~31u;
    const bool lastBounce = bounce + 1u >= rc.depth;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += stride) {
        const bool valid = i < count;
        uint32_t slot = 0;
        bool alive = false;
        Requests rq;
        rq.shadow = rq.probe = false;
        if (valid) {
            slot = bounce == 0u ? i : q[i];
            const float4 h = ldS(&w.hit[slot]);
            const float4 o4 = ldS(&w.orgRng[slot]), d4 = ldS(&w.dirFlags[slot]);
            float3 origin = f3(o4), dir = f3(d4);
            uint32_t flags = __float_as_uint(d4.w);
            Rng rng;
            rng.s = __float_as_uint(o4.w);
            rng.ld = (uint32_t)SAMPLER;
            rng.index = rng.pixelSeed = 0u;
            if (SAMPLER != 0) { /* the sampler's key is a function of the slot: (pixel, global sample index) */
                const uint32_t p = slot % rc.nPixLocal, pixel = rc.pixmap ? rc.pixmap[p] : p;
                rngRestore(rng, rng.ld, pixel % rc.width, pixel / rc.width, rc.width, firstSample + slot / rc.nPixLocal);
            }
            flags = (flags & ~PF_DEPTH_MASK) | (bounce & PF_DEPTH_MASK);
            const float4 beta4 = ldS(&w.beta[slot]);
            float3 beta = f3(beta4);
            float lastPdf = beta4.w; /* density of the direction this segment was sampled with (0 = camera ray) */
            float3 radiance = f3(0.0f); /* what THIS event adds (at most one term: first-hit emission or the miss); the slot's running
                                           sum is only read and written when there is something to add: 32 B of state traffic less per segment */
            bool radianceAdded = false;
            const float3 rayDir = dir;
            const int32_t triPos = __float_as_int(h.w);
            bool stop = false;
            bool doRoulette = true;

            /* A. free flight, for rays that hit and rays that leave alike (process_volume_hit.glsl via rchit :60-64 and rmiss :44-51) */
            bool sampledMedium = false;
            MediumHit mh{};
            if (VOLUMES && (flags & PF_INVOL)) {
                const float vtend = triPos >= 0 ? h.x : fminf((float)(uint32_t)rc.sd.volumes[2], 10000.0f);
                sampledMedium = freeFlight(sc, rng, flags, origin, dir, beta, 0.001f, vtend, mh);
            }
            /* B. what kind of event is this, and where does it sample its light from */
            bool surfaceEvent = false; /* a surface that scatters (not passed through, not the first-hit emitter) */
            float3 lightPoint = mh.sp;
            Surf s;
            Frame fr;
            float3 albedo = f3(0.0f), wo = f3(0.0f);
            Pbr pbr;
            bool lambert = false;
            if (!sampledMedium && triPos >= 0) {
                loadSurf(sc, triPos, sc.twoLevel ? w.hitInst[slot] : 0u, h.y, h.z, true, s);
                fr.n = s.n;
                fr.t = s.t;
                const bool flipped = fixFrame(fr, rayDir);
                const ptc_material *mat = s.mat;
                const float tu = s.uv.x * __ldg(&mat->uv_tiling[0]), tv = s.uv.y * __ldg(&mat->uv_tiling[1]);
                lambert = (int)__ldg(&mat->uv_tiling[2]) == PTC_MATERIAL_LAMBERT;
                bool passThrough = false;
                if (__ldg(&mat->metallic_roughness_ao[3]) > 0.0f) { /* stochastic transparency, rchit :66-96 */
                    const float alpha = __ldg(&mat->albedo[3]) * texFetch(sc, __ldg(&mat->tex2[3]), tu, tv).x;
                    const float r = rnd(rng);
                    if (alpha < PT_EPSILON || r > alpha) {
                        if (VOLUMES && s.inst->volFront != s.inst->volBack) volumeChange(s.inst, flipped, flags);
                        origin = s.pos;
                        dir = normalize(rayDir);
                        passThrough = true;
                        doRoulette = false;
                    }
                }
                if (!passThrough) {
                    const float4 normalTexel = texFetch(sc, __ldg(&mat->tex2[1]), tu, tv); /* .w: the roughness sample when the maps were packed */
                    applyNormal(fr, normalFromMap(f3(normalTexel)));
                    albedo = ld3(mat->albedo) * f3(texFetch(sc, __ldg(&mat->tex1[0]), tu, tv));
                    const float3 emissive = ld3(mat->emissive) * __ldg(&mat->emissive[3]) * f3(texFetch(sc, __ldg(&mat->tex2[0]), tu, tv));
                    if (LIGHTS && !VOLUMES && (flags & PF_PROBE_DEFERRED)) {
                        /* The previous event's BSDF probe (next_event_estimation.glsl:1-33) travels along this very ray.  With only
                         * opaque surfaces and no media its first candidate decides it (rayNEE.rahit.glsl:44-54, 73-131) and that candidate
                         * is this closest hit, so the probe was not traced: same terms, same place in the sum.  (The probe's range is
                         * [1e-4, zfar), the path ray's [1e-3, 1e4): the far end is checked, a surface in the first millimetre is not seen.) */
                        flags &= ~PF_PROBE_DEFERRED;
                        if (!isBlackEps(emissive, 0.05f) && h.x < rc.sd.volumes[2] && !(dot(s.n, rayDir) > 0.0f)) {
                            const float3 w0 = mulPoint(s.inst->m, s.p0), w1 = mulPoint(s.inst->m, s.p1), w2 = mulPoint(s.inst->m, s.p2);
                            const float area = 0.5f * length(cross(w1 - w0, w2 - w0));
                            const float pointPdf = (1.0f / (float)s.inst->numTriangles) * (1.0f / area);
                            const float dp = dot(-rayDir, s.n);
                            if (dp > 0.0f) {
                                const float3 ro = (rc.flags & PTC_FLAG_WORLD_ORIGIN_PROBE_PDF) ? origin : mulPoint(s.inst->w2o, origin);
                                const float dd = length(ro - s.pos);
                                const float pdfL = pointPdf * (dd * dd) / dp * (1.0f / (float)rc.totalLights);
                                const float4 bp = ldS(&w.prBetaPdf[slot]);
                                radiance = emissive * f3(bp) * powerHeuristic(bp.w, pdfL);
                                radianceAdded = !isBlack(emissive);
                                NAN_TRAP(BAD3(radiance), "[trap] deferred probe slot %u bounce %u: e %g %g %g bp %g %g %g %g pdfL %g area %g dd %g dp %g\n", slot, bounce, emissive.x, emissive.y,
                                         emissive.z, bp.x, bp.y, bp.z, bp.w, pdfL, area, dd, dp);
                            }
                        }
                    }
                    pbr.albedo = albedo;
                    pbr.metallic = 0.0f;
                    pbr.roughness = 1.0f;
                    if (!lambert) {
                        pbr.metallic = __ldg(&mat->metallic_roughness_ao[0]) * texFetch(sc, __ldg(&mat->tex1[1]), tu, tv).x;
                        const uint32_t roughTex = __ldg(&mat->tex1[2]);
                        const float roughTexel = roughTex == TEX_IN_NORMAL_ALPHA ? normalTexel.w : texFetch(sc, roughTex, tu, tv).x;
                        pbr.roughness = fmaxf(__ldg(&mat->metallic_roughness_ao[1]) * roughTexel, 0.035f);
                    }
                    const bool first = !(flags & PF_SURFACE);
                    if (first) {
                        stS(&w.aovAlbedo[slot], make_float4(albedo.x, albedo.y, albedo.z, 0.0f));
                        float3 nn = fr.n * 0.5f + f3(0.5f);
                        stS(&w.aovNormal[slot], make_float4(nn.x, nn.y, nn.z, 0.0f));
                    }
                    if (first && !isBlackEps(emissive, lambert ? 0.05f : 0.1f) && !flipped) {
                        radiance = emissive * beta; /* emission only at the first surface (trap T2) */
                        radianceAdded = true;
                        stop = true;
                        doRoulette = false;
                    } else {
                        flags |= PF_SURFACE;
                        wo = toLocal(fr, -rayDir);
                        surfaceEvent = true;
                        lightPoint = s.pos;
                    }
                }
            } else if (!sampledMedium) { /* rayPrimary.rmiss.glsl:53-106 */
                stop = true;
                doRoulette = false;
                const float *bg = rc.sd.background;
                const bool first = !(flags & PF_SURFACE);
                float3 col = f3(0.0f), aov = f3(0.0f);
                if (bg[3] == 0.0f) {
                    col = aov = f3(bg[0], bg[1], bg[2]);
                } else if (bg[3] == 1.0f) {
                    col = aov = envFetch(sc, rayDir) * rc.sd.exposure[1];
                } else if (bg[3] == 2.0f) {
                    aov = f3(bg[0], bg[1], bg[2]);
                    col = first ? aov : envFetch(sc, rayDir);
                }
                if (first) {
                    stS(&w.aovAlbedo[slot], make_float4(aov.x, aov.y, aov.z, 0.0f));
                    stS(&w.aovNormal[slot], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                }
                /* environment found by a sampled direction while it is also light-sampled: power heuristic */
                if (ENV && lastPdf > 0.0f && (bg[3] == 1.0f || (bg[3] == 2.0f && !first))) {
                    const float2 uv = envd::equirectUV(normalize(rayDir));
                    const envd::DeviceTables et{sc.envCdfV, sc.envCdfU};
                    const float pl = (1.0f / (float)rc.totalLights) * envd::solidAnglePdf(envd::pdfUV(et, uv.x, uv.y), uv.y);
                    col = col * powerHeuristic(lastPdf, pl);
                }
                radiance = col * beta;
                radianceAdded = true;
            }
            /* C. one light sample for every lane that scatters, in the medium or on a surface (lightSampling.glsl:1-106) */
            LightSample ls;
            ls.radiance = f3(0.0f);
            if (LIGHTS && (sampledMedium || surfaceEvent)) ls = sampleLight<ENV>(sc, rc, rng, lightPoint);
            /* D. its weight, the shadow request, and the next direction */
            if (VOLUMES && sampledMedium) { /* process_volume_hit.glsl:47-78 */
                if (LIGHTS && !isBlack(ls.radiance)) {
                    const float p = hg(dot(mh.wo, ls.dir), mh.g);
                    if (p != 0.0f) {
                        float3 c = ls.radiance * p * beta / ls.pdf;
                        if (!ls.delta) c = c * powerHeuristic(ls.pdf, p);
                        rq.shadow = true;
                        rq.shOrigin = mh.sp;
                        rq.shDir = ls.dir;
                        rq.shTmax = ls.tmax;
                        rq.shContrib = c;
                    }
                }
                const float2 h01 = rnd2(rng);
                float3 nd;
                const float spdf = hgSample(mh.wo, nd, h01.x, h01.y, mh.g);
                origin = mh.sp;
                dir = nd;
                rq.probe = true;
                rq.prBeta = beta;
                rq.prPdf = spdf;
            } else if (surfaceEvent) {
                if (LIGHTS && !isBlack(ls.radiance)) {
                    const float3 wi = toLocal(fr, ls.dir);
                    float3 F;
                    float bsdfPdf;
                    if (lambert) {
                        const float c = clampf(wi.y, 0.0f, 1.0f);
                        F = albedo * PT_INV_PI * c;
                        bsdfPdf = c * PT_INV_PI;
                    } else {
                        F = pbrEval(pbr, wi, wo);
                        bsdfPdf = ls.delta ? 0.0f : pbrPdf(wi, wo, pbr);
                    }
                    if (!isBlack(F)) {
                        float3 c = ls.radiance * F * beta / ls.pdf;
                        if (!ls.delta) c = c * powerHeuristic(ls.pdf, bsdfPdf);
                        NAN_TRAP(BAD3(c), "[trap] surface NEE slot %u bounce %u: c %g %g %g L %g %g %g F %g %g %g beta %g %g %g lpdf %g bpdf %g delta %d lambert %d wi %g %g %g wo %g %g %g\n", slot, bounce,
                                 c.x, c.y, c.z, ls.radiance.x, ls.radiance.y, ls.radiance.z, F.x, F.y, F.z, beta.x, beta.y, beta.z, ls.pdf, bsdfPdf, (int)ls.delta, (int)lambert, wi.x, wi.y, wi.z, wo.x, wo.y, wo.z);
                        rq.shadow = true;
                        rq.shOrigin = s.pos;
                        rq.shDir = ls.dir;
                        rq.shTmax = ls.tmax;
                        rq.shContrib = c;
                    }
                }
                float spdf;
                float3 wiL;
                if (lambert) {
                    const float2 a01 = rnd2(rng);
                    wiL = cosineHemisphere(a01.x, a01.y, spdf);
                    origin = s.pos;
                    dir = toWorld(fr, wiL);
                    beta *= albedo;
                } else {
                    const float2 a01 = rnd2(rng);
                    const float a2 = rnd(rng);
                    const float3 F = pbrSample(wiL, wo, spdf, pbr, a01.x, a01.y, a2);
                    origin = s.pos;
                    dir = toWorld(fr, wiL);
                    if (isBlack(F)) {
                        stop = true;
                        doRoulette = false;
                    } else {
                        beta *= clamp3(F / spdf, 0.0f, 1.0f);
                    }
                }
                if (!stop) {
                    rq.probe = true;
                    rq.prBeta = beta;
                    rq.prPdf = spdf;
                }
            }
            NAN_TRAP(BAD3(beta) || BAD3(dir) || BAD3(origin), "[trap] state slot %u bounce %u: beta %g %g %g dir %g %g %g origin %g %g %g surfaceEvent %d lambert %d\n", slot, bounce, beta.x, beta.y,
                     beta.z, dir.x, dir.y, dir.z, origin.x, origin.y, origin.z, (int)surfaceEvent, (int)lambert);
            NAN_TRAP(radianceAdded && BAD3(radiance), "[trap] radiance term slot %u bounce %u: %g %g %g\n", slot, bounce, radiance.x, radiance.y, radiance.z);
            if (doRoulette && !stop) stop = roulette(rng, bounce, beta);
            if (rq.probe) lastPdf = rq.prPdf; /* a new direction was sampled (surface or medium) */
            /* requests that can only return black are dropped (result-identical) */
            if (rq.probe && !sc.anyEmissive) rq.probe = false;
            if (rq.probe && sc.nEmissiveBoxes != 0u && !rayMeetsEmitters(sc, origin, dir)) rq.probe = false;
            alive = !stop && !lastBounce;
            flags &= ~PF_PROBE_DEFERRED; /* (a miss or a pass-through leaves it set; the probe of a missed ray adds nothing) */
            bool deferProbe = false;
            if (LIGHTS && !VOLUMES && rq.probe && rc.fuseProbe && alive) { /* the next path ray answers this probe */
                deferProbe = true;
                rq.probe = false;
                flags |= PF_PROBE_DEFERRED;
            }
            /* a finished path's state is never read again (a probe request still needs the new origin and direction) */
            if (alive || rq.probe) {
                stS(&w.orgRng[slot], make_float4(origin.x, origin.y, origin.z, __uint_as_float(rng.s)));
                stS(&w.dirFlags[slot], make_float4(dir.x, dir.y, dir.z, __uint_as_float(flags)));
            }
            if (alive) stS(&w.beta[slot], make_float4(beta.x, beta.y, beta.z, lastPdf));
            if (VOLUMES && alive) {
                /* how far the next segment has to be traced (ExtendPolicy): inside a medium its free-flight distance is already
                 * determined - the next event draws the same two numbers from the stored state (freeFlight) */
                float tmaxNext = 10000.0f;
                if (flags & PF_INVOL) {
                    Rng ahead = rng;
                    const Medium md = loadMedium(sc, flags >> PF_VOL_SHIFT);
                    const uint32_t channel = min((uint32_t)(rnd(ahead) * 3.0f), 2u);
                    const float hitDistance = -logf(1.0f - rnd(ahead)) / comp(md.sigma_t, (int)channel);
                    /* only when the path scatters before the far end the miss shader uses (rayPrimary.rmiss.glsl:44-51) */
                    if (hitDistance < fminf((float)(uint32_t)rc.sd.volumes[2], 10000.0f) - 0.001f) tmaxNext = fminf(0.001f + hitDistance, 10000.0f);
                }
                stS(&w.hit[slot], make_float4(tmaxNext, 0.0f, 0.0f, __int_as_float(-1)));
            }
            if (radianceAdded) {
                const float3 sum = f3(ldS(&w.radiance[slot])) + radiance; /* same single addition as before: bit-identical */
                stS(&w.radiance[slot], make_float4(sum.x, sum.y, sum.z, 0.0f));
            }
            if (rq.shadow) {
                stS(&w.shOrgTmax[slot], make_float4(rq.shOrigin.x, rq.shOrigin.y, rq.shOrigin.z, rq.shTmax));
                stS(&w.shDirVol[slot], make_float4(rq.shDir.x, rq.shDir.y, rq.shDir.z, __uint_as_float(flags)));
                stS(&w.shContrib[slot], make_float4(rq.shContrib.x, rq.shContrib.y, rq.shContrib.z, 0.0f));
            }
            if (rq.probe || deferProbe) stS(&w.prBetaPdf[slot], make_float4(rq.prBeta.x, rq.prBeta.y, rq.prBeta.z, rq.prPdf));
        }
        queuePush(qNext, cntNext, alive, slot);
        queuePush(w.qShadow, cntShadow, rq.shadow, slot);
        queuePush(w.qProbe, cntProbe, rq.probe, slot);
    }
}

/* ------------------------------------------------------------------ k_shadow */
/* lightSampling.glsl:108-144 + raySecondary.rahit/.rchit/.rmiss with nearest-first candidate order (trap T1), as a
 * per-lane state machine on top of traceLoop: one query = the next-nearest candidate of the current hop. */
template <bool OPAQUE, bool TL> /* OPAQUE: no material of the scene is transparent - every hit shadows, the candidate code is not compiled */
struct ShadowPolicy {
    static constexpr bool TWO_LEVEL = TL;
    /* Without media (no instance changes the volume, the camera is in none) the shadow chain is one ray whose transmittance is the
     * product of (1 - alpha) over the transparent surfaces it crosses, zero if any of them is opaque (raySecondary.rahit.glsl:31-72):
     * independent of the order, so one traversal that looks at every crossed triangle replaces one ordered query per surface. */
    static constexpr bool ALL_HITS = !OPAQUE;
    static constexpr bool opaqueScene = OPAQUE;
    const Wave &w;
    const DScene &sc;
    const RenderConst &rc;
    uint32_t slot, vol, hop, hops;
    float3 origin, thr; /* the direction lives in the traversal state */
    float distanceT, vtmin;
    bool everyHit;
    PTC_D ShadowPolicy(const Wave &w_, const DScene &sc_, const RenderConst &rc_)
        : w(w_), sc(sc_), rc(rc_), slot(0), vol(0), hop(0), hops(0), distanceT(0.0f), vtmin(0.0f),
          everyHit(!OPAQUE && !sc_.anyVolume && rc_.sd.volumes[0] == -1.0f) {}
    PTC_D bool anyHit() const { return opaqueScene || everyHit; } /* every surface is opaque: any hit shadows; all-hits: a decided ray is marked hit */
    PTC_D bool allHits() const { return everyHit; }
    /* a triangle the ray crosses (all-hits mode): true = the ray is shadowed */
    PTC_D bool candidate(int32_t pos, uint32_t hitInst, float u, float v) {
        const float4 *__restrict__ r = sc.shading + 9 * (size_t)pos;
        const ptc_material *mat = &sc.materials[sc.instances[TL ? hitInst : __float_as_uint(__ldg(r + 6).w)].material];
        if (!(__ldg(&mat->metallic_roughness_ao[3]) >= 0.99f)) return true; /* raySecondary.rahit.glsl:42-47 */
        const float w0 = 1.0f - u - v;
        const float uu = __ldg(r + 0).w * w0 + __ldg(r + 2).w * u + __ldg(r + 4).w * v, vv = __ldg(r + 1).w * w0 + __ldg(r + 3).w * u + __ldg(r + 5).w * v;
        const float alpha = __ldg(&mat->albedo[3]) * texFetch(sc, __ldg(&mat->tex2[3]), uu * __ldg(&mat->uv_tiling[0]), vv * __ldg(&mat->uv_tiling[1])).x;
        thr = thr * (1.0f - alpha);
        return !(max3(thr) > PT_EPSILON);
    }
    PTC_D void startHop(trv::Trav &tr) {
        const float tmin = 0.0001f;
        vtmin = tmin;
        tr.o = origin;
        tr.tmin = tr.t0 = tmin;
        tr.tmax = distanceT;
        tr.id0 = 0xffffffffu;
        hops++;
    }
    PTC_D bool begin(uint32_t i, trv::Trav &tr) {
        slot = w.qShadow[i];
        const float4 o4 = ldS(&w.shOrgTmax[slot]), d4 = ldS(&w.shDirVol[slot]);
        origin = f3(o4);
        tr.d = f3(d4);
        vol = __float_as_uint(d4.w);
        distanceT = o4.w - 0.0001f;
        thr = f3(1.0f);
        hop = 0;
        if (rc.depth == 0u) return finish(false);
        startHop(tr);
        return true;
    }
    PTC_D bool finish(bool shadowed) {
        if (!shadowed) {
            const float3 c = f3(ldS(&w.shContrib[slot])) * thr;
            float4 r = ldS(&w.radiance[slot]);
            r.x += c.x;
            r.y += c.y;
            r.z += c.z;
            stS(&w.radiance[slot], r);
        }
        return false;
    }
    PTC_D bool next(trv::Trav &tr) {
        const trv::HitRec h = tr.best;
        const float3 dir = tr.d;
        if (h.pos < 0) { /* raySecondary.rmiss.glsl:17-44 */
            bool shadowed = false;
            if (vol & PF_INVOL) {
                const float zfarTrunc = (float)(uint32_t)rc.sd.volumes[2];
                thr = thr * transmittance(sc, vol >> PF_VOL_SHIFT, vtmin, fminf(zfarTrunc, distanceT));
                shadowed = !(max3(thr) > PT_EPSILON);
            }
            return finish(shadowed);
        }
        if (opaqueScene || everyHit) return finish(true);
        if (!(__ldg(&hitMaterial(sc, h.pos, h.inst)->metallic_roughness_ao[3]) >= 0.99f)) return finish(true); /* raySecondary.rahit.glsl:42-47 */
        Surf s;
        loadSurf(sc, h.pos, h.inst, h.u, h.v, false, s);
        const ptc_material *mat = s.mat;
        const float alpha =
            __ldg(&mat->albedo[3]) * texFetch(sc, __ldg(&mat->tex2[3]), s.uv.x * __ldg(&mat->uv_tiling[0]), s.uv.y * __ldg(&mat->uv_tiling[1])).x;
        thr = thr * (1.0f - alpha);
        if (s.inst->volFront != s.inst->volBack) { /* accept: raySecondary.rchit.glsl:32-78 */
            const bool flipped = dot(s.n, dir) > 0.0f;
            if (vol & PF_INVOL) {
                thr = thr * transmittance(sc, vol >> PF_VOL_SHIFT, vtmin, h.t);
                if (max3(thr) < PT_EPSILON) return finish(true);
            }
            volumeChange(s.inst, flipped, vol);
            origin = s.pos;
            distanceT -= h.t;
            if (++hop >= rc.depth) return finish(false); /* lightSampling.glsl:118: the hop loop runs out */
            startHop(tr);
            return true;
        }
        if (max3(thr) > PT_EPSILON) { /* ignoreIntersectionEXT: same ray, next-nearest candidate */
            tr.t0 = h.t;
            tr.id0 = h.worldId;
            return true;
        }
        return finish(true);
    }
};
/* shadow / probe chains: 7 blocks per SM (72 registers) - measured on fog / progressive / Cornell: 5: 795 / 1462 / 1985, 6: 826 / 1500 / 2036,
 * 7: 847 / 1529 / 2071, 8: 845 / 1523 / 2027 Mseg/s */
#ifndef CHAIN_MINBLOCKS
#define CHAIN_MINBLOCKS 7
#endif
template <bool OPAQUE, bool TL>
__global__ void __launch_bounds__(TRV_BLOCK, TL ? TRV_TWO_LEVEL_MINBLOCKS : CHAIN_MINBLOCKS) k_shadow(Wave w, const __grid_constant__ DScene sc, const __grid_constant__ RenderConst rc, uint32_t bounce,
                                                      ExtendTune tune) {
    TRV_DECLARE_STACK(stack);
    __shared__ uint2 stashMem[EXTEND_STASH * TRV_BLOCK];
    ShadowPolicy<OPAQUE, TL> pol(w, sc, rc);
    TraceCounters cnt;
    traceLoop(sc, pol, w.counters[bounce * CNT_STRIDE + CNT_SHADOW], &w.counters[bounce * CNT_STRIDE + CNT_FETCH_SHADOW], tune, stack,
              stashMem + threadIdx.x, cnt);
    statAdd(&w.stats[ST_SHADOW_HOPS], pol.hops);
}

/* ------------------------------------------------------------------ k_probe */
/* next_event_estimation.glsl:1-33 + rayNEE.rahit/.rchit/.rmiss with nearest-first candidate order (trap T1) */
template <bool TL>
struct ProbePolicy {
    static constexpr bool TWO_LEVEL = TL;
    static constexpr bool ALL_HITS = false; /* the probe's any-hit decisions depend on the order of the candidates */
    const Wave &w;
    const DScene &sc;
    const RenderConst &rc;
    uint32_t slot, vol, hop, hops;
    float3 origin, thr; /* the direction lives in the traversal state */
    float vtmin;
    PTC_D ProbePolicy(const Wave &w_, const DScene &sc_, const RenderConst &rc_) : w(w_), sc(sc_), rc(rc_), slot(0), vol(0), hop(0), hops(0), vtmin(0.0f) {}
    PTC_D bool anyHit() const { return false; }
    PTC_D void startHop(trv::Trav &tr) {
        const float tmin = 0.0001f;
        vtmin = tmin;
        tr.o = origin;
        tr.tmin = tr.t0 = tmin;
        tr.tmax = rc.sd.volumes[2];
        tr.id0 = 0xffffffffu;
        hops++;
    }
    PTC_D bool begin(uint32_t i, trv::Trav &tr) {
        slot = w.qProbe[i];
        const float4 o4 = ldS(&w.orgRng[slot]), d4 = ldS(&w.dirFlags[slot]);
        origin = f3(o4);
        tr.d = f3(d4);
        vol = __float_as_uint(d4.w);
        thr = f3(1.0f);
        hop = 0;
        if (rc.depth == 0u) return false;
        startHop(tr);
        return true;
    }
    PTC_D bool finish(float3 emissive, float pdf) {
        if (!isBlack(emissive)) {
            const float4 bp = ldS(&w.prBetaPdf[slot]);
            const float wgt = powerHeuristic(bp.w, pdf);
            const float3 c = thr * emissive * f3(bp) * wgt;
            float4 r = ldS(&w.radiance[slot]);
            r.x += c.x;
            r.y += c.y;
            r.z += c.z;
            stS(&w.radiance[slot], r);
        }
        return false;
    }
    PTC_D bool next(trv::Trav &tr) {
        const trv::HitRec h = tr.best;
        const float3 dir = tr.d;
        if (h.pos < 0) return false; /* rayNEE.rmiss.glsl:12-19: nothing (the environment is not included, trap T6) */
        {
            /* rayNEE.rahit.glsl:44-54 decided from the material alone: the emissive texel is in [0, 1], so a material whose emissive
             * factor is already below the threshold cannot pass it, and an opaque one ends the ray (result-identical shortcut) */
            const ptc_material *m0 = hitMaterial(sc, h.pos, h.inst);
            if (!(__ldg(&m0->metallic_roughness_ao[3]) >= 0.99f) && isBlackEps(ld3(m0->emissive) * __ldg(&m0->emissive[3]), 0.05f)) return false;
        }
        Surf s;
        loadSurf(sc, h.pos, h.inst, h.u, h.v, false, s);
        const ptc_material *mat = s.mat;
        const float tu = s.uv.x * __ldg(&mat->uv_tiling[0]), tv = s.uv.y * __ldg(&mat->uv_tiling[1]);
        const float3 em = ld3(mat->emissive) * __ldg(&mat->emissive[3]) * f3(texFetch(sc, __ldg(&mat->tex2[0]), tu, tv));
        const bool isTransparent = __ldg(&mat->metallic_roughness_ao[3]) >= 0.99f;
        if (isBlackEps(em, 0.05f)) { /* rayNEE.rahit.glsl:44-71 */
            if (!isTransparent) return false;
            const float alpha = __ldg(&mat->albedo[3]) * texFetch(sc, __ldg(&mat->tex2[3]), tu, tv).x;
            thr = thr * (1.0f - alpha);
            if (s.inst->volFront != s.inst->volBack) { /* accept: rayNEE.rchit.glsl:33-82 */
                const bool flipped = dot(s.n, dir) > 0.0f;
                if (vol & PF_INVOL) {
                    thr = thr * transmittance(sc, vol >> PF_VOL_SHIFT, vtmin, h.t);
                    if (max3(thr) < PT_EPSILON) return false;
                }
                volumeChange(s.inst, flipped, vol);
                origin = s.pos;
                if (++hop >= rc.depth) return false; /* the hop loop runs out without an emitter */
                startHop(tr);
                return true;
            }
            tr.t0 = h.t; /* ignoreIntersectionEXT: same ray, next-nearest candidate */
            tr.id0 = h.worldId;
            return true;
        }
        /* emissive surface, rayNEE.rahit.glsl:73-131 */
        if (vol & PF_INVOL) {
            thr = thr * transmittance(sc, vol >> PF_VOL_SHIFT, vtmin, h.t);
            if (max3(thr) < PT_EPSILON) return false;
        }
        if (dot(s.n, dir) > 0.0f) return false; /* back face */
        const float3 w0 = mulPoint(s.inst->m, s.p0), w1 = mulPoint(s.inst->m, s.p1), w2 = mulPoint(s.inst->m, s.p2);
        const float area = 0.5f * length(cross(w1 - w0, w2 - w0));
        const float pointPdf = (1.0f / (float)s.inst->numTriangles) * (1.0f / area);
        const float dp = dot(-dir, s.n);
        if (!(dp > 0.0f)) return false;
        /* rayNEE.rahit.glsl:122 measures from gl_ObjectRayOriginEXT (trap T6) unless the flag asks otherwise */
        const float3 ro = (rc.flags & PTC_FLAG_WORLD_ORIGIN_PROBE_PDF) ? origin : mulPoint(s.inst->w2o, origin);
        const float dd = length(ro - s.pos);
        const float pdf = pointPdf * (dd * dd) / dp * (1.0f / (float)rc.totalLights);
        return finish(em, pdf);
    }
};
template <bool TL>
__global__ void __launch_bounds__(TRV_BLOCK, TL ? TRV_TWO_LEVEL_MINBLOCKS : CHAIN_MINBLOCKS) k_probe(Wave w, const __grid_constant__ DScene sc, const __grid_constant__ RenderConst rc, uint32_t bounce,
                                                     ExtendTune tune) {
    TRV_DECLARE_STACK(stack);
    __shared__ uint2 stashMem[EXTEND_STASH * TRV_BLOCK];
    ProbePolicy<TL> pol(w, sc, rc);
    TraceCounters cnt;
    traceLoop(sc, pol, w.counters[bounce * CNT_STRIDE + CNT_PROBE], &w.counters[bounce * CNT_STRIDE + CNT_FETCH_PROBE], tune, stack,
              stashMem + threadIdx.x, cnt);
    statAdd(&w.stats[ST_PROBE_HOPS], pol.hops);
}

/* ------------------------------------------------------------------ k_accumulate (raygen.rgen.glsl:126-144) */
__global__ void __launch_bounds__(256) k_accumulate(Wave w, const __grid_constant__ RenderConst rc, uint32_t nSamples, float4 *__restrict__ accR,
                                                    float4 *__restrict__ accA, float4 *__restrict__ accN) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rc.nPixLocal) return;
    const float total = (float)rc.totalSamples;
    float3 cr = f3(0.0f), ca = f3(0.0f), cn = f3(0.0f);
    for (uint32_t s = 0; s < nSamples; s++) {
        const size_t slot = (size_t)s * rc.nPixLocal + p;
        cr += f3(ldS(&w.radiance[slot])) / total;
        ca += f3(ldS(&w.aovAlbedo[slot])) / total;
        cn += f3(ldS(&w.aovNormal[slot])) / total;
    }
    const uint32_t pixel = rc.pixmap ? rc.pixmap[p] : p;
    /* alpha is not touched here: partitioned renders are SUMMED over ranks, so exactly one place writes the 1 (k_fill_alpha, after
     * the reduce / on the rank that owns the image) */
    float4 r = accR[pixel], a = accA[pixel], n = accN[pixel];
    accR[pixel] = make_float4(r.x + cr.x, r.y + cr.y, r.z + cr.z, r.w);
    accA[pixel] = make_float4(a.x + ca.x, a.y + ca.y, a.z + ca.z, a.w);
    accN[pixel] = make_float4(n.x + cn.x, n.y + cn.y, n.z + cn.z, n.w);
}

/* tile split: block k fills the pixels of this rank's k-th tile (tileIds[k], row-major tile index), row-major inside the tile;
 * offsets[k] = first local pixel index of tile k (edge tiles are smaller) */
__global__ void __launch_bounds__(256) k_tile_pixmap(uint32_t W, uint32_t H, uint32_t tile, const uint32_t *__restrict__ tileIds, const uint32_t *__restrict__ offsets,
                                                     uint32_t *__restrict__ pixmap) {
    const uint32_t tilesX = (W + tile - 1) / tile;
    const uint32_t t = tileIds[blockIdx.x], tx = t % tilesX, ty = t / tilesX;
    const uint32_t x0 = tx * tile, y0 = ty * tile, tw = min(tile, W - x0), th = min(tile, H - y0);
    const uint32_t base = offsets[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < tw * th; i += blockDim.x) pixmap[base + i] = (y0 + i / tw) * W + x0 + i % tw;
}

/* sums the per-bounce queue counters into the statistics block */
__global__ void k_collect_stats(Wave w, uint32_t depth) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long seg = 0, sh = 0, pr = 0;
    for (uint32_t d = 0; d < depth; d++) {
        seg += w.counters[d * CNT_STRIDE + CNT_ACTIVE];
        sh += w.counters[d * CNT_STRIDE + CNT_SHADOW];
        pr += w.counters[d * CNT_STRIDE + CNT_PROBE];
    }
    w.stats[ST_SEGMENTS] += seg;
    w.stats[ST_SHADOW_RAYS] += sh;
    w.stats[ST_PROBE_RAYS] += pr;
}

/* ------------------------------------------------------------------ parity-hook kernels */
/* closest hits of a caller's ray set, through the same warp-persistent loop as the render */
template <bool TL>
struct RaySetPolicy {
    static constexpr bool ALL_HITS = false;
    static constexpr bool TWO_LEVEL = TL;
    const DScene &sc;
    const float *__restrict__ rays;
    int *inst, *prim;
    float *t, *u, *v;
    uint32_t i;
    float tmaxRay;
    PTC_D RaySetPolicy(const DScene &sc_, const float *rays_, int *inst_, int *prim_, float *t_, float *u_, float *v_)
        : sc(sc_), rays(rays_), inst(inst_), prim(prim_), t(t_), u(u_), v(v_), i(0), tmaxRay(0.0f) {}
    PTC_D bool anyHit() const { return false; }
    PTC_D bool begin(uint32_t index, trv::Trav &tr) {
        i = index;
        const float *r = rays + (size_t)i * 8;
        tr.o = f3(r[0], r[1], r[2]);
        tr.d = f3(r[4], r[5], r[6]);
        tr.tmin = tr.t0 = r[3];
        tr.tmax = tmaxRay = r[7];
        tr.id0 = 0xffffffffu;
        return true;
    }
    PTC_D bool next(trv::Trav &tr) {
        const trv::HitRec &h = tr.best;
        if (h.pos >= 0) {
            inst[i] = TL ? (int)h.inst : (int)__float_as_uint(sc.tris[3 * (size_t)h.pos + 0].w);
            prim[i] = (int)__float_as_uint(sc.tris[3 * (size_t)h.pos + 1].w);
            t[i] = h.t;
            u[i] = h.u;
            v[i] = h.v;
        } else {
            inst[i] = -1;
            prim[i] = -1;
            t[i] = tmaxRay;
            u[i] = 0.0f;
            v[i] = 0.0f;
        }
        return false;
    }
};
template <bool TL>
__global__ void __launch_bounds__(TRV_BLOCK) k_trace_closest(const __grid_constant__ DScene sc, const float *__restrict__ rays, uint32_t n, uint32_t *fetchCounter, int *inst,
                                                             int *prim, float *t, float *u, float *v, ExtendTune tune) {
    TRV_DECLARE_STACK(stack);
    __shared__ uint2 stashMem[EXTEND_STASH * TRV_BLOCK];
    RaySetPolicy<TL> pol(sc, rays, inst, prim, t, u, v);
    TraceCounters cnt;
    traceLoop(sc, pol, n, fetchCounter, tune, stack, stashMem + threadIdx.x, cnt);
}

__global__ void k_bsdf_eval(int n, const float *params, const float *wi, const float *wo, float *outF, float *outPdf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pbr p{f3(params[i * 5], params[i * 5 + 1], params[i * 5 + 2]), params[i * 5 + 3], params[i * 5 + 4]};
    float3 a = f3(wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]), b = f3(wo[i * 3], wo[i * 3 + 1], wo[i * 3 + 2]);
    float3 F = pbrEval(p, a, b);
    outF[i * 3] = F.x;
    outF[i * 3 + 1] = F.y;
    outF[i * 3 + 2] = F.z;
    outPdf[i] = pbrPdf(a, b, p);
}
__global__ void k_bsdf_sample(int n, const float *params, const float *wo, const float *u, float *outWi, float *outF, float *outPdf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pbr p{f3(params[i * 5], params[i * 5 + 1], params[i * 5 + 2]), params[i * 5 + 3], params[i * 5 + 4]};
    float3 b = f3(wo[i * 3], wo[i * 3 + 1], wo[i * 3 + 2]), wi;
    float pdf;
    float3 F = pbrSample(wi, b, pdf, p, u[i * 3], u[i * 3 + 1], u[i * 3 + 2]);
    outWi[i * 3] = wi.x;
    outWi[i * 3 + 1] = wi.y;
    outWi[i * 3 + 2] = wi.z;
    outF[i * 3] = F.x;
    outF[i * 3 + 1] = F.y;
    outF[i * 3 + 2] = F.z;
    outPdf[i] = pdf;
}
__global__ void k_sampler_points(uint32_t px, uint32_t py, uint32_t width, uint32_t firstIndex, uint32_t count, uint32_t dimension, uint32_t flags,
                                 float *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Rng r = rngInit(px, py, width, firstIndex + i, samplerOfFlags(flags));
    if (r.ld == 2u)
        r.s += dimension; /* PMJ02BN: relative to the sample's start dimension; samplesPerPixel = firstIndex + count (g_pmj.spp) */
    else if (r.ld)
        r.s = dimension;
    else
        for (uint32_t k = 0; k < dimension; k++) rnd(r);
    float2 p;
    if (flags & PTC_SAMPLER_HOOK_1D) {
        p.x = rnd(r);
        p.y = rnd(r);
    } else {
        p = rnd2(r);
    }
    out[2 * i] = p.x;
    out[2 * i + 1] = p.y;
}
__global__ void k_srgb_table(cudaTextureObject_t tex, float *__restrict__ out) {
    const uint32_t i = threadIdx.x;
    out[i] = tex2DLayered<float4>(tex, ((float)i + 0.5f) / 256.0f, 0.5f, 0).x;
}
__global__ void k_env_lookup(const __grid_constant__ DScene sc, int n, const float *dirs, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 c = envFetch(sc, f3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]));
    out[i * 3] = c.x;
    out[i * 3 + 1] = c.y;
    out[i * 3 + 2] = c.z;
}

__global__ void k_env_sample(const __grid_constant__ DScene sc, int n, const float *u01, float *dirs, float *pdf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const envd::DeviceTables et{sc.envCdfV, sc.envCdfU};
    float u, v, puv;
    envd::sampleUV(et, u01[2 * i], u01[2 * i + 1], u, v, puv);
    const float3 d = envd::direction(u, v);
    dirs[i * 3] = d.x;
    dirs[i * 3 + 1] = d.y;
    dirs[i * 3 + 2] = d.z;
    pdf[i] = envd::solidAnglePdf(puv, v);
}
__global__ void k_env_pdf(const __grid_constant__ DScene sc, int n, const float *dirs, float *pdf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const envd::DeviceTables et{sc.envCdfV, sc.envCdfU};
    const float2 uv = envd::equirectUV(normalize(f3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2])));
    pdf[i] = envd::solidAnglePdf(envd::pdfUV(et, uv.x, uv.y), uv.y);
}

/* ------------------------------------------------------------------ equirect -> cubemap (skyboxCubemapWrite.frag.glsl:12-15) */
__global__ void k_equirect_to_cube(cudaTextureObject_t equirect, uint32_t N, float4 *__restrict__ faces) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (i >= N || j >= N) return;
    const float sc = 2.0f * ((float)i + 0.5f) / (float)N - 1.0f, tc = 2.0f * ((float)j + 0.5f) / (float)N - 1.0f;
    float3 d;
    switch (f) { /* inverse of the cubemap face selection rules */
        case 0: d = f3(1.0f, -tc, -sc); break;
        case 1: d = f3(-1.0f, -tc, sc); break;
        case 2: d = f3(sc, 1.0f, tc); break;
        case 3: d = f3(sc, -1.0f, -tc); break;
        case 4: d = f3(sc, -tc, 1.0f); break;
        default: d = f3(-sc, -tc, -1.0f); break;
    }
    d = normalize(d);
    /* include/environmentMap.glsl:1-10 */
    float ux = atan2f(d.z, d.x) * 0.1591f + 0.5f + 0.25f;
    ux = ux - floorf(ux);
    const float uy = asinf(clampf(d.y, -1.0f, 1.0f)) * 0.3183f + 0.5f;
    faces[((size_t)f * N + j) * N + i] = tex2D<float4>(equirect, ux, uy);
}

}  // namespace wf
