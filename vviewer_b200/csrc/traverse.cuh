/*
 * Software BVH traversal (B200 has no RT cores; this replaces traceRayEXT / the driver's ray-triangle test,
 * raygen.rgen.glsl:110, lightSampling.glsl:128, next_event_estimation.glsl:21).
 *
 * Hit rule (ours to define): among triangles with tmin < t < tmax the smallest (t, world
 * triangle id); no face culling (VulkanScene.cpp:364-377).  Moeller-Trumbore on (v0, e1, e2), same operation
 * order as oracle/accel.hpp::intersectTri.
 *
 * Node fetches are 128-bit (4 x float4 per binary node: both child boxes + child links), triangle fetches are
 * 3 x float4; the traversal stack lives in shared memory (one column per thread, bank-conflict free) with a
 * local-memory spill that is never reached by LBVH depths seen in practice.
 *
 * The traversal is a resumable state machine (Trav) so that kernels can run it warp-convergently: persistent warps
 * pull rays from a global counter with warp-aggregated atomics and replace finished rays while the other lanes keep
 * traversing (Aila & Laine 2009 "while-while" with dynamic fetch) instead of letting the warp fragment.
 */
#pragma once
#include "common.cuh"

namespace trv {

#define TRV_STACK 64
#define TRV_SHARED_STACK 24
#define TRV_BLOCK 128 /* threads per block of every kernel that traverses */
#define TRV_DONE 0x7fffffff

struct Ray {
    float3 o, d;
    float tmin, tmax;
};
struct HitRec {
    float t, u, v;
    int32_t pos;      /* position in the Morton-ordered triangle array, -1 = miss */
    uint32_t worldId; /* world triangle id (instance-major), the tie-break key */
};

PTC_D bool intersectTri(const float4 v0, const float4 e1, const float4 e2, const float3 o, const float3 d, float &t, float &u, float &v) {
    float3 E1 = f3(e1), E2 = f3(e2);
    float3 p = cross(d, E2);
    float det = dot(E1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float3 s = o - f3(v0);
    u = dot(s, p) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    float3 q = cross(s, E1);
    v = dot(d, q) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(E2, q) * inv;
    return true;
}

/* Ordered query state: finds the smallest (t, id) lexicographically greater than (t0, id0) with t < tmax.
 * Closest hit: (t0, id0) = (tmin, 0xffffffff). */
struct Trav {
    float3 o, d, idir, ood;
    float tmin, tmax, t0;
    uint32_t id0;
    HitRec best;
    int32_t node;
    int sp;
    int32_t spill[TRV_STACK];

    PTC_D void init(const DScene &sc, const Ray &ray, float t0_, uint32_t id0_) {
        o = ray.o;
        d = ray.d;
        tmin = ray.tmin;
        tmax = ray.tmax;
        t0 = t0_;
        id0 = id0_;
        const float ooeps = 1e-20f;
        idir = f3(1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x)), 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y)),
                  1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z)));
        ood = o * idir;
        best.t = tmax;
        best.u = best.v = 0.0f;
        best.pos = -1;
        best.worldId = 0xffffffffu;
        sp = 0;
        node = sc.nTris == 0 ? TRV_DONE : (sc.rootIsLeaf ? ~0 : 0);
    }
    PTC_D bool done() const { return node == TRV_DONE; }
    PTC_D int32_t pop(const int32_t *stack, int stride) {
        if (sp == 0) return TRV_DONE;
        --sp;
        return sp < TRV_SHARED_STACK ? stack[sp * stride] : spill[sp - TRV_SHARED_STACK];
    }
    PTC_D void push(int32_t *stack, int stride, int32_t v) {
        if (sp < TRV_SHARED_STACK)
            stack[sp * stride] = v;
        else if (sp - TRV_SHARED_STACK < TRV_STACK)
            spill[sp - TRV_SHARED_STACK] = v;
        ++sp;
    }

    /* one outer iteration of the while-while loop: descend to the next leaf, test it, pop. Returns done(). */
    template <bool ANY_HIT>
    PTC_D bool advance(const DScene &sc, int32_t *stack) {
        const float4 *__restrict__ nodes = sc.bvhNodes;
        const float4 *__restrict__ tris = sc.tris;
        const int stride = blockDim.x;
        while (node >= 0 && node != TRV_DONE) {
            const float4 n0 = __ldg(&nodes[4 * (size_t)node + 0]);
            const float4 n1 = __ldg(&nodes[4 * (size_t)node + 1]);
            const float4 n2 = __ldg(&nodes[4 * (size_t)node + 2]);
            const float4 n3 = __ldg(&nodes[4 * (size_t)node + 3]);
            /* slab test of both children; the interval is widened by 2 ulp so a boundary hit is never lost (Ize 2013) */
            float l0x = n0.x * idir.x - ood.x, l1x = n0.y * idir.x - ood.x;
            float l0y = n0.z * idir.y - ood.y, l1y = n0.w * idir.y - ood.y;
            float l0z = n2.x * idir.z - ood.z, l1z = n2.y * idir.z - ood.z;
            float lmin = fmaxf(fmaxf(fminf(l0x, l1x), fminf(l0y, l1y)), fmaxf(fminf(l0z, l1z), t0));
            float lmax = fminf(fminf(fmaxf(l0x, l1x), fmaxf(l0y, l1y)), fminf(fmaxf(l0z, l1z), best.t)) * 1.0000004f;
            float r0x = n1.x * idir.x - ood.x, r1x = n1.y * idir.x - ood.x;
            float r0y = n1.z * idir.y - ood.y, r1y = n1.w * idir.y - ood.y;
            float r0z = n2.z * idir.z - ood.z, r1z = n2.w * idir.z - ood.z;
            float rmin = fmaxf(fmaxf(fminf(r0x, r1x), fminf(r0y, r1y)), fmaxf(fminf(r0z, r1z), t0));
            float rmax = fminf(fminf(fmaxf(r0x, r1x), fmaxf(r0y, r1y)), fminf(fmaxf(r0z, r1z), best.t)) * 1.0000004f;
            const bool hl = lmin * 0.9999996f <= lmax, hr = rmin * 0.9999996f <= rmax;
            const int32_t cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
            if (!hl && !hr) {
                node = pop(stack, stride);
            } else {
                node = hl ? cl : cr;
                if (hl && hr) {
                    int32_t farNode = cr;
                    if (rmin < lmin) {
                        node = cr;
                        farNode = cl;
                    }
                    push(stack, stride, farNode);
                }
            }
        }
        if (node == TRV_DONE) return true;
        /* leaf: one triangle */
        {
            const int32_t pos = ~node;
            const float4 v0 = __ldg(&tris[3 * (size_t)pos + 0]);
            const float4 e1 = __ldg(&tris[3 * (size_t)pos + 1]);
            const float4 e2 = __ldg(&tris[3 * (size_t)pos + 2]);
            float t, u, v;
            if (intersectTri(v0, e1, e2, o, d, t, u, v)) {
                const uint32_t wid = __float_as_uint(e2.w);
                const bool after = t > t0 || (t == t0 && id0 != 0xffffffffu && wid > id0);
                const bool inRange = after && t < tmax && t > tmin;
                if (inRange && (t < best.t || (t == best.t && wid < best.worldId))) {
                    best.t = t;
                    best.u = u;
                    best.v = v;
                    best.pos = pos;
                    best.worldId = wid;
                    if (ANY_HIT) {
                        node = TRV_DONE;
                        return true;
                    }
                }
            }
        }
        node = pop(stack, stride);
        return node == TRV_DONE;
    }
};

/* run-to-completion wrappers (used by the chain kernels and the parity hooks) */
template <bool ANY_HIT>
PTC_D HitRec traverse(const DScene &sc, const Ray &ray, float t0, uint32_t id0, int32_t *stack) {
    Trav tr;
    tr.init(sc, ray, t0, id0);
    while (!tr.done()) tr.advance<ANY_HIT>(sc, stack);
    return tr.best;
}
PTC_D HitRec closestHit(const DScene &sc, const Ray &ray, int32_t *stack) { return traverse<false>(sc, ray, ray.tmin, 0xffffffffu, stack); }
PTC_D HitRec nextHit(const DScene &sc, const Ray &ray, float t0, uint32_t id0, int32_t *stack) { return traverse<false>(sc, ray, t0, id0, stack); }
PTC_D bool occluded(const DScene &sc, const Ray &ray, int32_t *stack) { return traverse<true>(sc, ray, ray.tmin, 0xffffffffu, stack).pos >= 0; }

/* ------------------------------------------------------------------ persistent-warp work distribution */
/* Each warp owns a chunk [pos, end) of the work list, refilled with ONE global atomic per chunk; lanes that need work
 * take consecutive items with a ballot/popc rank.  All 32 lanes must call fetch(). */
struct WarpFeeder {
    uint32_t pos = 0, end = 0;
    bool exhausted = false;
    static constexpr uint32_t CHUNK = 256;

    /* returns the work index for this lane or 0xffffffff */
    PTC_D uint32_t fetch(bool need, uint32_t *__restrict__ counter, uint32_t count) {
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m == 0u) return 0xffffffffu;
        const uint32_t lane = threadIdx.x & 31u;
        if (pos >= end && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(counter, CHUNK);
            base = __shfl_sync(0xffffffffu, base, 0);
            pos = base;
            end = min(base + CHUNK, count);
            if (base >= count) {
                exhausted = true;
                pos = end = 0;
            }
        }
        const uint32_t avail = end - pos;
        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
        const uint32_t n = min((uint32_t)__popc(m), avail);
        const uint32_t idx = (need && rank < avail) ? pos + rank : 0xffffffffu;
        pos += n;
        return idx;
    }
};

}  // namespace trv
