/*
 * Software BVH traversal (B200 has no RT cores; this replaces traceRayEXT / the driver's ray-triangle test,
 * raygen.rgen.glsl:110, lightSampling.glsl:128, next_event_estimation.glsl:21).
 *
 * Hit rule (ours to define): among triangles with tmin < t < tmax the smallest (t, world
 * triangle id); no face culling (VulkanScene.cpp:364-377).  Moeller-Trumbore on (v0, e1, e2), same operation
 * order as oracle/accel.hpp::intersectTri.
 *
 * Node fetches are 128-bit (5 x float4 per 8-wide compressed node), triangle fetches are 3 x float4; the traversal
 * stack lives in shared memory (one 8-byte column per thread) with a local-memory spill for pathological depths.
 *
 * The traversal is a resumable state machine (Trav) so that kernels can run it warp-convergently: persistent warps
 * pull rays from a global counter with warp-aggregated atomics and replace finished rays while the other lanes keep
 * traversing (Aila & Laine 2009 "while-while" with dynamic fetch) instead of letting the warp fragment.
 */
#pragma once
#include "common.cuh"

namespace trv {

#define TRV_STACK 88        /* local-memory spill entries (binary LBVH depth bounds the wide depth: <= 63 + 32 levels) */
#ifndef TRV_SHARED_STACK
#define TRV_SHARED_STACK 8  /* shared-memory entries per thread (8 B each) */
#endif
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
    uint32_t inst;    /* two-level structure only: the instance the hit triangle belongs to */
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

/* per-byte sign extension: every byte becomes 0xff when its top bit is set, else 0x00 (PRMT with replicate-sign selectors) */
PTC_D uint32_t signExtendBytes(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, 0x0000ba98;" : "=r"(r) : "r"(x));
    return r;
}

/* byte j of w as the float 32768 + byte: bits 0x47000000 | byte << 8 (exact), ONE byte permute.  `magic` must hold
 * 0x47000000 and come from kernel-parameter space (DScene::prmtMagic): PRMT takes a single immediate, and it has to be the
 * selector, otherwise every call pays an extra move of the selector into a register (ptxas folds any in-kernel constant). */
template <int J>
PTC_D float byteToFloat(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7504 | (J << 4)));
    return __uint_as_float(r);
}
/* Ordered query state: finds the smallest (t, id) lexicographically greater than (t0, id0) with t < tmax.
 * Closest hit: (t0, id0) = (tmin, 0xffffffff).
 *
 * Traversal of the 8-wide compressed BVH (layout in lbvh.cuh).  The state holds a "node group" (index of the first
 * internal child of the last visited node + an 8-bit hit mask in traversal order + the node's imask) and a "triangle
 * group" (first triangle + 24-bit hit mask); the stack holds postponed node groups only. */
struct Stack {
    uint2 *shared; /* this thread's column: entry k at shared[k * TRV_BLOCK] */
    uint2 *spill;  /* TRV_STACK local-memory entries */
};
/* declares the stack of the calling thread inside a kernel */
#define TRV_DECLARE_STACK(name)                                    \
    __shared__ uint2 name##Shared[TRV_SHARED_STACK * TRV_BLOCK];   \
    uint2 name##Spill[TRV_STACK];                                  \
    const trv::Stack name{name##Shared + threadIdx.x, name##Spill}

/* Two-level structure (DScene::twoLevel, lbvh.cuh::TwoLevel).  The same node / triangle machinery walks the top-level tree in world
 * space and the bottom-level tree of an instance in ITS object space: entering an instance replaces (o, d) by world->object * (o, d)
 * - the direction is not renormalised, so t means the same on both levels and best.t keeps pruning - and leaving restores them.
 * The stack then holds three kinds of entries:
 *   node group       (first child, hits | imask)              y > 0x00ffffff
 *   instance group   (first leaf entry | 0x80000000, mask)    the rest of a top-level leaf whose first instance is being visited
 *   marker           (0xffffffff, 0)                          below it lies the top level: popping it leaves the instance
 * TL is a compile-time switch of the kernels: the single-level instantiations never touch the extra state. */
#define TRV_MARKER_X 0xffffffffu
#define TRV_INSTANCE_FLAG 0x80000000u

struct Trav {
    float3 o, d, idir;
    float tmin, tmax, t0;
    uint32_t id0;
    HitRec best;
    uint2 ng, tg;
    uint32_t octinv4;
    int sp;
    /* two-level state */
    float3 wo, wd;          /* the world-space ray while the traversal is inside an instance */
    uint2 ig;               /* pending instances of the current top-level leaf: (first leaf entry, mask) */
    uint32_t inst, instFirstTri;
    bool top;               /* walking the top level */

    PTC_D void init(const DScene &sc, const Ray &ray, float t0_, uint32_t id0_) {
        o = ray.o;
        d = ray.d;
        tmin = ray.tmin;
        tmax = ray.tmax;
        t0 = t0_;
        id0 = id0_;
        start(sc);
    }
    /* reciprocal direction and octant of the current (o, d) */
    PTC_D void setDirection() {
        const float ooeps = 1e-20f;
        idir = f3(1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x)), 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y)),
                  1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z)));
        /* octant of the direction signs; slot (oct) of every node is visited first */
        const uint32_t oct = (idir.x < 0.0f ? 4u : 0u) | (idir.y < 0.0f ? 2u : 0u) | (idir.z < 0.0f ? 1u : 0u);
        octinv4 = (7u - oct) * 0x01010101u;
    }
    /* (re)starts the query described by o, d, tmin, tmax, t0, id0 */
    PTC_D void start(const DScene &sc) {
        setDirection();
        best.t = tmax;
        best.u = best.v = 0.0f;
        best.pos = -1;
        best.worldId = 0xffffffffu;
        sp = 0;
        tg = make_uint2(0u, 0u);
        /* the root is "child 0 of a virtual node group" whose only hit sits at the top bit */
        ng = sc.nTris == 0 ? make_uint2(0u, 0u) : make_uint2(0u, 0x80000000u);
        if (sc.nTris == 0) sp = -1;
    }
    template <bool TL>
    PTC_D void startLevel(const DScene &sc) {
        start(sc);
        if constexpr (TL) {
            top = true;
            ig = make_uint2(0u, 0u);
            inst = 0xffffffffu;
            instFirstTri = 0u;
            best.inst = 0xffffffffu;
            wo = o;
            wd = d;
        }
    }
    /* top level: visit the instance at leaf entry `entry` - transform the ray into its object space and start at its mesh's root */
    PTC_D void enterInstance(const DScene &sc, uint32_t entry) {
        inst = __ldg(&sc.tlasInst[entry]);
        const DInstance *I = &sc.instances[inst];
        const float *m = I->w2o;
        o = f3(m[0] * wo.x + m[1] * wo.y + m[2] * wo.z + m[3], m[4] * wo.x + m[5] * wo.y + m[6] * wo.z + m[7], m[8] * wo.x + m[9] * wo.y + m[10] * wo.z + m[11]);
        d = f3(m[0] * wd.x + m[1] * wd.y + m[2] * wd.z, m[4] * wd.x + m[5] * wd.y + m[6] * wd.z, m[8] * wd.x + m[9] * wd.y + m[10] * wd.z);
        setDirection();
        instFirstTri = I->firstWorldTri;
        ng = make_uint2(__ldg(&sc.meshRoot[I->mesh]), I->numTriangles ? 0x80000000u : 0u);
        tg = make_uint2(0u, 0u);
        top = false;
    }
    PTC_D void leaveInstance() {
        o = wo;
        d = wd;
        setDirection();
        top = true;
        inst = 0xffffffffu;
    }
    /* Two-level: the lane has no node work (and, on the top level, no pending instance): take the next stack entry.  canLeave = no
     * triangle of the current instance is still waiting for its test (they need the object-space ray). */
    PTC_D void popNext(const Stack &st, bool canLeave) {
        const uint2 e = sp <= TRV_SHARED_STACK ? st.shared[(sp - 1) * TRV_BLOCK] : st.spill[sp - 1 - TRV_SHARED_STACK];
        if (e.x == TRV_MARKER_X && e.y == 0u) {
            if (!canLeave) return;
            --sp;
            leaveInstance();
        } else if (e.y <= 0x00ffffffu) { /* instance group */
            --sp;
            ig = make_uint2(e.x & ~TRV_INSTANCE_FLAG, e.y);
        } else {
            --sp;
            ng = e;
        }
    }
    PTC_D bool done() const { return sp < 0; }
    /* the stack: TRV_SHARED_STACK entries in a shared-memory column, deeper entries in the caller's local spill array
     * (kept OUT of this struct so that the traversal state itself stays in registers) */
    PTC_D uint2 pop(const Stack &st) {
        --sp;
        return sp < TRV_SHARED_STACK ? st.shared[sp * TRV_BLOCK] : st.spill[sp - TRV_SHARED_STACK];
    }
    /* (no bounds test here: ptc_build_accel refuses a tree whose depth - one postponed group per level, plus the marker and the leaf rest
     * of the top level on two levels - does not fit TRV_SHARED_STACK + TRV_STACK entries; lbvh.cuh::checkStackDepth) */
    PTC_D void push(const Stack &st, uint2 v) {
        if (sp < TRV_SHARED_STACK)
            st.shared[sp * TRV_BLOCK] = v;
        else
            st.spill[sp - TRV_SHARED_STACK] = v;
        ++sp;
    }

    /* Visits the nearest pending child node (8 quantised boxes at once): updates the node group and returns the
     * triangle group (first triangle, 24-bit mask) the visit exposes.  Requires ng.y > 0x00ffffff. */
    PTC_D uint2 nodeStep(const DScene &sc, const Stack &stack) {
        const float4 *__restrict__ nodes = sc.bvhNodes;
        uint2 tgOut;
        {
            const uint32_t hits = ng.y;
            const uint32_t bit = 31u - (uint32_t)__clz(hits);
            ng.y &= ~(1u << bit);
            if (ng.y > 0x00ffffffu) push(stack, ng);
            const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xffu);
            const uint32_t rel = __popc(hits & ~(0xffffffffu << slot) & 0xffu);
            const float4 *np = nodes + 5 * (size_t)(ng.x + rel);
            const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            const uint32_t ei = __float_as_uint(n0.w);
            /* plane distance t = q * a + b with a = 2^e / d, b = (p - o) / d.  The byte q is turned into the float 32768 + q
             * by ONE byte permute (0x47000000 | q << 8), so t = (32768 + q) * a + (b - 32768 a).  Rounding of that constant
             * (<= 2^-9 |a|) and of b near the node is covered by moving the near planes down / the far planes up by 1/128 of
             * a quantisation step; rounding of b for far-away nodes (t ~ |b|) by the relative 1e-6 margin on the final test.
             * Near and far planes share a and c, so a flat box (q_near == q_far) can never come out inverted. */
            const float ax = __uint_as_float((ei & 0xffu) << 23) * idir.x, ay = __uint_as_float(((ei >> 8) & 0xffu) << 23) * idir.y,
                        az = __uint_as_float(((ei >> 16) & 0xffu) << 23) * idir.z;
            const float cx = fmaf(-32768.0f, ax, (n0.x - o.x) * idir.x), cy = fmaf(-32768.0f, ay, (n0.y - o.y) * idir.y),
                        cz = fmaf(-32768.0f, az, (n0.z - o.z) * idir.z);
            const float px = fabsf(ax) * 0.0078125f, py = fabsf(ay) * 0.0078125f, pz = fabsf(az) * 0.0078125f;
            const float bnx = cx - px, bny = cy - py, bnz = cz - pz, bfx = cx + px, bfy = cy + py, bfz = cz + pz;
            const float tlo = t0 * 0.999999f, thi = best.t * 1.000001f;
            const uint32_t magic = sc.prmtMagic;
            const uint32_t imask = ei >> 24;
            ng.x = __float_as_uint(n1.x);
            tgOut.x = __float_as_uint(n1.y);
            uint32_t hitmask = 0;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
                const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t innerMask4 = signExtendBytes(isInner4 << 3); /* 0xff in the bytes of internal children */
                const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
                const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t qlx = __float_as_uint(half ? n2.y : n2.x), qly = __float_as_uint(half ? n2.w : n2.z);
                const uint32_t qlz = __float_as_uint(half ? n3.y : n3.x), qhx = __float_as_uint(half ? n3.w : n3.z);
                const uint32_t qhy = __float_as_uint(half ? n4.y : n4.x), qhz = __float_as_uint(half ? n4.w : n4.z);
                const uint32_t nx = idir.x < 0.0f ? qhx : qlx, fx = idir.x < 0.0f ? qlx : qhx;
                const uint32_t ny = idir.y < 0.0f ? qhy : qly, fy = idir.y < 0.0f ? qly : qhy;
                const uint32_t nz = idir.z < 0.0f ? qhz : qlz, fz = idir.z < 0.0f ? qlz : qhz;
                /* Variants of this test that were measured and rejected (profiles/r1_v4_kernel_experiments.log; the code is in the
                 * history at "k_extend experiments recorded"): near / far plane of an axis in one packed FFMA2 (-4.5 % instructions, -0.7 %
                 * speed), bytes converted through fp16 halves on the FMA pipe (+-0), FMNMX3-only interval test with a sign-byte gather and
                 * an unpredicated hit mask (+0.3 %), next-node prefetch (-11 %); round 2 (profiles/r2_k_extend_cache_hints.log): triangle
                 * fetches that bypass L1 (-2.4 %, C4 -8 %), evict-last node fetches (+-0).  The kernel is bound by the dependency chain of its
                 * L1-missing fetches at the occupancy the register file allows, not by the instruction count of either pipe. */
#define TRV_CHILD(J)                                                                                                          \
    {                                                                                                                         \
        const float tnx = fmaf(byteToFloat<J>(nx, magic), ax, bnx), tny = fmaf(byteToFloat<J>(ny, magic), ay, bny),           \
                    tnz = fmaf(byteToFloat<J>(nz, magic), az, bnz);                                                           \
        const float tfx = fmaf(byteToFloat<J>(fx, magic), ax, bfx), tfy = fmaf(byteToFloat<J>(fy, magic), ay, bfy),           \
                    tfz = fmaf(byteToFloat<J>(fz, magic), az, bfz);                                                           \
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tlo));                                                             \
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, thi));                                                             \
        if (tn * 0.999999f <= tf) hitmask |= ((childBits4 >> (8 * J)) & 0xffu) << ((bitIndex4 >> (8 * J)) & 0xffu);           \
    }
                TRV_CHILD(0) TRV_CHILD(1) TRV_CHILD(2) TRV_CHILD(3)
#undef TRV_CHILD
            }
            ng.y = (hitmask & 0xff000000u) | imask;
            tgOut.y = hitmask & 0x00ffffffu;
            /* (measured: starting the next node's fetch here with prefetch.global.L1 - CCTL.PF1 x3 - costs 11 %, 2053 -> 1833 Mseg/s) */
        }
        return tgOut;
    }

    /* tests the triangle at position pos of the traversal order; returns true when it became the best hit */
    template <bool TL = false>
    PTC_D bool triTest(const DScene &sc, int32_t pos) {
        const float4 *__restrict__ tris = sc.tris;
        const float4 v0 = __ldg(&tris[3 * (size_t)pos + 0]);
        const float4 e1 = __ldg(&tris[3 * (size_t)pos + 1]);
        const float4 e2 = __ldg(&tris[3 * (size_t)pos + 2]);
        float t, u, v;
        if (intersectTri(v0, e1, e2, o, d, t, u, v)) {
            /* world triangle id (the tie-break key): stored with the triangle, or instance prefix + primitive on two levels */
            const uint32_t wid = TL ? instFirstTri + __float_as_uint(e1.w) : __float_as_uint(e2.w);
            const bool after = t > t0 || (t == t0 && id0 != 0xffffffffu && wid > id0);
            const bool inRange = after && t < tmax && t > tmin;
            if (inRange && (t < best.t || (t == best.t && wid < best.worldId))) {
                best.t = t;
                best.u = u;
                best.v = v;
                best.pos = pos;
                best.worldId = wid;
                if constexpr (TL) best.inst = inst;
                return true;
            }
        }
        return false;
    }

};

/* ------------------------------------------------------------------ persistent-warp work distribution */
/* Each warp owns a chunk [pos, end) of the work list, refilled with ONE global atomic per chunk; lanes that need work
 * take consecutive items with a ballot/popc rank.  All 32 lanes must call fetch(). */
struct WarpFeeder {
    uint32_t pos = 0, end = 0;
    bool exhausted = false;
/* items a warp takes per global atomic: small enough that the last chunks of a launch balance (measured 32: 2358, 64: 2365, 128: 2353,
 * 256: 2312, 512: 2237 Mseg/s on the bench scene) */
#ifndef TRV_FEED_CHUNK
#define TRV_FEED_CHUNK 64
#endif
    uint32_t chunk = TRV_FEED_CHUNK;
    /* small launches (late bounces) get smaller chunks, so that every resident warp still finds work: about four chunks per warp */
    PTC_D void sizeFor(uint32_t count) {
        const uint32_t warps = gridDim.x * (blockDim.x >> 5);
        chunk = min((uint32_t)TRV_FEED_CHUNK, max(8u, (count / (warps * 4u)) & ~7u));
    }

    /* returns the work index for this lane or 0xffffffff */
    PTC_D uint32_t fetch(bool need, uint32_t *__restrict__ counter, uint32_t count) {
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m == 0u) return 0xffffffffu;
        const uint32_t lane = threadIdx.x & 31u;
        if (pos >= end && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(counter, chunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            pos = base;
            end = min(base + chunk, count);
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
