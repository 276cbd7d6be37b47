/*
 * On-device LBVH build over world-space triangles (replaces the driver BLAS/TLAS builds,
 * src/lib/vengine/vulkan/resources/VulkanAccelerationStructure.cpp:137,258 and VulkanScene.cpp:306-381).
 *
 * B200-first choice: instances are flattened to world space (180 GB of HBM makes even the 50 M triangle
 * configuration a 2.4 GB triangle array), so traversal is single-level with no per-ray transform.
 *
 *   k_flatten   instance x primitive -> world triangle (v0, e1, e2) + bounds      [__f*_rn: no FMA contraction]
 *   k_bounds    scene AABB (order-preserving uint atomics: exact, order independent)
 *   k_morton    30-bit (<= 65 536 triangles), 48-bit (<= 2^26) or 63-bit Morton code of the bounds centre
 *   radix sort  stable LSD sort of (code, triangle id) pairs (radix.cuh: hand written, 8-bit digits)
 *   hierarchy   (a) PTC_HIERARCHY_LBVH: k_karras (Karras 2012, ties broken by sorted index) + k_fit (bottom-up AABB fit
 *               with per-node arrival counters);  (b) PTC_HIERARCHY_PLOC (default): parallel locally-ordered clustering
 *               over the same Morton order (Meister & Bittner 2018): every round each cluster finds the neighbour within
 *               +-radius positions that minimises the merged half-area (k_ploc_nn), mutual pairs merge (k_ploc_flags ->
 *               inclusive scan -> k_ploc_apply, which also compacts the cluster list).  Same inputs, 1.3-1.5x fewer node
 *               visits per ray than the Karras tree on the bench scene (profiles/README.md)
 *   collapse    binary LBVH -> 8-wide compressed BVH (80 B nodes: origin, per-axis power-of-two scale, 8-bit
 *               quantised child boxes, octant-ordered child slots; leaves of <= 3 triangles), level by level
 *               (k_wide_select -> inclusive scan -> k_wide_emit) so that node and triangle numbering is
 *               deterministic (breadth first); layout after Ylitie, Karras, Laine 2017
 *   k_gather    triangles in wide-node order
 *
 * The PLOC rounds and the collapse levels are data dependent loops (about 50 rounds / 10 levels): each runs inside ONE cooperative
 * kernel that keeps its counts on the device and separates its phases with grid-wide barriers, so the whole build needs two host
 * round trips (the node bound before the collapse buffers are sized, the final node count) instead of one per round and level.
 *
 * Every step is bit-exact against the CPU reference build in oracle/accel.hpp
 * (tests/test_gpu_parity.py::test_lbvh_bit_exact, ::test_wide_bvh_bit_exact).
 */
#pragma once
#include "common.cuh"
#include "radix.cuh"
#include <cooperative_groups.h>
#include <chrono>
#include <memory>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

namespace lbvh {

PTC_D uint32_t floatFlip(float f) { /* order-preserving float -> uint */
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
PTC_HD float floatUnflip(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

/* world = M * (p, 1) with the fixed order ((m0*x + m1*y) + m2*z) + m3, round-to-nearest, no contraction */
PTC_D float xformRow(const float *r, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], x), __fmul_rn(r[1], y)), __fmul_rn(r[2], z)), r[3]);
}

/* Fixed-size records (WORDS float4 each, one per thread) leave a block through shared memory: written record by record they would be
 * 16-byte stores at a stride of WORDS * 16 bytes (32 partly written sectors per warp instruction); staged, the block writes its
 * contiguous WORDS * 256 words with consecutive threads on consecutive words.  Every thread of the block must call this (barrier);
 * valid = the number of leading threads of the block that hold a record. */
#define RECORD_BLOCK 256
template <int WORDS>
PTC_D void storeRecords(float4 *__restrict__ out, uint32_t firstRecord, uint32_t valid, const float4 (&rec)[WORDS], float4 *stage) {
#pragma unroll
    for (int r = 0; r < WORDS; r++) stage[threadIdx.x * WORDS + r] = rec[r];
    __syncthreads();
    float4 *dst = out + (size_t)WORDS * firstRecord;
    const uint32_t total = (uint32_t)WORDS * valid;
    for (uint32_t j = threadIdx.x; j < total; j += RECORD_BLOCK) dst[j] = stage[j];
}

__global__ void __launch_bounds__(RECORD_BLOCK) k_flatten(const ptc_vertex *__restrict__ vertices, const uint32_t *__restrict__ indices,
                          const DInstance *__restrict__ instances, uint32_t nInstances, uint32_t nTris, float4 *__restrict__ triOut,
                          float4 *__restrict__ boundsLo, float4 *__restrict__ boundsHi, uint32_t *__restrict__ sceneBounds) {
    __shared__ float4 stage[3 * RECORD_BLOCK];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = f3(3.4e38f), hi = f3(-3.4e38f);
    float4 rec[3] = {};
    if (i < nTris) {
        /* binary search of the owning instance in the world-triangle prefix */
        uint32_t a = 0, b = nInstances;
        while (b - a > 1) {
            uint32_t m = (a + b) >> 1;
            if (instances[m].firstWorldTri <= i) a = m; else b = m;
        }
        const DInstance &I = instances[a];
        uint32_t prim = i - I.firstWorldTri;
        const uint32_t *ind = indices + I.firstIndex + 3 * (size_t)prim;
        float3 p[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const ptc_vertex &v = vertices[I.firstVertex + ind[k]];
            float x = v.position[0], y = v.position[1], z = v.position[2];
            p[k] = f3(xformRow(I.m, x, y, z), xformRow(I.m + 4, x, y, z), xformRow(I.m + 8, x, y, z));
        }
        float3 e1 = f3(__fsub_rn(p[1].x, p[0].x), __fsub_rn(p[1].y, p[0].y), __fsub_rn(p[1].z, p[0].z));
        float3 e2 = f3(__fsub_rn(p[2].x, p[0].x), __fsub_rn(p[2].y, p[0].y), __fsub_rn(p[2].z, p[0].z));
        rec[0] = make_float4(p[0].x, p[0].y, p[0].z, __uint_as_float(a));
        rec[1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(prim));
        rec[2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
        /* bounds over (v0, v0 + e1, v0 + e2), exactly what the oracle's triBounds does */
        float3 q1 = f3(__fadd_rn(p[0].x, e1.x), __fadd_rn(p[0].y, e1.y), __fadd_rn(p[0].z, e1.z));
        float3 q2 = f3(__fadd_rn(p[0].x, e2.x), __fadd_rn(p[0].y, e2.y), __fadd_rn(p[0].z, e2.z));
        lo = fmin3(p[0], fmin3(q1, q2));
        hi = fmax3(p[0], fmax3(q1, q2));
        boundsLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        boundsHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
    {
        const uint32_t first = blockIdx.x * RECORD_BLOCK;
        storeRecords<3>(triOut, first, first < nTris ? min((uint32_t)RECORD_BLOCK, nTris - first) : 0u, rec, stage);
    }
    /* warp reduce then one atomic per warp and component */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
        atomicMin(&sceneBounds[0], floatFlip(lo.x));
        atomicMin(&sceneBounds[1], floatFlip(lo.y));
        atomicMin(&sceneBounds[2], floatFlip(lo.z));
        atomicMax(&sceneBounds[3], floatFlip(hi.x));
        atomicMax(&sceneBounds[4], floatFlip(hi.y));
        atomicMax(&sceneBounds[5], floatFlip(hi.z));
    }
}

/* warp reduce of a box, then one atomic per warp and component into the order-preserving scene bounds */
PTC_D void boundsAtomic(float3 lo, float3 hi, uint32_t *__restrict__ sceneBounds) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
        atomicMin(&sceneBounds[0], floatFlip(lo.x));
        atomicMin(&sceneBounds[1], floatFlip(lo.y));
        atomicMin(&sceneBounds[2], floatFlip(lo.z));
        atomicMax(&sceneBounds[3], floatFlip(hi.x));
        atomicMax(&sceneBounds[4], floatFlip(hi.y));
        atomicMax(&sceneBounds[5], floatFlip(hi.z));
    }
}

/* bottom level of the two-level structure: the triangles of ONE mesh in object space, (v0, e1 = v1 - v0, e2 = v2 - v0) + bounds */
__global__ void k_mesh_tris(const ptc_vertex *__restrict__ vertices, const uint32_t *__restrict__ indices, uint32_t firstIndex, uint32_t firstVertex,
                            uint32_t nTris, float4 *__restrict__ triOut, float4 *__restrict__ boundsLo, float4 *__restrict__ boundsHi,
                            uint32_t *__restrict__ sceneBounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = f3(3.4e38f), hi = f3(-3.4e38f);
    if (i < nTris) {
        const uint32_t *ind = indices + firstIndex + 3 * (size_t)i;
        float3 p[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const ptc_vertex &v = vertices[firstVertex + ind[k]];
            p[k] = f3(v.position[0], v.position[1], v.position[2]);
        }
        const float3 e1 = f3(__fsub_rn(p[1].x, p[0].x), __fsub_rn(p[1].y, p[0].y), __fsub_rn(p[1].z, p[0].z));
        const float3 e2 = f3(__fsub_rn(p[2].x, p[0].x), __fsub_rn(p[2].y, p[0].y), __fsub_rn(p[2].z, p[0].z));
        triOut[3 * (size_t)i + 0] = make_float4(p[0].x, p[0].y, p[0].z, 0.0f);
        triOut[3 * (size_t)i + 1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(i));
        triOut[3 * (size_t)i + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
        const float3 q1 = f3(__fadd_rn(p[0].x, e1.x), __fadd_rn(p[0].y, e1.y), __fadd_rn(p[0].z, e1.z));
        const float3 q2 = f3(__fadd_rn(p[0].x, e2.x), __fadd_rn(p[0].y, e2.y), __fadd_rn(p[0].z, e2.z));
        lo = fmin3(p[0], fmin3(q1, q2));
        hi = fmax3(p[0], fmax3(q1, q2));
        boundsLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        boundsHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
    boundsAtomic(lo, hi, sceneBounds);
}

/* top level: the world box of every instance = bounds of the 8 transformed corners of its mesh's object-space box (corner order
 * x fastest, then y, then z; same fixed operation order as the flatten) */
__global__ void k_instance_boxes(const DInstance *__restrict__ instances, uint32_t nInstances, const float4 *__restrict__ meshLo, const float4 *__restrict__ meshHi,
                                 float4 *__restrict__ boundsLo, float4 *__restrict__ boundsHi, uint32_t *__restrict__ sceneBounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = f3(3.4e38f), hi = f3(-3.4e38f);
    if (i < nInstances) {
        const DInstance &I = instances[i];
        const float4 ml = meshLo[I.mesh], mh = meshHi[I.mesh];
        for (int corner = 0; corner < 8; corner++) {
            const float x = (corner & 1) ? mh.x : ml.x, y = (corner & 2) ? mh.y : ml.y, z = (corner & 4) ? mh.z : ml.z;
            const float3 w = f3(xformRow(I.m, x, y, z), xformRow(I.m + 4, x, y, z), xformRow(I.m + 8, x, y, z));
            lo = fmin3(lo, w);
            hi = fmax3(hi, w);
        }
        boundsLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        boundsHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
    boundsAtomic(lo, hi, sceneBounds);
}

PTC_HD uint64_t expandBits21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
PTC_HD uint64_t expandBits10(uint64_t v) {
    v &= 0x3ffull;
    v = (v * 0x00010001ull) & 0xFF0000FFull;
    v = (v * 0x00000101ull) & 0x0F00F00Full;
    v = (v * 0x00000011ull) & 0xC30C30C3ull;
    v = (v * 0x00000005ull) & 0x49249249ull;
    return v;
}
/* 30-bit codes up to 65 536 primitives, 48-bit (16 per axis: cells of 1 / 65 536 of the scene, six 8-bit sort passes) up to 2^26, the
 * full 63 bits beyond */
inline int mortonBitsPerAxis(uint64_t nTris) {
    if (const char *e = getenv("PTC_MORTON_BITS")) { /* tests: force a tier (10, 16 or 21) on both sides */
        const int b = atoi(e);
        if (b == 10 || b == 16 || b == 21) return b;
    }
    return nTris <= 65536ull ? 10 : (nTris <= (1ull << 26) ? 16 : 21);
}

__global__ void k_morton(const float4 *__restrict__ boundsLo, const float4 *__restrict__ boundsHi, const uint32_t *__restrict__ sceneBounds,
                         uint32_t nTris, int bits, uint64_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nTris) return;
    float3 slo = f3(floatUnflip(sceneBounds[0]), floatUnflip(sceneBounds[1]), floatUnflip(sceneBounds[2]));
    float3 shi = f3(floatUnflip(sceneBounds[3]), floatUnflip(sceneBounds[4]), floatUnflip(sceneBounds[5]));
    float ex = __fsub_rn(shi.x, slo.x), ey = __fsub_rn(shi.y, slo.y), ez = __fsub_rn(shi.z, slo.z);
    float ix = ex > 0.0f ? __fdiv_rn(1.0f, ex) : 0.0f, iy = ey > 0.0f ? __fdiv_rn(1.0f, ey) : 0.0f, iz = ez > 0.0f ? __fdiv_rn(1.0f, ez) : 0.0f;
    float4 lo = boundsLo[i], hi = boundsHi[i];
    float cx = __fmul_rn(__fadd_rn(lo.x, hi.x), 0.5f), cy = __fmul_rn(__fadd_rn(lo.y, hi.y), 0.5f), cz = __fmul_rn(__fadd_rn(lo.z, hi.z), 0.5f);
    float scale = (float)(1u << bits), qmax = scale - 1.0f;
    float qx = fminf(fmaxf(__fmul_rn(__fmul_rn(__fsub_rn(cx, slo.x), ix), scale), 0.0f), qmax);
    float qy = fminf(fmaxf(__fmul_rn(__fmul_rn(__fsub_rn(cy, slo.y), iy), scale), 0.0f), qmax);
    float qz = fminf(fmaxf(__fmul_rn(__fmul_rn(__fsub_rn(cz, slo.z), iz), scale), 0.0f), qmax);
    uint64_t x = (uint64_t)(uint32_t)qx, y = (uint64_t)(uint32_t)qy, z = (uint64_t)(uint32_t)qz;
    uint64_t code = bits == 10 ? ((expandBits10(x) << 2) | (expandBits10(y) << 1) | expandBits10(z))
                               : ((expandBits21(x) << 2) | (expandBits21(y) << 1) | expandBits21(z));
    keys[i] = code;
    ids[i] = i;
}

#ifndef WIDE_LEAF_TRIS
#define WIDE_LEAF_TRIS 3 /* triangles per leaf of the wide BVH */
#endif
#ifndef WIDE_PHASES
#define WIDE_PHASES 2 /* 2: free slots are filled by splitting leaves down to single triangles */
#endif

/* common-prefix length of sorted keys i and j; equal keys fall back to the index (Karras 2012, section 4) */
PTC_D int delta(const uint64_t *__restrict__ keys, int64_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

/* node numbering: internal 0..n-2, leaf k -> n-1+k */
__global__ void k_karras(const uint64_t *__restrict__ keys, uint32_t n, int32_t *__restrict__ parent, int32_t *__restrict__ left,
                         int32_t *__restrict__ right, uint32_t *__restrict__ subCount, uint32_t *__restrict__ bigNodes) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int64_t j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int64_t s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int64_t gamma = i + s * d + min(d, 0);
    int64_t lo = min(i, j), hi = max(i, j);
    int32_t L = (lo == gamma) ? (int32_t)(n - 1 + gamma) : (int32_t)gamma;
    int32_t R = (hi == gamma + 1) ? (int32_t)(n - 1 + gamma + 1) : (int32_t)(gamma + 1);
    left[i] = L;
    right[i] = R;
    parent[L] = (int32_t)i;
    parent[R] = (int32_t)i;
    subCount[i] = (uint32_t)(hi - lo + 1); /* node i covers the sorted range [min(i, j), max(i, j)] */
    if (hi - lo + 1 > WIDE_LEAF_TRIS) atomicAdd(bigNodes, 1u); /* upper bound of the wide node count */
}

/* bottom-up fit: the second thread to arrive at a node owns it (fmin/fmax are exact, so order is irrelevant) */
__global__ void k_fit(uint32_t n, const uint32_t *__restrict__ order, const float4 *__restrict__ triLo, const float4 *__restrict__ triHi,
                      const int32_t *__restrict__ parent, const int32_t *__restrict__ left, const int32_t *__restrict__ right,
                      float4 *__restrict__ nodeLo, float4 *__restrict__ nodeHi, uint32_t *__restrict__ arrivals) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t tri = order[k];
    int32_t node = (int32_t)(n - 1 + k);
    nodeLo[node] = triLo[tri];
    nodeHi[node] = triHi[tri];
    if (n == 1) return;
    __threadfence();
    int32_t p = parent[node];
    while (p >= 0) {
        if (atomicAdd(&arrivals[p], 1u) == 0u) return; /* first arrival: the sibling will finish */
        __threadfence();
        int32_t L = left[p], R = right[p];
        float4 a = __ldcg(&nodeLo[L]), b = __ldcg(&nodeLo[R]);
        float4 c = __ldcg(&nodeHi[L]), d = __ldcg(&nodeHi[R]);
        nodeLo[p] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.0f);
        nodeHi[p] = make_float4(fmaxf(c.x, d.x), fmaxf(c.y, d.y), fmaxf(c.z, d.z), 0.0f);
        __threadfence();
        p = parent[p];
    }
}

/* subCount[node] = triangles below a node (leaves: 1) */
PTC_D uint32_t subTris(int32_t node, uint32_t n, const uint32_t *__restrict__ subCount) {
    return node >= (int32_t)(n - 1) ? 1u : subCount[node];
}
/* sorted positions of the (at most WIDE_LEAF_TRIS) triangles below a small subtree, left to right */
PTC_D uint32_t subLeaves(int32_t node, uint32_t n, const int32_t *__restrict__ left, const int32_t *__restrict__ right, uint32_t *out) {
    int32_t st[WIDE_LEAF_TRIS + 1];
    int sp = 0;
    uint32_t k = 0;
    st[sp++] = node;
    while (sp > 0) {
        const int32_t x = st[--sp];
        if (x >= (int32_t)(n - 1)) {
            if (k < WIDE_LEAF_TRIS) out[k] = (uint32_t)(x - (int32_t)(n - 1));
            k++;
        } else {
            st[sp++] = right[x]; /* popped after the left child */
            st[sp++] = left[x];
        }
    }
    return k;
}

/* ------------------------------------------------------------------ PLOC hierarchy
 * Cluster list in Morton order: node id + box per position (double buffered).  Leaves are nodes n-1+k as in the Karras
 * numbering; internal nodes are numbered 0, 1, ... in creation order (deterministic: scan of the merge flags), so the root
 * is node n-2.  Ties in the nearest-neighbour search go to the smaller position. */
PTC_D float mergedHalfArea(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    const float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
    const float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
    const float ex = __fsub_rn(hi.x, lo.x), ey = __fsub_rn(hi.y, lo.y), ez = __fsub_rn(hi.z, lo.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(ex, ey), __fmul_rn(ey, ez)), __fmul_rn(ez, ex));
}

__global__ void k_ploc_init(uint32_t n, const uint32_t *__restrict__ order, const float4 *__restrict__ triLo, const float4 *__restrict__ triHi,
                            float4 *__restrict__ nodeLo, float4 *__restrict__ nodeHi, int32_t *__restrict__ cid, float4 *__restrict__ cLo,
                            float4 *__restrict__ cHi) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t t = order[k];
    const float4 lo = triLo[t], hi = triHi[t];
    nodeLo[n - 1 + k] = lo;
    nodeHi[n - 1 + k] = hi;
    cid[k] = (int32_t)(n - 1 + k);
    cLo[k] = lo;
    cHi[k] = hi;
}

#define PLOC_BLOCK 256
#define PLOC_MAX_RADIUS 32

/* block-wide exclusive scan of two counters per thread (creates / survives, internal children / triangles); returns the block totals */
PTC_D void blockScan2(uint32_t a, uint32_t b, uint32_t &exA, uint32_t &exB, uint32_t &totA, uint32_t &totB, uint32_t *warpA, uint32_t *warpB) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if ((int)lane >= o) ia += ta, ib += tb;
    }
    __syncthreads(); /* the arrays may still be read by the previous call */
    if (lane == 31u) warpA[warp] = ia, warpB[warp] = ib;
    __syncthreads();
    uint32_t baseA = 0, baseB = 0;
    totA = totB = 0;
    for (uint32_t w = 0; w < nWarps; w++) {
        const uint32_t va = warpA[w], vb = warpB[w];
        if (w < warp) baseA += va, baseB += vb;
        totA += va, totB += vb;
    }
    exA = baseA + ia - a;
    exB = baseB + ib - b;
}

/* sums blockSums[0 .. gridDim.x) (two counters each): prefix of the blocks before this one and the grid total, by every block */
PTC_D void gridPrefix2(const uint2 *__restrict__ blockSums, uint32_t &preA, uint32_t &preB, uint32_t &totA, uint32_t &totB, uint32_t *shA, uint32_t *shB) {
    uint32_t pa = 0, pb = 0, ta = 0, tb = 0;
    for (uint32_t k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
        const uint2 v = __ldcg(&blockSums[k]);
        if (k < blockIdx.x) pa += v.x, pb += v.y;
        ta += v.x, tb += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pa += __shfl_xor_sync(0xffffffffu, pa, o), pb += __shfl_xor_sync(0xffffffffu, pb, o);
        ta += __shfl_xor_sync(0xffffffffu, ta, o), tb += __shfl_xor_sync(0xffffffffu, tb, o);
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0u) shA[warp] = pa, shB[warp] = pb, shA[32 + warp] = ta, shB[32 + warp] = tb;
    __syncthreads();
    preA = preB = totA = totB = 0;
    for (uint32_t w = 0; w < nWarps; w++) preA += shA[w], preB += shB[w], totA += shA[32 + w], totB += shB[32 + w];
    __syncthreads();
}

struct PlocResult {
    uint32_t rounds, nodes, failed, pad;
};

/* The whole clustering in one cooperative launch.  Per round: (A) nearest neighbour of every cluster within +-radius positions
 * (shared-memory tile + halo), (B) mutual pairs: position i creates a node when nn[nn[i]] == i and i < nn[i], and survives unless it
 * is the larger position of a mutual pair; every block sums the flags of its contiguous chunk, (C) exclusive prefix over the blocks
 * + scan inside the chunk = node numbers (creation order = position order) and compacted positions; the pairs become nodes.  The
 * cluster count and the node counter live in registers of every thread (all blocks compute the same totals). */
#ifndef PLOC_MINBLOCKS
#define PLOC_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(PLOC_BLOCK, PLOC_MINBLOCKS) k_ploc_all(uint32_t n, int radius, int32_t *cid0, int32_t *cid1, float4 *cLo0, float4 *cLo1, float4 *cHi0, float4 *cHi1,
                                                         uint32_t *__restrict__ nn, int32_t *__restrict__ parent, int32_t *__restrict__ left, int32_t *__restrict__ right,
                                                         uint32_t *__restrict__ subCount, float4 *__restrict__ nodeLo, float4 *__restrict__ nodeHi, uint32_t *__restrict__ bigNodes,
                                                         uint2 *__restrict__ blockSums, PlocResult *__restrict__ result) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ float4 sLo[PLOC_BLOCK + 2 * PLOC_MAX_RADIUS], sHi[PLOC_BLOCK + 2 * PLOC_MAX_RADIUS];
    __shared__ uint32_t shA[64], shB[64];
    uint32_t c = n, nextNode = 0, rounds = 0, failed = 0;
    int cur = 0;
    while (c > 1u) {
        const float4 *__restrict__ cLo = cur ? cLo1 : cLo0, *__restrict__ cHi = cur ? cHi1 : cHi0;
        const int32_t *__restrict__ cid = cur ? cid1 : cid0;
        float4 *__restrict__ cLoOut = cur ? cLo0 : cLo1, *__restrict__ cHiOut = cur ? cHi0 : cHi1;
        int32_t *__restrict__ cidOut = cur ? cid0 : cid1;
        /* (A) */
        const uint32_t tiles = (c + PLOC_BLOCK - 1) / PLOC_BLOCK;
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            /* (positions are below 2^31 - the scene upload refuses more triangles - so 32-bit signed arithmetic is exact) */
            const int32_t base = (int32_t)(tile * PLOC_BLOCK) - radius;
            for (int k = threadIdx.x; k < PLOC_BLOCK + 2 * radius; k += PLOC_BLOCK) {
                const int32_t g = base + k;
                if (g >= 0 && g < (int32_t)c) {
                    sLo[k] = __ldcg(&cLo[g]);
                    sHi[k] = __ldcg(&cHi[g]);
                }
            }
            __syncthreads();
            const int32_t i = (int32_t)(tile * PLOC_BLOCK + threadIdx.x);
            if (i < (int32_t)c) {
                const float4 lo = sLo[threadIdx.x + radius], hi = sHi[threadIdx.x + radius];
                float bestA = 0.0f;
                int32_t best = -1;
                /* the window clipped to the array once, instead of two range tests per candidate */
                const int djLo = max(-radius, -i), djHi = min(radius, (int32_t)c - 1 - i);
                for (int dj = djLo; dj <= djHi; dj++) {
                    if (dj == 0) continue;
                    const float a = mergedHalfArea(lo, hi, sLo[threadIdx.x + radius + dj], sHi[threadIdx.x + radius + dj]);
                    if (best < 0 || a < bestA) { /* ascending j and strict <: the smallest position wins ties */
                        bestA = a;
                        best = i + dj;
                    }
                }
                nn[i] = (uint32_t)best;
            }
            __syncthreads();
        }
        grid.sync();
        /* (B) chunk sums; block b owns positions [b * chunk, (b + 1) * chunk) */
        const uint32_t chunk = ((c + gridDim.x - 1) / gridDim.x + PLOC_BLOCK - 1) / PLOC_BLOCK * PLOC_BLOCK;
        const uint32_t begin = min(blockIdx.x * chunk, c), end = min(begin + chunk, c);
        {
            uint32_t creates = 0, survives = 0;
            for (uint32_t i = begin + threadIdx.x; i < end; i += PLOC_BLOCK) {
                const uint32_t j = __ldcg(&nn[i]);
                const bool mutual = __ldcg(&nn[j]) == i;
                creates += (mutual && i < j) ? 1u : 0u;
                survives += (mutual && i > j) ? 0u : 1u;
            }
            uint32_t ea, eb, ta, tb;
            blockScan2(creates, survives, ea, eb, ta, tb, shA, shB);
            if (threadIdx.x == 0) blockSums[blockIdx.x] = make_uint2(ta, tb);
        }
        grid.sync();
        /* (C) */
        uint32_t preCreate, preSurvive, totCreate, totSurvive;
        gridPrefix2(blockSums, preCreate, preSurvive, totCreate, totSurvive, shA, shB);
        for (uint32_t piece = begin; piece < end; piece += PLOC_BLOCK) {
            const uint32_t i = piece + threadIdx.x;
            uint32_t j = 0, creates = 0, survives = 0;
            if (i < end) {
                j = __ldcg(&nn[i]);
                const bool mutual = __ldcg(&nn[j]) == i;
                creates = (mutual && i < j) ? 1u : 0u;
                survives = (mutual && i > j) ? 0u : 1u;
            }
            uint32_t exCreate, exSurvive, tc, ts;
            blockScan2(creates, survives, exCreate, exSurvive, tc, ts, shA, shB);
            if (i < end && survives) {
                const uint32_t pos = preSurvive + exSurvive;
                if (creates) {
                    const int32_t id = (int32_t)(nextNode + preCreate + exCreate);
                    const int32_t L = __ldcg(&cid[i]), R = __ldcg(&cid[j]);
                    const float4 alo = __ldcg(&cLo[i]), ahi = __ldcg(&cHi[i]), blo = __ldcg(&cLo[j]), bhi = __ldcg(&cHi[j]);
                    const float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
                    const float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
                    const uint32_t cnt = (L >= (int32_t)(n - 1) ? 1u : __ldcg(&subCount[L])) + (R >= (int32_t)(n - 1) ? 1u : __ldcg(&subCount[R]));
                    left[id] = L;
                    right[id] = R;
                    parent[L] = id;
                    parent[R] = id;
                    subCount[id] = cnt;
                    nodeLo[id] = lo;
                    nodeHi[id] = hi;
                    if (cnt > WIDE_LEAF_TRIS) atomicAdd(bigNodes, 1u);
                    cidOut[pos] = id;
                    cLoOut[pos] = lo;
                    cHiOut[pos] = hi;
                } else {
                    cidOut[pos] = __ldcg(&cid[i]);
                    cLoOut[pos] = __ldcg(&cLo[i]);
                    cHiOut[pos] = __ldcg(&cHi[i]);
                }
            }
            preCreate += tc;
            preSurvive += ts;
        }
        if (totCreate == 0u) { /* cannot happen (the globally closest pair is always mutual); never spin */
            failed = 1;
            break;
        }
        nextNode += totCreate;
        c = totSurvive;
        cur ^= 1;
        rounds++;
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *result = PlocResult{rounds, nextNode, failed, 0u};
}

/* ------------------------------------------------------------------ collapse to the 8-wide compressed BVH
 * Wide node = 80 B = 5 x 128-bit words (20 x u32):
 *   w0..2  origin p = node box lo (float bits)          w3   ex | ey << 8 | ez << 16 | imask << 24
 *   w4     index of the first internal child            w5   position of the node's first triangle
 *   w6,7   meta[8]: internal child in slot s -> 0x20 | (24 + s); leaf -> (unary triangle count) << 5 | offset; empty 0
 *   w8,9   qlo.x[8]   w10,11 qlo.y[8]   w12,13 qlo.z[8]   w14,15 qhi.x[8]   w16,17 qhi.y[8]   w18,19 qhi.z[8]
 * child box = p + q * 2^(e - 127), conservative.  ex/ey/ez are biased exponent bytes; imask marks the internal slots.
 * Internal children of a node are consecutive (slot order) starting at w4; triangles of its leaf children are
 * consecutive (slot order) starting at w5.  Node 0 is the root; numbering is breadth first.
 *
 * Children of a wide node: start from {binary root}; while fewer than 8, open the child of largest half-area with more
 * than 3 triangles (first wins ties; the opened child is replaced by its left child, the right child is appended);
 * then the same with "more than 1 triangle".  Children with more than 3 triangles become wide nodes, the rest leaves.
 * Slots: greedy assignment maximising dot(child centre - node centre, (+-1, +-1, +-1)_slot), so that the slot equal to
 * the ray's direction-sign octant holds the nearest child. */
struct WideTmp {
    int32_t slotChild[8];
};

PTC_D float halfArea(float4 lo, float4 hi) {
    const float ex = __fsub_rn(hi.x, lo.x), ey = __fsub_rn(hi.y, lo.y), ez = __fsub_rn(hi.z, lo.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(ex, ey), __fmul_rn(ey, ez)), __fmul_rn(ez, ex));
}

/* one thread per wide node of the current level: choose children and slots, count internal children / triangles */
PTC_D uint2 wideSelect(uint32_t n, int32_t root, const int32_t *__restrict__ left, const int32_t *__restrict__ right, const uint32_t *__restrict__ subCount,
                       const float4 *__restrict__ nodeLo, const float4 *__restrict__ nodeHi, WideTmp &t) {
    int32_t list[8];
    uint32_t tris[8];
    float area[8];
    int len = 1;
    list[0] = root;
    tris[0] = subTris(root, n, subCount);
    area[0] = halfArea(nodeLo[root], nodeHi[root]);
    for (int phase = 0; phase < WIDE_PHASES; phase++) {
        const uint32_t thr = phase == 0 ? (uint32_t)WIDE_LEAF_TRIS : 1u;
        while (len < 8) {
            int bi = -1;
            float ba = -1.0f;
            for (int i = 0; i < len; i++)
                if (tris[i] > thr && area[i] > ba) {
                    ba = area[i];
                    bi = i;
                }
            if (bi < 0) break;
            const int32_t c = list[bi], L = left[c], R = right[c];
            list[bi] = L;
            tris[bi] = subTris(L, n, subCount);
            area[bi] = halfArea(nodeLo[L], nodeHi[L]);
            list[len] = R;
            tris[len] = subTris(R, n, subCount);
            area[len] = halfArea(nodeLo[R], nodeHi[R]);
            len++;
        }
    }
    /* slot assignment */
    const float4 rlo = nodeLo[root], rhi = nodeHi[root];
    const float ncx = __fmul_rn(__fadd_rn(rlo.x, rhi.x), 0.5f), ncy = __fmul_rn(__fadd_rn(rlo.y, rhi.y), 0.5f), ncz = __fmul_rn(__fadd_rn(rlo.z, rhi.z), 0.5f);
    float vx[8], vy[8], vz[8];
    for (int i = 0; i < len; i++) {
        const float4 lo = nodeLo[list[i]], hi = nodeHi[list[i]];
        vx[i] = __fsub_rn(__fmul_rn(__fadd_rn(lo.x, hi.x), 0.5f), ncx);
        vy[i] = __fsub_rn(__fmul_rn(__fadd_rn(lo.y, hi.y), 0.5f), ncy);
        vz[i] = __fsub_rn(__fmul_rn(__fadd_rn(lo.z, hi.z), 0.5f), ncz);
    }
    int32_t slotChild[8];
    for (int s = 0; s < 8; s++) slotChild[s] = -1;
    uint32_t childDone = 0, slotDone = 0;
    for (int it = 0; it < len; it++) {
        int bc = -1, bs = -1;
        float bcost = 0.0f;
        for (int c = 0; c < len; c++) {
            if (childDone & (1u << c)) continue;
            for (int sl = 0; sl < 8; sl++) {
                if (slotDone & (1u << sl)) continue;
                const float cost = __fadd_rn(__fadd_rn((sl & 4) ? vx[c] : -vx[c], (sl & 2) ? vy[c] : -vy[c]), (sl & 1) ? vz[c] : -vz[c]);
                if (bc < 0 || cost > bcost) {
                    bcost = cost;
                    bc = c;
                    bs = sl;
                }
            }
        }
        childDone |= 1u << bc;
        slotDone |= 1u << bs;
        slotChild[bs] = list[bc];
    }
    uint32_t nInternal = 0, nTris = 0;
    for (int sl = 0; sl < 8; sl++) {
        const int32_t c = slotChild[sl];
        t.slotChild[sl] = c;
        if (c < 0) continue;
        const uint32_t ct = subTris(c, n, subCount);
        if (ct > (uint32_t)WIDE_LEAF_TRIS) nInternal++; else nTris += ct;
    }
    return make_uint2(nInternal, nTris);
}

/* biased exponent byte e with extent <= 255 * 2^(e - 127) */
PTC_D uint32_t wideExponent(float extent) {
    const float s = __fdiv_rn(extent, 255.0f);
    const uint32_t b = __float_as_uint(s);
    uint32_t e = (b >> 23) & 0xffu;
    if (b & 0x7fffffu) e++;
    if (e < 1u) e = 1u;
    if (e > 253u) e = 253u;
    return e;
}
PTC_D float pow2Biased(uint32_t e) { return __uint_as_float(e << 23); }

/* writes wide node `id` (rooted at binary node `root`, children chosen by wideSelect): its internal children are numbered from
 * childBase (their binary roots go to rootOf), the triangles of its leaf children from triBase */
PTC_D void wideEmit(uint32_t n, uint32_t id, int32_t root, const WideTmp &t, uint32_t childBase, uint32_t triBase, int32_t *__restrict__ rootOf,
                    const int32_t *__restrict__ left, const int32_t *__restrict__ right, const uint32_t *__restrict__ subCount,
                    const float4 *__restrict__ nodeLo, const float4 *__restrict__ nodeHi, uint4 *__restrict__ wide, uint32_t *__restrict__ triMap) {
    const float4 lo = nodeLo[root], hi = nodeHi[root];
    const float plo[3] = {lo.x, lo.y, lo.z}, phi[3] = {hi.x, hi.y, hi.z};
    uint32_t e[3];
    float scale[3], inv[3];
    for (int a = 0; a < 3; a++) {
        const float extent = __fsub_rn(phi[a], plo[a]);
        uint32_t ee = wideExponent(extent);
        while (ee < 253u && __fmul_rn(extent, pow2Biased(254u - ee)) > 255.0f) ee++;
        e[a] = ee;
        scale[a] = pow2Biased(ee);
        inv[a] = pow2Biased(254u - ee);
    }
    uint32_t meta[8], qlo[3][8], qhi[3][8];
    uint32_t imask = 0, rank = 0, off = 0;
    for (int sl = 0; sl < 8; sl++) {
        const int32_t c = t.slotChild[sl];
        if (c < 0) {
            meta[sl] = 0;
            for (int a = 0; a < 3; a++) {
                qlo[a][sl] = 255u;
                qhi[a][sl] = 0u;
            }
            continue;
        }
        const float4 clo4 = nodeLo[c], chi4 = nodeHi[c];
        const float clo[3] = {clo4.x, clo4.y, clo4.z}, chi[3] = {chi4.x, chi4.y, chi4.z};
        for (int a = 0; a < 3; a++) {
            float ql = floorf(__fmul_rn(__fsub_rn(clo[a], plo[a]), inv[a]));
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            while (ql > 0.0f && __fadd_rn(plo[a], __fmul_rn(ql, scale[a])) > clo[a]) ql -= 1.0f;
            float qh = ceilf(__fmul_rn(__fsub_rn(chi[a], plo[a]), inv[a]));
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            while (qh < 255.0f && __fadd_rn(plo[a], __fmul_rn(qh, scale[a])) < chi[a]) qh += 1.0f;
            qlo[a][sl] = (uint32_t)ql;
            qhi[a][sl] = (uint32_t)qh;
        }
        const uint32_t ct = subTris(c, n, subCount);
        if (ct > (uint32_t)WIDE_LEAF_TRIS) {
            meta[sl] = 0x20u | (24u + (uint32_t)sl);
            imask |= 1u << sl;
            rootOf[childBase + rank] = c;
            rank++;
        } else {
            meta[sl] = (((1u << ct) - 1u) << 5) | off;
            uint32_t leaves[WIDE_LEAF_TRIS];
            subLeaves(c, n, left, right, leaves);
            for (uint32_t j = 0; j < ct; j++) triMap[triBase + off + j] = leaves[j];
            off += ct;
        }
    }
    auto pack4 = [](const uint32_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
    uint4 w0, w1, w2, w3, w4;
    w0.x = __float_as_uint(lo.x);
    w0.y = __float_as_uint(lo.y);
    w0.z = __float_as_uint(lo.z);
    w0.w = e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24);
    w1.x = childBase;
    w1.y = triBase;
    w1.z = pack4(meta);
    w1.w = pack4(meta + 4);
    w2 = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    w3 = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    w4 = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    uint4 *out = wide + 5 * (size_t)id;
    out[0] = w0;
    out[1] = w1;
    out[2] = w2;
    out[3] = w3;
    out[4] = w4;
}

struct WideResult {
    uint32_t nWide, nTris, levels, failed;
};
#define WIDE_BLOCK 128
/* The whole collapse in one cooperative launch, level by level (numbering stays breadth first and deterministic).  Per level:
 * (1) every wide node of the level chooses its children (wideSelect); blocks own contiguous chunks and sum their (internal
 * children, triangles); (2) prefix over the blocks + scan inside the chunk = first child index / first triangle position of every
 * node; wideEmit writes the nodes and the roots of the next level. */
#ifndef WIDE_MINBLOCKS
#define WIDE_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(WIDE_BLOCK, WIDE_MINBLOCKS) k_wide_all(uint32_t n, int32_t binaryRoot, uint32_t maxWide, int32_t *__restrict__ rootOf, const int32_t *__restrict__ left,
                                                         const int32_t *__restrict__ right, const uint32_t *__restrict__ subCount, const float4 *__restrict__ nodeLo,
                                                         const float4 *__restrict__ nodeHi, WideTmp *__restrict__ tmp, uint2 *__restrict__ counts, uint4 *__restrict__ wide,
                                                         uint32_t *__restrict__ triMap, uint2 *__restrict__ blockSums, WideResult *__restrict__ result) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t shA[64], shB[64];
    if (blockIdx.x == 0 && threadIdx.x == 0) rootOf[0] = binaryRoot;
    grid.sync();
    uint32_t levelBase = 0, levelCount = 1, triBase = 0, levels = 0, failed = 0;
    while (levelCount > 0u) {
        if ((size_t)levelBase + levelCount > maxWide) {
            failed = 1;
            break;
        }
        const uint32_t chunk = ((levelCount + gridDim.x - 1) / gridDim.x + WIDE_BLOCK - 1) / WIDE_BLOCK * WIDE_BLOCK;
        const uint32_t begin = min(blockIdx.x * chunk, levelCount), end = min(begin + chunk, levelCount);
        {
            uint32_t a = 0, b = 0;
            for (uint32_t k = begin + threadIdx.x; k < end; k += WIDE_BLOCK) {
                WideTmp t;
                const uint2 cnt = wideSelect(n, __ldcg(&rootOf[levelBase + k]), left, right, subCount, nodeLo, nodeHi, t);
                tmp[levelBase + k] = t;
                counts[k] = cnt;
                a += cnt.x;
                b += cnt.y;
            }
            uint32_t ea, eb, ta, tb;
            blockScan2(a, b, ea, eb, ta, tb, shA, shB);
            if (threadIdx.x == 0) blockSums[blockIdx.x] = make_uint2(ta, tb);
        }
        grid.sync();
        uint32_t preInner, preTris, totInner, totTris;
        gridPrefix2(blockSums, preInner, preTris, totInner, totTris, shA, shB);
        const uint32_t nextBase = levelBase + levelCount;
        for (uint32_t piece = begin; piece < end; piece += WIDE_BLOCK) {
            const uint32_t k = piece + threadIdx.x;
            uint2 cnt = make_uint2(0u, 0u);
            if (k < end) cnt = counts[k];
            uint32_t exInner, exTris, ti, tt;
            blockScan2(cnt.x, cnt.y, exInner, exTris, ti, tt, shA, shB);
            if (k < end) {
                const uint32_t id = levelBase + k;
                const WideTmp t = tmp[id];
                wideEmit(n, id, __ldcg(&rootOf[id]), t, nextBase + preInner + exInner, triBase + preTris + exTris, rootOf, left, right, subCount, nodeLo, nodeHi, wide, triMap);
            }
            preInner += ti;
            preTris += tt;
        }
        triBase += totTris;
        levelBase = nextBase;
        levelCount = totInner;
        levels++;
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *result = WideResult{levelBase, triBase, levels, failed};
}

/* triangles in wide-node order: position k holds sorted triangle triMap[k] = world triangle order[triMap[k]] */
__global__ void __launch_bounds__(RECORD_BLOCK) k_gather_tris(uint32_t n, const uint32_t *__restrict__ triMap, const uint32_t *__restrict__ order,
                                                              const float4 *__restrict__ in, float4 *__restrict__ out, uint32_t *__restrict__ wideOrder) {
    __shared__ float4 stage[3 * RECORD_BLOCK];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    float4 rec[3] = {};
    if (k < n) {
        const uint32_t t = order[triMap[k]];
        rec[0] = in[3 * (size_t)t + 0], rec[1] = in[3 * (size_t)t + 1], rec[2] = in[3 * (size_t)t + 2];
        rec[2].w = __uint_as_float(t); /* world triangle id: the tie-break key of the hit rule */
        wideOrder[k] = t;
    }
    const uint32_t first = blockIdx.x * RECORD_BLOCK;
    storeRecords<3>(out, first, first < n ? min((uint32_t)RECORD_BLOCK, n - first) : 0u, rec, stage);
}

/* Shading records, one per triangle in traversal order, 9 x float4 = 144 B, object space (same numbers the reference's
 * closest-hit shaders fetch through InstanceData -> index buffer -> 3 x Vertex, process_hit.glsl:1-17, in one contiguous read):
 *   r0 = (p0, uv0.x) r1 = (p1, uv0.y) r2 = (p2, uv1.x) r3 = (n0, uv1.y) r4 = (n1, uv2.x) r5 = (n2, uv2.y)
 *   r6 = (tangent0, bits(instance)) r7 = (tangent1, bits(primitive)) r8 = (tangent2, 0) */
__global__ void __launch_bounds__(RECORD_BLOCK) k_gather_shading(uint32_t n, const uint32_t *__restrict__ wideOrder, const float4 *__restrict__ trisUnsorted,
                                                                 const ptc_vertex *__restrict__ vertices, const uint32_t *__restrict__ indices,
                                                                 const DInstance *__restrict__ instances, float4 *__restrict__ out) {
    __shared__ float4 stage[9 * RECORD_BLOCK];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    float4 r[9] = {};
    if (k < n) {
        const uint32_t t = wideOrder[k];
        const uint32_t inst = __float_as_uint(trisUnsorted[3 * (size_t)t + 0].w), prim = __float_as_uint(trisUnsorted[3 * (size_t)t + 1].w);
        const DInstance &I = instances[inst];
        const uint32_t *ind = indices + I.firstIndex + 3 * (size_t)prim;
        const ptc_vertex &a = vertices[I.firstVertex + ind[0]], &b = vertices[I.firstVertex + ind[1]], &c = vertices[I.firstVertex + ind[2]];
        r[0] = make_float4(a.position[0], a.position[1], a.position[2], a.uv[0]);
        r[1] = make_float4(b.position[0], b.position[1], b.position[2], a.uv[1]);
        r[2] = make_float4(c.position[0], c.position[1], c.position[2], b.uv[0]);
        r[3] = make_float4(a.normal[0], a.normal[1], a.normal[2], b.uv[1]);
        r[4] = make_float4(b.normal[0], b.normal[1], b.normal[2], c.uv[0]);
        r[5] = make_float4(c.normal[0], c.normal[1], c.normal[2], c.uv[1]);
        r[6] = make_float4(a.tangent[0], a.tangent[1], a.tangent[2], __uint_as_float(inst));
        r[7] = make_float4(b.tangent[0], b.tangent[1], b.tangent[2], __uint_as_float(prim));
        r[8] = make_float4(c.tangent[0], c.tangent[1], c.tangent[2], 0.0f);
    }
    const uint32_t first = blockIdx.x * RECORD_BLOCK;
    storeRecords<9>(out, first, first < n ? min((uint32_t)RECORD_BLOCK, n - first) : 0u, r, stage);
}

/* the traversal stack holds one postponed node group per level of the wide tree (traverse.cuh); its capacity is fixed */
inline void checkStackDepth(uint32_t levels) {
    const uint32_t capacity = 8u + 88u; /* TRV_SHARED_STACK + TRV_STACK */
    if (levels + 4u > capacity) throw CudaError{"acceleration structure too deep for the traversal stack (" + std::to_string(levels) + " levels)"};
}

struct Build {
    DBuf<float4> trisUnsorted, triLo, triHi, nodeLo, nodeHi;
    DBuf<float4> shading;          /* 9 x float4 per triangle, traversal order (k_gather_shading) */
    DBuf<float4> trav;             /* what traversal reads, ONE allocation so that one L2 access-policy window covers it:
                                      [5 x float4 per wide node, breadth first][3 x float4 per triangle, wide-node order] */
    DBuf<uint4> wide;              /* collapse output before compaction (sized by the node bound) */
    DBuf<uint64_t> keys, keysSorted;
    DBuf<uint32_t> ids, order, arrivals, sceneBounds, triMap, wideOrder, bigNodes, radixTable;
    DBuf<int32_t> parent, left, right, rootOf, cid[2];
    DBuf<uint32_t> subCount, nnIdx;
    DBuf<float4> cLo[2], cHi[2];
    DBuf<uint2> blockSums, wideCounts;
    DBuf<PlocResult> plocResult;
    DBuf<WideResult> wideResult;
    int32_t binaryRoot = 0;
    uint32_t hierarchy = PTC_HIERARCHY_PLOC, plocRadius = 16, plocRounds = 0;
    DBuf<WideTmp> wideTmp;
    uint32_t n = 0;
    uint32_t nWide = 0, wideLevels = 0;
    int bits = 0;
    int hostSyncs = 0; /* host round trips of the last build */

    size_t bytes() const {
        return trisUnsorted.bytes() + trav.bytes() + shading.bytes() + triLo.bytes() + triHi.bytes() + nodeLo.bytes() + nodeHi.bytes() + wide.bytes() +
               keys.bytes() + keysSorted.bytes() + ids.bytes() + order.bytes() + arrivals.bytes() + parent.bytes() + left.bytes() +
               right.bytes() + radixTable.bytes() + triMap.bytes() + wideOrder.bytes() + subCount.bytes() + rootOf.bytes() + wideTmp.bytes() +
               wideCounts.bytes() + cid[0].bytes() + cid[1].bytes() + cLo[0].bytes() + cLo[1].bytes() + cHi[0].bytes() + cHi[1].bytes() + nnIdx.bytes();
    }
    size_t traversalBytes() const { return (size_t)nWide * 80 + (size_t)n * 48; }
    const float4 *wideNodes() const { return trav.p; }
    const float4 *sortedTris() const { return trav.p ? trav.p + 5 * (size_t)nWide : nullptr; }

    /* blocks of a cooperative launch: all of them must be resident */
    static int cooperativeGrid(const void *kernel, int block, int capPerSm) {
        int dev = 0, sms = 0, perSm = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, block, 0));
        if (perSm < 1) throw CudaError{"cooperative kernel does not fit on an SM"};
        if (const char *e = getenv("PTC_COOP_BLOCKS")) capPerSm = std::max(1, atoi(e)); /* experiments: resident blocks per SM of the build's cooperative kernels */
        return sms * std::min(perSm, capPerSm);
    }

    /* phase 1: buffers for n primitives, cleared state */
    void begin(uint32_t nPrims, bool needTris, cudaStream_t s) {
        n = nPrims;
        nWide = wideLevels = 0;
        hostSyncs = 0;
        if (n == 0) return;
        size_t nn = 2 * (size_t)n - 1;
        if (needTris) trisUnsorted.alloc(3 * (size_t)n);
        triLo.alloc(n);
        triHi.alloc(n);
        keys.alloc(n);
        keysSorted.alloc(n);
        ids.alloc(n);
        order.alloc(n);
        parent.alloc(nn);
        left.alloc(nn);
        right.alloc(nn);
        subCount.alloc(n);
        nodeLo.alloc(nn);
        nodeHi.alloc(nn);
        arrivals.alloc(n);
        triMap.alloc(n);
        wideOrder.alloc(n);
        sceneBounds.alloc(6);
        bigNodes.alloc(1);
        radixTable.alloc(radix::tableEntries(n));
        uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        CUDA_TRY(cudaMemcpyAsync(sceneBounds.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemsetAsync(parent.p, 0xff, nn * sizeof(int32_t), s));
        CUDA_TRY(cudaMemsetAsync(left.p, 0xff, nn * sizeof(int32_t), s));
        CUDA_TRY(cudaMemsetAsync(right.p, 0xff, nn * sizeof(int32_t), s));
        CUDA_TRY(cudaMemsetAsync(arrivals.p, 0, n * sizeof(uint32_t), s));
        CUDA_TRY(cudaMemsetAsync(bigNodes.p, 0, sizeof(uint32_t), s));
    }

    /* phase 2 (after a prepare kernel has filled triLo / triHi / sceneBounds): Morton codes, sort, hierarchy, collapse.  Leaves the wide
     * nodes in `wide` (nWide of them, indices relative to this tree) and the primitive order in triMap / order.  Two host round trips. */
    int buildFromBounds(cudaStream_t s) {
        if (n == 0) return 0;
        const int B = 256;
        const uint32_t G = (n + B - 1) / B;
        int launches = 0;
        bits = mortonBitsPerAxis(n);
        k_morton<<<G, B, 0, s>>>(triLo.p, triHi.p, sceneBounds.p, n, bits, keys.p, ids.p);
        launches++;
        /* (code, id) pairs in ascending order: the sorted keys must end in keysSorted and the ids in order */
        const int endBit = bits * 3;
        const int passes = (endBit + 7) / 8;
        uint64_t *kA = keys.p, *kB = keysSorted.p;
        uint32_t *vA = ids.p, *vB = order.p;
        if (passes % 2 == 0) { /* an even number of passes ends in the buffers it started from: start from the final ones */
            CUDA_TRY(cudaMemcpyAsync(keysSorted.p, keys.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(order.p, ids.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
            kA = keysSorted.p, kB = keys.p, vA = order.p, vB = ids.p;
        }
        radix::sortPairs(kA, vA, kB, vB, radixTable.p, n, endBit, s, &launches);
        binaryRoot = 0;
        plocRounds = 0;
        const uint32_t blockSumEntries = 148u * 16u;
        blockSums.alloc(blockSumEntries);
        if (hierarchy == PTC_HIERARCHY_LBVH || n == 1) {
            if (n > 1) {
                k_karras<<<(n - 1 + B - 1) / B, B, 0, s>>>(keysSorted.p, n, parent.p, left.p, right.p, subCount.p, bigNodes.p);
                launches++;
            }
            k_fit<<<G, B, 0, s>>>(n, order.p, triLo.p, triHi.p, parent.p, left.p, right.p, nodeLo.p, nodeHi.p, arrivals.p);
            launches++;
        } else {
            int radius = (int)std::min<uint32_t>(std::max<uint32_t>(plocRadius, 1u), PLOC_MAX_RADIUS);
            for (int k = 0; k < 2; k++) {
                cid[k].alloc(n);
                cLo[k].alloc(n);
                cHi[k].alloc(n);
            }
            nnIdx.alloc(n);
            plocResult.alloc(1);
            k_ploc_init<<<G, B, 0, s>>>(n, order.p, triLo.p, triHi.p, nodeLo.p, nodeHi.p, cid[0].p, cLo[0].p, cHi[0].p);
            launches++;
            /* few resident blocks per SM keep the grid barriers cheap; small inputs need no more than one block per tile */
            int grid = cooperativeGrid((const void *)k_ploc_all, PLOC_BLOCK, n > (2u << 20) ? 8 : 4);
            grid = std::max(1, std::min<int>(grid, (int)((n + PLOC_BLOCK - 1) / PLOC_BLOCK)));
            if ((uint32_t)grid > blockSumEntries) grid = (int)blockSumEntries;
            uint32_t nArg = n;
            void *args[] = {&nArg, &radius, &cid[0].p, &cid[1].p, &cLo[0].p, &cLo[1].p, &cHi[0].p, &cHi[1].p, &nnIdx.p, &parent.p, &left.p, &right.p,
                            &subCount.p, &nodeLo.p, &nodeHi.p, &bigNodes.p, &blockSums.p, &plocResult.p};
            CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_ploc_all, dim3(grid), dim3(PLOC_BLOCK), args, 0, s));
            launches++;
            binaryRoot = (int32_t)(n - 2);
        }
        /* ---- first host round trip: the bound of the wide node count sizes the collapse buffers (and the PLOC verdict rides along) */
        uint32_t big = 0;
        PlocResult pr{};
        CUDA_TRY(cudaMemcpyAsync(&big, bigNodes.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (hierarchy == PTC_HIERARCHY_PLOC && n > 1) CUDA_TRY(cudaMemcpyAsync(&pr, plocResult.p, sizeof(pr), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        hostSyncs++;
        if (hierarchy == PTC_HIERARCHY_PLOC && n > 1) {
            if (pr.failed) throw CudaError{"PLOC round without a merge"};
            if (pr.nodes != n - 1) throw CudaError{"PLOC built a wrong number of nodes"};
            plocRounds = pr.rounds;
        }
        const size_t maxWide = (size_t)big + 1; /* every wide node but the root is rooted at a distinct binary node with > 3 triangles */
        wide.alloc(5 * maxWide);
        wideTmp.alloc(maxWide);
        rootOf.alloc(maxWide);
        wideCounts.alloc(maxWide);
        wideResult.alloc(1);
        {
            /* Karras: internal node 0 is the root; PLOC: the last node created; a single triangle is leaf node 0 = n - 1 */
            int grid = cooperativeGrid((const void *)k_wide_all, WIDE_BLOCK, n > (2u << 20) ? 16 : 4);
            grid = std::max(1, std::min<int>(grid, (int)((maxWide + WIDE_BLOCK - 1) / WIDE_BLOCK)));
            if ((uint32_t)grid > blockSumEntries) grid = (int)blockSumEntries;
            uint32_t nArg = n, maxWideArg = (uint32_t)maxWide;
            void *args[] = {&nArg, &binaryRoot, &maxWideArg, &rootOf.p, &left.p, &right.p, &subCount.p, &nodeLo.p, &nodeHi.p, &wideTmp.p, &wideCounts.p,
                            &wide.p, &triMap.p, &blockSums.p, &wideResult.p};
            CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_wide_all, dim3(grid), dim3(WIDE_BLOCK), args, 0, s));
            launches++;
        }
        /* ---- second host round trip: the node count sizes the traversal buffer */
        WideResult wr{};
        CUDA_TRY(cudaMemcpyAsync(&wr, wideResult.p, sizeof(wr), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        hostSyncs++;
        if (wr.failed) throw CudaError{"wide BVH collapse exceeded its node bound"};
        if (wr.nTris != n) throw CudaError{"wide BVH collapse lost triangles"};
        nWide = wr.nWide;
        wideLevels = wr.levels;
        checkStackDepth(wideLevels);
        return launches;
    }

    /* the single-level structure over world-space triangles; returns the number of kernel launches */
    int run(const ptc_vertex *vertices, const uint32_t *indices, const DInstance *instances, uint32_t nInstances, uint32_t nTris,
            cudaStream_t s) {
        const bool verbose = getenv("PTC_VERBOSE") != nullptr;
        auto now = [&] {
            if (verbose) cudaStreamSynchronize(s);
            return std::chrono::steady_clock::now();
        };
        auto msSince = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        const auto tStart = now();
        begin(nTris, true, s);
        if (n == 0) return 0;
        const int B = 256;
        const uint32_t G = (n + B - 1) / B;
        int launches = 0;
        k_flatten<<<G, RECORD_BLOCK, 0, s>>>(vertices, indices, instances, nInstances, n, trisUnsorted.p, triLo.p, triHi.p, sceneBounds.p);
        launches++;
        launches += buildFromBounds(s);
        const auto tCollapsed = now();
        trav.alloc(5 * (size_t)nWide + 3 * (size_t)n);
        CUDA_TRY(cudaMemcpyAsync(trav.p, wide.p, (size_t)nWide * 80, cudaMemcpyDeviceToDevice, s));
        k_gather_tris<<<G, RECORD_BLOCK, 0, s>>>(n, triMap.p, order.p, trisUnsorted.p, trav.p + 5 * (size_t)nWide, wideOrder.p);
        launches++;
        shading.alloc(9 * (size_t)n);
        k_gather_shading<<<G, RECORD_BLOCK, 0, s>>>(n, wideOrder.p, trisUnsorted.p, vertices, indices, instances, shading.p);
        launches++;
        CUDA_TRY(cudaGetLastError());
        if (verbose) {
            const auto tEnd = now();
            fprintf(stderr, "[ptc] build: %u triangles | flatten + morton + sort + hierarchy (%s, %u rounds) + collapse (%u levels, %u wide nodes) %.2f ms | gather %.2f ms | %d host syncs\n",
                    n, hierarchy == PTC_HIERARCHY_PLOC ? "PLOC" : "Karras", plocRounds, wideLevels, nWide, msSince(tStart, tCollapsed), msSince(tCollapsed, tEnd), hostSyncs);
        }
        return launches;
    }
};

/* ------------------------------------------------------------------ two-level structure (VulkanScene.cpp:306-381: one TLAS over instances with
 * 3x4 transforms, one BLAS per mesh in object space).  Chosen for heavily instanced scenes: C4's 1 250 instances of 24 meshes are
 * 42.5 M world triangles (2.5 GB of traversal data that no cache holds) but 0.6 M unique ones (30 MB, L2 resident).
 *   bottom level  per mesh: k_mesh_tris -> the same Morton / sort / hierarchy / collapse as above, over object-space triangles
 *   top level     k_instance_boxes (world box of every instance from its mesh's box) -> the same pipeline over the boxes; a leaf entry
 *                 is an instance
 *   assembly      one traversal buffer: [top-level nodes][nodes of mesh 0][mesh 1]...[triangles of mesh 0][mesh 1]...; child and triangle
 *                 indices of the bottom-level nodes are rebased to absolute positions (k_rebase_nodes), so the traversal needs no
 *                 per-tree offsets
 * Every tree is bit-exact against the oracle's restatement (oracle/accel.hpp, ptc_get_accel_level). */
__global__ void k_rebase_nodes(uint4 *__restrict__ nodes, uint32_t count, uint32_t nodeBase, uint32_t triBase) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint4 w1 = nodes[5 * (size_t)k + 1];
    w1.x += nodeBase;
    w1.y += triBase;
    nodes[5 * (size_t)k + 1] = w1;
}
/* bottom-level triangles in tree order: (v0, .) (e1, primitive) (e2, .) */
__global__ void k_gather_mesh_tris(uint32_t n, const uint32_t *__restrict__ triMap, const uint32_t *__restrict__ order, const float4 *__restrict__ in,
                                   float4 *__restrict__ out, uint32_t *__restrict__ wideOrder) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t t = order[triMap[k]];
    out[3 * (size_t)k + 0] = in[3 * (size_t)t + 0];
    out[3 * (size_t)k + 1] = in[3 * (size_t)t + 1];
    out[3 * (size_t)k + 2] = in[3 * (size_t)t + 2];
    wideOrder[k] = t;
}
/* shading records of a mesh's triangles in tree order (same 144 bytes as k_gather_shading; the instance word is not used: the hit
 * carries the instance) */
__global__ void k_gather_mesh_shading(uint32_t n, const uint32_t *__restrict__ wideOrder, const ptc_vertex *__restrict__ vertices,
                                      const uint32_t *__restrict__ indices, uint32_t firstIndex, uint32_t firstVertex, float4 *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t prim = wideOrder[k];
    const uint32_t *ind = indices + firstIndex + 3 * (size_t)prim;
    const ptc_vertex &a = vertices[firstVertex + ind[0]], &b = vertices[firstVertex + ind[1]], &c = vertices[firstVertex + ind[2]];
    float4 *r = out + 9 * (size_t)k;
    r[0] = make_float4(a.position[0], a.position[1], a.position[2], a.uv[0]);
    r[1] = make_float4(b.position[0], b.position[1], b.position[2], a.uv[1]);
    r[2] = make_float4(c.position[0], c.position[1], c.position[2], b.uv[0]);
    r[3] = make_float4(a.normal[0], a.normal[1], a.normal[2], b.uv[1]);
    r[4] = make_float4(b.normal[0], b.normal[1], b.normal[2], c.uv[0]);
    r[5] = make_float4(c.normal[0], c.normal[1], c.normal[2], c.uv[1]);
    r[6] = make_float4(a.tangent[0], a.tangent[1], a.tangent[2], __uint_as_float(0xffffffffu));
    r[7] = make_float4(b.tangent[0], b.tangent[1], b.tangent[2], __uint_as_float(prim));
    r[8] = make_float4(c.tangent[0], c.tangent[1], c.tangent[2], 0.0f);
}
__global__ void k_tlas_instances(uint32_t n, const uint32_t *__restrict__ triMap, const uint32_t *__restrict__ order, uint32_t *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = order[triMap[k]];
}

struct MeshRange { /* one mesh of the scene description */
    uint32_t firstIndex, triCount, firstVertex;
};

struct TwoLevel {
    struct Tree { /* one built tree, kept until assembly (and for the parity dump) */
        DBuf<uint4> nodes;      /* relative indices */
        DBuf<float4> tris;      /* bottom level only */
        DBuf<float4> shading;   /* bottom level only */
        DBuf<uint32_t> order;   /* tree position -> primitive (triangle of the mesh / instance) */
        uint32_t nNodes = 0, nPrims = 0, nodeBase = 0, triBase = 0, levels = 0;
        float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    };
    Build work; /* scratch of every tree build */
    std::vector<std::unique_ptr<Tree>> blas;
    Tree tlas;
    DBuf<float4> trav, shading, meshLo, meshHi;
    DBuf<uint32_t> tlasInst, meshRoot;
    uint32_t nNodes = 0, nTris = 0;
    int hostSyncs = 0;

    size_t traversalBytes() const { return (size_t)nNodes * 80 + (size_t)nTris * 48; }
    size_t bytes() const { return trav.bytes() + shading.bytes() + tlasInst.bytes() + meshRoot.bytes() + work.bytes(); }
    const float4 *nodes() const { return trav.p; }
    const float4 *tris() const { return trav.p ? trav.p + 5 * (size_t)nNodes : nullptr; }

    void keep(Tree &t, bool bottom, cudaStream_t s) {
        t.nNodes = work.nWide;
        t.nPrims = work.n;
        t.levels = work.wideLevels;
        t.nodes.alloc(5 * (size_t)std::max(1u, t.nNodes));
        CUDA_TRY(cudaMemcpyAsync(t.nodes.p, work.wide.p, (size_t)t.nNodes * 80, cudaMemcpyDeviceToDevice, s));
        t.order.alloc(std::max(1u, t.nPrims));
        uint32_t sb[6];
        CUDA_TRY(cudaMemcpyAsync(sb, work.sceneBounds.p, sizeof(sb), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        for (int a = 0; a < 3; a++) t.lo[a] = floatUnflip(sb[a]), t.hi[a] = floatUnflip(sb[3 + a]);
        (void)bottom;
    }

    int run(const ptc_vertex *vertices, const uint32_t *indices, const DInstance *instances, uint32_t nInstances, const std::vector<MeshRange> &meshes,
            uint32_t hierarchy, uint32_t plocRadius, cudaStream_t s) {
        const int B = 256;
        int launches = 0;
        hostSyncs = 0;
        work.hierarchy = hierarchy;
        work.plocRadius = plocRadius;
        blas.clear();
        std::vector<float4> hLo(meshes.size()), hHi(meshes.size());
        uint32_t nodeTotal = 0, triTotal = 0;
        /* ---- bottom level */
        for (size_t m = 0; m < meshes.size(); m++) {
            blas.emplace_back(new Tree());
            Tree &t = *blas.back();
            const MeshRange &mr = meshes[m];
            work.begin(mr.triCount, true, s);
            if (mr.triCount == 0) {
                hLo[m] = make_float4(0, 0, 0, 0), hHi[m] = make_float4(0, 0, 0, 0);
                continue;
            }
            const uint32_t G = (mr.triCount + B - 1) / B;
            k_mesh_tris<<<G, B, 0, s>>>(vertices, indices, mr.firstIndex, mr.firstVertex, mr.triCount, work.trisUnsorted.p, work.triLo.p, work.triHi.p, work.sceneBounds.p);
            launches += 1 + work.buildFromBounds(s);
            keep(t, true, s);
            t.tris.alloc(3 * (size_t)t.nPrims);
            k_gather_mesh_tris<<<G, B, 0, s>>>(t.nPrims, work.triMap.p, work.order.p, work.trisUnsorted.p, t.tris.p, t.order.p);
            t.shading.alloc(9 * (size_t)t.nPrims);
            k_gather_mesh_shading<<<G, B, 0, s>>>(t.nPrims, t.order.p, vertices, indices, mr.firstIndex, mr.firstVertex, t.shading.p);
            launches += 2;
            hostSyncs += work.hostSyncs + 1;
            hLo[m] = make_float4(t.lo[0], t.lo[1], t.lo[2], 0.0f);
            hHi[m] = make_float4(t.hi[0], t.hi[1], t.hi[2], 0.0f);
            nodeTotal += t.nNodes;
            triTotal += t.nPrims;
        }
        /* ---- top level over the instances' world boxes */
        meshLo.upload(hLo.data(), hLo.size(), s);
        meshHi.upload(hHi.data(), hHi.size(), s);
        work.begin(nInstances, false, s);
        if (nInstances) {
            k_instance_boxes<<<(nInstances + B - 1) / B, B, 0, s>>>(instances, nInstances, meshLo.p, meshHi.p, work.triLo.p, work.triHi.p, work.sceneBounds.p);
            launches += 1 + work.buildFromBounds(s);
            keep(tlas, false, s);
            k_tlas_instances<<<(nInstances + B - 1) / B, B, 0, s>>>(nInstances, work.triMap.p, work.order.p, tlas.order.p);
            launches++;
            hostSyncs += work.hostSyncs + 1;
        } else {
            tlas.nNodes = tlas.nPrims = 0;
        }
        {
            uint32_t deepest = 0;
            for (auto &t : blas) deepest = std::max(deepest, t->levels);
            checkStackDepth(tlas.levels + deepest + 2u); /* + the marker and the rest of the instance leaf */
        }
        /* ---- assembly */
        nNodes = tlas.nNodes + nodeTotal;
        nTris = triTotal;
        trav.alloc(std::max<size_t>(1, 5 * (size_t)nNodes + 3 * (size_t)nTris));
        shading.alloc(std::max<size_t>(1, 9 * (size_t)nTris));
        tlasInst.alloc(std::max(1u, nInstances));
        if (tlas.nNodes) CUDA_TRY(cudaMemcpyAsync(trav.p, tlas.nodes.p, (size_t)tlas.nNodes * 80, cudaMemcpyDeviceToDevice, s));
        if (nInstances) CUDA_TRY(cudaMemcpyAsync(tlasInst.p, tlas.order.p, (size_t)nInstances * 4, cudaMemcpyDeviceToDevice, s));
        std::vector<uint32_t> roots(std::max<size_t>(1, meshes.size()), 0u);
        uint32_t nodeBase = tlas.nNodes, triBase = 0;
        for (size_t m = 0; m < meshes.size(); m++) {
            Tree &t = *blas[m];
            t.nodeBase = nodeBase;
            t.triBase = triBase;
            roots[m] = nodeBase;
            if (t.nNodes) {
                float4 *dstNodes = trav.p + 5 * (size_t)nodeBase;
                CUDA_TRY(cudaMemcpyAsync(dstNodes, t.nodes.p, (size_t)t.nNodes * 80, cudaMemcpyDeviceToDevice, s));
                k_rebase_nodes<<<(t.nNodes + B - 1) / B, B, 0, s>>>((uint4 *)dstNodes, t.nNodes, nodeBase, triBase);
                CUDA_TRY(cudaMemcpyAsync(trav.p + 5 * (size_t)nNodes + 3 * (size_t)triBase, t.tris.p, (size_t)t.nPrims * 48, cudaMemcpyDeviceToDevice, s));
                CUDA_TRY(cudaMemcpyAsync(shading.p + 9 * (size_t)triBase, t.shading.p, (size_t)t.nPrims * 144, cudaMemcpyDeviceToDevice, s));
                launches++;
            }
            nodeBase += t.nNodes;
            triBase += t.nPrims;
        }
        meshRoot.upload(roots.data(), roots.size(), s);
        CUDA_TRY(cudaStreamSynchronize(s)); /* the host vectors die here */
        hostSyncs++;
        CUDA_TRY(cudaGetLastError());
        if (getenv("PTC_VERBOSE"))
            fprintf(stderr, "[ptc] two-level build: %zu meshes (%u triangles, %u nodes), %u instances (%u top-level nodes), traversal set %.1f MB, %d host syncs\n",
                    meshes.size(), nTris, nodeTotal, nInstances, tlas.nNodes, traversalBytes() / 1e6, hostSyncs);
        return launches;
    }
};

}  // namespace lbvh
