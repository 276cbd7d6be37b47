/*
 * On-device LBVH build over world-space triangles (replaces the driver BLAS/TLAS builds,
 * src/lib/vengine/vulkan/resources/VulkanAccelerationStructure.cpp:137,258 and VulkanScene.cpp:306-381).
 *
 * B200-first choice: instances are flattened to world space (180 GB of HBM makes even the 50 M triangle
 * configuration a 2.4 GB triangle array), so traversal is single-level with no per-ray transform.
 *
 *   k_flatten   instance x primitive -> world triangle (v0, e1, e2) + bounds      [__f*_rn: no FMA contraction]
 *   k_bounds    scene AABB (order-preserving uint atomics: exact, order independent)
 *   k_morton    30-bit (<= 65 536 triangles) or 63-bit Morton code of the bounds centre
 *   radix sort  stable LSD sort of (code, triangle id) pairs
 *   k_karras    Karras 2012 hierarchy, ties broken by sorted index
 *   k_fit       bottom-up AABB fit with per-node arrival counters
 *   k_emit      traversal nodes (both children's boxes in one 64 B record) + Morton-ordered triangles
 *
 * Every step is bit-exact against the CPU reference build in oracle/accel.hpp (tests/test_lbvh_parity.py).
 */
#pragma once
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

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

__global__ void k_flatten(const ptc_vertex *__restrict__ vertices, const uint32_t *__restrict__ indices,
                          const DInstance *__restrict__ instances, uint32_t nInstances, uint32_t nTris, float4 *__restrict__ triOut,
                          float4 *__restrict__ boundsLo, float4 *__restrict__ boundsHi, uint32_t *__restrict__ sceneBounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = f3(3.4e38f), hi = f3(-3.4e38f);
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
        triOut[3 * (size_t)i + 0] = make_float4(p[0].x, p[0].y, p[0].z, __uint_as_float(a));
        triOut[3 * (size_t)i + 1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(prim));
        triOut[3 * (size_t)i + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
        /* bounds over (v0, v0 + e1, v0 + e2), exactly what the oracle's triBounds does */
        float3 q1 = f3(__fadd_rn(p[0].x, e1.x), __fadd_rn(p[0].y, e1.y), __fadd_rn(p[0].z, e1.z));
        float3 q2 = f3(__fadd_rn(p[0].x, e2.x), __fadd_rn(p[0].y, e2.y), __fadd_rn(p[0].z, e2.z));
        lo = fmin3(p[0], fmin3(q1, q2));
        hi = fmax3(p[0], fmax3(q1, q2));
        boundsLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        boundsHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
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
inline int mortonBitsPerAxis(uint64_t nTris) { return nTris <= 65536ull ? 10 : 21; }

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

/* common-prefix length of sorted keys i and j; equal keys fall back to the index (Karras 2012, section 4) */
PTC_D int delta(const uint64_t *__restrict__ keys, int64_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

/* node numbering: internal 0..n-2, leaf k -> n-1+k */
__global__ void k_karras(const uint64_t *__restrict__ keys, uint32_t n, int32_t *__restrict__ parent, int32_t *__restrict__ left,
                         int32_t *__restrict__ right) {
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

/* traversal records. Internal node i = 4 x float4:
 *   n0 = (L.lo.x, L.hi.x, L.lo.y, L.hi.y)   n1 = (R.lo.x, R.hi.x, R.lo.y, R.hi.y)
 *   n2 = (L.lo.z, L.hi.z, R.lo.z, R.hi.z)   n3 = (bits(childL), bits(childR), 0, 0)
 * child >= 0: internal node index; child < 0: ~(sorted triangle position). */
__global__ void k_emit_nodes(uint32_t n, const int32_t *__restrict__ left, const int32_t *__restrict__ right, const float4 *__restrict__ nodeLo,
                             const float4 *__restrict__ nodeHi, float4 *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    int32_t L = left[i], R = right[i];
    float4 llo = nodeLo[L], lhi = nodeHi[L], rlo = nodeLo[R], rhi = nodeHi[R];
    int32_t cl = L >= (int32_t)(n - 1) ? ~(L - (int32_t)(n - 1)) : L;
    int32_t cr = R >= (int32_t)(n - 1) ? ~(R - (int32_t)(n - 1)) : R;
    out[4 * (size_t)i + 0] = make_float4(llo.x, lhi.x, llo.y, lhi.y);
    out[4 * (size_t)i + 1] = make_float4(rlo.x, rhi.x, rlo.y, rhi.y);
    out[4 * (size_t)i + 2] = make_float4(llo.z, lhi.z, rlo.z, rhi.z);
    out[4 * (size_t)i + 3] = make_float4(__int_as_float(cl), __int_as_float(cr), 0.0f, 0.0f);
}

__global__ void k_gather_tris(uint32_t n, const uint32_t *__restrict__ order, const float4 *__restrict__ in, float4 *__restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t t = order[k];
    float4 a = in[3 * (size_t)t + 0], b = in[3 * (size_t)t + 1], c = in[3 * (size_t)t + 2];
    c.w = __uint_as_float(t); /* world triangle id: the tie-break key of the hit rule */
    out[3 * (size_t)k + 0] = a;
    out[3 * (size_t)k + 1] = b;
    out[3 * (size_t)k + 2] = c;
}

struct Build {
    DBuf<float4> trisUnsorted, trisSorted, triLo, triHi, nodeLo, nodeHi, nodes;
    DBuf<uint64_t> keys, keysSorted;
    DBuf<uint32_t> ids, order, arrivals, sceneBounds;
    DBuf<int32_t> parent, left, right;
    DBuf<uint8_t> sortTemp;
    uint32_t n = 0;
    int bits = 0;

    size_t bytes() const {
        return trisUnsorted.bytes() + trisSorted.bytes() + triLo.bytes() + triHi.bytes() + nodeLo.bytes() + nodeHi.bytes() + nodes.bytes() +
               keys.bytes() + keysSorted.bytes() + ids.bytes() + order.bytes() + arrivals.bytes() + parent.bytes() + left.bytes() +
               right.bytes() + sortTemp.bytes();
    }

    /* returns the number of kernel launches */
    int run(const ptc_vertex *vertices, const uint32_t *indices, const DInstance *instances, uint32_t nInstances, uint32_t nTris,
            cudaStream_t s) {
        n = nTris;
        if (n == 0) return 0;
        const int B = 256;
        const uint32_t G = (n + B - 1) / B;
        size_t nn = 2 * (size_t)n - 1;
        trisUnsorted.alloc(3 * (size_t)n);
        trisSorted.alloc(3 * (size_t)n);
        triLo.alloc(n);
        triHi.alloc(n);
        keys.alloc(n);
        keysSorted.alloc(n);
        ids.alloc(n);
        order.alloc(n);
        parent.alloc(nn);
        left.alloc(nn);
        right.alloc(nn);
        nodeLo.alloc(nn);
        nodeHi.alloc(nn);
        arrivals.alloc(n);
        nodes.alloc(4 * (size_t)(n > 1 ? n - 1 : 1));
        sceneBounds.alloc(6);
        uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        CUDA_TRY(cudaMemcpyAsync(sceneBounds.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemsetAsync(parent.p, 0xff, nn * sizeof(int32_t), s));
        CUDA_TRY(cudaMemsetAsync(left.p, 0xff, nn * sizeof(int32_t), s));
        CUDA_TRY(cudaMemsetAsync(right.p, 0xff, nn * sizeof(int32_t), s));
        CUDA_TRY(cudaMemsetAsync(arrivals.p, 0, n * sizeof(uint32_t), s));
        int launches = 0;
        k_flatten<<<G, B, 0, s>>>(vertices, indices, instances, nInstances, n, trisUnsorted.p, triLo.p, triHi.p, sceneBounds.p);
        launches++;
        bits = mortonBitsPerAxis(n);
        k_morton<<<G, B, 0, s>>>(triLo.p, triHi.p, sceneBounds.p, n, bits, keys.p, ids.p);
        launches++;
        size_t tempBytes = 0;
        int endBit = bits * 3;
        cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys.p, keysSorted.p, ids.p, order.p, (int)n, 0, endBit, s);
        sortTemp.alloc(tempBytes);
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(sortTemp.p, tempBytes, keys.p, keysSorted.p, ids.p, order.p, (int)n, 0, endBit, s));
        launches += (endBit + 7) / 8 * 2 + 1;
        if (n > 1) {
            k_karras<<<(n - 1 + B - 1) / B, B, 0, s>>>(keysSorted.p, n, parent.p, left.p, right.p);
            launches++;
        }
        k_fit<<<G, B, 0, s>>>(n, order.p, triLo.p, triHi.p, parent.p, left.p, right.p, nodeLo.p, nodeHi.p, arrivals.p);
        launches++;
        if (n > 1) {
            k_emit_nodes<<<(n - 1 + B - 1) / B, B, 0, s>>>(n, left.p, right.p, nodeLo.p, nodeHi.p, nodes.p);
            launches++;
        }
        k_gather_tris<<<G, B, 0, s>>>(n, order.p, trisUnsorted.p, trisSorted.p);
        launches++;
        CUDA_TRY(cudaGetLastError());
        return launches;
    }
};

}  // namespace lbvh
