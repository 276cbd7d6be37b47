/*
 * Stable LSD radix sort of (64-bit key, 32-bit value) pairs, written for the LBVH build (Morton code, triangle id) - north star (1)
 * "30/63-bit Morton codes, radix sort".  8-bit digits, three kernels per pass:
 *
 *   k_radix_hist     every block counts the digits of ITS contiguous chunk of the input        -> table[digit][block]
 *   k_radix_scan     one block: exclusive scan of the table in (digit major, block minor) order  = where each block's run of each
 *                    digit starts in the output
 *   k_radix_scatter  every block walks its chunk again in sub-tiles, ranks the keys of a sub-tile STABLY (warp match ranks inside a
 *                    warp step, running counters across the steps of a warp, a scan across the warps, running bases across the
 *                    sub-tiles) and writes key + value to their final places
 *
 * The number of blocks is bounded (RADIX_MAX_BLOCKS), so the scan stays one small block however long the input is: 42.5 M keys are
 * 1024 chunks of 41.5 k keys.  Equal keys keep their input order, which together with ids 0..n-1 as the initial values makes the
 * result the unique ascending (key, id) order - the same the CPU reference build gets from std::stable_sort (oracle/accel.hpp).
 */
#pragma once
#include "common.cuh"

namespace radix {

#define RADIX_THREADS 256
#define RADIX_ITEMS 8                                  /* keys per thread and sub-tile */
#define RADIX_TILE (RADIX_THREADS * RADIX_ITEMS)       /* 2048 keys per sub-tile */
#define RADIX_MAX_BLOCKS 1024

struct Plan {
    uint32_t blocks = 0, chunk = 0; /* chunk: keys per block, a multiple of RADIX_TILE */
};
inline Plan plan(uint32_t n) {
    Plan p;
    const uint32_t tiles = (n + RADIX_TILE - 1) / RADIX_TILE;
    const uint32_t tilesPerBlock = (tiles + RADIX_MAX_BLOCKS - 1) / RADIX_MAX_BLOCKS;
    p.chunk = std::max(1u, tilesPerBlock) * RADIX_TILE;
    p.blocks = (n + p.chunk - 1) / p.chunk;
    return p;
}
inline size_t tableEntries(uint32_t n) { return (size_t)256 * plan(n).blocks; }

__global__ void __launch_bounds__(RADIX_THREADS) k_radix_hist(const uint64_t *__restrict__ keys, uint32_t n, uint32_t chunk, int shift, uint32_t *__restrict__ table) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t begin = blockIdx.x * chunk, end = min(begin + chunk, n);
    /* neighbouring keys of a spatially coherent input share their high digits: one shared-memory atomic per group of equal digits in a
     * warp instead of 32 serialised ones */
    const uint32_t rounded = begin + (end - begin + RADIX_THREADS - 1) / RADIX_THREADS * RADIX_THREADS;
    for (uint32_t i = begin + threadIdx.x; i < rounded; i += RADIX_THREADS) {
        const bool valid = i < end;
        const uint32_t digit = valid ? ((uint32_t)(keys[i] >> shift) & 0xffu) : 0x100u;
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        if (valid && (peers & ((1u << (threadIdx.x & 31u)) - 1u)) == 0u) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
    }
    __syncthreads();
    table[threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];
}

/* exclusive scan of `count` entries in place, one block of 1024 threads: each thread owns a contiguous slice */
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t *__restrict__ table, uint32_t count) {
    __shared__ uint32_t warpSums[32];
    const uint32_t per = (count + 1023u) / 1024u;
    const uint32_t begin = min(threadIdx.x * per, count), end = min(begin + per, count);
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; i++) sum += table[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)(threadIdx.x & 31u) >= o) incl += t;
    }
    if ((threadIdx.x & 31u) == 31u) warpSums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32u) {
        const uint32_t v = warpSums[threadIdx.x];
        uint32_t in2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, in2, o);
            if ((int)threadIdx.x >= o) in2 += t;
        }
        warpSums[threadIdx.x] = in2 - v;
    }
    __syncthreads();
    uint32_t run = warpSums[threadIdx.x >> 5] + incl - sum;
    for (uint32_t i = begin; i < end; i++) {
        const uint32_t v = table[i];
        table[i] = run;
        run += v;
    }
}

#ifndef RADIX_MINBLOCKS
#define RADIX_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(RADIX_THREADS, RADIX_MINBLOCKS) k_radix_scatter(const uint64_t *__restrict__ keysIn, const uint32_t *__restrict__ valsIn, uint32_t n, uint32_t chunk, int shift,
                                                                 const uint32_t *__restrict__ table, uint64_t *__restrict__ keysOut, uint32_t *__restrict__ valsOut) {
    constexpr int WARPS = RADIX_THREADS / 32;
    __shared__ uint32_t digitBase[256];          /* where the next key of each digit goes (global position) */
    __shared__ uint32_t warpCount[WARPS][256];   /* per warp: keys of each digit seen so far in this sub-tile, then their base */
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    digitBase[threadIdx.x] = table[threadIdx.x * gridDim.x + blockIdx.x];
    const uint32_t begin = blockIdx.x * chunk, end = min(begin + chunk, n);
    for (uint32_t tile = begin; tile < end; tile += RADIX_TILE) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) warpCount[w][threadIdx.x] = 0u;
        __syncthreads();
        uint64_t key[RADIX_ITEMS];
        uint32_t val[RADIX_ITEMS], rank[RADIX_ITEMS];
        /* warp `warp` owns the contiguous keys [tile + warp * 32 * ITEMS, + 32 * ITEMS): step i is 32 consecutive keys */
        const uint32_t warpBegin = tile + warp * 32u * RADIX_ITEMS;
#pragma unroll
        for (int i = 0; i < RADIX_ITEMS; i++) {
            const uint32_t idx = warpBegin + (uint32_t)i * 32u + lane;
            const bool valid = idx < end;
            key[i] = valid ? keysIn[idx] : 0ull;
            val[i] = valid ? valsIn[idx] : 0u;
            const uint32_t digit = (uint32_t)(key[i] >> shift) & 0xffu;
            /* lanes holding the same digit (invalid lanes form their own group) */
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? digit : 0x100u);
            const uint32_t below = __popc(peers & ((1u << lane) - 1u));
            uint32_t prior = 0;
            if (valid && below == 0u) { /* group leader: the only writer of this digit's counter in this warp step */
                prior = warpCount[warp][digit];
                warpCount[warp][digit] = prior + (uint32_t)__popc(peers);
            }
            prior = __shfl_sync(0xffffffffu, prior, __ffs(peers) - 1);
            rank[i] = prior + below;
            __syncwarp();
        }
        __syncthreads();
        { /* thread d: turn the per-warp counts of digit d into bases, in warp order, and advance the digit's global base */
            uint32_t run = digitBase[threadIdx.x];
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const uint32_t c = warpCount[w][threadIdx.x];
                warpCount[w][threadIdx.x] = run;
                run += c;
            }
            digitBase[threadIdx.x] = run;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RADIX_ITEMS; i++) {
            const uint32_t idx = warpBegin + (uint32_t)i * 32u + lane;
            if (idx < end) {
                const uint32_t pos = warpCount[warp][(uint32_t)(key[i] >> shift) & 0xffu] + rank[i];
                keysOut[pos] = key[i];
                valsOut[pos] = val[i];
            }
        }
        __syncthreads();
    }
}

/* Sorts n pairs by key bits [0, endBit).  keys / vals and keysAlt / valsAlt are ping-pong buffers of n entries; table holds
 * tableEntries(n) counters.  Returns 0 when the result is in keys / vals and 1 when it is in keysAlt / valsAlt.  launches counts
 * the kernels. */
inline int sortPairs(uint64_t *keys, uint32_t *vals, uint64_t *keysAlt, uint32_t *valsAlt, uint32_t *table, uint32_t n, int endBit, cudaStream_t s,
                     int *launches = nullptr) {
    if (n == 0) return 0;
    const Plan p = plan(n);
    int cur = 0;
    for (int shift = 0; shift < endBit; shift += 8) {
        const uint64_t *kin = cur ? keysAlt : keys;
        const uint32_t *vin = cur ? valsAlt : vals;
        uint64_t *kout = cur ? keys : keysAlt;
        uint32_t *vout = cur ? vals : valsAlt;
        k_radix_hist<<<p.blocks, RADIX_THREADS, 0, s>>>(kin, n, p.chunk, shift, table);
        k_radix_scan<<<1, 1024, 0, s>>>(table, 256u * p.blocks);
        k_radix_scatter<<<p.blocks, RADIX_THREADS, 0, s>>>(kin, vin, n, p.chunk, shift, table, kout, vout);
        if (launches) *launches += 3;
        cur ^= 1;
    }
    return cur;
}

}  // namespace radix
