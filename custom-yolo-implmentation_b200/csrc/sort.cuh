// One-CTA-per-image bitonic sort of 64-bit keys (ascending).  Used to order NMS / top-k candidates:
// key = (0xFFFFFFFF - score_bits) << 32 | anchor   =>   score descending, ties -> lowest anchor.
// n is a power of two (pad with ~0 sentinels); up to kSortTile keys are sorted entirely in shared
// memory, longer lists use global-memory steps for strides >= kSortTile.
#pragma once
#include "common.cuh"

namespace yb {

constexpr int kSortThreads = 1024;
constexpr int kSortTile = 16384;                 // 128 KB of dynamic shared memory
constexpr unsigned long long kSentinel = ~0ull;

__device__ __forceinline__ unsigned long long make_score_key(float score, unsigned int anchor) {
    return ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(score)) << 32) | anchor;
}
__device__ __forceinline__ float key_score(unsigned long long k) { return __uint_as_float(0xFFFFFFFFu - (unsigned int)(k >> 32)); }
__device__ __forceinline__ unsigned int key_anchor(unsigned long long k) { return (unsigned int)(k & 0xffffffffull); }

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// bitonic steps j = j_start .. 1 of merge size k on a tile in shared memory; `base` is the global
// index of s[0] (the sort direction depends on the global position).
__device__ __forceinline__ void bitonic_tile_steps(unsigned long long *s, int tile_n, int base, int k, int j_start) {
    for (int j = j_start; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (tile_n >> 1); t += blockDim.x) {
            const int i = 2 * t - (t & (j - 1));
            const int l = i + j;
            const bool asc = (((base + i) & k) == 0);
            const unsigned long long a = s[i], b = s[l];
            if ((a > b) == asc) { s[i] = b; s[l] = a; }
        }
        __syncthreads();
    }
}

// Sorts g[0..n) ascending.  smem must hold min(n, kSortTile) keys.  All threads of the CTA call it.
__device__ __forceinline__ void cta_bitonic_sort(unsigned long long *g, int n, unsigned long long *smem) {
    const int tile_n = n < kSortTile ? n : kSortTile;
    for (int base = 0; base < n; base += tile_n) {
        for (int t = threadIdx.x; t < tile_n; t += blockDim.x) smem[t] = g[base + t];
        __syncthreads();
        for (int k = 2; k <= tile_n; k <<= 1) bitonic_tile_steps(smem, tile_n, base, k, k >> 1);
        for (int t = threadIdx.x; t < tile_n; t += blockDim.x) g[base + t] = smem[t];
        __syncthreads();
    }
    for (int k = tile_n << 1; k <= n; k <<= 1) {
        for (int j = k >> 1; j >= tile_n; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1));
                const int l = i + j;
                const bool asc = ((i & k) == 0);
                const unsigned long long a = g[i], b = g[l];
                if ((a > b) == asc) { g[i] = b; g[l] = a; }
            }
            __syncthreads();
        }
        for (int base = 0; base < n; base += tile_n) {
            for (int t = threadIdx.x; t < tile_n; t += blockDim.x) smem[t] = g[base + t];
            __syncthreads();
            bitonic_tile_steps(smem, tile_n, base, k, tile_n >> 1);
            for (int t = threadIdx.x; t < tile_n; t += blockDim.x) g[base + t] = smem[t];
            __syncthreads();
        }
    }
}

}  // namespace yb
