// scan.cuh - in-place exclusive prefix sum over int32 (cell starts, CSR pointers).
// Three launches: per-chunk sums -> scan of chunk sums (one block) -> per-chunk rescan + base.
#pragma once
#include "nbpc_common.cuh"

#define NBPC_SCAN_THREADS 256
#define NBPC_SCAN_ITEMS 16
#define NBPC_SCAN_CHUNK (NBPC_SCAN_THREADS * NBPC_SCAN_ITEMS)

static inline size_t nbpc_scan_partials_count(int64_t n) { return (size_t)((n + NBPC_SCAN_CHUNK - 1) / NBPC_SCAN_CHUNK) + 1; }

#ifdef NBPC_HOST_EMU
static inline int nbpc_exclusive_scan_i32(int32_t *data, int64_t n, int32_t *partials, cudaStream_t) {
    (void)partials;
    int32_t run = 0;
    for (int64_t i = 0; i < n; ++i) { int32_t v = data[i]; data[i] = run; run += v; }
    return NBPC_OK;
}
#else
__device__ __forceinline__ int nbpc_warp_incl_scan(int v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, *total = block sum
template <int THREADS>
__device__ __forceinline__ int nbpc_block_excl_scan(int v, int *total) {
    __shared__ int warp_sums[THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = nbpc_warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < THREADS / 32) ? warp_sums[lane] : 0;
        int wi = nbpc_warp_incl_scan(w);
        if (lane < THREADS / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == THREADS / 32 - 1) *total = wi;
    }
    __syncthreads();
    int r = incl - v + warp_sums[wid];
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(NBPC_SCAN_THREADS)
nbpc_scan_chunk_sums(const int32_t *__restrict__ data, int64_t n, int32_t *__restrict__ partials) {
    __shared__ int total;
    const int64_t base = (int64_t)blockIdx.x * NBPC_SCAN_CHUNK + (int64_t)threadIdx.x * NBPC_SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int i = 0; i < NBPC_SCAN_ITEMS; ++i)
        if (base + i < n) s += data[base + i];
    (void)nbpc_block_excl_scan<NBPC_SCAN_THREADS>(s, &total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(1024) nbpc_scan_partials(int32_t *__restrict__ partials, int n) {
    __shared__ int total;
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < n) ? partials[i] : 0;
        int carry = carry_s;
        int ex = nbpc_block_excl_scan<1024>(v, &total);
        if (i < n) partials[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(NBPC_SCAN_THREADS)
nbpc_scan_apply(int32_t *__restrict__ data, int64_t n, const int32_t *__restrict__ partials) {
    __shared__ int total;
    const int64_t base = (int64_t)blockIdx.x * NBPC_SCAN_CHUNK + (int64_t)threadIdx.x * NBPC_SCAN_ITEMS;
    int v[NBPC_SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < NBPC_SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        s += v[i];
    }
    int run = nbpc_block_excl_scan<NBPC_SCAN_THREADS>(s, &total) + partials[blockIdx.x];
#pragma unroll
    for (int i = 0; i < NBPC_SCAN_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

// data[0..n) -> exclusive prefix sums, in place.  partials: nbpc_scan_partials_count(n) ints.
static inline int nbpc_exclusive_scan_i32(int32_t *data, int64_t n, int32_t *partials, cudaStream_t stream) {
    if (n <= 0) return NBPC_OK;
    const int nchunks = (int)((n + NBPC_SCAN_CHUNK - 1) / NBPC_SCAN_CHUNK);
    NBPC_LAUNCH(nbpc_scan_chunk_sums, nchunks, NBPC_SCAN_THREADS, 0, stream, data, n, partials);
    NBPC_LAUNCH(nbpc_scan_partials, 1, 1024, 0, stream, partials, nchunks);
    NBPC_LAUNCH(nbpc_scan_apply, nchunks, NBPC_SCAN_THREADS, 0, stream, data, n, partials);
    return nbpc_check_launch("nbpc_exclusive_scan_i32");
}
#endif
