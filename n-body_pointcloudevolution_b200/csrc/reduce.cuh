// reduce.cuh - deterministic (fixed-order, atomic-free) reductions shared by the layer kernels.
#pragma once
#include "nbpc_common.cuh"

#define GL_THREADS 256
#define GL_CUBE_CHUNK 256   // rows per partial in the per-sample (cube / set-mean) reductions

// partial[s][blk][ch] = sum over nodes [blk*CHUNK, ...) of sample s of X[node][ch]
static __global__ void cube_partial_kernel(const float *__restrict__ X, int ch_n, int N, int nblk, int B,
                                    float *__restrict__ partial) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * nblk * ch_n) return;
    const int ch = (int)(t % ch_n);
    const int blk = (int)((t / ch_n) % nblk);
    const int s = (int)(t / ((int64_t)ch_n * nblk));
    const int n0 = blk * GL_CUBE_CHUNK, n1 = nbpc_min(n0 + GL_CUBE_CHUNK, N);
    float acc = 0.f;
    for (int n = n0; n < n1; ++n) acc += X[((int64_t)s * N + n) * ch_n + ch];
    partial[t] = acc;
}

// out[s][ch] = (sum_blk partial[s][blk][ch]) * scale
static __global__ void cube_final_kernel(const float *__restrict__ partial, int ch_n, int nblk, int B, float divisor,
                                  float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * ch_n) return;
    const int s = t / ch_n, ch = t % ch_n;
    float acc = 0.f;
    for (int blk = 0; blk < nblk; ++blk) acc += partial[((int64_t)s * nblk + blk) * ch_n + ch];
    out[t] = acc / divisor;
}

// X^T Y over n rows, deterministic two-level: partial[chunk][kk][qo], then a fixed-order sum.
// X and Y are accessors (`float at(row, channel)`) so masks / mean-subtraction fuse into the read.
struct GlPlain {
    const float *p;
    int ld;
    __device__ __forceinline__ float at(int64_t r, int ch) const { return p[r * ld + ch]; }
};

template <class XAcc, class YAcc>
__global__ void xty_partial_kernel(XAcc X, YAcc Y, int64_t n, int rows_per_chunk, int nchunks, int k, int q,
                                   float *__restrict__ partial) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nchunks * k * q) return;
    const int qo = (int)(t % q);
    const int kk = (int)((t / q) % k);
    const int chunk = (int)(t / ((int64_t)k * q));
    const int64_t r0 = (int64_t)chunk * rows_per_chunk, r1 = nbpc_min(r0 + rows_per_chunk, n);
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc += X.at(r, kk) * Y.at(r, qo);
    partial[t] = acc;
}

static __global__ void xty_final_kernel(const float *__restrict__ partial, int nchunks, int kq, float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kq) return;
    float acc = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) acc += partial[(int64_t)ch * kq + t];
    out[t] = acc;
}

static void xty_plan(int64_t n, int k, int q, int *rows_per_chunk, int *nchunks) {
    int64_t cap = ((int64_t)1 << 24) / ((int64_t)k * q);
    if (cap < 16) cap = 16;
    if (cap > 2048) cap = 2048;
    int64_t nc = (n + 1023) / 1024;
    if (nc < 1) nc = 1;
    if (nc > cap) nc = cap;
    *rows_per_chunk = (int)((n + nc - 1) / nc);
    if (*rows_per_chunk < 1) *rows_per_chunk = 1;
    *nchunks = (int)((n + *rows_per_chunk - 1) / *rows_per_chunk);
    if (*nchunks < 1) *nchunks = 1;
}

template <class XAcc, class YAcc>
static void xty(const char *name, XAcc X, YAcc Y, int64_t n, int k, int q, float *partial, float *out, cudaStream_t stream) {
    (void)name;
    int rpc, nc;
    xty_plan(n, k, q, &rpc, &nc);
    void (*kern)(XAcc, YAcc, int64_t, int, int, int, int, float *) = xty_partial_kernel<XAcc, YAcc>;
    NBPC_LAUNCH_N(name, kern, nbpc_cdiv((int64_t)nc * k * q, GL_THREADS), GL_THREADS, 0, stream, X, Y, n, rpc, nc, k, q, partial);
    NBPC_LAUNCH(xty_final_kernel, nbpc_cdiv(k * q, GL_THREADS), GL_THREADS, 0, stream, partial, nc, k * q, out);
}
