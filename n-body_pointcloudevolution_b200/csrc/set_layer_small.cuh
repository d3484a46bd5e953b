// set_layer_small.cuh - CUDA-core kernels for the narrow ends of the set network (nn.py:10-28): the 6-wide input layer, the
// 16- and 3-wide output layers (utils.py:165).  With one side of the (k x q) weight at most 16 wide there is nothing for
// the tensor pipe: the kernels are float4 streams over the WIDE side with the narrow side held in registers.
//   thread = (row slot, 4-channel group of the wide side); a block walks a contiguous row range of ONE sample, so the
//   per-sample mean is a per-block constant; reductions over rows use a fixed order (deterministic).
#pragma once
#include "nbpc_common.cuh"
#ifndef NBPC_HOST_EMU

#define SGS_THREADS 256

// ------------------------------------------------------------------ forward, narrow input (K <= 16): out = act((x - mu) W + B)
template <int K>
__global__ void __launch_bounds__(SGS_THREADS) sgs_fwd_smallk_kernel(const float *__restrict__ X, const float *__restrict__ mu,
                                                                      const float *__restrict__ W, const float *__restrict__ bias, int N,
                                                                      int q, int rows_per_block, int relu, float *__restrict__ out) {
    const int G = q >> 2, slots = SGS_THREADS / G;
    const int g = threadIdx.x % G, slot = threadIdx.x / G, s = blockIdx.y;
    if (slot >= slots) return;
    float4 w[K];
    float m[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
        w[kk] = __ldg(reinterpret_cast<const float4 *>(W + kk * q + 4 * g));
        m[kk] = __ldg(&mu[s * K + kk]);
    }
    const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + 4 * g));
    const int r0 = blockIdx.x * rows_per_block, r1 = nbpc_min(r0 + rows_per_block, N);
    for (int r = r0 + slot; r < r1; r += slots) {
        const int64_t row = (int64_t)s * N + r;
        float4 o = b;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            const float x = __ldg(&X[row * K + kk]) - m[kk];
            o.x = fmaf(x, w[kk].x, o.x); o.y = fmaf(x, w[kk].y, o.y); o.z = fmaf(x, w[kk].z, o.z); o.w = fmaf(x, w[kk].w, o.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4 *>(out + row * q + 4 * g) = o;
    }
}

// ------------------------------------------------------------------ dH = (dZ - mean_s dZ) W^T [* (H > 0)], narrow output (Q <= 16)
template <int Q>
__global__ void __launch_bounds__(SGS_THREADS) sgs_bwd_in_smallq_kernel(const float *__restrict__ dZ, const float *__restrict__ colmean,
                                                                         const float *__restrict__ W, const float *__restrict__ Hmask,
                                                                         int N, int k, int rows_per_block, float *__restrict__ dH) {
    const int G = k >> 2, slots = SGS_THREADS / G;
    const int g = threadIdx.x % G, slot = threadIdx.x / G, s = blockIdx.y;
    if (slot >= slots) return;
    float w[4][Q], cm[Q];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int qo = 0; qo < Q; ++qo) w[j][qo] = __ldg(&W[(4 * g + j) * Q + qo]);
#pragma unroll
    for (int qo = 0; qo < Q; ++qo) cm[qo] = __ldg(&colmean[s * Q + qo]);
    const int r0 = blockIdx.x * rows_per_block, r1 = nbpc_min(r0 + rows_per_block, N);
    for (int r = r0 + slot; r < r1; r += slots) {
        const int64_t row = (int64_t)s * N + r;
        float z[Q];
        if constexpr (Q % 4 == 0) {
#pragma unroll
            for (int j = 0; j < Q / 4; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(dZ + row * Q + 4 * j));
                z[4 * j] = v.x - cm[4 * j]; z[4 * j + 1] = v.y - cm[4 * j + 1]; z[4 * j + 2] = v.z - cm[4 * j + 2]; z[4 * j + 3] = v.w - cm[4 * j + 3];
            }
        } else {
#pragma unroll
            for (int qo = 0; qo < Q; ++qo) z[qo] = __ldg(&dZ[row * Q + qo]) - cm[qo];
        }
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int qo = 0; qo < Q; ++qo)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaf(z[qo], w[j][qo], o[j]);
        if (Hmask) {
            const float4 h = __ldg(reinterpret_cast<const float4 *>(Hmask + row * k + 4 * g));
            o[0] = h.x > 0.f ? o[0] : 0.f; o[1] = h.y > 0.f ? o[1] : 0.f; o[2] = h.z > 0.f ? o[2] : 0.f; o[3] = h.w > 0.f ? o[3] : 0.f;
        }
        *reinterpret_cast<float4 *>(dH + row * k + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------ dW = (X - mu)^T dZ with one narrow side, per-block partials
// NARROW_IS_Q: the wide side is k (X float4 per thread, dZ row of NARROW values broadcast); else the wide side is q.
// partial[(s * gridDim.x + blk)][k][q]; the slots of a block are summed in ascending order by the slot-0 threads.
template <int NARROW, bool NARROW_IS_Q>
__global__ void __launch_bounds__(SGS_THREADS) sgs_xty_narrow_kernel(const float *__restrict__ X, const float *__restrict__ mu,
                                                                      const float *__restrict__ dZ, int N, int k, int q,
                                                                      int rows_per_block, float *__restrict__ partial) {
    __shared__ float4 red[SGS_THREADS];
    const int wide = NARROW_IS_Q ? k : q;
    const int G = wide >> 2, slots = SGS_THREADS / G;
    const int g = threadIdx.x % G, slot = threadIdx.x / G, s = blockIdx.y;
    float4 acc[NARROW];
#pragma unroll
    for (int j = 0; j < NARROW; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int r0 = blockIdx.x * rows_per_block, r1 = nbpc_min(r0 + rows_per_block, N);
    if (slot < slots) {
        if constexpr (NARROW_IS_Q) {
            const float4 m = __ldg(reinterpret_cast<const float4 *>(mu + s * k + 4 * g));
            for (int r = r0 + slot; r < r1; r += slots) {
                const int64_t row = (int64_t)s * N + r;
                float4 x = __ldg(reinterpret_cast<const float4 *>(X + row * k + 4 * g));
                x.x -= m.x; x.y -= m.y; x.z -= m.z; x.w -= m.w;
#pragma unroll
                for (int j = 0; j < NARROW; ++j) {
                    const float z = __ldg(&dZ[row * NARROW + j]);
                    acc[j].x = fmaf(x.x, z, acc[j].x); acc[j].y = fmaf(x.y, z, acc[j].y);
                    acc[j].z = fmaf(x.z, z, acc[j].z); acc[j].w = fmaf(x.w, z, acc[j].w);
                }
            }
        } else {
            float m[NARROW];
#pragma unroll
            for (int j = 0; j < NARROW; ++j) m[j] = __ldg(&mu[s * NARROW + j]);
            for (int r = r0 + slot; r < r1; r += slots) {
                const int64_t row = (int64_t)s * N + r;
                const float4 z = __ldg(reinterpret_cast<const float4 *>(dZ + row * q + 4 * g));
#pragma unroll
                for (int j = 0; j < NARROW; ++j) {
                    const float x = __ldg(&X[row * NARROW + j]) - m[j];
                    acc[j].x = fmaf(x, z.x, acc[j].x); acc[j].y = fmaf(x, z.y, acc[j].y);
                    acc[j].z = fmaf(x, z.z, acc[j].z); acc[j].w = fmaf(x, z.w, acc[j].w);
                }
            }
        }
    }
    float *dst = partial + ((int64_t)s * gridDim.x + blockIdx.x) * (int64_t)k * q;
#pragma unroll 1
    for (int j = 0; j < NARROW; ++j) {
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NARROW; ++i)
            if (i == j) mine = acc[i];
        red[threadIdx.x] = mine;
        __syncthreads();
        if (slot == 0) {
            float4 a = red[g];
            for (int sl = 1; sl < slots; ++sl) {
                const float4 b = red[sl * G + g];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if constexpr (NARROW_IS_Q) {   // acc[j] holds rows 4g..4g+3 of column j
                dst[(4 * g + 0) * q + j] = a.x; dst[(4 * g + 1) * q + j] = a.y; dst[(4 * g + 2) * q + j] = a.z; dst[(4 * g + 3) * q + j] = a.w;
            } else {                       // acc[j] holds columns 4g..4g+3 of row j
                *reinterpret_cast<float4 *>(dst + j * q + 4 * g) = a;
            }
        }
        __syncthreads();
    }
}

// dW[o] = sum over blocks of partial[blk][o], fixed order: a block owns 32 outputs, its 8 slices take the partials
// b = slice, slice + 8, ... (4 loads in flight each) and are combined in ascending slice order
__global__ void __launch_bounds__(256) sgs_sum_partials_kernel(const float *__restrict__ partial, int nparts, int kq,
                                                                float *__restrict__ dW) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (o < kq) {
        int b = slice;
        for (; b + 24 < nparts; b += 32) {
            a0 += __ldg(partial + (int64_t)b * kq + o); a1 += __ldg(partial + (int64_t)(b + 8) * kq + o);
            a2 += __ldg(partial + (int64_t)(b + 16) * kq + o); a3 += __ldg(partial + (int64_t)(b + 24) * kq + o);
        }
        for (; b < nparts; b += 8) a0 += __ldg(partial + (int64_t)b * kq + o);
    }
    red[slice][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (slice == 0 && o < kq) {
        float a = red[0][lane];
#pragma unroll
        for (int j = 1; j < 8; ++j) a += red[j][lane];
        dW[o] = a;
    }
}

// rows per block so that B * blocks-per-sample is about 8 blocks per SM (and the partial buffers stay small)
static int sgs_rows_per_block(int N, int B, int sms) {
    int64_t want = nbpc_max((int64_t)1, (int64_t)sms * 8 / B);
    int64_t rpb = (N + want - 1) / want;
    rpb = (rpb + 63) / 64 * 64;
    return (int)nbpc_max((int64_t)64, rpb);
}
static int sgs_max_blocks(int N, int B, int sms) { return B * nbpc_cdiv(N, sgs_rows_per_block(N, B, sms)); }

#define SGS_NARROW_LIST(X) X(1) X(2) X(3) X(4) X(6) X(8) X(16)
static bool sgs_narrow_ok(int n) {
#define X(V) if (n == V) return true;
    SGS_NARROW_LIST(X)
#undef X
    return false;
}
#endif  // !NBPC_HOST_EMU
