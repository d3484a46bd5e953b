// graph_layer_k3.cuh - first-layer edge kernels of the shift-invariant graph layer: k = 3 input channels (relative
// positions, graph.py:289-343) and k = 9 / 10 (positions + velocities of both ends (+ redshift), graph.py:245-275).
//
// With k <= 10 the projection is a handful of FMAs per output channel: nothing to tile, nothing for the tensor pipe,
// the kernels are pure HBM streams.  One thread owns a 4-channel group of one edge row: a row of Q outputs is written
// by Q/4 adjacent lanes as one contiguous 16-byte store each (a warp instruction writes 512 contiguous bytes), the
// gathers Q_col[col[e]] / Q_row[e / M] are 16-byte loads of the same shape, and no shared memory is used, so the SM
// holds its full complement of warps (the tiled kernels are limited to 12-16 by their staging buffers).
#pragma once
#include "nbpc_common.cuh"
#ifndef NBPC_HOST_EMU

#define GLK3_THREADS 256
#define GLK3_UNROLL 4
__host__ __device__ constexpr int glk3_unroll(int K) { return K <= 4 ? 4 : 2; }   // rows in flight per thread

// e / M for 0 <= e < 2^31 with magic = floor(2^32 / M): the estimate is exact or one short
__device__ __forceinline__ uint32_t glk3_div(uint32_t e, uint32_t M, uint32_t magic) {
    uint32_t qt = __umulhi(e, magic);
    if (e - qt * M >= M) ++qt;
    return qt;
}

//   out[e] = act( E[e] W1 + Q_col[col[e]] + Q_row[e / M] ),   E (c,K), W1 (K,Q)
// block = 256 threads = 256 / (Q/4) edge slots x Q/4 channel groups; a block walks UNROLL consecutive slot rows
template <int K, int Q, bool RELU>
__global__ void __launch_bounds__(GLK3_THREADS) glk3_edge_out_kernel(const float *__restrict__ E, const int32_t *__restrict__ col,
                                                                      const float *__restrict__ W1, const float *__restrict__ Q_col,
                                                                      const float *__restrict__ Q_row, uint32_t c, uint32_t M, uint32_t magic,
                                                                      float *__restrict__ out) {
    constexpr int G = Q / 4, EPB = GLK3_THREADS / G, U = glk3_unroll(K);   // channel groups per row, edges per slot row
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    float4 w[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) w[kk] = __ldg(reinterpret_cast<const float4 *>(W1 + kk * Q + 4 * g));
    const uint32_t e_base = blockIdx.x * (uint32_t)(EPB * U) + slot;
    float x[U][K];
    float4 qc[U], qr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {   // all loads of the U rows first: memory-level parallelism
        const uint32_t e = e_base + u * EPB;
        if (e < c) {
            const int cidx = __ldg(&col[e]);
#pragma unroll
            for (int kk = 0; kk < K; ++kk) x[u][kk] = __ldg(&E[K * (size_t)e + kk]);
            qc[u] = __ldg(reinterpret_cast<const float4 *>(Q_col + (size_t)cidx * Q + 4 * g));
            qr[u] = __ldg(reinterpret_cast<const float4 *>(Q_row + (size_t)glk3_div(e, M, magic) * Q + 4 * g));
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t e = e_base + u * EPB;
        if (e < c) {
            // (((x0 w0 + x1 w1) + x2 w2) + ...) + (Q_col + Q_row)
            float4 o = make_float4(x[u][0] * w[0].x, x[u][0] * w[0].y, x[u][0] * w[0].z, x[u][0] * w[0].w);
#pragma unroll
            for (int kk = 1; kk < K; ++kk) {
                o.x = fmaf(x[u][kk], w[kk].x, o.x); o.y = fmaf(x[u][kk], w[kk].y, o.y);
                o.z = fmaf(x[u][kk], w[kk].z, o.z); o.w = fmaf(x[u][kk], w[kk].w, o.w);
            }
            o.x += qc[u].x + qr[u].x; o.y += qc[u].y + qr[u].y; o.z += qc[u].z + qr[u].z; o.w += qc[u].w + qr[u].w;
            if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(out + (size_t)e * Q + 4 * g) = o;
        }
    }
}

//   dW1 = E^T dZ  (K x Q), dZ = dOut [* (H_out > 0)]: per-block partials, reduced over blocks in a fixed order afterwards.
// A thread accumulates its channel group over the edges  slot + i * EPB  of the block's contiguous range (fixed order),
// the EPB slots are then summed by a fixed shared-memory tree => bit-reproducible.
template <int K, int Q, bool RELU>
__global__ void __launch_bounds__(GLK3_THREADS) glk3_edge_dw_kernel(const float *__restrict__ E, const float *__restrict__ dOut,
                                                                     const float *__restrict__ Hout, uint32_t c, uint32_t edges_per_block,
                                                                     float *__restrict__ dW_partial) {
    constexpr int G = Q / 4, EPB = GLK3_THREADS / G, U = glk3_unroll(K);
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const uint32_t e_begin = blockIdx.x * edges_per_block;
    const uint32_t e_end = nbpc_min(e_begin + edges_per_block, c);
    float4 acc[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) acc[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t e0 = e_begin + slot; e0 < e_end; e0 += EPB * U) {
        float x[U][K];
        float4 z[U], h[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t e = e0 + u * EPB;
            if (e < e_end) {
#pragma unroll
                for (int kk = 0; kk < K; ++kk) x[u][kk] = __ldg(&E[K * (size_t)e + kk]);
                z[u] = __ldg(reinterpret_cast<const float4 *>(dOut + (size_t)e * Q + 4 * g));
                if (RELU) h[u] = __ldg(reinterpret_cast<const float4 *>(Hout + (size_t)e * Q + 4 * g));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t e = e0 + u * EPB;
            if (e < e_end) {
                float4 v = z[u];
                if (RELU) {
                    v.x = h[u].x > 0.f ? v.x : 0.f; v.y = h[u].y > 0.f ? v.y : 0.f;
                    v.z = h[u].z > 0.f ? v.z : 0.f; v.w = h[u].w > 0.f ? v.w : 0.f;
                }
#pragma unroll
                for (int kk = 0; kk < K; ++kk) {
                    acc[kk].x = fmaf(x[u][kk], v.x, acc[kk].x); acc[kk].y = fmaf(x[u][kk], v.y, acc[kk].y);
                    acc[kk].z = fmaf(x[u][kk], v.z, acc[kk].z); acc[kk].w = fmaf(x[u][kk], v.w, acc[kk].w);
                }
            }
        }
    }
    __shared__ float4 red[GLK3_THREADS];
#pragma unroll 1
    for (int kk = 0; kk < K; ++kk) {   // one channel row at a time: 4 KB of scratch for any K
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j == kk) mine = acc[j];
        red[threadIdx.x] = mine;
        __syncthreads();
        for (int stride = EPB / 2; stride >= 1; stride >>= 1) {   // slots s and s + stride, same channel group
            if (slot < stride) {
                float4 a = red[threadIdx.x];
                const float4 b = red[threadIdx.x + stride * G];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                red[threadIdx.x] = a;
            }
            __syncthreads();
        }
        if (slot == 0) *reinterpret_cast<float4 *>(dW_partial + ((size_t)blockIdx.x * K + kk) * Q + 4 * g) = red[threadIdx.x];
        __syncthreads();
    }
}
#endif  // !NBPC_HOST_EMU
