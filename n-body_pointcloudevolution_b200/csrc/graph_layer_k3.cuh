// graph_layer_k3.cuh - first-layer (3 input channels) edge kernels of the shift-invariant graph layer.
//
// With k = 3 the projection is 3 FMAs per output channel: there is nothing to tile and nothing for the tensor pipe,
// the kernels are pure HBM streams.  One thread owns a 4-channel group of one edge row: a row of Q outputs is written
// by Q/4 adjacent lanes as one contiguous 16-byte store each (a warp instruction writes 512 contiguous bytes), the
// gathers Q_col[col[e]] / Q_row[e / M] are 16-byte loads of the same shape, and no shared memory is used, so the SM
// holds its full complement of warps (the tiled kernels are limited to 12-16 by their staging buffers).
#pragma once
#include "nbpc_common.cuh"
#ifndef NBPC_HOST_EMU

#define GLK3_THREADS 256
#define GLK3_UNROLL 4

// e / M for 0 <= e < 2^31 with magic = floor(2^32 / M): the estimate is exact or one short
__device__ __forceinline__ uint32_t glk3_div(uint32_t e, uint32_t M, uint32_t magic) {
    uint32_t qt = __umulhi(e, magic);
    if (e - qt * M >= M) ++qt;
    return qt;
}

//   out[e] = act( E[e] W1 + Q_col[col[e]] + Q_row[e / M] ),   E (c,3), W1 (3,Q)
// block = 256 threads = 256 / (Q/4) edge slots x Q/4 channel groups; a block walks GLK3_UNROLL consecutive slot rows
template <int Q, bool RELU>
__global__ void __launch_bounds__(GLK3_THREADS) glk3_edge_out_kernel(const float *__restrict__ E, const int32_t *__restrict__ col,
                                                                      const float *__restrict__ W1, const float *__restrict__ Q_col,
                                                                      const float *__restrict__ Q_row, uint32_t c, uint32_t M, uint32_t magic,
                                                                      float *__restrict__ out) {
    constexpr int G = Q / 4, EPB = GLK3_THREADS / G;         // channel groups per row, edges per slot row
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    float4 w[3];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) w[kk] = __ldg(reinterpret_cast<const float4 *>(W1 + kk * Q + 4 * g));
    const uint32_t e_base = blockIdx.x * (uint32_t)(EPB * GLK3_UNROLL) + slot;
    float x[GLK3_UNROLL][3];
    float4 qc[GLK3_UNROLL], qr[GLK3_UNROLL];
#pragma unroll
    for (int u = 0; u < GLK3_UNROLL; ++u) {   // all loads of the GLK3_UNROLL rows first: memory-level parallelism
        const uint32_t e = e_base + u * EPB;
        if (e < c) {
            const int cidx = __ldg(&col[e]);
            x[u][0] = __ldg(&E[3 * (size_t)e]); x[u][1] = __ldg(&E[3 * (size_t)e + 1]); x[u][2] = __ldg(&E[3 * (size_t)e + 2]);
            qc[u] = __ldg(reinterpret_cast<const float4 *>(Q_col + (size_t)cidx * Q + 4 * g));
            qr[u] = __ldg(reinterpret_cast<const float4 *>(Q_row + (size_t)glk3_div(e, M, magic) * Q + 4 * g));
        }
    }
#pragma unroll
    for (int u = 0; u < GLK3_UNROLL; ++u) {
        const uint32_t e = e_base + u * EPB;
        if (e < c) {
            // same association as the tiled kernel: ((x0 w0 + x1 w1) + x2 w2) + (Q_col + Q_row)
            float4 o;
            o.x = fmaf(x[u][2], w[2].x, fmaf(x[u][1], w[1].x, x[u][0] * w[0].x)) + (qc[u].x + qr[u].x);
            o.y = fmaf(x[u][2], w[2].y, fmaf(x[u][1], w[1].y, x[u][0] * w[0].y)) + (qc[u].y + qr[u].y);
            o.z = fmaf(x[u][2], w[2].z, fmaf(x[u][1], w[1].z, x[u][0] * w[0].z)) + (qc[u].z + qr[u].z);
            o.w = fmaf(x[u][2], w[2].w, fmaf(x[u][1], w[1].w, x[u][0] * w[0].w)) + (qc[u].w + qr[u].w);
            if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(out + (size_t)e * Q + 4 * g) = o;
        }
    }
}

//   dW1 = E^T dZ  (3 x Q), dZ = dOut [* (H_out > 0)]: per-block partials, reduced over blocks in a fixed order afterwards.
// A thread accumulates its channel group over the edges  slot + i * EPB  of the block's contiguous range (fixed order),
// the EPB slots are then summed by a fixed shared-memory tree => bit-reproducible.
template <int Q, bool RELU>
__global__ void __launch_bounds__(GLK3_THREADS) glk3_edge_dw_kernel(const float *__restrict__ E, const float *__restrict__ dOut,
                                                                     const float *__restrict__ Hout, uint32_t c, uint32_t edges_per_block,
                                                                     float *__restrict__ dW_partial) {
    constexpr int G = Q / 4, EPB = GLK3_THREADS / G;
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const uint32_t e_begin = blockIdx.x * edges_per_block;
    const uint32_t e_end = nbpc_min(e_begin + edges_per_block, c);
    float4 acc[3];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) acc[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t e0 = e_begin + slot; e0 < e_end; e0 += EPB * GLK3_UNROLL) {
        float x[GLK3_UNROLL][3];
        float4 z[GLK3_UNROLL], h[GLK3_UNROLL];
#pragma unroll
        for (int u = 0; u < GLK3_UNROLL; ++u) {
            const uint32_t e = e0 + u * EPB;
            if (e < e_end) {
                x[u][0] = __ldg(&E[3 * (size_t)e]); x[u][1] = __ldg(&E[3 * (size_t)e + 1]); x[u][2] = __ldg(&E[3 * (size_t)e + 2]);
                z[u] = __ldg(reinterpret_cast<const float4 *>(dOut + (size_t)e * Q + 4 * g));
                if (RELU) h[u] = __ldg(reinterpret_cast<const float4 *>(Hout + (size_t)e * Q + 4 * g));
            }
        }
#pragma unroll
        for (int u = 0; u < GLK3_UNROLL; ++u) {
            const uint32_t e = e0 + u * EPB;
            if (e < e_end) {
                float4 v = z[u];
                if (RELU) {
                    v.x = h[u].x > 0.f ? v.x : 0.f; v.y = h[u].y > 0.f ? v.y : 0.f;
                    v.z = h[u].z > 0.f ? v.z : 0.f; v.w = h[u].w > 0.f ? v.w : 0.f;
                }
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    acc[kk].x = fmaf(x[u][kk], v.x, acc[kk].x); acc[kk].y = fmaf(x[u][kk], v.y, acc[kk].y);
                    acc[kk].z = fmaf(x[u][kk], v.z, acc[kk].z); acc[kk].w = fmaf(x[u][kk], v.w, acc[kk].w);
                }
            }
        }
    }
    __shared__ float4 red[3][GLK3_THREADS];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) red[kk][threadIdx.x] = acc[kk];
    __syncthreads();
    for (int stride = EPB / 2; stride >= 1; stride >>= 1) {   // slots s and s + stride, same channel group
        if (slot < stride) {
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
                float4 a = red[kk][threadIdx.x];
                const float4 b = red[kk][threadIdx.x + stride * G];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                red[kk][threadIdx.x] = a;
            }
        }
        __syncthreads();
    }
    if (slot == 0) {
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
            *reinterpret_cast<float4 *>(dW_partial + ((size_t)blockIdx.x * 3 + kk) * Q + 4 * g) = red[kk][threadIdx.x];
    }
}
#endif  // !NBPC_HOST_EMU
