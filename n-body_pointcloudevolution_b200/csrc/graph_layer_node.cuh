// graph_layer_node.cuh - node-level stages of the shift-invariant graph layer (graph.py:394-456), fused.
//
//   forward : pool (+ per-block column sums of P_row)  ->  cube (tiny: P_cube, Cq = P_cube W4 + B)
//             ->  project (Q_col = P_col W2, Q_row = P_row W3 + Cq)                       [-> edge kernel]
//   backward: bwd pool (+ column sums of dQ_row)  ->  cube (tiny: dCq, Gq = dCq W4^T / (N M))
//             ->  grad (G_col = dQ_col W2^T / indeg, G_row = dQ_row W3^T / M + Gq)        [-> edge kernel, X^T Y partials]
//             ->  final (dW1, dW2, dW3 = fixed-order sums of the per-block partials; dW4 = P_cube^T dCq; dB = sum_s dCq)
// The per-sample column sums ride along with the pooling kernels (a fixed shared-memory tree per block, then a fixed
// order over blocks), so nothing re-reads the node tensors and no float atomics are used: results are bit-reproducible.
// The projections are thread-per-node with the node's row in registers and the weights read from shared memory as
// warp-broadcast LDS.128 (one wavefront, 4 FMAs each): FMA-bound instead of shared-memory-bound.
#pragma once
#include "nbpc_common.cuh"
#ifndef NBPC_HOST_EMU

#define GLN_THREADS 256

// nodes a pooling block covers (threads = node slots x channel groups)
__host__ __device__ constexpr int gln_pool_nodes_per_block(int K) { return (K % 4 == 0) ? GLN_THREADS / (K / 4) : GLN_THREADS; }

// fixed tree over the node slots of a block: red[slot * G + g] (float4 per thread), result in slot 0
template <int G>
__device__ __forceinline__ void gln_block_colsum(float4 *red, float4 v) {
    constexpr int SLOTS = GLN_THREADS / G;
    const int slot = threadIdx.x / G;
    red[threadIdx.x] = v;
    __syncthreads();
#pragma unroll
    for (int stride = SLOTS / 2; stride >= 1; stride >>= 1) {
        if (slot < stride) {
            float4 a = red[threadIdx.x];
            const float4 b = red[threadIdx.x + stride * G];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            red[threadIdx.x] = a;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ forward pooling, K % 4 == 0
// thread per (node, 4-channel group); grid (blocks per sample, B).  P_row = mean of the node's M contiguous edge rows,
// P_col = mean over in-edges (CSR transpose, ascending edge id); partial[s][blk][K] = column sums of P_row over the block
// ROWGIVEN: P_row already holds the row means (handed over by the kernel that wrote H, glk3_edge_out_rowpool_kernel): H is
// read once, by the in-edge gather
template <int K, bool ROWGIVEN = false>
__global__ void __launch_bounds__(GLN_THREADS) gln_pool_kernel(const float *__restrict__ H, int M, int N,
                                                                const int32_t *__restrict__ csrT_ptr,
                                                                const int32_t *__restrict__ csrT_edge, float *__restrict__ P_row,
                                                                float *__restrict__ P_col, float *__restrict__ partial, int rev) {
    constexpr int G = K / 4, NPB = GLN_THREADS / G;
    __shared__ float4 red[GLN_THREADS];
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    // rev: walk the edge tensor from its END (the part the producing kernel wrote last is still in L2); results identical
    const int bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x, s = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
    const int local = bx * NPB + slot;
    float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (local < N) {
        const int64_t node = (int64_t)s * N + local;
        if constexpr (ROWGIVEN) {
            pr = *reinterpret_cast<const float4 *>(P_row + node * K + 4 * g);
        } else {
            const float *hr = H + (node * M) * K + 4 * g;
            float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int m = 0; m < M; ++m) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(hr + (int64_t)m * K));
                rs.x += v.x; rs.y += v.y; rs.z += v.z; rs.w += v.w;
            }
            const float fm = (float)M;
            pr = make_float4(rs.x / fm, rs.y / fm, rs.z / fm, rs.w / fm);
            *reinterpret_cast<float4 *>(P_row + node * K + 4 * g) = pr;
        }
        const int b = csrT_ptr[node], e = csrT_ptr[node + 1];
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = b;
        for (; p + 4 <= e; p += 4) {   // 4 independent row gathers in flight
            const int e0 = __ldg(&csrT_edge[p]), e1 = __ldg(&csrT_edge[p + 1]), e2 = __ldg(&csrT_edge[p + 2]), e3 = __ldg(&csrT_edge[p + 3]);
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(H + (int64_t)e0 * K + 4 * g));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(H + (int64_t)e1 * K + 4 * g));
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(H + (int64_t)e2 * K + 4 * g));
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(H + (int64_t)e3 * K + 4 * g));
            cs.x += v0.x; cs.y += v0.y; cs.z += v0.z; cs.w += v0.w;
            cs.x += v1.x; cs.y += v1.y; cs.z += v1.z; cs.w += v1.w;
            cs.x += v2.x; cs.y += v2.y; cs.z += v2.z; cs.w += v2.w;
            cs.x += v3.x; cs.y += v3.y; cs.z += v3.z; cs.w += v3.w;
        }
        for (; p < e; ++p) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(H + (int64_t)__ldg(&csrT_edge[p]) * K + 4 * g));
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
        }
        const float fc = (float)nbpc_max(e - b, 1);
        *reinterpret_cast<float4 *>(P_col + node * K + 4 * g) = make_float4(cs.x / fc, cs.y / fc, cs.z / fc, cs.w / fc);
    }
    gln_block_colsum<G>(red, pr);
    if (slot == 0) *reinterpret_cast<float4 *>(partial + ((int64_t)s * gridDim.x + bx) * K + 4 * g) = red[threadIdx.x];
}

// ------------------------------------------------------------------ forward pooling, runtime K <= 32 (the 3-channel
// input edges; any width that is not a multiple of 4): thread per (node, channel), KP = 2^ceil(log2 K) lanes per node
__global__ void __launch_bounds__(GLN_THREADS) gln_pool_generic_kernel(const float *__restrict__ H, int K, int KP, int M, int N,
                                                                        const int32_t *__restrict__ csrT_ptr,
                                                                        const int32_t *__restrict__ csrT_edge,
                                                                        float *__restrict__ P_row, float *__restrict__ P_col,
                                                                        float *__restrict__ partial) {
    __shared__ float red[GLN_THREADS];
    const int j = threadIdx.x % KP, slot = threadIdx.x / KP, npb = GLN_THREADS / KP;
    const int local = blockIdx.x * npb + slot, s = blockIdx.y;
    float pr = 0.f;
    if (local < N && j < K) {
        const int64_t node = (int64_t)s * N + local;
        const float *hr = H + node * M * K + j;
        float rs = 0.f;
        for (int m = 0; m < M; ++m) rs += __ldg(&hr[(int64_t)m * K]);
        pr = rs / (float)M;
        P_row[node * K + j] = pr;
        const int b = csrT_ptr[node], e = csrT_ptr[node + 1];
        float cs = 0.f;
        int p = b;
        for (; p + 4 <= e; p += 4) {
            const int e0 = __ldg(&csrT_edge[p]), e1 = __ldg(&csrT_edge[p + 1]), e2 = __ldg(&csrT_edge[p + 2]), e3 = __ldg(&csrT_edge[p + 3]);
            const float v0 = __ldg(&H[(int64_t)e0 * K + j]), v1 = __ldg(&H[(int64_t)e1 * K + j]);
            const float v2 = __ldg(&H[(int64_t)e2 * K + j]), v3 = __ldg(&H[(int64_t)e3 * K + j]);
            cs += v0; cs += v1; cs += v2; cs += v3;
        }
        for (; p < e; ++p) cs += __ldg(&H[(int64_t)__ldg(&csrT_edge[p]) * K + j]);
        P_col[node * K + j] = cs / (float)nbpc_max(e - b, 1);
    }
    red[threadIdx.x] = pr;
    __syncthreads();
    for (int stride = npb / 2; stride >= 1; stride >>= 1) {
        if (slot < stride) red[threadIdx.x] += red[threadIdx.x + stride * KP];
        __syncthreads();
    }
    if (slot == 0 && j < K) partial[((int64_t)s * gridDim.x + blockIdx.x) * K + j] = red[threadIdx.x];
}

// ------------------------------------------------------------------ backward pooling on dZ (c,Q), Q % 4 == 0:
// dQ_row = row sums, dQ_col = in-edge sums, partial[s][blk][Q] = column sums of dQ_row over the block
// ROWGIVEN: dQ_row already holds the row sums (handed over by the kernel that wrote dOut, glf_last_edge_in_rowsum_kernel)
template <int Q, bool RELU, bool ROWGIVEN = false>
__global__ void __launch_bounds__(GLN_THREADS) gln_bwd_pool_kernel(const float *__restrict__ dOut, const float *__restrict__ Hout, int M,
                                                                    int N, const int32_t *__restrict__ csrT_ptr,
                                                                    const int32_t *__restrict__ csrT_edge, float *__restrict__ dQ_row,
                                                                    float *__restrict__ dQ_col, float *__restrict__ partial, int rev) {
    constexpr int G = Q / 4, NPB = GLN_THREADS / G;
    __shared__ float4 red[GLN_THREADS];
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const int bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x, s = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
    const int local = bx * NPB + slot;
    auto dz = [&](int64_t e) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(dOut + e * Q + 4 * g));
        if (RELU) {
            const float4 h = __ldg(reinterpret_cast<const float4 *>(Hout + e * Q + 4 * g));
            v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f;
            v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
        }
        return v;
    };
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (local < N) {
        const int64_t node = (int64_t)s * N + local;
        if constexpr (ROWGIVEN) {
            rs = *reinterpret_cast<const float4 *>(dQ_row + node * Q + 4 * g);
        } else {
            for (int m = 0; m < M; ++m) {
                const float4 v = dz(node * M + m);
                rs.x += v.x; rs.y += v.y; rs.z += v.z; rs.w += v.w;
            }
            *reinterpret_cast<float4 *>(dQ_row + node * Q + 4 * g) = rs;
        }
        const int b = csrT_ptr[node], e = csrT_ptr[node + 1];
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = b;
        for (; p + 4 <= e; p += 4) {
            const int e0 = __ldg(&csrT_edge[p]), e1 = __ldg(&csrT_edge[p + 1]), e2 = __ldg(&csrT_edge[p + 2]), e3 = __ldg(&csrT_edge[p + 3]);
            const float4 v0 = dz(e0), v1 = dz(e1), v2 = dz(e2), v3 = dz(e3);
            cs.x += v0.x; cs.y += v0.y; cs.z += v0.z; cs.w += v0.w;
            cs.x += v1.x; cs.y += v1.y; cs.z += v1.z; cs.w += v1.w;
            cs.x += v2.x; cs.y += v2.y; cs.z += v2.z; cs.w += v2.w;
            cs.x += v3.x; cs.y += v3.y; cs.z += v3.z; cs.w += v3.w;
        }
        for (; p < e; ++p) {
            const float4 v = dz(__ldg(&csrT_edge[p]));
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
        }
        *reinterpret_cast<float4 *>(dQ_col + node * Q + 4 * g) = cs;
    }
    gln_block_colsum<G>(red, rs);
    if (slot == 0) *reinterpret_cast<float4 *>(partial + ((int64_t)s * gridDim.x + bx) * Q + 4 * g) = red[threadIdx.x];
}

// ------------------------------------------------------------------ tiny per-sample kernels (grid = B, runtime k, q)
// column sums over the blocks of a sample, fixed order: the 32 warps of a 1024-thread block take blocks w, w+32, ...
// (4 independent partial sums per warp keep 4 loads in flight), then warp 0 adds the 32 slices
#define GLN_TINY_THREADS 1024
__device__ __forceinline__ void gln_sum_partials(const float *__restrict__ partial, int nblk, int ch, float *out_smem /*[ch]*/) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int c0 = 0; c0 < ch; c0 += 32) {
        const int cc = c0 + lane;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
        if (cc < ch) {
            int b = w;
            for (; b + 224 < nblk; b += 256) {
                const float *p = partial + (int64_t)b * ch + cc;
                const float v0 = __ldg(p), v1 = __ldg(p + 32 * (int64_t)ch), v2 = __ldg(p + 64 * (int64_t)ch), v3 = __ldg(p + 96 * (int64_t)ch);
                const float v4 = __ldg(p + 128 * (int64_t)ch), v5 = __ldg(p + 160 * (int64_t)ch), v6 = __ldg(p + 192 * (int64_t)ch),
                            v7 = __ldg(p + 224 * (int64_t)ch);
                a0 += v0; a1 += v1; a2 += v2; a3 += v3; a4 += v4; a5 += v5; a6 += v6; a7 += v7;
            }
            for (; b < nblk; b += 32) a0 += __ldg(&partial[(int64_t)b * ch + cc]);
        }
        red[w][lane] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
        __syncthreads();
        if (w == 0 && cc < ch) {
            float tot = 0.f;
#pragma unroll
            for (int ww = 0; ww < 32; ++ww) tot += red[ww][lane];
            out_smem[cc] = tot;
        }
        __syncthreads();
    }
}

// P_cube[s] = (sum_blk partial[s][blk]) / N;   Cq[s] = P_cube[s] W4 + bias
__global__ void __launch_bounds__(GLN_TINY_THREADS) gln_cube_fwd_kernel(const float *__restrict__ partial, int nblk, int N, int k, int q,
                                                                    const float *__restrict__ W4, const float *__restrict__ bias,
                                                                    float *__restrict__ P_cube, float *__restrict__ Cq) {
    __shared__ float pc[256];
    const int s = blockIdx.x;
    gln_sum_partials(partial + (int64_t)s * nblk * k, nblk, k, pc);
    for (int kk = threadIdx.x; kk < k; kk += blockDim.x) {
        pc[kk] = pc[kk] / (float)N;
        P_cube[s * k + kk] = pc[kk];
    }
    __syncthreads();
    for (int qo = threadIdx.x; qo < q; qo += blockDim.x) {
        float a = 0.f;
        for (int kk = 0; kk < k; ++kk) a += pc[kk] * __ldg(&W4[kk * q + qo]);
        Cq[s * q + qo] = a + __ldg(&bias[qo]);
    }
}

// dCq[s] = sum_blk partial[s][blk];   Gq[s] = dCq[s] W4^T / (N M)
__global__ void __launch_bounds__(GLN_TINY_THREADS) gln_cube_bwd_kernel(const float *__restrict__ partial, int nblk, int N, int M, int k, int q,
                                                                    const float *__restrict__ W4, float *__restrict__ dCq,
                                                                    float *__restrict__ Gq) {
    __shared__ float dc[256];
    const int s = blockIdx.x;
    gln_sum_partials(partial + (int64_t)s * nblk * q, nblk, q, dc);
    for (int qo = threadIdx.x; qo < q; qo += blockDim.x) dCq[s * q + qo] = dc[qo];
    __syncthreads();
    if (Gq) {
        for (int kk = threadIdx.x; kk < k; kk += blockDim.x) {
            float a = 0.f;
            for (int qo = 0; qo < q; ++qo) a += dc[qo] * __ldg(&W4[kk * q + qo]);
            Gq[s * k + kk] = a / ((float)N * (float)M);
        }
    }
}

// ------------------------------------------------------------------ thread-per-node projections, compile-time K, Q
// row loads: 16-byte vectors when the width allows it
template <int C>
__device__ __forceinline__ void gln_load_row(const float *__restrict__ p, float *x) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (int j = 0; j < C / 4; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + 4 * j));
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < C; ++j) x[j] = __ldg(&p[j]);
    }
}
template <int C>
__device__ __forceinline__ void gln_store_row(float *__restrict__ p, const float *x) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (int j = 0; j < C / 4; ++j) *reinterpret_cast<float4 *>(p + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < C; ++j) p[j] = x[j];
    }
}
// y[QO] = x[KI] * Ws, Ws [KI][QP] in shared memory (QP = QO rounded up to 4, zero padded): warp-broadcast LDS.128
template <int KI, int QO>
__device__ __forceinline__ void gln_matvec(const float *x, const float *Ws, float *y) {
    constexpr int QP = (QO + 3) / 4 * 4;
    float acc[QP];
#pragma unroll
    for (int j = 0; j < QP; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KI; ++kk) {
#pragma unroll
        for (int j = 0; j < QP / 4; ++j) {
            const float4 w = *reinterpret_cast<const float4 *>(Ws + kk * QP + 4 * j);
            acc[4 * j] = fmaf(x[kk], w.x, acc[4 * j]); acc[4 * j + 1] = fmaf(x[kk], w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(x[kk], w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x[kk], w.w, acc[4 * j + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < QO; ++j) y[j] = acc[j];
}

// Q_col = P_col W2;  Q_row = P_row W3 + Cq[sample]
template <int K, int Q>
__global__ void __launch_bounds__(GLN_THREADS) gln_node_project_kernel(const float *__restrict__ P_col, const float *__restrict__ P_row,
                                                                        const float *__restrict__ Cq, const float *__restrict__ W, int BN,
                                                                        int N, float *__restrict__ Q_col, float *__restrict__ Q_row) {
    constexpr int QP = (Q + 3) / 4 * 4;
    __shared__ __align__(16) float W2s[K * QP], W3s[K * QP];
    for (int i = threadIdx.x; i < K * QP; i += GLN_THREADS) {
        const int kk = i / QP, qo = i % QP;
        W2s[i] = qo < Q ? __ldg(&W[(int64_t)K * Q + kk * Q + qo]) : 0.f;
        W3s[i] = qo < Q ? __ldg(&W[2 * (int64_t)K * Q + kk * Q + qo]) : 0.f;
    }
    __syncthreads();
    for (int node = blockIdx.x * GLN_THREADS + threadIdx.x; node < BN; node += gridDim.x * GLN_THREADS) {
        float x[K], y[Q];
        gln_load_row<K>(P_col + (int64_t)node * K, x);
        gln_matvec<K, Q>(x, W2s, y);
        gln_store_row<Q>(Q_col + (int64_t)node * Q, y);
        gln_load_row<K>(P_row + (int64_t)node * K, x);
        gln_matvec<K, Q>(x, W3s, y);
        const float *cq = Cq + (node / N) * Q;
#pragma unroll
        for (int j = 0; j < Q; ++j) y[j] += __ldg(&cq[j]);
        gln_store_row<Q>(Q_row + (int64_t)node * Q, y);
    }
}

// G_col = (dQ_col W2^T) / max(indeg, 1);  G_row = (dQ_row W3^T) / M + Gq[sample]
// add_w1 (last layer, whose output is the row mean of Z): G_row additionally carries the row term of dH,
// (dQ_row W1^T) / M, i.e. dQ_row is projected with (W3 + W1)^T
template <int K, int Q>
__global__ void __launch_bounds__(GLN_THREADS) gln_node_grad_kernel(const float *__restrict__ dQ_col, const float *__restrict__ dQ_row,
                                                                     const float *__restrict__ Gq, const float *__restrict__ W,
                                                                     const int32_t *__restrict__ csrT_ptr, int BN, int N, int M,
                                                                     int add_w1, float *__restrict__ G_col, float *__restrict__ G_row) {
    constexpr int KP = (K + 3) / 4 * 4;
    __shared__ __align__(16) float W2t[Q * KP], W3t[Q * KP];   // transposed: [q][k]
    for (int i = threadIdx.x; i < Q * KP; i += GLN_THREADS) {
        const int qo = i / KP, kk = i % KP;
        W2t[i] = kk < K ? __ldg(&W[(int64_t)K * Q + kk * Q + qo]) : 0.f;
        W3t[i] = kk < K ? __ldg(&W[2 * (int64_t)K * Q + kk * Q + qo]) + (add_w1 ? __ldg(&W[kk * Q + qo]) : 0.f) : 0.f;
    }
    __syncthreads();
    const float rm = 1.f / (float)M;
    for (int node = blockIdx.x * GLN_THREADS + threadIdx.x; node < BN; node += gridDim.x * GLN_THREADS) {
        float x[Q], y[K];
        gln_load_row<Q>(dQ_col + (int64_t)node * Q, x);
        gln_matvec<Q, K>(x, W2t, y);
        const float ri = 1.f / (float)nbpc_max(__ldg(&csrT_ptr[node + 1]) - __ldg(&csrT_ptr[node]), 1);
#pragma unroll
        for (int j = 0; j < K; ++j) y[j] = y[j] * ri;
        gln_store_row<K>(G_col + (int64_t)node * K, y);
        gln_load_row<Q>(dQ_row + (int64_t)node * Q, x);
        gln_matvec<Q, K>(x, W3t, y);
        const float *gq = Gq + (node / N) * K;
#pragma unroll
        for (int j = 0; j < K; ++j) y[j] = fmaf(y[j], rm, __ldg(&gq[j]));
        gln_store_row<K>(G_row + (int64_t)node * K, y);
    }
}

// ------------------------------------------------------------------ 8 threads per node (Q % 8 == 0 / K % 8 == 0)
// For the narrow first layer (K <= 10) a node's outputs are split over 8 adjacent lanes: every lane reads the whole (short)
// input row and its slice of the weights from shared memory (21 us instead of 31 us at k = 3, q = 32).  Measured and NOT kept
// for K >= 16: 8 lanes per node re-read every 128-byte row 8 times through L1 (59 us vs 45 us), 2 lanes per node (lane parity
// = col / row term) break the warp-broadcast weight loads (57 us).
template <int K, int Q>
__global__ void __launch_bounds__(GLN_THREADS) gln_node_project8_kernel(const float *__restrict__ P_col, const float *__restrict__ P_row,
                                                                         const float *__restrict__ Cq, const float *__restrict__ W, int BN,
                                                                         int N, float *__restrict__ Q_col, float *__restrict__ Q_row) {
    constexpr int QT = Q / 8;
    static_assert(Q % 8 == 0, "Q must be a multiple of 8");
    __shared__ __align__(16) float W2s[K * Q], W3s[K * Q];
    for (int i = threadIdx.x; i < K * Q; i += GLN_THREADS) {
        W2s[i] = __ldg(&W[(int64_t)K * Q + i]);
        W3s[i] = __ldg(&W[2 * (int64_t)K * Q + i]);
    }
    __syncthreads();
    const int sub = threadIdx.x & 7, q0 = sub * QT;
    for (int64_t node = ((int64_t)blockIdx.x * GLN_THREADS + threadIdx.x) >> 3; node < BN; node += ((int64_t)gridDim.x * GLN_THREADS) >> 3) {
        float x[K], y[K], a2[QT], a3[QT];
        gln_load_row<K>(P_col + node * K, x);
        gln_load_row<K>(P_row + node * K, y);
        const float *cq = Cq + (node / N) * Q + q0;
#pragma unroll
        for (int j = 0; j < QT; ++j) { a2[j] = 0.f; a3[j] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < K; ++kk)
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                a2[j] = fmaf(x[kk], W2s[kk * Q + q0 + j], a2[j]);
                a3[j] = fmaf(y[kk], W3s[kk * Q + q0 + j], a3[j]);
            }
#pragma unroll
        for (int j = 0; j < QT; ++j) a3[j] += __ldg(&cq[j]);
        gln_store_row<QT>(Q_col + node * Q + q0, a2);
        gln_store_row<QT>(Q_row + node * Q + q0, a3);
    }
}

// ------------------------------------------------------------------ final: every weight / bias gradient of the layer
//   dW[i] (i = 0,1,2) = sum over the n_i per-block partials at part_i (fixed order: 32 warps take blocks w, w+32, ...)
//   dW[3] = P_cube^T dCq;  dB = sum_s dCq[s]
// grid = ceil(k q / 32) blocks of 32 lanes x 32 warps; a null part_i (or n_i == 0) leaves dW[i] untouched
struct GlnFinalArgs {
    const float *part[3];
    int n[3];
    int tr[3];   // partial blocks are stored transposed, (q,k)
};
__global__ void __launch_bounds__(1024) gln_final_kernel(GlnFinalArgs a, const float *__restrict__ P_cube, const float *__restrict__ dCq,
                                                          int B, int k, int q, float *__restrict__ dW, float *__restrict__ dB) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int kq = k * q, o = blockIdx.x * 32 + lane, i = blockIdx.y;   // grid.y: 0..2 partial sets, 3: dW4 and dB
    if (i < 3) {
        const float *part = i == 0 ? a.part[0] : (i == 1 ? a.part[1] : a.part[2]);
        const int n = i == 0 ? a.n[0] : (i == 1 ? a.n[1] : a.n[2]), tr = i == 0 ? a.tr[0] : (i == 1 ? a.tr[1] : a.tr[2]);
        if (part == nullptr || n <= 0) return;   // uniform over the block
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
        if (o < kq) {
            const float *src = part + (tr ? (o % q) * k + (o / q) : o);
            int b = w;
            for (; b + 224 < n; b += 256) {
                const float *p = src + (int64_t)b * kq;
                const float v0 = __ldg(p), v1 = __ldg(p + 32 * (int64_t)kq), v2 = __ldg(p + 64 * (int64_t)kq), v3 = __ldg(p + 96 * (int64_t)kq);
                const float v4 = __ldg(p + 128 * (int64_t)kq), v5 = __ldg(p + 160 * (int64_t)kq), v6 = __ldg(p + 192 * (int64_t)kq),
                            v7 = __ldg(p + 224 * (int64_t)kq);
                a0 += v0; a1 += v1; a2 += v2; a3 += v3; a4 += v4; a5 += v5; a6 += v6; a7 += v7;
            }
            for (; b < n; b += 32) a0 += __ldg(&src[(int64_t)b * kq]);
        }
        red[w][lane] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
        __syncthreads();
        if (w == 0 && o < kq) {
            float tot = 0.f;
#pragma unroll
            for (int ww = 0; ww < 32; ++ww) tot += red[ww][lane];
            dW[(int64_t)i * kq + o] = tot;
        }
        return;
    }
    if (w == 0 && o < kq) {   // dW4
        const int kk = o / q, qo = o % q;
        float acc = 0.f;
        for (int s = 0; s < B; ++s) acc += __ldg(&P_cube[s * k + kk]) * __ldg(&dCq[s * q + qo]);
        dW[3 * (int64_t)kq + o] = acc;
    }
    if (blockIdx.x == 0 && w == 1) {   // dB
        for (int qo = lane; qo < q; qo += 32) {
            float acc = 0.f;
            for (int s = 0; s < B; ++s) acc += __ldg(&dCq[s * q + qo]);
            dB[qo] = acc;
        }
    }
}
#endif  // !NBPC_HOST_EMU
