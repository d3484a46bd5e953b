// graph_layer_vin.cuh - VIRTUAL layer input: inside a network the output of the FIRST graph layer,
//     H1[e] = relu( E[e] W1 + Q_col[col[e]] + Q_row[e / M] )          (graph.py:437-453 with the pooled terms at node level)
// is a function of the 12-byte edge feature row E[e] and two node tables that fit the L2 cache (32 channels: 2 x 33.5 MB at
// 8 x 32^3).  Materialising it costs one 128-byte write per edge and four 128-byte reads (pooling x 2, forward edge kernel,
// backward edge kernel): 2.35 GB of the 5.3 GB a [3,32,16,3] step moves.  The consumers below recompute it instead - 3 FMAs
// per channel from E plus two L2 gathers - with ONE shared expression, so every consumer sees bit-identical values (the same
// bits the materialising kernel glk3_edge_out_kernel writes).
#pragma once
#include "nbpc_common.cuh"
#ifndef NBPC_HOST_EMU

struct GlVin {
    const float *E;      // (c, K0) edge features
    const float *W1;     // (K0, K) first weight of the producing layer
    const float *Qc;     // (B*N, K) its node-level column term  Q_col = P_col W2
    const float *Qr;     // (B*N, K) its node-level row term     Q_row = P_row W3 + (P_cube W4 + B)
};

// channels [4g, 4g + 4) of the virtual row: x = E[e] (K0 values), w[kk] = W1[kk][4g .. 4g+3]
template <int K0>
__device__ __forceinline__ float4 glv_row4(const float *x, const float4 *w, const float4 qc, const float4 qr) {
    float4 o = make_float4(x[0] * w[0].x, x[0] * w[0].y, x[0] * w[0].z, x[0] * w[0].w);
#pragma unroll
    for (int kk = 1; kk < K0; ++kk) {
        o.x = fmaf(x[kk], w[kk].x, o.x); o.y = fmaf(x[kk], w[kk].y, o.y);
        o.z = fmaf(x[kk], w[kk].z, o.z); o.w = fmaf(x[kk], w[kk].w, o.w);
    }
    o.x += qc.x + qr.x; o.y += qc.y + qr.y; o.z += qc.z + qr.z; o.w += qc.w + qr.w;
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    return o;
}

// e / M for 0 <= e < 2^31 with magic = floor(2^32 / M) (2^32 - 1 for M = 1): exact or one short
__device__ __forceinline__ uint32_t glv_div(uint32_t e, uint32_t M, uint32_t magic) {
    uint32_t qt = __umulhi(e, magic);
    if (e - qt * M >= M) ++qt;
    return qt;
}
static inline uint32_t glv_magic(int M) { return M == 1 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)1 << 32) / (uint32_t)M); }

// ------------------------------------------------------------------ pooling of the virtual tensor (forward of the NEXT layer)
// thread per (node, 4-channel group); grid (blocks per sample, B): P_row = mean over the node's M out-edges, P_col = mean
// over its in-edges (CSR transpose, ascending edge id), partial[s][blk][K] = column sums of P_row over the block.
// Out-edges gather Q_col[col[e]] (Q_row is the node's own), in-edges gather Q_row[e / M] and E[e] (Q_col is the node's own).
#define GLV_THREADS 256
template <int K0, int K>
__global__ void __launch_bounds__(GLV_THREADS) gln_pool_vin_kernel(const GlVin V, const int32_t *__restrict__ col, int M, uint32_t magic, int N,
                                                                   const int32_t *__restrict__ csrT_ptr, const int32_t *__restrict__ csrT_edge,
                                                                   float *__restrict__ P_row, float *__restrict__ P_col,
                                                                   float *__restrict__ partial) {
    constexpr int G = K / 4, NPB = GLV_THREADS / G;
    __shared__ float4 red[GLV_THREADS];
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const int local = blockIdx.x * NPB + slot, s = blockIdx.y;
    float4 w[K0];
#pragma unroll
    for (int kk = 0; kk < K0; ++kk) w[kk] = __ldg(reinterpret_cast<const float4 *>(V.W1 + kk * K + 4 * g));
    float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (local < N) {
        const int64_t node = (int64_t)s * N + local;
        const float4 qr_own = __ldg(reinterpret_cast<const float4 *>(V.Qr + node * K + 4 * g));
        const float4 qc_own = __ldg(reinterpret_cast<const float4 *>(V.Qc + node * K + 4 * g));
        // ---- out-edges: contiguous rows of E / col
        float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t e0 = node * M;
        int m = 0;
        for (; m + 4 <= M; m += 4) {   // four edges in flight (two dependent L2 round trips each: col -> Q_col row)
            int cc[4];
            float x[4][K0];
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cc[u] = __ldg(&col[e0 + m + u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int kk = 0; kk < K0; ++kk) x[u][kk] = __ldg(&V.E[(e0 + m + u) * K0 + kk]);
                q[u] = __ldg(reinterpret_cast<const float4 *>(V.Qc + (int64_t)cc[u] * K + 4 * g));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 h = glv_row4<K0>(x[u], w, q[u], qr_own);
                rs.x += h.x; rs.y += h.y; rs.z += h.z; rs.w += h.w;
            }
        }
        for (; m < M; ++m) {
            const int c0 = __ldg(&col[e0 + m]);
            float x0[K0];
#pragma unroll
            for (int kk = 0; kk < K0; ++kk) x0[kk] = __ldg(&V.E[(e0 + m) * K0 + kk]);
            const float4 h0 = glv_row4<K0>(x0, w, __ldg(reinterpret_cast<const float4 *>(V.Qc + (int64_t)c0 * K + 4 * g)), qr_own);
            rs.x += h0.x; rs.y += h0.y; rs.z += h0.z; rs.w += h0.w;
        }
        const float fm = (float)M;
        pr = make_float4(rs.x / fm, rs.y / fm, rs.z / fm, rs.w / fm);
        *reinterpret_cast<float4 *>(P_row + node * K + 4 * g) = pr;
        // ---- in-edges (ascending edge id: the summation order of gln_pool_kernel)
        const int b = csrT_ptr[node], e = csrT_ptr[node + 1];
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = b;
        for (; p + 4 <= e; p += 4) {
            uint32_t ee[4];
            float x[4][K0];
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ee[u] = (uint32_t)__ldg(&csrT_edge[p + u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int kk = 0; kk < K0; ++kk) x[u][kk] = __ldg(&V.E[(int64_t)ee[u] * K0 + kk]);
                q[u] = __ldg(reinterpret_cast<const float4 *>(V.Qr + (int64_t)glv_div(ee[u], M, magic) * K + 4 * g));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 h = glv_row4<K0>(x[u], w, qc_own, q[u]);
                cs.x += h.x; cs.y += h.y; cs.z += h.z; cs.w += h.w;
            }
        }
        for (; p < e; ++p) {
            const uint32_t ea = (uint32_t)__ldg(&csrT_edge[p]);
            float x0[K0];
#pragma unroll
            for (int kk = 0; kk < K0; ++kk) x0[kk] = __ldg(&V.E[(int64_t)ea * K0 + kk]);
            const float4 h0 = glv_row4<K0>(x0, w, qc_own, __ldg(reinterpret_cast<const float4 *>(V.Qr + (int64_t)glv_div(ea, M, magic) * K + 4 * g)));
            cs.x += h0.x; cs.y += h0.y; cs.z += h0.z; cs.w += h0.w;
        }
        const float fc = (float)nbpc_max(e - b, 1);
        *reinterpret_cast<float4 *>(P_col + node * K + 4 * g) = make_float4(cs.x / fc, cs.y / fc, cs.z / fc, cs.w / fc);
    }
    // fixed tree over the node slots of the block (same as gln_block_colsum)
    red[threadIdx.x] = pr;
    __syncthreads();
#pragma unroll
    for (int stride = NPB / 2; stride >= 1; stride >>= 1) {
        if (slot < stride) {
            float4 a = red[threadIdx.x];
            const float4 bb = red[threadIdx.x + stride * G];
            a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
            red[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (slot == 0) *reinterpret_cast<float4 *>(partial + ((int64_t)s * gridDim.x + blockIdx.x) * K + 4 * g) = red[threadIdx.x];
}
#endif  // !NBPC_HOST_EMU
