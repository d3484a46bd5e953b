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

// e / M for 0 <= e < 2^31 with magic = floor(2^32 / M): the estimate is exact or one short.  M == 1 would need
// magic = 2^32; 2^32 - 1 gives the estimate e - 1 (0 for e = 0), which the correction step turns into e.
static inline uint32_t glk3_magic(int M) { return M == 1 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)1 << 32) / (uint32_t)M); }
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

// Same output, plus the row pool of the NEXT layer: P_row_next[i] = (1/M) sum_m out[i M + m] (graph.py:428-449 applied to
// this layer's output).  One thread owns a 4-channel group of one ROW node and walks its M contiguous edges in ascending
// order - the summation order of gln_pool_kernel, so the hand-over is bit-identical to pooling the stored tensor - with
// U edges in flight; Q_row[i] is loaded once per thread instead of once per edge.  A warp instruction stores Q/4-lane
// groups of full 16 Q-byte rows (whole 128-byte lines for Q = 32) M rows apart.
template <int K, int Q, bool RELU>
__global__ void __launch_bounds__(GLK3_THREADS) glk3_edge_out_rowpool_kernel(const float *__restrict__ E, const int32_t *__restrict__ col,
                                                                              const float *__restrict__ W1, const float *__restrict__ Q_col,
                                                                              const float *__restrict__ Q_row, uint32_t n_rows, uint32_t M,
                                                                              float *__restrict__ out, float *__restrict__ P_row_next) {
    constexpr int G = Q / 4, RPB = GLK3_THREADS / G, U = 2;
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const uint32_t row = blockIdx.x * (uint32_t)RPB + slot;
    if (row >= n_rows) return;
    float4 w[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) w[kk] = __ldg(reinterpret_cast<const float4 *>(W1 + kk * Q + 4 * g));
    const float4 qr = __ldg(reinterpret_cast<const float4 *>(Q_row + (size_t)row * Q + 4 * g));
    const uint32_t e0 = row * M;
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    auto emit = [&](uint32_t e, const float *x, const float4 &qc) {
        // (((x0 w0 + x1 w1) + x2 w2) + ...) + (Q_col + Q_row): the expression of glk3_edge_out_kernel
        float4 o = make_float4(x[0] * w[0].x, x[0] * w[0].y, x[0] * w[0].z, x[0] * w[0].w);
#pragma unroll
        for (int kk = 1; kk < K; ++kk) {
            o.x = fmaf(x[kk], w[kk].x, o.x); o.y = fmaf(x[kk], w[kk].y, o.y);
            o.z = fmaf(x[kk], w[kk].z, o.z); o.w = fmaf(x[kk], w[kk].w, o.w);
        }
        o.x += qc.x + qr.x; o.y += qc.y + qr.y; o.z += qc.z + qr.z; o.w += qc.w + qr.w;
        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4 *>(out + (size_t)e * Q + 4 * g) = o;
        rs.x += o.x; rs.y += o.y; rs.z += o.z; rs.w += o.w;
    };
    uint32_t m = 0;
    for (; m + U <= M; m += U) {
        float x[U][K];
        float4 qc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t e = e0 + m + u;
            const int cidx = __ldg(&col[e]);
#pragma unroll
            for (int kk = 0; kk < K; ++kk) x[u][kk] = __ldg(&E[K * (size_t)e + kk]);
            qc[u] = __ldg(reinterpret_cast<const float4 *>(Q_col + (size_t)cidx * Q + 4 * g));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) emit(e0 + m + u, x[u], qc[u]);
    }
    for (; m < M; ++m) {
        const uint32_t e = e0 + m;
        float x[K];
#pragma unroll
        for (int kk = 0; kk < K; ++kk) x[kk] = __ldg(&E[K * (size_t)e + kk]);
        emit(e, x, __ldg(reinterpret_cast<const float4 *>(Q_col + (size_t)__ldg(&col[e]) * Q + 4 * g)));
    }
    const float fm = (float)M;
    *reinterpret_cast<float4 *>(P_row_next + (size_t)row * Q + 4 * g) = make_float4(rs.x / fm, rs.y / fm, rs.z / fm, rs.w / fm);
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

// ------------------------------------------------------------------ first layer, whole backward in ONE pass over dZ
// The first layer needs no input gradient, so its node-level gradients can be folded into the edge stream:
//   dW1 = sum_e E[e]^T dZ[e]           dW2 = P_col^T dQ_col = sum_e P_col[col[e]]^T dZ[e]
//   dW3 = P_row^T dQ_row = sum_e P_row[e / M]^T dZ[e]        dCq[s] = sum_{e in s} dZ[e]  (-> dW4, dB)
// i.e. dZ (c,Q) is read once instead of three times (the backward pooling read it twice more) and the pooling, the node
// X^T Y kernels and their partial reductions disappear.  K = 3 only.  The 40 accumulator registers per thread leave too
// few loads in flight for a register-only stream (measured 2 TB/s), so the dZ / E / col tiles are staged through a
// 3-deep cp.async ring in shared memory: bytes in flight no longer depend on the register budget.
// grid (blocks per sample, B): a block's tiles belong to one sample, so its column sum is a partial of dCq[sample].
#define GLK3_FB_TILE 128
#define GLK3_FB_STAGES 3
__device__ __forceinline__ void glk3_cp_async4(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void glk3_cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
template <int Q, bool RELU>
struct Glk3FbCfg {
    static constexpr int ZF = GLK3_FB_TILE * Q;                                   // floats of one dZ tile
    static constexpr int STAGE_F = ZF * (RELU ? 2 : 1) + GLK3_FB_TILE * 3 + GLK3_FB_TILE;   // dZ | (H_out) | E | col
    static constexpr size_t SMEM = sizeof(float) * (size_t)STAGE_F * GLK3_FB_STAGES;
};
template <int Q, bool RELU>
__global__ void __launch_bounds__(GLK3_THREADS) glk3_first_layer_bwd_kernel(
    const float *__restrict__ E, const float *__restrict__ dOut, const float *__restrict__ Hout, const int32_t *__restrict__ col,
    const float *__restrict__ P_col, const float *__restrict__ P_row, uint32_t edges_per_sample, uint32_t tiles_per_block, uint32_t M,
    uint32_t magic, float *__restrict__ part1, float *__restrict__ part2, float *__restrict__ part3, float *__restrict__ colsum_partial,
    int rev) {
    using Cfg = Glk3FbCfg<Q, RELU>;
    constexpr int K = 3, G = Q / 4, SLOTS = GLK3_THREADS / G, TILE = GLK3_FB_TILE, S = GLK3_FB_STAGES;
    extern __shared__ __align__(16) float glk3_smem[];
    const int tid = threadIdx.x, g = tid % G, slot = tid / G;
    // rev: the blocks walk dZ from its end (what the producing kernel wrote last is still in L2); same partial slots
    const uint32_t bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x, s = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
    const uint32_t e_sample = s * edges_per_sample, e_sample_end = e_sample + edges_per_sample;
    const uint32_t ntiles_sample = (edges_per_sample + TILE - 1) / TILE;
    const uint32_t t_begin = bx * tiles_per_block, t_end = nbpc_min(t_begin + tiles_per_block, ntiles_sample);

    auto issue = [&](uint32_t t, int st) {
        float *Zs = glk3_smem + (size_t)st * Cfg::STAGE_F, *Hs = Zs + Cfg::ZF;
        float *Es = Zs + Cfg::ZF * (RELU ? 2 : 1);
        int *Cs = reinterpret_cast<int *>(Es + TILE * 3);
        const uint32_t e0 = e_sample + t * TILE;
        for (int i = tid; i < TILE * G; i += GLK3_THREADS) {
            const int r = i / G, ch = i % G;
            const uint32_t e = e0 + r;
            if (e < e_sample_end) {
                glk3_cp_async16(Zs + r * Q + 4 * ch, dOut + (size_t)e * Q + 4 * ch);
                if (RELU) glk3_cp_async16(Hs + r * Q + 4 * ch, Hout + (size_t)e * Q + 4 * ch);
            } else {
                *reinterpret_cast<float4 *>(Zs + r * Q + 4 * ch) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (RELU) *reinterpret_cast<float4 *>(Hs + r * Q + 4 * ch) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int i = tid; i < TILE * 4; i += GLK3_THREADS) {   // 3 floats of E and the column index per row
            const int r = i >> 2, w = i & 3;
            const uint32_t e = e0 + r;
            if (e < e_sample_end) {
                if (w < 3) glk3_cp_async4(Es + r * 3 + w, E + (size_t)e * 3 + w);
                else glk3_cp_async4(Cs + r, col + e);
            } else {
                if (w < 3) Es[r * 3 + w] = 0.f;
                else Cs[r] = (int)(s * (edges_per_sample / M));   // any valid node id of this sample
            }
        }
    };

    float4 a1[K], a2[K], a3[K], cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < K; ++kk) a1[kk] = a2[kk] = a3[kk] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll
    for (int p = 0; p < S - 1; ++p) {
        if (t_begin + p < t_end) issue(t_begin + p, p);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    int st = 0;
    for (uint32_t t = t_begin; t < t_end; ++t) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(S - 2) : "memory");
        __syncthreads();                       // tile t landed for everybody; the stage refilled below is no longer read
        if (t + S - 1 < t_end) issue(t + S - 1, (st + S - 1) % S);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        const float *Zs = glk3_smem + (size_t)st * Cfg::STAGE_F, *Hs = Zs + Cfg::ZF;
        const float *Es = Zs + Cfg::ZF * (RELU ? 2 : 1);
        const int *Cs = reinterpret_cast<const int *>(Es + TILE * 3);
        const uint32_t e0 = e_sample + t * TILE;
#pragma unroll 2
        for (int r = slot; r < TILE; r += SLOTS) {
            const int cidx = Cs[r];
            const uint32_t e = nbpc_min(e0 + r, e_sample_end - 1);
            const uint32_t ridx = glk3_div(e, M, magic);
            float pc[K], pr[K], x[K];
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                pc[kk] = __ldg(&P_col[K * (size_t)cidx + kk]);
                pr[kk] = __ldg(&P_row[K * (size_t)ridx + kk]);
                x[kk] = Es[r * 3 + kk];
            }
            float4 v = *reinterpret_cast<const float4 *>(Zs + r * Q + 4 * g);
            if (RELU) {
                const float4 h = *reinterpret_cast<const float4 *>(Hs + r * Q + 4 * g);
                v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
            }
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                a1[kk].x = fmaf(x[kk], v.x, a1[kk].x); a1[kk].y = fmaf(x[kk], v.y, a1[kk].y);
                a1[kk].z = fmaf(x[kk], v.z, a1[kk].z); a1[kk].w = fmaf(x[kk], v.w, a1[kk].w);
                a2[kk].x = fmaf(pc[kk], v.x, a2[kk].x); a2[kk].y = fmaf(pc[kk], v.y, a2[kk].y);
                a2[kk].z = fmaf(pc[kk], v.z, a2[kk].z); a2[kk].w = fmaf(pc[kk], v.w, a2[kk].w);
                a3[kk].x = fmaf(pr[kk], v.x, a3[kk].x); a3[kk].y = fmaf(pr[kk], v.y, a3[kk].y);
                a3[kk].z = fmaf(pr[kk], v.z, a3[kk].z); a3[kk].w = fmaf(pr[kk], v.w, a3[kk].w);
            }
        }
        st = (st + 1) % S;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    float4 *red = reinterpret_cast<float4 *>(glk3_smem);
    const size_t blk = (size_t)s * gridDim.x + bx;
    // 10 rows (3 x dW1, 3 x dW2, 3 x dW3, column sum) through the same fixed tree over the slots, one at a time
#pragma unroll 1
    for (int r = 0; r < 3 * K + 1; ++r) {
        float4 mine = cs;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (r == j) mine = a1[j];
            if (r == K + j) mine = a2[j];
            if (r == 2 * K + j) mine = a3[j];
        }
        red[tid] = mine;
        __syncthreads();
        for (int stride = SLOTS / 2; stride >= 1; stride >>= 1) {
            if (slot < stride) {
                float4 a = red[tid];
                const float4 b = red[tid + stride * G];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                red[tid] = a;
            }
            __syncthreads();
        }
        if (slot == 0) {
            float *dst = r < K ? part1 + (blk * K + r) * Q
                               : (r < 2 * K ? part2 + (blk * K + (r - K)) * Q
                                            : (r < 3 * K ? part3 + (blk * K + (r - 2 * K)) * Q : colsum_partial + blk * Q));
            *reinterpret_cast<float4 *>(dst + 4 * g) = red[tid];
        }
        __syncthreads();
    }
}
#endif  // !NBPC_HOST_EMU
