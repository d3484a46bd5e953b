// graph_layer_fast.cuh - tiled sm_100a kernels for the shift-invariant graph layer.
//
// Same math and the same fixed summation orders as the baseline kernels in graph_layer.cu, but laid
// out for HBM throughput:
//   * persistent blocks walk 128-edge tiles; the NEXT tile is prefetched into the other shared-memory
//     buffer with 16-byte cp.async copies (512 contiguous bytes per warp instruction) while the
//     current one is consumed; tile rows are padded so that a thread reading its own row with LDS.128
//     is bank-conflict free;
//   * the edge-level GEMM  Z = H W1  runs thread-per-edge-row with the row in registers and W1 read
//     from shared memory as broadcast LDS.128 (4 FMAs per shared load, no divergence); the epilogue
//     gathers (Q_col[col[e]], Q_row[e/M]) are issued before the FMA loop so their L2 latency hides;
//   * outputs leave through a per-warp shared staging area so every global store instruction writes
//     512 contiguous bytes;
//   * pooling gathers use float4 channel groups (a 128 B edge row is read by 8 adjacent lanes);
//   * X^T Y reductions (dW1 over edges, dW2/dW3 over nodes) accumulate 4x4 register micro-tiles per
//     block and are reduced over blocks in a fixed order (no float atomics => bit-reproducible).
// Compile-time channel widths: K in {3,16,32,64}, Q in {16,32,64} for the edge-level kernels; any
// other shape falls back to the baseline kernels.
#pragma once
#include "nbpc_common.cuh"

#ifndef NBPC_HOST_EMU

#define GLF_TE 128       // edges per tile == threads per block
#define GLF_THREADS 128

__device__ __forceinline__ void glf_cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void glf_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void glf_cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ float4 glf_ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// row stride (floats) for a C-wide tile row: multiple of 4, (stride/4) odd => conflict-free LDS.128 per row
__host__ __device__ constexpr int glf_stride(int C) { return (C % 8 == 0) ? C + 4 : C + 8; }

// stage rows [row0, row0+GLF_TE) of a (rows_total, C) row-major tensor into smem (row stride CS); C % 4 == 0
template <int C, int CS>
__device__ __forceinline__ void glf_stage_tile(float *smem, const float *__restrict__ g, int64_t row0, int64_t rows_total) {
    constexpr int CH = C / 4;
    for (int i = threadIdx.x; i < GLF_TE * CH; i += GLF_THREADS) {
        const int r = i / CH, ch = i % CH;
        float *dst = smem + r * CS + 4 * ch;
        if (row0 + r < rows_total) glf_cp_async16(dst, g + (row0 + r) * C + 4 * ch);
        else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// 3-wide rows (input edges): plain loads, zero-padded to 4 floats per row
template <int CS>
__device__ __forceinline__ void glf_stage_tile3(float *smem, const float *__restrict__ g, int64_t row0, int64_t rows_total) {
    for (int i = threadIdx.x; i < GLF_TE * 4; i += GLF_THREADS) {
        const int r = i / 4, kk = i % 4;
        smem[r * CS + kk] = (kk < 3 && row0 + r < rows_total) ? __ldg(&g[(row0 + r) * 3 + kk]) : 0.f;
    }
}

// coalesced store of a warp's 32 x C staged rows (smem row stride CS) to global rows [row0w, row0w+32)
template <int C, int CS>
__device__ __forceinline__ void glf_store_warp_rows(const float *swarp, float *__restrict__ g, int64_t row0w, int64_t rows_total) {
    constexpr int CH = C / 4;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int chunk = i * 32 + lane;
        const int r = chunk / CH, ch = chunk % CH;
        if (row0w + r < rows_total) {
            const float4 v = *reinterpret_cast<const float4 *>(swarp + r * CS + 4 * ch);
            *reinterpret_cast<float4 *>(g + (row0w + r) * C + 4 * ch) = v;
        }
    }
}

// ------------------------------------------------------------------ forward: pooling
// thread per (node, 4-channel group): P_row = mean of the node's M contiguous edge rows,
// P_col = mean over in-edges (CSR transpose, ascending edge id)
template <int K>
__global__ void __launch_bounds__(256) glf_pool_kernel(const float *__restrict__ H, int M, int BN,
                                                        const int32_t *__restrict__ csrT_ptr,
                                                        const int32_t *__restrict__ csrT_edge,
                                                        float *__restrict__ P_row, float *__restrict__ P_col) {
    constexpr int G = K / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * G) return;
    const int node = (int)(t / G), g = (int)(t % G);
    const float *hr = H + ((int64_t)node * M) * K + 4 * g;
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < M; ++m) {
        const float4 v = glf_ldg4(hr + (int64_t)m * K);
        rs.x += v.x; rs.y += v.y; rs.z += v.z; rs.w += v.w;
    }
    const float fm = (float)M;
    *reinterpret_cast<float4 *>(P_row + (int64_t)node * K + 4 * g) = make_float4(rs.x / fm, rs.y / fm, rs.z / fm, rs.w / fm);
    const int b = csrT_ptr[node], e = csrT_ptr[node + 1];
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = b;
    for (; p + 4 <= e; p += 4) {   // 4 independent row gathers in flight
        const int e0 = __ldg(&csrT_edge[p]), e1 = __ldg(&csrT_edge[p + 1]), e2 = __ldg(&csrT_edge[p + 2]), e3 = __ldg(&csrT_edge[p + 3]);
        const float4 v0 = glf_ldg4(H + (int64_t)e0 * K + 4 * g), v1 = glf_ldg4(H + (int64_t)e1 * K + 4 * g);
        const float4 v2 = glf_ldg4(H + (int64_t)e2 * K + 4 * g), v3 = glf_ldg4(H + (int64_t)e3 * K + 4 * g);
        cs.x += v0.x; cs.y += v0.y; cs.z += v0.z; cs.w += v0.w;
        cs.x += v1.x; cs.y += v1.y; cs.z += v1.z; cs.w += v1.w;
        cs.x += v2.x; cs.y += v2.y; cs.z += v2.z; cs.w += v2.w;
        cs.x += v3.x; cs.y += v3.y; cs.z += v3.z; cs.w += v3.w;
    }
    for (; p < e; ++p) {
        const float4 v = glf_ldg4(H + (int64_t)__ldg(&csrT_edge[p]) * K + 4 * g);
        cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
    const float fc = (float)nbpc_max(e - b, 1);
    *reinterpret_cast<float4 *>(P_col + (int64_t)node * K + 4 * g) = make_float4(cs.x / fc, cs.y / fc, cs.z / fc, cs.w / fc);
}

// same pooling on the (masked) gradient dZ (c,Q): dQ_row = row sums, dQ_col = in-edge sums
template <int Q, bool RELU>
__global__ void __launch_bounds__(256) glf_bwd_pool_kernel(const float *__restrict__ dOut, const float *__restrict__ Hout,
                                                            int M, int BN, const int32_t *__restrict__ csrT_ptr,
                                                            const int32_t *__restrict__ csrT_edge,
                                                            float *__restrict__ dQ_row, float *__restrict__ dQ_col) {
    constexpr int G = Q / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * G) return;
    const int node = (int)(t / G), g = (int)(t % G);
    auto dz = [&](int64_t e) {
        float4 v = glf_ldg4(dOut + e * Q + 4 * g);
        if (RELU) {
            const float4 h = glf_ldg4(Hout + e * Q + 4 * g);
            v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f;
            v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
        }
        return v;
    };
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < M; ++m) {
        const float4 v = dz((int64_t)node * M + m);
        rs.x += v.x; rs.y += v.y; rs.z += v.z; rs.w += v.w;
    }
    *reinterpret_cast<float4 *>(dQ_row + (int64_t)node * Q + 4 * g) = rs;
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
    *reinterpret_cast<float4 *>(dQ_col + (int64_t)node * Q + 4 * g) = cs;
}

// ------------------------------------------------------------------ forward: edge GEMM + epilogue
//   out[e] = act( H[e] W1 + Q_col[col[e]] + Q_row[e / M] )        persistent blocks, double-buffered tiles
template <int K, int Q, bool RELU>
__global__ void __launch_bounds__(GLF_THREADS) glf_edge_out_kernel(const float *__restrict__ H, const int32_t *__restrict__ col,
                                                                    const float *__restrict__ W1,
                                                                    const float *__restrict__ Q_col,
                                                                    const float *__restrict__ Q_row, int64_t c, int M,
                                                                    float *__restrict__ out) {
    constexpr int KS = (K == 3) ? 0 : glf_stride(K);      // K == 3: rows are read straight into registers
    constexpr int QS = glf_stride(Q);
    constexpr bool EARLY = (K + 2 * Q <= 112);            // prefetch the epilogue gathers before the FMA loop
    extern __shared__ __align__(16) float smem[];
    float *Ws = smem;                          // [K][Q]
    float *Os = Ws + ((K * Q + 3) / 4) * 4;    // [TE][QS]
    float *Hs = Os + GLF_TE * QS;              // [2][TE][KS]
    const int tid = threadIdx.x;
    const int ntiles = (int)((c + GLF_TE - 1) / GLF_TE);
    for (int i = tid; i < K * Q; i += GLF_THREADS) Ws[i] = __ldg(&W1[i]);
    int t = blockIdx.x, buf = 0;
    if constexpr (K != 3) {
        if (t < ntiles) glf_stage_tile<K, KS>(Hs, H, (int64_t)t * GLF_TE, c);
        glf_cp_async_commit();
    }
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        if constexpr (K != 3) glf_cp_async_wait0();
        __syncthreads();     // tile t has landed (and Ws on the first pass); the other buffer is free
        if constexpr (K != 3) {
            const int tn = t + gridDim.x;
            if (tn < ntiles) glf_stage_tile<K, KS>(Hs + (buf ^ 1) * GLF_TE * KS, H, (int64_t)tn * GLF_TE, c);
            glf_cp_async_commit();
        }
        const int64_t e0 = (int64_t)t * GLF_TE, e = e0 + tid;
        const bool valid = e < c;
        const float *qc = Q_col + (valid ? (int64_t)__ldg(&col[e]) : 0) * Q;
        const float *qr = Q_row + (valid ? e / M : 0) * Q;
        float acc[Q];
        if constexpr (EARLY) {
#pragma unroll
            for (int j = 0; j < Q / 4; ++j) {
                const float4 a = glf_ldg4(qc + 4 * j), b = glf_ldg4(qr + 4 * j);
                acc[4 * j] = a.x + b.x; acc[4 * j + 1] = a.y + b.y; acc[4 * j + 2] = a.z + b.z; acc[4 * j + 3] = a.w + b.w;
            }
        }
        float h[K];
        if constexpr (K == 3) {
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) h[kk] = valid ? __ldg(&H[e * 3 + kk]) : 0.f;
        } else {
            const float *hrow = Hs + buf * GLF_TE * KS + tid * KS;
#pragma unroll
            for (int j = 0; j < K / 4; ++j) {
                const float4 v = *reinterpret_cast<const float4 *>(hrow + 4 * j);
                h[4 * j] = v.x; h[4 * j + 1] = v.y; h[4 * j + 2] = v.z; h[4 * j + 3] = v.w;
            }
        }
        float z[Q];
#pragma unroll
        for (int qo = 0; qo < Q; ++qo) z[qo] = 0.f;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
#pragma unroll
            for (int j = 0; j < Q / 4; ++j) {
                const float4 w = *reinterpret_cast<const float4 *>(Ws + kk * Q + 4 * j);   // broadcast
                z[4 * j] += h[kk] * w.x; z[4 * j + 1] += h[kk] * w.y; z[4 * j + 2] += h[kk] * w.z; z[4 * j + 3] += h[kk] * w.w;
            }
        }
        if constexpr (!EARLY) {
#pragma unroll
            for (int j = 0; j < Q / 4; ++j) {
                const float4 a = glf_ldg4(qc + 4 * j), b = glf_ldg4(qr + 4 * j);
                acc[4 * j] = a.x + b.x; acc[4 * j + 1] = a.y + b.y; acc[4 * j + 2] = a.z + b.z; acc[4 * j + 3] = a.w + b.w;
            }
        }
#pragma unroll
        for (int j = 0; j < Q / 4; ++j) {
            float4 o = make_float4(z[4 * j] + acc[4 * j], z[4 * j + 1] + acc[4 * j + 1], z[4 * j + 2] + acc[4 * j + 2],
                                   z[4 * j + 3] + acc[4 * j + 3]);
            if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(Os + tid * QS + 4 * j) = o;
        }
        __syncwarp();
        const int w = tid >> 5;
        glf_store_warp_rows<Q, QS>(Os + w * 32 * QS, out, e0 + w * 32, c);
        __syncwarp();
    }
    if constexpr (K != 3) glf_cp_async_wait0();
}

// ------------------------------------------------------------------ X^T Y micro-tile machinery
// The (KP x Q) result is cut into 4x4 micro-tiles.  NT threads form an accumulator set (one or MT
// micro-tiles per thread); NSETS sets split a tile's rows (set s takes rows s, s+NSETS, ...).
template <int KP, int Q>
struct GlfXty {
    static constexpr int KG = KP / 4, QG = Q / 4, NMT = KG * QG;
    static constexpr int MT = (NMT > GLF_THREADS) ? NMT / GLF_THREADS : 1;
    static constexpr int NT = (NMT > GLF_THREADS) ? GLF_THREADS : NMT;
    static constexpr int NSETS = GLF_THREADS / NT;
    static_assert(GLF_THREADS % NT == 0 && NMT % NT == 0, "unsupported shape");
    float acc[MT][4][4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int j = 0; j < MT; ++j)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[j][a][b] = 0.f;
    }
    template <int KS, int QS>
    __device__ __forceinline__ void accumulate(const float *Xs, const float *Ys) {
        const int my_set = threadIdx.x / NT;
#pragma unroll 2
        for (int r = my_set; r < GLF_TE; r += NSETS) {
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const int mt = (threadIdx.x % NT) + j * NT, tk = mt % KG, tq = mt / KG;
                const float4 hv = *reinterpret_cast<const float4 *>(Xs + r * KS + 4 * tk);
                const float4 zv = *reinterpret_cast<const float4 *>(Ys + r * QS + 4 * tq);
                acc[j][0][0] += hv.x * zv.x; acc[j][0][1] += hv.x * zv.y; acc[j][0][2] += hv.x * zv.z; acc[j][0][3] += hv.x * zv.w;
                acc[j][1][0] += hv.y * zv.x; acc[j][1][1] += hv.y * zv.y; acc[j][1][2] += hv.y * zv.z; acc[j][1][3] += hv.y * zv.w;
                acc[j][2][0] += hv.z * zv.x; acc[j][2][1] += hv.z * zv.y; acc[j][2][2] += hv.z * zv.z; acc[j][2][3] += hv.z * zv.w;
                acc[j][3][0] += hv.w * zv.x; acc[j][3][1] += hv.w * zv.y; acc[j][3][2] += hv.w * zv.z; acc[j][3][3] += hv.w * zv.w;
            }
        }
    }
    // reduce the NSETS sets in a fixed order through `red` (>= NSETS*KP*Q floats of smem, all threads
    // must have finished with it) and write rows kk < K of this block's partial [K][Q]
    __device__ __forceinline__ void finish(float *red, int K, float *__restrict__ partial_block) {
        const int my_set = threadIdx.x / NT;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            const int mt = (threadIdx.x % NT) + j * NT, tk = mt % KG, tq = mt / KG;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) red[(my_set * KP + 4 * tk + a) * Q + 4 * tq + b] = acc[j][a][b];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < K * Q; i += GLF_THREADS) {
            float s = 0.f;
            for (int st = 0; st < NSETS; ++st) s += red[st * KP * Q + i];
            partial_block[i] = s;
        }
    }
};

// ------------------------------------------------------------------ backward: edge kernel
//   dZ[e]  = dOut[e] * [Hout[e] > 0]                      (RELU: mask of this layer's own activation)
//   dW1   += H[e]^T dZ[e]                                  (per-block partial, fixed-order reduce later)
//   dH[e]  = dZ[e] W1^T + G_col[col[e]] + G_row[e / M]    (HAS_DH; MASK_IN: times [H[e] > 0], i.e. the
//            ReLU backward of the layer that produced H is applied here, where H is already on chip)
// With HAS_DH = false and RELU = false this is a plain deterministic X^T Y (X = H, Y = dOut) and is
// reused for the node-level dW2 / dW3.
template <int K, int Q, bool RELU, bool HAS_DH, bool MASK_IN>
__device__ __forceinline__ void glf_edge_bwd_body(const float *__restrict__ dOut, const float *__restrict__ Hout,
                                                  const float *__restrict__ H, const int32_t *__restrict__ col,
                                                  const float *__restrict__ W1, const float *__restrict__ G_col,
                                                  const float *__restrict__ G_row, int64_t c, int M, int tiles_per_block,
                                                  float *__restrict__ dH, float *__restrict__ dW_partial) {
    constexpr int KP = (K == 3) ? 4 : K;           // 3-wide rows are zero-padded to 4 for the micro-tiles
    constexpr int KS = glf_stride(KP), QS = glf_stride(Q);
    constexpr int TILE = GLF_TE * (KS + QS + (RELU ? QS : 0));   // floats per pipeline stage: Hs | Zs | (Ms)
    extern __shared__ __align__(16) float smem[];
    float *Wt = smem;                              // [Q][KP]  (W1 transposed: 4 consecutive kk per LDS.128)
    float *stage0 = Wt + Q * KP;
    const int tid = threadIdx.x;
    if constexpr (HAS_DH) {
        for (int i = tid; i < Q * KP; i += GLF_THREADS) {
            const int qo = i / KP, kk = i % KP;
            Wt[i] = (kk < K) ? __ldg(&W1[kk * Q + qo]) : 0.f;
        }
    }
    GlfXty<KP, Q> xty;
    xty.clear();

    const int64_t ntiles = (c + GLF_TE - 1) / GLF_TE;
    const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_block;
    const int64_t t_end = nbpc_min(t_begin + tiles_per_block, ntiles);
    auto issue = [&](int64_t t, int b) {
        float *Hs = stage0 + b * TILE, *Zs = Hs + GLF_TE * KS;
        if constexpr (K == 3) glf_stage_tile3<KS>(Hs, H, t * GLF_TE, c);
        else glf_stage_tile<KP, KS>(Hs, H, t * GLF_TE, c);
        glf_stage_tile<Q, QS>(Zs, dOut, t * GLF_TE, c);
        if constexpr (RELU) glf_stage_tile<Q, QS>(Zs + GLF_TE * QS, Hout, t * GLF_TE, c);
    };
    if (t_begin < t_end) issue(t_begin, 0);
    glf_cp_async_commit();
    int buf = 0;
    for (int64_t t = t_begin; t < t_end; ++t, buf ^= 1) {
        float *Hs = stage0 + buf * TILE, *Zs = Hs + GLF_TE * KS;
        glf_cp_async_wait0();
        if constexpr (RELU) {   // mask in place; every thread masks exactly the chunks it copied itself
            constexpr int CH = Q / 4;
            const float *Ms = Zs + GLF_TE * QS;
            for (int i = tid; i < GLF_TE * CH; i += GLF_THREADS) {
                const int r = i / CH, ch = i % CH;
                const float4 ho = *reinterpret_cast<const float4 *>(Ms + r * QS + 4 * ch);
                float4 *zp = reinterpret_cast<float4 *>(Zs + r * QS + 4 * ch);
                float4 v = *zp;
                v.x = ho.x > 0.f ? v.x : 0.f; v.y = ho.y > 0.f ? v.y : 0.f;
                v.z = ho.z > 0.f ? v.z : 0.f; v.w = ho.w > 0.f ? v.w : 0.f;
                *zp = v;
            }
        }
        __syncthreads();       // tile t complete in smem; the other stage is no longer in use
        if (t + 1 < t_end) issue(t + 1, buf ^ 1);
        glf_cp_async_commit();

        const int64_t e0 = t * GLF_TE, e = e0 + tid;
        float dh[HAS_DH ? KP : 1];
        if constexpr (HAS_DH) {
            static_assert(!HAS_DH || K % 4 == 0, "dH path needs K % 4 == 0");
            const bool valid = e < c;
            const float *gc = G_col + (valid ? (int64_t)__ldg(&col[e]) : 0) * K;
            const float *gr = G_row + (valid ? e / M : 0) * K;
#pragma unroll
            for (int j = 0; j < K / 4; ++j) {   // issue the gathers first: their latency hides behind the FMAs
                const float4 a = glf_ldg4(gc + 4 * j), b = glf_ldg4(gr + 4 * j);
                dh[4 * j] = a.x + b.x; dh[4 * j + 1] = a.y + b.y; dh[4 * j + 2] = a.z + b.z; dh[4 * j + 3] = a.w + b.w;
            }
            float dz[Q];
#pragma unroll
            for (int j = 0; j < Q / 4; ++j) {
                const float4 v = *reinterpret_cast<const float4 *>(Zs + tid * QS + 4 * j);
                dz[4 * j] = v.x; dz[4 * j + 1] = v.y; dz[4 * j + 2] = v.z; dz[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < K / 4; ++j) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int qo = 0; qo < Q; ++qo) {
                    const float4 w = *reinterpret_cast<const float4 *>(Wt + qo * KP + 4 * j);   // broadcast
                    a.x += dz[qo] * w.x; a.y += dz[qo] * w.y; a.z += dz[qo] * w.z; a.w += dz[qo] * w.w;
                }
                dh[4 * j] += a.x; dh[4 * j + 1] += a.y; dh[4 * j + 2] += a.z; dh[4 * j + 3] += a.w;
                if constexpr (MASK_IN) {
                    const float4 hv = *reinterpret_cast<const float4 *>(Hs + tid * KS + 4 * j);
                    dh[4 * j] = hv.x > 0.f ? dh[4 * j] : 0.f; dh[4 * j + 1] = hv.y > 0.f ? dh[4 * j + 1] : 0.f;
                    dh[4 * j + 2] = hv.z > 0.f ? dh[4 * j + 2] : 0.f; dh[4 * j + 3] = hv.w > 0.f ? dh[4 * j + 3] : 0.f;
                }
            }
        }
        xty.template accumulate<KS, QS>(Hs, Zs);
        if constexpr (HAS_DH) {
            __syncthreads();   // everyone is done reading Hs: reuse it as the dH staging area
#pragma unroll
            for (int j = 0; j < K / 4; ++j)
                *reinterpret_cast<float4 *>(Hs + tid * KS + 4 * j) = make_float4(dh[4 * j], dh[4 * j + 1], dh[4 * j + 2], dh[4 * j + 3]);
            __syncwarp();
            const int w = tid >> 5;
            glf_store_warp_rows<KP, KS>(Hs + w * 32 * KS, dH, e0 + w * 32, c);
        }
    }
    glf_cp_async_wait0();
    static_assert(2 * TILE >= GlfXty<KP, Q>::NSETS * KP * Q, "reduction scratch too small");
    xty.finish(stage0, K, dW_partial + (int64_t)blockIdx.x * K * Q);
}
template <int K, int Q, bool RELU, bool HAS_DH, bool MASK_IN>
__global__ void __launch_bounds__(GLF_THREADS) glf_edge_bwd_kernel(const float *__restrict__ dOut, const float *__restrict__ Hout,
                                                                    const float *__restrict__ H, const int32_t *__restrict__ col,
                                                                    const float *__restrict__ W1,
                                                                    const float *__restrict__ G_col,
                                                                    const float *__restrict__ G_row, int64_t c, int M,
                                                                    int tiles_per_block, float *__restrict__ dH,
                                                                    float *__restrict__ dW_partial) {
    glf_edge_bwd_body<K, Q, RELU, HAS_DH, MASK_IN>(dOut, Hout, H, col, W1, G_col, G_row, c, M, tiles_per_block, dH, dW_partial);
}
// TWO independent X^T Y problems of the same shape in one launch (blockIdx.y selects): the node-level dW2 = P_col^T dQ_col and
// dW3 = P_row^T dQ_row of a layer's backward.  Each problem's blocks, partials and summation order are those of two separate
// launches of glf_edge_bwd_kernel<K, Q, false, false, false> (bit-identical); the kernels are latency-bound at low occupancy, so
// the pair takes about the time of one.
template <int K, int Q>
__global__ void __launch_bounds__(GLF_THREADS) glf_xty_pair_kernel(const float *__restrict__ Ya, const float *__restrict__ Xa,
                                                                    float *__restrict__ partial_a, const float *__restrict__ Yb,
                                                                    const float *__restrict__ Xb, float *__restrict__ partial_b,
                                                                    int64_t n, int tiles_per_block) {
    const bool second = blockIdx.y != 0;
    glf_edge_bwd_body<K, Q, false, false, false>(second ? Yb : Ya, nullptr, second ? Xb : Xa, nullptr, nullptr, nullptr, nullptr, n, 1,
                                                 tiles_per_block, nullptr, second ? partial_b : partial_a);
}

// out = sum over blocks of partial[b] (rows x cols), fixed order; transpose: out is (cols x rows).
// block = 32 lanes x 32 warps: 32 consecutive outputs, warp w sums blocks b = w, w+32, ...
static __global__ void __launch_bounds__(1024) glf_partial_reduce_kernel(const float *__restrict__ partial, int nblocks, int rows,
                                                                   int cols, int transpose, float *__restrict__ out) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = rows * cols;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < n)
        for (int b = w; b < nblocks; b += 32) s += partial[(int64_t)b * n + i];
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < n) {
        float tot = 0.f;
#pragma unroll
        for (int ww = 0; ww < 32; ++ww) tot += red[ww][lane];
        const int r = i / cols, cc = i % cols;
        out[transpose ? cc * rows + r : i] = tot;
    }
}

static __global__ void glf_copy_kernel(const float *__restrict__ src, float *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ------------------------------------------------------------------ node-level kernels (runtime k, q)
// persistent grid-stride blocks: the weights are staged in shared memory once per block
// Q_col = P_col W2;  Q_row = P_row W3 + (P_cube W4 + B)
static __global__ void __launch_bounds__(256) glf_node_project_kernel(const float *__restrict__ P_col, const float *__restrict__ P_row,
                                                                const float *__restrict__ P_cube, const float *__restrict__ W,
                                                                const float *__restrict__ bias, int BN, int N, int k, int q,
                                                                float *__restrict__ Q_col, float *__restrict__ Q_row) {
    extern __shared__ __align__(16) float smem[];   // W2, W3, W4: [k][q] each
    for (int i = threadIdx.x; i < 3 * k * q; i += blockDim.x) smem[i] = __ldg(&W[(int64_t)k * q + i]);
    __syncthreads();
    const float *W2 = smem, *W3 = smem + k * q, *W4 = smem + 2 * k * q;
    const int64_t total = (int64_t)BN * q, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int node = (int)(t / q), qo = (int)(t % q);
        const int s = node / N;
        const float *pc = P_col + (int64_t)node * k, *pr = P_row + (int64_t)node * k, *pq = P_cube + s * k;
        float a2 = 0.f, a3 = 0.f, a4 = 0.f;
        for (int kk = 0; kk < k; ++kk) {
            a2 += __ldg(&pc[kk]) * W2[kk * q + qo];
            a3 += __ldg(&pr[kk]) * W3[kk * q + qo];
            a4 += __ldg(&pq[kk]) * W4[kk * q + qo];
        }
        Q_col[t] = a2;
        Q_row[t] = a3 + (a4 + __ldg(&bias[qo]));
    }
}

// G_col = (dQ_col W2^T)/max(indeg,1);  G_row = (dQ_row W3^T)/M + (dCq W4^T)/(N M)
static __global__ void __launch_bounds__(256) glf_node_grad_kernel(const float *__restrict__ dQ_col, const float *__restrict__ dQ_row,
                                                             const float *__restrict__ dCq, const float *__restrict__ W,
                                                             const int32_t *__restrict__ csrT_ptr, int BN, int N, int M,
                                                             int k, int q, float *__restrict__ G_col, float *__restrict__ G_row) {
    extern __shared__ __align__(16) float smem[];   // W2t, W3t, W4t: [q][k] each
    float *W2t = smem, *W3t = smem + k * q, *W4t = smem + 2 * k * q;
    for (int i = threadIdx.x; i < k * q; i += blockDim.x) {
        const int qo = i / k, kk = i % k;
        W2t[i] = __ldg(&W[(int64_t)k * q + kk * q + qo]);
        W3t[i] = __ldg(&W[2 * (int64_t)k * q + kk * q + qo]);
        W4t[i] = __ldg(&W[3 * (int64_t)k * q + kk * q + qo]);
    }
    __syncthreads();
    const int64_t total = (int64_t)BN * k, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int node = (int)(t / k), kk = (int)(t % k);
        const int s = node / N;
        const float *dc = dQ_col + (int64_t)node * q, *dr = dQ_row + (int64_t)node * q, *dq = dCq + s * q;
        float a2 = 0.f, a3 = 0.f, a4 = 0.f;
        for (int qo = 0; qo < q; ++qo) {
            a2 += __ldg(&dc[qo]) * W2t[qo * k + kk];
            a3 += __ldg(&dr[qo]) * W3t[qo * k + kk];
            a4 += __ldg(&dq[qo]) * W4t[qo * k + kk];
        }
        const int indeg = csrT_ptr[node + 1] - csrT_ptr[node];
        G_col[t] = a2 / (float)nbpc_max(indeg, 1);
        G_row[t] = a3 / (float)M + a4 / ((float)N * (float)M);
    }
}

// ---- 4-wide node kernels (k % 4 == 0 and q % 4 == 0): float4 row loads, LDS.128 weights, 4 outputs / thread
// per-sample constants are hoisted into tiny kernels:  Cq[s] = P_cube[s] W4 + bias,  Gq[s] = dCq[s] W4^T / (N M)
static __global__ void glf_cube_project_kernel(const float *__restrict__ P_cube, const float *__restrict__ W4,
                                        const float *__restrict__ bias, int B, int k, int q, float *__restrict__ Cq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * q) return;
    const int s = t / q, qo = t % q;
    float a = 0.f;
    for (int kk = 0; kk < k; ++kk) a += P_cube[s * k + kk] * W4[kk * q + qo];
    Cq[t] = a + bias[qo];
}
static __global__ void glf_cube_grad_kernel(const float *__restrict__ dCq, const float *__restrict__ W4, int B, int N, int M, int k,
                                     int q, float *__restrict__ Gq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * k) return;
    const int s = t / k, kk = t % k;
    float a = 0.f;
    for (int qo = 0; qo < q; ++qo) a += dCq[s * q + qo] * W4[kk * q + qo];
    Gq[t] = a / ((float)N * (float)M);
}

static __global__ void __launch_bounds__(256) glf_node_project4_kernel(const float *__restrict__ P_col, const float *__restrict__ P_row,
                                                                 const float *__restrict__ Cq, const float *__restrict__ W,
                                                                 int BN, int N, int k, int q, float *__restrict__ Q_col,
                                                                 float *__restrict__ Q_row) {
    extern __shared__ __align__(16) float smem[];   // W2, W3: [k][q] each
    for (int i = threadIdx.x; i < 2 * k * q; i += blockDim.x) smem[i] = __ldg(&W[(int64_t)k * q + i]);
    __syncthreads();
    const float *W2 = smem, *W3 = smem + k * q;
    const int QG = q / 4;
    const int64_t total = (int64_t)BN * QG, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int node = (int)(t / QG), j = (int)(t % QG);
        const float *pc = P_col + (int64_t)node * k, *pr = P_row + (int64_t)node * k;
        float4 a2 = make_float4(0.f, 0.f, 0.f, 0.f), a3 = a2;
        for (int k4 = 0; k4 < k; k4 += 4) {
            const float4 c4 = glf_ldg4(pc + k4), r4 = glf_ldg4(pr + k4);
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 w2 = *reinterpret_cast<const float4 *>(W2 + (k4 + i) * q + 4 * j);
                const float4 w3 = *reinterpret_cast<const float4 *>(W3 + (k4 + i) * q + 4 * j);
                a2.x += cv[i] * w2.x; a2.y += cv[i] * w2.y; a2.z += cv[i] * w2.z; a2.w += cv[i] * w2.w;
                a3.x += rv[i] * w3.x; a3.y += rv[i] * w3.y; a3.z += rv[i] * w3.z; a3.w += rv[i] * w3.w;
            }
        }
        const float4 cq = glf_ldg4(Cq + (node / N) * q + 4 * j);
        *reinterpret_cast<float4 *>(Q_col + (int64_t)node * q + 4 * j) = a2;
        *reinterpret_cast<float4 *>(Q_row + (int64_t)node * q + 4 * j) = make_float4(a3.x + cq.x, a3.y + cq.y, a3.z + cq.z, a3.w + cq.w);
    }
}

static __global__ void __launch_bounds__(256) glf_node_grad4_kernel(const float *__restrict__ dQ_col, const float *__restrict__ dQ_row,
                                                              const float *__restrict__ Gq, const float *__restrict__ W,
                                                              const int32_t *__restrict__ csrT_ptr, int BN, int N, int M,
                                                              int k, int q, float *__restrict__ G_col, float *__restrict__ G_row) {
    extern __shared__ __align__(16) float smem[];   // W2t, W3t: [q][k] each
    float *W2t = smem, *W3t = smem + k * q;
    for (int i = threadIdx.x; i < k * q; i += blockDim.x) {
        const int qo = i / k, kk = i % k;
        W2t[i] = __ldg(&W[(int64_t)k * q + kk * q + qo]);
        W3t[i] = __ldg(&W[2 * (int64_t)k * q + kk * q + qo]);
    }
    __syncthreads();
    const int KG = k / 4;
    const int64_t total = (int64_t)BN * KG, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int node = (int)(t / KG), j = (int)(t % KG);
        const float *dc = dQ_col + (int64_t)node * q, *dr = dQ_row + (int64_t)node * q;
        float4 a2 = make_float4(0.f, 0.f, 0.f, 0.f), a3 = a2;
        for (int q4 = 0; q4 < q; q4 += 4) {
            const float4 c4 = glf_ldg4(dc + q4), r4 = glf_ldg4(dr + q4);
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 w2 = *reinterpret_cast<const float4 *>(W2t + (q4 + i) * k + 4 * j);
                const float4 w3 = *reinterpret_cast<const float4 *>(W3t + (q4 + i) * k + 4 * j);
                a2.x += cv[i] * w2.x; a2.y += cv[i] * w2.y; a2.z += cv[i] * w2.z; a2.w += cv[i] * w2.w;
                a3.x += rv[i] * w3.x; a3.y += rv[i] * w3.y; a3.z += rv[i] * w3.z; a3.w += rv[i] * w3.w;
            }
        }
        const float fi = (float)nbpc_max(csrT_ptr[node + 1] - csrT_ptr[node], 1), fm = (float)M;
        const float4 gq = glf_ldg4(Gq + (node / N) * k + 4 * j);
        *reinterpret_cast<float4 *>(G_col + (int64_t)node * k + 4 * j) = make_float4(a2.x / fi, a2.y / fi, a2.z / fi, a2.w / fi);
        *reinterpret_cast<float4 *>(G_row + (int64_t)node * k + 4 * j) =
            make_float4(a3.x / fm + gq.x, a3.y / fm + gq.y, a3.z / fm + gq.z, a3.w / fm + gq.w);
    }
}

// ---- per-sample column sums (cube pool / dCq): block = 32 lanes (channels) x 8 warps (row slices)
// partial[s][blk][ch] = sum of rows [blk*rpb, (blk+1)*rpb) of sample s;   grid (nblk, B)
static __global__ void __launch_bounds__(256) glf_colsum_partial_kernel(const float *__restrict__ X, int ch, int N, int rpb,
                                                                  float *__restrict__ partial) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int blk = blockIdx.x, s = blockIdx.y, nblk = gridDim.x;
    const int r0 = blk * rpb, r1 = nbpc_min(r0 + rpb, N);
    for (int c0 = 0; c0 < ch; c0 += 32) {
        const int cc = c0 + lane;
        float a0 = 0.f, a1 = 0.f;
        if (cc < ch) {
            int r = r0 + w;
            for (; r + 8 < r1; r += 16) {
                a0 += __ldg(&X[((int64_t)s * N + r) * ch + cc]);
                a1 += __ldg(&X[((int64_t)s * N + r + 8) * ch + cc]);
            }
            if (r < r1) a0 += __ldg(&X[((int64_t)s * N + r) * ch + cc]);
        }
        red[w][lane] = a0 + a1;
        __syncthreads();
        if (w == 0 && cc < ch) {
            float tot = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) tot += red[ww][lane];
            partial[((int64_t)s * nblk + blk) * ch + cc] = tot;
        }
        __syncthreads();
    }
}
// out[s][ch] = (sum_blk partial[s][blk][ch]) / divisor;   grid (B)
static __global__ void __launch_bounds__(256) glf_colsum_final_kernel(const float *__restrict__ partial, int ch, int nblk, float divisor,
                                                                float *__restrict__ out) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s = blockIdx.x;
    for (int c0 = 0; c0 < ch; c0 += 32) {
        const int cc = c0 + lane;
        float a = 0.f;
        if (cc < ch)
            for (int b = w; b < nblk; b += 8) a += partial[((int64_t)s * nblk + b) * ch + cc];
        red[w][lane] = a;
        __syncthreads();
        if (w == 0 && cc < ch) {
            float tot = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) tot += red[ww][lane];
            out[s * ch + cc] = tot / divisor;
        }
        __syncthreads();
    }
}

// generic X^T Y over n node rows (runtime k, q; used when no micro-tile instance fits)
#define GLF_XTY_ROWS 32
static __global__ void __launch_bounds__(256) glf_node_xty_kernel(const float *__restrict__ X, const float *__restrict__ Y, int64_t n,
                                                            int rows_per_block, int k, int q, float *__restrict__ partial) {
    extern __shared__ __align__(16) float smem[];   // Xs [ROWS][k], Ys [ROWS][q]
    float *Xs = smem, *Ys = smem + GLF_XTY_ROWS * k;
    const int kq = k * q;
    constexpr int MAXP = 16;                     // supports k*q <= 4096
    float acc[MAXP];
#pragma unroll
    for (int j = 0; j < MAXP; ++j) acc[j] = 0.f;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = nbpc_min(r_begin + rows_per_block, n);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += GLF_XTY_ROWS) {
        const int nr = (int)nbpc_min((int64_t)GLF_XTY_ROWS, r_end - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * k; i += blockDim.x) Xs[i] = __ldg(&X[r0 * k + i]);
        for (int i = threadIdx.x; i < nr * q; i += blockDim.x) Ys[i] = __ldg(&Y[r0 * q + i]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < MAXP; ++j) {
            const int p = threadIdx.x + j * 256;
            if (p < kq) {
                const int kk = p / q, qo = p % q;
                float a = acc[j];
                for (int r = 0; r < nr; ++r) a += Xs[r * k + kk] * Ys[r * q + qo];
                acc[j] = a;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
        const int p = threadIdx.x + j * 256;
        if (p < kq) partial[(int64_t)blockIdx.x * kq + p] = acc[j];
    }
}

// ------------------------------------------------------------------ last layer, node level
// out[i] = act( P_row[i] W1 + (1/M) sum_m Q_col[col[iM+m]] + Q_row[i] )  ==  row-mean of Z (graph.py:455)
static __global__ void __launch_bounds__(256) glf_last_out_kernel(const float *__restrict__ P_row, const int32_t *__restrict__ col,
                                                            const float *__restrict__ W1, const float *__restrict__ Q_col,
                                                            const float *__restrict__ Q_row, int BN, int M, int k, int q,
                                                            int relu, float *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * q) return;
    const int node = (int)(t / q), qo = (int)(t % q);
    float z = 0.f;
    for (int kk = 0; kk < k; ++kk) z += __ldg(&P_row[(int64_t)node * k + kk]) * __ldg(&W1[kk * q + qo]);
    float g = 0.f;
    for (int m = 0; m < M; ++m) g += __ldg(&Q_col[(int64_t)__ldg(&col[(int64_t)node * M + m]) * q + qo]);
    z += g / (float)M + Q_row[t];
    out[t] = (relu && z < 0.f) ? 0.f : z;
}

// last-layer backward, node level.  dOutM = dOut * [out > 0] (relu).
//   dQ_row[i] = dOutM[i];  dQ_col[j] = (1/M) sum_{e in csrT[j]} dOutM[e / M]
static __global__ void __launch_bounds__(256) glf_last_bwd_pool_kernel(const float *__restrict__ dOut, const float *__restrict__ Hout,
                                                                 int relu, int BN, int M, int q,
                                                                 const int32_t *__restrict__ csrT_ptr,
                                                                 const int32_t *__restrict__ csrT_edge,
                                                                 float *__restrict__ dQ_row, float *__restrict__ dQ_col) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * q) return;
    const int node = (int)(t / q), qo = (int)(t % q);
    auto dzm = [&](int64_t i) {
        float v = __ldg(&dOut[i * q + qo]);
        if (relu && !(__ldg(&Hout[i * q + qo]) > 0.f)) v = 0.f;
        return v;
    };
    dQ_row[t] = dzm(node);
    const int b = csrT_ptr[node], e = csrT_ptr[node + 1];
    float cs = 0.f;
    for (int p = b; p < e; ++p) cs += dzm(__ldg(&csrT_edge[p]) / M);
    dQ_col[t] = cs / (float)M;
}

// R[i] = (dOutM[i] W1^T)/M + G_row[i]   (in place on G_row)
static __global__ void __launch_bounds__(256) glf_last_rowterm_kernel(const float *__restrict__ dQ_row, const float *__restrict__ W1,
                                                                int BN, int M, int k, int q, float *__restrict__ G_row) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * k) return;
    const int node = (int)(t / k), kk = (int)(t % k);
    float a = 0.f;
    for (int qo = 0; qo < q; ++qo) a += __ldg(&dQ_row[(int64_t)node * q + qo]) * __ldg(&W1[kk * q + qo]);
    G_row[t] = a / (float)M + G_row[t];
}

// dH[e] = (R[e / M] + G_col[col[e]]) [* (H[e] > 0) if Hmask]     thread per (edge, 4-channel group); k % 4 == 0
static __global__ void __launch_bounds__(256) glf_last_edge_in_kernel(const int32_t *__restrict__ col, const float *__restrict__ R,
                                                                const float *__restrict__ G_col,
                                                                const float *__restrict__ Hmask, int64_t c, int M, int k,
                                                                float *__restrict__ dH) {
    const int G = k / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c * G) return;
    const int64_t e = t / G;
    const int g = (int)(t % G);
    const float4 r = glf_ldg4(R + (e / M) * k + 4 * g);
    const float4 gc = glf_ldg4(G_col + (int64_t)__ldg(&col[e]) * k + 4 * g);
    float4 o = make_float4(r.x + gc.x, r.y + gc.y, r.z + gc.z, r.w + gc.w);
    if (Hmask) {
        const float4 h = glf_ldg4(Hmask + e * k + 4 * g);
        o.x = h.x > 0.f ? o.x : 0.f; o.y = h.y > 0.f ? o.y : 0.f; o.z = h.z > 0.f ? o.z : 0.f; o.w = h.w > 0.f ? o.w : 0.f;
    }
    *reinterpret_cast<float4 *>(dH + e * k + 4 * g) = o;
}

// Same dH, plus the row sums the PREVIOUS layer's backward pools first: dQ_row_prev[i] = sum_m dH[i M + m] (ascending m, the
// order of gln_bwd_pool_kernel: bit-identical hand-over).  Thread per (row node, 4-channel group), two edges in flight;
// R[i] is loaded once per thread.
static __global__ void __launch_bounds__(256) glf_last_edge_in_rowsum_kernel(const int32_t *__restrict__ col, const float *__restrict__ R,
                                                                       const float *__restrict__ G_col,
                                                                       const float *__restrict__ Hmask, int64_t n_rows, int M, int k,
                                                                       float *__restrict__ dH, float *__restrict__ dQ_row_prev) {
    const int G = k / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * G) return;
    const int64_t row = t / G;
    const int g = (int)(t % G);
    const float4 r = glf_ldg4(R + row * k + 4 * g);
    const int64_t e0 = row * M;
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    auto emit = [&](int64_t e, const float4 &gc, const float4 &h) {
        float4 o = make_float4(r.x + gc.x, r.y + gc.y, r.z + gc.z, r.w + gc.w);
        if (Hmask) { o.x = h.x > 0.f ? o.x : 0.f; o.y = h.y > 0.f ? o.y : 0.f; o.z = h.z > 0.f ? o.z : 0.f; o.w = h.w > 0.f ? o.w : 0.f; }
        *reinterpret_cast<float4 *>(dH + e * k + 4 * g) = o;
        rs.x += o.x; rs.y += o.y; rs.z += o.z; rs.w += o.w;
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int m = 0;
    for (; m + 2 <= M; m += 2) {
        const int64_t ea = e0 + m, eb = ea + 1;
        const int ca = __ldg(&col[ea]), cb = __ldg(&col[eb]);
        const float4 ga = glf_ldg4(G_col + (int64_t)ca * k + 4 * g), gb = glf_ldg4(G_col + (int64_t)cb * k + 4 * g);
        const float4 ha = Hmask ? glf_ldg4(Hmask + ea * k + 4 * g) : zero, hb = Hmask ? glf_ldg4(Hmask + eb * k + 4 * g) : zero;
        emit(ea, ga, ha);
        emit(eb, gb, hb);
    }
    for (; m < M; ++m) {
        const int64_t e = e0 + m;
        emit(e, glf_ldg4(G_col + (int64_t)__ldg(&col[e]) * k + 4 * g), Hmask ? glf_ldg4(Hmask + e * k + 4 * g) : zero);
    }
    *reinterpret_cast<float4 *>(dQ_row_prev + row * k + 4 * g) = rs;
}

#endif  // !NBPC_HOST_EMU
