// graph_layer_fast.cuh - tiled sm_100a kernels for the shift-invariant graph layer.
//
// Same math and the same fixed summation orders as the baseline kernels in graph_layer.cu, but laid
// out for HBM throughput:
//   * edge tiles of 128 consecutive edges are staged into shared memory with 16-byte cp.async
//     copies (fully coalesced 512 B per warp instruction) into rows padded so that a thread reading
//     its own row with LDS.128 is bank-conflict free;
//   * the edge-level GEMM  Z = H W1  runs thread-per-edge-row with the row in registers and W1 read
//     from shared memory as broadcast LDS.128 (4 FMAs per shared load, no divergence);
//   * outputs leave through a per-warp shared staging buffer so that every global store instruction
//     writes 512 contiguous bytes;
//   * pooling gathers use float4 channel groups (a 128 B edge row is read by 8 adjacent lanes);
//   * dW1 = H^T dZ is accumulated per block in registers with 4x4 micro-tiles and reduced over
//     blocks in a fixed order (no float atomics => bit-reproducible).
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
__device__ __forceinline__ void glf_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
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
    for (; p + 2 <= e; p += 2) {
        const int e0 = __ldg(&csrT_edge[p]), e1 = __ldg(&csrT_edge[p + 1]);
        const float4 v0 = dz(e0), v1 = dz(e1);
        cs.x += v0.x; cs.y += v0.y; cs.z += v0.z; cs.w += v0.w;
        cs.x += v1.x; cs.y += v1.y; cs.z += v1.z; cs.w += v1.w;
    }
    for (; p < e; ++p) {
        const float4 v = dz(__ldg(&csrT_edge[p]));
        cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
    *reinterpret_cast<float4 *>(dQ_col + (int64_t)node * Q + 4 * g) = cs;
}

// ------------------------------------------------------------------ forward: edge GEMM + epilogue
//   out[e] = act( H[e] W1 + Q_col[col[e]] + Q_row[e / M] )
template <int K, int Q, bool RELU>
__global__ void __launch_bounds__(GLF_THREADS) glf_edge_out_kernel(const float *__restrict__ H, const int32_t *__restrict__ col,
                                                                    const float *__restrict__ W1,
                                                                    const float *__restrict__ Q_col,
                                                                    const float *__restrict__ Q_row, int64_t c, int M,
                                                                    float *__restrict__ out) {
    constexpr int KS = (K == 3) ? 3 : glf_stride(K);
    constexpr int QS = glf_stride(Q);
    extern __shared__ __align__(16) float smem[];
    float *Ws = smem;                         // [K][Q]
    float *Hs = Ws + K * Q;                   // [TE][KS]   (K == 3: dense, scalar reads)
    float *Os = Hs + GLF_TE * KS;             // [TE][QS]   (GLF_TE * KS and K * Q are multiples of 4)
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * GLF_TE;
    for (int i = tid; i < K * Q; i += GLF_THREADS) Ws[i] = __ldg(&W1[i]);
    if constexpr (K == 3) {
        for (int i = tid; i < GLF_TE * 3; i += GLF_THREADS) Hs[i] = (e0 * 3 + i < c * 3) ? __ldg(&H[e0 * 3 + i]) : 0.f;
    } else {
        glf_stage_tile<K, KS>(Hs, H, e0, c);
        glf_cp_async_wait_all();
    }
    __syncthreads();

    const int64_t e = e0 + tid;
    float h[K];
    if constexpr (K == 3) {
#pragma unroll
        for (int kk = 0; kk < K; ++kk) h[kk] = Hs[tid * 3 + kk];
    } else {
#pragma unroll
        for (int j = 0; j < K / 4; ++j) {
            const float4 v = *reinterpret_cast<const float4 *>(Hs + tid * KS + 4 * j);
            h[4 * j] = v.x; h[4 * j + 1] = v.y; h[4 * j + 2] = v.z; h[4 * j + 3] = v.w;
        }
    }
    float acc[Q];
#pragma unroll
    for (int qo = 0; qo < Q; ++qo) acc[qo] = 0.f;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
#pragma unroll
        for (int j = 0; j < Q / 4; ++j) {
            const float4 w = *reinterpret_cast<const float4 *>(Ws + kk * Q + 4 * j);   // broadcast
            acc[4 * j] += h[kk] * w.x; acc[4 * j + 1] += h[kk] * w.y;
            acc[4 * j + 2] += h[kk] * w.z; acc[4 * j + 3] += h[kk] * w.w;
        }
    }
    if (e < c) {
        const float *qc = Q_col + (int64_t)__ldg(&col[e]) * Q;
        const float *qr = Q_row + (e / M) * Q;
#pragma unroll
        for (int j = 0; j < Q / 4; ++j) {
            const float4 a = glf_ldg4(qc + 4 * j), b = glf_ldg4(qr + 4 * j);
            float4 z = make_float4(acc[4 * j] + (a.x + b.x), acc[4 * j + 1] + (a.y + b.y), acc[4 * j + 2] + (a.z + b.z),
                                   acc[4 * j + 3] + (a.w + b.w));
            if (RELU) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
            *reinterpret_cast<float4 *>(Os + tid * QS + 4 * j) = z;
        }
    }
    __syncwarp();
    const int w = tid >> 5;
    glf_store_warp_rows<Q, QS>(Os + w * 32 * QS, out, e0 + w * 32, c);
}

// ------------------------------------------------------------------ backward: edge kernel
//   dZ[e]  = dOut[e] * [Hout[e] > 0]                      (RELU: mask of this layer's activation)
//   dH[e]  = dZ[e] W1^T + G_col[col[e]] + G_row[e / M]    (HAS_DH)
//   dW1   += H[e]^T dZ[e]                                  (per-block partial, fixed-order reduce later)
template <int K, int Q, bool RELU, bool HAS_DH>
__global__ void __launch_bounds__(GLF_THREADS) glf_edge_bwd_kernel(const float *__restrict__ dOut, const float *__restrict__ Hout,
                                                                    const float *__restrict__ H, const int32_t *__restrict__ col,
                                                                    const float *__restrict__ W1,
                                                                    const float *__restrict__ G_col,
                                                                    const float *__restrict__ G_row, int64_t c, int M,
                                                                    int tiles_per_block, float *__restrict__ dH,
                                                                    float *__restrict__ dW_partial) {
    constexpr int KP = (K == 3) ? 4 : K;           // H tile rows are zero-padded to a multiple of 4 for the micro-tiles
    constexpr int KS = glf_stride(KP), QS = glf_stride(Q);
    constexpr int KG = KP / 4, QG = Q / 4;         // 4x4 micro-tiles of dW1
    constexpr int NMT = KG * QG;                   // micro-tiles of the (K x Q) result
    constexpr int MT = (NMT > GLF_THREADS) ? NMT / GLF_THREADS : 1;     // micro-tiles per thread
    constexpr int NT = (NMT > GLF_THREADS) ? GLF_THREADS : NMT;         // threads per accumulator set
    static_assert(GLF_THREADS % NT == 0 && NMT % NT == 0, "unsupported shape");
    constexpr int NSETS = GLF_THREADS / NT;        // sets split the tile's edges (set s: edges s, s+NSETS, ..)
    extern __shared__ __align__(16) float smem[];
    float *Wt = smem;                              // [Q][KP]  (W1 transposed: 4 consecutive kk per LDS.128)
    float *Hs = Wt + Q * KP;                       // [TE][KS]
    float *Zs = Hs + GLF_TE * KS;                  // [TE][QS]  masked dZ
    float *Ds = Zs + GLF_TE * QS;                  // [TE][KS]  dH staging (HAS_DH) / cross-set reduction scratch
    const int tid = threadIdx.x;
    for (int i = tid; i < Q * KP; i += GLF_THREADS) {
        const int qo = i / KP, kk = i % KP;
        Wt[i] = (kk < K) ? __ldg(&W1[kk * Q + qo]) : 0.f;
    }
    const int my_set = tid / NT;
    float wacc[MT][4][4];
#pragma unroll
    for (int j = 0; j < MT; ++j)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) wacc[j][a][b] = 0.f;

    for (int it = 0; it < tiles_per_block; ++it) {
        const int64_t e0 = ((int64_t)blockIdx.x * tiles_per_block + it) * GLF_TE;
        if (e0 >= c) break;
        __syncthreads();   // previous tile fully consumed
        // ---- stage H tile and dOut tile
        if constexpr (K == 3) {
            for (int i = tid; i < GLF_TE * KP; i += GLF_THREADS) {
                const int r = i / KP, kk = i % KP;
                Hs[r * KS + kk] = (kk < 3 && e0 + r < c) ? __ldg(&H[(e0 + r) * 3 + kk]) : 0.f;
            }
        } else {
            glf_stage_tile<KP, KS>(Hs, H, e0, c);
        }
        glf_stage_tile<Q, QS>(Zs, dOut, e0, c);
        glf_cp_async_wait_all();
        if constexpr (RELU) {   // mask own chunks in place (each thread masks exactly the chunks it staged)
            constexpr int CH = Q / 4;
            for (int i = tid; i < GLF_TE * CH; i += GLF_THREADS) {
                const int r = i / CH, ch = i % CH;
                if (e0 + r < c) {
                    const float4 ho = glf_ldg4(Hout + (e0 + r) * Q + 4 * ch);
                    float4 *z = reinterpret_cast<float4 *>(Zs + r * QS + 4 * ch);
                    float4 v = *z;
                    v.x = ho.x > 0.f ? v.x : 0.f; v.y = ho.y > 0.f ? v.y : 0.f;
                    v.z = ho.z > 0.f ? v.z : 0.f; v.w = ho.w > 0.f ? v.w : 0.f;
                    *z = v;
                }
            }
        }
        __syncthreads();

        // ---- dH: thread per edge row
        if constexpr (HAS_DH) {
            static_assert(K % 4 == 0, "dH path needs K % 4 == 0");
            const int64_t e = e0 + tid;
            float dz[Q];
#pragma unroll
            for (int j = 0; j < Q / 4; ++j) {
                const float4 v = *reinterpret_cast<const float4 *>(Zs + tid * QS + 4 * j);
                dz[4 * j] = v.x; dz[4 * j + 1] = v.y; dz[4 * j + 2] = v.z; dz[4 * j + 3] = v.w;
            }
            const int64_t ce = (e < c) ? (int64_t)__ldg(&col[e]) : 0;
            const int64_t re = (e < c) ? e / M : 0;
#pragma unroll
            for (int j = 0; j < K / 4; ++j) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int qo = 0; qo < Q; ++qo) {
                    const float4 w = *reinterpret_cast<const float4 *>(Wt + qo * KP + 4 * j);   // broadcast
                    a.x += dz[qo] * w.x; a.y += dz[qo] * w.y; a.z += dz[qo] * w.z; a.w += dz[qo] * w.w;
                }
                const float4 gc = glf_ldg4(G_col + ce * K + 4 * j), gr = glf_ldg4(G_row + re * K + 4 * j);
                a.x += gc.x + gr.x; a.y += gc.y + gr.y; a.z += gc.z + gr.z; a.w += gc.w + gr.w;
                *reinterpret_cast<float4 *>(Ds + tid * KS + 4 * j) = a;
            }
            __syncwarp();
            const int w = tid >> 5;
            glf_store_warp_rows<KP, KS>(Ds + w * 32 * KS, dH, e0 + w * 32, c);
        }

        // ---- dW1 micro-tiles: set s takes edges s, s + NSETS, ...
#pragma unroll 2
        for (int r = my_set; r < GLF_TE; r += NSETS) {
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const int mt = (tid % NT) + j * NT, tk = mt % KG, tq = mt / KG;
                const float4 hv = *reinterpret_cast<const float4 *>(Hs + r * KS + 4 * tk);
                const float4 zv = *reinterpret_cast<const float4 *>(Zs + r * QS + 4 * tq);
                wacc[j][0][0] += hv.x * zv.x; wacc[j][0][1] += hv.x * zv.y; wacc[j][0][2] += hv.x * zv.z; wacc[j][0][3] += hv.x * zv.w;
                wacc[j][1][0] += hv.y * zv.x; wacc[j][1][1] += hv.y * zv.y; wacc[j][1][2] += hv.y * zv.z; wacc[j][1][3] += hv.y * zv.w;
                wacc[j][2][0] += hv.z * zv.x; wacc[j][2][1] += hv.z * zv.y; wacc[j][2][2] += hv.z * zv.z; wacc[j][2][3] += hv.z * zv.w;
                wacc[j][3][0] += hv.w * zv.x; wacc[j][3][1] += hv.w * zv.y; wacc[j][3][2] += hv.w * zv.z; wacc[j][3][3] += hv.w * zv.w;
            }
        }
    }

    // ---- reduce the NSETS accumulator sets in a fixed order, write this block's partial [K][Q]
    __syncthreads();
    float *red = Hs;   // NSETS * KP * Q floats, spans the contiguous Hs|Zs region
    static_assert(GLF_TE * (KS + QS) >= NSETS * KP * Q, "reduction scratch too small");
#pragma unroll
    for (int j = 0; j < MT; ++j) {
        const int mt = (tid % NT) + j * NT, tk = mt % KG, tq = mt / KG;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) red[(my_set * KP + 4 * tk + a) * Q + 4 * tq + b] = wacc[j][a][b];
    }
    __syncthreads();
    for (int i = tid; i < K * Q; i += GLF_THREADS) {
        float s = 0.f;
        for (int st = 0; st < NSETS; ++st) s += red[st * KP * Q + i];   // i = kk*Q + qo with kk < K <= KP
        dW_partial[(int64_t)blockIdx.x * K * Q + i] = s;
    }
}

__global__ void glf_copy_kernel(const float *__restrict__ src, float *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// out[i] = sum_b partial[b][i], fixed order
__global__ void glf_partial_reduce_kernel(const float *__restrict__ partial, int nblocks, int n, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int b = 0;
    for (; b + 4 <= nblocks; b += 4) {
        s0 += partial[(int64_t)b * n + i]; s1 += partial[(int64_t)(b + 1) * n + i];
        s2 += partial[(int64_t)(b + 2) * n + i]; s3 += partial[(int64_t)(b + 3) * n + i];
    }
    for (; b < nblocks; ++b) s0 += partial[(int64_t)b * n + i];
    out[i] = (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------ node-level kernels (runtime k, q)
// G_col = (dQ_col W2^T)/max(indeg,1);  G_row = (dQ_row W3^T)/M + (dCq W4^T)/(N M)
__global__ void __launch_bounds__(256) glf_node_grad_kernel(const float *__restrict__ dQ_col, const float *__restrict__ dQ_row,
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
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * k) return;
    const int node = (int)(t / k), kk = (int)(t % k);
    const int s = node / N;
    float a2 = 0.f, a3 = 0.f, a4 = 0.f;
    for (int qo = 0; qo < q; ++qo) {
        a2 += __ldg(&dQ_col[(int64_t)node * q + qo]) * W2t[qo * k + kk];
        a3 += __ldg(&dQ_row[(int64_t)node * q + qo]) * W3t[qo * k + kk];
        a4 += __ldg(&dCq[s * q + qo]) * W4t[qo * k + kk];
    }
    const int indeg = csrT_ptr[node + 1] - csrT_ptr[node];
    G_col[t] = a2 / (float)nbpc_max(indeg, 1);
    G_row[t] = a3 / (float)M + a4 / ((float)N * (float)M);
}

// Q_col = P_col W2;  Q_row = P_row W3 + (P_cube W4 + B)        thread per (node, qo), W in smem
__global__ void __launch_bounds__(256) glf_node_project_kernel(const float *__restrict__ P_col, const float *__restrict__ P_row,
                                                                const float *__restrict__ P_cube, const float *__restrict__ W,
                                                                const float *__restrict__ bias, int BN, int N, int k, int q,
                                                                float *__restrict__ Q_col, float *__restrict__ Q_row) {
    extern __shared__ __align__(16) float smem[];   // W2, W3, W4: [k][q] each
    for (int i = threadIdx.x; i < 3 * k * q; i += blockDim.x) smem[i] = __ldg(&W[(int64_t)k * q + i]);
    __syncthreads();
    const float *W2 = smem, *W3 = smem + k * q, *W4 = smem + 2 * k * q;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * q) return;
    const int node = (int)(t / q), qo = (int)(t % q);
    const int s = node / N;
    float a2 = 0.f, a3 = 0.f, a4 = 0.f;
    for (int kk = 0; kk < k; ++kk) {
        a2 += __ldg(&P_col[(int64_t)node * k + kk]) * W2[kk * q + qo];
        a3 += __ldg(&P_row[(int64_t)node * k + kk]) * W3[kk * q + qo];
        a4 += __ldg(&P_cube[s * k + kk]) * W4[kk * q + qo];
    }
    Q_col[t] = a2;
    Q_row[t] = a3 + (a4 + __ldg(&bias[qo]));
}

// X^T Y over n node rows (runtime k, q): block handles a contiguous chunk of rows staged through smem,
// thread owns pairs p = tid, tid + 256, ... of the (k x q) result; per-block partial, fixed-order reduce later
#define GLF_XTY_ROWS 32
__global__ void __launch_bounds__(256) glf_node_xty_kernel(const float *__restrict__ X, const float *__restrict__ Y, int64_t n,
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
__global__ void __launch_bounds__(256) glf_last_out_kernel(const float *__restrict__ P_row, const int32_t *__restrict__ col,
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
__global__ void __launch_bounds__(256) glf_last_bwd_pool_kernel(const float *__restrict__ dOut, const float *__restrict__ Hout,
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
__global__ void __launch_bounds__(256) glf_last_rowterm_kernel(const float *__restrict__ dQ_row, const float *__restrict__ W1,
                                                                int BN, int M, int k, int q, float *__restrict__ G_row) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)BN * k) return;
    const int node = (int)(t / k), kk = (int)(t % k);
    float a = 0.f;
    for (int qo = 0; qo < q; ++qo) a += __ldg(&dQ_row[(int64_t)node * q + qo]) * __ldg(&W1[kk * q + qo]);
    G_row[t] = a / (float)M + G_row[t];
}

// dH[e] = R[e / M] + G_col[col[e]]      thread per (edge, 4-channel group); k % 4 == 0
__global__ void __launch_bounds__(256) glf_last_edge_in_kernel(const int32_t *__restrict__ col, const float *__restrict__ R,
                                                                const float *__restrict__ G_col, int64_t c, int M, int k,
                                                                float *__restrict__ dH) {
    const int G = k / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c * G) return;
    const int64_t e = t / G;
    const int g = (int)(t % G);
    const float4 r = glf_ldg4(R + (e / M) * k + 4 * g);
    const float4 gc = glf_ldg4(G_col + (int64_t)__ldg(&col[e]) * k + 4 * g);
    *reinterpret_cast<float4 *>(dH + e * k + 4 * g) = make_float4(r.x + gc.x, r.y + gc.y, r.z + gc.z, r.w + gc.w);
}

#endif  // !NBPC_HOST_EMU
