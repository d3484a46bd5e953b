// graph_layer_tc_bwd.cu - tcgen05 backward edge kernel (see graph_layer_tc.cuh)
#include "graph_layer_tc.cuh"
#ifndef NBPC_HOST_EMU
// ------------------------------------------------------------------ backward kernel
//
// dW1 = H^T dZ contracts over the edges, i.e. over the ROWS of both tiles: both operands are MN-major.  For 32-bit
// operands tcgen05 accepts an MN-major shared-memory operand only in the "128-byte swizzle, 32-byte atom" layout
// (descriptor layout type 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of exactly 128 bytes, 4 rows per swizzle
// atom.  Tensors with 16 channels are therefore viewed as (c/2, 32) - P = 2 edges per 128-byte "packed row" - and so
// is the other operand, (c/2, 2K): the MMA then yields the (P K) x (P Q) matrix [H_even H_odd]^T [Z_even Z_odd] whose
// P diagonal K x Q blocks sum to the tile's dW1 (the off-diagonal blocks are ignored; the tensor pipe is idle anyway).
// dH = dZ W1^T needs dZ K-major (standard swizzle), so the dZ tile is landed twice (the second copy comes from L2).
template <int K, int Q>
struct GltBwdCfg {
    using TZ = GltTile<Q>;
    static constexpr int P = (K == 16 || Q == 16) ? 2 : 1;               // edges per packed row
    static constexpr int R = GLT_TILE / P;                               // packed rows per tile
    static constexpr int HCH = K * P / 32, ZCH = Q * P / 32;             // 32-float column chunks of the packed tiles
    static constexpr int PCHUNK = R * 128;                               // bytes per packed chunk
    static constexpr int H_BYTES = HCH * PCHUNK;                         // packed H            (MN-major operand of D2)
    static constexpr int Z_BYTES = TZ::NCH * TZ::chunk_bytes(GLT_TILE);  // dZ, K-major         (A operand of D1)
    static constexpr int Z2_BYTES = ZCH * PCHUNK;                        // packed dZ           (MN-major operand of D2)
    static constexpr int STAGE = H_BYTES + Z_BYTES + Z2_BYTES;
    static constexpr int B_BYTES = TZ::NCH * TZ::chunk_bytes(K);         // W1 as (N = K rows) x Q, K(=q)-major
    static constexpr int KS = glf_stride(K);
    static constexpr int OS_BYTES = GLT_TILE * KS * 4;
    static constexpr int M2 = (K * P <= 64) ? 64 : 128, N2 = Q * P;      // D2 MMA shape
    static constexpr int ACC_COLS = K + N2;                              // D1 (dH tile) | D2 (dW1 tile partial blocks)
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512));
    static constexpr int SMEM_MAX = 227 * 1024;
    // epilogue warp sets per CTA: the split mode runs one CTA per SM (its hi / lo stages fill the shared memory), so it
    // gets two sets that take alternate tiles (= accumulator stages); the single-pass mode runs two CTAs per SM instead
    __host__ __device__ static constexpr int nsets(bool x3) { return x3 ? 2 : 1; }
    __host__ __device__ static constexpr int threads(bool x3) { return 32 * (2 + 4 * nsets(x3) + (x3 ? 4 : 0)); }
    __host__ __device__ static constexpr size_t fixed_bytes(bool x3) {
        return 1024 + (size_t)B_BYTES * (x3 ? 2 : 1) + (size_t)OS_BYTES * nsets(x3) + 512;
    }
    // deepest stage ring (<= 8) that fits when `ctas` CTAs share an SM
    // (the split mode also needs `lo` residual buffers of one stage each)
    __host__ __device__ static constexpr int stages_for(bool x3, int ctas, int lo = 1) {
        const int s = (int)((SMEM_MAX / ctas - 1024 - (int)fixed_bytes(x3) - (x3 ? lo * STAGE : 0)) / STAGE);
        return s > 8 ? 8 : s;
    }
    __host__ __device__ static constexpr size_t smem_bytes(bool x3, int S, int L) {
        return fixed_bytes(x3) + (size_t)S * STAGE + (x3 ? (size_t)L * STAGE : 0);
    }
    // byte offset of channel j (multiple of 4) of tile edge `row` inside the packed H tile
    __device__ static __forceinline__ int h_offset(int row, int j) {
        const int pr = row / P, colp = (row % P) * K + j, jj = colp & 31;
        return (colp >> 5) * PCHUNK + pr * 128 + ((((jj >> 3) ^ pr) & 3) << 5) + ((jj & 7) << 2);
    }
};

template <int K, int Q, bool MASK_IN, bool X3>
__global__ void __launch_bounds__(GltBwdCfg<K, Q>::threads(X3)) glt_edge_bwd_kernel(const __grid_constant__ CUtensorMap tmH,
                                                                    const __grid_constant__ CUtensorMap tmZ,
                                                                    const __grid_constant__ CUtensorMap tmZ2,
                                                                    const int32_t *__restrict__ col,
                                                                    const float *__restrict__ W1,
                                                                    const float *__restrict__ G_col,
                                                                    const float *__restrict__ G_row, int64_t c, int M,
                                                                    float *__restrict__ dH, float *__restrict__ dW_partial, const int S, const int L) {
    using Cfg = GltBwdCfg<K, Q>;
    using TZ = typename Cfg::TZ;
    constexpr int KS = Cfg::KS, P = Cfg::P, NSETS = Cfg::nsets(X3), EPI_END = 2 + 4 * NSETS;
    extern __shared__ __align__(16) unsigned char glt_smem_raw[];
    unsigned char *base = glt_smem_raw + ((1024 - (glt_smem_u32(glt_smem_raw) & 1023)) & 1023);
    unsigned char *St = base;                                       // [S][H' | Z | Z']  landed tiles (= the hi operands)
    unsigned char *Sl = St + S * Cfg::STAGE;                        // [L][H' | Z | Z']  residuals lo (X3 only)
    unsigned char *Bh = Sl + (X3 ? L * Cfg::STAGE : 0);
    unsigned char *Bl = Bh + Cfg::B_BYTES;                          // (X3 only)
    float *Os = reinterpret_cast<float *>(Bh + Cfg::B_BYTES * (X3 ? 2 : 1));   // [128][KS]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(Os) + Cfg::OS_BYTES * NSETS);
    const uint32_t bar0 = glt_smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (S + s); };
    auto CONV = [&](int l) { return bar0 + 8u * (2 * S + l); };
    auto LOFREE = [&](int l) { return bar0 + 8u * (2 * S + L + l); };
    auto TFULL = [&](int a) { return bar0 + 8u * (2 * S + 2 * L + a); };
    auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * S + 2 * L + 2 + a); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 2 * L + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (int)((c + GLT_TILE - 1) / GLT_TILE), G = gridDim.x;

    if (tid == 0) {
        // a stage is free again when its MMAs have read it AND (MASK_IN) the epilogue warps have read their H rows
        for (int s = 0; s < S; ++s) { glt_mbar_init(FULL(s), 1); glt_mbar_init(EMPTY(s), MASK_IN ? 5 : 1); }
        for (int l = 0; l < L; ++l) { glt_mbar_init(CONV(l), 4); glt_mbar_init(LOFREE(l), 1); }
        for (int a = 0; a < 2; ++a) { glt_mbar_init(TFULL(a), 1); glt_mbar_init(TEMPTY(a), 4); }
        glt_fence_barrier_init();
        glt_prefetch_tmap(&tmH);
        glt_prefetch_tmap(&tmZ);
        glt_prefetch_tmap(&tmZ2);
    }
    if (warp == 1) glt_tmem_alloc(glt_smem_u32(tmem_slot), Cfg::TMEM_COLS);
    // B operand of dH = dZ W1^T: row n = input channel k, column = output channel q: W1[k][q] (as stored)
    glt_fill_operand<Q>(reinterpret_cast<char *>(Bh), X3 ? reinterpret_cast<char *>(Bl) : nullptr, K,
                        [&](int n, int qq) { return __ldg(&W1[n * Q + qq]); }, tid, blockDim.x);
    glt_fence_proxy_async();
    glt_tc_fence_before();
    __syncthreads();
    glt_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---------------- TMA producer
            int s = 0, ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(EMPTY(s), ph ^ 1);
                glt_mbar_expect_tx(FULL(s), Cfg::STAGE);
                unsigned char *st = St + s * Cfg::STAGE;
#pragma unroll
                for (int ch = 0; ch < Cfg::HCH; ++ch)
                    glt_tma_load_2d(glt_smem_u32(st + ch * Cfg::PCHUNK), &tmH, FULL(s), ch * 32, t * Cfg::R);
#pragma unroll
                for (int ch = 0; ch < TZ::NCH; ++ch)
                    glt_tma_load_2d(glt_smem_u32(st + Cfg::H_BYTES + ch * TZ::chunk_bytes(GLT_TILE)), &tmZ, FULL(s), ch * TZ::CW, t * GLT_TILE);
#pragma unroll
                for (int ch = 0; ch < Cfg::ZCH; ++ch)
                    glt_tma_load_2d(glt_smem_u32(st + Cfg::H_BYTES + Cfg::Z_BYTES + ch * Cfg::PCHUNK), &tmZ2, FULL(s), ch * 32, t * Cfg::R);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---------------- MMA issuer
            constexpr uint32_t idesc1 = glt_idesc_tf32(GLT_TILE, K, 0, 0);   // D1 (128 x K) = dZ (K-major) * W1^T
            constexpr uint32_t idesc2 = glt_idesc_tf32(Cfg::M2, Cfg::N2, 1, 1);  // D2 (P K x P Q) = H'^T (MN-major) * dZ' (MN-major)
            int s = 0, ph = 0, a = 0, aph = 0, l = 0, lph = 0;
            for (int t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(TEMPTY(a), aph ^ 1);
                if constexpr (X3) glt_mbar_wait(CONV(l), lph);   // the residuals are ready (their producer had waited for FULL(s))
                else glt_mbar_wait(FULL(s), ph);
                glt_tc_fence_after();
                const uint32_t d1 = tmem_base + a * Cfg::ACC_COLS, d2 = d1 + K;
                const uint32_t h_hi = glt_smem_u32(St + s * Cfg::STAGE), z_hi = h_hi + Cfg::H_BYTES;
                const uint32_t h_lo = glt_smem_u32(Sl + l * Cfg::STAGE), z_lo = h_lo + Cfg::H_BYTES;
                const uint32_t z2_hi = z_hi + Cfg::Z_BYTES, z2_lo = z_lo + Cfg::Z_BYTES;
                const uint32_t b_hi = glt_smem_u32(Bh), b_lo = glt_smem_u32(Bl);
                uint32_t acc1 = 0, acc2 = 0;
#pragma unroll
                for (int pass = X3 ? 0 : 2; pass < 3; ++pass) {
                    const uint32_t zb = (pass == 0) ? z_lo : z_hi, bb = (pass == 1) ? b_lo : b_hi;
#pragma unroll
                    for (int ch = 0; ch < TZ::NCH; ++ch)
#pragma unroll
                        for (int k8 = 0; k8 < TZ::CW / 8; ++k8) {
                            const uint64_t da = glt_smem_desc(zb + ch * TZ::chunk_bytes(GLT_TILE) + k8 * 32, 16, TZ::ATOM, TZ::SWZ);
                            const uint64_t db = glt_smem_desc(bb + ch * TZ::chunk_bytes(K) + k8 * 32, 16, TZ::ATOM, TZ::SWZ);
                            glt_mma_tf32(d1, da, db, idesc1, acc1);
                            acc1 = 1;
                        }
                }
#pragma unroll
                for (int pass = X3 ? 0 : 2; pass < 3; ++pass) {
                    const uint32_t hb = (pass == 0) ? h_lo : h_hi, zb = (pass == 1) ? z2_lo : z2_hi;
                    // reduction dimension = the packed rows of the tile, 8 per MMA = two 4-row swizzle atoms (SBO = 512 B);
                    // MN groups of 32 floats are one column chunk apart (LBO)
#pragma unroll
                    for (int k8 = 0; k8 < Cfg::R / 8; ++k8) {
                        const uint64_t da = glt_smem_desc(hb + k8 * 1024, Cfg::PCHUNK, 512, 1);
                        const uint64_t db = glt_smem_desc(zb + k8 * 1024, Cfg::PCHUNK, 512, 1);
                        glt_mma_tf32(d2, da, db, idesc2, acc2);
                        acc2 = 1;
                    }
                }
                glt_tc_commit(EMPTY(s));
                if constexpr (X3) glt_tc_commit(LOFREE(l));
                glt_tc_commit(TFULL(a));
                if (++s == S) { s = 0; ph ^= 1; }
                if (++a == 2) { a = 0; aph ^= 1; }
                if (++l == L) { l = 0; lph ^= 1; }
            }
        }
    } else if (warp < EPI_END) {
        // ---------------- epilogue warps: set = (warp - 2) / 4 takes the tiles i = set, set + NSETS, ... of this CTA's
        // sequence (tile i uses accumulator i % 2 and stage i % S).  Software pipelined: the raw G_col rows of the set's
        // NEXT tile are gathered into registers (un-added) and its G_row lines prefetched into L1 before the accumulator
        // of the current tile is awaited.
        const int set = (warp - 2) >> 2, qd = warp & 3, row = qd * 32 + lane;
        float *Oset = Os + set * (Cfg::OS_BYTES / 4);
        // D2: M2 = 64 puts row r in lane (r % 16) + 32 (r / 16), M2 = 128 in lane r; row r = p K + k belongs to the
        // diagonal block p (warp-uniform) whose columns are [p Q, p Q + Q)
        constexpr int RPQ = (Cfg::M2 == 64) ? 16 : 32;
        const int r2 = qd * RPQ + lane, p2 = (qd * RPQ) / K, krow = r2 - p2 * K;
        const bool warp_has_dw = qd * RPQ < K * P;
        const bool has_dw = warp_has_dw && lane < RPQ;
        const bool small = c < ((int64_t)1 << 31);
        auto edge_row = [&](int64_t e) -> int64_t { return small ? (int64_t)((uint32_t)e / (uint32_t)M) : e / M; };
        auto load_col = [&](int t) -> int {
            const int64_t e = (int64_t)t * GLT_TILE + row;
            return (t < ntiles && e < c) ? __ldg(&col[e]) : -1;
        };
        auto gather_col = [&](int cidx, float *dst) {
            if (cidx >= 0) {
                const float *gc = G_col + (int64_t)cidx * K;
#pragma unroll
                for (int j = 0; j < K / 4; ++j) {
                    const float4 x = glf_ldg4(gc + 4 * j);
                    dst[4 * j] = x.x; dst[4 * j + 1] = x.y; dst[4 * j + 2] = x.z; dst[4 * j + 3] = x.w;
                }
            }
        };
        auto prefetch_row = [&](int t) {
            const int64_t e = (int64_t)t * GLT_TILE + row;
            if (t < ntiles && e < c) {
                const float *gr = G_row + edge_row(e) * K;
#pragma unroll
                for (int j = 0; j < K; j += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(gr + j));
            }
        };
        auto add_row = [&](int t, float *dst) {
            const int64_t e = (int64_t)t * GLT_TILE + row;
            if (e < c) {
                const float *gr = G_row + edge_row(e) * K;
#pragma unroll
                for (int j = 0; j < K / 4; ++j) {
                    const float4 y = glf_ldg4(gr + 4 * j);
                    dst[4 * j] += y.x; dst[4 * j + 1] += y.y; dst[4 * j + 2] += y.z; dst[4 * j + 3] += y.w;
                }
            }
        };
        float dw[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) dw[j] = 0.f;
        float g[K], gn[K];
#pragma unroll
        for (int j = 0; j < K; ++j) g[j] = gn[j] = 0.f;
        const int TS = NSETS * G;                       // tile stride of this set
        int i = set, t = blockIdx.x + set * G;
        gather_col(load_col(t), g);
        int c_next = load_col(t + TS);
        for (; t < ntiles; t += TS, i += NSETS) {
            gather_col(c_next, gn);
            c_next = load_col(t + 2 * TS);
            prefetch_row(t + TS);
            add_row(t, g);
            const int s = i % S, ph = (i / S) & 1, a = i & 1, aph = (i >> 1) & 1;
            unsigned char *st = St + s * Cfg::STAGE;
            const int64_t e0 = (int64_t)t * GLT_TILE;
            // ReLU mask of the producer of H: the sign bits of this thread's row are packed as soon as the tile has
            // landed (the landed tile is never modified), and the stage is released before the accumulator is awaited
            unsigned long long hmask = 0ull;
            if constexpr (MASK_IN) {
                glt_mbar_wait(FULL(s), ph);
#pragma unroll
                for (int jj = 0; jj < K; jj += 4) {
                    const float4 h = *reinterpret_cast<const float4 *>(st + Cfg::h_offset(row, jj));
                    hmask |= (unsigned long long)((h.x > 0.f ? 1u : 0u) | (h.y > 0.f ? 2u : 0u) | (h.z > 0.f ? 4u : 0u) | (h.w > 0.f ? 8u : 0u)) << jj;
                }
                __syncwarp();
                if (lane == 0) glt_mbar_arrive(EMPTY(s));
            }
            glt_mbar_wait(TFULL(a), aph);
            glt_tc_fence_after();
            const uint32_t tq = tmem_base + ((uint32_t)(qd * 32) << 16) + a * Cfg::ACC_COLS;
            float v[K], w[Q];
            glt_tmem_ld<K>(tq, v);                      // all accumulator loads in flight, one wait
            if (warp_has_dw) glt_tmem_ld<Q>(tq + K + p2 * Q, w);
            glt_tc_wait_ld();
#pragma unroll
            for (int jj = 0; jj < K; jj += 4) {
                float4 o = make_float4(v[jj] + g[jj], v[jj + 1] + g[jj + 1], v[jj + 2] + g[jj + 2], v[jj + 3] + g[jj + 3]);
                if constexpr (MASK_IN) {   // ReLU backward of the producer of H
                    const unsigned m4 = (unsigned)(hmask >> jj);
                    o.x = (m4 & 1u) ? o.x : 0.f; o.y = (m4 & 2u) ? o.y : 0.f; o.z = (m4 & 4u) ? o.z : 0.f; o.w = (m4 & 8u) ? o.w : 0.f;
                }
                *reinterpret_cast<float4 *>(Oset + row * KS + jj) = o;
            }
            if (has_dw) {
#pragma unroll
                for (int j = 0; j < Q; ++j) dw[j] += w[j];
            }
            glt_tc_fence_before();
            __syncwarp();
            if (lane == 0) glt_mbar_arrive(TEMPTY(a));
            glf_store_warp_rows<K, KS>(Oset + qd * 32 * KS, dH, e0 + qd * 32, c);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < K; ++j) g[j] = gn[j];
        }
        if (has_dw) {
            float *dst = dW_partial + ((((int64_t)blockIdx.x * NSETS + set) * P + p2) * K + krow) * Q;
#pragma unroll
            for (int j = 0; j < Q / 4; ++j)
                *reinterpret_cast<float4 *>(dst + 4 * j) = make_float4(dw[4 * j], dw[4 * j + 1], dw[4 * j + 2], dw[4 * j + 3]);
        }
    } else {
        // ---------------- converter warps (X3 only): split the landed tiles into TF32 hi / lo in place
        if constexpr (X3) {
            const int wtid = tid - 32 * EPI_END;
            int s = 0, ph = 0, l = 0, lph = 0;
            for (int t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(FULL(s), ph);
                glt_mbar_wait(LOFREE(l), lph ^ 1);
                glt_split_inplace<Cfg::STAGE / 4>(reinterpret_cast<const float *>(St + s * Cfg::STAGE),
                                                  reinterpret_cast<float *>(Sl + l * Cfg::STAGE), wtid);
                glt_fence_proxy_async();
                __syncwarp();
                if (lane == 0) glt_mbar_arrive(CONV(l));
                if (++s == S) { s = 0; ph ^= 1; }
                if (++l == L) { l = 0; lph ^= 1; }
            }
        }
    }
    glt_tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        glt_tc_fence_after();
        glt_tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// returns the number of per-block dW1 partials written (<= 0 on error)
template <int K, int Q, bool MASK_IN, bool X3>
static int glt_launch_edge_bwd_t(const float *dZ, const float *H, const int32_t *col, const float *W1, const float *Gc, const float *Gr,
                                 int64_t c, int M, float *dH, float *partial, cudaStream_t stream) {
    using Cfg = GltBwdCfg<K, Q>;
    if constexpr (Cfg::stages_for(X3, 1) < 2) return -1;
    else {
    if (c % Cfg::P) return -1;   // the packed view needs whole rows
    CUtensorMap tmH, tmZ, tmZ2;
    if (glt_make_tmap_packed(&tmH, H, c / Cfg::P, K * Cfg::P, Cfg::R) || glt_make_tmap<Q>(&tmZ, dZ, c) ||
        glt_make_tmap_packed(&tmZ2, dZ, c / Cfg::P, Q * Cfg::P, Cfg::R))
        return -1;
    auto kern = glt_edge_bwd_kernel<K, Q, MASK_IN, X3>;
    constexpr int threads = Cfg::threads(X3);
    static int grid_cache = 0, S = 0, L = 1;
    if (!grid_cache) {
        int ctas = 1;
        if (!X3)
            for (int t = 2; t >= 1; --t)
                if (Cfg::stages_for(X3, t) >= 2 && t * Cfg::TMEM_COLS <= 512) { ctas = t; break; }
        S = Cfg::stages_for(X3, ctas);
        S = S > 6 ? 6 : S;
        // split mode: this kernel is bound by shared-memory bandwidth (landing + residual pass + three operand passes +
        // output staging move ~0.4 MB per tile), not by bytes in flight: 2 landing stages + 2 residual buffers measured
        // best (355 us vs 385 us for 4 + 1 at k=32, q=16, 3.67 M edges)
        if (X3 && Cfg::stages_for(X3, 1, 2) >= 2) { S = 2; L = 2; }
        glt_env_cfg("NBPC_GLT_BWD", &ctas, &S, &L);
        if (S < 1 || S > 8 || L < 1 || L > 8 || Cfg::smem_bytes(X3, S, L) > (size_t)Cfg::SMEM_MAX || ctas * Cfg::TMEM_COLS > 512) return -1;
        grid_cache = glt_grid(kern, threads, Cfg::smem_bytes(X3, S, L), ctas);
    }
    const size_t smem = Cfg::smem_bytes(X3, S, L);
    if (grid_cache < 0) return -1;
    const int64_t ntiles = (c + GLT_TILE - 1) / GLT_TILE;
    const int grid = (int)(ntiles < grid_cache ? ntiles : grid_cache);
    NBPC_LAUNCH_N(NbpcKName(X3 ? "glt_edge_bwd_tf32x3" : "glt_edge_bwd_tf32", K, Q).c_str(), kern, grid, threads, smem, stream, tmH, tmZ, tmZ2,
                  col, W1, Gc, Gr, c, M, dH, partial, S, L);
    return grid * Cfg::P * Cfg::nsets(X3);
    }
}
template <int K, int Q>
static int glt_launch_edge_bwd(const float *dZ, const float *H, const int32_t *col, const float *W1, const float *Gc, const float *Gr,
                               int64_t c, int M, int mask_in, int x3, float *dH, float *partial, cudaStream_t stream) {
    int nb;
    if (mask_in) nb = x3 ? glt_launch_edge_bwd_t<K, Q, true, true>(dZ, H, col, W1, Gc, Gr, c, M, dH, partial, stream)
                         : glt_launch_edge_bwd_t<K, Q, true, false>(dZ, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
    else nb = x3 ? glt_launch_edge_bwd_t<K, Q, false, true>(dZ, H, col, W1, Gc, Gr, c, M, dH, partial, stream)
                 : glt_launch_edge_bwd_t<K, Q, false, false>(dZ, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
    return nb;
}

int glt_max_partial_blocks() { return 2 * 2 * 2 * gl_num_sms(); }

// (k, q, mode) combinations whose stage ring fits shared memory at least twice; c must be a multiple of the packing
bool glt_bwd_shape_ok(int k, int q, int x3, int64_t c) {
#define X(K_, Q_) if (k == K_ && q == Q_) return GltBwdCfg<K_, Q_>::stages_for(x3 != 0, 1) >= 2 && c % GltBwdCfg<K_, Q_>::P == 0;
    GLT_FOR_KQ(X)
#undef X
    return false;
}

int glt_edge_bwd(int k, int q, const float *dZ, const float *H, const int32_t *col, const float *W1, const float *Gc, const float *Gr,
                 int64_t c, int M, int mask_in, int x3, float *dH, float *partial, cudaStream_t stream) {
#define X(K_, Q_) if (k == K_ && q == Q_) return glt_launch_edge_bwd<K_, Q_>(dZ, H, col, W1, Gc, Gr, c, M, mask_in, x3, dH, partial, stream);
    GLT_FOR_KQ(X)
#undef X
    return -1;
}
#endif
