// graph_layer_tc_fwd.cu - tcgen05 forward edge kernel (see graph_layer_tc.cuh)
#include "graph_layer_tc.cuh"
#ifndef NBPC_HOST_EMU
// ------------------------------------------------------------------ forward kernel
template <int K, int Q>
struct GltFwdCfg {
    using TA = GltTile<K>;
    static constexpr int A_STAGE = TA::NCH * TA::chunk_bytes(GLT_TILE);    // raw / hi tile
    static constexpr int B_BYTES = TA::NCH * TA::chunk_bytes(Q);           // W1^T as (N = Q rows) x K, K-major
    static constexpr int QS = glf_stride(Q);
    static constexpr int OS_BYTES = GLT_TILE * QS * 4;
    static constexpr int TMEM_COLS = (2 * Q <= 32) ? 32 : (2 * Q <= 64 ? 64 : (2 * Q <= 128 ? 128 : 256));
    static constexpr int SMEM_MAX = 227 * 1024;
    __host__ __device__ static constexpr size_t fixed_bytes(bool x3) { return 1024 + (size_t)B_BYTES * (x3 ? 2 : 1) + OS_BYTES + 512; }
    // deepest stage ring (<= 8) that fits when `ctas` CTAs share an SM (each CTA also pays 1 KB of system shared memory)
    // (the split mode also needs `lo` residual buffers of one stage each)
    __host__ __device__ static constexpr int stages_for(bool x3, int ctas, int lo = 1) {
        const int s = (int)((SMEM_MAX / ctas - 1024 - (int)fixed_bytes(x3) - (x3 ? lo * A_STAGE : 0)) / A_STAGE);
        return s > 8 ? 8 : s;
    }
    __host__ __device__ static constexpr size_t smem_bytes(bool x3, int S, int L) {
        return fixed_bytes(x3) + (size_t)S * A_STAGE + (x3 ? (size_t)L * A_STAGE : 0);
    }
};

template <int K, int Q, bool RELU, bool X3>
__global__ void __launch_bounds__(X3 ? GLT_THREADS_X3 : GLT_THREADS) glt_edge_out_kernel(const __grid_constant__ CUtensorMap tmH,
                                                                                         const int32_t *__restrict__ col,
                                                                                         const float *__restrict__ W1,
                                                                                         const float *__restrict__ Q_col,
                                                                                         const float *__restrict__ Q_row, int64_t c, int M,
                                                                                         float *__restrict__ out, const int S, const int L) {
    using Cfg = GltFwdCfg<K, Q>;
    using TA = typename Cfg::TA;
    constexpr int QS = Cfg::QS;
    extern __shared__ __align__(16) unsigned char glt_smem_raw[];
    unsigned char *base = glt_smem_raw + ((1024 - (glt_smem_u32(glt_smem_raw) & 1023)) & 1023);
    unsigned char *As = base;                                      // [S][A_STAGE]   landed tiles (= the hi operand)
    unsigned char *Al = As + S * Cfg::A_STAGE;                     // [L][A_STAGE]   residuals lo (X3 only)
    unsigned char *Bh = Al + (X3 ? L * Cfg::A_STAGE : 0);          // [B_BYTES]
    unsigned char *Bl = Bh + Cfg::B_BYTES;                         // [B_BYTES]      (X3 only)
    float *Os = reinterpret_cast<float *>(Bh + Cfg::B_BYTES * (X3 ? 2 : 1));   // [128][QS]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(Os) + Cfg::OS_BYTES);
    // barriers: full[S] | empty[S] | conv[L] | lofree[L] | tmem_full[2] | tmem_empty[2]   (S, L <= 8)
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
        for (int s = 0; s < S; ++s) { glt_mbar_init(FULL(s), 1); glt_mbar_init(EMPTY(s), 1); }
        for (int l = 0; l < L; ++l) { glt_mbar_init(CONV(l), 4); glt_mbar_init(LOFREE(l), 1); }
        for (int a = 0; a < 2; ++a) { glt_mbar_init(TFULL(a), 1); glt_mbar_init(TEMPTY(a), 4); }
        glt_fence_barrier_init();
        glt_prefetch_tmap(&tmH);
    }
    if (warp == 1) glt_tmem_alloc(glt_smem_u32(tmem_slot), Cfg::TMEM_COLS);
    // B operand: row n = output channel, column kk = input channel: W1[kk][n]
    glt_fill_operand<K>(reinterpret_cast<char *>(Bh), X3 ? reinterpret_cast<char *>(Bl) : nullptr, Q,
                        [&](int n, int kk) { return __ldg(&W1[kk * Q + n]); }, tid, blockDim.x);
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
                glt_mbar_expect_tx(FULL(s), Cfg::A_STAGE);
#pragma unroll
                for (int ch = 0; ch < TA::NCH; ++ch)
                    glt_tma_load_2d(glt_smem_u32(As + s * Cfg::A_STAGE + ch * TA::chunk_bytes(GLT_TILE)), &tmH, FULL(s), ch * TA::CW, t * GLT_TILE);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---------------- MMA issuer
            constexpr uint32_t idesc = glt_idesc_tf32(GLT_TILE, Q, 0, 0);
            int s = 0, ph = 0, a = 0, aph = 0, l = 0, lph = 0;
            for (int t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(TEMPTY(a), aph ^ 1);
                if constexpr (X3) glt_mbar_wait(CONV(l), lph);   // the residual is ready (its producer had waited for FULL(s))
                else glt_mbar_wait(FULL(s), ph);
                glt_tc_fence_after();
                const uint32_t d = tmem_base + a * Q;
                const uint32_t a_hi = glt_smem_u32(As + s * Cfg::A_STAGE), a_lo = glt_smem_u32(Al + l * Cfg::A_STAGE);
                const uint32_t b_hi = glt_smem_u32(Bh), b_lo = glt_smem_u32(Bl);
                uint32_t acc = 0;
#pragma unroll
                for (int pass = X3 ? 0 : 2; pass < 3; ++pass) {   // small terms first: lo*hi, hi*lo, hi*hi
                    const uint32_t ab = (pass == 0) ? a_lo : a_hi, bb = (pass == 1) ? b_lo : b_hi;
#pragma unroll
                    for (int ch = 0; ch < TA::NCH; ++ch)
#pragma unroll
                        for (int k8 = 0; k8 < TA::CW / 8; ++k8) {
                            const uint64_t da = glt_smem_desc(ab + ch * TA::chunk_bytes(GLT_TILE) + k8 * 32, 16, TA::ATOM, TA::SWZ);
                            const uint64_t db = glt_smem_desc(bb + ch * TA::chunk_bytes(Q) + k8 * 32, 16, TA::ATOM, TA::SWZ);
                            glt_mma_tf32(d, da, db, idesc, acc);
                            acc = 1;
                        }
                }
                glt_tc_commit(EMPTY(s));    // the stage may be refilled once these MMAs have read it
                if constexpr (X3) glt_tc_commit(LOFREE(l));
                glt_tc_commit(TFULL(a));    // accumulator ready for the epilogue
                if (++s == S) { s = 0; ph ^= 1; }
                if (++a == 2) { a = 0; aph ^= 1; }
                if (++l == L) { l = 0; lph ^= 1; }
            }
        }
    } else if (warp < 6) {
        // ---------------- epilogue warps: quadrant = warp % 4 owns TMEM lanes / tile rows [32 qd, 32 qd + 32).
        // Software pipelined: the node-level terms of tile i+1 are gathered (and the column index of tile i+2 is loaded)
        // before the accumulator of tile i is awaited, so the dependent-load chain is off the critical path.
        const int qd = warp & 3, row = qd * 32 + lane;
        const bool small = c < ((int64_t)1 << 31);
        auto edge_row = [&](int64_t e) -> int64_t { return small ? (int64_t)((uint32_t)e / (uint32_t)M) : e / M; };
        auto load_col = [&](int t) -> int {
            const int64_t e = (int64_t)t * GLT_TILE + row;
            return (t < ntiles && e < c) ? __ldg(&col[e]) : -1;
        };
        // raw Q_col[col] rows of the NEXT tile are kept in registers un-added (an add would wait for the loads right
        // away); the Q_row[e / M] rows are shared by M consecutive edges and are re-read at use time (L1 hits)
        auto gather_col = [&](int cidx, float *dst) {
            if (cidx >= 0) {
                const float *qc = Q_col + (int64_t)cidx * Q;
#pragma unroll
                for (int j = 0; j < Q / 4; ++j) {
                    const float4 x = glf_ldg4(qc + 4 * j);
                    dst[4 * j] = x.x; dst[4 * j + 1] = x.y; dst[4 * j + 2] = x.z; dst[4 * j + 3] = x.w;
                }
            }
        };
        auto add_row = [&](int t, float *dst) {
            const int64_t e = (int64_t)t * GLT_TILE + row;
            if (e < c) {
                const float *qr = Q_row + edge_row(e) * Q;
#pragma unroll
                for (int j = 0; j < Q / 4; ++j) {
                    const float4 y = glf_ldg4(qr + 4 * j);
                    dst[4 * j] += y.x; dst[4 * j + 1] += y.y; dst[4 * j + 2] += y.z; dst[4 * j + 3] += y.w;
                }
            }
        };
        int a = 0, aph = 0;
        int t = blockIdx.x;
        float cur[Q], nxt[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) cur[j] = nxt[j] = 0.f;
        gather_col(load_col(t), cur);
        int c_next = load_col(t + G);
        for (; t < ntiles; t += G) {
            gather_col(c_next, nxt);
            c_next = load_col(t + 2 * G);
            add_row(t, cur);
            const int64_t e0 = (int64_t)t * GLT_TILE;
            glt_mbar_wait(TFULL(a), aph);
            glt_tc_fence_after();
            const uint32_t tq = tmem_base + ((uint32_t)(qd * 32) << 16) + a * Q;
#pragma unroll
            for (int cb = 0; cb < Q; cb += 16) {   // 16 accumulator columns at a time
                float z[16];
                glt_tmem_ld16(tq + cb, z);
                glt_tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 o = make_float4(z[4 * j] + cur[cb + 4 * j], z[4 * j + 1] + cur[cb + 4 * j + 1], z[4 * j + 2] + cur[cb + 4 * j + 2],
                                           z[4 * j + 3] + cur[cb + 4 * j + 3]);
                    if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4 *>(Os + row * QS + cb + 4 * j) = o;
                }
            }
            glt_tc_fence_before();
            __syncwarp();
            if (lane == 0) glt_mbar_arrive(TEMPTY(a));
            if (++a == 2) { a = 0; aph ^= 1; }
            glf_store_warp_rows<Q, QS>(Os + qd * 32 * QS, out, e0 + qd * 32, c);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < Q; ++j) cur[j] = nxt[j];
        }
    } else {
        // ---------------- converter warps (X3 only): split the landed tile into TF32 hi / lo in place
        if constexpr (X3) {
            const int wtid = tid - 192;
            int s = 0, ph = 0, l = 0, lph = 0;
            for (int t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(FULL(s), ph);
                glt_mbar_wait(LOFREE(l), lph ^ 1);
                glt_split_inplace<Cfg::A_STAGE / 4>(reinterpret_cast<const float *>(As + s * Cfg::A_STAGE),
                                                    reinterpret_cast<float *>(Al + l * Cfg::A_STAGE), wtid);
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

template <int K, int Q, bool RELU, bool X3>
static int glt_launch_edge_out_t(const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c, int M,
                                 float *out, cudaStream_t stream) {
    using Cfg = GltFwdCfg<K, Q>;
    if constexpr (Cfg::stages_for(X3, 1) < 2) return 1;
    else {
    CUtensorMap tm;
    if (glt_make_tmap<K>(&tm, H, c)) return 1;
    auto kern = glt_edge_out_kernel<K, Q, RELU, X3>;
    constexpr int threads = X3 ? GLT_THREADS_X3 : GLT_THREADS;
    // CTAs per SM / ring depth: as many CTAs (<= 4) as keep a 2-deep ring: each CTA is an independent
    // load -> MMA -> epilogue pipeline, so co-resident CTAs hide each other's latencies (measured on B200, k=32 q=16, 3.67 M
    // edges: 1 CTA x 4 stages 272 us, 2 x 4 170 us, 3 x 3 150 us, 4 x 2 148 us).  NBPC_GLT_FWD="ctas,stages" overrides.
    static int grid_cache = 0, S = 0, L = 1;
    if (!grid_cache) {
        int ctas = 1;
        for (int t = 4; t >= 1; --t)
            if (Cfg::stages_for(X3, t) >= 2 && t * Cfg::TMEM_COLS <= 512) { ctas = t; break; }
        S = Cfg::stages_for(X3, ctas);
        S = S > 4 ? 4 : S;
        glt_env_cfg("NBPC_GLT_FWD", &ctas, &S, &L);
        if (S < 1 || S > 8 || L < 1 || L > 8 || Cfg::smem_bytes(X3, S, L) > (size_t)Cfg::SMEM_MAX) return 1;
        grid_cache = glt_grid(kern, threads, Cfg::smem_bytes(X3, S, L), ctas);
    }
    const size_t smem = Cfg::smem_bytes(X3, S, L);
    if (grid_cache < 0) return 1;
    const int64_t ntiles = (c + GLT_TILE - 1) / GLT_TILE;
    const int grid = (int)(ntiles < grid_cache ? ntiles : grid_cache);
    NBPC_LAUNCH_N(NbpcKName(X3 ? "glt_edge_out_tf32x3" : "glt_edge_out_tf32", K, Q).c_str(), kern, grid, threads, smem, stream, tm, col, W1, Qc,
                  Qr, c, M, out, S, L);
    return 0;
    }
}
template <int K, int Q>
static int glt_launch_edge_out(const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c, int M,
                               int relu, int x3, float *out, cudaStream_t stream) {
    if (relu) return x3 ? glt_launch_edge_out_t<K, Q, true, true>(H, col, W1, Qc, Qr, c, M, out, stream)
                        : glt_launch_edge_out_t<K, Q, true, false>(H, col, W1, Qc, Qr, c, M, out, stream);
    return x3 ? glt_launch_edge_out_t<K, Q, false, true>(H, col, W1, Qc, Qr, c, M, out, stream)
              : glt_launch_edge_out_t<K, Q, false, false>(H, col, W1, Qc, Qr, c, M, out, stream);
}


bool glt_fwd_shape_ok(int k, int q, int x3) {
#define X(K_, Q_) if (k == K_ && q == Q_) return GltFwdCfg<K_, Q_>::stages_for(x3 != 0, 1) >= 2;
    GLT_FOR_KQ(X)
#undef X
    return false;
}

int glt_edge_out(int k, int q, const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c, int M,
                 int relu, int x3, float *out, cudaStream_t stream) {
#define X(K_, Q_) if (k == K_ && q == Q_) return glt_launch_edge_out<K_, Q_>(H, col, W1, Qc, Qr, c, M, relu, x3, out, stream);
    GLT_FOR_KQ(X)
#undef X
    return 1;
}
#endif
