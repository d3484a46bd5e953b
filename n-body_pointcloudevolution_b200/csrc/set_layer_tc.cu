// set_layer_tc.cu - tcgen05 / TMEM / TMA kernels of the permutation-equivariant set layer (nn.py:10-28):
//
//   forward   out = act( (H - mu_s) W + B )                     mu_s = mean over the N particles of sample s  (nn.py:25-27)
//   backward  dH  = ( (dZ - mean_s dZ) W^T ) [* (H > 0)]        (adjoint of the mean subtraction; optional fused ReLU
//             dW  = (H - mu_s)^T dZ                              backward of the layer that produced H)
//
// At the reference's default widths (utils.py:165: 6,64,128,128,256,64,128,16,3) these are (262 144 x k)(k x q) GEMMs of
// 4 - 42 flop/B: still HBM-bound, but far beyond what the CUDA-core issue slots can stream, so they run on the tensor pipe:
//
//   sgt_gemm_kernel : persistent CTAs walk 128-row tiles.  The weight tile B (all K x NT of it, TF32 hi and lo) stays
//       resident in shared memory; A streams through a ring of [128 rows x 32 floats] stages (TMA, 128-byte swizzle);
//       converter warps subtract the per-sample column mean IN PLACE (so the product is (H - mu) W, the reference's own
//       association, not H W - mu W) and write the TF32 residual; one thread issues tcgen05.mma kind::tf32
//       (lo*hi + hi*lo + hi*hi in the default tf32x3 mode), FP32 accumulators double-buffered in TMEM; 4 epilogue warps
//       read them with tcgen05.ld, add the bias, apply ReLU / the input mask and store 128-byte row segments.
//   sgt_dw_kernel   : dW contracts over the ROWS, so both operands are MN-major (TMA SWIZZLE_128B_ATOM_32B, 32-row stages);
//       every CTA keeps its (k x q) partial in TMEM across all of its tiles and writes it once; partials are summed over
//       CTAs in a fixed order (deterministic).
//   sgt_colsum_*    : per-sample column sums (mu, mean dZ, dB) - float4 streams, fixed shared-memory tree, fixed order over
//       blocks.
// Shapes outside (k % 32 == 0, q % 16 == 0, both <= 256) - the 6-wide input and 3-wide output layers - use the CUDA-core
// kernels of set_layer.cu.
#include "graph_layer_tc.cuh"
#ifndef NBPC_HOST_EMU
#include "set_layer_tc.h"

#define SGT_CHUNK_BYTES (GLT_TILE * 128)   // one A stage: 128 rows x 32 floats
#define SGT_DW_ROWS 32                     // rows per stage of the dW kernel
#define SGT_CONV_WARPS 8                   // converter warps (centre + TF32 residual): 4 left the stage cycle latency-bound
#define SGT_THREADS (192 + 32 * SGT_CONV_WARPS)
#define SGT_COLSUM_ROWS 256                // rows per block of the column-sum pass (N = 32^3, B = 8: 1024 blocks)

// x / d for 0 <= x < 2^31 with magic = floor(2^32 / d) (2^32 - 1 for d = 1): the estimate is exact or one short
__device__ __forceinline__ uint32_t sgt_div(uint32_t x, uint32_t d, uint32_t magic) {
    uint32_t qt = __umulhi(x, magic);
    if (x - qt * d >= d) ++qt;
    return qt;
}
static inline uint32_t sgt_magic(int d) { return d == 1 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)1 << 32) / (uint32_t)d); }

// ------------------------------------------------------------------ per-sample column sums
// partial[s][blk][C] = sum over the block's rows of sample s; grid (nblk, B), C % 4 == 0, C <= 1024
__global__ void __launch_bounds__(256) sgt_colsum_partial_kernel(const float *__restrict__ X, int C, int N, int rows_per_block,
                                                                 float *__restrict__ partial) {
    __shared__ float4 red[256];
    const int G = C >> 2, slots = 256 / G;
    const int g = threadIdx.x % G, slot = threadIdx.x / G, s = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block, r1 = nbpc_min(r0 + rows_per_block, N);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (slot < slots) {
        const float *p = X + ((int64_t)s * N) * C + 4 * g;
        int r = r0 + slot;
        for (; r + 3 * slots < r1; r += 4 * slots) {   // 4 independent loads in flight
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p + (int64_t)r * C));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(p + (int64_t)(r + slots) * C));
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(p + (int64_t)(r + 2 * slots) * C));
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(p + (int64_t)(r + 3 * slots) * C));
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
            a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        }
        for (; r < r1; r += slots) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + (int64_t)r * C));
            a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w;
        }
    }
    float4 a = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                           (a0.w + a1.w) + (a2.w + a3.w));
    red[threadIdx.x] = a;
    __syncthreads();
    // fixed order over the slots (slot counts are not powers of two for every C: sum sequentially in thread slot 0)
    if (slot == 0) {
        for (int sl = 1; sl < slots; ++sl) {
            const float4 b = red[sl * G + g];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        *reinterpret_cast<float4 *>(partial + ((int64_t)s * gridDim.x + blockIdx.x) * C + 4 * g) = a;
    }
}
// out[s][c] = (sum_blk partial[s][blk][c]) * scale: one WARP per (sample, column) - lane l adds blocks l, l + 32, ... and
// the lanes are combined by a fixed butterfly (deterministic)
__global__ void sgt_colsum_final_kernel(const float *__restrict__ partial, int C, int nblk, int B, float scale, float *__restrict__ out,
                                        float *__restrict__ sums) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= B * C) return;
    const int s = t / C, c = t % C;
    const float *p = partial + (int64_t)s * nblk * C + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = lane;
    for (; b + 96 < nblk; b += 128) {
        a0 += __ldg(p + (int64_t)b * C); a1 += __ldg(p + (int64_t)(b + 32) * C);
        a2 += __ldg(p + (int64_t)(b + 64) * C); a3 += __ldg(p + (int64_t)(b + 96) * C);
    }
    for (; b < nblk; b += 32) a0 += __ldg(p + (int64_t)b * C);
    float sum = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
        out[t] = sum * scale;
        if (sums) sums[t] = sum;
    }
}
// total[c] = sum_s sums[s][c]
__global__ void sgt_colsum_total_kernel(const float *__restrict__ sums, int C, int B, float *__restrict__ total) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float tot = 0.f;
    for (int s = 0; s < B; ++s) tot += __ldg(&sums[s * C + c]);
    total[c] = tot;
}

// out[o] = scale * sum over sets of parts[set][o] (o = sample * C + c), sums[o] = the unscaled sum (optional); fixed order:
// a block owns 32 outputs, its 32 slices take the sets slice, slice + 32, ... (4 loads in flight) and are combined in
// ascending slice order
__global__ void __launch_bounds__(1024) sgt_colsum_from_parts_kernel(const float *__restrict__ parts, int nsets, int BC, float scale,
                                                                      float *__restrict__ out, float *__restrict__ sums) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (o < BC) {
        int b = slice;
        for (; b + 96 < nsets; b += 128) {
            a0 += __ldg(parts + (int64_t)b * BC + o); a1 += __ldg(parts + (int64_t)(b + 32) * BC + o);
            a2 += __ldg(parts + (int64_t)(b + 64) * BC + o); a3 += __ldg(parts + (int64_t)(b + 96) * BC + o);
        }
        for (; b < nsets; b += 32) a0 += __ldg(parts + (int64_t)b * BC + o);
    }
    red[slice][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (slice == 0 && o < BC) {
        float a = red[0][lane];
#pragma unroll
        for (int j = 1; j < 32; ++j) a += red[j][lane];
        if (out) out[o] = a * scale;
        if (sums) sums[o] = a;
    }
}
void sgt_colsum_from_parts(const float *parts, int nsets, int B, int C, float scale, float *out, float *sums, cudaStream_t stream) {
    NBPC_LAUNCH(sgt_colsum_from_parts_kernel, nbpc_cdiv(B * C, 32), 1024, 0, stream, parts, nsets, B * C, scale, out, sums);
}

// ------------------------------------------------------------------ row GEMM: out = act((A - mu_s) Bm + bias) [* (mask > 0)]
struct SgtGemmArgs {
    const float *Bsrc;       // weights: (K, Ntot) row-major, or (Ntot, K) when b_transposed
    const float *mu;         // (samples, K) column means subtracted from A, or nullptr
    const float *bias;       // (Ntot) or nullptr
    const float *mask;       // (rows, Ntot): output multiplied by [mask > 0], or nullptr
    float *out;              // (rows, Ntot)
    int64_t rows;
    int rows_per_sample, K, NT, Ntot, n_ntiles, b_transposed, relu, S, L;
    uint32_t rps_magic;      // floor(2^32 / rows_per_sample) (2^32 - 1 for 1): sample of a row without a 64-bit division
    int dbg;                 // NBPC_SGT_DEBUG (profiling experiments only, results are WRONG): 1 no output stores, 2 no epilogue
                             // work at all, 4 converters only signal
    float *csum;             // optional: per-sample column sums of `out`, [gridDim.x / n_ntiles][samples][Ntot] partials
                             // (one set per CTA row; requires rows_per_sample % 128 == 0)
    int samples, tiles_per_sample;
    int obuf;                // staging tiles per epilogue warp (1 or 2)
    int lo_tmem;             // TF32x3: number of A-residual slots kept in TENSOR memory (0: the residual ring is in shared memory)
};

template <bool X3>
__global__ void __launch_bounds__(SGT_THREADS) sgt_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmO,
                                                               const SgtGemmArgs P) {
    extern __shared__ __align__(16) unsigned char glt_smem_raw[];
    unsigned char *base = glt_smem_raw + ((1024 - (glt_smem_u32(glt_smem_raw) & 1023)) & 1023);
    const int S = P.S, L = (X3 && !P.lo_tmem) ? P.L : 0, K = P.K, NT = P.NT, KC = K >> 5;
    const int LT = X3 ? P.lo_tmem : 0;                            // residual slots in tensor memory (32 columns each, after the accumulators)
    const int SW = NT < 32 ? NT : 32, PITCH = SW + 4;            // epilogue slab: SW accumulator columns at a time
    const int B_BYTES = K * NT * 4;
    unsigned char *As = base;                                     // [S][16 KB] landed (then centred) chunks = hi operand
    unsigned char *Al = As + S * SGT_CHUNK_BYTES;                 // [L][16 KB] residuals
    unsigned char *Bh = Al + L * SGT_CHUNK_BYTES;
    unsigned char *Bl = Bh + B_BYTES;
    unsigned char *Ob = Bh + B_BYTES * (X3 ? 2 : 1);              // [4 warps][obuf][32 rows x 128 B] output staging (TMA store, 128-B swizzle)
    float *Os = reinterpret_cast<float *>(Ob);                    // NT = 16 only: [4 warps][32][PITCH = 20] (aliases Ob)
    float *Cs = reinterpret_cast<float *>(Ob + 4 * P.obuf * 4096); // [4 warps][256] running column sums of the current sample
    float *Bs = Cs + 4 * 256;                                     // [256] bias of this CTA's columns
    uint64_t *bars = reinterpret_cast<uint64_t *>(Bs + 256);
    const uint32_t bar0 = glt_smem_u32(bars);
    const int LB = X3 ? (P.lo_tmem ? P.lo_tmem : P.L) : 1;        // barrier slots are laid out for max(residual slots, 1)
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (S + s); };
    auto CONV = [&](int s) { return bar0 + 8u * (2 * S + s); };   // per STAGE: centred (and split) data ready for the MMA
    auto LOFREE = [&](int l) { return bar0 + 8u * (3 * S + l); };
    auto TFULL = [&](int a) { return bar0 + 8u * (3 * S + LB + a); };
    auto TEMPTY = [&](int a) { return bar0 + 8u * (3 * S + LB + 2 + a); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * S + LB + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (P.rows + GLT_TILE - 1) / GLT_TILE;
    const int n_tile = blockIdx.x % P.n_ntiles, cta_m = blockIdx.x / P.n_ntiles, Gm = gridDim.x / P.n_ntiles;
    const int n0 = n_tile * NT;
    int tmem_cols = 32;
    while (tmem_cols < 2 * NT + 32 * LT) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { glt_mbar_init(FULL(s), 1); glt_mbar_init(EMPTY(s), 1); glt_mbar_init(CONV(s), SGT_CONV_WARPS); }
        for (int l = 0; l < LB; ++l) glt_mbar_init(LOFREE(l), 1);
        for (int a = 0; a < 2; ++a) { glt_mbar_init(TFULL(a), 1); glt_mbar_init(TEMPTY(a), 4); }
        glt_fence_barrier_init();
        glt_prefetch_tmap(&tmA);
        glt_prefetch_tmap(&tmO);
    }
    if (warp == 1) glt_tmem_alloc(glt_smem_u32(tmem_slot), tmem_cols);
    // B operand, K-major: row n = output column n0 + n, element kk; chunk c = 32 consecutive kk of all NT rows.
    // float4 loads along the contiguous direction of the source, 4 independent loads per thread and step (a
    // one-load-per-iteration loop cost ~30 us per launch in latency, the scalar version ~5 us)
    {
        const int V = P.b_transposed ? (K >> 2) : (NT >> 2);        // float4s per source row
        const int total4 = (K * NT) >> 2;
        auto put = [&](int n, int kk, float x) {
            const int off = (kk >> 5) * (NT * 128) + GltTile<32>::offset(n, kk & 31);
            *reinterpret_cast<float *>(Bh + off) = X3 ? x : glt_to_tf32(x);
            if (X3) *reinterpret_cast<float *>(Bl + off) = glt_residual(x);
        };
        for (int i0 = tid; i0 < total4; i0 += 4 * blockDim.x) {
            float4 x[4];
            int a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                a[u] = i / V;                       // source row: n (transposed) or kk
                b[u] = (i - a[u] * V) << 2;         // first of 4 consecutive kk (transposed) or n
                if (i < total4)
                    x[u] = glf_ldg4(P.b_transposed ? P.Bsrc + (int64_t)(n0 + a[u]) * K + b[u] : P.Bsrc + (int64_t)a[u] * P.Ntot + n0 + b[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * blockDim.x < total4) {
                    if (P.b_transposed) {           // 4 consecutive kk of row n: one 16-byte unit
                        const int off = (b[u] >> 5) * (NT * 128) + GltTile<32>::offset(a[u], b[u] & 31);
                        *reinterpret_cast<float4 *>(Bh + off) = X3 ? x[u] : make_float4(glt_to_tf32(x[u].x), glt_to_tf32(x[u].y), glt_to_tf32(x[u].z), glt_to_tf32(x[u].w));
                        if (X3) *reinterpret_cast<float4 *>(Bl + off) = make_float4(glt_residual(x[u].x), glt_residual(x[u].y), glt_residual(x[u].z), glt_residual(x[u].w));
                    } else {                        // 4 consecutive n of source row kk: 4 rows of the K-major tile
                        put(b[u], a[u], x[u].x); put(b[u] + 1, a[u], x[u].y); put(b[u] + 2, a[u], x[u].z); put(b[u] + 3, a[u], x[u].w);
                    }
                }
            }
        }
    }
    for (int i = tid; i < NT; i += blockDim.x) Bs[i] = P.bias ? __ldg(&P.bias[n0 + i]) : 0.f;
    glt_fence_proxy_async();
    glt_tc_fence_before();
    __syncthreads();
    glt_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---------------- TMA producer
            int s = 0, ph = 0;
            for (int64_t t = cta_m; t < ntiles; t += Gm)
                for (int c = 0; c < KC; ++c) {
                    glt_mbar_wait(EMPTY(s), ph ^ 1);
                    glt_mbar_expect_tx(FULL(s), SGT_CHUNK_BYTES);
                    glt_tma_load_2d(glt_smem_u32(As + s * SGT_CHUNK_BYTES), &tmA, FULL(s), c * 32, (int)(t * GLT_TILE));
                    if (++s == S) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---------------- MMA issuer
            const uint32_t idesc = glt_idesc_tf32(GLT_TILE, NT, 0, 0);
            // The issuing thread's own instruction stream bounds the chunk rate (one thread, ~5 cycles per dependent
            // instruction): descriptors are built ONCE and advanced by adding to their 14-bit address field (units of 16 bytes:
            // +2 per 8-float K step, +1024 per 16 KB stage, + NT * 8 per weight chunk)
            const uint64_t a_hi0 = glt_smem_desc(glt_smem_u32(As), 16, 1024, 2), a_lo0 = glt_smem_desc(glt_smem_u32(Al), 16, 1024, 2);
            const uint64_t b_hi0 = glt_smem_desc(glt_smem_u32(Bh), 16, 1024, 2), b_lo0 = glt_smem_desc(glt_smem_u32(Bl), 16, 1024, 2);
            const uint32_t b_step = (uint32_t)NT * 8;
            int s = 0, ph = 0, a = 0, aph = 0, l = 0;
            for (int64_t t = cta_m; t < ntiles; t += Gm) {
                glt_mbar_wait(TEMPTY(a), aph ^ 1);
                const uint32_t d = tmem_base + a * NT;
                uint32_t acc = 0;
                for (int c = 0; c < KC; ++c) {
                    glt_mbar_wait(CONV(s), ph);
                    glt_tc_fence_after();
                    const uint64_t a_hi = a_hi0 + (uint64_t)(s * (SGT_CHUNK_BYTES >> 4)), a_lo = a_lo0 + (uint64_t)(l * (SGT_CHUNK_BYTES >> 4));
                    const uint64_t b_hi = b_hi0 + (uint64_t)(c * b_step), b_lo = b_lo0 + (uint64_t)(c * b_step);
                    if (X3 && LT) {                                   // lo*hi with the residual read from tensor memory
                        const uint32_t a_t = tmem_base + 2 * NT + l * 32;
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {
                            glt_mma_tf32_ts(d, a_t + 8 * k8, b_hi + 2 * k8, idesc, acc);
                            acc = 1;
                        }
                    }
#pragma unroll
                    for (int pass = X3 ? 0 : 2; pass < 3; ++pass) {   // small terms first: lo*hi, hi*lo, hi*hi
                        if (pass == 0 && LT) continue;
                        const uint64_t ab = (pass == 0) ? a_lo : a_hi, bb = (pass == 1) ? b_lo : b_hi;
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {
                            glt_mma_tf32(d, ab + 2 * k8, bb + 2 * k8, idesc, acc);
                            acc = 1;
                        }
                    }
                    glt_tc_commit(EMPTY(s));
                    if (X3) { glt_tc_commit(LOFREE(l)); if (++l == (LT ? LT : L)) l = 0; }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                glt_tc_commit(TFULL(a));
                if (++a == 2) { a = 0; aph ^= 1; }
            }
        }
    } else if (warp < 6) {
        // ---------------- epilogue warps: quadrant qd owns TMEM lanes / tile rows [32 qd, 32 qd + 32)
        const int qd = warp & 3;
        float *Ow = Os + qd * 32 * 20;               // (16-column path: PITCH = 20)
        const int vshift = SW == 32 ? 3 : 2, vmask = (1 << vshift) - 1, iters = 1 << vshift;   // 32 rows * SW/4 float4 = 32 lanes * iters
        // (shifts: a runtime division by SW/4 in the two loops below cost more than the stores)
        int a = 0, aph = 0;
        // fused column sums of the output (the next layer's mean / the previous layer's bias gradient): per warp, per sample
        float *Cw = Cs + qd * 256;
        int cur_sample = -1;
        // (the four epilogue warps walk the same tiles, so they flush at the same points: named barrier 1, 128 threads;
        // their sums are combined in warp order 0..3, one partial set per CTA row)
        auto csum_flush = [&]() {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (cur_sample >= 0) {
                float *dst = P.csum + ((int64_t)cta_m * P.samples + cur_sample) * P.Ntot + n0;
                for (int i = qd * 32 + lane; i < NT; i += 128) dst[i] = ((Cs[i] + Cs[256 + i]) + Cs[512 + i]) + Cs[768 + i];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = lane; i < NT; i += 32) Cw[i] = 0.f;
            __syncwarp();
        };
        if (P.csum) csum_flush();
        if (SW == 32) {
            // ---- 32-column slabs: lane = tile row.  The lane adds the bias (shared memory), applies ReLU / the input mask
            // (its own 128-byte row segment, requested one slab ahead) and writes its row into a 128-byte-swizzled staging
            // tile; one elected lane hands the tile to the TMA store engine (coalescing, address generation and the
            // clipping of the last tile's rows are the engine's job).  Two staging tiles per warp: the store of slab i
            // reads its tile while slab i + 1 is assembled.  Column sums are taken from the staged tile (lane = column).
            unsigned char *ob = Ob + qd * P.obuf * 4096;
            const uint32_t bmask = (uint32_t)P.obuf - 1;
            const uint32_t ob_u32 = glt_smem_u32(ob);
            const int cq = lane >> 2, cw = (lane & 3) * 4;
            uint32_t nslab = 0;
            float4 mk[8];
            auto mask_issue = [&](int64_t grow, int col0) {
                if (grow < P.rows) {
                    const float *mp = P.mask + grow * P.Ntot + col0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) mk[j] = glf_ldg4(mp + 4 * j);
                }
            };
            if (P.mask && cta_m < ntiles) mask_issue((int64_t)cta_m * GLT_TILE + qd * 32 + lane, n0);
            for (int64_t t = cta_m; t < ntiles; t += Gm) {
                const int64_t row0 = t * GLT_TILE + qd * 32;
                if (P.csum) {
                    const int st = (int)t / P.tiles_per_sample;
                    if (st != cur_sample) { csum_flush(); cur_sample = st; }
                }
                glt_mbar_wait(TFULL(a), aph);
                glt_tc_fence_after();
                const uint32_t tq = tmem_base + ((uint32_t)(qd * 32) << 16) + a * NT;
                for (int cb = 0; cb < NT; cb += 32) {
                    float z[32];
                    glt_tmem_ld16(tq + cb, z);
                    glt_tmem_ld16(tq + cb + 16, z + 16);
                    glt_tc_wait_ld();
                    if (cb + 32 >= NT) {          // last accumulator columns are in registers: release the TMEM stage
                        glt_tc_fence_before();
                        __syncwarp();
                        if (lane == 0) glt_mbar_arrive(TEMPTY(a));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bv = *reinterpret_cast<const float4 *>(Bs + cb + 4 * j);
                        z[4 * j] += bv.x; z[4 * j + 1] += bv.y; z[4 * j + 2] += bv.z; z[4 * j + 3] += bv.w;
                    }
                    if (P.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) z[j] = fmaxf(z[j], 0.f);
                    }
                    if (P.mask) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            z[4 * j] = mk[j].x > 0.f ? z[4 * j] : 0.f; z[4 * j + 1] = mk[j].y > 0.f ? z[4 * j + 1] : 0.f;
                            z[4 * j + 2] = mk[j].z > 0.f ? z[4 * j + 2] : 0.f; z[4 * j + 3] = mk[j].w > 0.f ? z[4 * j + 3] : 0.f;
                        }
                        // the mask registers are free: request the next slab's (or the next tile's first) row segment now
                        if (cb + 32 < NT) mask_issue(row0 + lane, n0 + cb + 32);
                        else if (t + Gm < ntiles) mask_issue((t + Gm) * GLT_TILE + qd * 32 + lane, n0);
                    }
                    // the staging tile of this slab was last used two slabs ago: its store must have finished reading
                    if (lane == 0) { if (bmask) glt_bulk_wait_read<1>(); else glt_bulk_wait_read<0>(); }
                    __syncwarp();
                    unsigned char *buf = ob + (nslab & bmask) * 4096;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4 *>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
                    glt_fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && !(P.dbg & 1)) {
                        glt_tma_store_2d(&tmO, ob_u32 + (nslab & bmask) * 4096, n0 + cb, (int)row0);
                        glt_bulk_commit();
                    }
                    if (P.csum) {                 // column (cb + lane) of the staged tile, valid rows only
                        const int nv = (int)nbpc_min((int64_t)32, P.rows - row0);
                        const unsigned char *colp = buf + cw;
                        float sum = 0.f;
                        if (nv == 32) {
#pragma unroll
                            for (int r = 0; r < 32; ++r) sum += *reinterpret_cast<const float *>(colp + r * 128 + ((cq ^ (r & 7)) << 4));
                        } else {
                            for (int r = 0; r < nv; ++r) sum += *reinterpret_cast<const float *>(colp + r * 128 + ((cq ^ (r & 7)) << 4));
                        }
                        Cw[cb + lane] += sum;
                    }
                    ++nslab;
                }
                if (++a == 2) { a = 0; aph ^= 1; }
            }
            if (lane == 0) glt_bulk_wait_all();
            __syncwarp();
        } else
        for (int64_t t = cta_m; t < ntiles; t += Gm) {
            const int64_t row0 = t * GLT_TILE + qd * 32;
            if (P.csum) {
                const int st = (int)t / P.tiles_per_sample;
                if (st != cur_sample) { csum_flush(); cur_sample = st; }
            }
            glt_mbar_wait(TFULL(a), aph);
            glt_tc_fence_after();
            const uint32_t tq = tmem_base + ((uint32_t)(qd * 32) << 16) + a * NT;
            if (P.dbg & 2) {
                glt_tc_fence_before();
                __syncwarp();
                if (lane == 0) glt_mbar_arrive(TEMPTY(a));
                if (++a == 2) { a = 0; aph ^= 1; }
                continue;
            }
            for (int cb = 0; cb < NT; cb += SW) {
                // the input-mask rows of this slab are requested first: 8 independent loads in flight (issued one by one in
                // front of each store they cost one HBM round trip per row segment: 5.5 us per slab)
                float4 mk[8];
                if (P.mask) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int idx = i * 32 + lane, r = idx >> vshift, c4 = idx & vmask;
                        if (i < iters && row0 + r < P.rows) mk[i] = glf_ldg4(P.mask + (row0 + r) * P.Ntot + n0 + cb + 4 * c4);
                    }
                }
                float z[32];
                glt_tmem_ld16(tq + cb, z);
                if (SW == 32) glt_tmem_ld16(tq + cb + 16, z + 16);
                glt_tc_wait_ld();
                if (cb + SW >= NT) {          // last accumulator columns are in registers: release the TMEM stage
                    glt_tc_fence_before();
                    __syncwarp();
                    if (lane == 0) glt_mbar_arrive(TEMPTY(a));
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (j < SW) {
                        float4 o = make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]);
                        if (P.bias) {
                            const float4 bv = glf_ldg4(P.bias + n0 + cb + j);
                            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                        }
                        if (P.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4 *>(Ow + lane * PITCH + j) = o;
                    }
                }
                __syncwarp();
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);       // this lane's rows of column group (lane & vmask)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = i * 32 + lane, r = idx >> vshift, c4 = idx & vmask;
                    const int64_t grow = row0 + r;
                    if (i < iters && grow < P.rows && !(P.dbg & 1)) {
                        float4 v = *reinterpret_cast<const float4 *>(Ow + r * PITCH + 4 * c4);
                        if (P.mask) {
                            const float4 m = mk[i];
                            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                        }
                        *reinterpret_cast<float4 *>(P.out + grow * P.Ntot + n0 + cb + 4 * c4) = v;
                        if (P.csum) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
                    }
                }
                if (P.csum) {   // lanes with the same column group hold different rows: butterfly, then lanes 0..vmask accumulate
                    for (int m = vmask + 1; m < 32; m <<= 1) {
                        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, m); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, m);
                        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, m); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, m);
                    }
                    if (lane <= vmask) {
                        float4 *acc = reinterpret_cast<float4 *>(Cw + cb + 4 * lane);
                        float4 o = *acc;
                        o.x += cs.x; o.y += cs.y; o.z += cs.z; o.w += cs.w;
                        *acc = o;
                    }
                }
                __syncwarp();
            }
            if (++a == 2) { a = 0; aph ^= 1; }
        }
        if (P.csum) csum_flush();
    } else {
        // ---------------- converter warps: centre the landed chunk in place (A - mu_s) and write the TF32 residual
        const int wtid = tid - 192;
        int s = 0, ph = 0, l = 0, lph = 0;
        if (X3 && LT) {
            // ---- residuals to TENSOR memory: a warp may touch the 32 TMEM lanes of its quadrant (warp id % 4), so here a thread
            // owns tile row 32 (warp % 4) + lane and half of the chunk's 32 columns (warps 6..9: columns 0..15, 10..13: 16..31).
            // It centres its 4 granules in place (shared memory keeps the hi operand) and stores the 16 residuals into its
            // lane of the slot's 32 columns; the MMA warp reads them as the A operand of the lo*hi products.
            static_assert(SGT_CONV_WARPS == 8, "two converter warps per TMEM quadrant");
            const int qd = warp & 3, half = (warp - 6) >> 2, r = qd * 32 + lane;
            const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16) + 2 * NT + 16 * half;
            for (int64_t t = cta_m; t < ntiles; t += Gm) {
                const int64_t trow = t * GLT_TILE;
                const int64_t grow = nbpc_min(trow + r, P.rows - 1);                  // rows past the end: any valid mean row
                const float *mur = P.mu ? P.mu + (grow / P.rows_per_sample) * K + 16 * half : nullptr;
                for (int c = 0; c < KC; ++c) {
                    float4 m[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) m[j] = P.mu ? glf_ldg4(mur + c * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    glt_mbar_wait(FULL(s), ph);
                    glt_mbar_wait(LOFREE(l), lph ^ 1);
                    glt_tc_fence_after();
                    unsigned char *hi = As + s * SGT_CHUNK_BYTES + r * 128;
                    float res[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 *px = reinterpret_cast<float4 *>(hi + (((4 * half + j) ^ (r & 7)) << 4));
                        float4 x = *px;
                        if (P.mu) {
                            x.x -= m[j].x; x.y -= m[j].y; x.z -= m[j].z; x.w -= m[j].w;
                            *px = x;
                        }
                        res[4 * j] = glt_residual(x.x); res[4 * j + 1] = glt_residual(x.y);
                        res[4 * j + 2] = glt_residual(x.z); res[4 * j + 3] = glt_residual(x.w);
                    }
                    glt_tmem_st16(t_lane + l * 32, res);
                    glt_tc_wait_st();
                    glt_tc_fence_before();
                    glt_fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) glt_mbar_arrive(CONV(s));
                    if (++s == S) { s = 0; ph ^= 1; }
                    if (++l == LT) { l = 0; lph ^= 1; }
                }
            }
        } else {
        // a thread's granules g = wtid + 256 i all lie in the same logical 16-byte unit of their rows (rows advance by 32)
        const int lu_t = (wtid & 7) ^ ((wtid >> 3) & 7);
        for (int64_t t = cta_m; t < ntiles; t += Gm) {
            // sample of tile row r = s0 + (rem0 + r) / rows_per_sample: one 64-bit division per tile, none per granule
            const int64_t trow = t * GLT_TILE;
            const int64_t s0 = trow / P.rows_per_sample;
            const uint32_t rem0 = (uint32_t)(trow - s0 * P.rows_per_sample), last = (uint32_t)nbpc_min((int64_t)GLT_TILE - 1, P.rows - 1 - trow);
            const bool one_sample = rem0 + last < (uint32_t)P.rows_per_sample;   // the usual case: one mean row for the whole tile
            for (int c = 0; c < KC; ++c) {
                const float *mu0 = P.mu ? P.mu + s0 * K + c * 32 : nullptr;
                float4 mt = make_float4(0.f, 0.f, 0.f, 0.f);
                if (P.mu && one_sample) mt = glf_ldg4(mu0 + 4 * lu_t);           // requested before the wait for the data
                glt_mbar_wait(FULL(s), ph);
                if (X3) glt_mbar_wait(LOFREE(l), lph ^ 1);
                float *hi = reinterpret_cast<float *>(As + s * SGT_CHUNK_BYTES);
                float *lo = reinterpret_cast<float *>(Al + l * SGT_CHUNK_BYTES);
#pragma unroll 4
                for (int g = (P.dbg & 4) ? SGT_CHUNK_BYTES : wtid; g < SGT_CHUNK_BYTES / 16; g += 32 * SGT_CONV_WARPS) {   // 16-byte granules: row = g / 8
                    float4 x = *reinterpret_cast<const float4 *>(hi + 4 * g);
                    if (P.mu) {
                        float4 m = mt;
                        if (!one_sample) {
                            const int r = g >> 3;
                            const uint32_t ds = sgt_div(rem0 + nbpc_min((uint32_t)r, last), (uint32_t)P.rows_per_sample, P.rps_magic);
                            m = glf_ldg4(mu0 + ds * K + 4 * lu_t);
                        }
                        x.x -= m.x; x.y -= m.y; x.z -= m.z; x.w -= m.w;
                        *reinterpret_cast<float4 *>(hi + 4 * g) = X3 ? x : make_float4(glt_to_tf32(x.x), glt_to_tf32(x.y), glt_to_tf32(x.z), glt_to_tf32(x.w));
                    } else if (!X3) {
                        *reinterpret_cast<float4 *>(hi + 4 * g) = make_float4(glt_to_tf32(x.x), glt_to_tf32(x.y), glt_to_tf32(x.z), glt_to_tf32(x.w));
                    }
                    if (X3) *reinterpret_cast<float4 *>(lo + 4 * g) = make_float4(glt_residual(x.x), glt_residual(x.y), glt_residual(x.z), glt_residual(x.w));
                }
                glt_fence_proxy_async();
                __syncwarp();
                if (lane == 0) glt_mbar_arrive(CONV(s));
                if (++s == S) { s = 0; ph ^= 1; }
                if (X3 && ++l == L) { l = 0; lph ^= 1; }
            }
        }
        }
    }
    glt_tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        glt_tc_fence_after();
        glt_tmem_dealloc(tmem_base, tmem_cols);
    }
}

static size_t sgt_gemm_smem(bool x3, int K, int NT, int S, int L, int obuf = 1) {
    return 1024 + (size_t)(S + (x3 ? L : 0)) * SGT_CHUNK_BYTES + (size_t)K * NT * 4 * (x3 ? 2 : 1) + (size_t)4 * obuf * 4096 + 4 * 256 * 4 + 256 * 4 + 8 * (3 * 8 + 8 + 4) + 64;
}

// K-major tensor map over A (rows, K): box = 128 rows x 32 floats, 128-byte swizzle
static int sgt_make_tmap_a(CUtensorMap *tm, const float *ptr, int64_t rows, int K) {
    glt_encode_fn_t enc = glt_encode_fn();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)GLT_TILE};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

// tensor map over out (rows, Nout) for the epilogue's tile stores: box = 32 rows x 32 floats, 128-byte swizzle
static int sgt_make_tmap_out(CUtensorMap *tm, const float *ptr, int64_t rows, int Nout) {
    glt_encode_fn_t enc = glt_encode_fn();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)Nout, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Nout * 4};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

bool sgt_gemm_shape_ok(int K, int Nout) { return K % 32 == 0 && K >= 32 && K <= 256 && Nout % 16 == 0 && Nout >= 16 && Nout <= 256; }

// out (rows, Nout) = act((A - mu_s) Bm + bias) [* (mask > 0)];  0 on success
// csum_parts (optional, sgt_gemm_csum_ok): the kernel also leaves per-sample column sums of `out` as *n_sets partial sets
// [set][sample][Nout] (sgt_colsum_from_parts adds them up in a fixed order); sgt_gemm_csum_floats(samples, Nout) floats
bool sgt_gemm_csum_ok(int rows_per_sample) { return rows_per_sample % GLT_TILE == 0; }
size_t sgt_gemm_csum_floats(int samples, int Nout) { return (size_t)gl_num_sms() * samples * Nout; }

int sgt_gemm(const float *A, const float *Bsrc, int b_transposed, const float *mu, const float *bias, const float *mask, int64_t rows,
             int rows_per_sample, int K, int Nout, int relu, int x3, float *out, float *csum_parts, int *n_sets, cudaStream_t stream) {
    if (!sgt_gemm_shape_ok(K, Nout)) return 1;
    if (csum_parts && !sgt_gemm_csum_ok(rows_per_sample)) return 1;
    // N tile: the resident weight tile (hi + lo in the split mode) must leave room for >= 2 + 2 stages
    int NT = Nout;
    const size_t budget = 227 * 1024;
    while (NT > 16 && sgt_gemm_smem(x3, K, NT, 2, 2) > budget) NT >>= 1;
    if (Nout % NT || NT % 16 || sgt_gemm_smem(x3, K, NT, 2, 2) > budget) return 1;
    // tuning override NBPC_SGT="NT,S,L" (read once)
    static int env_nt = -1, env_s = 0, env_l = 0;
    if (env_nt < 0) {
        const char *e = getenv("NBPC_SGT");
        env_nt = 0;
        if (e) sscanf(e, "%d,%d,%d", &env_nt, &env_s, &env_l);
    }
    if (env_nt >= 16 && env_nt < NT && Nout % env_nt == 0 && env_nt % 16 == 0) NT = env_nt;
    int S = 2, L = 2;
    while (S < 6 && sgt_gemm_smem(x3, K, NT, S + 1, L) <= budget) ++S;
    if (x3 && S >= 4 && sgt_gemm_smem(x3, K, NT, S - 1, L + 1) <= budget) { --S; ++L; }
    if (env_s >= 1 && env_l >= 1 && env_s <= 8 && env_l <= 8 && sgt_gemm_smem(x3, K, NT, env_s, env_l) <= budget) { S = env_s; L = env_l; }
    // TF32x3: the A residuals can live in tensor memory (next to the 2 NT accumulator columns) instead of a shared-memory
    // ring, which leaves all of the ring's shared memory to landing stages (NBPC_SGT_LOTMEM=0 keeps the shared-memory ring)
    int lo_tmem = 0;
    static int env_lt = -1;
    if (env_lt < 0) {
        const char *e = getenv("NBPC_SGT_LOTMEM");
        env_lt = e ? atoi(e) : 4;
    }
    if (x3 && env_lt > 0 && 2 * NT + 32 * env_lt <= 512 && rows_per_sample >= 1) {
        lo_tmem = env_lt;
        const int s_old = S + L;
        L = 0;
        S = 2;
        while (S < 8 && sgt_gemm_smem(x3, K, NT, S + 1, 0) <= budget) ++S;
        if (env_s >= 1 && env_s <= 8 && sgt_gemm_smem(x3, K, NT, env_s, 0) <= budget) S = env_s;
        (void)s_old;
    }
    // a second staging tile per epilogue warp (the store of slab i overlaps the assembly of slab i + 1) if it costs no stage
    const int obuf = sgt_gemm_smem(x3, K, NT, S, L, 2) <= budget ? 2 : 1;
    const size_t smem = sgt_gemm_smem(x3, K, NT, S, L, obuf);
    auto kern = x3 ? sgt_gemm_kernel<true> : sgt_gemm_kernel<false>;
    static bool configured[64][2];   // per device: function attributes belong to a context
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev][x3 ? 1 : 0]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        if (dev >= 0 && dev < 64) configured[dev][x3 ? 1 : 0] = true;
    }
    CUtensorMap tm, tmo;
    if (sgt_make_tmap_a(&tm, A, rows, K)) return 1;
    if (NT >= 32) {
        if (sgt_make_tmap_out(&tmo, out, rows, Nout)) return 1;
    } else {
        tmo = tm;   // unused by the 16-column epilogue
    }
    SgtGemmArgs P;
    P.Bsrc = Bsrc; P.mu = mu; P.bias = bias; P.mask = mask; P.out = out; P.rows = rows; P.rows_per_sample = rows_per_sample;
    P.K = K; P.NT = NT; P.Ntot = Nout; P.n_ntiles = Nout / NT; P.b_transposed = b_transposed; P.relu = relu; P.S = S; P.L = L;
    P.rps_magic = sgt_magic(rows_per_sample);
    P.obuf = obuf;
    P.lo_tmem = lo_tmem;
    static int dbg = -1;
    if (dbg < 0) {
        const char *e = getenv("NBPC_SGT_DEBUG");
        dbg = e ? atoi(e) : 0;
    }
    P.dbg = dbg;
    const int64_t ntiles = (rows + GLT_TILE - 1) / GLT_TILE;
    int64_t gm = gl_num_sms() / P.n_ntiles;
    if (gm < 1) gm = 1;
    if (gm > ntiles) gm = ntiles;
    P.csum = csum_parts; P.samples = (int)(rows / rows_per_sample); P.tiles_per_sample = nbpc_max(rows_per_sample / GLT_TILE, 1);
    if (csum_parts) {
        *n_sets = (int)gm;
        // a CTA row only writes the samples its tiles visit: with fewer tiles per sample than CTA rows some sets stay unwritten
        if (P.tiles_per_sample < gm && cudaMemsetAsync(csum_parts, 0, sizeof(float) * (size_t)gm * P.samples * Nout, stream) != cudaSuccess) return 1;
    }
    const int grid = (int)(gm * P.n_ntiles);
    NBPC_LAUNCH_N(NbpcKName(x3 ? "sgt_gemm_tf32x3" : "sgt_gemm_tf32", K, Nout).c_str(), kern, grid, SGT_THREADS, smem, stream, tm, tmo, P);
    return 0;
}

// ------------------------------------------------------------------ dW = (H - mu_s)^T dZ, per-CTA partials
struct SgtDwArgs {
    const float *mu;          // (samples, k) or nullptr
    float *partial;           // [grid][k][q]
    int64_t rows;
    int rows_per_sample, k, q, S, L;
    uint32_t rps_magic;
};

template <bool X3>
__global__ void __launch_bounds__(SGT_THREADS) sgt_dw_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmZ,
                                                     const SgtDwArgs P) {
    constexpr int R = SGT_DW_ROWS, RCH = R * 128;                 // bytes of one [R rows x 32 floats] chunk
    static_assert(R == 32, "the converter's granule -> (chunk, row) mapping assumes 32-row chunks");
    extern __shared__ __align__(16) unsigned char glt_smem_raw[];
    unsigned char *base = glt_smem_raw + ((1024 - (glt_smem_u32(glt_smem_raw) & 1023)) & 1023);
    const int S = P.S, L = X3 ? P.L : 0, k = P.k, q = P.q, HCH = k >> 5, ZCH = q >> 5;
    const int H_BYTES = HCH * RCH, STAGE = (HCH + ZCH) * RCH;
    unsigned char *St = base, *Sl = St + S * STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(Sl + L * STAGE);
    const uint32_t bar0 = glt_smem_u32(bars);
    const int LB = X3 ? P.L : 1;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (S + s); };
    auto CONV = [&](int s) { return bar0 + 8u * (2 * S + s); };
    auto LOFREE = [&](int l) { return bar0 + 8u * (3 * S + l); };
    const uint32_t DONE = bar0 + 8u * (3 * S + LB);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * S + LB + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (P.rows + R - 1) / R;
    const int G = gridDim.x;
    const int MB = (k + 127) / 128, M2 = k < 128 ? k : 128;       // k in {64, 128, 256}
    int tmem_cols = 32;
    while (tmem_cols < MB * q) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { glt_mbar_init(FULL(s), 1); glt_mbar_init(EMPTY(s), 1); glt_mbar_init(CONV(s), SGT_CONV_WARPS); }
        for (int l = 0; l < LB; ++l) glt_mbar_init(LOFREE(l), 1);
        glt_mbar_init(DONE, 1);
        glt_fence_barrier_init();
        glt_prefetch_tmap(&tmH);
        glt_prefetch_tmap(&tmZ);
    }
    if (warp == 1) glt_tmem_alloc(glt_smem_u32(tmem_slot), tmem_cols);
    glt_tc_fence_before();
    __syncthreads();
    glt_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---------------- TMA producer
            int s = 0, ph = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(EMPTY(s), ph ^ 1);
                glt_mbar_expect_tx(FULL(s), STAGE);
                unsigned char *st = St + s * STAGE;
                for (int ch = 0; ch < HCH; ++ch) glt_tma_load_2d(glt_smem_u32(st + ch * RCH), &tmH, FULL(s), ch * 32, (int)(t * R));
                for (int ch = 0; ch < ZCH; ++ch) glt_tma_load_2d(glt_smem_u32(st + H_BYTES + ch * RCH), &tmZ, FULL(s), ch * 32, (int)(t * R));
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---------------- MMA issuer: D[mb] (M2 x q) += H'[mb]^T (MN-major) * dZ' (MN-major)
            const uint32_t idesc = glt_idesc_tf32(M2, q, 1, 1);
            // descriptors built once, advanced through the address field (16-byte units): +64 per 8-row K step,
            // + 4 chunks per M block, + STAGE per ring slot
            const uint64_t h_hi0 = glt_smem_desc(glt_smem_u32(St), RCH, 512, 1), h_lo0 = glt_smem_desc(glt_smem_u32(Sl), RCH, 512, 1);
            const uint32_t st_step = (uint32_t)STAGE >> 4, z_off = (uint32_t)H_BYTES >> 4, mb_step = (uint32_t)(4 * RCH) >> 4;
            int s = 0, ph = 0, l = 0;
            uint32_t acc = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += G) {
                glt_mbar_wait(CONV(s), ph);
                glt_tc_fence_after();
                const uint64_t h_hi = h_hi0 + (uint64_t)(s * st_step), z_hi = h_hi + z_off;
                const uint64_t h_lo = h_lo0 + (uint64_t)(l * st_step), z_lo = h_lo + z_off;
                for (int mb = 0; mb < MB; ++mb) {
                    uint32_t a2 = acc;
#pragma unroll
                    for (int pass = X3 ? 0 : 2; pass < 3; ++pass) {
                        const uint64_t hb = ((pass == 0) ? h_lo : h_hi) + (uint64_t)(mb * mb_step), zb = (pass == 1) ? z_lo : z_hi;
#pragma unroll
                        for (int k8 = 0; k8 < R / 8; ++k8) {   // 8 rows per MMA = two 4-row swizzle atoms (SBO 512); LBO = chunk stride
                            glt_mma_tf32(tmem_base + mb * q, hb + 64 * k8, zb + 64 * k8, idesc, a2);
                            a2 = 1;
                        }
                    }
                }
                acc = 1;
                glt_tc_commit(EMPTY(s));
                if (X3) { glt_tc_commit(LOFREE(l)); if (++l == L) l = 0; }
                if (++s == S) { s = 0; ph ^= 1; }
            }
            glt_tc_commit(DONE);
        }
    } else if (warp < 6) {
        // ---------------- epilogue warps: one read-out of the accumulated (k x q) partial at the end
        const int qd = warp & 3;
        glt_mbar_wait(DONE, 0);
        glt_tc_fence_after();
        const bool any_tile = blockIdx.x < ntiles;
        for (int mb = 0; mb < MB; ++mb) {
            // M = 128: row r in lane r;  M = 64: row r in lane (r % 16) + 32 (r / 16)
            const int row = (M2 == 128) ? qd * 32 + lane : (lane < 16 ? qd * 16 + lane : -1);
            const uint32_t tq = tmem_base + ((uint32_t)(qd * 32) << 16) + mb * q;
            float *dst = P.partial + ((int64_t)blockIdx.x * k + mb * 128 + (row < 0 ? 0 : row)) * q;
            for (int cb = 0; cb < q; cb += 16) {
                float z[16];
                glt_tmem_ld16(tq + cb, z);
                glt_tc_wait_ld();
                if (row >= 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4 *>(dst + cb + j) = any_tile ? make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    } else {
        // ---------------- converter warps: centre H in place, write the TF32 residuals of H and dZ
        const int wtid = tid - 192;
        int s = 0, ph = 0, l = 0, lph = 0;
        static_assert(32 * SGT_CONV_WARPS == 256, "one H granule per chunk and thread");
        // granule g = wtid + 256 i -> (chunk i, row r_t, 16-byte unit u_t): row and unit are per-thread constants; 32-byte swizzle
        // atoms: logical atom = atom ^ (row & 3)
        const int r_t = (wtid >> 3) & (R - 1), u_t = wtid & 7;
        const int col_t = (((u_t >> 1) ^ (r_t & 3)) << 3) | ((u_t & 1) << 2);
        for (int64_t t = blockIdx.x; t < ntiles; t += G) {
            const int n_gran = STAGE / 16;
            const int64_t trow = t * R;
            const int n_valid = (int)nbpc_min((int64_t)R, P.rows - trow);
            // this thread's row of the tile and its mean row, requested before the wait for the data (rows beyond the tensor
            // were zero-filled by TMA and must stay zero)
            const bool centre = P.mu && r_t < n_valid;
            float4 m[8];
            if (centre) {
                const float *mur = P.mu + ((trow + r_t) / P.rows_per_sample) * k + col_t;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < HCH) m[i] = glf_ldg4(mur + 32 * i);
            }
            glt_mbar_wait(FULL(s), ph);
            if (X3) glt_mbar_wait(LOFREE(l), lph ^ 1);
            float *hi = reinterpret_cast<float *>(St + s * STAGE);
            float *lo = reinterpret_cast<float *>(Sl + l * STAGE);
#pragma unroll
            for (int i = 0; i < 8; ++i) {                   // the H chunks
                if (i < HCH) {
                    const int g = wtid + 256 * i;
                    float4 x = *reinterpret_cast<const float4 *>(hi + 4 * g);
                    if (centre) { x.x -= m[i].x; x.y -= m[i].y; x.z -= m[i].z; x.w -= m[i].w; }
                    if (!X3) x = make_float4(glt_to_tf32(x.x), glt_to_tf32(x.y), glt_to_tf32(x.z), glt_to_tf32(x.w));
                    if (P.mu || !X3) *reinterpret_cast<float4 *>(hi + 4 * g) = x;
                    if (X3) *reinterpret_cast<float4 *>(lo + 4 * g) = make_float4(glt_residual(x.x), glt_residual(x.y), glt_residual(x.z), glt_residual(x.w));
                }
            }
#pragma unroll 4
            for (int g = HCH * 256 + wtid; g < n_gran; g += 256) {   // the dZ chunks
                float4 x = *reinterpret_cast<const float4 *>(hi + 4 * g);
                if (!X3) {
                    x = make_float4(glt_to_tf32(x.x), glt_to_tf32(x.y), glt_to_tf32(x.z), glt_to_tf32(x.w));
                    *reinterpret_cast<float4 *>(hi + 4 * g) = x;
                } else {
                    *reinterpret_cast<float4 *>(lo + 4 * g) = make_float4(glt_residual(x.x), glt_residual(x.y), glt_residual(x.z), glt_residual(x.w));
                }
            }
            glt_fence_proxy_async();
            __syncwarp();
            if (lane == 0) glt_mbar_arrive(CONV(s));
            if (++s == S) { s = 0; ph ^= 1; }
            if (X3 && ++l == L) { l = 0; lph ^= 1; }
        }
    }
    glt_tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        glt_tc_fence_after();
        glt_tmem_dealloc(tmem_base, tmem_cols);
    }
}

static size_t sgt_dw_smem(bool x3, int k, int q, int S, int L) {
    return 1024 + (size_t)(S + (x3 ? L : 0)) * (size_t)(k + q) * SGT_DW_ROWS * 4 + 8 * (3 * 8 + 8 + 2) + 64;
}

bool sgt_dw_shape_ok(int k, int q) { return (k == 64 || k == 128 || k == 256) && q % 32 == 0 && q >= 32 && q <= 256 && ((k + 127) / 128) * q <= 512; }
int sgt_dw_max_parts() { return gl_num_sms(); }

// dW (k, q) = (H - mu_s)^T dZ over `rows` rows; partial: workspace of sgt_dw_max_parts() * k * q floats.  0 on success
int sgt_dw(const float *H, const float *dZ, const float *mu, int64_t rows, int rows_per_sample, int k, int q, int x3, float *partial,
           float *dW, cudaStream_t stream) {
    if (!sgt_dw_shape_ok(k, q)) return 1;
    const size_t budget = 227 * 1024;
    int S = 2, L = 1;
    if (sgt_dw_smem(x3, k, q, S, L) > budget) return 1;
    while (S < 4 && sgt_dw_smem(x3, k, q, S + 1, L) <= budget) ++S;
    if (x3 && S >= 3 && sgt_dw_smem(x3, k, q, S - 1, L + 1) <= budget) { --S; ++L; }
    const size_t smem = sgt_dw_smem(x3, k, q, S, L);
    auto kern = x3 ? sgt_dw_kernel<true> : sgt_dw_kernel<false>;
    static bool configured[64][2];   // per device: function attributes belong to a context
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev][x3 ? 1 : 0]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        if (dev >= 0 && dev < 64) configured[dev][x3 ? 1 : 0] = true;
    }
    CUtensorMap tmH, tmZ;
    if (glt_make_tmap_packed(&tmH, H, rows, k, SGT_DW_ROWS) || glt_make_tmap_packed(&tmZ, dZ, rows, q, SGT_DW_ROWS)) return 1;
    SgtDwArgs P;
    P.mu = mu; P.partial = partial; P.rows = rows; P.rows_per_sample = rows_per_sample; P.k = k; P.q = q; P.S = S; P.L = L;
    P.rps_magic = sgt_magic(rows_per_sample);
    const int64_t ntiles = (rows + SGT_DW_ROWS - 1) / SGT_DW_ROWS;
    const int grid = (int)nbpc_min((int64_t)gl_num_sms(), ntiles);
    NBPC_LAUNCH_N(NbpcKName(x3 ? "sgt_dw_tf32x3" : "sgt_dw_tf32", k, q).c_str(), kern, grid, SGT_THREADS, smem, stream, tmH, tmZ, P);
    // fixed-order sum of the per-CTA partials (32 outputs x 32 slices per block)
    NBPC_LAUNCH_N("sgt_dw_final_kernel", sgt_colsum_from_parts_kernel, nbpc_cdiv(k * q, 32), 1024, 0, stream, (const float *)partial, grid, k * q, 1.0f, dW, (float *)nullptr);
    return 0;
}

// ------------------------------------------------------------------ per-sample column sums (C % 4 == 0, C <= 1024)
int sgt_colsum_blocks(int N) { return nbpc_cdiv(N, SGT_COLSUM_ROWS); }
void sgt_colsum(const float *X, int C, int N, int B, float scale, float *partial, float *out, float *total, cudaStream_t stream) {
    const int nblk = sgt_colsum_blocks(N);
    NBPC_LAUNCH(sgt_colsum_partial_kernel, dim3(nblk, B), 256, 0, stream, X, C, N, SGT_COLSUM_ROWS, partial);
    // the unscaled per-sample sums go behind the partials (the caller's buffer holds B * (nblk + 1) * C floats)
    float *sums = total ? partial + (size_t)B * nblk * C : nullptr;
    NBPC_LAUNCH(sgt_colsum_final_kernel, nbpc_cdiv((int64_t)B * C * 32, 256), 256, 0, stream, partial, C, nblk, B, scale, out, sums);
    if (total) NBPC_LAUNCH(sgt_colsum_total_kernel, nbpc_cdiv(C, 128), 128, 0, stream, sums, C, B, total);
}
#endif  // !NBPC_HOST_EMU
