// graph_layer_tc.cuh - tcgen05 / TMEM / TMA kernels for the edge-level GEMMs of the shift-invariant
// graph layer (graph.py:394-456) on sm_100a.
//
//   forward :  out[e] = act( H[e] W1 + Q_col[col[e]] + Q_row[e / M] )
//   backward:  dH[e]  = ( dZ[e] W1^T + G_col[col[e]] + G_row[e / M] ) [* (H[e] > 0)]
//              dW1    = H^T dZ                     (per-block partials, reduced in a fixed order afterwards)
//
// One persistent CTA walks 128-edge tiles (edge e = TMEM lane).  Warp roles:
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor tiles of H (and dZ) into a ring of shared-memory
//                      stages, 128-byte (64-byte for 16-wide rows) swizzled, mbarrier complete_tx;
//   warp 1 (one lane)  MMA issuer: tcgen05.mma.kind::tf32, A = tile (K-major for the projections, MN-major =
//                      "transposed in place" for H^T dZ), B = the weight matrix written once per CTA in the same
//                      canonical swizzled layout, FP32 accumulators in TMEM (double buffered);
//   warps 2..5         epilogue: tcgen05.ld 32x32b (thread = edge row), add the gathered node-level terms,
//                      activation / mask, stage through padded shared memory, 512-byte coalesced stores.
// Precision modes (nbpc_set_math_mode):
//   NBPC_MATH_TF32   one tensor-core pass on the FP32 bit patterns (10-bit mantissa operands, FP32 accumulate);
//   NBPC_MATH_TF32X3 error-compensated split x = hi + lo (both exactly representable in TF32):
//                    D = A_lo B_hi + A_hi B_lo + A_hi B_hi  -> FP32-class accuracy; the split is an in-place
//                    pass over the landed tile by the epilogue warps (position preserving, so it is
//                    independent of the swizzle), followed by fence.proxy.async.
// The GEMMs are HBM-bound by two orders of magnitude (5-8 flop/B): the tensor pipe is used to take the FMA
// work off the issue slots so that the kernel can run at memory speed, not for its flop rate.
#pragma once
#ifndef NBPC_HOST_EMU
#include <stdio.h>
#include <stdlib.h>
#include <cuda.h>   // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no -lcuda)

#include "nbpc_common.cuh"
#include "graph_layer_tc.h"

#define GLT_THREADS 192      // TMA warp, MMA warp, 4 epilogue warps
#define GLT_THREADS_X3 320   // + 4 converter warps (TF32 hi / lo split)
#define GLT_TILE 128
#define GLT_SPIN_LIMIT (1u << 22)


// small helpers shared with the CUDA-core kernels' conventions (graph_layer_fast.cuh)
__device__ __forceinline__ float4 glf_ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__host__ __device__ constexpr int glf_stride(int C) { return (C % 8 == 0) ? C + 4 : C + 8; }
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
static int gl_num_sms() {
    static int sms_d[NBPC_MAX_DEVICES];
    int &sms = sms_d[nbpc_device_slot()];
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t glt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void glt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void glt_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void glt_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void glt_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void glt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void glt_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < GLT_SPIN_LIMIT; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    printf("libnbpc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void glt_tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :
                 : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// smem -> global tile store (bulk async group of the issuing thread); OOB parts of the box are clipped
__device__ __forceinline__ void glt_tma_store_2d(const CUtensorMap *tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void glt_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void glt_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void glt_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void glt_prefetch_tmap(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void glt_tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void glt_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void glt_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void glt_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void glt_tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void glt_tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 operands, FP32 accumulate
__device__ __forceinline__ void glt_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M rows in the lanes, K 32-bit elements in consecutive columns) is read from
// tensor memory
__device__ __forceinline__ void glt_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp writes lane (quadrant base + i)
__device__ __forceinline__ void glt_tmem_st16(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void glt_tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (quadrant base + i)
__device__ __forceinline__ void glt_tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <int C>
__device__ __forceinline__ void glt_tmem_ld(uint32_t taddr, float *v) {
    static_assert(C % 16 == 0, "column count must be a multiple of 16");
#pragma unroll
    for (int j = 0; j < C / 16; ++j) glt_tmem_ld16(taddr + 16 * j, v + 16 * j);
}

// ------------------------------------------------------------------ descriptors
// shared-memory matrix descriptor (sm_100 "version 1"): start address, leading / stride byte offsets (all >> 4),
// swizzle mode in bits [61,64): 2 = 128-byte, 4 = 64-byte
__device__ __forceinline__ uint64_t glt_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swizzle) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swizzle << 61;
    return d;
}
// instruction descriptor for kind::tf32: FP32 accumulator, TF32 A/B, majors (0 = K, 1 = MN), N >> 3, M >> 4
__host__ __device__ constexpr uint32_t glt_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// geometry of a (rows x C) FP32 operand tile cut into column chunks of CW = min(C, 32) floats; each chunk is a
// [rows][CW] block whose 16-byte units are XOR-swizzled inside 8-row atoms (what TMA SWIZZLE_128B / _64B writes
// and what the UMMA descriptor modes 2 / 4 read)
template <int C>
struct GltTile {
    static_assert(C == 16 || C % 32 == 0, "row width must be 16 or a multiple of 32 floats");
    static constexpr int CW = C < 32 ? C : 32;          // floats per chunk row
    static constexpr int NCH = C / CW;                  // chunks
    static constexpr int PITCH = CW * 4;                // bytes per chunk row
    static constexpr int ATOM = 8 * PITCH;              // bytes per 8-row swizzle atom
    static constexpr uint32_t SWZ = (CW == 32) ? 2u : 4u;
    __host__ __device__ static constexpr int chunk_bytes(int rows) { return rows * PITCH; }
    // byte offset of float j of row r inside one chunk
    __device__ static __forceinline__ int offset(int r, int j) {
        const int unit = j >> 2;
        const int sw = (CW == 32) ? (r & 7) : ((r >> 1) & 3);
        return r * PITCH + ((unit ^ sw) << 4) + ((j & 3) << 2);
    }
};

// in-place error-compensated split of `bytes` of FP32 data at `hi` (16-byte granules, one per thread per step):
// hi <- rna_tf32(x), lo <- rna_tf32(x - hi); same positions in both buffers
__device__ __forceinline__ float glt_to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Error-compensated split of a landed FP32 tile for the TF32 tensor pipe.  tcgen05 kind::tf32 reads the top 19 bits of
// each 32-bit operand (the low 13 mantissa bits are ignored), so the landed tile itself already IS the "hi" operand
// trunc(x); only the residual lo = x - trunc(x) (exact in FP32, < 2^-10 |x|; the pipe keeps its top 11 bits) has to be
// produced, at the same position of a second buffer.  hi * hi + hi * lo + lo * hi then carries ~2^-20 relative error per
// product instead of 2^-11.  128 converter threads, one 16-byte granule per thread per step, 4 loads in flight.
__device__ __forceinline__ float glt_residual(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
template <int N_FLOATS, int NTHREADS = 128>
__device__ __forceinline__ void glt_split_inplace(const float *hi, float *lo, int tid) {
    constexpr int STEP = NTHREADS * 4, NSTEP = N_FLOATS / STEP;
    static_assert(N_FLOATS % STEP == 0, "tile size must be a multiple of 4 floats per converter thread");
    int s = 0;
#pragma unroll 1
    for (; s + 4 <= NSTEP; s += 4) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4 *>(hi + (s + u) * STEP + tid * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<float4 *>(lo + (s + u) * STEP + tid * 4) =
                make_float4(glt_residual(x[u].x), glt_residual(x[u].y), glt_residual(x[u].z), glt_residual(x[u].w));
    }
    for (; s < NSTEP; ++s) {
        const float4 x = *reinterpret_cast<const float4 *>(hi + s * STEP + tid * 4);
        *reinterpret_cast<float4 *>(lo + s * STEP + tid * 4) = make_float4(glt_residual(x.x), glt_residual(x.y), glt_residual(x.z), glt_residual(x.w));
    }
}

// write a (R x C) weight operand (element (r, j) = src(r, j)) in the canonical K-major swizzled layout, split hi / lo
template <int C, class F>
__device__ __forceinline__ void glt_fill_operand(char *hi, char *lo, int R, F src, int tid, int nthreads) {
    using T = GltTile<C>;
    for (int i = tid; i < R * C; i += nthreads) {
        const int r = i / C, jj = i % C;
        const int ch = jj / T::CW, j = jj % T::CW;
        const int off = ch * T::chunk_bytes(R) + T::offset(r, j);
        const float x = src(r, jj);
        // split mode: hi = the raw value (the pipe truncates it), lo = the exact residual; single pass: round to nearest
        *reinterpret_cast<float *>(hi + off) = lo ? x : glt_to_tf32(x);
        if (lo) *reinterpret_cast<float *>(lo + off) = glt_residual(x);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*glt_encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static glt_encode_fn_t glt_encode_fn() {
    static glt_encode_fn_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (glt_encode_fn_t)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// tensor map over a (rows, C) row-major FP32 tensor, box = 128 rows x min(C, 32) columns, swizzle = the box row width
template <int C>
static int glt_make_tmap(CUtensorMap *tm, const float *ptr, int64_t rows) {
    using T = GltTile<C>;
    glt_encode_fn_t enc = glt_encode_fn();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {(cuuint32_t)T::CW, (cuuint32_t)GLT_TILE};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = (T::CW == 32) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0 : 1;
}

// tensor map over the packed view (rows, W) of an FP32 tensor, W a multiple of 32 floats: box = box_rows x 32 floats,
// 128-byte swizzle with 32-byte atoms (the only layout tcgen05 reads MN-major 32-bit operands from)
static int glt_make_tmap_packed(CUtensorMap *tm, const float *ptr, int64_t rows, int W, int box_rows) {
    glt_encode_fn_t enc = glt_encode_fn();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0 : 1;
}

// tuning override "ctas,stages[,lo]" from the environment (unset or malformed: keep the defaults)
static void glt_env_cfg(const char *name, int *ctas, int *stages, int *lo) {
    const char *e = getenv(name);
    int a = 0, b = 0, l = 0;
    const int n = e ? sscanf(e, "%d,%d,%d", &a, &b, &l) : 0;
    if (n >= 2 && a >= 1 && b >= 1) { *ctas = a; *stages = b; }
    if (n == 3 && l >= 1) *lo = l;
}

// persistent grid = SMs x co-resident CTAs.  The co-residency is computed from the SM's budgets (shared memory with
// the carve-out forced to its maximum, registers, threads) instead of cudaOccupancyMaxActiveBlocksPerMultiprocessor,
// which answered 1 for these kernels (it evaluates the default carve-out).  NBPC_GLT_DEBUG=1 prints both.
template <class F>
static int glt_grid(F kern, int threads, size_t smem, int ctas_per_sm_cap) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int api = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&api, kern, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        api = 0;
    }
    const int regs_per_thread = (fa.numRegs + 7) / 8 * 8;
    const int by_regs = 65536 / (regs_per_thread * ((threads + 31) / 32 * 32));
    const int by_smem = (int)((228 * 1024) / (smem + fa.sharedSizeBytes + 1024));
    const int by_threads = 2048 / threads;
    int occ = by_regs < by_smem ? by_regs : by_smem;
    occ = occ < by_threads ? occ : by_threads;
    if (occ > ctas_per_sm_cap) occ = ctas_per_sm_cap;
    if (getenv("NBPC_GLT_DEBUG"))
        fprintf(stderr, "libnbpc: glt_grid threads=%d smem=%zu regs=%d: occupancy api=%d regs=%d smem=%d threads=%d cap=%d -> %d CTAs/SM\n",
                threads, smem, fa.numRegs, api, by_regs, by_smem, by_threads, ctas_per_sm_cap, occ);
    if (occ < 1) return -1;
    return gl_num_sms() * occ;
}

#define GLT_FOR_KQ(X) X(16, 16) X(16, 32) X(16, 64) X(32, 16) X(32, 32) X(32, 64) X(64, 16) X(64, 32) X(64, 64)
#endif  // !NBPC_HOST_EMU
