// set_layer.cu - permutation-equivariant set layer forward/backward (nn.py:10-28):
//   out = (H - mean_N H) W + B   on (B,N,k) -> (B,N,q), optional ReLU (nn.py:59, 65-66).
// Backward:  dZ = dOut * [out > 0];  dB = sum dZ;  dW = (H - mu)^T dZ;
//            dH = (dZ - colmean_s(dZ)) W^T   (the mean-subtraction's adjoint).
// Barrier-free baseline kernels (one thread per output element); all reductions fixed-order.
#include <stdlib.h>

#include "nbpc_common.cuh"
#include "reduce.cuh"
#include "set_layer_tc.h"
#include "set_layer_small.cuh"

// NBPC_BASELINE=1 (read once) keeps the barrier-free baseline kernels on the CUDA-core paths (cross-check)
static bool gl_set_fast() {
#ifdef NBPC_HOST_EMU
    return false;
#else
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("NBPC_BASELINE");
        cached = (e && e[0] == '1') ? 0 : 1;
    }
    return cached == 1;
#endif
}

struct SetXCentered {  // (H - mu) accessor
    const float *h, *mu;
    int k, N;
    __device__ __forceinline__ float at(int64_t r, int kk) const { return h[r * k + kk] - mu[(r / N) * k + kk]; }
};

struct SetDz {  // dOut masked by the ReLU of the forward output
    const float *g, *hout;
    int q, relu;
    __device__ __forceinline__ float at(int64_t r, int qo) const {
        float v = g[r * q + qo];
        if (relu && !(hout[r * q + qo] > 0.f)) v = 0.f;
        return v;
    }
};

__global__ void set_fwd_kernel(SetXCentered X, const float *__restrict__ W, const float *__restrict__ bias,
                               int64_t rows, int k, int q, int relu, float *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * q) return;
    const int64_t r = t / q;
    const int qo = (int)(t % q);
    float z = 0.f;
    for (int kk = 0; kk < k; ++kk) z += X.at(r, kk) * __ldg(&W[kk * q + qo]);
    z += __ldg(&bias[qo]);
    out[t] = (relu && z < 0.f) ? 0.f : z;
}

// partial[s][blk][qo] = sum over rows of the chunk of dZ
__global__ void set_dz_partial_kernel(SetDz dz, int q, int N, int nblk, int B, float *__restrict__ partial) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * nblk * q) return;
    const int qo = (int)(t % q);
    const int blk = (int)((t / q) % nblk);
    const int s = (int)(t / ((int64_t)q * nblk));
    const int n0 = blk * GL_CUBE_CHUNK, n1 = nbpc_min(n0 + GL_CUBE_CHUNK, N);
    float acc = 0.f;
    for (int n = n0; n < n1; ++n) acc += dz.at((int64_t)s * N + n, qo);
    partial[t] = acc;
}

__global__ void set_bias_kernel(const float *__restrict__ colsum, int B, int q, float *__restrict__ dB) {
    int qo = blockIdx.x * blockDim.x + threadIdx.x;
    if (qo >= q) return;
    float acc = 0.f;
    for (int s = 0; s < B; ++s) acc += colsum[s * q + qo];
    dB[qo] = acc;
}

__global__ void set_bwd_in_kernel(SetDz dz, const float *__restrict__ colsum, const float *__restrict__ W,
                                  int64_t rows, int N, int k, int q, float *__restrict__ dH) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * k) return;
    const int64_t r = t / k;
    const int kk = (int)(t % k);
    const int s = (int)(r / N);
    float a = 0.f;
    for (int qo = 0; qo < q; ++qo)
        a += (dz.at(r, qo) - __ldg(&colsum[s * q + qo]) / (float)N) * __ldg(&W[kk * q + qo]);
    dH[t] = a;
}

// dZ = dOut * [H_out > 0] materialised (tensor-core path of a ReLU layer whose gradient does not arrive pre-masked)
__global__ void set_mask_kernel(const float *__restrict__ g, const float *__restrict__ hout, int64_t n4, float *__restrict__ dz) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    float4 v = reinterpret_cast<const float4 *>(g)[t];
    const float4 h = reinterpret_cast<const float4 *>(hout)[t];
    v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
    reinterpret_cast<float4 *>(dz)[t] = v;
}
// ReLU backward of the layer that produced H_in, applied to dH_in (CUDA-core path)
__global__ void set_mask_input_kernel(const float *__restrict__ H, int64_t n, float *__restrict__ dH) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && !(H[t] > 0.f)) dH[t] = 0.f;
}

// same as set_bwd_in_kernel with the per-sample column MEANS of dZ precomputed
__global__ void set_bwd_in_mean_kernel(SetDz dz, const float *__restrict__ colmean, const float *__restrict__ W,
                                       int64_t rows, int N, int k, int q, float *__restrict__ dH) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * k) return;
    const int64_t r = t / k;
    const int kk = (int)(t % k);
    const int s = (int)(r / N);
    float a = 0.f;
    for (int qo = 0; qo < q; ++qo) a += (dz.at(r, qo) - __ldg(&colmean[s * q + qo])) * __ldg(&W[kk * q + qo]);
    dH[t] = a;
}

#ifndef NBPC_HOST_EMU
static int set_num_sms() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    return sms;
}
// dW = (H - mu)^T dZ when one side is narrow (<= 16); false if no instance exists
static bool set_xty_narrow(const float *H, const float *mu, const float *dZ, int B, int N, int k, int q, float *partial, float *dW,
                           cudaStream_t stream) {
    const int sms = set_num_sms();
    const int rpb = sgs_rows_per_block(N, B, sms), nblk = nbpc_cdiv(N, rpb);
    bool done = false;
    if (q <= 16 && k % 4 == 0 && k <= 1024 && sgs_narrow_ok(q)) {
#define X(V) if (q == V) NBPC_LAUNCH_N(NbpcKName("sgs_xty_narrow_q", k, q).c_str(), (sgs_xty_narrow_kernel<V, true>), dim3(nblk, B), SGS_THREADS, 0, stream, H, mu, dZ, N, k, q, rpb, partial);
        SGS_NARROW_LIST(X)
#undef X
        done = true;
    } else if (k <= 16 && q % 4 == 0 && q <= 1024 && sgs_narrow_ok(k)) {
#define X(V) if (k == V) NBPC_LAUNCH_N(NbpcKName("sgs_xty_narrow_k", k, q).c_str(), (sgs_xty_narrow_kernel<V, false>), dim3(nblk, B), SGS_THREADS, 0, stream, H, mu, dZ, N, k, q, rpb, partial);
        SGS_NARROW_LIST(X)
#undef X
        done = true;
    }
    if (done) NBPC_LAUNCH(sgs_sum_partials_kernel, nbpc_cdiv(k * q, 32), 256, 0, stream, partial, nblk * B, k * q, dW);
    return done;
}
#endif

#ifndef NBPC_HOST_EMU
// ---- 16-wide outputs on the tensor pipe by ROW PAIRING.  The tcgen05 kernels want 32-float (128-byte) rows on the
// contracted / narrow side.  Two consecutive rows of a (rows x 16) tensor ARE one row of its (rows/2 x 32) view, and
// two rows of the (rows x k) side one row of a (rows/2 x 2k) view; with the block-diagonal weight W2 = diag(W, W)
//   dH' = (dZ' - [m, m]) W2^T            is the (rows/2 x 2k) view of dH = (dZ - m) W^T, and
//   dW2 = (H' - [mu, mu])^T dZ'  (2k x 32) carries dW in its two diagonal blocks (even rows / odd rows of the sample).
// The zero blocks contribute exact zeros, so the arithmetic of every real term is that of the unpaired kernels.
static bool set_pair_ok(int N, int k, int q) {
    return q == 16 && N % 2 == 0 && sgt_gemm_shape_ok(32, 2 * k) && sgt_dw_shape_ok(2 * k, 32);
}
__global__ void set_pair_prepare_kernel(const float *__restrict__ W, const float *__restrict__ colmean, const float *__restrict__ mu,
                                        int B, int k, float *__restrict__ W2, float *__restrict__ cm2, float *__restrict__ mu2) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * k * 32) {                       // W2 (2k x 32) = diag(W, W), W (k x 16)
        const int a = t / 32, b = t % 32;
        const bool lo = a < k && b < 16, hi = a >= k && b >= 16;
        W2[t] = lo ? W[a * 16 + b] : (hi ? W[(a - k) * 16 + (b - 16)] : 0.f);
    }
    if (t < B * 32) cm2[t] = colmean[(t / 32) * 16 + (t % 16)];
    if (t < B * 2 * k) mu2[t] = mu[(t / (2 * k)) * k + (t % k)];
}
__global__ void set_pair_fold_kernel(const float *__restrict__ dW2, int k, float *__restrict__ dW) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * 16) return;
    const int a = t / 16, b = t % 16;
    dW[t] = dW2[a * 32 + b] + dW2[(k + a) * 32 + 16 + b];
}
#endif

struct SetWorkspace {
    float *csum_parts; // fused column-sum partials of the tcgen05 row GEMM (sgt_gemm_csum_floats(B, max(k,q)))
    float *pair;      // row-pairing scratch: W2 (2k x 32) | dW2 (2k x 32) | cm2 (B x 32) | mu2 (B x 2k)
    float *partial;   // (B, nblk, max(k,q))
    float *colsum;    // (B, q)
    float *xty_partial;
    float *dz;        // (B*N, q): masked dZ of the tensor-core path (q % 4 == 0 only)
    size_t bytes;
};

static SetWorkspace set_carve(void *ws, size_t ws_bytes, int B, int N, int k, int q) {
    NbpcArena a(ws, ws_bytes);
    SetWorkspace w;
    const int mx = k > q ? k : q;
    const int nblk = nbpc_cdiv(N, GL_CUBE_CHUNK);
    w.partial = a.take<float>((size_t)B * (nblk + 1) * mx);   // + the per-sample sums of sgt_colsum
    w.colsum = a.take<float>((size_t)B * mx);
    int rpc, nc;
    xty_plan((int64_t)B * N, k, q, &rpc, &nc);
    size_t nparts = (size_t)nc;
    size_t part_elems = (size_t)k * q, pair_elems = 0;
#ifndef NBPC_HOST_EMU
    nparts = nbpc_max(nparts, (size_t)sgt_dw_max_parts());
    nparts = nbpc_max(nparts, (size_t)sgs_max_blocks(N, B, set_num_sms()));
    if (set_pair_ok(N, k, q)) {
        part_elems = (size_t)2 * k * 32;         // the paired dW partials are (2k x 32)
        pair_elems = (size_t)2 * (2 * k * 32) + (size_t)B * 32 + (size_t)B * 2 * k;
    }
#endif
    w.pair = a.take<float>(pair_elems);
    size_t csum_elems = 0;
#ifndef NBPC_HOST_EMU
    if (sgt_gemm_csum_ok(N)) csum_elems = sgt_gemm_csum_floats(B, mx);
#endif
    w.csum_parts = a.take<float>(csum_elems);
    w.xty_partial = a.take<float>(nparts * part_elems);
    w.dz = a.take<float>(q % 4 == 0 ? (size_t)B * N * q : 0);
    w.bytes = a.off;
    return w;
}

// colmean[s][c] = sums[s][c] / N, dB[c] = sum_s sums[s][c] (per-sample column sums handed over by the producer of dOut)
__global__ void set_from_sums_kernel(const float *__restrict__ sums, int B, int q, float inv_n, float *__restrict__ colmean,
                                     float *__restrict__ dB) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * q) colmean[t] = sums[t] * inv_n;
    if (t < q) {
        float a = 0.f;
        for (int s = 0; s < B; ++s) a += sums[s * q + t];
        dB[t] = a;
    }
}

// out[s][c] = per-sample column mean (mean != 0) or sum of X (B, N, C)
static void set_colstat(const float *X, int C, int N, int B, int mean, float *partial, float *out, cudaStream_t stream) {
#ifndef NBPC_HOST_EMU
    if (C % 4 == 0 && C <= 1024) {
        sgt_colsum(X, C, N, B, mean ? 1.0f / (float)N : 1.0f, partial, out, nullptr, stream);
        return;
    }
#endif
    const int nblk = nbpc_cdiv(N, GL_CUBE_CHUNK);
    NBPC_LAUNCH(cube_partial_kernel, nbpc_cdiv((int64_t)B * nblk * C, GL_THREADS), GL_THREADS, 0, stream, X, C, N, nblk, B, partial);
    NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * C, GL_THREADS), GL_THREADS, 0, stream, partial, C, nblk, B, mean ? (float)N : 1.0f, out);
}

extern "C" {

size_t nbpc_set_layer_workspace_bytes(int B, int N, int k, int q) {
    if (B < 1 || N < 1 || k < 1 || q < 1) return 0;
    return set_carve(nullptr, 0, B, N, k, q).bytes;
}

int nbpc_set_layer_fwd(const float *H_in, int B, int N, int k, int q, const float *W, const float *bias, int relu,
                       float *H_out, float *mu, void *workspace, size_t ws_bytes, void *stream_) {
    return nbpc_set_layer_fwd_chained(H_in, B, N, k, q, W, bias, relu, H_out, mu, 0, nullptr, workspace, ws_bytes, stream_);
}

int nbpc_set_layer_fwd_chained(const float *H_in, int B, int N, int k, int q, const float *W, const float *bias, int relu,
                               float *H_out, float *mu, int mu_given, float *mean_out, void *workspace, size_t ws_bytes,
                               void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(H_in && W && bias && H_out && mu && workspace, "null pointer");
    NBPC_ARG(B >= 1 && N >= 1 && k >= 1 && q >= 1, "bad sizes");
    SetWorkspace w = set_carve(workspace, ws_bytes, B, N, k, q);
    if (w.bytes > ws_bytes) {
        nbpc_set_error("nbpc_set_layer_fwd: workspace too small");
        return NBPC_EWORKSPACE;
    }
    const int64_t rows = (int64_t)B * N;
    const int nblk = nbpc_cdiv(N, GL_CUBE_CHUNK);
#ifndef NBPC_HOST_EMU
    if (g_nbpc_math_mode != NBPC_MATH_FP32 && sgt_gemm_shape_ok(k, q)) {
        // tensor-core path (set_layer_tc.cu): column means (unless the producer of H_in handed them over), then
        // (H - mu) W + B on tcgen05; the epilogue leaves the column sums of H_out for the next layer when asked
        if (!mu_given) sgt_colsum(H_in, k, N, B, 1.0f / (float)N, w.partial, mu, nullptr, stream);
        float *parts = (mean_out && sgt_gemm_csum_ok(N)) ? w.csum_parts : nullptr;
        int nsets = 0;
        if (sgt_gemm(H_in, W, 0, mu, bias, nullptr, rows, N, k, q, relu, g_nbpc_math_mode == NBPC_MATH_TF32X3, H_out, parts, &nsets, stream)) {
            nbpc_set_error("nbpc_set_layer_fwd: could not set up the tensor-core kernel (tensor map / shared memory)");
            return NBPC_ELAUNCH;
        }
        if (parts) sgt_colsum_from_parts(parts, nsets, B, q, 1.0f / (float)N, mean_out, nullptr, stream);
        else if (mean_out) set_colstat(H_out, q, N, B, 1, w.partial, mean_out, stream);
        return nbpc_check_launch("nbpc_set_layer_fwd");
    }
#endif
    if (!mu_given) {
        NBPC_LAUNCH(cube_partial_kernel, nbpc_cdiv((int64_t)B * nblk * k, GL_THREADS), GL_THREADS, 0, stream, H_in, k, N, nblk,
                    B, w.partial);
        NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * k, GL_THREADS), GL_THREADS, 0, stream, w.partial, k, nblk, B, (float)N,
                    mu);  // nn.py:25
    }
#ifndef NBPC_HOST_EMU
    if (gl_set_fast() && k <= 16 && q % 4 == 0 && q <= 1024 && sgs_narrow_ok(k)) {   // narrow input layer: float4 stream over q
        const int rpb = sgs_rows_per_block(N, B, set_num_sms()), nb = nbpc_cdiv(N, rpb);
#define X(V) if (k == V) NBPC_LAUNCH_N(NbpcKName("sgs_fwd_smallk", k, q).c_str(), sgs_fwd_smallk_kernel<V>, dim3(nb, B), SGS_THREADS, 0, stream, H_in, mu, W, bias, N, q, rpb, relu, H_out);
        SGS_NARROW_LIST(X)
#undef X
        if (mean_out) set_colstat(H_out, q, N, B, 1, w.partial, mean_out, stream);
        return nbpc_check_launch("nbpc_set_layer_fwd");
    }
#endif
    SetXCentered X;
    X.h = H_in; X.mu = mu; X.k = k; X.N = N;
    NBPC_LAUNCH(set_fwd_kernel, nbpc_cdiv(rows * q, GL_THREADS), GL_THREADS, 0, stream, X, W, bias, rows, k, q, relu,
                H_out);  // nn.py:26-27
    if (mean_out) set_colstat(H_out, q, N, B, 1, w.partial, mean_out, stream);
    return nbpc_check_launch("nbpc_set_layer_fwd");
}

int nbpc_set_layer_bwd(const float *dOut, const float *H_in, const float *H_out, const float *mu, int B, int N,
                       int k, int q, const float *W, int relu, int mask_input, float *dH_in, float *dW, float *dB, void *workspace,
                       size_t ws_bytes, void *stream_) {
    return nbpc_set_layer_bwd_chained(dOut, H_in, H_out, mu, B, N, k, q, W, relu, mask_input, dH_in, dW, dB, nullptr, nullptr,
                                      workspace, ws_bytes, stream_);
}

int nbpc_set_layer_bwd_chained(const float *dOut, const float *H_in, const float *H_out, const float *mu, int B, int N,
                               int k, int q, const float *W, int relu, int mask_input, float *dH_in, float *dW, float *dB,
                               const float *dz_sums, float *dh_sums, void *workspace, size_t ws_bytes, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    NBPC_ARG(!dh_sums || dH_in, "dh_sums needs dH_in");
    if (relu) dz_sums = nullptr;   // sums of the UNMASKED gradient are of no use
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(dOut && H_in && mu && W && dW && dB && workspace, "null pointer");
    NBPC_ARG(!relu || H_out, "H_out is required when relu is set");
    NBPC_ARG(B >= 1 && N >= 1 && k >= 1 && q >= 1, "bad sizes");
    SetWorkspace w = set_carve(workspace, ws_bytes, B, N, k, q);
    if (w.bytes > ws_bytes) {
        nbpc_set_error("nbpc_set_layer_bwd: workspace too small");
        return NBPC_EWORKSPACE;
    }
    const int64_t rows = (int64_t)B * N;
    const int nblk = nbpc_cdiv(N, GL_CUBE_CHUNK);
    SetDz dz;
    dz.g = dOut; dz.hout = H_out; dz.q = q; dz.relu = relu;
#ifndef NBPC_HOST_EMU
    if (gl_set_fast() && q <= 1024 && !(relu && q % 4 != 0)) {
        const bool tc = g_nbpc_math_mode != NBPC_MATH_FP32;
        const int x3 = g_nbpc_math_mode == NBPC_MATH_TF32X3;
        const float *dZ = dOut;
        if (relu) {   // gradient not pre-masked by the consumer: materialise dZ = dOut * [H_out > 0]
            NBPC_LAUNCH(set_mask_kernel, nbpc_cdiv(rows * q / 4, 256), 256, 0, stream, dOut, H_out, rows * q / 4, w.dz);
            dZ = w.dz;
        }
        dz.g = dZ; dz.relu = 0;
        // per-sample column means of dZ (the adjoint of the mean subtraction) -> w.colsum, and dB = sum of dZ
        if (dz_sums) {   // handed over by the producer of dOut
            NBPC_LAUNCH(set_from_sums_kernel, nbpc_cdiv(B * q, 256), 256, 0, stream, dz_sums, B, q, 1.0f / (float)N, w.colsum, dB);
        } else if (q % 4 == 0) {
            sgt_colsum(dZ, q, N, B, 1.0f / (float)N, w.partial, w.colsum, dB, stream);
        } else {
            NBPC_LAUNCH(set_dz_partial_kernel, nbpc_cdiv((int64_t)B * nblk * q, GL_THREADS), GL_THREADS, 0, stream, dz, q, N, nblk, B, w.partial);
            NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * q, GL_THREADS), GL_THREADS, 0, stream, w.partial, q, nblk, B, 1.0f, w.colsum);
            NBPC_LAUNCH(set_bias_kernel, nbpc_cdiv(q, 64), 64, 0, stream, w.colsum, B, q, dB);
            NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * q, GL_THREADS), GL_THREADS, 0, stream, w.partial, q, nblk, B, (float)N, w.colsum);
        }
        int rc = 0;
        if (tc && set_pair_ok(N, k, q)) {   // 16-wide output: both products on the tensor pipe through the row-pair views
            float *W2 = w.pair, *dW2 = W2 + 2 * k * 32, *cm2 = dW2 + 2 * k * 32, *mu2 = cm2 + B * 32;
            const int nprep = nbpc_max(2 * k * 32, B * 2 * k);
            NBPC_LAUNCH(set_pair_prepare_kernel, nbpc_cdiv(nprep, 256), 256, 0, stream, W, w.colsum, mu, B, k, W2, cm2, mu2);
            rc = sgt_dw(H_in, dZ, mu2, rows / 2, N / 2, 2 * k, 32, x3, w.xty_partial, dW2, stream);
            if (!rc) NBPC_LAUNCH(set_pair_fold_kernel, nbpc_cdiv(k * 16, 256), 256, 0, stream, dW2, k, dW);
            if (!rc && dH_in)
                rc = sgt_gemm(dZ, W2, 1, cm2, nullptr, mask_input ? H_in : nullptr, rows / 2, N / 2, 32, 2 * k, 0, x3, dH_in, nullptr, nullptr, stream);
            if (rc) {
                nbpc_set_error("nbpc_set_layer_bwd: could not set up the tensor-core kernel (tensor map / shared memory)");
                return NBPC_ELAUNCH;
            }
            if (dh_sums) set_colstat(dH_in, k, N, B, 0, w.partial, dh_sums, stream);
            return nbpc_check_launch("nbpc_set_layer_bwd");
        }
        // ---- dW = (H - mu)^T dZ: tensor pipe, else the narrow-side stream, else the generic fixed-order reduction
        if (tc && sgt_dw_shape_ok(k, q)) {
            rc = sgt_dw(H_in, dZ, mu, rows, N, k, q, x3, w.xty_partial, dW, stream);
        } else if (!set_xty_narrow(H_in, mu, dZ, B, N, k, q, w.xty_partial, dW, stream)) {
            SetXCentered X;
            X.h = H_in; X.mu = mu; X.k = k; X.N = N;
            xty("xty_partial_set_dW", X, dz, rows, k, q, w.xty_partial, dW, stream);
        }
        // ---- dH = (dZ - mean_s dZ) W^T [* (H_in > 0)]
        if (!rc && dH_in) {
            bool sums_done = false;
            if (tc && sgt_gemm_shape_ok(q, k)) {
                float *parts = (dh_sums && sgt_gemm_csum_ok(N)) ? w.csum_parts : nullptr;
                int nsets = 0;
                rc = sgt_gemm(dZ, W, 1, w.colsum, nullptr, mask_input ? H_in : nullptr, rows, N, q, k, 0, x3, dH_in, parts, &nsets, stream);
                if (!rc && parts) {
                    sgt_colsum_from_parts(parts, nsets, B, k, 1.0f, nullptr, dh_sums, stream);
                    sums_done = true;
                }
            } else if (q <= 16 && k % 4 == 0 && k <= 1024 && sgs_narrow_ok(q)) {
                const int rpb = sgs_rows_per_block(N, B, set_num_sms()), nb = nbpc_cdiv(N, rpb);
#define X(V) if (q == V) NBPC_LAUNCH_N(NbpcKName("sgs_bwd_in_smallq", k, q).c_str(), sgs_bwd_in_smallq_kernel<V>, dim3(nb, B), SGS_THREADS, 0, stream, dZ, w.colsum, W, mask_input ? H_in : (const float *)nullptr, N, k, rpb, dH_in);
                SGS_NARROW_LIST(X)
#undef X
            } else {
                NBPC_LAUNCH(set_bwd_in_mean_kernel, nbpc_cdiv(rows * k, GL_THREADS), GL_THREADS, 0, stream, dz, w.colsum, W, rows, N, k, q, dH_in);
                if (mask_input) NBPC_LAUNCH(set_mask_input_kernel, nbpc_cdiv(rows * k, GL_THREADS), GL_THREADS, 0, stream, H_in, rows * k, dH_in);
            }
            if (!rc && dh_sums && !sums_done) set_colstat(dH_in, k, N, B, 0, w.partial, dh_sums, stream);
        }
        if (rc) {
            nbpc_set_error("nbpc_set_layer_bwd: could not set up the tensor-core kernel (tensor map / shared memory)");
            return NBPC_ELAUNCH;
        }
        return nbpc_check_launch("nbpc_set_layer_bwd");
    }
#endif
    NBPC_LAUNCH(set_dz_partial_kernel, nbpc_cdiv((int64_t)B * nblk * q, GL_THREADS), GL_THREADS, 0, stream, dz, q, N, nblk,
                B, w.partial);
    NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * q, GL_THREADS), GL_THREADS, 0, stream, w.partial, q, nblk, B, 1.0f,
                w.colsum);
    NBPC_LAUNCH(set_bias_kernel, nbpc_cdiv(q, 64), 64, 0, stream, w.colsum, B, q, dB);
    SetXCentered X;
    X.h = H_in; X.mu = mu; X.k = k; X.N = N;
    xty("xty_partial_set_dW", X, dz, rows, k, q, w.xty_partial, dW, stream);
    if (dH_in) {
        NBPC_LAUNCH(set_bwd_in_kernel, nbpc_cdiv(rows * k, GL_THREADS), GL_THREADS, 0, stream, dz, w.colsum, W, rows, N, k, q,
                    dH_in);
        if (mask_input) NBPC_LAUNCH(set_mask_input_kernel, nbpc_cdiv(rows * k, GL_THREADS), GL_THREADS, 0, stream, H_in, rows * k, dH_in);
        if (dh_sums) set_colstat(dH_in, k, N, B, 0, w.partial, dh_sums, stream);
    }
    return nbpc_check_launch("nbpc_set_layer_bwd");
}

}  // extern "C"
