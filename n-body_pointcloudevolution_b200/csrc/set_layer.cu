// set_layer.cu - permutation-equivariant set layer forward/backward (nn.py:10-28):
//   out = (H - mean_N H) W + B   on (B,N,k) -> (B,N,q), optional ReLU (nn.py:59, 65-66).
// Backward:  dZ = dOut * [out > 0];  dB = sum dZ;  dW = (H - mu)^T dZ;
//            dH = (dZ - colmean_s(dZ)) W^T   (the mean-subtraction's adjoint).
// Barrier-free baseline kernels (one thread per output element); all reductions fixed-order.
#include "nbpc_common.cuh"
#include "reduce.cuh"

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

struct SetWorkspace {
    float *partial;   // (B, nblk, max(k,q))
    float *colsum;    // (B, q)
    float *xty_partial;
    size_t bytes;
};

static SetWorkspace set_carve(void *ws, size_t ws_bytes, int B, int N, int k, int q) {
    NbpcArena a(ws, ws_bytes);
    SetWorkspace w;
    const int mx = k > q ? k : q;
    const int nblk = nbpc_cdiv(N, GL_CUBE_CHUNK);
    w.partial = a.take<float>((size_t)B * nblk * mx);
    w.colsum = a.take<float>((size_t)B * mx);
    int rpc, nc;
    xty_plan((int64_t)B * N, k, q, &rpc, &nc);
    w.xty_partial = a.take<float>((size_t)nc * k * q);
    w.bytes = a.off;
    return w;
}

extern "C" {

size_t nbpc_set_layer_workspace_bytes(int B, int N, int k, int q) {
    if (B < 1 || N < 1 || k < 1 || q < 1) return 0;
    return set_carve(nullptr, 0, B, N, k, q).bytes;
}

int nbpc_set_layer_fwd(const float *H_in, int B, int N, int k, int q, const float *W, const float *bias, int relu,
                       float *H_out, float *mu, void *workspace, size_t ws_bytes, void *stream_) {
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
    NBPC_LAUNCH(cube_partial_kernel, nbpc_cdiv((int64_t)B * nblk * k, GL_THREADS), GL_THREADS, 0, stream, H_in, k, N, nblk,
                B, w.partial);
    NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * k, GL_THREADS), GL_THREADS, 0, stream, w.partial, k, nblk, B, (float)N,
                mu);  // nn.py:25
    SetXCentered X;
    X.h = H_in; X.mu = mu; X.k = k; X.N = N;
    NBPC_LAUNCH(set_fwd_kernel, nbpc_cdiv(rows * q, GL_THREADS), GL_THREADS, 0, stream, X, W, bias, rows, k, q, relu,
                H_out);  // nn.py:26-27
    return nbpc_check_launch("nbpc_set_layer_fwd");
}

int nbpc_set_layer_bwd(const float *dOut, const float *H_in, const float *H_out, const float *mu, int B, int N,
                       int k, int q, const float *W, int relu, float *dH_in, float *dW, float *dB, void *workspace,
                       size_t ws_bytes, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
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
    NBPC_LAUNCH(set_dz_partial_kernel, nbpc_cdiv((int64_t)B * nblk * q, GL_THREADS), GL_THREADS, 0, stream, dz, q, N, nblk,
                B, w.partial);
    NBPC_LAUNCH(cube_final_kernel, nbpc_cdiv(B * q, GL_THREADS), GL_THREADS, 0, stream, w.partial, q, nblk, B, 1.0f,
                w.colsum);
    NBPC_LAUNCH(set_bias_kernel, nbpc_cdiv(q, 64), 64, 0, stream, w.colsum, B, q, dB);
    SetXCentered X;
    X.h = H_in; X.mu = mu; X.k = k; X.N = N;
    xty("xty_partial_set_dW", X, dz, rows, k, q, w.xty_partial, dW, stream);
    if (dH_in)
        NBPC_LAUNCH(set_bwd_in_kernel, nbpc_cdiv(rows * k, GL_THREADS), GL_THREADS, 0, stream, dz, w.colsum, W, rows, N, k, q,
                    dH_in);
    return nbpc_check_launch("nbpc_set_layer_bwd");
}

}  // extern "C"
