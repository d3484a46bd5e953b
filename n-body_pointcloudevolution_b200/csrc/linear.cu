// linear.cu - plain dense projection Y = X W (+ bias) over row-major FP32 rows and its deterministic weight gradient
// X^T dY: the building blocks of the 15-weight shift-invariant layer (graph.py:20-200), whose 15 projections act on
// pooled / gathered / transposed copies of the edge tensor (tf.matmul at graph.py:147-189).
#include "nbpc_common.cuh"
#include "reduce.cuh"

// thread per (row, output channel): X[row, :] is read as a warp-broadcast, W[:, qo] coalesced over qo
__global__ void linear_rows_kernel(const float *__restrict__ X, const float *__restrict__ W, const float *__restrict__ bias,
                                   int64_t n, int k, int q, int transpose_w, int accumulate, float *__restrict__ Y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * q) return;
    const int64_t row = t / q;
    const int qo = (int)(t % q);
    const float *x = X + row * k;
    float acc = 0.f;
    if (transpose_w)   // W is (q, k): Y = X W^T
        for (int kk = 0; kk < k; ++kk) acc += x[kk] * W[(int64_t)qo * k + kk];
    else               // W is (k, q)
        for (int kk = 0; kk < k; ++kk) acc += x[kk] * W[(int64_t)kk * q + qo];
    if (bias) acc += bias[qo];
    Y[t] = accumulate ? Y[t] + acc : acc;
}

struct LinWorkspace {
    float *partial;
    size_t bytes;
};
static LinWorkspace lin_carve(void *ws, size_t ws_bytes, int64_t n, int k, int q) {
    NbpcArena a(ws, ws_bytes);
    LinWorkspace w;
    int rpc, nc;
    xty_plan(n, k, q, &rpc, &nc);
    w.partial = a.take<float>((size_t)nc * k * q);
    w.bytes = a.off;
    return w;
}

extern "C" {

int nbpc_linear(const float *X, const float *W, const float *bias, int64_t n, int k, int q, int transpose_w, int accumulate,
                float *Y, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(X && W && Y, "null pointer");
    NBPC_ARG(n >= 0 && k >= 1 && q >= 1, "bad sizes");
    if (n == 0) return NBPC_OK;
    NBPC_LAUNCH(linear_rows_kernel, nbpc_cdiv(n * q, GL_THREADS), GL_THREADS, 0, stream, X, W, bias, n, k, q, transpose_w, accumulate, Y);
    return nbpc_check_launch("nbpc_linear");
}

size_t nbpc_xty_workspace_bytes(int64_t n, int k, int q) {
    if (n < 0 || k < 1 || q < 1) return 0;
    return lin_carve(nullptr, 0, n, k, q).bytes;
}

int nbpc_xty(const float *X, const float *Y, int64_t n, int k, int q, float *out, void *workspace, size_t ws_bytes, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(X && Y && out && workspace, "null pointer");
    NBPC_ARG(n >= 1 && k >= 1 && q >= 1, "bad sizes");
    LinWorkspace w = lin_carve(workspace, ws_bytes, n, k, q);
    if (w.bytes > ws_bytes) {
        nbpc_set_error("nbpc_xty: workspace too small");
        return NBPC_EWORKSPACE;
    }
    GlPlain x, y;
    x.p = X; x.ld = k;
    y.p = Y; y.ld = q;
    xty("nbpc_xty", x, y, n, k, q, w.partial, out, stream);
    return nbpc_check_launch("nbpc_xty");
}

}  // extern "C"
