// graph_layer_glf.cu - instantiations and launchers of the CUDA-core tiled edge kernels (graph_layer_fast.cuh): the
// FP32 math mode of the graph layer and the deterministic X^T Y micro-tile reductions.  Own translation unit so that
// the ~120 kernel instances here compile next to graph_layer.cu instead of inside it.
#include <stdlib.h>

#include "nbpc_common.cuh"
#include "reduce.cuh"
#include "graph_layer_fast.cuh"
#include "graph_layer_glf.h"

#ifndef NBPC_HOST_EMU
static int gl_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}
#else
static int gl_num_sms() { return 148; }
#endif
// upper bound on the number of per-block partials any X^T Y reduction writes (persistent grids never
// exceed 8 resident blocks per SM)
static int gl_max_partial_blocks() { return gl_num_sms() * 8; }
static int gl_node_xty_rows_per_block(int64_t n) {
    int64_t rpb = (n + 591) / 592;
    rpb = (rpb + 31) / 32 * 32;
    return (int)(rpb < 32 ? 32 : rpb);
}
static int gl_node_xty_grid(int64_t n) {
    const int rpb = gl_node_xty_rows_per_block(n);
    return (int)((n + rpb - 1) / rpb);
}


#ifndef NBPC_HOST_EMU
// ---- template dispatch over the compiled (K, Q) shapes of the edge-level kernels
#define GLF_FOR_KQ(X)  X(3, 16) X(3, 32) X(3, 64) X(16, 16) X(16, 32) X(16, 64) X(32, 16) X(32, 32) X(32, 64) X(64, 16) X(64, 32) X(64, 64)

bool glf_edge_shape_ok(int k, int q) {
#define X(K_, Q_) if (k == K_ && q == Q_) return true;
    GLF_FOR_KQ(X)
#undef X
    return false;
}

// persistent grid: blocks/SM from the occupancy calculator (cached per kernel instance)
template <class F>
static int glf_persistent_grid(F kern, int threads, size_t smem) {
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return -1;
    }
    if (occ > 8) occ = 8;
    return gl_num_sms() * occ;
}

template <int K, int Q, bool RELU>
static int glf_launch_edge_out_t(const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr,
                                 int64_t c, int M, float *out, cudaStream_t stream) {
    constexpr int KS = (K == 3) ? 0 : glf_stride(K), QS = glf_stride(Q);
    const size_t smem = sizeof(float) * (size_t)(((K * Q + 3) / 4) * 4 + GLF_TE * QS + 2 * GLF_TE * KS);
    auto kern = glf_edge_out_kernel<K, Q, RELU>;
    static int grid_cache_d[NBPC_MAX_DEVICES];   // per device
    int &grid_cache = grid_cache_d[nbpc_device_slot()];
    const int64_t ntiles = (c + GLF_TE - 1) / GLF_TE;
    if (!grid_cache) grid_cache = glf_persistent_grid(kern, GLF_THREADS, smem);
    if (grid_cache < 0) return 1;
    const int grid = (int)(ntiles < grid_cache ? ntiles : grid_cache);
    NBPC_LAUNCH_N(NbpcKName("glf_edge_out_kernel", K, Q).c_str(), kern, grid, GLF_THREADS, smem, stream, H, col, W1, Qc, Qr, c, M, out);
    return 0;
}
template <int K, int Q>
static int glf_launch_edge_out(const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr,
                               int64_t c, int M, int relu, float *out, cudaStream_t stream) {
    return relu ? glf_launch_edge_out_t<K, Q, true>(H, col, W1, Qc, Qr, c, M, out, stream)
                : glf_launch_edge_out_t<K, Q, false>(H, col, W1, Qc, Qr, c, M, out, stream);
}

// launches the edge backward kernel; returns the number of per-block partials written (<= 0 on error)
template <int K, int Q, bool RELU, bool HAS_DH, bool MASK_IN>
static int glf_launch_edge_bwd_t(const char *name, const float *dOut, const float *Hout, const float *H, const int32_t *col,
                                 const float *W1, const float *Gc, const float *Gr, int64_t c, int M, float *dH,
                                 float *partial, cudaStream_t stream) {
    constexpr int KP = (K == 3) ? 4 : K;
    constexpr int KS = glf_stride(KP), QS = glf_stride(Q);
    constexpr int TILE = GLF_TE * (KS + QS + (RELU ? QS : 0));
    const size_t smem = sizeof(float) * (size_t)(Q * KP + 2 * TILE);
    auto kern = glf_edge_bwd_kernel<K, Q, RELU, HAS_DH, MASK_IN>;
    static int grid_cache_d[NBPC_MAX_DEVICES];   // per device
    int &grid_cache = grid_cache_d[nbpc_device_slot()];
    if (!grid_cache) grid_cache = glf_persistent_grid(kern, GLF_THREADS, smem);
    if (grid_cache < 0) return -1;
    const int64_t ntiles = (c + GLF_TE - 1) / GLF_TE;
    int grid = (int)(ntiles < grid_cache ? ntiles : grid_cache);
    const int tpb = (int)((ntiles + grid - 1) / grid);
    grid = (int)((ntiles + tpb - 1) / tpb);
    NBPC_LAUNCH_N(name, kern, grid, GLF_THREADS, smem, stream, dOut, Hout, H, col, W1, Gc, Gr, c, M, tpb, dH, partial);
    return grid;
}

void glf_reduce_partials(const float *partial, int nblocks, int rows, int cols, int transpose, float *out, cudaStream_t stream) {
    NBPC_LAUNCH(glf_partial_reduce_kernel, nbpc_cdiv(rows * cols, 32), 1024, 0, stream, partial, nblocks, rows, cols, transpose, out);
}

// dW1 == nullptr: leave the nb per-block partials for the caller to reduce (*nb_out)
template <int K, int Q>
static int glf_launch_edge_bwd(const float *dOut, const float *Hout, const float *H, const int32_t *col, const float *W1,
                               const float *Gc, const float *Gr, int64_t c, int M, int relu, int mask_in, float *dH,
                               float *partial, float *dW1, cudaStream_t stream, int *nb_out = nullptr) {
    int nb = -1;
    const NbpcKName nm("glf_edge_bwd_kernel", K, Q);
    const char *n = nm.c_str();
    if constexpr (K % 4 == 0) {
        if (dH) {
            if (relu && mask_in) nb = glf_launch_edge_bwd_t<K, Q, true, true, true>(n, dOut, Hout, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
            else if (relu) nb = glf_launch_edge_bwd_t<K, Q, true, true, false>(n, dOut, Hout, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
            else if (mask_in) nb = glf_launch_edge_bwd_t<K, Q, false, true, true>(n, dOut, Hout, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
            else nb = glf_launch_edge_bwd_t<K, Q, false, true, false>(n, dOut, Hout, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
        }
    }
    if (!dH) {
        nb = relu ? glf_launch_edge_bwd_t<K, Q, true, false, false>(n, dOut, Hout, H, col, W1, Gc, Gr, c, M, dH, partial, stream)
                  : glf_launch_edge_bwd_t<K, Q, false, false, false>(n, dOut, Hout, H, col, W1, Gc, Gr, c, M, dH, partial, stream);
    }
    if (nb <= 0) return 1;
    if (nb_out) *nb_out = nb;
    if (dW1) glf_reduce_partials(partial, nb, K, Q, 0, dW1, stream);
    return 0;
}

// X^T Y over n node rows -> out (k,q); deterministic.  Uses the micro-tile kernel when (k,q) or (q,k)
// is a compiled shape, else the generic kernel.
// out == nullptr: leave the partials (count -> *nb_out, layout (q,k) instead of (k,q) -> *transposed) for the caller
int glf_node_xty(const char *name, const float *X, const float *Y, int64_t n, int k, int q, float *partial, float *out,
                 cudaStream_t stream, int *nb_out, int *transposed) {
    int nb = 0;
    if (transposed) *transposed = 0;
#define XN(K_, Q_)                                                                                                    \
    if (!nb && k == K_ && q == Q_) {                                                                                 \
        nb = glf_launch_edge_bwd_t<K_, Q_, false, false, false>(name, Y, nullptr, X, nullptr, nullptr, nullptr, nullptr, n, 1, \
                                                                nullptr, partial, stream);                          \
        if (nb > 0 && out) glf_reduce_partials(partial, nb, k, q, 0, out, stream);                                   \
    }                                                                                                                \
    if (!nb && k == Q_ && q == K_ && K_ != Q_) {                                                                     \
        nb = glf_launch_edge_bwd_t<K_, Q_, false, false, false>(name, X, nullptr, Y, nullptr, nullptr, nullptr, nullptr, n, 1, \
                                                                nullptr, partial, stream);                          \
        if (nb > 0 && out) glf_reduce_partials(partial, nb, q, k, 1, out, stream);                                   \
        if (nb > 0 && transposed) *transposed = 1;                                                                   \
    }
    GLF_FOR_KQ(XN)
#undef XN
    if (nb < 0) return 1;
    if (nb > 0) {
        if (nb_out) *nb_out = nb;
        return 0;
    }
    const int rpb = gl_node_xty_rows_per_block(n), grid = gl_node_xty_grid(n);
    const size_t smem = sizeof(float) * (size_t)GLF_XTY_ROWS * (k + q);
    NBPC_LAUNCH_N(name, glf_node_xty_kernel, grid, 256, smem, stream, X, Y, n, rpb, k, q, partial);
    if (out) glf_reduce_partials(partial, grid, k, q, 0, out, stream);
    if (nb_out) *nb_out = grid;
    return 0;
}

// dW2 / dW3 partials of one layer in ONE launch (glf_xty_pair_kernel): Xa^T Ya -> partial_a, Xb^T Yb -> partial_b, same partial
// count and layout as two glf_node_xty calls with out == nullptr; returns 1 when no micro-tile instance fits (the caller then
// launches the two problems separately)
template <int K, int Q>
static int glf_launch_xty_pair_t(const float *Ya, const float *Xa, float *pa, const float *Yb, const float *Xb, float *pb, int64_t n,
                                 cudaStream_t stream) {
    constexpr int KP = (K == 3) ? 4 : K;
    constexpr int KS = glf_stride(KP), QS = glf_stride(Q);
    constexpr int TILE = GLF_TE * (KS + QS);
    const size_t smem = sizeof(float) * (size_t)(Q * KP + 2 * TILE);
    auto kern = glf_xty_pair_kernel<K, Q>;
    static int grid_cache_d[NBPC_MAX_DEVICES];   // per device
    int &grid_cache = grid_cache_d[nbpc_device_slot()];
    // the grid of the single-problem kernel (same shared memory, same threads): the pair must produce the same partials
    if (!grid_cache) {
        grid_cache = glf_persistent_grid(glf_edge_bwd_kernel<K, Q, false, false, false>, GLF_THREADS, smem);
        if (grid_cache > 0 && glf_persistent_grid(kern, GLF_THREADS, smem) < 0) grid_cache = -1;
    }
    if (grid_cache < 0) return -1;
    const int64_t ntiles = (n + GLF_TE - 1) / GLF_TE;
    int grid = (int)(ntiles < grid_cache ? ntiles : grid_cache);
    const int tpb = (int)((ntiles + grid - 1) / grid);
    grid = (int)((ntiles + tpb - 1) / tpb);
    NBPC_LAUNCH_N(NbpcKName("glf_xty_pair_kernel", K, Q).c_str(), kern, dim3(grid, 2), GLF_THREADS, smem, stream, Ya, Xa, pa, Yb, Xb, pb, n, tpb);
    return grid;
}
int glf_node_xty_pair(const float *Xa, const float *Ya, float *partial_a, const float *Xb, const float *Yb, float *partial_b, int64_t n,
                      int k, int q, cudaStream_t stream, int *nb_out, int *transposed) {
    int nb = 0;
    *transposed = 0;
#define XN(K_, Q_)                                                                                                    \
    if (!nb && k == K_ && q == Q_) nb = glf_launch_xty_pair_t<K_, Q_>(Ya, Xa, partial_a, Yb, Xb, partial_b, n, stream);      \
    if (!nb && k == Q_ && q == K_ && K_ != Q_) {                                                                     \
        nb = glf_launch_xty_pair_t<K_, Q_>(Xa, Ya, partial_a, Xb, Yb, partial_b, n, stream);                                \
        if (nb > 0) *transposed = 1;                                                                                 \
    }
    GLF_FOR_KQ(XN)
#undef XN
    if (nb <= 0) return 1;
    *nb_out = nb;
    return 0;
}

int glf_dispatch_edge_out(int k, int q, const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c,
                          int M, int relu, float *out, cudaStream_t stream) {
    int rc = 1;
#define X(K_, Q_) if (k == K_ && q == Q_) rc = glf_launch_edge_out<K_, Q_>(H, col, W1, Qc, Qr, c, M, relu, out, stream);
    GLF_FOR_KQ(X)
#undef X
    return rc;
}

int glf_dispatch_edge_bwd(int k, int q, const float *dOut, const float *Hout, const float *H, const int32_t *col, const float *W1,
                          const float *Gc, const float *Gr, int64_t c, int M, int relu, int mask_in, float *dH, float *partial,
                          float *dW1, cudaStream_t stream, int *nb_out) {
    int rc = 1;
#define X(K_, Q_) if (k == K_ && q == Q_) rc = glf_launch_edge_bwd<K_, Q_>(dOut, Hout, H, col, W1, Gc, Gr, c, M, relu, mask_in, dH, partial, dW1, stream, nb_out);
    GLF_FOR_KQ(X)
#undef X
    return rc;
}
#endif  // !NBPC_HOST_EMU
