// graph_layer_glf.h - entry points of the CUDA-core tiled edge kernels (graph_layer_fast.cuh), which are instantiated in
// their own translation unit (graph_layer_glf.cu) so that the library builds in parallel.
#pragma once
#ifndef NBPC_HOST_EMU
#include <cuda_runtime.h>
#include <stdint.h>
// (k, q) with a compiled tiled instance: k in {3,16,32,64}, q in {16,32,64}
bool glf_edge_shape_ok(int k, int q);
// out[e] = act(H[e] W1 + Qc[col[e]] + Qr[e / M]); 0 on success
int glf_dispatch_edge_out(int k, int q, const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c,
                          int M, int relu, float *out, cudaStream_t stream);
// dH / dW1 partials of the tiled backward edge kernel; dW1 == nullptr leaves the *nb_out per-block partials to the caller
int glf_dispatch_edge_bwd(int k, int q, const float *dOut, const float *Hout, const float *H, const int32_t *col, const float *W1,
                          const float *Gc, const float *Gr, int64_t c, int M, int relu, int mask_in, float *dH, float *partial,
                          float *dW1, cudaStream_t stream, int *nb_out);
void glf_reduce_partials(const float *partial, int nblocks, int rows, int cols, int transpose, float *out, cudaStream_t stream);
// X^T Y over n node rows -> out (k,q), deterministic; out == nullptr: leave the partials (count -> *nb_out, layout (q,k)
// instead of (k,q) -> *transposed) for the caller
int glf_node_xty(const char *name, const float *X, const float *Y, int64_t n, int k, int q, float *partial, float *out,
                 cudaStream_t stream, int *nb_out = nullptr, int *transposed = nullptr);
// the dW2 / dW3 pair of a layer (Xa^T Ya, Xb^T Yb; same shapes) in one launch; partials as from two glf_node_xty(out = nullptr)
// calls.  Returns 1 if no micro-tile instance fits (launch them separately then)
int glf_node_xty_pair(const float *Xa, const float *Ya, float *partial_a, const float *Xb, const float *Yb, float *partial_b, int64_t n,
                      int k, int q, cudaStream_t stream, int *nb_out, int *transposed);
#endif
