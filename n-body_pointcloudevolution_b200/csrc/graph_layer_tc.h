// graph_layer_tc.h - entry points of the tcgen05 edge-GEMM kernels (graph_layer_tc_fwd.cu / graph_layer_tc_bwd.cu)
#pragma once
#ifndef NBPC_HOST_EMU
#include <cuda_runtime.h>
#include <stdint.h>
// channel widths with a tensor-core instance: k, q in {16, 32, 64}, where the stage ring fits shared memory
bool glt_fwd_shape_ok(int k, int q, int x3);
// out[e] = act(H[e] W1 + Qc[col[e]] + Qr[e / M]); returns 0 on success
int glt_edge_out(int k, int q, const float *H, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c, int M,
                 int relu, int x3, float *out, cudaStream_t stream);
// the same with a VIRTUAL input (graph_layer_vin.cuh; k0 = 3 input channels of the producing layer, split mode)
struct GlVin;
bool glt_fwd_vin_shape_ok(int k0, int k, int q);
int glt_edge_out_vin(int k0, int k, int q, const GlVin *vin, const int32_t *col, const float *W1, const float *Qc, const float *Qr, int64_t c,
                     int M, int relu, float *out, cudaStream_t stream);
// dH[e] = (dZ[e] W1^T + Gc[col[e]] + Gr[e / M]) [* (H[e] > 0)], per-block partials of dW1 = H^T dZ -> `partial`;
// returns the number of partial blocks written (<= 0 on error)
int glt_edge_bwd(int k, int q, const float *dZ, const float *H, const int32_t *col, const float *W1, const float *Gc, const float *Gr,
                 int64_t c, int M, int mask_in, int x3, float *dH, float *partial, cudaStream_t stream);
bool glt_bwd_vin_shape_ok(int k0, int k, int q, int64_t c);
int glt_edge_bwd_vin(int k0, int k, int q, const float *dZ, const GlVin *vin, const int32_t *col, const float *W1, const float *Gc,
                     const float *Gr, int64_t c, int M, float *dH, float *partial, cudaStream_t stream);
// whether glt_edge_bwd has an instance for (k, q) in this mode (shared-memory budget) and c edges (packing)
bool glt_bwd_shape_ok(int k, int q, int x3, int64_t c);
// upper bound on the partial blocks glt_edge_bwd writes
int glt_max_partial_blocks();
#endif
