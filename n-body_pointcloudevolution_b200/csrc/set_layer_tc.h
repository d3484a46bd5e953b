// set_layer_tc.h - entry points of the tcgen05 set-layer kernels (set_layer_tc.cu)
#pragma once
#ifndef NBPC_HOST_EMU
#include <cuda_runtime.h>
#include <stdint.h>
// row GEMM out (rows, Nout) = act((A - mu_s) Bm + bias) [* (mask > 0)], A (rows, K); Bm = Bsrc (K, Nout) or, with b_transposed,
// Bsrc^T for Bsrc (Nout, K); mu (samples, K) / bias (Nout) / mask (rows, Nout) may be null.  K % 32 == 0, Nout % 16 == 0, <= 256
bool sgt_gemm_shape_ok(int K, int Nout);
// csum_parts (optional; requires sgt_gemm_csum_ok, sgt_gemm_csum_floats(samples, Nout) floats): per-sample column sums of `out` as
// *n_sets partial sets [set][sample][Nout], to be added up by sgt_colsum_from_parts
bool sgt_gemm_csum_ok(int rows_per_sample);
size_t sgt_gemm_csum_floats(int samples, int Nout);
int sgt_gemm(const float *A, const float *Bsrc, int b_transposed, const float *mu, const float *bias, const float *mask, int64_t rows,
             int rows_per_sample, int K, int Nout, int relu, int x3, float *out, float *csum_parts, int *n_sets, cudaStream_t stream);
// out[s][c] = scale * sum over sets, sums[s][c] = the unscaled sum (either may be null)
void sgt_colsum_from_parts(const float *parts, int nsets, int B, int C, float scale, float *out, float *sums, cudaStream_t stream);
// dW (k, q) = (H - mu_s)^T dZ; k in {64, 128, 256}, q % 32 == 0, q <= 256; partial: sgt_dw_max_parts() * k * q floats
bool sgt_dw_shape_ok(int k, int q);
int sgt_dw_max_parts();
int sgt_dw(const float *H, const float *dZ, const float *mu, int64_t rows, int rows_per_sample, int k, int q, int x3, float *partial,
           float *dW, cudaStream_t stream);
// out[s][c] = scale * sum_n X[s][n][c] (C % 4 == 0, C <= 1024), total[c] = sum_s of the unscaled sums (optional);
// partial: B * (sgt_colsum_blocks(N) + 1) * C floats
int sgt_colsum_blocks(int N);
void sgt_colsum(const float *X, int C, int N, int B, float scale, float *partial, float *out, float *total, cudaStream_t stream);
#endif
