// loss.cu - readout / losses (nn.py:107-166) and the TF-form Adam step (train.py:70).
// Reductions are two-level with a fixed order (no atomics) => bit-reproducible losses.
#include "nbpc_common.cuh"

#define LOSS_THREADS 256
#define LOSS_CHUNK 16    // rows per first-level partial
#define LOSS_FAN 64      // first-level partials per second-level partial

__device__ __forceinline__ float sq_rn(float a) { return __fmul_rn(a, a); }

// nn.py:130-133, per axis
__device__ __forceinline__ float pbd_axis(float r, float t, int *which) {
    const float e1 = __fadd_rn(r, -t);
    const float e2 = __fadd_rn(r, -__fadd_rn(1.f, t));
    const float e3 = __fadd_rn(__fadd_rn(1.f, r), -t);
    const float d1 = sq_rn(e1), d2 = sq_rn(e2), d3 = sq_rn(e3);
    float best = d1;
    int w = 0;
    if (d2 < best) { best = d2; w = 1; }
    if (d3 < best) { best = d3; w = 2; }
    if (which) *which = w;
    return best;
}

__device__ __forceinline__ float pbd_diff(float r, float t, int which) {
    if (which == 0) return __fadd_rn(r, -t);
    if (which == 1) return __fadd_rn(r, -__fadd_rn(1.f, t));
    return __fadd_rn(__fadd_rn(1.f, r), -t);
}

template <int PBC>
__global__ void loss_partial_kernel(const float *__restrict__ pred, int ldp, const float *__restrict__ truth, int ldt,
                                    int64_t rows, int nchunks, float *__restrict__ partial) {
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= nchunks) return;
    const int64_t r0 = (int64_t)ch * LOSS_CHUNK, r1 = nbpc_min(r0 + LOSS_CHUNK, rows);
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        float rs = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float p = pred[r * ldp + d], t = truth[r * ldt + d];
            rs = __fadd_rn(rs, PBC ? pbd_axis(p, t, nullptr) : sq_rn(__fadd_rn(p, -t)));
        }
        acc = __fadd_rn(acc, rs);
    }
    partial[ch] = acc;
}

__global__ void loss_mid_kernel(const float *__restrict__ partial, int n1, int n2, float *__restrict__ partial2) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    const int i0 = j * LOSS_FAN, i1 = nbpc_min(i0 + LOSS_FAN, n1);
    float acc = 0.f;
    for (int i = i0; i < i1; ++i) acc = __fadd_rn(acc, partial[i]);
    partial2[j] = acc;
}

__global__ void loss_final_kernel(const float *__restrict__ partial, int nchunks, float rows, float scale,
                                  float *__restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float acc = 0.f;
    for (int i = 0; i < nchunks; ++i) acc = __fadd_rn(acc, partial[i]);
    out[0] = __fmul_rn(acc / rows, scale);
}

template <int PBC>
__global__ void loss_bwd_kernel(const float *__restrict__ pred, int ldp, const float *__restrict__ truth, int ldt,
                                int64_t rows, float scale, const float *__restrict__ dloss, float *__restrict__ dpred,
                                int ldd) {
    int64_t t_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= rows * 3) return;
    const int64_t r = t_ / 3;
    const int d = (int)(t_ % 3);
    const float p = pred[r * ldp + d], t = truth[r * ldt + d];
    float diff;
    if (PBC) {
        int which;
        (void)pbd_axis(p, t, &which);
        diff = pbd_diff(p, t, which);
    } else {
        diff = __fadd_rn(p, -t);
    }
    dpred[r * ldd + d] = dloss[0] * scale * 2.f * diff / (float)rows;
}

__global__ void pbd_kernel(const float *__restrict__ pred, int ldp, const float *__restrict__ truth, int ldt,
                           int64_t rows, float *__restrict__ out) {
    int64_t t_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= rows * 3) return;
    const int64_t r = t_ / 3;
    const int d = (int)(t_ % 3);
    out[t_] = pbd_axis(pred[r * ldp + d], truth[r * ldt + d], nullptr);
}

__device__ __forceinline__ float sign_f(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// nn.py:107-119, literal op order (keeps the reference's value 0.5 at exactly 0 and 1)
__global__ void readout_kernel(const float *__restrict__ h, int64_t rows, int C, float *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C) return;
    const int ch = (int)(t % C);
    const float c = h[t];
    if (ch >= 3) {
        out[t] = c;
        return;
    }
    const float gt_one = __fmul_rn(__fadd_rn(sign_f(__fadd_rn(c, -1.f)), 1.f), 0.5f);
    const float ls_zero = __fmul_rn(-__fadd_rn(sign_f(c), -1.f), 0.5f);
    const float rest = __fadd_rn(__fadd_rn(1.f, -gt_one), -ls_zero);
    const float a = __fmul_rn(rest, c);
    const float b = __fmul_rn(gt_one, __fadd_rn(c, -1.f));
    const float d = __fmul_rn(ls_zero, __fadd_rn(1.f, c));
    out[t] = __fadd_rn(__fadd_rn(a, b), d);
}

__global__ void adam_tf_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                               float *__restrict__ v, int64_t n, float lr_t, float b1, float b2, float eps,
                               float gscale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

// same update with the step counter in device memory (incremented by thread 0 AFTER every thread has read it is not
// possible in one launch, so the counter is advanced by a 1-thread kernel first): the launch parameters are then
// identical every step, which is what a CUDA-graph replay needs
__global__ void adam_tick_kernel(int64_t *step) { *step += 1; }
__global__ void adam_tf_dev_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                   float *__restrict__ v, int64_t n, const int64_t *__restrict__ step, float lr, float b1,
                                   float b2, float eps, float gscale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // a counter that is still <= 0 after the tick makes the call a no-op: a pipelined loop that issues the update of step
    // t at the START of step t+1 (train_utils.PipelinedStep) starts the counter at -1 so that its first replay updates nothing
    if (*step <= 0) return;
    const double t = (double)*step;
    const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

#ifndef NBPC_HOST_EMU
// The three reduction levels above in ONE launch, same summation order (bit-identical result): a block computes its 256 chunk
// partials and the 4 second-level sums they form; the block that takes the last ticket adds the second-level sums in index
// order.  The ticket is an integer atomic (no float atomics anywhere); it is zeroed by the launcher and left at zero.
template <int PBC>
__global__ void __launch_bounds__(LOSS_THREADS) loss_fused_kernel(const float *__restrict__ pred, int ldp, const float *__restrict__ truth,
                                                                  int ldt, int64_t rows, int nchunks, int n2, float scale,
                                                                  float *__restrict__ partial2, unsigned int *__restrict__ ticket,
                                                                  float *__restrict__ out) {
    static_assert(LOSS_THREADS % LOSS_FAN == 0, "a block's chunks form whole second-level groups");
    __shared__ float sp[LOSS_THREADS];
    __shared__ bool last;
    const int ch = blockIdx.x * LOSS_THREADS + threadIdx.x;
    float acc = 0.f;
    if (ch < nchunks) {
        const int64_t r0 = (int64_t)ch * LOSS_CHUNK, r1 = nbpc_min(r0 + LOSS_CHUNK, rows);
        for (int64_t r = r0; r < r1; ++r) {
            float rs = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float p = pred[r * ldp + d], t = truth[r * ldt + d];
                rs = __fadd_rn(rs, PBC ? pbd_axis(p, t, nullptr) : sq_rn(__fadd_rn(p, -t)));
            }
            acc = __fadd_rn(acc, rs);
        }
    }
    sp[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < LOSS_THREADS / LOSS_FAN) {
        const int j = blockIdx.x * (LOSS_THREADS / LOSS_FAN) + threadIdx.x;
        if (j < n2) {
            const int i0 = (int)threadIdx.x * LOSS_FAN, i1 = nbpc_min(i0 + LOSS_FAN, nchunks - (int)blockIdx.x * LOSS_THREADS);
            float a2 = 0.f;
            for (int i = i0; i < i1; ++i) a2 = __fadd_rn(a2, sp[i]);
            partial2[j] = a2;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    float tot = 0.f;
    for (int j0 = 0; j0 < n2; j0 += LOSS_THREADS) {         // staged through shared memory, added in index order by one thread
        const int j = j0 + threadIdx.x;
        __syncthreads();
        sp[threadIdx.x] = j < n2 ? __ldcg(&partial2[j]) : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = nbpc_min(LOSS_THREADS, n2 - j0);
            for (int i = 0; i < n; ++i) tot = __fadd_rn(tot, sp[i]);
        }
    }
    if (threadIdx.x == 0) {
        out[0] = __fmul_rn(tot / (float)rows, scale);
        *ticket = 0u;
    }
}
#endif

template <int PBC>
static int loss_fwd_impl(const float *pred, int ldp, const float *truth, int ldt, int64_t rows, float scale,
                         float *loss_out, void *workspace, size_t ws_bytes, cudaStream_t stream, const char *name) {
    NBPC_TRY(nbpc_require_sm100());
    NBPC_ARG(pred && truth && loss_out && workspace, "null pointer");
    NBPC_ARG(rows >= 1 && ldp >= 3 && ldt >= 3, "bad sizes");
    const int nchunks = nbpc_cdiv(rows, LOSS_CHUNK);
    const int n2 = nbpc_cdiv(nchunks, LOSS_FAN);
    if (nbpc_align_up((size_t)(nchunks + n2) * sizeof(float)) > ws_bytes) {
        nbpc_set_error(std::string(name) + ": workspace too small");
        return NBPC_EWORKSPACE;
    }
    float *partial = (float *)workspace;
    float *partial2 = partial + nchunks;
#ifndef NBPC_HOST_EMU
    {   // one launch; `partial` (unused by it) lends its first word to the ticket
        unsigned int *ticket = (unsigned int *)partial;
        if (nbpc_memset_async(ticket, 0, sizeof(unsigned int), stream)) {
            nbpc_set_error(std::string(name) + ": memset failed");
            return NBPC_ELAUNCH;
        }
        void (*fused)(const float *, int, const float *, int, int64_t, int, int, float, float *, unsigned int *, float *) = loss_fused_kernel<PBC>;
        NBPC_LAUNCH_N("loss_fused_kernel", fused, nbpc_cdiv(nchunks, LOSS_THREADS), LOSS_THREADS, 0, stream, pred, ldp, truth, ldt, rows,
                      nchunks, n2, scale, partial2, ticket, loss_out);
        return nbpc_check_launch(name);
    }
#endif
    void (*kern)(const float *, int, const float *, int, int64_t, int, float *) = loss_partial_kernel<PBC>;
    NBPC_LAUNCH_N("loss_partial_kernel", kern, nbpc_cdiv(nchunks, LOSS_THREADS), LOSS_THREADS, 0, stream, pred, ldp, truth, ldt, rows, nchunks, partial);
    NBPC_LAUNCH(loss_mid_kernel, nbpc_cdiv(n2, LOSS_THREADS), LOSS_THREADS, 0, stream, partial, nchunks, n2, partial2);
    NBPC_LAUNCH(loss_final_kernel, 1, 32, 0, stream, partial2, n2, (float)rows, scale, loss_out);
    return nbpc_check_launch(name);
}

template <int PBC>
static int loss_bwd_impl(const float *pred, int ldp, const float *truth, int ldt, int64_t rows, float scale,
                         const float *dloss, float *dpred, int ldd, cudaStream_t stream, const char *name) {
    NBPC_TRY(nbpc_require_sm100());
    NBPC_ARG(pred && truth && dloss && dpred, "null pointer");
    NBPC_ARG(rows >= 1 && ldp >= 3 && ldt >= 3 && ldd >= 3, "bad sizes");
    void (*kern)(const float *, int, const float *, int, int64_t, float, const float *, float *, int) = loss_bwd_kernel<PBC>;
    NBPC_LAUNCH_N("loss_bwd_kernel", kern, nbpc_cdiv(rows * 3, LOSS_THREADS), LOSS_THREADS, 0, stream, pred, ldp, truth, ldt, rows, scale, dloss,
                dpred, ldd);
    return nbpc_check_launch(name);
}

extern "C" {

size_t nbpc_loss_workspace_bytes(int64_t rows) {
    if (rows < 1) return 0;
    const size_t n1 = (size_t)nbpc_cdiv(rows, LOSS_CHUNK);
    return nbpc_align_up((n1 + (n1 + LOSS_FAN - 1) / LOSS_FAN) * sizeof(float));
}

int nbpc_loss_za_fwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows, float *loss_out,
                     void *workspace, size_t ws_bytes, void *stream) {
    return loss_fwd_impl<0>(pred, ld_pred, truth, ld_truth, rows, 1.f, loss_out, workspace, ws_bytes,
                            (cudaStream_t)stream, "nbpc_loss_za_fwd");
}

int nbpc_loss_za_bwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                     const float *dloss, float *dpred, int ld_dpred, void *stream) {
    return loss_bwd_impl<0>(pred, ld_pred, truth, ld_truth, rows, 1.f, dloss, dpred, ld_dpred, (cudaStream_t)stream,
                            "nbpc_loss_za_bwd");
}

int nbpc_pbc_loss_fwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                      int scale_error, float *loss_out, void *workspace, size_t ws_bytes, void *stream) {
    return loss_fwd_impl<1>(pred, ld_pred, truth, ld_truth, rows, scale_error ? 1e5f : 1.f, loss_out, workspace,
                            ws_bytes, (cudaStream_t)stream, "nbpc_pbc_loss_fwd");
}

int nbpc_pbc_loss_bwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                      int scale_error, const float *dloss, float *dpred, int ld_dpred, void *stream) {
    return loss_bwd_impl<1>(pred, ld_pred, truth, ld_truth, rows, scale_error ? 1e5f : 1.f, dloss, dpred, ld_dpred,
                            (cudaStream_t)stream, "nbpc_pbc_loss_bwd");
}

int nbpc_periodic_boundary_dist(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                                float *dist_out, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(pred && truth && dist_out, "null pointer");
    NBPC_ARG(rows >= 1 && ld_pred >= 3 && ld_truth >= 3, "bad sizes");
    NBPC_LAUNCH(pbd_kernel, nbpc_cdiv(rows * 3, LOSS_THREADS), LOSS_THREADS, 0, stream, pred, ld_pred, truth, ld_truth, rows,
                dist_out);
    return nbpc_check_launch("nbpc_periodic_boundary_dist");
}

// scaled residual update of the multi-redshift model (graph.py:558-566): X (rows, >= 6) = [loc, vel], net (rows, C), C = 3 | 6
//   out[:, :3] = net[:, :3] * loc_scalar + loc + vel * vel_scalar;   out[:, 3:6] = net[:, 3:6] * vel_scalar + vel   (C = 6)
__global__ void residual_update_kernel(const float *__restrict__ X, int ldx, const float *__restrict__ net, int C, int64_t rows,
                                       float loc_scalar, float vel_scalar, float *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C) return;
    const int64_t r = t / C;
    const int ch = (int)(t % C);
    const float n = net[t];
    if (ch < 3) out[t] = n * loc_scalar + X[r * ldx + ch] + X[r * ldx + 3 + ch] * vel_scalar;
    else out[t] = n * vel_scalar + X[r * ldx + ch];
}

int nbpc_residual_update(const float *X, int ldx, const float *net, int C, int64_t rows, float loc_scalar, float vel_scalar, float *out,
                         void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(X && net && out, "null pointer");
    NBPC_ARG(rows >= 1 && (C == 3 || C == 6) && ldx >= 6, "bad sizes");
    NBPC_LAUNCH(residual_update_kernel, nbpc_cdiv(rows * C, LOSS_THREADS), LOSS_THREADS, 0, stream, X, ldx, net, C, rows, loc_scalar,
                vel_scalar, out);
    return nbpc_check_launch("nbpc_residual_update");
}

int nbpc_readout(const float *h, int64_t rows, int C, float *out, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(h && out, "null pointer");
    NBPC_ARG(rows >= 1 && C >= 3, "bad sizes");
    NBPC_LAUNCH(readout_kernel, nbpc_cdiv(rows * C, LOSS_THREADS), LOSS_THREADS, 0, stream, h, rows, C, out);
    return nbpc_check_launch("nbpc_readout");
}

int nbpc_adam_tf(float *param, const float *grad, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                 float eps, int64_t step, float grad_scale, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(param && grad && m && v, "null pointer");
    NBPC_ARG(n >= 1 && step >= 1, "bad sizes");
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    NBPC_LAUNCH(adam_tf_kernel, nbpc_cdiv(n, LOSS_THREADS), LOSS_THREADS, 0, stream, param, grad, m, v, n, (float)lr_t,
                beta1, beta2, eps, grad_scale);
    return nbpc_check_launch("nbpc_adam_tf");
}

int nbpc_adam_tf_dev(float *param, const float *grad, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                     float eps, int64_t *step_counter, float grad_scale, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(param && grad && m && v && step_counter, "null pointer");
    NBPC_ARG(n >= 1, "bad sizes");
    NBPC_LAUNCH(adam_tick_kernel, 1, 1, 0, stream, step_counter);
    NBPC_LAUNCH(adam_tf_dev_kernel, nbpc_cdiv(n, LOSS_THREADS), LOSS_THREADS, 0, stream, param, grad, m, v, n, step_counter, lr,
                beta1, beta2, eps, grad_scale);
    return nbpc_check_launch("nbpc_adam_tf_dev");
}

}  // extern "C"
