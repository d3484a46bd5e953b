/* nbpc.h - C ABI of libnbpc.so: the B200 (sm_100a) replacement for the data-parallel hot path
 * of evdcush/N-Body_PointCloudEvolution (periodic-box kNN graph construction + set/graph layer
 * forward/backward).
 *
 * The reference is 100% Python and has NO FFI layer; its boundary for this path is the set of
 * Python call signatures in /root/reference/graph.py and nn.py.  Each entry point below names
 * the reference function(s) (file:line) whose work it replaces.  The Python package
 * `n-body_pointcloudevolution_b200` re-exports the reference's function names on top of these
 * symbols (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes; no allocation, no host synchronisation and no
 *     stream creation inside the library.  All pointers are DEVICE pointers owned by the caller
 *     (contiguous, row-major, 16-byte aligned unless a stride argument says otherwise);
 *     `stream` is a cudaStream_t passed as void*.
 *   - every call returns NBPC_OK (0) or a negative NBPC_E* code; nbpc_last_error_string() gives
 *     the thread-local message of the last failure.
 *   - workspaces: ask nbpc_*_workspace_bytes() first, pass a buffer at least that large.
 *   - the library refuses to run (NBPC_EARCH) on anything but compute capability 10.0.
 *   - c = B*N*M edges, stored row-major by (sample, particle, neighbour slot): edge e belongs to
 *     row node e / M.  Node ids in `col`/COO are global (sample*N + particle).
 */
#ifndef NBPC_H_
#define NBPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBPC_OK 0
#define NBPC_EINVAL (-1)     /* bad argument */
#define NBPC_EARCH (-2)      /* device is not sm_100 */
#define NBPC_ELAUNCH (-3)    /* CUDA launch / runtime error */
#define NBPC_EWORKSPACE (-4) /* workspace too small */

#define NBPC_ORDER_DISTANCE 0 /* ascending (d2, index): raw sklearn / get_pbc_kneighbors_csr order */
#define NBPC_ORDER_INDEX 1    /* ascending column index: get_kneighbor_list order (scipy astype) */

#define NBPC_KNN_MAX_K 64

/* ---------------------------------------------------------------- library */
int nbpc_version(void);
const char *nbpc_last_error_string(void);
/* NBPC_OK if the current CUDA device is compute capability 10.0, else NBPC_EARCH. */
int nbpc_device_check(void);

/* Arithmetic of the edge-level channel projections of the graph layer (H W1, dZ W1^T, H^T dZ) for channel
 * widths in {16,32,64}.  All modes run on the GPU; they differ in which pipe does the multiply-adds:
 *   NBPC_MATH_FP32    CUDA-core FP32 FMAs;
 *   NBPC_MATH_TF32X3  tcgen05 tensor cores, operands split x = hi + lo into two TF32 values and three MMAs
 *                     (lo*hi + hi*lo + hi*hi) accumulated in FP32 in TMEM: FP32-class accuracy (same test tolerance);
 *   NBPC_MATH_TF32    one tcgen05 pass on 10-bit-mantissa operands: ~1e-3 relative (better than BF16).
 * The process default comes from the environment variable NBPC_MATH (fp32 | tf32x3 | tf32), else
 * NBPC_MATH_DEFAULT.  The mode is process-global (not per stream). */
#define NBPC_MATH_FP32 0
#define NBPC_MATH_TF32X3 1
#define NBPC_MATH_TF32 2
#define NBPC_MATH_DEFAULT NBPC_MATH_TF32X3
int nbpc_set_math_mode(int mode);
int nbpc_get_math_mode(void);

/* Tracing (replaces the reference's single wall-clock timer around the training loop,
 * train.py:84, 122-124).  nbpc_launch_count(): kernels launched by this library so far in the
 * process.  nbpc_prof_enable(1) clears the records and starts bracketing every launch with CUDA
 * events on its stream; nbpc_prof_report() writes "kernel\tlaunches\ttotal_ms\n" lines (it waits
 * for the recorded events) and returns the buffer size needed.  Host-side only, not for hot loops. */
long long nbpc_launch_count(void);
int nbpc_prof_enable(int on);
long long nbpc_prof_report(char *buf, size_t cap);

/* ---------------------------------------------------------------- kNN graph
 * Replaces graph.get_kneighbor_list (graph.py:704-713; periodic=0) and
 * graph.get_pbc_kneighbors_csr + pad_cube_boundaries + get_pcube_csr (graph.py:801-917; periodic=1).
 *
 * xyz: float32 positions, element (b,n,d) at xyz[b*stride_b + n*stride_n + d], d in 0..2.
 * periodic=1: unit box; a particle has an image shifted by -1 (+1) along an axis iff
 *   x >= (float)(1-thr)  (x <= (float)thr), exactly the padded cloud of graph.py:842; images are
 *   formed as (double)x + shift and distances are float64, d2 = ((dx*dx)+(dy*dy))+(dz*dz) with
 *   separately rounded multiplies and adds (scikit-learn's rdist arithmetic).
 * include_self=0 drops the un-shifted query particle itself (sklearn queries k+1 and removes it).
 * Ties (exactly equal d2) are broken by ascending particle index - a documented total order;
 *   sklearn's own tie order is KD-tree traversal order and is not reproducible.
 * idx_out: int32 (B,N,k) neighbour indices LOCAL to the sample (image indices mapped back to the
 *   original particle, graph.py:888-893).  d2_out: optional float64 (B,N,k), always in
 *   distance order (NULL to skip).
 * status: optional int32[1] device (NULL to skip): periodic=1 assumes coordinates in the unit box [0,1]
 *   (graph.py:827-855 builds images for the unit box only); status[0] counts particles with a coordinate
 *   outside it, for which the result is unspecified (the reference would still run its KD-tree on them).
 *   The Python facade exposes it as KnnCSR.check().
 * Requires 1 <= k <= NBPC_KNN_MAX_K and k <= N - (include_self ? 0 : 1). */
size_t nbpc_knn_workspace_bytes(int B, int N, int k, int periodic);
/* Query kernel of nbpc_knn (same result bit for bit, different work distribution):
 *   NBPC_KNN_AUTO    thread-per-query for k <= 16, warp-per-query for 16 < k <= 32 (measured crossover)
 *   NBPC_KNN_THREAD  one thread per query, register-resident sorted list, warp-uniform loops with deferred insertion
 *   NBPC_KNN_WARP    one warp per query (k <= 32): FP32 keys sorted across the lanes, FP64 re-evaluation of the survivors
 * The process default comes from the environment variable NBPC_KNN_V (2 = thread, 3 = warp), else NBPC_KNN_AUTO. */
#define NBPC_KNN_AUTO 0
#define NBPC_KNN_THREAD 2
#define NBPC_KNN_WARP 3
int nbpc_set_knn_kernel(int kernel);
int nbpc_get_knn_kernel(void);
int nbpc_knn(const float *xyz, int64_t stride_b, int64_t stride_n, int B, int N, int k,
             int periodic, double boundary_threshold, int include_self, int order,
             int32_t *idx_out, double *d2_out, int32_t *status, void *workspace, size_t ws_bytes, void *stream);

/* The padded ("cloned") cube itself, for callers of graph.pad_cube_boundaries (graph.py:827-855; nbpc_knn never
 * materialises it).  One sample: xyz (N,>=3) float32, row stride stride_n.
 * nbpc_pad_cube_count: offsets int32 (N+1) = exclusive scan of the per-particle image counts (0/1/3/7,
 *   graph.py:801-825); offsets[N] = total number of images (read it back to size the outputs).
 * nbpc_pad_cube_emit: padded_out float64 (N + n_img, 3) = [particles; images in particle order, each particle's
 *   images in the reference's face / edge / corner pattern order], images = (pattern*bound) + particle in float64
 *   (NumPy's int64 + float32 promotion); idx_map_out int64 (n_img) = source particle of every image. */
size_t nbpc_pad_cube_workspace_bytes(int N);
int nbpc_pad_cube_count(const float *xyz, int64_t stride_n, int N, double boundary_threshold, int32_t *offsets,
                        void *workspace, size_t ws_bytes, void *stream);
int nbpc_pad_cube_emit(const float *xyz, int64_t stride_n, int N, double boundary_threshold, const int32_t *offsets,
                       double *padded_out, int64_t *idx_map_out, void *stream);

/* ---------------------------------------------------------------- adjacency
 * Replaces graph.to_coo_batch_ZA_diag / to_coo_batch / get_indices_from_list_CSR
 * (graph.py:593-697) and adds the CSR transpose used by every layer's col-pool and by backward.
 *
 * idx: int32 (B,N,M) local neighbour indices (output of nbpc_knn).
 * coo_out: int32 (3,c): [row + iN, col + iN, i]  (graph.py:643-652).
 * diag_out: int64 (B*N): flat edge position of the FIRST self edge (row == col) of each row, or
 *   -1 if the row has none (graph.py:655-656 for include_self graphs).
 * csrT_ptr int32 (B*N+1), csrT_edge int32 (c): in-edges of every node, edge ids ascending inside
 *   a node (=> deterministic summation order).
 * status: int32[2] device: [0] rows whose number of self edges != 1, [1] indices out of [0,N). */
size_t nbpc_adjacency_workspace_bytes(int B, int N, int M);
int nbpc_adjacency(const int32_t *idx, int B, int N, int M, int32_t *coo_out, int64_t *diag_out,
                   int32_t *csrT_ptr, int32_t *csrT_edge, int32_t *status, void *workspace,
                   size_t ws_bytes, void *stream);

/* Generic: members of every segment, ascending, for arbitrary int32 ids in [0,num_segs)
 * (what tf.unsorted_segment_mean needs to be deterministic).  status: int32[1] = ids out of range. */
size_t nbpc_segment_csr_workspace_bytes(int64_t n_items, int num_segs);
int nbpc_segment_csr(const int32_t *ids, int64_t n_items, int num_segs, int32_t *seg_ptr,
                     int32_t *seg_members, int32_t *status, void *workspace, size_t ws_bytes,
                     void *stream);

/* ---------------------------------------------------------------- edge input features
 * graph.get_input_features_shift_inv_ZA (graph.py:289-343):
 *   edges[e] = pos[col[e]] - pos[e / M];  edges[diag[n]] += za[n]  for n < n_diag (diag[n] < 0 skipped).
 * pos/za: (BN, 3) with leading dimension ld_pos / ld_za (floats). */
int nbpc_edge_features_za(const float *pos, int ld_pos, const float *za, int ld_za,
                          const int32_t *col, const int64_t *diag, int64_t n_diag, int BN, int M,
                          float *edges_out, void *stream);
/* graph.get_input_features_shift_inv (graph.py:346-364): raw differences of X[:, :3] (za NULL) -
 * same kernel, exported under its own name. */
int nbpc_edge_features(const float *pos, int ld_pos, const int32_t *col, int BN, int M,
                       float *edges_out, void *stream);
/* graph.include_node_features (graph.py:245-275):
 *   out[e] = [edges[e] (E), nodes[e / M] (F), nodes[col[e]] (F), redshift[e] (R = 0 or 1)]. */
int nbpc_include_node_features(const float *edges, int E, const float *nodes, int ld_nodes, int F,
                               const int32_t *col, const float *redshift, int BN, int M, float *out,
                               void *stream);

/* ---------------------------------------------------------------- pooling primitive
 * graph.shift_inv_conv (graph.py:367-391) = tf.unsorted_segment_mean (+ tf.gather_nd).
 * nbpc_segment_reduce: out[s] = sum_{e in seg s} h[e] (/ max(count,1) if mean), members ascending.
 * nbpc_gather_rows:    out[i] = src[ids[i]] (* 1/max(count(ids[i]),1) if seg_ptr given). */
int nbpc_segment_reduce(const float *h, int k, const int32_t *seg_ptr, const int32_t *seg_members,
                        int num_segs, int mean, float *out, void *stream);
int nbpc_gather_rows(const float *src, int k, const int32_t *ids, int64_t n_ids,
                     const int32_t *seg_ptr, float *out, void *stream);

/* ---------------------------------------------------------------- dense projection (15-weight layer, graph.py:20-200)
 * nbpc_linear: Y (n,q) = X (n,k) W [+ bias]; W is (k,q), or (q,k) with transpose_w (Y = X W^T, the input gradient);
 *              accumulate adds into Y.  nbpc_xty: out (k,q) = X^T Y over n rows, fixed summation order (the weight
 *              gradient); workspace from nbpc_xty_workspace_bytes. */
int nbpc_linear(const float *X, const float *W, const float *bias, int64_t n, int k, int q, int transpose_w,
                int accumulate, float *Y, void *stream);
size_t nbpc_xty_workspace_bytes(int64_t n, int k, int q);
int nbpc_xty(const float *X, const float *Y, int64_t n, int k, int q, float *out, void *workspace, size_t ws_bytes,
             void *stream);

/* ---------------------------------------------------------------- shift-invariant graph layer
 * graph.shift_inv_layer (graph.py:394-456):
 *   Z = H W1 + pool_col(H) W2 + pool_row(H) W3 + pool_cube(H) W4 + B        (c, q)
 *   is_last: out = row-mean(Z) (BN, q);  relu: out = max(Z, 0) (the activation the network
 *   functions apply between layers, graph.py:466, 474-475).
 * W: float32 (4, k, q) = [W1, W2, W3, W4]; bias (q).
 * Saved for backward (caller-allocated): P_col (BN,k), P_row (BN,k), P_cube (B,k).
 * Backward returns dH_in (c,k; NULL to skip - first layer), dW (4,k,q), dB (q); all sums run in
 * a fixed order (CSR transpose + two-level trees; no float atomics) => bit-reproducible.
 * relu (bwd): dOut is masked by [H_out > 0] first (pass 0 if the caller already did).
 * mask_input: dH_in is multiplied by [H_in > 0], i.e. the ReLU backward of the layer that produced
 *   H_in is fused here, where H_in is already on chip (its producer then passes relu = 0). */
size_t nbpc_graph_layer_workspace_bytes(int B, int N, int M, int k, int q);
int nbpc_graph_layer_fwd(const float *H_in, const int32_t *col, const int32_t *csrT_ptr,
                         const int32_t *csrT_edge, int B, int N, int M, int k, int q,
                         const float *W, const float *bias, int is_last, int relu, float *H_out,
                         float *P_col, float *P_row, float *P_cube, void *workspace,
                         size_t ws_bytes, void *stream);
int nbpc_graph_layer_bwd(const float *dOut, const float *H_in, const float *H_out,
                         const int32_t *col, const int32_t *csrT_ptr, const int32_t *csrT_edge,
                         int B, int N, int M, int k, int q, const float *W, const float *P_col,
                         const float *P_row, const float *P_cube, int is_last, int relu,
                         int mask_input, float *dH_in, float *dW, float *dB, void *workspace,
                         size_t ws_bytes, void *stream);

/* Row-pool hand-over between consecutive layers of a network (graph.py:463-515 calls shift_inv_layer in a loop; every layer
 * starts by pooling its input over the row, graph.py:428-449, and its backward starts by summing dZ over the row).  The kernel
 * that WRITES an edge tensor can emit those row reductions on the way, so that the consumer reads the tensor once (in-edge gather)
 * instead of twice.  Same summation order as the pooling kernels: results are bit-identical to the plain entry points.
 *   forward : layer l   (direction NBPC_ROWPOOL_FWD_EMIT) writes P_row_next (B*N, q) = row means of its H_out;
 *             layer l+1 (NBPC_ROWPOOL_FWD_TAKE) is called with p_row_given = 1 and that tensor as P_row (it is then an INPUT);
 *   backward: layer l+1 (NBPC_ROWPOOL_BWD_EMIT) writes dQ_row_prev (B*N, k) = row sums of its dH_in;
 *             layer l   (NBPC_ROWPOOL_BWD_TAKE, relu = 0: dOut arrives masked) is called with dQ_row_given = that tensor.
 * nbpc_graph_layer_rowpool_supported(k, q, is_last, direction) says which layers can do which (today: the 3-channel first layer
 * emits forward, the last layer emits backward, layers with 16 / 32 / 64 channels on the pooled side take); NULL / 0 arguments
 * make the _rp entry points identical to nbpc_graph_layer_fwd / _bwd. */
#define NBPC_ROWPOOL_FWD_EMIT 0
#define NBPC_ROWPOOL_FWD_TAKE 1
#define NBPC_ROWPOOL_BWD_EMIT 2
#define NBPC_ROWPOOL_BWD_TAKE 3
int nbpc_graph_layer_rowpool_supported(int k, int q, int is_last, int direction);
int nbpc_graph_layer_fwd_rp(const float *H_in, const int32_t *col, const int32_t *csrT_ptr, const int32_t *csrT_edge, int B, int N,
                            int M, int k, int q, const float *W, const float *bias, int is_last, int relu, float *H_out,
                            float *P_col, float *P_row, float *P_cube, int p_row_given, float *P_row_next, void *workspace,
                            size_t ws_bytes, void *stream);
int nbpc_graph_layer_bwd_rp(const float *dOut, const float *H_in, const float *H_out, const int32_t *col, const int32_t *csrT_ptr,
                            const int32_t *csrT_edge, int B, int N, int M, int k, int q, const float *W, const float *P_col,
                            const float *P_row, const float *P_cube, int is_last, int relu, int mask_input, float *dH_in,
                            float *dW, float *dB, const float *dQ_row_given, float *dQ_row_prev, void *workspace,
                            size_t ws_bytes, void *stream);

/* Virtual layer input.  Inside a network the output of the FIRST graph layer, H1[e] = relu(E[e] W1 + Q_col[col[e]] +
 * Q_row[e / M]), is a function of the 12-byte edge feature row and two L2-resident node tables; materialising it is 44 % of
 * the bytes a [3,32,16,3] training step moves.  With node_only = 1 nbpc_graph_layer_fwd_v stops after the node-level terms
 * of that layer (H_out may be NULL) and leaves Q_col / Q_row (B*N, q) in the caller's buffers; the NEXT layer is then called
 * with `vin` describing that virtual tensor instead of H_in (NULL): its pooling, forward and backward edge kernels
 * recompute the rows with one shared expression (bit-identical to the materialised tensor).  vin->k = 3 input channels,
 * hidden layers of width 16 / 32 / 64, split-mode arithmetic (NBPC_MATH_TF32X3); nbpc_graph_layer_vin_supported says whether
 * the current configuration has the kernels.  The backward of a virtual-input layer needs mask_input = 1 (the virtual tensor
 * is a ReLU output, its mask is recomputed) and relu = 0 (gradient pre-masked by the consumer). */
typedef struct nbpc_virtual_input {
    const float *E;       /* (c, k) edge features of the producing layer */
    const float *W1;      /* (k, width) its first weight */
    const float *Q_col;   /* (B*N, width) */
    const float *Q_row;   /* (B*N, width) */
    int k;                /* 3 */
} nbpc_virtual_input;
int nbpc_graph_layer_vin_supported(int k0, int k, int q, int64_t c);
int nbpc_graph_layer_fwd_v(const float *H_in, const nbpc_virtual_input *vin, const int32_t *col, const int32_t *csrT_ptr,
                           const int32_t *csrT_edge, int B, int N, int M, int k, int q, const float *W, const float *bias,
                           int is_last, int relu, int node_only, float *H_out, float *P_col, float *P_row, float *P_cube,
                           float *Q_col_out, float *Q_row_out, void *workspace, size_t ws_bytes, void *stream);
int nbpc_graph_layer_bwd_v(const float *dOut, const float *H_in, const nbpc_virtual_input *vin, const float *H_out,
                           const int32_t *col, const int32_t *csrT_ptr, const int32_t *csrT_edge, int B, int N, int M, int k,
                           int q, const float *W, const float *P_col, const float *P_row, const float *P_cube, int is_last,
                           int relu, int mask_input, float *dH_in, float *dW, float *dB, void *workspace, size_t ws_bytes,
                           void *stream);

/* ---------------------------------------------------------------- 15-weight layer on a symmetrised adjacency
 * graph.shift_inv_15op_layer (graph.py:20-200).  The reference ships no builder for its `adj` dict; nbpc_sym_adjacency_*
 * builds the canonical one from a kNN graph: the union A u A^T of every sample, edges sorted row-major, with
 *   row, col (S): node ids of every edge;  all (S): sample of every edge;  tra (S): position of the transposed edge;
 *   dia (B*N): position of the self edge of every node;  dal (B*N): sample of every node;  row_ptr (B*N+1): edge range
 *   of every row.  idx (B,N,M) are the kNN lists (self edge included), csrT_* the in-edge lists of nbpc_adjacency.
 * nbpc_sym_adjacency_count fills row_ptr (row_ptr[B*N] = S, to be read back); nbpc_sym_adjacency_emit the rest;
 * status[0] counts rows without a self edge, status[1] edges without a transposed partner (both must be 0).
 * nbpc_graph15_layer_fwd/bwd: H (S,k) -> H_out (S,q) [+ ReLU]; W (15,k,q), Bias (2,q) as graph.py:135-192; the pooled
 * node tensors Hr, Hc, Hd (B*N,k) and Ha, Hp (B,k) are saved for backward.  Pooled operands are projected once per node
 * and gathered by ONE edge kernel per direction; all reductions have a fixed order. */
size_t nbpc_sym_adjacency_workspace_bytes(int B, int N);
int nbpc_sym_adjacency_count(const int32_t *idx, const int32_t *csrT_ptr, const int32_t *csrT_edge, int B, int N, int M,
                             int32_t *row_ptr, void *workspace, size_t ws_bytes, void *stream);
int nbpc_sym_adjacency_emit(const int32_t *idx, const int32_t *csrT_ptr, const int32_t *csrT_edge, const int32_t *row_ptr,
                            int B, int N, int M, int64_t S, int32_t *row, int32_t *col, int32_t *all, int32_t *tra,
                            int32_t *dia, int32_t *dal, int32_t *status, void *stream);
size_t nbpc_graph15_workspace_bytes(int B, int N, int64_t S, int k, int q);
int nbpc_graph15_layer_fwd(const float *H, const int32_t *row, const int32_t *col, const int32_t *tra, const int32_t *dia,
                           const int32_t *row_ptr, int B, int N, int64_t S, int k, int q, const float *W, const float *Bias,
                           int relu, float *H_out, float *Hr, float *Hc, float *Hd, float *Ha, float *Hp, void *workspace,
                           size_t ws_bytes, void *stream);
int nbpc_graph15_layer_bwd(const float *dOut, const float *H, const float *H_out, const int32_t *row, const int32_t *col,
                           const int32_t *tra, const int32_t *dia, const int32_t *row_ptr, int B, int N, int64_t S, int k,
                           int q, const float *W, const float *Hr, const float *Hc, const float *Hd, const float *Ha,
                           const float *Hp, int relu, float *dH, float *dW, float *dB, void *workspace, size_t ws_bytes,
                           void *stream);

/* ---------------------------------------------------------------- set layer
 * nn.set_layer (nn.py:10-28): out = (H - mean_N H) W + B on (B,N,k) -> (B,N,q); relu optional
 * (nn.py:59, 65-66).  mu (B,k) is saved for backward.
 * Widths with k % 32 == 0, q % 16 == 0, both <= 256 run on tcgen05 (TF32x3 by default, nbpc_set_math_mode): the mean is
 * subtracted from the landed tile before the product, i.e. the association is the reference's (H - mu) W.
 * Backward: relu masks dOut by [H_out > 0]; mask_input multiplies dH_in by [H_in > 0] (the ReLU backward of the layer
 * that produced H_in, fused - nn.network_func_set chains the layers this way so that no masked copy is written). */
size_t nbpc_set_layer_workspace_bytes(int B, int N, int k, int q);
int nbpc_set_layer_fwd(const float *H_in, int B, int N, int k, int q, const float *W,
                       const float *bias, int relu, float *H_out, float *mu, void *workspace,
                       size_t ws_bytes, void *stream);
int nbpc_set_layer_bwd(const float *dOut, const float *H_in, const float *H_out, const float *mu,
                       int B, int N, int k, int q, const float *W, int relu, int mask_input, float *dH_in,
                       float *dW, float *dB, void *workspace, size_t ws_bytes, void *stream);
/* Chained variants for stacks of set layers (nn.network_func_set), where every hidden tensor has exactly one consumer:
 * the kernel that WRITES a tensor also leaves its per-sample column sums, so that the consumer does not re-read the
 * tensor for its mean pass (the sums come out of the tcgen05 epilogue when N % 128 == 0, otherwise from a separate pass -
 * the outputs are filled either way).
 *   fwd: mu_given != 0: mu (B,k) is an INPUT (column means of H_in: the previous layer's mean_out); else it is computed.
 *        mean_out (B,q), optional OUTPUT: per-sample column means of H_out (after the activation).
 *   bwd: dz_sums (B,q), optional INPUT: per-sample column SUMS of dOut; ignored when relu != 0 (the sums of the unmasked
 *        gradient are useless), so pass a pre-masked gradient.  dh_sums (B,k), optional OUTPUT: per-sample column sums of
 *        dH_in (after the input mask) - the dz_sums of the layer below. */
int nbpc_set_layer_fwd_chained(const float *H_in, int B, int N, int k, int q, const float *W, const float *bias, int relu,
                               float *H_out, float *mu, int mu_given, float *mean_out, void *workspace, size_t ws_bytes,
                               void *stream);
int nbpc_set_layer_bwd_chained(const float *dOut, const float *H_in, const float *H_out, const float *mu, int B, int N,
                               int k, int q, const float *W, int relu, int mask_input, float *dH_in, float *dW, float *dB,
                               const float *dz_sums, float *dh_sums, void *workspace, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------- readout / losses
 * nn.loss_ZA (nn.py:151-166): loss = mean_rows(sum_3 (pred - truth)^2); rows = B*N.
 * nn.pbc_loss / periodic_boundary_dist (nn.py:123-148): per-axis min over {d, d-1, d+1} images,
 *   optionally * 1e5.  nn.get_readout (nn.py:107-119): wrap the first 3 channels into [0,1).
 * pred/truth have leading dimensions (floats) ld_pred / ld_truth >= 3; only columns 0..2 are read.
 * *_bwd: dpred[:, :3] = dloss * d loss / d pred (dloss is a DEVICE scalar). */
size_t nbpc_loss_workspace_bytes(int64_t rows);
int nbpc_loss_za_fwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                     float *loss_out, void *workspace, size_t ws_bytes, void *stream);
int nbpc_loss_za_bwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                     const float *dloss, float *dpred, int ld_dpred, void *stream);
int nbpc_pbc_loss_fwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                      int scale_error, float *loss_out, void *workspace, size_t ws_bytes,
                      void *stream);
int nbpc_pbc_loss_bwd(const float *pred, int ld_pred, const float *truth, int ld_truth, int64_t rows,
                      int scale_error, const float *dloss, float *dpred, int ld_dpred, void *stream);
int nbpc_periodic_boundary_dist(const float *pred, int ld_pred, const float *truth, int ld_truth,
                                int64_t rows, float *dist_out, void *stream);
int nbpc_readout(const float *h, int64_t rows, int C, float *out, void *stream);
/* Scaled residual update of the multi-redshift model (graph.py:558-566, the reference's legacy block): X (rows, ldx >= 6) =
 * [position, velocity], net (rows, C) with C = 3 | 6:  out[:, :3] = net[:, :3] * loc_scalar + loc + vel * vel_scalar,
 * out[:, 3:6] = net[:, 3:6] * vel_scalar + vel.  Forward only (inference / rollout; with trainable scalars the facade
 * composes the update from differentiable ops). */
int nbpc_residual_update(const float *X, int ldx, const float *net, int C, int64_t rows, float loc_scalar, float vel_scalar,
                         float *out, void *stream);

/* ---------------------------------------------------------------- optimiser
 * tf.train.AdamOptimizer as used by train.py:70 (TF "epsilon-hat" form):
 *   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m,v updates;  p -= lr_t * m / (sqrt(v) + eps).
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce). */
int nbpc_adam_tf(float *param, const float *grad, float *m, float *v, int64_t n, float lr,
                 float beta1, float beta2, float eps, int64_t step, float grad_scale, void *stream);

/* Same step with the step count t kept in device memory (*step_counter is incremented first, then used): every launch
 * parameter is then identical from step to step, so a whole training step can be captured in a CUDA graph and
 * replayed.  If the incremented counter is still <= 0 the call updates nothing: a pipelined loop that applies the update
 * of step t at the start of step t+1 (overlapping the gradient all-reduce with the next graph build) starts it at -1. */
int nbpc_adam_tf_dev(float *param, const float *grad, float *m, float *v, int64_t n, float lr,
                     float beta1, float beta2, float eps, int64_t *step_counter, float grad_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NBPC_H_ */
