// pad_cube.cu - the padded ("cloned") cube of the reference's periodic kNN as an explicit product:
// graph.pad_cube_boundaries (/root/reference/graph.py:827-855) with face_outer / edge_outer / corner_outer
// (graph.py:801-825).  nbpc_knn(periodic=1) never materialises this cloud (it generates the images on the fly
// from the same rule); these two entry points exist for callers that want the padded cloud or idx_map itself.
//
//   bound  = where(x >= upper, -1, where(x <= lower, +1, 0))            float32 compare, graph.py:842
//   images = rows of (pattern * bound) + particle                       int64 + float32 -> FLOAT64, graph.py:801-816
//   order  = particles in ascending index, each followed by its 1 / 3 / 7 images in the reference's pattern order
#include "nbpc_common.cuh"
#include "scan.cuh"

__device__ __forceinline__ int pc_bound(float x, float lower, float upper) { return x >= upper ? -1 : (x <= lower ? 1 : 0); }

// counts[i] = number of images of particle i (0, 1, 3 or 7); counts[N] = 0 (so that the exclusive scan over N+1
// entries leaves the total at [N])
__global__ void pad_cube_count_kernel(const float *__restrict__ xyz, int64_t stride_n, int N, float lower, float upper,
                                      int32_t *__restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    if (i == N) { counts[N] = 0; return; }
    const float *p = xyz + (int64_t)i * stride_n;
    const int nb = (pc_bound(p[0], lower, upper) != 0) + (pc_bound(p[1], lower, upper) != 0) + (pc_bound(p[2], lower, upper) != 0);
    counts[i] = nb ? (1 << nb) - 1 : 0;
}

// padded (N + n_img, 3) float64: rows [0, N) = the particles, rows N + offsets[i] + r = image r of particle i;
// idx_map (n_img,) int64 = source particle of every image
__global__ void pad_cube_emit_kernel(const float *__restrict__ xyz, int64_t stride_n, int N, float lower, float upper,
                                     const int32_t *__restrict__ offsets, double *__restrict__ padded,
                                     int64_t *__restrict__ idx_map) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *p = xyz + (int64_t)i * stride_n;
    const double x[3] = {(double)p[0], (double)p[1], (double)p[2]};
    const int b[3] = {pc_bound(p[0], lower, upper), pc_bound(p[1], lower, upper), pc_bound(p[2], lower, upper)};
    padded[(int64_t)i * 3 + 0] = x[0];
    padded[(int64_t)i * 3 + 1] = x[1];
    padded[(int64_t)i * 3 + 2] = x[2];
    const int nb = (b[0] != 0) + (b[1] != 0) + (b[2] != 0);
    if (!nb) return;
    // pattern rows as 3-bit masks, bit d = coordinate d is relocated
    int rows[7], n_rows;
    if (nb == 1) {                       // face_outer: bound + particle
        n_rows = 1;
        rows[0] = 7;
    } else if (nb == 2) {                // edge_outer: roll([[0,1,1],[0,1,0],[0,0,1]], zero_idx, axis=1)
        n_rows = 3;
        const int z = b[0] == 0 ? 0 : (b[1] == 0 ? 1 : 2);
        const int base[3][3] = {{0, 1, 1}, {0, 1, 0}, {0, 0, 1}};
        for (int r = 0; r < 3; ++r) {
            int m = 0;
            for (int j = 0; j < 3; ++j)
                if (base[r][j]) m |= 1 << ((j + z) % 3);
            rows[r] = m;
        }
    } else {                             // corner_outer
        n_rows = 7;
        const int base[7][3] = {{1, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 0, 0}, {0, 1, 1}, {0, 1, 0}, {0, 0, 1}};
        for (int r = 0; r < 7; ++r) rows[r] = base[r][0] | (base[r][1] << 1) | (base[r][2] << 2);
    }
    const int64_t o = offsets[i];
    for (int r = 0; r < n_rows; ++r) {
        double *dst = padded + ((int64_t)N + o + r) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) dst[d] = (double)(((rows[r] >> d) & 1) * b[d]) + x[d];   // (pattern * bound) + particle
        idx_map[o + r] = i;
    }
}

extern "C" {

size_t nbpc_pad_cube_workspace_bytes(int N) {
    if (N < 1) return 0;
    return nbpc_align_up(sizeof(int32_t) * nbpc_scan_partials_count((int64_t)N + 1));
}

int nbpc_pad_cube_count(const float *xyz, int64_t stride_n, int N, double boundary_threshold, int32_t *offsets,
                        void *workspace, size_t ws_bytes, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(xyz && offsets && workspace, "null pointer");
    NBPC_ARG(N >= 1 && stride_n >= 3, "bad sizes");
    NBPC_ARG(boundary_threshold >= 0.0 && boundary_threshold <= 1.0, "boundary_threshold must be in [0,1]");
    if (ws_bytes < nbpc_pad_cube_workspace_bytes(N)) {
        nbpc_set_error("nbpc_pad_cube_count: workspace too small");
        return NBPC_EWORKSPACE;
    }
    const float lower = (float)boundary_threshold, upper = (float)(1.0 - boundary_threshold);
    NBPC_LAUNCH(pad_cube_count_kernel, nbpc_cdiv((int64_t)N + 1, 256), 256, 0, stream, xyz, stride_n, N, lower, upper, offsets);
    NBPC_TRY(nbpc_exclusive_scan_i32(offsets, (int64_t)N + 1, (int32_t *)workspace, stream));
    return nbpc_check_launch("nbpc_pad_cube_count");
}

int nbpc_pad_cube_emit(const float *xyz, int64_t stride_n, int N, double boundary_threshold, const int32_t *offsets,
                       double *padded_out, int64_t *idx_map_out, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(xyz && offsets && padded_out && idx_map_out, "null pointer");
    NBPC_ARG(N >= 1 && stride_n >= 3, "bad sizes");
    const float lower = (float)boundary_threshold, upper = (float)(1.0 - boundary_threshold);
    NBPC_LAUNCH(pad_cube_emit_kernel, nbpc_cdiv(N, 256), 256, 0, stream, xyz, stride_n, N, lower, upper, offsets, padded_out, idx_map_out);
    return nbpc_check_launch("nbpc_pad_cube_emit");
}

}  // extern "C"
