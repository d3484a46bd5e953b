// adjacency.cu - adjacency format conversion (graph.py:593-697) and the CSR transpose that makes
// col-pooling and the backward scatter deterministic (members of every segment in ascending order).
#include "nbpc_common.cuh"
#include "scan.cuh"

// ------------------------------------------------------------------ segment CSR (generic)
// counts -> seg_ptr (shifted by one so that the scan can run in place), range/sortedness flags
__global__ void seg_count_kernel(const int32_t *__restrict__ ids, int64_t n, int num_segs,
                                 int32_t *__restrict__ seg_ptr, int32_t *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool unsorted = false;
    if (i < n) {
        const int id = ids[i];
        if (id < 0 || id >= num_segs) {
            atomicAdd(&flags[1], 1);
        } else {
            atomicAdd(&seg_ptr[id], 1);
            unsorted = i > 0 && ids[i - 1] > id;   // not monotone
        }
    }
    // one flag update per block at most (a kNN column list is unsorted nearly everywhere: a per-thread atomicOr
    // serialised ~2 M updates of one address and cost 80 us of this kernel's 96 us at 3.67 M edges)
#ifndef NBPC_HOST_EMU
    if (__syncthreads_or(unsorted) && threadIdx.x == 0 && *(volatile int32_t *)&flags[0] == 0) atomicOr(&flags[0], 1);
#else
    if (unsorted) atomicOr(&flags[0], 1);
#endif
}

__global__ void seg_fill_kernel(const int32_t *__restrict__ ids, int64_t n, int num_segs,
                                const int32_t *__restrict__ seg_ptr, int32_t *__restrict__ fill,
                                int32_t *__restrict__ members, const int32_t *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    if (id < 0 || id >= num_segs) return;
    if (!flags[0] && !flags[1]) {  // monotone ids: the stable order is the identity
        members[i] = (int32_t)i;
        return;
    }
    const int pos = seg_ptr[id] + atomicAdd(&fill[id], 1);
    members[pos] = (int32_t)i;
}

// one thread per segment: insertion sort of its (short) member list; O(len^2), len ~ in-degree
__global__ void seg_sort_kernel(const int32_t *__restrict__ seg_ptr, int32_t *__restrict__ members,
                                int num_segs, const int32_t *__restrict__ flags) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= num_segs) return;
    if (!flags[0] && !flags[1]) return;
    const int b = seg_ptr[s], e = seg_ptr[s + 1];
    for (int i = b + 1; i < e; ++i) {
        const int v = members[i];
        int j = i - 1;
        while (j >= b && members[j] > v) {
            members[j + 1] = members[j];
            --j;
        }
        members[j + 1] = v;
    }
}

struct SegWorkspace {
    int32_t *fill, *partials, *flags, *rank;
    size_t bytes;
};
static SegWorkspace seg_carve(void *ws, size_t ws_bytes, int64_t n_items, int num_segs, bool with_rank = false) {
    NbpcArena a(ws, ws_bytes);
    SegWorkspace w;
    w.fill = a.take<int32_t>((size_t)num_segs);
    w.partials = a.take<int32_t>(nbpc_scan_partials_count((int64_t)num_segs + 1));
    w.flags = a.take<int32_t>(2);
    w.rank = a.take<int32_t>(with_rank ? (size_t)n_items : 0);
    w.bytes = a.off;
    return w;
}

static int segment_csr_impl(const int32_t *ids, int64_t n, int num_segs, int32_t *seg_ptr, int32_t *members,
                            int32_t *status, const SegWorkspace &w, cudaStream_t stream) {
    if (nbpc_memset_async(seg_ptr, 0, sizeof(int32_t) * ((size_t)num_segs + 1), stream) ||
        nbpc_memset_async(w.fill, 0, sizeof(int32_t) * (size_t)num_segs, stream) ||
        nbpc_memset_async(w.flags, 0, sizeof(int32_t) * 2, stream)) {
        nbpc_set_error("segment_csr: memset failed");
        return NBPC_ELAUNCH;
    }
    const int T = 256;
    NBPC_LAUNCH(seg_count_kernel, nbpc_cdiv(n, T), T, 0, stream, ids, n, num_segs, seg_ptr, w.flags);
    NBPC_TRY(nbpc_exclusive_scan_i32(seg_ptr, (int64_t)num_segs + 1, w.partials, stream));
    NBPC_LAUNCH(seg_fill_kernel, nbpc_cdiv(n, T), T, 0, stream, ids, n, num_segs, seg_ptr, w.fill, members, w.flags);
    NBPC_LAUNCH(seg_sort_kernel, nbpc_cdiv(num_segs, T), T, 0, stream, seg_ptr, members, num_segs, w.flags);
    (void)status;
    return nbpc_check_launch("segment_csr");
}

__global__ void seg_status_kernel(const int32_t *flags, int32_t *status) {
    if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = flags[1];
}

// ------------------------------------------------------------------ kNN index list -> COO (+diag)
__global__ void adj_coo_kernel(const int32_t *__restrict__ idx, int B, int N, int M,
                               int32_t *__restrict__ coo, int32_t *__restrict__ status) {
    const int64_t c = (int64_t)B * N * M;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= c) return;
    const int row = (int)(e / M);  // global node id = sample*N + particle
    const int s = row / N;
    int col = idx[e];
    if (col < 0 || col >= N) {
        atomicAdd(&status[1], 1);
        col = col < 0 ? 0 : N - 1;
    }
    coo[e] = row;              // graph.py:644
    coo[c + e] = s * N + col;  // graph.py:645
    coo[2 * c + e] = s;        // graph.py:646
}

__global__ void adj_diag_kernel(const int32_t *__restrict__ coo_col, int BN, int M,
                                int64_t *__restrict__ diag, int32_t *__restrict__ status) {
    int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= BN) return;
    int64_t first = -1;
    int cnt = 0;
    for (int m = 0; m < M; ++m) {
        const int64_t e = (int64_t)node * M + m;
        if (coo_col[e] == node) {
            if (first < 0) first = e;
            ++cnt;
        }
    }
    diag[node] = first;  // graph.py:655-656
    if (cnt != 1) atomicAdd(&status[0], 1);
}

// ------------------------------------------------------------------ kNN adjacency, fused pipeline
// One pass over the neighbour lists writes the COO rows AND counts the in-degrees; the value the counting atomic returns
// is the edge's (arbitrary but unique) rank inside its column bucket, so the fill pass is a plain scatter without atomics.
__global__ void adj_coo_count_kernel(const int32_t *__restrict__ idx, int B, int N, int M, int32_t *__restrict__ coo,
                                     int32_t *__restrict__ seg_ptr, int32_t *__restrict__ rank, int32_t *__restrict__ status) {
    const int64_t c = (int64_t)B * N * M;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= c) return;
    const int row = (int)(e / M);
    const int s = row / N;
    int col = idx[e];
    if (col < 0 || col >= N) {
        atomicAdd(&status[1], 1);
        col = col < 0 ? 0 : N - 1;
    }
    const int gcol = s * N + col;
    coo[e] = row;              // graph.py:644
    coo[c + e] = gcol;         // graph.py:645
    coo[2 * c + e] = s;        // graph.py:646
    rank[e] = atomicAdd(&seg_ptr[gcol], 1);
}

__global__ void adj_fill_kernel(const int32_t *__restrict__ gcol, const int32_t *__restrict__ rank, int64_t c,
                                const int32_t *__restrict__ seg_ptr, int32_t *__restrict__ members) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= c) return;
    members[__ldg(&seg_ptr[gcol[e]]) + rank[e]] = (int32_t)e;
}

#ifndef NBPC_HOST_EMU
// ascending edge ids inside every bucket: a block stages the contiguous member range of SEGS consecutive buckets in
// shared memory, every thread insertion-sorts its own (short) bucket there, the range goes back coalesced.  Ranges that
// do not fit (pathological in-degrees) are sorted in global memory.
#define ADJ_SORT_SEGS 256
#define ADJ_SORT_CAP 8192
__global__ void __launch_bounds__(ADJ_SORT_SEGS) adj_sort_kernel(const int32_t *__restrict__ seg_ptr, int32_t *__restrict__ members,
                                                                 int num_segs) {
    __shared__ int32_t buf[ADJ_SORT_CAP];
    const int s0 = blockIdx.x * ADJ_SORT_SEGS, s1 = nbpc_min(s0 + ADJ_SORT_SEGS, num_segs);
    const int lo = seg_ptr[s0], hi = seg_ptr[s1], n = hi - lo;
    const int s = s0 + threadIdx.x;
    const bool fits = n <= ADJ_SORT_CAP;
    int32_t *base = fits ? buf - lo : members;
    if (fits) {
        for (int i = threadIdx.x; i < n; i += ADJ_SORT_SEGS) buf[i] = members[lo + i];
        __syncthreads();
    }
    if (s < s1) {
        const int b = seg_ptr[s], e = seg_ptr[s + 1];
        for (int i = b + 1; i < e; ++i) {
            const int v = base[i];
            int j = i - 1;
            while (j >= b && base[j] > v) {
                base[j + 1] = base[j];
                --j;
            }
            base[j + 1] = v;
        }
    }
    if (fits) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += ADJ_SORT_SEGS) members[lo + i] = buf[i];
    }
}
#endif

extern "C" {

size_t nbpc_segment_csr_workspace_bytes(int64_t n_items, int num_segs) {
    if (n_items < 0 || num_segs < 1) return 0;
    return seg_carve(nullptr, 0, n_items, num_segs).bytes;
}

int nbpc_segment_csr(const int32_t *ids, int64_t n_items, int num_segs, int32_t *seg_ptr, int32_t *seg_members,
                     int32_t *status, void *workspace, size_t ws_bytes, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(ids && seg_ptr && seg_members && status && workspace, "null pointer");
    NBPC_ARG(n_items >= 0 && n_items < ((int64_t)1 << 31) && num_segs >= 1, "bad sizes");
    SegWorkspace w = seg_carve(workspace, ws_bytes, n_items, num_segs);
    if (w.bytes > ws_bytes) {
        nbpc_set_error("nbpc_segment_csr: workspace too small");
        return NBPC_EWORKSPACE;
    }
    NBPC_TRY(segment_csr_impl(ids, n_items, num_segs, seg_ptr, seg_members, status, w, stream));
    NBPC_LAUNCH(seg_status_kernel, 1, 32, 0, stream, w.flags, status);
    return nbpc_check_launch("nbpc_segment_csr");
}

size_t nbpc_adjacency_workspace_bytes(int B, int N, int M) {
    if (B < 1 || N < 1 || M < 1) return 0;
    return seg_carve(nullptr, 0, (int64_t)B * N * M, B * N, true).bytes;
}

int nbpc_adjacency(const int32_t *idx, int B, int N, int M, int32_t *coo_out, int64_t *diag_out,
                   int32_t *csrT_ptr, int32_t *csrT_edge, int32_t *status, void *workspace, size_t ws_bytes,
                   void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(idx && coo_out && diag_out && csrT_ptr && csrT_edge && status && workspace, "null pointer");
    NBPC_ARG(B >= 1 && N >= 1 && M >= 1, "B, N, M must be positive");
    const int64_t c = (int64_t)B * N * M;
    NBPC_ARG(c < ((int64_t)1 << 31), "B*N*M must fit int32");
    SegWorkspace w = seg_carve(workspace, ws_bytes, c, B * N, true);
    if (w.bytes > ws_bytes) {
        nbpc_set_error("nbpc_adjacency: workspace too small");
        return NBPC_EWORKSPACE;
    }
    if (nbpc_memset_async(status, 0, sizeof(int32_t) * 2, stream)) {
        nbpc_set_error("nbpc_adjacency: memset failed");
        return NBPC_ELAUNCH;
    }
    const int T = 256;
#ifndef NBPC_HOST_EMU
    {
        // fused pipeline: COO rows + in-degree count (+ bucket rank) -> scan -> atomic-free fill -> shared-memory bucket sort
        const int num_segs = B * N;
        if (nbpc_memset_async(csrT_ptr, 0, sizeof(int32_t) * ((size_t)num_segs + 1), stream)) {
            nbpc_set_error("nbpc_adjacency: memset failed");
            return NBPC_ELAUNCH;
        }
        NBPC_LAUNCH(adj_coo_count_kernel, nbpc_cdiv(c, T), T, 0, stream, idx, B, N, M, coo_out, csrT_ptr, w.rank, status);
        NBPC_LAUNCH(adj_diag_kernel, nbpc_cdiv(num_segs, T), T, 0, stream, coo_out + c, num_segs, M, diag_out, status);
        NBPC_TRY(nbpc_exclusive_scan_i32(csrT_ptr, (int64_t)num_segs + 1, w.partials, stream));
        NBPC_LAUNCH(adj_fill_kernel, nbpc_cdiv(c, T), T, 0, stream, coo_out + c, w.rank, c, csrT_ptr, csrT_edge);
        NBPC_LAUNCH(adj_sort_kernel, nbpc_cdiv(num_segs, ADJ_SORT_SEGS), ADJ_SORT_SEGS, 0, stream, csrT_ptr, csrT_edge, num_segs);
        return nbpc_check_launch("nbpc_adjacency");
    }
#endif
    NBPC_LAUNCH(adj_coo_kernel, nbpc_cdiv(c, T), T, 0, stream, idx, B, N, M, coo_out, status);
    NBPC_LAUNCH(adj_diag_kernel, nbpc_cdiv(B * N, T), T, 0, stream, coo_out + c, B * N, M, diag_out, status);
    NBPC_TRY(segment_csr_impl(coo_out + c, c, B * N, csrT_ptr, csrT_edge, status, w, stream));
    return nbpc_check_launch("nbpc_adjacency");
}

}  // extern "C"
