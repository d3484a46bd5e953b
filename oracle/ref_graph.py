"""TEST INFRASTRUCTURE ONLY - the CPU oracle for the graph-construction half of the path.

A NumPy/SciPy restatement of the reference's host-side graph functions, each
citing the /root/reference file:line it follows.  The kNN arithmetic itself is
NOT in the reference tree: it is scikit-learn's `kneighbors_graph` (version
unpinned by the reference; this image has scikit-learn 1.9.0, SciPy 1.18.1,
NumPy 2.3.5).  Two kNN back-ends are offered:

  * backend="sklearn": calls the same third-party entry point the reference
    calls (graph.py:709, 887) - this is what `bench.py`'s cpu_baseline times.
  * backend="exact":  `oracle/knn_exact.c` - brute-force float64 kNN with the
    published sklearn distance arithmetic (d = ((dx*dx)+(dy*dy))+(dz*dz), separate
    multiply and add; sklearn/metrics/_dist_metrics.pxd.tp `rdist`) and the
    documented (d2, index) tie-break of the CUDA kernel.  Independent of sklearn.

Parity pinned: tests/test_oracle_golden.py checks every function here against
golden vectors produced by running the unmodified reference through
oracle/load_reference.py (script: oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import numpy as np
from scipy.sparse import csr_matrix


# --------------------------------------------------------------------------- kNN
def _knn_indices_sklearn(x, k, include_self):
    from sklearn.neighbors import kneighbors_graph
    return kneighbors_graph(x, k, include_self=include_self)


def _knn_indices_exact(x64, k, include_self, n_query):
    """x64: (P,3) float64 cloud; rows [0,n_query) are the queries. Returns (n_query,k) int64,
    ascending (d2, index); self excluded when include_self is False."""
    from oracle.knn_exact import knn_exact
    return knn_exact(x64, n_query, k, include_self)


def _csr_from_indices(ind, n_cols, dtype=np.float64, index_dtype=np.int64):
    n, k = ind.shape
    indptr = np.arange(0, n * k + 1, k).astype(index_dtype)
    return csr_matrix((np.ones(n * k, dtype=dtype), ind.reshape(-1).astype(index_dtype), indptr),
                      shape=(n, n_cols))


def get_kneighbor_list(X_in, M, offset_idx=False, include_self=True, backend="sklearn"):
    """graph.py:704-713.  Per sample: kneighbors_graph(X[i,:,:3], M).astype(float32).
    SciPy's astype goes through `_deduped_data()` which sorts each row's columns and
    downcasts indices to int32 => per-row ASCENDING COLUMN order (SURVEY.md §7 hard part 3)."""
    b, N, D = X_in.shape
    lst_csrs = []
    for i in range(b):
        pts = X_in[i, :, :3]
        if backend == "sklearn":
            kgraph = _knn_indices_sklearn(pts, M, include_self).astype(np.float32)
        else:
            ind = _knn_indices_exact(np.asarray(pts, dtype=np.float64), M, include_self, N)
            kgraph = _csr_from_indices(ind, N).astype(np.float32)
        if offset_idx:
            kgraph.indices = kgraph.indices + (N * i)
        lst_csrs.append(kgraph)
    return lst_csrs


# boundary images ------------------------------------------------------------
_EDGE = np.array([[0, 1, 1], [0, 1, 0], [0, 0, 1]])                                   # graph.py:809
_CORNER = np.array([[1, 1, 1], [1, 1, 0], [1, 0, 1], [1, 0, 0], [0, 1, 1], [0, 1, 0], [0, 0, 1]])  # graph.py:815


def pad_cube_boundaries(x, boundary_threshold):
    """graph.py:827-855 (+ face/edge/corner_outer 801-825), O(N) instead of the
    reference's O(N^2) repeated np.concatenate, same rows in the same order.

    bound = where(x >= 1-thr, -1, where(x <= thr, +1, 0)); a particle with 1/2/3 axes in
    the boundary layer gets 1/3/7 images `mask * bound + particle`.  The sum int64 +
    float32 promotes to float64, so the padded cloud is float64 (original rows upcast
    exactly)."""
    N, D = x.shape
    lower = boundary_threshold
    upper = 1 - boundary_threshold
    bound_x = np.where(x >= upper, -1, np.where(x <= lower, 1, 0))
    bound_x_count = np.count_nonzero(bound_x, axis=-1)
    outer = []
    idx_list = []
    for idx in np.nonzero(bound_x_count)[0]:
        nb = bound_x_count[idx]
        bound = bound_x[idx]
        particle = x[idx]
        if nb == 1:
            img = (bound + particle)[None, :]
        elif nb == 2:
            zero_idx = list(bound).index(0)
            img = (np.roll(_EDGE, zero_idx, 1) * bound) + particle
        else:
            img = (_CORNER * bound) + particle
        outer.append(img.astype(np.float64))
        idx_list.extend([idx] * img.shape[0])
    if outer:
        padded = np.concatenate([x.astype(np.float64)] + outer, axis=0)
    else:
        padded = x
    return padded, np.array(idx_list, dtype=np.int64)


def get_pcube_csr(x, idx_map, N, K, include_self=False, backend="sklearn"):
    """graph.py:877-894: kNN on the padded cloud, first N rows, image columns mapped back."""
    if backend == "sklearn":
        kgraph = _knn_indices_sklearn(x, K, include_self)[:N]
    else:
        ind = _knn_indices_exact(np.asarray(x, dtype=np.float64), K, include_self, N)
        kgraph = _csr_from_indices(ind, x.shape[0])
    ind = kgraph.indices
    outer = ind >= N
    if outer.any():
        ind[outer] = idx_map[ind[outer] - N]
    kgraph.indices = ind
    return kgraph


def get_pbc_kneighbors_csr(X, K, boundary_threshold, include_self=False, backend="sklearn"):
    """graph.py:896-917.  Output rows stay DISTANCE-sorted (no astype), shape (N, N_padded)."""
    mb_size, N, D = X.shape
    csr_list = []
    clone = np.copy(X[..., :3])
    for b in range(mb_size):
        padded_cube, idx_map = pad_cube_boundaries(clone[b], boundary_threshold)
        csr_list.append(get_pcube_csr(padded_cube, idx_map, N, K, include_self, backend=backend))
    return csr_list


# ------------------------------------------------------------------- adjacency format
def get_indices_from_list_CSR(A, offset=True):
    """graph.py:593-610"""
    b = len(A)
    N = A[0].shape[0]
    M = A[0].indices.shape[0] // N
    out = np.zeros((b * N * M), dtype=np.int32)
    for i in range(b):
        out[i * N * M:(i + 1) * N * M] = A[i].indices + i * N
    return out


def _rows_cols(csr):
    # scipy `nonzero()` / `tocoo()` keep CSR storage order; data are all ones
    N = csr.shape[0]
    counts = np.diff(csr.indptr)
    r = np.repeat(np.arange(N, dtype=np.int32), counts)
    return r, csr.indices.astype(np.int32)


def to_coo_batch_ZA_diag(A):
    """graph.py:621-662.  COO[0]=row+iN, COO[1]=col+iN, COO[2]=i (int32);
    diagonals = flat edge positions where row == col (int64)."""
    b = len(A)
    N = A[0].shape[0]
    M = A[0].indices.shape[0] // N
    dia = []
    COO_feats = np.zeros((3, b * N * M), dtype=np.int32)
    for i in range(b):
        r, c = _rows_cols(A[i])
        k, q = i * N * M, (i + 1) * N * M
        COO_feats[0, k:q] = r + i * N
        COO_feats[1, k:q] = c + i * N
        COO_feats[2, k:q] = i
        d = np.where(r == c)[0]
        dia.extend(d + i * len(r))
    return COO_feats, np.array(dia)


def to_coo_batch(A):
    """graph.py:664-697"""
    return to_coo_batch_ZA_diag(A)[0]


def get_symmetrized_adjacency(A):
    """The `adj` dict shift_inv_15op_layer documents (graph.py:46-59); the reference ships no builder for it.
    Symmetrised kNN graph A u A^T per sample, edges in row-major (row, col) order, ids shifted by i*N."""
    b = len(A)
    N = A[0].shape[0]
    M = A[0].indices.shape[0] // N
    BN = b * N
    rows, cols = [], []
    for i, a in enumerate(A):
        r = np.repeat(np.arange(N, dtype=np.int64), M) + i * N
        c = a.indices.astype(np.int64) + i * N
        rows += [r, c]
        cols += [c, r]
    keys = np.unique(np.concatenate(rows) * BN + np.concatenate(cols))
    row, col = keys // BN, keys % BN
    nodes = np.arange(BN, dtype=np.int64)
    tra = np.searchsorted(keys, col * BN + row)
    dia = np.searchsorted(keys, nodes * BN + nodes)
    assert np.array_equal(keys[dia], nodes * BN + nodes), "self edges required (include_self=True)"
    i32 = lambda t: t.astype(np.int32)
    return {"row": i32(row), "col": i32(col), "all": i32(row // N), "tra": i32(tra), "dia": i32(dia), "dal": i32(nodes // N)}
