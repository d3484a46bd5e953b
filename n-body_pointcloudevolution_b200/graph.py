"""Drop-in for the hot-path functions of /root/reference/graph.py - same names, argument order
and return structure, with torch CUDA tensors where the reference had NumPy arrays / TF tensors.

Graph construction (reference: scikit-learn KD-tree on the host, every step) and every layer op
(reference: TF stock ops) run as hand-written sm_100a kernels in libnbpc.so.  No CPU fallback.

Differences a caller can observe (all documented in DESIGN.md):
  * kNN results are `KnnCSR` objects exposing the SciPy-CSR attributes the reference touches
    (`.indices`, `.indptr`, `.shape`, `.nonzero()`, `.tocoo()`) as device tensors;
  * exact-distance ties are broken by ascending particle index (sklearn: KD-tree traversal order);
  * `COO_feats` returned by `to_coo_batch*` carries the precomputed CSR transpose as an attribute;
    a plain (3,c) tensor/array is accepted too (the transpose is then built on first use).
"""
from types import SimpleNamespace

import os as _os
import weakref

import numpy as np
import torch

from . import ops

__all__ = [
    "KnnCSR", "Adjacency", "get_kneighbor_list", "get_pbc_kneighbors_csr", "pad_cube_boundaries", "get_pcube_csr",
    "get_pcube_adjacency_list",
    "get_indices_from_list_CSR", "to_coo_batch_ZA_diag", "to_coo_batch", "confirm_CSR_to_COO_index_integrity",
    "include_node_features", "get_input_features_shift_inv_ZA", "get_input_features_shift_inv",
    "shift_inv_conv", "shift_inv_layer", "network_func_shift_inv_za", "model_func_shift_inv_za",
    "network_func_shift_inv", "model_func_shift_inv", "rollout_shift_inv",
    "get_symmetrized_adjacency", "shift_inv_15op_layer", "network_func_15op_shift_inv_za", "model_func_15op_shift_inv_za",
]


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("n-body_pointcloudevolution_b200 needs an sm_100 (B200) GPU; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(x, dtype=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        x = x.to(_dev(), non_blocking=True)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x


# =============================================================================== kNN results
class KnnCSR:
    """One sample's kNN graph: the subset of scipy.sparse.csr_matrix the reference uses."""

    def __init__(self, batch_idx, sample, n_cols=None, offset=0, status=None):
        self._status = status            # int32[1] device: out-of-unit-box particles of a periodic build
        self._batch = batch_idx          # (B, N, M) int32, shared by the whole batch (local indices)
        self._sample = sample
        self._offset = offset            # graph.py:710-711 offset_idx
        self._n_cols = n_cols
        _, self.N, self.M = batch_idx.shape

    @property
    def indices(self):
        ind = self._batch[self._sample].reshape(-1)
        return ind + self._offset if self._offset else ind

    @property
    def indptr(self):
        return torch.arange(0, self.N * self.M + 1, self.M, dtype=torch.int32, device=self._batch.device)

    @property
    def data(self):
        return torch.ones(self.N * self.M, dtype=torch.float32, device=self._batch.device)

    @property
    def shape(self):
        n_cols = self._n_cols() if callable(self._n_cols) else (self._n_cols or self.N)
        return (self.N, int(n_cols))

    def nonzero(self):
        rows = torch.arange(self.N, dtype=torch.int32, device=self._batch.device).repeat_interleave(self.M)
        return rows, self.indices

    def tocoo(self):
        r, c = self.nonzero()
        return SimpleNamespace(row=r, col=c, data=self.data, shape=self.shape)

    def check(self):
        """Host-synchronising validation of a periodic build: get_pbc_kneighbors_csr assumes the unit box (like the
        reference's pad_cube_boundaries, graph.py:827-855); raises if some particle had a coordinate outside [0,1]."""
        if self._status is not None and int(self._status.item()):
            raise ValueError(f"periodic kNN: {int(self._status.item())} particles have coordinates outside the unit box [0,1]; "
                             "wrap them first (nn.get_readout)")
        return True

    def toarray_indices(self):
        """(N, M) view of the neighbour indices."""
        return self.indices.view(self.N, self.M)


def _as_batch(A):
    """list of KnnCSR produced by one kNN call -> the shared (B,N,M) tensor (no copy), else stack."""
    if isinstance(A, torch.Tensor) and A.dim() == 3:
        return _to_cuda(A, torch.int32)
    if all(isinstance(a, KnnCSR) for a in A):
        first = A[0]._batch
        if (first.shape[0] == len(A) and all(a._batch is first and a._sample == i and not a._offset
                                              for i, a in enumerate(A))):
            return first
        return torch.stack([a._batch[a._sample] for a in A])
    # scipy CSR matrices (e.g. produced by the reference itself)
    N = A[0].shape[0]
    return _to_cuda(np.stack([np.asarray(a.indices).reshape(N, -1) for a in A]), torch.int32)


def get_kneighbor_list(X_in, M, offset_idx=False, include_self=True):
    """graph.py:704-713.  X_in (b,N,D>=3) -> list of b KnnCSR of shape (N,N); per-row indices in
    ASCENDING COLUMN order, int32 (what `.astype(np.float32)` does to the sklearn result)."""
    X = _to_cuda(X_in, torch.float32)
    b, N, D = X.shape
    idx, _, _ = ops.knn(X, int(M), False, 0.0, bool(include_self), ops._lib.ORDER_INDEX, False)
    return [KnnCSR(idx, i, N, offset=(N * i if offset_idx else 0)) for i in range(b)]


def _n_padded(X, thr):
    """Number of points of the reference's padded cloud (graph.py:827-855): N + sum(2^nb - 1)."""
    def f():
        x = X[..., :3]
        upper = torch.tensor(1 - thr, dtype=torch.float32, device=X.device)
        lower = torch.tensor(thr, dtype=torch.float32, device=X.device)
        nb = ((x >= upper) | (x <= lower)).sum(-1)
        return X.shape[0] + int(((2 ** nb) - 1).sum().item())
    return f


def get_pbc_kneighbors_csr(X, K, boundary_threshold, include_self=False):
    """graph.py:896-917 (+ pad_cube_boundaries 827-855, get_pcube_csr 877-894).  Unit periodic box;
    rows stay DISTANCE-sorted, image columns are mapped back to the original particle."""
    Xc = _to_cuda(X, torch.float32)
    mb_size, N, D = Xc.shape
    idx, _, status = ops.knn(Xc, int(K), True, float(boundary_threshold), bool(include_self), ops._lib.ORDER_DISTANCE, False)
    return [KnnCSR(idx, i, _n_padded(Xc[i], float(boundary_threshold)), status=status) for i in range(mb_size)]


class PaddedCube(torch.Tensor):
    """The padded cloud returned by pad_cube_boundaries: a float64 (N + n_img, 3) device tensor that remembers the
    threshold it was built with, so that get_pcube_csr can run the periodic kNN kernel on the original particles
    (images are regenerated on the fly from the same rule) instead of a KD-tree over the materialised cloud."""

    @staticmethod
    def wrap(t, n_particles, threshold):
        out = t.as_subclass(PaddedCube)
        out.n_particles, out.boundary_threshold = int(n_particles), float(threshold)
        return out


def pad_cube_boundaries(x, boundary_threshold):
    """graph.py:827-855 (+ face / edge / corner_outer, 801-825).  x (N, 3) float32 in the unit box ->
    (padded (N + n_img, 3) float64, idx_map (n_img,) int64): every particle with nb coordinates within the threshold
    of a wall gets 2^nb - 1 images, appended in particle order and, per particle, in the reference's pattern order.
    get_pbc_kneighbors_csr does NOT go through this (its kernel generates the same images on the fly)."""
    X = _to_cuda(x, torch.float32)
    if X.dim() != 2:
        raise ValueError("pad_cube_boundaries: x must be (N, 3) - one sample, as in the reference")
    padded, idx_map = ops.pad_cube(X, float(boundary_threshold))
    return PaddedCube.wrap(padded, X.shape[0], boundary_threshold), idx_map


def get_pcube_csr(x, idx_map, N, K, include_self=False):
    """graph.py:877-894: kNN graph of the first N rows of the padded cloud with image columns mapped back through
    idx_map; rows distance-sorted.  `x` must be the padded cloud returned by pad_cube_boundaries (it carries its
    threshold): the neighbours are computed by the periodic kernel on the N original particles, which is the same
    search - an image exists in the padded cloud iff the kernel generates it."""
    if not isinstance(x, PaddedCube) or x.n_particles != int(N):
        raise TypeError("get_pcube_csr: x must be the padded cloud returned by pad_cube_boundaries(x, thr) for the same N")
    thr = x.boundary_threshold
    Xc = x[:N].to(torch.float32).as_subclass(torch.Tensor).unsqueeze(0)          # exact: rows [0, N) are float32 values
    idx, _, status = ops.knn(Xc, int(K), True, thr, bool(include_self), ops._lib.ORDER_DISTANCE, False)
    return KnnCSR(idx, 0, int(x.shape[0]), status=status)


def get_pcube_adjacency_list(x, idx_map, N, K):
    """graph.py:857-874: (N, K) neighbour indices, self included (kneighbors_graph(..., include_self=True))."""
    return get_pcube_csr(x, idx_map, N, K, include_self=True).toarray_indices()


# =============================================================================== adjacency
class Adjacency:
    """COO (3,c) + diagonals + CSR transpose of a fixed-degree batch graph (c = b*N*M edges)."""

    def __init__(self, coo, diag, csrT_ptr, csrT_edge, status, b, N, M):
        self._coo, self._coo_ref = coo, None
        self.diag, self.csrT_ptr, self.csrT_edge, self.status = diag, csrT_ptr, csrT_edge, status
        self.b, self.N, self.M = b, N, M
        self._seg_cache = {}

    @property
    def coo(self):
        return self._coo if self._coo_ref is None else self._coo_ref()

    def weaken(self):
        """Called when this object is hung on its own COO tensor (`_attach`): keep only a weak reference back, or
        tensor and adjacency form a cycle that pins ~60 MB of device memory per step until the cyclic GC runs."""
        self._coo_ref, self._coo = weakref.ref(self._coo), None

    @property
    def col(self):
        return self.coo[1]

    def check(self):
        """Host-synchronising sanity check: one self edge per row, all indices in range."""
        bad_diag, bad_idx = self.status.tolist()
        if bad_idx:
            raise ValueError(f"adjacency: {bad_idx} neighbour indices out of range")
        return bad_diag == 0


def _attach(coo, adj):
    coo._nbpc_adjacency = adj
    adj.weaken()
    return coo


def _build_adjacency(A):
    idx = _as_batch(A)
    b, N, M = idx.shape
    coo, diag, ptr, edge, status = ops.adjacency(idx)
    return Adjacency(coo, diag, ptr, edge, status, b, N, M)


def to_coo_batch_ZA_diag(A):
    """graph.py:621-662 -> (COO_feats (3, b*N*M) int32, diagonals (b*N,) int64).
    `diagonals[n]` is the flat position of row n's self edge (-1 if a row has none; the reference
    would return a shorter array in that case)."""
    adj = _build_adjacency(A)
    return _attach(adj.coo, adj), adj.diag


def to_coo_batch(A):
    """graph.py:664-697"""
    return to_coo_batch_ZA_diag(A)[0]


def get_indices_from_list_CSR(A, offset=True):
    """graph.py:593-610"""
    idx = _as_batch(A)
    b, N, M = idx.shape
    off = (torch.arange(b, dtype=torch.int32, device=idx.device) * N).view(b, 1, 1)
    return (idx + off).reshape(-1)


def confirm_CSR_to_COO_index_integrity(A, COO_feats):
    """graph.py:612-618"""
    assert bool((get_indices_from_list_CSR(A) == _to_cuda(COO_feats[1], torch.int32)).all())


_ADJ_CACHE = {}


def _adjacency_of(COO_feats, b, N, need_cube=True):
    """need_cube: the caller pools per sample (cube pool, graph.py:447) with node // N, so the (b, N) factorisation
    must be the one the graph was built with (COO_feats[2] == COO_feats[0] // N); gathers only need b*N."""
    adj = getattr(COO_feats, "_nbpc_adjacency", None)
    if adj is not None and adj.b * adj.N == b * N:
        if need_cube and (adj.b, adj.N) != (b, N):
            raise ValueError(f"adjacency was built for (b, N) = ({adj.b}, {adj.N}) but the layer was called with ({b}, {N}): "
                             "the per-sample pool (COO_feats[2]) would differ from the reference's")
        return adj
    # plain (3,c) array/tensor: rebuild the CSR transpose (cached on identity of the storage)
    coo = _to_cuda(COO_feats, torch.int32).contiguous()
    key = (coo.data_ptr(), tuple(coo.shape), coo._version, b, N)
    adj = _ADJ_CACHE.get(key)
    if adj is None:
        c = coo.shape[1]
        if coo.shape[0] != 3 or c % (b * N) != 0:
            raise ValueError(f"COO_feats must be (3, b*N*M); got {tuple(coo.shape)} for b*N={b * N}")
        M = c // (b * N)
        rows_ok = bool((coo[0] == torch.arange(c, device=coo.device, dtype=torch.int32) // M).all())
        if not rows_ok:
            raise ValueError("COO_feats[0] must be the fixed-degree CSR row index e // M (kNN graph layout)")
        if need_cube and not bool((coo[2] == coo[0] // N).all()):
            raise ValueError("COO_feats[2] must be the sample id COO_feats[0] // N (graph.py:646)")
        ptr, edge, status = ops.segment_csr(coo[1], b * N)
        if int(status.item()):
            raise ValueError("COO_feats[1] has indices outside [0, b*N)")
        diag = torch.full((b * N,), -1, dtype=torch.int64, device=coo.device)
        adj = Adjacency(coo, diag, ptr, edge, torch.zeros(2, dtype=torch.int32, device=coo.device), b, N, M)
        if len(_ADJ_CACHE) > 8:
            _ADJ_CACHE.clear()
        _ADJ_CACHE[key] = adj
    return adj


# =============================================================================== input features
def include_node_features(X_in_edges, X_in_nodes, COO_feats, redshift=None):
    """graph.py:245-275: concat [edges, nodes[row], nodes[col], (redshift)] -> (c, 9|10)."""
    nodes = _to_cuda(X_in_nodes, torch.float32)
    adj = _adjacency_of(COO_feats, 1, nodes.shape[0], need_cube=False)
    rs = None if redshift is None else _to_cuda(redshift, torch.float32)
    return ops.include_node_features(_to_cuda(X_in_edges, torch.float32), nodes, adj.col, rs, adj.M)


def get_input_features_shift_inv_ZA(init_pos, ZA_displacement, coo, diag, dims):
    """graph.py:289-343: edges[e] = pos[col[e]] - pos[row[e]], ZA displacement added on the self edge."""
    b, N, M = dims
    pos = _to_cuda(init_pos, torch.float32).reshape(b * N, -1)
    za = _to_cuda(ZA_displacement, torch.float32).reshape(b * N, -1)
    adj = _adjacency_of(coo, b, N, need_cube=False)
    return ops.edge_features(pos, za, adj.col, _to_cuda(diag, torch.int64), M)


def get_input_features_shift_inv(X_in, coo, dims):
    """graph.py:346-364: raw (non minimum-image) relative positions (c,3) and the node features X[...,3:]."""
    b, N, M = dims
    X = _to_cuda(X_in, torch.float32).reshape(b * N, -1)
    adj = _adjacency_of(coo, b, N, need_cube=False)
    edges = ops.edge_features(X, None, adj.col, None, M)
    return edges, X[:, 3:]


# =============================================================================== layers
def shift_inv_conv(h, pool_idx, num_segs, broadcast):
    """graph.py:367-391: unsorted segment mean (+ gather back).  Deterministic: members of every
    segment are summed in ascending order."""
    h = _to_cuda(h, torch.float32)
    ids = _to_cuda(pool_idx, torch.int32).contiguous()
    ptr, members, _ = ops.segment_csr(ids, int(num_segs))
    return ops.SegmentPool.apply(h, ids, ptr, members, bool(broadcast))


def _is_relu(fn):
    return fn in (torch.relu, torch.nn.functional.relu) or getattr(fn, "__name__", "") == "relu"


def _layer(H_in, COO_feats, bN, layer_vars, is_last, relu, input_relu=False, grad_premasked=False, chain=None, idx=0):
    b, N = bN
    weights, B = layer_vars
    adj = _adjacency_of(COO_feats, b, N)
    W = weights if isinstance(weights, torch.Tensor) and weights.dim() == 3 else torch.stack(list(weights[:4]))
    return ops.GraphLayer.apply(H_in, W, B, adj.col, adj.csrT_ptr, adj.csrT_edge, b, N, adj.M, bool(is_last), relu,
                                input_relu, grad_premasked, chain, idx)


# Row-pool hand-over (csrc: glk3_edge_out_rowpool_kernel, glf_last_edge_in_rowsum_kernel): inside a fused-ReLU network the
# kernel that writes a hidden edge tensor also emits the row reduction the next layer (forward) / previous layer (backward)
# starts with, so that layer reads the tensor once.  Bit-identical to the plain path; NBPC_ROWPOOL_CHAIN=0 disables.
_ROWPOOL_CHAIN = _os.environ.get("NBPC_ROWPOOL_CHAIN", "1") != "0"


def set_rowpool_chain(on):
    """Opt out of / back into the row-pool hand-over between layers; returns the previous setting."""
    global _ROWPOOL_CHAIN
    old, _ROWPOOL_CHAIN = _ROWPOOL_CHAIN, bool(on)
    return old


def _rowpool_chain(widths, first, num_layers):
    """RowPoolChain for layers first..num_layers-1 of a fused-ReLU network with channel widths `widths` (None: nothing to hand over)."""
    chain = ops.RowPoolChain()
    for l in range(first, num_layers - 1):
        (k, q), (k2, q2) = widths[l], widths[l + 1]
        last2 = l + 1 == num_layers - 1
        if ops.graph_layer_rowpool_supported(k, q, False, ops.ROWPOOL_FWD_EMIT) and \
                ops.graph_layer_rowpool_supported(k2, q2, last2, ops.ROWPOOL_FWD_TAKE):
            chain.fwd_emit.add(l)
        if l >= 1 and ops.graph_layer_rowpool_supported(k2, q2, last2, ops.ROWPOOL_BWD_EMIT) and \
                ops.graph_layer_rowpool_supported(k, q, False, ops.ROWPOOL_BWD_TAKE):
            chain.bwd_emit.add(l + 1)
    return chain if (chain.fwd_emit or chain.bwd_emit) else None


def shift_inv_layer(H_in, COO_feats, bN, layer_vars, is_last=False):
    """graph.py:394-456.  H_in (c,k); layer_vars = ([W1..W4] each (k,q), B (q,)) -> (c,q), or (b,N,q) if is_last."""
    return _layer(_to_cuda(H_in, torch.float32), COO_feats, bN, layer_vars, is_last, False)


# Virtual first layer (csrc/graph_layer_vin.cuh): layer 1's (c, q) output is not materialised, layer 2's kernels recompute its
# rows.  It removes 44 % of the step's HBM bytes and is bit-identical to the materialised path, but measured SLOWER on B200
# (2.71 vs 2.00 ms per step at 8 x 32^3: four generator warps per CTA cannot keep enough L2 gathers in flight - the
# materialising kernel does the same gathers with 2048 threads per SM), so it is opt-in: NBPC_VIRTUAL_FIRST_LAYER=1 or
# set_virtual_first_layer(True).
_VIRTUAL_FIRST_LAYER = _os.environ.get("NBPC_VIRTUAL_FIRST_LAYER", "0") == "1"


def set_virtual_first_layer(on):
    """Opt in / out of the recomputing (virtual first layer) kernels; returns the previous setting."""
    global _VIRTUAL_FIRST_LAYER
    old, _VIRTUAL_FIRST_LAYER = _VIRTUAL_FIRST_LAYER, bool(on)
    return old


def _network(H0, coo, num_layers, dims, activation, model_vars):
    """Layer loop shared by the network functions.  A ReLU activation is fused into the layer kernel; any other
    callable is applied to the un-activated layer output."""
    fuse = _is_relu(activation)
    # inside this function every hidden tensor has exactly one consumer (the next layer), so the ReLU
    # backward of layer l is applied by layer l+1's edge kernel (input_relu) and layer l skips its own mask
    chain = fuse and num_layers > 1
    first = 1
    if fuse and num_layers >= 3 and not H0.requires_grad and _VIRTUAL_FIRST_LAYER:
        # the first layer's (c, q) output is never materialised: layer 1 computes its node-level terms only and layer 2
        # recomputes the rows it needs from the 12-byte edge features (ops.GraphLayerVirtualIn, csrc/graph_layer_vin.cuh)
        b, N = dims
        adj = _adjacency_of(coo, b, N)
        (W0, B0), (W1, B1) = model_vars.get_layer_vars(0), model_vars.get_layer_vars(1)
        W0 = W0 if isinstance(W0, torch.Tensor) and W0.dim() == 3 else torch.stack(list(W0[:4]))
        W1 = W1 if isinstance(W1, torch.Tensor) and W1.dim() == 3 else torch.stack(list(W1[:4]))
        k0, q0, q1 = W0.shape[1], W0.shape[2], W1.shape[2]
        if H0.shape[1] == k0 and W1.shape[1] == q0 and ops.graph_layer_vin_supported(k0, q0, q1, H0.shape[0]):
            Hv, Qc, Qr = ops.GraphLayerNodeOnly.apply(H0, W0, B0, adj.col, adj.csrT_ptr, adj.csrT_edge, b, N, adj.M)
            H = ops.GraphLayerVirtualIn.apply(Hv, W1, B1, adj.col, adj.csrT_ptr, adj.csrT_edge, b, N, adj.M, True, True, H0, W0[0], Qc, Qr)
            first = 2
    rp = None
    if chain and _ROWPOOL_CHAIN and first == 1:
        widths = []
        for l in range(num_layers):
            Wl = model_vars.get_layer_vars(l)[0]
            Wl = Wl if isinstance(Wl, torch.Tensor) else Wl[0]
            widths.append((int(Wl.shape[-2]), int(Wl.shape[-1])))
        if H0.shape[1] == widths[0][0] and all(widths[l][1] == widths[l + 1][0] for l in range(num_layers - 1)):
            rp = _rowpool_chain(widths, 0, num_layers)
    if first == 1:
        H = _layer(H0, coo, dims, model_vars.get_layer_vars(0), False, fuse, input_relu=False, grad_premasked=chain, chain=rp, idx=0)
        if not fuse:
            H = activation(H)
    for layer_idx in range(first, num_layers):
        is_last = layer_idx == num_layers - 1
        H = _layer(H, coo, dims, model_vars.get_layer_vars(layer_idx), is_last, fuse and not is_last,
                   input_relu=fuse, grad_premasked=fuse and not is_last, chain=rp, idx=layer_idx)
        if not is_last and not fuse:
            H = activation(H)
    return H


def network_func_shift_inv_za(edges, coo, num_layers, dims, activation, model_vars):
    """graph.py:463-476."""
    return _network(_to_cuda(edges, torch.float32), coo, num_layers, dims, activation, model_vars)


def network_func_shift_inv(X_in_edges, X_in_nodes, COO_feats, num_layers, dims, activation, model_vars, redshift=None):
    """graph.py:517-533 (the multi-redshift network; kept in a commented-out block by the reference): the input layer
    sees [relative position, velocity of the row node, velocity of the column node (, redshift)] = 9 | 10 channels."""
    H_in = include_node_features(X_in_edges, X_in_nodes, COO_feats, redshift=redshift)
    return _network(H_in, COO_feats, num_layers, dims, activation, model_vars)


def model_func_shift_inv(X_in, COO_feats, model_vars, dims, activation=torch.relu, redshift=None):
    """graph.py:536-567 (commented-out block): X_in (b,N,6) = [position, velocity] at one redshift -> the prediction
    at the next one, (b,N,6) (or (b,N,3) for a 3-channel network): loc' = net[:3]*loc_scalar + loc + vel*vel_scalar,
    vel' = net[3:]*vel_scalar + vel, with (loc_scalar, vel_scalar) = model_vars.get_scalars()."""
    num_layers = len(model_vars.channels) - 1
    X = _to_cuda(X_in, torch.float32)
    edges, nodes = get_input_features_shift_inv(X, COO_feats, dims)
    net_out = network_func_shift_inv(edges, nodes, COO_feats, num_layers, dims[:-1], activation, model_vars, redshift)
    loc_scalar, vel_scalar = model_vars.get_scalars()
    if (not torch.is_grad_enabled() or not (net_out.requires_grad or isinstance(loc_scalar, torch.Tensor))) and \
            not isinstance(loc_scalar, torch.Tensor) and not isinstance(vel_scalar, torch.Tensor) and net_out.shape[-1] in (3, 6):
        return ops.residual_update(X, net_out, loc_scalar, vel_scalar)      # inference / rollout: one kernel
    loc, vel = X[..., :3], X[..., 3:]
    H_out = net_out[..., :3] * loc_scalar + loc + vel * vel_scalar
    if net_out.shape[-1] > 3:
        H_out = torch.cat([H_out, net_out[..., 3:] * vel_scalar + vel], dim=-1)
    return H_out


# =============================================================================== 15-weight layer (graph.py:20-229)
class SymAdjacency(dict):
    """The `adj` dict of shift_inv_15op_layer (graph.py:46-59): row / col / all / tra (S,), dia / dal (b*N,) as int32
    device tensors, plus the CSR of every index list (built once, on demand) that makes the pooling forward and the
    gather backward deterministic segment sums."""

    def __init__(self, fields, b, N):
        super().__init__(fields)
        self.b, self.N = b, N
        self._csr = {}
        self.row_ptr = None     # set by get_symmetrized_adjacency: row-major sorted, symmetric, self edges present

    def csr(self, name, num_segs):
        key = (name, num_segs)
        if key not in self._csr:
            ptr, members, status = ops.segment_csr(self[name], int(num_segs))
            self._csr[key] = (ptr, members)
        return self._csr[key]


def get_symmetrized_adjacency(A):
    """The adjacency the 15-weight layer expects and for which the reference ships no builder: the
    SYMMETRISED kNN graph A u A^T of every sample, edges in row-major (row, col) order, with
      row, col : node ids of every edge (shifted by i*N across the batch),  all : sample id of every edge,
      tra      : position of the transposed edge (col, row),                dia : position of the self edge of every node,
      dal      : sample id of every node.
    The kNN lists must include the self edge (include_self=True).  Built by libnbpc kernels (csrc/graph15.cu): per node, a
    merge of its sorted out-neighbours with the rows of its in-edges (count -> scan -> emit), then a binary search per
    edge for the transposed position."""
    idx = _as_batch(A)
    b, N, M = idx.shape
    _, _, ptr, edge, _ = ops.adjacency(idx)
    fields, status = ops.sym_adjacency(idx, ptr, edge)
    no_self, no_partner = status.tolist()
    if no_self or no_partner:
        raise ValueError("get_symmetrized_adjacency: every node needs its self edge (build the kNN graph with include_self=True)")
    adj = SymAdjacency({k: v for k, v in fields.items() if k != "row_ptr"}, b, N)
    adj.row_ptr = fields["row_ptr"]         # canonical layout: enables the fused layer kernels
    return adj


def _as_sym(adj, b, N):
    if isinstance(adj, SymAdjacency):
        return adj
    return SymAdjacency({k: _to_cuda(v, torch.int32).contiguous().reshape(-1) for k, v in adj.items()}, b, N)


def shift_inv_15op_layer(H_in, adj, bN, layer_vars, is_last=False, _relu=False):
    """graph.py:20-200: the 15-weight permutation-equivariant basis on a symmetrised adjacency.  W (15, k, q), B (2, q);
    H_in (S, k) -> (S, q), or (b, N, q) if is_last (pooled over adj["row"]).
    Same sums as the reference, associated at node level: every pooled operand is projected once per NODE and then
    gathered to the edges (the reference projects after broadcasting), i.e.
        out[e] = H[e] W0 + H[tra[e]] W1 + T_col[col[e]] + T_row[row[e]] + T_all[all[e]] + [e is diagonal] T_dia[node]
        T_col = Hr W3 + Hc W7 + Hd W13          T_row = Hr W4 + Hc W6 + Hd W14          T_all = Ha W9 + Hp W11 + B[1]
        T_dia = Hd W2 + Hr W5 + Hc W8 + (Ha W10 + Hp W12)[dal] + B[0]
    with Hr / Hc / Ha = segment means of H over col / row / all, Hd = H[dia], Hp = segment mean of Hd over dal."""
    b, N = bN
    BN = b * N
    W, B = layer_vars
    W = W if isinstance(W, torch.Tensor) and W.dim() == 3 else torch.stack(list(W))
    B = B if isinstance(B, torch.Tensor) else torch.stack(list(B))
    H = _to_cuda(H_in, torch.float32)
    adj = _as_sym(adj, b, N)
    S = H.shape[0]
    if adj.row_ptr is not None and (adj.b, adj.N) == (b, N):
        # canonical adjacency (get_symmetrized_adjacency): fused node-level + edge-level kernels
        out = ops.Graph15Layer.apply(H, W, B, adj["row"], adj["col"], adj["tra"], adj["dia"], adj.row_ptr, b, N, bool(_relu))
        if is_last:
            key = ("_rows", BN)
            if key not in adj._csr:
                adj._csr[key] = (adj.row_ptr, torch.arange(S, dtype=torch.int32, device=H.device))
            ptr, mem = adj._csr[key]
            return ops.SegmentPool.apply(out, adj["row"], ptr, mem, False).reshape(b, N, -1)
        return out
    if _relu:
        raise ValueError("the fused ReLU needs the canonical adjacency of get_symmetrized_adjacency")
    lin = ops.Linear.apply

    def pool(h, name, nseg):
        ptr, mem = adj.csr(name, nseg)
        return ops.SegmentPool.apply(h, adj[name], ptr, mem, False)

    def gather(src, name):
        ptr, mem = adj.csr(name, src.shape[0])
        return ops.GatherRows.apply(src, adj[name], ptr, mem)

    Hr, Hc, Ha = pool(H, "col", BN), pool(H, "row", BN), pool(H, "all", b)
    Hd = gather(H, "dia")
    Hp = pool(Hd, "dal", b)
    T_col = lin(Hr, W[3]) + lin(Hc, W[7]) + lin(Hd, W[13])
    T_row = lin(Hr, W[4]) + lin(Hc, W[6]) + lin(Hd, W[14])
    T_all = lin(Ha, W[9]) + lin(Hp, W[11]) + B[1]
    T_dia = lin(Hd, W[2]) + lin(Hr, W[5]) + lin(Hc, W[8]) + gather(lin(Ha, W[10]) + lin(Hp, W[12]), "dal") + B[0]
    out = lin(H, W[0]) + lin(gather(H, "tra"), W[1]) + gather(T_col, "col") + gather(T_row, "row") + gather(T_all, "all")
    # broadcast to the diagonal (tf.scatter_nd, graph.py:106): edge e receives T_dia[i] iff e == dia[i]
    inv = torch.full((S,), BN, dtype=torch.int32, device=H.device)
    inv[adj["dia"].long()] = torch.arange(BN, dtype=torch.int32, device=H.device)
    key = ("_inv_dia", BN + 1)
    if key not in adj._csr:
        adj["_inv_dia"] = inv
        adj.csr("_inv_dia", BN + 1)
    out = out + gather(torch.cat([T_dia, torch.zeros_like(T_dia[:1])]), "_inv_dia")
    if is_last:
        return pool(out, "row", BN).reshape(b, N, -1)
    return out


def network_func_15op_shift_inv_za(edges, adj, num_layers, dims, activation, sess_mgr):
    """graph.py:202-216"""
    adj = _as_sym(adj, dims[0], dims[1])
    fuse = _is_relu(activation) and adj.row_ptr is not None      # ReLU inside the edge kernel (canonical adjacency only)
    H = shift_inv_15op_layer(edges, adj, dims, sess_mgr.get_layer_vars(0), _relu=fuse)
    if not fuse:
        H = activation(H)
    for layer_idx in range(1, num_layers):
        is_last = layer_idx == num_layers - 1
        H = shift_inv_15op_layer(H, adj, dims, sess_mgr.get_layer_vars(layer_idx), is_last=is_last, _relu=fuse and not is_last)
        if not is_last and not fuse:
            H = activation(H)
    return H


def model_func_15op_shift_inv_za(edges, adj_map, sess_mgr, dims, activation=torch.relu):
    """graph.py:219-229"""
    num_layers = len(sess_mgr.channels) - 1
    return network_func_15op_shift_inv_za(edges, adj_map, num_layers, dims[:-1], activation, sess_mgr)


def rollout_shift_inv(X0, model_vars_per_step, K, boundary_threshold, redshifts=None, activation=torch.relu,
                      include_self=False, trajectory=False):
    """Multi-redshift rollout (BASELINE config 5): for every step the periodic kNN graph is REBUILT on the
    current positions (get_pbc_kneighbors_csr, graph.py:896-917), the network predicts the next [position, velocity]
    (model_func_shift_inv) and nn.get_readout (nn.py:107-119) wraps the positions back into the unit box.
    model_vars_per_step: one model_vars per step (or a single one used for every step); redshifts: optional sequence
    of scalars appended as a 10th input channel.  Returns the final state (b,N,6), or all states if trajectory."""
    from . import nn as _nn
    X = _to_cuda(X0, torch.float32)
    b, N = X.shape[0], X.shape[1]
    steps = len(model_vars_per_step) if isinstance(model_vars_per_step, (list, tuple)) else (len(redshifts) if redshifts is not None else 1)
    states = []
    for i in range(steps):
        mv = model_vars_per_step[i] if isinstance(model_vars_per_step, (list, tuple)) else model_vars_per_step
        A = get_pbc_kneighbors_csr(X, K, boundary_threshold, include_self=include_self)
        coo = to_coo_batch(A)
        rs = None
        if redshifts is not None:
            rs = torch.full((b * N * K, 1), float(redshifts[i]), dtype=torch.float32, device=X.device)
        X = _nn.get_readout(model_func_shift_inv(X, coo, mv, (b, N, K), activation, rs))
        if trajectory:
            states.append(X)
    return states if trajectory else X


def model_func_shift_inv_za(init_pos, COO_feats, ZA_displacement, ZA_diagonal, model_vars, dims,
                            activation=torch.relu):
    """graph.py:479-515 -> predicted displacement error (b, N, q_last)."""
    num_layers = len(model_vars.channels) - 1
    edges = get_input_features_shift_inv_ZA(init_pos, ZA_displacement, COO_feats, ZA_diagonal, dims)
    return network_func_shift_inv_za(edges, COO_feats, num_layers, dims[:-1], activation, model_vars)
