"""PyTorch custom ops (`torch.ops.nbpc.*`) over the C ABI of libnbpc.so, plus the autograd
Functions that pair each forward with its hand-written backward.

torch is plumbing here: device memory, the current CUDA stream and autograd bookkeeping.  Every
op launches hand-written sm_100a kernels through ctypes; none has a CPU or eager-PyTorch
fallback (a CPU tensor raises).
"""
import ctypes
from typing import List, Optional, Tuple

import torch

from . import _lib

_vp = ctypes.c_void_p


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nbpc ops need CUDA tensors on an sm_100 (B200) device; there is no CPU fallback")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _i32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ================================================================== kNN
@torch.library.custom_op("nbpc::knn", mutates_args=())
def knn(xyz: torch.Tensor, k: int, periodic: bool, boundary_threshold: float, include_self: bool,
        order: int, want_d2: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """xyz (B,N,D>=3) float32 (any strides with unit stride on the last dim) -> idx (B,N,k) int32,
    d2 (B,N,k) float64 in distance order (empty if not requested), status int32[1] (periodic: particles with a
    coordinate outside the unit box, for which the result is unspecified)."""
    _need_cuda(xyz)
    L = _lib.load()
    if xyz.dtype != torch.float32:
        xyz = xyz.to(torch.float32)
    if xyz.dim() != 3 or xyz.shape[-1] < 3:
        raise RuntimeError("knn: xyz must be (B, N, D>=3)")
    if xyz.stride(-1) != 1:
        xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    idx = torch.empty((B, N, k), dtype=torch.int32, device=xyz.device)
    d2 = torch.empty((B, N, k) if want_d2 else (0,), dtype=torch.float64, device=xyz.device)
    status = torch.empty((1,), dtype=torch.int32, device=xyz.device)
    ws = _workspace(L.nbpc_knn_workspace_bytes(B, N, k, int(periodic)), xyz.device)
    with torch.cuda.device(xyz.device):
        rc = L.nbpc_knn(_ptr(xyz), xyz.stride(0), xyz.stride(1), B, N, k, int(periodic), float(boundary_threshold),
                        int(include_self), int(order), _ptr(idx), _ptr(d2) if want_d2 else None, _ptr(status), _ptr(ws),
                        ws.numel(), _stream())
    _lib.check(rc, "nbpc_knn")
    return idx, d2, status


def pad_cube(xyz: torch.Tensor, boundary_threshold: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """One sample xyz (N, D>=3) float32 -> padded cloud (N + n_img, 3) float64, idx_map (n_img,) int64
    (graph.py:827-855).  Host-synchronising: the number of images is read back to size the outputs."""
    _need_cuda(xyz)
    L = _lib.load()
    if xyz.dtype != torch.float32:
        xyz = xyz.to(torch.float32)
    if xyz.dim() != 2 or xyz.shape[-1] < 3:
        raise RuntimeError("pad_cube: xyz must be (N, D>=3)")
    if xyz.stride(-1) != 1:
        xyz = xyz.contiguous()
    N = xyz.shape[0]
    offsets = torch.empty((N + 1,), dtype=torch.int32, device=xyz.device)
    ws = _workspace(L.nbpc_pad_cube_workspace_bytes(N), xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(L.nbpc_pad_cube_count(_ptr(xyz), xyz.stride(0), N, float(boundary_threshold), _ptr(offsets), _ptr(ws),
                                         ws.numel(), _stream()), "nbpc_pad_cube_count")
        n_img = int(offsets[N].item())
        padded = torch.empty((N + n_img, 3), dtype=torch.float64, device=xyz.device)
        idx_map = torch.empty((max(n_img, 1),), dtype=torch.int64, device=xyz.device)
        _lib.check(L.nbpc_pad_cube_emit(_ptr(xyz), xyz.stride(0), N, float(boundary_threshold), _ptr(offsets), _ptr(padded),
                                        _ptr(idx_map), _stream()), "nbpc_pad_cube_emit")
    return padded, idx_map[:n_img]


# ================================================================== adjacency
@torch.library.custom_op("nbpc::adjacency", mutates_args=())
def adjacency(idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """idx (B,N,M) int32 local -> coo (3,c) int32, diag (B*N) int64, csrT_ptr (B*N+1), csrT_edge (c), status (2)."""
    _need_cuda(idx)
    L = _lib.load()
    idx = _i32c(idx)
    B, N, M = idx.shape
    c = B * N * M
    dev = idx.device
    coo = torch.empty((3, c), dtype=torch.int32, device=dev)
    diag = torch.empty((B * N,), dtype=torch.int64, device=dev)
    ptr = torch.empty((B * N + 1,), dtype=torch.int32, device=dev)
    edge = torch.empty((c,), dtype=torch.int32, device=dev)
    status = torch.empty((2,), dtype=torch.int32, device=dev)
    ws = _workspace(L.nbpc_adjacency_workspace_bytes(B, N, M), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_adjacency(_ptr(idx), B, N, M, _ptr(coo), _ptr(diag), _ptr(ptr), _ptr(edge), _ptr(status),
                              _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_adjacency")
    return coo, diag, ptr, edge, status


@torch.library.custom_op("nbpc::segment_csr", mutates_args=())
def segment_csr(ids: torch.Tensor, num_segs: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _need_cuda(ids)
    L = _lib.load()
    ids = _i32c(ids).reshape(-1)
    n = ids.numel()
    dev = ids.device
    ptr = torch.empty((num_segs + 1,), dtype=torch.int32, device=dev)
    mem = torch.empty((n,), dtype=torch.int32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    ws = _workspace(L.nbpc_segment_csr_workspace_bytes(n, num_segs), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_segment_csr(_ptr(ids), n, num_segs, _ptr(ptr), _ptr(mem), _ptr(status), _ptr(ws), ws.numel(),
                                _stream())
    _lib.check(rc, "nbpc_segment_csr")
    return ptr, mem, status


# ================================================================== input features
@torch.library.custom_op("nbpc::edge_features", mutates_args=())
def edge_features(pos: torch.Tensor, za: Optional[torch.Tensor], col: torch.Tensor,
                  diag: Optional[torch.Tensor], M: int) -> torch.Tensor:
    """pos (BN, ld>=3), za (BN, ld>=3)|None, col (c) int32, diag (n) int64|None -> edges (c,3)."""
    _need_cuda(pos, za, col, diag)
    L = _lib.load()
    pos = _f32c(pos)
    BN = pos.shape[0]
    col = _i32c(col)
    out = torch.empty((BN * M, 3), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        if za is None:
            rc = L.nbpc_edge_features(_ptr(pos), pos.shape[1], _ptr(col), BN, M, _ptr(out), _stream())
        else:
            za = _f32c(za)
            diag = diag.to(torch.int64).contiguous()
            rc = L.nbpc_edge_features_za(_ptr(pos), pos.shape[1], _ptr(za), za.shape[1], _ptr(col), _ptr(diag),
                                         diag.numel(), BN, M, _ptr(out), _stream())
    _lib.check(rc, "nbpc_edge_features")
    return out


@torch.library.custom_op("nbpc::include_node_features", mutates_args=())
def include_node_features(edges: torch.Tensor, nodes: torch.Tensor, col: torch.Tensor,
                          redshift: Optional[torch.Tensor], M: int) -> torch.Tensor:
    _need_cuda(edges, nodes, col, redshift)
    L = _lib.load()
    edges, nodes, col = _f32c(edges), _f32c(nodes), _i32c(col)
    c, E = edges.shape
    BN, F = nodes.shape
    if redshift is not None:
        redshift = _f32c(redshift).reshape(-1)
    out = torch.empty((c, E + 2 * F + (1 if redshift is not None else 0)), dtype=torch.float32, device=edges.device)
    with torch.cuda.device(edges.device):
        rc = L.nbpc_include_node_features(_ptr(edges), E, _ptr(nodes), F, F, _ptr(col), _ptr(redshift), BN, M,
                                          _ptr(out), _stream())
    _lib.check(rc, "nbpc_include_node_features")
    return out


# ================================================================== pooling primitive
@torch.library.custom_op("nbpc::segment_reduce", mutates_args=())
def segment_reduce(h: torch.Tensor, seg_ptr: torch.Tensor, seg_members: torch.Tensor, mean: bool) -> torch.Tensor:
    _need_cuda(h, seg_ptr, seg_members)
    L = _lib.load()
    h = _f32c(h)
    k = h.shape[1]
    num_segs = seg_ptr.numel() - 1
    out = torch.empty((num_segs, k), dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        rc = L.nbpc_segment_reduce(_ptr(h), k, _ptr(seg_ptr), _ptr(seg_members), num_segs, int(mean), _ptr(out),
                                   _stream())
    _lib.check(rc, "nbpc_segment_reduce")
    return out


@torch.library.custom_op("nbpc::gather_rows", mutates_args=())
def gather_rows(src: torch.Tensor, ids: torch.Tensor, seg_ptr: Optional[torch.Tensor]) -> torch.Tensor:
    _need_cuda(src, ids, seg_ptr)
    L = _lib.load()
    src, ids = _f32c(src), _i32c(ids).reshape(-1)
    k = src.shape[1]
    out = torch.empty((ids.numel(), k), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        rc = L.nbpc_gather_rows(_ptr(src), k, _ptr(ids), ids.numel(), _ptr(seg_ptr), _ptr(out), _stream())
    _lib.check(rc, "nbpc_gather_rows")
    return out


class SegmentPool(torch.autograd.Function):
    """shift_inv_conv (graph.py:367-391) with a deterministic backward."""

    @staticmethod
    def forward(ctx, h, ids, seg_ptr, seg_members, broadcast):
        pooled = segment_reduce(h, seg_ptr, seg_members, True)
        ctx.save_for_backward(ids, seg_ptr, seg_members)
        ctx.broadcast = broadcast
        return gather_rows(pooled, ids, None) if broadcast else pooled

    @staticmethod
    def backward(ctx, g):
        ids, seg_ptr, seg_members = ctx.saved_tensors
        g = g.contiguous()
        if ctx.broadcast:  # adjoint of the gather = segment sum
            g = segment_reduce(g, seg_ptr, seg_members, False)
        # adjoint of the mean = gather of g / max(count, 1)
        return gather_rows(g, ids, seg_ptr), None, None, None, None


class GatherRows(torch.autograd.Function):
    """y = src[ids] (tf.gather / tf.gather_nd on rows).  Backward = deterministic segment sum over the CSR of ids
    (members of a row in ascending order), not an atomic scatter."""

    @staticmethod
    def forward(ctx, src, ids, seg_ptr, seg_members):
        ctx.save_for_backward(seg_ptr, seg_members)
        return gather_rows(src, ids, None)

    @staticmethod
    def backward(ctx, g):
        seg_ptr, seg_members = ctx.saved_tensors
        return segment_reduce(g.contiguous(), seg_ptr, seg_members, False), None, None, None


# ================================================================== dense projection (15-weight layer)
@torch.library.custom_op("nbpc::linear", mutates_args=())
def linear(X: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], transpose_w: bool) -> torch.Tensor:
    """Y = X W (+ bias), W (k,q); with transpose_w W is (q,k) and Y = X W^T."""
    _need_cuda(X, W, bias)
    L = _lib.load()
    X, W = _f32c(X), _f32c(W)
    n, k = X.shape
    q = W.shape[0] if transpose_w else W.shape[1]
    if (W.shape[1] if transpose_w else W.shape[0]) != k:
        raise RuntimeError(f"linear: X is {tuple(X.shape)} but W is {tuple(W.shape)} (transpose_w={transpose_w})")
    Y = torch.empty((n, q), dtype=torch.float32, device=X.device)
    b = None if bias is None else _f32c(bias)
    with torch.cuda.device(X.device):
        rc = L.nbpc_linear(_ptr(X), _ptr(W), _ptr(b), n, k, q, int(transpose_w), 0, _ptr(Y), _stream())
    _lib.check(rc, "nbpc_linear")
    return Y


@torch.library.custom_op("nbpc::xty", mutates_args=())
def xty(X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """X^T Y over the rows, fixed summation order (bit-reproducible)."""
    _need_cuda(X, Y)
    L = _lib.load()
    X, Y = _f32c(X), _f32c(Y)
    n, k = X.shape
    q = Y.shape[1]
    out = torch.zeros((k, q), dtype=torch.float32, device=X.device)
    if n == 0:
        return out
    ws = _workspace(L.nbpc_xty_workspace_bytes(n, k, q), X.device)
    with torch.cuda.device(X.device):
        rc = L.nbpc_xty(_ptr(X), _ptr(Y), n, k, q, _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_xty")
    return out


class Linear(torch.autograd.Function):
    """tf.matmul(X, W) (graph.py:147-189) with dX = dY W^T and the deterministic dW = X^T dY."""

    @staticmethod
    def forward(ctx, X, W):
        ctx.save_for_backward(X, W)
        return linear(X, W, None, False)

    @staticmethod
    def backward(ctx, g):
        X, W = ctx.saved_tensors
        g = g.contiguous()
        dX = linear(g, W, None, True) if ctx.needs_input_grad[0] else None
        dW = xty(X, g) if ctx.needs_input_grad[1] else None
        return dX, dW


# ================================================================== graph layer
@torch.library.custom_op("nbpc::graph_layer_fwd", mutates_args=())
def graph_layer_fwd(H_in: torch.Tensor, col: torch.Tensor, csrT_ptr: torch.Tensor, csrT_edge: torch.Tensor,
                    W: torch.Tensor, bias: torch.Tensor, B: int, N: int, M: int, is_last: bool,
                    relu: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> H_out (c,q) | (B*N,q), P_col (BN,k), P_row (BN,k), P_cube (B,k)."""
    _need_cuda(H_in, col, csrT_ptr, csrT_edge, W, bias)
    L = _lib.load()
    H_in, W, bias = _f32c(H_in), _f32c(W), _f32c(bias)
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    if H_in.shape != (c, k) or W.shape[0] != 4 or bias.shape != (q,):
        raise RuntimeError(f"graph_layer_fwd: shape mismatch H_in {tuple(H_in.shape)} vs (c={c}, k={k}); W {tuple(W.shape)}")
    dev = H_in.device
    out = torch.empty(((B * N) if is_last else c, q), dtype=torch.float32, device=dev)
    P_col = torch.empty((B * N, k), dtype=torch.float32, device=dev)
    P_row = torch.empty((B * N, k), dtype=torch.float32, device=dev)
    P_cube = torch.empty((B, k), dtype=torch.float32, device=dev)
    ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_graph_layer_fwd(_ptr(H_in), _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge), B, N, M, k, q, _ptr(W),
                                    _ptr(bias), int(is_last), int(relu), _ptr(out), _ptr(P_col), _ptr(P_row),
                                    _ptr(P_cube), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_graph_layer_fwd")
    return out, P_col, P_row, P_cube


@torch.library.custom_op("nbpc::graph_layer_bwd", mutates_args=())
def graph_layer_bwd(dOut: torch.Tensor, H_in: torch.Tensor, H_out: torch.Tensor, col: torch.Tensor,
                    csrT_ptr: torch.Tensor, csrT_edge: torch.Tensor, W: torch.Tensor, P_col: torch.Tensor,
                    P_row: torch.Tensor, P_cube: torch.Tensor, B: int, N: int, M: int, is_last: bool, relu: bool,
                    mask_input: bool, need_dH: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> dH_in (c,k) (empty if not needed), dW (4,k,q), dB (q).
    relu: mask dOut by [H_out > 0]; mask_input: multiply dH_in by [H_in > 0] (fused ReLU backward of the
    layer that produced H_in)."""
    _need_cuda(dOut, H_in, H_out, W)
    L = _lib.load()
    dOut = _f32c(dOut)
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    dev = H_in.device
    dH = torch.empty((c, k) if need_dH else (0,), dtype=torch.float32, device=dev)
    dW = torch.empty((4, k, q), dtype=torch.float32, device=dev)
    dB = torch.empty((q,), dtype=torch.float32, device=dev)
    ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_graph_layer_bwd(_ptr(dOut), _ptr(H_in), _ptr(H_out), _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge),
                                    B, N, M, k, q, _ptr(W), _ptr(P_col), _ptr(P_row), _ptr(P_cube), int(is_last),
                                    int(relu), int(mask_input), _ptr(dH) if need_dH else None, _ptr(dW), _ptr(dB),
                                    _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_graph_layer_bwd")
    return dH, dW, dB


# ---- row-pool hand-over between consecutive layers (include/nbpc.h: nbpc_graph_layer_fwd_rp / _bwd_rp)
ROWPOOL_FWD_EMIT, ROWPOOL_FWD_TAKE, ROWPOOL_BWD_EMIT, ROWPOOL_BWD_TAKE = 0, 1, 2, 3


def graph_layer_rowpool_supported(k: int, q: int, is_last: bool, direction: int) -> bool:
    return bool(_lib.load().nbpc_graph_layer_rowpool_supported(int(k), int(q), int(is_last), int(direction)))


class RowPoolChain:
    """Which layers of a network hand their row reductions to the neighbouring layer (graph._network fills `fwd_emit` /
    `bwd_emit` with layer indices), and the tensors in flight: p_row[l] = row means of layer l's INPUT written by layer
    l-1's edge kernel, dq_row[l] = row sums of layer l's dOut written by layer l+1's backward edge kernel."""

    def __init__(self):
        self.fwd_emit, self.bwd_emit = set(), set()
        self.p_row, self.dq_row = {}, {}


@torch.library.custom_op("nbpc::graph_layer_fwd_rp", mutates_args=())
def graph_layer_fwd_rp(H_in: torch.Tensor, col: torch.Tensor, csrT_ptr: torch.Tensor, csrT_edge: torch.Tensor,
                       W: torch.Tensor, bias: torch.Tensor, B: int, N: int, M: int, is_last: bool, relu: bool,
                       P_row_given: Optional[torch.Tensor], emit_next: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor,
                                                                                     torch.Tensor, torch.Tensor]:
    """graph_layer_fwd with the hand-over: -> H_out, P_col, P_row (empty if P_row_given), P_cube, P_row_next (B*N,q) (empty
    unless emit_next)."""
    _need_cuda(H_in, col, csrT_ptr, csrT_edge, W, bias)
    L = _lib.load()
    H_in, W, bias = _f32c(H_in), _f32c(W), _f32c(bias)
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    if H_in.shape != (c, k) or W.shape[0] != 4 or bias.shape != (q,):
        raise RuntimeError(f"graph_layer_fwd_rp: shape mismatch H_in {tuple(H_in.shape)} vs (c={c}, k={k}); W {tuple(W.shape)}")
    if P_row_given is not None and (P_row_given.shape != (B * N, k) or P_row_given.dtype != torch.float32 or not P_row_given.is_contiguous()):
        raise RuntimeError("graph_layer_fwd_rp: P_row_given must be a contiguous float32 (B*N, k) tensor")
    dev = H_in.device
    out = torch.empty(((B * N) if is_last else c, q), dtype=torch.float32, device=dev)
    P_col = torch.empty((B * N, k), dtype=torch.float32, device=dev)
    P_row = torch.empty((B * N, k) if P_row_given is None else (0,), dtype=torch.float32, device=dev)
    P_cube = torch.empty((B, k), dtype=torch.float32, device=dev)
    P_next = torch.empty((B * N, q) if emit_next else (0,), dtype=torch.float32, device=dev)
    ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_graph_layer_fwd_rp(_ptr(H_in), _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge), B, N, M, k, q, _ptr(W),
                                       _ptr(bias), int(is_last), int(relu), _ptr(out), _ptr(P_col),
                                       _ptr(P_row if P_row_given is None else P_row_given), _ptr(P_cube),
                                       int(P_row_given is not None), _ptr(P_next) if emit_next else None, _ptr(ws), ws.numel(),
                                       _stream())
    _lib.check(rc, "nbpc_graph_layer_fwd_rp")
    return out, P_col, P_row, P_cube, P_next


@torch.library.custom_op("nbpc::graph_layer_bwd_rp", mutates_args=())
def graph_layer_bwd_rp(dOut: torch.Tensor, H_in: torch.Tensor, H_out: torch.Tensor, col: torch.Tensor,
                       csrT_ptr: torch.Tensor, csrT_edge: torch.Tensor, W: torch.Tensor, P_col: torch.Tensor,
                       P_row: torch.Tensor, P_cube: torch.Tensor, B: int, N: int, M: int, is_last: bool, relu: bool,
                       mask_input: bool, need_dH: bool, dQ_row_given: Optional[torch.Tensor],
                       emit_prev: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """graph_layer_bwd with the hand-over: -> dH_in, dW, dB, dQ_row_prev (B*N,k) = row sums of dH_in (empty unless emit_prev)."""
    _need_cuda(dOut, H_in, H_out, W)
    L = _lib.load()
    dOut = _f32c(dOut)
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    dev = H_in.device
    if dQ_row_given is not None and (dQ_row_given.shape != (B * N, q) or dQ_row_given.dtype != torch.float32 or not dQ_row_given.is_contiguous()):
        raise RuntimeError("graph_layer_bwd_rp: dQ_row_given must be a contiguous float32 (B*N, q) tensor")
    dH = torch.empty((c, k) if need_dH else (0,), dtype=torch.float32, device=dev)
    dW = torch.empty((4, k, q), dtype=torch.float32, device=dev)
    dB = torch.empty((q,), dtype=torch.float32, device=dev)
    dQp = torch.empty((B * N, k) if emit_prev else (0,), dtype=torch.float32, device=dev)
    ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_graph_layer_bwd_rp(_ptr(dOut), _ptr(H_in), _ptr(H_out), _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge),
                                       B, N, M, k, q, _ptr(W), _ptr(P_col), _ptr(P_row), _ptr(P_cube), int(is_last),
                                       int(relu), int(mask_input), _ptr(dH) if need_dH else None, _ptr(dW), _ptr(dB),
                                       _ptr(dQ_row_given), _ptr(dQp) if emit_prev else None, _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_graph_layer_bwd_rp")
    return dH, dW, dB, dQp


def _vin_struct(vin):
    E, W1, Qc, Qr = vin
    return _lib.VirtualInput(E.data_ptr(), W1.data_ptr(), Qc.data_ptr(), Qr.data_ptr(), int(E.shape[1]))


def graph_layer_vin_supported(k0: int, k: int, q: int, c: int) -> bool:
    """Whether the next layer can consume a (k0 -> k) first layer's output as a VIRTUAL input (include/nbpc.h)."""
    return bool(_lib.load().nbpc_graph_layer_vin_supported(int(k0), int(k), int(q), int(c)))


class GraphLayerNodeOnly(torch.autograd.Function):
    """First graph layer of a network whose output is consumed as a VIRTUAL input by the next layer: only the node-level
    terms are computed (pooling of the edge features, Q_col, Q_row); the (c, q) edge tensor is never written.  Returns a
    zero-stride placeholder of the output's shape (it carries the gradient dH1 back) plus Q_col, Q_row."""

    @staticmethod
    def forward(ctx, E, W, bias, col, csrT_ptr, csrT_edge, B, N, M):
        _need_cuda(E, W, bias, col)
        L = _lib.load()
        E, W, bias = _f32c(E), _f32c(W), _f32c(bias)
        k, q = W.shape[1], W.shape[2]
        c, BN, dev = B * N * M, B * N, E.device
        P_col = torch.empty((BN, k), dtype=torch.float32, device=dev)
        P_row = torch.empty((BN, k), dtype=torch.float32, device=dev)
        P_cube = torch.empty((B, k), dtype=torch.float32, device=dev)
        Qc = torch.empty((BN, q), dtype=torch.float32, device=dev)
        Qr = torch.empty((BN, q), dtype=torch.float32, device=dev)
        ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
        with torch.cuda.device(dev):
            rc = L.nbpc_graph_layer_fwd_v(_ptr(E), None, _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge), B, N, M, k, q, _ptr(W), _ptr(bias), 0, 1,
                                          1, None, _ptr(P_col), _ptr(P_row), _ptr(P_cube), _ptr(Qc), _ptr(Qr), _ptr(ws), ws.numel(),
                                          _stream())
        _lib.check(rc, "nbpc_graph_layer_fwd_v")
        ctx.save_for_backward(E, W, col, csrT_ptr, csrT_edge, P_col, P_row, P_cube)
        ctx.cfg = (B, N, M)
        placeholder = torch.zeros(1, dtype=torch.float32, device=dev).expand(c, q)
        ctx.mark_non_differentiable(Qc, Qr)
        return placeholder, Qc, Qr

    @staticmethod
    def backward(ctx, g, _gqc, _gqr):
        E, W, col, csrT_ptr, csrT_edge, P_col, P_row, P_cube = ctx.saved_tensors
        B, N, M = ctx.cfg
        # g = dH1, already masked by the consumer (which recomputed H1): plain first-layer backward, no input gradient
        _, dW, dB = graph_layer_bwd(g.contiguous(), E, g, col, csrT_ptr, csrT_edge, W, P_col, P_row, P_cube, B, N, M, False, False,
                                    False, False)
        return None, dW, dB, None, None, None, None, None, None


class GraphLayerVirtualIn(torch.autograd.Function):
    """Hidden graph layer whose input is the virtual tensor of GraphLayerNodeOnly: pooling, forward and backward edge
    kernels recompute H1[e] = relu(E[e] W1 + Q_col[col[e]] + Q_row[e / M]) (csrc/graph_layer_vin.cuh)."""

    @staticmethod
    def forward(ctx, H_placeholder, W, bias, col, csrT_ptr, csrT_edge, B, N, M, relu, grad_premasked, E, W1a, Qc1, Qr1):
        L = _lib.load()
        W, bias = _f32c(W), _f32c(bias)
        vin = (_f32c(E), _f32c(W1a), Qc1, Qr1)
        k, q = W.shape[1], W.shape[2]
        c, BN, dev = B * N * M, B * N, W.device
        out = torch.empty((c, q), dtype=torch.float32, device=dev)
        P_col = torch.empty((BN, k), dtype=torch.float32, device=dev)
        P_row = torch.empty((BN, k), dtype=torch.float32, device=dev)
        P_cube = torch.empty((B, k), dtype=torch.float32, device=dev)
        ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
        vs = _vin_struct(vin)
        with torch.cuda.device(dev):
            rc = L.nbpc_graph_layer_fwd_v(None, ctypes.byref(vs), _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge), B, N, M, k, q, _ptr(W),
                                          _ptr(bias), 0, int(relu), 0, _ptr(out), _ptr(P_col), _ptr(P_row), _ptr(P_cube), None, None,
                                          _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "nbpc_graph_layer_fwd_v")
        ctx.save_for_backward(out, W, col, csrT_ptr, csrT_edge, P_col, P_row, P_cube, *vin)
        ctx.cfg = (B, N, M, relu and not grad_premasked)
        return out

    @staticmethod
    def backward(ctx, g):
        out, W, col, csrT_ptr, csrT_edge, P_col, P_row, P_cube, E, W1a, Qc1, Qr1 = ctx.saved_tensors
        B, N, M, relu = ctx.cfg
        if relu:
            raise RuntimeError("virtual-input layer: the gradient must arrive pre-masked (graph._network chains the layers)")
        L = _lib.load()
        g = _f32c(g)
        k, q = W.shape[1], W.shape[2]
        c, dev = B * N * M, W.device
        dH = torch.empty((c, k), dtype=torch.float32, device=dev)
        dW = torch.empty((4, k, q), dtype=torch.float32, device=dev)
        dB = torch.empty((q,), dtype=torch.float32, device=dev)
        ws = _workspace(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q), dev)
        vs = _vin_struct((E, W1a, Qc1, Qr1))
        with torch.cuda.device(dev):
            rc = L.nbpc_graph_layer_bwd_v(_ptr(g), None, ctypes.byref(vs), _ptr(out), _ptr(col), _ptr(csrT_ptr), _ptr(csrT_edge), B, N, M, k,
                                          q, _ptr(W), _ptr(P_col), _ptr(P_row), _ptr(P_cube), 0, 0, 1, _ptr(dH), _ptr(dW), _ptr(dB),
                                          _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "nbpc_graph_layer_bwd_v")
        return dH, dW, dB, None, None, None, None, None, None, None, None, None, None, None, None


class GraphLayer(torch.autograd.Function):
    """shift_inv_layer (graph.py:394-456) [+ fused ReLU], backward through the CSR transpose.

    input_relu: H_in is the output of a ReLU-fused layer whose ONLY consumer is this layer; the ReLU
      backward of that layer is then applied to dH_in inside this layer's edge kernel (H_in is on chip).
    grad_premasked: the converse - this layer fused a ReLU and its only consumer applies the mask, so
      backward does not re-read H_out.  graph.network_func_shift_inv_za sets both consistently."""

    @staticmethod
    def forward(ctx, H_in, W, bias, col, csrT_ptr, csrT_edge, B, N, M, is_last, relu, input_relu=False,
                grad_premasked=False, chain=None, idx=0):
        """chain / idx: row-pool hand-over inside a network (RowPoolChain; this layer is layer `idx`)."""
        H_in = _f32c(H_in)
        p_given = chain.p_row.pop(idx, None) if chain is not None else None
        emit = chain is not None and idx in chain.fwd_emit
        if p_given is not None or emit:
            out, P_col, P_row, P_cube, P_next = graph_layer_fwd_rp(H_in, col, csrT_ptr, csrT_edge, W, bias, B, N, M, is_last, relu,
                                                                   p_given, emit)
            if p_given is not None:
                P_row = p_given
            if emit:
                chain.p_row[idx + 1] = P_next
        else:
            out, P_col, P_row, P_cube = graph_layer_fwd(H_in, col, csrT_ptr, csrT_edge, W, bias, B, N, M, is_last, relu)
        ctx.save_for_backward(H_in, out, W, col, csrT_ptr, csrT_edge, P_col, P_row, P_cube)
        ctx.cfg = (B, N, M, is_last, relu and not grad_premasked, input_relu)
        ctx.chain = (chain, idx)
        if is_last:
            out = out.view(B, N, -1)
        return out

    @staticmethod
    def backward(ctx, g):
        H_in, out, W, col, csrT_ptr, csrT_edge, P_col, P_row, P_cube = ctx.saved_tensors
        B, N, M, is_last, relu, input_relu = ctx.cfg
        chain, idx = ctx.chain
        need_dH = ctx.needs_input_grad[0]
        g = g.reshape(out.shape)
        dq_given = chain.dq_row.pop(idx, None) if chain is not None else None
        emit = chain is not None and need_dH and idx in chain.bwd_emit
        if dq_given is not None or emit:
            dH, dW, dB, dQp = graph_layer_bwd_rp(g, H_in, out, col, csrT_ptr, csrT_edge, W, P_col, P_row, P_cube, B, N, M,
                                                 is_last, relu, input_relu, need_dH, dq_given, emit)
            if emit:
                chain.dq_row[idx - 1] = dQp
        else:
            dH, dW, dB = graph_layer_bwd(g, H_in, out, col, csrT_ptr, csrT_edge, W, P_col, P_row, P_cube, B, N, M,
                                         is_last, relu, input_relu, need_dH)
        return (dH if need_dH else None), dW, dB, None, None, None, None, None, None, None, None, None, None, None, None


# ================================================================== 15-weight layer (graph.py:20-200)
def sym_adjacency(idx: torch.Tensor, csrT_ptr: torch.Tensor, csrT_edge: torch.Tensor):
    """kNN lists idx (B,N,M) + their in-edge lists -> the canonical symmetrised adjacency: dict of int32 device tensors
    row, col, all, tra (S), dia, dal (B*N), row_ptr (B*N+1), and status int32[2].  Host-synchronising (S is read back)."""
    _need_cuda(idx, csrT_ptr, csrT_edge)
    L = _lib.load()
    idx = _i32c(idx)
    B, N, M = idx.shape
    BN, dev = B * N, idx.device
    row_ptr = torch.empty((BN + 1,), dtype=torch.int32, device=dev)
    ws = _workspace(L.nbpc_sym_adjacency_workspace_bytes(B, N), dev)
    with torch.cuda.device(dev):
        _lib.check(L.nbpc_sym_adjacency_count(_ptr(idx), _ptr(csrT_ptr), _ptr(csrT_edge), B, N, M, _ptr(row_ptr), _ptr(ws), ws.numel(),
                                              _stream()), "nbpc_sym_adjacency_count")
        S = int(row_ptr[BN].item())
        out = {n: torch.empty((S,), dtype=torch.int32, device=dev) for n in ("row", "col", "all", "tra")}
        out.update({n: torch.empty((BN,), dtype=torch.int32, device=dev) for n in ("dia", "dal")})
        status = torch.empty((2,), dtype=torch.int32, device=dev)
        _lib.check(L.nbpc_sym_adjacency_emit(_ptr(idx), _ptr(csrT_ptr), _ptr(csrT_edge), _ptr(row_ptr), B, N, M, S, _ptr(out["row"]),
                                             _ptr(out["col"]), _ptr(out["all"]), _ptr(out["tra"]), _ptr(out["dia"]), _ptr(out["dal"]),
                                             _ptr(status), _stream()), "nbpc_sym_adjacency_emit")
    out["row_ptr"] = row_ptr
    return out, status


class Graph15Layer(torch.autograd.Function):
    """shift_inv_15op_layer (graph.py:20-200) [+ fused ReLU] on the canonical symmetrised adjacency: node-level pooling and
    projections + ONE edge kernel per direction (csrc/graph15.cu); deterministic backward."""

    @staticmethod
    def forward(ctx, H, W, Bias, row, col, tra, dia, row_ptr, B, N, relu):
        _need_cuda(H, W, Bias, row, col, tra, dia, row_ptr)
        L = _lib.load()
        H, W, Bias = _f32c(H), _f32c(W), _f32c(Bias)
        S, k = H.shape
        q = W.shape[2]
        if W.shape[0] != 15 or W.shape[1] != k or Bias.shape != (2, q) or row.numel() != S:
            raise RuntimeError(f"graph15: shape mismatch H {tuple(H.shape)}, W {tuple(W.shape)}, B {tuple(Bias.shape)}, S {row.numel()}")
        dev, BN = H.device, B * N
        out = torch.empty((S, q), dtype=torch.float32, device=dev)
        Hr, Hc, Hd = (torch.empty((BN, k), dtype=torch.float32, device=dev) for _ in range(3))
        Ha, Hp = (torch.empty((B, k), dtype=torch.float32, device=dev) for _ in range(2))
        ws = _workspace(L.nbpc_graph15_workspace_bytes(B, N, S, k, q), dev)
        with torch.cuda.device(dev):
            rc = L.nbpc_graph15_layer_fwd(_ptr(H), _ptr(row), _ptr(col), _ptr(tra), _ptr(dia), _ptr(row_ptr), B, N, S, k, q, _ptr(W),
                                          _ptr(Bias), int(relu), _ptr(out), _ptr(Hr), _ptr(Hc), _ptr(Hd), _ptr(Ha), _ptr(Hp), _ptr(ws),
                                          ws.numel(), _stream())
        _lib.check(rc, "nbpc_graph15_layer_fwd")
        ctx.save_for_backward(H, out, W, row, col, tra, dia, row_ptr, Hr, Hc, Hd, Ha, Hp)
        ctx.cfg = (B, N, relu)
        return out

    @staticmethod
    def backward(ctx, g):
        H, out, W, row, col, tra, dia, row_ptr, Hr, Hc, Hd, Ha, Hp = ctx.saved_tensors
        B, N, relu = ctx.cfg
        L = _lib.load()
        g = _f32c(g)
        S, k = H.shape
        q = W.shape[2]
        dev = H.device
        need_dH = ctx.needs_input_grad[0]
        dH = torch.empty((S, k), dtype=torch.float32, device=dev) if need_dH else None
        dW = torch.empty((15, k, q), dtype=torch.float32, device=dev)
        dB = torch.empty((2, q), dtype=torch.float32, device=dev)
        ws = _workspace(L.nbpc_graph15_workspace_bytes(B, N, S, k, q), dev)
        with torch.cuda.device(dev):
            rc = L.nbpc_graph15_layer_bwd(_ptr(g), _ptr(H), _ptr(out), _ptr(row), _ptr(col), _ptr(tra), _ptr(dia), _ptr(row_ptr), B, N, S,
                                          k, q, _ptr(W), _ptr(Hr), _ptr(Hc), _ptr(Hd), _ptr(Ha), _ptr(Hp), int(relu), _ptr(dH), _ptr(dW),
                                          _ptr(dB), _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "nbpc_graph15_layer_bwd")
        return dH, dW, dB, None, None, None, None, None, None, None, None


# ================================================================== set layer
@torch.library.custom_op("nbpc::set_layer_fwd", mutates_args=())
def set_layer_fwd(H_in: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, relu: bool, mu_in: Optional[torch.Tensor] = None,
                  want_mean_out: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (H_out, mu, mean_out).  mu_in: column means of H_in handed over by the layer that wrote H_in (skips the mean
    pass); want_mean_out: also return the per-sample column means of H_out (else an empty tensor)."""
    _need_cuda(H_in, W, bias)
    L = _lib.load()
    H_in, W, bias = _f32c(H_in), _f32c(W), _f32c(bias)
    B, N, k = H_in.shape
    q = W.shape[1]
    dev = H_in.device
    out = torch.empty((B, N, q), dtype=torch.float32, device=dev)
    mu = _f32c(mu_in).reshape(B, k) if mu_in is not None else torch.empty((B, k), dtype=torch.float32, device=dev)
    mean_out = torch.empty((B, q) if want_mean_out else (0,), dtype=torch.float32, device=dev)
    ws = _workspace(L.nbpc_set_layer_workspace_bytes(B, N, k, q), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_set_layer_fwd_chained(_ptr(H_in), B, N, k, q, _ptr(W), _ptr(bias), int(relu), _ptr(out), _ptr(mu),
                                          int(mu_in is not None), _ptr(mean_out) if want_mean_out else None,
                                          _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_set_layer_fwd")
    return out, (mu.clone() if mu_in is not None else mu), mean_out


@torch.library.custom_op("nbpc::set_layer_bwd", mutates_args=())
def set_layer_bwd(dOut: torch.Tensor, H_in: torch.Tensor, H_out: torch.Tensor, mu: torch.Tensor, W: torch.Tensor,
                  relu: bool, need_dH: bool, mask_input: bool = False, dz_sums: Optional[torch.Tensor] = None,
                  want_dh_sums: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (dH, dW, dB, dh_sums).  dz_sums: per-sample column sums of dOut handed over by the layer that wrote dOut;
    want_dh_sums: also return the per-sample column sums of dH (else an empty tensor)."""
    _need_cuda(dOut, H_in, H_out, mu, W)
    L = _lib.load()
    dOut = _f32c(dOut)
    B, N, k = H_in.shape
    q = W.shape[1]
    dev = H_in.device
    want_dh_sums = want_dh_sums and need_dH
    dH = torch.empty((B, N, k) if need_dH else (0,), dtype=torch.float32, device=dev)
    dW = torch.empty((k, q), dtype=torch.float32, device=dev)
    dB = torch.empty((q,), dtype=torch.float32, device=dev)
    dh_sums = torch.empty((B, k) if want_dh_sums else (0,), dtype=torch.float32, device=dev)
    ws = _workspace(L.nbpc_set_layer_workspace_bytes(B, N, k, q), dev)
    with torch.cuda.device(dev):
        rc = L.nbpc_set_layer_bwd_chained(_ptr(dOut), _ptr(H_in), _ptr(H_out), _ptr(mu), B, N, k, q, _ptr(W), int(relu), int(mask_input),
                                          _ptr(dH) if need_dH else None, _ptr(dW), _ptr(dB),
                                          _ptr(_f32c(dz_sums)) if dz_sums is not None else None,
                                          _ptr(dh_sums) if want_dh_sums else None, _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_set_layer_bwd")
    return dH, dW, dB, dh_sums


class SetChain:
    """Side channel between the set layers of one nn.network_func_set stack (every hidden tensor has exactly one consumer,
    the next layer): layer l's forward leaves the column means of its output for layer l+1's forward, layer l+1's
    backward leaves the column sums of its dH for layer l's backward - the kernels that WRITE those tensors produce the
    sums, so nobody re-reads a (B,N,C) tensor for a mean pass.  Entries are popped by their single consumer; a missing
    entry just means the consumer computes the statistic itself."""

    def __init__(self):
        self.mean = {}     # layer index -> (B,k) column means of that layer's input
        self.dsum = {}     # layer index -> (B,q) column sums of that layer's output gradient


class SetLayer(torch.autograd.Function):
    """set_layer (nn.py:10-28) [+ fused ReLU].  input_relu / grad_premasked: as in GraphLayer - inside a network whose
    hidden tensors have exactly one consumer, the ReLU backward of layer l is applied by layer l+1's backward kernel
    (which has H_in in hand) and layer l skips its own mask (nn.network_func_set sets both consistently).
    chain / idx / last: SetChain of the stack, this layer's index and whether it is the last one."""

    @staticmethod
    def forward(ctx, H_in, W, bias, relu, input_relu=False, grad_premasked=False, chain=None, idx=0, last=True):
        H_in = _f32c(H_in)
        mu_in = chain.mean.pop(idx, None) if chain is not None else None
        fwd_fuse = chain is not None and not last
        out, mu, mean_out = set_layer_fwd(H_in, W, bias, relu, mu_in, fwd_fuse)
        if fwd_fuse:
            chain.mean[idx + 1] = mean_out
        ctx.save_for_backward(H_in, out, mu, W)
        ctx.cfg = (relu and not grad_premasked, input_relu)
        ctx.chain = (chain, idx)
        return out

    @staticmethod
    def backward(ctx, g):
        H_in, out, mu, W = ctx.saved_tensors
        relu, input_relu = ctx.cfg
        chain, idx = ctx.chain
        need_dH = ctx.needs_input_grad[0]
        dz_sums = chain.dsum.pop(idx, None) if (chain is not None and not relu) else None
        want = chain is not None and idx > 0 and need_dH
        dH, dW, dB, dh_sums = set_layer_bwd(g, H_in, out, mu, W, relu, need_dH, input_relu, dz_sums, want)
        if want:
            chain.dsum[idx - 1] = dh_sums
        return (dH if need_dH else None), dW, dB, None, None, None, None, None, None


# ================================================================== losses / readout
def _rows_ld(t: torch.Tensor):
    t = _f32c(t)
    ld = t.shape[-1]
    return t, t.numel() // ld, ld


@torch.library.custom_op("nbpc::loss_fwd", mutates_args=())
def loss_fwd(pred: torch.Tensor, truth: torch.Tensor, pbc: bool, scale_error: bool) -> torch.Tensor:
    _need_cuda(pred, truth)
    L = _lib.load()
    pred, rows, ldp = _rows_ld(pred)
    truth, rows_t, ldt = _rows_ld(truth)
    if rows != rows_t or ldp < 3 or ldt < 3:
        raise RuntimeError("loss: pred/truth must have the same leading shape and >= 3 channels")
    out = torch.empty((), dtype=torch.float32, device=pred.device)
    ws = _workspace(L.nbpc_loss_workspace_bytes(rows), pred.device)
    with torch.cuda.device(pred.device):
        if pbc:
            rc = L.nbpc_pbc_loss_fwd(_ptr(pred), ldp, _ptr(truth), ldt, rows, int(scale_error), _ptr(out), _ptr(ws),
                                     ws.numel(), _stream())
        else:
            rc = L.nbpc_loss_za_fwd(_ptr(pred), ldp, _ptr(truth), ldt, rows, _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbpc_loss_fwd")
    return out


@torch.library.custom_op("nbpc::loss_bwd", mutates_args=())
def loss_bwd(pred: torch.Tensor, truth: torch.Tensor, dloss: torch.Tensor, pbc: bool, scale_error: bool) -> torch.Tensor:
    _need_cuda(pred, truth, dloss)
    L = _lib.load()
    pred, rows, ldp = _rows_ld(pred)
    truth, _, ldt = _rows_ld(truth)
    dloss = _f32c(dloss).reshape(1)
    dpred = torch.zeros_like(pred) if ldp > 3 else torch.empty_like(pred)
    with torch.cuda.device(pred.device):
        if pbc:
            rc = L.nbpc_pbc_loss_bwd(_ptr(pred), ldp, _ptr(truth), ldt, rows, int(scale_error), _ptr(dloss),
                                     _ptr(dpred), ldp, _stream())
        else:
            rc = L.nbpc_loss_za_bwd(_ptr(pred), ldp, _ptr(truth), ldt, rows, _ptr(dloss), _ptr(dpred), ldp, _stream())
    _lib.check(rc, "nbpc_loss_bwd")
    return dpred


class Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, truth, pbc, scale_error):
        ctx.save_for_backward(pred, truth)
        ctx.cfg = (pbc, scale_error)
        return loss_fwd(pred, truth, pbc, scale_error)

    @staticmethod
    def backward(ctx, g):
        pred, truth = ctx.saved_tensors
        pbc, scale_error = ctx.cfg
        return loss_bwd(pred, truth, g, pbc, scale_error).reshape(pred.shape), None, None, None


@torch.library.custom_op("nbpc::periodic_boundary_dist", mutates_args=())
def periodic_boundary_dist(pred: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    _need_cuda(pred, truth)
    L = _lib.load()
    pred, rows, ldp = _rows_ld(pred)
    truth, _, ldt = _rows_ld(truth)
    out = torch.empty(pred.shape[:-1] + (3,), dtype=torch.float32, device=pred.device)
    with torch.cuda.device(pred.device):
        rc = L.nbpc_periodic_boundary_dist(_ptr(pred), ldp, _ptr(truth), ldt, rows, _ptr(out), _stream())
    _lib.check(rc, "nbpc_periodic_boundary_dist")
    return out


@torch.library.custom_op("nbpc::readout", mutates_args=())
def readout(h: torch.Tensor) -> torch.Tensor:
    _need_cuda(h)
    L = _lib.load()
    h, rows, C = _rows_ld(h)
    out = torch.empty_like(h)
    with torch.cuda.device(h.device):
        rc = L.nbpc_readout(_ptr(h), rows, C, _ptr(out), _stream())
    _lib.check(rc, "nbpc_readout")
    return out


def residual_update(X: torch.Tensor, net: torch.Tensor, loc_scalar: float, vel_scalar: float) -> torch.Tensor:
    """graph.py:558-566 in one kernel (forward only): X (..., >= 6) = [loc, vel], net (..., 3 | 6)."""
    _need_cuda(X, net)
    L = _lib.load()
    X, net = _f32c(X), _f32c(net)
    C, ldx = net.shape[-1], X.shape[-1]
    rows = net.numel() // C
    out = torch.empty_like(net)
    with torch.cuda.device(net.device):
        rc = L.nbpc_residual_update(_ptr(X), ldx, _ptr(net), C, rows, float(loc_scalar), float(vel_scalar), _ptr(out), _stream())
    _lib.check(rc, "nbpc_residual_update")
    return out


class Readout(torch.autograd.Function):
    """get_readout (nn.py:107-119): d readout / d h = 1 almost everywhere (tf.sign has zero gradient)."""

    @staticmethod
    def forward(ctx, h):
        return readout(h)

    @staticmethod
    def backward(ctx, g):
        return g


# ================================================================== optimiser
def adam_tf_(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
             beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    """In-place tf.train.AdamOptimizer step (train.py:70) on flat float32 buffers."""
    _need_cuda(param, grad, m, v)
    L = _lib.load()
    for t in (param, grad, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("adam_tf_: buffers must be contiguous float32")
    with torch.cuda.device(param.device):
        rc = L.nbpc_adam_tf(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(), lr, beta1, beta2, eps, step,
                            grad_scale, _stream())
    _lib.check(rc, "nbpc_adam_tf")


def adam_tf_dev_(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step_counter: torch.Tensor,
                 lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    """adam_tf_ with the step count in a device int64 tensor (incremented by the call): CUDA-graph capturable."""
    _need_cuda(param, grad, m, v, step_counter)
    L = _lib.load()
    for t in (param, grad, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("adam_tf_dev_: buffers must be contiguous float32")
    if step_counter.dtype != torch.int64 or step_counter.numel() != 1:
        raise RuntimeError("adam_tf_dev_: step_counter must be one int64")
    with torch.cuda.device(param.device):
        rc = L.nbpc_adam_tf_dev(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(), lr, beta1, beta2, eps,
                                _ptr(step_counter), grad_scale, _stream())
    _lib.check(rc, "nbpc_adam_tf_dev")


def device_check() -> None:
    """Raise unless the current CUDA device is an sm_100 part."""
    _lib.check(_lib.load().nbpc_device_check(), "nbpc_device_check")
