"""TEST INFRASTRUCTURE ONLY: host-emulation harness.

Builds tests/_emu/libnbpc_emu.so (the barrier-free baseline kernels of csrc/*.cu compiled as
plain C++ with -DNBPC_HOST_EMU; every launch becomes a sequential loop) and drives the SAME C
ABI on NumPy arrays, so the kernels' index algebra and the host orchestration are checked on
the CPU-only box.  The product package never loads this library.
"""
import ctypes
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "n-body_pointcloudevolution_b200", "csrc")
EMU_SO = os.path.join(ROOT, "tests", "_emu", "libnbpc_emu.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", CSRC, "emu"])
        binding = importlib.import_module("n-body_pointcloudevolution_b200._lib")
        _lib = binding.bind(ctypes.CDLL(EMU_SO))
    return _lib


def P(a):
    return None if a is None else a.ctypes.data


def ok(rc):
    assert rc == 0, lib().nbpc_last_error_string().decode()


def ws(nbytes):
    return np.zeros(max(int(nbytes), 8), dtype=np.uint8)


def knn(x, k, periodic=False, thr=0.0, include_self=True, order=0, want_d2=False):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, D = x.shape
    idx = np.full((B, N, k), -7, dtype=np.int32)
    d2 = np.zeros((B, N, k), dtype=np.float64) if want_d2 else None
    w = ws(L.nbpc_knn_workspace_bytes(B, N, k, int(periodic)))
    ok(L.nbpc_knn(P(x), N * D, D, B, N, k, int(periodic), float(thr), int(include_self), order,
                  P(idx), P(d2), None, P(w), w.nbytes, None))
    return (idx, d2) if want_d2 else idx


def adjacency(idx):
    L = lib()
    B, N, M = idx.shape
    c = B * N * M
    coo = np.zeros((3, c), dtype=np.int32)
    diag = np.zeros((B * N,), dtype=np.int64)
    ptr = np.zeros((B * N + 1,), dtype=np.int32)
    edge = np.zeros((c,), dtype=np.int32)
    status = np.zeros((2,), dtype=np.int32)
    w = ws(L.nbpc_adjacency_workspace_bytes(B, N, M))
    ok(L.nbpc_adjacency(P(np.ascontiguousarray(idx)), B, N, M, P(coo), P(diag), P(ptr), P(edge), P(status),
                        P(w), w.nbytes, None))
    return coo, diag, ptr, edge, status


def segment_csr(ids, num_segs):
    L = lib()
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    n = ids.shape[0]
    ptr = np.zeros((num_segs + 1,), dtype=np.int32)
    mem = np.zeros((n,), dtype=np.int32)
    status = np.zeros((1,), dtype=np.int32)
    w = ws(L.nbpc_segment_csr_workspace_bytes(n, num_segs))
    ok(L.nbpc_segment_csr(P(ids), n, num_segs, P(ptr), P(mem), P(status), P(w), w.nbytes, None))
    return ptr, mem, status


def edge_features_za(pos, za, col, diag, M):
    L = lib()
    pos = np.ascontiguousarray(pos.reshape(-1, pos.shape[-1]), dtype=np.float32)
    BN = pos.shape[0]
    out = np.zeros((BN * M, 3), dtype=np.float32)
    if za is None:
        ok(L.nbpc_edge_features(P(pos), pos.shape[1], P(col), BN, M, P(out), None))
    else:
        za = np.ascontiguousarray(za.reshape(-1, za.shape[-1]), dtype=np.float32)
        ok(L.nbpc_edge_features_za(P(pos), pos.shape[1], P(za), za.shape[1], P(col), P(diag), diag.shape[0], BN, M,
                                   P(out), None))
    return out


def graph_layer_fwd(H, col, ptr, edge, B, N, M, W, bias, is_last, relu):
    L = lib()
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    out = np.zeros(((B * N) if is_last else c, q), dtype=np.float32)
    Pc = np.zeros((B * N, k), dtype=np.float32)
    Pr = np.zeros((B * N, k), dtype=np.float32)
    Pq = np.zeros((B, k), dtype=np.float32)
    w = ws(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q))
    ok(L.nbpc_graph_layer_fwd(P(H), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(bias), int(is_last), int(relu),
                              P(out), P(Pc), P(Pr), P(Pq), P(w), w.nbytes, None))
    return out, (Pc, Pr, Pq)


def graph_layer_bwd(dOut, H, Hout, col, ptr, edge, B, N, M, W, saved, is_last, relu, need_dH=True, mask_input=False):
    L = lib()
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    dH = np.zeros((c, k), dtype=np.float32) if need_dH else None
    dW = np.zeros((4, k, q), dtype=np.float32)
    dB = np.zeros((q,), dtype=np.float32)
    Pc, Pr, Pq = saved
    w = ws(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q))
    ok(L.nbpc_graph_layer_bwd(P(dOut), P(H), P(Hout), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(Pc), P(Pr),
                              P(Pq), int(is_last), int(relu), int(mask_input), P(dH), P(dW), P(dB), P(w), w.nbytes, None))
    return dH, dW, dB


def set_layer_fwd(H, W, bias, relu):
    L = lib()
    B, N, k = H.shape
    q = W.shape[1]
    out = np.zeros((B, N, q), dtype=np.float32)
    mu = np.zeros((B, k), dtype=np.float32)
    w = ws(L.nbpc_set_layer_workspace_bytes(B, N, k, q))
    ok(L.nbpc_set_layer_fwd(P(H), B, N, k, q, P(W), P(bias), int(relu), P(out), P(mu), P(w), w.nbytes, None))
    return out, mu


def set_layer_bwd(dOut, H, Hout, mu, W, relu, need_dH=True, mask_input=False):
    L = lib()
    B, N, k = H.shape
    q = W.shape[1]
    dH = np.zeros((B, N, k), dtype=np.float32) if need_dH else None
    dW = np.zeros((k, q), dtype=np.float32)
    dB = np.zeros((q,), dtype=np.float32)
    w = ws(L.nbpc_set_layer_workspace_bytes(B, N, k, q))
    ok(L.nbpc_set_layer_bwd(P(dOut), P(H), P(Hout), P(mu), B, N, k, q, P(W), int(relu), int(mask_input), P(dH), P(dW), P(dB),
                            P(w), w.nbytes, None))
    return dH, dW, dB


def set_layer_fwd_chained(H, W, bias, relu, mu_in=None, want_mean_out=True):
    L = lib()
    B, N, k = H.shape
    q = W.shape[1]
    out = np.zeros((B, N, q), dtype=np.float32)
    mu = np.ascontiguousarray(mu_in, dtype=np.float32).copy() if mu_in is not None else np.zeros((B, k), dtype=np.float32)
    mean_out = np.zeros((B, q), dtype=np.float32) if want_mean_out else None
    w = ws(L.nbpc_set_layer_workspace_bytes(B, N, k, q))
    ok(L.nbpc_set_layer_fwd_chained(P(H), B, N, k, q, P(W), P(bias), int(relu), P(out), P(mu), int(mu_in is not None), P(mean_out),
                                    P(w), w.nbytes, None))
    return out, mu, mean_out


def set_layer_bwd_chained(dOut, H, Hout, mu, W, relu, mask_input=False, dz_sums=None, want_dh_sums=True):
    L = lib()
    B, N, k = H.shape
    q = W.shape[1]
    dH = np.zeros((B, N, k), dtype=np.float32)
    dW = np.zeros((k, q), dtype=np.float32)
    dB = np.zeros((q,), dtype=np.float32)
    dh_sums = np.zeros((B, k), dtype=np.float32) if want_dh_sums else None
    w = ws(L.nbpc_set_layer_workspace_bytes(B, N, k, q))
    ok(L.nbpc_set_layer_bwd_chained(P(dOut), P(H), P(Hout), P(mu), B, N, k, q, P(W), int(relu), int(mask_input), P(dH), P(dW), P(dB),
                                    P(dz_sums), P(dh_sums), P(w), w.nbytes, None))
    return dH, dW, dB, dh_sums


def loss(pred, truth, pbc=False, scale=True):
    L = lib()
    rows = pred.shape[0] * pred.shape[1]
    out = np.zeros((1,), dtype=np.float32)
    w = ws(L.nbpc_loss_workspace_bytes(rows))
    if pbc:
        ok(L.nbpc_pbc_loss_fwd(P(pred), pred.shape[-1], P(truth), truth.shape[-1], rows, int(scale), P(out), P(w),
                               w.nbytes, None))
    else:
        ok(L.nbpc_loss_za_fwd(P(pred), pred.shape[-1], P(truth), truth.shape[-1], rows, P(out), P(w), w.nbytes, None))
    return out[0]


def loss_bwd(pred, truth, pbc=False, scale=True, dloss=1.0):
    L = lib()
    rows = pred.shape[0] * pred.shape[1]
    dl = np.array([dloss], dtype=np.float32)
    dp = np.zeros(pred.shape[:2] + (3,), dtype=np.float32)
    if pbc:
        ok(L.nbpc_pbc_loss_bwd(P(pred), pred.shape[-1], P(truth), truth.shape[-1], rows, int(scale), P(dl), P(dp), 3, None))
    else:
        ok(L.nbpc_loss_za_bwd(P(pred), pred.shape[-1], P(truth), truth.shape[-1], rows, P(dl), P(dp), 3, None))
    return dp


def knn_status(x, k, thr):
    """periodic build -> number of particles outside the unit box (nbpc_knn status word)"""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, D = x.shape
    idx = np.zeros((B, N, k), dtype=np.int32)
    status = np.full((1,), -1, dtype=np.int32)
    w = ws(L.nbpc_knn_workspace_bytes(B, N, k, 1))
    ok(L.nbpc_knn(P(x), N * D, D, B, N, k, 1, float(thr), 1, 0, P(idx), None, P(status), P(w), w.nbytes, None))
    return int(status[0])


def pad_cube(x, thr):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, D = x.shape
    offsets = np.zeros(N + 1, dtype=np.int32)
    w = ws(L.nbpc_pad_cube_workspace_bytes(N))
    ok(L.nbpc_pad_cube_count(P(x), D, N, float(thr), P(offsets), P(w), w.nbytes, None))
    n_img = int(offsets[N])
    padded = np.zeros((N + n_img, 3), dtype=np.float64)
    idx_map = np.zeros(max(n_img, 1), dtype=np.int64)
    ok(L.nbpc_pad_cube_emit(P(x), D, N, float(thr), P(offsets), P(padded), P(idx_map), None))
    return padded, idx_map[:n_img]


def sym_adjacency(idx):
    """canonical symmetrised adjacency of kNN lists idx (B,N,M) through the C ABI (host emulation)"""
    L = lib()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, N, M = idx.shape
    _, _, ptr, edge, _ = adjacency(idx)
    BN = B * N
    row_ptr = np.zeros(BN + 1, dtype=np.int32)
    w = ws(L.nbpc_sym_adjacency_workspace_bytes(B, N))
    ok(L.nbpc_sym_adjacency_count(P(idx), P(ptr), P(edge), B, N, M, P(row_ptr), P(w), w.nbytes, None))
    S = int(row_ptr[BN])
    out = {n: np.zeros(S, dtype=np.int32) for n in ("row", "col", "all", "tra")}
    out.update({n: np.zeros(BN, dtype=np.int32) for n in ("dia", "dal")})
    status = np.zeros(2, dtype=np.int32)
    ok(L.nbpc_sym_adjacency_emit(P(idx), P(ptr), P(edge), P(row_ptr), B, N, M, S, P(out["row"]), P(out["col"]), P(out["all"]),
                                 P(out["tra"]), P(out["dia"]), P(out["dal"]), P(status), None))
    out["row_ptr"] = row_ptr
    return out, status


def graph15_fwd(H, adj, B, N, W, Bias, relu=False):
    L = lib()
    H, W, Bias = (np.ascontiguousarray(a, dtype=np.float32) for a in (H, W, Bias))
    S, k = H.shape
    q = W.shape[2]
    out = np.zeros((S, q), dtype=np.float32)
    Hr, Hc, Hd = (np.zeros((B * N, k), dtype=np.float32) for _ in range(3))
    Ha, Hp = (np.zeros((B, k), dtype=np.float32) for _ in range(2))
    w = ws(L.nbpc_graph15_workspace_bytes(B, N, S, k, q))
    ok(L.nbpc_graph15_layer_fwd(P(H), P(adj["row"]), P(adj["col"]), P(adj["tra"]), P(adj["dia"]), P(adj["row_ptr"]), B, N, S, k, q, P(W),
                                P(Bias), int(relu), P(out), P(Hr), P(Hc), P(Hd), P(Ha), P(Hp), P(w), w.nbytes, None))
    return out, (Hr, Hc, Hd, Ha, Hp)


def graph15_bwd(dOut, H, Hout, adj, B, N, W, saved, relu=False, need_dH=True):
    L = lib()
    dOut, H, Hout, W = (np.ascontiguousarray(a, dtype=np.float32) for a in (dOut, H, Hout, W))
    S, k = H.shape
    q = W.shape[2]
    Hr, Hc, Hd, Ha, Hp = saved
    dH = np.zeros((S, k), dtype=np.float32) if need_dH else None
    dW = np.zeros((15, k, q), dtype=np.float32)
    dB = np.zeros((2, q), dtype=np.float32)
    w = ws(L.nbpc_graph15_workspace_bytes(B, N, S, k, q))
    ok(L.nbpc_graph15_layer_bwd(P(dOut), P(H), P(Hout), P(adj["row"]), P(adj["col"]), P(adj["tra"]), P(adj["dia"]), P(adj["row_ptr"]), B, N,
                                S, k, q, P(W), P(Hr), P(Hc), P(Hd), P(Ha), P(Hp), int(relu), P(dH), P(dW), P(dB), P(w), w.nbytes, None))
    return dH, dW, dB
