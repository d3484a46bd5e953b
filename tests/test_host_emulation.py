"""CPU: host emulation of the barrier-free baseline kernels (tests/emu.py) through the real C ABI,
against the golden vectors of the unmodified reference and the oracle.  This checks index algebra,
the kNN traversal/termination logic, bit-exactness of the float64 distance path and the backward
formulas before any GPU time is spent; the GPU parity tests (-m gpu) repeat all of it on device."""
import numpy as np
import pytest

import emu
from conftest import load_golden
from oracle import ref_graph
from oracle.knn_exact import knn_exact


def col_sorted(idx):
    return np.sort(idx, axis=-1)


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_emu_knn_16_matches_reference(kind):
    g = load_golden("knn_16.npz")
    for seed in (0, 1):
        tag = f"{kind}_s{seed}"
        x = g[f"x_{tag}"]
        assert np.array_equal(emu.knn(x, 14, order=1), g[f"knl_{tag}"])                       # get_kneighbor_list
        assert np.array_equal(emu.knn(x, 14, periodic=True, thr=0.1, include_self=False), g[f"pbc_{tag}"])
        if seed == 0:
            assert np.array_equal(emu.knn(x, 14, periodic=True, thr=0.1), g[f"pbcself_{tag}"])
            assert np.array_equal(emu.knn(x, 8, periodic=True, thr=0.3, include_self=False), g[f"pbc03_{tag}"])
            assert np.array_equal(emu.knn(x, 8, include_self=False, order=1), g[f"knlnoself_{tag}"])


def test_emu_knn_d2_and_full_wrap():
    rng = np.random.default_rng(3)
    x = rng.random((1, 700, 3)).astype(np.float32)
    # thr = 0.5: every particle has all 7 images (8N padded points) -> exact minimum-image kNN
    idx, d2 = emu.knn(x, 33, periodic=True, thr=0.5, include_self=True, want_d2=True)
    ref = ref_graph.get_pbc_kneighbors_csr(x, 33, 0.5, include_self=True, backend="exact")[0]
    assert np.array_equal(idx[0], ref.indices.reshape(700, 33))
    padded, _ = ref_graph.pad_cube_boundaries(x[0], 0.5)
    _, rd2 = knn_exact(padded, 700, 33, True, return_d2=True)
    assert np.array_equal(d2[0], rd2)                     # bit-exact float64 distances
    # non-periodic, arbitrary box (reference feeds Mpc/h coordinates to get_kneighbor_list)
    y = (x * 128.0 - 3.0).astype(np.float32)
    idx = emu.knn(y, 5)
    assert np.array_equal(idx[0], knn_exact(y[0].astype(np.float64), 700, 5, True))


def test_emu_knn_tiny_and_ties():
    # tiny N (grid of 1..2 cells), k = N, duplicates (exact ties -> index order)
    x = np.array([[[0.1, 0.1, 0.1], [0.9, 0.9, 0.9], [0.1, 0.1, 0.1], [0.5, 0.5, 0.5]]], dtype=np.float32)
    idx = emu.knn(x, 4)
    assert np.array_equal(idx[0], knn_exact(x[0].astype(np.float64), 4, 4, True))
    assert idx[0, 0].tolist()[:2] == [0, 2] and idx[0, 2].tolist()[:2] == [0, 2]
    g = load_golden("lattice_8.npz")
    xl = g["x"]
    idx, d2 = emu.knn(xl, 14, want_d2=True)
    assert np.array_equal(d2[0], g["knl_sorted_d2"])     # distances are tie-independent
    assert np.array_equal(idx[0], knn_exact(xl[0].astype(np.float64), 512, 14, True))


def test_emu_adjacency():
    g = load_golden("layers_small.npz")
    x = g["x"]; k = int(g["k"])
    idx = emu.knn(x, k, order=1)
    coo, diag, ptr, edge, status = emu.adjacency(idx)
    assert np.array_equal(coo, g["coo"]) and np.array_equal(diag, g["diag"]) and status.tolist() == [0, 0]
    BN = x.shape[0] * x.shape[1]
    assert ptr[0] == 0 and ptr[-1] == coo.shape[1]
    ref_order = np.argsort(coo[1], kind="stable")
    assert np.array_equal(edge, ref_order) and np.array_equal(np.diff(ptr), np.bincount(coo[1], minlength=BN))
    # generic segment CSR on unsorted and on monotone ids
    ids = np.random.default_rng(0).integers(0, 37, size=5000).astype(np.int32)
    p2, m2, st = emu.segment_csr(ids, 37)
    assert np.array_equal(m2, np.argsort(ids, kind="stable")) and st[0] == 0
    p3, m3, st = emu.segment_csr(np.sort(ids), 37)
    assert np.array_equal(m3, np.arange(5000)) and np.array_equal(p3, p2)


def test_emu_features_and_layers_vs_golden():
    g = load_golden("layers_small.npz")
    x, za = g["x"], g["za"]
    B, N = x.shape[:2]; M = int(g["k"]); ch = list(g["channels"])
    coo, diag, ptr, edge, _ = emu.adjacency(emu.knn(x, M, order=1))
    col = np.ascontiguousarray(coo[1])
    H = emu.edge_features_za(x, za, col, diag, M)
    assert np.array_equal(H, g["f32_edges"])
    X6 = g["feat_X6"]
    assert np.array_equal(emu.edge_features_za(X6, None, col, None, M), g["feat_edges"])

    acts, saved = [H], []
    L = len(ch) - 1
    Ws = [np.stack([g[f"W{li}_{wi}"] for wi in range(4)]) for li in range(L)]
    Bs = [g[f"B{li}"] for li in range(L)]
    for li in range(L):
        last = li == L - 1
        out, sv = emu.graph_layer_fwd(acts[-1], col, ptr, edge, B, N, M, Ws[li], Bs[li], last, not last)
        ref = g[f"f32_H{li}"].reshape(out.shape)
        np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)
        # and closer to (or as close as) the float64 run of the reference than float32 noise allows
        ref64 = g[f"f64_H{li}"].reshape(out.shape)
        assert np.abs(out - ref64).max() <= 4 * max(np.abs(ref - ref64).max(), 1e-7)
        acts.append(out); saved.append(sv)
    pred = acts[-1].reshape(B, N, 3)
    tgt = g["tgt"]
    np.testing.assert_allclose(emu.loss(pred, tgt), g["f32_loss"], rtol=1e-5)
    dOut = emu.loss_bwd(pred, tgt).reshape(B * N, 3)
    for li in reversed(range(L)):
        last = li == L - 1
        dH, dW, dB = emu.graph_layer_bwd(dOut, acts[li], acts[li + 1], col, ptr, edge, B, N, M, Ws[li], saved[li],
                                         last, not last, need_dH=li > 0)
        for wi in range(4):
            np.testing.assert_allclose(dW[wi], g[f"f64_gW{li}_{wi}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(dB, g[f"f64_gB{li}"], rtol=2e-4, atol=1e-7)
        dOut = dH


def test_emu_layer_odd_widths():
    g = load_golden("layers_small.npz")
    B, N = g["x"].shape[:2]; M = int(g["k"])
    coo, diag, ptr, edge, _ = emu.adjacency(emu.knn(g["x"], M, order=1))
    col = np.ascontiguousarray(coo[1])
    H = g["odd_H_in"]; W = np.stack([g[f"odd_W{i}"] for i in range(4)]); Bv = g["odd_B"]
    for last, t in ((False, "mid"), (True, "last")):
        out, sv = emu.graph_layer_fwd(H, col, ptr, edge, B, N, M, W, Bv, last, False)
        np.testing.assert_allclose(out, g[f"odd_{t}_out"].reshape(out.shape), rtol=2e-5, atol=2e-6)
        gout = np.ascontiguousarray(g[f"odd_{t}_gout"].reshape(out.shape))
        dH, dW, dB = emu.graph_layer_bwd(gout, H, out, col, ptr, edge, B, N, M, W, sv, last, False)
        np.testing.assert_allclose(dH, g[f"odd_{t}_gH"], rtol=1e-4, atol=1e-5)
        for i in range(4):
            np.testing.assert_allclose(dW[i], g[f"odd_{t}_gW{i}"], rtol=1e-4, atol=2e-4)
        np.testing.assert_allclose(dB, g[f"odd_{t}_gB"], rtol=1e-4, atol=2e-4)


def test_emu_rowpool_entry_points_degrade_to_the_plain_layer():
    """nbpc_graph_layer_fwd_rp / _bwd_rp (row-pool hand-over between layers, include/nbpc.h): with no hand-over requested they
    ARE the plain entry points; the host build has none of the emitting kernels, says so (`_rowpool_supported` == 0) and rejects
    a hand-over with NBPC_EINVAL instead of ignoring it."""
    g = load_golden("layers_small.npz")
    B, N = g["x"].shape[:2]; M = int(g["k"])
    coo, diag, ptr, edge, _ = emu.adjacency(emu.knn(g["x"], M, order=1))
    col = np.ascontiguousarray(coo[1])
    H = g["odd_H_in"]; W = np.stack([g[f"odd_W{i}"] for i in range(4)]); Bv = g["odd_B"]
    L, P = emu.lib(), emu.P
    k, q = W.shape[1], W.shape[2]
    c = B * N * M
    for direction in range(4):
        assert L.nbpc_graph_layer_rowpool_supported(3, 32, 0, direction) == 0
    out, (Pc, Pr, Pq) = emu.graph_layer_fwd(H, col, ptr, edge, B, N, M, W, Bv, False, False)
    out2, Pc2, Pr2, Pq2 = np.zeros_like(out), np.zeros_like(Pc), np.zeros_like(Pr), np.zeros_like(Pq)
    w = emu.ws(L.nbpc_graph_layer_workspace_bytes(B, N, M, k, q))
    emu.ok(L.nbpc_graph_layer_fwd_rp(P(H), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(Bv), 0, 0, P(out2), P(Pc2), P(Pr2), P(Pq2),
                                     0, None, P(w), w.nbytes, None))
    assert np.array_equal(out, out2) and np.array_equal(Pc, Pc2) and np.array_equal(Pr, Pr2) and np.array_equal(Pq, Pq2)
    nxt = np.zeros((B * N, q), dtype=np.float32)
    assert L.nbpc_graph_layer_fwd_rp(P(H), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(Bv), 0, 0, P(out2), P(Pc2), P(Pr2), P(Pq2),
                                     0, P(nxt), P(w), w.nbytes, None) != 0
    assert L.nbpc_graph_layer_fwd_rp(P(H), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(Bv), 0, 0, P(out2), P(Pc2), P(Pr2), P(Pq2),
                                     1, None, P(w), w.nbytes, None) != 0
    gout = np.ascontiguousarray(g["odd_mid_gout"].reshape(out.shape))
    dH, dW, dB = emu.graph_layer_bwd(gout, H, out, col, ptr, edge, B, N, M, W, (Pc, Pr, Pq), False, False)
    dH2, dW2, dB2 = np.zeros_like(dH), np.zeros_like(dW), np.zeros_like(dB)
    emu.ok(L.nbpc_graph_layer_bwd_rp(P(gout), P(H), P(out), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(Pc), P(Pr), P(Pq), 0, 0, 0,
                                     P(dH2), P(dW2), P(dB2), None, None, P(w), w.nbytes, None))
    assert np.array_equal(dH, dH2) and np.array_equal(dW, dW2) and np.array_equal(dB, dB2)
    dq = np.zeros((B * N, q), dtype=np.float32)
    assert L.nbpc_graph_layer_bwd_rp(P(gout), P(H), P(out), P(col), P(ptr), P(edge), B, N, M, k, q, P(W), P(Pc), P(Pr), P(Pq), 0, 0, 0,
                                     P(dH2), P(dW2), P(dB2), P(dq), None, P(w), w.nbytes, None) != 0


def test_emu_set_layers_and_losses():
    g = load_golden("set_small.npz")
    ch = list(g["channels"]); L = len(ch) - 1
    acts, mus = [g["X"]], []
    for li in range(L):
        out, mu = emu.set_layer_fwd(acts[-1], g[f"W{li}_0"], g[f"B{li}"], li < L - 1)
        np.testing.assert_allclose(out, g[f"f32_H{li}"], rtol=2e-5, atol=2e-6)
        acts.append(out); mus.append(mu)
    np.testing.assert_allclose(emu.loss(acts[-1], g["Y"]), g["f32_loss"], rtol=1e-5)
    dOut = emu.loss_bwd(acts[-1], g["Y"])
    for li in reversed(range(L)):
        dH, dW, dB = emu.set_layer_bwd(dOut, acts[li], acts[li + 1], mus[li], g[f"W{li}_0"], li < L - 1, need_dH=li > 0)
        np.testing.assert_allclose(dW, g[f"f64_gW{li}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(dB, g[f"f64_gB{li}"], rtol=2e-4, atol=1e-7)
        dOut = dH
    gl = load_golden("losses.npz")
    pred, truth = gl["pred"], gl["truth"]
    L_ = emu.lib()
    ro = np.zeros_like(pred)
    emu.ok(L_.nbpc_readout(emu.P(pred), pred.shape[0] * pred.shape[1], pred.shape[2], emu.P(ro), None))
    assert np.array_equal(ro, gl["f32_readout"])
    pbd = np.zeros(pred.shape[:2] + (3,), dtype=np.float32)
    emu.ok(L_.nbpc_periodic_boundary_dist(emu.P(ro), 6, emu.P(truth), 6, pred.shape[0] * pred.shape[1], emu.P(pbd), None))
    assert np.array_equal(pbd, gl["f32_pbd"])
    np.testing.assert_allclose(emu.loss(ro, truth, pbc=True), gl["f32_pbc_loss"], rtol=1e-5)
    np.testing.assert_allclose(emu.loss(ro, truth, pbc=True, scale=False), gl["f32_pbc_loss_unscaled"], rtol=1e-5)
    np.testing.assert_allclose(emu.loss_bwd(ro, truth, pbc=True), gl["f32_pbc_loss_gpred"][..., :3], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(emu.loss(pred, truth), gl["f32_loss_za"], rtol=1e-5)
    np.testing.assert_allclose(emu.loss_bwd(pred, truth), gl["f32_loss_za_gpred"], rtol=1e-5, atol=1e-9)
