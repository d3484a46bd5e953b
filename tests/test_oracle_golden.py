"""CPU: the oracle restatements (oracle/ref_graph.py, oracle/ref_layers.py, oracle/knn_exact.c)
against golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import hashlib
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_graph, ref_layers


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def idx_of(A):
    N = A[0].shape[0]
    return np.stack([a.indices.reshape(N, -1) for a in A]).astype(np.int32)


def mv_for(g, channels, dtype, prefix_w="W", n_w=4):
    tp = []
    for li in range(len(channels) - 1):
        Ws = [torch.tensor(g[f"{prefix_w}{li}_{wi}"], dtype=dtype, requires_grad=True) for wi in range(n_w)]
        B = torch.tensor(g[f"B{li}"], dtype=dtype, requires_grad=True)
        tp.append((Ws, B))
    mv = types.SimpleNamespace(var_scope="params", channels=list(channels), num_layers=len(channels) - 1,
                               activation=torch.relu, get_layer_vars=lambda i: tp[i])
    return mv, tp


@pytest.mark.parametrize("backend", ["sklearn", "exact"])
@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_knn_16_both_backends(kind, backend):
    g = load_golden("knn_16.npz")
    for seed in (0, 1, 2):
        tag = f"{kind}_s{seed}"
        x = g[f"x_{tag}"]
        assert (idx_of(ref_graph.get_kneighbor_list(x, 14, backend=backend)) == g[f"knl_{tag}"]).all()
        assert (idx_of(ref_graph.get_pbc_kneighbors_csr(x, 14, 0.1, backend=backend)) == g[f"pbc_{tag}"]).all()
        if seed == 0:
            assert (idx_of(ref_graph.get_pbc_kneighbors_csr(x, 14, 0.1, include_self=True, backend=backend))
                    == g[f"pbcself_{tag}"]).all()
            assert (idx_of(ref_graph.get_pbc_kneighbors_csr(x, 8, 0.3, backend=backend)) == g[f"pbc03_{tag}"]).all()
            assert (idx_of(ref_graph.get_kneighbor_list(x, 8, include_self=False, backend=backend))
                    == g[f"knlnoself_{tag}"]).all()


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_coo_diag_16(kind):
    g = load_golden("knn_16.npz")
    tag = f"{kind}_s0"
    A = ref_graph.get_kneighbor_list(g[f"x_{tag}"], 14, backend="exact")
    coo, diag = ref_graph.to_coo_batch_ZA_diag(A)
    assert coo.dtype == np.int32 and coo.shape == (3, 2 * 4096 * 14)
    assert sha(coo) == str(g[f"coo_sha_{tag}"])
    assert (diag == g[f"diag_{tag}"]).all() and len(diag) == 2 * 4096
    assert (ref_graph.to_coo_batch(A) == coo).all()
    assert (ref_graph.get_indices_from_list_CSR(A) == coo[1]).all()


def test_knn_32_hashes(syn):
    g = load_golden("knn_32.npz")
    for kind in ("uniform", "clustered"):
        x = syn.make_box(kind, 1, 32768, 0)
        assert sha(x) == str(g[f"x_sha_{kind}"]), "synthetic generator drifted from the golden run"
        knl = idx_of(ref_graph.get_kneighbor_list(x, 14, backend="exact"))
        assert sha(knl) == str(g[f"knl_sha_{kind}_k14"])
        assert (knl[:, :64] == g[f"knl_head_{kind}_k14"]).all()


def test_lattice_known_answer():
    """interior lattice rows have sorted distances [0, h x6, sqrt2 h x7...]; ties make
    the index order implementation-defined, distances are not."""
    g = load_golden("lattice_8.npz")
    x = g["x"]
    from oracle.knn_exact import knn_exact
    idx, d2, ties = knn_exact(x[0].astype(np.float64), 512, 14, True, return_d2=True, return_ties=True)
    assert np.array_equal(np.sort(d2, axis=1), g["knl_sorted_d2"])
    assert ties.all()  # every lattice row is tie-heavy
    h2 = (1.0 / 8) ** 2
    inner = np.all((x[0] > 0.2) & (x[0] < 0.8), axis=1)
    np.testing.assert_allclose(d2[inner][:, 1:7], h2, rtol=1e-6)
    np.testing.assert_allclose(d2[inner][:, 7:14], 2 * h2, rtol=1e-6)


def test_layers_small_per_layer():
    g = load_golden("layers_small.npz")
    ch = list(g["channels"]); k = int(g["k"]); b, N = g["x"].shape[:2]
    coo, diag = g["coo"], g["diag"]
    for dt, tag, tol in ((torch.float32, "f32", 2e-6), (torch.float64, "f64", 1e-13)):
        mv, tp = mv_for(g, ch, dt)
        pos = torch.tensor(g["x"], dtype=dt); za = torch.tensor(g["za"], dtype=dt)
        H = ref_layers.get_input_features_shift_inv_ZA(pos, za, coo, diag, (b, N, k))
        np.testing.assert_allclose(H.numpy(), g[f"{tag}_edges"], rtol=0, atol=tol)
        for li in range(len(ch) - 1):
            last = li == len(ch) - 2
            H = ref_layers.shift_inv_layer(H, coo, (b, N), tp[li], is_last=last)
            if not last:
                H = torch.relu(H)
            np.testing.assert_allclose(H.detach().numpy(), g[f"{tag}_H{li}"], rtol=tol * 10, atol=tol)
        pred = ref_layers.model_func_shift_inv_za(pos, coo, za, diag, mv, (b, N, k))
        loss = ref_layers.loss_ZA(pred, torch.tensor(g["tgt"], dtype=dt))
        loss.backward()
        np.testing.assert_allclose(pred.detach().numpy(), g[f"{tag}_pred"], rtol=tol * 10, atol=tol)
        np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"], rtol=tol * 10)
        for li, (Ws, B) in enumerate(tp):
            for wi, w in enumerate(Ws):
                np.testing.assert_allclose(w.grad.numpy(), g[f"{tag}_gW{li}_{wi}"], rtol=1e-4, atol=tol)
            np.testing.assert_allclose(B.grad.numpy(), g[f"{tag}_gB{li}"], rtol=1e-4, atol=tol)


def test_layer_odd_widths_and_conv():
    g = load_golden("layers_small.npz")
    b, N = g["x"].shape[:2]
    coo = g["coo"]
    for last, t in ((False, "mid"), (True, "last")):
        Ht = torch.tensor(g["odd_H_in"], requires_grad=True)
        Wt = [torch.tensor(g[f"odd_W{i}"], requires_grad=True) for i in range(4)]
        Bt = torch.tensor(g["odd_B"], requires_grad=True)
        o = ref_layers.shift_inv_layer(Ht, coo, (b, N), (Wt, Bt), is_last=last)
        (o * torch.tensor(g[f"odd_{t}_gout"])).sum().backward()
        np.testing.assert_allclose(o.detach().numpy(), g[f"odd_{t}_out"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Ht.grad.numpy(), g[f"odd_{t}_gH"], rtol=1e-4, atol=1e-6)
        for i in range(4):
            np.testing.assert_allclose(Wt[i].grad.numpy(), g[f"odd_{t}_gW{i}"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(Bt.grad.numpy(), g[f"odd_{t}_gB"], rtol=1e-4, atol=1e-4)
    H = torch.tensor(g["odd_H_in"])
    for ci, nm in ((0, "row"), (1, "col"), (2, "cube")):
        np.testing.assert_allclose(ref_layers.shift_inv_conv(H, coo[ci], b * N, True).numpy(),
                                   g[f"conv_{nm}_bc"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ref_layers.shift_inv_conv(H, coo[ci], b * N, False).numpy(),
                                   g[f"conv_{nm}_nobc"], rtol=1e-6, atol=1e-7)


def test_legacy_features():
    g = load_golden("layers_small.npz")
    b, N = g["x"].shape[:2]; k = int(g["k"]); coo = g["coo"]
    e, n = ref_layers.get_input_features_shift_inv(torch.tensor(g["feat_X6"]), coo, (b, N, k))
    assert np.array_equal(e.numpy(), g["feat_edges"]) and np.array_equal(n.numpy(), g["feat_nodes"])
    assert np.array_equal(ref_layers.include_node_features(e, n, coo).numpy(), g["feat_nodes9"])
    rs = torch.full((b * N * k, 1), 0.75)
    assert np.array_equal(ref_layers.include_node_features(e, n, coo, redshift=rs).numpy(), g["feat_nodes10"])


def test_set_model():
    g = load_golden("set_small.npz")
    ch = list(g["channels"])
    for dt, tag, tol in ((torch.float32, "f32", 2e-6), (torch.float64, "f64", 1e-13)):
        mv, tp = mv_for(g, ch, dt)
        pred = ref_layers.model_func_set(torch.tensor(g["X"], dtype=dt), mv)
        loss = ref_layers.loss_ZA(pred, torch.tensor(g["Y"], dtype=dt))
        loss.backward()
        np.testing.assert_allclose(pred.detach().numpy(), g[f"{tag}_pred"], rtol=tol * 10, atol=tol)
        for li, (Ws, B) in enumerate(tp):
            np.testing.assert_allclose(Ws[0].grad.numpy(), g[f"{tag}_gW{li}"], rtol=1e-4, atol=tol)
            np.testing.assert_allclose(B.grad.numpy(), g[f"{tag}_gB{li}"], rtol=1e-4, atol=tol)
            assert all(w.grad is None for w in Ws[1:])


def test_losses_readout():
    g = load_golden("losses.npz")
    for dt, tag, tol in ((torch.float32, "f32", 1e-6), (torch.float64, "f64", 1e-14)):
        p = torch.tensor(g["pred"], dtype=dt, requires_grad=True)
        t = torch.tensor(g["truth"], dtype=dt)
        ro = ref_layers.get_readout(p)
        np.testing.assert_allclose(ro.detach().numpy(), g[f"{tag}_readout"], rtol=0, atol=tol)
        np.testing.assert_allclose(ref_layers.get_readout(p[..., :3]).detach().numpy(), g[f"{tag}_readout3"], atol=tol)
        np.testing.assert_allclose(ref_layers.periodic_boundary_dist(ro, t).detach().numpy(), g[f"{tag}_pbd"], atol=tol)
        l1 = ref_layers.pbc_loss(ro, t); l1.backward()
        np.testing.assert_allclose(l1.item(), g[f"{tag}_pbc_loss"], rtol=1e-5)
        np.testing.assert_allclose(p.grad.numpy(), g[f"{tag}_pbc_loss_gpred"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(ref_layers.pbc_loss(ro, t, scale_error=False).item(),
                                   g[f"{tag}_pbc_loss_unscaled"], rtol=1e-5)
        p2 = torch.tensor(g["pred"][..., :3], dtype=dt, requires_grad=True)
        l2 = ref_layers.loss_ZA(p2, t[..., :3]); l2.backward()
        np.testing.assert_allclose(l2.item(), g[f"{tag}_loss_za"], rtol=1e-5)
        np.testing.assert_allclose(p2.grad.numpy(), g[f"{tag}_loss_za_gpred"], rtol=1e-5, atol=1e-9)


def _rollout_case(g, tag, dtype):
    import types
    ch = [int(v) for v in g[f"{tag}_channels"]]
    tp = []
    for li in range(len(ch) - 1):
        tp.append(([torch.tensor(g[f"{tag}_W{li}_{wi}"], dtype=dtype) for wi in range(4)], torch.tensor(g[f"{tag}_B{li}"], dtype=dtype)))
    scalars = tuple(float(v) for v in g[f"{tag}_scalars"])
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j], get_scalars=lambda: scalars)
    return ch, mv


@pytest.mark.parametrize("tag", ["v9", "v10"])
def test_legacy_multi_redshift_model_restatement(tag):
    """oracle.ref_layers.model_func_shift_inv vs the reference's own (commented-out) graph.py:517-567 block, executed
    unmodified by oracle/make_golden_rollout.py."""
    from oracle import ref_layers
    g = load_golden("rollout_small.npz")
    X, coo = g[f"{tag}_X"], g[f"{tag}_coo"]
    b, N = X.shape[0], X.shape[1]
    K = coo.shape[1] // (b * N)
    for dtype, name, tol in ((torch.float32, "f32", 1e-6), (torch.float64, "f64", 1e-12)):
        ch, mv = _rollout_case(g, tag, dtype)
        rs = torch.full((b * N * K, 1), 2.5, dtype=dtype) if ch[0] == 10 else None
        y = ref_layers.model_func_shift_inv(torch.tensor(X, dtype=dtype), coo, mv, (b, N, K), torch.relu, rs)
        np.testing.assert_allclose(y.numpy(), g[f"{tag}_{name}_out"], rtol=tol, atol=tol)


def test_15op_layer_restatement_and_adjacency():
    """oracle.ref_layers.shift_inv_15op_layer / network_func_15op_shift_inv_za vs the UNMODIFIED reference
    (graph.py:20-216, oracle/make_golden_15op.py), and the properties of the symmetrised-adjacency builder."""
    import types
    from oracle import ref_layers
    g = load_golden("layer15_small.npz")
    ch = [int(v) for v in g["channels"]]
    x = g["x"]
    b, N, K = x.shape[0], x.shape[1], int(g["K"])
    adj = ref_graph.get_symmetrized_adjacency(ref_graph.get_kneighbor_list(x, K))
    for name in ("row", "col", "all", "tra", "dia", "dal"):
        assert np.array_equal(adj[name], g[f"adj_{name}"]), name
    S = adj["row"].shape[0]
    assert np.array_equal(adj["tra"][adj["tra"]], np.arange(S))                       # transposition is an involution
    assert np.array_equal(adj["row"][adj["tra"]], adj["col"]) and np.array_equal(adj["col"][adj["tra"]], adj["row"])
    assert np.array_equal(adj["row"][adj["dia"]], np.arange(b * N)) and np.array_equal(adj["col"][adj["dia"]], np.arange(b * N))
    assert np.all(np.diff(adj["row"].astype(np.int64) * b * N + adj["col"]) > 0)      # row-major, no duplicates
    for dt, name, tol in ((torch.float32, "f32", 2e-5), (torch.float64, "f64", 1e-11)):
        tp = [(torch.tensor(g[f"W{li}"], dtype=dt, requires_grad=True), torch.tensor(g[f"B{li}"], dtype=dt, requires_grad=True))
              for li in range(len(ch) - 1)]
        H = torch.tensor(g["H"], dtype=dt, requires_grad=True)
        mgr = types.SimpleNamespace(channels=ch, get_layer_vars=lambda j: tp[j])
        lay0 = ref_layers.shift_inv_15op_layer(H, adj, (b, N), tp[0])
        net = ref_layers.network_func_15op_shift_inv_za(H, adj, len(ch) - 1, (b, N), torch.relu, mgr)
        loss = ((net - torch.tensor(g["tgt"], dtype=dt)) ** 2).sum(-1).mean()
        loss.backward()
        np.testing.assert_allclose(lay0.detach().numpy(), g[f"{name}_layer0"], rtol=tol, atol=tol)
        np.testing.assert_allclose(net.detach().numpy(), g[f"{name}_net"], rtol=tol, atol=tol)
        np.testing.assert_allclose(H.grad.numpy(), g[f"{name}_gH"], rtol=tol * 10, atol=tol)
        for li, (W, B) in enumerate(tp):
            np.testing.assert_allclose(W.grad.numpy(), g[f"{name}_gW{li}"], rtol=tol * 10, atol=tol)
            np.testing.assert_allclose(B.grad.numpy(), g[f"{name}_gB{li}"], rtol=tol * 10, atol=tol)
