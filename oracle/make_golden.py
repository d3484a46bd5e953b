"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by executing the UNMODIFIED
reference (/root/reference/graph.py, nn.py) through oracle/tf_shim.py.

Run in the authoring container (the only place /root/reference exists):

    python -m oracle.make_golden

The reference ships no tests/golden vectors of its own, so these files
are what pins the oracle and the CUDA path.  Library versions are recorded in every
file because the kNN arithmetic is scikit-learn's, unpinned by the reference.
"""
import hashlib
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle.load_reference import load_reference  # noqa: E402

syn = importlib.import_module("n-body_pointcloudevolution_b200.synthetic")


def versions():
    import scipy
    import sklearn
    return np.array([f"sklearn={sklearn.__version__}", f"scipy={scipy.__version__}",
                     f"numpy={np.__version__}", f"torch={torch.__version__}"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def csr_list_to_idx(A):
    N = A[0].shape[0]
    return np.stack([a.indices.reshape(N, -1) for a in A]).astype(np.int32)


def model_vars_for(params, channels, dtype):
    tp = [([torch.tensor(w, dtype=dtype, requires_grad=True) for w in Ws],
           torch.tensor(B, dtype=dtype, requires_grad=True)) for Ws, B in params]
    mv = types.SimpleNamespace(var_scope="params", channels=channels, num_layers=len(channels) - 1,
                               activation=torch.relu, get_layer_vars=lambda i: tp[i])
    return mv, tp


def save(name, **arrs):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, versions=versions(), **arrs)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def run_graph_model(rg, rnn, x, za, tgt, k, channels, dtype, keep_layers=False):
    """Reference graph-model step: kNN -> COO -> model -> loss -> backward."""
    b, N, _ = x.shape
    A = rg.get_kneighbor_list(x, k)
    coo, diag = rg.to_coo_batch_ZA_diag(A)
    params = syn.glorot_params(channels)
    mv, tp = model_vars_for(params, channels, dtype)
    out = {}
    pos = torch.tensor(x, dtype=dtype)
    zat = torch.tensor(za, dtype=dtype)
    if keep_layers:
        edges = rg.get_input_features_shift_inv_ZA(pos, zat, coo, diag, (b, N, k))
        out["edges"] = edges.detach().numpy()
        H = edges
        L = len(channels) - 1
        for li in range(L):
            last = li == L - 1
            H = rg.shift_inv_layer(H, coo, (b, N), tp[li], is_last=last)
            if not last:
                H = torch.relu(H)
            out[f"H{li}"] = H.detach().numpy()
    pred = rg.model_func_shift_inv_za(pos, coo, zat, diag, mv, (b, N, k))
    loss = rnn.loss_ZA(pred, torch.tensor(tgt, dtype=dtype))
    loss.backward()
    out["pred"] = pred.detach().numpy()
    out["loss"] = loss.detach().numpy()
    for li, (Ws, B) in enumerate(tp):
        for wi, w in enumerate(Ws):
            out[f"gW{li}_{wi}"] = w.grad.numpy()
        out[f"gB{li}"] = B.grad.numpy()
    return out, coo, diag


def main():
    rg, rnn = load_reference()

    # ------------------------------------------------------------- A. kNN, 16^3 (C1 size)
    N, k = 4096, 14
    arrs = {}
    for kind in ("uniform", "clustered"):
        for seed in (0, 1, 2):
            b = 2 if seed == 0 else 1
            x = syn.make_box(kind, b, N, seed)
            tag = f"{kind}_s{seed}"
            arrs[f"x_{tag}"] = x
            arrs[f"knl_{tag}"] = csr_list_to_idx(rg.get_kneighbor_list(x, k)).astype(np.uint16)
            arrs[f"pbc_{tag}"] = csr_list_to_idx(rg.get_pbc_kneighbors_csr(x, k, 0.1)).astype(np.uint16)
            if seed == 0:
                arrs[f"pbcself_{tag}"] = csr_list_to_idx(
                    rg.get_pbc_kneighbors_csr(x, k, 0.1, include_self=True)).astype(np.uint16)
                arrs[f"pbc03_{tag}"] = csr_list_to_idx(rg.get_pbc_kneighbors_csr(x, 8, 0.3)).astype(np.uint16)
                arrs[f"knlnoself_{tag}"] = csr_list_to_idx(
                    rg.get_kneighbor_list(x, 8, include_self=False)).astype(np.uint16)
                A = rg.get_kneighbor_list(x, k)
                coo, diag = rg.to_coo_batch_ZA_diag(A)
                arrs[f"coo_sha_{tag}"] = np.array(sha(coo))
                arrs[f"diag_{tag}"] = diag.astype(np.int64)
                assert (rg.to_coo_batch(A) == coo).all()
                assert (rg.get_indices_from_list_CSR(A) == coo[1]).all()
    save("knn_16.npz", **arrs)

    # ------------------------------------------------------------- B. kNN, 32^3: hashes + head rows
    N = 32768
    arrs = {}
    for kind in ("uniform", "clustered"):
        x = syn.make_box(kind, 1, N, 0)
        arrs[f"x_sha_{kind}"] = np.array(sha(x))
        for k in (8, 14, 32):
            knl = csr_list_to_idx(rg.get_kneighbor_list(x, k))
            arrs[f"knl_sha_{kind}_k{k}"] = np.array(sha(knl))
            arrs[f"knl_head_{kind}_k{k}"] = knl[:, :64]
        pbc = csr_list_to_idx(rg.get_pbc_kneighbors_csr(x, 14, 0.1, include_self=True))
        arrs[f"pbc_sha_{kind}_k14"] = np.array(sha(pbc))
        arrs[f"pbc_head_{kind}_k14"] = pbc[:, :64]
    save("knn_32.npz", **arrs)

    # ------------------------------------------------------------- C. lattice known answer (ties)
    xl = syn.lattice_box(1, 8, seed=0, jitter=0.0)
    A = rg.get_kneighbor_list(xl, 14)
    ind = csr_list_to_idx(A)[0]
    d2 = ((xl[0].astype(np.float64)[ind] - xl[0].astype(np.float64)[:, None, :]) ** 2).sum(-1)
    P = rg.get_pbc_kneighbors_csr(xl, 14, 0.2, include_self=True)
    save("lattice_8.npz", x=xl, knl_sorted_d2=np.sort(d2, axis=1),
         pbc_idx=csr_list_to_idx(P).astype(np.uint16))

    # ------------------------------------------------------------- D. layers, small (per-layer outputs)
    b, N, k = 2, 512, 8
    channels = syn.DEFAULT_GRAPH_CHANNELS
    x = syn.uniform_box(b, N, 0)
    za, tgt = syn.za_features(b, N, 0)
    arrs = {"x": x, "za": za, "tgt": tgt, "channels": np.array(channels), "k": np.array(k)}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        out, coo, diag = run_graph_model(rg, rnn, x, za, tgt, k, channels, dt, keep_layers=True)
        for key, v in out.items():
            arrs[f"{tag}_{key}"] = v
    arrs["coo"] = coo
    arrs["diag"] = diag
    for li, (Ws, B) in enumerate(syn.glorot_params(channels)):
        for wi, w in enumerate(Ws):
            arrs[f"W{li}_{wi}"] = w
        arrs[f"B{li}"] = B

    # single-layer checks at odd widths (k=5 -> q=7), both is_last settings, and the pooling primitive
    rng = np.random.default_rng(5)
    c = b * N * k
    H_in = rng.standard_normal((c, 5)).astype(np.float32)
    lw = [(0.3 * rng.standard_normal((5, 7))).astype(np.float32) for _ in range(4)]
    lb = (0.1 * rng.standard_normal(7)).astype(np.float32)
    arrs["odd_H_in"] = H_in
    for i, w in enumerate(lw):
        arrs[f"odd_W{i}"] = w
    arrs["odd_B"] = lb
    for last in (False, True):
        Ht = torch.tensor(H_in, requires_grad=True)
        Wt = [torch.tensor(w, requires_grad=True) for w in lw]
        Bt = torch.tensor(lb, requires_grad=True)
        o = rg.shift_inv_layer(Ht, coo, (b, N), (Wt, Bt), is_last=last)
        g = torch.tensor(np.random.default_rng(6).standard_normal(tuple(o.shape)).astype(np.float32))
        (o * g).sum().backward()
        t = "last" if last else "mid"
        arrs[f"odd_{t}_out"] = o.detach().numpy()
        arrs[f"odd_{t}_gout"] = g.numpy()
        arrs[f"odd_{t}_gH"] = Ht.grad.numpy()
        for i, w in enumerate(Wt):
            arrs[f"odd_{t}_gW{i}"] = w.grad.numpy()
        arrs[f"odd_{t}_gB"] = Bt.grad.numpy()
    for ci, nm in ((0, "row"), (1, "col"), (2, "cube")):
        arrs[f"conv_{nm}_bc"] = rg.shift_inv_conv(torch.tensor(H_in), coo[ci], b * N, True).numpy()
        arrs[f"conv_{nm}_nobc"] = rg.shift_inv_conv(torch.tensor(H_in), coo[ci], b * N, False).numpy()

    # legacy input features (graph.py:346-364, 245-275)
    X6 = np.concatenate([x, za * 3], axis=-1)
    e, n = rg.get_input_features_shift_inv(torch.tensor(X6), coo, (b, N, k))
    arrs["feat_X6"] = X6
    arrs["feat_edges"] = e.numpy()
    arrs["feat_nodes"] = n.numpy()
    rs = np.full((c, 1), 0.75, dtype=np.float32)
    arrs["feat_nodes9"] = rg.include_node_features(e, n, coo).numpy()
    arrs["feat_nodes10"] = rg.include_node_features(e, n, coo, redshift=torch.tensor(rs)).numpy()
    save("layers_small.npz", **arrs)

    # ------------------------------------------------------------- E. whole model at C1 (16^3, b=2, k=14)
    b, N, k = 2, 4096, 14
    arrs = {"channels": np.array(channels), "k": np.array(k)}
    for kind in ("uniform", "clustered"):
        x = syn.make_box(kind, b, N, 0)
        za, tgt = syn.za_features(b, N, 0)
        arrs[f"x_sha_{kind}"] = np.array(sha(x))
        arrs[f"za_sha_{kind}"] = np.array(sha(za))
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            out, coo, diag = run_graph_model(rg, rnn, x, za, tgt, k, channels, dt)
            for key, v in out.items():
                arrs[f"{kind}_{tag}_{key}"] = v if tag == "f32" or key != "pred" else v.astype(np.float64)
    save("model_16.npz", **arrs)

    # ------------------------------------------------------------- F. set model
    b, N = 2, 512
    ch = [6, 16, 32, 3]
    rng = np.random.default_rng(11)
    X = rng.standard_normal((b, N, 6)).astype(np.float32)
    Y = (0.1 * rng.standard_normal((b, N, 3))).astype(np.float32)
    arrs = {"X": X, "Y": Y, "channels": np.array(ch)}
    params = syn.glorot_params(ch, seed=123)
    for li, (Ws, B) in enumerate(params):
        for wi, w in enumerate(Ws):
            arrs[f"W{li}_{wi}"] = w
        arrs[f"B{li}"] = B
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        mv, tp = model_vars_for(params, ch, dt)
        H = torch.tensor(X, dtype=dt)
        for li in range(len(ch) - 1):
            H = rnn.set_layer(H, tp[li])
            if li < len(ch) - 2:
                H = torch.relu(H)
            arrs[f"{tag}_H{li}"] = H.detach().numpy()
        pred = rnn.model_func_set(torch.tensor(X, dtype=dt), mv)
        loss = rnn.loss_ZA(pred, torch.tensor(Y, dtype=dt))
        loss.backward()
        arrs[f"{tag}_pred"] = pred.detach().numpy()
        arrs[f"{tag}_loss"] = loss.detach().numpy()
        for li, (Ws, B) in enumerate(tp):
            arrs[f"{tag}_gW{li}"] = Ws[0].grad.numpy()
            arrs[f"{tag}_gB{li}"] = B.grad.numpy()
            assert all(w.grad is None for w in Ws[1:])  # nn.py:22: only W[0] is used
    save("set_small.npz", **arrs)

    # ------------------------------------------------------------- G. losses / readout
    rng = np.random.default_rng(21)
    pred = (rng.random((2, 512, 6)) * 1.4 - 0.2).astype(np.float32)   # spans <0 and >1
    truth = rng.random((2, 512, 6)).astype(np.float32)
    arrs = {"pred": pred, "truth": truth}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        p = torch.tensor(pred, dtype=dt, requires_grad=True)
        t = torch.tensor(truth, dtype=dt)
        ro = rnn.get_readout(p)
        arrs[f"{tag}_readout"] = ro.detach().numpy()
        arrs[f"{tag}_readout3"] = rnn.get_readout(p[..., :3]).detach().numpy()
        arrs[f"{tag}_pbd"] = rnn.periodic_boundary_dist(ro, t).detach().numpy()
        l1 = rnn.pbc_loss(ro, t)
        l1.backward()
        arrs[f"{tag}_pbc_loss"] = l1.detach().numpy()
        arrs[f"{tag}_pbc_loss_gpred"] = p.grad.numpy().copy()
        arrs[f"{tag}_pbc_loss_unscaled"] = rnn.pbc_loss(ro, t, scale_error=False).detach().numpy()
        p2 = torch.tensor(pred[..., :3], dtype=dt, requires_grad=True)
        l2 = rnn.loss_ZA(p2, t[..., :3])
        l2.backward()
        arrs[f"{tag}_loss_za"] = l2.detach().numpy()
        arrs[f"{tag}_loss_za_gpred"] = p2.grad.numpy()
    save("losses.npz", **arrs)


if __name__ == "__main__":
    main()
