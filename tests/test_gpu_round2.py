"""GPU parity tests added in round 2 (run on the B200 box, through the reference-named API -> C ABI):
  * BASELINE config 2 ITSELF (32^3, batch 8, k = 14, [3,32,16,3], default tf32x3 math) and one 64^3 sample against a float64
    golden produced by the UNMODIFIED reference (tests/golden/headline_32.npz, oracle/make_golden_r2.py);
  * the set model at the reference's default widths (utils.py:165), b = 8, N = 32^3;
  * pad_cube_boundaries / get_pcube_csr (graph.py:827-894), bit-exact;
  * experiment.py's attention / residual / batch-norm net (experiment.py:83-157);
  * train.py on a ZA_###.npy file, the periodic-kNN unit-box check, M = 1, 2-GPU data parallelism.
Tolerances are written next to each assertion."""
import hashlib
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_graph, ref_layers

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# =============================================================================== BASELINE config 2, numerically
@pytest.mark.parametrize("kind,seed,b,side", [("uniform", 0, 8, 32), ("clustered", 0, 8, 32), ("uniform", 1, 2, 32),
                                              ("clustered", 1, 2, 32), ("uniform", 2, 2, 32), ("clustered", 2, 2, 32),
                                              ("uniform", 0, 1, 64)])
def test_headline_config_golden(nb, syn, kind, seed, b, side):
    """kNN lists bit-exact (sha of COO[1] / diagonals); prediction rows rtol 5e-5 / atol 5e-6, loss rtol 2e-5, every one of
    the 2675 gradients rtol 5e-4 (atol 5e-5 of the tensor's max) against the float64 reference run."""
    g = load_golden("headline_32.npz")
    ch, k, N = [int(v) for v in g["channels"]], int(g["k"]), side ** 3
    tag = f"{kind}_s{seed}_b{b}_n{side}"
    x = syn.make_box(kind, b, N, seed)
    za, tgt = syn.za_features(b, N, seed)
    assert sha(x) == str(g[f"{tag}_x_sha"]) and sha(za) == str(g[f"{tag}_za_sha"])
    store = nb.train_utils.ParamStore(ch, device=DEV)
    store.load_numpy(syn.glorot_params(ch))
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
    xt = torch.tensor(x, device=DEV)
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(xt, k))
    assert sha(coo[1].cpu().numpy()) == str(g[f"{tag}_col_sha"])
    assert sha(diag.cpu().numpy()) == str(g[f"{tag}_diag_sha"])
    pred = nb.graph.model_func_shift_inv_za(xt, coo, torch.tensor(za, device=DEV), diag, mv, (b, N, k))
    loss = nb.nn.loss_ZA(pred, torch.tensor(tgt, device=DEV))
    store.zero_grad()
    loss.backward()
    rows = g[f"{tag}_rows"]
    np.testing.assert_allclose(pred.detach().reshape(b * N, -1)[torch.tensor(rows, device=DEV)].cpu().numpy(), g[f"{tag}_pred"],
                               rtol=5e-5, atol=5e-6)
    np.testing.assert_allclose(loss.item(), float(g[f"{tag}_loss"]), rtol=2e-5)
    for li in range(len(ch) - 1):
        W, B = store.get_layer_vars(li)
        for wi in range(4):
            ref = g[f"{tag}_gW{li}_{wi}"]
            np.testing.assert_allclose(W.grad[wi].cpu().numpy(), ref, rtol=5e-4, atol=5e-5 * float(np.abs(ref).max()))
        ref = g[f"{tag}_gB{li}"]
        np.testing.assert_allclose(B.grad.cpu().numpy(), ref, rtol=5e-4, atol=5e-5 * float(np.abs(ref).max()) + 1e-12)


# =============================================================================== set model, default widths
def test_set_model_default_widths_golden(nb, syn):
    """nn.model_func_set at channels [6,64,128,128,256,64,128,16,3] (utils.py:165), b = 8, N = 32^3, against the float64
    reference run: prediction rows rtol 2e-4 / atol 2e-5 of max, loss rtol 5e-5, gradients rtol 1e-3 / atol 4e-3 of max."""
    g = load_golden("set_default.npz")
    ch, b, N = [int(v) for v in g["channels"]], int(g["b"]), int(g["N"])
    rng = np.random.default_rng(31)
    X = rng.standard_normal((b, N, 6)).astype(np.float32)
    Y = (0.1 * rng.standard_normal((b, N, 3))).astype(np.float32)
    assert sha(X) == str(g["x_sha"])
    store = nb.train_utils.ParamStore(ch, device=DEV)
    store.load_numpy(syn.glorot_params(ch, seed=321))
    mv = store.model_vars(torch.relu)
    pred = nb.nn.model_func_set(torch.tensor(X, device=DEV), mv)
    loss = nb.nn.loss_ZA(pred, torch.tensor(Y, device=DEV))
    store.zero_grad()
    loss.backward()
    ref = g["pred"]
    got = pred.detach().reshape(b * N, -1)[torch.tensor(g["rows"], device=DEV)].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5 * float(np.abs(ref).max()))
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=5e-5)
    for li in range(len(ch) - 1):
        W, B = store.get_layer_vars(li)
        # every gradient entry is a sum of 262 144 float32 terms of both signs accumulated in FP32: the absolute tolerance
        # is a fraction of the tensor's largest entry (4e-3 of max; the relative one, 1e-3, covers the large entries)
        ref = g[f"gW{li}"]
        np.testing.assert_allclose(W.grad[0].cpu().numpy(), ref, rtol=1e-3, atol=4e-3 * float(np.abs(ref).max()))
        assert float(W.grad[1:].abs().max()) == 0.0                      # nn.py:22: only W[0] is used
        ref = g[f"gB{li}"]
        np.testing.assert_allclose(B.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-3 * float(np.abs(ref).max()))


@pytest.mark.parametrize("mode", ["tf32x3", "fp32", "tf32"])
@pytest.mark.parametrize("b,N,ch", [(3, 1000, [6, 64, 128, 32, 3]), (2, 333, [6, 32, 256, 64, 16, 3]), (1, 4096, [6, 128, 128, 3]),
                                    (2, 50, [6, 64, 32, 3]), (5, 7, [6, 32, 32, 3])])
def test_set_model_tensor_core_shapes_vs_oracle(nb, mode, b, N, ch):
    """The tcgen05 set-layer kernels (row GEMM with in-place mean subtraction, MN-major dW GEMM, fused input mask) on ragged
    shapes: N not a multiple of the 128-row tile (tiles straddle samples), two N tiles (k=32 -> q=256 backward), the narrow
    16-wide layer - forward, loss and all gradients against the float64 oracle (tf32x3 / fp32: rtol 2e-4, atol 2e-5 of max;
    tf32 single pass: 6e-2 in Frobenius norm)."""
    rng = np.random.default_rng(b * 100 + N)
    X = rng.standard_normal((b, N, ch[0])).astype(np.float32)
    X[..., :3] += 3.0                                                     # non-zero column means
    Y = (0.1 * rng.standard_normal((b, N, ch[-1]))).astype(np.float32)
    params = [([(rng.standard_normal((kk, qq)) * np.sqrt(2.0 / (kk + qq))).astype(np.float32)], (0.05 * rng.standard_normal(qq)).astype(np.float32))
              for kk, qq in zip(ch[:-1], ch[1:])]
    tp = [([torch.tensor(Ws[0], device=DEV, requires_grad=True)], torch.tensor(B, device=DEV, requires_grad=True)) for Ws, B in params]
    rp = [([torch.tensor(Ws[0], dtype=torch.float64, requires_grad=True)], torch.tensor(B, dtype=torch.float64, requires_grad=True)) for Ws, B in params]
    mv = types.SimpleNamespace(num_layers=len(ch) - 1, activation=torch.relu, get_layer_vars=lambda i: tp[i])
    rmv = types.SimpleNamespace(num_layers=len(ch) - 1, activation=torch.relu, get_layer_vars=lambda i: rp[i])
    old = nb.get_math_mode()
    nb.set_math_mode(mode)
    try:
        Xt = torch.tensor(X, device=DEV, requires_grad=True)
        pred = nb.nn.model_func_set(Xt, mv)
        loss = nb.nn.loss_ZA(pred, torch.tensor(Y, device=DEV))
        loss.backward()
        torch.cuda.synchronize()
    finally:
        nb.set_math_mode(old)
    Xr = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    rpred = ref_layers.model_func_set(Xr, rmv)
    rloss = ref_layers.loss_ZA(rpred, torch.tensor(Y, dtype=torch.float64))
    rloss.backward()
    tol = 2e-2 if mode == "tf32" else 2e-5

    def close(got, ref, what):
        ref = ref.detach().numpy()
        got = got.detach().cpu().numpy()
        if mode == "tf32":    # 2^-11 operand rounding flips ReLU masks of near-zero pre-activations: compare in norm
            err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
            assert err <= 3 * tol, (what, err)
            return
        np.testing.assert_allclose(got, ref, rtol=10 * tol, atol=tol * float(np.abs(ref).max()) + 1e-12, err_msg=what)
    close(pred, rpred, "pred")
    np.testing.assert_allclose(loss.item(), rloss.item(), rtol=1e-2 if mode == "tf32" else 2e-5)
    close(Xt.grad, Xr.grad, "dX")
    for li in range(len(ch) - 1):
        close(tp[li][0][0].grad, rp[li][0][0].grad, f"dW{li}")
        close(tp[li][1].grad, rp[li][1].grad, f"dB{li}")


# =============================================================================== padded cube
@pytest.mark.parametrize("thr", [0.1, 0.3])
def test_pad_cube_boundaries_golden(nb, thr):
    """graph.pad_cube_boundaries / get_pcube_csr / get_pcube_adjacency_list (graph.py:827-894): padded cloud (float64),
    idx_map and the neighbour lists bit-exact against the unmodified reference."""
    g = load_golden("pad_cube.npz")
    t = str(thr).replace(".", "")
    x = g["x"]
    padded, idx_map = nb.graph.pad_cube_boundaries(x, thr)
    assert padded.dtype == torch.float64 and idx_map.dtype == torch.int64
    assert np.array_equal(padded.cpu().numpy(), g[f"padded_{t}"])
    assert np.array_equal(idx_map.cpu().numpy(), g[f"idx_map_{t}"])
    csr = nb.graph.get_pcube_csr(padded, idx_map, x.shape[0], 8, include_self=False)
    assert csr.shape == (x.shape[0], padded.shape[0])
    assert np.array_equal(csr.toarray_indices().cpu().numpy(), g[f"pcube_csr_{t}"])
    assert np.array_equal(nb.graph.get_pcube_adjacency_list(padded, idx_map, x.shape[0], 8).cpu().numpy(), g[f"pcube_adj_{t}"])
    with pytest.raises(TypeError):
        nb.graph.get_pcube_csr(padded.cpu().numpy(), idx_map, x.shape[0], 8)
    # no boundary particles at all
    p0, m0 = nb.graph.pad_cube_boundaries(np.full((5, 3), 0.5, dtype=np.float32), 0.1)
    assert tuple(p0.shape) == (5, 3) and m0.numel() == 0


def test_periodic_knn_flags_out_of_box(nb):
    x = np.random.default_rng(0).random((1, 500, 3)).astype(np.float32)
    assert nb.graph.get_pbc_kneighbors_csr(x, 4, 0.1)[0].check()
    x[0, 17, 1] = 1.25
    x[0, 99, 2] = -0.01
    with pytest.raises(ValueError, match="2 particles"):
        nb.graph.get_pbc_kneighbors_csr(x, 4, 0.1)[0].check()


# =============================================================================== experiment.py
def test_experiment_net_golden(nb):
    """experiment.py's attn_layer / res_layer / net_fwd (experiment.py:83-157, executed unmodified to make the golden):
    forward rtol 1e-4 / atol 1e-5 vs the float32 run, loss rtol 1e-4 and gradients rtol 2e-3 / atol 1e-3 of max vs float64."""
    import experiment as ex
    g = load_golden("experiment.npz")
    ch = [int(v) for v in g["channels"]]
    nl = len(ch) - 1
    ex.init_model(ch, device=DEV)
    with torch.no_grad():
        for name, mod in (("Wf", ex.Wf), ("Wg", ex.Wg), ("Wh", ex.Wh), ("Rset", ex.Rset), ("Bset", ex.Bset)):
            for i in range(nl):
                mod[i].copy_(torch.tensor(g[f"{name}{i}"]))
        for i in range(nl - 1):
            ex.Gamma[i].copy_(torch.tensor(g[f"gamma{i}"]))
            ex.Beta[i].copy_(torch.tensor(g[f"beta{i}"]))
    X, Y = torch.tensor(g["X"], device=DEV), torch.tensor(g["Y"], device=DEV)
    ex.X_in = X
    np.testing.assert_allclose(ex.attn_layer(X, 0).detach().cpu().numpy(), g["f32_attn0"], rtol=1e-4, atol=1e-5)
    pred = ex.net_fwd(X)
    err = ex.loss(pred, Y)
    ex.params.zero_grad()
    err.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["f32_pred"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(err.item(), float(g["f64_loss"]), rtol=1e-4)
    for name, mod in (("Wf", ex.Wf), ("Wg", ex.Wg), ("Wh", ex.Wh), ("Rset", ex.Rset), ("Bset", ex.Bset), ("gamma", ex.Gamma),
                      ("beta", ex.Beta)):
        for i, p in enumerate(mod):
            ref = g[f"f64_g{name}{i}"]
            np.testing.assert_allclose(p.grad.cpu().numpy(), ref, rtol=2e-3, atol=1e-3 * float(np.abs(ref).max()) + 1e-7,   # + 1e-7: gradients that are exactly 0 in exact arithmetic (beta under a mean subtraction)
                                       err_msg=f"{name}{i}")


def test_experiment_entry_point_runs(tmp_path):
    """experiment.py's command line (-i -b -n) on the synthetic stand-in: trains, validates, writes test_cubes / test_error."""
    import experiment as ex
    hist, preds, train_hist = ex.main(["-i", "20", "-b", "2", "-n", "E", "--side", "8", "--num_test", "4", "--channels", "6", "16", "16", "3",
                                       "--checkpoint", "10", "--out_dir", str(tmp_path), "--data_dir", str(tmp_path / "none")])
    assert hist.shape == (2,) and np.isfinite(hist).all() and train_hist.shape == (2,) and np.isfinite(train_hist).all()
    assert preds.shape == (2, 4, 512, 3)
    assert np.load(tmp_path / "E" / "test_cubes.npy").shape == (2, 4, 512, 3) and (tmp_path / "E" / "test_error.npy").is_file()


# =============================================================================== train.py on the reference's file format
def test_train_py_reads_za_npy_and_writes_prediction_file(tmp_path):
    """A ZA_001.npy with the reference's layout (S, side, side, side, 19) (utils.py:538-544, 594-621) is loaded through
    train.py; Results/X_0_prediction.npy has shape (2, num_test, N, 3) with the truth in [0] (train.py:131-132)."""
    import train
    rng = np.random.default_rng(3)
    S, side = 112, 8
    data = rng.standard_normal((S, side, side, side, 19)).astype(np.float32)
    np.save(tmp_path / "ZA_001.npy", data)
    err = train.main(["-i", "4", "-b", "2", "-t", "4", "-k", "6", "-c", "3", "16", "3", "-d", "0", "--data_dir", str(tmp_path),
                      "--out_dir", str(tmp_path), "-n", "z", "--checkpoint", "2"])
    assert err.shape == (2,) and np.isfinite(err).all()
    preds = np.load(tmp_path / "z" / "Results" / "X_0_prediction.npy")
    assert preds.shape == (2, 4, side ** 3, 3)
    X = train.Dataset.load_data(str(tmp_path / "ZA_001.npy"))
    perm = np.random.RandomState(train.DATASET_SEED).permutation(S)
    assert np.array_equal(preds[0], X[perm][-4:, :, 6:])              # truth = FPM - ZA of the seeded test split
    assert np.load(tmp_path / "z" / "Results" / "error_test.npy").shape == (2,)


def test_train_py_forward_matches_oracle(nb):
    """train.py's graph-model wiring (ONE position set, q + ZA displacement, for the kNN graph and for the edge features,
    graph.py:289-343) against the float64 oracle on the same minibatch."""
    import train
    args = train.parser().parse_args(["-k", "6", "-c", "3", "16", "3", "--side", "8"])
    store = nb.train_utils.ParamStore(args.channels, seed=5, device=DEV)
    fwd = train.build_step(args, store, torch.device(DEV))
    batch = train.Dataset.synthetic(2, 8)
    pred, loss = fwd(batch)
    b, N, K = 2, 512, 6
    pos = (batch[..., :3] + batch[..., 3:6]).astype(np.float32)
    rA = ref_graph.get_kneighbor_list(pos, K, backend="exact")
    rcoo, rdiag = ref_graph.to_coo_batch_ZA_diag(rA)
    rp = [([w.detach().double().cpu() for w in store.get_layer_vars(i)[0]], store.get_layer_vars(i)[1].detach().double().cpu())
          for i in range(store.num_layers)]
    rmv = types.SimpleNamespace(channels=args.channels, get_layer_vars=lambda i: rp[i])
    rpred = ref_layers.model_func_shift_inv_za(torch.tensor(pos, dtype=torch.float64), rcoo, torch.tensor(batch[..., 3:6], dtype=torch.float64),
                                               rdiag, rmv, (b, N, K))
    rloss = ref_layers.loss_ZA(rpred, torch.tensor(batch[..., 6:], dtype=torch.float64))
    np.testing.assert_allclose(pred.detach().cpu().numpy(), rpred.numpy(), rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(loss.item(), rloss.item(), rtol=2e-5)


# =============================================================================== virtual first layer (graph_layer_vin.cuh)
@pytest.mark.parametrize("b,N,M,ch", [(2, 4096, 14, [3, 32, 16, 3]), (3, 601, 10, [3, 16, 32, 3]), (1, 1000, 7, [3, 64, 16, 16, 3])])
def test_virtual_first_layer_is_bit_identical(nb, syn, b, N, M, ch):
    """The recomputing kernels (layer 1's output never materialised: virtual-input pooling, tcgen05 forward and backward with
    generator warps) against the materialising path on the same inputs: prediction, loss and every gradient BIT-identical
    (one shared expression produces the rows in every consumer)."""
    x = torch.tensor(syn.make_box("clustered", b, N, 3), device=DEV)
    za, tgt = (torch.tensor(t, device=DEV) for t in syn.za_features(b, N, 3))
    outs = []
    for virtual in (False, True):
        old = nb.graph.set_virtual_first_layer(virtual)
        try:
            store = nb.train_utils.ParamStore(ch, device=DEV)
            store.load_numpy(syn.glorot_params(ch))
            mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
            coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, M))
            n0 = nb._lib.launch_count()
            pred = nb.graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, M))
            loss = nb.nn.loss_ZA(pred, tgt)
            store.zero_grad()
            loss.backward()
            torch.cuda.synchronize()
            outs.append((pred.detach().clone(), float(loss.detach()), store.flat_grad.clone(), nb._lib.launch_count() - n0))
        finally:
            nb.graph.set_virtual_first_layer(old)
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] and torch.equal(outs[0][2], outs[1][2])
    assert outs[1][3] == outs[0][3] - 1                       # one launch less: the first layer's edge kernel


# =============================================================================== row-pool hand-over between layers
@pytest.mark.parametrize("b,N,M,ch", [(2, 4096, 14, [3, 32, 16, 3]), (3, 601, 9, [3, 16, 32, 3]), (1, 1000, 1, [3, 64, 16, 16, 3]),
                                      (2, 777, 5, [3, 32, 3]), (1, 500, 14, [3, 16, 3])])
def test_rowpool_chain_is_bit_identical(nb, syn, b, N, M, ch):
    """nbpc_graph_layer_fwd_rp / _bwd_rp: the first layer's edge kernel emits the row means the second layer pools first, the last
    layer's backward edge kernel emits the row sums the previous layer's backward pools first (same summation order) -
    prediction, loss and every gradient BIT-identical to the plain entry points, which read the edge tensors twice."""
    x = torch.tensor(syn.make_box("clustered", b, N, 5), device=DEV)
    za, tgt = (torch.tensor(t, device=DEV) for t in syn.za_features(b, N, 5))
    outs = []
    for chained in (False, True):
        old = nb.graph.set_rowpool_chain(chained)
        try:
            store = nb.train_utils.ParamStore(ch, device=DEV)
            store.load_numpy(syn.glorot_params(ch))
            mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
            coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, M))
            nb._lib.prof_enable(True)
            pred = nb.graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, M))
            loss = nb.nn.loss_ZA(pred, tgt)
            store.zero_grad()
            loss.backward()
            torch.cuda.synchronize()
            names = set(nb._lib.prof_report())
            nb._lib.prof_enable(False)
            outs.append((pred.detach().clone(), float(loss.detach()), store.flat_grad.clone(), names))
        finally:
            nb._lib.prof_enable(False)
            nb.graph.set_rowpool_chain(old)
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] and torch.equal(outs[0][2], outs[1][2])
    assert not any("rowpool" in n or "rowsum" in n or "colonly" in n for n in outs[0][3])
    if len(ch) > 3:                                       # a hidden layer exists: both hand-overs must have run
        assert any(n.startswith("glk3_edge_out_rowpool_kernel") for n in outs[1][3]), sorted(outs[1][3])
        assert any(n.startswith("gln_pool_colonly_kernel") for n in outs[1][3])
        assert any(n.startswith("glf_last_edge_in_rowsum_kernel") for n in outs[1][3])
        assert any(n.startswith("gln_bwd_pool_colonly_kernel") for n in outs[1][3])


# =============================================================================== M = 1 (ADVICE: magic divisor overflow)
@pytest.mark.parametrize("ch", [[3, 16, 3], [3, 32, 16, 3]])
def test_graph_model_single_neighbour(nb, ch):
    b, N, M = 2, 300, 1
    rng = np.random.default_rng(12)
    x = rng.random((b, N, 3)).astype(np.float32)
    za = (0.01 * rng.standard_normal((b, N, 3))).astype(np.float32)
    tgt = (0.01 * rng.standard_normal((b, N, 3))).astype(np.float32)
    params = [([(rng.standard_normal((kk, qq)) * np.sqrt(2.0 / (kk + qq))).astype(np.float32) for _ in range(4)],
               (0.01 * rng.standard_normal(qq)).astype(np.float32)) for kk, qq in zip(ch[:-1], ch[1:])]
    tp = [([torch.tensor(w, device=DEV, requires_grad=True) for w in Ws], torch.tensor(B, device=DEV, requires_grad=True)) for Ws, B in params]
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda i: tp[i])
    xt = torch.tensor(x, device=DEV)
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(xt, M))
    pred = nb.graph.model_func_shift_inv_za(xt, coo, torch.tensor(za, device=DEV), diag, mv, (b, N, M))
    loss = nb.nn.loss_ZA(pred, torch.tensor(tgt, device=DEV))
    loss.backward()
    rcoo, rdiag = ref_graph.to_coo_batch_ZA_diag(ref_graph.get_kneighbor_list(x, M, backend="exact"))
    rp = [([torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws], torch.tensor(B, dtype=torch.float64, requires_grad=True))
          for Ws, B in params]
    rmv = types.SimpleNamespace(channels=ch, get_layer_vars=lambda i: rp[i])
    rpred = ref_layers.model_func_shift_inv_za(torch.tensor(x, dtype=torch.float64), rcoo, torch.tensor(za, dtype=torch.float64), rdiag, rmv,
                                               (b, N, M))
    rloss = ref_layers.loss_ZA(rpred, torch.tensor(tgt, dtype=torch.float64))
    rloss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), rpred.detach().numpy(), rtol=5e-5, atol=5e-6)
    for li in range(len(ch) - 1):
        for wi in range(4):
            ref = rp[li][0][wi].grad.numpy()
            np.testing.assert_allclose(tp[li][0][wi].grad.cpu().numpy(), ref, rtol=5e-4, atol=5e-5 * float(np.abs(ref).max()) + 1e-12)


def test_layer_rejects_mismatched_batch_factorisation(nb):
    x = np.random.default_rng(1).random((2, 64, 3)).astype(np.float32)
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, 4))
    H = torch.randn(2 * 64 * 4, 3, device=DEV)
    W, B = [torch.randn(3, 16, device=DEV) for _ in range(4)], torch.zeros(16, device=DEV)
    nb.graph.shift_inv_layer(H, coo, (2, 64), (W, B))
    with pytest.raises(ValueError):
        nb.graph.shift_inv_layer(H, coo, (4, 32), (W, B))       # same b*N, different per-sample pool


# =============================================================================== data parallelism on real GPUs
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_dp2_gradients_match_global_batch(tmp_path):
    """2 ranks x 2 samples over NCCL == 1 GPU x 4 samples: the all-reduced flat gradient (scaled by 1/world) equals the
    gradient of the global-batch mean loss (rtol 1e-5, atol 1e-6 of max: the loss mean is associated differently), and the
    parameters after 3 Adam steps agree to 1e-5."""
    out = tmp_path / "dp.npz"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29731", os.path.join(ROOT, "tests", "dp_worker.py"), str(out)], check=True, env=env, timeout=600)
    r = np.load(out)
    scale = float(np.abs(r["g_1"]).max())
    np.testing.assert_allclose(r["g_dp"], r["g_1"], rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(r["p_dp"], r["p_1"], rtol=1e-5, atol=1e-4)
    # eager DP loop (adam.step, host-side step count) vs the CUDA-graph form with the all-reduce captured (device-side count)
    np.testing.assert_allclose(r["p_pipe"], r["p_dp"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(r["p_over"], r["p_pipe"])       # two-stream form (train_utils.OverlappedStep): same updates
