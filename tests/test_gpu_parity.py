"""GPU parity tests (run on the B200 box): the CUDA path, called through the reference-named Python
API (which goes through the C ABI of libnbpc.so), against
  * golden vectors produced by the UNMODIFIED reference (tests/golden, oracle/make_golden.py),
  * the CPU oracle (oracle/) on seeded inputs,
  * size-independent properties at the full BASELINE sizes.
Tolerances: kNN indices / COO / diag / float64 distances / input features / readout: bit-exact.
Layer outputs: rtol 2e-5, atol 2e-6 vs the float32 reference run (different association of the same
sums, SURVEY.md §7 hard part 7); gradients: rtol 2e-4 vs the float64 reference run."""
import hashlib
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_graph, ref_layers
from oracle.knn_exact import knn_exact

pytestmark = pytest.mark.gpu

DEV = "cuda"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def idx_of(A):
    return torch.stack([a.indices.view(a.N, a.M) for a in A]).cpu().numpy()


def cuda_model_vars(g, channels, n_w=4):
    tp = []
    for li in range(len(channels) - 1):
        Ws = [torch.tensor(g[f"W{li}_{wi}"], device=DEV, requires_grad=True) for wi in range(n_w)]
        B = torch.tensor(g[f"B{li}"], device=DEV, requires_grad=True)
        tp.append((Ws, B))
    mv = types.SimpleNamespace(var_scope="params", channels=list(channels), num_layers=len(channels) - 1,
                               activation=torch.relu, get_layer_vars=lambda i: tp[i])
    return mv, tp


# =============================================================================== kNN
@pytest.fixture(params=["thread", "warp"])
def knn_kernel(nb, request):
    """Both query kernels must produce the reference's answer bit for bit (the default picks one by k)."""
    nb.set_knn_kernel(request.param)
    yield request.param
    nb.set_knn_kernel("auto")


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_knn_16_golden(nb, kind, knn_kernel):
    g = load_golden("knn_16.npz")
    for seed in (0, 1, 2):
        tag = f"{kind}_s{seed}"
        x = g[f"x_{tag}"]
        A = nb.graph.get_kneighbor_list(x, 14)
        assert A[0].shape == (4096, 4096) and A[0].indices.dtype == torch.int32
        assert np.array_equal(idx_of(A), g[f"knl_{tag}"])
        assert np.array_equal(idx_of(nb.graph.get_pbc_kneighbors_csr(x, 14, 0.1)), g[f"pbc_{tag}"])
        if seed == 0:
            P = nb.graph.get_pbc_kneighbors_csr(x, 14, 0.1, include_self=True)
            assert np.array_equal(idx_of(P), g[f"pbcself_{tag}"])
            padded, _ = ref_graph.pad_cube_boundaries(x[0], 0.1)
            assert P[0].shape == (4096, padded.shape[0])
            assert np.array_equal(idx_of(nb.graph.get_pbc_kneighbors_csr(x, 8, 0.3)), g[f"pbc03_{tag}"])
            assert np.array_equal(idx_of(nb.graph.get_kneighbor_list(x, 8, include_self=False)), g[f"knlnoself_{tag}"])
            # offset_idx (graph.py:710-711)
            Ao = nb.graph.get_kneighbor_list(x, 14, offset_idx=True)
            assert np.array_equal(Ao[1].indices.cpu().numpy(), g[f"knl_{tag}"][1].reshape(-1).astype(np.int64) + 4096)


def test_knn_32_golden_hashes(nb, syn, knn_kernel):
    g = load_golden("knn_32.npz")
    for kind in ("uniform", "clustered"):
        x = syn.make_box(kind, 1, 32768, 0)
        assert sha(x) == str(g[f"x_sha_{kind}"])
        for k in (8, 14, 32):
            knl = idx_of(nb.graph.get_kneighbor_list(x, k))
            assert np.array_equal(knl[:, :64], g[f"knl_head_{kind}_k{k}"])
            assert sha(knl.astype(np.int32)) == str(g[f"knl_sha_{kind}_k{k}"]), (kind, k)
        pbc = idx_of(nb.graph.get_pbc_kneighbors_csr(x, 14, 0.1, include_self=True))
        assert sha(pbc.astype(np.int32)) == str(g[f"pbc_sha_{kind}_k14"]), kind


@pytest.mark.parametrize("N,k,periodic,thr,inc", [
    (5000, 1, False, 0.0, True), (5000, 64, False, 0.0, True), (777, 33, True, 0.5, True),
    (3000, 14, True, 0.05, False), (64, 63, False, 0.0, False), (9, 8, True, 0.5, False), (1, 1, False, 0.0, True),
])
def test_knn_vs_exact_oracle(nb, N, k, periodic, thr, inc, knn_kernel):
    x = np.random.default_rng(N + k).random((2, N, 3)).astype(np.float32)
    idx, d2, _ = nb.ops.knn(torch.tensor(x, device=DEV), k, periodic, thr, inc, 0, True)
    for s in range(2):
        cloud = ref_graph.pad_cube_boundaries(x[s], thr) if periodic else (x[s].astype(np.float64), None)
        ridx, rd2 = knn_exact(cloud[0], N, k, inc, return_d2=True)
        if periodic and len(cloud[1]):
            ridx = np.where(ridx >= N, cloud[1][np.maximum(ridx - N, 0)], ridx)
        assert np.array_equal(d2[s].cpu().numpy(), rd2)          # bit-exact float64 distances
        assert np.array_equal(idx[s].cpu().numpy(), ridx)


def test_knn_arbitrary_box_strided_and_ties(nb, knn_kernel):
    rng = np.random.default_rng(4)
    X = (rng.random((2, 3000, 9)) * 128 - 5).astype(np.float32)    # reference layout (b,N,9), Mpc/h coordinates
    A = nb.graph.get_kneighbor_list(torch.tensor(X, device=DEV), 14)
    ref = ref_graph.get_kneighbor_list(X, 14, backend="exact")
    assert np.array_equal(idx_of(A), np.stack([r.indices.reshape(3000, 14) for r in ref]))
    g = load_golden("lattice_8.npz")       # tie-heavy lattice: distances are defined, order is (d2, index)
    idx, d2, _ = nb.ops.knn(torch.tensor(g["x"], device=DEV), 14, False, 0.0, True, 0, True)
    assert np.array_equal(d2[0].cpu().numpy(), g["knl_sorted_d2"])
    assert np.array_equal(idx[0].cpu().numpy(), knn_exact(g["x"][0].astype(np.float64), 512, 14, True))


@pytest.mark.parametrize("case", ["uniform", "clustered", "lattice", "duplicates", "far_box", "tiny", "dense_blob"])
@pytest.mark.parametrize("k", [5, 14, 17, 32])
def test_knn_warp_kernel_equals_thread_kernel(nb, syn, case, k):
    """The warp-per-query kernel (FP32 keys + FP64 re-evaluation of the survivors) against the thread-per-query kernel:
    indices AND float64 distances identical, open and periodic, both output orders.  Lattice / duplicate inputs have
    more than 32 candidates within the FP32 margin of the k-th distance: those queries take the hand-over path;
    dense_blob exercises the long-run (no staging) path and the keep-list compaction."""
    rng = np.random.default_rng(11)
    if case in ("uniform", "clustered"):
        x = syn.make_box(case, 2, 6000, 3)
    elif case == "lattice":
        g = (np.arange(12, dtype=np.float32) + 0.5) / 12
        x = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3).repeat(2, 0)
    elif case == "duplicates":
        x = rng.random((2, 40, 3)).astype(np.float32).repeat(50, 1)          # every point 50 times
    elif case == "far_box":
        x = (rng.random((2, 3000, 3)) * 1000 + 5e4).astype(np.float32)       # coarse float32 grid far from the origin
    elif case == "tiny":
        x = rng.random((3, 40, 3)).astype(np.float32)
    else:
        x = rng.random((1, 8000, 3)).astype(np.float32)
        x[0, :3000] = 0.5 + 1e-3 * rng.standard_normal((3000, 3)).astype(np.float32)
    xt = torch.tensor(np.ascontiguousarray(x), device=DEV)
    for periodic, thr in ((False, 0.0), (True, 0.1)):
        if periodic and case == "far_box":
            continue                                                          # periodic mode is defined on the unit box
        for order in (0, 1):
            out = {}
            for kern in ("thread", "warp"):
                nb.set_knn_kernel(kern)
                try:
                    idx, d2, _ = nb.ops.knn(xt, k, periodic, thr, order == 0, order, True)
                finally:
                    nb.set_knn_kernel("auto")
                out[kern] = (idx.cpu().numpy(), d2.cpu().numpy())
            assert np.array_equal(out["thread"][1], out["warp"][1]), (case, k, periodic, order)
            assert np.array_equal(out["thread"][0], out["warp"][0]), (case, k, periodic, order)


def test_knn_kernel_selection(nb):
    assert nb.get_knn_kernel() == "auto"
    nb.set_knn_kernel("warp")
    assert nb.get_knn_kernel() == "warp"
    nb.set_knn_kernel("auto")
    with pytest.raises(RuntimeError, match="unknown kernel"):
        nb.set_knn_kernel(7)


def test_knn_errors(nb):
    x = torch.rand(1, 10, 3, device=DEV)
    with pytest.raises(RuntimeError, match="k exceeds"):
        nb.graph.get_kneighbor_list(x, 11)
    with pytest.raises(RuntimeError, match="k exceeds"):
        nb.graph.get_kneighbor_list(x, 10, include_self=False)
    with pytest.raises(RuntimeError, match="k must be"):
        nb.ops.knn(torch.rand(1, 100, 3, device=DEV), 65, False, 0.0, True, 0, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.ops.knn(torch.rand(1, 100, 3), 4, False, 0.0, True, 0, False)


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_knn_128_cubed_properties(nb, syn, kind, knn_kernel):
    """BASELINE size (2 097 152 particles, k=14, periodic): properties + brute-force spot check."""
    N, k = 128 ** 3, 14
    x = torch.tensor(syn.make_box(kind, 1, N, 0), device=DEV)
    idx, d2, _ = nb.ops.knn(x, k, True, 0.5, True, 0, True)
    idx, d2 = idx[0], d2[0]
    assert bool((idx[:, 0] == torch.arange(N, device=DEV)).all()) and bool((d2[:, 0] == 0).all())
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    # recompute the reported distances from the indices (minimum image, float64)
    xd = x[0].double()
    diff = xd[idx.long()] - xd[:, None, :]
    diff = diff - torch.round(diff)
    assert torch.allclose((diff * diff).sum(-1), d2, rtol=1e-12, atol=1e-18)
    # brute force on 256 random queries
    q = torch.randperm(N, device=DEV)[:256]
    for chunk in q.split(32):
        dd = xd[None, :, :] - xd[chunk][:, None, :]
        dd = dd - torch.round(dd)
        dist = (dd * dd).sum(-1)
        ref = torch.topk(dist, k, dim=1, largest=False).values
        assert torch.allclose(ref, d2[chunk], rtol=1e-12, atol=1e-18)


# =============================================================================== adjacency + features
def test_adjacency_golden(nb):
    g = load_golden("knn_16.npz")
    for kind in ("uniform", "clustered"):
        tag = f"{kind}_s0"
        A = nb.graph.get_kneighbor_list(g[f"x_{tag}"], 14)
        coo, diag = nb.graph.to_coo_batch_ZA_diag(A)
        assert coo.dtype == torch.int32 and tuple(coo.shape) == (3, 2 * 4096 * 14) and diag.dtype == torch.int64
        assert sha(coo.cpu().numpy()) == str(g[f"coo_sha_{tag}"])
        assert np.array_equal(diag.cpu().numpy(), g[f"diag_{tag}"])
        assert torch.equal(nb.graph.to_coo_batch(A), coo)
        assert torch.equal(nb.graph.get_indices_from_list_CSR(A), coo[1])
        nb.graph.confirm_CSR_to_COO_index_integrity(A, coo)
        adj = coo._nbpc_adjacency
        assert adj.check()
        order = np.argsort(coo[1].cpu().numpy(), kind="stable")
        assert np.array_equal(adj.csrT_edge.cpu().numpy(), order)
        assert np.array_equal(np.diff(adj.csrT_ptr.cpu().numpy()), np.bincount(coo[1].cpu().numpy(), minlength=8192))


def test_input_features_golden(nb):
    g = load_golden("layers_small.npz")
    b, N = g["x"].shape[:2]; k = int(g["k"])
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(g["x"], k))
    assert np.array_equal(coo.cpu().numpy(), g["coo"]) and np.array_equal(diag.cpu().numpy(), g["diag"])
    e = nb.graph.get_input_features_shift_inv_ZA(torch.tensor(g["x"], device=DEV), torch.tensor(g["za"], device=DEV),
                                                 coo, diag, (b, N, k))
    assert np.array_equal(e.cpu().numpy(), g["f32_edges"])
    e2, n2 = nb.graph.get_input_features_shift_inv(torch.tensor(g["feat_X6"], device=DEV), coo, (b, N, k))
    assert np.array_equal(e2.cpu().numpy(), g["feat_edges"]) and np.array_equal(n2.cpu().numpy(), g["feat_nodes"])
    assert np.array_equal(nb.graph.include_node_features(e2, n2, coo).cpu().numpy(), g["feat_nodes9"])
    rs = torch.full((b * N * k, 1), 0.75, device=DEV)
    assert np.array_equal(nb.graph.include_node_features(e2, n2, coo, redshift=rs).cpu().numpy(), g["feat_nodes10"])
    # NumPy COO (as the reference feeds it) is accepted too
    e3 = nb.graph.get_input_features_shift_inv_ZA(g["x"], g["za"], g["coo"], g["diag"], (b, N, k))
    assert np.array_equal(e3.cpu().numpy(), g["f32_edges"])


# =============================================================================== graph layers
def test_graph_layers_per_layer_golden(nb):
    g = load_golden("layers_small.npz")
    ch = list(g["channels"]); k = int(g["k"]); b, N = g["x"].shape[:2]
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(g["x"], k))
    mv, tp = cuda_model_vars(g, ch)
    H = torch.tensor(g["f32_edges"], device=DEV)
    L = len(ch) - 1
    for li in range(L):
        last = li == L - 1
        H = nb.graph.shift_inv_layer(H, coo, (b, N), tp[li], is_last=last)
        if not last:
            H = torch.relu(H)
        ref32, ref64 = g[f"f32_H{li}"], g[f"f64_H{li}"]
        out = H.detach().cpu().numpy()
        assert out.shape == ref32.shape
        np.testing.assert_allclose(out, ref32, rtol=2e-5, atol=2e-6)
        assert np.abs(out - ref64).max() <= 4 * max(np.abs(ref32 - ref64).max(), 1e-7)


@pytest.mark.parametrize("plain_coo", [False, True])
def test_graph_model_small_golden_fwd_bwd(nb, plain_coo):
    g = load_golden("layers_small.npz")
    ch = list(g["channels"]); k = int(g["k"]); b, N = g["x"].shape[:2]
    if plain_coo:   # a bare (3,c) tensor, as a reference caller would feed it
        coo, diag = torch.tensor(g["coo"], device=DEV), torch.tensor(g["diag"], device=DEV)
    else:
        coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(g["x"], k))
    mv, tp = cuda_model_vars(g, ch)
    pred = nb.graph.model_func_shift_inv_za(torch.tensor(g["x"], device=DEV), coo, torch.tensor(g["za"], device=DEV),
                                            diag, mv, (b, N, k))
    loss = nb.nn.loss_ZA(pred, torch.tensor(g["tgt"], device=DEV))
    loss.backward()
    assert tuple(pred.shape) == (b, N, 3)
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["f32_pred"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(loss.item(), g["f64_loss"], rtol=1e-5)
    for li, (Ws, B) in enumerate(tp):
        for wi, w in enumerate(Ws):
            np.testing.assert_allclose(w.grad.cpu().numpy(), g[f"f64_gW{li}_{wi}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(B.grad.cpu().numpy(), g[f"f64_gB{li}"], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_graph_model_c1_golden(nb, syn, kind):
    """BASELINE config 1: 16^3 particles, k=14, 3-layer graph net, batch 2, one fwd+bwd step."""
    g = load_golden("model_16.npz")
    ch = list(g["channels"]); k = int(g["k"]); b, N = 2, 4096
    x = syn.make_box(kind, b, N, 0)
    za, tgt = syn.za_features(b, N, 0)
    assert sha(x) == str(g[f"x_sha_{kind}"]) and sha(za) == str(g[f"za_sha_{kind}"])
    store = nb.train_utils.ParamStore(ch, device=DEV)
    store.load_numpy(syn.glorot_params(ch))
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
    xt = torch.tensor(x, device=DEV)
    A = nb.graph.get_kneighbor_list(xt, k)
    coo, diag = nb.graph.to_coo_batch_ZA_diag(A)
    pred = nb.graph.model_func_shift_inv_za(xt, coo, torch.tensor(za, device=DEV), diag, mv, (b, N, k))
    loss = nb.nn.loss_ZA(pred, torch.tensor(tgt, device=DEV))
    loss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g[f"{kind}_f32_pred"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(loss.item(), g[f"{kind}_f64_loss"], rtol=1e-5)
    for li in range(len(ch) - 1):
        W, B = store.get_layer_vars(li)
        for wi in range(4):
            np.testing.assert_allclose(W.grad[wi].cpu().numpy(), g[f"{kind}_f64_gW{li}_{wi}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(B.grad.cpu().numpy(), g[f"{kind}_f64_gB{li}"], rtol=2e-4, atol=1e-7)


def test_layer_odd_widths_and_conv_golden(nb):
    g = load_golden("layers_small.npz")
    b, N = g["x"].shape[:2]
    coo = torch.tensor(g["coo"], device=DEV)
    for last, t in ((False, "mid"), (True, "last")):
        Ht = torch.tensor(g["odd_H_in"], device=DEV, requires_grad=True)
        Wt = [torch.tensor(g[f"odd_W{i}"], device=DEV, requires_grad=True) for i in range(4)]
        Bt = torch.tensor(g["odd_B"], device=DEV, requires_grad=True)
        o = nb.graph.shift_inv_layer(Ht, coo, (b, N), (Wt, Bt), is_last=last)
        (o * torch.tensor(g[f"odd_{t}_gout"], device=DEV)).sum().backward()
        np.testing.assert_allclose(o.detach().cpu().numpy(), g[f"odd_{t}_out"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(Ht.grad.cpu().numpy(), g[f"odd_{t}_gH"], rtol=1e-4, atol=1e-5)
        for i in range(4):
            np.testing.assert_allclose(Wt[i].grad.cpu().numpy(), g[f"odd_{t}_gW{i}"], rtol=1e-4, atol=2e-4)
        np.testing.assert_allclose(Bt.grad.cpu().numpy(), g[f"odd_{t}_gB"], rtol=1e-4, atol=2e-4)
    H = torch.tensor(g["odd_H_in"], device=DEV)
    for ci, nm in ((0, "row"), (1, "col"), (2, "cube")):
        for bc, suffix in ((True, "bc"), (False, "nobc")):
            out = nb.graph.shift_inv_conv(H, coo[ci], b * N, bc)
            np.testing.assert_allclose(out.cpu().numpy(), g[f"conv_{nm}_{suffix}"], rtol=2e-5, atol=2e-6)


def test_shift_inv_conv_autograd_vs_oracle(nb):
    rng = np.random.default_rng(8)
    h = rng.standard_normal((4000, 6)).astype(np.float32)
    ids = rng.integers(0, 300, size=4000).astype(np.int32)      # arbitrary, unsorted, some segments empty
    gout = rng.standard_normal((4000, 6)).astype(np.float32)
    for bc in (True, False):
        ht = torch.tensor(h, device=DEV, requires_grad=True)
        o = nb.graph.shift_inv_conv(ht, torch.tensor(ids, device=DEV), 320, bc)
        hc = torch.tensor(h, dtype=torch.float64, requires_grad=True)
        oc = ref_layers.shift_inv_conv(hc, ids, 320, bc)
        go = gout if bc else gout[:320]
        (o * torch.tensor(go, device=DEV)).sum().backward()
        (oc * torch.tensor(go, dtype=torch.float64)).sum().backward()
        np.testing.assert_allclose(o.detach().cpu().numpy(), oc.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ht.grad.cpu().numpy(), hc.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_graph_step_is_deterministic(nb, syn):
    ch = [3, 32, 16, 3]; b, N, k = 2, 4096, 14
    x = torch.tensor(syn.clustered_box(b, N, 3), device=DEV)
    za, tgt = (torch.tensor(a, device=DEV) for a in syn.za_features(b, N, 3))
    outs = []
    for _ in range(2):
        store = nb.train_utils.ParamStore(ch, device=DEV)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
        coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, k))
        pred = nb.graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k))
        loss = nb.nn.loss_ZA(pred, tgt)
        loss.backward()
        outs.append((pred.detach().clone(), loss.detach().clone(), store.flat_grad.clone()))
    for a, c in zip(outs[0], outs[1]):
        assert torch.equal(a, c)     # bit-identical: no float atomics anywhere


def test_graph_model_properties_c2(nb, syn):
    """BASELINE config 2 size (32^3, b=8, k=14): shift invariance and particle-relabelling equivariance."""
    ch = [3, 32, 16, 3]; b, N, k = 8, 32768, 14
    # coordinates on a 2^-20 lattice inside [0.25, 0.75): adding 0.125 is then exact in float32
    x = torch.round(torch.tensor(syn.uniform_box(b, N, 1), device=DEV) * 2 ** 19) / 2 ** 20 + 0.25
    za = torch.tensor(syn.za_features(b, N, 1)[0], device=DEV)
    store = nb.train_utils.ParamStore(ch, device=DEV)
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)

    def run(xx, zz):
        coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(xx, k))
        with torch.no_grad():
            return nb.graph.model_func_shift_inv_za(xx, coo, zz, diag, mv, (b, N, k))
    base = run(x, za)
    assert tuple(base.shape) == (b, N, 3) and bool(torch.isfinite(base).all())
    perm = torch.randperm(N, device=DEV)
    assert torch.allclose(run(x[:, perm], za[:, perm]), base[:, perm], rtol=1e-4, atol=1e-6)
    assert torch.equal(run(x + 0.125, za), base)      # exact shift => identical graph, edges and output


# =============================================================================== set model + losses
def test_set_model_golden(nb):
    g = load_golden("set_small.npz")
    ch = list(g["channels"]); L = len(ch) - 1
    mv, tp = cuda_model_vars(g, ch)
    H = torch.tensor(g["X"], device=DEV)
    for li in range(L):
        H = nb.nn.set_layer(H, tp[li])
        if li < L - 1:
            H = torch.relu(H)
        np.testing.assert_allclose(H.detach().cpu().numpy(), g[f"f32_H{li}"], rtol=2e-5, atol=2e-6)
    pred = nb.nn.model_func_set(torch.tensor(g["X"], device=DEV), mv)
    loss = nb.nn.loss_ZA(pred, torch.tensor(g["Y"], device=DEV))
    loss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["f32_pred"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(loss.item(), g["f64_loss"], rtol=1e-5)
    for li, (Ws, B) in enumerate(tp):
        np.testing.assert_allclose(Ws[0].grad.cpu().numpy(), g[f"f64_gW{li}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(B.grad.cpu().numpy(), g[f"f64_gB{li}"], rtol=2e-4, atol=1e-7)
        assert all(w.grad is None for w in Ws[1:])      # nn.py:22: only W[0] takes part


def test_set_layer_properties(nb):
    """with B = 0 the output has zero mean over N; permutation equivariance."""
    H = torch.randn(3, 1000, 7, device=DEV) * 3 + 5
    W = [torch.randn(7, 11, device=DEV)]
    out = nb.nn.set_layer(H, (W, torch.zeros(11, device=DEV)))
    assert float(out.mean(dim=1).abs().max()) < 1e-4
    perm = torch.randperm(1000, device=DEV)
    assert torch.allclose(nb.nn.set_layer(H[:, perm], (W, torch.zeros(11, device=DEV))), out[:, perm], atol=1e-5)


def test_losses_and_readout_golden(nb):
    g = load_golden("losses.npz")
    p = torch.tensor(g["pred"], device=DEV, requires_grad=True)
    t = torch.tensor(g["truth"], device=DEV)
    ro = nb.nn.get_readout(p)
    assert np.array_equal(ro.detach().cpu().numpy(), g["f32_readout"])
    assert np.array_equal(nb.nn.get_readout(p[..., :3]).detach().cpu().numpy(), g["f32_readout3"])
    assert np.array_equal(nb.nn.periodic_boundary_dist(ro.detach(), t).cpu().numpy(), g["f32_pbd"])
    l1 = nb.nn.pbc_loss(ro, t)
    l1.backward()
    np.testing.assert_allclose(l1.item(), g["f64_pbc_loss"], rtol=1e-5)
    np.testing.assert_allclose(p.grad.cpu().numpy(), g["f64_pbc_loss_gpred"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(nb.nn.pbc_loss(ro, t, scale_error=False).item(), g["f64_pbc_loss_unscaled"], rtol=1e-5)
    p2 = torch.tensor(g["pred"][..., :3], device=DEV, requires_grad=True)
    l2 = nb.nn.loss_ZA(p2, t[..., :3])
    l2.backward()
    np.testing.assert_allclose(l2.item(), g["f64_loss_za"], rtol=1e-5)
    np.testing.assert_allclose(p2.grad.cpu().numpy(), g["f64_loss_za_gpred"], rtol=1e-5, atol=1e-9)
    # reference quirk preserved: readout(0) = readout(1) = 0.5 (nn.py:112-115 with sign(0) = 0)
    q = nb.nn.get_readout(torch.tensor([[[0.0, 1.0, 0.25]]], device=DEV))
    assert q.flatten().tolist() == [0.5, 0.5, 0.25]


def test_adam_tf_matches_tf_formula(nb):
    rng = np.random.default_rng(0)
    p = rng.standard_normal(1000).astype(np.float32); m = np.zeros_like(p); v = np.zeros_like(p)
    pt, mt, vt = (torch.tensor(a, device=DEV) for a in (p, m, v))
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    p64, m64, v64 = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for t in range(1, 6):
        gr = rng.standard_normal(1000).astype(np.float32)
        nb.ops.adam_tf_(pt, torch.tensor(gr, device=DEV), mt, vt, t, lr, b1, b2, eps, 0.5)
        g64 = gr.astype(np.float64) * 0.5
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        m64 = b1 * m64 + (1 - b1) * g64
        v64 = b2 * v64 + (1 - b2) * g64 * g64
        p64 = p64 - lr_t * m64 / (np.sqrt(v64) + eps)
    np.testing.assert_allclose(pt.cpu().numpy(), p64, rtol=1e-5, atol=1e-6)


# =============================================================================== every compiled layer shape
@pytest.mark.parametrize("k,q", [(3, 16), (3, 32), (3, 64), (16, 16), (16, 32), (16, 64), (32, 16), (32, 32), (32, 64),
                                 (64, 16), (64, 32), (64, 64), (16, 3), (32, 3), (9, 32), (10, 8), (8, 128), (128, 8)])
@pytest.mark.parametrize("is_last,relu", [(False, True), (False, False), (True, False)])
def test_graph_layer_shapes_vs_oracle(nb, k, q, is_last, relu):
    """Tiled kernels (and the baseline fallback for shapes they do not cover) against the float64 oracle,
    forward + all gradients, on a graph whose edge count is not a multiple of the 128-edge tile."""
    b, N, M = 2, 601, 10
    rng = np.random.default_rng(k * 131 + q)
    x = rng.random((b, N, 3)).astype(np.float32)
    coo, _ = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, M))
    coo_np = coo.cpu().numpy()
    c = b * N * M
    H = rng.standard_normal((c, k)).astype(np.float32)
    Ws = [(rng.standard_normal((k, q)) / np.sqrt(k)).astype(np.float32) for _ in range(4)]
    Bv = (0.1 * rng.standard_normal(q)).astype(np.float32)
    gout = rng.standard_normal(((b, N, q) if is_last else (c, q))).astype(np.float32)

    Ht = torch.tensor(H, device=DEV, requires_grad=True)
    Wt = [torch.tensor(w, device=DEV, requires_grad=True) for w in Ws]
    Bt = torch.tensor(Bv, device=DEV, requires_grad=True)
    if relu:   # the fused-activation path the network functions use
        o = nb.graph._layer(Ht, coo, (b, N), (Wt, Bt), is_last, True)
    else:
        o = nb.graph.shift_inv_layer(Ht, coo, (b, N), (Wt, Bt), is_last=is_last)
    (o * torch.tensor(gout, device=DEV)).sum().backward()

    Hc = torch.tensor(H, dtype=torch.float64, requires_grad=True)
    Wc = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws]
    Bc = torch.tensor(Bv, dtype=torch.float64, requires_grad=True)
    oc = ref_layers.shift_inv_layer(Hc, coo_np, (b, N), (Wc, Bc), is_last=is_last)
    if relu:
        # ReLU'(z) is discontinuous at 0: a pre-activation within rounding noise of zero may legitimately land on the
        # other side, and ONE flipped mask bit moves every gradient of its sample through the cube pool.  The oracle
        # therefore uses the device's mask, after checking that it differs from its own only at such |z| < 2e-5.
        mask_dev = torch.tensor(o.detach().cpu().numpy() > 0)
        flips = mask_dev != (oc.detach() > 0)
        assert int(flips.sum()) <= 4 and (not flips.any() or float(oc.detach().abs()[flips].max()) < 2e-5)
        oc = oc * mask_dev.double()
    (oc * torch.tensor(gout, dtype=torch.float64)).sum().backward()

    np.testing.assert_allclose(o.detach().cpu().numpy(), oc.detach().numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(Ht.grad.cpu().numpy(), Hc.grad.numpy(), rtol=2e-4, atol=2e-5)
    scale = float(np.abs(Wc[0].grad.numpy()).max())
    for i in range(4):
        np.testing.assert_allclose(Wt[i].grad.cpu().numpy(), Wc[i].grad.numpy(), rtol=2e-4, atol=2e-5 * max(scale, 1.0))
    np.testing.assert_allclose(Bt.grad.cpu().numpy(), Bc.grad.numpy(), rtol=2e-4, atol=2e-4)


def test_fast_and_baseline_kernels_agree(nb, syn):
    """NBPC_BASELINE=1 (barrier-free baseline kernels) vs the tiled kernels, whole model, in a subprocess."""
    import os
    import subprocess
    import sys
    code = r'''
import importlib, sys, types, torch, numpy as np
sys.path.insert(0, %r)
nb = importlib.import_module("n-body_pointcloudevolution_b200")
syn = nb.synthetic
ch = [3, 32, 16, 3]; b, N, k = 2, 4096, 14
x = torch.tensor(syn.uniform_box(b, N, 2), device="cuda"); za, tgt = (torch.tensor(a, device="cuda") for a in syn.za_features(b, N, 2))
store = nb.train_utils.ParamStore(ch, device="cuda")
mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, k))
pred = nb.graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k)); loss = nb.nn.loss_ZA(pred, tgt); loss.backward()
np.savez(sys.argv[1], pred=pred.detach().cpu().numpy(), grad=store.flat_grad.cpu().numpy())
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    outs = []
    for flag in ("0", "1"):
        path = f"/tmp/nbpc_fb_{flag}.npz"
        env = dict(os.environ, NBPC_BASELINE=flag)
        subprocess.check_call([sys.executable, "-c", code, path], env=env)
        outs.append(np.load(path))
    np.testing.assert_allclose(outs[0]["pred"], outs[1]["pred"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(outs[0]["grad"], outs[1]["grad"], rtol=1e-4, atol=1e-7)


# =============================================================================== tensor-core (tcgen05) math modes
@pytest.fixture
def math_mode(nb):
    """Sets the arithmetic of the edge-level projections for one test and restores the process default."""
    before = nb._lib.get_math_mode()

    def setter(mode):
        nb._lib.set_math_mode(mode)
    yield setter
    nb._lib.set_math_mode(before)


@pytest.mark.parametrize("mode", ["tf32x3", "tf32"])
@pytest.mark.parametrize("k,q", [(16, 16), (16, 32), (16, 64), (32, 16), (32, 32), (32, 64), (64, 16), (64, 32), (64, 64)])
@pytest.mark.parametrize("chain", [False, True])
def test_tensor_core_layer_vs_oracle(nb, math_mode, mode, k, q, chain):
    """tcgen05 edge kernels (forward, dH, dW1) against the float64 oracle.  tf32x3 (error-compensated) must meet the
    FP32 tolerances of the CUDA-core path; tf32 (one pass, 10-bit mantissa operands) is held to 4e-3 of the tensor scale.
    chain=True exercises the fused ReLU-backward mask of the input (the path the network functions take)."""
    math_mode(mode)
    assert nb._lib.get_math_mode() == mode
    b, N, M = 2, 601, 10           # 12020 edges: not a multiple of the 128-edge tile
    rng = np.random.default_rng(k * 17 + q)
    x = rng.random((b, N, 3)).astype(np.float32)
    coo, _ = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, M))
    coo_np = coo.cpu().numpy()
    c = b * N * M
    H0 = rng.standard_normal((c, k)).astype(np.float32)
    Ws = [(rng.standard_normal((k, q)) / np.sqrt(k)).astype(np.float32) for _ in range(4)]
    Bv = (0.1 * rng.standard_normal(q)).astype(np.float32)
    gout = rng.standard_normal((c, q)).astype(np.float32)

    H0t = torch.tensor(H0, device=DEV, requires_grad=True)
    Wt = [torch.tensor(w, device=DEV, requires_grad=True) for w in Ws]
    Bt = torch.tensor(Bv, device=DEV, requires_grad=True)
    if chain:    # H = relu(H0) with the mask applied by this layer's backward kernel (input_relu)
        Ht = torch.relu(H0t.detach()).requires_grad_(True)
        o = nb.graph._layer(Ht, coo, (b, N), (Wt, Bt), False, False, input_relu=True)
    else:
        Ht = H0t
        o = nb.graph.shift_inv_layer(Ht, coo, (b, N), (Wt, Bt))
    (o * torch.tensor(gout, device=DEV)).sum().backward()

    Hc0 = torch.tensor(H0, dtype=torch.float64, requires_grad=True)
    Hc = torch.relu(Hc0) if chain else Hc0
    Wc = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws]
    Bc = torch.tensor(Bv, dtype=torch.float64, requires_grad=True)
    oc = ref_layers.shift_inv_layer(Hc, coo_np, (b, N), (Wc, Bc), is_last=False)
    (oc * torch.tensor(gout, dtype=torch.float64)).sum().backward()

    got = [o.detach().cpu().numpy(), Ht.grad.cpu().numpy()] + [w.grad.cpu().numpy() for w in Wt] + [Bt.grad.cpu().numpy()]
    ref = [oc.detach().numpy(), Hc0.grad.numpy()] + [w.grad.numpy() for w in Wc] + [Bc.grad.numpy()]
    if mode == "tf32x3":
        np.testing.assert_allclose(got[0], ref[0], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(got[1], ref[1], rtol=2e-4, atol=2e-5)
        scale = max(float(np.abs(ref[2]).max()), 1.0)
        for i in range(2, 6):
            np.testing.assert_allclose(got[i], ref[i], rtol=2e-4, atol=2e-5 * scale)
        np.testing.assert_allclose(got[6], ref[6], rtol=2e-4, atol=2e-4)
    else:
        for g_, r_ in zip(got, ref):
            assert np.abs(g_ - r_).max() <= 4e-3 * max(float(np.abs(r_).max()), 1e-6)


def test_tensor_core_odd_edge_count_falls_back(nb, math_mode):
    """An odd number of edges cannot be viewed as packed 128-byte rows: the backward must take the CUDA-core kernel
    (still on the GPU) and stay within the FP32 tolerance."""
    math_mode("tf32x3")
    b, N, M, k, q = 1, 601, 9, 32, 16
    rng = np.random.default_rng(5)
    x = rng.random((b, N, 3)).astype(np.float32)
    coo, _ = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, M))
    c = b * N * M
    assert c % 2 == 1
    H = rng.standard_normal((c, k)).astype(np.float32)
    Ws = [(rng.standard_normal((k, q)) / np.sqrt(k)).astype(np.float32) for _ in range(4)]
    Bv = np.zeros(q, np.float32)
    Ht = torch.tensor(H, device=DEV, requires_grad=True)
    Wt = [torch.tensor(w, device=DEV, requires_grad=True) for w in Ws]
    o = nb.graph.shift_inv_layer(Ht, coo, (b, N), (Wt, torch.tensor(Bv, device=DEV, requires_grad=True)))
    o.sum().backward()
    Hc = torch.tensor(H, dtype=torch.float64, requires_grad=True)
    Wc = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws]
    oc = ref_layers.shift_inv_layer(Hc, coo.cpu().numpy(), (b, N), (Wc, torch.tensor(Bv, dtype=torch.float64)), is_last=False)
    oc.sum().backward()
    np.testing.assert_allclose(o.detach().cpu().numpy(), oc.detach().numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(Ht.grad.cpu().numpy(), Hc.grad.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(Wt[0].grad.cpu().numpy(), Wc[0].grad.numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("mode", ["tf32x3", "tf32"])
def test_tensor_core_model_c1_golden(nb, syn, math_mode, mode):
    """BASELINE config 1 in the tensor-core modes against the reference's own outputs (model_16 golden)."""
    math_mode(mode)
    kind = "uniform"
    g = load_golden("model_16.npz")
    ch = list(g["channels"]); k = int(g["k"]); b, N = 2, 4096
    x = syn.make_box(kind, b, N, 0)
    za, tgt = syn.za_features(b, N, 0)
    store = nb.train_utils.ParamStore(ch, device=DEV)
    store.load_numpy(syn.glorot_params(ch))
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
    xt = torch.tensor(x, device=DEV)
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(xt, k))
    pred = nb.graph.model_func_shift_inv_za(xt, coo, torch.tensor(za, device=DEV), diag, mv, (b, N, k))
    loss = nb.nn.loss_ZA(pred, torch.tensor(tgt, device=DEV))
    loss.backward()
    x3 = mode == "tf32x3"
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g[f"{kind}_f32_pred"], rtol=2e-5 if x3 else 5e-3, atol=2e-6 if x3 else 2e-4)
    np.testing.assert_allclose(loss.item(), g[f"{kind}_f64_loss"], rtol=1e-5 if x3 else 2e-3)
    for li in range(len(ch) - 1):
        W, B = store.get_layer_vars(li)
        for wi in range(4):
            ref = g[f"{kind}_f64_gW{li}_{wi}"]
            if x3:
                np.testing.assert_allclose(W.grad[wi].cpu().numpy(), ref, rtol=2e-4, atol=1e-7)
            else:
                assert np.abs(W.grad[wi].cpu().numpy() - ref).max() <= 5e-3 * float(np.abs(ref).max()) + 1e-9, (li, wi)


# =============================================================================== multi-redshift rollout
@pytest.mark.parametrize("with_redshift", [False, True])
def test_rollout_shift_inv_vs_oracle(nb, with_redshift):
    """model_func_shift_inv / rollout_shift_inv (graph.py:517-567, legacy multi-redshift path): periodic kNN rebuilt every
    step, 9|10-channel input edges, scaled residual update, readout wrap.  Every step is compared with the float64
    oracle fed with the device's state of the previous step (teacher forcing: the kNN of a drifted copy could differ)."""
    b, N, K, thr = 2, 512, 8, 0.25
    rng = np.random.default_rng(11)
    X0 = np.concatenate([rng.random((b, N, 3)), 0.02 * rng.standard_normal((b, N, 3))], axis=-1).astype(np.float32)
    ch = [10 if with_redshift else 9, 16, 16, 6]
    redshifts = [0.9, 0.4, 0.1] if with_redshift else None   # small: one readout wrap must bring positions back
    steps = 3
    params = []
    for _ in range(steps):
        layers = []
        for kk, qq in zip(ch[:-1], ch[1:]):
            layers.append(([(rng.standard_normal((kk, qq)) * np.sqrt(2.0 / (kk + qq))).astype(np.float32) for _ in range(4)],
                           (0.01 * rng.standard_normal(qq)).astype(np.float32)))
        params.append(layers)
    scalars = (0.05, 0.02)

    def mv_dev(i):
        tp = [([torch.tensor(w, device=DEV) for w in Ws], torch.tensor(B, device=DEV)) for Ws, B in params[i]]
        return types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j], get_scalars=lambda: scalars)

    def mv_ref(i):
        tp = [([torch.tensor(w, dtype=torch.float64) for w in Ws], torch.tensor(B, dtype=torch.float64)) for Ws, B in params[i]]
        return types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j], get_scalars=lambda: scalars)

    states = nb.graph.rollout_shift_inv(X0, [mv_dev(i) for i in range(steps)], K, thr, redshifts=redshifts, trajectory=True)
    assert len(states) == steps and tuple(states[-1].shape) == (b, N, 6)
    prev = X0
    for i in range(steps):
        A = ref_graph.get_pbc_kneighbors_csr(prev, K, thr)
        dev_idx = idx_of(nb.graph.get_pbc_kneighbors_csr(prev, K, thr))
        assert np.array_equal(dev_idx, np.stack([a.indices.reshape(N, K) for a in A]))
        coo = ref_graph.to_coo_batch(A)
        rs = None if redshifts is None else torch.full((b * N * K, 1), redshifts[i], dtype=torch.float64)
        out = ref_layers.model_func_shift_inv(torch.tensor(prev, dtype=torch.float64), coo, mv_ref(i), (b, N, K), torch.relu, rs)
        ref = ref_layers.get_readout(out).numpy()
        got = states[i].cpu().numpy()
        assert got[..., :3].min() >= 0.0 and got[..., :3].max() <= 1.0
        d = np.abs(got - ref)
        d[..., :3] = np.minimum(d[..., :3], 1.0 - d[..., :3])      # positions live on the unit torus
        assert d.max() < 2e-5, (i, d.max())
        prev = got


@pytest.mark.parametrize("tag", ["v9", "v10"])
def test_legacy_multi_redshift_model_golden(nb, tag):
    """graph.model_func_shift_inv against the reference's own legacy block (tests/golden/rollout_small.npz)."""
    g = load_golden("rollout_small.npz")
    X, coo = g[f"{tag}_X"], g[f"{tag}_coo"]
    b, N = X.shape[0], X.shape[1]
    K = coo.shape[1] // (b * N)
    ch = [int(v) for v in g[f"{tag}_channels"]]
    tp = [([torch.tensor(g[f"{tag}_W{li}_{wi}"], device=DEV) for wi in range(4)], torch.tensor(g[f"{tag}_B{li}"], device=DEV))
          for li in range(len(ch) - 1)]
    scalars = tuple(float(v) for v in g[f"{tag}_scalars"])
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j], get_scalars=lambda: scalars)
    rs = torch.full((b * N * K, 1), 2.5, device=DEV) if ch[0] == 10 else None
    y = nb.graph.model_func_shift_inv(X, torch.tensor(coo, device=DEV), mv, (b, N, K), torch.relu, rs)
    np.testing.assert_allclose(y.cpu().numpy(), g[f"{tag}_f32_out"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(y.cpu().numpy(), g[f"{tag}_f64_out"], rtol=2e-5, atol=2e-6)


# =============================================================================== drop-in train.py
@pytest.mark.parametrize("argv", [["-k", "8", "-c", "3", "16", "3"], ["-k", "-1", "-c", "6", "16", "8", "3"],
                                  ["-k", "8", "-c", "3", "16", "3", "--graph"]])
def test_train_entry_point_runs(tmp_path, argv):
    """train.py with the reference's flags (utils.py:242-271): graph model (-k 8) and set model (-k -1) on the synthetic
    data set, checkpoints written, evaluation loop returns finite errors, and training reduces the loss."""
    import train
    common = ["-i", "40", "-b", "2", "-t", "4", "-l", "0.01", "--side", "8", "--checkpoint", "20", "--out_dir", str(tmp_path), "-n", "t"]
    err = train.main(argv + common)
    assert err is not None and err.shape == (2,) and np.isfinite(err).all()
    assert (tmp_path / "t" / "Session" / "chkpt-40.pt").is_file() and (tmp_path / "t" / "Results" / "error_test.npy").is_file()
    err0 = train.main([v for v in argv if v != "--graph"] + ["-i", "0"] + common[2:])          # untrained parameters, same test set
    assert err.mean() < err0.mean()


def test_graphed_step_matches_eager(nb, syn):
    """train_utils.GraphedStep (whole training step captured in one CUDA graph, Adam step count in device memory) must
    reproduce the eager loop bit for bit: same losses, same parameters after 6 steps on changing inputs."""
    tu, graph, nn_ = nb.train_utils, nb.graph, nb.nn
    ch, b, N, k = [3, 32, 16, 3], 2, 1000, 8
    batches = []
    for i in range(3):
        x = torch.tensor(syn.make_box("clustered", b, N, 20 + i), device=DEV)
        za, tgt = (torch.tensor(t, device=DEV) for t in syn.za_features(b, N, 20 + i))
        batches.append((x, za, tgt))

    def make():
        store = tu.ParamStore(ch, device=DEV)
        store.load_numpy(syn.glorot_params(ch))
        adam = tu.AdamTF(store, lr=0.01)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)

        def step(x, za, tgt, dev_step):
            coo, diag = graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))
            loss = nn_.loss_ZA(graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k)), tgt)
            store.zero_grad()
            loss.backward()
            adam.step_dev() if dev_step else adam.step()
            return loss
        return store, step

    store_e, step_e = make()
    eager = [float(step_e(*batches[i % 3], False).detach()) for i in range(6)]
    store_g, step_g = make()
    gs = tu.GraphedStep(lambda x, za, tgt: step_g(x, za, tgt, True), batches[0], warmup=0)
    assert gs.kernels_per_replay > 20
    graphed = [float(gs(*batches[i % 3]).detach()) for i in range(6)]
    assert graphed == eager
    assert torch.equal(store_g.flat, store_e.flat) and int(store_g.step_dev.item()) == 6
    assert eager[-1] < eager[0]


def test_pipelined_step_matches_eager(nb, syn):
    """train_utils.PipelinedStep (update of step i-1 applied on a forked stream at the start of replay i, next to the kNN
    build; the form the data-parallel runs capture, here with world = 1) reproduces the eager loop bit for bit."""
    tu, graph, nn_ = nb.train_utils, nb.graph, nb.nn
    ch, b, N, k = [3, 32, 16, 3], 2, 1000, 8
    batches = []
    for i in range(3):
        x = torch.tensor(syn.make_box("uniform", b, N, 30 + i), device=DEV)
        za, tgt = (torch.tensor(t, device=DEV) for t in syn.za_features(b, N, 30 + i))
        batches.append((x, za, tgt))

    def make():
        store = tu.ParamStore(ch, device=DEV)
        store.load_numpy(syn.glorot_params(ch))
        adam = tu.AdamTF(store, lr=0.01)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)

        def prep(x, za, tgt):
            return graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))

        def grad(ctx, x, za, tgt):
            loss = nn_.loss_ZA(graph.model_func_shift_inv_za(x, ctx[0], za, ctx[1], mv, (b, N, k)), tgt)
            store.zero_grad()
            loss.backward()
            return loss
        return store, adam, prep, grad

    store_e, adam_e, prep, grad = make()
    eager = []
    for i in range(6):
        eager.append(float(grad(prep(*batches[i % 3]), *batches[i % 3]).detach()))
        adam_e.step_dev()
    store_p, adam_p, prep, grad = make()
    ps = tu.PipelinedStep(prep, grad, store_p, adam_p, 1, batches[0])
    piped = [float(ps(*batches[i % 3]).detach()) for i in range(6)]
    ps.flush()
    ps.close()
    assert piped == eager
    assert torch.equal(store_p.flat, store_e.flat) and int(store_p.step_dev.item()) == 6 == int(store_e.step_dev.item())


@pytest.mark.parametrize("deferred", [False, True])
def test_overlapped_step_matches_eager(nb, syn, deferred):
    """train_utils.OverlappedStep (the graph of batch i+1 built on a second stream next to forward / backward / Adam of batch i,
    four CUDA graphs ordered by events) reproduces the eager loop bit for bit: losses (one call late), parameters, step count.
    deferred: with a head_fn the update of a step is captured at the start of the NEXT training graph, next to the
    parameter-free edge features."""
    tu, graph, nn_ = nb.train_utils, nb.graph, nb.nn
    ch, b, N, k = [3, 32, 16, 3], 2, 1000, 8
    batches = []
    for i in range(3):
        x = torch.tensor(syn.make_box("uniform", b, N, 40 + i), device=DEV)
        za, tgt = (torch.tensor(t, device=DEV) for t in syn.za_features(b, N, 40 + i))
        batches.append((x, za, tgt))

    def make():
        store = tu.ParamStore(ch, device=DEV)
        store.load_numpy(syn.glorot_params(ch))
        adam = tu.AdamTF(store, lr=0.01)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)

        def prep(x, za, tgt):
            return graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))

        def grad(ctx, x, za, tgt):
            loss = nn_.loss_ZA(graph.model_func_shift_inv_za(x, ctx[0], za, ctx[1], mv, (b, N, k)), tgt)
            store.zero_grad()
            loss.backward()
            return loss
        return store, adam, prep, grad

    store_e, adam_e, prep, grad = make()
    eager = []
    for i in range(7):
        eager.append(float(grad(prep(*batches[i % 3]), *batches[i % 3]).detach()))
        adam_e.step_dev()
    store_o, adam_o, prep, grad = make()
    if deferred:
        mv_o = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store_o.get_layer_vars)

        def head(ctx, x, za, tgt):
            return graph.get_input_features_shift_inv_ZA(x, za, ctx[0], ctx[1], (b, N, k))

        def grad_h(ctx, edges, x, za, tgt):
            loss = nn_.loss_ZA(graph.network_func_shift_inv_za(edges, ctx[0], len(ch) - 1, (b, N), torch.relu, mv_o), tgt)
            store_o.zero_grad()
            loss.backward()
            return loss
        ov = tu.OverlappedStep(prep, grad_h, store_o, adam_o, 1, batches[0], head_fn=head)
    else:
        ov = tu.OverlappedStep(prep, grad, store_o, adam_o, 1, batches[0])
    got = []
    for i in range(7):
        out = ov(*batches[i % 3])
        assert (out is None) == (i == 0)
        if out is not None:
            got.append(float(out.detach()))                 # (synchronises: the static loss tensor is overwritten two calls later)
    got.append(float(ov.flush().detach()))
    ov.close()
    assert got == eager
    assert torch.equal(store_o.flat, store_e.flat) and int(store_o.step_dev.item()) == 7 == int(store_e.step_dev.item())


def test_train_step_128_cubed_finite_and_deterministic(nb, syn):
    """BASELINE's largest box as a TRAINING step (29.4 M edges, 3.8 GB edge tensors: byte offsets beyond 2^32): loss and
    every gradient finite, two runs bit-identical, and sample-independence: the same box inside a batch of one and as
    sample 0 of nothing else gives the gradients the 32^3 tests pin (size-independent property: the mean of the
    per-node loss terms equals the loss)."""
    ch, b, N, k = [3, 32, 16, 3], 1, 128 ** 3, 14
    x = torch.tensor(syn.make_box("uniform", b, N, 1), device=DEV)
    za, tgt = (torch.tensor(t, device=DEV) for t in syn.za_features(b, N, 1))
    outs = []
    for _ in range(2):
        store = nb.train_utils.ParamStore(ch, device=DEV)
        store.load_numpy(syn.glorot_params(ch))
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
        coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, k))
        pred = nb.graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k))
        loss = nb.nn.loss_ZA(pred, tgt)
        store.zero_grad()
        loss.backward()
        outs.append((float(loss.detach()), store.flat_grad.clone(), pred.detach()))
        del coo, diag, pred, loss
    assert np.isfinite(outs[0][0]) and bool(torch.isfinite(outs[0][1]).all()) and float(outs[0][1].abs().max()) > 0
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    per_node = ((outs[0][2] - tgt) ** 2).sum(-1).double().mean()
    assert abs(float(per_node) - outs[0][0]) <= 1e-5 * outs[0][0]


# =============================================================================== 15-weight layer
def test_15op_layer_golden(nb):
    """graph.get_symmetrized_adjacency / shift_inv_15op_layer / network_func_15op_shift_inv_za (graph.py:20-216) against
    the unmodified reference run (tests/golden/layer15_small.npz): adjacency bit-exact, outputs rtol 2e-5 vs the float32
    run, gradients rtol 2e-4 vs the float64 run."""
    g = load_golden("layer15_small.npz")
    ch = [int(v) for v in g["channels"]]
    x = g["x"]
    b, N, K = x.shape[0], x.shape[1], int(g["K"])
    adj = nb.graph.get_symmetrized_adjacency(nb.graph.get_kneighbor_list(x, K))
    for name in ("row", "col", "all", "tra", "dia", "dal"):
        assert np.array_equal(adj[name].cpu().numpy(), g[f"adj_{name}"]), name
    tp = [(torch.tensor(g[f"W{li}"], device=DEV, requires_grad=True), torch.tensor(g[f"B{li}"], device=DEV, requires_grad=True))
          for li in range(len(ch) - 1)]
    H = torch.tensor(g["H"], device=DEV, requires_grad=True)
    mgr = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j])
    lay0 = nb.graph.shift_inv_15op_layer(H, adj, (b, N), tp[0])
    # a plain dict of arrays (what a reference caller would pass) must work too
    lay0_d = nb.graph.shift_inv_15op_layer(H, {k: g[f"adj_{k}"] for k in ("row", "col", "all", "tra", "dia", "dal")}, (b, N), tp[0])
    net = nb.graph.model_func_15op_shift_inv_za(H, adj, mgr, (b, N, K))
    loss = ((net - torch.tensor(g["tgt"], device=DEV)) ** 2).sum(-1).mean()
    loss.backward()
    np.testing.assert_allclose(lay0.detach().cpu().numpy(), g["f32_layer0"], rtol=2e-5, atol=2e-5)
    # the plain-dict path composes the layer from primitives (any adjacency); the canonical path runs the fused kernels
    np.testing.assert_allclose(lay0_d.detach().cpu().numpy(), lay0.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(net.detach().cpu().numpy(), g["f32_net"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(loss.item(), float(g["f64_loss"]), rtol=1e-5)
    np.testing.assert_allclose(H.grad.cpu().numpy(), g["f64_gH"], rtol=2e-4, atol=2e-6)
    for li, (W, B) in enumerate(tp):
        scale = float(np.abs(g[f"f64_gW{li}"]).max())
        np.testing.assert_allclose(W.grad.cpu().numpy(), g[f"f64_gW{li}"], rtol=2e-4, atol=2e-5 * scale)
        np.testing.assert_allclose(B.grad.cpu().numpy(), g[f"f64_gB{li}"], rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("b,N,M,ch", [(3, 601, 10, [3, 32, 16, 3]), (1, 257, 5, [3, 16, 32, 3]), (2, 1000, 32, [3, 64, 16, 3]),
                                      (5, 130, 7, [3, 32, 3])])
def test_graph_model_odd_shapes_vs_oracle(nb, b, N, M, ch):
    """Whole model (default math mode) on shapes where nothing divides anything: edges per sample not a multiple of the
    128-edge tile (the one-pass first-layer backward walks per-sample tiles), odd batch sizes, k = 5 / 7 / 32, two- and
    three-layer nets - forward, loss and every gradient against the float64 oracle."""
    rng = np.random.default_rng(b * 1000 + N)
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

    rA = ref_graph.get_kneighbor_list(x, M, backend="exact")
    rcoo, rdiag = ref_graph.to_coo_batch_ZA_diag(rA)
    assert np.array_equal(coo.cpu().numpy(), rcoo) and np.array_equal(diag.cpu().numpy(), rdiag)
    rp = [([torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws], torch.tensor(B, dtype=torch.float64, requires_grad=True))
          for Ws, B in params]
    rmv = types.SimpleNamespace(channels=ch, get_layer_vars=lambda i: rp[i])
    rpred = ref_layers.model_func_shift_inv_za(torch.tensor(x, dtype=torch.float64), rcoo, torch.tensor(za, dtype=torch.float64), rdiag,
                                               rmv, (b, N, M))
    rloss = ref_layers.loss_ZA(rpred, torch.tensor(tgt, dtype=torch.float64))
    rloss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), rpred.detach().numpy(), rtol=5e-5, atol=5e-6)
    np.testing.assert_allclose(loss.item(), rloss.item(), rtol=2e-5)
    for li in range(len(ch) - 1):
        for wi in range(4):
            ref = rp[li][0][wi].grad.numpy()
            np.testing.assert_allclose(tp[li][0][wi].grad.cpu().numpy(), ref, rtol=5e-4, atol=5e-5 * float(np.abs(ref).max()))
        ref = rp[li][1].grad.numpy()
        np.testing.assert_allclose(tp[li][1].grad.cpu().numpy(), ref, rtol=5e-4, atol=5e-5 * float(np.abs(ref).max()) + 1e-12)
