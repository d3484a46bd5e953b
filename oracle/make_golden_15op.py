"""TEST INFRASTRUCTURE ONLY - golden vectors for the 15-weight layer from the UNMODIFIED reference
(graph.shift_inv_15op_layer / network_func_15op_shift_inv_za, graph.py:20-216, run through oracle/tf_shim.py) on a
symmetrised adjacency built by oracle.ref_graph.get_symmetrized_adjacency (the reference has no builder).
Run here only:  python -m oracle.make_golden_15op   ->  tests/golden/layer15_small.npz"""
import types

import numpy as np
import torch

from oracle import ref_graph
from oracle.load_reference import load_reference
from oracle.make_golden import save


def main():
    rg, _ = load_reference()
    rng = np.random.default_rng(15)
    b, N, K = 2, 200, 6
    x = rng.random((b, N, 3)).astype(np.float32)
    A = ref_graph.get_kneighbor_list(x, K)
    adj = ref_graph.get_symmetrized_adjacency(A)
    S = adj["row"].shape[0]
    ch = [5, 8, 3]
    H = rng.standard_normal((S, ch[0])).astype(np.float32)
    params = [((rng.standard_normal((15, kk, qq)) * np.sqrt(2.0 / (kk + qq))).astype(np.float32),
               (0.05 * rng.standard_normal((2, qq))).astype(np.float32)) for kk, qq in zip(ch[:-1], ch[1:])]
    tgt = rng.standard_normal((b, N, ch[-1])).astype(np.float32)
    out = {"x": x, "H": H, "channels": np.array(ch), "tgt": tgt, "K": np.array(K)}
    for k_, v in adj.items():
        out[f"adj_{k_}"] = v
    for li, (W, B) in enumerate(params):
        out[f"W{li}"] = W
        out[f"B{li}"] = B
    tadj = {k_: torch.tensor(v.astype(np.int64)) for k_, v in adj.items()}
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        tp = [(torch.tensor(W, dtype=dt, requires_grad=True), torch.tensor(B, dtype=dt, requires_grad=True)) for W, B in params]
        Ht = torch.tensor(H, dtype=dt, requires_grad=True)
        mgr = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j])
        lay0 = rg.shift_inv_15op_layer(Ht, tadj, (b, N), tp[0])
        net = rg.network_func_15op_shift_inv_za(Ht, tadj, len(ch) - 1, (b, N), torch.relu, mgr)
        loss = ((net - torch.tensor(tgt, dtype=dt)) ** 2).sum(-1).mean()
        loss.backward()
        out[f"{name}_layer0"] = lay0.detach().numpy()
        out[f"{name}_net"] = net.detach().numpy()
        out[f"{name}_loss"] = np.array(loss.item())
        out[f"{name}_gH"] = Ht.grad.numpy()
        for li, (W, B) in enumerate(tp):
            out[f"{name}_gW{li}"] = W.grad.numpy()
            out[f"{name}_gB{li}"] = B.grad.numpy()
    save("layer15_small.npz", **out)


if __name__ == "__main__":
    main()
