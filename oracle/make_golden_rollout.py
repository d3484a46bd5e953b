"""TEST INFRASTRUCTURE ONLY - golden vectors for the legacy multi-redshift model (graph.py:517-567).

The reference keeps `_network_func_shift_inv` / `_model_func_shift_inv` inside a commented-out block.  This script
reads that block from the UNMODIFIED /root/reference/graph.py (text between the two ''' fences at lines 516-568), renames
the inner call to the underscored definition it refers to, executes it inside the loaded reference module (so that it runs
on the reference's own include_node_features / shift_inv_layer / get_input_features_shift_inv through oracle/tf_shim.py)
and stores inputs and outputs in tests/golden/rollout_small.npz.  Run here only:  python -m oracle.make_golden_rollout
"""
import os
import types

import numpy as np
import torch

from oracle import ref_graph
from oracle.load_reference import REFERENCE_DIR, load_reference
from oracle.make_golden import save


def legacy_block():
    lines = open(os.path.join(REFERENCE_DIR, "graph.py")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def _network_func_shift_inv("))
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == "'''")
    src = "\n".join(lines[start:end])
    # the block calls network_func_shift_inv (no underscore), which exists nowhere else in the file
    return src.replace("net_out = network_func_shift_inv(", "net_out = _network_func_shift_inv("), (start + 1, end)


def main():
    rg, _ = load_reference()
    src, span = legacy_block()
    exec(compile(src, "reference graph.py legacy block", "exec"), rg.__dict__)
    rng = np.random.default_rng(5)
    b, N, K = 2, 256, 6
    out = {"lines": np.array(span)}
    for tag, kin, with_rs in (("v9", 9, False), ("v10", 10, True)):
        ch = [kin, 8, 6]
        X = np.concatenate([rng.random((b, N, 3)), 0.02 * rng.standard_normal((b, N, 3))], axis=-1).astype(np.float32)
        A = ref_graph.get_pbc_kneighbors_csr(X, K, 0.3)
        coo = ref_graph.to_coo_batch(A)
        params = [([(rng.standard_normal((kk, qq)) * np.sqrt(2.0 / (kk + qq))).astype(np.float32) for _ in range(4)],
                   (0.01 * rng.standard_normal(qq)).astype(np.float32)) for kk, qq in zip(ch[:-1], ch[1:])]
        scalars = (0.05, 0.02)
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            tp = [([torch.tensor(w, dtype=dt) for w in Ws], torch.tensor(B, dtype=dt)) for Ws, B in params]
            mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda j: tp[j], get_scalars=lambda: scalars)
            rs = torch.full((b * N * K, 1), 2.5, dtype=dt) if with_rs else None
            y = rg._model_func_shift_inv(torch.tensor(X, dtype=dt), torch.tensor(coo), mv, (b, N, K), activation=torch.relu, redshift=rs)
            out[f"{tag}_{name}_out"] = y.detach().numpy()
        out[f"{tag}_X"] = X
        out[f"{tag}_coo"] = coo
        out[f"{tag}_channels"] = np.array(ch)
        out[f"{tag}_scalars"] = np.array(scalars)
        for li, (Ws, B) in enumerate(params):
            for wi, w in enumerate(Ws):
                out[f"{tag}_W{li}_{wi}"] = w
            out[f"{tag}_B{li}"] = B
    save("rollout_small.npz", **out)


if __name__ == "__main__":
    main()
