"""kNN build timing (device time, CUDA events): python tools/knn_bench.py [--sizes 32 64 128] [--k 14]
Reads NBPC_KNN_RHO (particles per cell) from the environment - one process per setting."""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[32, 64, 128])
ap.add_argument("--k", type=int, nargs="+", default=[14])
ap.add_argument("--batch32", type=int, default=8)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
out = {"rho": os.environ.get("NBPC_KNN_RHO", "default")}
for n in a.sizes:
    b = a.batch32 if n == 32 else 1
    for kind in ("uniform", "clustered"):
        x = torch.from_numpy(nb.synthetic.make_box(kind, b, n ** 3, 0)).cuda()
        for k in a.k:
            for periodic in (False, True):
                f = (lambda: nb.ops.knn(x, k, True, 0.05, True, 0, False)) if periodic else \
                    (lambda: nb.ops.knn(x, k, False, 0.0, True, 1, False))
                for _ in range(3):
                    f()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    f()
                e1.record()
                torch.cuda.synchronize()
                out[f"{n}^3 b={b} {kind} k={k} {'pbc' if periodic else 'open'}"] = round(e0.elapsed_time(e1) / a.reps, 4)
print(json.dumps(out))
