"""Per-kernel device time of the tensor-core edge kernels for one layer shape (use with NBPC_GLT_FWD / NBPC_GLT_BWD
= "ctas,stages" to sweep the launch configuration; one process per configuration).
  NBPC_MATH=tf32 NBPC_GLT_FWD=3,3 python tools/tc_tune.py --shape 32,16"""
import argparse, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
ops, lib = nb.ops, nb._lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
ap.add_argument("--b", type=int, default=8)
ap.add_argument("--m", type=int, default=14)
ap.add_argument("--shape", default="32,16")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--nomask", action="store_true")
a = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
B, N, M = a.b, a.n, a.m
k, q = (int(t) for t in a.shape.split(","))
x = torch.rand(B, N, 3, device=dev)
idx = ops.knn(x, M, False, 0.0, True, 1, False)[0]
coo, diag, csrT_ptr, csrT_edge, status = ops.adjacency(idx)
col = coo[1].contiguous()
c = B * N * M
H = torch.relu(torch.randn(c, k, device=dev))
W = torch.randn(4, k, q, device=dev) * (2.0 / (k + q)) ** 0.5
bias = torch.randn(q, device=dev) * 0.1
g = torch.randn(c, q, device=dev) * 0.01
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    Z, Pc, Pr, Pq = ops.graph_layer_fwd(H, col, csrT_ptr, csrT_edge, W, bias, B, N, M, False, True)
    flush.zero_()
    ops.graph_layer_bwd(g, H, Z, col, csrT_ptr, csrT_edge, W, Pc, Pr, Pq, B, N, M, False, False, not a.nomask, True)
    flush.zero_()
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.prof_enable(True)
for _ in range(a.reps):
    run()
torch.cuda.synchronize()
rep = lib.prof_report()
lib.prof_enable(False)
tag = f"math={lib.get_math_mode()} FWD={os.environ.get('NBPC_GLT_FWD','-')} BWD={os.environ.get('NBPC_GLT_BWD','-')}"
for name, (cnt, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    if "edge" in name:
        print(f"{tag}  {name:40s} {ms / cnt * 1e3:8.1f} us")
