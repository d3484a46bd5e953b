"""One kNN build per query kernel (for ncu captures): python tools/knn_one.py [k] [uniform|clustered]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
nb = importlib.import_module("n-body_pointcloudevolution_b200")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
x = torch.from_numpy(nb.synthetic.make_box(kind, 8, 32 ** 3, 0)).cuda()
for kern in ("thread", "warp"):
    nb.set_knn_kernel(kern)
    for _ in range(2):
        idx, _, _ = nb.ops.knn(x, k, False, 0.0, True, 1, False)
    torch.cuda.synchronize()
    print(kern, int(idx.sum()))
