import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
from oracle import ref_layers
DEV = "cuda"
k, q, is_last, relu = 32, 32, False, True
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
nb._lib.set_math_mode(mode)
b, N, M = 2, 601, 10
rng = np.random.default_rng(k * 131 + q)
x = rng.random((b, N, 3)).astype(np.float32)
coo, _ = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, M))
coo_np = coo.cpu().numpy()
c = b * N * M
H = rng.standard_normal((c, k)).astype(np.float32)
Ws = [(rng.standard_normal((k, q)) / np.sqrt(k)).astype(np.float32) for _ in range(4)]
Bv = (0.1 * rng.standard_normal(q)).astype(np.float32)
gout = rng.standard_normal((c, q)).astype(np.float32)
Ht = torch.tensor(H, device=DEV, requires_grad=True)
Wt = [torch.tensor(w, device=DEV, requires_grad=True) for w in Ws]
Bt = torch.tensor(Bv, device=DEV, requires_grad=True)
o = nb.graph._layer(Ht, coo, (b, N), (Wt, Bt), is_last, True)
(o * torch.tensor(gout, device=DEV)).sum().backward()
Hc = torch.tensor(H, dtype=torch.float64, requires_grad=True)
Wc = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in Ws]
Bc = torch.tensor(Bv, dtype=torch.float64, requires_grad=True)
oc = torch.relu(ref_layers.shift_inv_layer(Hc, coo_np, (b, N), (Wc, Bc), is_last=is_last))
(oc * torch.tensor(gout, dtype=torch.float64)).sum().backward()
d = np.abs(Ht.grad.cpu().numpy() - Hc.grad.numpy())
bad = np.where(d.max(axis=1) > 1e-4)[0]
print(mode, "fwd maxerr", np.abs(o.detach().cpu().numpy() - oc.detach().numpy()).max(), "bad rows", len(bad), bad[:40], "max", d.max())
for i in range(4):
    print("dW", i, np.abs(Wt[i].grad.cpu().numpy() - Wc[i].grad.numpy()).max())
mask_diff = ((o.detach().cpu().numpy() > 0) != (oc.detach().numpy() > 0)).sum()
print("mask flips", mask_diff)
