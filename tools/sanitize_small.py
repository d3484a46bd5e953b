"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): kNN (open + periodic), adjacency,
one training step of the [3,32,16,3] graph net in the three math modes, the set model, the rollout step.
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib, os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
syn, graph, nn_, tu = nb.synthetic, nb.graph, nb.nn, nb.train_utils
dev = "cuda"
b, N, k = 2, 1331, 14
x = torch.from_numpy(syn.make_box("clustered", b, N, 3)).to(dev)
za, tgt = (torch.from_numpy(t).to(dev) for t in syn.za_features(b, N, 3))
for mode in ("fp32", "tf32x3", "tf32"):
    nb.set_math_mode(mode)
    ch = [3, 32, 16, 3]
    store = tu.ParamStore(ch, device=dev)
    mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
    coo, diag = graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))
    loss = nn_.loss_ZA(graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k)), tgt)
    store.zero_grad(); loss.backward(); tu.AdamTF(store).step()
    print(mode, "graph loss", float(loss.detach()))
nb.set_math_mode("tf32x3")
chs = [6, 16, 8, 3]
st = tu.ParamStore(chs, device=dev)
mvs = types.SimpleNamespace(channels=chs, var_scope="params", num_layers=len(chs) - 1, get_layer_vars=st.get_layer_vars, activation=torch.relu)
ls = nn_.loss_ZA(nn_.model_func_set(torch.cat([x, za], -1), mvs), tgt); ls.backward()
print("set loss", float(ls.detach()))
X6 = torch.cat([x, 0.01 * za], -1)
ch5 = [9, 32, 16, 6]
s5 = tu.ParamStore(ch5, device=dev)
mv5 = types.SimpleNamespace(channels=ch5, var_scope="params", get_layer_vars=s5.get_layer_vars, get_scalars=lambda: (0.01, 0.01))
with torch.no_grad():
    out = graph.rollout_shift_inv(X6, mv5, 8, 0.2)
print("rollout", tuple(out.shape), bool(torch.isfinite(out).all()))
torch.cuda.synchronize()
