"""Per-step device times of the benchmark training step, queued back to back (no host sync between steps).
  NBPC_MATH=tf32 python tools/step_probe.py [--steps 30]"""
import argparse, importlib, os, sys, types, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
syn, graph, nn_, tu, lib = nb.synthetic, nb.graph, nb.nn, nb.train_utils, nb._lib
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--nogc", action="store_true")
ap.add_argument("--pool", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
N, b, k, ch = 32 ** 3, a.batch, 14, [3, 32, 16, 3]
store = tu.ParamStore(ch, device=dev)
store.load_numpy(syn.glorot_params(ch))
adam = tu.AdamTF(store, lr=0.01)
mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
x = torch.from_numpy(syn.make_box("uniform", b, N, 0)).to(dev)
za, tgt = (torch.from_numpy(t).to(dev) for t in syn.za_features(b, N, 0))

def step():
    A = graph.get_kneighbor_list(x, k)
    coo, diag = graph.to_coo_batch_ZA_diag(A)
    pred = graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k))
    loss = nn_.loss_ZA(pred, tgt)
    store.zero_grad()
    loss.backward()
    adam.step(grad_scale=1.0)
    return loss

import gc
if a.nogc:
    gc.disable()
for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
host = []
ev[0].record()
t0 = time.perf_counter()
for i in range(a.steps):
    step()
    ev[i + 1].record()
    host.append((time.perf_counter() - t0) * 1e3)
    if i % 10 == 0:
        print(i, "alloc GB", round(torch.cuda.memory_allocated() / 2**30, 2), "reserved GB", round(torch.cuda.memory_reserved() / 2**30, 2),
              "gc", gc.get_count(), "mallocs", torch.cuda.memory_stats()["num_device_alloc"], flush=True)
torch.cuda.synchronize()
print("math", lib.get_math_mode())
print("device ms/step:", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(a.steps)])
print("host enqueue ms (cumulative):", [round(h, 1) for h in host])
