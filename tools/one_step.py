"""A few training steps of the BASELINE config-2 workload (for ncu / sanitizer runs): python tools/one_step.py [steps]"""
import importlib
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
b, N, k, ch = 8, 32 ** 3, 14, [3, 32, 16, 3]
syn = nb.synthetic
store = nb.train_utils.ParamStore(ch, device="cuda")
adam = nb.train_utils.AdamTF(store)
mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
x = torch.from_numpy(syn.uniform_box(b, N, 0)).cuda()
za, tgt = (torch.from_numpy(t).cuda() for t in syn.za_features(b, N, 0))
for _ in range(steps):
    coo, diag = nb.graph.to_coo_batch_ZA_diag(nb.graph.get_kneighbor_list(x, k))
    loss = nb.nn.loss_ZA(nb.graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k)), tgt)
    store.zero_grad()
    loss.backward()
    adam.step()
torch.cuda.synchronize()
print("loss", float(loss))
