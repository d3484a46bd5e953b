"""One eager training step of the default set model (for ncu captures of the sgt_* kernels)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
nb = importlib.import_module("n-body_pointcloudevolution_b200")
ch = [6, 64, 128, 128, 256, 64, 128, 16, 3]
b, N = 8, 32 ** 3
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.standard_normal((b, N, 6)).astype(np.float32)).cuda()
Y = torch.from_numpy((0.1 * rng.standard_normal((b, N, 3))).astype(np.float32)).cuda()
store = nb.train_utils.ParamStore(ch, device="cuda")
mv = store.model_vars(torch.relu)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    loss = nb.nn.loss_ZA(nb.nn.model_func_set(X, mv), Y)
    store.zero_grad()
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
