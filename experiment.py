"""Drop-in for the reference's experiment.py (experiment.py:1-325) on the B200-native set-layer path.

The reference script is the experimental variant of the set model: every layer is an ATTENTION layer built from three
set transforms (experiment.py:110-136), hidden layers are wrapped in leaky-ReLU + batch normalisation, tanh residual
projections of the input are computed next to them (experiment.py:96-107) and only the last one is added before the
output layer (experiment.py:139-157).  Same module-level configuration, function names and command line
(-i / -b / -n) as the reference; nothing runs at import (the reference builds its TF graph and opens a session at
import, experiment.py:159-175 - here `main()` does).

Where the arithmetic runs:
  * set_transform  (x - mean_N x) W [+ b]   -> libnbpc set-layer kernels        (ops.SetLayer; nn.py:10-28 / experiment.py:83-89)
  * xf^T xg over the B*N rows               -> libnbpc deterministic X^T Y      (ops.xty; experiment.py:131)
  * xh fg                                   -> libnbpc projection kernel        (ops.Linear; experiment.py:132)
  * Adam (tf.train.AdamOptimizer(lr))       -> libnbpc adam kernel on one flat buffer
  * softmax of the (k_out, k_out) gate, leaky-ReLU, tanh, the batch-norm affine map and the bias adds are plain torch
    elementwise device ops: they are outside the hot path SURVEY.md §8 scopes (a 16 x 16 softmax, O(B N k) elementwise).
tf.layers.batch_normalization is called with its default training=False (experiment.py:142), i.e. it is the affine map
gamma * x / sqrt(1 + 0.001) + beta with trainable gamma (1) and beta (0) - the moving statistics are never updated.
There is no CPU fallback.
"""
import argparse
import math
import os
import types

import numpy as np
import torch

import nbpc

####  data  ####
didx = 0
num_test = 200

####  vars  ####
lr = 0.006
channels = [6, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 3]   # experiment.py:27
kdims = list(zip(channels[:-1], channels[1:]))
rdims = [(6, k) for _, k in kdims]
num_layers = len(kdims)

####  train  ####
num_iters = 100000
batch_size = 10
num_particles = 32 ** 3

s1 = 77743196   # experiment.py:65
BN_EPS = 0.001  # tf.layers.batch_normalization default epsilon

# model variables (experiment.py:72-78) + the batch-norm gamma / beta tf.layers creates; set by init_model()
Wf = Wg = Wh = Rset = Bset = Gamma = Beta = None
params = None
X_in = None     # the reference's input placeholder: res_layer reads it (experiment.py:107)


####  loss  ####
def loss(yhat, y):
    """experiment.py:37-40 (= nn.loss_ZA)."""
    return nbpc.nn.loss_ZA(yhat, y)


####  init  ####
def glorot_normal(kdims, scale=1.0):
    """experiment.py:43-47"""
    fan = sum(kdims)
    dv = scale * np.sqrt(2 / fan)
    return np.random.normal(scale=dv, size=kdims).astype(np.float32)


def init_model(layer_channels=None, seed=s1, device="cuda"):
    """experiment.py:58-78: seeded glorot-normal Wf / Wg / Wh (k_in, k_out), Rset (6, k_out), bias 1e-6, in the
    reference's draw order; batch-norm gamma = 1, beta = 0 for the hidden layers."""
    global channels, kdims, rdims, num_layers, Wf, Wg, Wh, Rset, Bset, Gamma, Beta, params
    if layer_channels is not None:
        channels = list(layer_channels)
    kdims = list(zip(channels[:-1], channels[1:]))
    rdims = [(channels[0], k) for _, k in kdims]
    num_layers = len(kdims)
    np.random.seed(seed)
    spec = {"Wf": [glorot_normal(k) for k in kdims], "Wg": [glorot_normal(k) for k in kdims], "Wh": [glorot_normal(k) for k in kdims],
            "Rset": [glorot_normal(r) for r in rdims],
            "Bset": [np.ones((k[-1],), dtype=np.float32) * 1e-6 for k in kdims],
            "Gamma": [np.ones((k[-1],), dtype=np.float32) for k in kdims[:-1]],
            "Beta": [np.zeros((k[-1],), dtype=np.float32) for k in kdims[:-1]]}
    params = nbpc.train_utils.FlatParams(spec, device=device)
    Wf, Wg, Wh, Rset, Bset, Gamma, Beta = (params[g] for g in ("Wf", "Wg", "Wh", "Rset", "Bset", "Gamma", "Beta"))
    return params


# -----------------------------------------------------------------------------#
#                                   network                                   #
# -----------------------------------------------------------------------------#
class _XtY(torch.autograd.Function):
    """X^T Y over the rows (fixed summation order); dX = Y G^T, dY = X G."""

    @staticmethod
    def forward(ctx, X, Y):
        ctx.save_for_backward(X, Y)
        return nbpc.ops.xty(X, Y)

    @staticmethod
    def backward(ctx, G):
        X, Y = ctx.saved_tensors
        G = G.contiguous()
        return nbpc.ops.linear(Y, G, None, True), nbpc.ops.linear(X, G, None, False)


def set_transform(x_in, w, b=None):
    """experiment.py:83-89: (x - mean_N x) w [+ b]."""
    bias = b if b is not None else torch.zeros(w.shape[1], dtype=torch.float32, device=w.device)
    return nbpc.ops.SetLayer.apply(x_in, w, bias, False)


def res_layer(idx):
    """experiment.py:96-107: skip connection from the INPUT, weights (6, S)."""
    return set_transform(X_in, Rset[idx])


def attn_layer(x_in, idx):
    """experiment.py:110-136: o = xh softmax(xf^T xg) + b with xf, xg, xh = set_transform(x_in, wf | wg | wh)."""
    wf, wg, wh, b = Wf[idx], Wg[idx], Wh[idx], Bset[idx]
    xf, xg, xh = set_transform(x_in, wf), set_transform(x_in, wg), set_transform(x_in, wh)
    k = kdims[idx][-1]
    xfr, xgr, xhr = xf.reshape(-1, k), xg.reshape(-1, k), xh.reshape(-1, k)
    fg = torch.softmax(_XtY.apply(xfr, xgr), dim=-1)          # (k_out, k_out)
    o = nbpc.ops.Linear.apply(xhr, fg)                        # (BN, k_out)
    bdim, n = x_in.shape[0], x_in.shape[1]
    return o.reshape(bdim, n, k) + b


def _norm(x, i):
    """tf.layers.batch_normalization(x) with training=False: gamma * x / sqrt(moving_var + eps) + beta."""
    return Gamma[i] * (x * (1.0 / math.sqrt(1.0 + BN_EPS))) + Beta[i]


def net_fwd(x_in):
    """experiment.py:139-157."""
    global X_in
    X_in = x_in
    act_set = lambda t: torch.nn.functional.leaky_relu(t, negative_slope=0.2)    # tf.nn.leaky_relu default alpha
    act_res = torch.tanh
    H = _norm(act_set(attn_layer(x_in, 0)), 0)
    R = act_res(res_layer(0))
    for i in range(1, num_layers - 1):
        H = _norm(act_set(attn_layer(H, i)), i)
        R = act_res(res_layer(i))
    return attn_layer(H + R, num_layers - 1)


# -----------------------------------------------------------------------------#
#                                    utils                                    #
# -----------------------------------------------------------------------------#
def savestuff(name, err, data, dpath=os.path.expanduser("~/.Data/Nbody/za_misc")):
    """experiment.py:184-192"""
    spath = f"{dpath}/{name}"
    os.makedirs(spath, exist_ok=True)
    np.save(f"{spath}/test_cubes", data)
    np.save(f"{spath}/test_error", err)
    print("saved to " + spath)


def print_evaluation_results(err, label="Test", retstring=False):
    """experiment.py:194-207"""
    tbody = [f'\n# {label} Error\n# {"=" * 17}', f"  median : {np.median(err) : .5f}",
             f"    mean : {np.mean(err) : .5f} +- {np.std(err) : .4f} stdv\n"]
    eval_results = "\n".join(tbody)
    print(eval_results)
    if retstring:
        return eval_results


class Model:
    """The session-level state of the reference script (data split, placeholders, optimiser) as an object."""

    def __init__(self, X_train, X_val, X_test, learnrate=lr, device="cuda"):
        self.X_train, self.X_val, self.X_test, self.dev = X_train, X_val, X_test, device
        self.lr = learnrate

    def feed(self, x, b, i=None):
        """experiment.py:211-225 get_feed_dict"""
        idx = np.random.choice(x.shape[0], b, replace=False) if i is None else np.arange(i * b, (i + 1) * b)
        batch = torch.from_numpy(np.copy(x[idx])).to(self.dev, non_blocking=True)
        return batch[..., :6].contiguous(), batch[..., 6:].contiguous()

    def error(self, x_za, y):
        return loss(net_fwd(x_za), y)

    def train_step(self, x_za, y):
        err = self.error(x_za, y)
        params.zero_grad()
        err.backward()
        params.step_count += 1
        nbpc.ops.adam_tf_(params.flat, params.flat_grad, params.m, params.v, params.step_count, self.lr)
        return err

    def model_validation(self, bsize):
        """experiment.py:238-245"""
        nval = self.X_val.shape[0] // bsize
        val_hist = np.zeros((nval,), dtype=np.float32)
        with torch.no_grad():
            for i in range(nval):
                val_hist[i] = float(self.error(*self.feed(self.X_val, bsize, i)))
        return val_hist

    def model_test(self, bsize):
        """experiment.py:247-261"""
        ntest = self.X_test.shape[0] // bsize
        test_hist = np.zeros((ntest,), dtype=np.float32)
        test_preds = np.zeros((2,) + self.X_test.shape[:-1] + (channels[-1],), dtype=np.float32)
        with torch.no_grad():
            for i in range(ntest):
                j, k = i * bsize, (i + 1) * bsize
                x_za, y = self.feed(self.X_test, bsize, i)
                pred = net_fwd(x_za)
                test_hist[i] = float(loss(pred, y))
                test_preds[1, j:k] = pred.cpu().numpy()
        test_preds[0] = self.X_test[..., 6:]
        print_evaluation_results(test_hist, "Test")
        return test_hist, test_preds

    def model_train(self, n_iters, bsize, chkpt):
        """experiment.py:263-281"""
        train_hist = np.zeros((n_iters // chkpt,), dtype=np.float32)
        for step in range(n_iters):
            self.train_step(*self.feed(self.X_train, bsize))
            if (step + 1) % chkpt == 0:
                vmu = self.model_validation(bsize).mean() if self.X_val.shape[0] >= bsize else float("nan")
                print(f"{step + 1:>6}: Validation Error = {vmu:.6f}")
                train_hist[(step + 1) // chkpt - 1] = vmu
        return train_hist


# -----------------------------------------------------------------------------#
#                                     RUN                                     #
# -----------------------------------------------------------------------------#
cli = argparse.ArgumentParser()
cli.add_argument("-i", "--num_iters", type=int, default=num_iters)
cli.add_argument("-b", "--batch_size", type=int, default=batch_size)
cli.add_argument("-n", "--name", type=str, default="TEST")
# additions (not in the reference): where the data lives / synthetic stand-in, output directory, layer widths
cli.add_argument("--data_dir", type=str, default=os.path.expanduser("~/.Data/nbody_simulations"))
cli.add_argument("--out_dir", type=str, default=os.path.expanduser("~/.Data/Nbody/za_misc"))
cli.add_argument("--side", type=int, default=32, help="particles per axis of the synthetic data set")
cli.add_argument("--num_samples", type=int, default=0)
cli.add_argument("--num_test", type=int, default=num_test)
cli.add_argument("--channels", type=int, nargs="+", default=None)
cli.add_argument("--checkpoint", type=int, default=100)


def main(argv=None):
    args = cli.parse_args(argv)
    from train import Dataset                                      # same ZA_###.npy loader / synthetic stand-in as train.py
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    nbpc.ops.device_check()
    dargs = types.SimpleNamespace(data_dir=args.data_dir, data_idx=didx, num_samples=args.num_samples, num_test=args.num_test,
                                  batch_size=args.batch_size, side=args.side)
    dataset = Dataset(dargs)
    init_model(args.channels, device=dev)
    model = Model(np.copy(dataset.X_train), np.copy(dataset.X_val), np.copy(dataset.X_test), device=dev)
    train_hist = model.model_train(args.num_iters, args.batch_size, args.checkpoint)
    test_hist, test_preds = model.model_test(args.batch_size)
    savestuff(args.name, test_hist, test_preds, dpath=args.out_dir)
    return test_hist, test_preds, train_hist


if __name__ == "__main__":
    main()
