"""Tensor-core (tcgen05) edge kernels vs the CUDA-core FP32 kernels and a float64 torch reference, through the C ABI.

  python tools/tc_check.py [--n 4096] [--b 2] [--m 14] [--shapes 32,16 16,32 ...] [--time]
Prints max abs / relative-to-scale errors of the layer output, dH, dW1 for every math mode."""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
ops, lib = nb.ops, nb._lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--b", type=int, default=2)
ap.add_argument("--m", type=int, default=14)
ap.add_argument("--shapes", nargs="+", default=["32,16"])
ap.add_argument("--modes", nargs="+", default=["fp32", "tf32", "tf32x3"])
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
B, N, M = a.b, a.n, a.m
x = torch.rand(B, N, 3, device=dev)
idx = ops.knn(x, M, False, 0.0, True, 1, False)[0]
coo, diag, csrT_ptr, csrT_edge, status = ops.adjacency(idx)
col = coo[1].contiguous()
c = B * N * M


def ref64(H, W, bias, g, relu_in):
    """float64 reference of the non-last layer and its backward (graph.py:394-456)."""
    H = H.double().requires_grad_(True)
    W = W.double().requires_grad_(True)
    row = torch.arange(c, device=dev) // M
    cube = row // N
    def pool(ids, n):
        s = torch.zeros(n, H.shape[1], dtype=torch.float64, device=dev).index_add(0, ids, H)
        cnt = torch.zeros(n, dtype=torch.float64, device=dev).index_add(0, ids, torch.ones(c, dtype=torch.float64, device=dev))
        return (s / cnt.clamp(min=1)[:, None])[ids]
    Z = H @ W[0] + pool(col.long(), B * N) @ W[1] + pool(row, B * N) @ W[2] + pool(cube, B) @ W[3] + bias.double()
    Z.backward(g.double())
    dH = H.grad
    if relu_in:
        dH = dH * (H.detach() > 0)
    return Z.detach(), dH, W.grad


def err(x, r):
    d = (x.double() - r).abs()
    return float(d.max()), float(d.max() / r.abs().max())


out = {}
for sh in a.shapes:
    k, q = (int(t) for t in sh.split(","))
    H = torch.randn(c, k, device=dev)
    H = torch.relu(H) if True else H
    W = torch.randn(4, k, q, device=dev) * (2.0 / (k + q)) ** 0.5
    bias = torch.randn(q, device=dev) * 0.1
    g = torch.randn(c, q, device=dev) * 0.01
    Zr, dHr, dWr = ref64(H, W, bias, g, True)
    for mode in a.modes:
        lib.set_math_mode(mode)
        Z, Pc, Pr, Pq = ops.graph_layer_fwd(H, col, csrT_ptr, csrT_edge, W, bias, B, N, M, False, False)
        dH, dW, dB = ops.graph_layer_bwd(g, H, Z, col, csrT_ptr, csrT_edge, W, Pc, Pr, Pq, B, N, M, False, False, True, True)
        torch.cuda.synchronize()
        rec = {"Z": err(Z, Zr), "dH": err(dH, dHr), "dW1": err(dW[0], dWr[0]), "dW2": err(dW[1], dWr[1])}
        if a.time:
            for name, fn in (("fwd_ms", lambda: ops.graph_layer_fwd(H, col, csrT_ptr, csrT_edge, W, bias, B, N, M, False, False)),
                             ("bwd_ms", lambda: ops.graph_layer_bwd(g, H, Z, col, csrT_ptr, csrT_edge, W, Pc, Pr, Pq, B, N, M, False, False, True, True))):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                rec[name] = round(e0.elapsed_time(e1) / 10, 4)
        out[f"{sh} {mode}"] = rec
        print(sh, mode, json.dumps(rec), flush=True)
lib.set_math_mode("fp32")
