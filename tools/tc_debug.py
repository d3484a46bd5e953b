import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("n-body_pointcloudevolution_b200")
ops, lib = nb.ops, nb._lib
torch.manual_seed(0)
dev = "cuda"
B, N, M = 1, 1024, 8
k, q = (int(t) for t in (sys.argv[1] if len(sys.argv) > 1 else "32,16").split(","))
x = torch.rand(B, N, 3, device=dev)
idx = ops.knn(x, M, False, 0.0, True, 1, False)[0]
coo, diag, csrT_ptr, csrT_edge, status = ops.adjacency(idx)
col = coo[1].contiguous()
c = B * N * M
# structured inputs: H[e, j] = (j + 1) for e in tile 0 only...; dZ[e, n] = (n + 1) * 0.01
H = torch.zeros(c, k, device=dev)
H[:, :] = torch.arange(1, k + 1, device=dev, dtype=torch.float32)[None, :]
g = torch.zeros(c, q, device=dev)
g[:, :] = torch.arange(1, q + 1, device=dev, dtype=torch.float32)[None, :] * 0.5
W = torch.randn(4, k, q, device=dev) * 0.1
bias = torch.zeros(q, device=dev)
torch.set_printoptions(linewidth=250, precision=3, sci_mode=False)
for mode in ("fp32", "tf32"):
    lib.set_math_mode(mode)
    Z, Pc, Pr, Pq = ops.graph_layer_fwd(H, col, csrT_ptr, csrT_edge, W, bias, B, N, M, False, False)
    dH, dW, dB = ops.graph_layer_bwd(g, H, Z, col, csrT_ptr, csrT_edge, W, Pc, Pr, Pq, B, N, M, False, False, False, True)
    torch.cuda.synchronize()
    print(mode, "dW1 / c:\n", (dW[0] / c).cpu())
# second probe: H[e, j] = 1 only for j == 3; g[e, n] = 1 only for n == 5, varying with e
H2 = torch.zeros(c, k, device=dev); H2[:, 3] = 1.0
g2 = torch.zeros(c, q, device=dev); g2[:, 5] = 1.0
for mode in ("fp32", "tf32"):
    lib.set_math_mode(mode)
    Z, Pc, Pr, Pq = ops.graph_layer_fwd(H2, col, csrT_ptr, csrT_edge, W, bias, B, N, M, False, False)
    dH, dW, dB = ops.graph_layer_bwd(g2, H2, Z, col, csrT_ptr, csrT_edge, W, Pc, Pr, Pq, B, N, M, False, False, False, True)
    torch.cuda.synchronize()
    print(mode, "probe2 dW1 / c nonzeros:", [(int(i), int(j), round(float(dW[0][i, j] / c), 3)) for i, j in (dW[0].abs() > 1e-3).nonzero().tolist()][:40])
