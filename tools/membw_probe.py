"""Probe (torch library kernels only; not product code): write-only / read-only / copy HBM bandwidth and
L2-resident re-read bandwidth on this B200.  Used to set the floors quoted in DESIGN.md."""
import torch, json
dev = "cuda"
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
out = {}
for mb in (64, 470, 2048):
    n = mb * 1024 * 1024 // 4
    x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
    x.normal_()
    out[f"fill_{mb}MB_GBs"] = n * 4 / t(lambda: y.fill_(1.0)) / 1e6
    out[f"sum_{mb}MB_GBs"] = n * 4 / t(lambda: x.sum()) / 1e6
    out[f"copy_{mb}MB_GBs(rd+wr)"] = 2 * n * 4 / t(lambda: y.copy_(x)) / 1e6
    out[f"relu_inplace_{mb}MB_GBs(rd+wr)"] = 2 * n * 4 / t(lambda: x.relu_()) / 1e6
    del x, y
# gather of 128-byte rows from a table of T MB (random rows), 470 MB of output
for tmb in (8, 32, 64, 128, 512):
    rows = tmb * 1024 * 1024 // 128
    tab = torch.randn(rows, 32, device=dev)
    idx = torch.randint(0, rows, (3670016,), device=dev)
    o = torch.empty(3670016, 32, device=dev)
    out[f"gather128B_table{tmb}MB_us"] = t(lambda: torch.index_select(tab, 0, idx, out=o)) * 1e3
print(json.dumps(out, indent=1))
