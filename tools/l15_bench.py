import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
nb = importlib.import_module("n-body_pointcloudevolution_b200")
nb._lib.prof_enable(True)
r = bench.layer15_bench(nb, torch.device("cuda"), bench.measured_peak_gbs()[0])
rep = nb._lib.prof_report()
print(json.dumps(r, indent=1))
for n, (c, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{n:40s} {c:5d} {t:9.3f} ms")
