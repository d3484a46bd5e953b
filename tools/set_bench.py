"""Set-model training step at the reference's default widths (bench.py extras.set_model), stand-alone."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
nb = importlib.import_module("n-body_pointcloudevolution_b200")
for mode in sys.argv[1:] or ["tf32x3"]:
    nb.set_math_mode(mode)
    print(json.dumps(bench.set_model_bench(nb, torch.device("cuda"), bench.measured_peak_gbs()[0]), indent=1))
