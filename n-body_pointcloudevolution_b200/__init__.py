"""B200-native (sm_100a) hot path of evdcush/N-Body_PointCloudEvolution: periodic-box kNN graph
construction and set/graph-layer forward+backward, behind the reference's own function names.

    import importlib; nb = importlib.import_module("n-body_pointcloudevolution_b200")   # or: import nbpc as nb
    nb.graph.get_kneighbor_list(...), nb.graph.shift_inv_layer(...), nb.nn.set_layer(...), ...

All compute runs in libnbpc.so (hand-written CUDA, C ABI in include/nbpc.h).  There is no CPU
fallback: calls raise RuntimeError without an sm_100 GPU.
"""
from . import _lib  # noqa: F401
from . import synthetic  # noqa: F401
from ._lib import get_knn_kernel, set_knn_kernel  # noqa: F401  (thread-per-query | warp-per-query kNN kernel)
from ._lib import get_math_mode, set_math_mode  # noqa: F401  (fp32 | tf32x3 | tf32 edge-level arithmetic)

__all__ = ["_lib", "synthetic", "ops", "graph", "nn", "train_utils"]


def __getattr__(name):
    # ops / graph / nn import torch; keep `import package` light for tooling
    if name in ("ops", "graph", "nn", "train_utils"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
