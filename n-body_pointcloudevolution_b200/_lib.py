"""ctypes binding of the C ABI in include/nbpc.h (libnbpc.so, sm_100a).

The product path has NO fallback: if libnbpc.so is missing, or a call returns a
non-zero status (no GPU, not an sm_100 device, bad argument, launch error), a
RuntimeError is raised with the library's own message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbpc.so")

NBPC_OK = 0
ORDER_DISTANCE = 0
ORDER_INDEX = 1
KNN_MAX_K = 64

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t
_f = ctypes.c_float
_d = ctypes.c_double

# name -> (restype, argtypes); one entry per symbol declared in include/nbpc.h
SIGNATURES = {
    "nbpc_version": (_i, []),
    "nbpc_last_error_string": (ctypes.c_char_p, []),
    "nbpc_device_check": (_i, []),
    "nbpc_set_math_mode": (_i, [_i]),
    "nbpc_get_math_mode": (_i, []),
    "nbpc_set_knn_kernel": (_i, [_i]),
    "nbpc_get_knn_kernel": (_i, []),
    "nbpc_launch_count": (ctypes.c_longlong, []),
    "nbpc_prof_enable": (_i, [_i]),
    "nbpc_prof_report": (ctypes.c_longlong, [ctypes.c_char_p, _sz]),
    "nbpc_knn_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "nbpc_knn": (_i, [_p, _i64, _i64, _i, _i, _i, _i, _d, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "nbpc_pad_cube_workspace_bytes": (_sz, [_i]),
    "nbpc_pad_cube_count": (_i, [_p, _i64, _i, _d, _p, _p, _sz, _p]),
    "nbpc_pad_cube_emit": (_i, [_p, _i64, _i, _d, _p, _p, _p, _p]),
    "nbpc_adjacency_workspace_bytes": (_sz, [_i, _i, _i]),
    "nbpc_adjacency": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "nbpc_segment_csr_workspace_bytes": (_sz, [_i64, _i]),
    "nbpc_segment_csr": (_i, [_p, _i64, _i, _p, _p, _p, _p, _sz, _p]),
    "nbpc_edge_features_za": (_i, [_p, _i, _p, _i, _p, _p, _i64, _i, _i, _p, _p]),
    "nbpc_edge_features": (_i, [_p, _i, _p, _i, _i, _p, _p]),
    "nbpc_include_node_features": (_i, [_p, _i, _p, _i, _i, _p, _p, _i, _i, _p, _p]),
    "nbpc_segment_reduce": (_i, [_p, _i, _p, _p, _i, _i, _p, _p]),
    "nbpc_gather_rows": (_i, [_p, _i, _p, _i64, _p, _p, _p]),
    "nbpc_graph_layer_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "nbpc_graph_layer_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "nbpc_graph_layer_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i,
                                  _p, _p, _p, _p, _sz, _p]),
    "nbpc_graph_layer_rowpool_supported": (_i, [_i, _i, _i, _i]),
    "nbpc_graph_layer_fwd_rp": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _i, _p, _p, _sz, _p]),
    "nbpc_graph_layer_bwd_rp": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i,
                                     _p, _p, _p, _p, _p, _p, _sz, _p]),
    "nbpc_sym_adjacency_workspace_bytes": (_sz, [_i, _i]),
    "nbpc_sym_adjacency_count": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "nbpc_sym_adjacency_emit": (_i, [_p, _p, _p, _p, _i, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "nbpc_graph15_workspace_bytes": (_sz, [_i, _i, _i64, _i, _i]),
    "nbpc_graph15_layer_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "nbpc_graph15_layer_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i64, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p,
                                    _p, _sz, _p]),
    "nbpc_graph_layer_vin_supported": (_i, [_i, _i, _i, _i64]),
    "nbpc_graph_layer_fwd_v": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "nbpc_graph_layer_bwd_v": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i,
                                    _p, _p, _p, _p, _sz, _p]),
    "nbpc_set_layer_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "nbpc_set_layer_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _sz, _p]),
    "nbpc_set_layer_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "nbpc_set_layer_fwd_chained": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p, _p, _sz, _p]),
    "nbpc_set_layer_bwd_chained": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "nbpc_loss_workspace_bytes": (_sz, [_i64]),
    "nbpc_loss_za_fwd": (_i, [_p, _i, _p, _i, _i64, _p, _p, _sz, _p]),
    "nbpc_loss_za_bwd": (_i, [_p, _i, _p, _i, _i64, _p, _p, _i, _p]),
    "nbpc_pbc_loss_fwd": (_i, [_p, _i, _p, _i, _i64, _i, _p, _p, _sz, _p]),
    "nbpc_pbc_loss_bwd": (_i, [_p, _i, _p, _i, _i64, _i, _p, _p, _i, _p]),
    "nbpc_periodic_boundary_dist": (_i, [_p, _i, _p, _i, _i64, _p, _p]),
    "nbpc_readout": (_i, [_p, _i64, _i, _p, _p]),
    "nbpc_residual_update": (_i, [_p, _i, _p, _i, _i64, _f, _f, _p, _p]),
    "nbpc_linear": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _p, _p]),
    "nbpc_xty_workspace_bytes": (_sz, [_i64, _i, _i]),
    "nbpc_xty": (_i, [_p, _p, _i64, _i, _i, _p, _p, _sz, _p]),
    "nbpc_adam_tf": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i64, _f, _p]),
    "nbpc_adam_tf_dev": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _p, _f, _p]),
}


def bind(cdll):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


class VirtualInput(ctypes.Structure):
    """nbpc_virtual_input (include/nbpc.h)"""
    _fields_ = [("E", ctypes.c_void_p), ("W1", ctypes.c_void_p), ("Q_col", ctypes.c_void_p), ("Q_row", ctypes.c_void_p),
                ("k", ctypes.c_int)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C n-body_pointcloudevolution_b200/csrc`). There is no CPU fallback.")
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


def last_error():
    return load().nbpc_last_error_string().decode("utf-8", "replace")


def check(rc, what):
    if rc != NBPC_OK:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


MATH_MODES = {"fp32": 0, "tf32x3": 1, "tf32": 2}


def set_math_mode(mode):
    """Arithmetic of the edge-level channel projections: 'fp32' (CUDA cores), 'tf32x3' (tcgen05, error-compensated,
    FP32-class accuracy) or 'tf32' (tcgen05, one pass).  Process-global; see include/nbpc.h."""
    check(load().nbpc_set_math_mode(MATH_MODES[mode] if isinstance(mode, str) else int(mode)), "nbpc_set_math_mode")


def get_math_mode():
    m = int(load().nbpc_get_math_mode())
    return {v: k for k, v in MATH_MODES.items()}[m]


KNN_KERNELS = {"auto": 0, "thread": 2, "warp": 3}


def set_knn_kernel(kernel):
    """Query kernel of nbpc_knn: 'auto' (by k), 'thread' (one thread per query) or 'warp' (one warp per query, k <= 32).
    Same result bit for bit; process-global; see include/nbpc.h."""
    check(load().nbpc_set_knn_kernel(KNN_KERNELS[kernel] if isinstance(kernel, str) else int(kernel)), "nbpc_set_knn_kernel")


def get_knn_kernel():
    k = int(load().nbpc_get_knn_kernel())
    return {v: n for n, v in KNN_KERNELS.items()}.get(k, str(k))


def launch_count():
    """Kernels launched by libnbpc.so so far in this process."""
    return int(load().nbpc_launch_count())


def prof_enable(on=True):
    load().nbpc_prof_enable(int(bool(on)))


def prof_report():
    """{kernel name: (launches, total_ms)} since prof_enable(True); waits for the recorded events."""
    L = load()
    need = L.nbpc_prof_report(None, 0)
    buf = ctypes.create_string_buffer(int(need) + 16)
    L.nbpc_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name] = (int(cnt), float(ms))
    return out
