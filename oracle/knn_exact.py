"""TEST INFRASTRUCTURE ONLY - ctypes wrapper for oracle/knn_exact.c (builds it on demand)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libknn_exact.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "knn_exact.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libknn_exact.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.knn_exact.restype = ctypes.c_int
        _lib.knn_exact.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p]
    return _lib


def knn_exact(x64, n_query, k, include_self=True, return_d2=False, return_ties=False):
    """x64 (P,3) float64; queries = rows [0,n_query).  Returns (n_query,k) int64 indices in
    ascending (d2, index) order [, d2 (n_query,k) float64][, row_has_tie (n_query,) bool]."""
    lib = _load()
    x64 = np.ascontiguousarray(x64, dtype=np.float64)
    P = x64.shape[0]
    idx = np.empty((n_query, k), dtype=np.int64)
    d2 = np.empty((n_query, k), dtype=np.float64)
    ties = np.zeros((n_query,), dtype=np.uint8)
    nthreads = max(1, min(os.cpu_count() or 1, n_query // 256 or 1))
    bounds = np.linspace(0, n_query, nthreads + 1).astype(np.int64)

    def run(t):
        return lib.knn_exact(x64.ctypes.data, P, int(bounds[t]), int(bounds[t + 1]), k,
                             int(bool(include_self)), idx.ctypes.data, d2.ctypes.data,
                             ties.ctypes.data)
    if nthreads == 1:
        rcs = [run(0)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthreads) as ex:
            rcs = list(ex.map(run, range(nthreads)))
    if any(rcs):
        raise ValueError(f"knn_exact failed rc={rcs} (k={k}, P={P}, n_query={n_query})")
    out = (idx,)
    if return_d2:
        out += (d2,)
    if return_ties:
        out += (ties.astype(bool),)
    return out[0] if len(out) == 1 else out
