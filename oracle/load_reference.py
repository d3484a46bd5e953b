"""TEST INFRASTRUCTURE ONLY.

Imports the *unmodified* reference modules /root/reference/graph.py and
/root/reference/nn.py with `oracle.tf_shim` standing in for TensorFlow
(recipe: SURVEY.md Appendix A).  Only usable in the authoring container, where
/root/reference exists; the GPU box never calls this (golden vectors produced
with it are committed under tests/golden/).
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("NBPC_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "graph.py"))


def load_reference():
    """Returns (graph_module, nn_module) of the real reference."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
    from oracle import tf_shim

    saved = {k: sys.modules.get(k) for k in ("tensorflow", "utils", "graph", "nn")}
    sys.modules["tensorflow"] = tf_shim
    # real utils.py touches $HOME, yaml and a data dir at import (utils.py:92-105)
    sys.modules["utils"] = types.ModuleType("utils")
    sys.path.insert(0, REFERENCE_DIR)
    try:
        for name in ("graph", "nn"):
            sys.modules.pop(name, None)
        ref_graph = importlib.import_module("graph")
        ref_nn = importlib.import_module("nn")
        assert os.path.dirname(ref_graph.__file__) == REFERENCE_DIR
        assert os.path.dirname(ref_nn.__file__) == REFERENCE_DIR
    finally:
        sys.path.remove(REFERENCE_DIR)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref_graph, ref_nn


def load_experiment_functions(names=("loss", "set_transform", "res_layer", "attn_layer", "net_fwd")):
    """experiment.py builds its TF graph, loads a data set and opens a session at import, so it cannot be imported.
    Its layer / model FUNCTIONS (experiment.py:37-157) are extracted with `ast` and compiled UNMODIFIED into a fresh
    namespace whose `tf` is oracle.tf_shim; the caller fills in the module-level variables those functions read
    (Wf, Wg, Wh, Rset, Bset, kdims, num_layers, X_in)."""
    import ast

    import numpy as np

    from oracle import tf_shim
    path = os.path.join(REFERENCE_DIR, "experiment.py")
    if not os.path.isfile(path):
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in keep} == set(names), "experiment.py no longer defines the expected functions"
    ns = {"tf": tf_shim, "np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns
