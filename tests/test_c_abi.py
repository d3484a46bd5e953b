"""CPU: the C-ABI shared library loads and exports every symbol include/nbpc.h declares, the Python
signature table matches the header, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nbpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(nb):
    lib = nb._lib.load()                      # raises if libnbpc.so is missing
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libnbpc.so does not export {s}"
    assert sorted(nb._lib.SIGNATURES) == syms, "python signature table out of sync with include/nbpc.h"
    assert lib.nbpc_version() >= 100


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "nbpc.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text and 'extern "C"' in text


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu(nb):
    lib = nb._lib.load()
    assert lib.nbpc_device_check() == -2      # NBPC_EARCH
    assert b"no CPU fallback" in lib.nbpc_last_error_string()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.graph.get_kneighbor_list(torch.rand(1, 64, 3), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.nn.set_layer(torch.rand(1, 8, 3), ([torch.rand(3, 2)], torch.zeros(2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.ops.graph_layer_fwd(torch.rand(8, 3), torch.zeros(8, dtype=torch.int32), torch.zeros(3, dtype=torch.int32),
                               torch.zeros(8, dtype=torch.int32), torch.rand(4, 3, 2), torch.zeros(2), 1, 2, 4, False, False)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "n-body_pointcloudevolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libnbpc_emu" not in src or f == "nbpc_common.cuh", f
