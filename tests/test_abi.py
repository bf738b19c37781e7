"""The C-ABI library loads and exports every symbol include/zkgpu.h declares (no compute calls: CPU only)."""
import ctypes
import os
import pytest
import zk_evm_b200
from zk_evm_b200 import _lib


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    names = _lib.declared_symbols()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/zkgpu.h but not exported: %s" % missing


def test_version_and_error_string():
    lib = zk_evm_b200.lib()
    assert b"zkgpu" in lib.zkgpu_version()
    assert isinstance(lib.zkgpu_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zk_evm_b200.ZkGpuError):
        zk_evm_b200.Context(0)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dp, _, fns in os.walk(os.path.join(root, "zk_evm_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if "oracle_lib" in txt or "liboracle" in txt or 'include "oracle' in txt or "../oracle" in txt:
                    bad.append(fn)
    assert not bad, "product code references the oracle: %s" % bad
