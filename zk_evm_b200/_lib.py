"""ctypes binding of libzkgpu.so (the C ABI declared in include/zkgpu.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable every call fails loudly.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkgpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "zkgpu.h")

u64p = C.POINTER(C.c_uint64)
_lib = None


class ZkGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("zkgpu error %d: %s" % (code, msg))
        self.code = code


class StarkConfig(C.Structure):
    """zkgpu_stark_config == starky StarkConfig + FriConfig."""
    _fields_ = [("security_bits", C.c_uint32), ("num_challenges", C.c_uint32), ("rate_bits", C.c_uint32),
                ("cap_height", C.c_uint32), ("proof_of_work_bits", C.c_uint32), ("fri_arity_bits", C.c_uint32),
                ("fri_final_poly_bits", C.c_uint32), ("num_query_rounds", C.c_uint32)]

    @classmethod
    def standard_fast(cls):
        # StarkConfig::standard_fast_config (zero/src/prover_state/mod.rs:283)
        return cls(100, 2, 1, 4, 16, 4, 5, 84)

    @classmethod
    def test(cls):
        # TEST_STARK_CONFIG (evm_arithmetization/src/testing_utils.rs:41-52)
        return cls(1, 1, 1, 4, 1, 4, 5, 1)


class KernelLabels(C.Structure):
    _fields_ = [("halt_final", C.c_uint64), ("init", C.c_uint64), ("syscall_jumptable", C.c_uint64),
                ("exception_jumptable", C.c_uint64)]


def declared_symbols():
    """Every function name include/zkgpu.h declares (used by the ABI test)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(zkgpu_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZkGpuError(-2, "libzkgpu.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                 "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.zkgpu_last_error.restype = C.c_char_p
        _lib.zkgpu_version.restype = C.c_char_p
        _lib.zkgpu_ctx_destroy.restype = None
        _lib.zkgpu_batch_free.restype = None
        if hasattr(_lib, "zkgpu_dev_trace_ptr"):
            _lib.zkgpu_dev_trace_ptr.restype = C.c_void_p
        for name in ("zkgpu_ctl_free", "zkgpu_proof_free", "zkgpu_challenger_free", "zkgpu_table_job_free", "zkgpu_upload_free",
                     "zkgpu_dev_trace_free"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = None
    return _lib


def check(rc):
    if rc != 0:
        raise ZkGpuError(rc, lib().zkgpu_last_error().decode())
