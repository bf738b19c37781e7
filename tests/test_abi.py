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


# ---- error behaviour of the entry points that need no device: status codes instead of exceptions / aborts (include/zkgpu.h:11-18) ----
ZKGPU_ERR_INVALID = -1
u64p = ctypes.POINTER(ctypes.c_uint64)


def _arr(vals):
    return (ctypes.c_uint64 * len(vals))(*vals)


def test_table_info_shapes_and_bad_ids():
    """the auxiliary-column layout of SURVEY 8(a)'s table (trace width; lookup columns + CTL helpers + CTL Zs for two challenges)"""
    lib = zk_evm_b200.lib()
    want = {0: (116, 100), 1: (71, 70), 2: (85, 24), 3: (2431, 4), 4: (438, 290), 5: (523, 2), 6: (30, 16), 7: (12, 4), 8: (12, 2)}
    for t, (c, a) in want.items():
        nc, nl, nh, nz = (ctypes.c_uint32() for _ in range(4))
        assert lib.zkgpu_table_info(t, 2, ctypes.byref(nc), ctypes.byref(nl), ctypes.byref(nh), ctypes.byref(nz)) == 0
        assert (nc.value, nl.value + nh.value + nz.value) == (c, a), t
        # with one challenge (TEST_STARK_CONFIG) every per-challenge count halves
        nl1, nh1, nz1 = (ctypes.c_uint32() for _ in range(3))
        assert lib.zkgpu_table_info(t, 1, ctypes.byref(nc), ctypes.byref(nl1), ctypes.byref(nh1), ctypes.byref(nz1)) == 0
        assert 2 * (nl1.value + nh1.value + nz1.value) == a, t
    nc = ctypes.c_uint32()
    assert lib.zkgpu_table_info(9, 2, ctypes.byref(nc), ctypes.byref(nc), ctypes.byref(nc), ctypes.byref(nc)) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_last_error()                      # a message is kept for the caller
    assert lib.zkgpu_table_info(0, 2, None, ctypes.byref(nc), ctypes.byref(nc), ctypes.byref(nc)) in (0, ZKGPU_ERR_INVALID)   # never a crash


def test_merkle_block_words_and_bad_splits():
    lib = zk_evm_b200.lib()
    lib.zkgpu_merkle_block_words.argtypes = [ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_size_t)]
    w = ctypes.c_size_t()
    # 2^10 leaves, cap height 4: levels 1024 .. 16 digests of 4 words, split over k blocks
    full = 4 * sum(1024 >> l for l in range(7))
    for k in (1, 2, 4, 8, 16):
        assert lib.zkgpu_merkle_block_words(1024, 4, k, ctypes.byref(w)) == 0 and w.value == full // k
    for bad in (0, 3, 32):                             # not a power of two / more blocks than cap entries
        assert lib.zkgpu_merkle_block_words(1024, 4, bad, ctypes.byref(w)) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_merkle_block_words(8, 4, 1, ctypes.byref(w)) == ZKGPU_ERR_INVALID      # cap higher than the tree
    assert lib.zkgpu_merkle_block_words(1024, 4, 1, None) == ZKGPU_ERR_INVALID


def test_host_transcript_entry_points_reject_bad_arguments():
    lib = zk_evm_b200.lib()
    ch = ctypes.c_void_p()
    assert lib.zkgpu_challenger_new(ctypes.byref(ch)) == 0
    assert lib.zkgpu_challenger_new(None) == ZKGPU_ERR_INVALID
    P = 0xFFFFFFFF00000001
    assert lib.zkgpu_challenger_observe(ch, _arr([1, 2, 3]), ctypes.c_size_t(3)) == 0
    assert lib.zkgpu_challenger_observe(ch, _arr([P]), ctypes.c_size_t(1)) == ZKGPU_ERR_INVALID          # not canonical
    assert lib.zkgpu_challenger_observe(ch, None, ctypes.c_size_t(1)) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_challenger_observe(ch, None, ctypes.c_size_t(0)) == 0                                  # nothing to observe
    st = (ctypes.c_uint64 * 12)()
    assert lib.zkgpu_challenger_compact(ch, st) == 0 and all(0 <= x < P for x in st)
    assert lib.zkgpu_challenger_compact(ch, None) == ZKGPU_ERR_INVALID
    # a state word is a field element: p stands for 0 (the host permutation works on any u64 congruent mod p), and what comes back is canonical
    a, b = (ctypes.c_uint64 * 12)(), (ctypes.c_uint64 * 12)()
    assert lib.zkgpu_challenger_set_state(ch, _arr([P] + [5] * 11)) == 0 and lib.zkgpu_challenger_observe(ch, _arr([7]), ctypes.c_size_t(1)) == 0
    assert lib.zkgpu_challenger_compact(ch, a) == 0
    assert lib.zkgpu_challenger_set_state(ch, _arr([0] + [5] * 11)) == 0 and lib.zkgpu_challenger_observe(ch, _arr([7]), ctypes.c_size_t(1)) == 0
    assert lib.zkgpu_challenger_compact(ch, b) == 0
    assert list(a) == list(b) and all(x < P for x in a)
    assert lib.zkgpu_challenger_set_state(ch, None) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_challenger_get_challenges(ch, None, ctypes.c_size_t(2)) == ZKGPU_ERR_INVALID
    lib.zkgpu_challenger_free(ch)
    lib.zkgpu_challenger_free(None)                    # freeing nothing is fine, as for every *_free
    for fn in ("zkgpu_batch_free", "zkgpu_ctl_free", "zkgpu_proof_free", "zkgpu_table_job_free", "zkgpu_upload_free", "zkgpu_dev_trace_free"):
        getattr(lib, fn)(None)
    # handles that were never made
    n = ctypes.c_size_t()
    assert lib.zkgpu_batch_dims(None, ctypes.byref(n), ctypes.byref(n), None, None) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_batch_cap(None, _arr([0] * 64)) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_proof_serialize(None, None, ctypes.byref(n)) == ZKGPU_ERR_INVALID
    assert lib.zkgpu_ctl_export(None, None) == ZKGPU_ERR_INVALID
