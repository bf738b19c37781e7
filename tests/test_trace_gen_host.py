"""CPU checks of the device-side trace finishing code: the single-source Keccak row generator
(zk_evm_b200/csrc/stark/keccak_trace.h) built for the host, against the Python restatement of the reference's generate_trace_rows
(tests/traces.py keccak_trace, keccak_stark.rs:70-250), against hashlib's Keccak-f through SHA3, and through the oracle's prover +
verifier (the generated trace satisfies every KeccakStark constraint)."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from tests import traces

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "zk_evm_b200", "csrc")
u64p = C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "native", "keccak_trace_host.cpp")
    lib = os.path.join(HERE, "native", "libkeccak_trace_host.so")
    deps = [src] + [os.path.join(CSRC, "stark", f) for f in ("keccak_trace.h", "table_keccak.h", "logic_trace.h", "table_logic.h", "memory_trace.h", "table_memory.h")] + \
           [os.path.join(CSRC, "gl.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", CSRC, "-o", lib, src])
    h = C.CDLL(lib)
    h.keccak_trace_rows.restype = C.c_uint32
    h.logic_trace_rows.restype = C.c_uint32
    return h


def _rows(host, inputs, ts, n):
    out = np.full((2431, n), 0xDEADBEEF, dtype=np.uint64)
    cells = host.keccak_trace_rows(inputs.ctypes.data_as(u64p), ts.ctypes.data_as(u64p), C.c_uint64(inputs.shape[0]), C.c_size_t(n),
                                   out.ctypes.data_as(u64p))
    assert cells == 2431
    return out


@pytest.mark.parametrize("nperm,log_n", [(1, 5), (5, 7), (10, 8), (0, 4)])
def test_keccak_rows_match_reference_restatement(host, nperm, log_n):
    rng = np.random.default_rng(100 + nperm)
    inputs = rng.integers(0, 1 << 64, size=(nperm, 25), dtype=np.uint64)
    ts = rng.integers(1, 1 << 30, size=(nperm,), dtype=np.uint64)
    got = _rows(host, inputs, ts, 1 << log_n)
    want, _ = traces.keccak_trace(log_n, inputs, ts) if nperm else (np.zeros((2431, 1 << log_n), dtype=np.uint64), None)
    assert np.array_equal(got, want)      # every cell written exactly once (the 0xDEADBEEF fill is gone), padding rows zero


def test_keccak_permutation_is_keccak_f1600(host):
    """the plain round function the rows are advanced with is Keccak-f[1600]: SHA3-256 of a one-block message from it == hashlib"""
    msg = b"zk_evm_b200 device-side trace finishing"
    block = bytearray(136)
    block[:len(msg)] = msg
    block[len(msg)] ^= 0x06
    block[135] ^= 0x80
    state = np.zeros(25, dtype=np.uint64)
    state[:17] = np.frombuffer(bytes(block), dtype="<u8")
    out = np.zeros(25, dtype=np.uint64)
    host.keccak_permutation_output(state.ctypes.data_as(u64p), out.ctypes.data_as(u64p))
    assert out[:4].tobytes() == hashlib.sha3_256(msg).digest()


def test_generated_keccak_trace_proves_and_verifies(host):
    """the finished trace satisfies the KeccakStark constraints: the oracle proves it and its verifier accepts (table 3)"""
    from tests import oracle_lib
    orc = oracle_lib.load()
    rng = np.random.default_rng(7)
    inputs = rng.integers(0, 1 << 64, size=(2, 25), dtype=np.uint64)
    ts = np.array([3, 9], dtype=np.uint64)
    tr = _rows(host, inputs, ts, 64)
    bg = np.array([11, 22], dtype=np.uint64)
    st0 = np.arange(12, dtype=np.uint64)
    proof, st = oracle_lib.orc_prove_table(orc, traces.T_KECCAK, oracle_lib.TEST_CONFIG, tr, bg, st0)
    ok, err, st2 = oracle_lib.orc_verify_table(orc, traces.T_KECCAK, oracle_lib.TEST_CONFIG, proof, bg, st0)
    assert ok, err
    assert np.array_equal(st, st2)


# ---- LogicStark ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nops,log_n", [(1, 0), (5, 3), (200, 8), (0, 2)])
def test_logic_rows_match_reference_restatement(host, nops, log_n):
    ops = traces.logic_ops(nops, 50 + nops)
    n = 1 << log_n
    out = np.full((523, n), 0xDEADBEEF, dtype=np.uint64)
    assert host.logic_trace_rows(ops.ctypes.data_as(u64p), C.c_uint64(nops), C.c_size_t(n), out.ctypes.data_as(u64p)) == 523
    assert np.array_equal(out, traces.logic_trace_from_ops(log_n, ops))
    # and operation by operation against Python integers (Op::result, logic.rs:131-139)
    for i in range(min(nops, 20)):
        a = sum(int(ops[i, 1 + l]) << (64 * l) for l in range(4))
        b = sum(int(ops[i, 5 + l]) << (64 * l) for l in range(4))
        r = (a & b, a | b, a ^ b)[int(ops[i, 0])]
        assert sum(int(out[515 + k, i]) << (32 * k) for k in range(8)) == r
        assert sum(int(out[3 + k, i]) << k for k in range(256)) == a and sum(int(out[259 + k, i]) << k for k in range(256)) == b


def test_generated_logic_trace_proves_and_verifies(host):
    from tests import oracle_lib
    orc = oracle_lib.load()
    ops = traces.logic_ops(40, 3)
    out = np.zeros((523, 64), dtype=np.uint64)
    assert host.logic_trace_rows(ops.ctypes.data_as(u64p), C.c_uint64(40), C.c_size_t(64), out.ctypes.data_as(u64p)) == 523
    bg = np.array([11, 22], dtype=np.uint64)
    st0 = np.arange(12, dtype=np.uint64)
    proof, st = oracle_lib.orc_prove_table(orc, traces.T_LOGIC, oracle_lib.TEST_CONFIG, out, bg, st0)
    ok, err, st2 = oracle_lib.orc_verify_table(orc, traces.T_LOGIC, oracle_lib.TEST_CONFIG, proof, bg, st0)
    assert ok, err
    assert np.array_equal(st, st2)


# ---- MemoryStark ---------------------------------------------------------------------------------------------------------------------------
def _mem_finish(host, ops, stale=()):
    n = ops.shape[1]
    ops = np.ascontiguousarray(ops, dtype=np.uint64)
    st = np.array(list(stale) or [0], dtype=np.uint64)
    out = np.zeros((30, n), dtype=np.uint64)
    ok = host.memory_finish_rows(ops.ctypes.data_as(u64p), C.c_size_t(n), st.ctypes.data_as(u64p), C.c_size_t(len(stale)), out.ctypes.data_as(u64p))
    return ok, out


@pytest.mark.parametrize("seed,log_n,stale", [(1, 6, ()), (2, 6, (2,)), (3, 7, (1, 2)), (4, 6, (0, 1, 2))])
def test_memory_rows_match_reference_restatement(host, seed, log_n, stale):
    ops = traces.memory_sorted_ops(log_n, seed)
    ok, got = _mem_finish(host, ops, stale)
    assert ok == 1
    assert np.array_equal(got, traces.memory_finish_reference(ops, stale))


def test_memory_finish_reproduces_hand_built_valid_traces_and_flags_bad_order(host):
    t = traces.memory_trace_simple(6)
    ok, got = _mem_finish(host, t[list(traces.MEM_OP_COLS)])
    assert ok == 1 and np.array_equal(got, t)
    ops = traces.memory_sorted_ops(6, 5)
    ops[1, 10] += 1000                      # a timestamp gap no 64-row table can range-check
    ok, _ = _mem_finish(host, ops)
    assert ok == 0


def test_finished_memory_trace_proves_and_verifies(host):
    from tests import oracle_lib
    orc = oracle_lib.load()
    ok, tr = _mem_finish(host, traces.memory_sorted_ops(6, 8), (2,))
    assert ok == 1
    bg = np.array([11, 22], dtype=np.uint64)
    st0 = np.arange(12, dtype=np.uint64)
    proof, st = oracle_lib.orc_prove_table(orc, traces.T_MEMORY, oracle_lib.TEST_CONFIG, tr, bg, st0)
    ok, err, st2 = oracle_lib.orc_verify_table(orc, traces.T_MEMORY, oracle_lib.TEST_CONFIG, proof, bg, st0)
    assert ok, err
