"""CPU: the device's fast Poseidon permutation (csrc/poseidon_fast.cuh) built for the host — portable bodies stand in for the inline
PTX — against the oracle's plain 30-round permutation and the reference's known answers.  Pins the algorithm of the hot kernel
(lazy reductions, frequency-domain MDS on 22-bit limbs, compact round loop) without a GPU."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from tests import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "poseidon_fast_host.cpp")
LIB = os.path.join(HERE, "native", "libposeidon_fast_host.so")
CSRC = os.path.join(HERE, "..", "zk_evm_b200", "csrc")
P = oracle_lib.P
HASH_ZEROS = [4330397376401421145, 14124799381142128323, 8742572140681234676, 14345658006221440202]   # smt_trie/src/keys.rs:10-15


@pytest.fixture(scope="module")
def host():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("poseidon_fast.cuh", "poseidon.cuh", "gl.cuh", "poseidon_constants.h", "poseidon_fast_constants.h")]
    deps = [d for d in deps if os.path.exists(d)]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-Wno-unknown-pragmas", "-I", CSRC, "-o", LIB, SRC])
    return C.CDLL(LIB)


def test_fast_permutation_matches_oracle_and_kat(host):
    orc = oracle_lib.load()
    rng = np.random.default_rng(7)
    x = oracle_lib.rand_field(rng, (4096, 12))
    x[0] = 0
    x[1] = P - 1
    x[2] = 0xFFFFFFFF
    x[3] = 0xFFFFFFFF00000000
    x[4, ::2] = P - 1
    got = x.copy()
    host.pf_permute_host(got.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(got.shape[0]))
    want = orc.poseidon(x)
    assert np.array_equal(got, want)
    assert [int(v) for v in got[0, :4]] == HASH_ZEROS
    # non-canonical inputs (>= p) are legal for the fast form: same result as their canonical representatives
    y = x[:64].copy()
    y[:, :] = (y % np.uint64(2 ** 32 - 1))           # small values v, so that v + p still fits in 64 bits
    z = y + np.uint64(P)
    a, b = y.copy(), z.copy()
    host.pf_permute_host(a.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(64))
    host.pf_permute_host(b.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(64))
    assert np.array_equal(a, b)


def test_fast_field_ops_on_edge_values(host):
    host.pf_ops_host.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    edge = [0, 1, 2, 0xFFFFFFFF, 0x100000000, P - 1, P, P + 1, 2 ** 64 - 1, 2 ** 63, 0xFFFFFFFF00000000, 0xFFFFFFFEFFFFFFFF]
    rng = np.random.default_rng(8)
    vals = edge + [int(v) for v in rng.integers(0, 2 ** 64, size=40, dtype=np.uint64)]
    out = (C.c_uint64 * 4)()
    for a in vals:
        for b in vals[:16]:
            host.pf_ops_host(a, b, out)
            assert out[0] == a * b % P and out[1] == a * a % P and out[2] == pow(a, 7, P) and out[3] == (a + b) % P, (a, b)
