"""CPU: the product's NTT tile code (csrc/ntt_tile.cuh is host+device) executed on the host by tests/native/ntt_tile_host.cpp with
ntt.cu's pass plan, against the oracle's fft / ifft / coset_fft (plonky2_field fft.rs restatement).  Pins the pass/round index
arithmetic and the power-of-two butterfly network before any GPU time is spent; the CUDA kernel wraps the same header."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from tests import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "ntt_tile_host.cpp")
LIB = os.path.join(HERE, "native", "libntt_tile_host.so")
CSRC = os.path.join(HERE, "..", "zk_evm_b200", "csrc")
GENERATOR = 14293326489335486720
P = oracle_lib.P


@pytest.fixture(scope="module")
def host():
    deps = [SRC, os.path.join(CSRC, "ntt_tile.cuh"), os.path.join(CSRC, "gl.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", CSRC, "-o", LIB, SRC])
    return C.CDLL(LIB)


def bitrev_perm(lg):
    i = np.arange(1 << lg, dtype=np.uint64)
    r = np.zeros_like(i)
    for b in range(lg):
        r |= ((i >> np.uint64(b)) & np.uint64(1)) << np.uint64(lg - 1 - b)
    return r.astype(np.int64)


def run(host, x, lg, inverse, prescale=None):
    a = np.ascontiguousarray(x, dtype=np.uint64).copy()
    u64p = C.POINTER(C.c_uint64)
    pre = None if prescale is None else np.ascontiguousarray(prescale, dtype=np.uint64)
    host.ntt_tile_host(a.ctypes.data_as(u64p), C.c_size_t(a.shape[0]), C.c_uint(lg), int(inverse),
                       pre.ctypes.data_as(u64p) if pre is not None else None)
    return a


@pytest.mark.parametrize("lg", [1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 16, 17])
def test_tile_transform_matches_oracle(host, lg):
    orc = oracle_lib.load()
    rng = np.random.default_rng(100 + lg)
    x = oracle_lib.rand_field(rng, (2, 1 << lg))
    x[1, : min(4, 1 << lg)] = P - 1          # edge values
    br = bitrev_perm(lg)
    # DIF: natural in -> bit-reversed out
    assert np.array_equal(run(host, x, lg, False)[:, br], orc.ntt(x, 0))
    ninv = pow(1 << lg, P - 2, P)
    want = orc.ntt(x, 1)
    got = run(host, x, lg, True)[:, br]
    got = np.array([[int(v) * ninv % P for v in row] for row in got], dtype=np.uint64) if lg <= 12 else None
    if got is not None:
        assert np.array_equal(got, want)
    # coset transform through the prescale table
    pre = np.array([pow(GENERATOR, j, P) for j in range(1 << lg)], dtype=np.uint64) if lg <= 14 else None
    if pre is not None:
        assert np.array_equal(run(host, x, lg, False, pre)[:, br], orc.ntt(x, 2, GENERATOR))


@pytest.mark.parametrize("lg", [21])
def test_tile_transform_three_passes(host, lg):
    """L = 21 takes two strided passes and a final one; checked through linearity + an impulse (no big oracle transform needed)"""
    x = np.zeros((1, 1 << lg), dtype=np.uint64)
    x[0, 1] = 1                              # fft of delta_1 = w^k
    out = run(host, x, lg, False)
    orc = oracle_lib.load()
    w = int(orc.lib.orc_root_of_unity(lg))
    br = bitrev_perm(lg)
    nat = out[0, br]
    ks = [0, 1, 2, 3, 12345, (1 << lg) - 1, 1 << 20, (1 << 20) + 7]
    for k in ks:
        assert int(nat[k]) == pow(w, k, P)
