"""CPU: the lookup / cross-table-lookup constraint interpreter the CUDA quotient kernels run per point (csrc/stark/checks.h over the flat
descriptors of stark/lookup.h, incl. the plain-cell column encoding) executed on the host on arbitrary rows, against the oracle's
eval_vanishing_poly, for all nine tables and both challenge counts."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from tests import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "flat_eval_host.cpp")
LIB = os.path.join(HERE, "native", "libflat_eval_host.so")
CSRC = os.path.join(HERE, "..", "zk_evm_b200", "csrc")
ORACLE = os.path.join(HERE, "..", "oracle")
NUM_COLUMNS = (116, 71, 85, 2431, 438, 523, 30, 12, 12)


@pytest.fixture(scope="module")
def host():
    deps = [SRC] + [os.path.join(CSRC, "stark", f) for f in os.listdir(os.path.join(CSRC, "stark"))] + \
           [os.path.join(ORACLE, f) for f in os.listdir(ORACLE) if f.endswith(".h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-Wno-unknown-pragmas", "-I", CSRC, "-I", ORACLE,
                               "-o", LIB, SRC])
    return C.CDLL(LIB)


@pytest.mark.parametrize("table", range(9))
@pytest.mark.parametrize("nch", [1, 2])
def test_flat_interpreter_matches_oracle(host, table, nch):
    rng = np.random.default_rng(1000 + 10 * table + nch)
    u64p = C.POINTER(C.c_uint64)
    p = lambda a: a.ctypes.data_as(u64p)
    nc = NUM_COLUMNS[table]
    lv, nv = oracle_lib.rand_field(rng, (nc,)), oracle_lib.rand_field(rng, (nc,))
    naux = C.c_uint32()
    dummy = np.zeros(4096, dtype=np.uint64)
    al, be, ga, sel = (oracle_lib.rand_field(rng, (k,)) for k in (2, 2, 2, 3))
    lab = np.array(oracle_lib.DEFAULT_LABELS, dtype=np.uint64)
    a, b = np.zeros(2, np.uint64), np.zeros(2, np.uint64)
    alv, anv = oracle_lib.rand_field(rng, (4096,)), oracle_lib.rand_field(rng, (4096,))
    r = host.flat_eval_agrees(C.c_uint32(table), C.c_uint32(nch), p(lv), p(nv), p(alv), p(anv), p(al), p(be), p(ga), p(sel), p(lab),
                              p(a), p(b), C.byref(naux))
    assert r == 1, (r, a, b)
    assert (naux.value & 0x7FFFFFFF) <= 4096 and a[0] != 0
    # with two challenges every table's CTL items come in twins and the shared-walk evaluator is the one that ran
    assert bool(naux.value >> 31) == (nch == 2)
