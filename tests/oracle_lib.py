"""ctypes access to oracle/liboracle.so — the CPU restatement used as the parity checker.

Test infrastructure only: nothing under zk_evm_b200/ imports this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
u64p = C.POINTER(C.c_uint64)
P = 0xFFFFFFFF00000001
GENERATOR = 14293326489335486720


def _ptr(a):
    return a.ctypes.data_as(u64p)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        for f in ("orc_gl_add", "orc_gl_sub", "orc_gl_mul", "orc_gl_inv", "orc_gl_pow", "orc_root_of_unity"):
            getattr(lib, f).restype = C.c_uint64
        lib.orc_gl_add.argtypes = lib.orc_gl_sub.argtypes = lib.orc_gl_mul.argtypes = lib.orc_gl_pow.argtypes = [C.c_uint64, C.c_uint64]
        lib.orc_gl_inv.argtypes = [C.c_uint64]
        lib.orc_challenger_run.restype = C.c_size_t

    def poseidon(self, states):
        s = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 12).copy()
        self.lib.orc_poseidon(_ptr(s), C.c_size_t(s.shape[0]))
        return s

    def hash_no_pad(self, xs):
        a = np.ascontiguousarray(xs, dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        self.lib.orc_hash_no_pad(_ptr(a), C.c_size_t(a.size), _ptr(out))
        return out

    def hash_or_noop(self, xs):
        a = np.ascontiguousarray(xs, dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        self.lib.orc_hash_or_noop(_ptr(a), C.c_size_t(a.size), _ptr(out))
        return out

    def two_to_one(self, l, r):
        l = np.ascontiguousarray(l, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        self.lib.orc_two_to_one(_ptr(l), _ptr(r), _ptr(out))
        return out

    def hash_rows_colmajor(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint64)
        out = np.empty((a.shape[1], 4), dtype=np.uint64)
        self.lib.orc_hash_rows_colmajor(_ptr(a), C.c_size_t(a.shape[1]), C.c_size_t(a.shape[0]), _ptr(out))
        return out

    def ntt(self, data, kind=0, shift=0):
        a = np.ascontiguousarray(data, dtype=np.uint64).copy()
        if a.ndim == 1:
            a = a.reshape(1, -1)
        self.lib.orc_ntt(_ptr(a), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]), int(kind), C.c_uint64(shift))
        return a

    def naive_dft(self, x):
        a = np.ascontiguousarray(x, dtype=np.uint64)
        out = np.empty_like(a)
        self.lib.orc_naive_dft(_ptr(a), _ptr(out), C.c_size_t(a.size))
        return out

    def commit(self, cols, rate_bits=1, cap_height=4, from_coeffs=False):
        a = np.ascontiguousarray(cols, dtype=np.uint64)
        ncols, n = a.shape
        N = n << rate_bits
        coeffs = np.empty((ncols, n), dtype=np.uint64)
        leaves = np.empty((N, ncols), dtype=np.uint64)
        digests = np.empty((2 * (N - (1 << cap_height)), 4), dtype=np.uint64)
        cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
        rc = self.lib.orc_commit(_ptr(a), C.c_size_t(ncols), C.c_size_t(n), C.c_uint(rate_bits), C.c_uint(cap_height),
                                 int(from_coeffs), _ptr(coeffs), _ptr(leaves), _ptr(digests), _ptr(cap))
        if rc != 0:
            raise RuntimeError("oracle commit failed")
        return coeffs, leaves, digests, cap

    def challenger_run(self, ops):
        """ops: list of ('o', v) | ('c',) | ('k',) -> (challenges, final state)"""
        enc = []
        for op in ops:
            if op[0] == 'o':
                enc += [0, op[1]]
            elif op[0] == 'c':
                enc += [1, 0]
            else:
                enc += [2, 0]
        a = np.array(enc, dtype=np.uint64)
        out = np.zeros(len(ops) + 1, dtype=np.uint64)
        st = np.zeros(12, dtype=np.uint64)
        k = self.lib.orc_challenger_run(_ptr(a), C.c_size_t(len(ops)), _ptr(out), _ptr(st))
        return out[:k].copy(), st


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def load():
    if not os.path.exists(LIB):
        build()
    return Oracle(C.CDLL(LIB))


def rand_field(rng, shape):
    """uniform canonical Goldilocks elements"""
    x = rng.integers(0, P, size=shape, dtype=np.uint64, endpoint=False)
    return x


# ---- per-table STARK prover / verifier --------------------------------------------------------------------------
STANDARD_FAST = (100, 2, 1, 4, 16, 4, 5, 84)     # StarkConfig::standard_fast_config
TEST_CONFIG = (1, 1, 1, 4, 1, 4, 5, 1)           # TEST_STARK_CONFIG, testing_utils.rs:41-52
DEFAULT_LABELS = (0x1234, 0x77, 0x4000, 0x5000)   # halt_final, init, syscall_jumptable, exception_jumptable (arbitrary)


def _cfg(cfg):
    return (C.c_uint32 * 8)(*cfg)


def orc_prove_table(orc, table, cfg, trace, beta_gamma, state, labels=DEFAULT_LABELS, forced_pow=None, debug=False):
    """returns (proof words, new challenger state[, aux values, quotient chunk coeffs, fri values])"""
    lib = orc.lib
    lib.orc_prove_table.restype = C.c_long
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_table_num_aux.restype = C.c_size_t
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    ncols, n = t.shape
    bg = np.ascontiguousarray(beta_gamma, dtype=np.uint64)
    st = np.ascontiguousarray(state, dtype=np.uint64).copy()
    lab = np.array(labels, dtype=np.uint64)
    fp = np.array([forced_pow if forced_pow is not None else 0], dtype=np.uint64)
    na = lib.orc_table_num_aux(C.c_uint32(table), C.c_uint32(cfg[1]))
    aux = np.zeros((na, n), dtype=np.uint64) if debug else None
    quot = np.zeros((2 * cfg[1], n), dtype=np.uint64) if debug else None
    fri = np.zeros((n << cfg[2], 2), dtype=np.uint64) if debug else None
    cap = 1 << 16
    while True:
        out = np.zeros(cap, dtype=np.uint64)
        st2 = st.copy()
        r = lib.orc_prove_table(C.c_uint32(table), _cfg(cfg), _ptr(t), C.c_size_t(ncols), C.c_size_t(n), _ptr(bg), _ptr(st2),
                                _ptr(lab), _ptr(fp) if forced_pow is not None else None, _ptr(out), C.c_size_t(cap),
                                _ptr(aux) if debug and na else None, _ptr(quot) if debug else None, _ptr(fri) if debug else None)
        if r < 0:
            raise RuntimeError("oracle prove_table failed: " + lib.orc_last_error().decode())
        if r <= cap:
            break
        cap = int(r)
    if debug:
        return out[:r].copy(), st2, aux, quot, fri
    return out[:r].copy(), st2


def orc_check_table_rows(orc, table, trace, labels=DEFAULT_LABELS, max_pairs=64):
    """the reference's generator-test pattern: every constraint of the table on every pair of consecutive trace rows ->
    list of violated (row, constraint index); [] for a valid trace"""
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    lab = np.array(labels, dtype=np.uint64)
    out = np.zeros(2 * max_pairs, dtype=np.uint64)
    orc.lib.orc_check_table_rows.restype = C.c_long
    k = orc.lib.orc_check_table_rows(C.c_uint32(table), _ptr(t), C.c_size_t(t.shape[0]), C.c_size_t(t.shape[1]), _ptr(lab), _ptr(out), C.c_size_t(max_pairs))
    if k < 0:
        orc.lib.orc_last_error.restype = C.c_char_p
        raise RuntimeError(orc.lib.orc_last_error().decode())
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(k)]


def orc_table_constraint_degree(orc, table, log_n, seed=1, labels=DEFAULT_LABELS):
    """starky's test_stark_low_degree restated (oracle_api.cpp): degree of the alpha-combined constraint polynomial of a random
    degree-<n trace, interpolated over the subgroup of size 4 n; -2 when it is identically zero"""
    lab = np.array(labels, dtype=np.uint64)
    orc.lib.orc_table_constraint_degree.restype = C.c_long
    d = orc.lib.orc_table_constraint_degree(C.c_uint32(table), C.c_uint(log_n), C.c_uint64(seed), _ptr(lab))
    if d == -1:
        orc.lib.orc_last_error.restype = C.c_char_p
        raise RuntimeError(orc.lib.orc_last_error().decode())
    return int(d)


def orc_table_eval_consistency(orc, table, seed=1, labels=DEFAULT_LABELS):
    """base-field vs extension-field evaluation of the same constraint templates (oracle_api.cpp) -> 0 when consistent"""
    lab = np.array(labels, dtype=np.uint64)
    orc.lib.orc_table_eval_consistency.restype = C.c_long
    r = orc.lib.orc_table_eval_consistency(C.c_uint32(table), C.c_uint64(seed), _ptr(lab))
    if r < 0:
        orc.lib.orc_last_error.restype = C.c_char_p
        raise RuntimeError(orc.lib.orc_last_error().decode())
    return int(r)


def orc_check_ctls(orc, traces, extra_rows=()):
    """starky's debug check_ctls restated (prover.rs:165-184): per cross-table lookup, the number of distinct rows whose multiplicity
    on the looking side (plus `extra_rows` for the Memory lookup) differs from the looked side -> list of 10 counts"""
    arrs = [None if t is None else np.ascontiguousarray(t, dtype=np.uint64) for t in traces]
    ptrs = (u64p * 9)(*[None if a is None else _ptr(a) for a in arrs])
    ns = (C.c_size_t * 9)(*[0 if a is None else a.shape[1] for a in arrs])
    ex = np.array([list(r) for r in extra_rows], dtype=np.uint64).reshape(-1, 13)
    out = (C.c_size_t * 10)()
    orc.lib.orc_check_ctls.restype = C.c_long
    r = orc.lib.orc_check_ctls(ptrs, ns, _ptr(ex) if ex.size else None, C.c_size_t(ex.shape[0]), out)
    if r < 0:
        orc.lib.orc_last_error.restype = C.c_char_p
        raise RuntimeError(orc.lib.orc_last_error().decode())
    return [int(x) for x in out]


def orc_verify_table(orc, table, cfg, proof, beta_gamma, state, labels=DEFAULT_LABELS):
    lib = orc.lib
    lib.orc_last_error.restype = C.c_char_p
    p = np.ascontiguousarray(proof, dtype=np.uint64)
    bg = np.ascontiguousarray(beta_gamma, dtype=np.uint64)
    st = np.ascontiguousarray(state, dtype=np.uint64).copy()
    lab = np.array(labels, dtype=np.uint64)
    r = lib.orc_verify_table(C.c_uint32(table), _cfg(cfg), _ptr(p), C.c_size_t(p.size), _ptr(bg), _ptr(st), _ptr(lab))
    return r == 1, lib.orc_last_error().decode(), st


# ---- whole segment ---------------------------------------------------------------------------------------------------------
def orc_prove_segment(orc, cfg, traces, public_values, labels=DEFAULT_LABELS, forced_pows=None):
    """traces: list of 9 (ncols, n) arrays or None.  Returns (list of 9 proof-word arrays or None, beta_gamma, caps (9,16,4))."""
    lib = orc.lib
    lib.orc_prove_segment.restype = C.c_long
    lib.orc_last_error.restype = C.c_char_p
    arrs = [None if t is None else np.ascontiguousarray(t, dtype=np.uint64) for t in traces]
    ptrs = (u64p * 9)(*[None if a is None else _ptr(a) for a in arrs])
    ns = (C.c_size_t * 9)(*[0 if a is None else a.shape[1] for a in arrs])
    pv = np.ascontiguousarray(public_values, dtype=np.uint64)
    lab = np.array(labels, dtype=np.uint64)
    fp = None if forced_pows is None else np.ascontiguousarray(forced_pows, dtype=np.uint64)
    offs = (C.c_size_t * 10)()
    bg = np.zeros(2 * cfg[1], dtype=np.uint64)
    caps = np.zeros((9, 1 << cfg[3], 4), dtype=np.uint64)
    cap = 1 << 18
    while True:
        out = np.zeros(cap, dtype=np.uint64)
        r = lib.orc_prove_segment(_cfg(cfg), ptrs, ns, _ptr(pv), C.c_size_t(pv.size), _ptr(lab), _ptr(fp) if fp is not None else None,
                                  _ptr(out), C.c_size_t(cap), offs, _ptr(bg), _ptr(caps))
        if r < 0:
            raise RuntimeError("oracle prove_segment failed: " + lib.orc_last_error().decode())
        if r <= cap:
            break
        cap = int(r)
    proofs = [out[offs[t]:offs[t + 1]].copy() if offs[t + 1] > offs[t] else None for t in range(9)]
    return proofs, bg, caps


def orc_verify_segment(orc, cfg, proofs, public_values, labels=DEFAULT_LABELS, extra_looking_sums=None):
    lib = orc.lib
    lib.orc_last_error.restype = C.c_char_p
    offs = (C.c_size_t * 10)()
    tot = 0
    for t in range(9):
        offs[t] = tot
        tot += 0 if proofs[t] is None else len(proofs[t])
    offs[9] = tot
    buf = np.concatenate([np.asarray(p, dtype=np.uint64) for p in proofs if p is not None]) if tot else np.zeros(1, dtype=np.uint64)
    pv = np.ascontiguousarray(public_values, dtype=np.uint64)
    lab = np.array(labels, dtype=np.uint64)
    ex = None if extra_looking_sums is None else np.ascontiguousarray(extra_looking_sums, dtype=np.uint64)
    r = lib.orc_verify_segment(_cfg(cfg), _ptr(buf), offs, _ptr(pv), C.c_size_t(pv.size), _ptr(lab), _ptr(ex) if ex is not None else None)
    return r == 1, lib.orc_last_error().decode()
