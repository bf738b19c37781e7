"""Committed golden vectors (tests/golden/golden_r1.json, made by tests/golden/make_golden.py): the oracle still reproduces them
(CPU), and the device reproduces them without the oracle in the loop (GPU)."""
import hashlib
import json
import os
import numpy as np
import pytest
from tests import traces, oracle_lib
from tests.golden.make_golden import TABLE_CASES, table_trace, fp, BG2, STATE0, PUBLIC_VALUES
from tests.oracle_lib import STANDARD_FAST, TEST_CONFIG, DEFAULT_LABELS

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r1.json")))
OV = G["oracle_vectors"]


def test_reference_kats_against_oracle(oracle):
    k = G["reference_kats"]
    assert [int(v) for v in oracle.poseidon(np.zeros(12, dtype=np.uint64))[0][:4]] == k["poseidon_hash_zeros"]["value"]
    assert [int(v) for v in oracle.hash_no_pad(np.zeros(2048, dtype=np.uint64))] == k["empty_consolidated_blockhash"]["value"]
    assert oracle.lib.orc_gl_inv(65536) == k["goldilocks_inverse_65536"]["value"]


def test_oracle_reproduces_golden_commit_and_ntt(oracle):
    cols = oracle_lib.rand_field(np.random.default_rng(1234), (12, 1 << 10))
    co, le, di, cap = oracle.commit(cols, 1, 4)
    e = OV["commit_12x1024_seed1234"]
    assert [int(x) for x in cap.ravel()] == e["cap"] and fp(co) == e["coeffs"] and fp(le) == e["leaves"] and fp(di) == e["digests"]
    x = oracle_lib.rand_field(np.random.default_rng(1), 1 << 16)
    assert fp(oracle.ntt(x, 0)) == OV["ntt_65536_seed1"]["fft"]
    assert fp(oracle.ntt(x, 2, oracle_lib.GENERATOR)) == OV["ntt_65536_seed1"]["coset_fft"]


@pytest.mark.parametrize("case", TABLE_CASES, ids=lambda c: c[0])
def test_oracle_reproduces_golden_proofs(oracle, case):
    name, table, lg, cfgname, kind = case
    cfg = STANDARD_FAST if cfgname == "std" else TEST_CONFIG
    proof, st = oracle_lib.orc_prove_table(oracle, table, cfg, table_trace(table, lg, kind), BG2[:2 * cfg[1]], STATE0)
    e = OV["prove_table_%s_2^%d_%s_%s" % (name, lg, cfgname, kind)]
    assert fp(proof) == e["proof"] and [int(v) for v in st] == e["state_after"]


# ---- device vs the committed vectors ----------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_reference_kats(ctx):
    import zk_evm_b200  # noqa: F401
    k = G["reference_kats"]
    assert [int(v) for v in ctx.poseidon_permute(np.zeros(12, dtype=np.uint64))[0][:4]] == k["poseidon_hash_zeros"]["value"]
    # hash_no_pad([0; 2048]) == hash_or_noop of one 2048-wide row
    out = ctx.poseidon_hash_rows(np.zeros((2048, 1), dtype=np.uint64))
    assert [int(v) for v in out[0]] == k["empty_consolidated_blockhash"]["value"]


@pytest.mark.gpu
def test_gpu_golden_commit_ntt_poseidon(ctx):
    import zk_evm_b200 as zk
    cols = oracle_lib.rand_field(np.random.default_rng(1234), (12, 1 << 10))
    b = zk.PolynomialBatch.from_values(ctx, cols, rate_bits=1, cap_height=4)
    co, le, di = b.export()
    e = OV["commit_12x1024_seed1234"]
    assert [int(x) for x in b.cap.ravel()] == e["cap"] and fp(co) == e["coeffs"] and fp(le) == e["leaves"] and fp(di) == e["digests"]
    x = oracle_lib.rand_field(np.random.default_rng(1), 1 << 16)
    assert fp(ctx.ntt(x.reshape(1, -1))) == OV["ntt_65536_seed1"]["fft"]
    assert fp(ctx.ntt(x.reshape(1, -1), coset_shift=oracle_lib.GENERATOR)) == OV["ntt_65536_seed1"]["coset_fft"]
    st = oracle_lib.rand_field(np.random.default_rng(2), (64, 12))
    assert fp(ctx.poseidon_permute(st)) == OV["poseidon_64_states_seed2"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", TABLE_CASES, ids=lambda c: c[0])
def test_gpu_golden_table_proofs(ctx, case):
    import zk_evm_b200 as zk
    name, table, lg, cfgname, kind = case
    cfg = STANDARD_FAST if cfgname == "std" else TEST_CONFIG
    tr = table_trace(table, lg, kind)
    tb = zk.PolynomialBatch.from_values(ctx, tr, rate_bits=cfg[2], cap_height=cfg[3], keep_values=True)
    ctl = zk.get_ctl_data(ctx, table, tb, BG2[:2 * cfg[1]], cfg[1])
    proof, st = zk.prove_single_table(ctx, table, zk.StarkConfig(*cfg), tb, ctl, STATE0, labels=zk.KernelLabels(*DEFAULT_LABELS))
    e = OV["prove_table_%s_2^%d_%s_%s" % (name, lg, cfgname, kind)]
    assert len(proof.words) == e["proof_words"] and int(proof.words[-1]) == e["pow_witness"]
    assert fp(proof.words) == e["proof"] and [int(v) for v in st] == e["state_after"]


@pytest.mark.gpu
def test_gpu_golden_segment(ctx):
    import zk_evm_b200 as zk
    ap = zk.prove_with_traces(ctx, traces.valid_segment(seed=11), PUBLIC_VALUES, zk.StarkConfig(*TEST_CONFIG), zk.KernelLabels(*DEFAULT_LABELS))
    e = OV["prove_segment_valid_seed11_test"]
    assert [int(v) for v in ap.ctl_challenges] == e["ctl_challenges"] and fp(ap.trace_caps) == e["caps"]
    assert [None if p is None else fp(p) for p in ap.stark_proofs] == e["proofs"]
