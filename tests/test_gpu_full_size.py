"""GPU, at the sizes BASELINE.json names (where the CPU oracle prover would take minutes): the proofs the CUDA path produces for VALID
traces are checked by the oracle's restated verifier (verifier.rs:172-313: constraint identity at zeta, FRI, proof of work, and for a
segment the cross-table-lookup sums) — the size-independent property the reference's own tests rely on (verify_proof after prove)."""
import numpy as np
import pytest
from tests import traces
from tests.oracle_lib import orc_verify_table, orc_verify_segment, STANDARD_FAST, DEFAULT_LABELS
import zk_evm_b200 as zk

pytestmark = pytest.mark.gpu

BG2 = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
STATE0 = np.arange(1, 13, dtype=np.uint64)
PUBLIC_VALUES = np.arange(1000, 1000 + 37, dtype=np.uint64)


def test_config2_cpu_table_2e20_verifies(ctx, oracle):
    """BASELINE config #2: single CpuStark table, 2^20 rows (valid all-padding trace), standard_fast_config"""
    cfg = STANDARD_FAST
    tr = traces.cpu_padding_trace(20, halt_final=DEFAULT_LABELS[0])
    tb = zk.PolynomialBatch.from_values(ctx, tr, rate_bits=cfg[2], cap_height=cfg[3], keep_values=True)
    ctl = zk.get_ctl_data(ctx, traces.T_CPU, tb, BG2, cfg[1])
    sp, st = zk.prove_single_table(ctx, traces.T_CPU, zk.StarkConfig(*cfg), tb, ctl, STATE0, labels=zk.KernelLabels(*DEFAULT_LABELS))
    proof = np.array(sp.words, dtype=np.uint64)
    ok, err, st2 = orc_verify_table(oracle, traces.T_CPU, cfg, proof, BG2, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)          # prover and verifier leave the transcript in the same state
    bad = proof.copy()
    bad[len(bad) // 2] ^= np.uint64(1)
    ok, err, _ = orc_verify_table(oracle, traces.T_CPU, cfg, bad, BG2, STATE0)
    assert not ok
    tb.free()


def test_config4_sized_valid_segment_verifies(ctx, oracle):
    """a valid segment at the heights of BASELINE config #4 for the tables a witness-free generator can fill (Cpu 2^19, Memory 2^21,
    MemBefore / MemAfter 2^19, Arithmetic 2^16; 400 000 memory cells carried from MemBefore to MemAfter through Memory)"""
    tr = traces.valid_segment(seed=21, log_cpu=19, log_mem=21, log_memcont=19, k=400000)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*STANDARD_FAST), zk.KernelLabels(*DEFAULT_LABELS))
    ok, err = orc_verify_segment(oracle, STANDARD_FAST, ap.stark_proofs, PUBLIC_VALUES)
    assert ok, err
    # dropping one MemAfter row breaks exactly the cross-table lookup between Memory and MemAfter
    tr[traces.T_MEM_AFTER][0, 399999] = 0                   # filter off: the row is no longer looked
    ap2 = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*STANDARD_FAST), zk.KernelLabels(*DEFAULT_LABELS))
    ok, err = orc_verify_segment(oracle, STANDARD_FAST, ap2.stark_proofs, PUBLIC_VALUES)
    assert not ok and "Cross-table lookup" in err
