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


@pytest.mark.parametrize("cfgname", ["standard_fast", "test"])
def test_config3_b3_b6_segment_matches_oracle_word_for_word(ctx, oracle, cfgname):
    """BASELINE config #3 at ITS heights (CI ranges of witness_b3_b6, scripts/prove_stdio.rs:102-114: Arithmetic 2^16, BytePacking 2^10,
    Cpu 2^16, Keccak 2^12, KeccakSponge 2^8, Logic 2^10, Memory 2^18, MemBefore 2^16, MemAfter 2^7), all nine tables, under
    standard_fast_config and under TEST_STARK_CONFIG (what CI uses, ci.yml:196): device proofs == oracle proofs, bit for bit"""
    import bench
    from tests.oracle_lib import orc_prove_segment, TEST_CONFIG
    cfg = STANDARD_FAST if cfgname == "standard_fast" else TEST_CONFIG
    tr = traces.random_segment(list(bench.SEGMENT_CONFIGS["b3_b6"]), seed=3)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*cfg), zk.KernelLabels(*DEFAULT_LABELS))
    want, bg, caps = orc_prove_segment(oracle, cfg, tr, PUBLIC_VALUES)
    assert np.array_equal(ap.ctl_challenges, bg) and np.array_equal(ap.trace_caps, caps)
    for t in range(9):
        assert np.array_equal(ap.stark_proofs[t], want[t]), "table %d proof differs from the oracle" % t


@pytest.mark.parametrize("table,lg", [(traces.T_KECCAK_SPONGE, 13), (traces.T_LOGIC, 16), (traces.T_KECCAK, 17)])
def test_zz_bench_sized_wide_tables_match_oracle_word_for_word(ctx, oracle, table, lg):
    """the three wide tables at the heights of the bench segment (config #4: KeccakSponge 2^13 x 438, Logic 2^16 x 523, Keccak 2^17 x 2431 —
    45 % of a segment's bytes): commitment cap, auxiliary values, quotient chunks and the whole proof == the oracle's, word for word.
    (The oracle needs about a minute for the Keccak table on 16 host threads; these run last.)"""
    from tests.oracle_lib import orc_prove_table
    cfg = STANDARD_FAST
    tr = traces.random_trace(table, lg, 900 + table)
    tb = zk.PolynomialBatch.from_values(ctx, tr, rate_bits=cfg[2], cap_height=cfg[3], keep_values=True)
    ctl = zk.get_ctl_data(ctx, table, tb, BG2, cfg[1])
    sp, st = zk.prove_single_table(ctx, table, zk.StarkConfig(*cfg), tb, ctl, STATE0, labels=zk.KernelLabels(*DEFAULT_LABELS))
    got = np.array(sp.words, dtype=np.uint64)
    sp.free(); ctl.free(); tb.free()
    want, st_want = orc_prove_table(oracle, table, cfg, tr, BG2, STATE0)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(st, st_want)


def test_zz_keccak_at_the_top_of_its_default_range_proves_and_verifies(oracle):
    """The widest table at the tallest height the reference's default ranges allow (`.env` KECCAK_CIRCUIT_SIZE 4..20: 2^19 rows x 2431
    columns = 10 GB of trace, a 20 GB LDE, 2.5e9 elements — past 2^31, where a 32-bit index anywhere in the path would wrap): a VALID trace
    of 21845 permutations finished on the device, committed, proved; the restated verifier accepts the proof.  Own context, closed
    afterwards, so that the 50 GB of pool do not stay with the suite's shared one."""
    cfg = STANDARD_FAST
    rng = np.random.default_rng(19)
    nperm = (1 << 19) // 24
    inputs = rng.integers(0, 1 << 63, size=(nperm, 25), dtype=np.uint64)
    ts = np.arange(1, nperm + 1, dtype=np.uint64)
    ctx = zk.Context(0)
    try:
        dt = zk.keccak_generate_trace(ctx, inputs, ts)
        assert (dt.ncols, dt.n) == (2431, 1 << 19)
        batch = zk.PolynomialBatch.from_device_values(ctx, dt.device_ptr, dt.ncols, dt.n, cfg[2], cfg[3], keep_values=True)
        dt.free()
        ctl = zk.get_ctl_data(ctx, traces.T_KECCAK, batch, BG2, cfg[1])
        sp, st = zk.prove_single_table(ctx, traces.T_KECCAK, zk.StarkConfig(*cfg), batch, ctl, STATE0, zk.KernelLabels(*DEFAULT_LABELS))
        proof = np.array(sp.words, dtype=np.uint64)
        assert ctx.stats()["bytes_peak"] > 40 << 30
    finally:
        ctx.close()
    ok, err, st2 = orc_verify_table(oracle, traces.T_KECCAK, cfg, proof, BG2, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)
