"""GPU parity of the device-side trace finishing: zkgpu_keccak_generate_trace == the restated
KeccakStark::generate_trace (keccak_stark.rs:70-250; tests/traces.py) and zkgpu_logic_generate_trace == LogicStark::generate_trace
(logic.rs:165-237), bit for bit, and a segment proved from the device-resident Keccak / Logic traces next to host traces of the other tables (ZKGPU_MEM_AUTO) == the oracle's proofs from the host trace."""
import numpy as np
import pytest
from tests import traces
from tests.oracle_lib import orc_prove_segment, orc_prove_table, orc_verify_table, TEST_CONFIG, STANDARD_FAST, DEFAULT_LABELS
import zk_evm_b200 as zk

pytestmark = pytest.mark.gpu
PUBLIC_VALUES = np.arange(1000, 1000 + 37, dtype=np.uint64)


@pytest.mark.parametrize("nperm,min_rows,log_n", [(1, 0, 5), (5, 0, 7), (10, 16, 8), (2, 256, 8), (0, 16, 4), (170, 0, 12), (1365, 0, 15)])
def test_keccak_trace_matches_reference_restatement(ctx, nperm, min_rows, log_n):
    rng = np.random.default_rng(300 + nperm)
    inputs = rng.integers(0, 1 << 64, size=(nperm, 25), dtype=np.uint64)
    ts = rng.integers(1, 1 << 40, size=(nperm,), dtype=np.uint64)
    dt = zk.keccak_generate_trace(ctx, inputs, ts, min_rows)
    assert (dt.ncols, dt.n) == (2431, 1 << log_n)
    got = dt.export()
    dt.free()
    want = traces.keccak_trace(log_n, inputs, ts)[0] if nperm else np.zeros((2431, 1 << log_n), dtype=np.uint64)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("nops,min_rows,log_n", [(1, 0, 0), (5, 0, 3), (200, 0, 8), (3, 64, 6), (0, 4, 2), (5000, 0, 13)])
def test_logic_trace_matches_reference_restatement(ctx, nops, min_rows, log_n):
    ops = traces.logic_ops(nops, 400 + nops)
    dt = zk.logic_generate_trace(ctx, ops, min_rows)
    assert (dt.ncols, dt.n) == (523, 1 << log_n)
    got = dt.export()
    dt.free()
    assert np.array_equal(got, traces.logic_trace_from_ops(log_n, ops))


def test_logic_trace_rejects_unknown_operator(ctx):
    ops = traces.logic_ops(4, 1)
    ops[2, 0] = 3
    with pytest.raises(zk.ZkGpuError):
        zk.logic_generate_trace(ctx, ops)


@pytest.mark.parametrize("log_n", [16, 17])
def test_arithmetic_range_checks_match_reference_restatement(ctx, log_n):
    want = traces.arithmetic_addcy_trace(log_n, seed=log_n, nops=3000)
    tr = want.copy()
    tr[114] = 12345          # whatever the two columns held is ignored
    tr[115] = 678
    dt = zk.arithmetic_generate_range_checks(ctx, zk.upload_trace(ctx, tr))
    got = dt.export()
    dt.free()
    assert np.array_equal(got, want)


def test_arithmetic_range_checks_reject_large_values_and_short_traces(ctx):
    tr = traces.arithmetic_addcy_trace(16, seed=1, nops=10)
    tr[40, 999] = 1 << 16
    dt = zk.upload_trace(ctx, tr)
    with pytest.raises(zk.ZkGpuError):
        zk.arithmetic_generate_range_checks(ctx, dt)
    dt.free()
    dt = zk.upload_trace(ctx, np.zeros((116, 1 << 10), dtype=np.uint64))
    with pytest.raises(zk.ZkGpuError):
        zk.arithmetic_generate_range_checks(ctx, dt)
    dt.free()


@pytest.mark.parametrize("seed,log_n,nops,stale", [(1, 6, 40, ()), (2, 6, 40, (2,)), (3, 10, 900, (1, 2)), (4, 5, 20, (0, 1, 2))])
def test_memory_finish_matches_reference_restatement(ctx, seed, log_n, nops, stale):
    ops = traces.memory_sorted_ops(log_n, seed, nops=nops)
    dt = zk.memory_finish_trace(ctx, ops, stale)
    assert (dt.ncols, dt.n) == (30, 1 << log_n)
    got = dt.export()
    dt.free()
    want = traces.memory_finish_reference(ops, stale)
    assert np.array_equal(got, want), [c for c in range(30) if not np.array_equal(got[c], want[c])]


def test_memory_finish_rejects_unsorted_operations(ctx):
    ops = traces.memory_sorted_ops(6, 5)
    ops[1, 10] += 1000
    with pytest.raises(zk.ZkGpuError):
        zk.memory_finish_trace(ctx, ops)
    with pytest.raises(zk.ZkGpuError):
        zk.memory_finish_trace(ctx, traces.memory_sorted_ops(6, 5), stale_contexts=(64,))


def test_device_finished_memory_trace_proves_and_verifies(ctx, oracle):
    cfg = zk.StarkConfig(*TEST_CONFIG)
    dt = zk.memory_finish_trace(ctx, traces.memory_sorted_ops(6, 8), (2,))
    batch = zk.PolynomialBatch.from_device_values(ctx, dt.device_ptr, dt.ncols, dt.n, cfg.rate_bits, cfg.cap_height, keep_values=True)
    dt.free()
    bg = np.array([11, 22], dtype=np.uint64)
    st0 = np.arange(12, dtype=np.uint64)
    ctl = zk.get_ctl_data(ctx, traces.T_MEMORY, batch, bg, cfg.num_challenges)
    proof, st = zk.prove_single_table(ctx, traces.T_MEMORY, cfg, batch, ctl, st0, zk.KernelLabels(*DEFAULT_LABELS))
    ok, err, st2 = orc_verify_table(oracle, traces.T_MEMORY, TEST_CONFIG, proof.words, bg, st0)
    assert ok, err


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_valid_keccak_sponge_trace_on_gpu_verifies(ctx, oracle, cfg):
    """KeccakSpongeStark on a VALID trace (tests/traces.py keccak_sponge_trace = the restated generate_trace): the GPU proof equals the
    oracle's and the restated verifier accepts it (the sponge cases of test_gpu_stark.py are random traces)"""
    rng = np.random.default_rng(6)
    ops = [(1, 2, 100, 7, b""), (0, 3, 5, 9, rng.bytes(135)), (2, 1, 0, 11, rng.bytes(136)), (0, 0, 50, 13, rng.bytes(300)), (0, 2, 7, 21, b"abc")]
    tr, _ = traces.keccak_sponge_trace(8, ops)
    c = zk.StarkConfig(*cfg)
    bg = np.array([11, 22, 33, 44], dtype=np.uint64)[:2 * cfg[1]]
    st0 = np.arange(12, dtype=np.uint64)
    batch = zk.PolynomialBatch.from_values(ctx, tr, c.rate_bits, c.cap_height, keep_values=True)
    ctl = zk.get_ctl_data(ctx, traces.T_KECCAK_SPONGE, batch, bg, c.num_challenges)
    proof, st = zk.prove_single_table(ctx, traces.T_KECCAK_SPONGE, c, batch, ctl, st0, zk.KernelLabels(*DEFAULT_LABELS))
    want, st_want = orc_prove_table(oracle, traces.T_KECCAK_SPONGE, cfg, tr, bg, st0)
    assert np.array_equal(proof.words, want) and np.array_equal(st, st_want)
    ok, err, _ = orc_verify_table(oracle, traces.T_KECCAK_SPONGE, cfg, proof.words, bg, st0)
    assert ok, err


@pytest.mark.parametrize("table,cfg", [(traces.T_CPU, TEST_CONFIG), (traces.T_CPU, STANDARD_FAST), (traces.T_ARITHMETIC, TEST_CONFIG),
                                       (traces.T_ARITHMETIC + 100, TEST_CONFIG)])
def test_valid_traces_with_active_rows_on_gpu(ctx, oracle, table, cfg):
    """CpuStark with active instruction rows, ArithmeticStark with MUL / SHL / BYTE rows and with the two-row modular / division
    operations: GPU proof == oracle proof, the restated verifier accepts"""
    if table == traces.T_CPU:
        program, inputs = traces.cpu_demo_program()          # 72 executed rows: every instruction family cpu_program_trace knows
        tr = traces.cpu_program_trace(7, program, halt_final=DEFAULT_LABELS[0], inputs=inputs)
    elif table == traces.T_ARITHMETIC:
        tr = traces.arithmetic_mul_trace(16, 9, nops=60)
    else:                                                     # the two-row operations: ADDMOD ... DIV, MOD, SHR
        table, tr = traces.T_ARITHMETIC, traces.arithmetic_modular_trace(16, 9, nops=30)
    c = zk.StarkConfig(*cfg)
    bg = np.array([11, 22, 33, 44], dtype=np.uint64)[:2 * cfg[1]]
    st0 = np.arange(12, dtype=np.uint64)
    batch = zk.PolynomialBatch.from_values(ctx, tr, c.rate_bits, c.cap_height, keep_values=True)
    ctl = zk.get_ctl_data(ctx, table, batch, bg, c.num_challenges)
    proof, st = zk.prove_single_table(ctx, table, c, batch, ctl, st0, zk.KernelLabels(*DEFAULT_LABELS))
    want, st_want = orc_prove_table(oracle, table, cfg, tr, bg, st0)
    assert np.array_equal(proof.words, want) and np.array_equal(st, st_want)
    ok, err, _ = orc_verify_table(oracle, table, cfg, proof.words, bg, st0)
    assert ok, err


def test_valid_eight_table_segment_on_gpu_matches_oracle_and_verifies(ctx, oracle):
    """PROVER_INPUT + KECCAK_GENERAL program: Arithmetic, Cpu, Keccak, KeccakSponge, Logic, Memory, MemBefore, MemAfter in use, all ten
    lookups balanced (tests/test_oracle_stark.py::test_cpu_segment_with_keccak_general_and_prover_input_verifies)"""
    from tests.oracle_lib import orc_verify_segment
    rng = np.random.default_rng(5)
    addr = lambda c, sg, v: v | (sg << 32) | (c << 64)
    data = {(1, 0, 10): rng.bytes(150), (1, 0, 300): b"abc"}
    tr, labels = traces.cpu_segment("IIKIIKXXJ", inputs=[150, addr(1, 0, 10), 3, addr(1, 0, 300)], keccak_inputs=data, log_mem=10)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*TEST_CONFIG), zk.KernelLabels(*labels))
    want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PUBLIC_VALUES, labels=labels)
    assert np.array_equal(ap.ctl_challenges, bg) and np.array_equal(ap.trace_caps, caps)
    for t in range(9):
        assert (ap.stark_proofs[t] is None) == (want[t] is None)
        assert want[t] is None or np.array_equal(ap.stark_proofs[t], want[t]), "table %s proof differs" % zk.TABLE_NAMES[t]
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, ap.stark_proofs, PUBLIC_VALUES, labels=labels)
    assert ok, err


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_valid_cpu_segment_on_gpu_matches_oracle_and_verifies(ctx, oracle, cfg):
    """the valid multi-table segment around an executing Cpu program (tests/traces.py cpu_segment: Cpu -> Memory / Arithmetic / Logic
    lookups, MemBefore / MemAfter): GPU proofs == oracle proofs, verify_proof incl. the cross-table-lookup sums accepts"""
    from tests.oracle_lib import orc_verify_segment
    tr, labels = traces.cpu_segment(traces.CPU_SEGMENT_PROGRAM)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*cfg), zk.KernelLabels(*labels))
    want, bg, caps = orc_prove_segment(oracle, cfg, tr, PUBLIC_VALUES, labels=labels)
    assert np.array_equal(ap.ctl_challenges, bg) and np.array_equal(ap.trace_caps, caps)
    for t in range(9):
        assert (ap.stark_proofs[t] is None) == (want[t] is None)
        assert want[t] is None or np.array_equal(ap.stark_proofs[t], want[t]), "table %s proof differs" % zk.TABLE_NAMES[t]
    ok, err = orc_verify_segment(oracle, cfg, ap.stark_proofs, PUBLIC_VALUES, labels=labels)
    assert ok, err


def test_segment_from_device_finished_keccak_and_logic_traces(ctx, oracle):
    rng = np.random.default_rng(9)
    inputs = rng.integers(0, 1 << 64, size=(5, 25), dtype=np.uint64)
    ts = np.arange(1, 6, dtype=np.uint64) * np.uint64(7)
    tr = traces.valid_segment(seed=13, k=17)
    host_keccak, _ = traces.keccak_trace(7, inputs, ts)
    cfg = zk.StarkConfig(*TEST_CONFIG)
    ops = traces.logic_ops(50, 2)
    dt = zk.keccak_generate_trace(ctx, inputs, ts)
    dl = zk.logic_generate_trace(ctx, ops)
    da = zk.arithmetic_generate_range_checks(ctx, zk.upload_trace(ctx, tr[traces.T_ARITHMETIC]))
    tr_dev = list(tr)
    tr_dev[traces.T_ARITHMETIC] = da
    tr_dev[traces.T_KECCAK] = dt
    tr_dev[traces.T_LOGIC] = dl
    ap = zk.prove_with_traces(ctx, tr_dev, PUBLIC_VALUES, cfg, zk.KernelLabels(*DEFAULT_LABELS))
    dt.free()
    dl.free()
    da.free()
    tr_host = list(tr)
    tr_host[traces.T_KECCAK] = host_keccak
    tr_host[traces.T_LOGIC] = traces.logic_trace_from_ops(6, ops)
    want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr_host, PUBLIC_VALUES)
    assert np.array_equal(ap.ctl_challenges, bg) and np.array_equal(ap.trace_caps, caps)
    for t in range(9):
        assert (ap.stark_proofs[t] is None) == (want[t] is None)
        assert want[t] is None or np.array_equal(ap.stark_proofs[t], want[t]), "table %s proof differs" % zk.TABLE_NAMES[t]


def test_device_finished_keccak_trace_proves_and_verifies(ctx, oracle):
    """commit straight from the device trace, prove the table, the oracle's verifier accepts (the constraints hold on generated rows)"""
    rng = np.random.default_rng(10)
    inputs = rng.integers(0, 1 << 64, size=(2, 25), dtype=np.uint64)
    ts = np.array([4, 8], dtype=np.uint64)
    cfg = zk.StarkConfig(*TEST_CONFIG)
    dt = zk.keccak_generate_trace(ctx, inputs, ts)
    batch = zk.PolynomialBatch.from_device_values(ctx, dt.device_ptr, dt.ncols, dt.n, cfg.rate_bits, cfg.cap_height, keep_values=True)
    dt.free()
    bg = np.array([11, 22], dtype=np.uint64)
    st0 = np.arange(12, dtype=np.uint64)
    ctl = zk.get_ctl_data(ctx, traces.T_KECCAK, batch, bg, cfg.num_challenges)
    proof, st = zk.prove_single_table(ctx, traces.T_KECCAK, cfg, batch, ctl, st0, zk.KernelLabels(*DEFAULT_LABELS))
    ok, err, st2 = orc_verify_table(oracle, traces.T_KECCAK, TEST_CONFIG, proof.words, bg, st0)
    assert ok, err
    assert np.array_equal(st, st2)
