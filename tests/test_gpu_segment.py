"""GPU parity at the top of the path: prove_with_traces (all tables of a segment, shared transcript) through the C ABI
vs the oracle's prove_segment, bit for bit, and the oracle's verify_proof (incl. cross-table-lookup sums) on the result."""
import numpy as np
import pytest
from tests import traces
from tests.oracle_lib import orc_prove_segment, orc_verify_segment, STANDARD_FAST, TEST_CONFIG, DEFAULT_LABELS
import zk_evm_b200 as zk

pytestmark = pytest.mark.gpu
PUBLIC_VALUES = np.arange(1000, 1000 + 37, dtype=np.uint64)


def _same(ap, want, bg, caps):
    assert np.array_equal(ap.ctl_challenges, bg)
    assert np.array_equal(ap.trace_caps, caps)
    for t in range(9):
        assert (ap.stark_proofs[t] is None) == (want[t] is None), "table %d presence" % t
        if want[t] is not None:
            assert np.array_equal(ap.stark_proofs[t], want[t]), "table %s proof differs" % zk.TABLE_NAMES[t]


def test_executing_segment_with_32_byte_memory_operations(ctx, oracle):
    """an executing Cpu program with MLOAD_32BYTES / MSTORE_32BYTES rows, seven tables in use, Cpu -> BytePacking carrying non-zero sums:
    device proofs == oracle proofs word for word, and the restated verifier (incl. the cross-table-lookup sums) accepts them"""
    from tests.test_oracle_stark import PACK_PROGRAM, PACK_INPUTS
    tr, labels = traces.cpu_segment(PACK_PROGRAM, inputs=PACK_INPUTS, log_mem=10)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*STANDARD_FAST), zk.KernelLabels(*labels))
    want, bg, caps = orc_prove_segment(oracle, STANDARD_FAST, tr, PUBLIC_VALUES, labels=labels)
    _same(ap, want, bg, caps)
    ok, err = orc_verify_segment(oracle, STANDARD_FAST, ap.stark_proofs, PUBLIC_VALUES, labels=labels)
    assert ok, err


@pytest.mark.parametrize("which", ["syscall_exception_exit_kernel", "set_context_pruning", "user_mode_push"])
def test_executing_segments_of_the_remaining_cpu_families(ctx, oracle, which):
    """the executing segments that give the last Cpu constraint families active rows and the last lookups non-zero sums (syscall /
    exception / EXIT_KERNEL with jump-table reads; SET_CONTEXT with context pruning; a user-mode PUSH read through BytePacking):
    device proofs == oracle proofs word for word, restated verifier accepts"""
    from tests import test_oracle_stark as tos
    if which == "syscall_exception_exit_kernel":
        prog, log_mem = tos.SYS_PROGRAM, 15
        inputs = tos._sys_inputs(len(prog) + 8)
    elif which == "set_context_pruning":
        prog, inputs, log_mem = tos.CTX_PROGRAM, tos.CTX_INPUTS, 10
    else:
        prog, log_mem = tos.USER_PROGRAM, 15
        inputs = tos._user_inputs(len(prog) + 8)
    tr, labels = traces.cpu_segment(prog, inputs=inputs, log_mem=log_mem)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*TEST_CONFIG), zk.KernelLabels(*labels))
    want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PUBLIC_VALUES, labels=labels)
    _same(ap, want, bg, caps)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, ap.stark_proofs, PUBLIC_VALUES, labels=labels)
    assert ok, err


def test_segment_with_the_public_value_writes_through_the_whole_boundary(ctx, oracle):
    """PublicValues -> flatten_public_values (what the transcript observes, all 2217 elements) + the 301 memory writes the kernel makes
    of them (verifier.rs:547-737) in the Memory table: device proofs == oracle proofs word for word, and the restated verify_proof
    accepts them with get_memory_extra_looking_sum (verifier.rs:319-512) computed from the same PublicValues and the DEVICE's (beta, gamma)"""
    from zk_evm_b200.public_values import PublicValues, RegistersData, flatten_public_values, memory_extra_looking_values, memory_extra_looking_sum
    rng = np.random.default_rng(12)
    pv = PublicValues()
    pv.block_metadata.block_beneficiary, pv.block_metadata.block_number, pv.block_metadata.block_random = rng.bytes(20), 19807080, rng.bytes(32)
    pv.block_metadata.block_bloom = [int.from_bytes(rng.bytes(32), "big") for _ in range(8)]
    pv.block_hashes.prev_hashes, pv.block_hashes.cur_hash = [rng.bytes(32) for _ in range(256)], rng.bytes(32)
    pv.trie_roots_before.state_root, pv.trie_roots_after.state_root = rng.bytes(32), rng.bytes(32)
    pv.registers_after = RegistersData(program_counter=0x1234, is_kernel=1, stack_len=2, stack_top=1 << 200, gas_used=77)
    kernel_hash, kernel_len = rng.bytes(32), 60123
    rows = memory_extra_looking_values(pv, kernel_hash, kernel_len)
    flat = flatten_public_values(pv)
    tr, labels = traces.cpu_segment("PPMXJ", log_mem=10, extra_memory_rows=rows)
    cfg = STANDARD_FAST
    ap = zk.prove_with_traces(ctx, tr, flat, zk.StarkConfig(*cfg), zk.KernelLabels(*labels))
    want, bg, caps = orc_prove_segment(oracle, cfg, tr, flat, labels=labels)
    _same(ap, want, bg, caps)
    nc = cfg[1]
    sums = np.zeros(10 * nc, dtype=np.uint64)
    for c in range(nc):
        sums[6 * nc + c] = memory_extra_looking_sum(pv, int(ap.ctl_challenges[2 * c]), int(ap.ctl_challenges[2 * c + 1]), kernel_hash, kernel_len)
    ok, err = orc_verify_segment(oracle, cfg, ap.stark_proofs, flat, labels=labels, extra_looking_sums=sums)
    assert ok, err
    ok, err = orc_verify_segment(oracle, cfg, ap.stark_proofs, flat, labels=labels)
    assert not ok and "lookup 6" in err


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_valid_segment_matches_oracle_and_verifies(ctx, oracle, cfg):
    tr = traces.valid_segment(seed=11)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*cfg), zk.KernelLabels(*DEFAULT_LABELS))
    want, bg, caps = orc_prove_segment(oracle, cfg, tr, PUBLIC_VALUES)
    _same(ap, want, bg, caps)
    ok, err = orc_verify_segment(oracle, cfg, ap.stark_proofs, PUBLIC_VALUES)
    assert ok, err
    assert np.array_equal(ap.mem_before_cap, caps[7]) and np.array_equal(ap.mem_after_cap, caps[8])


def test_random_nine_table_segment_matches_oracle(ctx, oracle):
    log_ns = [9, 8, 10, 6, 7, 8, 11, 9, 7]
    tr = traces.random_segment(log_ns, seed=5)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*STANDARD_FAST), zk.KernelLabels(*DEFAULT_LABELS))
    want, bg, caps = orc_prove_segment(oracle, STANDARD_FAST, tr, PUBLIC_VALUES)
    _same(ap, want, bg, caps)
    # forcing the witnesses reproduces the same proofs; a wrong one is refused
    pows = np.array([int(p[-1]) if p is not None else 0 for p in want], dtype=np.uint64)
    ap2 = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*STANDARD_FAST), zk.KernelLabels(*DEFAULT_LABELS),
                               forced_pow_witnesses=pows)
    _same(ap2, want, bg, caps)


def test_sharded_path_on_one_gpu_and_device_resident_traces(ctx, oracle):
    import torch
    tr = traces.valid_segment(seed=12, k=23)
    in_use = [t is not None for t in tr]
    cfg = zk.StarkConfig(*TEST_CONFIG)
    backend = zk.ZkGpuBackend(ctx, cfg, zk.KernelLabels(*DEFAULT_LABELS))
    ap = zk.prove_with_traces_sharded(backend, zk.LocalComm(), tr, in_use, PUBLIC_VALUES)
    want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PUBLIC_VALUES)
    _same(ap, want, bg, caps)
    # traces already in HBM (torch tensors): same proofs through both entry points
    dev = [None if t is None else torch.from_numpy(t.view(np.int64)).cuda() for t in tr]
    ptrs = [None if d is None else (d.data_ptr(), d.shape[1]) for d in dev]
    torch.cuda.synchronize()
    ap2 = zk.prove_with_traces(ctx, None, PUBLIC_VALUES, cfg, zk.KernelLabels(*DEFAULT_LABELS), device_ptrs=ptrs)
    _same(ap2, want, bg, caps)
    ap3 = zk.prove_with_traces_sharded(backend, zk.LocalComm(), ptrs, in_use, PUBLIC_VALUES)
    _same(ap3, want, bg, caps)


def test_seams_s1_s2_s3_compose_to_prove_with_traces(ctx, oracle):
    """the reference's own decomposition (prover.rs:72-194): PolynomialBatch::from_values per table (S1) -> the transcript replay ->
    get_ctl_data per table (S2) -> prove_with_commitments = prove_single_table per table with one challenger (S3, :211-293) gives the
    proofs, challenges and memory caps of the one-call path, word for word"""
    tr, labels = traces.cpu_segment("PP|PP^PPaXXJ")
    cfg = zk.StarkConfig(*STANDARD_FAST)
    lab = zk.KernelLabels(*labels)
    ap = zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, cfg, lab)
    in_use = [t is not None for t in tr]
    commits = [None if t is None else zk.PolynomialBatch.from_values(ctx, t, rate_bits=cfg.rate_bits, cap_height=cfg.cap_height, keep_values=True)
               for t in tr]
    capw = 4 << cfg.cap_height
    caps = np.zeros((9, capw), dtype=np.uint64)
    for t, b in enumerate(commits):
        if b is not None:
            caps[t] = np.array(b.cap, dtype=np.uint64).reshape(-1)
    bg, st = zk.segment_challenges(caps, in_use, cfg.cap_height, PUBLIC_VALUES, cfg.num_challenges)
    assert np.array_equal(bg, ap.ctl_challenges) and np.array_equal(caps.reshape(ap.trace_caps.shape), ap.trace_caps)
    ctls = [None if b is None else zk.get_ctl_data(ctx, t, b, bg, cfg.num_challenges) for t, b in enumerate(commits)]
    proofs, st_end, mem_before, mem_after = zk.prove_with_commitments(ctx, cfg, commits, in_use, ctls, st, lab)
    for t in range(9):
        assert (proofs[t] is None) == (ap.stark_proofs[t] is None)
        if proofs[t] is not None:
            assert np.array_equal(np.array(proofs[t].words, dtype=np.uint64), ap.stark_proofs[t]), zk.TABLE_NAMES[t]
    assert np.array_equal(mem_before, np.asarray(ap.mem_before_cap).reshape(-1)) and np.array_equal(mem_after, np.asarray(ap.mem_after_cap).reshape(-1))
    ok, err = orc_verify_segment(oracle, STANDARD_FAST, [None if p is None else np.array(p.words, dtype=np.uint64) for p in proofs], PUBLIC_VALUES, labels=labels)
    assert ok, err
    with pytest.raises(ValueError):
        zk.prove_with_commitments(ctx, cfg, commits, in_use, [None] * 9, st, lab)
    for b in commits:
        if b is not None:
            b.free()


def test_missing_mandatory_table_is_an_error(ctx):
    tr = traces.valid_segment(seed=1)
    tr[traces.T_CPU] = None
    with pytest.raises(zk.ZkGpuError):
        zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, zk.StarkConfig(*TEST_CONFIG), zk.KernelLabels(*DEFAULT_LABELS))


def test_upload_ahead_of_prove_gives_the_same_proofs(ctx, oracle):
    """zkgpu_segment_upload for segment s+1 queued before zkgpu_prove_segment_uploaded for segment s (the prefetch pattern of a
    stream of segments): same proofs as the one-call form, uploads consumed exactly once, an unused upload can be dropped"""
    cfg = zk.StarkConfig(*TEST_CONFIG)
    lab = zk.KernelLabels(*DEFAULT_LABELS)
    segs = [traces.valid_segment(seed=31), traces.valid_segment(seed=32, k=17), traces.valid_segment(seed=33, k=29)]
    want = [zk.prove_with_traces(ctx, tr, PUBLIC_VALUES, cfg, lab) for tr in segs]
    got = []
    nxt = zk.upload_traces(ctx, segs[0], cfg)
    for s in range(len(segs)):
        cur, nxt = nxt, (zk.upload_traces(ctx, segs[s + 1], cfg) if s + 1 < len(segs) else None)
        got.append(zk.prove_with_traces(ctx, None, PUBLIC_VALUES, cfg, lab, upload=cur))
        with pytest.raises(Exception):
            zk.prove_with_traces(ctx, None, PUBLIC_VALUES, cfg, lab, upload=cur)      # consumed
    for a, b in zip(got, want):
        _same(a, b.stark_proofs, b.ctl_challenges, b.trace_caps)
    zk.upload_traces(ctx, segs[1], cfg).free()                                          # never proved
    # another StarkConfig than the one the upload was made for is refused
    up = zk.upload_traces(ctx, segs[0], cfg)
    with pytest.raises(zk.ZkGpuError):
        zk.prove_with_traces(ctx, None, PUBLIC_VALUES, zk.StarkConfig(*(TEST_CONFIG[:3] + (3,) + TEST_CONFIG[4:])), lab, upload=up)
