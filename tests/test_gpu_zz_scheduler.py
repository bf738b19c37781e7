"""GPU: the segment scheduler's product path (zk_evm_b200/scheduler.py: one Context per worker thread, prove_with_traces through the C ABI)
on a stream of small segments, two in flight, against the oracle; the abort signal reaches the running proofs.
(File name: sorts after the GPU files whose code already ran on a B200 in round 1 — this one has not yet.)"""
import threading

import numpy as np
import pytest
from tests import traces
from tests.oracle_lib import orc_prove_segment, TEST_CONFIG, DEFAULT_LABELS
import zk_evm_b200 as zk
from zk_evm_b200.scheduler import SegmentProver, SegmentAborted

pytestmark = pytest.mark.gpu
PV = np.arange(1000, 1037, dtype=np.uint64)


def test_stream_of_segments_two_in_flight_matches_oracle(oracle):
    segs = [(traces.random_segment([7, 6, 8, 6, 6, 6, 9, 7, 6 + i % 2], seed=60 + i), PV + np.uint64(i)) for i in range(6)]
    prover = SegmentProver(device=0, streams=2, config=zk.StarkConfig(*TEST_CONFIG), labels=DEFAULT_LABELS)
    got = prover.prove_all(iter(segs))
    assert len(got) == 6
    for (tr, pv), ap in zip(segs, got):
        want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, pv)
        assert np.array_equal(ap.ctl_challenges, bg) and np.array_equal(ap.trace_caps, caps)
        for t in range(9):
            assert np.array_equal(ap.stark_proofs[t], want[t]), "table %s" % zk.TABLE_NAMES[t]


def test_abort_stops_the_stream():
    prover = SegmentProver(device=0, streams=2, config=zk.StarkConfig(*TEST_CONFIG), labels=DEFAULT_LABELS)
    seg = (traces.random_segment([10, 9, 12, 8, 8, 9, 13, 10, 9], seed=70), PV)
    threading.Timer(0.2, prover.abort).start()
    with pytest.raises(SegmentAborted):
        prover.prove_all(seg for _ in range(400))


def test_pinned_host_memory_through_the_abi(ctx, oracle):
    """zkgpu_host_alloc / zkgpu_host_register: a host that does not link the CUDA runtime pins its trace buffers through the library"""
    tr = traces.random_segment([7, 6, 8, 6, 6, 6, 9, 7, 6], seed=80)
    cfg = zk.StarkConfig(*TEST_CONFIG)
    want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PV)
    pinned = []
    for t in tr:
        p = zk.PinnedArray(t.shape)
        p.array[...] = t
        pinned.append(p)
    ap = zk.prove_with_traces(ctx, [p.array for p in pinned], PV, cfg, zk.KernelLabels(*DEFAULT_LABELS))
    for t in range(9):
        assert np.array_equal(ap.stark_proofs[t], want[t])
    for p in pinned:
        p.free()
    big = np.ascontiguousarray(tr[6])
    zk.host_register(big)
    try:
        tr2 = list(tr)
        tr2[6] = big
        ap = zk.prove_with_traces(ctx, tr2, PV, cfg, zk.KernelLabels(*DEFAULT_LABELS))
        assert np.array_equal(ap.stark_proofs[6], want[6])
    finally:
        zk.host_unregister(big)


def test_stage_spans(ctx):
    """zkgpu_ctx_set_timing / zkgpu_ctx_timing_report: the TimingTree counterpart"""
    tr = traces.random_segment([7, 6, 8, 6, 6, 6, 9, 7, 6], seed=81)
    ctx.set_timing(True)
    try:
        zk.prove_with_traces(ctx, tr, PV, zk.StarkConfig(*TEST_CONFIG), zk.KernelLabels(*DEFAULT_LABELS))
        spans = ctx.timing_report()
    finally:
        ctx.set_timing(False)
    names = [n for n, _ in spans]
    assert names.count("trace upload") == 9 and names.count("ctl data") == 9 and names.count("quotient eval") == 9
    assert names.count("fri commit phase") == 9 and names.count("pow") == 9
    assert all(ms >= 0 for _, ms in spans) and sum(ms for _, ms in spans) > 0
    assert ctx.timing_report() == []
