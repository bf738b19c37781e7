"""Host logic of the segment scheduler (zk_evm_b200/scheduler.py; zero/src/prover.rs:205-236, zero/src/ops.rs:24-66) on the CPU, with
the oracle as the stand-in for the device: proofs come back in segment order and equal the sequential ones, the segment source is
consumed lazily, a failing segment / the abort signal stops the others."""
import threading
import time

import numpy as np
import pytest

from tests import traces
from tests.oracle_lib import orc_prove_segment, TEST_CONFIG
from zk_evm_b200.scheduler import SegmentProver, SegmentAborted

PV = np.arange(1000, 1037, dtype=np.uint64)


def _oracle_prover(oracle, streams, **kw):
    made = []

    def make_worker(device):
        made.append(threading.current_thread().name)
        return {"device": device}

    def prove(state, tr, pv, labels, abort_flag):
        if abort_flag.value:
            raise SegmentAborted("aborted")
        return orc_prove_segment(oracle, TEST_CONFIG, tr, pv)[0]
    return SegmentProver(device=0, streams=streams, make_worker=make_worker, prove=prove, **kw), made


def test_proofs_come_back_in_segment_order_and_equal_sequential_ones(oracle):
    segs = [(traces.random_segment([6, 5, 6, 4, 5, 5, 7, 6, 5 + i % 2], seed=30 + i), PV + np.uint64(i)) for i in range(5)]
    prover, made = _oracle_prover(oracle, streams=3)
    got = prover.prove_all(iter(segs))
    assert len(made) == 3 and len(got) == 5
    for (tr, pv), proofs in zip(segs, got):
        want = orc_prove_segment(oracle, TEST_CONFIG, tr, pv)[0]
        for t in range(9):
            assert (proofs[t] is None) == (want[t] is None)
            assert want[t] is None or np.array_equal(proofs[t], want[t])


def test_segment_source_is_consumed_lazily():
    produced, gate = [], threading.Event()

    def source():
        for i in range(20):
            produced.append(i)
            yield ([None] * 9, PV)

    def prove(state, tr, pv, labels, abort_flag):
        gate.wait(5)
        return "proof"
    prover = SegmentProver(streams=2, make_worker=lambda d: None, prove=prove)
    th = threading.Thread(target=lambda: prover.prove_all(source()))
    th.start()
    time.sleep(0.5)
    # 2 segments being proved + 2 queued + the one the feeder is blocked on
    assert len(produced) <= 5, produced
    gate.set()
    th.join(10)
    assert len(produced) == 20


def test_failing_segment_stops_the_others_and_is_reraised():
    started = []

    def prove(state, tr, pv, labels, abort_flag):
        started.append(int(pv[0]))
        if int(pv[0]) == 3:
            raise ValueError("segment 3 is broken")
        for _ in range(200):
            if abort_flag.value:
                raise SegmentAborted("aborted")
            time.sleep(0.005)
        return "proof"
    prover = SegmentProver(streams=2, make_worker=lambda d: None, prove=prove)
    with pytest.raises(ValueError, match="segment 3"):
        prover.prove_all(([None] * 9, np.array([i], dtype=np.uint64)) for i in range(50))
    assert len(started) < 50


def test_abort_signal():
    def prove(state, tr, pv, labels, abort_flag):
        for _ in range(400):
            if abort_flag.value:
                raise SegmentAborted("aborted")
            time.sleep(0.005)
        return "proof"
    prover = SegmentProver(streams=2, make_worker=lambda d: None, prove=prove)
    threading.Timer(0.3, prover.abort).start()
    t0 = time.time()
    with pytest.raises(SegmentAborted):
        prover.prove_all(([None] * 9, PV) for _ in range(10))
    assert time.time() - t0 < 5
    # the prover is reusable afterwards
    assert prover.prove_all(([None] * 9, PV) for _ in range(1)) == ["proof"]


def test_trace_file_segments_are_accepted(tmp_path, oracle):
    from zk_evm_b200 import trace_file
    tr = traces.random_segment([5, 5, 6, 4, 5, 5, 6, 6, 5], seed=40)
    path = tmp_path / "s.trace"
    trace_file.save(path, tr, PV, (1, 2, 3, 4))
    seen = []

    def prove(state, tr_, pv, labels, abort_flag):
        seen.append(tuple(labels))
        return orc_prove_segment(oracle, TEST_CONFIG, list(tr_), pv, labels=tuple(labels))[0]
    got = SegmentProver(streams=1, make_worker=lambda d: None, prove=prove).prove_all([trace_file.load(path)])
    assert seen == [(1, 2, 3, 4)] and len(got) == 1


# ---- admission by device memory: segments at the top of the reference's ranges do not fit four at a time in 180 GB ---------------------
class _Shape:
    def __init__(self, c, lg):
        self.shape = (c, 1 << lg)


_WIDTHS = [116, 71, 85, 2431, 438, 523, 30, 12, 12]


def _segment_of(lgs):
    return [_Shape(_WIDTHS[t], lg) for t, lg in enumerate(lgs)]


def test_estimate_of_device_memory_per_segment():
    from zk_evm_b200.scheduler import estimate_segment_bytes
    bench = estimate_segment_bytes(_segment_of([17, 14, 19, 17, 13, 16, 21, 19, 19]))         # BASELINE config #4 heights
    # 3.94 GB of traces -> values + coefficients + LDE = 15.8 GB, + the widest table's auxiliary / quotient / FRI buffers, + headroom
    assert 18e9 < bench < 24e9
    top = estimate_segment_bytes(_segment_of([20, 20, 20, 19, 16, 20, 23, 22, 22]))           # top of the default ranges (.env:2-10)
    assert 95e9 < top < 120e9                                                                  # one at a time on a 180 GB device
    assert estimate_segment_bytes([None] * 9) == 0
    some = _segment_of([17, 14, 19, 17, 13, 16, 21, 19, 19])
    some[3] = None                                                                             # a segment without Keccak
    assert estimate_segment_bytes(some) < bench / 2


def _counting_prover(streams, budget, hold=0.05):
    state = {"now": 0, "max": 0, "order": []}
    lock = threading.Lock()

    def prove(st, tr, pv, labels, abort_flag):
        with lock:
            state["now"] += 1
            state["max"] = max(state["max"], state["now"])
        time.sleep(hold)
        with lock:
            state["now"] -= 1
        return int(pv)
    return SegmentProver(streams=streams, make_worker=lambda d: None, prove=prove, memory_budget=budget), state


def test_memory_budget_bounds_the_segments_in_flight():
    from zk_evm_b200.scheduler import estimate_segment_bytes
    seg = _segment_of([17, 14, 19, 17, 13, 16, 21, 19, 19])
    need = estimate_segment_bytes(seg)
    segs = [(seg, i) for i in range(12)]
    # room for all four streams
    prover, st = _counting_prover(4, 5 * need)
    assert prover.prove_all(iter(segs)) == list(range(12)) and st["max"] == 4
    # room for two at a time only
    prover, st = _counting_prover(4, int(2.5 * need))
    assert prover.prove_all(iter(segs)) == list(range(12)) and st["max"] == 2
    # a segment larger than the whole budget still runs, alone
    prover, st = _counting_prover(4, need // 2)
    assert prover.prove_all(iter(segs)) == list(range(12)) and st["max"] == 1
    # no budget: the stream count alone bounds it
    prover, st = _counting_prover(3, None)
    assert prover.prove_all(iter(segs)) == list(range(12)) and st["max"] == 3


def test_memory_budget_mixed_sizes_and_abort():
    small = _segment_of([16, 8, 12, 7, 8, 5, 17, 16, 7])
    big = _segment_of([20, 20, 20, 19, 16, 20, 23, 22, 22])
    from zk_evm_b200.scheduler import estimate_segment_bytes
    budget = estimate_segment_bytes(big) + 3 * estimate_segment_bytes(small)
    segs = [(big if i % 3 == 0 else small, i) for i in range(9)]
    prover, st = _counting_prover(4, budget, hold=0.03)
    assert prover.prove_all(iter(segs)) == list(range(9)) and 2 <= st["max"] <= 4
    # the abort signal wakes the workers that wait for memory
    prover, st = _counting_prover(4, estimate_segment_bytes(big), hold=0.2)
    th_err = []

    def run():
        try:
            prover.prove_all(iter([(big, i) for i in range(8)]))
        except SegmentAborted as e:
            th_err.append(e)
    th = threading.Thread(target=run)
    th.start()
    time.sleep(0.1)
    prover.abort()
    th.join(5)
    assert not th.is_alive() and th_err and st["max"] == 1
