"""Segment scheduler: a stream of segments proved on one GPU with several segments in flight.

The reference dispatches every segment of a batch as an independent proving job and collects the proofs by segment index
(zero/src/prover.rs:205-236: `Directive::map(IndexedStream::from([segment_data]), &seg_prove_ops)` per segment, results sorted by
index; the job itself is `SegmentProof::execute`, zero/src/ops.rs:24-66, i.e. `prove_all_segments` -> `prove`).  Here the workers are
host threads of one process, each with its own `Context` (own CUDA stream, copy stream and memory pool): four segments in flight is
the measured optimum on a B200 (2: 3.94, 3: 4.13, 4: 4.17, 5: 4.18 proofs/s) — the latency-bound phases of one segment fall under the throughput-bound
phases of another, and the library orders the trace uploads of the contexts of one device one chain at a time.

    prover = SegmentProver(device=0, streams=4, config=StarkConfig.standard_fast())
    proofs = prover.prove_all(segments)            # segments: iterable of SegmentTraces / (traces, public_values[, labels])

The segment source is consumed lazily, at most `streams` segments ahead of the proofs (the bounded channel between the reference's
segment generation task and its proving task, zero/src/prover.rs:141-149), so a generator that builds traces on the fly never holds
more than that many segments in memory.  `abort()` raises the abort signal every running proof polls (prover.rs:346-354); a failing
segment aborts the others and its exception is re-raised by `prove_all`.
"""
import ctypes as C
import queue
import threading


class SegmentAborted(RuntimeError):
    pass


# trace width c and auxiliary width a (two challenges) of the nine tables, SURVEY.md 8(a) / zkgpu_table_info
_TABLE_WIDTHS = ((116, 100), (71, 70), (85, 24), (2431, 4), (438, 290), (523, 2), (30, 16), (12, 4), (12, 2))


def estimate_segment_bytes(traces):
    """Device memory one segment proof holds at its peak: every table's trace commitment is made before the first
    table is finished — values + coefficients + the blow-up-2 LDE = 32 n c bytes and ~128 n bytes of digests per table — plus the
    auxiliary and quotient commitments, openings and FRI buffers of the table being finished (the widest one bounds it), 15 % headroom.
    Measured pool peaks on a B200 (tools/memory_peak.py, profiles/r2y_memory_peak.jsonl): 1.32 GB for the b3_b6 heights (estimate 1.65),
    16.8 GB for the b19807080 heights (estimate 21.8) — the estimate is 25-30 % above the peak, on the safe side.
    `traces`: nine (ncols, n) arrays / objects with a `shape`, or None for a table left out."""
    total, transient = 0, 0
    for t, tr in enumerate(traces):
        if tr is None:
            continue
        shape = getattr(tr, "shape", None)
        c, n = (int(shape[0]), int(shape[1])) if shape is not None else (_TABLE_WIDTHS[t][0], int(tr))
        total += 32 * n * c + 128 * n
        transient = max(transient, 32 * n * (_TABLE_WIDTHS[t][1] + 8) + 512 * n)
    return int(1.15 * (total + transient))


class _MemoryBudget:
    """admission by bytes: a segment starts when its estimate fits next to the ones in flight — or when nothing is in flight (a segment
    larger than the whole budget is not refused here; the library reports ZKGPU_ERR_NOMEM if it really does not fit)"""

    def __init__(self, limit):
        self.limit, self.used, self.running = limit, 0, 0
        self.cv = threading.Condition()

    def acquire(self, nbytes, abort):
        with self.cv:
            while self.running and self.used + nbytes > self.limit and not abort.value:
                self.cv.wait(timeout=0.05)
            self.used += nbytes
            self.running += 1

    def release(self, nbytes):
        with self.cv:
            self.used -= nbytes
            self.running -= 1
            self.cv.notify_all()


def _unpack(segment, default_labels):
    if hasattr(segment, "traces"):          # trace_file.SegmentTraces
        return segment.traces, segment.public_values, getattr(segment, "labels", None) or default_labels
    if len(segment) == 3:
        return segment
    traces, public_values = segment
    return traces, public_values, default_labels


class SegmentProver:
    def __init__(self, device=0, streams=4, config=None, labels=None, make_worker=None, prove=None, memory_budget=150 << 30):
        """make_worker(device) -> per-thread state (default: a zk_evm_b200.Context); prove(state, traces, public_values, labels,
        abort_flag) -> proof (default: prove_with_traces through the C ABI, CUDA only).  The two hooks exist so that the scheduling
        logic can be exercised without a GPU (tests use the oracle as the stand-in); the product path has no CPU fallback.
        memory_budget: bytes of device memory the segments in flight may hold together (default 150 GB of a B200's 180 GB; None: no
        limit).  Segments at the bench heights are estimated at 22 GB each (measured peak 16.8), so four run side by side; at the top of
        the reference's default ranges (Keccak 2^19 x 2431, Logic 2^20 x 523, Memory 2^23) one segment is estimated at 100 GB and the scheduler runs them one at a time
        instead of failing with ZKGPU_ERR_NOMEM."""
        if streams < 1:
            raise ValueError("at least one segment in flight")
        self.device, self.streams, self.config, self.labels = device, streams, config, labels
        self._budget = None if memory_budget is None else _MemoryBudget(int(memory_budget))
        self._make_worker = make_worker or self._default_worker
        self._prove = prove or self._default_prove
        self._abort = C.c_int(0)

    def _default_worker(self, device):
        from .prover import Context
        return Context(device)

    def _default_prove(self, ctx, traces, public_values, labels, abort_flag):
        from ._lib import KernelLabels
        from .segment import prove_with_traces
        lab = labels if isinstance(labels, KernelLabels) or labels is None else KernelLabels(*labels)
        return prove_with_traces(ctx, list(traces), public_values, self.config, lab, abort_flag=abort_flag)

    def abort(self):
        """Option<Arc<AtomicBool>> abort_signal: running proofs stop at their next check, queued segments are dropped"""
        self._abort.value = 1

    def prove_all(self, segments):
        """-> list of proofs in segment order"""
        self._abort.value = 0
        work = queue.Queue(maxsize=self.streams)
        results, errors = {}, []
        lock = threading.Lock()

        def worker():
            state = None
            try:
                state = self._make_worker(self.device)
                while True:
                    item = work.get()
                    if item is None:
                        return
                    idx, seg = item
                    if self._abort.value:
                        continue                      # drain
                    traces, pv, labels = _unpack(seg, self.labels)
                    need = estimate_segment_bytes(traces) if self._budget is not None else 0
                    if self._budget is not None:
                        self._budget.acquire(need, self._abort)
                    try:
                        if self._abort.value:
                            continue
                        proof = self._prove(state, traces, pv, labels, self._abort)
                    finally:
                        if self._budget is not None:
                            self._budget.release(need)
                    with lock:
                        results[idx] = proof
            except BaseException as e:               # noqa: BLE001  (re-raised by prove_all)
                with lock:
                    errors.append(e)
                self._abort.value = 1
                while work.get() is not None:         # keep draining so that the feeder never blocks
                    pass
            finally:
                if state is not None and hasattr(state, "close"):
                    state.close()

        threads = [threading.Thread(target=worker, name="segment-prover-%d" % i) for i in range(self.streams)]
        for th in threads:
            th.start()
        count = 0
        try:
            for idx, seg in enumerate(segments):
                if self._abort.value:
                    break
                work.put((idx, seg))
                count += 1
        except BaseException as e:                    # the segment source failed
            errors.append(e)
            self._abort.value = 1
        finally:
            for _ in threads:
                work.put(None)
            for th in threads:
                th.join()
        if errors:
            first = errors[0]
            if getattr(first, "code", None) == -3:    # ZKGPU_ERR_ABORTED caused by abort()
                raise SegmentAborted("abort signal observed") from first
            raise first
        if self._abort.value:
            raise SegmentAborted("abort signal observed")
        return [results[i] for i in range(count)]
