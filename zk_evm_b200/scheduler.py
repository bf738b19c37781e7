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


def _unpack(segment, default_labels):
    if hasattr(segment, "traces"):          # trace_file.SegmentTraces
        return segment.traces, segment.public_values, getattr(segment, "labels", None) or default_labels
    if len(segment) == 3:
        return segment
    traces, public_values = segment
    return traces, public_values, default_labels


class SegmentProver:
    def __init__(self, device=0, streams=4, config=None, labels=None, make_worker=None, prove=None):
        """make_worker(device) -> per-thread state (default: a zk_evm_b200.Context); prove(state, traces, public_values, labels,
        abort_flag) -> proof (default: prove_with_traces through the C ABI, CUDA only).  The two hooks exist so that the scheduling
        logic can be exercised without a GPU (tests use the oracle as the stand-in); the product path has no CPU fallback."""
        if streams < 1:
            raise ValueError("at least one segment in flight")
        self.device, self.streams, self.config, self.labels = device, streams, config, labels
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
                    proof = self._prove(state, traces, pv, labels, self._abort)
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
