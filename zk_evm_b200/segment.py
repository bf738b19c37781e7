"""Segment-level host mirror: `prove_with_traces` / `prove_with_commitments` of evm_arithmetization/src/prover.rs:72-293.

Two ways to prove the tables of one segment:
  * `prove_with_traces`         one device, one call into the C ABI (zkgpu_prove_segment);
  * `prove_with_traces_sharded` the tables of the segment spread over the ranks of a process group (one process per GPU):
        phase 1 (parallel)  every rank commits the traces of its tables;
        exchange 1          all-gather of the 9 trace caps (512 B each) -> every rank replays the transcript
                            (caps, public values) and draws the same CTL challenges (prover.rs:118-144);
        phase 2 (parallel)  CTL / lookup auxiliary columns and their commitment for the local tables (table_job_begin);
        phase 3 (relay)     in Table order the owner of table t finishes its proof from the shared transcript state and
                            broadcasts the 12-word state it ends with (prover.rs:251-259 proves the tables one after the
                            other with the same challenger, so this chain is part of the protocol, not of the implementation).
The compute steps go through a small backend object so the sharding logic can be exercised on CPU (gloo) with the test
oracle as the backend; the product backend is `ZkGpuBackend` (CUDA only, no fallback).
"""
import ctypes as C
import numpy as np
from ._lib import u64p, check, lib
from . import prover as _p

NUM_TABLES = 9
TABLE_NAMES = ("arithmetic", "byte_packing", "cpu", "keccak", "keccak_sponge", "logic", "memory", "mem_before", "mem_after")
OPTIONAL_TABLES = (1, 3, 4, 5, 8)          # OPTIONAL_TABLE_INDICES, all_stark.rs:110-117


def _ptr(a):
    return a.ctypes.data_as(u64p)


class Challenger:
    """plonky2 Challenger<F, PoseidonHash> (host side)."""

    def __init__(self):
        self._h = C.c_void_p()
        check(lib().zkgpu_challenger_new(C.byref(self._h)))

    def observe_elements(self, xs):
        a = np.ascontiguousarray(xs, dtype=np.uint64).ravel()
        check(lib().zkgpu_challenger_observe(self._h, _ptr(a), C.c_size_t(a.size)))

    def get_n_challenges(self, n):
        out = np.empty(n, dtype=np.uint64)
        check(lib().zkgpu_challenger_get_challenges(self._h, _ptr(out), C.c_size_t(n)))
        return out

    def compact(self):
        st = np.empty(12, dtype=np.uint64)
        check(lib().zkgpu_challenger_compact(self._h, _ptr(st)))
        return st

    def set_state(self, st):
        a = np.ascontiguousarray(st, dtype=np.uint64)
        check(lib().zkgpu_challenger_set_state(self._h, _ptr(a)))

    def __del__(self):
        try:
            if self._h:
                lib().zkgpu_challenger_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def segment_challenges(trace_caps, table_in_use, cap_height, public_values, num_challenges):
    """prover.rs:118-144: (beta_gamma, compacted challenger state before the first table)."""
    caps = np.ascontiguousarray(trace_caps, dtype=np.uint64).reshape(NUM_TABLES, -1)
    assert caps.shape[1] == 4 << cap_height
    use = np.ascontiguousarray(table_in_use, dtype=np.uint8)
    pv = np.ascontiguousarray(public_values, dtype=np.uint64).ravel()
    bg = np.empty(2 * num_challenges, dtype=np.uint64)
    st = np.empty(12, dtype=np.uint64)
    check(lib().zkgpu_segment_challenges(_ptr(caps), use.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(cap_height), _ptr(pv),
                                         C.c_size_t(pv.size), C.c_uint32(num_challenges), _ptr(bg), _ptr(st)))
    return bg, st


class AllProof:
    """AllProof / MultiProof (proof.rs:29-54): per-table StarkProofWithMetadata words (None = table not in use), the CTL
    challenges and the trace caps (MemBefore / MemAfter caps feed PublicValues.mem_before / mem_after, prover.rs:261-271)."""

    def __init__(self, stark_proofs, ctl_challenges, trace_caps, table_in_use):
        self.stark_proofs, self.ctl_challenges, self.trace_caps, self.table_in_use = stark_proofs, ctl_challenges, trace_caps, table_in_use

    @property
    def mem_before_cap(self):
        return self.trace_caps[7]

    @property
    def mem_after_cap(self):
        return self.trace_caps[8] if self.table_in_use[8] else np.zeros_like(self.trace_caps[8])


class _Trace(C.Structure):
    _fields_ = [("cols", u64p), ("n", C.c_size_t)]


def _trace_array(traces, device_ptrs, config):
    arr = (_Trace * NUM_TABLES)()
    keep = []
    in_use = [False] * NUM_TABLES
    for t in range(NUM_TABLES):
        if device_ptrs is not None:
            if device_ptrs[t] is not None:
                arr[t].cols = C.cast(C.c_void_p(int(device_ptrs[t][0])), u64p)
                arr[t].n = int(device_ptrs[t][1])
                in_use[t] = True
        elif isinstance(traces[t], _p.DeviceTrace):      # finished on the device (keccak_generate_trace)
            info = _p.table_info(t, config.num_challenges)
            if traces[t].ncols != info["num_columns"]:
                raise ValueError("table %s: expected %d columns" % (TABLE_NAMES[t], info["num_columns"]))
            keep.append(traces[t])
            arr[t].cols = C.cast(C.c_void_p(traces[t].device_ptr), u64p)
            arr[t].n = traces[t].n
            in_use[t] = True
        elif traces[t] is not None:
            a = np.ascontiguousarray(traces[t], dtype=np.uint64)
            info = _p.table_info(t, config.num_challenges)
            if a.ndim != 2 or a.shape[0] != info["num_columns"]:
                raise ValueError("table %s: expected a (%d, n) trace" % (TABLE_NAMES[t], info["num_columns"]))
            keep.append(a)
            arr[t].cols = _ptr(a)
            arr[t].n = a.shape[1]
            in_use[t] = True
    return arr, keep, in_use


class SegmentUpload:
    """The traces of one segment queued for (or already in) device memory: `upload_traces` for segment s+1 BEFORE
    `prove_with_traces(..., upload=...)` for segment s puts the whole H2D chain of s+1 under the proof of s."""

    def __init__(self, ctx, handle, keep, in_use):
        self.ctx, self._h, self._keep, self.in_use = ctx, handle, keep, in_use
        ctx._children.add(self)

    def free(self):
        if self._h:
            lib().zkgpu_upload_free(self._h)
            self._h = C.c_void_p()
            self._keep = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _mem_kind(traces, device_ptrs):
    """ZKGPU_MEM_DEVICE for device_ptrs, ZKGPU_MEM_AUTO when a DeviceTrace sits among host arrays, else ZKGPU_MEM_HOST"""
    if device_ptrs is not None:
        return 1
    return 2 if any(isinstance(t, _p.DeviceTrace) for t in traces) else 0


def upload_traces(ctx, traces, config, device_ptrs=None):
    arr, keep, in_use = _trace_array(traces, device_ptrs, config)
    h = C.c_void_p()
    check(lib().zkgpu_segment_upload(ctx._h, arr, _mem_kind(traces, device_ptrs), C.byref(config), C.byref(h)))
    return SegmentUpload(ctx, h, keep, in_use)


def prove_with_traces(ctx, traces, public_values, config, labels=None, forced_pow_witnesses=None, abort_flag=None, device_ptrs=None,
                      upload=None):
    """prove_with_traces on one device.  traces: list of 9 (ncols, n) uint64 arrays (or DeviceTrace objects: traces finished on the
    device, e.g. keccak_generate_trace), None for an optional table not in use.
    device_ptrs: instead of host arrays, list of 9 (device address, n) or None (traces already resident in HBM).
    upload: a SegmentUpload made earlier by upload_traces (consumed by this call) instead of traces / device_ptrs."""
    pv = np.ascontiguousarray(public_values, dtype=np.uint64).ravel()
    fp = None if forced_pow_witnesses is None else np.ascontiguousarray(forced_pow_witnesses, dtype=np.uint64)
    outs = (C.c_void_p * NUM_TABLES)()
    bg = np.zeros(4, dtype=np.uint64)
    caps = np.zeros((NUM_TABLES, 1 << config.cap_height, 4), dtype=np.uint64)
    tail = (_ptr(pv), C.c_size_t(pv.size), C.byref(labels) if labels is not None else None, C.byref(config),
            _ptr(fp) if fp is not None else None, C.byref(abort_flag) if abort_flag is not None else None, outs, _ptr(bg), _ptr(caps))
    if upload is not None:
        in_use = upload.in_use
        try:
            check(lib().zkgpu_prove_segment_uploaded(ctx._h, upload._h, *tail))
        finally:
            upload.free()
    else:
        arr, keep, in_use = _trace_array(traces, device_ptrs, config)
        check(lib().zkgpu_prove_segment(ctx._h, arr, _mem_kind(traces, device_ptrs), *tail))
    proofs = []
    for t in range(NUM_TABLES):
        if outs[t]:
            sp = _p.StarkProof(ctx, C.c_void_p(outs[t]))
            proofs.append(sp.words)
            sp.free()
        else:
            proofs.append(None)
    return AllProof(proofs, bg[:2 * config.num_challenges].copy(), caps, in_use)


# ---- table-sharded proving ------------------------------------------------------------------------------------------------
def prove_with_commitments(ctx, config, trace_commitments, table_in_use, ctl_data_per_table, challenger_state, labels=None,
                           forced_pow_witnesses=None, abort_flag=None):
    """prover.rs:211-293: the tables in Table order, ONE challenger handed from table to table (:251-259), `None` for a table that is not in
    use.  trace_commitments[t]: the PolynomialBatch of table t's trace (made with keep_values — the seam S1 of SURVEY 8b), or None;
    ctl_data_per_table[t]: its CtlData (get_ctl_data, seam S2); challenger_state: the compacted transcript state after the trace caps, the
    public values and the lookup challenges (segment_challenges).  -> (proofs[9], challenger state after the last table, mem_before cap,
    mem_after cap — zeros when MemAfter is not in use, as ProofWithMemCaps)."""
    st = np.ascontiguousarray(challenger_state, dtype=np.uint64).copy()
    proofs = [None] * NUM_TABLES
    for t in range(NUM_TABLES):
        if not table_in_use[t]:
            continue
        if trace_commitments[t] is None or ctl_data_per_table[t] is None:
            raise ValueError("table %s is in use: it needs its trace commitment and its CtlData" % TABLE_NAMES[t])
        fp = None if forced_pow_witnesses is None else int(forced_pow_witnesses[t])
        proofs[t], st = _p.prove_single_table(ctx, t, config, trace_commitments[t], ctl_data_per_table[t], st, labels, fp, abort_flag)
    capw = 4 << config.cap_height
    mem_before = np.array(trace_commitments[7].cap, dtype=np.uint64).reshape(-1)[:capw] if trace_commitments[7] is not None else np.zeros(capw, np.uint64)
    mem_after = np.array(trace_commitments[8].cap, dtype=np.uint64).reshape(-1)[:capw] if table_in_use[8] and trace_commitments[8] is not None \
        else np.zeros(capw, np.uint64)
    return proofs, st, mem_before, mem_after


class ZkGpuBackend:
    """The compute steps of one rank on its GPU (through the C ABI)."""

    def __init__(self, ctx, config, labels=None, precompute_constraints=False):
        self.ctx, self.config, self.labels = ctx, config, labels
        # table-sharded segments: evaluate the alpha-independent constraint values in begin(), i.e. while the transcript is with another
        # table, so that finish() — the serial relay — only combines them (include/zkgpu.h zkgpu_ctx_set_precompute_constraints)
        self.precompute_constraints = precompute_constraints
        self.phase_ms = None      # a dict: prove_with_traces_sharded adds the wall time of its phases (synchronising at every mark)

    def commit(self, table, trace):
        if isinstance(trace, tuple):      # (device address, n)
            info = _p.table_info(table, self.config.num_challenges)
            return _p.PolynomialBatch.from_device_values(self.ctx, trace[0], info["num_columns"], trace[1], self.config.rate_bits,
                                                         self.config.cap_height, keep_values=True)
        return _p.PolynomialBatch.from_values(self.ctx, trace, self.config.rate_bits, self.config.cap_height, keep_values=True)

    def cap(self, handle):
        return handle.cap

    def torch_stream(self):
        """the context's stream as torch sees it (collectives issued under it are ordered with the library's kernels on the device)"""
        if getattr(self, "_tstream", None) is None:
            import torch
            self._tstream = torch.cuda.ExternalStream(self.ctx.stream_handle(), device=torch.device("cuda", self.ctx.device))
        return self._tstream

    def buffer(self, key, shape):
        """a persistent int64 device buffer of this backend (one per key; reallocated when the shape changes)"""
        import torch
        bufs = self.__dict__.setdefault("_bufs", {})
        t = bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            with torch.cuda.stream(self.torch_stream()):
                t = bufs[key] = torch.empty(tuple(shape), dtype=torch.int64, device=torch.device("cuda", self.ctx.device))
        return t

    def split_commit(self, comm, table, trace):
        return SplitCommit(self, comm, table, trace)

    def sync(self):
        self.ctx.sync()

    def segment_challenges(self, caps, in_use, public_values):
        return segment_challenges(caps, in_use, self.config.cap_height, public_values, self.config.num_challenges)

    def begin(self, table, handle, beta_gamma, on_critical_path=False):
        """on_critical_path: nobody is ahead of this table in the relay, so there is no wait to hide the constraint recording in"""
        ctl = _p.get_ctl_data(self.ctx, table, handle, beta_gamma, self.config.num_challenges)
        job = C.c_void_p()
        self.ctx.set_precompute_constraints(self.precompute_constraints and not on_critical_path)
        try:
            check(lib().zkgpu_table_job_begin(self.ctx._h, C.c_uint32(table), C.byref(self.labels) if self.labels is not None else None,
                                              C.byref(self.config), handle._h, ctl._h, None, C.byref(job)))
        finally:
            self.ctx.set_precompute_constraints(False)
        return (job, ctl, handle)

    def finish(self, job, state, forced_pow=None):
        jh, ctl, handle = job
        st = np.ascontiguousarray(state, dtype=np.uint64).copy()
        fp = C.c_uint64(forced_pow) if forced_pow is not None else None
        h = C.c_void_p()
        try:
            check(lib().zkgpu_table_job_finish(jh, _ptr(st), C.byref(fp) if fp is not None else None, None, C.byref(h)))
        finally:
            lib().zkgpu_table_job_free(jh)
        sp = _p.StarkProof(self.ctx, h)
        words = sp.words
        sp.free(); ctl.free(); handle.free()
        return words, st


class TorchComm:
    """torch.distributed plumbing of the exchanges.  Two kinds of data move between the ranks of a table-sharded segment:
      * host data — the 9 trace caps (512 B each) and the 12-word transcript state of the relay: these go over a HOST channel (a gloo
        group next to the NCCL one).  They are host values on both sides (the transcript is host code, as in the reference), and a
        NCCL collective for them means an H2D copy, a kernel that spins on the device until the peer arrives, and a D2H copy — with the
        receiver's kernel spinning while that rank launches the auxiliary-polynomial work of its next table (measured: sporadic
        stalls of 0.1 - 1.8 s per proof, profiles/r2i);
      * device data — the LDE / coefficient / digest slices of the split commitments: NCCL all-gathers over NVLink
        (all_gather_device), ordered on the library's stream."""

    def __init__(self, group=None, device=None, host_group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.hgroup = host_group
        if self.hgroup is None and device is not None and dist.get_backend(group) != "gloo":
            ranks = None if group is None else dist.get_process_group_ranks(group)
            self.hgroup = dist.new_group(ranks=ranks, backend="gloo")      # collective: every rank builds its TorchComm at the same point
        if self.hgroup is None:
            self.hgroup = group

    def _t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())

    def _gsrc(self, src):
        # `src` is a rank of this communicator's group; torch.distributed wants the global rank
        return src if self.group is None else self.dist.get_global_rank(self.group, src)

    def all_gather(self, a):
        """(world, *a.shape) uint64"""
        t = self._t(a)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.hgroup)
        return np.stack([o.numpy().view(np.uint64) for o in out])

    def broadcast(self, a, src):
        t = self._t(a)
        self.dist.broadcast(t, src=self._gsrc(src), group=self.hgroup)
        return t.numpy().view(np.uint64)

    def broadcast_async(self, a, src):
        """the same broadcast, not waited for: .done() polls, .result() waits and returns the array"""
        t = self._t(a)
        work = self.dist.broadcast(t, src=self._gsrc(src), group=self.hgroup, async_op=True)

        class Pending:
            def done(self_):
                return work.is_completed()

            def result(self_):
                work.wait()
                return t.numpy().view(np.uint64)
        return Pending()

    def all_to_all_device(self, out, inp):
        """NCCL all-to-all of device tensors whose first dimension is the rank: out[i] = rank i's inp[my rank]; async work handle"""
        return self.dist.all_to_all_single(out, inp, group=self.group, async_op=True)

    def all_gather_device(self, out, inp):
        """NCCL all-gather of device tensors (inp may be the rank's own slot of out: in place); returns the async work handle"""
        return self.dist.all_gather_into_tensor(out, inp, group=self.group, async_op=True)

    def gather_proofs(self, proofs, owner):
        """every rank ends up with all proofs (object all-gather: proofs are a few hundred kB)"""
        mine = {t: p for t, p in enumerate(proofs) if p is not None and owner[t] == self.rank}
        out = [None] * self.world
        self.dist.all_gather_object(out, mine, group=self.group)
        res = [None] * NUM_TABLES
        for d in out:
            for t, p in d.items():
                res[t] = p
        return res


class LocalComm:
    """world of one (same code path as the sharded run, no process group)"""
    rank, world = 0, 1

    def all_gather(self, a):
        return np.asarray(a)[None]

    def broadcast(self, a, src):
        return np.asarray(a)

    def gather_proofs(self, proofs, owner):
        return proofs


NUM_COLUMNS = (116, 71, 85, 2431, 438, 523, 30, 12, 12)       # trace widths (zkgpu_table_info), for planning only
# auxiliary columns with two challenges: the per-row weight of phases 2-3 is ~ (c + a)
_NUM_AUX2 = (100, 70, 24, 4, 290, 2, 16, 4, 2)


class ShardPlan:
    """Who does what when the tables of ONE segment are spread over `world` GPUs.
    owner[t]  the rank that proves table t (auxiliary polynomials, quotient, openings, FRI: everything after the trace commitment);
    split[t]  True: the trace commitment of table t is computed by ALL ranks — column slices through ifft + LDE, NCCL all-gather of
              the slices over NVLink, every rank hashes one block of leaves and builds the Merkle levels under its cap entries,
              all-gather of the digests, the owner assembles the batch (include/zkgpu.h "S1 split over several devices").
              False: the owner commits the table alone."""

    def __init__(self, world, log_ns, owner, split):
        self.world, self.log_ns, self.owner, self.split = world, list(log_ns), list(owner), list(split)

    def needs(self, rank):
        """tables whose trace (for split tables: a column slice of it) rank `rank` reads"""
        return [lg is not None and (self.split[t] or self.owner[t] == rank) for t, lg in enumerate(self.log_ns)]

    def describe(self):
        return {"owner": list(self.owner), "split_over_all_gpus": [TABLE_NAMES[t] for t in range(NUM_TABLES) if self.split[t]]}


def shard_plan(world, log_ns, split_min_bytes=32 << 20, cap_height=4):
    """Split the trace commitment of every table whose trace has at least split_min_bytes (the all-gather of a small table costs more
    than hashing it alone); owners by longest-processing-time packing of the per-table weight of the phases that stay with the owner.
    A split commitment gives every rank one block of whole cap subtrees (zkgpu_merkle_block): the number of ranks must be a power of
    two and at most 2^cap_height.  Any other world size (3, 6, ... GPUs) still shards the tables by owner, without splitting."""
    in_use = [lg is not None for lg in log_ns]
    rows = [(1 << lg) if lg is not None else 0 for lg in log_ns]
    can_split = world > 1 and (world & (world - 1)) == 0 and world <= (1 << cap_height)
    split = [can_split and in_use[t] and 8 * NUM_COLUMNS[t] * rows[t] >= split_min_bytes for t in range(NUM_TABLES)]
    w = []
    for t in range(NUM_TABLES):
        own = rows[t] * (NUM_COLUMNS[t] + _NUM_AUX2[t] + 4) * 2.0                       # phases 2-3: streaming passes over trace + aux + quotient
        own += rows[t] * ((_NUM_AUX2[t] + 7) // 8) * 40.0                               # aux leaf hashing
        if not split[t]:
            own += rows[t] * ((NUM_COLUMNS[t] + 7) // 8) * 40.0 + rows[t] * NUM_COLUMNS[t] * 3.0   # the whole trace commitment as well
        w.append(own if in_use[t] else 0.0)
    return ShardPlan(world, log_ns, default_owner(world, weights=w), split)


class SplitCommit:
    """The trace commitment of one table computed by all ranks of the communicator (see ShardPlan.split).  Every device step is
    enqueued on the context's own stream and the NCCL collectives are ordered against it on the device (torch sees the stream as an
    ExternalStream): begin() of the next table runs while the all-gather of this one is in flight."""

    def __init__(self, backend, comm, table, trace):
        import torch
        self.torch, self.be, self.comm, self.table = torch, backend, comm, table
        ctx, cfg = backend.ctx, backend.config
        k, r = comm.world, comm.rank
        self.ncols = ncols = NUM_COLUMNS[table]
        device = isinstance(trace, tuple)
        self.n = n = int(trace[1]) if device else int(trace.shape[1])
        self.N = N = n << cfg.rate_bits
        self.cpr = cpr = -(-ncols // k)
        c0, c1 = min(ncols, r * cpr), min(ncols, (r + 1) * cpr)
        dev = torch.device("cuda", ctx.device)
        self.stream = backend.torch_stream()
        with torch.cuda.stream(self.stream):
            # the exchange buffers persist in the backend from proof to proof (backend.buffer): a fresh torch allocation per proof is
            # held back by the collectives' record_stream bookkeeping, the caching allocator then grows by ~10 GB per proof until it
            # has to cudaFree — measured as sporadic stalls of 80 ms to 1 s (profiles/r2h)
            self.lde = backend.buffer(("lde", table), (k * cpr, N))
            self.coef = backend.buffer(("coef", table), (k * cpr, n))
            # values: resident tables stay where they are (the owner reads its own copy); host tables are uploaded slice by slice —
            # 1/k of the PCIe time per device — and gathered like the rest
            self.vals = None if device else backend.buffer(("vals", table), (k * cpr, n))
            self.values_ptr = int(trace[0]) if device else self.vals.data_ptr()
            if device:
                src = int(trace[0]) + 8 * c0 * n
            else:
                a = trace if (trace.dtype == np.uint64 and trace.flags.c_contiguous) else np.ascontiguousarray(trace, dtype=np.uint64)
                self._keep = a
                src = a.ctypes.data + 8 * c0 * n
            slot = 8 * r * cpr
            check(lib().zkgpu_lde_slice(ctx._h, C.cast(C.c_void_p(src), u64p), 1 if device else 0, C.c_size_t(c1 - c0), C.c_size_t(n),
                                        C.c_uint32(cfg.rate_bits),
                                        None if device else C.cast(C.c_void_p(self.vals.data_ptr() + slot * n), u64p),
                                        C.cast(C.c_void_p(self.coef.data_ptr() + slot * n), u64p),
                                        C.cast(C.c_void_p(self.lde.data_ptr() + slot * N), u64p)))
            # what the hashing needs first: every rank's columns of THIS rank's block of rows (1/k of an all-gather, so the leaf hashing starts
            # almost at once); the all-gathers that give the owner the whole LDE / coefficients / values then run under the hashing
            self.per = per = N // k
            send = backend.buffer(("rows_out", table), (k, cpr, per))
            send.copy_(self.lde[r * cpr:(r + 1) * cpr].view(cpr, k, per).transpose(0, 1))
            self.rows = backend.buffer(("rows_in", table), (k, cpr, per))
            self.w_rows = comm.all_to_all_device(self.rows, send)
            self.w_rest = [comm.all_gather_device(self.lde, self.lde[r * cpr:(r + 1) * cpr]),
                           comm.all_gather_device(self.coef, self.coef[r * cpr:(r + 1) * cpr])]
            if self.vals is not None:
                self.w_rest.append(comm.all_gather_device(self.vals, self.vals[r * cpr:(r + 1) * cpr]))

    def hash_block(self):
        torch, ctx, cfg, k, r = self.torch, self.be.ctx, self.be.config, self.comm.world, self.comm.rank
        words = C.c_size_t()
        check(lib().zkgpu_merkle_block_words(C.c_size_t(self.N), C.c_uint32(cfg.cap_height), C.c_uint32(k), C.byref(words)))
        with torch.cuda.stream(self.stream):
            self.w_rows.wait()
            self.packed = self.be.buffer(("packed", self.table), (k, words.value))
            # self.rows is the block's (k * cpr) x per column-major LDE: the library addresses block r of an LDE with column pitch `per`
            # whose rows start r * per before it
            base = self.rows.data_ptr() - 8 * r * self.per
            check(lib().zkgpu_merkle_block(ctx._h, C.cast(C.c_void_p(base), u64p), C.c_size_t(self.per), C.c_size_t(self.ncols),
                                           C.c_size_t(self.N), C.c_uint32(cfg.cap_height), C.c_uint32(k), C.c_uint32(r),
                                           C.cast(C.c_void_p(self.packed.data_ptr() + 8 * r * words.value), u64p)))
            self.w_packed = self.comm.all_gather_device(self.packed, self.packed[r])

    def finish(self, is_owner):
        """-> the assembled PolynomialBatch on the owner (it borrows the gathered buffers), None elsewhere"""
        torch, ctx, cfg = self.torch, self.be.ctx, self.be.config
        with torch.cuda.stream(self.stream):
            for w in self.w_rest + [self.w_packed]:
                w.wait()
            if not is_owner:
                return None
            h = C.c_void_p()
            check(lib().zkgpu_batch_assemble(ctx._h, C.cast(C.c_void_p(self.values_ptr), u64p), C.cast(C.c_void_p(self.coef.data_ptr()), u64p),
                                             C.cast(C.c_void_p(self.lde.data_ptr()), u64p), C.cast(C.c_void_p(self.packed.data_ptr()), u64p),
                                             C.c_uint32(self.comm.world), C.c_size_t(self.ncols), C.c_size_t(self.n), C.c_uint32(cfg.rate_bits),
                                             C.c_uint32(cfg.cap_height), C.byref(h)))
        b = _p.PolynomialBatch(ctx, h)
        b._keepalive = (self.lde, self.coef, self.vals, self.packed, getattr(self, "_keep", None))
        return b


def _rows(trace):
    return 0 if trace is None else int(trace[1]) if isinstance(trace, tuple) else int(trace.shape[1])


def default_owner(world, weights=None):
    """table -> rank.  Longest-processing-time bin packing on per-table weights (default: a static cost model
    rows-independent ~ columns * ceil(columns / 8), dominated by Keccak, KeccakSponge, Logic, Arithmetic, Cpu)."""
    if weights is None:
        weights = [116 * 15, 71 * 9, 85 * 11, 2431 * 304, 438 * 55, 523 * 66, 30 * 4, 12 * 2, 12 * 2]
    load = [0.0] * world
    owner = [0] * NUM_TABLES
    for t in sorted(range(NUM_TABLES), key=lambda t: -weights[t]):
        r = min(range(world), key=lambda r: load[r])
        owner[t] = r
        load[r] += weights[t]
    return owner


def prove_with_traces_sharded(backend, comm, traces, table_in_use, public_values, owner=None, forced_pow_witnesses=None,
                              gather=True, cap_height=None, plan=None):
    """traces[t] is needed on owner[t] (and, for the tables the plan splits, on every rank: a column slice of it is read); host array
    or (device address, n); table_in_use must agree on every rank.  plan: a ShardPlan (shard_plan(world, log_ns)); without one the
    tables are only distributed whole (owner, default: default_owner)."""
    import time as _time
    if plan is not None:
        owner = plan.owner
    owner = owner if owner is not None else default_owner(comm.world)
    split = plan.split if plan is not None and comm.world > 1 else [False] * NUM_TABLES
    if cap_height is None:
        cap_height = getattr(getattr(backend, "config", None), "cap_height", 4)
    cap_words = 4 << cap_height
    phases = getattr(backend, "phase_ms", None)
    t_mark = [_time.perf_counter()]

    def mark(name):
        if phases is None:
            return
        if not phases.get("__nosync__"):
            backend.sync()
        now = _time.perf_counter()
        phases[name] = phases.get(name, 0.0) + (now - t_mark[0]) * 1e3
        t_mark[0] = now
    # phase 1: trace commitments — the split tables by all ranks (largest first, so that the all-gather of one runs under the LDE of
    # the next), then the tables this rank commits alone
    handles = {}
    caps = np.zeros((NUM_TABLES, cap_words), dtype=np.uint64)
    order = sorted((t for t in range(NUM_TABLES) if table_in_use[t] and split[t]), key=lambda t: -NUM_COLUMNS[t] * _rows(traces[t]))
    for t in order:
        if traces[t] is None:
            raise ValueError("rank %d takes part in the commitment of table %s but has no trace for it" % (comm.rank, TABLE_NAMES[t]))
    jobs1 = [backend.split_commit(comm, t, traces[t]) for t in order]
    mark("phase 1a: ifft + LDE of the column slices")
    for j in jobs1:
        j.hash_block()
    mark("phase 1b: all-gather of the LDEs, leaf blocks + Merkle levels")
    for t in range(NUM_TABLES):
        if table_in_use[t] and not split[t] and owner[t] == comm.rank:
            if traces[t] is None:
                raise ValueError("rank %d owns table %s but has no trace for it" % (comm.rank, TABLE_NAMES[t]))
            handles[t] = backend.commit(t, traces[t])
    mark("phase 1c: tables committed alone")
    for j in jobs1:
        b = j.finish(owner[j.table] == comm.rank)
        if b is not None:
            handles[j.table] = b
    for t, hnd in handles.items():
        caps[t] = np.asarray(backend.cap(hnd), dtype=np.uint64).ravel()
    mark("phase 1d: all-gather of coefficients / digests, assembly")
    # exchange 1: all-gather of the caps; row t of the result comes from the owner of table t
    allcaps = comm.all_gather(caps)
    caps = np.stack([allcaps[owner[t], t] for t in range(NUM_TABLES)])
    # transcript replay (identical on every rank)
    beta_gamma, state = backend.segment_challenges(caps, table_in_use, public_values)
    mark("exchange: caps + transcript")
    # phases 2 + 3.  Phase 3 is the relay of the transcript state in Table order (prover.rs:251-259: one challenger, the tables one
    # after the other), so at any time ONE rank is finishing a table and the others wait for the state: the auxiliary polynomials
    # (phase 2, challenger-independent) of a rank's tables are computed in those waits — up front only the first one, so that the
    # relay starts as early as it can; a table whose turn comes before its auxiliary polynomials exist gets them right then.
    mine = [t for t in range(NUM_TABLES) if t in handles]
    jobs, begun = {}, set()
    first_table = next(t for t in range(NUM_TABLES) if table_in_use[t])

    def begin(t):
        begun.add(t)
        if t == first_table and getattr(backend, "precompute_constraints", False):
            return backend.begin(t, handles[t], beta_gamma, on_critical_path=True)
        return backend.begin(t, handles[t], beta_gamma)

    def begin_next():
        for t in mine:
            if t not in begun:
                jobs[t] = begin(t)
                return True
        return False
    import os as _os
    mode = _os.environ.get("ZK_SHARD_BEGIN", "lazy")
    if mode == "upfront":
        while begin_next():
            pass
    else:
        begin_next()
    mark("phase 2: auxiliary commitment of the first local table")
    proofs = [None] * NUM_TABLES
    for t in range(NUM_TABLES):
        if not table_in_use[t]:
            continue
        if owner[t] == comm.rank:
            if t not in begun:
                jobs[t] = begin(t)
            fp = None if forced_pow_witnesses is None else int(forced_pow_witnesses[t])
            proofs[t], state = backend.finish(jobs.pop(t), state, fp)
            state = comm.broadcast(state, src=owner[t])
        elif hasattr(comm, "broadcast_async") and mode != "blocking":
            pending = comm.broadcast_async(state, src=owner[t])
            while not pending.done() and begin_next():
                pass
            state = pending.result()
        else:
            state = comm.broadcast(state, src=owner[t])
        mark("phase 3: relay, " + TABLE_NAMES[t])
    if gather:
        proofs = comm.gather_proofs(proofs, owner)
    return AllProof(proofs, np.asarray(beta_gamma), caps.reshape(NUM_TABLES, -1, 4), list(table_in_use))
