"""Host-side mirror of the reference's proving interface over the C ABI.

Names follow the reference: `PolynomialBatch.from_values` (plonky2 fri/oracle.rs, called at
evm_arithmetization/src/prover.rs:100-107), `prove_single_table` / `prove_with_traces` (prover.rs:72,301).
numpy arrays stand in for Vec<PolynomialValues<F>>: shape (ncols, n), dtype uint64, C-contiguous == column-major
columns of n canonical Goldilocks elements.
"""
import ctypes as C
import weakref
import numpy as np
from . import _lib
from ._lib import u64p, check, lib


def _as_cols(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2:
        raise ValueError("expected a (ncols, n) uint64 array")
    return a


def _ptr(a):
    return a.ctypes.data_as(u64p)


class Context:
    """One per (host thread, GPU): owns the stream, the device memory pool and the cached twiddle tables."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().zkgpu_ctx_create(int(device), C.byref(self._h)))
        self.device = device
        self._children = weakref.WeakSet()   # device objects must be released before the context (its stream frees them)

    def close(self):
        if self._h:
            for ch in list(self._children):
                ch.free()
            lib().zkgpu_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().zkgpu_ctx_sync(self._h))

    def stream_handle(self):
        """the context's cudaStream_t as an integer (torch.cuda.ExternalStream(handle) orders a caller's device work with the library's)"""
        p = C.c_void_p()
        check(lib().zkgpu_ctx_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def set_precompute_constraints(self, on=True):
        """table jobs begun on this context evaluate the alpha-independent constraint values ahead of the transcript (table-sharded segments)"""
        check(lib().zkgpu_ctx_set_precompute_constraints(self._h, int(bool(on))))

    def set_timing(self, on=True):
        """stage spans (the reference's TimingTree): record a CUDA event at every stage boundary of the prover"""
        check(lib().zkgpu_ctx_set_timing(self._h, int(bool(on))))

    def timing_report(self):
        """-> [(stage name, milliseconds)] in order, for everything proved since set_timing(True)"""
        n = C.c_size_t(0)
        check(lib().zkgpu_ctx_timing_report(self._h, None, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        check(lib().zkgpu_ctx_timing_report(self._h, buf, C.byref(n)))
        return [(name, float(ms)) for name, ms in (line.split("\t") for line in buf.value.decode().splitlines())]

    def timer_start(self):
        check(lib().zkgpu_ctx_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().zkgpu_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def stats(self):
        k, u, p = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().zkgpu_ctx_stats(self._h, C.byref(k), C.byref(u), C.byref(p)))
        return {"kernel_launches": k.value, "bytes_in_use": u.value, "bytes_peak": p.value}

    # ---- fine-grained kernels -------------------------------------------------------------------------------
    def ntt(self, data, inverse=False, coset_shift=0):
        """fft / ifft / coset_fft / coset_ifft of every row of `data` (ncols, n); natural order in and out."""
        a = _as_cols(data).copy()
        check(lib().zkgpu_ntt(self._h, _ptr(a), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]), int(bool(inverse)),
                              C.c_uint64(int(coset_shift))))
        return a

    def poseidon_permute(self, states):
        s = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 12).copy()
        check(lib().zkgpu_poseidon_permute(self._h, _ptr(s), C.c_size_t(s.shape[0])))
        return s

    def poseidon_hash_rows(self, data):
        """hash_or_noop of every row; `data` is (width, nrows) i.e. column-major rows."""
        a = _as_cols(data)
        out = np.empty((a.shape[1], 4), dtype=np.uint64)
        check(lib().zkgpu_poseidon_hash_rows(self._h, _ptr(a), C.c_size_t(a.shape[1]), C.c_size_t(a.shape[0]), _ptr(out)))
        return out

    def bench_ntt(self, ncols, n, iters=5):
        ms, l = C.c_float(), C.c_uint64()
        check(lib().zkgpu_bench_ntt(self._h, C.c_size_t(ncols), C.c_size_t(n), int(iters), C.byref(ms), C.byref(l)))
        return ms.value, l.value

    def bench_leaf_hash(self, ncols, nrows, iters=5):
        ms = C.c_float()
        check(lib().zkgpu_bench_leaf_hash(self._h, C.c_size_t(ncols), C.c_size_t(nrows), int(iters), C.byref(ms)))
        return ms.value

    def bench_merkle_levels(self, nleaves, iters=5):
        ms = C.c_float()
        check(lib().zkgpu_bench_merkle_levels(self._h, C.c_size_t(nleaves), int(iters), C.byref(ms)))
        return ms.value


class PolynomialBatch:
    """Device-resident PolynomialBatch (polynomials + Merkle tree of the blown-up evaluations)."""

    def __init__(self, ctx, handle, borrowed=False):
        self.ctx = ctx
        self._h = handle
        self._borrowed = borrowed
        if not borrowed:
            ctx._children.add(self)
        nc, n, rb, ch = C.c_size_t(), C.c_size_t(), C.c_uint32(), C.c_uint32()
        check(lib().zkgpu_batch_dims(self._h, C.byref(nc), C.byref(n), C.byref(rb), C.byref(ch)))
        self.ncols, self.n, self.rate_bits, self.cap_height = nc.value, n.value, rb.value, ch.value

    @classmethod
    def from_values(cls, ctx, values, rate_bits=1, cap_height=4, keep_values=False):
        a = _as_cols(values)
        h = C.c_void_p()
        check(lib().zkgpu_commit_values_contig(ctx._h, _ptr(a), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]),
                                               C.c_uint32(rate_bits), C.c_uint32(cap_height), 0, int(keep_values),
                                               C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_device_values(cls, ctx, device_ptr, ncols, n, rate_bits=1, cap_height=4, keep_values=False):
        """Same as from_values for a trace already resident in HBM: `device_ptr` is the address of ncols*n contiguous
        column-major u64 (e.g. a torch.int64 CUDA tensor's data_ptr()); the data is copied device-to-device."""
        h = C.c_void_p()
        check(lib().zkgpu_commit_values_contig(ctx._h, C.cast(C.c_void_p(int(device_ptr)), u64p), C.c_size_t(ncols), C.c_size_t(n),
                                               C.c_uint32(rate_bits), C.c_uint32(cap_height), 1, int(keep_values), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_coeffs(cls, ctx, coeffs, rate_bits=1, cap_height=4):
        a = _as_cols(coeffs)
        ptrs = (u64p * a.shape[0])(*[a[i].ctypes.data_as(u64p) for i in range(a.shape[0])])
        h = C.c_void_p()
        check(lib().zkgpu_commit_coeffs(ctx._h, ptrs, C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]),
                                        C.c_uint32(rate_bits), C.c_uint32(cap_height), 0, C.byref(h)))
        return cls(ctx, h)

    def free(self):
        if self._h:
            if not self._borrowed:
                lib().zkgpu_batch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def cap(self):
        out = np.empty((1 << self.cap_height, 4), dtype=np.uint64)
        check(lib().zkgpu_batch_cap(self._h, _ptr(out)))
        return out

    def export(self, coeffs=True, leaves=True, digests=True):
        """Host copies of PolynomialBatch.{polynomials, merkle_tree.leaves, merkle_tree.digests} (plonky2 layouts)."""
        N = self.n << self.rate_bits
        co = np.empty((self.ncols, self.n), dtype=np.uint64) if coeffs else None
        le = np.empty((N, self.ncols), dtype=np.uint64) if leaves else None
        nd = 2 * (N - (1 << self.cap_height))
        di = np.empty((nd, 4), dtype=np.uint64) if digests else None
        check(lib().zkgpu_batch_export(self._h, _ptr(co) if coeffs else None, _ptr(le) if leaves else None,
                                       _ptr(di) if digests else None))
        return co, le, di


class CtlData:
    """Device-resident CtlData of one table (the per-table slice of starky get_ctl_data, prover.rs:137-143)."""

    def __init__(self, ctx, handle, table, n, num_challenges):
        self.ctx, self._h, self.table, self.n, self.num_challenges = ctx, handle, table, n, num_challenges
        ctx._children.add(self)

    def export(self):
        info = table_info(self.table, self.num_challenges)
        k = info["num_ctl_helper_columns"] + info["num_ctl_zs"]
        out = np.empty((k, self.n), dtype=np.uint64)
        if k:
            check(lib().zkgpu_ctl_export(self._h, _ptr(out)))
        return out

    def free(self):
        if self._h:
            lib().zkgpu_ctl_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class StarkProof:
    """StarkProofWithMetadata of one table; `words` is the canonical serialisation (csrc/stark/proof.h)."""

    def __init__(self, ctx, handle):
        self._h = handle
        ctx._children.add(self)
        n = C.c_size_t(0)
        check(lib().zkgpu_proof_serialize(self._h, None, C.byref(n)))
        self.words = np.empty(n.value, dtype=np.uint64)
        check(lib().zkgpu_proof_serialize(self._h, _ptr(self.words), C.byref(n)))

    def debug_batch(self, ctx, which):
        h = C.c_void_p()
        check(lib().zkgpu_proof_debug_batch(self._h, int(which), C.byref(h)))
        return PolynomialBatch(ctx, h, borrowed=True)

    def debug_fri_values(self):
        n = C.c_size_t(0)
        check(lib().zkgpu_proof_debug_fri_values(self._h, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint64)
        check(lib().zkgpu_proof_debug_fri_values(self._h, _ptr(out), C.byref(n)))
        return out.reshape(-1, 2)

    def free(self):
        if self._h:
            lib().zkgpu_proof_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """A page-locked (ncols, n) uint64 host array (zkgpu_host_alloc): trace uploads from it run at PCIe speed, asynchronously.
    `array` is a numpy view of the buffer; keep the PinnedArray alive while the view is in use."""

    def __init__(self, shape):
        self._p = C.c_void_p()
        nbytes = int(np.prod(shape)) * 8
        check(lib().zkgpu_host_alloc(C.c_size_t(nbytes), C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, u64p), shape=(int(np.prod(shape)),)).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().zkgpu_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_register(a):
    """page-lock an existing contiguous numpy array in place (zkgpu_host_register); pair with host_unregister"""
    check(lib().zkgpu_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))


def host_unregister(a):
    check(lib().zkgpu_host_unregister(C.c_void_p(a.ctypes.data)))


class DeviceTrace:
    """A table's trace finished in device memory (zkgpu_dev_trace): ncols x n, column-major.  Goes into `prove_with_traces` /
    `upload_traces` in place of a host array, or into PolynomialBatch.from_device_values via `device_ptr`."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        ctx._children.add(self)
        nc, n = C.c_size_t(), C.c_size_t()
        check(lib().zkgpu_dev_trace_dims(self._h, C.byref(nc), C.byref(n)))
        self.ncols, self.n = nc.value, n.value
        self.device_ptr = int(lib().zkgpu_dev_trace_ptr(self._h) or 0)

    def export(self):
        out = np.empty((self.ncols, self.n), dtype=np.uint64)
        check(lib().zkgpu_dev_trace_export(self._h, _ptr(out)))
        return out

    def free(self):
        if self._h:
            lib().zkgpu_dev_trace_free(self._h)
            self._h = C.c_void_p()
            self.device_ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def keccak_generate_trace(ctx, inputs, timestamps, min_rows=0):
    """KeccakStark::generate_trace on the device (keccak_stark.rs:70-250): inputs (num_perms, 25) uint64 (lane y*5 + x), one timestamp
    per permutation -> DeviceTrace of 2431 columns x max(24 * num_perms, min_rows).next_power_of_two() rows."""
    a = np.ascontiguousarray(inputs, dtype=np.uint64).reshape(-1, 25)
    ts = np.ascontiguousarray(timestamps, dtype=np.uint64).ravel()
    if ts.size != a.shape[0]:
        raise ValueError("one timestamp per permutation")
    h = C.c_void_p()
    check(lib().zkgpu_keccak_generate_trace(ctx._h, _ptr(a), _ptr(ts), C.c_size_t(a.shape[0]), C.c_size_t(min_rows), C.byref(h)))
    return DeviceTrace(ctx, h)


def logic_generate_trace(ctx, ops, min_rows=0):
    """LogicStark::generate_trace on the device (logic.rs:165-237): ops (num_ops, 9) uint64 = operator (0 AND, 1 OR, 2 XOR), input0 and
    input1 as 4 little-endian u64 limbs each -> DeviceTrace of 523 columns x max(num_ops, min_rows).next_power_of_two() rows."""
    a = np.ascontiguousarray(ops, dtype=np.uint64).reshape(-1, 9)
    h = C.c_void_p()
    check(lib().zkgpu_logic_generate_trace(ctx._h, _ptr(a), C.c_size_t(a.shape[0]), C.c_size_t(min_rows), C.byref(h)))
    return DeviceTrace(ctx, h)


def upload_trace(ctx, trace):
    """a host trace (ncols, n) moved to the device as a DeviceTrace (for the in-place finishing steps)"""
    a = _as_cols(trace)
    h = C.c_void_p()
    check(lib().zkgpu_dev_trace_upload(ctx._h, _ptr(a), C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]), C.byref(h)))
    ctx.sync()          # `a` may be a temporary
    return DeviceTrace(ctx, h)


def arithmetic_generate_range_checks(ctx, device_trace):
    """ArithmeticStark::generate_range_checks (arithmetic_stark.rs:130-156) in place on a device-resident Arithmetic trace"""
    check(lib().zkgpu_arithmetic_generate_range_checks(ctx._h, device_trace._h))
    return device_trace


def memory_finish_trace(ctx, ops, stale_contexts=()):
    """The data-parallel tail of MemoryStark::generate_trace on the device (memory_stark.rs:104-294): ops (14, n) uint64 = filter,
    timestamp, is_read, context, segment, virtual, 8 value limbs of the sorted / gap-filled / padded operations; stale_contexts as
    insert_stale_contexts takes them -> DeviceTrace of 30 columns x n."""
    a = _as_cols(ops)
    if a.shape[0] != 14:
        raise ValueError("expected the 14 operation columns")
    st = np.ascontiguousarray(list(stale_contexts), dtype=np.uint64)
    h = C.c_void_p()
    check(lib().zkgpu_memory_finish_trace(ctx._h, _ptr(a), C.c_size_t(a.shape[1]), _ptr(st) if st.size else None, C.c_size_t(st.size), C.byref(h)))
    return DeviceTrace(ctx, h)


def table_info(table, num_challenges):
    a, b, c_, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(lib().zkgpu_table_info(C.c_uint32(table), C.c_uint32(num_challenges), C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
    return {"num_columns": a.value, "num_lookup_columns": b.value, "num_ctl_helper_columns": c_.value, "num_ctl_zs": d.value}


def get_ctl_data(ctx, table, trace_batch, beta_gamma, num_challenges):
    bg = np.ascontiguousarray(beta_gamma, dtype=np.uint64)
    h = C.c_void_p()
    check(lib().zkgpu_ctl_data(ctx._h, C.c_uint32(table), trace_batch._h, _ptr(bg), C.c_uint32(num_challenges), C.byref(h)))
    return CtlData(ctx, h, table, trace_batch.n, num_challenges)


def prove_single_table(ctx, table, config, trace_batch, ctl_data, challenger_state, labels=None, forced_pow_witness=None,
                       abort_flag=None):
    """prover.rs:301-341.  Returns (StarkProof, challenger state after the proof)."""
    st = np.ascontiguousarray(challenger_state, dtype=np.uint64).copy()
    h = C.c_void_p()
    fp = C.c_uint64(forced_pow_witness) if forced_pow_witness is not None else None
    check(lib().zkgpu_prove_table(ctx._h, C.c_uint32(table), C.byref(labels) if labels is not None else None, C.byref(config),
                                  trace_batch._h, ctl_data._h, _ptr(st), C.byref(fp) if fp is not None else None,
                                  C.byref(abort_flag) if abort_flag is not None else None, C.byref(h)))
    return StarkProof(ctx, h), st


def set_debug(ctx, on=True):
    check(lib().zkgpu_ctx_set_debug(ctx._h, int(bool(on))))
