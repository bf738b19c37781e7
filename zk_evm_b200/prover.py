"""Host-side mirror of the reference's proving interface over the C ABI.

Names follow the reference: `PolynomialBatch.from_values` (plonky2 fri/oracle.rs, called at
evm_arithmetization/src/prover.rs:100-107), `prove_single_table` / `prove_with_traces` (prover.rs:72,301).
numpy arrays stand in for Vec<PolynomialValues<F>>: shape (ncols, n), dtype uint64, C-contiguous == column-major
columns of n canonical Goldilocks elements.
"""
import ctypes as C
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

    def close(self):
        if self._h:
            lib().zkgpu_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().zkgpu_ctx_sync(self._h))

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

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._h = handle
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
    def from_coeffs(cls, ctx, coeffs, rate_bits=1, cap_height=4):
        a = _as_cols(coeffs)
        ptrs = (u64p * a.shape[0])(*[a[i].ctypes.data_as(u64p) for i in range(a.shape[0])])
        h = C.c_void_p()
        check(lib().zkgpu_commit_coeffs(ctx._h, ptrs, C.c_size_t(a.shape[0]), C.c_size_t(a.shape[1]),
                                        C.c_uint32(rate_bits), C.c_uint32(cap_height), 0, C.byref(h)))
        return cls(ctx, h)

    def free(self):
        if self._h:
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
