"""Segment trace container: the hand-off format between the Rust witness generator and this prover.

What `prove_with_traces` consumes is `[Vec<PolynomialValues<F>>; NUM_TABLES]` (witness/traces.rs:243-258) plus `table_in_use`, the flattened
`PublicValues` (proof.rs:68-90, observed in the order of get_challenges.rs:202-227) and the four kernel labels CpuStark's constraints need.
A Rust host that cannot link libzkgpu directly writes this file; `load` memory-maps it so the column data goes to the GPU without a copy
through Python (`prove_with_traces(ctx, seg.traces, seg.public_values, ...)`).

Layout (little endian, every section 64-byte aligned so that a mapped column block can be pinned / DMA'd as is):
    header   magic "ZKSEGTR1", u32 version = 1, u32 num_tables = 9, u64 n_public_values, u64 labels[4]
             (halt_final, init, syscall_jumptable, exception_jumptable), then per table: u32 in_use, u32 num_columns, u64 n, u64 offset
    public   n_public_values x u64
    tables   for each table in use, at `offset`: num_columns x n canonical Goldilocks u64, column-major
             (column c = elements [c*n, (c+1)*n), exactly PolynomialValues<F>.values of trace column c)
"""
import struct
import numpy as np

MAGIC = b"ZKSEGTR1"
VERSION = 1
NUM_TABLES = 9
NUM_COLUMNS = (116, 71, 85, 2431, 438, 523, 30, 12, 12)       # table_num_columns of csrc/stark/all_stark.h
OPTIONAL_TABLES = (1, 3, 4, 5, 8)                            # OPTIONAL_TABLE_INDICES, all_stark.rs:110-117
P = 0xFFFFFFFF00000001
_ALIGN = 64
_HEAD = struct.Struct("<8sIIQ4Q")
_ENTRY = struct.Struct("<IIQQ")


class TraceFileError(ValueError):
    pass


class SegmentTraces:
    def __init__(self, traces, public_values, labels):
        self.traces, self.public_values, self.labels = traces, public_values, labels

    @property
    def table_in_use(self):
        return [t is not None for t in self.traces]


def _align(x):
    return (x + _ALIGN - 1) // _ALIGN * _ALIGN


def save(path, traces, public_values, labels):
    """traces: list of 9 (num_columns, n) uint64 arrays or None (optional table not in use)."""
    if len(traces) != NUM_TABLES or len(labels) != 4:
        raise TraceFileError("expected 9 tables and 4 kernel labels")
    pv = np.ascontiguousarray(public_values, dtype="<u8").ravel()
    arrs, entries = [], []
    off = _align(_HEAD.size + NUM_TABLES * _ENTRY.size) + _align(pv.nbytes)
    for t, tr in enumerate(traces):
        if tr is None:
            if t not in OPTIONAL_TABLES:
                raise TraceFileError("table %d is mandatory (all_stark.rs:110-117)" % t)
            arrs.append(None); entries.append((0, NUM_COLUMNS[t], 0, 0))
            continue
        a = np.ascontiguousarray(tr, dtype="<u8")
        if a.ndim != 2 or a.shape[0] != NUM_COLUMNS[t]:
            raise TraceFileError("table %d: expected a (%d, n) trace" % (t, NUM_COLUMNS[t]))
        n = a.shape[1]
        if n == 0 or n & (n - 1):
            raise TraceFileError("table %d: the number of rows must be a power of two" % t)
        if a.size and int(a.max()) >= P:
            raise TraceFileError("table %d: non-canonical field element" % t)
        arrs.append(a); entries.append((1, NUM_COLUMNS[t], n, off))
        off = _align(off + a.nbytes)
    with open(path, "wb") as f:
        f.write(_HEAD.pack(MAGIC, VERSION, NUM_TABLES, pv.size, *[int(x) for x in labels]))
        for e in entries:
            f.write(_ENTRY.pack(*e))
        f.seek(_align(_HEAD.size + NUM_TABLES * _ENTRY.size))
        f.write(pv.tobytes())
        for a, e in zip(arrs, entries):
            if a is not None:
                f.seek(e[3])
                f.write(a.tobytes())
        f.truncate(off)


def load(path, mmap=True, check_canonical=True):
    """-> SegmentTraces; with mmap the column blocks are read-only views of the file.  check_canonical=False only for a trusted
    producer: the device field forms assume every element is < p, a larger word silently changes commitments and proofs."""
    with open(path, "rb") as f:
        head = f.read(_HEAD.size)
        if len(head) < _HEAD.size:
            raise TraceFileError("truncated header")
        magic, version, ntab, npv, *labels = _HEAD.unpack(head)
        if magic != MAGIC:
            raise TraceFileError("not a segment trace file")
        if version != VERSION or ntab != NUM_TABLES:
            raise TraceFileError("unsupported version / table count")
        raw = f.read(NUM_TABLES * _ENTRY.size)
        if len(raw) < NUM_TABLES * _ENTRY.size:
            raise TraceFileError("truncated table directory")
        entries = [_ENTRY.unpack_from(raw, i * _ENTRY.size) for i in range(NUM_TABLES)]
        f.seek(0, 2)
        size = f.tell()
    pv_off = _align(_HEAD.size + NUM_TABLES * _ENTRY.size)
    if pv_off + 8 * npv > size:
        raise TraceFileError("truncated public values")
    data = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    pv = np.frombuffer(data, dtype="<u8", count=npv, offset=pv_off).astype(np.uint64)
    traces = []
    for t, (in_use, nc, n, off) in enumerate(entries):
        if nc != NUM_COLUMNS[t]:
            raise TraceFileError("table %d: %d columns, this prover expects %d" % (t, nc, NUM_COLUMNS[t]))
        if not in_use:
            if t not in OPTIONAL_TABLES:
                raise TraceFileError("table %d is mandatory" % t)
            traces.append(None)
            continue
        if n == 0 or n & (n - 1) or off % _ALIGN or off + 8 * nc * n > size:
            raise TraceFileError("table %d: bad shape or offset" % t)
        a = np.frombuffer(data, dtype="<u8", count=nc * n, offset=off).reshape(nc, n)
        if check_canonical and int(a.max()) >= P:
            raise TraceFileError("table %d: non-canonical field element" % t)
        traces.append(a)
    return SegmentTraces(traces, pv, tuple(labels))
