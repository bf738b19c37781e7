"""Synthetic trace generators (numpy) for the STARK tables: valid traces for the tables whose row semantics are simple
enough to restate here, random traces for throughput.  Layouts: (ncols, n) uint64, column-major.

Sources: LogicStark rows /root/reference/evm_arithmetization/src/logic.rs:165-237; MemoryContinuationStark
memory_continuation_stark.rs:54-100; MemoryStark columns memory/columns.rs:10-51 + generate_first_change_flags_and_rc
memory_stark.rs:134-200; padding CPU rows generation/mod.rs:646-663."""
import numpy as np

P = 0xFFFFFFFF00000001
T_ARITHMETIC, T_BYTE_PACKING, T_CPU, T_KECCAK, T_KECCAK_SPONGE, T_LOGIC, T_MEMORY, T_MEM_BEFORE, T_MEM_AFTER = range(9)
NUM_COLUMNS = {T_ARITHMETIC: 116, T_BYTE_PACKING: 71, T_CPU: 85, T_KECCAK: 2431, T_KECCAK_SPONGE: 438, T_LOGIC: 523,
               T_MEMORY: 30, T_MEM_BEFORE: 12, T_MEM_AFTER: 12}


def random_trace(table, log_n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, P, size=(NUM_COLUMNS[table], 1 << log_n), dtype=np.uint64)


def memcont_trace(log_n, seed, fill=0.7):
    """filter in {0,1}; address and value limbs arbitrary 32-bit values"""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    t = rng.integers(0, 1 << 32, size=(12, n), dtype=np.uint64)
    t[0] = (rng.random(n) < fill).astype(np.uint64)
    return t


def logic_trace(log_n, seed, fill=0.8):
    """random AND/OR/XOR rows, the rest padding (all-zero rows)"""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    t = np.zeros((523, n), dtype=np.uint64)
    nops = int(n * fill)
    op = rng.integers(0, 3, size=nops)
    a = rng.integers(0, 2, size=(256, nops), dtype=np.uint64)
    b = rng.integers(0, 2, size=(256, nops), dtype=np.uint64)
    for k in range(3):
        t[k, :nops] = (op == k)
    t[3:259, :nops] = a
    t[259:515, :nops] = b
    res = np.where(op == 0, a & b, np.where(op == 1, a | b, a ^ b))
    w = (np.uint64(1) << np.arange(32, dtype=np.uint64))[:, None]
    for l in range(8):
        t[515 + l, :nops] = (res[32 * l:32 * l + 32] * w).sum(axis=0)
    return t


def memory_trace_simple(log_n):
    """A valid MemoryStark trace made of dummy reads at (0, 0, r), timestamp 0: every ordering, initialisation and
    range-check constraint holds with range_check == 0 and all frequencies on counter 0."""
    n = 1 << log_n
    t = np.zeros((30, n), dtype=np.uint64)
    t[3] = 1                                   # is_read
    t[6] = np.arange(n, dtype=np.uint64)       # addr_virtual
    t[17] = 1                                  # virtual_first_change (also on the last row: wraps to row 0)
    t[20] = 34 * 35                            # preinitialized_segments_aux = (0-34)(0-35)
    t[28] = np.arange(n, dtype=np.uint64)      # counter
    t[29, 0] = n                               # frequencies: range_check == 0 on every row
    return t
