// Host build of the single-source Keccak / Logic row generators (zk_evm_b200/csrc/stark/{keccak,logic}_trace.h) for tests/test_trace_gen_host.py:
// the same code the device kernel keccak_trace_kernel runs, one "thread" per row, writing the column-major trace.
#include <stdint.h>
#include <stddef.h>
#include "stark/keccak_trace.h"
#include "stark/logic_trace.h"
#include "stark/memory_trace.h"
#include <string.h>

struct HostStore {
    uint64_t* out; size_t n; uint32_t count;
    void operator()(uint32_t col, uint64_t v) { out[(size_t)col * n] = v; count++; }
};

// returns the number of cells written per row (must be 2431 for every row) or 0 on a mismatch
extern "C" uint32_t keccak_trace_rows(const uint64_t* inputs, const uint64_t* timestamps, uint64_t num_perms, size_t n, uint64_t* out) {
    uint32_t cells = 0;
    for (size_t row = 0; row < n; row++) {
        HostStore st{out + row, n, 0};
        zkstark::keccak::generate_row(inputs, timestamps, num_perms, row, st);
        if (row && st.count != cells) return 0;
        cells = st.count;
    }
    return cells;
}

extern "C" void keccak_permutation_output(const uint64_t* input, uint64_t* out) { zkstark::keccak::permutation_output(input, out); }

extern "C" uint32_t logic_trace_rows(const uint64_t* ops, uint64_t num_ops, size_t n, uint64_t* out) {
    uint32_t cells = 0;
    for (size_t row = 0; row < n; row++) {
        HostStore st{out + row, n, 0};
        zkstark::logic::generate_row(ops, num_ops, row, st);
        if (row && st.count != cells) return 0;
        cells = st.count;
    }
    return cells;
}

// MemoryStark finishing as the device does it: the 14 operation columns and the stale list in place, then finish_row per row and the
// histogram contributions applied; returns 0 if a row reports a range-check overflow
extern "C" int memory_finish_rows(const uint64_t* ops, size_t n, const uint64_t* stale, size_t num_stale, uint64_t* t) {
    namespace mem = zkstark::memory;
    memset(t, 0xAB, 30 * n * 8);                                  // every cell must be written by the finishing code
    memcpy(t + mem::FILTER * n, ops, 2 * n * 8);
    memcpy(t + mem::IS_READ * n, ops + 2 * n, 12 * n * 8);
    memset(t + mem::STALE_CONTEXTS * n, 0, 3 * n * 8);
    memset(t + mem::FREQUENCIES * n, 0, n * 8);
    for (size_t k = 0; k < num_stale; k++) { t[mem::STALE_CONTEXTS * n + stale[k]] = stale[k] + 1; t[mem::IS_PRUNED * n + stale[k]] = 1; }
    for (size_t i = 0; i < n; i++) {
        mem::RowCounts rc;
        if (!mem::finish_row(t, n, i, rc)) return 0;
        t[mem::FREQUENCIES * n + rc.freq_a]++;
        if (rc.freq_b != mem::NONE) t[mem::FREQUENCIES * n + rc.freq_b]++;
        if (rc.stale_ctx != mem::NONE) t[mem::STALE_CONTEXT_FREQUENCIES * n + rc.stale_ctx]++;
    }
    return 1;
}
