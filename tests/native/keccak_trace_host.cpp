// Host build of the single-source Keccak / Logic row generators (zk_evm_b200/csrc/stark/{keccak,logic}_trace.h) for tests/test_trace_gen_host.py:
// the same code the device kernel keccak_trace_kernel runs, one "thread" per row, writing the column-major trace.
#include <stdint.h>
#include <stddef.h>
#include "stark/keccak_trace.h"
#include "stark/logic_trace.h"

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
