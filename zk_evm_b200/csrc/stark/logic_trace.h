// LogicStark trace rows, one row at a time (single source: the device trace-finishing kernel and its host test).
// Follows /root/reference/evm_arithmetization/src/logic.rs:165-237 (Operation::into_row, generate_trace_rows); column map: table_logic.h.
#pragma once
#include "hd.h"
#include "table_logic.h"

namespace zkstark { namespace logic {

static const uint32_t OP_WORDS = 9;    // operator (0 AND, 1 OR, 2 XOR: logic.rs Op), input0 and input1 as 4 little-endian u64 limbs each (U256)

// Row `row` of the trace of `num_ops` operations: st(column, value) for each of the 523 cells; rows past num_ops are all zero (:228-231).
template <class Store>
ZKS_HD void generate_row(const uint64_t* ops, uint64_t num_ops, uint64_t row, Store& st) {
    if (row >= num_ops) {
        ZKS_NOUNROLL
        for (uint32_t c = 0; c < NUM_COLUMNS; c++) st(c, (uint64_t)0);
        return;
    }
    const uint64_t* o = ops + row * OP_WORDS;
    const uint64_t op = o[0];
    st(IS_AND, (uint64_t)(op == 0));
    st(IS_OR, (uint64_t)(op == 1));
    st(IS_XOR, (uint64_t)(op == 2));
    ZKS_UNROLL
    for (uint32_t l = 0; l < 4; l++) {
        const uint64_t a = o[1 + l], b = o[5 + l];
        ZKS_NOUNROLL
        for (uint32_t z = 0; z < 64; z++) st(INPUT0 + 64 * l + z, (a >> z) & 1);
        ZKS_NOUNROLL
        for (uint32_t z = 0; z < 64; z++) st(INPUT1 + 64 * l + z, (b >> z) & 1);
        const uint64_t r = op == 0 ? (a & b) : op == 1 ? (a | b) : (a ^ b);
        st(RESULT + 2 * l, r & 0xFFFFFFFFull);
        st(RESULT + 2 * l + 1, r >> 32);
    }
}

}}  // namespace zkstark::logic
