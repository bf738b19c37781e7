// BytePackingStark.
// Source: /root/reference/evm_arithmetization/src/byte_packing/{columns.rs:10-26, byte_packing_stark.rs:57-150 (CTL), 296-352
// (constraints), 426-437 (lookup)}.
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace byte_packing {

static const uint32_t NUM_BYTES = 32;
enum : uint32_t { IS_READ = 0, INDEX_LEN = 1, ADDR_CONTEXT = 33, ADDR_SEGMENT = 34, ADDR_VIRTUAL = 35, TIMESTAMP = 36, VALUE_BYTES = 37,
                  RANGE_COUNTER = 69, RC_FREQUENCIES = 70, NUM_COLUMNS = 71 };
static const uint64_t BYTE_RANGE_MAX = 256;

template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    // range column: starts at 0, ends at 255, increments by 0 or 1
    P rc1 = lv[RANGE_COUNTER], rc2 = nv[RANGE_COUNTER];
    yc.constraint_first_row(rc1);
    P incr = rc2 - rc1;
    yc.constraint_transition(incr * incr - incr);
    yc.constraint_last_row(rc1 - P::from_u64(BYTE_RANGE_MAX - 1));
    P current_filter = P::zero();
    for (uint32_t i = 0; i < NUM_BYTES; i++) current_filter = current_filter + lv[INDEX_LEN + i];
    yc.constraint(current_filter * (current_filter - one));
    yc.constraint_first_row(current_filter - one);
    P current_is_read = lv[IS_READ];
    yc.constraint(current_is_read * (current_is_read - one));
    for (uint32_t i = 0; i < NUM_BYTES; i++) { P idx = lv[INDEX_LEN + i]; yc.constraint(idx * (idx - one)); }
    P next_filter = P::zero();
    for (uint32_t i = 0; i < NUM_BYTES; i++) next_filter = next_filter + nv[INDEX_LEN + i];
    yc.constraint_transition(next_filter * (next_filter - current_filter));
    // all bytes after the final length are 0
    for (uint32_t i = 0; i + 1 < NUM_BYTES; i++) {
        P idx = lv[INDEX_LEN + i];
        for (uint32_t j = i + 1; j < NUM_BYTES; j++) yc.constraint(idx * lv[VALUE_BYTES + j]);
    }
}

inline std::vector<Column> ctl_looked_data() {
    std::vector<Column> res = Column::singles({IS_READ, ADDR_CONTEXT, ADDR_SEGMENT, ADDR_VIRTUAL});
    std::vector<std::pair<uint32_t, uint64_t>> len;
    for (uint32_t i = 0; i < NUM_BYTES; i++) len.push_back({INDEX_LEN + i, i + 1});
    res.push_back(Column::linear_combination(len));
    res.push_back(Column::single(TIMESTAMP));
    for (uint32_t i = 0; i < 8; i++) {
        std::vector<std::pair<uint32_t, uint64_t>> t;
        for (uint32_t j = 0; j < 4; j++) t.push_back({VALUE_BYTES + 4 * i + j, 1ULL << (8 * j)});
        res.push_back(Column::linear_combination(t));
    }
    return res;
}
inline Filter ctl_looked_filter() {
    std::vector<uint32_t> cs; for (uint32_t i = 0; i < NUM_BYTES; i++) cs.push_back(INDEX_LEN + i);
    return Filter::new_simple(Column::sum(cs));
}
inline std::vector<Column> ctl_looking_memory(uint32_t i) {
    std::vector<Column> res = Column::singles({IS_READ, ADDR_CONTEXT, ADDR_SEGMENT});
    // virtual address: ADDR_VIRTUAL + sequence_len - 1 - i
    std::vector<std::pair<uint32_t, uint64_t>> t = {{ADDR_VIRTUAL, 1}};
    for (uint32_t j = 0; j < NUM_BYTES; j++) t.push_back({INDEX_LEN + j, j});
    res.push_back(Column::linear_combination_with_constant(t, neg_const(i)));
    res.push_back(Column::single(VALUE_BYTES + i));
    for (uint32_t k = 1; k < 8; k++) res.push_back(Column::zero());
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline Filter ctl_looking_memory_filter(uint32_t i) {
    std::vector<uint32_t> cs; for (uint32_t k = i; k < NUM_BYTES; k++) cs.push_back(INDEX_LEN + k);
    return Filter::new_simple(Column::sum(cs));
}
inline std::vector<Lookup> lookups() {
    Lookup l;
    for (uint32_t i = 0; i < NUM_BYTES; i++) { l.columns.push_back(Column::single(VALUE_BYTES + i)); l.filter_columns.push_back(Filter()); }
    l.table_column = Column::single(RANGE_COUNTER);
    l.frequencies_column = Column::single(RC_FREQUENCIES);
    return {l};
}

}}  // namespace zkstark::byte_packing
