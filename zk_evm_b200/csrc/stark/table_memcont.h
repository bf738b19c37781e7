// MemoryContinuationStark (MemBefore / MemAfter tables).
// Source: /root/reference/evm_arithmetization/src/memory_continuation/{columns.rs:7-23, memory_continuation_stark.rs:28-53,110-122}
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace memcont {

enum : uint32_t { FILTER = 0, ADDR_CONTEXT = 1, ADDR_SEGMENT = 2, ADDR_VIRTUAL = 3, VALUE_START = 4, NUM_COLUMNS = 12 };
inline uint32_t value_limb(uint32_t i) { return VALUE_START + i; }

template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& /*nv*/, CC& yc) {
    // The filter must be binary.
    P filter = lv[FILTER];
    yc.constraint(filter * (filter - P::one()));
}

inline std::vector<Column> ctl_data() {
    std::vector<Column> res = Column::singles({ADDR_CONTEXT, ADDR_SEGMENT, ADDR_VIRTUAL});
    for (uint32_t i = 0; i < 8; i++) res.push_back(Column::single(value_limb(i)));
    return res;
}
inline Filter ctl_filter() { return Filter::new_simple(Column::single(FILTER)); }
inline std::vector<Column> ctl_data_memory() {
    std::vector<Column> res = {Column::zero()};   // IS_READ
    for (auto& c : Column::singles({ADDR_CONTEXT, ADDR_SEGMENT, ADDR_VIRTUAL})) res.push_back(c);
    for (uint32_t i = 0; i < 8; i++) res.push_back(Column::single(value_limb(i)));
    res.push_back(Column::zero());                // TIMESTAMP
    return res;
}
inline std::vector<Lookup> lookups() { return {}; }

}}  // namespace zkstark::memcont
