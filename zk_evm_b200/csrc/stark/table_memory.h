// MemoryStark.
// Source: /root/reference/evm_arithmetization/src/memory/{columns.rs:10-88, memory_stark.rs:35-95 (CTL), 474-626 (constraints),
// 859-887 (lookups)}; segment numbers memory/segments.rs:14-88.
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace memory {

enum : uint32_t {
    FILTER = 0, TIMESTAMP = 1, TIMESTAMP_INV = 2, IS_READ = 3, ADDR_CONTEXT = 4, ADDR_SEGMENT = 5, ADDR_VIRTUAL = 6,
    VALUE_LIMBS0 = 7,   // 7..14
    CONTEXT_FIRST_CHANGE = 15, SEGMENT_FIRST_CHANGE = 16, VIRTUAL_FIRST_CHANGE = 17, INITIALIZE_AUX = 18,
    PREINITIALIZED_SEGMENTS = 19, PREINITIALIZED_SEGMENTS_AUX = 20, STALE_CONTEXTS = 21, IS_PRUNED = 22,
    STALE_CONTEXT_FREQUENCIES = 23, IS_STALE = 24, MAYBE_IN_MEM_AFTER = 25, MEM_AFTER_FILTER = 26, RANGE_CHECK = 27,
    COUNTER = 28, FREQUENCIES = 29, NUM_COLUMNS = 30
};
enum : uint64_t { SEG_CODE = 0, SEG_TRIE_DATA = 12, SEG_ACCOUNTS_LINKED_LIST = 34, SEG_STORAGE_LINKED_LIST = 35 };

template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    P timestamp = lv[TIMESTAMP], addr_context = lv[ADDR_CONTEXT], addr_segment = lv[ADDR_SEGMENT],
      addr_virtual = lv[ADDR_VIRTUAL], timestamp_inv = lv[TIMESTAMP_INV], is_stale = lv[IS_STALE],
      maybe_in_mem_after = lv[MAYBE_IN_MEM_AFTER], mem_after_filter = lv[MEM_AFTER_FILTER],
      initialize_aux = lv[INITIALIZE_AUX], preinitialized_segments = lv[PREINITIALIZED_SEGMENTS],
      preinitialized_segments_aux = lv[PREINITIALIZED_SEGMENTS_AUX];
    P next_timestamp = nv[TIMESTAMP], next_is_read = nv[IS_READ], next_addr_context = nv[ADDR_CONTEXT],
      next_addr_segment = nv[ADDR_SEGMENT], next_addr_virtual = nv[ADDR_VIRTUAL];

    // The filter must be 0 or 1.
    P filter = lv[FILTER];
    yc.constraint(filter * (filter - one));
    // Dummy rows must be reads.
    P is_dummy = one - filter;
    P is_write = one - lv[IS_READ];
    yc.constraint(is_dummy * is_write);

    P context_first_change = lv[CONTEXT_FIRST_CHANGE], segment_first_change = lv[SEGMENT_FIRST_CHANGE],
      virtual_first_change = lv[VIRTUAL_FIRST_CHANGE];
    P address_unchanged = one - context_first_change - segment_first_change - virtual_first_change;
    P range_check = lv[RANGE_CHECK];
    P not_context_first_change = one - context_first_change;
    P not_segment_first_change = one - segment_first_change;
    P not_virtual_first_change = one - virtual_first_change;
    P not_address_unchanged = one - address_unchanged;

    // First set of ordering constraints: first_change flags are boolean.
    yc.constraint(context_first_change * not_context_first_change);
    yc.constraint(segment_first_change * not_segment_first_change);
    yc.constraint(virtual_first_change * not_virtual_first_change);
    yc.constraint(address_unchanged * not_address_unchanged);

    // Second set: no change before the column corresponding to the nonzero first_change flag.
    yc.constraint_transition(segment_first_change * (next_addr_context - addr_context));
    yc.constraint_transition(virtual_first_change * (next_addr_context - addr_context));
    yc.constraint_transition(virtual_first_change * (next_addr_segment - addr_segment));
    yc.constraint_transition(address_unchanged * (next_addr_context - addr_context));
    yc.constraint_transition(address_unchanged * (next_addr_segment - addr_segment));
    yc.constraint_transition(address_unchanged * (next_addr_virtual - addr_virtual));

    // Third set: range-check the difference in the column that should be increasing.
    P computed_range_check = context_first_change * (next_addr_context - addr_context - one) +
                             segment_first_change * (next_addr_segment - addr_segment - one) +
                             virtual_first_change * (next_addr_virtual - addr_virtual - one) +
                             address_unchanged * (next_timestamp - timestamp);
    yc.constraint_transition(range_check - computed_range_check);

    // Validate `preinitialized_segments_aux`.
    yc.constraint_transition(preinitialized_segments_aux -
                             (next_addr_segment - P::from_u64(SEG_ACCOUNTS_LINKED_LIST)) *
                                 (next_addr_segment - P::from_u64(SEG_STORAGE_LINKED_LIST)));
    // Validate `preinitialized_segments`.
    yc.constraint_transition(preinitialized_segments -
                             (next_addr_segment - P::from_u64(SEG_CODE)) *
                                 (next_addr_segment - P::from_u64(SEG_TRIE_DATA)) * preinitialized_segments_aux);
    // Validate `initialize_aux`.
    yc.constraint_transition(initialize_aux - preinitialized_segments * not_address_unchanged * next_is_read);

    for (uint32_t i = 0; i < 8; i++) {
        // Enumerate purportedly-ordered log.
        yc.constraint_transition(next_is_read * address_unchanged * (nv[VALUE_LIMBS0 + i] - lv[VALUE_LIMBS0 + i]));
        // Zero-initialisation of everything but the preinitialized segments.
        yc.constraint_transition(initialize_aux * nv[VALUE_LIMBS0 + i]);
    }

    // Validate `maybe_in_mem_after`.
    yc.constraint_transition(maybe_in_mem_after + filter * not_address_unchanged * (is_stale - one));
    // `mem_after_filter` must be binary.
    yc.constraint(mem_after_filter * (mem_after_filter - one));
    for (uint32_t i = 0; i < 8; i++)
        yc.constraint((mem_after_filter - maybe_in_mem_after) * preinitialized_segments * lv[VALUE_LIMBS0 + i]);

    // Validate timestamp_inv.
    yc.constraint(timestamp * (timestamp * timestamp_inv - one));

    // Range column: first value 0, increments by 1.
    P rc1 = lv[COUNTER], rc2 = nv[COUNTER];
    yc.constraint_first_row(rc1);
    P incr = rc2 - rc1;
    yc.constraint_transition(incr - one);
}

inline std::vector<Column> ctl_data() {
    std::vector<Column> res = Column::singles({IS_READ, ADDR_CONTEXT, ADDR_SEGMENT, ADDR_VIRTUAL});
    for (uint32_t i = 0; i < 8; i++) res.push_back(Column::single(VALUE_LIMBS0 + i));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline Filter ctl_filter() { return Filter::new_simple(Column::single(FILTER)); }
inline std::vector<Column> ctl_looking_mem() {
    std::vector<Column> res = Column::singles({ADDR_CONTEXT, ADDR_SEGMENT, ADDR_VIRTUAL});
    for (uint32_t i = 0; i < 8; i++) res.push_back(Column::single(VALUE_LIMBS0 + i));
    return res;
}
inline TableWithColumns ctl_context_pruning_looking() {
    return TableWithColumns(6, {Column::linear_combination_with_constant({{STALE_CONTEXTS, 1}}, GL_MOD - 1)},
                            Filter({}, {Column::single(IS_PRUNED)}));
}
// 1 - timestamp * timestamp_inv
inline Filter ctl_filter_mem_before() {
    return Filter({{Column::single(TIMESTAMP), Column::linear_combination({{TIMESTAMP_INV, GL_MOD - 1}})}},
                  {Column::one()});
}
inline Filter ctl_filter_mem_after() { return Filter::new_simple(Column::single(MEM_AFTER_FILTER)); }

inline std::vector<Lookup> lookups() {
    Lookup a;
    a.columns = {Column::single(RANGE_CHECK), Column::single_next_row(ADDR_VIRTUAL)};
    a.table_column = Column::single(COUNTER);
    a.frequencies_column = Column::single(FREQUENCIES);
    a.filter_columns = {Filter(), Filter::new_simple(Column::sum({CONTEXT_FIRST_CHANGE, SEGMENT_FIRST_CHANGE}))};
    Lookup b;
    b.columns = {Column::linear_combination_with_constant({{ADDR_CONTEXT, 1}}, 1)};
    b.table_column = Column::single(STALE_CONTEXTS);
    b.frequencies_column = Column::single(STALE_CONTEXT_FREQUENCIES);
    b.filter_columns = {Filter::new_simple(Column::single(IS_STALE))};
    return {a, b};
}

}}  // namespace zkstark::memory
