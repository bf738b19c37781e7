// MemoryStark trace finishing, one row at a time (single source: the device kernel and its host test).
// Follows /root/reference/evm_arithmetization/src/memory/memory_stark.rs: MemoryOp::into_row (:104-131, the timestamp inverse),
// generate_first_change_flags_and_rc (:134-213), insert_stale_contexts (:387-404), generate_trace_col_major (:240-294).
// Input: the operations AFTER the host's sort / fill_gaps / pad_memory_ops (:215-236: data-dependent insertions, they stay on the host),
// i.e. the 14 columns filter, timestamp, is_read, context, segment, virtual, value limbs of every row; output: the other 16 columns.
#pragma once
#include "hd.h"
#include "table_memory.h"
#include "../gl.cuh"

namespace zkstark { namespace memory {

// what a row contributes to the two histogram columns (applied by the caller: atomics on the device, plain adds on the host)
struct RowCounts {
    uint64_t freq_a;            // frequencies[range_check]                                   (:247-248)
    uint64_t freq_b;            // frequencies[next virtual] when the context or segment changes, NONE otherwise   (:249-259)
    uint64_t stale_ctx;         // stale_context_frequencies[context] when the row is stale, NONE otherwise        (:263-265)
};
static const uint64_t NONE = ~0ull;

// t: the column-major trace (column c at t + c*n) with the 14 operation columns, STALE_CONTEXTS and IS_PRUNED already in place.
// Writes the row's cells of the other columns except the three accumulated ones (FREQUENCIES, STALE_CONTEXT_FREQUENCIES are histograms).
// Returns false if a range-checked value does not fit the table (the reference asserts, :190-194).
ZKS_HD bool finish_row(uint64_t* t, uint64_t n, uint64_t i, RowCounts& rc) {
    using zk::gl_sub; using zk::gl_mul; using zk::gl_add;
    const uint64_t j = i + 1 == n ? 0 : i + 1;                       // the last row looks at the first one (:138-143)
    const uint64_t ctx = t[ADDR_CONTEXT * n + i], seg = t[ADDR_SEGMENT * n + i], virt = t[ADDR_VIRTUAL * n + i], ts = t[TIMESTAMP * n + i];
    const uint64_t nctx = t[ADDR_CONTEXT * n + j], nseg = t[ADDR_SEGMENT * n + j], nvirt = t[ADDR_VIRTUAL * n + j], nts = t[TIMESTAMP * n + j];
    const uint64_t next_is_read = t[IS_READ * n + j], filter = t[FILTER * n + i];
    t[TIMESTAMP_INV * n + i] = ts ? zk::gl_inv(ts) : 0;              // try_inverse().unwrap_or_default()
    const bool cfc = ctx != nctx;
    const bool sfc = seg != nseg && !cfc;
    const bool vfc = virt != nvirt && !sfc && !cfc;
    t[CONTEXT_FIRST_CHANGE * n + i] = cfc;
    t[SEGMENT_FIRST_CHANGE * n + i] = sfc;
    t[VIRTUAL_FIRST_CHANGE * n + i] = vfc;
    const uint64_t one = 1;
    const uint64_t range_check = i + 1 == n ? 0
                                 : cfc ? gl_sub(gl_sub(nctx, ctx), one)
                                 : sfc ? gl_sub(gl_sub(nseg, seg), one)
                                 : vfc ? gl_sub(gl_sub(nvirt, virt), one)
                                       : gl_sub(nts, ts);
    t[RANGE_CHECK * n + i] = range_check;
    const uint64_t aux = gl_mul(gl_sub(nseg, SEG_ACCOUNTS_LINKED_LIST), gl_sub(nseg, SEG_STORAGE_LINKED_LIST));
    const uint64_t pre = gl_mul(gl_mul(gl_sub(nseg, SEG_CODE), gl_sub(nseg, SEG_TRIE_DATA)), aux);
    t[PREINITIALIZED_SEGMENTS_AUX * n + i] = aux;
    t[PREINITIALIZED_SEGMENTS * n + i] = pre;
    const bool address_changed = cfc || sfc || vfc;                  // the three flags are exclusive: their sum is 0 or 1
    t[INITIALIZE_AUX * n + i] = address_changed ? gl_mul(pre, next_is_read) : 0;
    t[COUNTER * n + i] = i;
    // generate_trace_col_major
    rc.freq_a = range_check;
    rc.freq_b = (cfc || sfc) ? (i + 1 < n ? nvirt : 0) : NONE;
    bool ok = range_check < n && (rc.freq_b == NONE || rc.freq_b < n);
    // addr_ctx + 1 == stale_contexts[addr_ctx]: the context is in the stale list (the reference indexes the column with the context
    // number; a context past the end of the trace cannot be listed)
    const bool stale = ctx < n && t[STALE_CONTEXTS * n + ctx] == gl_add(ctx, one);
    uint64_t maybe = 0, after = 0;
    rc.stale_ctx = NONE;
    if (stale) {
        rc.stale_ctx = ctx;
    } else if (filter == 1 && address_changed) {
        maybe = 1;
        bool nonzero = false;
        for (uint32_t l = 0; l < 8; l++) nonzero |= t[(VALUE_LIMBS0 + l) * n + i] != 0;
        if (nonzero || seg == SEG_CODE || seg == SEG_TRIE_DATA || seg == SEG_ACCOUNTS_LINKED_LIST || seg == SEG_STORAGE_LINKED_LIST) after = 1;
    }
    t[IS_STALE * n + i] = stale;
    t[MAYBE_IN_MEM_AFTER * n + i] = maybe;
    t[MEM_AFTER_FILTER * n + i] = after;
    return ok;
}

}}  // namespace zkstark::memory
