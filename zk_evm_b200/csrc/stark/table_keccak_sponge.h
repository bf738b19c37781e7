// KeccakSpongeStark.
// Source: /root/reference/evm_arithmetization/src/keccak_sponge/{columns.rs:9-64, keccak_sponge_stark.rs:34-250 (CTL), 546-715
// (constraints), 946-953 (lookup)}.
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace keccak_sponge {

static const uint32_t KECCAK_WIDTH_BYTES = 200, KECCAK_DIGEST_BYTES = 32, KECCAK_RATE_BYTES = 136, KECCAK_RATE_U32S = 34,
                      KECCAK_CAPACITY_U32S = 16, KECCAK_DIGEST_U32S = 8, KECCAK_WIDTH_MINUS_DIGEST_U32S = 42;
enum : uint32_t {
    IS_FULL_INPUT_BLOCK = 0, CONTEXT = 1, SEGMENT = 2, VIRT = 3, TIMESTAMP = 4, ALREADY_ABSORBED_BYTES = 5,
    IS_PADDING_BYTE = 6,                 // 136
    ORIGINAL_RATE_U32S = 142,            // 34
    ORIGINAL_CAPACITY_U32S = 176,        // 16
    BLOCK_BYTES = 192,                   // 136
    XORED_RATE_U32S = 328,               // 34
    PARTIAL_UPDATED_STATE_U32S = 362,    // 42
    UPDATED_DIGEST_STATE_BYTES = 404,    // 32
    RANGE_COUNTER = 436, RC_FREQUENCIES = 437, NUM_COLUMNS = 438
};
static const uint64_t BYTE_RANGE_MAX = 256;

template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    P rc1 = lv[RANGE_COUNTER], rc2 = nv[RANGE_COUNTER];
    yc.constraint_first_row(rc1);
    P incr = rc2 - rc1;
    yc.constraint_transition(incr * incr - incr);
    yc.constraint_last_row(rc1 - P::from_u64(BYTE_RANGE_MAX - 1));

    P is_full_input_block = lv[IS_FULL_INPUT_BLOCK];
    yc.constraint(is_full_input_block * (is_full_input_block - one));
    for (uint32_t i = 0; i < KECCAK_RATE_BYTES; i++) { P b = lv[IS_PADDING_BYTE + i]; yc.constraint(b * (b - one)); }
    P is_final_block = lv[IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1];
    for (uint32_t i = 1; i < KECCAK_RATE_BYTES; i++) yc.constraint(lv[IS_PADDING_BYTE + i - 1] * (lv[IS_PADDING_BYTE + i] - one));
    yc.constraint(is_final_block * is_full_input_block);

    P already_absorbed_bytes = lv[ALREADY_ABSORBED_BYTES];
    yc.constraint_first_row(already_absorbed_bytes);
    for (uint32_t i = 0; i < KECCAK_RATE_U32S; i++) yc.constraint_first_row(lv[ORIGINAL_RATE_U32S + i]);
    for (uint32_t i = 0; i < KECCAK_CAPACITY_U32S; i++) yc.constraint_first_row(lv[ORIGINAL_CAPACITY_U32S + i]);

    yc.constraint_transition(is_final_block * nv[ALREADY_ABSORBED_BYTES]);
    for (uint32_t i = 0; i < KECCAK_RATE_U32S; i++) yc.constraint_transition(is_final_block * nv[ORIGINAL_RATE_U32S + i]);
    for (uint32_t i = 0; i < KECCAK_CAPACITY_U32S; i++) yc.constraint_transition(is_final_block * nv[ORIGINAL_CAPACITY_U32S + i]);

    yc.constraint_transition(is_full_input_block * (lv[CONTEXT] - nv[CONTEXT]));
    yc.constraint_transition(is_full_input_block * (lv[SEGMENT] - nv[SEGMENT]));
    yc.constraint_transition(is_full_input_block * (lv[VIRT] - nv[VIRT]));
    yc.constraint_transition(is_full_input_block * (lv[TIMESTAMP] - nv[TIMESTAMP]));

    // the next row's "before" state matches our "after" state
    for (uint32_t k = 0; k < KECCAK_DIGEST_U32S; k++) {
        P current_after = lv[UPDATED_DIGEST_STATE_BYTES + 4 * k];
        for (uint32_t i = 1; i < 4; i++) current_after = current_after + lv[UPDATED_DIGEST_STATE_BYTES + 4 * k + i] * P::from_u64(1ULL << (8 * i));
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_RATE_U32S + k] - current_after));
    }
    for (uint32_t k = 0; k < KECCAK_RATE_U32S - KECCAK_DIGEST_U32S; k++)
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_RATE_U32S + KECCAK_DIGEST_U32S + k] - lv[PARTIAL_UPDATED_STATE_U32S + k]));
    for (uint32_t k = 0; k < KECCAK_CAPACITY_U32S; k++)
        yc.constraint_transition(is_full_input_block *
                                 (nv[ORIGINAL_CAPACITY_U32S + k] - lv[PARTIAL_UPDATED_STATE_U32S + (KECCAK_RATE_U32S - KECCAK_DIGEST_U32S) + k]));

    yc.constraint_transition(is_full_input_block * (already_absorbed_bytes + P::from_u64(KECCAK_RATE_BYTES) - nv[ALREADY_ABSORBED_BYTES]));

    P has_single_padding_byte = lv[IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1] - lv[IS_PADDING_BYTE + KECCAK_RATE_BYTES - 2];
    yc.constraint_transition(has_single_padding_byte * (lv[BLOCK_BYTES + KECCAK_RATE_BYTES - 1] - P::from_u64(0x81)));
    for (uint32_t i = 0; i + 1 < KECCAK_RATE_BYTES; i++) {
        P is_first_padding_byte = i > 0 ? lv[IS_PADDING_BYTE + i] - lv[IS_PADDING_BYTE + i - 1] : lv[IS_PADDING_BYTE + i];
        yc.constraint_transition(is_first_padding_byte * (lv[BLOCK_BYTES + i] - one));
        yc.constraint_transition(lv[IS_PADDING_BYTE + i] * (is_first_padding_byte - one) * lv[BLOCK_BYTES + i]);
    }
    yc.constraint_transition(is_final_block * (has_single_padding_byte - one) * (lv[BLOCK_BYTES + KECCAK_RATE_BYTES - 1] - P::from_u64(0x80)));

    P is_dummy = one - is_full_input_block - is_final_block;
    P next_is_final_block = nv[IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1];
    yc.constraint_transition(is_dummy * (nv[IS_FULL_INPUT_BLOCK] + next_is_final_block));
}

inline std::vector<Column> ctl_looked_data() {
    std::vector<Column> outputs;
    for (uint32_t i = 8; i-- > 0;) {
        std::vector<std::pair<uint32_t, uint64_t>> t;
        for (uint32_t j = 0; j < 4; j++) t.push_back({UPDATED_DIGEST_STATE_BYTES + 4 * i + j, 1ULL << (24 - 8 * j)});
        outputs.push_back(Column::linear_combination(t));
    }
    std::vector<std::pair<uint32_t, uint64_t>> len = {{ALREADY_ABSORBED_BYTES, 1}};
    for (uint32_t i = 0; i < KECCAK_RATE_BYTES; i++) len.push_back({IS_PADDING_BYTE + i, GL_MOD - 1});
    std::vector<Column> res = Column::singles({CONTEXT, SEGMENT, VIRT});
    res.push_back(Column::linear_combination_with_constant(len, KECCAK_RATE_BYTES));
    res.push_back(Column::single(TIMESTAMP));
    res.insert(res.end(), outputs.begin(), outputs.end());
    return res;
}
inline std::vector<Column> ctl_looking_keccak_inputs() {
    std::vector<Column> res;
    for (uint32_t i = 0; i < KECCAK_RATE_U32S; i++) res.push_back(Column::single(XORED_RATE_U32S + i));
    for (uint32_t i = 0; i < KECCAK_CAPACITY_U32S; i++) res.push_back(Column::single(ORIGINAL_CAPACITY_U32S + i));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline std::vector<Column> ctl_looking_keccak_outputs() {
    std::vector<Column> res;
    for (uint32_t k = 0; k < KECCAK_DIGEST_U32S; k++) {
        std::vector<std::pair<uint32_t, uint64_t>> t;
        for (uint32_t i = 0; i < 4; i++) t.push_back({UPDATED_DIGEST_STATE_BYTES + 4 * k + i, 1ULL << (8 * i)});
        res.push_back(Column::linear_combination(t));
    }
    for (uint32_t i = 0; i < KECCAK_WIDTH_MINUS_DIGEST_U32S; i++) res.push_back(Column::single(PARTIAL_UPDATED_STATE_U32S + i));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline std::vector<Column> ctl_looking_memory(uint32_t i) {
    std::vector<Column> res = {Column::constant_col(1), Column::single(CONTEXT), Column::single(SEGMENT)};
    res.push_back(Column::linear_combination_with_constant({{VIRT, 1}, {ALREADY_ABSORBED_BYTES, 1}}, i));
    res.push_back(Column::single(BLOCK_BYTES + i));
    for (uint32_t k = 1; k < 8; k++) res.push_back(Column::zero());
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline uint32_t num_logic_ctls() { return (KECCAK_RATE_BYTES + 31) / 32; }
inline std::vector<Column> ctl_looking_logic(uint32_t i) {
    const uint32_t U32S_PER_CTL = 8, U8S_PER_CTL = 32;
    std::vector<Column> res = {Column::constant_col(0x18)};   // is_xor
    for (uint32_t k = 0; k < U32S_PER_CTL; k++) {
        uint32_t idx = i * U32S_PER_CTL + k;
        res.push_back(idx < KECCAK_RATE_U32S ? Column::single(ORIGINAL_RATE_U32S + idx) : Column::zero());
    }
    for (uint32_t k = 0; k < U32S_PER_CTL; k++) {
        uint32_t b0 = i * U8S_PER_CTL + 4 * k;
        if (b0 < KECCAK_RATE_BYTES) {
            std::vector<uint32_t> cs;
            for (uint32_t j = 0; j < 4 && b0 + j < KECCAK_RATE_BYTES; j++) cs.push_back(BLOCK_BYTES + b0 + j);
            res.push_back(Column::le_bytes(cs));
        } else res.push_back(Column::zero());
    }
    for (uint32_t k = 0; k < U32S_PER_CTL; k++) {
        uint32_t idx = i * U32S_PER_CTL + k;
        res.push_back(idx < KECCAK_RATE_U32S ? Column::single(XORED_RATE_U32S + idx) : Column::zero());
    }
    return res;
}
inline Filter ctl_looked_filter() { return Filter::new_simple(Column::single(IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1)); }
inline Filter ctl_looking_memory_filter(uint32_t i) {
    if (i == KECCAK_RATE_BYTES - 1) return Filter::new_simple(Column::single(IS_FULL_INPUT_BLOCK));
    return Filter::new_simple(Column::linear_combination({{IS_FULL_INPUT_BLOCK, 1}, {IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1, 1},
                                                          {IS_PADDING_BYTE + i, GL_MOD - 1}}));
}
inline Filter ctl_looking_logic_filter() { return Filter::new_simple(Column::sum({IS_FULL_INPUT_BLOCK, IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1})); }
inline Filter ctl_looking_keccak_filter() { return Filter::new_simple(Column::sum({IS_FULL_INPUT_BLOCK, IS_PADDING_BYTE + KECCAK_RATE_BYTES - 1})); }
inline std::vector<Lookup> lookups() {
    Lookup l;
    for (uint32_t i = 0; i < KECCAK_RATE_BYTES; i++) { l.columns.push_back(Column::single(BLOCK_BYTES + i)); l.filter_columns.push_back(Filter()); }
    l.table_column = Column::single(RANGE_COUNTER);
    l.frequencies_column = Column::single(RC_FREQUENCIES);
    return {l};
}

}}  // namespace zkstark::keccak_sponge
