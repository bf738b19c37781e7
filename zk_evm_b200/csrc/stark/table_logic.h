// LogicStark.
// Source: /root/reference/evm_arithmetization/src/logic.rs:30-71 (columns), 84-113 (CTL), 249-303 (constraints).
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace logic {

enum : uint32_t { IS_AND = 0, IS_OR = 1, IS_XOR = 2, INPUT0 = 3, INPUT1 = 3 + 256, RESULT = 3 + 512, NUM_COLUMNS = 3 + 512 + 8 };
static const uint32_t VAL_BITS = 256, PACKED_LIMB_BITS = 32, PACKED_LEN = 8;

template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& /*nv*/, CC& yc) {
    const P one = P::one();
    P is_and = lv[IS_AND], is_or = lv[IS_OR], is_xor = lv[IS_XOR];
    // Flags must be boolean.
    yc.constraint(is_and * (is_and - one));
    yc.constraint(is_or * (is_or - one));
    yc.constraint(is_xor * (is_xor - one));
    // Only a single flag must be activated at once.
    P all_flags = is_and + is_or + is_xor;
    yc.constraint(all_flags * (all_flags - one));
    // in0 OP in1 = sum_coeff * (in0 + in1) + and_coeff * (in0 AND in1)
    P sum_coeff = is_or + is_xor;
    P and_coeff = is_and - is_or - is_xor * P::from_u64(2);
    // All bits are bits.
    for (uint32_t i = 0; i < VAL_BITS; i++) { P bit = lv[INPUT0 + i]; yc.constraint(bit * (bit - one)); }
    for (uint32_t i = 0; i < VAL_BITS; i++) { P bit = lv[INPUT1 + i]; yc.constraint(bit * (bit - one)); }
    // Form the result.
    for (uint32_t l = 0; l < PACKED_LEN; l++) {
        P x = P::zero(), y = P::zero(), x_land_y = P::zero();
        for (uint32_t i = 0; i < PACKED_LIMB_BITS; i++) {
            P xb = lv[INPUT0 + l * PACKED_LIMB_BITS + i], yb = lv[INPUT1 + l * PACKED_LIMB_BITS + i];
            P w = P::from_u64(1ULL << i);
            x = x + xb * w;
            y = y + yb * w;
            x_land_y = x_land_y + xb * yb * w;
        }
        P x_op_y = sum_coeff * (x + y) + and_coeff * x_land_y;
        yc.constraint(lv[RESULT + l] - x_op_y);
    }
}

inline std::vector<Column> ctl_data() {
    std::vector<Column> res = {Column::linear_combination({{IS_AND, 0x16}, {IS_OR, 0x17}, {IS_XOR, 0x18}})};
    for (uint32_t base : {(uint32_t)INPUT0, (uint32_t)INPUT1})
        for (uint32_t l = 0; l < PACKED_LEN; l++) {
            std::vector<uint32_t> bits;
            for (uint32_t i = 0; i < PACKED_LIMB_BITS; i++) bits.push_back(base + l * PACKED_LIMB_BITS + i);
            res.push_back(Column::le_bits(bits));
        }
    for (uint32_t l = 0; l < PACKED_LEN; l++) res.push_back(Column::single(RESULT + l));
    return res;
}
inline Filter ctl_filter() { return Filter::new_simple(Column::sum({IS_AND, IS_OR, IS_XOR})); }
inline std::vector<Lookup> lookups() { return {}; }

}}  // namespace zkstark::logic
