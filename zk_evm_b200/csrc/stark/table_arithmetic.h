// ArithmeticStark.
// Source: /root/reference/evm_arithmetization/src/arithmetic/{columns.rs:5-119, arithmetic_stark.rs:37-115 (CTL), 203-252 (dispatcher),
// 318-327 (lookup), utils.rs (limb-polynomial helpers), mul.rs:124-190, addcy.rs:67-166, divmod.rs:86-147, modular.rs:382-612,
// byte.rs:201-300, shift.rs:85-127}.
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace arithmetic {

static const uint32_t LIMB_BITS = 16, N_LIMBS = 16;
enum : uint32_t {
    IS_ADD = 0, IS_MUL = 1, IS_SUB = 2, IS_DIV = 3, IS_MOD = 4, IS_ADDMOD = 5, IS_MULMOD = 6, IS_ADDFP254 = 7, IS_MULFP254 = 8,
    IS_SUBFP254 = 9, IS_SUBMOD = 10, IS_LT = 11, IS_GT = 12, IS_BYTE = 13, IS_SHL = 14, IS_SHR = 15, IS_RANGE_CHECK = 16,
    OPCODE_COL = 17, START_SHARED_COLS = 18, NUM_SHARED_COLS = 96,
    INPUT_REGISTER_0 = 18, INPUT_REGISTER_1 = 34, INPUT_REGISTER_2 = 50, OUTPUT_REGISTER = 66,
    AUX_INPUT_REGISTER_0 = 82, AUX_INPUT_REGISTER_1 = 98, AUX_INPUT_REGISTER_DBL = 82,
    AUX_REGISTER_0 = 18,       // 16
    AUX_REGISTER_1 = 34,       // 32
    AUX_REGISTER_2 = 66,       // 31
    MUL_AUX_INPUT_LO = 82, MUL_AUX_INPUT_HI = 98,
    MODULAR_INPUT_0 = 18, MODULAR_INPUT_1 = 34, MODULAR_MODULUS = 50, MODULAR_OUTPUT = 66, MODULAR_QUO_INPUT = 82,
    MODULAR_OUT_AUX_RED = 18, MODULAR_MOD_IS_ZERO = 34,
    MODULAR_AUX_INPUT_LO = 35,   // 31 columns
    MODULAR_AUX_INPUT_HI = 66,   // 31 columns
    MODULAR_DIV_DENOM_IS_ZERO = 97,
    RANGE_COUNTER = 114, RC_FREQUENCIES = 115, NUM_COLUMNS = 116
};
static const uint64_t AUX_COEFF_ABS_MAX = 1ULL << 20;
static const uint64_t RANGE_MAX = 1ULL << 16;
static const uint64_t GOLDILOCKS_INVERSE_65536 = 18446462594437939201ULL;   // addcy.rs:67

// pol_adjoin_root: (x - root) * a(x), truncated to N coefficients
template <class P, int N> ZKS_HD void pol_adjoin_root(const P (&a)[N], P root, P (&res)[N]) {
    res[0] = P::zero() - root * a[0];
    for (int d = 1; d < N; d++) res[d] = a[d - 1] - root * a[d];
}

// mul.rs:124-154
template <class P, class V, class CC>
ZKS_HD void eval_mul(const V& lv, P filter, uint32_t left_reg, uint32_t right_reg, CC& yc) {
    const P base = P::from_u64(1ULL << LIMB_BITS), offset = P::from_u64(AUX_COEFF_ABS_MAX);
    P aux[N_LIMBS], adj[N_LIMBS];
    for (uint32_t i = 0; i < N_LIMBS; i++) aux[i] = lv[MUL_AUX_INPUT_LO + i] + (lv[MUL_AUX_INPUT_HI + i] * base - offset);
    pol_adjoin_root<P, (int)N_LIMBS>(aux, base, adj);
    for (uint32_t deg = 0; deg < N_LIMBS; deg++) {
        // pol_mul_lo
        P c = P::zero();
        for (uint32_t i = 0; i <= deg; i++) c = c + lv[left_reg + i] * lv[right_reg + deg - i];
        c = c - lv[OUTPUT_REGISTER + deg];
        c = c - adj[deg];
        yc.constraint(filter * c);
    }
}

// addcy.rs:69-133; x, y, z, given_cy are register getters (so that rows of lv / nv and synthetic values can be mixed)
template <class P, class FX, class FY, class FZ, class FC, class CC>
ZKS_HD void eval_addcy(CC& yc, P filter, FX x, FY y, FZ z, FC given_cy, bool is_two_row_op) {
    const P overflow = P::from_u64(1ULL << LIMB_BITS), overflow_inv = P::from_u64(GOLDILOCKS_INVERSE_65536);
    P cy = P::zero();
    for (uint32_t i = 0; i < N_LIMBS; i++) {
        P t = cy + x(i) + y(i) - z(i);
        if (is_two_row_op) yc.constraint_transition(filter * t * (overflow - t));
        else yc.constraint(filter * t * (overflow - t));
        cy = t * overflow_inv;
    }
    if (is_two_row_op) {
        yc.constraint_transition(filter * (cy - given_cy(0)));
        for (uint32_t i = 1; i < N_LIMBS; i++) yc.constraint_transition(filter * given_cy(i));
    } else {
        yc.constraint(filter * given_cy(0) * (given_cy(0) - P::one()));
        yc.constraint(filter * (cy - given_cy(0)));
        for (uint32_t i = 1; i < N_LIMBS; i++) yc.constraint(filter * given_cy(i));
    }
}
template <class P, class V> struct Reg { const V* v; uint32_t base; ZKS_HD P operator()(uint32_t i) const { return (*v)[base + i]; } };

// modular.rs:426-497 modular_constr_poly.  output / modulus: 16 limbs, quot: 32 limbs (all by value: they are patched locally).
template <class P, class V, class CC>
ZKS_HD void modular_constr_poly(const V& lv, const V& nv, CC& yc, P filter, P (&output)[N_LIMBS], P (&modulus)[N_LIMBS],
                                const P (&quot)[2 * N_LIMBS], P (&constr_poly)[2 * N_LIMBS]) {
    const P one = P::one();
    P mod_is_zero = nv[MODULAR_MOD_IS_ZERO];
    yc.constraint_transition(filter * (mod_is_zero * mod_is_zero - mod_is_zero));
    P limb_sum = P::zero();
    for (uint32_t i = 0; i < N_LIMBS; i++) limb_sum = limb_sum + modulus[i];
    yc.constraint_transition(filter * limb_sum * mod_is_zero);
    modulus[0] = modulus[0] + mod_is_zero;
    P div_denom_is_zero = nv[MODULAR_DIV_DENOM_IS_ZERO];
    yc.constraint_transition(filter * (mod_is_zero * (lv[IS_DIV] + lv[IS_SHR]) - div_denom_is_zero));
    output[0] = output[0] + div_denom_is_zero;
    {   // check_reduced (modular.rs:382-412): modulus + out_aux_red == output + is_less_than * 2^256
        P is_less_than0 = one - mod_is_zero * (lv[IS_DIV] + lv[IS_SHR]);
        auto fx = [&](uint32_t i) { return modulus[i]; };
        auto fy = [&](uint32_t i) { return nv[MODULAR_OUT_AUX_RED + i]; };
        auto fz = [&](uint32_t i) { return output[i]; };
        auto fc = [&](uint32_t i) { return i == 0 ? is_less_than0 : P::zero(); };
        eval_addcy<P>(yc, filter, fx, fy, fz, fc, true);
    }
    output[0] = output[0] - div_denom_is_zero;
    // prod = q(x) * m(x): 47 coefficients, the top 15 must vanish
    for (uint32_t k = 2 * N_LIMBS; k < 3 * N_LIMBS - 1; k++) {
        P x = P::zero();
        for (uint32_t j = 0; j < N_LIMBS; j++) if (k >= j && k - j < 2 * N_LIMBS) x = x + quot[k - j] * modulus[j];
        yc.constraint_transition(filter * x);
    }
    for (uint32_t k = 0; k < 2 * N_LIMBS; k++) {
        P x = P::zero();
        for (uint32_t j = 0; j < N_LIMBS && j <= k; j++) x = x + quot[k - j] * modulus[j];
        constr_poly[k] = x;
    }
    for (uint32_t i = 0; i < N_LIMBS; i++) constr_poly[i] = constr_poly[i] + output[i];
    const P base = P::from_u64(1ULL << LIMB_BITS), offset = P::from_u64(AUX_COEFF_ABS_MAX);
    P aux[2 * N_LIMBS], adj[2 * N_LIMBS];
    for (uint32_t i = 0; i < 2 * N_LIMBS; i++) aux[i] = P::zero();
    for (uint32_t i = 0; i < 2 * N_LIMBS - 1; i++) aux[i] = nv[MODULAR_AUX_INPUT_LO + i] - offset;
    for (uint32_t i = 0; i < 2 * N_LIMBS - 1; i++) aux[i] = aux[i] + base * nv[MODULAR_AUX_INPUT_HI + i];
    pol_adjoin_root<P, (int)(2 * N_LIMBS)>(aux, base, adj);
    for (uint32_t i = 0; i < 2 * N_LIMBS; i++) constr_poly[i] = constr_poly[i] + adj[i];
}

// divmod.rs:86-118
template <class P, class V, class CC>
ZKS_HD void eval_divmod_helper(const V& lv, const V& nv, CC& yc, P filter, uint32_t num, uint32_t den, uint32_t quo, uint32_t rem) {
    yc.constraint_last_row(filter);
    P output[N_LIMBS], modulus[N_LIMBS], quot[2 * N_LIMBS], cp[2 * N_LIMBS];
    for (uint32_t i = 0; i < N_LIMBS; i++) { output[i] = lv[rem + i]; modulus[i] = lv[den + i]; quot[i] = lv[quo + i]; quot[N_LIMBS + i] = P::zero(); }
    modular_constr_poly<P>(lv, nv, yc, filter, output, modulus, quot, cp);
    for (uint32_t i = 0; i < N_LIMBS; i++) cp[i] = cp[i] - lv[num + i];
    for (uint32_t i = 0; i < 2 * N_LIMBS; i++) yc.constraint_transition(filter * cp[i]);
}

// modular.rs:539-612
template <class P, class V, class CC>
ZKS_HD void eval_modular(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    const uint64_t BN254[N_LIMBS] = {0xfd47, 0xd87c, 0x8c16, 0x3c20, 0xca8d, 0x6871, 0x6a91, 0x9781,
                                     0x585d, 0x8181, 0x45b6, 0xb850, 0xa029, 0xe131, 0x4e72, 0x3064};
    P bn254_filter = lv[IS_ADDFP254] + lv[IS_MULFP254] + lv[IS_SUBFP254];
    P filter = lv[IS_ADDMOD] + lv[IS_SUBMOD] + lv[IS_MULMOD] + bn254_filter;
    yc.constraint_last_row(filter);
    for (uint32_t i = 0; i < N_LIMBS; i++) yc.constraint_transition(bn254_filter * (lv[MODULAR_MODULUS + i] - P::from_u64(BN254[i])));
    P add_filter = lv[IS_ADDMOD] + lv[IS_ADDFP254];
    P sub_filter = lv[IS_SUBMOD] + lv[IS_SUBFP254];
    P mul_filter = lv[IS_MULMOD] + lv[IS_MULFP254];
    P addmul_filter = add_filter + mul_filter;

    P output[N_LIMBS], modulus[N_LIMBS], quot[2 * N_LIMBS];
    P sub_cp[2 * N_LIMBS], mod_cp[2 * N_LIMBS];
    {   // submod_constr_poly (modular.rs:512-536)
        for (uint32_t i = 0; i < N_LIMBS; i++) { output[i] = lv[MODULAR_OUTPUT + i]; modulus[i] = lv[MODULAR_MODULUS + i]; }
        for (uint32_t i = 0; i < 2 * N_LIMBS; i++) quot[i] = lv[MODULAR_QUO_INPUT + i];
        P sign = quot[N_LIMBS];
        yc.constraint(sub_filter * sign * (sign - one));
        P offset = P::from_u64(0xFFFF);
        for (uint32_t i = 0; i < N_LIMBS; i++) quot[i] = quot[i] - offset * sign;
        quot[N_LIMBS] = P::zero();
        for (uint32_t i = N_LIMBS; i < 2 * N_LIMBS; i++) yc.constraint(sub_filter * quot[i]);
        modular_constr_poly<P>(lv, nv, yc, sub_filter, output, modulus, quot, sub_cp);
    }
    ZKS_SYNC();
    {
        for (uint32_t i = 0; i < N_LIMBS; i++) { output[i] = lv[MODULAR_OUTPUT + i]; modulus[i] = lv[MODULAR_MODULUS + i]; }
        for (uint32_t i = 0; i < 2 * N_LIMBS; i++) quot[i] = lv[MODULAR_QUO_INPUT + i];
        modular_constr_poly<P>(lv, nv, yc, addmul_filter, output, modulus, quot, mod_cp);
    }
    ZKS_SYNC();
    // add: constr_poly - (a + b)
    for (uint32_t k = 0; k < 2 * N_LIMBS; k++) {
        P c = mod_cp[k];
        if (k < N_LIMBS) c = c - (lv[MODULAR_INPUT_0 + k] + lv[MODULAR_INPUT_1 + k]);
        yc.constraint_transition(add_filter * c);
    }
    // sub: submod constr_poly - (a - b)
    for (uint32_t k = 0; k < 2 * N_LIMBS; k++) {
        P c = sub_cp[k];
        if (k < N_LIMBS) c = c - (lv[MODULAR_INPUT_0 + k] - lv[MODULAR_INPUT_1 + k]);
        yc.constraint_transition(sub_filter * c);
    }
    // mul: constr_poly - a * b (31 coefficients)
    for (uint32_t k = 0; k < 2 * N_LIMBS; k++) {
        P c = mod_cp[k];
        if (k < 2 * N_LIMBS - 1) {
            P m = P::zero();
            for (uint32_t i = 0; i < N_LIMBS; i++) if (k >= i && k - i < N_LIMBS) m = m + lv[MODULAR_INPUT_0 + i] * lv[MODULAR_INPUT_1 + k - i];
            c = c - m;
        }
        yc.constraint_transition(mul_filter * c);
    }
}

// byte.rs:201-300
template <class P, class V, class CC>
ZKS_HD void eval_byte(const V& lv, CC& yc) {
    const P one = P::one();
    const uint32_t idx = INPUT_REGISTER_0, val = INPUT_REGISTER_1, out = OUTPUT_REGISTER, idx_decomp = AUX_INPUT_REGISTER_0,
                   tree = AUX_INPUT_REGISTER_1;
    const uint32_t BYTE_IDX_DECOMP_HI = AUX_INPUT_REGISTER_0 + 5, BYTE_LAST_LIMB_LO = AUX_INPUT_REGISTER_0 + 6,
                   BYTE_LAST_LIMB_HI = AUX_INPUT_REGISTER_0 + 7, BYTE_IDX_IS_LARGE = AUX_INPUT_REGISTER_0 + 8,
                   BYTE_IDX_HI_LIMB_SUM_INV_0 = AUX_INPUT_REGISTER_0 + 9;
    P is_byte = lv[IS_BYTE];
    P idx0_lo5 = P::zero();
    for (uint32_t i = 0; i < 5; i++) {
        P bit = lv[idx_decomp + i];
        yc.constraint(is_byte * (bit * bit - bit));
        idx0_lo5 = idx0_lo5 + bit * P::from_u64(1ULL << i);
    }
    P idx0_hi = lv[idx_decomp + 5] * P::from_u64(32);
    yc.constraint(is_byte * (lv[idx] - (idx0_lo5 + idx0_hi)));
    P bit = lv[idx_decomp + 4];
    for (uint32_t i = 0; i < 8; i++) {
        P limb = bit * lv[val + i] + (one - bit) * lv[val + i + 8];
        yc.constraint(is_byte * (lv[tree + i] - limb));
    }
    bit = lv[idx_decomp + 3];
    for (uint32_t i = 0; i < 4; i++) {
        P limb = bit * lv[tree + i] + (one - bit) * lv[tree + i + 4];
        yc.constraint(is_byte * (lv[tree + i + 8] - limb));
    }
    bit = lv[idx_decomp + 2];
    for (uint32_t i = 0; i < 2; i++) {
        P limb = bit * lv[tree + i + 8] + (one - bit) * lv[tree + i + 10];
        yc.constraint(is_byte * (lv[tree + i + 12] - limb));
    }
    bit = lv[idx_decomp + 1];
    P limb = bit * lv[tree + 12] + (one - bit) * lv[tree + 13];
    yc.constraint(is_byte * (lv[tree + 14] - limb));
    const P base8 = P::from_u64(256);
    P lo_byte = lv[BYTE_LAST_LIMB_LO], hi_byte = lv[BYTE_LAST_LIMB_HI];
    yc.constraint(is_byte * (lo_byte + base8 * (base8 * hi_byte - limb)));
    bit = lv[idx_decomp];
    P t = bit * lo_byte + (one - bit) * base8 * hi_byte;
    yc.constraint(is_byte * (base8 * lv[tree + 15] - t));
    P expected_out_byte = lv[tree + 15];
    P hi_limb_sum = lv[BYTE_IDX_DECOMP_HI];
    for (uint32_t i = 1; i < N_LIMBS; i++) hi_limb_sum = hi_limb_sum + lv[idx + i];
    P idx_is_large = lv[BYTE_IDX_IS_LARGE];
    yc.constraint(is_byte * (idx_is_large * idx_is_large - idx_is_large));
    yc.constraint(is_byte * hi_limb_sum * (idx_is_large - one));
    P hi_limb_sum_inv = lv[BYTE_IDX_HI_LIMB_SUM_INV_0] + lv[BYTE_IDX_HI_LIMB_SUM_INV_0 + 1] * P::from_u64(1ULL << 16) +
                        lv[BYTE_IDX_HI_LIMB_SUM_INV_0 + 2] * P::from_u64(1ULL << 32) + lv[BYTE_IDX_HI_LIMB_SUM_INV_0 + 3] * P::from_u64(1ULL << 48);
    yc.constraint(is_byte * (hi_limb_sum * hi_limb_sum_inv - idx_is_large));
    P out_byte = lv[out];
    P check = out_byte - (one - idx_is_large) * expected_out_byte;
    yc.constraint(is_byte * check);
    for (uint32_t i = 1; i < N_LIMBS; i++) yc.constraint(is_byte * lv[out + i]);
}

// arithmetic_stark.rs:203-252
template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    for (uint32_t i = IS_ADD; i <= IS_RANGE_CHECK; i++) { P flag = lv[i]; yc.constraint(flag * (flag - one)); }
    P all_flags = P::zero();
    for (uint32_t i = IS_ADD; i <= IS_RANGE_CHECK; i++) all_flags = all_flags + lv[i];
    yc.constraint(all_flags * (all_flags - one));
    yc.constraint((one - lv[IS_RANGE_CHECK]) * lv[OPCODE_COL]);
    P rc1 = lv[RANGE_COUNTER], rc2 = nv[RANGE_COUNTER];
    yc.constraint_first_row(rc1);
    P incr = rc2 - rc1;
    yc.constraint_transition(incr * incr - incr);
    yc.constraint_last_row(rc1 - P::from_u64(RANGE_MAX - 1));

    // ZKS_SYNC: the warps of the (one) block of an SM walk this megabyte of straight-line code together and share the fetched lines
    ZKS_SYNC();
    // MUL
    eval_mul<P>(lv, lv[IS_MUL], INPUT_REGISTER_0, INPUT_REGISTER_1, yc);
    ZKS_SYNC();
    // ADD, SUB, LT, GT (addcy.rs:135-153): x + y = z + cy * 2^256
    {
        Reg<P, V> in0{&lv, INPUT_REGISTER_0}, in1{&lv, INPUT_REGISTER_1}, out{&lv, OUTPUT_REGISTER}, aux{&lv, AUX_INPUT_REGISTER_0};
        eval_addcy<P>(yc, lv[IS_ADD], in0, in1, out, aux, false);
        eval_addcy<P>(yc, lv[IS_SUB], in1, out, in0, aux, false);
        eval_addcy<P>(yc, lv[IS_LT], in1, aux, in0, out, false);
        eval_addcy<P>(yc, lv[IS_GT], in0, aux, in1, out, false);
    }
    ZKS_SYNC();
    // DIV, MOD
    eval_divmod_helper<P>(lv, nv, yc, lv[IS_DIV], INPUT_REGISTER_0, INPUT_REGISTER_1, OUTPUT_REGISTER, AUX_INPUT_REGISTER_0);
    ZKS_SYNC();
    eval_divmod_helper<P>(lv, nv, yc, lv[IS_MOD], INPUT_REGISTER_0, INPUT_REGISTER_1, AUX_INPUT_REGISTER_0, OUTPUT_REGISTER);
    ZKS_SYNC();
    // ADDMOD, SUBMOD, MULMOD and the FP254 variants
    eval_modular<P>(lv, nv, yc);
    ZKS_SYNC();
    // BYTE
    eval_byte<P>(lv, yc);
    ZKS_SYNC();
    // SHL (== MUL on registers 1, 2), SHR (== DIV on registers 1, 2)
    eval_mul<P>(lv, lv[IS_SHL], INPUT_REGISTER_1, INPUT_REGISTER_2, yc);
    ZKS_SYNC();
    eval_divmod_helper<P>(lv, nv, yc, lv[IS_SHR], INPUT_REGISTER_1, INPUT_REGISTER_2, OUTPUT_REGISTER, AUX_INPUT_REGISTER_0);
}

// arithmetic_stark.rs:37-115
inline TableWithColumns ctl_arithmetic_rows() {
    const std::vector<std::pair<uint32_t, uint64_t>> COMBINED_OPS = {
        {IS_ADD, 0x01}, {IS_MUL, 0x02}, {IS_SUB, 0x03}, {IS_DIV, 0x04}, {IS_MOD, 0x06}, {IS_ADDMOD, 0x08}, {IS_MULMOD, 0x09},
        {IS_ADDFP254, 0x0c}, {IS_MULFP254, 0x0d}, {IS_SUBFP254, 0x0e}, {IS_SUBMOD, 0x0f}, {IS_LT, 0x10}, {IS_GT, 0x11},
        {IS_BYTE, 0x1a}, {IS_SHL, 0x1b}, {IS_SHR, 0x1c}};
    std::vector<uint32_t> filter_cols;
    for (auto& p : COMBINED_OPS) filter_cols.push_back(p.first);
    filter_cols.push_back(IS_RANGE_CHECK);
    std::vector<std::pair<uint32_t, uint64_t>> all_combined = COMBINED_OPS;
    all_combined.push_back({OPCODE_COL, 0x01});
    std::vector<Column> cols = {Column::linear_combination(all_combined)};
    const uint32_t REGS[4] = {INPUT_REGISTER_0, INPUT_REGISTER_1, INPUT_REGISTER_2, OUTPUT_REGISTER};
    for (uint32_t r : REGS)
        for (uint32_t i = 0; i < N_LIMBS / 2; i++) cols.push_back(Column::linear_combination({{r + 2 * i, 1}, {r + 2 * i + 1, 1ULL << LIMB_BITS}}));
    return TableWithColumns(0, cols, Filter::new_simple(Column::sum(filter_cols)));
}
inline std::vector<Lookup> lookups() {
    Lookup l;
    for (uint32_t i = 0; i < NUM_SHARED_COLS; i++) { l.columns.push_back(Column::single(START_SHARED_COLS + i)); l.filter_columns.push_back(Filter()); }
    l.table_column = Column::single(RANGE_COUNTER);
    l.frequencies_column = Column::single(RC_FREQUENCIES);
    return {l};
}

}}  // namespace zkstark::arithmetic
