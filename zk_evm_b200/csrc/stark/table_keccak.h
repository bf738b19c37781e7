// KeccakStark (one Keccak-f[1600] round per row, 24 rows per permutation).
// Source: /root/reference/evm_arithmetization/src/keccak/{columns.rs:6-134, keccak_stark.rs:38-62 (CTL), 266-426 (constraints),
// round_flags.rs:14-56, logic.rs:21-66, constants.rs (RC)}.
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark { namespace keccak {

static const uint32_t NUM_ROUNDS = 24, NUM_INPUTS = 25;
ZKS_HD uint32_t reg_step(uint32_t i) { return i; }
static const uint32_t TIMESTAMP = NUM_ROUNDS;
static const uint32_t START_A = TIMESTAMP + 1;
ZKS_HD uint32_t reg_a(uint32_t x, uint32_t y) { return START_A + (x * 5 + y) * 2; }
static const uint32_t START_C = START_A + 5 * 5 * 2;
ZKS_HD uint32_t reg_c(uint32_t x, uint32_t z) { return START_C + x * 64 + z; }
static const uint32_t START_C_PRIME = START_C + 5 * 64;
ZKS_HD uint32_t reg_c_prime(uint32_t x, uint32_t z) { return START_C_PRIME + x * 64 + z; }
static const uint32_t START_A_PRIME = START_C_PRIME + 5 * 64;
ZKS_HD uint32_t reg_a_prime(uint32_t x, uint32_t y, uint32_t z) { return START_A_PRIME + x * 64 * 5 + y * 64 + z; }
ZKS_HD uint32_t reg_b(uint32_t x, uint32_t y, uint32_t z) {
    // B[x, y] = ROT(A'[a, b], r[a, b]) with a = (x + 3y) % 5, b = x
    const uint8_t R[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
    uint32_t a = (x + 3 * y) % 5, b = x;
    uint32_t rot = R[a][b];
    return reg_a_prime(a, b, (z + 64 - rot) % 64);
}
static const uint32_t START_A_PRIME_PRIME = START_A_PRIME + 5 * 5 * 64;
ZKS_HD uint32_t reg_a_prime_prime(uint32_t x, uint32_t y) { return START_A_PRIME_PRIME + x * 2 * 5 + y * 2; }
static const uint32_t START_A_PRIME_PRIME_0_0_BITS = START_A_PRIME_PRIME + 5 * 5 * 2;
ZKS_HD uint32_t reg_a_prime_prime_0_0_bit(uint32_t i) { return START_A_PRIME_PRIME_0_0_BITS + i; }
static const uint32_t REG_A_PRIME_PRIME_PRIME_0_0_LO = START_A_PRIME_PRIME_0_0_BITS + 64;
static const uint32_t REG_A_PRIME_PRIME_PRIME_0_0_HI = REG_A_PRIME_PRIME_PRIME_0_0_LO + 1;
ZKS_HD uint32_t reg_a_prime_prime_prime(uint32_t x, uint32_t y) {
    return (x == 0 && y == 0) ? REG_A_PRIME_PRIME_PRIME_0_0_LO : reg_a_prime_prime(x, y);
}
static const uint32_t NUM_COLUMNS = REG_A_PRIME_PRIME_PRIME_0_0_HI + 1;   // 2431
ZKS_HD uint32_t reg_output_limb(uint32_t i) { uint32_t w = i / 2, y = w / 5, x = w % 5; return reg_a_prime_prime_prime(x, y) + (i % 2); }
ZKS_HD uint32_t reg_input_limb(uint32_t i) { uint32_t w = i / 2, y = w / 5, x = w % 5; return reg_a(x, y) + (i % 2); }

ZKS_HD uint64_t round_constant(uint32_t r) {
    const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
                             0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                             0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
                             0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                             0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
                             0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    return RC[r];
}

template <class P> ZKS_HD P xor_gen(P x, P y) { return x + y - x * (y + y); }
template <class P> ZKS_HD P xor3_gen(P x, P y, P z) { return xor_gen<P>(x, xor_gen<P>(y, z)); }
template <class P> ZKS_HD P andn_gen(P x, P y) { return (P::one() - x) * y; }

template <class P, class V, class CC>
ZKS_HD void eval_head(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    // ---- eval_round_flags ----
    for (uint32_t i = 0; i < NUM_ROUNDS; i++) { P f = lv[reg_step(i)]; yc.constraint(f * (f - one)); }
    P local_any_flag = P::zero(), next_any_flag = P::zero();
    for (uint32_t i = 0; i < NUM_ROUNDS; i++) { local_any_flag = local_any_flag + lv[reg_step(i)]; next_any_flag = next_any_flag + nv[reg_step(i)]; }
    yc.constraint_first_row(local_any_flag * (lv[reg_step(0)] - one));
    for (uint32_t i = 1; i < NUM_ROUNDS; i++) yc.constraint_first_row(local_any_flag * lv[reg_step(i)]);
    P current_any_flag = local_any_flag;
    P last_round_flag = lv[reg_step(NUM_ROUNDS - 1)];
    P padding_constraint = (next_any_flag - one) * current_any_flag * (last_round_flag - one);
    for (uint32_t i = 0; i < NUM_ROUNDS; i++) {
        P current_round_flag = lv[reg_step(i)];
        P next_round_flag = nv[reg_step((i + 1) % NUM_ROUNDS)];
        yc.constraint_transition(next_any_flag * (next_round_flag - current_round_flag) + padding_constraint);
    }
    yc.constraint_transition(next_any_flag * (current_any_flag - one));

    // ---- keccak_stark.rs:276-426 ----
    P not_final_step = one - lv[reg_step(NUM_ROUNDS - 1)];
    P sum_round_flags = local_any_flag;
    yc.constraint(sum_round_flags * not_final_step * (nv[TIMESTAMP] - lv[TIMESTAMP]));
}

// the four big constraint families (740 constraints), in the reference's emission order
template <class P, class V, class CC>
ZKS_HD void eval_middle(const V& lv, const V& nv, CC& yc) {
    (void)nv;
    // C'[x, z] = xor(C[x, z], C[x - 1, z], C[x + 1, z - 1])
    ZKS_NOUNROLL for (uint32_t x = 0; x < 5; x++)
        ZKS_NOUNROLL for (uint32_t z0 = 0; z0 < 64; z0 += 8) {
            P c0[8], c1[8], c2[8], cp[8];
            ZKS_UNROLL for (uint32_t k = 0; k < 8; k++) {
                uint32_t z = z0 + k;
                c0[k] = lv[reg_c(x, z)]; c1[k] = lv[reg_c((x + 4) % 5, z)]; c2[k] = lv[reg_c((x + 1) % 5, (z + 63) % 64)]; cp[k] = lv[reg_c_prime(x, z)];
            }
            ZKS_UNROLL for (uint32_t k = 0; k < 8; k++) yc.constraint(cp[k] - xor3_gen<P>(c0[k], c1[k], c2[k]));
        }
    // A[x, y, z] = xor(A'[x, y, z], C[x, z], C'[x, z])
    ZKS_NOUNROLL for (uint32_t x = 0; x < 5; x++)
        ZKS_NOUNROLL for (uint32_t y = 0; y < 5; y++) {
            P a_lo = lv[reg_a(x, y)], a_hi = lv[reg_a(x, y) + 1];
            P computed_lo = P::zero(), computed_hi = P::zero();
            // bits from the top down, eight at a time: the 24 column loads of a batch are issued before any of them is used
            ZKS_NOUNROLL for (uint32_t z0 = 64; z0 > 0; z0 -= 8) {
                P ap[8], cc[8], cp[8];
                ZKS_UNROLL for (uint32_t k = 0; k < 8; k++) {
                    uint32_t z = z0 - 1 - k;
                    ap[k] = lv[reg_a_prime(x, y, z)]; cc[k] = lv[reg_c(x, z)]; cp[k] = lv[reg_c_prime(x, z)];
                }
                P acc = z0 > 32 ? computed_hi : computed_lo;
                ZKS_UNROLL for (uint32_t k = 0; k < 8; k++) acc = acc + acc + xor3_gen<P>(ap[k], cc[k], cp[k]);
                if (z0 > 32) computed_hi = acc; else computed_lo = acc;
            }
            yc.constraint(computed_lo - a_lo);
            yc.constraint(computed_hi - a_hi);
        }
    // xor_{i<5} A'[x, i, z] = C'[x, z]: diff (diff - 2) (diff - 4) = 0
    ZKS_NOUNROLL for (uint32_t x = 0; x < 5; x++)
        ZKS_NOUNROLL for (uint32_t z0 = 0; z0 < 64; z0 += 4) {
            P ap[4][5], cp[4];
            ZKS_UNROLL for (uint32_t k = 0; k < 4; k++) {
                ZKS_UNROLL for (uint32_t i = 0; i < 5; i++) ap[k][i] = lv[reg_a_prime(x, i, z0 + k)];
                cp[k] = lv[reg_c_prime(x, z0 + k)];
            }
            ZKS_UNROLL for (uint32_t k = 0; k < 4; k++) {
                P sum = P::zero();
                ZKS_UNROLL for (uint32_t i = 0; i < 5; i++) sum = sum + ap[k][i];
                P diff = sum - cp[k];
                yc.constraint(diff * (diff - P::from_u64(2)) * (diff - P::from_u64(4)));
            }
        }
    // A''[x, y] = xor(B[x, y], andn(B[x + 1, y], B[x + 2, y]))
    ZKS_NOUNROLL for (uint32_t x = 0; x < 5; x++)
        ZKS_NOUNROLL for (uint32_t y = 0; y < 5; y++) {
            P lo = lv[reg_a_prime_prime(x, y)], hi = lv[reg_a_prime_prime(x, y) + 1];
            P computed_lo = P::zero(), computed_hi = P::zero();
            ZKS_NOUNROLL for (uint32_t z0 = 64; z0 > 0; z0 -= 8) {
                P b0[8], b1[8], b2[8];
                ZKS_UNROLL for (uint32_t k = 0; k < 8; k++) {
                    uint32_t z = z0 - 1 - k;
                    b0[k] = lv[reg_b(x, y, z)]; b1[k] = lv[reg_b((x + 1) % 5, y, z)]; b2[k] = lv[reg_b((x + 2) % 5, y, z)];
                }
                P acc = z0 > 32 ? computed_hi : computed_lo;
                ZKS_UNROLL for (uint32_t k = 0; k < 8; k++) acc = acc + acc + xor_gen<P>(b0[k], andn_gen<P>(b1[k], b2[k]));
                if (z0 > 32) computed_hi = acc; else computed_lo = acc;
            }
            yc.constraint(computed_lo - lo);
            yc.constraint(computed_hi - hi);
        }
}

// The same 740 constraints, emitted by index (Consumer::block_put) so that the columns are walked once per phase instead of once
// per family: the sequential form reads the 2240 columns of C, C' and A' 12800 times per point (28 GB of DRAM traffic for the
// 5 GB LDE of a 2^17-row table, profiles/r1h), this one 4480 times.  Index map = position in eval_middle's emission order:
//   [0, 320)    C' (x, z)          -> 64 x + z
//   [320, 370)  A  (x, y) lo, hi   -> 320 + 2 (5 x + y) + {0, 1}
//   [370, 690)  xor5 (x, z)        -> 370 + 64 x + z
//   [690, 740)  A'' (x, y) lo, hi  -> 690 + 2 (5 x + y) + {0, 1}
static const uint32_t MIDDLE_CONSTRAINTS = 740;
template <class P, class V, class CC>
ZKS_HD void eval_middle_blocked(const V& lv, const V& nv, CC& yc) {
    (void)nv;
    yc.block_begin(MIDDLE_CONSTRAINTS);
    const P two = P::from_u64(2), four = P::from_u64(4);
    // phase 1, per x: one walk over z (top bit first) feeds C'[x, z], the five A[x, y] words and the xor5 check
    ZKS_NOUNROLL for (uint32_t x = 0; x < 5; x++) {
        P acc[5];
        ZKS_UNROLL for (uint32_t y = 0; y < 5; y++) acc[y] = P::zero();
        const uint32_t xm = (x + 4) % 5, xp = (x + 1) % 5;
        ZKS_NOUNROLL for (uint32_t z0 = 64; z0 > 0; z0 -= 4) {
            ZKS_SYNC();   // the warps of a block walk the loop body together (instruction-cache sharing, see quotient_kernel.cuh)
            P c0[4], c1[4], c2[4], cp[4], ap[4][5];
            ZKS_UNROLL for (uint32_t k = 0; k < 4; k++) {
                const uint32_t z = z0 - 1 - k;
                c0[k] = lv[reg_c(x, z)]; c1[k] = lv[reg_c(xm, z)]; c2[k] = lv[reg_c(xp, (z + 63) % 64)]; cp[k] = lv[reg_c_prime(x, z)];
                ZKS_UNROLL for (uint32_t y = 0; y < 5; y++) ap[k][y] = lv[reg_a_prime(x, y, z)];
            }
            ZKS_UNROLL for (uint32_t k = 0; k < 4; k++) {
                const uint32_t z = z0 - 1 - k;
                yc.block_put(64 * x + z, cp[k] - xor3_gen<P>(c0[k], c1[k], c2[k]));
                const P ccp = xor_gen<P>(c0[k], cp[k]);          // xor3(A', C, C') = xor(A', xor(C, C'))
                P sum = P::zero();
                ZKS_UNROLL for (uint32_t y = 0; y < 5; y++) {
                    acc[y] = acc[y] + acc[y] + xor_gen<P>(ap[k][y], ccp);
                    sum = sum + ap[k][y];
                }
                const P diff = sum - cp[k];
                yc.block_put(370 + 64 * x + z, diff * (diff - two) * (diff - four));
            }
            if (z0 == 36 || z0 == 4) {                            // bits 63..32 (hi) resp. 31..0 (lo) are complete
                const uint32_t h = z0 == 36 ? 1 : 0;
                ZKS_UNROLL for (uint32_t y = 0; y < 5; y++) {
                    yc.block_put(320 + 2 * (5 * x + y) + h, acc[y] - lv[reg_a(x, y) + h]);
                    acc[y] = P::zero();
                }
            }
        }
    }
    // phase 2, per y: the five B[., y, z] feed the five A''[x, y] words
    ZKS_NOUNROLL for (uint32_t y = 0; y < 5; y++) {
        uint32_t base[5], rot[5];
        ZKS_UNROLL for (uint32_t x = 0; x < 5; x++) { base[x] = reg_b(x, y, 0); rot[x] = (base[x] - START_A_PRIME) & 63; base[x] -= rot[x]; }
        P acc[5];
        ZKS_UNROLL for (uint32_t x = 0; x < 5; x++) acc[x] = P::zero();
        ZKS_NOUNROLL for (uint32_t z0 = 64; z0 > 0; z0 -= 4) {
            ZKS_SYNC();
            P b[4][5];
            ZKS_UNROLL for (uint32_t k = 0; k < 4; k++) {
                const uint32_t z = z0 - 1 - k;
                ZKS_UNROLL for (uint32_t x = 0; x < 5; x++) b[k][x] = lv[base[x] + ((z + rot[x]) & 63)];   // == reg_b(x, y, z)
            }
            ZKS_UNROLL for (uint32_t k = 0; k < 4; k++)
                ZKS_UNROLL for (uint32_t x = 0; x < 5; x++)
                    acc[x] = acc[x] + acc[x] + xor_gen<P>(b[k][x], andn_gen<P>(b[k][(x + 1) % 5], b[k][(x + 2) % 5]));
            if (z0 == 36 || z0 == 4) {
                const uint32_t h = z0 == 36 ? 1 : 0;
                ZKS_UNROLL for (uint32_t x = 0; x < 5; x++) {
                    yc.block_put(690 + 2 * (5 * x + y) + h, acc[x] - lv[reg_a_prime_prime(x, y) + h]);
                    acc[x] = P::zero();
                }
            }
        }
    }
}

template <class P, class V, class CC>
ZKS_HD void eval_tail(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    // A'''[0, 0] = A''[0, 0] XOR RC
    {
        P c_lo = P::zero(), c_hi = P::zero();
        for (uint32_t z = 32; z-- > 0;) c_lo = c_lo + c_lo + lv[reg_a_prime_prime_0_0_bit(z)];
        for (uint32_t z = 64; z-- > 32;) c_hi = c_hi + c_hi + lv[reg_a_prime_prime_0_0_bit(z)];
        yc.constraint(c_lo - lv[reg_a_prime_prime(0, 0)]);
        yc.constraint(c_hi - lv[reg_a_prime_prime(0, 0) + 1]);
        P x_lo = P::zero(), x_hi = P::zero();
        for (uint32_t z = 32; z-- > 0;) {
            P rc_bit = P::zero();
            for (uint32_t r = 0; r < NUM_ROUNDS; r++) if ((round_constant(r) >> z) & 1) rc_bit = rc_bit + lv[reg_step(r)];
            x_lo = x_lo + x_lo + xor_gen<P>(lv[reg_a_prime_prime_0_0_bit(z)], rc_bit);
        }
        for (uint32_t z = 64; z-- > 32;) {
            P rc_bit = P::zero();
            for (uint32_t r = 0; r < NUM_ROUNDS; r++) if ((round_constant(r) >> z) & 1) rc_bit = rc_bit + lv[reg_step(r)];
            x_hi = x_hi + x_hi + xor_gen<P>(lv[reg_a_prime_prime_0_0_bit(z)], rc_bit);
        }
        yc.constraint(x_lo - lv[reg_a_prime_prime_prime(0, 0)]);
        yc.constraint(x_hi - lv[reg_a_prime_prime_prime(0, 0) + 1]);
    }
    // this round's output equals the next round's input
    ZKS_NOUNROLL for (uint32_t x = 0; x < 5; x++)
        ZKS_NOUNROLL for (uint32_t y = 0; y < 5; y++) {
            P output_lo = lv[reg_a_prime_prime_prime(x, y)], output_hi = lv[reg_a_prime_prime_prime(x, y) + 1];
            P input_lo = nv[reg_a(x, y)], input_hi = nv[reg_a(x, y) + 1];
            P not_last_round = one - lv[reg_step(NUM_ROUNDS - 1)];
            yc.constraint_transition(not_last_round * (output_lo - input_lo));
            yc.constraint_transition(not_last_round * (output_hi - input_hi));
        }
}

// reference emission order (oracle prover + verifier)
template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& nv, CC& yc) {
    eval_head<P>(lv, nv, yc);
    eval_middle<P>(lv, nv, yc);
    eval_tail<P>(lv, nv, yc);
}
// same value, columns walked once per phase (device quotient kernel)
template <class P, class V, class CC>
ZKS_HD void eval_blocked(const V& lv, const V& nv, CC& yc) {
    eval_head<P>(lv, nv, yc);
    eval_middle_blocked<P>(lv, nv, yc);
    eval_tail<P>(lv, nv, yc);
}

inline std::vector<Column> ctl_data_inputs() {
    std::vector<Column> res;
    for (uint32_t i = 0; i < 2 * NUM_INPUTS; i++) res.push_back(Column::single(reg_input_limb(i)));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline std::vector<Column> ctl_data_outputs() {
    std::vector<Column> res;
    for (uint32_t i = 0; i < 2 * NUM_INPUTS; i++) res.push_back(Column::single(reg_output_limb(i)));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline Filter ctl_filter_inputs() { return Filter::new_simple(Column::single(reg_step(0))); }
inline Filter ctl_filter_outputs() { return Filter::new_simple(Column::single(reg_step(NUM_ROUNDS - 1))); }
inline std::vector<Lookup> lookups() { return {}; }

}}  // namespace zkstark::keccak
