// KeccakStark trace rows, one row at a time (single source: the device trace-finishing kernel and its host test).
// Follows /root/reference/evm_arithmetization/src/keccak/keccak_stark.rs:70-250 (generate_trace_rows, generate_trace_rows_for_perm,
// copy_output_to_input, generate_trace_row_for_round); column map: keccak/columns.rs (table_keccak.h reg_*).
//
// The reference fills the 24 rows of a permutation one after the other (row r+1 copies its A from row r's A''').  A row depends on its
// permutation's input only through the 1600-bit state entering its round, so here EVERY ROW IS INDEPENDENT: row 24p + r advances the
// input of permutation p by r plain Keccak-f rounds (a few hundred word operations) and then expands round r into the 2431 cells.
// On the device that is one thread per row, so that the 32 threads of a warp store 32 consecutive rows of the same column (the trace
// is column-major): every store instruction writes 256 contiguous bytes.
#pragma once
#include "hd.h"
#include "table_keccak.h"

namespace zkstark { namespace keccak {

ZKS_HD uint64_t rotl64(uint64_t x, uint32_t r) { return r ? (x << r) | (x >> (64 - r)) : x; }
// rotation offsets r[x][y] (keccak_stark.rs reg_b / columns.rs:84-100)
ZKS_HD uint32_t rho_offset(uint32_t x, uint32_t y) {
    const uint8_t R[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
    return R[x][y];
}

// the words of one round; lane (x, y) of a 25-word array lives at index y*5 + x (the order of the permutation's input, keccak_stark.rs:108)
struct RoundWords {
    uint64_t C[5], Cp[5];       // C[x] = xor_y A[x,y];  C'[x] = C[x] ^ C[x-1] ^ rotl(C[x+1], 1)           (:135-160)
    uint64_t Ap[25];            // A'[x,y] = A[x,y] ^ C[x] ^ C'[x]                                          (:162-178)
    uint64_t App[25];           // A''[x,y] = B[x,y] ^ (~B[x+1,y] & B[x+2,y]),  B[x,y] = rotl(A'[a,b], r[a,b]), a = (x+3y)%5, b = x   (:180-205)
    uint64_t out00;             // A'''[0,0] = A''[0,0] ^ RC[round]                                         (:224-233)
};

ZKS_HD void round_words(const uint64_t* A, uint32_t round, RoundWords& w) {
    ZKS_UNROLL
    for (uint32_t x = 0; x < 5; x++) w.C[x] = A[x] ^ A[5 + x] ^ A[10 + x] ^ A[15 + x] ^ A[20 + x];
    ZKS_UNROLL
    for (uint32_t x = 0; x < 5; x++) w.Cp[x] = w.C[x] ^ w.C[(x + 4) % 5] ^ rotl64(w.C[(x + 1) % 5], 1);
    ZKS_UNROLL
    for (uint32_t y = 0; y < 5; y++) {
        ZKS_UNROLL
        for (uint32_t x = 0; x < 5; x++) w.Ap[y * 5 + x] = A[y * 5 + x] ^ w.C[x] ^ w.Cp[x];
    }
    ZKS_UNROLL
    for (uint32_t y = 0; y < 5; y++) {
        uint64_t B[5];
        ZKS_UNROLL
        for (uint32_t x = 0; x < 5; x++) {
            const uint32_t a = (x + 3 * y) % 5, b = x;
            B[x] = rotl64(w.Ap[b * 5 + a], rho_offset(a, b));
        }
        ZKS_UNROLL
        for (uint32_t x = 0; x < 5; x++) w.App[y * 5 + x] = B[x] ^ (~B[(x + 1) % 5] & B[(x + 2) % 5]);
    }
    w.out00 = w.App[0] ^ round_constant(round);
}

// A of the next round = A''' of this one (copy_output_to_input, :118-131)
ZKS_HD void next_round_input(const RoundWords& w, uint64_t* A) {
    ZKS_UNROLL
    for (uint32_t i = 0; i < 25; i++) A[i] = w.App[i];
    A[0] = w.out00;
}

// Row `round` of a permutation whose state ENTERING that round is A: st(column, value) for each of the 2431 cells, every value
// canonical (bits, 32-bit limbs, the timestamp as given: F::from_canonical_usize).
template <class Store>
ZKS_HD void emit_row(const uint64_t* A, uint64_t timestamp, uint32_t round, const RoundWords& w, Store& st) {
    ZKS_NOUNROLL
    for (uint32_t i = 0; i < NUM_ROUNDS; i++) st(reg_step(i), (uint64_t)(i == round));
    st(TIMESTAMP, timestamp);
    ZKS_UNROLL
    for (uint32_t x = 0; x < 5; x++) {
        ZKS_UNROLL
        for (uint32_t y = 0; y < 5; y++) {
            st(reg_a(x, y), A[y * 5 + x] & 0xFFFFFFFFull);
            st(reg_a(x, y) + 1, A[y * 5 + x] >> 32);
        }
    }
    ZKS_UNROLL
    for (uint32_t x = 0; x < 5; x++) {
        const uint64_t c = w.C[x], cp = w.Cp[x];
        ZKS_NOUNROLL
        for (uint32_t z = 0; z < 64; z++) st(reg_c(x, z), (c >> z) & 1);
        ZKS_NOUNROLL
        for (uint32_t z = 0; z < 64; z++) st(reg_c_prime(x, z), (cp >> z) & 1);
    }
    ZKS_UNROLL
    for (uint32_t x = 0; x < 5; x++) {
        ZKS_UNROLL
        for (uint32_t y = 0; y < 5; y++) {
            const uint64_t ap = w.Ap[y * 5 + x];
            ZKS_NOUNROLL
            for (uint32_t z = 0; z < 64; z++) st(reg_a_prime(x, y, z), (ap >> z) & 1);
        }
    }
    ZKS_UNROLL
    for (uint32_t x = 0; x < 5; x++) {
        ZKS_UNROLL
        for (uint32_t y = 0; y < 5; y++) {
            st(reg_a_prime_prime(x, y), w.App[y * 5 + x] & 0xFFFFFFFFull);
            st(reg_a_prime_prime(x, y) + 1, w.App[y * 5 + x] >> 32);
        }
    }
    const uint64_t app00 = w.App[0];
    ZKS_NOUNROLL
    for (uint32_t z = 0; z < 64; z++) st(reg_a_prime_prime_0_0_bit(z), (app00 >> z) & 1);
    st(REG_A_PRIME_PRIME_PRIME_0_0_LO, w.out00 & 0xFFFFFFFFull);
    st(REG_A_PRIME_PRIME_PRIME_0_0_HI, w.out00 >> 32);
}

// Row `row` of the trace of `num_perms` permutations (inputs: 25 words each, lane y*5 + x; timestamps: one per permutation): rows past
// 24 * num_perms are the all-zero padding rows (generate_trace_rows, :84-87).
template <class Store>
ZKS_HD void generate_row(const uint64_t* inputs, const uint64_t* timestamps, uint64_t num_perms, uint64_t row, Store& st) {
    const uint64_t p = row / NUM_ROUNDS;
    const uint32_t round = (uint32_t)(row % NUM_ROUNDS);
    if (p >= num_perms) {
        ZKS_NOUNROLL
        for (uint32_t c = 0; c < NUM_COLUMNS; c++) st(c, (uint64_t)0);
        return;
    }
    uint64_t A[25];
    ZKS_UNROLL
    for (uint32_t i = 0; i < 25; i++) A[i] = inputs[p * NUM_INPUTS + i];
    RoundWords w;
    ZKS_NOUNROLL
    for (uint32_t r = 0; r < round; r++) { round_words(A, r, w); next_round_input(w, A); }
    round_words(A, round, w);
    emit_row(A, timestamps[p], round, w, st);
}

// the permutation's output (what the KeccakSponge table looks up): the state after 24 rounds, lane y*5 + x
ZKS_HD void permutation_output(const uint64_t* input, uint64_t* out) {
    uint64_t A[25];
    for (uint32_t i = 0; i < 25; i++) A[i] = input[i];
    RoundWords w;
    for (uint32_t r = 0; r < NUM_ROUNDS; r++) { round_words(A, r, w); next_round_input(w, A); }
    for (uint32_t i = 0; i < 25; i++) out[i] = A[i];
}

}}  // namespace zkstark::keccak
