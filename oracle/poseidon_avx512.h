// ORACLE (test infrastructure, NOT product code): eight Poseidon permutations at a time in AVX-512 registers.
//
// The CPU baseline's hot spot is Poseidon (60 % of a segment proof with the scalar form).  plonky2 itself ships packed AVX-512
// Goldilocks arithmetic for exactly this reason (plonky2_field 1.0.0 arch/x86_64/avx512_goldilocks_field.rs); this is our own
// restatement of the same idea, kept deliberately plain: the same 30-round structure as orc::poseidon (oracle_core.h), canonical
// values between operations, one lane per independent permutation.  Checked against the scalar form on random states and through
// it against the reference's known answers (tests/test_oracle_kats.py); used only when the CPU has AVX-512F/DQ.
#pragma once
#include <immintrin.h>
#include "oracle_core.h"

namespace orc {

#define ORC_AVX512 __attribute__((target("avx512f,avx512dq"), always_inline)) static inline

typedef __m512i v8;
ORC_AVX512 v8 v8_set1(uint64_t x) { return _mm512_set1_epi64((long long)x); }
// canonical a, b -> canonical a + b
ORC_AVX512 v8 v8_add(v8 a, v8 b) {
    const v8 p = v8_set1(P);
    v8 s = _mm512_add_epi64(a, b);
    __mmask8 m = _mm512_cmplt_epu64_mask(s, a) | _mm512_cmpge_epu64_mask(s, p);
    return _mm512_mask_sub_epi64(s, m, s, p);
}
// lo + 2^64 hi -> canonical  (n0 + (2^32-1) n1 - n2, field.md:9-20)
ORC_AVX512 v8 v8_reduce128(v8 lo, v8 hi) {
    const v8 eps = v8_set1(EPS), p = v8_set1(P);
    v8 n2 = _mm512_srli_epi64(hi, 32), n1 = _mm512_and_si512(hi, eps);
    v8 t = _mm512_sub_epi64(lo, n2);
    __mmask8 b = _mm512_cmplt_epu64_mask(lo, n2);
    t = _mm512_mask_sub_epi64(t, b, t, eps);
    v8 u = _mm512_sub_epi64(_mm512_slli_epi64(n1, 32), n1);
    v8 r = _mm512_add_epi64(t, u);
    __mmask8 c = _mm512_cmplt_epu64_mask(r, u);
    r = _mm512_mask_add_epi64(r, c, r, eps);
    __mmask8 g = _mm512_cmpge_epu64_mask(r, p);
    return _mm512_mask_sub_epi64(r, g, r, p);
}
ORC_AVX512 v8 v8_mul(v8 a, v8 b) {
    const v8 eps = v8_set1(EPS);
    v8 ah = _mm512_srli_epi64(a, 32), bh = _mm512_srli_epi64(b, 32);
    v8 ll = _mm512_mul_epu32(a, b), lh = _mm512_mul_epu32(a, bh), hl = _mm512_mul_epu32(ah, b), hh = _mm512_mul_epu32(ah, bh);
    v8 mid = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));                  // <= (2^32-1)^2 + 2^32-1
    v8 mid2 = _mm512_add_epi64(hl, _mm512_and_si512(mid, eps));
    v8 lo = _mm512_or_si512(_mm512_and_si512(ll, eps), _mm512_slli_epi64(mid2, 32));
    v8 hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(mid, 32), _mm512_srli_epi64(mid2, 32)));
    return v8_reduce128(lo, hi);
}
ORC_AVX512 v8 v8_sbox7(v8 x) {
    v8 x2 = v8_mul(x, x), x4 = v8_mul(x2, x2), x3 = v8_mul(x, x2);
    return v8_mul(x3, x4);
}
// MDS layer without multiplications, the way plonky2's own CPU code does it (plonky2 1.0.0 hash/poseidon_goldilocks.rs
// `mds_multiply_freq`): the circulant part of the matrix is a cyclic convolution of length 12; splitting the index as j = b + 3a and
// taking a 4-point DFT over a (roots 1, i, -1, -i: additions only) leaves three 3x3 twisted convolutions whose kernels are powers of
// two for this matrix — DFT(circ)/4 = [16, 32, 16], [-1, -8, 2] and (2 + i, -4 - i, 16 - i) for the complex pair — so the layer is ~90
// additions and shifts per half.  Done on the 32-bit halves of the state in wrap-around 64-bit lanes (every true output is < 2^42,
// intermediates may be "negative"), then one 96-bit reduction per output.  Checked against orc::mds_layer through poseidon_x8 == poseidon.
#define V8ADD(a, b) _mm512_add_epi64(a, b)
#define V8SUB(a, b) _mm512_sub_epi64(a, b)
#define V8SHL(a, k) _mm512_slli_epi64(a, k)
ORC_AVX512 void v8_mds_freq_half(const v8 s[12], v8 o[12]) {
    v8 F1[3], Fm[3], Fc[3], Fd[3];
    for (int b = 0; b < 3; b++) {
        v8 x0 = s[b], x1 = s[b + 3], x2 = s[b + 6], x3 = s[b + 9];
        v8 A = V8ADD(x0, x2), B = V8ADD(x1, x3);
        Fc[b] = V8SUB(x0, x2); Fd[b] = V8SUB(x1, x3);
        F1[b] = V8ADD(A, B); Fm[b] = V8SUB(A, B);
    }
    // real block, kernel [16, 32, 16]:  G1_b = 16 (F1_0 + F1_1 + F1_2 + F1_(b+2))
    v8 T = V8ADD(V8ADD(F1[0], F1[1]), F1[2]);
    v8 G1[3] = {V8SHL(V8ADD(T, F1[2]), 4), V8SHL(V8ADD(T, F1[0]), 4), V8SHL(V8ADD(T, F1[1]), 4)};
    // alternating block, kernel [-1, -8, 2] under the twisted (sign-changing) convolution
    v8 Gm[3] = {V8SUB(V8SUB(V8SHL(Fm[2], 3), V8SHL(Fm[1], 1)), Fm[0]),
                V8SUB(V8SUB(V8SUB(_mm512_setzero_si512(), V8SHL(Fm[0], 3)), V8SHL(Fm[2], 1)), Fm[1]),
                V8SUB(V8SUB(V8SHL(Fm[0], 1), V8SHL(Fm[1], 3)), Fm[2])};
    // complex block, kernel k0 = 2 + i, k1 = -4 - i, k2 = 16 - i:  (c + d i) k0 = (2c - d) + (2d + c) i,  k1: (-4c + d) + (-4d - c) i,
    // k2: (16c + d) + (16d - c) i;   Gi_0 = k0 F0 + i (k1 F2 + k2 F1),  Gi_1 = k0 F1 + k1 F0 + i k2 F2,  Gi_2 = k0 F2 + k1 F1 + k2 F0
    v8 u[3], v[3];
    {
        v8 re = V8ADD(V8ADD(V8SUB(V8SHL(Fc[1], 4), V8SHL(Fc[2], 2)), Fd[2]), Fd[1]);
        v8 im = V8SUB(V8SUB(V8SUB(V8SHL(Fd[1], 4), V8SHL(Fd[2], 2)), Fc[2]), Fc[1]);
        u[0] = V8SUB(V8SUB(V8SHL(Fc[0], 1), Fd[0]), im);
        v[0] = V8ADD(V8ADD(V8SHL(Fd[0], 1), Fc[0]), re);
    }
    u[1] = V8ADD(V8SUB(V8ADD(V8SUB(V8SUB(V8SHL(Fc[1], 1), Fd[1]), V8SHL(Fc[0], 2)), Fd[0]), V8SHL(Fd[2], 4)), Fc[2]);
    v[1] = V8ADD(V8ADD(V8SUB(V8SUB(V8ADD(V8SHL(Fd[1], 1), Fc[1]), V8SHL(Fd[0], 2)), Fc[0]), V8SHL(Fc[2], 4)), Fd[2]);
    u[2] = V8ADD(V8ADD(V8ADD(V8SUB(V8SUB(V8SHL(Fc[2], 1), Fd[2]), V8SHL(Fc[1], 2)), Fd[1]), V8SHL(Fc[0], 4)), Fd[0]);
    v[2] = V8SUB(V8ADD(V8SUB(V8SUB(V8ADD(V8SHL(Fd[2], 1), Fc[2]), V8SHL(Fd[1], 2)), Fc[1]), V8SHL(Fd[0], 4)), Fc[0]);
    for (int b = 0; b < 3; b++) {
        v8 Pp = V8ADD(G1[b], Gm[b]), Q = V8SUB(G1[b], Gm[b]);
        o[b] = V8ADD(Pp, u[b]); o[b + 3] = V8ADD(Q, v[b]); o[b + 6] = V8SUB(Pp, u[b]); o[b + 9] = V8SUB(Q, v[b]);
    }
    o[0] = V8ADD(o[0], V8SHL(s[0], 3));          // + diag(8, 0, ..., 0)
}
ORC_AVX512 void v8_mds(v8 s[12]) {
    const v8 eps = v8_set1(EPS);
    v8 lo[12], hi[12], ol[12], oh[12];
    for (int i = 0; i < 12; i++) { lo[i] = _mm512_and_si512(s[i], eps); hi[i] = _mm512_srli_epi64(s[i], 32); }
    v8_mds_freq_half(lo, ol);
    v8_mds_freq_half(hi, oh);
    for (int r = 0; r < 12; r++) {
        // ol + 2^32 oh as a 96-bit number: low word and the carries above it
        v8 sh = _mm512_slli_epi64(oh[r], 32);
        v8 low = _mm512_add_epi64(ol[r], sh);
        __mmask8 c = _mm512_cmplt_epu64_mask(low, sh);
        v8 top = _mm512_srli_epi64(oh[r], 32);
        top = _mm512_mask_add_epi64(top, c, top, v8_set1(1));
        s[r] = v8_reduce128(low, top);
    }
}
#undef V8ADD
#undef V8SUB
#undef V8SHL
// st[lane of the state][permutation]: eight independent permutations
__attribute__((target("avx512f,avx512dq"))) static inline void poseidon_x8(uint64_t st[12][8]) {
    v8 s[12];
    for (int i = 0; i < 12; i++) s[i] = _mm512_loadu_si512((const void*)st[i]);
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = v8_add(s[i], v8_set1(POSEIDON_RC[12 * r + i]));
        if (r < 4 || r >= 26) { for (int i = 0; i < 12; i++) s[i] = v8_sbox7(s[i]); }
        else s[0] = v8_sbox7(s[0]);
        v8_mds(s);
    }
    for (int i = 0; i < 12; i++) _mm512_storeu_si512((void*)st[i], s[i]);
}

static inline bool have_avx512() {
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && !getenv("ORC_NO_AVX512");
    return ok;
}

// hash_or_noop of eight rows of the same length (row j at rows[j]); len > 4
static inline void hash_no_pad_x8(const uint64_t* const rows[8], size_t len, Hash out[8]) {
    alignas(64) uint64_t st[12][8];
    memset(st, 0, sizeof st);
    for (size_t i = 0; i < len; i += 8) {
        const size_t l = len - i < 8 ? len - i : 8;
        for (size_t j = 0; j < l; j++)
            for (int k = 0; k < 8; k++) st[j][k] = rows[k][i + j];
        poseidon_x8(st);
    }
    for (int k = 0; k < 8; k++) for (int j = 0; j < 4; j++) out[k].e[j] = st[j][k];
}
// eight two_to_one compressions: out[k] = H(children[2k], children[2k+1])
static inline void two_to_one_x8(const Hash* children, Hash* out) {
    alignas(64) uint64_t st[12][8];
    for (int k = 0; k < 8; k++) {
        for (int j = 0; j < 4; j++) { st[j][k] = children[2 * k].e[j]; st[4 + j][k] = children[2 * k + 1].e[j]; }
        for (int j = 8; j < 12; j++) st[j][k] = 0;
    }
    poseidon_x8(st);
    for (int k = 0; k < 8; k++) for (int j = 0; j < 4; j++) out[k].e[j] = st[j][k];
}

}  // namespace orc
