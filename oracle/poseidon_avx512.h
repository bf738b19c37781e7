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
// MDS layer on the 32-bit halves of the state: every matrix entry is < 2^6, so the 13 products of an output sum to < 2^42 per half
ORC_AVX512 void v8_mds(v8 s[12]) {
    const v8 eps = v8_set1(EPS);
    v8 lo[12], hi[12], out[12];
    for (int i = 0; i < 12; i++) { lo[i] = _mm512_and_si512(s[i], eps); hi[i] = _mm512_srli_epi64(s[i], 32); }
    for (int r = 0; r < 12; r++) {
        v8 al = _mm512_setzero_si512(), ah = _mm512_setzero_si512();
        for (int i = 0; i < 12; i++) {
            const v8 c = v8_set1(MDS_CIRC[i]);
            al = _mm512_add_epi64(al, _mm512_mul_epu32(lo[(i + r) % 12], c));
            ah = _mm512_add_epi64(ah, _mm512_mul_epu32(hi[(i + r) % 12], c));
        }
        if (r == 0) {
            const v8 d = v8_set1(ZK_POSEIDON_MDS_DIAG0);
            al = _mm512_add_epi64(al, _mm512_mul_epu32(lo[0], d));
            ah = _mm512_add_epi64(ah, _mm512_mul_epu32(hi[0], d));
        }
        // al + 2^32 ah as a 96-bit number: low word and the carries above it
        v8 sh = _mm512_slli_epi64(ah, 32);
        v8 low = _mm512_add_epi64(al, sh);
        __mmask8 c = _mm512_cmplt_epu64_mask(low, sh);
        v8 top = _mm512_srli_epi64(ah, 32);
        top = _mm512_mask_add_epi64(top, c, top, v8_set1(1));
        out[r] = v8_reduce128(low, top);
    }
    for (int r = 0; r < 12; r++) s[r] = out[r];
}
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
