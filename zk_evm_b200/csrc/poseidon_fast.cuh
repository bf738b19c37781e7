// Device-only fast path of the Poseidon permutation (same function as poseidon.cuh's poseidon_permute, which stays the
// host+device reference form).  plonky2 1.0.0 `Poseidon::poseidon` for GoldilocksField, width 12.
//
// What is different from the plain form:
//  * the state is kept NON-CANONICAL between operations: any u64 congruent to the value mod p.  Every reduction then
//    ends with carry-flag corrections (2^64 == 2^32-1 mod p) instead of compare/select chains; callers canonicalise
//    once at the very end (pf_canon).
//  * 64x64->128 products are 4 IMAD.WIDE.U32 + one carry add; the 128->64 reduction is 3 short carry chains.
//  * the round constants of round r+1 are the initial value of the MDS accumulators of round r (M(s)+c costs nothing
//    extra), so only the first round's constants are added explicitly.
//  * the dense MDS layer runs on the 32-bit halves of the state with 64-bit IMAD.WIDE accumulators (every matrix
//    entry is < 2^6), one 73-bit -> 64-bit fold per output.
#pragma once
#include "poseidon.cuh"

namespace zk {

// The file is device code; a plain C++ compiler sees the same functions with portable bodies where the device ones are inline PTX
// (tests/native/poseidon_fast_host.cpp runs the permutation's algorithm — lazy reductions, limb-domain MDS, round loop — on the CPU).
#if defined(__CUDACC__)
#define PF_FN __device__ __forceinline__
#else
#define PF_FN inline
#endif

struct U64 { uint32_t lo, hi; };

#if defined(__CUDA_ARCH__)
PF_FN uint64_t pf_pack(uint32_t lo, uint32_t hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
PF_FN void pf_unpack(uint64_t x, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}
PF_FN uint64_t pf_mulwide(uint32_t a, uint32_t b) {
    uint64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
PF_FN uint64_t pf_madwide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
#else
PF_FN uint64_t pf_pack(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
PF_FN void pf_unpack(uint64_t x, uint32_t& lo, uint32_t& hi) { lo = (uint32_t)x; hi = (uint32_t)(x >> 32); }
PF_FN uint64_t pf_mulwide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
PF_FN uint64_t pf_madwide(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }
PF_FN uint32_t pf_funnelshift_r(uint32_t lo, uint32_t hi, unsigned s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> s); }
PF_FN uint32_t pf_funnelshift_l(uint32_t lo, uint32_t hi, unsigned s) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << s) >> 32); }
#endif
#if defined(__CUDA_ARCH__)
#define pf_funnelshift_r __funnelshift_r
#define pf_funnelshift_l __funnelshift_l
#endif

// w0 + 2^32 w1 + 2^64 w2 + 2^96 w3  ->  some u64 congruent to it mod p
PF_FN uint64_t pf_reduce128(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
#if !defined(__CUDA_ARCH__)
    // the same steps with 64-bit words: t = (w1:w0) - w3 (wrap: -EPS), u = w2 EPS, r = t + u (wrap: +EPS)
    const uint64_t lo = pf_pack(w0, w1);
    uint64_t t = lo - w3;
    if (lo < w3) t -= GL_EPS;
    const uint64_t u = ((uint64_t)w2 << 32) - w2;
    uint64_t r = t + u;
    if (r < t) r += GL_EPS;
    return r;
#else
    uint32_t t0, t1, m, u0, u1;
    // t = (w1:w0) - w3 ; a borrow wrapped by 2^64 == EPS, so take EPS off again (cannot borrow twice)
    asm("sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, 0;\n\t"
        "subc.u32 %2, 0, 0;\n\t"          // m = 0xFFFFFFFF on borrow, else 0
        : "=r"(t0), "=r"(t1), "=r"(m) : "r"(w0), "r"(w1), "r"(w3));
    asm("sub.cc.u32 %0, %0, %2;\n\t"
        "subc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(m));
    // u = w2 * EPS = (w2 << 32) - w2
    asm("sub.cc.u32 %0, 0, %2;\n\t"
        "subc.u32 %1, %2, 0;\n\t" : "=r"(u0), "=r"(u1) : "r"(w2));
    // r = t + u ; a carry wrapped by 2^64 == EPS, add it back (cannot carry twice)
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, 0, 0;\n\t"          // m = carry (0/1)
        : "+r"(t0), "+r"(t1), "=r"(m) : "r"(u0), "r"(u1));
    m = 0u - m;
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(m));
    return pf_pack(t0, t1);
#endif
}

PF_FN uint64_t pf_mul(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1;
    pf_unpack(a, a0, a1); pf_unpack(b, b0, b1);
    uint64_t p00 = pf_mulwide(a0, b0);
    uint32_t w0, c0; pf_unpack(p00, w0, c0);
    uint64_t mid = pf_madwide(a0, b1, (uint64_t)c0);           // <= (2^32-1)^2 + 2^32-1 : no overflow
    uint32_t m0, m1; pf_unpack(mid, m0, m1);
    uint64_t mid2 = pf_madwide(a1, b0, (uint64_t)m0);
    uint32_t w1, m2; pf_unpack(mid2, w1, m2);
    uint64_t hi = pf_madwide(a1, b1, (uint64_t)m1 + (uint64_t)m2);
    uint32_t w2, w3; pf_unpack(hi, w2, w3);
    return pf_reduce128(w0, w1, w2, w3);
}

PF_FN uint64_t pf_sqr(uint64_t a) {
    uint32_t a0, a1;
    pf_unpack(a, a0, a1);
    uint64_t p00 = pf_mulwide(a0, a0);
    uint64_t p01 = pf_mulwide(a0, a1);
    uint64_t p11 = pf_mulwide(a1, a1);
    // a^2 = p00 + 2^33 p01 + 2^64 p11
    uint32_t w0, w1, w2, w3, q0, q1;
    pf_unpack(p00, w0, w1); pf_unpack(p11, w2, w3); pf_unpack(p01, q0, q1);
    uint32_t s0 = q0 << 1, s1 = pf_funnelshift_l(q0, q1, 1), s2 = q1 >> 31;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, %5;\n\t" : "+r"(w1), "+r"(w2), "+r"(w3) : "r"(s0), "r"(s1), "r"(s2));
#else
    { uint64_t c = (uint64_t)w1 + s0; w1 = (uint32_t)c; c = (uint64_t)w2 + s1 + (c >> 32); w2 = (uint32_t)c; w3 = w3 + s2 + (uint32_t)(c >> 32); }
#endif
    return pf_reduce128(w0, w1, w2, w3);
}

PF_FN uint64_t pf_sbox7(uint64_t x) {
    uint64_t x2 = pf_sqr(x), x4 = pf_sqr(x2), x3 = pf_mul(x, x2);
    return pf_mul(x3, x4);
}

// x + c for canonical c (x any u64): result any u64 congruent
PF_FN uint64_t pf_add_canon(uint64_t x, uint64_t c) {
#if !defined(__CUDA_ARCH__)
    uint64_t r = x + c;
    if (r < x) r += GL_EPS;
    return r;
#else
    uint32_t x0, x1, c0, c1, m;
    pf_unpack(x, x0, x1); pf_unpack(c, c0, c1);
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, 0, 0;\n\t" : "+r"(x0), "+r"(x1), "=r"(m) : "r"(c0), "r"(c1));
    m = 0u - m;
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, 0;\n\t" : "+r"(x0), "+r"(x1) : "r"(m));
    return pf_pack(x0, x1);
#endif
}

PF_FN uint64_t pf_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

#if defined(__CUDACC__)
// MDS layer + constants of the following round (rc == nullptr: none).  In/out non-canonical.  (dense form, micro-benchmark only)
template <bool ADD_RC>
PF_FN void pf_mds(uint64_t s[12], const uint64_t* __restrict__ rc) {
    constexpr uint32_t C[12] = ZK_POSEIDON_MDS_CIRC_INIT;
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) pf_unpack(s[i], lo[i], hi[i]);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
        if (ADD_RC) { uint32_t c0, c1; pf_unpack(rc[r], c0, c1); al = c0; ah = c1; }
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al = pf_madwide(lo[(i + r) % 12], C[i], al);
            ah = pf_madwide(hi[(i + r) % 12], C[i], ah);
        }
        if (r == 0) { al = pf_madwide(lo[0], ZK_POSEIDON_MDS_DIAG0, al); ah = pf_madwide(hi[0], ZK_POSEIDON_MDS_DIAG0, ah); }
        // value = al + 2^32 ah, al, ah < 2^42.   2^64 ah.hi == EPS ah.hi:  t = al + EPS*ah.hi (< 2^44), then add ah.lo << 32
        uint32_t ah0, ah1, t0, t1, m;
        pf_unpack(ah, ah0, ah1);
        uint64_t t = pf_madwide(ah1, 0xFFFFFFFFu, al);
        pf_unpack(t, t0, t1);
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.u32 %1, 0, 0;\n\t" : "+r"(t1), "=r"(m) : "r"(ah0));
        m = 0u - m;
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(m));
        s[r] = pf_pack(t0, t1);
    }
}
#endif

// ---- MDS layer in the frequency domain, on three 22/22/20-bit limbs with wrap-around 32-bit arithmetic ----------------
// The circulant part of the MDS matrix is a cyclic convolution of length 12.  Splitting the index as j = b + 3a and taking
// a 4-point DFT over a (roots 1, i, -1, -i: only additions) leaves three 3x3 "twisted" convolutions whose kernels are, for
// this matrix, all powers of two (DFT of circ/4: [16,32,16], [-1,-8,2], and (2+i, -4-i, 16-i) for the complex pair), so the
// whole layer is ~90 additions / shift-additions per limb and needs no multiplier at all (IMAD.WIDE costs 4 issue slots on
// sm_100a, IADD3/LEA one).  All arithmetic is mod 2^32: every true output limb is < 2^22 * 285 < 2^31, so wrap-around in
// the signed intermediates is harmless.
struct PfRc3 { uint32_t v[372 * 3]; };   // 30 rounds + one all-zero round (constants "after" the last MDS)
constexpr PfRc3 pf_make_rc3() {
    PfRc3 r{};
    constexpr uint64_t rc[360] = ZK_POSEIDON_RC_INIT;
    for (int i = 0; i < 360; i++) {
        r.v[3 * i] = (uint32_t)(rc[i] & 0x3FFFFFu);
        r.v[3 * i + 1] = (uint32_t)((rc[i] >> 22) & 0x3FFFFFu);
        r.v[3 * i + 2] = (uint32_t)(rc[i] >> 44);
    }
    return r;
}
#if defined(__CUDACC__)
static __device__ __constant__ PfRc3 POSEIDON_RC3_DEV = pf_make_rc3();
#endif
#if defined(__CUDA_ARCH__)
#define PF_RC3 POSEIDON_RC3_DEV.v
#else
static const PfRc3 POSEIDON_RC3_HOST = pf_make_rc3();
#define PF_RC3 POSEIDON_RC3_HOST.v
#endif

PF_FN void pf_mds_fft_limb(const uint32_t s[12], uint32_t o[12]) { poseidon_mds_freq<uint32_t>(s, o); }

// rc3: the next round's constants pre-split into the same limbs ([12][3]), or nullptr
template <bool ADD_RC>
PF_FN void pf_mds_fft(uint64_t s[12], const uint32_t* __restrict__ rc3) {
    uint32_t a0[12], a1[12], a2[12], o0[12], o1[12], o2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t lo, hi; pf_unpack(s[i], lo, hi);
        a0[i] = lo & 0x3FFFFFu;
        a1[i] = pf_funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
        a2[i] = hi >> 12;
    }
    pf_mds_fft_limb(a0, o0); pf_mds_fft_limb(a1, o1); pf_mds_fft_limb(a2, o2);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t x0 = o0[i], x1 = o1[i], x2 = o2[i];
        if (ADD_RC) { x0 += rc3[3 * i]; x1 += rc3[3 * i + 1]; x2 += rc3[3 * i + 2]; }
        // value = x0 + 2^22 x1 + 2^44 x2 (each < 2^31).  x2 = 2^20 h + l : 2^64 h == (2^32 - 1) h
        uint32_t h = x2 >> 20, l = x2 & 0xFFFFFu;
        x1 += h << 10;                                  // 2^32 h = 2^22 (2^10 h); x1 stays < 2^32
        uint32_t t0, t1, m;
        // t = x0 + 2^22 x1 - h   (>= 0 because x1 >= 2^10 h)
        uint32_t s0 = x1 << 22, s1 = x1 >> 10;
#if !defined(__CUDA_ARCH__)
        {   // the same steps with 64-bit words
            (void)t0; (void)t1; (void)m;
            uint64_t t = (uint64_t)x0 + pf_pack(s0, s1) - h;
            const uint64_t add = (uint64_t)(l << 12) << 32;
            uint64_t r = t + add;
            if (r < t) r += GL_EPS;
            s[i] = r;
            continue;
        }
#else
        asm("add.cc.u32 %0, %2, %3;\n\t"
            "addc.u32 %1, %4, 0;\n\t" : "=r"(t0), "=r"(t1) : "r"(x0), "r"(s0), "r"(s1));
        asm("sub.cc.u32 %0, %0, %2;\n\t"
            "subc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(h));
        // + 2^44 l : may carry out of 64 bits (2^64 == EPS)
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.u32 %1, 0, 0;\n\t" : "+r"(t1), "=r"(m) : "r"(l << 12));
        m = 0u - m;
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(m));
        s[i] = pf_pack(t0, t1);
#endif
    }
}

// ---- partial rounds with lanes 1..11 kept in limb form ------------------------------------------------------------------
// Only lane 0 goes through the S-box in a partial round, so only lane 0 has to be a 64-bit word there: lanes 1..11 stay in the
// three 22/22/20-bit limbs across all 22 partial rounds and are merely re-normalised (carries pushed up, the bits above 2^64
// folded back through 2^64 == 2^32 - 1) instead of folded to 64 bits and split again every round.
// (Round 1's first version of this mis-compiled: ptxas folded a negation into the funnel shift of h (2^32 - 1), `LEA.HI Rd, -Ra, ..`.
// The limb is written without a negated shifted operand here: (h (2^32 - 1)) >> 22 = (h << 10) - [h != 0].  Checked on a B200 against
// the 64-bit form on 4 Mi random states, profiles/r2a_pbench.txt: 1126 -> 1179 M perm/s.)
PF_FN void pfl_split(uint64_t v, uint32_t& x0, uint32_t& x1, uint32_t& x2) {
    uint32_t lo, hi; pf_unpack(v, lo, hi);
    x0 = lo & 0x3FFFFFu;
    x1 = pf_funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
    x2 = hi >> 12;
}
// x0 + 2^22 x1 + 2^44 x2 (x0 < 2^32, x1 + 2^10 (x2 >> 20) < 2^32) -> some u64 congruent mod p (the same steps as pf_mds_fft's fold)
PF_FN uint64_t pfl_fold(uint32_t x0, uint32_t x1, uint32_t x2) {
    const uint32_t h = x2 >> 20, l = x2 & 0xFFFFFu;
    x1 += h << 10;
    const uint64_t t = (uint64_t)x0 + (((uint64_t)x1) << 22) - h;      // >= 0 because x1 >= 2^10 h; < 2^55
    return pf_add_canon(t, (uint64_t)l << 44);                           // l 2^44 < p: a canonical addend
}
// limbs < 2^32 (x2 < 2^30) -> x0 < 2^23, x1 < 2^22 + 2^19, x2 < 2^20, same value mod p, every limb >= 0
PF_FN void pfl_norm(uint32_t& x0, uint32_t& x1, uint32_t& x2) {
    x1 += x0 >> 22; x0 &= 0x3FFFFFu;
    x2 += x1 >> 22; x1 &= 0x3FFFFFu;
    const uint32_t h = x2 >> 20; x2 &= 0xFFFFFu;
    // 2^64 h == h (2^32 - 1) = 2^22 ((h << 10) - [h != 0]) + ((2^22 - h) mod 2^22)
    x0 += (0x400000u - h) & 0x3FFFFFu;
    x1 += (h << 10) - (h != 0 ? 1u : 0u);
}
PF_FN void pfl_partial_rounds(uint64_t s[12]) {
    uint32_t a0[12], a1[12], a2[12];
#pragma unroll
    for (int i = 1; i < 12; i++) pfl_split(s[i], a0[i], a1[i], a2[i]);
    uint64_t s0 = s[0];
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s0 = pf_sbox7(s0);
        pfl_split(s0, a0[0], a1[0], a2[0]);
        uint32_t o0[12], o1[12], o2[12];
        pf_mds_fft_limb(a0, o0); pf_mds_fft_limb(a1, o1); pf_mds_fft_limb(a2, o2);
        const uint32_t* __restrict__ rc3 = PF_RC3 + 36 * (r + 1);
        s0 = pfl_fold(o0[0] + rc3[0], o1[0] + rc3[1], o2[0] + rc3[2]);
#pragma unroll
        for (int i = 1; i < 12; i++) {
            a0[i] = o0[i] + rc3[3 * i]; a1[i] = o1[i] + rc3[3 * i + 1]; a2[i] = o2[i] + rc3[3 * i + 2];
            pfl_norm(a0[i], a1[i], a2[i]);
        }
    }
    s[0] = s0;
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = pfl_fold(a0[i], a1[i], a2[i]);
}

// s: canonical or not on input; NON-canonical on output (apply pf_canon to the words that are stored).
//
// Code size matters more than instruction count here: the fully unrolled 30-round body is ~90 KB of SASS, far beyond the
// 32 KB L1.5 instruction cache, and ncu showed warps stalled on instruction fetch ("no_instructions") for most cycles.
// So the full rounds are ONE loop body used by both halves: the full S-box layer is 3 iterations of "4 S-boxes + rotate the state
// by 4 lanes" (static register indices, 24 moves per iteration), the MDS layer (with the next round's constants folded in) appears
// once; the partial rounds are the second loop (pfl_partial_rounds).
PF_FN void pf_permute(uint64_t s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = pf_add_canon(s[i], poseidon_rc(i));
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            const int r = half * 26 + q;
#pragma unroll 1
            for (int k = 0; k < 3; k++) {
                uint64_t t0 = pf_sbox7(s[0]), t1 = pf_sbox7(s[1]), t2 = pf_sbox7(s[2]), t3 = pf_sbox7(s[3]);
#pragma unroll
                for (int i = 0; i < 8; i++) s[i] = s[i + 4];
                s[8] = t0; s[9] = t1; s[10] = t2; s[11] = t3;
            }
            pf_mds_fft<true>(s, PF_RC3 + 36 * (r + 1));
        }
        if (half == 0) pfl_partial_rounds(s);
    }
}

// the round-1 form (every lane folded to 64 bits after every MDS layer), kept for the micro-benchmark (tools/pbench.cu variant 3)
PF_FN void pf_permute_r1(uint64_t s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = pf_add_canon(s[i], poseidon_rc(i));
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        if (r < 4 || r >= 26) {
#pragma unroll 1
            for (int k = 0; k < 3; k++) {
                uint64_t t0 = pf_sbox7(s[0]), t1 = pf_sbox7(s[1]), t2 = pf_sbox7(s[2]), t3 = pf_sbox7(s[3]);
#pragma unroll
                for (int i = 0; i < 8; i++) s[i] = s[i + 4];
                s[8] = t0; s[9] = t1; s[10] = t2; s[11] = t3;
            }
        } else {
            s[0] = pf_sbox7(s[0]);
        }
        pf_mds_fft<true>(s, PF_RC3 + 36 * (r + 1));
    }
}

#if defined(__CUDACC__)
// the fully unrolled / dense-MDS forms, kept for the micro-benchmark (tools/pbench.cu)
template <int MDS>
PF_FN void pf_permute_unrolled(uint64_t s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = pf_add_canon(s[i], POSEIDON_RC_DEV[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = pf_sbox7(s[i]);
        if (MDS == 0) pf_mds<true>(s, POSEIDON_RC_DEV + 12 * (r + 1)); else pf_mds_fft<true>(s, POSEIDON_RC3_DEV.v + 36 * (r + 1));
    }
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s[0] = pf_sbox7(s[0]);
        if (MDS == 0) pf_mds<true>(s, POSEIDON_RC_DEV + 12 * (r + 1)); else pf_mds_fft<true>(s, POSEIDON_RC3_DEV.v + 36 * (r + 1));
    }
#pragma unroll 1
    for (int r = 26; r < 29; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = pf_sbox7(s[i]);
        if (MDS == 0) pf_mds<true>(s, POSEIDON_RC_DEV + 12 * (r + 1)); else pf_mds_fft<true>(s, POSEIDON_RC3_DEV.v + 36 * (r + 1));
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = pf_sbox7(s[i]);
    if (MDS == 0) pf_mds<false>(s, nullptr); else pf_mds_fft<false>(s, nullptr);
}

#endif  // __CUDACC__ (micro-benchmark forms)

}  // namespace zk
