// EXPERIMENT (tools/pbench.cu variant 4, not used by the library): Poseidon with lanes 1..11 kept in three 22/22/20-bit limbs across the
// 22 partial rounds, so that only lane 0 (the one that goes through the S-box) is folded back to 64 bits every round.
// History: the first version of this was right as a host build and at PTX level (tools/ptx_emu.py) but wrong on the
// device, and ~3 % faster; the suspect is ptxas folding a negation into LEA.HI for the funnel shift of h (2^32 - 1).  This version
// writes that limb without a negated funnel-shift operand: (h (2^32 - 1)) >> 22 = (h << 10) - (h != 0).
#pragma once
#include "poseidon_fast.cuh"

namespace zk {

PF_FN void pfl_split(uint64_t v, uint32_t& x0, uint32_t& x1, uint32_t& x2) {
    uint32_t lo, hi; pf_unpack(v, lo, hi);
    x0 = lo & 0x3FFFFFu;
    x1 = pf_funnelshift_r(lo, hi, 22) & 0x3FFFFFu;
    x2 = hi >> 12;
}
// x0 + 2^22 x1 + 2^44 x2 (x0 < 2^32, x1 + 2^10 (x2 >> 20) < 2^32) -> some u64 congruent mod p (same steps as pf_mds_fft's fold)
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

PF_FN void pf_permute_limb(uint64_t s[12]) {
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

}  // namespace zk
