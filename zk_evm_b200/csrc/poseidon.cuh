// Poseidon permutation over Goldilocks, width 12 (plonky2 1.0.0 `PoseidonHash`, the `C::Hasher` of
// PoseidonGoldilocksConfig: /root/reference/evm_arithmetization/src/public_types.rs:9-27), host+device.
//
// Round structure: 4 full + 22 partial + 4 full, S-box x^7, MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0..0).
// We evaluate the plain 30-round form: because every MDS entry is < 2^6 the dense layer is done on the 32-bit
// halves of the state with 64-bit accumulators (no modular multiplications), one 96-bit reduction per output.
// This computes the same function as plonky2's "fast partial round" factorisation.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

namespace zk {

#if defined(__CUDACC__)
static __device__ __constant__ uint64_t POSEIDON_RC_DEV[360] = ZK_POSEIDON_RC_INIT;
#endif
static const uint64_t POSEIDON_RC_HOST[360] = ZK_POSEIDON_RC_INIT;

ZK_HD uint64_t poseidon_rc(int i) {
#if defined(__CUDA_ARCH__)
    return POSEIDON_RC_DEV[i];
#else
    return POSEIDON_RC_HOST[i];
#endif
}

ZK_HD uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

// MDS layer: out[r] = sum_i s[(i+r)%12] * C[i] + (r==0 ? 8*s[0] : 0).
// Evaluated in the frequency domain (derivation in poseidon_fast.cuh): the circulant part is a cyclic convolution of length 12;
// a 4-point DFT over the index split j = b + 3a leaves three 3x3 twisted convolutions whose kernels are powers of two for this
// matrix, so the layer is ~90 additions / shifts in wrap-around integer arithmetic of type T.  T = uint32_t on 22-bit limbs
// (device), T = uint64_t on the 32-bit halves of the state (host transcript): every true output is < 2^bits(T).
template <class T>
ZK_HD void poseidon_mds_freq(const T s[12], T o[12]) {
    T F1[3], Fm[3], Fc[3], Fd[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        T x0 = s[b], x1 = s[b + 3], x2 = s[b + 6], x3 = s[b + 9];
        T A = x0 + x2, B = x1 + x3;
        Fc[b] = x0 - x2; Fd[b] = x1 - x3;
        F1[b] = A + B; Fm[b] = A - B;
    }
    T Tsum = F1[0] + F1[1] + F1[2];
    T G1[3] = {(Tsum + F1[2]) << 4, (Tsum + F1[0]) << 4, (Tsum + F1[1]) << 4};
    T Gm[3] = {(Fm[2] << 3) - (Fm[1] << 1) - Fm[0], (T)0 - (Fm[0] << 3) - (Fm[2] << 1) - Fm[1], (Fm[0] << 1) - (Fm[1] << 3) - Fm[2]};
    // complex products with k0 = 2+i, k1 = -4-i, k2 = 16-i :  (p+qi)(c+di) = (pc - qd) + (pd + qc) i
    //   k0 F = (2c - d, 2d + c) ; k1 F = (-4c + d, -4d - c) ; k2 F = (16c + d, 16d - c)
    // Gi_0 = k0 F0 + i (k1 F2 + k2 F1) ; Gi_1 = k0 F1 + k1 F0 + i k2 F2 ; Gi_2 = k0 F2 + k1 F1 + k2 F0
    T u[3], v[3];
    {
        // k1 F2 + k2 F1 = (-4c2 + d2 + 16c1 + d1, -4d2 - c2 + 16d1 - c1); times i -> (-(im), re)
        T re = (Fc[1] << 4) - (Fc[2] << 2) + Fd[2] + Fd[1];
        T im = (Fd[1] << 4) - (Fd[2] << 2) - Fc[2] - Fc[1];
        u[0] = (Fc[0] << 1) - Fd[0] - im;
        v[0] = (Fd[0] << 1) + Fc[0] + re;
    }
    {
        // k0 F1 + k1 F0 + i k2 F2, k2 F2 = (16c2 + d2, 16d2 - c2) -> i * = (-(16d2 - c2), 16c2 + d2)
        u[1] = (Fc[1] << 1) - Fd[1] - (Fc[0] << 2) + Fd[0] - (Fd[2] << 4) + Fc[2];
        v[1] = (Fd[1] << 1) + Fc[1] - (Fd[0] << 2) - Fc[0] + (Fc[2] << 4) + Fd[2];
    }
    {
        u[2] = (Fc[2] << 1) - Fd[2] - (Fc[1] << 2) + Fd[1] + (Fc[0] << 4) + Fd[0];
        v[2] = (Fd[2] << 1) + Fc[2] - (Fd[1] << 2) - Fc[1] + (Fd[0] << 4) - Fc[0];
    }
#pragma unroll
    for (int b = 0; b < 3; b++) {
        T P = G1[b] + Gm[b], Q = G1[b] - Gm[b];
        o[b] = P + u[b]; o[b + 3] = Q + v[b]; o[b + 6] = P - u[b]; o[b + 9] = Q - v[b];
    }
    o[0] += s[0] << 3;
}


ZK_HD void poseidon_mds(uint64_t s[12]) {
    uint64_t lo[12], hi[12], ol[12], oh[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = (uint32_t)s[i]; hi[i] = s[i] >> 32; }
    poseidon_mds_freq<uint64_t>(lo, ol);
    poseidon_mds_freq<uint64_t>(hi, oh);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        // value = ol + 2^32 * oh  (ol, oh < 2^41) -> 96-bit number
        uint64_t al = ol[r], ah = oh[r];
        uint64_t low = al + (ah << 32);
        uint32_t top = (uint32_t)(ah >> 32) + (low < al ? 1u : 0u);
        s[r] = gl_reduce96(low, top);
    }
}

// the full S-box layer, power by power across the twelve lanes: twelve independent products per step instead of twelve dependent
// chains one after the other (the host transcript's permutation is latency-bound otherwise: 3.0 -> 1.0 us for the eight layers)
ZK_HD void poseidon_sbox_layer(uint64_t s[12], const int rc_base) {
    uint64_t x[12], x2[12], x3[12], x4[12];
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = gl_add(s[i], poseidon_rc(rc_base + i));
#pragma unroll
    for (int i = 0; i < 12; i++) x2[i] = gl_sqr(x[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) x4[i] = gl_sqr(x2[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) x3[i] = gl_mul(x[i], x2[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_mul(x3[i], x4[i]);
}

ZK_HD void poseidon_permute(uint64_t s[12]) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        poseidon_sbox_layer(s, 12 * r);
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], poseidon_rc(12 * r + i));
        s[0] = sbox7(s[0]);
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 26; r < 30; r++) {
        poseidon_sbox_layer(s, 12 * r);
        poseidon_mds(s);
    }
}

// PoseidonHash::two_to_one
ZK_HD void poseidon_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    poseidon_permute(s);
    out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; out[3] = s[3];
}

}  // namespace zk
