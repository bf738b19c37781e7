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

// out[r] = sum_i s[(i+r)%12] * C[i] + (r==0 ? 8*s[0] : 0)
ZK_HD void poseidon_mds(uint64_t s[12]) {
    constexpr uint32_t C[12] = ZK_POSEIDON_MDS_CIRC_INIT;
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = (uint32_t)s[i]; hi[i] = (uint32_t)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (uint64_t)lo[(i + r) % 12] * C[i];
            ah += (uint64_t)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) { al += (uint64_t)lo[0] * ZK_POSEIDON_MDS_DIAG0; ah += (uint64_t)hi[0] * ZK_POSEIDON_MDS_DIAG0; }
        // value = al + 2^32 * ah  (al, ah < 2^41) -> 96-bit number
        uint64_t low = al + (ah << 32);
        uint32_t top = (uint32_t)(ah >> 32) + (low < al ? 1u : 0u);
        s[r] = gl_reduce96(low, top);
    }
}

ZK_HD void poseidon_permute(uint64_t s[12]) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add(s[i], poseidon_rc(12 * r + i)));
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
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add(s[i], poseidon_rc(12 * r + i)));
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
