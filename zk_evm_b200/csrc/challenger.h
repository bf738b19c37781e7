// Host-side Fiat-Shamir transcript: plonky2 1.0.0 iop/challenger.rs `Challenger<F, PoseidonHash>` as driven by
// /root/reference/evm_arithmetization/src/prover.rs:118-144,320 and get_challenges.rs:202-227.
// Duplex sponge in overwrite mode, rate 8; challenges are popped from the back of the squeezed block.
#pragma once
#include "poseidon.cuh"
#include <vector>
#include <string.h>

namespace zk {

// Host form of the permutation for the transcript (a few thousand permutations per segment proof): the state stays
// non-canonical (any u64 congruent mod p) between operations, reductions are branch-free carry corrections.
namespace hostp {
inline uint64_t reduce128(unsigned __int128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0;
    bool borrow = __builtin_sub_overflow(lo, hi_hi, &t0);
    t0 -= borrow ? GL_EPS : 0;                       // 2^64 == EPS: cannot borrow twice
    uint64_t t1 = hi_lo * GL_EPS, t2;
    bool carry = __builtin_add_overflow(t0, t1, &t2);
    return t2 + (carry ? GL_EPS : 0);                // cannot carry twice
}
inline uint64_t mul(uint64_t a, uint64_t b) { return reduce128((unsigned __int128)a * b); }
inline uint64_t sbox7(uint64_t x) { uint64_t x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x, x2); return mul(x3, x4); }
inline uint64_t add_canon(uint64_t x, uint64_t c) {  // c < p
    uint64_t s; bool carry = __builtin_add_overflow(x, c, &s);
    return s + (carry ? GL_EPS : 0);
}
inline void mds(uint64_t s[12], const uint64_t* rc_next) {
    uint64_t lo[12], hi[12], ol[12], oh[12];
    for (int i = 0; i < 12; i++) { lo[i] = (uint32_t)s[i]; hi[i] = s[i] >> 32; }
    poseidon_mds_freq<uint64_t>(lo, ol);
    poseidon_mds_freq<uint64_t>(hi, oh);
    for (int r = 0; r < 12; r++) {
        unsigned __int128 v = (unsigned __int128)ol[r] + ((unsigned __int128)oh[r] << 32);
        if (rc_next) v += rc_next[r];
        s[r] = reduce128(v);
    }
}
inline void permute(uint64_t s[12]) {
    for (int i = 0; i < 12; i++) s[i] = add_canon(s[i], POSEIDON_RC_HOST[i]);
    for (int r = 0; r < 30; r++) {
        if (r < 4 || r >= 26) for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        else s[0] = sbox7(s[0]);
        mds(s, r < 29 ? POSEIDON_RC_HOST + 12 * (r + 1) : nullptr);
    }
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}
}  // namespace hostp

struct Challenger {
    uint64_t state[12];
    std::vector<uint64_t> in, out;
    Challenger() { memset(state, 0, sizeof(state)); }
    void duplex() {
        for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
        in.clear();
        hostp::permute(state);
        out.assign(state, state + 8);
    }
    void observe(uint64_t x) { out.clear(); in.push_back(x); if (in.size() == 8) duplex(); }
    void observe_n(const uint64_t* x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
    void observe_vec(const std::vector<uint64_t>& v) { observe_n(v.data(), v.size()); }
    uint64_t challenge() {
        if (!in.empty() || out.empty()) duplex();
        uint64_t r = out.back(); out.pop_back(); return r;
    }
    Fp2 ext_challenge() { uint64_t a = challenge(); uint64_t b = challenge(); return Fp2(a, b); }
    // Challenger::compact: absorb pending input, drop buffered outputs (prover.rs:320)
    void compact() { if (!in.empty()) duplex(); out.clear(); }
    void set_state(const uint64_t s[12]) { memcpy(state, s, sizeof(state)); in.clear(); out.clear(); }
};

}  // namespace zk
