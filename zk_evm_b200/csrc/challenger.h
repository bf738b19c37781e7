// Host-side Fiat-Shamir transcript: plonky2 1.0.0 iop/challenger.rs `Challenger<F, PoseidonHash>` as driven by
// /root/reference/evm_arithmetization/src/prover.rs:118-144,320 and get_challenges.rs:202-227.
// Duplex sponge in overwrite mode, rate 8; challenges are popped from the back of the squeezed block.
#pragma once
#include "poseidon.cuh"
#include <vector>
#include <string.h>

namespace zk {

struct Challenger {
    uint64_t state[12];
    std::vector<uint64_t> in, out;
    Challenger() { memset(state, 0, sizeof(state)); }
    void duplex() {
        for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
        in.clear();
        poseidon_permute(state);
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
