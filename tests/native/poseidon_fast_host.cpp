// TEST INFRASTRUCTURE: the product's fast Poseidon permutation (zk_evm_b200/csrc/poseidon_fast.cuh — non-canonical state, lazy
// 128 -> 64 reductions, MDS layer in the frequency domain on 22-bit limbs, one compact round loop) compiled for the host, where
// the inline-PTX carry chains are replaced by their portable equivalents.  Pins the ALGORITHM of the device permutation against the
// oracle on the CPU (tests/test_poseidon_fast_host.py); the PTX itself is covered by the GPU parity tests.
#include <stddef.h>
#include "poseidon_fast.cuh"
using namespace zk;

extern "C" void pf_permute_host(uint64_t* states, size_t count) {
    for (size_t i = 0; i < count; i++) {
        uint64_t s[12];
        for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
        pf_permute(s);
        for (int k = 0; k < 12; k++) states[12 * i + k] = pf_canon(s[k]);
    }
}
// the building blocks on arbitrary (non-canonical) 64-bit inputs: out = {mul, sqr, sbox7, add_canon(a, canon(b))}, all canonicalised
extern "C" void pf_ops_host(uint64_t a, uint64_t b, uint64_t out[4]) {
    out[0] = pf_canon(pf_mul(a, b));
    out[1] = pf_canon(pf_sqr(a));
    out[2] = pf_canon(pf_sbox7(a));
    out[3] = pf_canon(pf_add_canon(a, b >= GL_P ? b - GL_P : b));
}
