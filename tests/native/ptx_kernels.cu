// TEST INFRASTRUCTURE: single-thread wrapper kernels around the product's device functions, compiled to PTX (nvcc -ptx, no GPU needed)
// and executed by tools/ptx_emu.py in tests/test_ptx_emulation.py — the device instruction streams (inline PTX carry chains
// included) of the field arithmetic, the fast Poseidon permutation and the radix-16 butterfly network, checked on the CPU.
#include "poseidon_fast.cuh"
#include "ntt_tile.cuh"
using namespace zk;

extern "C" __global__ void k_perm(const uint64_t* in, uint64_t* out) {
    uint64_t s[12];
    for (int k = 0; k < 12; k++) s[k] = in[k];
    pf_permute(s);
    for (int k = 0; k < 12; k++) out[k] = pf_canon(s[k]);
}
// canonical a, b -> add, sub, mul, neg(a), reduce128(lo = a, hi = b)
extern "C" __global__ void k_field(const uint64_t* in, uint64_t* out) {
    const uint64_t a = in[0], b = in[1];
    out[0] = gl_add(a, b); out[1] = gl_sub(a, b); out[2] = gl_mul(a, b); out[3] = gl_neg(a); out[4] = gl_reduce128(a, b);
}
// any 64-bit a, b -> the lazy forms of the permutation, canonicalised
extern "C" __global__ void k_lazy(const uint64_t* in, uint64_t* out) {
    const uint64_t a = in[0], b = in[1];
    out[0] = pf_canon(pf_mul(a, b)); out[1] = pf_canon(pf_sqr(a)); out[2] = pf_canon(pf_sbox7(a));
    out[3] = pf_canon(pf_add_canon(a, b >= GL_P ? b - GL_P : b));
}
template <int S> __device__ void pow2_one(const uint64_t* in, uint64_t* out, int slot) { out[slot] = gl_mul_pow2<S>(in[0]); }
extern "C" __global__ void k_pow2(const uint64_t* in, uint64_t* out) {
    pow2_one<1>(in, out, 0); pow2_one<12>(in, out, 1); pow2_one<24>(in, out, 2); pow2_one<31>(in, out, 3); pow2_one<32>(in, out, 4);
    pow2_one<36>(in, out, 5); pow2_one<48>(in, out, 6); pow2_one<60>(in, out, 7); pow2_one<63>(in, out, 8); pow2_one<64>(in, out, 9);
    pow2_one<72>(in, out, 10); pow2_one<84>(in, out, 11); pow2_one<95>(in, out, 12);
}
extern "C" __global__ void k_dft16(const uint64_t* in, uint64_t* out, int inverse) {
    uint64_t x[16];
    for (int k = 0; k < 16; k++) x[k] = in[k];
    if (inverse) ntt_dft_regs<4, true>(x); else ntt_dft_regs<4, false>(x);
    for (int k = 0; k < 16; k++) out[k] = x[k];
}
