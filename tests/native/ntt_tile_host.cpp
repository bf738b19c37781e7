// TEST INFRASTRUCTURE: runs the product's NTT tile code (zk_evm_b200/csrc/ntt_tile.cuh, host+device) on the CPU — one "thread"
// walking every item of every phase in program order — with the same pass plan and tables as ntt.cu's launcher, so that the index
// arithmetic and the power-of-two butterfly network can be checked against the oracle transform without a GPU
// (tests/test_ntt_tile_host.py).  Not part of the product; the product path is the CUDA kernel that wraps the same header.
#include "ntt_tile.cuh"
#include <vector>
using namespace zk;

extern "C" int ntt_tile_host(uint64_t* data, size_t ncols, unsigned L, int inverse, const uint64_t* prescale) {
    const size_t n = (size_t)1 << L;
    if (L == 0) return 0;
    std::vector<unsigned> digits;
    ntt_plan_passes(L, digits);
    std::vector<uint64_t> roots((size_t)1 << NTT_ROOT_LOG);
    uint64_t w = gl_root_of_unity(NTT_ROOT_LOG);
    if (inverse) w = gl_inv(w);
    roots[0] = 1;
    for (size_t k = 1; k < roots.size(); k++) roots[k] = gl_mul(roots[k - 1], w);
    std::vector<uint64_t> sm(NTT_TILE_WORDS);
    unsigned m = L;
    for (size_t pi = 0; pi < digits.size(); pi++) {
        const bool first = pi == 0, last = pi + 1 == digits.size();
        PassParams p;
        p.src = data; p.dst = data; p.src_stride = n; p.dst_stride = n; p.src_shift = 0;
        p.log_n = L; p.m = m; p.r = digits[pi];
        p.strided = last ? 0 : 1;
        p.t = ntt_pass_t(L, m, p.r, last);
        p.roots = roots.data();
        p.prescale0 = first ? prescale : nullptr; p.prescale1 = nullptr; p.prescale_mask = 0;
        std::vector<uint64_t> ip;
        if (!last) {
            // tab[kd * M' + j'] = w_M^(j' * kd)
            const size_t M = (size_t)1 << m;
            const unsigned mp = m - p.r;
            uint64_t wm = gl_root_of_unity(m);
            if (inverse) wm = gl_inv(wm);
            ip.resize(M);
            for (size_t idx = 0; idx < M; idx++) {
                size_t kd = idx >> mp, jp = idx & (((size_t)1 << mp) - 1);
                ip[idx] = gl_pow(wm, (kd * jp) & (M - 1));
            }
        }
        p.interpass = last ? nullptr : ip.data();
        const unsigned tile_log = p.r + p.t;
        const size_t tiles = (size_t)1 << (L - tile_log);
        for (size_t c = 0; c < ncols; c++)
            for (size_t tile = 0; tile < tiles; tile++) {
                if (inverse) ntt_pass_tile<true>(p, tile, c, sm.data(), 0, 1);
                else ntt_pass_tile<false>(p, tile, c, sm.data(), 0, 1);
            }
        m -= p.r;
    }
    return 0;
}
