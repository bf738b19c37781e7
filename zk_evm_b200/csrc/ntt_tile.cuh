// One tile of one pass of the batched Goldilocks NTT (see ntt.cu), host+device so the index arithmetic and the butterfly
// network can be executed on the CPU by tests/native/ntt_tile_host.cpp (tid = 0, nthreads = 1 walks every item of every phase in
// program order) and compared with the oracle transform without a GPU.
//
// Replaces plonky2_field 1.0.0 fft.rs `fft_classic` (radix-2, one twiddle multiplication per butterfly).  Here a pass over an
// r-bit index digit runs in ROUNDS of up to four radix-2 stages on 16 elements held in registers.  In Goldilocks 2 has order 192
// (2^96 = -1) and plonky2's primitive roots satisfy w_64 = 2^3, w_16 = 2^12, w_8 = 2^24, w_4 = 2^48: every twiddle INSIDE a round
// is a power of two, i.e. a compile-time shift followed by the 96-bit shift-add reduction — no multiplier.  One general
// multiplication per element remains per round (the twiddle w^(kf * lo) between rounds), instead of one per two stages.
#pragma once
#include "gl.cuh"
#include <utility>
#include <vector>

namespace zk {

static constexpr unsigned NTT_ROOT_LOG = 12;       // roots[k] = w_4096^(+-k), k < 4096
static constexpr unsigned NTT_MAX_TILE_LOG = 12;   // 4096 elements (+ padding) = 34 KB shared memory per CTA
static constexpr unsigned NTT_STRIDED_T = 4;       // 16 contiguous elements = 128 B per row of a strided tile

struct PassParams {
    const uint64_t* src;
    uint64_t* dst;
    size_t src_stride, dst_stride;   // elements between consecutive transforms
    unsigned src_shift;              // transform t reads source column t >> src_shift
    unsigned log_n, m, r, t;         // transform size, size of the sub-transforms this pass starts from, digit bits, tile columns
    unsigned strided;                // 1: tile = 2^r digit values x 2^t contiguous elements; 0: final pass, 2^t consecutive blocks of 2^r
    const uint64_t* roots;           // w_4096^(+-k)
    const uint64_t* interpass;       // [kd * M' + j'] or nullptr
    const uint64_t* prescale0;       // per natural index j, or nullptr
    const uint64_t* prescale1;
    unsigned prescale_mask;          // table = (t & mask) ? prescale1 : prescale0
};

// digit plan of a size-2^L transform, most significant digit first; the last entry is the final (contiguous) pass
static constexpr unsigned NTT_MAX_STRIDED_R = 8;
inline void ntt_plan_passes(unsigned L, std::vector<unsigned>& digits) {
    // strided passes always take 8 bits (a full 2^8 x 16 tile: 256 threads, one radix-16 item each, two rounds); the final
    // contiguous pass takes what is left (<= 12 bits, as many consecutive sub-transforms per tile as fit in 4096 elements)
    digits.clear();
    unsigned rem = L;
    while (rem > NTT_MAX_TILE_LOG) { digits.push_back(NTT_MAX_STRIDED_R); rem -= NTT_MAX_STRIDED_R; }
    digits.push_back(rem);
}
// tile columns of pass `pi` of the plan: strided passes stage 16 contiguous elements per digit value, the final pass as many
// consecutive sub-transforms as fit
inline unsigned ntt_pass_t(unsigned L, unsigned m, unsigned r, bool last) {
    if (!last) return NTT_STRIDED_T;
    unsigned t = NTT_MAX_TILE_LOG > r ? NTT_MAX_TILE_LOG - r : 0;
    if (t > L - m) t = L - m;
    return t;
}

ZK_HD uint64_t ntt_ld(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// one padding element per 16: the 16 lanes of a half-warp that walk the tile with a stride of 16 (or 256, ...) elements then
// fall into distinct 8-byte banks
ZK_HD unsigned ntt_pad(unsigned idx) { return idx + (idx >> 4); }
static constexpr unsigned NTT_TILE_WORDS = (1u << NTT_MAX_TILE_LOG) + (1u << (NTT_MAX_TILE_LOG - 4));

// x * 2^S for canonical x, 0 <= S < 96; canonical result
template <int S>
ZK_HD uint64_t gl_mul_pow2(uint64_t x) {
    static_assert(S >= 0 && S < 96, "shift out of range");
    if constexpr (S == 0) return x;
#if defined(__CUDA_ARCH__)
    constexpr int q = S / 32, r = S % 32;
    uint32_t lo, hi;
    gl_unpack(x, lo, hi);
    uint32_t v0, v1, v2;   // x << r as three words
    if constexpr (r == 0) { v0 = lo; v1 = hi; v2 = 0; }
    else { v0 = lo << r; v1 = __funnelshift_l(lo, hi, r); v2 = hi >> (32 - r); }
    if constexpr (q == 0) {
        // (v1:v0) + v2 * 2^64, 2^64 == 2^32 - 1
        uint32_t u0, u1, m;
        asm("sub.cc.u32 %0, 0, %2;\n\t"
            "subc.u32 %1, %2, 0;\n\t" : "=r"(u0), "=r"(u1) : "r"(v2));
        asm("add.cc.u32 %0, %0, %3;\n\t"
            "addc.cc.u32 %1, %1, %4;\n\t"
            "addc.u32 %2, 0, 0;\n\t" : "+r"(v0), "+r"(v1), "=r"(m) : "r"(u0), "r"(u1));
        m = 0u - m;
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.u32 %1, %1, 0;\n\t" : "+r"(v0), "+r"(v1) : "r"(m));
        return gl_canon(gl_pack(v0, v1));
    } else if constexpr (q == 1) {
        return gld_reduce_words(0, v0, v1, v2);
    } else {
        // v0 2^64 + v1 2^96 + v2 2^128 == v0 (2^32 - 1) - (v1 + 2^32 v2);  the subtrahend is < 2^63
        uint32_t u0, u1;
        asm("sub.cc.u32 %0, 0, %2;\n\t"
            "subc.u32 %1, %2, 0;\n\t" : "=r"(u0), "=r"(u1) : "r"(v0));
        return gld_sub(gl_pack(u0, u1), gl_pack(v1, v2));
    }
#else
    const uint64_t c = S < 64 ? ((uint64_t)1 << (S & 63)) : gl_mul(GL_EPS, (uint64_t)1 << ((S - 64) & 63));
    return gl_mul(x, c);
#endif
}

// butterfly I of stage J of the radix-2^K DIF network on x[0 .. 2^K): pair (k, k + hk), twiddle w_{2^(K-J)}^pos = 2^(+-(192 >> (K-J)) pos)
template <int K, bool INV, int J, int I>
ZK_HD void ntt_bfly(uint64_t (&x)[1 << K]) {
    constexpr int hk = 1 << (K - 1 - J);
    constexpr int pos = I % hk, k = (I / hk) * 2 * hk + pos;
    constexpr int E = (192 >> (K - J)) * pos;            // forward exponent of 2, < 96
    constexpr int e = INV ? (192 - E) % 192 : E;         // 2^-E = 2^(192 - E)
    const uint64_t a = x[k], b = x[k + hk];
    x[k] = gl_add(a, b);
    if constexpr (e == 0) x[k + hk] = gl_sub(a, b);
    else if constexpr (e < 96) x[k + hk] = gl_mul_pow2<e>(gl_sub(a, b));
    else x[k + hk] = gl_mul_pow2<e - 96>(gl_sub(b, a));   // 2^96 = -1
}
template <int K, bool INV, int J, int... I>
ZK_HD void ntt_stage(uint64_t (&x)[1 << K], std::integer_sequence<int, I...>) { (ntt_bfly<K, INV, J, I>(x), ...); }
template <int K, bool INV, int... J>
ZK_HD void ntt_stages(uint64_t (&x)[1 << K], std::integer_sequence<int, J...>) {
    (ntt_stage<K, INV, J>(x, std::make_integer_sequence<int, (1 << K) / 2>()), ...);
}
// size-2^K DFT with root w_{2^K}^(+-1): natural order in, bit-reversed order out
template <int K, bool INV>
ZK_HD void ntt_dft_regs(uint64_t (&x)[1 << K]) { ntt_stages<K, INV>(x, std::make_integer_sequence<int, K>()); }

// where the tile's elements live in global memory (tile index idx = (jd << t) + u for a strided pass, the offset inside the tile for
// the final pass)
struct TileIO {
    const uint64_t* src;
    uint64_t* dst;
    const uint64_t* prescale;    // per transform index, or nullptr
    const uint64_t* interpass;   // already offset to this tile's low index bits, or nullptr
    size_t base;
    unsigned strided, t, mp, r;
    ZK_HD size_t gaddr(unsigned idx) const {
        return strided ? base + ((size_t)(idx >> t) << mp) + (idx & ((1u << t) - 1)) : base + idx;
    }
};

// One round: the K DIF stages over bits s .. s-K+1 of the digit index jd of every sub-transform in the tile, followed by the
// twiddle w_{2^(s+1)}^(kf * lo) (kf = frequency index produced by this round, lo = the bits of jd below the round) when lo exists.
// Tile element (jd, u) lives at ntt_pad((jd << t) + u).  Items are numbered u fastest, then lo, then the bits above the round.
// from_global: the round reads its 2^K elements straight from global memory (+ coset prescale) — the first round of a pass, 2^K
// independent loads in flight per thread; to_global: it writes them straight back (+ inter-pass twiddle) — the last round of a
// strided pass, where the 16 lanes of a half-warp still cover one 128-byte row.
template <int K, bool INV>
ZK_HD void ntt_round(uint64_t* sm, const uint64_t* __restrict__ roots, int s, unsigned tile_log, unsigned t, unsigned tid, unsigned nthreads,
                     const TileIO& io, bool from_global, bool to_global, const uint64_t* stage = nullptr) {
    constexpr int E = 1 << K;
    const int low = s - K + 1;
    const unsigned items = 1u << (tile_log - K);
    for (unsigned w = tid; w < items; w += nthreads) {
        const unsigned u = w & ((1u << t) - 1), g = w >> t;
        const unsigned lo = g & ((1u << low) - 1), hi = g >> low;
        const unsigned base = ((((hi << K) << low) | lo) << t) + u;
        uint64_t x[E];
        if (from_global) {
            // `stage`: the tile as the TMA unit laid it down in shared memory (unpadded, tile index order) — see ntt_pass_tma_kernel
            if (stage) {
#pragma unroll
                for (int k = 0; k < E; k++) x[k] = stage[base + ((unsigned)k << (low + t))];
            } else {
#pragma unroll
                for (int k = 0; k < E; k++) x[k] = io.src[io.gaddr(base + ((unsigned)k << (low + t)))];
            }
            if (io.prescale) {
#pragma unroll
                for (int k = 0; k < E; k++) x[k] = gl_mul(x[k], ntt_ld(io.prescale + io.gaddr(base + ((unsigned)k << (low + t)))));
            }
        } else {
#pragma unroll
            for (int k = 0; k < E; k++) x[k] = sm[ntt_pad(base + ((unsigned)k << (low + t)))];
        }
        ntt_dft_regs<K, INV>(x);
        if (low > 0) {
            const unsigned sh = NTT_ROOT_LOG - (unsigned)(low + K);
#pragma unroll
            for (int kp = 1; kp < E; kp++) {
                const unsigned kf = bitrev32((uint32_t)kp, K);
                x[kp] = gl_mul(x[kp], ntt_ld(roots + ((size_t)(kf * lo) << sh)));
            }
        }
        if (to_global) {
            // low == 0 here: element k is digit position jd = (hi << K) | k, holding frequency kd = rev_r(jd)
#pragma unroll
            for (int k = 0; k < E; k++) {
                const unsigned idx = base + ((unsigned)k << (low + t));
                uint64_t v = x[k];
                if (io.interpass) {
                    const unsigned kd = bitrev32(idx >> t, io.r);
                    v = gl_mul(v, ntt_ld(io.interpass + ((size_t)kd << io.mp) + (idx & ((1u << t) - 1))));
                }
                io.dst[io.gaddr(idx)] = v;
            }
        } else {
#pragma unroll
            for (int k = 0; k < E; k++) sm[ntt_pad(base + ((unsigned)k << (low + t)))] = x[k];
        }
    }
}

#if defined(__CUDA_ARCH__)
#define NTT_TILE_SYNC() __syncthreads()
#else
#define NTT_TILE_SYNC() ((void)0)
#endif

// the whole tile: the rounds of the r-bit digit (first one loading from global memory, last one of a strided pass storing to it),
// in place: digit position jd ends up holding frequency kd = rev_r(jd)
// where tile `tile` of transform `trans` lives (the TileIO of the pass)
ZK_HD TileIO ntt_tile_io(const PassParams& p, size_t tile, size_t trans) {
    const unsigned r = p.r, t = p.t, m = p.m;
    const unsigned mp = m - r;   // log M'
    TileIO io;
    io.src = p.src + (trans >> p.src_shift) * p.src_stride;
    io.dst = p.dst + trans * p.dst_stride;
    io.prescale = p.prescale0 ? ((trans & p.prescale_mask) ? p.prescale1 : p.prescale0) : nullptr;
    io.strided = p.strided; io.t = t; io.mp = mp; io.r = r;
    if (p.strided) {
        // tile id = hi * 2^(mp - t) + lo_hi; element (jd, lo_t) is transform index hi 2^m + jd 2^mp + lo_hi 2^t + lo_t
        const size_t hi = tile >> (mp - t), lo_hi = tile & (((size_t)1 << (mp - t)) - 1);
        io.base = (hi << m) + (lo_hi << t);
        io.interpass = p.interpass ? p.interpass + (lo_hi << t) : nullptr;
    } else {
        // final pass (mp == 0): 2^t consecutive sub-transforms of 2^r elements, tile index = u 2^r + jd
        io.base = tile * ((size_t)1 << (r + t));
        io.interpass = nullptr;
    }
    return io;
}

// the whole tile: the rounds of the r-bit digit (first one loading from global memory — or from `stage`, the tile already brought to
// shared memory by the TMA unit —, last one of a strided pass storing to global memory), in place: digit position jd ends up holding
// frequency kd = rev_r(jd)
template <bool INV>
ZK_HD void ntt_pass_rounds(const PassParams& p, const TileIO& io, uint64_t* sm, unsigned tid, unsigned nthreads, const uint64_t* stage) {
    const unsigned r = p.r, t = p.t;
    const unsigned tile_log = r + t, tile_elems = 1u << tile_log;
    const unsigned tt = p.strided ? t : 0;
    int s = (int)r - 1;
    const int first = (r % 4) ? (int)(r % 4) : 4;
    const bool single = first == (int)r;
    const bool last_to_global = p.strided != 0;
    if (first == 1) ntt_round<1, INV>(sm, p.roots, s, tile_log, tt, tid, nthreads, io, true, single && last_to_global, stage);
    else if (first == 2) ntt_round<2, INV>(sm, p.roots, s, tile_log, tt, tid, nthreads, io, true, single && last_to_global, stage);
    else if (first == 3) ntt_round<3, INV>(sm, p.roots, s, tile_log, tt, tid, nthreads, io, true, single && last_to_global, stage);
    else ntt_round<4, INV>(sm, p.roots, s, tile_log, tt, tid, nthreads, io, true, single && last_to_global, stage);
    for (s -= first; s >= 0; s -= 4) {
        NTT_TILE_SYNC();
        ntt_round<4, INV>(sm, p.roots, s, tile_log, tt, tid, nthreads, io, false, s < 4 && last_to_global);
    }
    if (!last_to_global) {
        NTT_TILE_SYNC();
        for (unsigned e = tid; e < tile_elems; e += nthreads) io.dst[io.base + e] = sm[ntt_pad(e)];
    }
}

template <bool INV>
ZK_HD void ntt_pass_tile(const PassParams& p, size_t tile, size_t trans, uint64_t* sm, unsigned tid, unsigned nthreads) {
    const TileIO io = ntt_tile_io(p, tile, trans);
    ntt_pass_rounds<INV>(p, io, sm, tid, nthreads, nullptr);
}

}  // namespace zk
