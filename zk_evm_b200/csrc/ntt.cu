// Batched Goldilocks NTT for sm_100a.
//
// Replaces plonky2_field 1.0.0 fft.rs (`PolynomialValues::ifft`, `PolynomialCoeffs::{lde, coset_fft, coset_ifft}`),
// reached from /root/reference/evm_arithmetization/src/prover.rs:100-107 through PolynomialBatch::from_values.
//
// One kernel family: decimation-in-frequency, natural order in -> bit-reversed order out.  A transform of size 2^L is cut into
// passes over index digits (most significant first): strided passes of 8 bits (tile = 2^8 digit values x 16 contiguous 8-byte
// elements) until <= 12 bits remain for the final contiguous pass.  Inside a pass the digit is processed in rounds of four radix-2
// stages on 16 registers whose twiddles are powers of two (w_16 = 2^12 in Goldilocks: shifts, no multiplier), with one general
// twiddle per element between rounds and the inter-pass twiddle w_M^(j'*kd) (a cached table) on the way out — all of it in
// ntt_tile.cuh, which is host+device so that tests/test_ntt_tile_host.py runs the same tile code on the CPU.  Natural-order results
// (coefficients) are produced by a tiled bit-reversal permutation that also applies the 1/n (and coset) scaling.  The LDE with
// blow-up 2 is two size-n coset transforms (shifts g and g*w_2n) read from the same coefficient column:
// lde_br[h*n + p] = DIF_n(c_j (g w_2n^h)^j)[p].
#include "internal.h"
#include "ntt.h"
#include "ntt_tile.cuh"
#include <stdlib.h>

namespace zk {

// 80 registers per thread: three 256-thread blocks (or the equivalent in smaller ones) per SM
template <int THREADS, bool INV>
__global__ void __launch_bounds__(THREADS, 768 / THREADS) ntt_pass_kernel(PassParams p) {
    extern __shared__ uint64_t sm[];
    ntt_pass_tile<INV>(p, blockIdx.x, blockIdx.y, sm, threadIdx.x, THREADS);
}
template <int THREADS>
static void launch_pass(dim3 grid, size_t smem, cudaStream_t stream, const PassParams& q, bool inverse) {
    if (inverse) ntt_pass_kernel<THREADS, true><<<grid, THREADS, smem, stream>>>(q);
    else ntt_pass_kernel<THREADS, false><<<grid, THREADS, smem, stream>>>(q);
}

// ---- the same pass with the tile staged through shared memory by the TMA unit ------------------------------------------------
// A full tile is 2^12 elements = 256 rows of 16 contiguous elements (128 B): the 2^8 digit values of a strided pass (rows 2^mp
// elements apart in global memory) or 32 KB of consecutive elements in the final pass.  Thread r asks the TMA unit for row r
// (`cp.async.bulk.shared::cluster.global`, 128 B, completing on one mbarrier armed with the tile's 32 KB); the tile arrives in shared
// memory in tile-index order, unpadded (a bulk copy wants 16-byte aligned rows; the first round's reads walk it with consecutive
// lanes on consecutive words, which needs no padding), and the rounds then run as in ntt_pass_kernel with the padded working tile
// next to it.  No address arithmetic, no 16 dependent global loads per thread, no registers held across the load latency.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool INV>
__global__ void __launch_bounds__(256, 3) ntt_pass_tma_kernel(PassParams p) {
    extern __shared__ __align__(128) uint64_t sm[];
    uint64_t* stage = sm + NTT_TILE_WORDS + 2;                    // keeps the staging area 16-byte aligned (NTT_TILE_WORDS is even)
    uint64_t* mbar = sm + NTT_TILE_WORDS;
    const TileIO io = ntt_tile_io(p, blockIdx.x, blockIdx.y);
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8u << NTT_MAX_TILE_LOG) : "memory");
    {
        const unsigned row = threadIdx.x;                         // tile indices [16 row, 16 row + 16)
        const uint64_t* src = io.src + io.gaddr(row << 4);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                     ::"r"(smem_u32(stage + (row << 4))), "l"(src), "r"(bar) : "memory");
    }
    {   // every thread waits for phase 0 of the barrier: all 32 KB have landed
        asm volatile("{\n\t.reg .pred p;\n\tNTT_TMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra NTT_TMA_DONE;\n\tbra NTT_TMA_WAIT;\n\tNTT_TMA_DONE:\n\t}" ::"r"(bar) : "memory");
    }
    ntt_pass_rounds<INV>(p, io, sm, threadIdx.x, 256, stage);
}

// out[j] = c0 * base^j
__global__ void powers_kernel(uint64_t* out, size_t len, uint64_t base, uint64_t c0) {
    size_t chunk = 16;
    size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * chunk;
    if (start >= len) return;
    uint64_t v = gl_mul(c0, gl_pow(base, start));
    for (size_t i = 0; i < chunk && start + i < len; i++) { out[start + i] = v; v = gl_mul(v, base); }
}

// tab[kd * M' + j'] = w^(j' * kd), w = root of order M = 2^m
__global__ void interpass_kernel(uint64_t* tab, unsigned m, unsigned r, uint64_t w) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t M = (size_t)1 << m;
    if (idx >= M) return;
    unsigned mp = m - r;
    size_t kd = idx >> mp, jp = idx & (((size_t)1 << mp) - 1);
    tab[idx] = gl_pow(w, (kd * jp) & (M - 1));
}

// out[c][rev(p)] = in[c][p] * scale * (tab ? tab[rev(p)] : 1); tiled so both sides are coalesced
__global__ void __launch_bounds__(256) bitrev_permute_kernel(const uint64_t* in, size_t in_stride, uint64_t* out,
                                                             size_t out_stride, unsigned L, unsigned a, uint64_t scale,
                                                             const uint64_t* tab) {
    __shared__ uint64_t sm[32][33];
    const unsigned A = 1u << a;
    const unsigned midbits = L - 2 * a;
    const size_t mid = blockIdx.x;
    const uint64_t* src = in + blockIdx.y * in_stride;
    uint64_t* dst = out + blockIdx.y * out_stride;
    for (unsigned e = threadIdx.x; e < A * A; e += 256) {
        unsigned lo = e & (A - 1), hi = e >> a;
        sm[hi][lo] = src[((size_t)hi << (L - a)) | (mid << a) | lo];
    }
    __syncthreads();
    const size_t rmid = midbits ? (size_t)(__brev((unsigned)mid) >> (32 - midbits)) : 0;
    for (unsigned e = threadIdx.x; e < A * A; e += 256) {
        unsigned rhi = e & (A - 1), rlo = e >> a;
        unsigned hi = a ? (__brev(rhi) >> (32 - a)) : 0, lo = a ? (__brev(rlo) >> (32 - a)) : 0;
        size_t o = ((size_t)rlo << (L - a)) | (rmid << a) | rhi;
        uint64_t v = sm[hi][lo];
        if (scale != 1) v = gl_mul(v, scale);
        if (tab) v = gl_mul(v, tab[o]);
        dst[o] = v;
    }
}

// -------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------
static const uint64_t* get_roots(Ctx& c, bool inverse) {
    DevBuf& b = inverse ? c.ntt.roots_inv : c.ntt.roots_fwd;
    if (!b.get()) {
        size_t len = (size_t)1 << NTT_ROOT_LOG;
        b = DevBuf(&c, len * 8);
        uint64_t w = gl_root_of_unity(NTT_ROOT_LOG);
        if (inverse) w = gl_inv(w);
        powers_kernel<<<(unsigned)((len / 16 + 127) / 128), 128, 0, c.stream>>>(b.get(), len, w, 1);
        c.count_launch();
        c.check_launch("powers_kernel(roots)");
    }
    return b.get();
}

// out[j] = c0 * base^j, j < len (uncached)
void fill_powers(Ctx& c, uint64_t* out, size_t len, uint64_t base, uint64_t c0) {
    powers_kernel<<<(unsigned)(((len + 15) / 16 + 127) / 128), 128, 0, c.stream>>>(out, len, base, c0);
    c.count_launch();
    c.check_launch("powers_kernel");
}

const uint64_t* get_power_table(Ctx& c, uint64_t base, uint64_t c0, size_t len) {
    std::string key = "pow:" + std::to_string(base) + ":" + std::to_string(c0) + ":" + std::to_string(len);
    auto it = c.table_cache.find(key);
    if (it != c.table_cache.end()) return it->second.get();
    DevBuf b(&c, len * 8);
    powers_kernel<<<(unsigned)(((len + 15) / 16 + 127) / 128), 128, 0, c.stream>>>(b.get(), len, base, c0);
    c.count_launch();
    c.check_launch("powers_kernel");
    const uint64_t* p = b.get();
    c.table_cache.emplace(key, std::move(b));
    return p;
}

static const uint64_t* get_interpass(Ctx& c, unsigned m, unsigned r, bool inverse) {
    std::string key = std::string("ip:") + (inverse ? "i" : "f") + std::to_string(m) + ":" + std::to_string(r);
    auto it = c.table_cache.find(key);
    if (it != c.table_cache.end()) return it->second.get();
    size_t M = (size_t)1 << m;
    DevBuf b(&c, M * 8);
    uint64_t w = gl_root_of_unity(m);
    if (inverse) w = gl_inv(w);
    interpass_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c.stream>>>(b.get(), m, r, w);
    c.count_launch();
    c.check_launch("interpass_kernel");
    const uint64_t* p = b.get();
    c.table_cache.emplace(key, std::move(b));
    return p;
}

void ntt_dif(Ctx& c, const uint64_t* src, size_t src_stride, unsigned src_shift, uint64_t* dst, size_t dst_stride,
             size_t ntrans, unsigned L, bool inverse, const uint64_t* prescale0, const uint64_t* prescale1,
             unsigned prescale_mask) {
    if (ntrans == 0) return;
    KernelScope ks(c, KF_NTT, 16.0 * (double)ntrans * (double)((size_t)1 << L));
    if (L == 0) {
        // size-1 transforms: copy (with prescale == 1 at j = 0)
        for (size_t t = 0; t < ntrans; t++)
            ZK_CUDA(cudaMemcpyAsync(dst + t * dst_stride, src + (t >> src_shift) * src_stride, 8,
                                    cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    std::vector<unsigned> digits;
    ntt_plan_passes(L, digits);
    const uint64_t* roots = get_roots(c, inverse);
    unsigned m = L;
    for (size_t pi = 0; pi < digits.size(); pi++) {
        PassParams p;
        bool first = pi == 0, last = pi + 1 == digits.size();
        p.src = first ? src : dst;
        p.src_stride = first ? src_stride : dst_stride;
        p.src_shift = first ? src_shift : 0;
        p.dst = dst;
        p.dst_stride = dst_stride;
        p.log_n = L;
        p.m = m;
        p.r = digits[pi];
        p.roots = roots;
        p.prescale0 = first ? prescale0 : nullptr;
        p.prescale1 = first ? prescale1 : nullptr;
        p.prescale_mask = prescale_mask;
        p.strided = last ? 0 : 1;
        p.t = ntt_pass_t(L, m, p.r, last);
        p.interpass = last ? nullptr : get_interpass(c, m, p.r, inverse);
        unsigned tile_log = p.r + p.t;
        size_t tiles = (size_t)1 << (L - tile_log);
        size_t smem = (size_t)8 * ntt_pad(1u << tile_log);
        for (size_t t0 = 0; t0 < ntrans; t0 += 65535) {
            size_t cnt = ntrans - t0 < 65535 ? ntrans - t0 : 65535;
            PassParams q = p;
            // advance by whole transforms (t0 is a multiple of 65535 which is odd: keep src_shift semantics by
            // requiring src_shift == 0 when chunking)
            if (t0) {
                ZK_REQUIRE(q.src_shift == 0 && q.prescale_mask == 0, "ntt_dif: too many transforms for a shifted source");
                q.src += t0 * q.src_stride;
                q.dst += t0 * q.dst_stride;
            }
            dim3 grid((unsigned)tiles, (unsigned)cnt);
            // one radix-16 item (16 elements) per thread and round
            static const bool tma = [] { const char* e = getenv("ZKGPU_NTT_TMA"); return !(e && *e == '0'); }();
            if (tile_log == NTT_MAX_TILE_LOG && tma && (!q.strided || q.t == NTT_STRIDED_T)) {
                // full tiles of rows of >= 16 contiguous elements: staged by the TMA unit
                const size_t smem_tma = 8 * ((size_t)NTT_TILE_WORDS + 2 + (1u << NTT_MAX_TILE_LOG));
                static bool attr = false;
                if (!attr) {
                    ZK_CUDA(cudaFuncSetAttribute(ntt_pass_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
                    ZK_CUDA(cudaFuncSetAttribute(ntt_pass_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
                    attr = true;
                }
                if (inverse) ntt_pass_tma_kernel<true><<<grid, 256, smem_tma, c.stream>>>(q);
                else ntt_pass_tma_kernel<false><<<grid, 256, smem_tma, c.stream>>>(q);
            } else if (tile_log >= 12) launch_pass<256>(grid, smem, c.stream, q, inverse);
            else if (tile_log >= 11) launch_pass<128>(grid, smem, c.stream, q, inverse);
            else if (tile_log >= 10) launch_pass<64>(grid, smem, c.stream, q, inverse);
            else launch_pass<32>(grid, smem, c.stream, q, inverse);
            c.count_launch();
        }
        c.check_launch("ntt_pass_kernel");
        m -= p.r;
    }
}

void bitrev_permute(Ctx& c, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, size_t ncols,
                    unsigned L, uint64_t scale, const uint64_t* tab) {
    if (ncols == 0) return;
    KernelScope ks(c, KF_NTT, 0.0);   // part of the inverse transform: its algorithmic bytes are counted by ntt_dif
    ZK_REQUIRE(in != out, "bitrev_permute must be out of place");
    unsigned a = L / 2 < 5 ? L / 2 : 5;
    size_t mids = (size_t)1 << (L - 2 * a);
    for (size_t c0 = 0; c0 < ncols; c0 += 65535) {
        size_t cnt = ncols - c0 < 65535 ? ncols - c0 : 65535;
        dim3 grid((unsigned)mids, (unsigned)cnt);
        bitrev_permute_kernel<<<grid, 256, 0, c.stream>>>(in + c0 * in_stride, in_stride, out + c0 * out_stride,
                                                          out_stride, L, a, scale, tab);
        c.count_launch();
    }
    c.check_launch("bitrev_permute_kernel");
}

// values (natural) -> coefficients (natural).  `scratch` holds ncols*n elements and may alias nothing else.
void intt_natural(Ctx& c, const uint64_t* values, uint64_t* scratch, uint64_t* coeffs, size_t ncols, unsigned L,
                  uint64_t coset_shift) {
    size_t n = (size_t)1 << L;
    ntt_dif(c, values, n, 0, scratch, n, ncols, L, true, nullptr, nullptr, 0);
    uint64_t ninv = gl_inv(gl_canon(n));
    const uint64_t* tab = nullptr;
    if (coset_shift > 1) tab = get_power_table(c, gl_inv(coset_shift), 1, n);
    bitrev_permute(c, scratch, n, coeffs, n, ncols, L, ninv, tab);
}

// coefficients (natural, n per column) -> LDE evaluations on g*<w_N>, N = n << rate_bits, bit-reversed order.
// lde column stride is N.
void lde_bitrev(Ctx& c, const uint64_t* coeffs, uint64_t* lde, size_t ncols, unsigned L, unsigned rate_bits,
                uint64_t shift) {
    ZK_REQUIRE(rate_bits <= 4, "blow-up factors up to 16 are implemented (StarkConfig rate_bits = 1, plonky2 recursion configs rate_bits = 3)");
    size_t n = (size_t)1 << L;
    if (rate_bits == 0) {
        const uint64_t* t0 = shift > 1 ? get_power_table(c, shift, 1, n) : nullptr;
        ntt_dif(c, coeffs, n, 0, lde, n, ncols, L, false, t0, nullptr, 0);
        return;
    }
    if (rate_bits == 1) {
        // half h evaluates on the coset (shift * w_2n^h) * <w_n>: both halves in one launch per pass, read from the same coefficient column
        const uint64_t* t0 = get_power_table(c, shift, 1, n);
        const uint64_t* t1 = get_power_table(c, gl_mul(shift, gl_root_of_unity(L + 1)), 1, n);
        ntt_dif(c, coeffs, n, 1, lde, n, 2 * ncols, L, false, t0, t1, 1);
        return;
    }
    // Blow-up 2^r in general (the widths plonky2's recursion circuits commit with: PolynomialBatch::from_values(.., rate_bits = 3, ..)).
    // Natural index i = c + 2^r m of the size-N domain is the point (shift w_N^c) w_n^m, and it is stored at bitrev_N(i) = bitrev_r(c) n +
    // bitrev_n(m): block h of n storage positions is the size-n transform on the coset shift * w_N^bitrev_r(h), natural in, bit-reversed out.
    const size_t N = n << rate_bits;
    const uint64_t wN = gl_root_of_unity(L + rate_bits);
    for (unsigned h = 0; h < (1u << rate_bits); h++) {
        const uint64_t sh = gl_mul(shift, gl_pow(wN, bitrev32(h, rate_bits)));
        ntt_dif(c, coeffs, n, 0, lde + (size_t)h * n, N, ncols, L, false, get_power_table(c, sh, 1, n), nullptr, 0);
    }
}

}  // namespace zk
