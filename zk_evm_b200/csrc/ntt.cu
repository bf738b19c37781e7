// Batched Goldilocks NTT for sm_100a.
//
// Replaces plonky2_field 1.0.0 fft.rs (`PolynomialValues::ifft`, `PolynomialCoeffs::{lde, coset_fft, coset_ifft}`),
// reached from /root/reference/evm_arithmetization/src/prover.rs:100-107 through PolynomialBatch::from_values.
//
// One kernel family: decimation-in-frequency, natural order in -> bit-reversed order out, mixed radix.
// A transform of size 2^L is cut into passes over index digits (most significant first).  A pass stages a tile of
// 2^r (digit) x 2^t (contiguous 8-byte elements, coalesced) in shared memory, runs the r radix-2 stages there with
// pure w_{2^r} twiddles, multiplies by the inter-pass twiddle w_M^(j'*kd) (a cached table, coalesced) and writes
// the tile back in place.  Natural-order results (coefficients) are produced by a tiled bit-reversal permutation
// that also applies the 1/n (and coset) scaling.  The LDE with blow-up 2 is two size-n coset transforms
// (shifts g and g*w_2n) read from the same coefficient column: lde_br[h*n + p] = DIF_n(c_j (g w_2n^h)^j)[p].
#include "internal.h"
#include "ntt.h"

namespace zk {

static constexpr unsigned ROOT_LOG = 13;      // roots tables hold w_{2^13}^k, k < 2^12
static constexpr unsigned MAX_TILE_LOG = 12;  // 4096 elements = 32 KB shared memory per CTA
static constexpr unsigned MAX_STRIDED_R = 8;
static constexpr unsigned STRIDED_T = 4;      // 16 contiguous elements = 128 B per row of a strided tile

struct PassParams {
    const uint64_t* src;
    uint64_t* dst;
    size_t src_stride, dst_stride;   // elements between consecutive transforms
    unsigned src_shift;              // transform t reads source column t >> src_shift
    unsigned log_n, m, r, t;
    const uint64_t* roots;           // w_{2^ROOT_LOG}^(+-k)
    const uint64_t* interpass;       // [kd * M' + j'] or nullptr
    const uint64_t* prescale0;       // per natural index j, or nullptr
    const uint64_t* prescale1;
    unsigned prescale_mask;          // table = (t & mask) ? prescale1 : prescale0
};

// K consecutive DIF stages (halves 2^s ... 2^(s-K+1)) of the size-2^r transforms held in sm[jd * 2^t + u]: every work item
// loads the 2^K elements jd = (hi << (s+1)) | (k << (s-K+1)) | lo, k < 2^K, runs the K stages in registers and writes them
// back, so shared memory is touched once per K stages.  Items are numbered u fastest, then lo, then hi: for a fixed k a
// warp reads 32 consecutive elements whenever 2^t * 2^(s-K+1) >= 32.
template <int K, int THREADS>
__device__ __forceinline__ void dif_round(uint64_t* sm, const uint64_t* __restrict__ roots, int s, unsigned r, unsigned t) {
    constexpr int E = 1 << K;
    const int low = s - K + 1;
    const unsigned items = (1u << (r - K)) << t;
    for (unsigned w = threadIdx.x; w < items; w += THREADS) {
        const unsigned u = w & ((1u << t) - 1), g = w >> t;
        const unsigned lo = g & ((1u << low) - 1), hi = g >> low;
        const unsigned base = ((((hi << K) << low) | lo) << t) + u;     // element k lives at base + (k << (low + t))
        uint64_t x[E];
#pragma unroll
        for (int k = 0; k < E; k++) x[k] = sm[base + ((unsigned)k << (low + t))];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int lh = s - j;                  // this stage's log2(half), in units of jd
            constexpr int dummy = 0; (void)dummy;
            const int hk = 1 << (K - 1 - j);       // the partner's distance in k
#pragma unroll
            for (int k = 0; k < E; k++) {
                if (k & hk) continue;
                uint64_t a = x[k], b = x[k + hk];
                x[k] = gl_add(a, b);
                uint64_t d = gl_sub(a, b);
                if (lh > 0) {
                    const unsigned j_in = ((unsigned)(k & (hk - 1)) << low) | lo;
                    d = gl_mul(d, __ldg(roots + ((size_t)j_in << (ROOT_LOG - 1 - lh))));
                }
                x[k + hk] = d;
            }
        }
#pragma unroll
        for (int k = 0; k < E; k++) sm[base + ((unsigned)k << (low + t))] = x[k];
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) ntt_dif_pass_kernel(PassParams p) {
    extern __shared__ uint64_t sm[];
    const unsigned r = p.r, t = p.t, m = p.m;
    const unsigned R = 1u << r, T = 1u << t;
    const unsigned tile_elems = R << t;
    const size_t trans = blockIdx.y;
    const uint64_t* src = p.src + (trans >> p.src_shift) * p.src_stride;
    uint64_t* dst = p.dst + trans * p.dst_stride;
    const uint64_t* prescale = p.prescale0 ? ((trans & p.prescale_mask) ? p.prescale1 : p.prescale0) : nullptr;
    const size_t tile = blockIdx.x;
    const unsigned mp = m - r;   // log M'
    const bool strided = mp >= t;   // else: final pass (mp == 0), tile = 2^t consecutive blocks of R

    // ---- load ----
    size_t base;
    if (strided) {
        // tile id = hi * 2^(mp - t) + lo_hi
        size_t hi = tile >> (mp - t), lo_hi = tile & (((size_t)1 << (mp - t)) - 1);
        base = (hi << m) + (lo_hi << t);
        for (unsigned e = threadIdx.x; e < tile_elems; e += THREADS) {
            unsigned lo_t = e & (T - 1), jd = e >> t;
            size_t g = base + ((size_t)jd << mp) + lo_t;
            uint64_t v = src[g];
            if (prescale) v = gl_mul(v, prescale[g]);
            sm[e] = v;   // layout [jd][lo_t]
        }
    } else {
        base = tile * (size_t)tile_elems;
        for (unsigned e = threadIdx.x; e < tile_elems; e += THREADS) {
            unsigned jd = e & (R - 1), u = e >> r;
            size_t g = base + e;
            uint64_t v = src[g];
            if (prescale) v = gl_mul(v, prescale[g]);
            sm[(jd << t) + u] = v;   // layout [jd][u]; (bank-conflicted store, r stages amortise it)
        }
    }
    __syncthreads();

    // ---- r radix-2 DIF stages over jd, K (<= 3) stages per shared-memory round trip on a register tile of 2^K elements ----
    {
        int s = (int)r - 1;                       // log2(half) of the next stage
        const int first = (r % 3) ? (int)(r % 3) : 3;
        if (first == 1) dif_round<1, THREADS>(sm, p.roots, s, r, t);
        else if (first == 2) dif_round<2, THREADS>(sm, p.roots, s, r, t);
        else dif_round<3, THREADS>(sm, p.roots, s, r, t);
        __syncthreads();
        for (s -= first; s >= 0; s -= 3) {
            dif_round<3, THREADS>(sm, p.roots, s, r, t);
            __syncthreads();
        }
    }

    // ---- inter-pass twiddle + store (in place: digit position jd_pos holds kd = rev_r(jd_pos)) ----
    if (strided) {
        const size_t lo_base = (tile & (((size_t)1 << (mp - t)) - 1)) << t;
        for (unsigned e = threadIdx.x; e < tile_elems; e += THREADS) {
            unsigned lo_t = e & (T - 1), jd = e >> t;
            uint64_t v = sm[e];
            if (p.interpass) {
                unsigned kd = __brev(jd) >> (32 - r);
                v = gl_mul(v, __ldg(p.interpass + ((size_t)kd << mp) + lo_base + lo_t));
            }
            dst[base + ((size_t)jd << mp) + lo_t] = v;
        }
    } else {
        for (unsigned e = threadIdx.x; e < tile_elems; e += THREADS) {
            unsigned jd = e & (R - 1), u = e >> r;
            dst[base + e] = sm[(jd << t) + u];
        }
    }
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
        size_t len = (size_t)1 << (ROOT_LOG - 1);
        b = DevBuf(&c, len * 8);
        uint64_t w = gl_root_of_unity(ROOT_LOG);
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

// digit plan, most significant digit first; the last entry is the final (contiguous) pass
static void plan_passes(unsigned L, std::vector<unsigned>& digits) {
    digits.clear();
    if (L <= MAX_TILE_LOG) { digits.push_back(L); return; }
    if (L <= MAX_TILE_LOG + MAX_STRIDED_R) {
        // two passes: keep both tiles large (r1 + STRIDED_T and L - r1 close to MAX_TILE_LOG)
        unsigned r1 = L - MAX_TILE_LOG;
        unsigned half = L / 2 < MAX_STRIDED_R ? L / 2 : MAX_STRIDED_R;
        if (r1 < half) r1 = half;
        digits.push_back(r1);
        digits.push_back(L - r1);
        return;
    }
    unsigned rest = L - MAX_TILE_LOG;
    unsigned ns = (rest + MAX_STRIDED_R - 1) / MAX_STRIDED_R;
    unsigned per = rest / ns, extra = rest % ns;
    for (unsigned i = 0; i < ns; i++) digits.push_back(per + (i < extra ? 1 : 0));
    digits.push_back(MAX_TILE_LOG);
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
    plan_passes(L, digits);
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
        if (last) {
            // final pass: m == r; fill the tile with 2^t consecutive blocks
            unsigned t = MAX_TILE_LOG > p.r ? MAX_TILE_LOG - p.r : 0;
            if (t > L - m) t = L - m;
            p.t = t;
            p.interpass = nullptr;
        } else {
            p.t = STRIDED_T;
            p.interpass = get_interpass(c, m, p.r, inverse);
        }
        unsigned tile_log = p.r + p.t;
        size_t tiles = (size_t)1 << (L - tile_log);
        size_t smem = ((size_t)8) << tile_log;
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
            if (tile_log >= 10)
                ntt_dif_pass_kernel<256><<<grid, 256, smem, c.stream>>>(q);
            else
                ntt_dif_pass_kernel<64><<<grid, 64, smem, c.stream>>>(q);
            c.count_launch();
        }
        c.check_launch("ntt_dif_pass_kernel");
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
    ZK_REQUIRE(rate_bits <= 1, "only blow-up 1 or 2 is implemented (StarkConfig rate_bits = 1)");
    size_t n = (size_t)1 << L;
    if (rate_bits == 0) {
        const uint64_t* t0 = shift > 1 ? get_power_table(c, shift, 1, n) : nullptr;
        ntt_dif(c, coeffs, n, 0, lde, n, ncols, L, false, t0, nullptr, 0);
        return;
    }
    // half h evaluates on the coset (shift * w_2n^h) * <w_n>
    const uint64_t* t0 = get_power_table(c, shift, 1, n);
    const uint64_t* t1 = get_power_table(c, gl_mul(shift, gl_root_of_unity(L + 1)), 1, n);
    ntt_dif(c, coeffs, n, 1, lde, n, 2 * ncols, L, false, t0, t1, 1);
}

}  // namespace zk
