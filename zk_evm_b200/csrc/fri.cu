// Openings and the FRI commit phase on the device.
//
// Replaces starky 1.0.0 proof.rs `StarkOpeningSet::new` and plonky2 1.0.0 fri/oracle.rs `PolynomialBatch::prove_openings`
// + fri/prover.rs `fri_committed_trees` / `fri_proof_of_work` / `fri_prover_query_rounds`, reached from
// /root/reference/evm_arithmetization/src/prover.rs:322-334 through prove_with_commitment.
//
// Differences in HOW (results are identical, the arithmetic is exact):
//  * openings: every coefficient column is read once; zeta, g*zeta and 1 are evaluated in the same pass
//    (thread t takes coefficients t, t+T, ... of a segment: Horner in zeta^T, then a block reduction).
//  * batch reduction: the reference combines coefficient vectors (sum_j alpha^j f_j), runs synthetic division by
//    (X - z) per batch and then a size-2n coset FFT.  Here the same polynomial's VALUES on the coset are produced
//    directly from the LDE columns that are already resident: v(x) = sum_b alpha^(..) (sum_j alpha^j f_j(x) - sum_j alpha^j
//    f_j(z_b)) / (x - z_b), one streaming pass, no scan; its coefficients come from one small inverse NTT.
//  * proof of work: grid search with atomicMin -> the smallest valid witness (the reference's rayon find_any
//    returns an arbitrary one; zkgpu_prove_table accepts a forced witness to reproduce a given proof).
#include "stark_dev.h"
#include <algorithm>
#include "poseidon_fast.cuh"

namespace zk {

// ---- openings ---------------------------------------------------------------------------------------------------
// f(zeta), f(g zeta) and f(1) for every coefficient column.  The coefficients are base-field elements, so with a table of the
// powers zeta^i and (g zeta)^i (built once per table proof, shared by all columns, L2-resident) each coefficient costs four
// 64x64 products accumulated WITHOUT reduction: the partial products a0*b0, a0*b1 + a1*b0, a1*b1 of every term go to three
// 96-bit accumulators per component (no carry counters: a 96-bit accumulator holds 2^32 64-bit terms), reduced once per
// thread, then a block reduction.  The previous form (extension Horner in zeta^T per thread) cost ~10 field products per coefficient.
static constexpr int EVAL_THREADS = 256;
static constexpr int EVAL_K = 16;                 // coefficient indices per thread
static constexpr int EVAL_G = 2;                  // columns per block

struct Acc96 { uint32_t w0, w1, w2; };
__device__ __forceinline__ void acc96_add(Acc96& a, uint64_t v) {
    uint32_t lo, hi; gl_unpack(v, lo, hi);
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, 0;\n\t" : "+r"(a.w0), "+r"(a.w1), "+r"(a.w2) : "r"(lo), "r"(hi));
}
// sum_i c_i * b_i for 64-bit c, b: lo += c0*b0, mid += c0*b1 + c1*b0 (weight 2^32), hi += c1*b1 (weight 2^64)
struct Dot { Acc96 lo, mid, hi; };
__device__ __forceinline__ void dot_init(Dot& d) { d.lo = {0, 0, 0}; d.mid = {0, 0, 0}; d.hi = {0, 0, 0}; }
__device__ __forceinline__ void dot_mac(Dot& d, uint32_t c0, uint32_t c1, uint64_t b) {
    uint32_t b0, b1; gl_unpack(b, b0, b1);
    uint64_t p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(c0), "r"(b0)); acc96_add(d.lo, p);
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(c0), "r"(b1)); acc96_add(d.mid, p);
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(c1), "r"(b0)); acc96_add(d.mid, p);
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(c1), "r"(b1)); acc96_add(d.hi, p);
}
// a 96-bit value mod p: w0 + 2^32 w1 + 2^64 w2
__device__ __forceinline__ uint64_t acc96_reduce(const Acc96& a) { return gld_reduce_words(a.w0, a.w1, a.w2, 0); }
__device__ __forceinline__ uint64_t dot_reduce(const Dot& d) {
    // lo + 2^32 mid + 2^64 hi ;  2^32 and 2^64 == 2^32 - 1 are field constants
    uint64_t r = acc96_reduce(d.lo);
    r = gl_add(r, gl_mul(acc96_reduce(d.mid), (uint64_t)1 << 32));
    r = gl_add(r, gl_mul(acc96_reduce(d.hi), GL_EPS));
    return r;
}

// tab[i] = (zeta^i.a, zeta^i.b, (g zeta)^i.a, (g zeta)^i.b), i < n; each thread starts from a power and walks 16 steps
__global__ void eval_powers_kernel(ulonglong4* __restrict__ tab, size_t n, Fp2 zeta, Fp2 zeta_next) {
    size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (start >= n) return;
    Fp2 a = fp2_pow(zeta, start), b = fp2_pow(zeta_next, start);
    for (size_t i = start; i < start + 16 && i < n; i++) {
        tab[i] = make_ulonglong4(a.a, a.b, b.a, b.b);
        a = a * zeta; b = b * zeta_next;
    }
}

__global__ void __launch_bounds__(EVAL_THREADS, 2) eval_columns_kernel(const uint64_t* __restrict__ coeffs, size_t ncols, size_t n,
                                                                    const ulonglong4* __restrict__ tab, uint64_t* __restrict__ partial,
                                                                    size_t nseg) {
    __shared__ uint64_t sm[EVAL_G * 5][EVAL_THREADS / 32];
    const size_t seg = blockIdx.x, col0 = (size_t)blockIdx.y * EVAL_G;
    const unsigned t = threadIdx.x;
    const size_t seg_len = (size_t)EVAL_THREADS * EVAL_K;
    Dot d[EVAL_G][4];
    Acc96 one[EVAL_G];
#pragma unroll
    for (int g = 0; g < EVAL_G; g++) { one[g] = {0, 0, 0}; for (int q = 0; q < 4; q++) dot_init(d[g][q]); }
    // the operands of iteration k + 1 are requested before the products of iteration k are formed
    const size_t i0 = seg * seg_len + t;
    ulonglong2 pw0 = make_ulonglong2(0, 0), pw1 = make_ulonglong2(0, 0);
    uint64_t cv[EVAL_G];
#pragma unroll
    for (int g = 0; g < EVAL_G; g++) cv[g] = 0;
    if (i0 < n) {
        pw0 = __ldg(reinterpret_cast<const ulonglong2*>(tab + i0)); pw1 = __ldg(reinterpret_cast<const ulonglong2*>(tab + i0) + 1);
#pragma unroll
        for (int g = 0; g < EVAL_G; g++) if (col0 + g < ncols) cv[g] = __ldg(coeffs + (col0 + g) * n + i0);
    }
#pragma unroll 1
    for (int k = 0; k < EVAL_K; k++) {
        const size_t i = i0 + (size_t)k * EVAL_THREADS;
        if (i >= n) break;
        const ulonglong2 q0 = pw0, q1 = pw1;
        uint64_t cc[EVAL_G];
#pragma unroll
        for (int g = 0; g < EVAL_G; g++) cc[g] = cv[g];
        const size_t inx = i + EVAL_THREADS;
        if (k + 1 < EVAL_K && inx < n) {
            pw0 = __ldg(reinterpret_cast<const ulonglong2*>(tab + inx)); pw1 = __ldg(reinterpret_cast<const ulonglong2*>(tab + inx) + 1);
#pragma unroll
            for (int g = 0; g < EVAL_G; g++) if (col0 + g < ncols) cv[g] = __ldg(coeffs + (col0 + g) * n + inx);
        }
#pragma unroll
        for (int g = 0; g < EVAL_G; g++) {
            if (col0 + g >= ncols) break;
            uint32_t c0, c1; gl_unpack(cc[g], c0, c1);
            dot_mac(d[g][0], c0, c1, q0.x); dot_mac(d[g][1], c0, c1, q0.y);
            dot_mac(d[g][2], c0, c1, q1.x); dot_mac(d[g][3], c0, c1, q1.y);
            acc96_add(one[g], cc[g]);
        }
    }
    // per-thread reduction to field elements, warp shuffle tree, then one value per warp through shared memory
#pragma unroll
    for (int g = 0; g < EVAL_G; g++) {
#pragma unroll
        for (int q = 0; q < 5; q++) {
            uint64_t v = q < 4 ? dot_reduce(d[g][q]) : acc96_reduce(one[g]);
            for (int off = 16; off > 0; off >>= 1) v = gl_add(v, __shfl_down_sync(0xFFFFFFFFu, v, off));
            if ((t & 31) == 0) sm[g * 5 + q][t >> 5] = v;
        }
    }
    __syncthreads();
    if (t < EVAL_G * 5) {
        uint64_t v = 0;
        for (int w = 0; w < EVAL_THREADS / 32; w++) v = gl_add(v, sm[t][w]);
        size_t col = col0 + t / 5;
        if (col < ncols) partial[(col * nseg + seg) * 5 + (t % 5)] = v;
    }
}

void eval_columns(Ctx& c, const uint64_t* coeffs, size_t ncols, size_t n, Fp2 zeta, Fp2 zeta_next, std::vector<uint64_t>& out5) {
    KernelScope ks(c, KF_OPENINGS, 8.0 * n * ncols);
    out5.assign(ncols * 5, 0);
    if (!ncols) return;
    // power table: cached per (zeta, n) for the three calls (trace, aux, quotient) of one table proof
    std::string key = "evalpow";
    auto it = c.table_cache.find(key);
    if (it == c.table_cache.end() || c.evalpow_n != n || !(c.evalpow_zeta[0] == zeta.a && c.evalpow_zeta[1] == zeta.b)) {
        if (it != c.table_cache.end()) c.table_cache.erase(it);
        DevBuf b(&c, n * 32);
        eval_powers_kernel<<<(unsigned)(((n + 15) / 16 + 127) / 128), 128, 0, c.stream>>>((ulonglong4*)b.get(), n, zeta, zeta_next);
        c.count_launch();
        c.check_launch("eval_powers_kernel");
        it = c.table_cache.emplace(key, std::move(b)).first;
        c.evalpow_n = n; c.evalpow_zeta[0] = zeta.a; c.evalpow_zeta[1] = zeta.b;
    }
    const ulonglong4* tab = (const ulonglong4*)it->second.get();
    const size_t seg_len = (size_t)EVAL_THREADS * EVAL_K;
    size_t nseg = (n + seg_len - 1) / seg_len;
    DevBuf partial(&c, ncols * nseg * 5 * 8);
    size_t groups = (ncols + EVAL_G - 1) / EVAL_G;
    for (size_t g0 = 0; g0 < groups; g0 += 65535) {
        unsigned cnt = (unsigned)std::min<size_t>(65535, groups - g0);
        dim3 grid((unsigned)nseg, cnt);
        eval_columns_kernel<<<grid, EVAL_THREADS, 0, c.stream>>>(coeffs + g0 * EVAL_G * n, ncols - g0 * EVAL_G, n, tab,
                                                                 partial.get() + g0 * EVAL_G * nseg * 5, nseg);
        c.count_launch();
    }
    c.check_launch("eval_columns_kernel");
    std::vector<uint64_t> h(ncols * nseg * 5);
    c.d2h(h.data(), partial.get(), h.size() * 8);
    for (size_t col = 0; col < ncols; col++)
        for (size_t s = 0; s < nseg; s++)
            for (int q = 0; q < 5; q++) out5[col * 5 + q] = gl_add(out5[col * 5 + q], h[(col * nseg + s) * 5 + q]);
}

__device__ __forceinline__ Fp2 fp2_mul_base(Fp2 x, uint64_t s) { return Fp2(gl_mul(x.a, s), gl_mul(x.b, s)); }

// ---- batch reduction straight to coset values --------------------------------------------------------------------
struct CombineKernelArgs {
    const uint64_t* lde[3];
    uint32_t ncols[3];
    uint32_t zs_begin;
    uint32_t log_N;
    const uint64_t* apow;     // alpha^j as (a, b) pairs, j < ncols[0]+ncols[1]+ncols[2]
    Fp2 zeta, zeta_next, v0, v1, v2;
    Fp2 shift1, shift2;       // alpha^(#polys of batch 1), alpha^(#polys of batch 2)
    uint64_t w_N;
    int has_b2;
    uint64_t* out_re; uint64_t* out_im;
};

// sum_j alpha^j f_j(x) per batch: extension x base products, accumulated WITHOUT reduction in three 96-bit accumulators per
// component (Dot above: 4 IMAD.WIDE + 12 carry adds per product instead of a full multiply-reduce-add), reduced once per point.
// The three denominators (x - zeta), (x - g zeta) [extension] and (x - 1) share ONE Fermat inversion: 1/(x - z) = conj / norm with
// norm in the base field, and the three base-field values are inverted together (Montgomery's trick).
struct DotExt { Dot re, im; };
__device__ __forceinline__ void dotext_init(DotExt& d) { dot_init(d.re); dot_init(d.im); }
__device__ __forceinline__ void dotext_mac(DotExt& d, uint64_t v, const uint64_t* __restrict__ ap) {
    uint32_t v0, v1; gl_unpack(v, v0, v1);
    const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2*>(ap));
    dot_mac(d.re, v0, v1, w.x); dot_mac(d.im, v0, v1, w.y);
}
__device__ __forceinline__ Fp2 dotext_reduce(const DotExt& d) { return Fp2(dot_reduce(d.re), dot_reduce(d.im)); }

__global__ void __launch_bounds__(256) fri_combine_kernel(CombineKernelArgs a) {
    const size_t N = (size_t)1 << a.log_N;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const uint32_t i = bitrev32((uint32_t)j, a.log_N);
    const uint64_t x = gl_mul(GL_GENERATOR, gl_pow(a.w_N, i));
    DotExt d_ta, d_q, d_z;
    dotext_init(d_ta); dotext_init(d_q); dotext_init(d_z);
    uint32_t k = 0;
#pragma unroll 8
    for (uint32_t c = 0; c < a.ncols[0]; c++, k++) dotext_mac(d_ta, __ldg(a.lde[0] + (size_t)c * N + j), a.apow + 2 * k);
    {
        const uint32_t nz = a.zs_begin < a.ncols[1] ? a.zs_begin : a.ncols[1];
#pragma unroll 4
        for (uint32_t c = 0; c < nz; c++, k++) dotext_mac(d_ta, __ldg(a.lde[1] + (size_t)c * N + j), a.apow + 2 * k);
        for (uint32_t c = nz; c < a.ncols[1]; c++, k++) {
            const uint64_t v = __ldg(a.lde[1] + (size_t)c * N + j);
            dotext_mac(d_ta, v, a.apow + 2 * k);
            dotext_mac(d_z, v, a.apow + 2 * (c - a.zs_begin));
        }
    }
    for (uint32_t c = 0; c < a.ncols[2]; c++, k++) dotext_mac(d_q, __ldg(a.lde[2] + (size_t)c * N + j), a.apow + 2 * k);
    const Fp2 s_ta = dotext_reduce(d_ta), s_q = dotext_reduce(d_q), s_z = dotext_reduce(d_z);
    // 1 / (x - z) for z = zeta, g zeta (extension) and 1 (base) from one inversion
    const Fp2 e0 = Fp2(x, 0) - a.zeta, e1 = Fp2(x, 0) - a.zeta_next;
    const uint64_t n0 = gl_sub(gl_sqr(e0.a), gl_mul7(gl_sqr(e0.b))), n1 = gl_sub(gl_sqr(e1.a), gl_mul7(gl_sqr(e1.b)));
    const uint64_t n2 = a.has_b2 ? gl_sub(x, 1) : 1;
    const uint64_t n01 = gl_mul(n0, n1);
    const uint64_t inv_all = gl_inv(gl_mul(n01, n2));
    const uint64_t i2 = gl_mul(inv_all, n01), i01 = gl_mul(inv_all, n2);
    const uint64_t i0 = gl_mul(i01, n1), i1 = gl_mul(i01, n0);
    const Fp2 inv0(gl_mul(e0.a, i0), gl_mul(gl_neg(e0.b), i0)), inv1(gl_mul(e1.a, i1), gl_mul(gl_neg(e1.b), i1));
    Fp2 q0 = (s_ta + s_q - a.v0) * inv0;
    Fp2 q1 = (s_ta - a.v1) * inv1;
    Fp2 fin = q0 * a.shift1 + q1;
    if (a.has_b2) {
        Fp2 q2 = fp2_mul_base(s_z - a.v2, i2);
        fin = fin * a.shift2 + q2;
    }
    a.out_re[j] = fin.a;
    a.out_im[j] = fin.b;
}

void fri_combine(Ctx& c, const CombineArgs& a) {
    KernelScope ks(c, KF_FRI, 8.0 * ((size_t)1 << a.log_N) * (a.ncols[0] + a.ncols[1] + a.ncols[2] + 2));
    size_t total = a.ncols[0] + a.ncols[1] + a.ncols[2];
    std::vector<uint64_t> apow(2 * total);
    Fp2 p(1, 0);
    for (size_t k = 0; k < total; k++) { apow[2 * k] = p.a; apow[2 * k + 1] = p.b; p = p * a.alpha; }
    DevBuf dap(&c, apow.size() * 8 + 16);
    c.h2d(dap.get(), apow.data(), apow.size() * 8);
    CombineKernelArgs k;
    for (int o = 0; o < 3; o++) { k.lde[o] = a.lde[o]; k.ncols[o] = (uint32_t)a.ncols[o]; }
    k.zs_begin = (uint32_t)a.zs_begin; k.log_N = a.log_N; k.apow = dap.get();
    k.zeta = a.zeta; k.zeta_next = a.zeta_next; k.v0 = a.v0; k.v1 = a.v1; k.v2 = a.v2;
    k.shift1 = fp2_pow(a.alpha, a.ncols[0] + a.ncols[1]);
    k.shift2 = fp2_pow(a.alpha, a.ncols[1] - a.zs_begin);
    k.w_N = gl_root_of_unity(a.log_N);
    k.has_b2 = a.has_b2 ? 1 : 0;
    k.out_re = a.out_re; k.out_im = a.out_im;
    size_t N = (size_t)1 << a.log_N;
    fri_combine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c.stream>>>(k);
    c.count_launch();
    c.check_launch("fri_combine_kernel");
    c.sync();
}

// ---- commit-phase layers -----------------------------------------------------------------------------------------
__global__ void fri_leaves_kernel(const uint64_t* __restrict__ re, const uint64_t* __restrict__ im, size_t rows, unsigned arity_bits,
                                  uint64_t* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t width = (size_t)2 << arity_bits;
    if (idx >= rows * width) return;
    size_t k = idx / rows, r = idx % rows;
    const uint64_t* src = (k & 1) ? im : re;
    out[idx] = src[(r << arity_bits) + (k >> 1)];
}
void fri_leaves(Ctx& c, const uint64_t* re, const uint64_t* im, size_t M, unsigned arity_bits, uint64_t* out) {
    KernelScope ks(c, KF_FRI, 32.0 * M);
    size_t rows = M >> arity_bits, total = rows * ((size_t)2 << arity_bits);
    fri_leaves_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(re, im, rows, arity_bits, out);
    c.count_launch();
    c.check_launch("fri_leaves_kernel");
}

__global__ void fri_fold_kernel(const uint64_t* __restrict__ re, const uint64_t* __restrict__ im, size_t out_len, unsigned arity_bits,
                                Fp2 beta, uint64_t* __restrict__ out_re, uint64_t* __restrict__ out_im) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_len) return;
    size_t arity = (size_t)1 << arity_bits;
    Fp2 acc(0, 0);
    for (size_t t = arity; t-- > 0;) acc = acc * beta + Fp2(re[i * arity + t], im[i * arity + t]);
    out_re[i] = acc.a;
    out_im[i] = acc.b;
}
void fri_fold(Ctx& c, const uint64_t* re, const uint64_t* im, size_t M, unsigned arity_bits, Fp2 beta, uint64_t* out_re,
              uint64_t* out_im) {
    KernelScope ks(c, KF_FRI, 16.0 * M + 16.0 * (M >> arity_bits));
    size_t out_len = M >> arity_bits;
    fri_fold_kernel<<<(unsigned)((out_len + 127) / 128), 128, 0, c.stream>>>(re, im, out_len, arity_bits, beta, out_re, out_im);
    c.count_launch();
    c.check_launch("fri_fold_kernel");
}

// ---- proof of work -------------------------------------------------------------------------------------------------
struct PowState { uint64_t s[12]; };
__global__ void __launch_bounds__(128) pow_kernel(PowState st, unsigned pos, unsigned bits, uint64_t base, unsigned long long* best) {
    uint64_t cand = base + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cand >= GL_P) return;
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = st.s[i];
#pragma unroll
    for (int i = 0; i < 12; i++) if (i == (int)pos) s[i] = cand;
    pf_permute(s);
    if (bits == 0 || (pf_canon(s[7]) >> (64 - bits)) == 0) atomicMin(best, (unsigned long long)cand);
}
uint64_t pow_grind(Ctx& c, const uint64_t state[12], unsigned pos, unsigned bits) {
    ZK_REQUIRE(pos < 8 && bits < 40, "pow_grind: bad arguments");
    KernelScope ks(c, KF_POW, 0.0);
    PowState st;
    memcpy(st.s, state, 96);
    DevBuf best(&c, 8);
    unsigned long long init = ~0ULL;
    c.h2d(best.get(), &init, 8);
    // expected number of candidates is 2^bits: a batch of 2 * 2^bits (one wave of the GPU for 16 bits) finds a witness with
    // probability 1 - e^-2; batches are scanned in order, so the result is still the smallest witness
    const uint64_t batch = std::max<uint64_t>((uint64_t)1 << 14, std::min<uint64_t>((uint64_t)2 << bits, (uint64_t)1 << 22));
    for (uint64_t base = 0;; base += batch) {
        pow_kernel<<<(unsigned)(batch / 128), 128, 0, c.stream>>>(st, pos, bits, base, (unsigned long long*)best.get());
        c.count_launch();
        c.check_launch("pow_kernel");
        unsigned long long r;
        c.d2h(&r, best.get(), 8);
        if (r != ~0ULL) return r;
        if (base > ((uint64_t)1 << 44)) throw ZkError(ZKGPU_ERR_PROOF, "proof of work search exhausted");
    }
}

// ---- gather ------------------------------------------------------------------------------------------------------
__global__ void gather_kernel(const uint64_t* __restrict__ src, const uint64_t* __restrict__ offs, size_t cnt, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) out[i] = src[offs[i]];
}
__global__ void gather_addr_kernel(const uint64_t* const* __restrict__ addrs, size_t cnt, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) out[i] = *addrs[i];
}
// out_host[i] = *addrs[i] for device addresses: every leaf and every Merkle sibling the query rounds of one table open, from all its
// oracles, in ONE launch and one round trip
void gather_addrs(Ctx& c, const std::vector<const uint64_t*>& addrs, uint64_t* out_host) {
    if (addrs.empty()) return;
    DevBuf o(&c, addrs.size() * 8), d(&c, addrs.size() * 8);
    c.h2d(o.get(), addrs.data(), addrs.size() * 8);
    gather_addr_kernel<<<(unsigned)((addrs.size() + 255) / 256), 256, 0, c.stream>>>(reinterpret_cast<const uint64_t* const*>(o.get()), addrs.size(), d.get());
    c.count_launch();
    c.check_launch("gather_addr_kernel");
    c.d2h(out_host, d.get(), addrs.size() * 8);
}
void gather_words(Ctx& c, const uint64_t* src, const std::vector<uint64_t>& offsets, uint64_t* out_host) {
    if (offsets.empty()) return;
    DevBuf o(&c, offsets.size() * 8), d(&c, offsets.size() * 8);
    c.h2d(o.get(), offsets.data(), offsets.size() * 8);
    gather_kernel<<<(unsigned)((offsets.size() + 255) / 256), 256, 0, c.stream>>>(src, o.get(), offsets.size(), d.get());
    c.count_launch();
    c.check_launch("gather_kernel");
    c.d2h(out_host, d.get(), offsets.size() * 8);
}

}  // namespace zk
