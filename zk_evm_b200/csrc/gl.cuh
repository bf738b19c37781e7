// Goldilocks field p = 2^64 - 2^32 + 1 and its quadratic extension F[X]/(X^2 - 7), host+device.
//
// Replaces plonky2_field 1.0.0 `GoldilocksField` / `QuadraticExtension` as used by the reference's proving path
// (/root/reference/evm_arithmetization/src/prover.rs:8-15; reduction spec book/src/framework/field.md:3-20).
// Every value stored in memory is canonical (< p); the 128->64 reduction is the 96-bit shift-add form
// n0 + (2^32-1) n1 - n2 with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_D inline
#endif

namespace zk {

static constexpr uint64_t GL_P = 0xFFFFFFFF00000001ULL;
static constexpr uint64_t GL_EPS = 0xFFFFFFFFULL;
static constexpr uint64_t GL_GENERATOR = 14293326489335486720ULL;   // multiplicative generator == FRI coset shift
static constexpr uint64_t GL_P2_GENERATOR = 7277203076849721926ULL; // generator of the 2^32 subgroup

ZK_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

#if defined(__CUDACC__)
// ---- device forms: carry-flag arithmetic instead of compare/select chains ------------------------------------------------
// Measured on sm_100a (tools/pipebench.cu): IADD3/LOP3/LEA issue at 1 warp-instr/clk/SMSP, IMAD at 1/2, IMAD.WIDE and
// IMAD.HI at 1/4 and not overlapped with the ALU pipe — so a product is four IMAD.WIDE and everything else is IADD3 carry chains.
__device__ __forceinline__ void gl_unpack(uint64_t x, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ uint64_t gl_pack(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }

// (a - b) mod p for any a and b <= p; canonical when a < p
__device__ __forceinline__ uint64_t gld_sub(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1, m;
    gl_unpack(a, a0, a1); gl_unpack(b, b0, b1);
    asm("sub.cc.u32 %0, %0, %3;\n\t"
        "subc.cc.u32 %1, %1, %4;\n\t"
        "subc.u32 %2, 0, 0;\n\t" : "+r"(a0), "+r"(a1), "=r"(m) : "r"(b0), "r"(b1));     // m = 0xFFFFFFFF on borrow
    asm("sub.cc.u32 %0, %0, %2;\n\t"
        "subc.u32 %1, %1, 0;\n\t" : "+r"(a0), "+r"(a1) : "r"(m));                        // + p == - EPS (mod 2^64)
    return gl_pack(a0, a1);
}
__device__ __forceinline__ uint64_t gld_neg(uint64_t a) { return gld_sub(0, a); }
// a + b = a - (p - b); p - b is computed without reduction (b <= p), p - 0 = p is a valid second operand of gl_sub
__device__ __forceinline__ uint64_t gld_add(uint64_t a, uint64_t b) { return gld_sub(a, GL_P - b); }
__device__ __forceinline__ uint64_t gld_dbl(uint64_t a) { return gld_add(a, a); }

// w0 + 2^32 w1 + 2^64 w2 + 2^96 w3 -> canonical
__device__ __forceinline__ uint64_t gld_reduce_words(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint32_t t0, t1, m, u0, u1;
    asm("sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, 0;\n\t"
        "subc.u32 %2, 0, 0;\n\t" : "=r"(t0), "=r"(t1), "=r"(m) : "r"(w0), "r"(w1), "r"(w3));   // (w1:w0) - w3, 2^96 == -1
    asm("sub.cc.u32 %0, %0, %2;\n\t"
        "subc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(m));
    asm("sub.cc.u32 %0, 0, %2;\n\t"
        "subc.u32 %1, %2, 0;\n\t" : "=r"(u0), "=r"(u1) : "r"(w2));                               // w2 * (2^32 - 1), 2^64 == 2^32 - 1
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, 0, 0;\n\t" : "+r"(t0), "+r"(t1), "=r"(m) : "r"(u0), "r"(u1));
    m = 0u - m;
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, 0;\n\t" : "+r"(t0), "+r"(t1) : "r"(m));
    uint64_t r = gl_pack(t0, t1);
    return r >= GL_P ? r - GL_P : r;
}
__device__ __forceinline__ uint64_t gld_reduce128(uint64_t lo, uint64_t hi) {
    uint32_t w0, w1, w2, w3;
    gl_unpack(lo, w0, w1); gl_unpack(hi, w2, w3);
    return gld_reduce_words(w0, w1, w2, w3);
}
__device__ __forceinline__ uint64_t gld_mul(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1;
    gl_unpack(a, a0, a1); gl_unpack(b, b0, b1);
    uint64_t p00, mid, mid2, hi;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p00) : "r"(a0), "r"(b0));
    uint32_t w0, c0; gl_unpack(p00, w0, c0);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(mid) : "r"(a0), "r"(b1), "l"((uint64_t)c0));
    uint32_t m0, m1; gl_unpack(mid, m0, m1);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(mid2) : "r"(a1), "r"(b0), "l"((uint64_t)m0));
    uint32_t w1, m2; gl_unpack(mid2, w1, m2);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(hi) : "r"(a1), "r"(b1), "l"((uint64_t)m1 + (uint64_t)m2));
    uint32_t w2, w3; gl_unpack(hi, w2, w3);
    return gld_reduce_words(w0, w1, w2, w3);
}
#endif  // __CUDACC__

ZK_HD uint64_t gl_add(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return gld_add(a, b);
#endif
    uint64_t s = a + b;
    // a, b < p so the true sum is < 2p: one conditional subtraction (mask form: on pseudo-random values the comparison is an
    // unpredictable branch, and this is the host transcript's inner loop)
    return s - ((uint64_t)(-(int64_t)((s < a) | (s >= GL_P))) & GL_P);
}
ZK_HD uint64_t gl_sub(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return gld_sub(a, b);
#endif
    uint64_t d = a - b;
    return d + ((uint64_t)(-(int64_t)(a < b)) & GL_P);
}
ZK_HD uint64_t gl_neg(uint64_t a) {
#if defined(__CUDA_ARCH__)
    return gld_neg(a);
#endif
    return a ? GL_P - a : 0;
}
ZK_HD uint64_t gl_dbl(uint64_t a) { return gl_add(a, a); }

// reduce lo + 2^64 * hi to a canonical element
ZK_HD uint64_t gl_reduce128(uint64_t lo, uint64_t hi) {
#if defined(__CUDA_ARCH__)
    return gld_reduce128(lo, hi);
#endif
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    t0 -= (uint64_t)(-(int64_t)(lo < hi_hi)) & GL_EPS;      // borrowed 2^64 == EPS (mod p)
    uint64_t t1 = (hi_lo << 32) - hi_lo;                    // hi_lo * EPS < 2^64 - 2^33 + 1
    uint64_t t2 = t0 + t1;
    t2 += (uint64_t)(-(int64_t)(t2 < t1)) & GL_EPS;         // carried 2^64 == EPS; cannot carry again
    return t2 - ((uint64_t)(-(int64_t)(t2 >= GL_P)) & GL_P);
}
ZK_HD uint64_t gl_mul(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return gld_mul(a, b);
#endif
    return gl_reduce128(a * b, mulhi64(a, b));
}
ZK_HD uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }
// reduce a 96-bit value lo + 2^64 * hi32 (hi32 < 2^32)
ZK_HD uint64_t gl_reduce96(uint64_t lo, uint32_t hi32) {
    uint64_t t1 = ((uint64_t)hi32 << 32) - hi32;
    uint64_t t2 = lo + t1;
    t2 += (uint64_t)(-(int64_t)(t2 < t1)) & GL_EPS;
    return t2 - ((uint64_t)(-(int64_t)(t2 >= GL_P)) & GL_P);
}
ZK_HD uint64_t gl_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_sqr(b); e >>= 1; }
    return r;
}
// Fermat inverse x^(p-2); p-2 = 0xFFFFFFFEFFFFFFFF.  inv(0) = 0.
ZK_HD uint64_t gl_inv(uint64_t x) {
    // addition chain: x^(2^32-1) then compose
    uint64_t x2 = gl_mul(gl_sqr(x), x);                    // 2^2-1
    uint64_t x4 = x2; for (int i = 0; i < 2; i++) x4 = gl_sqr(x4); x4 = gl_mul(x4, x2);     // 2^4-1
    uint64_t x8 = x4; for (int i = 0; i < 4; i++) x8 = gl_sqr(x8); x8 = gl_mul(x8, x4);     // 2^8-1
    uint64_t x16 = x8; for (int i = 0; i < 8; i++) x16 = gl_sqr(x16); x16 = gl_mul(x16, x8); // 2^16-1
    uint64_t x31 = x16; for (int i = 0; i < 8; i++) x31 = gl_sqr(x31); x31 = gl_mul(x31, x8);  // 2^24-1
    for (int i = 0; i < 4; i++) x31 = gl_sqr(x31); x31 = gl_mul(x31, x4);                      // 2^28-1
    for (int i = 0; i < 2; i++) x31 = gl_sqr(x31); x31 = gl_mul(x31, x2);                      // 2^30-1
    x31 = gl_mul(gl_sqr(x31), x);                                                              // 2^31-1
    // p-2 = (2^31-1)*2^33 + (2^32-1) : bits 63..33 ones, bit 32 zero, bits 31..0 ones
    uint64_t r = x31;
    for (int i = 0; i < 33; i++) r = gl_sqr(r);
    uint64_t x32 = gl_mul(gl_sqr(x31), x);                                                     // 2^32-1
    return gl_mul(r, x32);
}
ZK_HD uint64_t gl_root_of_unity(unsigned log_n) {
    uint64_t r = GL_P2_GENERATOR;
    for (unsigned i = log_n; i < 32; i++) r = gl_sqr(r);
    return r;
}
ZK_HD uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

// ------------------------------------------------------------------------------------------------
// base-field wrapper with operators, used as the packed type P of the constraint templates
// ------------------------------------------------------------------------------------------------
struct Fp {
    uint64_t v;
    ZK_HD Fp() : v(0) {}
    ZK_HD explicit Fp(uint64_t x) : v(x) {}              // x must be canonical
    static ZK_HD Fp from_u64(uint64_t x) { return Fp(gl_canon(x)); }
    static ZK_HD Fp zero() { return Fp(0); }
    static ZK_HD Fp one() { return Fp(1); }
};
ZK_HD Fp operator+(Fp a, Fp b) { return Fp(gl_add(a.v, b.v)); }
ZK_HD Fp operator-(Fp a, Fp b) { return Fp(gl_sub(a.v, b.v)); }
ZK_HD Fp operator*(Fp a, Fp b) { return Fp(gl_mul(a.v, b.v)); }
ZK_HD Fp operator-(Fp a) { return Fp(gl_neg(a.v)); }
ZK_HD Fp& operator+=(Fp& a, Fp b) { a = a + b; return a; }
ZK_HD Fp& operator-=(Fp& a, Fp b) { a = a - b; return a; }
ZK_HD Fp& operator*=(Fp& a, Fp b) { a = a * b; return a; }
ZK_HD bool operator==(Fp a, Fp b) { return a.v == b.v; }
ZK_HD Fp scalar_mul(Fp a, uint64_t s) { return Fp(gl_mul(a.v, s)); }

// ------------------------------------------------------------------------------------------------
// quadratic extension, W = 7
// ------------------------------------------------------------------------------------------------
struct Fp2 {
    uint64_t a, b;
    ZK_HD Fp2() : a(0), b(0) {}
    ZK_HD Fp2(uint64_t a_, uint64_t b_) : a(a_), b(b_) {}
    ZK_HD explicit Fp2(uint64_t x) : a(x), b(0) {}
    static ZK_HD Fp2 from_u64(uint64_t x) { return Fp2(gl_canon(x), 0); }
    static ZK_HD Fp2 zero() { return Fp2(0, 0); }
    static ZK_HD Fp2 one() { return Fp2(1, 0); }
};
ZK_HD Fp2 operator+(Fp2 x, Fp2 y) { return Fp2(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
ZK_HD Fp2 operator-(Fp2 x, Fp2 y) { return Fp2(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
ZK_HD Fp2 operator-(Fp2 x) { return Fp2(gl_neg(x.a), gl_neg(x.b)); }
ZK_HD uint64_t gl_mul7(uint64_t x) {   // 7x = 8x - x
    uint64_t x2 = gl_dbl(x), x4 = gl_dbl(x2), x8 = gl_dbl(x4);
    return gl_sub(x8, x);
}
ZK_HD Fp2 operator*(Fp2 x, Fp2 y) {
    uint64_t aa = gl_mul(x.a, y.a), bb = gl_mul(x.b, y.b);
    // Karatsuba: (a+b)(c+d) - ac - bd
    uint64_t cross = gl_sub(gl_sub(gl_mul(gl_add(x.a, x.b), gl_add(y.a, y.b)), aa), bb);
    return Fp2(gl_add(aa, gl_mul7(bb)), cross);
}
ZK_HD Fp2& operator+=(Fp2& a, Fp2 b) { a = a + b; return a; }
ZK_HD Fp2& operator-=(Fp2& a, Fp2 b) { a = a - b; return a; }
ZK_HD Fp2& operator*=(Fp2& a, Fp2 b) { a = a * b; return a; }
ZK_HD bool operator==(Fp2 x, Fp2 y) { return x.a == y.a && x.b == y.b; }
ZK_HD Fp2 scalar_mul(Fp2 x, uint64_t s) { return Fp2(gl_mul(x.a, s), gl_mul(x.b, s)); }
ZK_HD Fp2 fp2_inv(Fp2 x) {
    uint64_t d = gl_inv(gl_sub(gl_sqr(x.a), gl_mul7(gl_sqr(x.b))));
    return Fp2(gl_mul(x.a, d), gl_mul(gl_neg(x.b), d));
}
ZK_HD Fp2 fp2_pow(Fp2 b, uint64_t e) {
    Fp2 r(1, 0);
    while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
    return r;
}

ZK_HD uint32_t bitrev32(uint32_t x, unsigned bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
#endif
}

}  // namespace zk
