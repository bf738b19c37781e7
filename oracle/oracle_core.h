// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the arithmetic that the reference's STARK proving path runs through its
// un-vendored dependencies plonky2 1.0.0 / plonky2_field 1.0.0 / starky 1.0.0 (pinned in
// /root/reference/Cargo.lock:3702-3755,4740-4752; call sites evm_arithmetization/src/prover.rs:100-107,
// 118-144, 322-334).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or call this.  The product library (zk_evm_b200/csrc) never includes it.
//
// Parity status: Poseidon / sponge are PINNED by the reference's in-tree known answers
// (smt_trie/src/keys.rs:10-15, evm_arithmetization/src/proof.rs:505-510, smt_trie/src/code.rs:57-84) and the
// field by arithmetic/addcy.rs:67; FFT / Merkle layout / transcript conventions are restated from the
// published upstream crates ("parity unpinned" — no golden vectors exist in the reference for them; they
// are cross-checked by the restated verifier in stark_oracle.cpp).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <vector>
#include <array>
#include <stdexcept>
#include <stdlib.h>
#include "poseidon_constants.h"

#include <map>
#include <string>
#include <chrono>
#include <atomic>

namespace orc {

// wall-clock stage accounting of the prover (main thread only; read through orc_stage_report): where the CPU baseline spends its time
struct StageClock {
    static std::map<std::string, double>& acc() { static std::map<std::string, double> m; return m; }
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    const char* name; double t0;
    explicit StageClock(const char* n) : name(n), t0(now()) {}
    ~StageClock() { acc()[name] += now() - t0; }
};
#define ORC_STAGE(name) orc::StageClock _stage_clock_##__LINE__(name)

typedef unsigned __int128 u128;
static const uint64_t P = 0xFFFFFFFF00000001ULL;
static const uint64_t EPS = 0xFFFFFFFFULL;  // 2^32 - 1 = 2^64 mod p

// ---------------------------------------------------------------------------------------------
// Goldilocks base field, canonical representatives only (book/src/framework/field.md:3-20).
// ---------------------------------------------------------------------------------------------
// (branch-free forms: the comparisons of modular arithmetic on pseudo-random values are unpredictable branches)
static inline uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    return s - ((uint64_t)(-(int64_t)((s < a) | (s >= P))) & P);
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return a - b + ((uint64_t)(-(int64_t)(a < b)) & P); }
static inline uint64_t gl_neg(uint64_t a) { return a ? P - a : 0; }
// n = n0 + 2^64 n1 + 2^96 n2  ->  n0 + (2^32-1) n1 - n2   (field.md:9-20)
static inline uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t n1 = hi & EPS, n2 = hi >> 32;
    uint64_t t = lo - n2;                                   // a borrow means + 2^64 = + EPS (mod p) too much
    t -= (uint64_t)(-(int64_t)(lo < n2)) & EPS;
    uint64_t r = t + n1 * EPS;                              // n1*EPS < 2^64 - 2^33 + 1; a carry means 2^64 = EPS (mod p) is missing
    r += (uint64_t)(-(int64_t)(r < t)) & EPS;
    return r - ((uint64_t)(-(int64_t)(r >= P)) & P);
}
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
static inline uint64_t gl_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_mul(b, b); e >>= 1; }
    return r;
}
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, P - 2); }  // inv(0) = 0
static inline uint64_t gl_from_u64(uint64_t x) { return x >= P ? x - P : x; }

static const uint64_t GL_GENERATOR = 14293326489335486720ULL;      // MULTIPLICATIVE_GROUP_GENERATOR = coset shift
static const uint64_t GL_P2_GENERATOR = 7277203076849721926ULL;    // POWER_OF_TWO_GENERATOR, order 2^32
static inline uint64_t gl_root_of_unity(unsigned log_n) {           // Field::primitive_root_of_unity
    uint64_t r = GL_P2_GENERATOR;
    for (unsigned i = log_n; i < 32; i++) r = gl_mul(r, r);
    return r;
}

// ---------------------------------------------------------------------------------------------
// Quadratic extension F[X]/(X^2 - 7)
// ---------------------------------------------------------------------------------------------
struct Ext {
    uint64_t a, b;
    Ext() : a(0), b(0) {}
    Ext(uint64_t a_, uint64_t b_ = 0) : a(a_), b(b_) {}
    bool operator==(const Ext& o) const { return a == o.a && b == o.b; }
    bool operator!=(const Ext& o) const { return !(*this == o); }
};
static inline Ext operator+(Ext x, Ext y) { return Ext(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
static inline Ext operator-(Ext x, Ext y) { return Ext(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
static inline Ext operator*(Ext x, Ext y) {
    return Ext(gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b))),
               gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
static inline Ext ext_scalar(Ext x, uint64_t s) { return Ext(gl_mul(x.a, s), gl_mul(x.b, s)); }
static inline Ext ext_inv(Ext x) {  // 1/(a + bX) = (a - bX)/(a^2 - 7 b^2)
    uint64_t d = gl_inv(gl_sub(gl_mul(x.a, x.a), gl_mul(7, gl_mul(x.b, x.b))));
    return Ext(gl_mul(x.a, d), gl_mul(gl_neg(x.b), d));
}
static inline Ext ext_pow(Ext b, uint64_t e) {
    Ext r(1, 0);
    while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
    return r;
}

// ---------------------------------------------------------------------------------------------
// Poseidon permutation (width 12) in its naive 30-round form; equivalent to plonky2's optimised
// partial rounds (in-tree restatement of the round structure: evm_arithmetization/src/poseidon/
// poseidon_stark.rs:329-405).  Known answers: see tests/test_oracle_kats.py.
// ---------------------------------------------------------------------------------------------
static const uint64_t POSEIDON_RC[360] = ZK_POSEIDON_RC_INIT;
static const uint64_t MDS_CIRC[12] = ZK_POSEIDON_MDS_CIRC_INIT;

static inline uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}
static inline void mds_layer(uint64_t s[12]) {
    uint64_t d[24], out[12];
    memcpy(d, s, 96); memcpy(d + 12, s, 96);                // d[i + r] = s[(i + r) % 12]
    for (int r = 0; r < 12; r++) {
        u128 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u128)d[i + r] * MDS_CIRC[i];
        if (r == 0) acc += (u128)s[0] * ZK_POSEIDON_MDS_DIAG0;
        out[r] = gl_reduce128(acc);
    }
    memcpy(s, out, sizeof(out));
}
static inline void poseidon(uint64_t s[12]) {
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC[12 * r + i]);
        if (r < 4 || r >= 26) { for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]); }
        else s[0] = sbox7(s[0]);
        mds_layer(s);
    }
}

struct Hash { uint64_t e[4]; };
static inline bool operator==(const Hash& x, const Hash& y) { return !memcmp(x.e, y.e, 32); }

// PoseidonHash::hash_no_pad: overwrite-mode sponge, rate 8, output 4 (KAT proof.rs:505-510)
static inline Hash hash_no_pad(const uint64_t* in, size_t n) {
    uint64_t st[12] = {0};
    for (size_t i = 0; i < n; i += 8) {
        size_t l = n - i < 8 ? n - i : 8;
        for (size_t j = 0; j < l; j++) st[j] = in[i + j];
        poseidon(st);
    }
    Hash h; memcpy(h.e, st, 32); return h;
}
// hash_or_noop: inputs of <= 4 elements are copied (zero padded), not hashed
static inline Hash hash_or_noop(const uint64_t* in, size_t n) {
    if (n <= 4) { Hash h = {{0, 0, 0, 0}}; for (size_t i = 0; i < n; i++) h.e[i] = in[i]; return h; }
    return hash_no_pad(in, n);
}
static inline Hash two_to_one(const Hash& l, const Hash& r) {
    uint64_t st[12] = {l.e[0], l.e[1], l.e[2], l.e[3], r.e[0], r.e[1], r.e[2], r.e[3], 0, 0, 0, 0};
    poseidon(st);
    Hash h; memcpy(h.e, st, 32); return h;
}

// ---------------------------------------------------------------------------------------------
// FFT (plonky2_field fft.rs semantics: natural order in, natural order out)
// ---------------------------------------------------------------------------------------------
static inline size_t bitrev(size_t x, unsigned bits) {
    if (!bits) return 0;
    uint64_t v = x;
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
    v = __builtin_bswap64(v);
    return (size_t)(v >> (64 - bits));
}
// out[i] = sum_j in[j] w^(ij), w = primitive_root_of_unity(log_n)   (textbook iterative radix-2 DIT)
// Twiddles: one table per transform size, w^j for j < n/2, built once and shared by every column and thread (the stage of
// span m uses every (n/m)-th entry), as plonky2's fft_root_table does.
struct FftRoots {
    std::vector<uint64_t> w[33];
    std::atomic<bool> ready[33];
    FftRoots() { for (auto& r : ready) r.store(false); }
    const uint64_t* get(unsigned log_n) {
        if (!log_n) return nullptr;
        if (!ready[log_n].load(std::memory_order_acquire)) {
            #pragma omp critical(orc_fft_roots)
            if (!ready[log_n].load(std::memory_order_relaxed)) {
                std::vector<uint64_t> t((size_t)1 << (log_n - 1));
                uint64_t r = gl_root_of_unity(log_n), x = 1;
                for (auto& e : t) { e = x; x = gl_mul(x, r); }
                w[log_n].swap(t);
                ready[log_n].store(true, std::memory_order_release);
            }
        }
        return w[log_n].data();
    }
};
static inline FftRoots& fft_roots() { static FftRoots r; return r; }
// one radix-2 DIT stage (span m = 2^s) on the elements [lo, hi) of a (hi - lo a multiple of m)
static inline void fft_stage(uint64_t* a, size_t lo, size_t hi, unsigned s, size_t n, const uint64_t* tw) {
    const size_t m = (size_t)1 << s, half = m >> 1, step = n >> s;
    if (half == 1) {
        for (size_t k = lo; k < hi; k += 2) { uint64_t u = a[k], t = a[k + 1]; a[k] = gl_add(u, t); a[k + 1] = gl_sub(u, t); }
        return;
    }
    for (size_t k = lo; k < hi; k += m)
        for (size_t j = 0; j < half; j++) {
            uint64_t t = gl_mul(tw[j * step], a[k + j + half]), u = a[k + j];
            a[k + j] = gl_add(u, t);
            a[k + j + half] = gl_sub(u, t);
        }
}
static inline void fft_inplace(uint64_t* a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev(i, log_n); if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; } }
    const uint64_t* tw = fft_roots().get(log_n);
    // the stages whose span fits a 32 KB block are done block by block (all of them while the block is in L1), the wider ones sweep the
    // whole array: the same butterflies in another order
    const unsigned BLK = 12;
    const unsigned inner = log_n < BLK ? log_n : BLK;
    for (size_t b = 0; b < n; b += (size_t)1 << inner)
        for (unsigned s = 1; s <= inner; s++) fft_stage(a, b, b + ((size_t)1 << inner), s, n, tw);
    for (unsigned s = inner + 1; s <= log_n; s++) fft_stage(a, 0, n, s, n, tw);
}
// The same transform with its output LEFT IN BIT-REVERSED ORDER: a[bitrev(k)] = sum_j in[j] w^(jk) (decimation in frequency: the wide
// stages first, then block by block; no permutation pass).  PolynomialBatch wants exactly this order — leaves[j] = lde[bitrev(j)] — so
// the commitment never pays for a bit reversal.  `par`: split every stage over the OpenMP threads (tall, narrow batches).
static inline void dif_stage(uint64_t* a, size_t lo, size_t hi, unsigned s, size_t n, const uint64_t* tw) {
    const size_t m = (size_t)1 << s, half = m >> 1, step = n >> s;
    if (half == 1) {
        for (size_t k = lo; k < hi; k += 2) { uint64_t u = a[k], t = a[k + 1]; a[k] = gl_add(u, t); a[k + 1] = gl_sub(u, t); }
        return;
    }
    for (size_t k = lo; k < hi; k += m)
        for (size_t j = 0; j < half; j++) {
            uint64_t u = a[k + j], v = a[k + j + half];
            a[k + j] = gl_add(u, v);
            a[k + j + half] = gl_mul(gl_sub(u, v), tw[j * step]);
        }
}
static inline void fft_dif_bitrev_out(uint64_t* a, unsigned log_n, bool par = false) {
    const size_t n = (size_t)1 << log_n;
    if (!log_n) return;
    const uint64_t* tw = fft_roots().get(log_n);
    const unsigned BLK = 12;
    const unsigned inner = log_n < BLK ? log_n : BLK;
    const size_t blk = (size_t)1 << inner;
    for (unsigned s = log_n; s > inner; s--) {
        // a wide stage: 2^(log_n - s) independent groups of span 2^s; cut into pieces of 2^inner butterflies each
        const size_t m = (size_t)1 << s, half = m >> 1, step = n >> s, pieces = n / 2 / blk * 2;
        #pragma omp parallel for schedule(static) if (par)
        for (size_t pc = 0; pc < pieces; pc++) {
            const size_t per = half / (blk / 2) ? half / (blk / 2) : 1;          // pieces per group
            const size_t g = pc / per, j0 = (pc % per) * (blk / 2), k = g * m;
            for (size_t j = j0; j < j0 + blk / 2 && j < half; j++) {
                uint64_t u = a[k + j], v = a[k + j + half];
                a[k + j] = gl_add(u, v);
                a[k + j + half] = gl_mul(gl_sub(u, v), tw[j * step]);
            }
        }
    }
    #pragma omp parallel for schedule(static) if (par)
    for (size_t b = 0; b < n; b += blk)
        for (unsigned s = inner; s >= 1; s--) dif_stage(a, b, b + blk, s, n, tw);
}
// ifft(v)[j] = n^-1 sum_i v_i w^(-ij)
static inline void ifft_inplace(uint64_t* a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    fft_inplace(a, log_n);
    uint64_t ninv = gl_inv(gl_from_u64(n));
    for (size_t i = 1; i < n / 2; i++) { uint64_t t = a[i]; a[i] = a[n - i]; a[n - i] = t; }
    for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv);
}
// coset_fft(c, s) = fft(c_j s^j)
static inline void coset_fft_inplace(uint64_t* a, unsigned log_n, uint64_t shift) {
    size_t n = (size_t)1 << log_n;
    uint64_t w = 1;
    for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], w); w = gl_mul(w, shift); }
    fft_inplace(a, log_n);
}
// coset_ifft(v, s)_j = ifft(v)_j s^-j
static inline void coset_ifft_inplace(uint64_t* a, unsigned log_n, uint64_t shift) {
    size_t n = (size_t)1 << log_n;
    ifft_inplace(a, log_n);
    uint64_t si = gl_inv(shift), w = 1;
    for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], w); w = gl_mul(w, si); }
}

// ---------------------------------------------------------------------------------------------
// Merkle tree over rows ("leaves"), cap of 2^cap_height subtree roots (plonky2 hash/merkle_tree.rs).
// We keep digests level by level (level 0 = leaf digests); `digests_plonky2_layout` re-creates the
// host layout of MerkleTree::digests for the PolynomialBatch export.
// ---------------------------------------------------------------------------------------------
// eight-at-a-time forms (poseidon_avx512.h), used when the CPU has AVX-512
static inline bool have_avx512();
static inline void hash_no_pad_x8(const uint64_t* const rows[8], size_t len, Hash out[8]);
static inline void two_to_one_x8(const Hash* children, Hash* out);

// storage whose elements are NOT zeroed on allocation: the leaves of a large commitment are 10 GB that every element of is written once
template <class T> struct DefaultInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
    template <class U, class... A> void construct(U* p, A&&... args) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(args)...);
    }
};
using LeafVec = std::vector<uint64_t, DefaultInitAlloc<uint64_t>>;

struct MerkleTree {
    size_t num_leaves = 0, leaf_len = 0;
    unsigned cap_height = 0;
    LeafVec leaves;                               // row-major num_leaves x leaf_len
    std::vector<std::vector<Hash>> levels;        // levels[0] = leaf digests ... last = cap
    const Hash* cap() const { return levels.back().data(); }
    size_t cap_len() const { return levels.back().size(); }
    const uint64_t* leaf(size_t i) const { return &leaves[i * leaf_len]; }

    void build(LeafVec&& rows, size_t n_leaves, size_t len, unsigned cap_h) {
        leaves = std::move(rows); num_leaves = n_leaves; leaf_len = len; cap_height = cap_h;
        if (((size_t)1 << cap_h) > n_leaves) throw std::runtime_error("cap height too large for tree");
        levels.clear();
        levels.emplace_back(n_leaves);
        const bool x8 = have_avx512();
        if (x8 && len > 4 && n_leaves >= 8) {
            #pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n_leaves; i += 8) {
                const uint64_t* rows8[8];
                for (int k = 0; k < 8; k++) rows8[k] = &leaves[(i + k) * len];
                hash_no_pad_x8(rows8, len, &levels[0][i]);
            }
        } else {
            #pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n_leaves; i++) levels[0][i] = hash_or_noop(&leaves[i * len], len);
        }
        while (levels.back().size() > ((size_t)1 << cap_h)) {
            const std::vector<Hash>& prev = levels.back();
            std::vector<Hash> next(prev.size() / 2);
            if (x8 && next.size() >= 8) {
                #pragma omp parallel for schedule(static)
                for (size_t i = 0; i < next.size(); i += 8) two_to_one_x8(&prev[2 * i], &next[i]);
            } else {
                #pragma omp parallel for schedule(static)
                for (size_t i = 0; i < next.size(); i++) next[i] = two_to_one(prev[2 * i], prev[2 * i + 1]);
            }
            levels.push_back(std::move(next));
        }
    }
    // MerkleTree::prove: sibling digests bottom-up, log2(num_leaves) - cap_height of them
    std::vector<Hash> prove(size_t leaf_index) const {
        std::vector<Hash> sib;
        size_t idx = leaf_index;
        for (size_t l = 0; l + 1 < levels.size(); l++) { sib.push_back(levels[l][idx ^ 1]); idx >>= 1; }
        return sib;
    }
    // plonky2 `digests` vector: per cap subtree, recursively  buf(left) | digest(left child) | digest(right child) | buf(right)
    void digests_plonky2_layout(std::vector<Hash>& out) const {
        out.clear();
        size_t ncap = (size_t)1 << cap_height;
        size_t per = num_leaves / ncap;
        if (per < 2) return;
        out.resize(2 * (num_leaves - ncap));
        size_t sub = 2 * per - 2;
        for (size_t c = 0; c < ncap; c++) fill_rec(&out[c * sub], levels.size() - 1, c);
    }
  private:
    // writes the buffer (2*m - 2 entries) of the subtree rooted at node `idx` of level `lvl`
    void fill_rec(Hash* buf, size_t lvl, size_t idx) const {
        if (lvl == 0) return;
        size_t m = (size_t)1 << lvl;        // leaves under this node
        size_t half_buf = m - 2;            // entries of each child's buffer (2*(m/2) - 2)
        fill_rec(buf, lvl - 1, 2 * idx);
        buf[half_buf] = levels[lvl - 1][2 * idx];
        buf[half_buf + 1] = levels[lvl - 1][2 * idx + 1];
        fill_rec(buf + half_buf + 2, lvl - 1, 2 * idx + 1);
    }
};
static inline Hash merkle_root_from_proof(const uint64_t* leaf, size_t leaf_len, size_t index, const std::vector<Hash>& sib) {
    Hash h = hash_or_noop(leaf, leaf_len);
    for (const Hash& s : sib) { h = (index & 1) ? two_to_one(s, h) : two_to_one(h, s); index >>= 1; }
    return h;
}

// ---------------------------------------------------------------------------------------------
// Challenger (plonky2 iop/challenger.rs): duplex sponge, overwrite mode, outputs popped from the back.
// ---------------------------------------------------------------------------------------------
struct Challenger {
    uint64_t state[12];
    std::vector<uint64_t> in, out;
    Challenger() { memset(state, 0, sizeof(state)); }
    void duplex() {
        for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
        in.clear();
        poseidon(state);
        out.assign(state, state + 8);
    }
    void observe(uint64_t x) { out.clear(); in.push_back(x); if (in.size() == 8) duplex(); }
    void observe_n(const uint64_t* x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
    void observe_hash(const Hash& h) { observe_n(h.e, 4); }
    void observe_cap(const Hash* cap, size_t n) { for (size_t i = 0; i < n; i++) observe_hash(cap[i]); }
    void observe_ext(Ext e) { observe(e.a); observe(e.b); }
    uint64_t challenge() {
        if (!in.empty() || out.empty()) duplex();
        uint64_t r = out.back(); out.pop_back(); return r;
    }
    Ext ext_challenge() { uint64_t a = challenge(); uint64_t b = challenge(); return Ext(a, b); }
    // Challenger::compact (prover.rs:320): absorb pending input, drop buffered outputs
    void compact() { if (!in.empty()) duplex(); out.clear(); }
    void set_state(const uint64_t s[12]) { memcpy(state, s, sizeof(state)); in.clear(); out.clear(); }
};

}  // namespace orc

#include "poseidon_avx512.h"
