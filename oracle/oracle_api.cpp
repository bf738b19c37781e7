// ORACLE (test infrastructure, NOT product code) — C entry points for ctypes (tests/, smoke(), bench cpu_baseline).
#include "oracle_core.h"
#include "oracle_poly.h"
#include <omp.h>

using namespace orc;

extern "C" {

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

uint64_t orc_gl_add(uint64_t a, uint64_t b) { return gl_add(a, b); }
uint64_t orc_gl_sub(uint64_t a, uint64_t b) { return gl_sub(a, b); }
uint64_t orc_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t orc_gl_inv(uint64_t a) { return gl_inv(a); }
uint64_t orc_gl_pow(uint64_t a, uint64_t e) { return gl_pow(a, e); }
uint64_t orc_root_of_unity(unsigned log_n) { return gl_root_of_unity(log_n); }

void orc_poseidon(uint64_t* states, size_t count) {
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < count; i++) poseidon(states + 12 * i);
}
void orc_hash_no_pad(const uint64_t* in, size_t n, uint64_t out[4]) { Hash h = hash_no_pad(in, n); memcpy(out, h.e, 32); }
void orc_hash_or_noop(const uint64_t* in, size_t n, uint64_t out[4]) { Hash h = hash_or_noop(in, n); memcpy(out, h.e, 32); }
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    Hash a, b; memcpy(a.e, l, 32); memcpy(b.e, r, 32);
    Hash h = two_to_one(a, b); memcpy(out, h.e, 32);
}
// rows given column-major (col c at data + c*nrows), like the device entry point
void orc_hash_rows_colmajor(const uint64_t* data, size_t nrows, size_t width, uint64_t* out) {
    #pragma omp parallel for schedule(static)
    for (size_t j = 0; j < nrows; j++) {
        std::vector<uint64_t> row(width);
        for (size_t c = 0; c < width; c++) row[c] = data[c * nrows + j];
        Hash h = hash_or_noop(row.data(), width);
        memcpy(out + 4 * j, h.e, 32);
    }
}

// batched transforms over ncols contiguous columns of length n; kind: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft
void orc_ntt(uint64_t* data, size_t ncols, size_t n, int kind, uint64_t shift) {
    unsigned lg = 0; while (((size_t)1 << lg) < n) lg++;
    #pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < ncols; c++) {
        uint64_t* a = data + c * n;
        switch (kind) {
            case 0: fft_inplace(a, lg); break;
            case 1: ifft_inplace(a, lg); break;
            case 2: coset_fft_inplace(a, lg, shift); break;
            default: coset_ifft_inplace(a, lg, shift); break;
        }
    }
}
// O(n^2) DFT for pinning the FFT itself on small sizes: out[i] = sum_j in[j] w^(ij)
void orc_naive_dft(const uint64_t* in, uint64_t* out, size_t n) {
    unsigned lg = 0; while (((size_t)1 << lg) < n) lg++;
    uint64_t w = gl_root_of_unity(lg);
    for (size_t i = 0; i < n; i++) {
        uint64_t wi = gl_pow(w, i), x = 1, acc = 0;
        for (size_t j = 0; j < n; j++) { acc = gl_add(acc, gl_mul(in[j], x)); x = gl_mul(x, wi); }
        out[i] = acc;
    }
}

// PolynomialBatch::from_values / from_coeffs.  Outputs (any may be null):
//   coeffs ncols*n col-major; leaves N*ncols row-major (bit-reversed rows); digests plonky2 layout 2*(N-2^cap)*4; cap 2^cap*4
int orc_commit(const uint64_t* cols_contig, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int from_coeffs,
               uint64_t* coeffs, uint64_t* leaves, uint64_t* digests, uint64_t* cap) {
    try {
        PolyBatch b;
        if (from_coeffs) {
            std::vector<uint64_t> cf(cols_contig, cols_contig + ncols * n);
            b.from_coeffs(std::move(cf), ncols, n, rate_bits, cap_height);
        } else {
            std::vector<const uint64_t*> ptr(ncols);
            for (size_t c = 0; c < ncols; c++) ptr[c] = cols_contig + c * n;
            b.from_values(ptr.data(), ncols, n, rate_bits, cap_height);
        }
        if (coeffs) memcpy(coeffs, b.coeffs.data(), ncols * n * 8);
        if (leaves) memcpy(leaves, b.tree.leaves.data(), b.tree.leaves.size() * 8);
        if (digests) { std::vector<Hash> d; b.tree.digests_plonky2_layout(d); if (!d.empty()) memcpy(digests, d.data(), d.size() * 32); }
        if (cap) memcpy(cap, b.tree.cap(), b.tree.cap_len() * 32);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// Challenger replay for tests: ops encoded as (op, value): 0 observe value, 1 get_challenge -> appended to out, 2 compact
size_t orc_challenger_run(const uint64_t* ops, size_t nops, uint64_t* out, uint64_t state_out[12]) {
    Challenger ch; size_t k = 0;
    for (size_t i = 0; i < nops; i++) {
        uint64_t op = ops[2 * i], v = ops[2 * i + 1];
        if (op == 0) ch.observe(v); else if (op == 1) out[k++] = ch.challenge(); else ch.compact();
    }
    memcpy(state_out, ch.state, 96);
    return k;
}

}  // extern "C"
