// ORACLE (test infrastructure, NOT product code) — PolynomialBatch restatement.
// Follows plonky2 1.0.0 fri/oracle.rs `PolynomialBatch::{from_values, from_coeffs, get_lde_values}` as called from
// /root/reference/evm_arithmetization/src/prover.rs:100-107 and verifier.rs:69-76.
// Parity unpinned (no golden vectors in the reference); validated by the restated verifier.
#pragma once
#include <algorithm>
#include <omp.h>
#include "oracle_core.h"

namespace orc {

struct PolyBatch {
    size_t ncols = 0, n = 0;
    unsigned log_n = 0, rate_bits = 0;
    std::vector<uint64_t> coeffs;   // column-major: coeffs[c*n + j]
    MerkleTree tree;                // leaves[j] = LDE row bitrev(j), row-major

    size_t lde_size() const { return n << rate_bits; }
    const uint64_t* col(size_t c) const { return &coeffs[c * n]; }
    // get_lde_values(i, step): row of the LDE at natural index i*step
    const uint64_t* lde_row(size_t i, size_t step = 1) const { return tree.leaf(bitrev(i * step, log_n + rate_bits)); }

    // from_coeffs: lde_values[c][i] = poly_c(g * w_{k+r}^i); leaves = transpose, bit-reversed rows; MerkleTree::new
    void from_coeffs(std::vector<uint64_t>&& cf, size_t ncols_, size_t n_, unsigned rate_bits_, unsigned cap_height) {
        coeffs = std::move(cf); ncols = ncols_; n = n_; rate_bits = rate_bits_;
        log_n = 0; while (((size_t)1 << log_n) < n) log_n++;
        if (((size_t)1 << log_n) != n) throw std::runtime_error("n must be a power of two");
        size_t N = lde_size(); unsigned log_N = log_n + rate_bits;
        StageClock* sc_l = new StageClock(" lde + transpose");
        LeafVec rows(N * ncols);
        // the coset transform of a column is left in bit-reversed order (fft_dif_bitrev_out): row j of the leaves IS position j of it
        std::vector<uint64_t> shift_pow(N);
        { uint64_t w = 1; for (size_t i = 0; i < n; i++) { shift_pow[i] = w; w = gl_mul(w, GL_GENERATOR); } }
        auto lde_column = [&](size_t c, uint64_t* t, bool par) {
            const uint64_t* src = &coeffs[c * n];
            for (size_t i = 0; i < n; i++) t[i] = gl_mul(src[i], shift_pow[i]);
            memset(t + n, 0, (N - n) * 8);
            fft_dif_bitrev_out(t, log_N, par);
        };
        const size_t CB = 8, nblocks = (ncols + CB - 1) / CB;
        if (nblocks >= (size_t)omp_get_max_threads()) {
            // wide batches: columns in blocks of 8 per thread; the rows are then written one 64-byte line at a time
            #pragma omp parallel
            {
                std::vector<uint64_t> tmp(CB * N);
                #pragma omp for schedule(dynamic, 1)
                for (size_t blk = 0; blk < nblocks; blk++) {
                    const size_t c0 = blk * CB, nb = std::min(CB, ncols - c0);
                    for (size_t b = 0; b < nb; b++) lde_column(c0 + b, tmp.data() + b * N, false);
                    for (size_t j = 0; j < N; j++) {
                        uint64_t* dst = &rows[j * ncols + c0];
                        for (size_t b = 0; b < nb; b++) dst[b] = tmp[b * N + j];
                    }
                }
            }
        } else {
            // tall, narrow batches (Memory, the quotient and auxiliary polynomials): every transform uses all threads, the transposition
            // is split over the rows
            std::vector<uint64_t> tmp(ncols * N);
            for (size_t c = 0; c < ncols; c++) lde_column(c, tmp.data() + c * N, true);
            #pragma omp parallel for schedule(static)
            for (size_t j = 0; j < N; j++) {
                uint64_t* dst = &rows[j * ncols];
                for (size_t c = 0; c < ncols; c++) dst[c] = tmp[c * N + j];
            }
        }
        delete sc_l;
        ORC_STAGE(" merkle tree");
        tree.build(std::move(rows), N, ncols, cap_height);
    }
    // from_values: per-column ifft first
    void from_values(const uint64_t* const* cols, size_t ncols_, size_t n_, unsigned rate_bits_, unsigned cap_height) {
        StageClock* sc_i = new StageClock(" ifft");
        std::vector<uint64_t> cf(ncols_ * n_);
        unsigned lg = 0; while (((size_t)1 << lg) < n_) lg++;
        #pragma omp parallel for schedule(dynamic, 1)
        for (size_t c = 0; c < ncols_; c++) {
            memcpy(&cf[c * n_], cols[c], n_ * 8);
            ifft_inplace(&cf[c * n_], lg);
        }
        delete sc_i;
        from_coeffs(std::move(cf), ncols_, n_, rate_bits_, cap_height);
    }
    // evaluate column c at an extension point (PolynomialCoeffs::to_extension().eval)
    Ext eval_ext(size_t c, Ext z) const {
        Ext acc;
        const uint64_t* p = col(c);
        for (size_t j = n; j-- > 0;) acc = acc * z + Ext(p[j], 0);
        return acc;
    }
    uint64_t eval_base(size_t c, uint64_t z) const {
        uint64_t acc = 0;
        const uint64_t* p = col(c);
        for (size_t j = n; j-- > 0;) acc = gl_add(gl_mul(acc, z), p[j]);
        return acc;
    }
};

}  // namespace orc
