// ORACLE (test infrastructure, NOT product code) — C entry points for ctypes (tests/, smoke(), bench cpu_baseline).
#include "oracle_core.h"
#include "oracle_poly.h"
#include <omp.h>
#include <map>

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
// the AVX-512 form on the same interface (count rounded down to a multiple of 8 is done eight at a time, the rest by the scalar
// form); returns 0 when the CPU has no AVX-512 (nothing is computed)
int orc_poseidon_x8(uint64_t* states, size_t count) {
    if (!have_avx512()) return 0;
    const size_t full = count / 8 * 8;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < full; i += 8) {
        alignas(64) uint64_t st[12][8];
        for (int k = 0; k < 8; k++) for (int j = 0; j < 12; j++) st[j][k] = states[12 * (i + k) + j];
        poseidon_x8(st);
        for (int k = 0; k < 8; k++) for (int j = 0; j < 12; j++) states[12 * (i + k) + j] = st[j][k];
    }
    for (size_t i = full; i < count; i++) poseidon(states + 12 * i);
    return 1;
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

// ---- per-table STARK prover / verifier (oracle_stark.h) ---------------------------------------------------------
#include "oracle_stark.h"

static Config cfg_from(const uint32_t c[8]) {
    Config k;
    k.security_bits = c[0]; k.num_challenges = c[1]; k.rate_bits = c[2]; k.cap_height = c[3]; k.pow_bits = c[4];
    k.arity_bits = c[5]; k.final_poly_bits = c[6]; k.num_queries = c[7];
    return k;
}
static TableParams params_from(const uint64_t l[4]) {
    TableParams p; p.halt_final = l[0]; p.init = l[1]; p.syscall_jumptable = l[2]; p.exception_jumptable = l[3];
    return p;
}
static thread_local std::string g_orc_err;

extern "C" {

const char* orc_last_error(void) { return g_orc_err.c_str(); }
uint32_t orc_table_num_columns(uint32_t table) { return zkstark::table_num_columns(table); }
int orc_table_supported(uint32_t table) { return zkstark::table_supported(table) ? 1 : 0; }
size_t orc_table_num_aux(uint32_t table, uint32_t num_challenges) {
    return aux_shape(table, zkstark::all_cross_table_lookups(), num_challenges, zkstark::CONSTRAINT_DEGREE).num_aux();
}

// The device's index-addressed (reordered) evaluators must give what the reference's emission order gives on ANY two rows, not
// only on valid traces: evaluates both forms of a table's constraints on the given rows (base field, two alphas) and returns 1
// when the accumulators agree.  Only Keccak has a reordered form so far.
int orc_eval_forms_agree(uint32_t table, const uint64_t* lv, const uint64_t* nv, const uint64_t alphas[2], const uint64_t sel[3],
                         uint64_t out_seq[2], uint64_t out_blk[2]) {
    ConsumerT<OF> a, b;
    for (ConsumerT<OF>* y : {&a, &b}) {
        for (int j = 0; j < 2; j++) { y->alphas.push_back(OF(alphas[j])); y->acc.push_back(OF(0)); }
        y->z_last = OF(sel[0]); y->lagrange_first = OF(sel[1]); y->lagrange_last = OF(sel[2]);
    }
    RowOF l{lv}, n{nv};
    if (table != zkstark::T_KECCAK) return -1;
    zkstark::keccak::eval<OF>(l, n, a);
    zkstark::keccak::eval_blocked<OF>(l, n, b);
    for (int j = 0; j < 2; j++) { out_seq[j] = a.acc[j].v; out_blk[j] = b.acc[j].v; }
    return (a.acc[0].v == b.acc[0].v && a.acc[1].v == b.acc[1].v) ? 1 : 0;
}

// prove_single_table (prover.rs:301-341) for one table given its trace and the CTL challenges.
// trace: ncols*n column-major.  beta_gamma: [beta_0, gamma_0, beta_1, gamma_1 ...].  chal_state: compacted transcript
// state in, state after this table's proof out.  Returns the number of proof words written (or needed when out is too
// small / null), negative on failure.
long orc_prove_table(uint32_t table, const uint32_t cfgw[8], const uint64_t* trace, size_t ncols, size_t n,
                     const uint64_t* beta_gamma, uint64_t chal_state[12], const uint64_t labels[4], const uint64_t* forced_pow,
                     uint64_t* out, size_t out_cap, uint64_t* aux_out, uint64_t* quot_out, uint64_t* fri_values_out) {
    try {
        Config cfg = cfg_from(cfgw);
        if (ncols != zkstark::table_num_columns(table)) throw std::runtime_error("wrong number of trace columns");
        std::vector<const uint64_t*> cols(ncols);
        for (size_t c = 0; c < ncols; c++) cols[c] = trace + c * n;
        PolyBatch tc;
        tc.from_values(cols.data(), ncols, n, cfg.rate_bits, cfg.cap_height);
        std::vector<uint64_t> betas, gammas;
        for (unsigned i = 0; i < cfg.num_challenges; i++) { betas.push_back(beta_gamma[2 * i]); gammas.push_back(beta_gamma[2 * i + 1]); }
        auto ctls = zkstark::all_cross_table_lookups();
        CtlData ctl = ctl_data_for_table(table, cols.data(), n, ctls, betas, gammas, zkstark::CONSTRAINT_DEGREE);
        Challenger ch; ch.set_state(chal_state);
        ProveDebug dbg;
        StarkProofData p = prove_table(table, cfg, cols.data(), n, tc, ctl, betas, gammas, ch, params_from(labels), forced_pow, ctls, &dbg);
        ch.compact();
        memcpy(chal_state, ch.state, 96);
        if (aux_out) for (size_t c = 0; c < dbg.aux_values.size(); c++) memcpy(aux_out + c * n, dbg.aux_values[c].data(), n * 8);
        if (quot_out) memcpy(quot_out, dbg.quotient_chunk_coeffs.data(), dbg.quotient_chunk_coeffs.size() * 8);
        if (fri_values_out) for (size_t i = 0; i < dbg.fri_final_values.size(); i++) { fri_values_out[2 * i] = dbg.fri_final_values[i].a; fri_values_out[2 * i + 1] = dbg.fri_final_values[i].b; }
        Words w = zkstark::serialize_proof(p);
        if (out && out_cap >= w.size()) memcpy(out, w.data(), w.size() * 8);
        return (long)w.size();
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// verify one table's proof; returns 1 if valid, 0 if rejected (reason in orc_last_error), -1 on malformed input
int orc_verify_table(uint32_t table, const uint32_t cfgw[8], const uint64_t* proof, size_t len, const uint64_t* beta_gamma,
                     uint64_t chal_state[12], const uint64_t labels[4]) {
    try {
        Config cfg = cfg_from(cfgw);
        StarkProofData p = zkstark::deserialize_proof(proof, len);
        std::vector<uint64_t> betas, gammas;
        for (unsigned i = 0; i < cfg.num_challenges; i++) { betas.push_back(beta_gamma[2 * i]); gammas.push_back(beta_gamma[2 * i + 1]); }
        Challenger ch; ch.set_state(chal_state);
        std::string err;
        bool ok = verify_table(table, cfg, p, betas, gammas, ch, params_from(labels), zkstark::all_cross_table_lookups(), err);
        ch.compact();
        memcpy(chal_state, ch.state, 96);
        g_orc_err = err;
        return ok ? 1 : 0;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

}  // extern "C"

// ---- whole-segment prover / verifier: prove_with_traces (prover.rs:72-194) and verify_proof (verifier.rs:172-313) ------------
static void segment_transcript(Challenger& ch, const std::vector<Words>& caps, const uint8_t in_use[9], size_t cap_words,
                               const uint64_t* pv, size_t npv, unsigned nch, std::vector<uint64_t>& betas, std::vector<uint64_t>& gammas) {
    for (uint32_t t = 0; t < zkstark::NUM_TABLES; t++) {
        if (!in_use[t]) {
            if (!zkstark::table_is_optional(t)) throw std::runtime_error("only optional tables may be left out");
            for (size_t i = 0; i < cap_words; i++) ch.observe(0);      // zero cap, prover.rs:120-123
        } else ch.observe_n(caps[t].data(), cap_words);
    }
    ch.observe_n(pv, npv);                                             // observe_public_values, flattened by the caller
    for (unsigned i = 0; i < nch; i++) { betas.push_back(ch.challenge()); gammas.push_back(ch.challenge()); }
}

extern "C" {

// traces[t]: ncols(t)*n[t] column-major, or null (table not in use).  Proof words of table t are written at
// out[offsets[t] .. offsets[t+1]) (empty for unused tables).  Returns total words (needed if > out_cap), negative on failure.
// ---- row-level constraint check -----------------------------------------------------------------------------------------------
// The reference's own generator tests (e.g. logic.rs:426-472, keccak_stark.rs:631-690): evaluate the table's constraints on
// every pair of consecutive TRACE rows and require all of them to vanish (transition constraints not on the last row, first / last
// row constraints only there).  Reports every violated (row, constraint index), the index counting yield_constr calls in the
// evaluator's order — which is what a new trace generator needs in order to be debugged.  Table constraints only (the lookup / CTL
// checks need the auxiliary columns; the restated verifier covers those).
struct RowCheckConsumer {
    size_t row = 0, n = 0;
    uint32_t idx = 0;
    std::vector<uint64_t>* out = nullptr;     // pairs (row, constraint index)
    size_t max_out = 0;
    void hit(OF c) { if (c.v != 0 && out->size() < 2 * max_out) { out->push_back(row); out->push_back(idx); } idx++; }
    void constraint(OF c) { hit(c); }
    void constraint_transition(OF c) { if (row + 1 < n) hit(c); else idx++; }
    void constraint_first_row(OF c) { if (row == 0) hit(c); else idx++; }
    void constraint_last_row(OF c) { if (row + 1 == n) hit(c); else idx++; }
    // index-addressed blocks (the device's reordered evaluators are not used here, but the interface must exist)
    uint32_t blk_base = 0;
    void block_begin(uint32_t M) { blk_base = idx; idx += M; }
    void block_put(uint32_t i, OF c) { if (c.v != 0 && out->size() < 2 * max_out) { out->push_back(row); out->push_back(blk_base + i); } }
};
// -> number of violations written (<= max_pairs) as (row, constraint index) pairs; -1 on error
long orc_check_table_rows(uint32_t table, const uint64_t* trace, size_t ncols, size_t n, const uint64_t labels[4], uint64_t* out_pairs, size_t max_pairs) {
    try {
        if (ncols != zkstark::table_num_columns(table)) throw std::runtime_error("trace width does not match the table");
        std::vector<uint64_t> out;
        std::vector<uint64_t> lrow(ncols), nrow(ncols);
        RowCheckConsumer yc;
        yc.n = n; yc.out = &out; yc.max_out = max_pairs;
        for (size_t r = 0; r < n; r++) {
            for (size_t c = 0; c < ncols; c++) { lrow[c] = trace[c * n + r]; nrow[c] = trace[c * n + (r + 1) % n]; }
            RowOF lv{lrow.data()}, nv{nrow.data()};
            yc.row = r; yc.idx = 0;
            zkstark::eval_table<OF>(table, lv, nv, yc, params_from(labels));
        }
        size_t k = out.size() / 2 < max_pairs ? out.size() / 2 : max_pairs;
        if (out_pairs) memcpy(out_pairs, out.data(), 2 * k * 8);
        return (long)k;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// number of constraints the table's evaluator yields per row pair (its yield_constr calls; lookup / CTL checks not included)
long orc_table_num_constraints(uint32_t table) {
    try {
        const size_t ncols = zkstark::table_num_columns(table);
        std::vector<uint64_t> row(ncols, 0), out;
        RowCheckConsumer yc;
        yc.n = 4; yc.out = &out; yc.max_out = 0;
        RowOF lv{row.data()}, nv{row.data()};
        uint64_t labels[4] = {1, 2, 3, 4};
        zkstark::eval_table<OF>(table, lv, nv, yc, params_from(labels));
        return (long)yc.idx;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// starky 1.0.0 stark_testing.rs `test_stark_low_degree`, which the reference runs for every table (cpu_stark.rs:679-703,
// memory_stark.rs:902-925, logic.rs:400-424, keccak_stark.rs:631-655, keccak_sponge_stark.rs:973-993, byte_packing_stark.rs:453-476,
// memory_continuation_stark.rs:160-180, arithmetic_stark.rs:334+), restated: a random trace of degree < n is extended to the
// subgroup of size 4n (rate_bits = log2_ceil(constraint_degree + 1) = 2), the table's constraints are evaluated at every point
// (next row = 4 points on; x - w_n^-1, L_0, L_{n-1} as the consumer's selectors; Horner in one random alpha) and the values are
// interpolated.  -> the degree of that polynomial (the reference asserts degree <= 3 n - 1), -2 when it is identically zero, -1 on error.
long orc_table_constraint_degree(uint32_t table, unsigned log_n, uint64_t seed, const uint64_t labels[4]) {
    try {
        const unsigned rate_bits = 2, L = log_n + rate_bits;
        const size_t n = (size_t)1 << log_n, N = n << rate_bits, ncols = zkstark::table_num_columns(table);
        if (ncols == 0 || log_n < 1 || L > 20) throw std::runtime_error("bad table or size");
        uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        auto rnd = [&]() { st += 0x9E3779B97F4A7C15ull; uint64_t z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                           return gl_from_u64(z ^ (z >> 31)); };
        auto extend = [&](std::vector<uint64_t>& v) { ifft_inplace(v.data(), log_n); v.resize(N, 0); fft_inplace(v.data(), L); };
        std::vector<std::vector<uint64_t>> lde(ncols);
        for (auto& col : lde) { col.resize(n); for (auto& x : col) x = rnd(); extend(col); }
        std::vector<uint64_t> first(n, 0), last(n, 0);
        first[0] = 1; last[n - 1] = 1;
        extend(first); extend(last);
        const uint64_t alpha = rnd(), wN = gl_root_of_unity(L), last_pt = gl_inv(gl_root_of_unity(log_n));
        std::vector<OF> apow(1025);
        apow[0] = OF(1);
        for (size_t e = 1; e < apow.size(); e++) apow[e] = apow[e - 1] * OF(alpha);
        std::vector<uint64_t> evals(N), lrow(ncols), nrow(ncols);
        uint64_t x = 1;
        for (size_t i = 0; i < N; i++, x = gl_mul(x, wN)) {
            const size_t inext = (i + ((size_t)1 << rate_bits)) % N;
            for (size_t c = 0; c < ncols; c++) { lrow[c] = lde[c][i]; nrow[c] = lde[c][inext]; }
            zkstark::Consumer<OF, 1> yc;
            yc.nc = 1; yc.alpha[0] = OF(alpha); yc.acc[0] = OF(0); yc.apow[0] = apow.data();
            yc.z_last = OF(gl_sub(x, last_pt)); yc.lagrange_first = OF(first[i]); yc.lagrange_last = OF(last[i]);
            RowOF lv{lrow.data()}, nv{nrow.data()};
            zkstark::eval_table<OF>(table, lv, nv, yc, params_from(labels));
            evals[i] = yc.acc[0].v;
        }
        ifft_inplace(evals.data(), L);
        for (size_t d = N; d-- > 0;) if (evals[d] != 0) return (long)d;
        return -2;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// starky 1.0.0 cross_table_lookup.rs debug_utils `check_ctls`, which the reference runs over the raw traces of every segment when
// debug assertions are on (prover.rs:165-184; CI builds with -Cdebug-assertions, ci.yml:93), restated: for every cross-table lookup
// the multiset of rows the looking tables send (rows whose filter is 1) must equal the multiset the looked table holds; the Memory
// lookup's looking side also gets `extra_rows` (n_extra x 13: get_memory_extra_looking_values, verifier.rs:547-737).  A filter that is
// neither 0 nor 1 is an error, as there.  -> total number of distinct rows whose multiplicities differ, per lookup in mismatches[10];
// -1 on error.  Tables left out (null trace) send and hold nothing.
long orc_check_ctls(const uint64_t* const* traces, const size_t* ns, const uint64_t* extra_rows, size_t n_extra, size_t mismatches[10]) {
    try {
        auto ctls = zkstark::all_cross_table_lookups();
        if (ctls.size() != zkstark::NUM_CTLS) throw std::runtime_error("all_cross_table_lookups().len() != NUM_CTLS");   // all_stark.rs:451-454 check_num_ctls
        std::vector<std::vector<const uint64_t*>> cols(9);
        for (uint32_t t = 0; t < 9; t++) {
            if (!traces[t]) continue;
            size_t nc = zkstark::table_num_columns(t);
            cols[t].resize(nc);
            for (size_t c = 0; c < nc; c++) cols[t][c] = traces[t] + c * ns[t];
        }
        long total = 0;
        for (size_t ci = 0; ci < ctls.size(); ci++) {
            std::map<std::vector<uint64_t>, long> diff;
            auto process = [&](const TableWithColumns& tc, long sign) {
                if (tc.table >= 9 || !traces[tc.table]) return;
                const uint64_t* const* tr = cols[tc.table].data();
                const size_t n = ns[tc.table];
                std::vector<uint64_t> row(tc.columns.size());
                for (size_t r = 0; r < n; r++) {
                    const uint64_t f = filter_eval_table(tc.filter, tr, n, r);
                    if (f == 0) continue;
                    if (f != 1) throw std::runtime_error("Non-binary filter? (lookup " + std::to_string(ci) + ", table " + std::to_string(tc.table) + ", row " + std::to_string(r) + ")");
                    for (size_t k = 0; k < row.size(); k++) row[k] = col_eval_table(tc.columns[k], tr, n, r);
                    diff[row] += sign;
                }
            };
            for (auto& l : ctls[ci].looking_tables) process(l, +1);
            process(ctls[ci].looked_table, -1);
            if (ci == zkstark::MEMORY_CTL_IDX)
                for (size_t e = 0; e < n_extra; e++) diff[std::vector<uint64_t>(extra_rows + 13 * e, extra_rows + 13 * e + 13)] += 1;
            size_t bad = 0;
            for (auto& kv : diff) bad += kv.second != 0;
            if (mismatches) mismatches[ci] = bad;
            total += (long)bad;
        }
        return total;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// The reference's second per-table sanity test, starky `test_stark_circuit_constraints`, compares the packed evaluator with the
// recursion-circuit one; the circuit evaluators are outside this path, but the property it protects — ONE set of constraints, whatever
// type it is evaluated in — has an analogue here, where the same templates are evaluated over the base field (prover: LDE points) and
// over the quadratic extension (verifier: at zeta).  On a random frame: (1) evaluating in the extension at base-field points gives the
// embedded base-field result; (2) the constraints have base-field coefficients, so they commute with the extension's automorphism
// a + bX -> a - bX: C(conj(frame)) = conj(C(frame)) for a random EXTENSION frame and extension selectors.
// -> 0 when both hold, bit 0 / bit 1 for a failure of (1) / (2); -1 on error.
long orc_table_eval_consistency(uint32_t table, uint64_t seed, const uint64_t labels[4]) {
    try {
        const size_t ncols = zkstark::table_num_columns(table);
        if (ncols == 0) throw std::runtime_error("bad table");
        uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        auto rnd = [&]() { st += 0x9E3779B97F4A7C15ull; uint64_t z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                           return gl_from_u64(z ^ (z >> 31)); };
        auto conj = [](Ext x) { return Ext(x.a, gl_neg(x.b)); };
        const uint64_t alpha = rnd();
        auto run_ext = [&](const std::vector<Ext>& l, const std::vector<Ext>& n, Ext zl, Ext lf, Ext ll) {
            std::vector<OE> lv(ncols), nv(ncols);
            for (size_t c = 0; c < ncols; c++) { lv[c] = OE(l[c]); nv[c] = OE(n[c]); }
            ConsumerT<OE> yc;
            yc.alphas.push_back(OE(Ext(alpha, 0))); yc.acc.push_back(OE::zero());
            yc.z_last = OE(zl); yc.lagrange_first = OE(lf); yc.lagrange_last = OE(ll);
            RowOE rl{lv.data()}, rn{nv.data()};
            zkstark::eval_table<OE>(table, rl, rn, yc, params_from(labels));
            return yc.acc[0].e;
        };
        long bad = 0;
        // (1)
        std::vector<uint64_t> lb(ncols), nb(ncols);
        for (auto& x : lb) x = rnd();
        for (auto& x : nb) x = rnd();
        const uint64_t zl = rnd(), lf = rnd(), ll = rnd();
        ConsumerT<OF> yb;
        yb.alphas.push_back(OF(alpha)); yb.acc.push_back(OF(0));
        yb.z_last = OF(zl); yb.lagrange_first = OF(lf); yb.lagrange_last = OF(ll);
        RowOF rl{lb.data()}, rn{nb.data()};
        zkstark::eval_table<OF>(table, rl, rn, yb, params_from(labels));
        std::vector<Ext> le(ncols), ne(ncols);
        for (size_t c = 0; c < ncols; c++) { le[c] = Ext(lb[c], 0); ne[c] = Ext(nb[c], 0); }
        if (run_ext(le, ne, Ext(zl, 0), Ext(lf, 0), Ext(ll, 0)) != Ext(yb.acc[0].v, 0)) bad |= 1;
        if (yb.acc[0].v == 0) bad |= 4;      // a random frame does not satisfy the constraints
        // (2)
        for (size_t c = 0; c < ncols; c++) { le[c] = Ext(rnd(), rnd()); ne[c] = Ext(rnd(), rnd()); }
        const Ext ezl(rnd(), rnd()), elf(rnd(), rnd()), ell(rnd(), rnd());
        const Ext v = run_ext(le, ne, ezl, elf, ell);
        for (size_t c = 0; c < ncols; c++) { le[c] = conj(le[c]); ne[c] = conj(ne[c]); }
        if (run_ext(le, ne, conj(ezl), conj(elf), conj(ell)) != conj(v)) bad |= 2;
        if (v.b == 0) bad |= 4;              // the extension part is really exercised
        return bad;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// "stage\tseconds\n" lines accumulated since the last call with reset != 0 (main-thread wall clock of the prover's stages)
size_t orc_stage_report(char* buf, size_t cap, int reset) {
    std::string out;
    for (auto& kv : StageClock::acc()) { char line[128]; snprintf(line, sizeof line, "%s\t%.4f\n", kv.first.c_str(), kv.second); out += line; }
    if (reset) StageClock::acc().clear();
    if (buf && cap) { size_t l = out.size() < cap - 1 ? out.size() : cap - 1; memcpy(buf, out.data(), l); buf[l] = 0; }
    return out.size();
}

long orc_prove_segment(const uint32_t cfgw[8], const uint64_t* const* traces, const size_t* ns, const uint64_t* public_values, size_t npv,
                       const uint64_t labels[4], const uint64_t* forced_pows, uint64_t* out, size_t out_cap, size_t offsets[10],
                       uint64_t* beta_gamma_out, uint64_t* caps_out) {
    try {
        Config cfg = cfg_from(cfgw);
        const size_t capw = (size_t)4 << cfg.cap_height;
        auto ctls = zkstark::all_cross_table_lookups();
        uint8_t in_use[9];
        std::vector<std::vector<const uint64_t*>> cols(9);
        std::vector<PolyBatch> commits(9);
        std::vector<Words> caps(9);
        for (uint32_t t = 0; t < 9; t++) {
            in_use[t] = traces[t] != nullptr;
            if (!in_use[t]) { caps[t].assign(capw, 0); continue; }
            size_t nc = zkstark::table_num_columns(t);
            cols[t].resize(nc);
            for (size_t c = 0; c < nc; c++) cols[t][c] = traces[t] + c * ns[t];
            { ORC_STAGE("trace commit"); commits[t].from_values(cols[t].data(), nc, ns[t], cfg.rate_bits, cfg.cap_height); }
            caps[t] = cap_words(commits[t].tree);
        }
        if (caps_out) for (uint32_t t = 0; t < 9; t++) memcpy(caps_out + t * capw, caps[t].data(), capw * 8);
        Challenger ch;
        std::vector<uint64_t> betas, gammas;
        segment_transcript(ch, caps, in_use, capw, public_values, npv, cfg.num_challenges, betas, gammas);
        if (beta_gamma_out) for (unsigned i = 0; i < cfg.num_challenges; i++) { beta_gamma_out[2 * i] = betas[i]; beta_gamma_out[2 * i + 1] = gammas[i]; }
        std::vector<Words> proofs(9);
        for (uint32_t t = 0; t < 9; t++) {
            if (!in_use[t]) continue;
            CtlData ctl;
            { ORC_STAGE("ctl data"); ctl = ctl_data_for_table(t, cols[t].data(), ns[t], ctls, betas, gammas, zkstark::CONSTRAINT_DEGREE); }
            StarkProofData p = prove_table(t, cfg, cols[t].data(), ns[t], commits[t], ctl, betas, gammas, ch, params_from(labels),
                                           forced_pows ? &forced_pows[t] : nullptr, ctls, nullptr);
            proofs[t] = zkstark::serialize_proof(p);
            commits[t] = PolyBatch();
        }
        size_t total = 0;
        for (uint32_t t = 0; t < 9; t++) { offsets[t] = total; total += proofs[t].size(); }
        offsets[9] = total;
        if (out && out_cap >= total) for (uint32_t t = 0; t < 9; t++) if (!proofs[t].empty()) memcpy(out + offsets[t], proofs[t].data(), proofs[t].size() * 8);
        return (long)total;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

// verify_proof: re-derive every challenge from the proofs (get_challenges.rs:273-314), verify each table's proof, then check the
// cross-table-lookup sums (starky verify_cross_table_lookups) with the caller's extra looking sums (verifier.rs:240-262:
// extra[ctl_index * num_challenges + c], only the memory CTL is non-zero in the reference).  1 valid, 0 rejected, -1 malformed.
int orc_verify_segment(const uint32_t cfgw[8], const uint64_t* proofs, const size_t offsets[10], const uint64_t* public_values, size_t npv,
                       const uint64_t labels[4], const uint64_t* extra_looking_sums) {
    try {
        Config cfg = cfg_from(cfgw);
        const size_t capw = (size_t)4 << cfg.cap_height;
        auto ctls = zkstark::all_cross_table_lookups();
        uint8_t in_use[9];
        std::vector<StarkProofData> ps(9);
        std::vector<Words> caps(9);
        for (uint32_t t = 0; t < 9; t++) {
            in_use[t] = offsets[t + 1] > offsets[t];
            if (!in_use[t]) continue;
            ps[t] = zkstark::deserialize_proof(proofs + offsets[t], offsets[t + 1] - offsets[t]);
            if (ps[t].trace_cap.size() != capw) throw std::runtime_error("trace cap has the wrong size");
            caps[t] = ps[t].trace_cap;
        }
        Challenger ch;
        std::vector<uint64_t> betas, gammas;
        segment_transcript(ch, caps, in_use, capw, public_values, npv, cfg.num_challenges, betas, gammas);
        for (uint32_t t = 0; t < 9; t++) {
            if (!in_use[t]) continue;
            std::string err;
            if (!verify_table(t, cfg, ps[t], betas, gammas, ch, params_from(labels), ctls, err)) {
                g_orc_err = std::string(zkstark::table_name(t)) + ": " + err;
                return 0;
            }
        }
        // verify_cross_table_lookups
        std::vector<size_t> pos(9, 0);
        std::vector<Words> zs(9);
        for (uint32_t t = 0; t < 9; t++) {
            if (in_use[t]) zs[t] = ps[t].ctl_zs_first;
            else zs[t].assign(aux_shape(t, ctls, cfg.num_challenges, zkstark::CONSTRAINT_DEGREE).ctl_entries.size(), 0);   // verifier.rs:283-293
        }
        std::vector<size_t> failed;          // every failing lookup is reported (the reference stops at the first one): tests of partial instances
        std::string first_msg;
        for (size_t ci = 0; ci < ctls.size(); ci++) {
            std::vector<uint32_t> lookers;
            for (auto& l : ctls[ci].looking_tables) { bool seen = false; for (uint32_t x : lookers) seen |= x == l.table; if (!seen) lookers.push_back(l.table); }
            for (unsigned c = 0; c < cfg.num_challenges; c++) {
                uint64_t sum = extra_looking_sums ? extra_looking_sums[ci * cfg.num_challenges + c] : 0;
                for (uint32_t t : lookers) { if (pos[t] >= zs[t].size()) throw std::runtime_error("ctl_zs_first too short"); sum = gl_add(sum, zs[t][pos[t]++]); }
                uint32_t lt = ctls[ci].looked_table.table;
                if (pos[lt] >= zs[lt].size()) throw std::runtime_error("ctl_zs_first too short");
                uint64_t looked = zs[lt][pos[lt]++];
                if (sum != looked) {
                    if (failed.empty()) first_msg = "Cross-table lookup " + std::to_string(ci) + " verification failed (challenge " + std::to_string(c) + ")";
                    if (failed.empty() || failed.back() != ci) failed.push_back(ci);
                }
            }
        }
        if (!failed.empty()) {
            g_orc_err = first_msg + "; failing lookups:";
            for (size_t ci : failed) g_orc_err += " " + std::to_string(ci);
            return 0;
        }
        for (uint32_t t = 0; t < 9; t++) if (pos[t] != zs[t].size()) throw std::runtime_error("unused ctl_zs_first entries");
        g_orc_err.clear();
        return 1;
    } catch (const std::exception& e) { g_orc_err = e.what(); return -1; }
}

}  // extern "C"
