// ORACLE (test infrastructure, NOT product code) — CPU restatement of the per-table STARK prover and verifier.
//
// Follows starky 1.0.0 (prover.rs `prove_with_commitment`, `compute_quotient_polys`; lookup.rs; cross_table_lookup.rs
// `get_ctl_data`, `cross_table_lookup_data`, `partial_sums`, `eval_cross_table_lookup_checks`; verifier.rs
// `verify_stark_proof_with_challenges`) and plonky2 1.0.0 (fri/oracle.rs `prove_openings`; fri/prover.rs; fri/verifier.rs)
// as reached from /root/reference/evm_arithmetization/src/prover.rs:72-341 and verifier.rs:172-313.  Those crates are not
// vendored in the reference (Cargo.lock:3702-3755,4740-4752): the algorithms are restated from their published sources.
// Parity unpinned (the reference holds no golden proofs); the verifier below is the self-consistency check.
//
// The per-table constraint definitions (stark/table_*.h) are single-source templates shared with the product; everything
// else here (aux-column generation, lookup/CTL checks, quotient loop, openings, FRI, transcript) is written independently
// of the product's CUDA path, directly against the Column/Filter/Lookup/CrossTableLookup descriptions.
#pragma once
#include "oracle_core.h"
#include "oracle_poly.h"
#include "stark/all_stark.h"
#include "stark/proof.h"
#include <string>
#include <algorithm>
#include <functional>

namespace orc {

using zkstark::Column; using zkstark::Filter; using zkstark::Lookup; using zkstark::CrossTableLookup;
using zkstark::TableWithColumns; using zkstark::StarkProofData; using zkstark::Words; using zkstark::Config;
using zkstark::TableParams;

// ---- evaluation types for the constraint templates -----------------------------------------------------------
struct OF {   // base field
    uint64_t v;
    OF() : v(0) {}
    explicit OF(uint64_t x) : v(x) {}
    static OF zero() { return OF(0); }
    static OF one() { return OF(1); }
    static OF from_u64(uint64_t x) { return OF(gl_from_u64(x)); }
};
static inline OF operator+(OF a, OF b) { return OF(gl_add(a.v, b.v)); }
static inline OF operator-(OF a, OF b) { return OF(gl_sub(a.v, b.v)); }
static inline OF operator*(OF a, OF b) { return OF(gl_mul(a.v, b.v)); }
static inline OF operator-(OF a) { return OF(gl_neg(a.v)); }
static inline OF& operator+=(OF& a, OF b) { a = a + b; return a; }
static inline OF& operator-=(OF& a, OF b) { a = a - b; return a; }
static inline OF& operator*=(OF& a, OF b) { a = a * b; return a; }
struct OE {   // extension field (verifier side)
    Ext e;
    OE() {}
    explicit OE(Ext x) : e(x) {}
    static OE zero() { return OE(Ext(0, 0)); }
    static OE one() { return OE(Ext(1, 0)); }
    static OE from_u64(uint64_t x) { return OE(Ext(gl_from_u64(x), 0)); }
};
static inline OE operator+(OE a, OE b) { return OE(a.e + b.e); }
static inline OE operator-(OE a, OE b) { return OE(a.e - b.e); }
static inline OE operator*(OE a, OE b) { return OE(a.e * b.e); }
static inline OE operator-(OE a) { return OE(Ext(0, 0) - a.e); }
static inline OE& operator+=(OE& a, OE b) { a = a + b; return a; }
static inline OE& operator-=(OE& a, OE b) { a = a - b; return a; }
static inline OE& operator*=(OE& a, OE b) { a = a * b; return a; }

struct RowOF { const uint64_t* p; OF operator[](uint32_t c) const { return OF(p[c]); } };
struct RowOE { const OE* p; OE operator[](uint32_t c) const { return p[c]; } };

// ---- Column / Filter evaluation -------------------------------------------------------------------------------
// Column::eval_table: next-row terms count as 0 on the last row
static inline uint64_t col_eval_table(const Column& c, const uint64_t* const* trace, size_t n, size_t row) {
    uint64_t r = c.constant;
    for (auto& t : c.lin) r = gl_add(r, gl_mul(trace[t.first][row], t.second));
    if (!c.next.empty() && row + 1 < n)
        for (auto& t : c.next) r = gl_add(r, gl_mul(trace[t.first][row + 1], t.second));
    return r;
}
static inline uint64_t filter_eval_table(const Filter& f, const uint64_t* const* trace, size_t n, size_t row) {
    uint64_t r = 0;
    for (auto& pr : f.products) r = gl_add(r, gl_mul(col_eval_table(pr.first, trace, n, row), col_eval_table(pr.second, trace, n, row)));
    for (auto& c : f.constants) r = gl_add(r, col_eval_table(c, trace, n, row));
    return r;
}
// Column::eval_with_next on an evaluation frame
template <class P, class V> static inline P col_eval(const Column& c, const V& lv, const V& nv) {
    P r = P::from_u64(c.constant);
    for (auto& t : c.lin) r = r + lv[t.first] * P::from_u64(t.second);
    for (auto& t : c.next) r = r + nv[t.first] * P::from_u64(t.second);
    return r;
}
template <class P, class V> static inline P filter_eval(const Filter& f, const V& lv, const V& nv) {
    P r = P::zero();
    for (auto& pr : f.products) r = r + col_eval<P>(pr.first, lv, nv) * col_eval<P>(pr.second, lv, nv);
    for (auto& c : f.constants) r = r + col_eval<P>(c, lv, nv);
    return r;
}
template <class P, class V> static inline P combine(const std::vector<Column>& cols, P beta, P gamma, const V& lv, const V& nv) {
    P acc = P::zero();
    for (size_t k = cols.size(); k-- > 0;) acc = acc * beta + col_eval<P>(cols[k], lv, nv);
    return acc + gamma;
}

// ---- CTL data ------------------------------------------------------------------------------------------------
typedef std::pair<std::vector<Column>, Filter> ColumnsFilter;
struct CtlZData {
    std::vector<std::vector<uint64_t>> helpers;
    std::vector<uint64_t> z;
    unsigned challenge = 0;
    std::vector<ColumnsFilter> entries;
};
typedef std::vector<CtlZData> CtlData;

// get_helper_cols: h(r) = sum over the chunk of filter/combined (filters are 0/1 in valid traces; a non-binary filter,
// which the reference rejects with "Non-binary filter?", is generalised here as f * 1/combined so that random traces
// can be used for throughput runs — identical on every input the reference accepts)
static inline std::vector<std::vector<uint64_t>> get_helper_cols(const uint64_t* const* trace, size_t n,
                                                                 const std::vector<ColumnsFilter>& cf, uint64_t beta,
                                                                 uint64_t gamma, unsigned constraint_degree) {
    size_t chunk = constraint_degree - 1;
    size_t nh = (cf.size() + chunk - 1) / chunk;
    std::vector<std::vector<uint64_t>> helpers(nh, std::vector<uint64_t>(n, 0));
    // denominators of one (helper, term) for a block of rows, inverted together (Montgomery's trick, as the reference's
    // batch_multiplicative_inverse does); rows whose filter is 0 contribute nothing and are skipped as before
    const size_t BLK = 4096;
    for (size_t h = 0; h < nh; h++) {
        #pragma omp parallel
        {
            std::vector<uint64_t> den(BLK), fil(BLK), pre(BLK);
            #pragma omp for schedule(static)
            for (size_t r0 = 0; r0 < n; r0 += BLK) {
                const size_t cnt = std::min(BLK, n - r0);
                for (size_t k = h * chunk; k < std::min(cf.size(), (h + 1) * chunk); k++) {
                    for (size_t i = 0; i < cnt; i++) {
                        const size_t r = r0 + i;
                        uint64_t f = filter_eval_table(cf[k].second, trace, n, r);
                        fil[i] = f;
                        uint64_t comb = 1;
                        if (f != 0) {
                            comb = 0;
                            for (size_t j = cf[k].first.size(); j-- > 0;) comb = gl_add(gl_mul(comb, beta), col_eval_table(cf[k].first[j], trace, n, r));
                            comb = gl_add(comb, gamma);
                        }
                        den[i] = comb;
                    }
                    // prefix products over the non-zero denominators (inverse(0) = 0 as Field::inverse_or_zero would not be hit here:
                    // a zero denominator keeps its slot out of the chain and yields 0)
                    uint64_t acc = 1;
                    for (size_t i = 0; i < cnt; i++) { pre[i] = acc; if (den[i] != 0) acc = gl_mul(acc, den[i]); }
                    uint64_t inv = gl_inv(acc);
                    for (size_t i = cnt; i-- > 0;) {
                        if (den[i] == 0) continue;
                        const uint64_t di = gl_mul(inv, pre[i]);
                        inv = gl_mul(inv, den[i]);
                        if (fil[i] != 0) helpers[h][r0 + i] = gl_add(helpers[h][r0 + i], gl_mul(fil[i], di));
                    }
                }
            }
        }
    }
    return helpers;
}
// partial_sums: Z(n-1) = sum_t h_t(n-1), Z(r) = Z(r+1) + sum_t h_t(r); helpers kept only when there is more than one pair
static inline void partial_sums(const uint64_t* const* trace, size_t n, const std::vector<ColumnsFilter>& cf, uint64_t beta,
                                uint64_t gamma, unsigned constraint_degree, CtlZData& out) {
    auto helpers = get_helper_cols(trace, n, cf, beta, gamma, constraint_degree);
    out.z.assign(n, 0);
    uint64_t acc = 0;
    for (size_t r = n; r-- > 0;) {
        for (auto& h : helpers) acc = gl_add(acc, h[r]);
        out.z[r] = acc;
    }
    if (cf.size() > 1) out.helpers = std::move(helpers); else out.helpers.clear();
    out.entries = cf;
}
// cross_table_lookup_data for ONE table (the per-table slice of what get_ctl_data returns)
static inline CtlData ctl_data_for_table(uint32_t table, const uint64_t* const* trace, size_t n,
                                         const std::vector<CrossTableLookup>& ctls, const std::vector<uint64_t>& betas,
                                         const std::vector<uint64_t>& gammas, unsigned constraint_degree) {
    CtlData data;
    for (const CrossTableLookup& ctl : ctls) {
        for (size_t ch = 0; ch < betas.size(); ch++) {
            // itertools group_by over consecutive looking tables
            size_t i = 0;
            const auto& lt = ctl.looking_tables;
            while (i < lt.size()) {
                size_t j = i;
                std::vector<ColumnsFilter> group;
                while (j < lt.size() && lt[j].table == lt[i].table) { group.push_back({lt[j].columns, lt[j].filter}); j++; }
                if (lt[i].table == table) {
                    CtlZData z; z.challenge = (unsigned)ch;
                    partial_sums(trace, n, group, betas[ch], gammas[ch], constraint_degree, z);
                    data.push_back(std::move(z));
                }
                i = j;
            }
            if (ctl.looked_table.table == table) {
                CtlZData z; z.challenge = (unsigned)ch;
                partial_sums(trace, n, {{ctl.looked_table.columns, ctl.looked_table.filter}}, betas[ch], gammas[ch], constraint_degree, z);
                data.push_back(std::move(z));
            }
        }
    }
    return data;
}

// lookup_helper_columns (starky lookup.rs): helpers..., Z with Z(0) = 0, Z(r+1) = Z(r) + sum_t h_t(r) - freq(r)/(table(r)+challenge)
static inline std::vector<std::vector<uint64_t>> lookup_helper_columns(const Lookup& l, const uint64_t* const* trace, size_t n,
                                                                       uint64_t challenge, unsigned constraint_degree) {
    std::vector<ColumnsFilter> cf;
    for (size_t i = 0; i < l.columns.size(); i++) cf.push_back({{l.columns[i]}, l.filter_columns[i]});
    auto cols = get_helper_cols(trace, n, cf, 1, challenge, constraint_degree);
    std::vector<uint64_t> z(n, 0), tinv(n);
    #pragma omp parallel for schedule(static)
    for (size_t r0 = 0; r0 < n; r0 += 4096) {
        const size_t cnt = std::min((size_t)4096, n - r0);
        std::vector<uint64_t> den(cnt), pre(cnt);
        uint64_t acc = 1;
        for (size_t i = 0; i < cnt; i++) {
            den[i] = gl_add(challenge, col_eval_table(l.table_column, trace, n, r0 + i));
            pre[i] = acc;
            if (den[i] != 0) acc = gl_mul(acc, den[i]);
        }
        uint64_t inv = gl_inv(acc);
        for (size_t i = cnt; i-- > 0;) {
            if (den[i] == 0) { tinv[r0 + i] = 0; continue; }
            tinv[r0 + i] = gl_mul(inv, pre[i]);
            inv = gl_mul(inv, den[i]);
        }
    }
    for (size_t r = 0; r + 1 < n; r++) {
        uint64_t x = 0;
        for (auto& h : cols) x = gl_add(x, h[r]);
        x = gl_sub(x, gl_mul(col_eval_table(l.frequencies_column, trace, n, r), tinv[r]));
        z[r + 1] = gl_add(z[r], x);
    }
    cols.push_back(std::move(z));
    return cols;
}

// ---- constraint consumer ---------------------------------------------------------------------------------------
template <class P> struct ConsumerT {
    std::vector<P> alphas, acc;
    P z_last, lagrange_first, lagrange_last;
    void constraint(P c) { for (size_t j = 0; j < acc.size(); j++) acc[j] = acc[j] * alphas[j] + c; }
    // index-addressed block (see stark/consumer.h): only used to check the device's reordered evaluators against the sequential ones
    std::vector<std::vector<P>> apow;
    uint32_t blk_top = 0;
    void block_begin(uint32_t M) {
        apow.assign(acc.size(), std::vector<P>());
        for (size_t j = 0; j < acc.size(); j++) {
            P w = P::one();
            for (uint32_t e = 0; e <= M; e++) { apow[j].push_back(w); w = w * alphas[j]; }
            acc[j] = acc[j] * apow[j][M];
        }
        blk_top = M - 1;
    }
    void block_put(uint32_t idx, P c) { for (size_t j = 0; j < acc.size(); j++) acc[j] = acc[j] + c * apow[j][blk_top - idx]; }
    void constraint_transition(P c) { constraint(c * z_last); }
    void constraint_first_row(P c) { constraint(c * lagrange_first); }
    void constraint_last_row(P c) { constraint(c * lagrange_last); }
};

// eval_helper_columns
template <class P, class V, class CC>
static inline void eval_helper_columns(const std::vector<ColumnsFilter>& cf, const std::vector<P>& helpers, P beta, P gamma,
                                       const V& lv, const V& nv, CC& yc) {
    for (size_t t = 0; t < helpers.size(); t++) {
        size_t k0 = 2 * t;
        P c0 = combine<P>(cf[k0].first, beta, gamma, lv, nv), f0 = filter_eval<P>(cf[k0].second, lv, nv);
        if (k0 + 1 < cf.size()) {
            P c1 = combine<P>(cf[k0 + 1].first, beta, gamma, lv, nv), f1 = filter_eval<P>(cf[k0 + 1].second, lv, nv);
            yc.constraint(c1 * c0 * helpers[t] - f0 * c1 - f1 * c0);
        } else {
            yc.constraint(c0 * helpers[t] - f0);
        }
    }
}

// Shape of one table's auxiliary columns
struct AuxShape {
    std::vector<Lookup> lookups;
    size_t num_lookup_cols = 0;
    std::vector<size_t> ctl_helper_counts;   // per CtlZData
    std::vector<std::vector<ColumnsFilter>> ctl_entries;
    std::vector<unsigned> ctl_challenge;
    size_t num_ctl_helpers = 0;
    size_t num_aux() const { return num_lookup_cols + num_ctl_helpers + ctl_helper_counts.size(); }
};
static inline AuxShape aux_shape(uint32_t table, const std::vector<CrossTableLookup>& ctls, unsigned num_challenges, unsigned cd) {
    AuxShape s;
    s.lookups = zkstark::table_lookups(table);
    for (auto& l : s.lookups) s.num_lookup_cols += num_challenges * l.num_helper_columns(cd);
    // same walk as ctl_data_for_table, shapes only
    for (const CrossTableLookup& ctl : ctls)
        for (unsigned ch = 0; ch < num_challenges; ch++) {
            size_t i = 0; const auto& lt = ctl.looking_tables;
            while (i < lt.size()) {
                size_t j = i; std::vector<ColumnsFilter> group;
                while (j < lt.size() && lt[j].table == lt[i].table) { group.push_back({lt[j].columns, lt[j].filter}); j++; }
                if (lt[i].table == table) {
                    size_t nh = group.size() > 1 ? (group.size() + cd - 2) / (cd - 1) : 0;
                    s.ctl_helper_counts.push_back(nh); s.num_ctl_helpers += nh; s.ctl_entries.push_back(group); s.ctl_challenge.push_back(ch);
                }
                i = j;
            }
            if (ctl.looked_table.table == table) {
                s.ctl_helper_counts.push_back(0); s.ctl_entries.push_back({{ctl.looked_table.columns, ctl.looked_table.filter}});
                s.ctl_challenge.push_back(ch);
            }
        }
    return s;
}

// eval_vanishing_poly: table constraints, then lookup checks, then CTL checks
template <class P, class V, class A>
static inline void eval_vanishing_poly(uint32_t table, const AuxShape& sh, const std::vector<P>& betas, const std::vector<P>& gammas,
                                       const V& lv, const V& nv, const A& aux_lv, const A& aux_nv, ConsumerT<P>& yc,
                                       const TableParams& prm, unsigned cd) {
    zkstark::eval_table<P>(table, lv, nv, yc, prm);
    size_t start = 0;
    for (const Lookup& l : sh.lookups) {
        size_t nhc = l.num_helper_columns(cd);
        std::vector<ColumnsFilter> cf;
        for (size_t i = 0; i < l.columns.size(); i++) cf.push_back({{l.columns[i]}, l.filter_columns[i]});
        for (size_t ch = 0; ch < betas.size(); ch++) {
            P challenge = betas[ch];
            std::vector<P> helpers;
            for (size_t t = 0; t + 1 < nhc; t++) helpers.push_back(aux_lv[start + t]);
            eval_helper_columns<P>(cf, helpers, P::one(), challenge, lv, nv, yc);
            P z = aux_lv[start + nhc - 1], next_z = aux_nv[start + nhc - 1];
            P twc = col_eval<P>(l.table_column, lv, nv) + challenge;
            P hsum = P::zero();
            for (auto& h : helpers) hsum = hsum + h;
            P y = hsum * twc - col_eval<P>(l.frequencies_column, lv, nv);
            yc.constraint_first_row(z);
            yc.constraint((next_z - z) * twc - y);
            start += nhc;
        }
    }
    size_t hstart = sh.num_lookup_cols, zstart = sh.num_lookup_cols + sh.num_ctl_helpers;
    for (size_t i = 0; i < sh.ctl_entries.size(); i++) {
        size_t nh = sh.ctl_helper_counts[i];
        P beta = betas[sh.ctl_challenge[i]], gamma = gammas[sh.ctl_challenge[i]];
        P local_z = aux_lv[zstart + i], next_z = aux_nv[zstart + i];
        std::vector<P> helpers;
        for (size_t t = 0; t < nh; t++) helpers.push_back(aux_lv[hstart + t]);
        eval_helper_columns<P>(sh.ctl_entries[i], helpers, beta, gamma, lv, nv, yc);
        if (nh) {
            P hsum = P::zero();
            for (auto& h : helpers) hsum = hsum + h;
            yc.constraint_last_row(local_z - hsum);
            yc.constraint_transition(local_z - next_z - hsum);
        } else {
            P c0 = combine<P>(sh.ctl_entries[i][0].first, beta, gamma, lv, nv);
            P f0 = filter_eval<P>(sh.ctl_entries[i][0].second, lv, nv);
            yc.constraint_last_row(c0 * local_z - f0);
            yc.constraint_transition(c0 * (local_z - next_z) - f0);
        }
        hstart += nh;
    }
}

// ---- helpers ---------------------------------------------------------------------------------------------------
static inline void push_ext(Words& w, Ext e) { w.push_back(e.a); w.push_back(e.b); }
static inline void observe_words(Challenger& ch, const Words& w) { ch.observe_n(w.data(), w.size()); }
static inline Words cap_words(const MerkleTree& t) { Words w(t.cap_len() * 4); memcpy(w.data(), t.cap(), w.size() * 8); return w; }
static inline Words path_words(const std::vector<Hash>& p) { Words w(p.size() * 4); if (!p.empty()) memcpy(w.data(), p.data(), w.size() * 8); return w; }
static inline Ext ext_from_base(uint64_t x) { return Ext(x, 0); }

// fft over the extension field = componentwise base-field fft
static inline void ext_coset_fft(std::vector<Ext>& a, unsigned log_n, uint64_t shift) {
    size_t n = a.size();
    std::vector<uint64_t> re(n), im(n);
    for (size_t i = 0; i < n; i++) { re[i] = a[i].a; im[i] = a[i].b; }
    coset_fft_inplace(re.data(), log_n, shift);
    coset_fft_inplace(im.data(), log_n, shift);
    for (size_t i = 0; i < n; i++) a[i] = Ext(re[i], im[i]);
}

// ---- prover ----------------------------------------------------------------------------------------------------
struct ProveDebug {   // optional intermediate outputs for stage-by-stage parity tests
    std::vector<std::vector<uint64_t>> aux_values;      // aux columns (values)
    std::vector<uint64_t> quotient_chunk_coeffs;        // q columns x n, column-major
    std::vector<Ext> fri_final_values;                  // values of the FRI input polynomial, bit-reversed order
};

// starky prove_with_commitment (called by prove_single_table, prover.rs:301-341).  `ch` is the shared transcript.
static inline StarkProofData prove_table(uint32_t table, const Config& cfg, const uint64_t* const* trace, size_t n,
                                         const PolyBatch& trace_commit, const CtlData& ctl, const std::vector<uint64_t>& betas,
                                         const std::vector<uint64_t>& gammas, Challenger& ch, const TableParams& prm,
                                         const uint64_t* forced_pow, const std::vector<CrossTableLookup>& ctls,
                                         ProveDebug* dbg = nullptr) {
    const unsigned cd = zkstark::CONSTRAINT_DEGREE;
    const unsigned k = trace_commit.log_n, rate_bits = cfg.rate_bits;
    const size_t N = n << rate_bits;
    const size_t ncols = trace_commit.ncols;
    StarkProofData proof;
    proof.table_id = table; proof.degree_bits = k;
    ch.compact();
    memcpy(proof.init_challenger_state, ch.state, 96);
    std::vector<unsigned> arities = zkstark::fri_reduction_arity_bits(cfg, k);
    {
        unsigned tot = 0; for (unsigned a : arities) tot += a;
        if (tot > k + rate_bits - cfg.cap_height) throw std::runtime_error("FRI total arity is too large");
    }
    AuxShape sh = aux_shape(table, ctls, cfg.num_challenges, cd);

    // 1. auxiliary polynomials: lookup columns, CTL helpers, CTL Zs
    std::vector<std::vector<uint64_t>> aux;
    { ORC_STAGE("aux columns");
    for (const Lookup& l : sh.lookups)
        for (unsigned c = 0; c < cfg.num_challenges; c++) {
            auto cols = lookup_helper_columns(l, trace, n, betas[c], cd);
            for (auto& col : cols) aux.push_back(std::move(col));
        }
    if (ctl.size() != sh.ctl_entries.size()) throw std::runtime_error("ctl data does not match the table's CTL shape");
    for (auto& z : ctl) for (auto& h : z.helpers) aux.push_back(h);
    for (auto& z : ctl) aux.push_back(z.z);
    }
    const size_t na = aux.size();
    if (na != sh.num_aux()) throw std::runtime_error("aux column count mismatch");
    if (dbg) dbg->aux_values = aux;
    PolyBatch aux_commit;
    if (na) {
        ORC_STAGE("aux commit");
        std::vector<const uint64_t*> ptr(na);
        for (size_t i = 0; i < na; i++) ptr[i] = aux[i].data();
        aux_commit.from_values(ptr.data(), na, n, rate_bits, cfg.cap_height);
        proof.aux_cap = cap_words(aux_commit.tree);
        observe_words(ch, proof.aux_cap);
    }
    proof.trace_cap = cap_words(trace_commit.tree);

    // 2. alphas
    std::vector<uint64_t> alphas(cfg.num_challenges);
    for (auto& a : alphas) a = ch.challenge();

    // 3. quotient polynomials (compute_quotient_polys): quotient_degree_bits == rate_bits == 1 -> step 1, next_step 2
    if (rate_bits != 1) throw std::runtime_error("oracle prover implements rate_bits = 1 (quotient_degree_factor 2)");
    const unsigned logN = k + rate_bits;
    const size_t next_step = 2;
    uint64_t wN = gl_root_of_unity(logN), wn = gl_root_of_unity(k);
    uint64_t last = gl_inv(wn);
    // Lagrange selectors on the coset via their LDE, as the reference does
    std::vector<uint64_t> lag_first(N, 0), lag_last(N, 0);
    {
        std::vector<uint64_t> e(n, 0);
        e[0] = 1; ifft_inplace(e.data(), k); std::copy(e.begin(), e.end(), lag_first.begin());
        coset_fft_inplace(lag_first.data(), logN, GL_GENERATOR);
        std::fill(e.begin(), e.end(), 0);
        e[n - 1] = 1; ifft_inplace(e.data(), k); std::copy(e.begin(), e.end(), lag_last.begin());
        coset_fft_inplace(lag_last.data(), logN, GL_GENERATOR);
    }
    uint64_t gn = gl_pow(GL_GENERATOR, n);
    uint64_t zh_inv[2] = {gl_inv(gl_sub(gn, 1)), gl_inv(gl_sub(gl_neg(gn), 1))};   // Z_H(g w^i) = g^n (-1)^i - 1
    std::vector<OF> betasP, gammasP;
    for (auto b : betas) betasP.push_back(OF(b));
    for (auto g : gammas) gammasP.push_back(OF(g));
    std::vector<std::vector<uint64_t>> qvals(cfg.num_challenges, std::vector<uint64_t>(N));
    std::vector<uint64_t> xs(N);
    { ORC_STAGE("quotient eval");
    { uint64_t x = GL_GENERATOR; for (size_t i = 0; i < N; i++) { xs[i] = x; x = gl_mul(x, wN); } }
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) {
        size_t inext = (i + next_step) % N;
        ConsumerT<OF> yc;
        for (auto a : alphas) { yc.alphas.push_back(OF(a)); yc.acc.push_back(OF(0)); }
        yc.z_last = OF(gl_sub(xs[i], last));
        yc.lagrange_first = OF(lag_first[i]);
        yc.lagrange_last = OF(lag_last[i]);
        RowOF lv{trace_commit.lde_row(i)}, nv{trace_commit.lde_row(inext)};
        RowOF alv{na ? aux_commit.lde_row(i) : nullptr}, anv{na ? aux_commit.lde_row(inext) : nullptr};
        eval_vanishing_poly<OF>(table, sh, betasP, gammasP, lv, nv, alv, anv, yc, prm, cd);
        for (unsigned j = 0; j < cfg.num_challenges; j++) qvals[j][i] = gl_mul(yc.acc[j].v, zh_inv[i & 1]);
    }
    }
    const size_t nq = 2 * cfg.num_challenges;
    std::vector<uint64_t> qcoeffs(nq * n);
    StageClock* sc_q = new StageClock("quotient commit");
    for (unsigned j = 0; j < cfg.num_challenges; j++) {
        coset_ifft_inplace(qvals[j].data(), logN, GL_GENERATOR);
        memcpy(&qcoeffs[(2 * j) * n], qvals[j].data(), n * 8);
        memcpy(&qcoeffs[(2 * j + 1) * n], qvals[j].data() + n, n * 8);
    }
    if (dbg) dbg->quotient_chunk_coeffs = qcoeffs;
    PolyBatch quot_commit;
    quot_commit.from_coeffs(std::move(qcoeffs), nq, n, rate_bits, cfg.cap_height);
    proof.quotient_cap = cap_words(quot_commit.tree);
    observe_words(ch, proof.quotient_cap);
    delete sc_q;

    // 4. zeta and the openings
    StageClock* sc_o = new StageClock("openings");
    Ext zeta = ch.ext_challenge();
    if (ext_pow(zeta, n) == Ext(1, 0)) throw std::runtime_error("Opening point is in the subgroup.");
    Ext zeta_next = ext_scalar(zeta, wn);
    std::vector<Ext> local(ncols), next(ncols), auxz(na), auxn(na), quot(nq);
    std::vector<uint64_t> aux_first(na);
    #pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < ncols; c++) { local[c] = trace_commit.eval_ext(c, zeta); next[c] = trace_commit.eval_ext(c, zeta_next); }
    #pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < na; c++) {
        auxz[c] = aux_commit.eval_ext(c, zeta); auxn[c] = aux_commit.eval_ext(c, zeta_next); aux_first[c] = aux_commit.eval_base(c, 1);
    }
    for (size_t c = 0; c < nq; c++) quot[c] = quot_commit.eval_ext(c, zeta);
    for (auto& e : local) push_ext(proof.local_values, e);
    for (auto& e : next) push_ext(proof.next_values, e);
    for (auto& e : auxz) push_ext(proof.aux_polys, e);
    for (auto& e : auxn) push_ext(proof.aux_polys_next, e);
    for (auto& e : quot) push_ext(proof.quotient_polys, e);
    const size_t zs_begin = sh.num_lookup_cols + sh.num_ctl_helpers;
    for (size_t c = zs_begin; c < na; c++) proof.ctl_zs_first.push_back(aux_first[c]);
    // observe_openings(to_fri_openings): batch zeta = local, aux, quotient; batch zeta_next = next, aux_next; batch 1 = ctl_zs_first
    observe_words(ch, proof.local_values); observe_words(ch, proof.aux_polys); observe_words(ch, proof.quotient_polys);
    observe_words(ch, proof.next_values); observe_words(ch, proof.aux_polys_next);
    for (uint64_t v : proof.ctl_zs_first) { ch.observe(v); ch.observe(0); }

    delete sc_o;
    // 5./6. FRI batch reduction in coefficient space (PolynomialBatch::prove_openings)
    StageClock* sc_f = new StageClock("fri batch reduce");
    Ext alpha = ch.ext_challenge();
    struct PolyRef { const PolyBatch* b; size_t c; };
    std::vector<PolyRef> all_trace, all_aux, all_quot, zs;
    for (size_t c = 0; c < ncols; c++) all_trace.push_back({&trace_commit, c});
    for (size_t c = 0; c < na; c++) all_aux.push_back({&aux_commit, c});
    for (size_t c = 0; c < nq; c++) all_quot.push_back({&quot_commit, c});
    for (size_t c = zs_begin; c < na; c++) zs.push_back({&aux_commit, c});
    std::vector<std::pair<Ext, std::vector<PolyRef>>> batches;
    { std::vector<PolyRef> b0 = all_trace; b0.insert(b0.end(), all_aux.begin(), all_aux.end()); b0.insert(b0.end(), all_quot.begin(), all_quot.end());
      batches.push_back({zeta, b0}); }
    { std::vector<PolyRef> b1 = all_trace; b1.insert(b1.end(), all_aux.begin(), all_aux.end()); batches.push_back({zeta_next, b1}); }
    if (!zs.empty()) batches.push_back({Ext(1, 0), zs});
    std::vector<Ext> final_poly(n, Ext(0, 0));
    for (auto& bt : batches) {
        const Ext z = bt.first;
        const auto& polys = bt.second;
        // composition = sum_j alpha^j f_j  (rows split over the threads; the alpha powers are tabulated first)
        std::vector<Ext> comp(n, Ext(0, 0)), apows(polys.size());
        { Ext apow(1, 0); for (size_t j = 0; j < polys.size(); j++) { apows[j] = apow; apow = apow * alpha; } }
        const size_t CH = 1024;
        #pragma omp parallel for schedule(static)
        for (size_t i0 = 0; i0 < n; i0 += CH) {
            const size_t i1 = std::min(n, i0 + CH);
            for (size_t j = 0; j < polys.size(); j++) {
                const uint64_t* cf = polys[j].b->col(polys[j].c);
                const Ext ap = apows[j];
                for (size_t i = i0; i < i1; i++) comp[i] = comp[i] + ext_scalar(ap, cf[i]);
            }
        }
        // divide_by_linear(z): synthetic division, remainder dropped, padded back to n
        std::vector<Ext> quo(n, Ext(0, 0));
        Ext acc(0, 0);
        for (size_t i = n; i-- > 1;) { acc = acc * z + comp[i]; quo[i - 1] = acc; }
        // final = final * alpha^(#polys) + quo
        Ext shiftf = ext_pow(alpha, polys.size());
        for (size_t i = 0; i < n; i++) final_poly[i] = final_poly[i] * shiftf + quo[i];
    }
    std::vector<Ext> coeffs(N, Ext(0, 0));
    std::copy(final_poly.begin(), final_poly.end(), coeffs.begin());
    std::vector<Ext> values = coeffs;
    ext_coset_fft(values, logN, GL_GENERATOR);
    delete sc_f;

    // 7. FRI commit phase (fri_committed_trees)
    StageClock* sc_c = new StageClock("fri commit phase");
    std::vector<MerkleTree> trees;
    uint64_t shift = GL_GENERATOR;
    bool first_layer = true;
    for (unsigned ab : arities) {
        size_t arity = (size_t)1 << ab;
        size_t M = values.size();
        unsigned logM = 0; while (((size_t)1 << logM) < M) logM++;
        std::vector<Ext> br(M);
        for (size_t i = 0; i < M; i++) br[i] = values[bitrev(i, logM)];
        if (first_layer && dbg) dbg->fri_final_values = br;
        first_layer = false;
        LeafVec rows(2 * M);
        for (size_t i = 0; i < M; i++) { rows[2 * i] = br[i].a; rows[2 * i + 1] = br[i].b; }
        MerkleTree t;
        t.build(std::move(rows), M / arity, 2 * arity, cfg.cap_height);
        proof.commit_phase_caps.push_back(cap_words(t));
        observe_words(ch, proof.commit_phase_caps.back());
        trees.push_back(std::move(t));
        Ext beta = ch.ext_challenge();
        std::vector<Ext> nc(coeffs.size() / arity);
        for (size_t i = 0; i < nc.size(); i++) {
            Ext acc(0, 0);
            for (size_t t2 = arity; t2-- > 0;) acc = acc * beta + coeffs[i * arity + t2];
            nc[i] = acc;
        }
        coeffs = std::move(nc);
        shift = gl_pow(shift, arity);
        values = coeffs;
        ext_coset_fft(values, logM - ab, shift);
    }
    if (first_layer && dbg) {   // no reduction layers: still expose the bit-reversed values
        size_t M = values.size(); unsigned logM = 0; while (((size_t)1 << logM) < M) logM++;
        dbg->fri_final_values.resize(M);
        for (size_t i = 0; i < M; i++) dbg->fri_final_values[i] = values[bitrev(i, logM)];
    }
    coeffs.resize(coeffs.size() >> rate_bits);
    for (auto& e : coeffs) push_ext(proof.final_poly, e);
    observe_words(ch, proof.final_poly);

    delete sc_c;
    ORC_STAGE("pow + queries");
    // 8. proof of work: smallest witness (the reference takes any: rayon find_any), or the forced one
    {
        Challenger base = ch;
        uint64_t w = 0;
        if (forced_pow) w = *forced_pow;
        else {
            // candidates in blocks, all host threads per block (the reference grinds with rayon), smallest valid one of the first
            // block that holds any
            const uint64_t BLOCK = 1 << 12;
            for (uint64_t w0 = 0;; w0 += BLOCK) {
                uint64_t best = ~0ull;
                #pragma omp parallel for reduction(min : best) schedule(static)
                for (uint64_t k = 0; k < BLOCK; k++) {
                    Challenger t = base;
                    t.observe(w0 + k);
                    uint64_t r = t.challenge();
                    if ((cfg.pow_bits == 0 || (r >> (64 - cfg.pow_bits)) == 0) && w0 + k < best) best = w0 + k;
                }
                if (best != ~0ull) { w = best; break; }
            }
        }
        ch.observe(w);
        uint64_t r = ch.challenge();
        if (cfg.pow_bits && (r >> (64 - cfg.pow_bits)) != 0) throw std::runtime_error("forced proof-of-work witness is invalid");
        proof.pow_witness = w;
    }

    // 9. query rounds
    std::vector<uint64_t> rands(cfg.num_queries);
    for (auto& r : rands) r = ch.challenge();
    std::vector<const MerkleTree*> initial = {&trace_commit.tree};
    if (na) initial.push_back(&aux_commit.tree);
    initial.push_back(&quot_commit.tree);
    for (uint64_t r : rands) {
        size_t x = (size_t)(r % N);
        zkstark::FriQueryRound qr;
        for (const MerkleTree* t : initial) {
            zkstark::FriInitialProof ip;
            ip.leaf.assign(t->leaf(x), t->leaf(x) + t->leaf_len);
            ip.path = path_words(t->prove(x));
            qr.initial.push_back(std::move(ip));
        }
        for (size_t li = 0; li < trees.size(); li++) {
            size_t xi = x >> arities[li];
            zkstark::FriQueryStep st;
            st.evals.assign(trees[li].leaf(xi), trees[li].leaf(xi) + trees[li].leaf_len);
            st.path = path_words(trees[li].prove(xi));
            qr.steps.push_back(std::move(st));
            x = xi;
        }
        proof.queries.push_back(std::move(qr));
    }
    return proof;
}

// ---- verifier ----------------------------------------------------------------------------------------------------
static inline Ext get_ext(const Words& w, size_t i) { return Ext(w[2 * i], w[2 * i + 1]); }
static inline bool verify_merkle(const uint64_t* leaf, size_t leaf_len, size_t index, const Words& path, const Words& cap) {
    Hash h = hash_or_noop(leaf, leaf_len);
    for (size_t s = 0; s + 4 <= path.size(); s += 4) {
        Hash sib; memcpy(sib.e, &path[s], 32);
        h = (index & 1) ? two_to_one(sib, h) : two_to_one(h, sib);
        index >>= 1;
    }
    if (4 * (index + 1) > cap.size()) return false;
    return !memcmp(h.e, &cap[4 * index], 32);
}

// starky verify_stark_proof_with_challenges + plonky2 verify_fri_proof, with the challenges re-derived from the shared
// transcript exactly as get_challenges.rs:297-311 does (the trace cap was observed by the caller).
static inline bool verify_table(uint32_t table, const Config& cfg, const StarkProofData& p, const std::vector<uint64_t>& betas,
                                const std::vector<uint64_t>& gammas, Challenger& ch, const TableParams& prm,
                                const std::vector<CrossTableLookup>& ctls, std::string& err) {
    const unsigned cd = zkstark::CONSTRAINT_DEGREE;
    const unsigned k = (unsigned)p.degree_bits, rate_bits = cfg.rate_bits, logN = k + rate_bits;
    const size_t n = (size_t)1 << k, N = n << rate_bits;
    const size_t ncols = zkstark::table_num_columns(table);
    AuxShape sh = aux_shape(table, ctls, cfg.num_challenges, cd);
    const size_t na = sh.num_aux(), nq = 2 * cfg.num_challenges;
    auto fail = [&](const std::string& m) { err = m; return false; };
    if (p.local_values.size() != 2 * ncols || p.next_values.size() != 2 * ncols) return fail("wrong number of trace openings");
    if (p.aux_polys.size() != 2 * na || p.aux_polys_next.size() != 2 * na) return fail("wrong number of aux openings");
    if (p.quotient_polys.size() != 2 * nq) return fail("wrong number of quotient openings");
    if (p.ctl_zs_first.size() != sh.ctl_entries.size()) return fail("wrong number of ctl_zs_first");
    std::vector<unsigned> arities = zkstark::fri_reduction_arity_bits(cfg, k);
    if (p.commit_phase_caps.size() != arities.size()) return fail("wrong number of FRI layers");

    ch.compact();
    if (memcmp(ch.state, p.init_challenger_state, 96)) return fail("init_challenger_state mismatch");
    if (na) observe_words(ch, p.aux_cap);
    std::vector<uint64_t> alphas(cfg.num_challenges);
    for (auto& a : alphas) a = ch.challenge();
    observe_words(ch, p.quotient_cap);
    Ext zeta = ch.ext_challenge();
    observe_words(ch, p.local_values); observe_words(ch, p.aux_polys); observe_words(ch, p.quotient_polys);
    observe_words(ch, p.next_values); observe_words(ch, p.aux_polys_next);
    for (uint64_t v : p.ctl_zs_first) { ch.observe(v); ch.observe(0); }
    Ext fri_alpha = ch.ext_challenge();
    std::vector<Ext> fri_betas;
    for (auto& cap : p.commit_phase_caps) { observe_words(ch, cap); fri_betas.push_back(ch.ext_challenge()); }
    observe_words(ch, p.final_poly);
    ch.observe(p.pow_witness);
    uint64_t pow_response = ch.challenge();
    if (cfg.pow_bits && (pow_response >> (64 - cfg.pow_bits)) != 0) return fail("invalid proof of work");
    std::vector<size_t> qidx(cfg.num_queries);
    for (auto& q : qidx) q = (size_t)(ch.challenge() % N);

    // constraint identity at zeta
    std::vector<OE> lv(ncols), nv(ncols), alv(na), anv(na);
    for (size_t c = 0; c < ncols; c++) { lv[c] = OE(get_ext(p.local_values, c)); nv[c] = OE(get_ext(p.next_values, c)); }
    for (size_t c = 0; c < na; c++) { alv[c] = OE(get_ext(p.aux_polys, c)); anv[c] = OE(get_ext(p.aux_polys_next, c)); }
    uint64_t wn = gl_root_of_unity(k);
    Ext zeta_pow = ext_pow(zeta, n);
    Ext z_x = zeta_pow - Ext(1, 0);
    uint64_t nf = gl_from_u64(n);
    Ext l0 = z_x * ext_inv(ext_scalar(zeta - Ext(1, 0), nf));
    Ext llast = z_x * ext_inv(ext_scalar(ext_scalar(zeta, wn) - Ext(1, 0), nf));
    ConsumerT<OE> yc;
    for (auto a : alphas) { yc.alphas.push_back(OE(Ext(a, 0))); yc.acc.push_back(OE::zero()); }
    yc.z_last = OE(zeta - Ext(gl_inv(wn), 0));
    yc.lagrange_first = OE(l0);
    yc.lagrange_last = OE(llast);
    std::vector<OE> betasP, gammasP;
    for (auto b : betas) betasP.push_back(OE(Ext(b, 0)));
    for (auto g : gammas) gammasP.push_back(OE(Ext(g, 0)));
    RowOE rlv{lv.data()}, rnv{nv.data()}, ralv{alv.data()}, ranv{anv.data()};
    eval_vanishing_poly<OE>(table, sh, betasP, gammasP, rlv, rnv, ralv, ranv, yc, prm, cd);
    for (unsigned j = 0; j < cfg.num_challenges; j++) {
        Ext t = get_ext(p.quotient_polys, 2 * j) + zeta_pow * get_ext(p.quotient_polys, 2 * j + 1);
        if (yc.acc[j].e != z_x * t) return fail("Mismatch between evaluation and opening of quotient polynomial (challenge " + std::to_string(j) + ")");
    }
    // ctl_zs_first must match nothing here (cross-table sums are checked by verify_ctl_sums)

    // FRI
    Ext zeta_next = ext_scalar(zeta, wn);
    const size_t zs_begin = sh.num_lookup_cols + sh.num_ctl_helpers;
    struct Ref { int oracle; size_t idx; };
    int o_trace = 0, o_aux = na ? 1 : -1, o_quot = na ? 2 : 1;
    std::vector<Ref> b0, b1, b2;
    for (size_t c = 0; c < ncols; c++) { b0.push_back({o_trace, c}); b1.push_back({o_trace, c}); }
    for (size_t c = 0; c < na; c++) { b0.push_back({o_aux, c}); b1.push_back({o_aux, c}); }
    for (size_t c = 0; c < nq; c++) b0.push_back({o_quot, c});
    for (size_t c = zs_begin; c < na; c++) b2.push_back({o_aux, c});
    std::vector<Ext> open0, open1, open2;
    for (size_t c = 0; c < ncols; c++) open0.push_back(get_ext(p.local_values, c));
    for (size_t c = 0; c < na; c++) open0.push_back(get_ext(p.aux_polys, c));
    for (size_t c = 0; c < nq; c++) open0.push_back(get_ext(p.quotient_polys, c));
    for (size_t c = 0; c < ncols; c++) open1.push_back(get_ext(p.next_values, c));
    for (size_t c = 0; c < na; c++) open1.push_back(get_ext(p.aux_polys_next, c));
    for (uint64_t v : p.ctl_zs_first) open2.push_back(Ext(v, 0));
    auto reduce = [&](const std::vector<Ext>& v) { Ext acc(0, 0); for (size_t i = v.size(); i-- > 0;) acc = acc * fri_alpha + v[i]; return acc; };
    struct BatchV { Ext point; std::vector<Ref> polys; Ext reduced; };
    std::vector<BatchV> batches = {{zeta, b0, reduce(open0)}, {zeta_next, b1, reduce(open1)}};
    if (!b2.empty()) batches.push_back({Ext(1, 0), b2, reduce(open2)});
    size_t final_len = p.final_poly.size() / 2;
    {
        unsigned tot = 0; for (unsigned a : arities) tot += a;
        if (final_len != ((size_t)1 << (k - tot))) return fail("final polynomial has the wrong length");
    }
    size_t n_oracles = na ? 3 : 2;
    std::vector<const Words*> caps = {&p.trace_cap};
    if (na) caps.push_back(&p.aux_cap);
    caps.push_back(&p.quotient_cap);
    std::vector<size_t> widths = {ncols};
    if (na) widths.push_back(na);
    widths.push_back(nq);
    if (p.queries.size() != cfg.num_queries) return fail("wrong number of query rounds");
    uint64_t wN = gl_root_of_unity(logN);
    for (size_t qi = 0; qi < p.queries.size(); qi++) {
        const zkstark::FriQueryRound& qr = p.queries[qi];
        size_t x_index = qidx[qi];
        if (qr.initial.size() != n_oracles || qr.steps.size() != arities.size()) return fail("malformed query round");
        for (size_t o = 0; o < n_oracles; o++) {
            if (qr.initial[o].leaf.size() != widths[o]) return fail("wrong initial leaf width");
            if (qr.initial[o].path.size() != 4 * (logN - cfg.cap_height)) return fail("wrong initial path length");
            if (!verify_merkle(qr.initial[o].leaf.data(), widths[o], x_index, qr.initial[o].path, *caps[o])) return fail("invalid initial Merkle proof");
        }
        uint64_t subgroup_x = gl_mul(GL_GENERATOR, gl_pow(wN, bitrev(x_index, logN)));
        // fri_combine_initial
        Ext sum(0, 0);
        for (auto& b : batches) {
            Ext red(0, 0);
            for (size_t i = b.polys.size(); i-- > 0;) red = red * fri_alpha + Ext(qr.initial[b.polys[i].oracle].leaf[b.polys[i].idx], 0);
            Ext numerator = red - b.reduced;
            Ext denominator = Ext(subgroup_x, 0) - b.point;
            sum = sum * ext_pow(fri_alpha, b.polys.size()) + numerator * ext_inv(denominator);
        }
        Ext old_eval = sum;
        unsigned cur_log = logN;
        for (size_t li = 0; li < arities.size(); li++) {
            unsigned ab = arities[li];
            size_t arity = (size_t)1 << ab;
            const zkstark::FriQueryStep& st = qr.steps[li];
            if (st.evals.size() != 2 * arity) return fail("wrong number of evals in a query step");
            size_t coset_index = x_index >> ab, within = x_index & (arity - 1);
            if (get_ext(st.evals, within) != old_eval) return fail("FRI consistency check failed at layer " + std::to_string(li));
            // compute_evaluation: interpolate {(x g^i, P(x g^i))} over the coset and evaluate at beta
            uint64_t g = gl_root_of_unity(ab);
            size_t rev_within = bitrev(within, ab);
            uint64_t coset_start = gl_mul(subgroup_x, gl_pow(g, arity - rev_within));
            std::vector<Ext> pts(arity), vals(arity);
            { uint64_t y = 1; for (size_t i = 0; i < arity; i++) { pts[i] = Ext(gl_mul(coset_start, y), 0); vals[i] = get_ext(st.evals, bitrev(i, ab)); y = gl_mul(y, g); } }
            Ext res(0, 0);
            for (size_t i = 0; i < arity; i++) {
                Ext num(1, 0), den(1, 0);
                for (size_t j = 0; j < arity; j++) if (j != i) { num = num * (fri_betas[li] - pts[j]); den = den * (pts[i] - pts[j]); }
                res = res + vals[i] * num * ext_inv(den);
            }
            old_eval = res;
            if (st.path.size() != 4 * (cur_log - ab - cfg.cap_height)) return fail("wrong FRI path length");
            if (!verify_merkle(st.evals.data(), 2 * arity, coset_index, st.path, p.commit_phase_caps[li])) return fail("invalid FRI layer Merkle proof");
            subgroup_x = gl_pow(subgroup_x, arity);
            x_index = coset_index;
            cur_log -= ab;
        }
        Ext fe(0, 0);
        for (size_t i = final_len; i-- > 0;) fe = fe * Ext(subgroup_x, 0) + get_ext(p.final_poly, i);
        if (fe != old_eval) return fail("Final polynomial evaluation is invalid.");
    }
    return true;
}

}  // namespace orc
