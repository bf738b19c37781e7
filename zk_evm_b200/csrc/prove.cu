// Per-table STARK proof on the device: the host-side orchestration of starky 1.0.0 `prove_with_commitment` as called by
// `prove_single_table` (/root/reference/evm_arithmetization/src/prover.rs:301-341), plus `get_ctl_data` for one table
// (prover.rs:137-143) and the C ABI around them.  All polynomial data stays in HBM; only caps, openings, challenges and
// query answers cross to the host.
#include "stark_dev.h"
#include "challenger.h"
#include "ntt.h"
#include "merkle.h"
#include <string.h>

namespace zk {

using zkstark::StarkProofData; using zkstark::Words; using zkstark::Config;

Config config_from(const zkgpu_stark_config* k) {
    Config c;
    c.security_bits = k->security_bits; c.num_challenges = k->num_challenges; c.rate_bits = k->rate_bits;
    c.cap_height = k->cap_height; c.pow_bits = k->proof_of_work_bits; c.arity_bits = k->fri_arity_bits;
    c.final_poly_bits = k->fri_final_poly_bits; c.num_queries = k->num_query_rounds;
    return c;
}

zkstark::TableParams params_from(const zkgpu_kernel_labels* labels) {
    zkstark::TableParams prm = {0, 0, 0, 0};
    if (labels) {
        prm.halt_final = labels->halt_final; prm.init = labels->init;
        prm.syscall_jumptable = labels->syscall_jumptable; prm.exception_jumptable = labels->exception_jumptable;
    }
    return prm;
}

static void check_abort(volatile const int* flag) {
    if (flag && *flag) throw ZkError(ZKGPU_ERR_ABORTED, "abort signal observed (prover.rs:346-354)");
}

static Words cap_words(const Batch& b) { return b.cap_host; }
static void push_ext(Words& w, uint64_t a, uint64_t b) { w.push_back(a); w.push_back(b); }

// The query answers of one table are collected in two steps: every oracle appends the device addresses of the words it opens
// (leaf row x of a column-major batch, the siblings of its Merkle path), ONE gather brings them all to the host, and the words are
// dealt out in the same order.
struct QueryPlan {
    std::vector<const uint64_t*> addrs;
    struct Part { size_t ncols, nlev, first; };
    std::vector<Part> parts;
    // leaf row xs[q] and its path, for every q, of the oracle (lde: ncols x N column-major, digests with level offsets level_off)
    void add(const uint64_t* lde, size_t ncols, size_t N, const uint64_t* digests, const std::vector<size_t>& level_off, const std::vector<size_t>& xs) {
        const size_t nlev = level_off.size() - 1;   // siblings per path
        parts.push_back({ncols, nlev, addrs.size()});
        for (size_t x : xs) {
            for (size_t col = 0; col < ncols; col++) addrs.push_back(lde + col * N + x);
            size_t idx = x;
            for (size_t l = 0; l < nlev; l++) { for (int w = 0; w < 4; w++) addrs.push_back(digests + level_off[l] + 4 * (idx ^ 1) + w); idx >>= 1; }
        }
    }
    // words of query q of part `part`: leaf then path
    void take(const std::vector<uint64_t>& vals, size_t part, size_t q, Words& leaf, Words& path) const {
        const Part& p = parts[part];
        const uint64_t* v = vals.data() + p.first + q * (p.ncols + 4 * p.nlev);
        leaf.assign(v, v + p.ncols);
        path.assign(v + p.ncols, v + p.ncols + 4 * p.nlev);
    }
};

struct FriLayer {
    DevBuf leaves;    // column-major (2*arity) x rows
    DevBuf digests;
    std::vector<size_t> off, cnt;
    size_t rows = 0, width = 0;
};

// Challenger-independent half of prove_single_table: the auxiliary polynomials (lookup columns ++ CTL helpers ++ CTL Zs)
// and their commitment depend only on the trace and the CTL challenges, so a table-sharded run computes them for all
// tables in parallel before the serial transcript relay reaches the table (DESIGN.md "multi-GPU").
void prove_table_begin(Ctx& c, uint32_t table, const zkstark::TableParams& prm, const Config& cfg, const Batch& trace, const Ctl& ctl,
                       volatile const int* abort_flag, TableJob& job) {
    ZK_REQUIRE(zkstark::table_supported(table), "table id not supported");
    ZK_REQUIRE(cfg.rate_bits == 1, "only rate_bits = 1 (quotient degree factor 2) is implemented");
    ZK_REQUIRE(cfg.num_challenges >= 1 && cfg.num_challenges <= 2, "num_challenges must be 1 or 2");
    ZK_REQUIRE(trace.ncols == zkstark::table_num_columns(table), "trace width does not match the table");
    ZK_REQUIRE(trace.rate_bits == cfg.rate_bits && trace.cap_height == cfg.cap_height, "trace commitment made with another config");
    ZK_REQUIRE(trace.values.get() != nullptr, "trace batch must be committed with keep_values (needed for the lookup columns)");
    ZK_REQUIRE(ctl.table == table && ctl.n == trace.n && ctl.num_challenges == cfg.num_challenges, "ctl data does not match the table");
    check_abort(abort_flag);
    job.ctx = &c; job.table = table; job.prm = prm; job.cfg = cfg; job.trace = &trace; job.ctl = &ctl;
    const size_t n = trace.n;
    const TableDev& td = get_table_dev(c, table, cfg.num_challenges);
    const zkstark::Flat& fl = td.flat;
    const size_t na = fl.num_aux();
    job.aux.reset(new zkgpu_batch());
    Batch& aux = job.aux->b;
    StageLog lg(c);
    if (na) {
        init_batch(c, aux, na, n, cfg.rate_bits, cfg.cap_height);
        aux.values = DevBuf(&c, na * n * 8);
        if (fl.num_lookup_cols) lookup_columns(c, td, trace.values.get(), n, ctl.betas, aux.values.get());
        size_t nctl = fl.num_ctl_helpers + fl.num_ctl_zs;
        if (nctl) ZK_CUDA(cudaMemcpyAsync(aux.values.get() + (size_t)fl.num_lookup_cols * n, ctl.cols.get(), nctl * n * 8,
                                          cudaMemcpyDeviceToDevice, c.stream));
        lg.mark("aux columns");
        commit_from_device_values(c, aux, c.debug);
        lg.mark("aux commit");
    }
    if (c.precompute) {
        // everything of the quotient evaluation that does not depend on the alphas, while the transcript is still with another table
        job.ncons = total_constraints(td);
        const size_t need = (size_t)job.ncons * trace.N * 8;
        DevBuf& buf = c.cons_cache[table];
        if (buf.bytes < need) { buf.release(); buf = DevBuf(&c, need); }
        job.cons = buf.get();
        QuotientArgs qa;
        qa.table = table; qa.trace_lde = trace.lde.get(); qa.aux_lde = na ? aux.lde.get() : nullptr; qa.log_n = trace.log_n;
        qa.num_challenges = cfg.num_challenges;
        for (int i = 0; i < 4; i++) { qa.alphas[i] = 0; qa.betas[i] = ctl.betas[i]; qa.gammas[i] = ctl.gammas[i]; }
        qa.prm = prm; qa.out = nullptr;
        constraints_record(c, td, qa, job.cons, job.ncons);
        lg.mark("constraint values");
    }
    job.begun = true;
}

void prove_table_finish(Ctx& c, TableJob& job, uint64_t challenger_state[12], const uint64_t* forced_pow,
                        volatile const int* abort_flag, Proof& out) {
    ZK_REQUIRE(job.begun && job.ctx == &c, "table job was not begun on this context");
    const uint32_t table = job.table;
    const zkstark::TableParams& prm = job.prm;
    const Config& cfg = job.cfg;
    const Batch& trace = *job.trace;
    const Ctl& ctl = *job.ctl;
    check_abort(abort_flag);
    const unsigned k = trace.log_n, logN = k + 1;
    const size_t n = trace.n, N = trace.N, ncols = trace.ncols;
    const TableDev& td = get_table_dev(c, table, cfg.num_challenges);
    const zkstark::Flat& fl = td.flat;
    const size_t na = fl.num_aux(), nq = 2 * cfg.num_challenges;
    std::vector<unsigned> arities = zkstark::fri_reduction_arity_bits(cfg, k);
    {
        unsigned tot = 0; for (unsigned a : arities) tot += a;
        ZK_REQUIRE(tot <= k + cfg.rate_bits - cfg.cap_height, "FRI total arity is too large");
    }
    StageLog lg(c);
    StarkProofData& p = out.data;
    p = StarkProofData();
    p.table_id = table; p.degree_bits = k;
    Challenger ch;
    ch.set_state(challenger_state);
    ch.compact();
    memcpy(p.init_challenger_state, ch.state, 96);
    p.trace_cap = cap_words(trace);

    // 1. auxiliary polynomials: committed by prove_table_begin
    std::unique_ptr<zkgpu_batch> auxh = std::move(job.aux);
    job.begun = false;
    Batch& aux = auxh->b;
    if (na) {
        p.aux_cap = cap_words(aux);
        ch.observe_vec(p.aux_cap);
    }
    check_abort(abort_flag);

    // 2. alphas
    uint64_t alphas[4] = {0, 0, 0, 0};
    for (unsigned i = 0; i < cfg.num_challenges; i++) alphas[i] = ch.challenge();

    // 3. quotient: values on the coset (natural order) -> coset iNTT -> chunks of n coefficients -> commitment
    std::unique_ptr<zkgpu_batch> quoth(new zkgpu_batch());
    Batch& quot = quoth->b;
    {
        DevBuf qv(&c, cfg.num_challenges * N * 8), scratch(&c, cfg.num_challenges * N * 8);
        QuotientArgs qa;
        qa.table = table; qa.trace_lde = trace.lde.get(); qa.aux_lde = na ? aux.lde.get() : nullptr; qa.log_n = k;
        qa.num_challenges = cfg.num_challenges;
        for (int i = 0; i < 4; i++) { qa.alphas[i] = alphas[i]; qa.betas[i] = ctl.betas[i]; qa.gammas[i] = ctl.gammas[i]; }
        qa.prm = prm; qa.out = qv.get();
        if (job.ncons) {
            quotient_from_constraints(c, job.cons, job.ncons, k, cfg.num_challenges, alphas, qv.get());
            job.cons = nullptr;
            job.ncons = 0;
        } else {
            quotient_values(c, td, qa);
        }
        lg.mark("quotient eval");
        init_batch(c, quot, nq, n, cfg.rate_bits, cfg.cap_height);
        quot.coeffs = DevBuf(&c, nq * n * 8);
        // coset_ifft of each challenge's 2n values; the 2n coefficients of challenge j are chunks 2j, 2j+1
        intt_natural(c, qv.get(), scratch.get(), quot.coeffs.get(), cfg.num_challenges, logN, GL_GENERATOR);
        lg.mark("quotient intt");
        commit_from_device_coeffs(c, quot);
        lg.mark("quotient commit");
    }
    p.quotient_cap = cap_words(quot);
    ch.observe_vec(p.quotient_cap);
    check_abort(abort_flag);

    // 4. zeta, openings
    Fp2 zeta = ch.ext_challenge();
    if (fp2_pow(zeta, n) == Fp2(1, 0)) throw ZkError(ZKGPU_ERR_PROOF, "Opening point is in the subgroup.");
    const uint64_t wn = gl_root_of_unity(k);
    Fp2 zeta_next = scalar_mul(zeta, wn);
    std::vector<uint64_t> ev_t, ev_a, ev_q;
    eval_columns(c, trace.coeffs.get(), ncols, n, zeta, zeta_next, ev_t);
    if (na) eval_columns(c, aux.coeffs.get(), na, n, zeta, zeta_next, ev_a);
    eval_columns(c, quot.coeffs.get(), nq, n, zeta, zeta_next, ev_q);
    lg.mark("openings eval");
    for (size_t i = 0; i < ncols; i++) { push_ext(p.local_values, ev_t[5 * i], ev_t[5 * i + 1]); push_ext(p.next_values, ev_t[5 * i + 2], ev_t[5 * i + 3]); }
    for (size_t i = 0; i < na; i++) { push_ext(p.aux_polys, ev_a[5 * i], ev_a[5 * i + 1]); push_ext(p.aux_polys_next, ev_a[5 * i + 2], ev_a[5 * i + 3]); }
    for (size_t i = 0; i < nq; i++) push_ext(p.quotient_polys, ev_q[5 * i], ev_q[5 * i + 1]);
    const size_t zs_begin = fl.num_lookup_cols + fl.num_ctl_helpers;
    for (size_t i = zs_begin; i < na; i++) p.ctl_zs_first.push_back(ev_a[5 * i + 4]);
    ch.observe_vec(p.local_values); ch.observe_vec(p.aux_polys); ch.observe_vec(p.quotient_polys);
    ch.observe_vec(p.next_values); ch.observe_vec(p.aux_polys_next);
    for (uint64_t v : p.ctl_zs_first) { ch.observe(v); ch.observe(0); }

    lg.mark("observe openings (host)");
    // 5./6. FRI: alpha, reduced openings, values of the combined quotient on the coset
    Fp2 alpha = ch.ext_challenge();
    auto reduce = [&](std::initializer_list<const Words*> parts) {
        std::vector<Fp2> v;
        for (const Words* w : parts) for (size_t i = 0; i + 1 < w->size(); i += 2) v.push_back(Fp2((*w)[i], (*w)[i + 1]));
        Fp2 acc(0, 0);
        for (size_t i = v.size(); i-- > 0;) acc = acc * alpha + v[i];
        return acc;
    };
    Words zs_ext;
    for (uint64_t v : p.ctl_zs_first) { zs_ext.push_back(v); zs_ext.push_back(0); }
    CombineArgs ca;
    ca.lde[0] = trace.lde.get(); ca.ncols[0] = ncols;
    ca.lde[1] = na ? aux.lde.get() : nullptr; ca.ncols[1] = na;
    ca.lde[2] = quot.lde.get(); ca.ncols[2] = nq;
    ca.zs_begin = zs_begin; ca.log_N = logN; ca.alpha = alpha; ca.zeta = zeta; ca.zeta_next = zeta_next;
    ca.v0 = reduce({&p.local_values, &p.aux_polys, &p.quotient_polys});
    ca.v1 = reduce({&p.next_values, &p.aux_polys_next});
    ca.v2 = reduce({&zs_ext});
    ca.has_b2 = !p.ctl_zs_first.empty();
    DevBuf vals(&c, 2 * N * 8);          // re | im, bit-reversed order
    ca.out_re = vals.get(); ca.out_im = vals.get() + N;
    fri_combine(c, ca);
    if (c.debug) {
        std::vector<uint64_t> h(2 * N);
        c.d2h(h.data(), vals.get(), 2 * N * 8);
        out.fri_values.resize(2 * N);
        for (size_t i = 0; i < N; i++) { out.fri_values[2 * i] = h[i]; out.fri_values[2 * i + 1] = h[N + i]; }
    }
    // coefficients of the FRI polynomial: natural-order values -> coset iNTT (2 columns)
    DevBuf coeffs(&c, 2 * N * 8);
    {
        DevBuf nat(&c, 2 * N * 8), scratch(&c, 2 * N * 8);
        bitrev_permute(c, vals.get(), N, nat.get(), N, 2, logN, 1, nullptr);
        intt_natural(c, nat.get(), scratch.get(), coeffs.get(), 2, logN, GL_GENERATOR);
    }
    check_abort(abort_flag);
    lg.mark("fri combine + intt");

    // 7. commit phase
    std::vector<FriLayer> layers;
    uint64_t shift = GL_GENERATOR;
    size_t M = N;            // current number of values; coefficient vector has the same length
    unsigned logM = logN;
    for (unsigned ab : arities) {
        FriLayer L;
        L.rows = M >> ab; L.width = (size_t)2 << ab;
        L.leaves = DevBuf(&c, L.rows * L.width * 8);
        fri_leaves(c, vals.get(), vals.get() + M, M, ab, L.leaves.get());
        merkle_build(c, L.leaves.get(), L.rows, L.width, L.rows, cfg.cap_height, L.digests, L.off, L.cnt);
        Words cap(4 * L.cnt.back());
        c.d2h(cap.data(), L.digests.get() + L.off.back(), cap.size() * 8);
        p.commit_phase_caps.push_back(cap);
        ch.observe_vec(cap);
        Fp2 beta = ch.ext_challenge();
        size_t M2 = M >> ab;
        DevBuf nc(&c, 2 * M2 * 8);
        fri_fold(c, coeffs.get(), coeffs.get() + M, M, ab, beta, nc.get(), nc.get() + M2);
        coeffs = std::move(nc);
        shift = gl_pow(shift, (uint64_t)1 << ab);
        M = M2; logM -= ab;
        // values of the folded polynomial on shift * <w_M>, bit-reversed order (what the next layer's leaves hold)
        DevBuf nv(&c, 2 * M * 8);
        const uint64_t* pre = get_power_table(c, shift, 1, M);
        ntt_dif(c, coeffs.get(), M, 0, nv.get(), M, 2, logM, false, pre, nullptr, 0);
        vals = std::move(nv);
        layers.push_back(std::move(L));
    }
    {
        size_t keep = M >> cfg.rate_bits;
        std::vector<uint64_t> h(2 * M);
        c.d2h(h.data(), coeffs.get(), 2 * M * 8);
        for (size_t i = 0; i < keep; i++) push_ext(p.final_poly, h[i], h[M + i]);
    }
    ch.observe_vec(p.final_poly);
    lg.mark("fri commit phase");

    // 8. proof of work
    {
        uint64_t w;
        if (forced_pow) w = *forced_pow;
        else {
            uint64_t st[12];
            memcpy(st, ch.state, 96);
            for (size_t i = 0; i < ch.in.size(); i++) st[i] = ch.in[i];
            w = pow_grind(c, st, (unsigned)ch.in.size(), cfg.pow_bits);
        }
        ch.observe(w);
        uint64_t r = ch.challenge();
        if (cfg.pow_bits && (r >> (64 - cfg.pow_bits)) != 0) throw ZkError(ZKGPU_ERR_PROOF, "proof-of-work witness is invalid");
        p.pow_witness = w;
    }
    check_abort(abort_flag);

    lg.mark("pow");
    // 9. query rounds
    std::vector<size_t> xs(cfg.num_queries);
    for (auto& x : xs) x = (size_t)(ch.challenge() % N);
    p.queries.assign(cfg.num_queries, zkstark::FriQueryRound());
    {
        std::vector<const Batch*> oracles = {&trace};
        if (na) oracles.push_back(&aux);
        oracles.push_back(&quot);
        QueryPlan qp;
        for (const Batch* b : oracles) qp.add(b->lde.get(), b->ncols, b->N, b->digests.get(), b->level_off, xs);
        std::vector<size_t> cur = xs;
        for (size_t li = 0; li < layers.size(); li++) {
            for (auto& x : cur) x >>= arities[li];
            qp.add(layers[li].leaves.get(), layers[li].width, layers[li].rows, layers[li].digests.get(), layers[li].off, cur);
        }
        std::vector<uint64_t> vals(qp.addrs.size());
        gather_addrs(c, qp.addrs, vals.data());
        for (size_t q = 0; q < xs.size(); q++) {
            for (size_t o = 0; o < oracles.size(); o++) {
                zkstark::FriInitialProof ip;
                qp.take(vals, o, q, ip.leaf, ip.path);
                p.queries[q].initial.push_back(std::move(ip));
            }
            for (size_t li = 0; li < layers.size(); li++) {
                zkstark::FriQueryStep st;
                qp.take(vals, oracles.size() + li, q, st.evals, st.path);
                p.queries[q].steps.push_back(std::move(st));
            }
        }
    }
    lg.mark("queries");
    ch.compact();
    memcpy(challenger_state, ch.state, 96);
    out.words = zkstark::serialize_proof(p);
    if (c.debug) { if (na) out.aux = std::move(auxh); out.quot = std::move(quoth); }
    c.sync();
}

void make_ctl_data(Ctx& c, uint32_t table_id, const Batch& b, const uint64_t* beta_gamma, uint32_t num_challenges, Ctl& k) {
    ZK_REQUIRE(num_challenges >= 1 && num_challenges <= 2, "num_challenges must be 1 or 2");
    ZK_REQUIRE(b.values.get() != nullptr, "trace batch must be committed with keep_values");
    ZK_REQUIRE(zkstark::table_supported(table_id) && b.ncols == zkstark::table_num_columns(table_id), "trace width does not match the table");
    const TableDev& td = get_table_dev(c, table_id, num_challenges);
    k.ctx = &c; k.table = table_id; k.n = b.n; k.num_challenges = num_challenges;
    for (uint32_t i = 0; i < num_challenges; i++) { k.betas[i] = gl_canon(beta_gamma[2 * i]); k.gammas[i] = gl_canon(beta_gamma[2 * i + 1]); }
    size_t nctl = td.flat.num_ctl_helpers + td.flat.num_ctl_zs;
    k.cols = DevBuf(&c, nctl * b.n * 8);
    if (nctl) ctl_columns(c, td, b.values.get(), b.n, k.betas, k.gammas, k.cols.get());
}

void prove_table(Ctx& c, uint32_t table, const zkstark::TableParams& prm, const Config& cfg, const Batch& trace, const Ctl& ctl,
                 uint64_t challenger_state[12], const uint64_t* forced_pow, volatile const int* abort_flag, Proof& out) {
    TableJob job;
    prove_table_begin(c, table, prm, cfg, trace, ctl, abort_flag, job);
    prove_table_finish(c, job, challenger_state, forced_pow, abort_flag, out);
}

}  // namespace zk

using namespace zk;

extern "C" {

int zkgpu_ctx_set_precompute_constraints(zkgpu_ctx* h, int on) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "ctx is null");
    h->c.precompute = on != 0;
    ZK_API_END
}

int zkgpu_ctx_set_debug(zkgpu_ctx* h, int on) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "ctx is null");
    h->c.debug = on != 0;
    ZK_API_END
}

int zkgpu_table_info(uint32_t table_id, uint32_t num_challenges, uint32_t* num_columns, uint32_t* num_lookup_columns,
                     uint32_t* num_ctl_helper_columns, uint32_t* num_ctl_zs) {
    ZK_API_BEGIN
    ZK_REQUIRE(zkstark::table_supported(table_id), "table id not supported");
    auto items = zkstark::table_ctl_items(table_id, zkstark::all_cross_table_lookups(), num_challenges);
    zkstark::Flat f = zkstark::build_table_flat(zkstark::table_lookups(table_id), items, num_challenges, zkstark::CONSTRAINT_DEGREE);
    if (num_columns) *num_columns = zkstark::table_num_columns(table_id);
    if (num_lookup_columns) *num_lookup_columns = f.num_lookup_cols;
    if (num_ctl_helper_columns) *num_ctl_helper_columns = f.num_ctl_helpers;
    if (num_ctl_zs) *num_ctl_zs = f.num_ctl_zs;
    ZK_API_END
}

int zkgpu_ctl_data(zkgpu_ctx* h, uint32_t table_id, const zkgpu_batch* trace, const uint64_t* beta_gamma, uint32_t num_challenges,
                   zkgpu_ctl** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && trace && beta_gamma && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_ctl> hc(new zkgpu_ctl());
    make_ctl_data(c, table_id, trace->b, beta_gamma, num_challenges, hc->c);
    c.sync();
    *out = hc.release();
    ZK_API_END
}

void zkgpu_ctl_free(zkgpu_ctl* k) {
    if (!k) return;
    if (k->c.ctx) cudaSetDevice(k->c.ctx->device);
    delete k;
}

int zkgpu_ctl_export(const zkgpu_ctl* k, uint64_t* out_cols) {
    ZK_API_BEGIN
    ZK_REQUIRE(k && out_cols, "null argument");
    Ctx& c = *k->c.ctx;
    ZK_CUDA(cudaSetDevice(c.device));
    const TableDev& td = get_table_dev(c, k->c.table, k->c.num_challenges);
    size_t nctl = td.flat.num_ctl_helpers + td.flat.num_ctl_zs;
    if (nctl) c.d2h(out_cols, k->c.cols.get(), nctl * k->c.n * 8);
    ZK_API_END
}

int zkgpu_prove_table(zkgpu_ctx* h, uint32_t table_id, const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config,
                      const zkgpu_batch* trace, const zkgpu_ctl* ctl, uint64_t challenger_state[12], const uint64_t* forced_pow_witness,
                      volatile const int* abort_flag, zkgpu_proof** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && config && trace && ctl && challenger_state && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    zkstark::TableParams prm = params_from(labels);
    std::unique_ptr<zkgpu_proof> hp(new zkgpu_proof());
    uint64_t st[12];
    memcpy(st, challenger_state, 96);
    prove_table(c, table_id, prm, config_from(config), trace->b, ctl->c, st, forced_pow_witness, abort_flag, hp->p);
    memcpy(challenger_state, st, 96);
    *out = hp.release();
    ZK_API_END
}

void zkgpu_proof_free(zkgpu_proof* p) { delete p; }

int zkgpu_proof_serialize(const zkgpu_proof* p, uint64_t* buf, size_t* len_words) {
    ZK_API_BEGIN
    ZK_REQUIRE(p && len_words, "null argument");
    size_t need = p->p.words.size();
    if (buf && *len_words >= need) memcpy(buf, p->p.words.data(), need * 8);
    *len_words = need;
    ZK_API_END
}

int zkgpu_proof_debug_batch(const zkgpu_proof* p, int which, const zkgpu_batch** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(p && out, "null argument");
    const zkgpu_batch* b = which == 0 ? p->p.aux.get() : p->p.quot.get();
    ZK_REQUIRE(b != nullptr, "no debug batch retained (enable zkgpu_ctx_set_debug before proving)");
    *out = b;
    ZK_API_END
}

int zkgpu_proof_debug_fri_values(const zkgpu_proof* p, uint64_t* out, size_t* len_words) {
    ZK_API_BEGIN
    ZK_REQUIRE(p && len_words, "null argument");
    size_t need = p->p.fri_values.size();
    if (out && *len_words >= need) memcpy(out, p->p.fri_values.data(), need * 8);
    *len_words = need;
    ZK_API_END
}

}  // extern "C"
