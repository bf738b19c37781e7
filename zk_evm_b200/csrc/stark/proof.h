// StarkProof containers and their canonical serialisation (host only, plain data — no arithmetic lives here).
//
// Mirrors starky 1.0.0 proof.rs `StarkProof` / `StarkOpeningSet`, plonky2 1.0.0 fri/proof.rs `FriProof` /
// `FriQueryRound` / `FriInitialTreeProof` / `FriQueryStep`, and `StarkProofWithMetadata`
// (/root/reference/evm_arithmetization/src/proof.rs:29-54, prover.rs:335-338).
//
// Serialised form (the bytes zkgpu_proof_serialize returns; little-endian u64 words, every field element canonical):
//   magic 'ZKSTARK1', table_id, degree_bits,
//   init_challenger_state[12],
//   vec(trace_cap) vec(aux_cap) vec(quotient_cap)                    -- vec(x) = length word then the words
//   vec(local) vec(next) vec(aux) vec(aux_next) vec(ctl_zs_first) vec(quotient)   -- ext values as (c0, c1) pairs
//   num_layers, then per layer vec(cap)
//   num_queries, then per query:  num_oracles, per oracle vec(leaf) vec(path);  num_steps, per step vec(evals) vec(path)
//   vec(final_poly)  pow_witness
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <stdexcept>

namespace zkstark {

typedef std::vector<uint64_t> Words;

struct FriInitialProof { Words leaf; Words path; };          // path = siblings bottom-up, 4 words each
struct FriQueryStep { Words evals; Words path; };            // evals = arity ext values, (c0, c1) pairs
struct FriQueryRound { std::vector<FriInitialProof> initial; std::vector<FriQueryStep> steps; };

struct StarkProofData {
    uint64_t table_id = 0, degree_bits = 0;
    uint64_t init_challenger_state[12] = {0};
    Words trace_cap, aux_cap, quotient_cap;
    Words local_values, next_values, aux_polys, aux_polys_next, ctl_zs_first, quotient_polys;
    std::vector<Words> commit_phase_caps;
    std::vector<FriQueryRound> queries;
    Words final_poly;
    uint64_t pow_witness = 0;
};

static const uint64_t PROOF_MAGIC = 0x314b524154534b5aULL;   // "ZKSTARK1"

inline void put_vec(Words& o, const Words& v) { o.push_back(v.size()); o.insert(o.end(), v.begin(), v.end()); }

inline Words serialize_proof(const StarkProofData& p) {
    Words o;
    o.push_back(PROOF_MAGIC); o.push_back(p.table_id); o.push_back(p.degree_bits);
    for (int i = 0; i < 12; i++) o.push_back(p.init_challenger_state[i]);
    put_vec(o, p.trace_cap); put_vec(o, p.aux_cap); put_vec(o, p.quotient_cap);
    put_vec(o, p.local_values); put_vec(o, p.next_values); put_vec(o, p.aux_polys); put_vec(o, p.aux_polys_next);
    put_vec(o, p.ctl_zs_first); put_vec(o, p.quotient_polys);
    o.push_back(p.commit_phase_caps.size());
    for (auto& c : p.commit_phase_caps) put_vec(o, c);
    o.push_back(p.queries.size());
    for (auto& q : p.queries) {
        o.push_back(q.initial.size());
        for (auto& ip : q.initial) { put_vec(o, ip.leaf); put_vec(o, ip.path); }
        o.push_back(q.steps.size());
        for (auto& s : q.steps) { put_vec(o, s.evals); put_vec(o, s.path); }
    }
    put_vec(o, p.final_poly);
    o.push_back(p.pow_witness);
    return o;
}

struct WordReader {
    const uint64_t* p; size_t n, pos = 0;
    WordReader(const uint64_t* p_, size_t n_) : p(p_), n(n_) {}
    uint64_t get() { if (pos >= n) throw std::runtime_error("proof truncated"); return p[pos++]; }
    Words vec() {
        uint64_t l = get();
        if (l > n - pos) throw std::runtime_error("proof truncated");
        Words v(p + pos, p + pos + l); pos += l; return v;
    }
};

inline StarkProofData deserialize_proof(const uint64_t* w, size_t n, size_t* consumed = nullptr) {
    WordReader r(w, n);
    StarkProofData p;
    if (r.get() != PROOF_MAGIC) throw std::runtime_error("bad proof magic");
    p.table_id = r.get(); p.degree_bits = r.get();
    for (int i = 0; i < 12; i++) p.init_challenger_state[i] = r.get();
    p.trace_cap = r.vec(); p.aux_cap = r.vec(); p.quotient_cap = r.vec();
    p.local_values = r.vec(); p.next_values = r.vec(); p.aux_polys = r.vec(); p.aux_polys_next = r.vec();
    p.ctl_zs_first = r.vec(); p.quotient_polys = r.vec();
    size_t nl = r.get();
    for (size_t i = 0; i < nl; i++) p.commit_phase_caps.push_back(r.vec());
    size_t nq = r.get();
    for (size_t i = 0; i < nq; i++) {
        FriQueryRound q;
        size_t no = r.get();
        for (size_t k = 0; k < no; k++) { FriInitialProof ip; ip.leaf = r.vec(); ip.path = r.vec(); q.initial.push_back(ip); }
        size_t ns = r.get();
        for (size_t k = 0; k < ns; k++) { FriQueryStep s; s.evals = r.vec(); s.path = r.vec(); q.steps.push_back(s); }
        p.queries.push_back(q);
    }
    p.final_poly = r.vec();
    p.pow_witness = r.get();
    if (consumed) *consumed = r.pos;
    return p;
}

// StarkConfig / FriConfig (zkgpu_stark_config with host-friendly names)
struct Config {
    unsigned security_bits = 100, num_challenges = 2, rate_bits = 1, cap_height = 4, pow_bits = 16, arity_bits = 4,
             final_poly_bits = 5, num_queries = 84;
};
// FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits).reduction_arity_bits(degree_bits, rate_bits, cap_height)
inline std::vector<unsigned> fri_reduction_arity_bits(const Config& c, unsigned degree_bits) {
    std::vector<unsigned> r;
    unsigned d = degree_bits;
    while (d > c.final_poly_bits && d + c.rate_bits >= c.cap_height + c.arity_bits) {
        r.push_back(c.arity_bits);
        if (d < c.arity_bits) throw std::runtime_error("degree_bits < arity_bits");
        d -= c.arity_bits;
    }
    return r;
}

}  // namespace zkstark
