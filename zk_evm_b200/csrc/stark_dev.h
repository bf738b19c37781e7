// Internal interfaces of the STARK proving path on the device (aux.cu, quotient.cu, fri.cu, prove.cu).
#pragma once
#include "internal.h"
#include "stark/all_stark.h"
#include "stark/checks.h"
#include "stark/proof.h"

namespace zk {

// lookup / CTL descriptors of one table, uploaded once per context
struct TableDev {
    uint32_t table = 0, num_challenges = 0;
    zkstark::Flat flat;
    DevBuf buf;
    zkstark::FlatView view;   // device pointers into buf
};
const TableDev& get_table_dev(Ctx& c, uint32_t table, unsigned num_challenges);

// CtlData of one table on the device: CTL helper columns then CTL Z columns, each n values (natural row order)
struct Ctl {
    Ctx* ctx = nullptr;
    uint32_t table = 0;
    size_t n = 0;
    unsigned num_challenges = 0;
    uint64_t betas[4] = {0}, gammas[4] = {0};
    DevBuf cols;   // (num_ctl_helpers + num_ctl_zs) x n
};

struct Proof {
    zkstark::StarkProofData data;
    std::vector<uint64_t> words;
    // debug only
    std::unique_ptr<zkgpu_batch> aux, quot;
    std::vector<uint64_t> fri_values;   // N x (re, im), bit-reversed order
};

// prove_single_table split at the point where the shared transcript is first needed (prove.cu)
struct TableJob {
    Ctx* ctx = nullptr;
    uint32_t table = 0;
    zkstark::TableParams prm = {0, 0, 0, 0};
    zkstark::Config cfg;
    const Batch* trace = nullptr;
    const Ctl* ctl = nullptr;
    std::unique_ptr<zkgpu_batch> aux;   // committed auxiliary polynomials
    uint64_t* cons = nullptr;            // alpha-independent constraint values, ncons x N (the context's per-table buffer; only when it precomputes them)
    uint32_t ncons = 0;
    bool begun = false;
};
void prove_table_begin(Ctx& c, uint32_t table, const zkstark::TableParams& prm, const zkstark::Config& cfg, const Batch& trace,
                       const Ctl& ctl, volatile const int* abort_flag, TableJob& job);
void prove_table_finish(Ctx& c, TableJob& job, uint64_t challenger_state[12], const uint64_t* forced_pow,
                        volatile const int* abort_flag, Proof& out);
void prove_table(Ctx& c, uint32_t table, const zkstark::TableParams& prm, const zkstark::Config& cfg, const Batch& trace, const Ctl& ctl,
                 uint64_t challenger_state[12], const uint64_t* forced_pow, volatile const int* abort_flag, Proof& out);
zkstark::Config config_from(const zkgpu_stark_config* k);
zkstark::TableParams params_from(const zkgpu_kernel_labels* labels);
// get_ctl_data for one table (prove.cu)
void make_ctl_data(Ctx& c, uint32_t table, const Batch& trace, const uint64_t* beta_gamma, uint32_t num_challenges, Ctl& out);

// ---- aux.cu ---------------------------------------------------------------------------------------------------
// CTL helper + Z columns of a table (starky cross_table_lookup_data / partial_sums) into out[(helpers+zs) x n]
void ctl_columns(Ctx& c, const TableDev& t, const uint64_t* values, size_t n, const uint64_t* betas, const uint64_t* gammas,
                 uint64_t* out);
// logUp helper + Z columns (starky lookup_helper_columns) into out[num_lookup_cols x n]
void lookup_columns(Ctx& c, const TableDev& t, const uint64_t* values, size_t n, const uint64_t* betas, uint64_t* out);

// ---- quotient.cu ----------------------------------------------------------------------------------------------
struct QuotientArgs {
    uint32_t table;
    const uint64_t* trace_lde;   // ncols x N, bit-reversed rows
    const uint64_t* aux_lde;     // naux x N
    unsigned log_n;              // trace length 2^log_n, N = 2n
    unsigned num_challenges;
    uint64_t alphas[4], betas[4], gammas[4];
    zkstark::TableParams prm;
    uint64_t* out;               // num_challenges x N, NATURAL order: out[j*N + i] = quotient_j(g w_N^i)
};
void quotient_values(Ctx& c, const TableDev& t, const QuotientArgs& a);
// the same evaluation split at the alphas: every constraint value to its column of cons (expect x N) ...
uint32_t total_constraints(const TableDev& t);
void constraints_record(Ctx& c, const TableDev& t, const QuotientArgs& a, uint64_t* cons, uint32_t expect);
// ... and the quotient values as their Horner combination in alpha (out: num_challenges x N, natural order)
void quotient_from_constraints(Ctx& c, const uint64_t* cons, uint32_t T, unsigned log_n, unsigned num_challenges, const uint64_t* alphas, uint64_t* out);

// ---- fri.cu ---------------------------------------------------------------------------------------------------
// evaluate every coefficient column at zeta and zeta_next (extension) and at 1 (base): out[col] = {z.a,z.b,zn.a,zn.b,one}
void eval_columns(Ctx& c, const uint64_t* coeffs, size_t ncols, size_t n, Fp2 zeta, Fp2 zeta_next, std::vector<uint64_t>& out5);
struct CombineArgs {
    const uint64_t* lde[3]; size_t ncols[3];   // trace, aux, quotient oracles (column-major, stride N, bit-reversed rows)
    size_t zs_begin;                            // first CTL-Z column inside aux
    unsigned log_N;
    Fp2 alpha, zeta, zeta_next;
    Fp2 v0, v1, v2;                             // reduced openings sum_j alpha^j opening_j of the three batches
    bool has_b2;
    uint64_t* out_re; uint64_t* out_im;         // N values each, bit-reversed order
};
void fri_combine(Ctx& c, const CombineArgs& a);
// column-major leaf matrix of a commit-phase layer: out[k*(M/arity) + r] = (k&1 ? im : re)[arity*r + (k>>1)]
void fri_leaves(Ctx& c, const uint64_t* re, const uint64_t* im, size_t M, unsigned arity_bits, uint64_t* out);
// coefficient folding: out[i] = sum_t beta^t in[arity*i + t]
void fri_fold(Ctx& c, const uint64_t* re, const uint64_t* im, size_t M, unsigned arity_bits, Fp2 beta, uint64_t* out_re,
              uint64_t* out_im);
// smallest w such that permute(state with w at position pos)[7] has >= bits leading zeros
uint64_t pow_grind(Ctx& c, const uint64_t state[12], unsigned pos, unsigned bits);
// out[i] = src[offsets[i]] (device gather of scattered words)
void gather_words(Ctx& c, const uint64_t* src, const std::vector<uint64_t>& offsets, uint64_t* out_host);
void gather_addrs(Ctx& c, const std::vector<const uint64_t*>& addrs, uint64_t* out_host);

// ---- api.cu ---------------------------------------------------------------------------------------------------
void commit_from_device_values(Ctx& c, Batch& b, bool keep_values);
void commit_from_device_coeffs(Ctx& c, Batch& b);
void init_batch(Ctx& c, Batch& b, size_t ncols, size_t n, uint32_t rate_bits, uint32_t cap_height);

}  // namespace zk

struct zkgpu_ctl { zk::Ctl c; };
struct zkgpu_proof { zk::Proof p; };
struct zkgpu_table_job { zk::TableJob j; };
