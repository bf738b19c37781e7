// C++17 host side over the C ABI (include/zkgpu.h): the reference's interface for the proving path, name for name.
//
// The reference is compiled code (Rust); its toolchain is not in this image, so this header is the compiled-language host mirror: the
// same call shapes as evm_arithmetization/src/prover.rs — prove_with_traces (:72-194), prove_single_table (:301-341),
// PolynomialBatch::from_values (:100-107), get_ctl_data (:137-143), Challenger (:118-129, 320), check_abort_signal (:346-354) — and the
// same types: StarkProof / StarkProofWithMetadata / AllProof (proof.rs:29-54, prover.rs:335-338), MemCap (prover.rs:261-271),
// StarkConfig::standard_fast_config (zero/src/prover_state/mod.rs:283).  Rust's Result<T> is an exception here (zkgpu::Error carries
// the zkgpu_status); Option<Arc<AtomicBool>> is a `const std::atomic<int>*`.  Header-only; link with -lzkgpu.  No arithmetic lives here:
// every computation is a call into libzkgpu.so (CUDA), and there is no CPU fallback — without a usable device the calls throw
// Error{ZKGPU_ERR_CUDA}.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <exception>
#include <functional>
#include <map>
#include <mutex>
#include <condition_variable>
#include <thread>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "zkgpu.h"
#include "../zk_evm_b200/csrc/stark/proof.h"   // StarkProofData: the typed StarkProof fields + the canonical word layout (plain data)

namespace zkgpu {

using F = uint64_t;                                   // GoldilocksField: canonical u64 (#[repr(transparent)] over u64)
constexpr size_t NUM_TABLES = ZKGPU_NUM_TABLES;       // all_stark.rs:74-86
enum class Table : uint32_t { Arithmetic = 0, BytePacking, Cpu, Keccak, KeccakSponge, Logic, Memory, MemBefore, MemAfter };
constexpr std::array<uint32_t, 5> OPTIONAL_TABLE_INDICES = {1, 3, 4, 5, 8};   // all_stark.rs:110-117

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != ZKGPU_OK) throw Error(rc, zkgpu_last_error());
}

struct StarkConfig : zkgpu_stark_config {
    static StarkConfig standard_fast_config() { return StarkConfig{{100, 2, 1, 4, 16, 4, 5, 84}}; }
    static StarkConfig test_config() { return StarkConfig{{1, 1, 1, 4, 1, 4, 5, 1}}; }          // TEST_STARK_CONFIG, testing_utils.rs:41-52
};
using KernelLabels = zkgpu_kernel_labels;
using AbortSignal = const std::atomic<int>*;           // Option<Arc<AtomicBool>>: nullptr = None
inline volatile const int* abort_ptr(AbortSignal a) {
    static_assert(sizeof(std::atomic<int>) == sizeof(int), "atomic<int> must be layout-compatible with int");
    return reinterpret_cast<volatile const int*>(a);
}

// PolynomialValues<F>: one column of evaluations
using PolynomialValues = std::vector<F>;

class Context {
public:
    explicit Context(int device = 0) { check(zkgpu_ctx_create(device, &h_)); }
    ~Context() { if (h_) zkgpu_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    void sync() { check(zkgpu_ctx_sync(h_)); }
    // stage spans, the counterpart of the reference's TimingTree: (stage name, milliseconds) in order since set_timing(true)
    void set_timing(bool on) { check(zkgpu_ctx_set_timing(h_, on)); }
    std::vector<std::pair<std::string, double>> timing_report() {
        size_t len = 0;
        check(zkgpu_ctx_timing_report(h_, nullptr, &len));
        std::string buf(len, '\0');
        check(zkgpu_ctx_timing_report(h_, buf.data(), &len));
        std::vector<std::pair<std::string, double>> spans;
        size_t pos = 0;
        for (;;) {
            const size_t tab = buf.find('\t', pos), nl = buf.find('\n', pos);
            if (tab == std::string::npos || nl == std::string::npos) break;
            spans.emplace_back(buf.substr(pos, tab - pos), std::stod(buf.substr(tab + 1, nl - tab - 1)));
            pos = nl + 1;
        }
        return spans;
    }
    // table jobs begun on this context also evaluate the alpha-independent constraint values (table-sharded segments: the evaluator then runs
    // while the transcript is with another table, TableJob::finish only combines the stored columns)
    void set_precompute_constraints(bool on) { check(zkgpu_ctx_set_precompute_constraints(h_, on)); }
    // the context's cudaStream_t: a host that issues its own device work (the NCCL all-gathers of a commitment split over several
    // devices) orders it with the library's kernels on this stream instead of synchronising
    void* stream() const { void* s = nullptr; check(zkgpu_ctx_stream(h_, &s)); return s; }
    zkgpu_ctx* handle() const { return h_; }
private:
    zkgpu_ctx* h_ = nullptr;
};

// a page-locked host buffer of field elements (zkgpu_host_alloc): uploads from it run at PCIe speed and asynchronously
class PinnedBuffer {
public:
    explicit PinnedBuffer(size_t n) : n_(n) { void* p = nullptr; check(zkgpu_host_alloc(n * sizeof(F), &p)); p_ = (F*)p; }
    ~PinnedBuffer() { zkgpu_host_free(p_); }
    PinnedBuffer(const PinnedBuffer&) = delete;
    PinnedBuffer& operator=(const PinnedBuffer&) = delete;
    F* data() { return p_; }
    const F* data() const { return p_; }
    size_t size() const { return n_; }
private:
    F* p_ = nullptr;
    size_t n_;
};

using Hash = std::array<F, 4>;
using MerkleCap = std::vector<Hash>;                   // plonky2 MerkleCap<F, PoseidonHash>
inline MerkleCap cap_from_words(const uint64_t* w, size_t n_words) {
    MerkleCap c(n_words / 4);
    for (size_t i = 0; i < c.size(); i++) for (int k = 0; k < 4; k++) c[i][k] = w[4 * i + k];
    return c;
}

// PolynomialBatch<F, C, D>: resident on the device; the host fields are filled on request (export_fields)
class PolynomialBatch {
public:
    struct HostFields {                                // PolynomialBatch{polynomials, merkle_tree{leaves, digests, cap}}
        std::vector<F> polynomials;                    // ncols x n coefficients, column-major
        std::vector<F> leaves;                         // (n << rate_bits) rows x ncols, row-major, row j = LDE row bitrev(j)
        std::vector<F> digests;                        // plonky2's recursive layout
        MerkleCap cap;
    };
    // PolynomialBatch::from_values(values, rate_bits, blinding = false, cap_height, timing, fft_root_table = None)
    static PolynomialBatch from_values(Context& ctx, const std::vector<PolynomialValues>& values, uint32_t rate_bits, uint32_t cap_height,
                                       bool keep_values = true) {
        if (values.empty()) throw Error(ZKGPU_ERR_INVALID, "no columns");
        std::vector<const uint64_t*> cols(values.size());
        for (size_t c = 0; c < values.size(); c++) {
            if (values[c].size() != values[0].size()) throw Error(ZKGPU_ERR_INVALID, "columns of different lengths");
            cols[c] = values[c].data();
        }
        zkgpu_batch* b = nullptr;
        check(zkgpu_commit_values(ctx.handle(), cols.data(), cols.size(), values[0].size(), rate_bits, cap_height, ZKGPU_MEM_HOST, keep_values, &b));
        return PolynomialBatch(b);
    }
    // one contiguous column-major block (host, or device with mem_kind = ZKGPU_MEM_DEVICE)
    static PolynomialBatch from_values_contig(Context& ctx, const F* base, size_t ncols, size_t n, uint32_t rate_bits, uint32_t cap_height,
                                              int mem_kind = ZKGPU_MEM_HOST, bool keep_values = true) {
        zkgpu_batch* b = nullptr;
        check(zkgpu_commit_values_contig(ctx.handle(), base, ncols, n, rate_bits, cap_height, mem_kind, keep_values, &b));
        return PolynomialBatch(b);
    }
    // PolynomialBatch::from_coeffs
    static PolynomialBatch from_coeffs(Context& ctx, const std::vector<std::vector<F>>& coeffs, uint32_t rate_bits, uint32_t cap_height) {
        if (coeffs.empty()) throw Error(ZKGPU_ERR_INVALID, "no columns");
        std::vector<const uint64_t*> cols(coeffs.size());
        for (size_t c = 0; c < coeffs.size(); c++) cols[c] = coeffs[c].data();
        zkgpu_batch* b = nullptr;
        check(zkgpu_commit_coeffs(ctx.handle(), cols.data(), cols.size(), coeffs[0].size(), rate_bits, cap_height, ZKGPU_MEM_HOST, &b));
        return PolynomialBatch(b);
    }
    size_t num_polys() const { return dims().ncols; }
    size_t degree() const { return dims().n; }
    MerkleCap cap() const {                            // merkle_tree.cap
        const Dims d = dims();
        std::vector<uint64_t> w((size_t)4 << d.cap_height);
        check(zkgpu_batch_cap(h_.get(), w.data()));
        return cap_from_words(w.data(), w.size());
    }
    HostFields export_fields() const {
        const Dims d = dims();
        const size_t N = d.n << d.rate_bits;
        HostFields f;
        f.polynomials.resize(d.ncols * d.n);
        f.leaves.resize(N * d.ncols);
        f.digests.resize(2 * (N - ((size_t)1 << d.cap_height)) * 4);
        check(zkgpu_batch_export(h_.get(), f.polynomials.data(), f.leaves.data(), f.digests.empty() ? nullptr : f.digests.data()));
        f.cap = cap();
        return f;
    }
    const zkgpu_batch* handle() const { return h_.get(); }

    // ---- from_values split over the k devices of a table-sharded segment (include/zkgpu.h "S1 split over several devices") ----
    // All pointers are DEVICE memory of the caller (the buffers an all-gather fills).  Step 1 on every device: its column slice
    // through ifft + LDE; exchange the slices; step 2 on every device: the leaf digests and Merkle levels of ITS block of leaves;
    // exchange the packed digests; step 3 on the owner: the batch over the gathered buffers (borrowed until it is destroyed).
    static void lde_slice(Context& ctx, const F* values, int mem_kind, size_t ncols, size_t n, uint32_t rate_bits, F* values_out, F* coeffs_out, F* lde_out) {
        check(zkgpu_lde_slice(ctx.handle(), values, mem_kind, ncols, n, rate_bits, values_out, coeffs_out, lde_out));
    }
    static size_t merkle_block_words(size_t nleaves, uint32_t cap_height, uint32_t nblocks) {
        size_t w = 0;
        check(zkgpu_merkle_block_words(nleaves, cap_height, nblocks, &w));
        return w;
    }
    static void merkle_block(Context& ctx, const F* lde, size_t stride, size_t ncols, size_t nleaves, uint32_t cap_height, uint32_t nblocks,
                             uint32_t block, F* packed_out) {
        check(zkgpu_merkle_block(ctx.handle(), lde, stride, ncols, nleaves, cap_height, nblocks, block, packed_out));
    }
    static PolynomialBatch assemble(Context& ctx, const F* values, const F* coeffs, const F* lde, const F* packed, uint32_t nblocks, size_t ncols,
                                    size_t n, uint32_t rate_bits, uint32_t cap_height) {
        zkgpu_batch* b = nullptr;
        check(zkgpu_batch_assemble(ctx.handle(), values, coeffs, lde, packed, nblocks, ncols, n, rate_bits, cap_height, &b));
        return PolynomialBatch(b);
    }
private:
    struct Dims { size_t ncols, n; uint32_t rate_bits, cap_height; };
    Dims dims() const { Dims d{}; check(zkgpu_batch_dims(h_.get(), &d.ncols, &d.n, &d.rate_bits, &d.cap_height)); return d; }
    struct Del { void operator()(zkgpu_batch* b) const { zkgpu_batch_free(b); } };
    explicit PolynomialBatch(zkgpu_batch* b) : h_(b) {}
    std::unique_ptr<zkgpu_batch, Del> h_;
};

// GrandProductChallengeSet: (beta, gamma) per challenge, flattened [beta_0, gamma_0, beta_1, gamma_1]
struct GrandProductChallengeSet { std::vector<F> beta_gamma; };

// plonky2 Challenger<F, PoseidonHash>
class Challenger {
public:
    Challenger() { check(zkgpu_challenger_new(&h_)); }
    ~Challenger() { if (h_) zkgpu_challenger_free(h_); }
    Challenger(const Challenger&) = delete;
    Challenger& operator=(const Challenger&) = delete;
    static std::unique_ptr<Challenger> from_state(const std::array<F, 12>& st) {
        auto c = std::make_unique<Challenger>();
        check(zkgpu_challenger_set_state(c->h_, st.data()));
        return c;
    }
    void observe_element(F x) { check(zkgpu_challenger_observe(h_, &x, 1)); }
    void observe_elements(const std::vector<F>& xs) { check(zkgpu_challenger_observe(h_, xs.data(), xs.size())); }
    void observe_cap(const MerkleCap& cap) { for (const Hash& hsh : cap) check(zkgpu_challenger_observe(h_, hsh.data(), 4)); }
    F get_challenge() { F x; check(zkgpu_challenger_get_challenges(h_, &x, 1)); return x; }
    std::vector<F> get_n_challenges(size_t n) { std::vector<F> v(n); check(zkgpu_challenger_get_challenges(h_, v.data(), n)); return v; }
    std::array<F, 12> compact() { std::array<F, 12> st; check(zkgpu_challenger_compact(h_, st.data())); return st; }
    void set_state(const std::array<F, 12>& st) { check(zkgpu_challenger_set_state(h_, st.data())); }
private:
    zkgpu_challenger* h_ = nullptr;
};

// CtlData<F> of one table (device-resident helper and Z columns)
class CtlData {
public:
    const zkgpu_ctl* handle() const { return h_.get(); }
private:
    friend CtlData get_ctl_data(Context&, Table, const PolynomialBatch&, const GrandProductChallengeSet&, uint32_t);
    struct Del { void operator()(zkgpu_ctl* c) const { zkgpu_ctl_free(c); } };
    explicit CtlData(zkgpu_ctl* c) : h_(c) {}
    std::unique_ptr<zkgpu_ctl, Del> h_;
};
// the per-table slice of starky get_ctl_data (prover.rs:137-143); the trace batch must have been committed with keep_values
inline CtlData get_ctl_data(Context& ctx, Table table, const PolynomialBatch& trace, const GrandProductChallengeSet& ch, uint32_t num_challenges) {
    if (ch.beta_gamma.size() != 2 * (size_t)num_challenges) throw Error(ZKGPU_ERR_INVALID, "beta_gamma must hold 2 * num_challenges elements");
    zkgpu_ctl* c = nullptr;
    check(zkgpu_ctl_data(ctx.handle(), (uint32_t)table, trace.handle(), ch.beta_gamma.data(), num_challenges, &c));
    return CtlData(c);
}

// StarkProofWithMetadata{proof: StarkProof{trace_cap, auxiliary_polys_cap, quotient_polys_cap, openings, opening_proof}, init_challenger_state}
// as typed fields (zkstark::StarkProofData, stark/proof.h) plus the canonical words they were decoded from
struct StarkProofWithMetadata {
    zkstark::StarkProofData proof;
    std::vector<uint64_t> words;
};
inline StarkProofWithMetadata take_proof(zkgpu_proof* p) {
    std::unique_ptr<zkgpu_proof, void (*)(zkgpu_proof*)> own(p, zkgpu_proof_free);
    size_t len = 0;
    check(zkgpu_proof_serialize(p, nullptr, &len));
    StarkProofWithMetadata out;
    out.words.resize(len);
    check(zkgpu_proof_serialize(p, out.words.data(), &len));
    out.proof = zkstark::deserialize_proof(out.words.data(), out.words.size());
    return out;
}

// prove_single_table (prover.rs:301-341): the shared challenger is compacted, handed over as its 12-word state and restored
inline StarkProofWithMetadata prove_single_table(Context& ctx, Table table, const StarkConfig& config, const PolynomialBatch& trace_commitment,
                                                 const CtlData& ctl_data, Challenger& challenger, const KernelLabels* labels = nullptr,
                                                 AbortSignal abort_signal = nullptr, const F* forced_pow_witness = nullptr) {
    std::array<F, 12> st = challenger.compact();
    zkgpu_proof* p = nullptr;
    check(zkgpu_prove_table(ctx.handle(), (uint32_t)table, labels, &config, trace_commitment.handle(), ctl_data.handle(), st.data(),
                            forced_pow_witness, abort_ptr(abort_signal), &p));
    challenger.set_state(st);
    return take_proof(p);
}

// prove_single_table split where the shared transcript is first needed (tables of one segment on several GPUs): begin = auxiliary
// polynomials + their commitment (needs only the CTL challenges), finish = everything that consumes the transcript
class TableJob {
public:
    static TableJob begin(Context& ctx, Table table, const StarkConfig& config, const PolynomialBatch& trace_commitment, const CtlData& ctl_data,
                          const KernelLabels* labels = nullptr, AbortSignal abort_signal = nullptr) {
        zkgpu_table_job* j = nullptr;
        check(zkgpu_table_job_begin(ctx.handle(), (uint32_t)table, labels, &config, trace_commitment.handle(), ctl_data.handle(),
                                    abort_ptr(abort_signal), &j));
        return TableJob(j);
    }
    StarkProofWithMetadata finish(std::array<F, 12>& challenger_state, AbortSignal abort_signal = nullptr, const F* forced_pow_witness = nullptr) {
        zkgpu_proof* p = nullptr;
        check(zkgpu_table_job_finish(h_.get(), challenger_state.data(), forced_pow_witness, abort_ptr(abort_signal), &p));
        return take_proof(p);
    }
private:
    struct Del { void operator()(zkgpu_table_job* j) const { zkgpu_table_job_free(j); } };
    explicit TableJob(zkgpu_table_job* j) : h_(j) {}
    std::unique_ptr<zkgpu_table_job, Del> h_;
};

// prover.rs:118-144: a fresh challenger observes the nine trace caps in Table order (a zero cap for an optional table that is not in
// use, :120-123) and the public values (observe_public_values), then draws the CTL challenges; returns them and leaves `challenger`
// where get_ctl_data leaves the reference's
inline GrandProductChallengeSet segment_challenges(Challenger& challenger, const std::array<std::optional<MerkleCap>, NUM_TABLES>& trace_caps,
                                                   const std::vector<F>& public_values, const StarkConfig& config) {
    const size_t cap_words = (size_t)4 << config.cap_height;
    std::vector<uint64_t> caps(NUM_TABLES * cap_words, 0);
    std::array<uint8_t, NUM_TABLES> in_use{};
    for (size_t t = 0; t < NUM_TABLES; t++) {
        in_use[t] = trace_caps[t].has_value();
        if (!in_use[t]) continue;
        if (trace_caps[t]->size() * 4 != cap_words) throw Error(ZKGPU_ERR_INVALID, "cap of the wrong height");
        for (size_t i = 0; i < trace_caps[t]->size(); i++) for (int k = 0; k < 4; k++) caps[t * cap_words + 4 * i + k] = (*trace_caps[t])[i][k];
    }
    GrandProductChallengeSet ch;
    ch.beta_gamma.resize(2 * config.num_challenges);
    std::array<F, 12> st;
    check(zkgpu_segment_challenges(caps.data(), in_use.data(), config.cap_height, public_values.data(), public_values.size(), config.num_challenges,
                                   ch.beta_gamma.data(), st.data()));
    challenger.set_state(st);
    return ch;
}

// prove_with_commitments (prover.rs:211-293): the tables in Table order, one shared challenger (:251-259); None for a table not in use
inline std::array<std::optional<StarkProofWithMetadata>, NUM_TABLES>
prove_with_commitments(Context& ctx, const StarkConfig& config, const std::array<const PolynomialBatch*, NUM_TABLES>& trace_commitments,
                       const std::array<const CtlData*, NUM_TABLES>& ctl_data_per_table, Challenger& challenger, const KernelLabels& labels,
                       AbortSignal abort_signal = nullptr) {
    std::array<std::optional<StarkProofWithMetadata>, NUM_TABLES> proofs;
    for (size_t t = 0; t < NUM_TABLES; t++) {
        if (!trace_commitments[t]) continue;
        if (!ctl_data_per_table[t]) throw Error(ZKGPU_ERR_INVALID, "a table in use needs its CtlData");
        proofs[t] = prove_single_table(ctx, (Table)t, config, *trace_commitments[t], *ctl_data_per_table[t], challenger, &labels, abort_signal);
    }
    return proofs;
}

// AllProof / MultiProof (proof.rs:29-54): stark_proofs[t] is None for an optional table that is not in use
struct AllProof {
    std::array<std::optional<StarkProofWithMetadata>, NUM_TABLES> stark_proofs;
    GrandProductChallengeSet ctl_challenges;
    std::array<MerkleCap, NUM_TABLES> trace_caps;
    const MerkleCap& mem_before_cap() const { return trace_caps[(size_t)Table::MemBefore]; }       // MemCap, prover.rs:261-271
    const MerkleCap& mem_after_cap() const { return trace_caps[(size_t)Table::MemAfter]; }
};

// one table's trace: Vec<PolynomialValues<F>> as one contiguous column-major block (column c at data + c*n); empty = table not in use
struct TableTrace {
    const F* cols = nullptr;
    size_t n = 0;
};

// prove_with_traces (prover.rs:72-194).  public_values: the elements observe_public_values feeds the challenger, in order
// (get_challenges.rs:202-227).  mem_kind: ZKGPU_MEM_HOST, ZKGPU_MEM_DEVICE or ZKGPU_MEM_AUTO (per-table pointers of either kind).
inline AllProof prove_with_traces(Context& ctx, const std::array<TableTrace, NUM_TABLES>& trace_poly_values, const std::vector<F>& public_values,
                                  const StarkConfig& config, const KernelLabels& labels, AbortSignal abort_signal = nullptr,
                                  int mem_kind = ZKGPU_MEM_HOST, const std::array<F, NUM_TABLES>* forced_pow_witnesses = nullptr) {
    std::array<zkgpu_table_trace, NUM_TABLES> tr;
    for (size_t t = 0; t < NUM_TABLES; t++) tr[t] = {trace_poly_values[t].cols, trace_poly_values[t].n};
    std::array<zkgpu_proof*, NUM_TABLES> proofs{};
    const size_t cap_words = (size_t)4 << config.cap_height;
    std::vector<uint64_t> bg(4, 0), caps(NUM_TABLES * cap_words, 0);
    check(zkgpu_prove_segment(ctx.handle(), tr.data(), mem_kind, public_values.data(), public_values.size(), &labels, &config,
                              forced_pow_witnesses ? forced_pow_witnesses->data() : nullptr, abort_ptr(abort_signal), proofs.data(), bg.data(),
                              caps.data()));
    AllProof out;
    out.ctl_challenges.beta_gamma.assign(bg.begin(), bg.begin() + 2 * config.num_challenges);
    for (size_t t = 0; t < NUM_TABLES; t++) {
        out.trace_caps[t] = cap_from_words(&caps[t * cap_words], cap_words);
        if (proofs[t]) out.stark_proofs[t] = take_proof(proofs[t]);
    }
    return out;
}

// ---- PublicValues -> the elements observe_public_values feeds the challenger (get_challenges.rs:11-227) --------------------------------
// proof.rs:68-91, 314-321, 357-364, 398-425, 471-488.  H256 / Address are big-endian byte strings, U256 is four little-endian u64 limbs
// (ethereum_types).  2217 elements with the default eth_mainnet feature (24 * 2 + 97 + 2056 + 16, proof.rs:652-655); the registers and
// the memory caps are not observed.
using H256 = std::array<uint8_t, 32>;
using Address = std::array<uint8_t, 20>;
struct U256 {
    std::array<uint64_t, 4> limbs{};                   // U256.0
    U256() = default;
    U256(uint64_t x) { limbs[0] = x; }
    static U256 from_big_endian(const uint8_t* b, size_t len) {
        U256 r;
        for (size_t i = 0; i < len; i++) r.limbs[(len - 1 - i) / 8] |= (uint64_t)b[i] << (8 * ((len - 1 - i) % 8));
        return r;
    }
};
struct TrieRoots { H256 state_root{}, transactions_root{}, receipts_root{}; };
struct BlockMetadata {
    Address block_beneficiary{};
    U256 block_timestamp, block_number, block_difficulty;
    H256 block_random{};
    U256 block_gaslimit, block_chain_id, block_base_fee, block_gas_used, block_blob_gas_used, block_excess_blob_gas;
    H256 parent_beacon_block_root{};
    std::array<U256, 8> block_bloom;
};
struct BlockHashes { std::vector<H256> prev_hashes = std::vector<H256>(256); H256 cur_hash{}; };
struct ExtraBlockData {
    H256 checkpoint_state_trie_root{};
    std::array<F, 4> checkpoint_consolidated_hash{};
    U256 txn_number_before, txn_number_after, gas_used_before, gas_used_after;
};
struct RegistersData { U256 program_counter, is_kernel, stack_len, stack_top, context, gas_used; };   // proof.rs:537-550 (written to memory, not observed)
struct PublicValues {
    TrieRoots trie_roots_before, trie_roots_after;
    BlockMetadata block_metadata;
    BlockHashes block_hashes;
    ExtraBlockData extra_block_data;
    std::optional<U256> burn_addr;                     // cdk_erigon only
    RegistersData registers_before, registers_after;
};
namespace detail {
inline void u256_limbs(std::vector<F>& o, const U256& x, size_t count = 8) {                     // util.rs:101-113
    for (size_t i = 0; i < count; i++) o.push_back((x.limbs[i / 2] >> (32 * (i % 2))) & 0xFFFFFFFFull);
}
inline void h256_limbs(std::vector<F>& o, const H256& h) { u256_limbs(o, U256::from_big_endian(h.data(), 32)); }   // util.rs:116-126, observe_root
inline void u256_to_u32(std::vector<F>& o, const U256& x) {                                      // util.rs:40-46
    if (x.limbs[1] | x.limbs[2] | x.limbs[3] | (x.limbs[0] >> 32)) throw Error(ZKGPU_ERR_INVALID, "IntegerTooLarge: public value does not fit 32 bits");
    o.push_back(x.limbs[0]);
}
inline void u256_to_u64(std::vector<F>& o, const U256& x) {                                      // util.rs:50-59
    if (x.limbs[1] | x.limbs[2] | x.limbs[3]) throw Error(ZKGPU_ERR_INVALID, "IntegerTooLarge: public value does not fit 64 bits");
    o.push_back(x.limbs[0] & 0xFFFFFFFFull);
    o.push_back(x.limbs[0] >> 32);
}
inline void trie_roots(std::vector<F>& o, const TrieRoots& r) { h256_limbs(o, r.state_root); h256_limbs(o, r.transactions_root); h256_limbs(o, r.receipts_root); }
}  // namespace detail
// observe_public_values (get_challenges.rs:202-227) as the list of observed elements, ready for prove_with_traces
inline std::vector<F> flatten_public_values(const PublicValues& pv, bool eth_mainnet = true, bool cdk_erigon = false) {
    using namespace detail;
    std::vector<F> o;
    o.reserve(2232);
    trie_roots(o, pv.trie_roots_before);
    trie_roots(o, pv.trie_roots_after);
    const BlockMetadata& m = pv.block_metadata;                                                  // observe_block_metadata, :45-81
    u256_limbs(o, U256::from_big_endian(m.block_beneficiary.data(), 20), 5);
    u256_to_u32(o, m.block_timestamp); u256_to_u32(o, m.block_number); u256_to_u32(o, m.block_difficulty);
    h256_limbs(o, m.block_random);
    u256_to_u32(o, m.block_gaslimit); u256_to_u32(o, m.block_chain_id);
    u256_to_u64(o, m.block_base_fee);
    u256_to_u32(o, m.block_gas_used);
    if (eth_mainnet) { u256_to_u64(o, m.block_blob_gas_used); u256_to_u64(o, m.block_excess_blob_gas); h256_limbs(o, m.parent_beacon_block_root); }
    for (const U256& w : m.block_bloom) u256_limbs(o, w);
    if (pv.block_hashes.prev_hashes.size() != 256) throw Error(ZKGPU_ERR_INVALID, "256 previous block hashes");   // observe_block_hashes, :172-184
    for (const H256& h : pv.block_hashes.prev_hashes) h256_limbs(o, h);
    h256_limbs(o, pv.block_hashes.cur_hash);
    const ExtraBlockData& e = pv.extra_block_data;                                               // observe_extra_block_data, :108-123
    h256_limbs(o, e.checkpoint_state_trie_root);
    for (F x : e.checkpoint_consolidated_hash) o.push_back(x);
    u256_to_u32(o, e.txn_number_before); u256_to_u32(o, e.txn_number_after); u256_to_u32(o, e.gas_used_before); u256_to_u32(o, e.gas_used_after);
    if (cdk_erigon) {
        if (!pv.burn_addr) throw Error(ZKGPU_ERR_INVALID, "There should be an address set in cdk_erigon.");
        u256_limbs(o, *pv.burn_addr);
    }
    return o;
}

// ---- the memory writes the kernel makes of the public values, which no Cpu row sends (verifier.rs:319-512, 536-737) --------------------
// The verifier adds them to the looking side of the Memory lookup; a host that builds Memory traces needs the same rows.
// GlobalMetadata indices inside Segment::GlobalMetadata (cpu/kernel/constants/global_metadata.rs:11-115, `unscale`), segments unscaled
// (memory/segments.rs:10-91).
namespace global_metadata {
enum : uint64_t { StateTrieRootDigestBefore = 6, TransactionTrieRootDigestBefore, ReceiptTrieRootDigestBefore, StateTrieRootDigestAfter,
                  TransactionTrieRootDigestAfter, ReceiptTrieRootDigestAfter, BlockBeneficiary, BlockTimestamp, BlockNumber, BlockDifficulty,
                  BlockRandom, BlockGasLimit, BlockChainId, BlockBaseFee, BlockBlobGasUsed, BlockExcessBlobGas, BlockGasUsed, BlockGasUsedBefore,
                  BlockGasUsedAfter, BlockCurrentHash, ParentBeaconBlockRoot, TxnNumberBefore = 42, TxnNumberAfter, KernelHash = 45, KernelLen,
                  BurnAddr = 53, COUNT = 54 };
}
namespace segment { enum : uint64_t { GlobalMetadata = 5, GlobalBlockBloom = 24, BlockHashes = 32, RegistersStates = 33 }; }
using MemoryLookupRow = std::array<F, 13>;             // is_read, context, segment, virt, eight 32-bit value limbs, timestamp
// get_memory_extra_looking_values (verifier.rs:547-737), in its order.  kernel_code_hash / kernel_code_len = KERNEL.code_hash /
// KERNEL.code.len(): the kernel is assembled by the host, like the four kernel labels.
inline std::vector<MemoryLookupRow> memory_extra_looking_values(const PublicValues& pv, const H256& kernel_code_hash, uint64_t kernel_code_len,
                                                                bool eth_mainnet = true, bool cdk_erigon = false) {
    namespace gm = global_metadata;
    std::vector<MemoryLookupRow> rows;
    auto row = [&](uint64_t seg, uint64_t index, const U256& val) {            // add_extra_looking_row, verifier.rs:720-735
        MemoryLookupRow r{};
        r[2] = seg; r[3] = index;
        for (size_t j = 0; j < 8; j++) r[4 + j] = (val.limbs[j / 2] >> (32 * (j % 2))) & 0xFFFFFFFFull;
        r[12] = 2;
        rows.push_back(r);
    };
    auto h2u = [](const H256& h) { return U256::from_big_endian(h.data(), 32); };
    auto meta = [&](uint64_t field, const U256& val) { row(segment::GlobalMetadata, field, val); };
    const BlockMetadata& m = pv.block_metadata;
    const ExtraBlockData& e = pv.extra_block_data;
    meta(gm::BlockBeneficiary, U256::from_big_endian(m.block_beneficiary.data(), 20));
    if (cdk_erigon) {
        if (!pv.burn_addr) throw Error(ZKGPU_ERR_INVALID, "There should be an address set in cdk_erigon.");
        meta(gm::BurnAddr, *pv.burn_addr);
    }
    meta(gm::BlockTimestamp, m.block_timestamp); meta(gm::BlockNumber, m.block_number); meta(gm::BlockRandom, h2u(m.block_random));
    meta(gm::BlockDifficulty, m.block_difficulty); meta(gm::BlockGasLimit, m.block_gaslimit); meta(gm::BlockChainId, m.block_chain_id);
    meta(gm::BlockBaseFee, m.block_base_fee); meta(gm::BlockCurrentHash, h2u(pv.block_hashes.cur_hash)); meta(gm::BlockGasUsed, m.block_gas_used);
    if (eth_mainnet) {
        meta(gm::BlockBlobGasUsed, m.block_blob_gas_used); meta(gm::BlockExcessBlobGas, m.block_excess_blob_gas);
        meta(gm::ParentBeaconBlockRoot, h2u(m.parent_beacon_block_root));
    }
    meta(gm::TxnNumberBefore, e.txn_number_before); meta(gm::TxnNumberAfter, e.txn_number_after);
    meta(gm::BlockGasUsedBefore, e.gas_used_before); meta(gm::BlockGasUsedAfter, e.gas_used_after);
    meta(gm::StateTrieRootDigestBefore, h2u(pv.trie_roots_before.state_root));
    meta(gm::TransactionTrieRootDigestBefore, h2u(pv.trie_roots_before.transactions_root));
    meta(gm::ReceiptTrieRootDigestBefore, h2u(pv.trie_roots_before.receipts_root));
    meta(gm::StateTrieRootDigestAfter, h2u(pv.trie_roots_after.state_root));
    meta(gm::TransactionTrieRootDigestAfter, h2u(pv.trie_roots_after.transactions_root));
    meta(gm::ReceiptTrieRootDigestAfter, h2u(pv.trie_roots_after.receipts_root));
    meta(gm::KernelHash, h2u(kernel_code_hash)); meta(gm::KernelLen, U256(kernel_code_len));
    for (size_t i = 0; i < 8; i++) row(segment::GlobalBlockBloom, i, m.block_bloom[i]);
    if (pv.block_hashes.prev_hashes.size() != 256) throw Error(ZKGPU_ERR_INVALID, "256 previous block hashes");
    for (size_t i = 0; i < 256; i++) row(segment::BlockHashes, i, h2u(pv.block_hashes.prev_hashes[i]));
    const RegistersData* regs[2] = {&pv.registers_before, &pv.registers_after};
    for (size_t k = 0; k < 2; k++) {
        const U256 v[6] = {regs[k]->program_counter, regs[k]->is_kernel, regs[k]->stack_len, regs[k]->stack_top, regs[k]->context, regs[k]->gas_used};
        for (size_t i = 0; i < 6; i++) row(segment::RegistersStates, 6 * k + i, v[i]);
    }
    return rows;
}
namespace detail {
constexpr F GL_P = 0xFFFFFFFF00000001ull;
inline F gl_mul(F a, F b) { return (F)((unsigned __int128)a * b % GL_P); }
inline F gl_add(F a, F b) { return (F)(((unsigned __int128)a + b) % GL_P); }
inline F gl_inv(F a) { F r = 1, e = GL_P - 2; while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; } return r; }
}  // namespace detail
// get_memory_extra_looking_sum (verifier.rs:319-512) for one challenge: sum over those rows of 1 / (gamma + sum_i row_i beta^i)
inline F memory_extra_looking_sum(const PublicValues& pv, F beta, F gamma, const H256& kernel_code_hash, uint64_t kernel_code_len,
                                  bool eth_mainnet = true, bool cdk_erigon = false) {
    F sum = 0;
    for (const MemoryLookupRow& r : memory_extra_looking_values(pv, kernel_code_hash, kernel_code_len, eth_mainnet, cdk_erigon)) {
        F acc = 0;
        for (size_t i = r.size(); i-- > 0;) acc = detail::gl_add(detail::gl_mul(acc, beta), r[i] % detail::GL_P);
        sum = detail::gl_add(sum, detail::gl_inv(detail::gl_add(acc, gamma)));
    }
    return sum;
}

// ---- a stream of segments, several in flight on one GPU (zero/src/prover.rs:205-236 dispatches every segment as its own proving job and
// collects the proofs by index; zero/src/ops.rs:24-66 is the job) ---------------------------------------------------------------------
// `streams` worker threads, each with its own worker state (default: a Context = own CUDA stream, copy stream and memory pool; four in
// flight is the measured optimum on a B200).  next() is called under a lock and returns std::nullopt at the end of the stream, so the
// source is consumed lazily: at most `streams` segments are alive at a time.  A failing segment raises the abort signal the running
// proofs poll and its exception is rethrown by prove_all; abort() does the same from outside (-> Error{ZKGPU_ERR_ABORTED}).
template <class Segment, class Proof, class Worker>
class SegmentStream {
public:
    using MakeWorker = std::function<std::unique_ptr<Worker>()>;
    using Prove = std::function<Proof(Worker&, const Segment&, AbortSignal)>;
    using Estimate = std::function<size_t(const Segment&)>;
    SegmentStream(unsigned streams, MakeWorker make_worker, Prove prove) : streams_(streams ? streams : 1), make_worker_(std::move(make_worker)), prove_(std::move(prove)) {}
    SegmentStream(unsigned streams, MakeWorker make_worker, Prove prove, size_t budget_bytes, Estimate estimate)
        : streams_(streams ? streams : 1), make_worker_(std::move(make_worker)), prove_(std::move(prove)), estimate_(std::move(estimate)), budget_(budget_bytes) {}
    // Admission by device memory: a segment starts when estimate(segment) bytes fit next to the ones in flight, or when nothing is in
    // flight (a segment larger than the whole budget still runs, alone; the library reports ZKGPU_ERR_NOMEM if it really does not fit).
    SegmentStream& with_memory_budget(size_t budget_bytes, Estimate estimate) { budget_ = budget_bytes; estimate_ = std::move(estimate); return *this; }
    void abort() { abort_.store(1); }
    // proofs in segment order
    std::vector<Proof> prove_all(const std::function<std::optional<Segment>()>& next) {
        abort_.store(0);
        std::mutex mu;
        std::map<size_t, Proof> results;
        std::exception_ptr first_error;
        size_t count = 0;
        bool done = false;
        auto worker = [&]() {
            try {
                std::unique_ptr<Worker> w = make_worker_();
                for (;;) {
                    std::optional<Segment> seg;
                    size_t idx;
                    {
                        std::lock_guard<std::mutex> g(mu);
                        if (done || abort_.load()) return;
                        seg = next();
                        if (!seg) { done = true; return; }
                        idx = count++;
                    }
                    const size_t need = estimate_ ? estimate_(*seg) : 0;
                    if (estimate_) {
                        std::unique_lock<std::mutex> g(mu);
                        while (running_ && used_ + need > budget_ && !abort_.load()) cv_.wait_for(g, std::chrono::milliseconds(50));
                        used_ += need; running_++;
                    }
                    struct Release {
                        SegmentStream* s; std::mutex& mu; size_t need; bool on;
                        ~Release() { if (on) { std::lock_guard<std::mutex> g(mu); s->used_ -= need; s->running_--; s->cv_.notify_all(); } }
                    } release{this, mu, need, (bool)estimate_};
                    if (abort_.load()) return;
                    Proof p = prove_(*w, *seg, &abort_);
                    std::lock_guard<std::mutex> g(mu);
                    results.emplace(idx, std::move(p));
                }
            } catch (...) {
                std::lock_guard<std::mutex> g(mu);
                if (!first_error) first_error = std::current_exception();
                abort_.store(1);
            }
        };
        std::vector<std::thread> threads;
        for (unsigned i = 0; i < streams_; i++) threads.emplace_back(worker);
        for (std::thread& t : threads) t.join();
        if (first_error) std::rethrow_exception(first_error);
        if (abort_.load()) throw Error(ZKGPU_ERR_ABORTED, "abort signal observed");
        std::vector<Proof> out;
        for (auto& kv : results) out.push_back(std::move(kv.second));
        return out;
    }
private:
    unsigned streams_;
    MakeWorker make_worker_;
    Prove prove_;
    std::atomic<int> abort_{0};
    Estimate estimate_;
    size_t budget_ = 0, used_ = 0, running_ = 0;
    std::condition_variable cv_;
};

// the product instance: segments as trace blocks + flattened public values, proved by prove_with_traces on `device`
struct SegmentInput {
    std::array<TableTrace, NUM_TABLES> traces;
    std::vector<F> public_values;
    int mem_kind = ZKGPU_MEM_HOST;
};
// Device memory one segment proof holds at its peak: every table's trace commitment (values + coefficients + blow-up-2 LDE = 32 n c
// bytes, ~128 n of digests) is alive before the first table is finished, plus the auxiliary / quotient / FRI buffers of the widest
// table, 15 % headroom.  Measured on a B200: 16.8 GB at the b19807080 heights against an estimate of 21.8 (profiles/r2y_memory_peak.jsonl).
inline size_t estimate_segment_bytes(const std::array<TableTrace, NUM_TABLES>& traces) {
    static const unsigned width[NUM_TABLES][2] = {{116, 100}, {71, 70}, {85, 24}, {2431, 4}, {438, 290}, {523, 2}, {30, 16}, {12, 4}, {12, 2}};
    double total = 0, transient = 0;
    for (size_t t = 0; t < NUM_TABLES; t++) {
        if (!traces[t].cols || !traces[t].n) continue;
        const double n = (double)traces[t].n;
        total += 32.0 * n * width[t][0] + 128.0 * n;
        transient = std::max(transient, 32.0 * n * (width[t][1] + 8) + 512.0 * n);
    }
    return (size_t)(1.15 * (total + transient));
}
inline SegmentStream<SegmentInput, AllProof, Context> segment_prover(int device, unsigned streams, StarkConfig config, KernelLabels labels,
                                                                     size_t memory_budget = (size_t)150 << 30) {
    using Stream = SegmentStream<SegmentInput, AllProof, Context>;
    return Stream(
        streams, [device]() { return std::make_unique<Context>(device); },
        [config, labels](Context& ctx, const SegmentInput& s, AbortSignal a) { return prove_with_traces(ctx, s.traces, s.public_values, config, labels, a, s.mem_kind); },
        memory_budget, memory_budget ? Stream::Estimate([](const SegmentInput& s) { return estimate_segment_bytes(s.traces); }) : Stream::Estimate());
}

// a trace finished in device memory (zkgpu_dev_trace): KeccakStark / LogicStark generate_trace, the Arithmetic range-check columns,
// the derived Memory columns (include/zkgpu.h "device-side trace finishing")
class DeviceTrace {
public:
    // KeccakStark::generate_trace(inputs_and_timestamps, min_rows)   keccak/keccak_stark.rs:70-250
    static DeviceTrace keccak_generate_trace(Context& ctx, const std::vector<std::pair<std::array<uint64_t, 25>, size_t>>& inputs_and_timestamps,
                                             size_t min_rows) {
        std::vector<uint64_t> in(inputs_and_timestamps.size() * 25), ts(inputs_and_timestamps.size());
        for (size_t i = 0; i < ts.size(); i++) {
            for (int k = 0; k < 25; k++) in[25 * i + k] = inputs_and_timestamps[i].first[k];
            ts[i] = inputs_and_timestamps[i].second;
        }
        zkgpu_dev_trace* t = nullptr;
        check(zkgpu_keccak_generate_trace(ctx.handle(), in.data(), ts.data(), ts.size(), min_rows, &t));
        ctx.sync();                                     // `in` / `ts` are temporaries
        return DeviceTrace(t);
    }
    // LogicStark::generate_trace(operations, min_rows)   logic.rs:165-237; one operation = {operator, input0 limbs, input1 limbs}
    static DeviceTrace logic_generate_trace(Context& ctx, const std::vector<std::array<uint64_t, 9>>& operations, size_t min_rows) {
        zkgpu_dev_trace* t = nullptr;
        check(zkgpu_logic_generate_trace(ctx.handle(), operations.empty() ? nullptr : operations[0].data(), operations.size(), min_rows, &t));
        ctx.sync();
        return DeviceTrace(t);
    }
    TableTrace as_table_trace() const {
        size_t nc = 0, n = 0;
        check(zkgpu_dev_trace_dims(h_.get(), &nc, &n));
        return TableTrace{zkgpu_dev_trace_ptr(h_.get()), n};
    }
    std::vector<F> to_host() const {
        size_t nc = 0, n = 0;
        check(zkgpu_dev_trace_dims(h_.get(), &nc, &n));
        std::vector<F> v(nc * n);
        check(zkgpu_dev_trace_export(h_.get(), v.data()));
        return v;
    }
    zkgpu_dev_trace* handle() const { return h_.get(); }
private:
    struct Del { void operator()(zkgpu_dev_trace* t) const { zkgpu_dev_trace_free(t); } };
    explicit DeviceTrace(zkgpu_dev_trace* t) : h_(t) {}
    std::unique_ptr<zkgpu_dev_trace, Del> h_;
};

}  // namespace zkgpu
