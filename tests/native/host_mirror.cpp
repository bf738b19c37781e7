// C++ host harness over include/zkgpu.hpp (the compiled-language mirror of the reference interface).  Driven by tests/test_host_mirror.py:
//   host_mirror nodevice                         Context(0) must throw Error{ZKGPU_ERR_CUDA} on a box without a GPU (no CPU path)
//   host_mirror challenger <n>                   observe 1..n, draw 3 challenges, compact -> prints the challenges and the state
//   host_mirror pubvals <file>                   PublicValues read from a packed byte file -> the flattened elements observe_public_values feeds
//   host_mirror stream                           SegmentStream (the scheduler template) with a stand-in prover: order, laziness, failure, abort
//   host_mirror decode <proof.words>             typed StarkProof fields of a serialised proof, and re-serialisation equality
//   host_mirror prove <segment.trace> <out> [test|fast]   prove_with_traces on cuda:0 from a segment trace file (zk_evm_b200/trace_file.py
//                                                layout) -> writes ctl challenges, trace caps and every table's proof words to <out>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <iterator>
#include "zkgpu.hpp"

using namespace zkgpu;

static std::vector<uint64_t> read_words(const char* path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    const size_t bytes = (size_t)f.tellg();
    std::vector<uint64_t> w(bytes / 8);
    f.seekg(0);
    f.read((char*)w.data(), (std::streamsize)(w.size() * 8));
    return w;
}

// segment trace container (trace_file.py): header "ZKSEGTR1", u32 version, u32 num_tables, u64 n_public, u64 labels[4], then per table
// u32 in_use, u32 num_columns, u64 n, u64 offset; public values at the next 64-byte boundary; column-major blocks at `offset`
struct Segment {
    std::vector<uint8_t> bytes;
    std::array<TableTrace, NUM_TABLES> traces;
    std::vector<F> public_values;
    KernelLabels labels;
};
static Segment load_segment(const char* path) {
    Segment s;
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    s.bytes.resize((size_t)f.tellg());
    f.seekg(0);
    f.read((char*)s.bytes.data(), (std::streamsize)s.bytes.size());
    const uint8_t* b = s.bytes.data();
    if (s.bytes.size() < 56 + 24 * NUM_TABLES || memcmp(b, "ZKSEGTR1", 8)) throw std::runtime_error("not a segment trace file");
    uint32_t version, ntab; uint64_t npv, lab[4];
    memcpy(&version, b + 8, 4); memcpy(&ntab, b + 12, 4); memcpy(&npv, b + 16, 8); memcpy(lab, b + 24, 32);
    if (version != 1 || ntab != NUM_TABLES) throw std::runtime_error("unsupported trace file version");
    s.labels = KernelLabels{lab[0], lab[1], lab[2], lab[3]};
    const size_t dir = 56, pv_off = (dir + 24 * NUM_TABLES + 63) / 64 * 64;
    s.public_values.resize(npv);
    memcpy(s.public_values.data(), b + pv_off, npv * 8);
    for (size_t t = 0; t < NUM_TABLES; t++) {
        uint32_t in_use, nc; uint64_t n, off;
        memcpy(&in_use, b + dir + 24 * t, 4); memcpy(&nc, b + dir + 24 * t + 4, 4); memcpy(&n, b + dir + 24 * t + 8, 8); memcpy(&off, b + dir + 24 * t + 16, 8);
        if (in_use) {
            if (off + 8 * (uint64_t)nc * n > s.bytes.size()) throw std::runtime_error("truncated trace file");
            s.traces[t] = TableTrace{(const F*)(b + off), (size_t)n};
        }
    }
    return s;
}

int main(int argc, char** argv) {
    try {
        const std::string mode = argc > 1 ? argv[1] : "";
        if (mode == "nodevice") {
            try {
                Context ctx(0);
                printf("device present\n");
                return 0;
            } catch (const Error& e) {
                printf("error %d: %s\n", e.code, e.what());
                return e.code == ZKGPU_ERR_CUDA ? 0 : 1;
            }
        }
        if (mode == "challenger" && argc > 2) {
            Challenger ch;
            std::vector<F> xs;
            for (int i = 1; i <= atoi(argv[2]); i++) xs.push_back((F)i);
            ch.observe_elements(xs);
            for (F c : ch.get_n_challenges(3)) printf("%llu\n", (unsigned long long)c);
            ch.observe_element(7);
            for (F c : ch.compact()) printf("%llu\n", (unsigned long long)c);
            auto ch2 = Challenger::from_state(ch.compact());
            printf("%llu\n", (unsigned long long)ch2->get_challenge());
            return 0;
        }
        if (mode == "pubvals" && argc > 2) {
            std::ifstream f(argv[2], std::ios::binary);
            std::vector<uint8_t> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            size_t pos = 0;
            auto h256 = [&]() { H256 h; memcpy(h.data(), &b.at(pos + 31) - 31, 32); pos += 32; return h; };
            auto u256 = [&]() { U256 x = U256::from_big_endian(&b.at(pos + 31) - 31, 32); pos += 32; return x; };
            PublicValues pv;
            for (TrieRoots* r : {&pv.trie_roots_before, &pv.trie_roots_after}) { r->state_root = h256(); r->transactions_root = h256(); r->receipts_root = h256(); }
            BlockMetadata& m = pv.block_metadata;
            memcpy(m.block_beneficiary.data(), &b.at(pos + 19) - 19, 20); pos += 20;
            m.block_timestamp = u256(); m.block_number = u256(); m.block_difficulty = u256(); m.block_random = h256();
            m.block_gaslimit = u256(); m.block_chain_id = u256(); m.block_base_fee = u256(); m.block_gas_used = u256();
            m.block_blob_gas_used = u256(); m.block_excess_blob_gas = u256(); m.parent_beacon_block_root = h256();
            for (U256& w : m.block_bloom) w = u256();
            for (H256& h : pv.block_hashes.prev_hashes) h = h256();
            pv.block_hashes.cur_hash = h256();
            ExtraBlockData& e = pv.extra_block_data;
            e.checkpoint_state_trie_root = h256();
            memcpy(e.checkpoint_consolidated_hash.data(), &b.at(pos + 31) - 31, 32); pos += 32;
            e.txn_number_before = u256(); e.txn_number_after = u256(); e.gas_used_before = u256(); e.gas_used_after = u256();
            if (pos == b.size()) {
                for (F x : flatten_public_values(pv)) printf("%llu\n", (unsigned long long)x);
                return 0;
            }
            // the longer form: registers before / after, kernel hash, kernel length, beta, gamma -> the extra Memory lookup rows and their sum
            for (RegistersData* r : {&pv.registers_before, &pv.registers_after}) {
                r->program_counter = u256(); r->is_kernel = u256(); r->stack_len = u256(); r->stack_top = u256(); r->context = u256(); r->gas_used = u256();
            }
            H256 kh = h256();
            uint64_t klen = u256().limbs[0], beta = u256().limbs[0], gamma = u256().limbs[0];
            if (pos != b.size()) throw std::runtime_error("trailing bytes");
            for (const auto& r : memory_extra_looking_values(pv, kh, klen)) for (F x : r) printf("%llu\n", (unsigned long long)x);
            printf("%llu\n", (unsigned long long)memory_extra_looking_sum(pv, beta, gamma, kh, klen));
            return 0;
        }
        if (mode == "stream") {
            // the scheduling logic without a device: the "proof" of segment k is k * k, the workers count what they were given
            struct W { int made = 0; };
            std::atomic<int> workers{0}, alive{0}, max_alive{0};
            SegmentStream<int, long, W> st(3, [&]() { workers++; return std::make_unique<W>(); },
                                           [&](W&, const int& k, AbortSignal a) -> long {
                                               int now = ++alive, m = max_alive.load();
                                               while (now > m && !max_alive.compare_exchange_weak(m, now)) {}
                                               for (int i = 0; i < 20; i++) {
                                                   if (a->load()) { alive--; throw Error(ZKGPU_ERR_ABORTED, "aborted"); }
                                                   std::this_thread::sleep_for(std::chrono::milliseconds(1));
                                               }
                                               if (k == 1000) { alive--; throw std::runtime_error("segment 1000 is broken"); }
                                               alive--;
                                               return (long)k * k;
                                           });
            int produced = 0;
            auto src = [&](int n, int poison) { produced = 0; return [&produced, n, poison]() -> std::optional<int> {
                if (produced >= n) return std::nullopt;
                int k = produced++;
                return k == poison ? 1000 : k; }; };
            const std::vector<long> got = st.prove_all(src(17, -1));
            bool ok = got.size() == 17 && workers == 3 && max_alive <= 3;
            for (size_t k = 0; k < got.size(); k++) ok = ok && got[k] == (long)(k * k);
            printf("order %d\n", (int)ok);
            bool failed = false;
            try { st.prove_all(src(500, 5)); } catch (const std::runtime_error& e) { failed = std::string(e.what()).find("1000") != std::string::npos; }
            printf("failure %d consumed_at_most %d\n", (int)failed, produced);
            std::thread killer([&]() { std::this_thread::sleep_for(std::chrono::milliseconds(50)); st.abort(); });
            bool aborted = false;
            try { st.prove_all(src(100000, -1)); } catch (const Error& e) { aborted = e.code == ZKGPU_ERR_ABORTED; }
            killer.join();
            printf("abort %d consumed_at_most %d\n", (int)aborted, produced);
            // admission by device memory: every "segment" needs 10 units; a budget of 25 lets two run side by side, 5 one at a time
            bool budget_ok = true;
            for (auto [budget, want_max] : {std::pair<size_t, int>{25, 2}, {5, 1}, {1000, 3}}) {
                max_alive = 0;
                st.with_memory_budget(budget, [](const int&) { return (size_t)10; });
                const std::vector<long> g2 = st.prove_all(src(12, -1));
                budget_ok = budget_ok && g2.size() == 12 && max_alive == want_max;
                for (size_t k = 0; k < g2.size(); k++) budget_ok = budget_ok && g2[k] == (long)(k * k);
            }
            std::array<TableTrace, NUM_TABLES> shapes{};
            static const F dummy = 0;
            const unsigned lg[NUM_TABLES] = {17, 14, 19, 17, 13, 16, 21, 19, 19};
            for (size_t t = 0; t < NUM_TABLES; t++) shapes[t] = TableTrace{&dummy, (size_t)1 << lg[t]};
            const size_t est = estimate_segment_bytes(shapes);
            printf("budget %d estimate %zu\n", (int)budget_ok, est);
            return ok && failed && aborted && budget_ok ? 0 : 1;
        }
        if (mode == "decode" && argc > 2) {
            const std::vector<uint64_t> w = read_words(argv[2]);
            size_t used = 0;
            const zkstark::StarkProofData p = zkstark::deserialize_proof(w.data(), w.size(), &used);
            const bool same = zkstark::serialize_proof(p) == w && used == w.size();
            printf("table %llu degree_bits %llu trace_cap %zu aux_cap %zu quotient_cap %zu local %zu next %zu aux %zu ctl_zs_first %zu quotient %zu "
                   "layers %zu queries %zu final_poly %zu pow %llu roundtrip %d\n",
                   (unsigned long long)p.table_id, (unsigned long long)p.degree_bits, p.trace_cap.size() / 4, p.aux_cap.size() / 4,
                   p.quotient_cap.size() / 4, p.local_values.size() / 2, p.next_values.size() / 2, p.aux_polys.size() / 2, p.ctl_zs_first.size(),
                   p.quotient_polys.size() / 2, p.commit_phase_caps.size(), p.queries.size(), p.final_poly.size() / 2,
                   (unsigned long long)p.pow_witness, (int)same);
            return same ? 0 : 1;
        }
        if (mode == "prove" && argc > 3) {
            const Segment seg = load_segment(argv[2]);
            const StarkConfig cfg = (argc > 4 && std::string(argv[4]) == "fast") ? StarkConfig::standard_fast_config() : StarkConfig::test_config();
            Context ctx(0);
            std::atomic<int> abort_signal{0};
            const AllProof ap = prove_with_traces(ctx, seg.traces, seg.public_values, cfg, seg.labels, &abort_signal);
            // the same Cpu trace through the per-table seam: from_values on its columns gives the cap prove_with_traces observed
            {
                const TableTrace& ct = seg.traces[(size_t)Table::Cpu];
                std::vector<PolynomialValues> cols(85);
                for (size_t c = 0; c < cols.size(); c++) cols[c].assign(ct.cols + c * ct.n, ct.cols + (c + 1) * ct.n);
                const PolynomialBatch batch = PolynomialBatch::from_values(ctx, cols, cfg.rate_bits, cfg.cap_height);
                if (batch.cap() != ap.trace_caps[(size_t)Table::Cpu]) { printf("from_values cap differs from the segment's\n"); return 1; }
                if (batch.num_polys() != 85 || batch.degree() != ct.n) { printf("batch dims\n"); return 1; }
            }
            // the same segment through the reference's seams: from_values per table (S1) -> transcript + CTL challenges -> get_ctl_data (S2)
            // -> prove_with_commitments (S3, one shared challenger) must give the proofs of the one-call path, word for word
            {
                std::vector<std::unique_ptr<PolynomialBatch>> batches(NUM_TABLES);
                std::vector<std::unique_ptr<CtlData>> ctls(NUM_TABLES);
                std::array<std::optional<MerkleCap>, NUM_TABLES> caps;
                const uint32_t widths[NUM_TABLES] = {116, 71, 85, 2431, 438, 523, 30, 12, 12};
                for (size_t t = 0; t < NUM_TABLES; t++) {
                    if (!seg.traces[t].cols) continue;
                    batches[t] = std::make_unique<PolynomialBatch>(PolynomialBatch::from_values_contig(ctx, seg.traces[t].cols, widths[t], seg.traces[t].n,
                                                                                                       cfg.rate_bits, cfg.cap_height));
                    caps[t] = batches[t]->cap();
                }
                Challenger challenger;
                const GrandProductChallengeSet ctl_challenges = segment_challenges(challenger, caps, seg.public_values, cfg);
                if (ctl_challenges.beta_gamma != ap.ctl_challenges.beta_gamma) { printf("CTL challenges differ\n"); return 1; }
                std::array<const PolynomialBatch*, NUM_TABLES> bp{};
                std::array<const CtlData*, NUM_TABLES> cp{};
                for (size_t t = 0; t < NUM_TABLES; t++) {
                    if (!batches[t]) continue;
                    ctls[t] = std::make_unique<CtlData>(get_ctl_data(ctx, (Table)t, *batches[t], ctl_challenges, cfg.num_challenges));
                    bp[t] = batches[t].get(); cp[t] = ctls[t].get();
                }
                const auto proofs = prove_with_commitments(ctx, cfg, bp, cp, challenger, seg.labels);
                for (size_t t = 0; t < NUM_TABLES; t++) {
                    if (proofs[t].has_value() != ap.stark_proofs[t].has_value() || (proofs[t] && proofs[t]->words != ap.stark_proofs[t]->words)) {
                        printf("prove_with_commitments differs from prove_with_traces on table %zu\n", t);
                        return 1;
                    }
                }
            }
            // a stream of four copies of the segment, two in flight (segment_prover = SegmentStream over Context + prove_with_traces)
            {
                auto sp = segment_prover(0, 2, cfg, seg.labels);
                int left = 4;
                const std::vector<AllProof> all = sp.prove_all([&]() -> std::optional<SegmentInput> {
                    if (left-- <= 0) return std::nullopt;
                    return SegmentInput{seg.traces, seg.public_values, ZKGPU_MEM_HOST};
                });
                if (all.size() != 4) { printf("segment stream returned %zu proofs\n", all.size()); return 1; }
                for (const AllProof& a : all)
                    for (size_t t = 0; t < NUM_TABLES; t++)
                        if (a.stark_proofs[t].has_value() != ap.stark_proofs[t].has_value() || (a.stark_proofs[t] && a.stark_proofs[t]->words != ap.stark_proofs[t]->words)) {
                            printf("segment stream differs from prove_with_traces on table %zu\n", t);
                            return 1;
                        }
            }
            // an abort signal that is already raised stops the proof with ZKGPU_ERR_ABORTED (check_abort_signal, prover.rs:346-354)
            abort_signal = 1;
            try {
                prove_with_traces(ctx, seg.traces, seg.public_values, cfg, seg.labels, &abort_signal);
                printf("abort signal ignored\n");
                return 1;
            } catch (const Error& e) {
                if (e.code != ZKGPU_ERR_ABORTED) throw;
            }
            std::ofstream out(argv[3], std::ios::binary);
            auto put = [&](const std::vector<uint64_t>& v) { uint64_t l = v.size(); out.write((const char*)&l, 8); out.write((const char*)v.data(), (std::streamsize)(8 * v.size())); };
            put(ap.ctl_challenges.beta_gamma);
            for (size_t t = 0; t < NUM_TABLES; t++) {
                std::vector<uint64_t> cap;
                for (const Hash& h : ap.trace_caps[t]) cap.insert(cap.end(), h.begin(), h.end());
                put(cap);
                put(ap.stark_proofs[t] ? ap.stark_proofs[t]->words : std::vector<uint64_t>());
            }
            printf("proved %zu tables\n", (size_t)std::count_if(ap.stark_proofs.begin(), ap.stark_proofs.end(), [](const auto& p) { return p.has_value(); }));
            return 0;
        }
        fprintf(stderr, "usage: host_mirror nodevice | challenger <n> | decode <proof.words> | prove <segment.trace> <out> [test|fast]\n");
        return 2;
    } catch (const std::exception& e) {
        fprintf(stderr, "host_mirror: %s\n", e.what());
        return 1;
    }
}
