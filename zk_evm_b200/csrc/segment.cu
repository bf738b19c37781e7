// Segment-level entry points: the shared Fiat-Shamir transcript, the split per-table job (for table-sharded multi-GPU
// runs) and the single-device `prove_with_traces`.
//
// Replaces /root/reference/evm_arithmetization/src/prover.rs:72-194 (prove_with_traces: commit every trace, observe the
// caps and the public values, draw the CTL challenges, get_ctl_data) and :211-293 (prove_with_commitments: the tables
// proven one after the other with the same challenger, in Table order).
#include "stark_dev.h"
#include "challenger.h"
#include <string.h>
#include <algorithm>

#include <mutex>
#include <condition_variable>

struct zkgpu_challenger { zk::Challenger ch; };

namespace zk {
// One host->device upload chain at a time per device.  With several segments in flight (one context + host thread each) their
// trace uploads would otherwise interleave on the one H2D copy engine: every segment then waits ~N times longer for its first
// tables and all of them reach the compute phase together.  Serialised, the first segment's tables arrive at full PCIe speed and
// the next segment uploads underneath the first one's commitments (measured, two segments in flight: 645-915 ms -> see profiles/).
struct UploadGate { std::mutex mu; std::condition_variable cv; bool busy = false; };
static UploadGate g_upload_gate[16];
static void CUDART_CB upload_gate_release(void* p) {
    UploadGate* g = static_cast<UploadGate*>(p);
    { std::lock_guard<std::mutex> l(g->mu); g->busy = false; }
    g->cv.notify_one();
}
}  // namespace zk

namespace zk {

// observe the trace caps in table order (zero cap for an optional table that is not in use, prover.rs:118-127), then the
// public values (already flattened by the host in the order of get_challenges.rs:202-227), then draw (beta, gamma) per challenge
// (starky get_grand_product_challenge_set, reached through get_ctl_data at prover.rs:137-143)
static void segment_transcript(Challenger& ch, const uint64_t* caps /*[9][cap_len*4]*/, const uint8_t in_use[ZKGPU_NUM_TABLES], size_t cap_words,
                               const uint64_t* public_values, size_t n_public, unsigned num_challenges, uint64_t* beta_gamma) {
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) {
        if (!in_use[t]) {
            ZK_REQUIRE(zkstark::table_is_optional(t), "only optional tables may be left out (all_stark.rs:110-117)");
            for (size_t i = 0; i < cap_words; i++) ch.observe(0);
        } else {
            ch.observe_n(caps + (size_t)t * cap_words, cap_words);
        }
    }
    ch.observe_n(public_values, n_public);
    for (unsigned i = 0; i < num_challenges; i++) { beta_gamma[2 * i] = ch.challenge(); beta_gamma[2 * i + 1] = ch.challenge(); }
}

}  // namespace zk

using namespace zk;

extern "C" {

// ---- transcript ----------------------------------------------------------------------------------------------------------
int zkgpu_challenger_new(zkgpu_challenger** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(out, "null argument");
    *out = new zkgpu_challenger();
    ZK_API_END
}
void zkgpu_challenger_free(zkgpu_challenger* ch) { delete ch; }
int zkgpu_challenger_observe(zkgpu_challenger* ch, const uint64_t* elements, size_t n) {
    ZK_API_BEGIN
    ZK_REQUIRE(ch && (elements || n == 0), "null argument");
    for (size_t i = 0; i < n; i++) {
        ZK_REQUIRE(elements[i] < GL_P, "observed element is not canonical");
        ch->ch.observe(elements[i]);
    }
    ZK_API_END
}
int zkgpu_challenger_get_challenges(zkgpu_challenger* ch, uint64_t* out, size_t n) {
    ZK_API_BEGIN
    ZK_REQUIRE(ch && (out || n == 0), "null argument");
    for (size_t i = 0; i < n; i++) out[i] = ch->ch.challenge();
    ZK_API_END
}
int zkgpu_challenger_compact(zkgpu_challenger* ch, uint64_t state_out[12]) {
    ZK_API_BEGIN
    ZK_REQUIRE(ch && state_out, "null argument");
    ch->ch.compact();
    memcpy(state_out, ch->ch.state, 96);
    ZK_API_END
}
int zkgpu_challenger_set_state(zkgpu_challenger* ch, const uint64_t state[12]) {
    ZK_API_BEGIN
    ZK_REQUIRE(ch && state, "null argument");
    ch->ch.set_state(state);
    ZK_API_END
}
int zkgpu_segment_challenges(const uint64_t* trace_caps, const uint8_t* table_in_use, uint32_t cap_height, const uint64_t* public_values,
                             size_t n_public_values, uint32_t num_challenges, uint64_t* beta_gamma, uint64_t challenger_state[12]) {
    ZK_API_BEGIN
    ZK_REQUIRE(trace_caps && table_in_use && beta_gamma && challenger_state && (public_values || n_public_values == 0), "null argument");
    ZK_REQUIRE(num_challenges >= 1 && num_challenges <= 2 && cap_height <= 16, "bad config");
    Challenger ch;
    segment_transcript(ch, trace_caps, table_in_use, (size_t)4 << cap_height, public_values, n_public_values, num_challenges, beta_gamma);
    // no input is pending after a challenge was drawn, so compact() only drops the buffered outputs (as prover.rs:320 will)
    ch.compact();
    memcpy(challenger_state, ch.state, 96);
    ZK_API_END
}

// ---- split per-table job -------------------------------------------------------------------------------------------------
int zkgpu_table_job_begin(zkgpu_ctx* h, uint32_t table_id, const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config,
                          const zkgpu_batch* trace, const zkgpu_ctl* ctl, volatile const int* abort_flag, zkgpu_table_job** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && config && trace && ctl && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_table_job> j(new zkgpu_table_job());
    prove_table_begin(c, table_id, params_from(labels), config_from(config), trace->b, ctl->c, abort_flag, j->j);
    *out = j.release();
    ZK_API_END
}
int zkgpu_table_job_aux_cap(const zkgpu_table_job* job, uint64_t* out_cap, size_t* len_words) {
    ZK_API_BEGIN
    ZK_REQUIRE(job && len_words, "null argument");
    const std::vector<uint64_t>& cap = job->j.aux ? job->j.aux->b.cap_host : std::vector<uint64_t>();
    if (out_cap && *len_words >= cap.size() && !cap.empty()) memcpy(out_cap, cap.data(), cap.size() * 8);
    *len_words = cap.size();
    ZK_API_END
}
int zkgpu_table_job_finish(zkgpu_table_job* job, uint64_t challenger_state[12], const uint64_t* forced_pow_witness,
                           volatile const int* abort_flag, zkgpu_proof** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(job && challenger_state && out, "null argument");
    ZK_REQUIRE(job->j.begun, "table job already finished");
    Ctx& c = *job->j.ctx;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_proof> hp(new zkgpu_proof());
    uint64_t st[12];
    memcpy(st, challenger_state, 96);
    prove_table_finish(c, job->j, st, forced_pow_witness, abort_flag, hp->p);
    memcpy(challenger_state, st, 96);
    *out = hp.release();
    ZK_API_END
}
void zkgpu_table_job_free(zkgpu_table_job* job) {
    if (!job) return;
    if (job->j.ctx) cudaSetDevice(job->j.ctx->device);
    delete job;
}

// ---- prove_with_traces on one device ---------------------------------------------------------------------------------------
int zkgpu_prove_segment(zkgpu_ctx* h, const zkgpu_table_trace* traces, int mem_kind, const uint64_t* public_values, size_t n_public_values,
                        const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config, const uint64_t* forced_pow_witnesses,
                        volatile const int* abort_flag, zkgpu_proof** proofs_out, uint64_t* ctl_challenges_out, uint64_t* trace_caps_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && traces && config && proofs_out && (public_values || n_public_values == 0), "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    const zkstark::Config cfg = config_from(config);
    const zkstark::TableParams prm = params_from(labels);
    const size_t cap_words = (size_t)4 << cfg.cap_height;
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) proofs_out[t] = nullptr;

    // 1. trace commitments (prover.rs:92-116).  The caps enter the transcript in Table order only after every table is
    // committed, so the tables are uploaded and committed smallest first: all uploads are queued on the copy stream up front
    // (one event per table) and the upload of table k+1 runs under the commitment of table k — hashing a trace takes longer than
    // moving it over PCIe, so after the first (smallest) table the copies are hidden.
    StageLog lg(c);
    std::unique_ptr<zkgpu_batch> tb[ZKGPU_NUM_TABLES];
    std::vector<uint64_t> caps(ZKGPU_NUM_TABLES * cap_words, 0);
    uint8_t in_use[ZKGPU_NUM_TABLES];
    std::vector<uint32_t> order;
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) {
        in_use[t] = traces[t].cols != nullptr;
        if (!in_use[t]) continue;
        tb[t].reset(new zkgpu_batch());
        Batch& b = tb[t]->b;
        init_batch(c, b, zkstark::table_num_columns(t), traces[t].n, cfg.rate_bits, cfg.cap_height);
        b.values = DevBuf(&c, b.ncols * b.n * 8);
        order.push_back(t);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return tb[x]->b.ncols * tb[x]->b.n < tb[y]->b.ncols * tb[y]->b.n; });
    struct EventGuard {
        std::vector<cudaEvent_t> ev;
        ~EventGuard() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
        cudaEvent_t make() { cudaEvent_t e; ZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ev.push_back(e); return e; }
    } events;
    cudaEvent_t uploaded[ZKGPU_NUM_TABLES] = {nullptr};
    {
        // the buffers are stream-ordered allocations of c.stream: the copy stream may touch them only after that point
        cudaEvent_t allocated = events.make();
        ZK_CUDA(cudaEventRecord(allocated, c.stream));
        ZK_CUDA(cudaStreamWaitEvent(c.copy_stream, allocated, 0));
        UploadGate* gate = mem_kind == ZKGPU_MEM_DEVICE ? nullptr : &g_upload_gate[c.device & 15];
        if (gate) {
            std::unique_lock<std::mutex> l(gate->mu);
            gate->cv.wait(l, [&] { return !gate->busy; });
            gate->busy = true;
        }
        try {
            for (uint32_t t : order) {
                Batch& b = tb[t]->b;
                ZK_CUDA(cudaMemcpyAsync(b.values.get(), traces[t].cols, b.ncols * b.n * 8,
                                        mem_kind == ZKGPU_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c.copy_stream));
                uploaded[t] = events.make();
                ZK_CUDA(cudaEventRecord(uploaded[t], c.copy_stream));
            }
            // the gate opens when the last copy of this chain has executed (host callback in copy-stream order)
            if (gate) ZK_CUDA(cudaLaunchHostFunc(c.copy_stream, upload_gate_release, gate));
        } catch (...) {
            if (gate) upload_gate_release(gate);
            throw;
        }
    }
    struct CopyDrain {   // on any exit (errors included) the copy stream is drained before the buffers it writes are freed
        Ctx& c; ~CopyDrain() { cudaStreamSynchronize(c.copy_stream); }
    } drain{c};
    for (uint32_t t : order) {
        if (abort_flag && *abort_flag) throw ZkError(ZKGPU_ERR_ABORTED, "abort signal observed (prover.rs:346-354)");
        Batch& b = tb[t]->b;
        ZK_CUDA(cudaStreamWaitEvent(c.stream, uploaded[t], 0));
        lg.mark("trace upload");
        commit_from_device_values(c, b, true);
        lg.mark(zkstark::table_name(t));
        memcpy(&caps[t * cap_words], b.cap_host.data(), cap_words * 8);
    }
    if (trace_caps_out) memcpy(trace_caps_out, caps.data(), caps.size() * 8);

    // 2. transcript: caps, public values, CTL challenges (prover.rs:118-144)
    Challenger ch;
    uint64_t bg[4] = {0, 0, 0, 0};
    segment_transcript(ch, caps.data(), in_use, cap_words, public_values, n_public_values, cfg.num_challenges, bg);
    if (ctl_challenges_out) memcpy(ctl_challenges_out, bg, 2 * cfg.num_challenges * 8);
    ch.compact();
    uint64_t st[12];
    memcpy(st, ch.state, 96);

    // 3. per table: CTL data, then the proof, chained through the transcript state in Table order (prover.rs:251-259)
    std::unique_ptr<zkgpu_proof> proofs[ZKGPU_NUM_TABLES];
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) {
        if (!in_use[t]) continue;
        Ctl ctl;
        lg.mark("--");
        make_ctl_data(c, t, tb[t]->b, bg, cfg.num_challenges, ctl);
        lg.mark("ctl data");
        proofs[t].reset(new zkgpu_proof());
        prove_table(c, t, prm, cfg, tb[t]->b, ctl, st, forced_pow_witnesses ? &forced_pow_witnesses[t] : nullptr, abort_flag, proofs[t]->p);
        tb[t].reset();   // release the table's device memory before the next one
        lg.mark(zkstark::table_name(t));
    }
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) proofs_out[t] = proofs[t].release();
    ZK_API_END
}

}  // extern "C"
