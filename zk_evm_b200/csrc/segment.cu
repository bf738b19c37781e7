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

struct zkgpu_challenger { zk::Challenger ch; };

namespace zk {
// One host->device upload chain at a time per device, ordered ON THE DEVICE.  With several segments in flight (one context + host
// thread each) their trace uploads would otherwise interleave on the one H2D copy engine: every segment then waits ~N times longer
// for its first tables.  Each chain's first copy waits (cudaStreamWaitEvent on the context's copy stream) for the event that ends
// the previously queued chain of the device; the host never blocks, so a thread can queue the uploads of its next segment and go
// straight on to prove the current one.
struct UploadGate { std::mutex mu; cudaEvent_t last = nullptr; };
static UploadGate g_upload_gate[16];
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
            for (size_t i = 0; i < cap_words; i++) ZK_REQUIRE(caps[(size_t)t * cap_words + i] < GL_P, "cap element is not canonical");
            ch.observe_n(caps + (size_t)t * cap_words, cap_words);
        }
    }
    // the same rule as zkgpu_challenger_observe: a non-canonical word would give a transcript no reference verifier reproduces
    for (size_t i = 0; i < n_public; i++) ZK_REQUIRE(public_values[i] < GL_P, "public value is not canonical");
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
}  // extern "C"

// the traces of one segment on their way to (or already in) device memory: one values buffer + one "uploaded" event per table
struct zkgpu_upload {
    zk::Ctx* ctx = nullptr;
    std::unique_ptr<zkgpu_batch> tb[ZKGPU_NUM_TABLES];
    uint8_t in_use[ZKGPU_NUM_TABLES] = {0};
    std::vector<uint32_t> order;                       // tables in upload / commit order (smallest first)
    cudaEvent_t uploaded[ZKGPU_NUM_TABLES] = {nullptr};
    std::vector<cudaEvent_t> events;
    uint32_t rate_bits = 0, cap_height = 0;
    bool consumed = false;                             // set by the one prove call that takes the buffers (zkgpu.h contract)
    cudaEvent_t make_event() { cudaEvent_t e; ZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); events.push_back(e); return e; }
    ~zkgpu_upload() {
        // the copy stream is drained before the buffers it writes are freed (errors included)
        if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->copy_stream); }
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
};

namespace zk {

// Queue the uploads of all tables on the copy stream (smallest table first, one event per table).  The caps enter the transcript in
// Table order only after every table is committed, so the commit order is free: the upload of table k+1 runs under the commitment
// of table k — hashing a trace takes longer than moving it over PCIe, so after the first (smallest) table the copies are hidden.
// Called one segment AHEAD (zkgpu_segment_upload for segment s+1 before zkgpu_prove_segment_uploaded for segment s) the whole chain
// runs under the previous segment's proof.
static void segment_upload(Ctx& c, const zkgpu_table_trace* traces, int mem_kind, const zkstark::Config& cfg, zkgpu_upload& u) {
    u.ctx = &c; u.rate_bits = cfg.rate_bits; u.cap_height = cfg.cap_height;
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) {
        u.in_use[t] = traces[t].cols != nullptr;
        if (!u.in_use[t]) {
            ZK_REQUIRE(zkstark::table_is_optional(t), "only optional tables may be left out (all_stark.rs:110-117)");
            continue;
        }
        u.tb[t].reset(new zkgpu_batch());
        Batch& b = u.tb[t]->b;
        init_batch(c, b, zkstark::table_num_columns(t), traces[t].n, cfg.rate_bits, cfg.cap_height);
        b.values = DevBuf(&c, b.ncols * b.n * 8);
        u.order.push_back(t);
    }
    std::stable_sort(u.order.begin(), u.order.end(),
                     [&](uint32_t x, uint32_t y) { return u.tb[x]->b.ncols * u.tb[x]->b.n < u.tb[y]->b.ncols * u.tb[y]->b.n; });
    // the buffers are stream-ordered allocations of c.stream: the copy stream may touch them only after that point
    cudaEvent_t allocated = u.make_event();
    ZK_CUDA(cudaEventRecord(allocated, c.stream));
    ZK_CUDA(cudaStreamWaitEvent(c.copy_stream, allocated, 0));
    UploadGate* gate = mem_kind == ZKGPU_MEM_DEVICE ? nullptr : &g_upload_gate[c.device & 15];
    std::unique_lock<std::mutex> lock;
    if (gate) {
        lock = std::unique_lock<std::mutex>(gate->mu);      // held while the chain is queued only
        if (gate->last) ZK_CUDA(cudaStreamWaitEvent(c.copy_stream, gate->last, 0));
    }
    for (uint32_t t : u.order) {
        Batch& b = u.tb[t]->b;
        ZK_CUDA(cudaMemcpyAsync(b.values.get(), traces[t].cols, b.ncols * b.n * 8,
                                mem_kind == ZKGPU_MEM_DEVICE ? cudaMemcpyDeviceToDevice
                                : mem_kind == ZKGPU_MEM_AUTO ? cudaMemcpyDefault : cudaMemcpyHostToDevice, c.copy_stream));
        u.uploaded[t] = u.make_event();
        ZK_CUDA(cudaEventRecord(u.uploaded[t], c.copy_stream));
    }
    if (gate) {
        // the next chain of this device starts after the last copy of this one (the old event is released once it has completed)
        cudaEvent_t end;
        ZK_CUDA(cudaEventCreateWithFlags(&end, cudaEventDisableTiming));
        ZK_CUDA(cudaEventRecord(end, c.copy_stream));
        if (gate->last) cudaEventDestroy(gate->last);
        gate->last = end;
    }
}

static void segment_prove(Ctx& c, zkgpu_upload& u, const uint64_t* public_values, size_t n_public_values, const zkgpu_kernel_labels* labels,
                          const zkstark::Config& cfg, const uint64_t* forced_pow_witnesses, volatile const int* abort_flag,
                          zkgpu_proof** proofs_out, uint64_t* ctl_challenges_out, uint64_t* trace_caps_out) {
    ZK_REQUIRE(u.ctx == &c, "the upload belongs to another context");
    ZK_REQUIRE(u.rate_bits == cfg.rate_bits && u.cap_height == cfg.cap_height, "the upload was made for another StarkConfig");
    ZK_REQUIRE(!u.consumed, "the upload was already consumed by a prove call (its table buffers are released as the tables are proved)");
    u.consumed = true;
    const zkstark::TableParams prm = params_from(labels);
    const size_t cap_words = (size_t)4 << cfg.cap_height;
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) proofs_out[t] = nullptr;

    // 1. trace commitments (prover.rs:92-116)
    StageLog lg(c);
    std::vector<uint64_t> caps(ZKGPU_NUM_TABLES * cap_words, 0);
    for (uint32_t t : u.order) {
        if (abort_flag && *abort_flag) throw ZkError(ZKGPU_ERR_ABORTED, "abort signal observed (prover.rs:346-354)");
        Batch& b = u.tb[t]->b;
        ZK_CUDA(cudaStreamWaitEvent(c.stream, u.uploaded[t], 0));
        lg.mark("trace upload");
        commit_from_device_values(c, b, true);
        lg.mark(zkstark::table_name(t));
        memcpy(&caps[t * cap_words], b.cap_host.data(), cap_words * 8);
    }
    if (trace_caps_out) memcpy(trace_caps_out, caps.data(), caps.size() * 8);

    // 2. transcript: caps, public values, CTL challenges (prover.rs:118-144)
    Challenger ch;
    uint64_t bg[4] = {0, 0, 0, 0};
    segment_transcript(ch, caps.data(), u.in_use, cap_words, public_values, n_public_values, cfg.num_challenges, bg);
    if (ctl_challenges_out) memcpy(ctl_challenges_out, bg, 2 * cfg.num_challenges * 8);
    ch.compact();
    uint64_t st[12];
    memcpy(st, ch.state, 96);

    // 3. per table: CTL data, then the proof, chained through the transcript state in Table order (prover.rs:251-259)
    std::unique_ptr<zkgpu_proof> proofs[ZKGPU_NUM_TABLES];
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) {
        if (!u.in_use[t]) continue;
        Ctl ctl;
        lg.mark("--");
        make_ctl_data(c, t, u.tb[t]->b, bg, cfg.num_challenges, ctl);
        lg.mark("ctl data");
        proofs[t].reset(new zkgpu_proof());
        prove_table(c, t, prm, cfg, u.tb[t]->b, ctl, st, forced_pow_witnesses ? &forced_pow_witnesses[t] : nullptr, abort_flag, proofs[t]->p);
        u.tb[t].reset();   // release the table's device memory before the next one
        lg.mark(zkstark::table_name(t));
    }
    for (uint32_t t = 0; t < ZKGPU_NUM_TABLES; t++) proofs_out[t] = proofs[t].release();
}

}  // namespace zk

extern "C" {

int zkgpu_segment_upload(zkgpu_ctx* h, const zkgpu_table_trace* traces, int mem_kind, const zkgpu_stark_config* config, zkgpu_upload** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && traces && config && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_upload> u(new zkgpu_upload());
    segment_upload(c, traces, mem_kind, config_from(config), *u);
    *out = u.release();
    ZK_API_END
}
void zkgpu_upload_free(zkgpu_upload* u) { delete u; }

int zkgpu_prove_segment_uploaded(zkgpu_ctx* h, zkgpu_upload* upload, const uint64_t* public_values, size_t n_public_values,
                                 const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config, const uint64_t* forced_pow_witnesses,
                                 volatile const int* abort_flag, zkgpu_proof** proofs_out, uint64_t* ctl_challenges_out, uint64_t* trace_caps_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && upload && config && proofs_out && (public_values || n_public_values == 0), "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    segment_prove(c, *upload, public_values, n_public_values, labels, config_from(config), forced_pow_witnesses, abort_flag, proofs_out,
                  ctl_challenges_out, trace_caps_out);
    ZK_API_END
}

int zkgpu_prove_segment(zkgpu_ctx* h, const zkgpu_table_trace* traces, int mem_kind, const uint64_t* public_values, size_t n_public_values,
                        const zkgpu_kernel_labels* labels, const zkgpu_stark_config* config, const uint64_t* forced_pow_witnesses,
                        volatile const int* abort_flag, zkgpu_proof** proofs_out, uint64_t* ctl_challenges_out, uint64_t* trace_caps_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && traces && config && proofs_out && (public_values || n_public_values == 0), "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    const zkstark::Config cfg = config_from(config);
    zkgpu_upload u;
    segment_upload(c, traces, mem_kind, cfg, u);
    segment_prove(c, u, public_values, n_public_values, labels, cfg, forced_pow_witnesses, abort_flag, proofs_out, ctl_challenges_out,
                  trace_caps_out);
    ZK_API_END
}

}  // extern "C"
