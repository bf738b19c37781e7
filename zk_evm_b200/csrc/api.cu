// libzkgpu C ABI: context, device memory, commitments (S1) and the fine-grained kernel entry points.
#include "internal.h"
#include "ntt.h"
#include "merkle.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <functional>
#include <memory>

namespace zk {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }

DevBuf::DevBuf(Ctx* c, size_t nbytes) : ctx(c), bytes(nbytes ? nbytes : 8) {
    void* q = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&q, bytes, c->pool, c->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw ZkError(e == cudaErrorMemoryAllocation ? ZKGPU_ERR_NOMEM : ZKGPU_ERR_CUDA,
                      std::string("cudaMallocAsync(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    p = (uint64_t*)q;
    c->bytes_in_use += bytes;
    if (c->bytes_in_use > c->bytes_peak) c->bytes_peak = c->bytes_in_use;
}
void DevBuf::release() {
    if (p && !ctx) { p = nullptr; bytes = 0; return; }     // borrowed view
    if (p) {
        cudaFreeAsync(p, ctx->stream);
        ctx->bytes_in_use -= bytes;
        p = nullptr;
        bytes = 0;
    }
}

// Small transfers between pageable host memory and the device go through a pinned staging buffer of the context: a pageable
// cudaMemcpyAsync is staged by the driver and ordered against every other copy in flight, so the proof's many small read-backs
// (caps, openings, PoW results) would queue behind the multi-GB trace uploads of the next segment.
static void* ctx_staging(Ctx& c, size_t bytes) {
    if (bytes > c.staging_bytes) {
        if (c.staging) { cudaStreamSynchronize(c.stream); cudaFreeHost(c.staging); c.staging = nullptr; c.staging_bytes = 0; }
        size_t want = bytes < ((size_t)1 << 20) ? ((size_t)1 << 20) : bytes;
        ZK_CUDA(cudaHostAlloc(&c.staging, want, cudaHostAllocDefault));
        c.staging_bytes = want;
    }
    return c.staging;
}
static constexpr size_t STAGING_MAX = (size_t)64 << 20;

void Ctx::h2d(void* dst, const void* src, size_t bytes) {
    if (bytes && bytes <= STAGING_MAX) {
        // the staging buffer is reused by the next small transfer: wait for earlier users, then copy through it
        ZK_CUDA(cudaStreamSynchronize(stream));
        void* st = ctx_staging(*this, bytes);
        memcpy(st, src, bytes);
        ZK_CUDA(cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, stream));
        ZK_CUDA(cudaStreamSynchronize(stream));
        return;
    }
    ZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
}
void Ctx::d2h(void* dst, const void* src, size_t bytes) {
    if (bytes && bytes <= STAGING_MAX) {
        void* st = ctx_staging(*this, bytes);
        ZK_CUDA(cudaMemcpyAsync(st, src, bytes, cudaMemcpyDeviceToHost, stream));
        ZK_CUDA(cudaStreamSynchronize(stream));
        memcpy(dst, st, bytes);
        return;
    }
    ZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    ZK_CUDA(cudaStreamSynchronize(stream));
}
double StageLog::now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static void timing_mark(Ctx& c, const char* what) {
    // bounded: a caller that never asks for the report does not grow the list for ever (the oldest spans are dropped)
    if (c.timing_marks.size() >= 65536) {
        for (size_t i = 0; i < 32768; i++) cudaEventDestroy(c.timing_marks[i].second);
        c.timing_marks.erase(c.timing_marks.begin(), c.timing_marks.begin() + 32768);
    }
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c.stream);
    c.timing_marks.emplace_back(what, e);
}
StageLog::StageLog(Ctx& c_) : c(c_) {
    const char* e = getenv("ZKGPU_TRACE"); on = e && *e == '1'; t0 = on ? now() : 0;
    if (c.timing) timing_mark(c, "");          // an empty name starts a chain of spans
}
void StageLog::mark(const char* what) {
    if (c.timing) timing_mark(c, what);
    if (!on) return;
    double a = now();
    cudaStreamSynchronize(c.stream);
    double b = now();
    fprintf(stderr, "[zkgpu] %-28s host %8.3f ms  +drain %8.3f ms\n", what, a - t0, b - a);
    t0 = b;
}
void Ctx::prof_collect() {
    if (prof_pending.empty()) return;
    ZK_CUDA(cudaStreamSynchronize(stream));
    for (ProfRec& r : prof_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { prof_ms[r.fam] += ms; prof_bytes[r.fam] += r.bytes; prof_launches[r.fam] += r.launches; }
        prof_free_events.push_back(r.e0); prof_free_events.push_back(r.e1);
    }
    prof_pending.clear();
}
void Ctx::check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw ZkError(ZKGPU_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ---- commitments ---------------------------------------------------------------------------------------------
// coefficients already in b.coeffs: LDE + Merkle tree
static void finish_commit(Ctx& c, Batch& b) {
    lde_bitrev(c, b.coeffs.get(), b.lde.get(), b.ncols, b.log_n, b.rate_bits, GL_GENERATOR);
    merkle_build(c, b.lde.get(), b.N, b.ncols, b.N, b.cap_height, b.digests, b.level_off, b.level_cnt);
    b.cap_host.resize(4 * b.level_cnt.back());
    c.d2h(b.cap_host.data(), b.cap_dev(), b.cap_host.size() * 8);
}

void init_batch(Ctx& c, Batch& b, size_t ncols, size_t n, uint32_t rate_bits, uint32_t cap_height) {
    ZK_REQUIRE(ncols > 0 && n > 0, "empty batch");
    b.ctx = &c;
    b.ncols = ncols;
    b.n = n;
    b.log_n = log2_exact(n);
    b.rate_bits = rate_bits;
    b.cap_height = cap_height;
    b.N = n << rate_bits;
    ZK_REQUIRE(((size_t)1 << cap_height) <= b.N, "cap_height too large for the LDE size");
}

// values on the device (b.values filled, natural order) -> full batch
void commit_from_device_values(Ctx& c, Batch& b, bool keep_values) {
    b.coeffs = DevBuf(&c, b.ncols * b.n * 8);
    b.lde = DevBuf(&c, b.ncols * b.N * 8);
    // the LDE buffer doubles as scratch for the bit-reversed intermediate of the inverse transform
    {
        DevBuf scratch;
        uint64_t* s = b.lde.get();
        if (b.rate_bits == 0) { scratch = DevBuf(&c, b.ncols * b.n * 8); s = scratch.get(); }
        intt_natural(c, b.values.get(), s, b.coeffs.get(), b.ncols, b.log_n, 0);
    }
    if (!keep_values) b.values.release();
    finish_commit(c, b);
}

void commit_from_device_coeffs(Ctx& c, Batch& b) {
    b.lde = DevBuf(&c, b.ncols * b.N * 8);
    finish_commit(c, b);
}

static void upload_cols(Ctx& c, uint64_t* dst, const uint64_t* const* cols, size_t ncols, size_t n, int mem_kind) {
    cudaMemcpyKind k = mem_kind == ZKGPU_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    for (size_t i = 0; i < ncols; i++) {
        ZK_REQUIRE(cols[i] != nullptr, "null column pointer");
        ZK_CUDA(cudaMemcpyAsync(dst + i * n, cols[i], n * 8, k, c.stream));
    }
}

}  // namespace zk

using namespace zk;

extern "C" {

const char* zkgpu_last_error(void) { return g_last_error.c_str(); }
const char* zkgpu_version(void) { return "zkgpu 0.1 (sm_100a)"; }

int zkgpu_ctx_create(int device, zkgpu_ctx** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(out != nullptr, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw ZkError(ZKGPU_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                                          " (libzkgpu has no CPU path)");
    ZK_REQUIRE(device >= 0 && device < count, "device index out of range");
    ZK_CUDA(cudaSetDevice(device));
    {
        // Keep the device's local-memory (stack / spill) backing at its high-water mark.  The Arithmetic quotient kernel needs 1.8 KB
        // of stack per thread, the others ~0.5 KB; by default the driver shrinks the backing store after the big kernel and grows
        // it again at its next launch, which needs the whole context idle — with two segments in flight that showed up as sporadic
        // stalls of hundreds of milliseconds (profiles/r1q: 559 vs 852 ms for the same step).
        unsigned flags = 0;
        if (cudaGetDeviceFlags(&flags) == cudaSuccess && !(flags & cudaDeviceLmemResizeToMax))
            if (cudaSetDeviceFlags(flags | cudaDeviceLmemResizeToMax) != cudaSuccess) cudaGetLastError();   // older drivers: best effort
    }
    zkgpu_ctx* h = new zkgpu_ctx();
    Ctx& c = h->c;
    c.device = device;
    ZK_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    ZK_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    ZK_CUDA(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, device));
    // One memory pool PER CONTEXT, never trimmed.  With the device's default pool shared by several contexts (segments in flight),
    // a block freed stream-ordered by one context can be handed to another whose work then waits for the first stream to reach
    // that free — an invisible cross-stream dependency that made two-stream steps take anything from 550 to 1350 ms
    // (profiles/r1s).  Own pools: allocations and frees of a context are ordered on its one stream only.
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        ZK_CUDA(cudaMemPoolCreate(&c.pool, &props));
        uint64_t thresh = UINT64_MAX;
        ZK_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    }
    *out = h;
    ZK_API_END
}

void zkgpu_ctx_destroy(zkgpu_ctx* h) {
    if (!h) return;
    cudaSetDevice(h->c.device);
    cudaStreamSynchronize(h->c.stream);
    h->c.table_cache.clear();
    h->c.cons_cache.clear();
    h->c.stark_tables.clear();
    h->c.ntt.roots_fwd.release();
    h->c.ntt.roots_inv.release();
    cudaStreamSynchronize(h->c.stream);
    if (h->c.ev0) { cudaEventDestroy(h->c.ev0); cudaEventDestroy(h->c.ev1); }
    for (auto& m : h->c.timing_marks) cudaEventDestroy(m.second);
    h->c.timing_marks.clear();
    for (auto& r : h->c.prof_pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (cudaEvent_t e : h->c.prof_free_events) cudaEventDestroy(e);
    cudaStreamDestroy(h->c.stream);
    cudaStreamDestroy(h->c.copy_stream);
    if (h->c.pool) cudaMemPoolDestroy(h->c.pool);
    if (h->c.staging) cudaFreeHost(h->c.staging);
    delete h;
}

int zkgpu_ctx_stream(zkgpu_ctx* h, void** stream_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && stream_out, "null argument");
    *stream_out = (void*)h->c.stream;
    ZK_API_END
}

int zkgpu_ctx_sync(zkgpu_ctx* h) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "ctx is null");
    h->c.sync();
    ZK_API_END
}

// CUDA-event timer on the context's stream (bench.py: device time of a step, events on the launching stream)
int zkgpu_ctx_timer_start(zkgpu_ctx* h) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "ctx is null");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    if (!c.ev0) { ZK_CUDA(cudaEventCreate(&c.ev0)); ZK_CUDA(cudaEventCreate(&c.ev1)); }
    ZK_CUDA(cudaEventRecord(c.ev0, c.stream));
    ZK_API_END
}
int zkgpu_ctx_timer_stop(zkgpu_ctx* h, float* ms) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && ms && h->c.ev0, "timer not started");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    ZK_CUDA(cudaEventRecord(c.ev1, c.stream));
    ZK_CUDA(cudaEventSynchronize(c.ev1));
    ZK_CUDA(cudaEventElapsedTime(ms, c.ev0, c.ev1));
    ZK_API_END
}

int zkgpu_ctx_stats(zkgpu_ctx* h, uint64_t* kernel_launches, uint64_t* bytes_in_use, uint64_t* bytes_peak) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "ctx is null");
    if (kernel_launches) *kernel_launches = h->c.launches;
    if (bytes_in_use) *bytes_in_use = h->c.bytes_in_use;
    if (bytes_peak) *bytes_peak = h->c.bytes_peak;
    ZK_API_END
}

int zkgpu_ctx_set_profiling(zkgpu_ctx* h, int on) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "ctx is null");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    c.prof_collect();
    c.profiling = on != 0;
    if (on) for (int i = 0; i < KF_COUNT; i++) { c.prof_ms[i] = 0; c.prof_bytes[i] = 0; c.prof_launches[i] = 0; }
    ZK_API_END
}
int zkgpu_ctx_set_timing(zkgpu_ctx* h, int on) {
    ZK_API_BEGIN
    ZK_REQUIRE(h, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    for (auto& m : c.timing_marks) cudaEventDestroy(m.second);
    c.timing_marks.clear();
    c.timing = on != 0;
    ZK_API_END
}
int zkgpu_ctx_timing_report(zkgpu_ctx* h, char* buf, size_t* len) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && len, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    ZK_CUDA(cudaStreamSynchronize(c.stream));
    std::string out;
    for (size_t i = 0; i < c.timing_marks.size(); i++) {
        const auto& m = c.timing_marks[i];
        if (m.first.empty() || i == 0) continue;                 // chain start
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c.timing_marks[i - 1].second, m.second) != cudaSuccess) ms = 0;
        char line[160];
        snprintf(line, sizeof line, "%s\t%.6f\n", m.first.c_str(), ms);
        out += line;
    }
    const size_t need = out.size() + 1;
    if (buf && *len >= need) memcpy(buf, out.c_str(), need);
    *len = need;
    ZK_API_END
}
int zkgpu_ctx_kernel_stats(zkgpu_ctx* h, uint32_t family, uint64_t* launches, double* ms_total, double* algorithmic_bytes) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && family < KF_COUNT, "bad argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    c.prof_collect();
    if (launches) *launches = c.prof_launches[family];
    if (ms_total) *ms_total = c.prof_ms[family];
    if (algorithmic_bytes) *algorithmic_bytes = c.prof_bytes[family];
    ZK_API_END
}

int zkgpu_commit_values(zkgpu_ctx* h, const uint64_t* const* cols, size_t ncols, size_t n, uint32_t rate_bits,
                        uint32_t cap_height, int mem_kind, int keep_values, zkgpu_batch** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && cols && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_batch> hb(new zkgpu_batch());
    Batch& b = hb->b;
    init_batch(c, b, ncols, n, rate_bits, cap_height);
    b.values = DevBuf(&c, ncols * n * 8);
    upload_cols(c, b.values.get(), cols, ncols, n, mem_kind);
    commit_from_device_values(c, b, keep_values != 0);
    *out = hb.release();
    ZK_API_END
}

int zkgpu_commit_values_contig(zkgpu_ctx* h, const uint64_t* base, size_t ncols, size_t n, uint32_t rate_bits,
                               uint32_t cap_height, int mem_kind, int keep_values, zkgpu_batch** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && base && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_batch> hb(new zkgpu_batch());
    Batch& b = hb->b;
    init_batch(c, b, ncols, n, rate_bits, cap_height);
    b.values = DevBuf(&c, ncols * n * 8);
    ZK_CUDA(cudaMemcpyAsync(b.values.get(), base, ncols * n * 8,
                            mem_kind == ZKGPU_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c.stream));
    commit_from_device_values(c, b, keep_values != 0);
    *out = hb.release();
    ZK_API_END
}

int zkgpu_commit_coeffs(zkgpu_ctx* h, const uint64_t* const* coeff_cols, size_t ncols, size_t n, uint32_t rate_bits,
                        uint32_t cap_height, int mem_kind, zkgpu_batch** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && coeff_cols && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_batch> hb(new zkgpu_batch());
    Batch& b = hb->b;
    init_batch(c, b, ncols, n, rate_bits, cap_height);
    b.coeffs = DevBuf(&c, ncols * n * 8);
    upload_cols(c, b.coeffs.get(), coeff_cols, ncols, n, mem_kind);
    commit_from_device_coeffs(c, b);
    *out = hb.release();
    ZK_API_END
}

void zkgpu_batch_free(zkgpu_batch* b) {
    if (!b) return;
    if (b->b.ctx) cudaSetDevice(b->b.ctx->device);
    delete b;
}

int zkgpu_batch_dims(const zkgpu_batch* b, size_t* ncols, size_t* n, uint32_t* rate_bits, uint32_t* cap_height) {
    ZK_API_BEGIN
    ZK_REQUIRE(b, "batch is null");
    if (ncols) *ncols = b->b.ncols;
    if (n) *n = b->b.n;
    if (rate_bits) *rate_bits = b->b.rate_bits;
    if (cap_height) *cap_height = b->b.cap_height;
    ZK_API_END
}

int zkgpu_batch_cap(const zkgpu_batch* b, uint64_t* out_cap) {
    ZK_API_BEGIN
    ZK_REQUIRE(b && out_cap, "null argument");
    memcpy(out_cap, b->b.cap_host.data(), b->b.cap_host.size() * 8);
    ZK_API_END
}

int zkgpu_batch_export(const zkgpu_batch* hb, uint64_t* coeffs, uint64_t* leaves, uint64_t* digests) {
    ZK_API_BEGIN
    ZK_REQUIRE(hb, "batch is null");
    const Batch& b = hb->b;
    Ctx& c = *b.ctx;
    ZK_CUDA(cudaSetDevice(c.device));
    if (coeffs) c.d2h(coeffs, b.coeffs.get(), b.ncols * b.n * 8);
    if (leaves) {
        // device layout is column-major; the host PolynomialBatch wants row-major leaves
        std::vector<uint64_t> tmp(b.ncols * b.N);
        c.d2h(tmp.data(), b.lde.get(), tmp.size() * 8);
        for (size_t col = 0; col < b.ncols; col++)
            for (size_t j = 0; j < b.N; j++) leaves[j * b.ncols + col] = tmp[col * b.N + j];
    }
    if (digests) {
        size_t total = b.level_off.back() + 4 * b.level_cnt.back();
        std::vector<uint64_t> lv(total);
        c.d2h(lv.data(), b.digests.get(), total * 8);
        merkle_export_plonky2(lv, b.level_off, b.level_cnt, digests);
    }
    ZK_API_END
}

// ---- fine-grained entry points ---------------------------------------------------------------------------
int zkgpu_ntt(zkgpu_ctx* h, uint64_t* data, size_t ncols, size_t n, int inverse, uint64_t coset_shift) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && data, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    unsigned L = log2_exact(n);
    DevBuf a(&c, ncols * n * 8), w(&c, ncols * n * 8), o(&c, ncols * n * 8);
    c.h2d(a.get(), data, ncols * n * 8);
    if (inverse) {
        intt_natural(c, a.get(), w.get(), o.get(), ncols, L, coset_shift > 1 ? gl_canon(coset_shift) : 0);
    } else {
        const uint64_t* pre = coset_shift > 1 ? get_power_table(c, gl_canon(coset_shift), 1, n) : nullptr;
        ntt_dif(c, a.get(), n, 0, w.get(), n, ncols, L, false, pre, nullptr, 0);
        bitrev_permute(c, w.get(), n, o.get(), n, ncols, L, 1, nullptr);
    }
    c.d2h(data, o.get(), ncols * n * 8);
    ZK_API_END
}

int zkgpu_poseidon_permute(zkgpu_ctx* h, uint64_t* states, size_t count) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && states, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    if (count == 0) return ZKGPU_OK;
    DevBuf d(&c, count * 96);
    c.h2d(d.get(), states, count * 96);
    poseidon_states(c, d.get(), count);
    c.d2h(states, d.get(), count * 96);
    ZK_API_END
}

int zkgpu_poseidon_hash_rows(zkgpu_ctx* h, const uint64_t* data_colmajor, size_t nrows, size_t width, uint64_t* out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && data_colmajor && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    if (nrows == 0) return ZKGPU_OK;
    DevBuf d(&c, nrows * width * 8), o(&c, nrows * 32);
    c.h2d(d.get(), data_colmajor, nrows * width * 8);
    leaf_hash(c, d.get(), nrows, width, nrows, o.get());
    c.d2h(out, o.get(), nrows * 32);
    ZK_API_END
}

__global__ void fill_kernel(uint64_t* p, size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t z = seed + i * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        p[i] = zk::gl_canon(z);
    }
}

static float time_loop(Ctx& c, int iters, const std::function<void()>& body) {
    cudaEvent_t e0, e1;
    ZK_CUDA(cudaEventCreate(&e0));
    ZK_CUDA(cudaEventCreate(&e1));
    body();   // warm-up (also builds cached tables)
    c.sync();
    ZK_CUDA(cudaEventRecord(e0, c.stream));
    for (int i = 0; i < iters; i++) body();
    ZK_CUDA(cudaEventRecord(e1, c.stream));
    ZK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / iters;
}

int zkgpu_bench_ntt(zkgpu_ctx* h, size_t ncols, size_t n, int iters, float* ms_per_iter, uint64_t* launches) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && ms_per_iter && iters > 0, "bad argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    unsigned L = log2_exact(n);
    DevBuf a(&c, ncols * n * 8), w(&c, ncols * n * 8);
    fill_kernel<<<1024, 256, 0, c.stream>>>(a.get(), ncols * n, 1);
    uint64_t l0 = 0;
    *ms_per_iter = time_loop(c, iters, [&]() {
        l0 = c.launches;
        ntt_dif(c, a.get(), n, 0, w.get(), n, ncols, L, false, nullptr, nullptr, 0);
    });
    if (launches) *launches = c.launches - l0;
    ZK_API_END
}

int zkgpu_bench_leaf_hash(zkgpu_ctx* h, size_t ncols, size_t nrows, int iters, float* ms_per_iter) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && ms_per_iter && iters > 0, "bad argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    DevBuf a(&c, ncols * nrows * 8), o(&c, nrows * 32);
    fill_kernel<<<1024, 256, 0, c.stream>>>(a.get(), ncols * nrows, 2);
    *ms_per_iter = time_loop(c, iters, [&]() { leaf_hash(c, a.get(), nrows, ncols, nrows, o.get()); });
    ZK_API_END
}

int zkgpu_bench_merkle_levels(zkgpu_ctx* h, size_t nleaves, int iters, float* ms_per_iter) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && ms_per_iter && iters > 0, "bad argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::vector<size_t> off, cnt;
    merkle_layout(nleaves, 4, off, cnt);
    DevBuf d(&c, (off.back() + 4 * cnt.back()) * 8);
    fill_kernel<<<1024, 256, 0, c.stream>>>(d.get(), 4 * nleaves, 3);
    *ms_per_iter = time_loop(c, iters, [&]() { merkle_inner_levels(c, d.get(), off, cnt); });
    ZK_API_END
}

}  // extern "C"
