// Internal context / device-memory plumbing of libzkgpu (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include <memory>
#include "../../include/zkgpu.h"
#include "gl.cuh"

namespace zk {

struct ZkError : std::runtime_error {
    int code;
    ZkError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);

#define ZK_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw zk::ZkError(ZKGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +      \
                                                  " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)
#define ZK_REQUIRE(cond, msg)                                                   \
    do {                                                                        \
        if (!(cond)) throw zk::ZkError(ZKGPU_ERR_INVALID, std::string(msg));    \
    } while (0)

// wraps every ABI body: translate exceptions to status codes
#define ZK_API_BEGIN try {
#define ZK_API_END                                              \
    }                                                           \
    catch (const zk::ZkError& e) {                              \
        zk::set_last_error(e.what());                           \
        return e.code;                                          \
    }                                                           \
    catch (const std::bad_alloc&) {                             \
        zk::set_last_error("host allocation failed");           \
        return ZKGPU_ERR_NOMEM;                                 \
    }                                                           \
    catch (const std::exception& e) {                           \
        zk::set_last_error(e.what());                           \
        return ZKGPU_ERR_INVALID;                               \
    }                                                           \
    return ZKGPU_OK;

static inline unsigned log2_exact(size_t n) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) l++;
    if (((size_t)1 << l) != n) throw ZkError(ZKGPU_ERR_INVALID, "length must be a power of two");
    return l;
}

struct Ctx;
struct TableDev;

// stream-ordered device buffer
struct DevBuf {
    Ctx* ctx = nullptr;
    uint64_t* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(Ctx* c, size_t nbytes);
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); ctx = o.ctx; p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release();
    // a view of caller-owned device memory (never freed here)
    static DevBuf borrowed(const uint64_t* q, size_t nbytes) { DevBuf d; d.p = const_cast<uint64_t*>(q); d.bytes = nbytes; return d; }
    uint64_t* get() const { return p; }
};

// twiddle / scaling tables cached per transform size
struct NttTables {
    // roots_fwd[k] = w^k, roots_inv[k] = w^-k for w = primitive 2^ROOT_LOG-th root, k < 2^(ROOT_LOG-1)
    DevBuf roots_fwd, roots_inv;
};

// kernel families the in-library profiler accounts for (zkgpu_ctx_kernel_stats)
enum KernelFamily { KF_LEAF_HASH = 0, KF_MERKLE_LEVELS, KF_NTT, KF_QUOTIENT, KF_AUX, KF_OPENINGS, KF_FRI, KF_POW, KF_TRACE_GEN, KF_COUNT };

struct ProfRec { int fam; cudaEvent_t e0, e1; double bytes; uint64_t launches; };

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaMemPool_t pool = nullptr;
    int num_sms = 148;
    uint64_t launches = 0;
    uint64_t bytes_in_use = 0, bytes_peak = 0;
    NttTables ntt;
    // cache: key -> device table (inter-pass twiddles, coset power tables, lagrange selectors ...)
    std::map<std::string, DevBuf> table_cache;
    // per (table, num_challenges): lookup / CTL descriptors uploaded to the device (stark_dev.h)
    std::map<uint32_t, std::shared_ptr<TableDev>> stark_tables;
    // in-library profiler: CUDA events on `stream` around each launch group, with its algorithmic bytes
    bool profiling = false;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_free_events;
    double prof_ms[KF_COUNT] = {0}, prof_bytes[KF_COUNT] = {0};
    uint64_t prof_launches[KF_COUNT] = {0};
    void prof_collect();   // resolve finished event pairs (synchronises the stream)
    size_t evalpow_n = 0; uint64_t evalpow_zeta[2] = {0, 0};   // key of the cached zeta-power table (fri.cu eval_columns)
    // stage spans (zkgpu_ctx_set_timing): every StageLog mark records an event on `stream`; resolved by zkgpu_ctx_timing_report
    bool timing = false;
    std::vector<std::pair<std::string, cudaEvent_t>> timing_marks;
    // per table: the buffer of precomputed constraint values, kept from proof to proof (multi-GB: allocating it per proof makes the pool
    // grow and remap, measured as 0.1 - 0.7 s stalls, profiles/r2m)
    std::map<uint32_t, DevBuf> cons_cache;
    bool precompute = false;   // table jobs evaluate the alpha-independent constraint values in their first half (zkgpu_ctx_set_precompute_constraints)
    bool debug = false;   // proofs keep their aux / quotient batches and FRI input values for stage-by-stage parity tests
    // pinned staging buffer for H2D / D2H of pageable memory
    void* staging = nullptr;
    size_t staging_bytes = 0;

    void sync() { ZK_CUDA(cudaStreamSynchronize(stream)); }
    void count_launch(uint64_t k = 1) { launches += k; }
    void h2d(void* dst, const void* src, size_t bytes);
    void d2h(void* dst, const void* src, size_t bytes);   // synchronous on return
    void check_launch(const char* what);
};

// ZKGPU_TRACE=1: wall-clock stage log on stderr (synchronises at every mark; diagnosis only)
struct StageLog {
    Ctx& c; bool on; double t0;
    static double now();
    explicit StageLog(Ctx& c_);
    void mark(const char* what);
};

// RAII scope: times the launches issued inside it as one group of `fam` (no-op unless profiling is on)
struct KernelScope {
    Ctx& c; int fam; double bytes; uint64_t l0; cudaEvent_t e0 = nullptr;
    KernelScope(Ctx& c_, int fam_, double algorithmic_bytes) : c(c_), fam(fam_), bytes(algorithmic_bytes), l0(c_.launches) {
        if (!c.profiling) return;
        auto get = [&]() { cudaEvent_t e; if (!c.prof_free_events.empty()) { e = c.prof_free_events.back(); c.prof_free_events.pop_back(); } else cudaEventCreate(&e); return e; };
        e0 = get();
        cudaEventRecord(e0, c.stream);
    }
    ~KernelScope() {
        if (!e0) return;
        cudaEvent_t e1; if (!c.prof_free_events.empty()) { e1 = c.prof_free_events.back(); c.prof_free_events.pop_back(); } else cudaEventCreate(&e1);
        cudaEventRecord(e1, c.stream);
        c.prof_pending.push_back({fam, e0, e1, bytes, c.launches - l0});
    }
};

// ---- PolynomialBatch on the device ----------------------------------------------------------------------
// Layout (DESIGN.md "data layout"): everything column-major, 8-byte elements.
//   values  ncols x n      trace values, natural row order (optional)
//   coeffs  ncols x n      polynomial coefficients, natural order
//   lde     ncols x N      N = n << rate_bits; evaluations on the coset g*<w_N> stored in BIT-REVERSED index order,
//                          i.e. lde[c*N + j] = poly_c(g * w_N^bitrev(j)) == plonky2 merkle_tree.leaves[j][c]
//   digests level 0 = N leaf digests (4 u64 each), level l+1 = N >> (l+1) nodes, ..., last level = the cap
struct Batch {
    Ctx* ctx = nullptr;
    size_t ncols = 0, n = 0, N = 0;
    unsigned log_n = 0, rate_bits = 0, cap_height = 0;
    DevBuf values, coeffs, lde, digests;
    std::vector<size_t> level_off;   // offset (in u64) of each digest level inside `digests`
    std::vector<size_t> level_cnt;   // number of digests per level
    std::vector<uint64_t> cap_host;  // (1<<cap_height)*4
    const uint64_t* cap_dev() const { return digests.get() + level_off.back(); }
};

}  // namespace zk

struct zkgpu_ctx { zk::Ctx c; };
struct zkgpu_batch { zk::Batch b; };
