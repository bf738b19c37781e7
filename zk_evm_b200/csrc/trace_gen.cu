// GPU-side trace finishing: the data-parallel tail of trace generation runs on the device, so that the widest trace of a
// segment never crosses PCIe — KeccakStark is 2431 columns (2.5 GB at 2^17 rows, 65 % of a config-#4 segment's upload) generated from
// 208 bytes per permutation.
//
// Replaces /root/reference/evm_arithmetization/src/keccak/keccak_stark.rs:70-250 (KeccakStark::generate_trace: generate_trace_rows +
// trace_rows_to_poly_values), called from witness/traces.rs:243-258 (`into_tables`).  The row code is stark/keccak_trace.h (host + device).
//
// and src/logic.rs:165-237 (LogicStark::generate_trace: 523 columns from 72 bytes per operation; stark/logic_trace.h).
//
// and arithmetic/arithmetic_stark.rs:130-156 (ArithmeticStark::generate_range_checks: the RANGE_COUNTER column and the RC_FREQUENCIES histogram
// of the 96 shared columns, in place on a device-resident trace).
//
// and memory/memory_stark.rs:104-213,240-294,387-404 (MemoryStark::generate_trace after the host's sort / fill_gaps / padding: 16 of the 30
// columns derived on the device from the 14 operation columns; stark/memory_trace.h).
//
// keccak_trace_kernel: one thread per trace ROW.  Rows are independent given the permutation's input (keccak_trace.h), the trace is
// column-major, so a warp stores 32 consecutive rows of one column per instruction: 256 contiguous bytes.  Write-bound: 8 * 2431 bytes per
// row against ~1.5 k word operations; algorithmic bytes = 8 * 2431 * n written + 208 * num_perms read.
#include "internal.h"
#include "stark/keccak_trace.h"
#include "stark/logic_trace.h"
#include "stark/table_arithmetic.h"
#include "stark/memory_trace.h"

namespace zk {

struct ColStore {
    uint64_t* out; size_t n;     // out already points at the thread's row
    __device__ __forceinline__ void operator()(uint32_t col, uint64_t v) const { out[(size_t)col * n] = v; }
};

__global__ void __launch_bounds__(128) keccak_trace_kernel(const uint64_t* __restrict__ inputs, const uint64_t* __restrict__ timestamps,
                                                           uint64_t num_perms, size_t n, uint64_t* __restrict__ out) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    ColStore st{out + row, n};
    zkstark::keccak::generate_row(inputs, timestamps, num_perms, row, st);
}

// logic_trace_kernel: one thread per row as well (each row is one operation); 8 * 523 bytes written per row
__global__ void __launch_bounds__(128) logic_trace_kernel(const uint64_t* __restrict__ ops, uint64_t num_ops, size_t n, uint64_t* __restrict__ out) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    ColStore st{out + row, n};
    zkstark::logic::generate_row(ops, num_ops, row, st);
}

// ArithmeticStark::generate_range_checks.  RANGE_MAX = 2^16 (arithmetic_stark.rs:126).
static constexpr uint64_t ARITH_RANGE_MAX = 1ull << 16;
__global__ void arith_rc_init_kernel(uint64_t* __restrict__ counter, uint64_t* __restrict__ freq, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    counter[i] = i < ARITH_RANGE_MAX ? i : ARITH_RANGE_MAX - 1;     // :136-141
    freq[i] = 0;
}
// Histogram of the shared columns into freq[x] (:144-155).  A thread walks its cells with a grid stride (consecutive threads read
// consecutive rows of a column) and counts ZERO cells privately — padding rows and unused registers make zero the bin nearly every
// cell of a sparse trace falls into — adding that count once at the end; the other values go to the L2 as reductions (red.add.u64;
// a count never reaches p, so the field addition of the reference is a plain integer addition).  A cell >= 2^16 raises *bad (the
// reference asserts).
__global__ void __launch_bounds__(256) arith_rc_hist_kernel(const uint64_t* __restrict__ shared_cols, size_t cells, uint64_t* __restrict__ freq,
                                                            unsigned* __restrict__ bad) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long zeros = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        const uint64_t x = shared_cols[i];
        if (x == 0) zeros++;
        else if (x < ARITH_RANGE_MAX) atomicAdd((unsigned long long*)&freq[x], 1ull);
        else atomicOr(bad, 1u);
    }
    if (zeros) atomicAdd((unsigned long long*)&freq[0], zeros);
}

// MemoryStark: the stale-context list into its two columns (insert_stale_contexts), then one thread per row (memory_trace.h) with the
// two histogram columns accumulated by reductions into the L2
__global__ void memory_stale_kernel(const uint64_t* __restrict__ stale, size_t num_stale, size_t n, uint64_t* __restrict__ t) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= num_stale) return;
    const uint64_t ctx = stale[k];
    t[(size_t)zkstark::memory::STALE_CONTEXTS * n + ctx] = ctx + 1;
    t[(size_t)zkstark::memory::IS_PRUNED * n + ctx] = 1;
}
__global__ void __launch_bounds__(256) memory_finish_kernel(uint64_t* __restrict__ t, size_t n, unsigned* __restrict__ bad) {
    namespace mem = zkstark::memory;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mem::RowCounts rc;
    if (!mem::finish_row(t, n, i, rc)) { atomicOr(bad, 1u); return; }
    unsigned long long* freq = (unsigned long long*)(t + (size_t)mem::FREQUENCIES * n);
    atomicAdd(&freq[rc.freq_a], 1ull);
    if (rc.freq_b != mem::NONE) atomicAdd(&freq[rc.freq_b], 1ull);
    if (rc.stale_ctx != mem::NONE) atomicAdd((unsigned long long*)(t + (size_t)mem::STALE_CONTEXT_FREQUENCIES * n) + rc.stale_ctx, 1ull);
}

static size_t padded_rows(size_t rows, size_t min_rows) {
    ZK_REQUIRE(rows <= ((size_t)1 << 40) && min_rows <= ((size_t)1 << 40), "too many rows");
    size_t want = rows < min_rows ? min_rows : rows, n = 1;
    while (n < want) n <<= 1;
    return n;
}

}  // namespace zk

struct zkgpu_dev_trace {
    zk::DevBuf buf;
    size_t ncols = 0, n = 0;
};

extern "C" {

int zkgpu_keccak_generate_trace(zkgpu_ctx* h, const uint64_t* inputs, const uint64_t* timestamps, size_t num_perms, size_t min_rows,
                                zkgpu_dev_trace** out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(h && out && ((inputs && timestamps) || num_perms == 0), "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    // num_rows = max(24 * len, min_rows).next_power_of_two()   (keccak_stark.rs:76-78)
    ZK_REQUIRE(num_perms <= ((size_t)1 << 40), "too many rows");
    const size_t n = padded_rows(num_perms * zkstark::keccak::NUM_ROUNDS, min_rows);
    std::unique_ptr<zkgpu_dev_trace> t(new zkgpu_dev_trace());
    t->ncols = zkstark::keccak::NUM_COLUMNS;
    t->n = n;
    t->buf = DevBuf(&c, t->ncols * n * 8);
    DevBuf in(&c, (num_perms ? num_perms : 1) * 26 * 8);      // 25 input words per permutation, then the timestamps
    if (num_perms) {
        c.h2d(in.get(), inputs, num_perms * 25 * 8);
        c.h2d(in.get() + num_perms * 25, timestamps, num_perms * 8);
    }
    {
        KernelScope ks(c, KF_TRACE_GEN, 8.0 * (double)t->ncols * (double)n + 208.0 * (double)num_perms);
        const unsigned T = 128;
        keccak_trace_kernel<<<(unsigned)((n + T - 1) / T), T, 0, c.stream>>>(in.get(), in.get() + num_perms * 25, num_perms, n, t->buf.get());
        c.count_launch();
        c.check_launch("keccak_trace_kernel");
    }
    // `in` is released in stream order (after the kernel); pageable inputs were staged synchronously by h2d
    *out = t.release();
    ZK_API_END
}

int zkgpu_logic_generate_trace(zkgpu_ctx* h, const uint64_t* ops, size_t num_ops, size_t min_rows, zkgpu_dev_trace** out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(h && out && (ops || num_ops == 0), "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    const size_t W = zkstark::logic::OP_WORDS;
    for (size_t i = 0; i < num_ops; i++) ZK_REQUIRE(ops[i * W] <= 2, "operator must be 0 (AND), 1 (OR) or 2 (XOR)");
    // padded_len = len.max(min_rows).next_power_of_two()   (logic.rs:218-220)
    const size_t n = padded_rows(num_ops, min_rows);
    std::unique_ptr<zkgpu_dev_trace> t(new zkgpu_dev_trace());
    t->ncols = zkstark::logic::NUM_COLUMNS;
    t->n = n;
    t->buf = DevBuf(&c, t->ncols * n * 8);
    DevBuf in(&c, (num_ops ? num_ops : 1) * W * 8);
    if (num_ops) c.h2d(in.get(), ops, num_ops * W * 8);
    {
        KernelScope ks(c, KF_TRACE_GEN, 8.0 * (double)t->ncols * (double)n + 8.0 * W * (double)num_ops);
        const unsigned T = 128;
        logic_trace_kernel<<<(unsigned)((n + T - 1) / T), T, 0, c.stream>>>(in.get(), num_ops, n, t->buf.get());
        c.count_launch();
        c.check_launch("logic_trace_kernel");
    }
    *out = t.release();
    ZK_API_END
}

int zkgpu_dev_trace_upload(zkgpu_ctx* h, const uint64_t* cols, size_t ncols, size_t n, zkgpu_dev_trace** out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(h && cols && out && ncols && n, "null argument");
    log2_exact(n);
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_dev_trace> t(new zkgpu_dev_trace());
    t->ncols = ncols;
    t->n = n;
    t->buf = DevBuf(&c, ncols * n * 8);
    c.h2d(t->buf.get(), cols, ncols * n * 8);
    *out = t.release();
    ZK_API_END
}

int zkgpu_arithmetic_generate_range_checks(zkgpu_ctx* h, zkgpu_dev_trace* t) {
    ZK_API_BEGIN
    using namespace zk;
    namespace ar = zkstark::arithmetic;
    ZK_REQUIRE(h && t, "null argument");
    Ctx& c = h->c;
    ZK_REQUIRE(t->buf.ctx == &c, "the trace belongs to another context");
    ZK_REQUIRE(t->ncols == ar::NUM_COLUMNS, "not an Arithmetic trace (116 columns)");
    ZK_REQUIRE(t->n >= ARITH_RANGE_MAX, "the Arithmetic trace needs at least 2^16 rows (arithmetic_stark.rs:171-180)");
    ZK_CUDA(cudaSetDevice(c.device));
    const size_t n = t->n, cells = (size_t)ar::NUM_SHARED_COLS * n;
    uint64_t* counter = t->buf.get() + (size_t)ar::RANGE_COUNTER * n;
    uint64_t* freq = t->buf.get() + (size_t)ar::RC_FREQUENCIES * n;
    DevBuf bad(&c, 8);
    ZK_CUDA(cudaMemsetAsync(bad.get(), 0, 8, c.stream));
    {
        KernelScope ks(c, KF_TRACE_GEN, 8.0 * (double)(cells + 2 * n));
        arith_rc_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(counter, freq, n);
        // the SHARED_COLS are one contiguous block of the column-major trace
        const size_t want = (cells + 256 * 64 - 1) / (256 * 64);
        const unsigned blocks = (unsigned)(want < 1 ? 1 : want > (size_t)c.num_sms * 8 ? (size_t)c.num_sms * 8 : want);
        arith_rc_hist_kernel<<<blocks, 256, 0, c.stream>>>(t->buf.get() + (size_t)ar::START_SHARED_COLS * n, cells, freq, (unsigned*)bad.get());
        c.count_launch(2);
        c.check_launch("arith_rc kernels");
    }
    uint64_t flag = 0;
    c.d2h(&flag, bad.get(), 8);
    ZK_REQUIRE((flag & 0xFFFFFFFFull) == 0, "a shared column value exceeds the max range value 65536 (arithmetic_stark.rs:147-152)");
    ZK_API_END
}

int zkgpu_memory_finish_trace(zkgpu_ctx* h, const uint64_t* ops, size_t n, const uint64_t* stale_contexts, size_t num_stale,
                              zkgpu_dev_trace** out) {
    ZK_API_BEGIN
    using namespace zk;
    namespace mem = zkstark::memory;
    ZK_REQUIRE(h && ops && out && n && (stale_contexts || num_stale == 0), "null argument");
    log2_exact(n);
    for (size_t k = 0; k < num_stale; k++) ZK_REQUIRE(stale_contexts[k] < n, "stale context past the end of the trace (memory_stark.rs:401)");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_dev_trace> t(new zkgpu_dev_trace());
    t->ncols = mem::NUM_COLUMNS;
    t->n = n;
    t->buf = DevBuf(&c, t->ncols * n * 8);
    uint64_t* d = t->buf.get();
    // the 14 operation columns: filter, timestamp | is_read, context, segment, virtual, 8 value limbs
    c.h2d(d + (size_t)mem::FILTER * n, ops, 2 * n * 8);
    c.h2d(d + (size_t)mem::IS_READ * n, ops + 2 * n, 12 * n * 8);
    ZK_CUDA(cudaMemsetAsync(d + (size_t)mem::STALE_CONTEXTS * n, 0, 3 * n * 8, c.stream));      // stale_contexts, is_pruned, stale_context_frequencies
    ZK_CUDA(cudaMemsetAsync(d + (size_t)mem::FREQUENCIES * n, 0, n * 8, c.stream));
    DevBuf aux(&c, (num_stale + 1) * 8);          // [0] = error flag, then the stale list
    ZK_CUDA(cudaMemsetAsync(aux.get(), 0, 8, c.stream));
    if (num_stale) c.h2d(aux.get() + 1, stale_contexts, num_stale * 8);
    {
        KernelScope ks(c, KF_TRACE_GEN, 8.0 * 30.0 * (double)n);
        if (num_stale) {
            memory_stale_kernel<<<(unsigned)((num_stale + 255) / 256), 256, 0, c.stream>>>(aux.get() + 1, num_stale, n, d);
            c.count_launch();
        }
        memory_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(d, n, (unsigned*)aux.get());
        c.count_launch();
        c.check_launch("memory_finish_kernel");
    }
    uint64_t flag = 0;
    c.d2h(&flag, aux.get(), 8);
    ZK_REQUIRE((flag & 0xFFFFFFFFull) == 0, "a range-checked difference does not fit the table: the operations are not sorted / gap-filled "
                                            "(memory_stark.rs:190-194)");
    *out = t.release();
    ZK_API_END
}

int zkgpu_dev_trace_dims(const zkgpu_dev_trace* t, size_t* ncols, size_t* n) {
    ZK_API_BEGIN
    ZK_REQUIRE(t, "null argument");
    if (ncols) *ncols = t->ncols;
    if (n) *n = t->n;
    ZK_API_END
}

const uint64_t* zkgpu_dev_trace_ptr(const zkgpu_dev_trace* t) { return t ? t->buf.get() : nullptr; }

int zkgpu_dev_trace_export(const zkgpu_dev_trace* t, uint64_t* host_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(t && host_out, "null argument");
    zk::Ctx& c = *t->buf.ctx;
    ZK_CUDA(cudaSetDevice(c.device));
    c.d2h(host_out, t->buf.get(), t->ncols * t->n * 8);
    ZK_API_END
}

// pinned host memory for hosts that do not link the CUDA runtime (include/zkgpu.h)
int zkgpu_host_register(const void* ptr, size_t bytes) {
    ZK_API_BEGIN
    ZK_REQUIRE(ptr && bytes, "null argument");
    ZK_CUDA(cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable));
    ZK_API_END
}
int zkgpu_host_unregister(const void* ptr) {
    ZK_API_BEGIN
    ZK_REQUIRE(ptr, "null argument");
    ZK_CUDA(cudaHostUnregister(const_cast<void*>(ptr)));
    ZK_API_END
}
int zkgpu_host_alloc(size_t bytes, void** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(out && bytes, "null argument");
    ZK_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    ZK_API_END
}
int zkgpu_host_free(void* ptr) {
    ZK_API_BEGIN
    if (ptr) ZK_CUDA(cudaFreeHost(ptr));
    ZK_API_END
}

void zkgpu_dev_trace_free(zkgpu_dev_trace* t) {
    if (!t) return;
    if (t->buf.ctx) cudaSetDevice(t->buf.ctx->device);
    delete t;
}

}  // extern "C"
