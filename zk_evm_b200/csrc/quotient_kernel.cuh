// Quotient-polynomial evaluation: one fused pass over a table's LDE that evaluates every constraint of the table, its
// logUp checks and its cross-table-lookup checks, Horner-accumulates them in alpha and divides by Z_H.
//
// Replaces starky 1.0.0 prover.rs `compute_quotient_polys` + `eval_vanishing_poly` + `ConstraintConsumer` and the
// `Stark::eval_packed_generic` of each table (evm_arithmetization/src/{cpu,memory,arithmetic,keccak,keccak_sponge,
// byte_packing,memory_continuation}/*_stark.rs, logic.rs), reached from
// /root/reference/evm_arithmetization/src/prover.rs:322-334.
//
// Layout: the LDE is column-major with bit-reversed rows (exactly the Merkle leaf order), so thread j owns storage
// position j == natural coset index i = bitrev(j).  A warp reads 32 consecutive elements of a column (one 256-byte
// request) for the local row, and — because i+2 only disturbs the high bits of j — another contiguous 32-element run
// for the next row.  The two running accumulators per challenge live in registers; columns are streamed straight from
// global memory / L2 (each is touched by a handful of constraints).
// This header holds the kernel template; quotient_tN.cu instantiate it for one table each (separate translation units so the
// nine large kernels compile in parallel), quotient.cu holds the launcher.
#pragma once
#include "stark_dev.h"
#include <stdlib.h>

namespace zk {

using namespace zkstark;

struct DevRow {
    const uint64_t* p;
    size_t stride;
    __device__ __forceinline__ Fp operator[](uint32_t c) const { return Fp(__ldg(p + (size_t)c * stride)); }
};

static constexpr unsigned APOW_MAX = 1023;

struct QuotKernelArgs {
    const uint64_t* trace_lde;
    const uint64_t* aux_lde;
    uint64_t* out;
    size_t N;
    unsigned log_N, nc;
    uint64_t alphas[2], betas[2], gammas[2];
    // per-point domain values in storage order j (they depend on log_N only; built once per size and cached):
    // dom[j] = x - w_n^-1 (vanishes on the last row), dom[N + j] = L_first(x), dom[2N + j] = L_last(x), x = g w_N^bitrev(j)
    const uint64_t* dom;
    uint64_t zh_inv[2];      // 1 / (g^n (-1)^i - 1)
    // alpha_j^e, e <= APOW_MAX, for the index-addressed constraint blocks (Consumer::block_put): apow + j * (APOW_MAX + 1)
    const uint64_t* apow;
    FlatView flat;
    TableParams prm;
};

// 128-thread blocks per SM the kernel is compiled for (i.e. its register budget, 65536 / (128 * blocks)).  Measured (profiles/r1h,
// r1k): the Keccak kernel is a load stream (2431 columns per point) and wants ~40 loads in flight per thread at 128 registers; the
// Cpu evaluator is instruction-fetch bound and gains from more resident warps: 12.8 ms at 3 blocks (168 registers), 8.0 ms at 5
// (96 registers, ~200 bytes of spills); the Arithmetic evaluator keeps long-lived limb arrays and does not: 6.9 ms at 3 blocks, 8.6 ms at
// 5 (2.2 KB of stack per thread) — profiles/r1p, r1x ncu rows; letting ptxas take 255 registers (min blocks = 1) made Cpu 25 % and
// Arithmetic 34 % slower.  The others are small and run best at high occupancy.
constexpr int quotient_min_blocks(uint32_t table) {
    return table == T_KECCAK ? 4 : table == T_CPU ? 5 : table == T_ARITHMETIC ? 3 : table == T_MEMORY ? 3 : 8;
}
// 128-thread blocks (Memory: 256).  Measured alternative (profiles/r1m): ONE 384-thread block per SM for the Arithmetic / Cpu evaluators
// (12 warps in lock-step sharing the instruction stream) is no faster in isolation (Cpu 14.6 vs 14.7 ms, Arithmetic 7.7 vs 6.7 ms), and a
// block that needs the whole register file of an SM can only start on an empty SM, which is the wrong shape when a second segment's
// small blocks are in flight.  (The 1040 ms two-stream step measured with it was later traced to the shared memory pool, see api.cu.)
// (Memory: 256-thread blocks at 85 registers — a quarter of an SM — for the same code sharing; its constraint code is descriptor loops.)
constexpr unsigned quotient_block_threads(uint32_t table) { return table == T_MEMORY ? 256 : 128; }

template <uint32_t TABLE>
__global__ void __launch_bounds__(quotient_block_threads(TABLE), quotient_min_blocks(TABLE)) quotient_kernel(QuotKernelArgs a) {
    // every thread runs the whole evaluator (the constraint code contains block-wide barriers, ZKS_SYNC): threads past the end of
    // the domain recompute the last point and skip the store
    const size_t j0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = j0 < a.N;
    const size_t j = live ? j0 : a.N - 1;
    const uint32_t i = bitrev32((uint32_t)j, a.log_N);
    const uint32_t inext = (i + 2) & (uint32_t)(a.N - 1);
    const size_t jn = bitrev32(inext, a.log_N);

    Consumer<Fp, 2> yc;
    yc.nc = (int)a.nc;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        yc.alpha[k] = Fp(a.alphas[k]); yc.acc[k] = Fp(0);
        yc.apow[k] = reinterpret_cast<const Fp*>(a.apow) + k * (APOW_MAX + 1);
    }
    yc.z_last = Fp(__ldg(a.dom + j));
    yc.lagrange_first = Fp(__ldg(a.dom + a.N + j));
    yc.lagrange_last = Fp(__ldg(a.dom + 2 * a.N + j));

    DevRow lv{a.trace_lde + j, a.N}, nv{a.trace_lde + jn, a.N};
    DevRow alv{a.aux_lde + j, a.N}, anv{a.aux_lde + jn, a.N};

    if constexpr (TABLE == T_LOGIC) logic::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_MEMORY) memory::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_MEM_BEFORE || TABLE == T_MEM_AFTER) memcont::eval<Fp>(lv, nv, yc);
#if ZKS_ALL_TABLES
    else if constexpr (TABLE == T_ARITHMETIC) arithmetic::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_BYTE_PACKING) byte_packing::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_CPU) cpu::eval<Fp>(lv, nv, yc, a.prm);
    else if constexpr (TABLE == T_KECCAK) keccak::eval_blocked<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_KECCAK_SPONGE) keccak_sponge::eval<Fp>(lv, nv, yc);
#endif

    Fp betas[2] = {Fp(a.betas[0]), Fp(a.betas[1])}, gammas[2] = {Fp(a.gammas[0]), Fp(a.gammas[1])};
    flat_eval_lookups<Fp>(a.flat, betas, lv, nv, alv, anv, yc);
    flat_eval_ctls<Fp>(a.flat, betas, gammas, lv, nv, alv, anv, yc);

    if (live)
        for (unsigned k = 0; k < a.nc; k++) a.out[(size_t)k * a.N + i] = gl_mul(yc.acc[k].v, a.zh_inv[i & 1]);
}

// ---- the alpha-independent half of the same evaluation (RecordConsumer, stark/consumer.h) --------------------------------------
// Same point <-> thread map, same evaluators; every constraint value goes to its column of a.cons (T x N, point index = storage
// position j, so a warp writes 32 consecutive words per constraint).  Thread 0 reports the number of constraints it yielded.
struct RecKernelArgs {
    const uint64_t* trace_lde;
    const uint64_t* aux_lde;
    uint64_t* cons;
    uint32_t* count_out;
    size_t N;
    unsigned log_N;
    uint64_t betas[2], gammas[2];
    const uint64_t* dom;
    FlatView flat;
    TableParams prm;
};

template <uint32_t TABLE>
__global__ void __launch_bounds__(quotient_block_threads(TABLE), quotient_min_blocks(TABLE)) constraints_record_kernel(RecKernelArgs a) {
    const size_t j0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = j0 < a.N;
    const size_t j = live ? j0 : a.N - 1;      // threads past the end redo the last point (the evaluators contain block-wide barriers)
    const uint32_t i = bitrev32((uint32_t)j, a.log_N);
    const uint32_t inext = (i + 2) & (uint32_t)(a.N - 1);
    const size_t jn = bitrev32(inext, a.log_N);

    RecordConsumer<Fp> yc;
    yc.out = a.cons + j; yc.stride = a.N;
    yc.z_last = Fp(__ldg(a.dom + j));
    yc.lagrange_first = Fp(__ldg(a.dom + a.N + j));
    yc.lagrange_last = Fp(__ldg(a.dom + 2 * a.N + j));

    DevRow lv{a.trace_lde + j, a.N}, nv{a.trace_lde + jn, a.N};
    DevRow alv{a.aux_lde + j, a.N}, anv{a.aux_lde + jn, a.N};

    if constexpr (TABLE == T_LOGIC) logic::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_MEMORY) memory::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_MEM_BEFORE || TABLE == T_MEM_AFTER) memcont::eval<Fp>(lv, nv, yc);
#if ZKS_ALL_TABLES
    else if constexpr (TABLE == T_ARITHMETIC) arithmetic::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_BYTE_PACKING) byte_packing::eval<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_CPU) cpu::eval<Fp>(lv, nv, yc, a.prm);
    else if constexpr (TABLE == T_KECCAK) keccak::eval_blocked<Fp>(lv, nv, yc);
    else if constexpr (TABLE == T_KECCAK_SPONGE) keccak_sponge::eval<Fp>(lv, nv, yc);
#endif

    Fp betas[2] = {Fp(a.betas[0]), Fp(a.betas[1])}, gammas[2] = {Fp(a.gammas[0]), Fp(a.gammas[1])};
    flat_eval_lookups<Fp>(a.flat, betas, lv, nv, alv, anv, yc);
    flat_eval_ctls<Fp>(a.flat, betas, gammas, lv, nv, alv, anv, yc);
    if (j0 == 0) *a.count_out = yc.idx;
}
template <uint32_t TABLE> void launch_record(const RecKernelArgs& a, cudaStream_t stream);

// domain table of the quotient kernels (see QuotKernelArgs::dom)
struct DomArgs { uint64_t* dom; size_t N; unsigned log_N; uint64_t w_N, last, c_first[2], c_last[2]; };

// one launcher per table, defined in quotient_t<N>.cu
template <uint32_t TABLE> void launch_quotient(const QuotKernelArgs& a, cudaStream_t stream);
#define ZK_INSTANTIATE_QUOTIENT(TABLE)                                                                       \
    template <> void launch_quotient<TABLE>(const QuotKernelArgs& a, cudaStream_t stream) {                  \
        constexpr unsigned T = quotient_block_threads(TABLE);                                                \
        quotient_kernel<TABLE><<<(unsigned)((a.N + T - 1) / T), T, 0, stream>>>(a);                          \
    }                                                                                                        \
    template <> void launch_record<TABLE>(const RecKernelArgs& a, cudaStream_t stream) {                    \
        constexpr unsigned T = quotient_block_threads(TABLE);                                                \
        constraints_record_kernel<TABLE><<<(unsigned)((a.N + T - 1) / T), T, 0, stream>>>(a);                \
    }

}  // namespace zk
