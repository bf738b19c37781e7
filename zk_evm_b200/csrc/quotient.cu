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
#include "quotient_kernel.cuh"
#include "ntt.h"

namespace zk {

using namespace zkstark;

__global__ void quotient_domain_kernel(DomArgs a) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.N) return;
    const uint32_t i = bitrev32((uint32_t)j, a.log_N);
    const uint64_t x = gl_mul(GL_GENERATOR, gl_pow(a.w_N, i));
    const uint64_t xm1 = gl_sub(x, 1), xml = gl_sub(x, a.last);
    const uint64_t d = gl_inv(gl_mul(xm1, xml));
    a.dom[j] = xml;
    a.dom[a.N + j] = gl_mul(a.c_first[i & 1], gl_mul(d, xml));      // L_first(x) = (Z_H(x) / n) / (x - 1)
    a.dom[2 * a.N + j] = gl_mul(a.c_last[i & 1], gl_mul(d, xm1));   // L_last(x) = (Z_H(x) w_n^-1 / n) / (x - w_n^-1)
}

static const uint64_t* get_domain(Ctx& c, unsigned k, const uint64_t zh[2]) {
    std::string key = "qdom:" + std::to_string(k);
    auto it = c.table_cache.find(key);
    if (it != c.table_cache.end()) return it->second.get();
    DomArgs a;
    a.log_N = k + 1; a.N = (size_t)2 << k;
    DevBuf b(&c, 3 * a.N * 8);
    a.dom = b.get();
    a.w_N = gl_root_of_unity(k + 1);
    a.last = gl_inv(gl_root_of_unity(k));
    uint64_t ninv = gl_inv(gl_canon((uint64_t)1 << k));
    for (int h = 0; h < 2; h++) { a.c_first[h] = gl_mul(zh[h], ninv); a.c_last[h] = gl_mul(gl_mul(zh[h], a.last), ninv); }
    quotient_domain_kernel<<<(unsigned)((a.N + 127) / 128), 128, 0, c.stream>>>(a);
    c.count_launch();
    c.check_launch("quotient_domain_kernel");
    const uint64_t* p = b.get();
    c.table_cache.emplace(key, std::move(b));
    return p;
}

// ---- alpha-independent constraint values + their Horner combination ------------------------------------------------------------------
// constraints the table's own evaluator yields per point (its yield_constr calls; the oracle counts the same: orc_table_num_constraints)
static const uint32_t TABLE_CONSTRAINTS[9] = {707, 535, 514, 868, 706, 524, 45, 1, 1};
uint32_t total_constraints(const TableDev& t) {
    // + per lookup and challenge: one per helper column, Z(first row) and the Z transition; + the CTL section
    return TABLE_CONSTRAINTS[t.table] + t.flat.num_lookup_cols + t.view.n_lookups + t.flat.ctl_num_constraints;
}

void constraints_record(Ctx& c, const TableDev& t, const QuotientArgs& q, uint64_t* cons, uint32_t expect) {
    const unsigned k = q.log_n;
    KernelScope ks(c, KF_QUOTIENT, 8.0 * ((size_t)2 << k) * (zkstark::table_num_columns(q.table) + t.flat.num_aux() + expect));
    RecKernelArgs a;
    a.trace_lde = q.trace_lde; a.aux_lde = q.aux_lde; a.cons = cons;
    a.log_N = k + 1; a.N = (size_t)2 << k;
    for (unsigned i = 0; i < 2; i++) { a.betas[i] = i < q.num_challenges ? q.betas[i] : 0; a.gammas[i] = i < q.num_challenges ? q.gammas[i] : 0; }
    uint64_t gn = gl_pow(GL_GENERATOR, (uint64_t)1 << k);
    uint64_t zh[2] = {gl_sub(gn, 1), gl_sub(gl_neg(gn), 1)};
    a.dom = get_domain(c, k, zh);
    a.flat = t.view;
    a.prm = q.prm;
    // the kernel's own count of what it yielded is checked against the expected width of the buffer once per (table, challenges)
    const std::string key = "ncons:" + std::to_string(q.table) + ":" + std::to_string(q.num_challenges);
    const bool verify = c.table_cache.find(key) == c.table_cache.end();
    DevBuf cnt(&c, 8);
    a.count_out = reinterpret_cast<uint32_t*>(cnt.get());
    switch (q.table) {
        case T_LOGIC: launch_record<T_LOGIC>(a, c.stream); break;
        case T_MEMORY: launch_record<T_MEMORY>(a, c.stream); break;
        case T_MEM_BEFORE: case T_MEM_AFTER: launch_record<T_MEM_BEFORE>(a, c.stream); break;
#if ZKS_ALL_TABLES
        case T_ARITHMETIC: launch_record<T_ARITHMETIC>(a, c.stream); break;
        case T_BYTE_PACKING: launch_record<T_BYTE_PACKING>(a, c.stream); break;
        case T_CPU: launch_record<T_CPU>(a, c.stream); break;
        case T_KECCAK: launch_record<T_KECCAK>(a, c.stream); break;
        case T_KECCAK_SPONGE: launch_record<T_KECCAK_SPONGE>(a, c.stream); break;
#endif
        default: throw ZkError(ZKGPU_ERR_INVALID, "constraints_record: table id not supported");
    }
    c.count_launch();
    c.check_launch("constraints_record_kernel");
    if (verify) {
        uint32_t got = 0;         // the kernel writes one 32-bit word (initcheck, profiles/r2v: reading 8 bytes here read 4 nobody wrote)
        c.d2h(&got, cnt.get(), 4);
        ZK_REQUIRE(got == expect, "constraints_record: the evaluator yielded " + std::to_string(got) + " constraints, the buffer has " +
                                                std::to_string(expect) + " columns");
        c.table_cache.emplace(key, DevBuf(&c, 8));
    }
}

// out[k * N + i] = zh_inv[i & 1] * sum_t alpha_k^(T - 1 - t) cons[t * N + j],  i = bitrev(j)
__global__ void __launch_bounds__(256) quotient_combine_kernel(const uint64_t* __restrict__ cons, uint32_t T, size_t N, unsigned log_N, unsigned nc,
                                                               uint64_t a0, uint64_t a1, uint64_t zh0, uint64_t zh1, uint64_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    uint64_t acc0 = 0, acc1 = 0;
    const uint64_t* p = cons + j;
    uint32_t t = 0;
    for (; t + 4 <= T; t += 4) {      // four loads in flight per thread
        const uint64_t c0 = __ldg(p), c1 = __ldg(p + N), c2 = __ldg(p + 2 * N), c3 = __ldg(p + 3 * N);
        p += 4 * N;
        acc0 = gl_add(gl_mul(acc0, a0), c0); acc0 = gl_add(gl_mul(acc0, a0), c1); acc0 = gl_add(gl_mul(acc0, a0), c2); acc0 = gl_add(gl_mul(acc0, a0), c3);
        if (nc > 1) { acc1 = gl_add(gl_mul(acc1, a1), c0); acc1 = gl_add(gl_mul(acc1, a1), c1); acc1 = gl_add(gl_mul(acc1, a1), c2); acc1 = gl_add(gl_mul(acc1, a1), c3); }
    }
    for (; t < T; t++) {
        const uint64_t cv = __ldg(p);
        p += N;
        acc0 = gl_add(gl_mul(acc0, a0), cv);
        if (nc > 1) acc1 = gl_add(gl_mul(acc1, a1), cv);
    }
    const uint32_t i = bitrev32((uint32_t)j, log_N);
    const uint64_t zh = (i & 1) ? zh1 : zh0;
    out[i] = gl_mul(acc0, zh);
    if (nc > 1) out[N + i] = gl_mul(acc1, zh);
}

void quotient_from_constraints(Ctx& c, const uint64_t* cons, uint32_t T, unsigned log_n, unsigned num_challenges, const uint64_t* alphas, uint64_t* out) {
    const size_t N = (size_t)2 << log_n;
    KernelScope ks(c, KF_QUOTIENT, 8.0 * N * (T + num_challenges));
    uint64_t gn = gl_pow(GL_GENERATOR, (uint64_t)1 << log_n);
    const uint64_t zh0 = gl_inv(gl_sub(gn, 1)), zh1 = gl_inv(gl_sub(gl_neg(gn), 1));
    quotient_combine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c.stream>>>(cons, T, N, log_n + 1, num_challenges, alphas[0],
                                                                               num_challenges > 1 ? alphas[1] : 0, zh0, zh1, out);
    c.count_launch();
    c.check_launch("quotient_combine_kernel");
}

void quotient_values(Ctx& c, const TableDev& t, const QuotientArgs& q) {
    ZK_REQUIRE(q.num_challenges >= 1 && q.num_challenges <= 2, "num_challenges must be 1 or 2");
    // each LDE row of trace + aux read once, num_challenges values per point written (16n(c+a) + 16n*nc)
    KernelScope ks(c, KF_QUOTIENT, 8.0 * ((size_t)2 << q.log_n) * (zkstark::table_num_columns(q.table) + t.flat.num_aux() + q.num_challenges));
    QuotKernelArgs a;
    const unsigned k = q.log_n;
    a.trace_lde = q.trace_lde; a.aux_lde = q.aux_lde; a.out = q.out;
    a.log_N = k + 1; a.N = (size_t)2 << k; a.nc = q.num_challenges;
    for (unsigned i = 0; i < 2; i++) {
        a.alphas[i] = i < a.nc ? q.alphas[i] : 0; a.betas[i] = i < a.nc ? q.betas[i] : 0; a.gammas[i] = i < a.nc ? q.gammas[i] : 0;
    }
    uint64_t gn = gl_pow(GL_GENERATOR, (uint64_t)1 << k);
    uint64_t zh[2] = {gl_sub(gn, 1), gl_sub(gl_neg(gn), 1)};
    for (int b = 0; b < 2; b++) a.zh_inv[b] = gl_inv(zh[b]);
    a.dom = get_domain(c, k, zh);
    a.flat = t.view;
    a.prm = q.prm;
    // powers of the alphas for the index-addressed constraint blocks (2 x 1024 elements, rebuilt per proof: alpha is fresh)
    static_assert(zkstark::keccak::MIDDLE_CONSTRAINTS <= APOW_MAX, "alpha power table too short");
    ZK_REQUIRE(t.flat.ctl_num_constraints <= APOW_MAX, "CTL constraint section longer than the alpha power table (APOW_MAX)");
    DevBuf apow(&c, 2 * (size_t)(APOW_MAX + 1) * 8);
    for (unsigned i = 0; i < 2; i++) fill_powers(c, apow.get() + i * (size_t)(APOW_MAX + 1), APOW_MAX + 1, a.alphas[i], 1);
    a.apow = apow.get();
    switch (q.table) {
        case T_LOGIC: launch_quotient<T_LOGIC>(a, c.stream); break;
        case T_MEMORY: launch_quotient<T_MEMORY>(a, c.stream); break;
        case T_MEM_BEFORE: case T_MEM_AFTER: launch_quotient<T_MEM_BEFORE>(a, c.stream); break;
#if ZKS_ALL_TABLES
        case T_ARITHMETIC: launch_quotient<T_ARITHMETIC>(a, c.stream); break;
        case T_BYTE_PACKING: launch_quotient<T_BYTE_PACKING>(a, c.stream); break;
        case T_CPU: launch_quotient<T_CPU>(a, c.stream); break;
        case T_KECCAK: launch_quotient<T_KECCAK>(a, c.stream); break;
        case T_KECCAK_SPONGE: launch_quotient<T_KECCAK_SPONGE>(a, c.stream); break;
#endif
        default: throw ZkError(ZKGPU_ERR_INVALID, "quotient: table id not supported");
    }
    c.count_launch();
    c.check_launch("quotient_kernel");
}

}  // namespace zk
