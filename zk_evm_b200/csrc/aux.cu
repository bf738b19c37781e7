// Auxiliary columns of a STARK table: logUp range-check helpers + running sum, cross-table-lookup helpers + running sum.
//
// Replaces starky 1.0.0 lookup.rs `lookup_helper_columns` and cross_table_lookup.rs `get_ctl_data` /
// `cross_table_lookup_data` / `partial_sums` / `get_helper_cols` (called at
// /root/reference/evm_arithmetization/src/prover.rs:137-143 and inside prove_with_commitment; in-tree spec
// book/src/framework/ctls.md:17-25, range_check.md:55-120).
//
// Three small kernels over the raw trace values (column-major, natural row order):
//   1. helper_kernel   per helper column, a thread takes AUX_ROWS rows (a block-stride apart, so every load stays coalesced):
//                      h = sum over <= 2 (columns, filter) pairs of filter/combined; the AUX_ROWS denominators share ONE Fermat
//                      inversion (Montgomery's trick: 3 products per row + 76 per thread instead of 76 per row), as the CPU code
//                      batch-inverts a whole column
//   2. zsum_kernel     per running-sum column the per-row increment (sum of helpers [- freq/(table+challenge)])
//   3. scan            modular prefix / suffix sums (block scan + scan of block totals + offset add)
#include "stark_dev.h"

namespace zk {

using zkstark::FlatView; using zkstark::ColRec; using zkstark::FilterRec; using zkstark::EntryRec;

// Column::eval_table: next-row terms count as zero on the last row
__device__ __forceinline__ uint64_t col_eval_table(const FlatView& f, uint32_t id, const uint64_t* __restrict__ v, size_t n, size_t r) {
    if (id & zkstark::FLAT_CELL) return v[(size_t)(id & ~zkstark::FLAT_CELL) * n + r];
    const ColRec c = f.cols[id];
    uint64_t acc = c.constant;
    for (uint32_t t = c.lin_begin; t < c.lin_end; t++) acc = gl_add(acc, gl_mul(v[(size_t)f.term_col[t] * n + r], f.term_coef[t]));
    if (r + 1 < n)
        for (uint32_t t = c.next_begin; t < c.next_end; t++) acc = gl_add(acc, gl_mul(v[(size_t)f.term_col[t] * n + r + 1], f.term_coef[t]));
    return acc;
}
__device__ __forceinline__ uint64_t filter_eval_table(const FlatView& f, uint32_t id, const uint64_t* __restrict__ v, size_t n, size_t r) {
    const FilterRec fr = f.filters[id];
    uint64_t acc = 0;
    for (uint32_t k = fr.prod_begin; k < fr.prod_end; k += 2)
        acc = gl_add(acc, gl_mul(col_eval_table(f, f.prod_ids[k], v, n, r), col_eval_table(f, f.prod_ids[k + 1], v, n, r)));
    for (uint32_t k = fr.const_begin; k < fr.const_end; k++) acc = gl_add(acc, col_eval_table(f, f.const_ids[k], v, n, r));
    return acc;
}
__device__ __forceinline__ uint64_t combine_table(const FlatView& f, const EntryRec& e, uint64_t beta, uint64_t gamma,
                                                  const uint64_t* __restrict__ v, size_t n, size_t r) {
    uint64_t acc = 0;
    for (uint32_t k = e.col_end; k-- > e.col_begin;) acc = gl_add(gl_mul(acc, beta), col_eval_table(f, f.col_ids[k], v, n, r));
    return gl_add(acc, gamma);
}

// out_col2 != NO_COL: the same (columns, filter) pairs under a second challenge (beta2, gamma2) -> second output column; the filters
// and the column values are then evaluated once for both
struct HelperJob { uint32_t entry_begin, entry_end, out_col, out_col2; uint64_t beta, gamma, beta2, gamma2; };
static constexpr uint32_t NO_COL = 0xFFFFFFFFu;
__device__ __forceinline__ void combine_table2(const FlatView& f, const EntryRec& e, const HelperJob& j, const uint64_t* __restrict__ v, size_t n,
                                               size_t r, uint64_t& c0, uint64_t& c1) {
    uint64_t a0 = 0, a1 = 0;
    for (uint32_t k = e.col_end; k-- > e.col_begin;) {
        const uint64_t x = col_eval_table(f, f.col_ids[k], v, n, r);
        a0 = gl_add(gl_mul(a0, j.beta), x);
        a1 = gl_add(gl_mul(a1, j.beta2), x);
    }
    c0 = gl_add(a0, j.gamma);
    c1 = gl_add(a1, j.gamma2);
}

static constexpr int AUX_ROWS = 8;
// inv[i] = 1 / d[i] for the rows < cnt (0 where d[i] == 0, as gl_inv(0) = 0): one inversion for the whole batch
__device__ __forceinline__ void batch_inverse(const uint64_t (&d)[AUX_ROWS], uint64_t (&inv)[AUX_ROWS]) {
    uint64_t pre[AUX_ROWS];
    uint64_t run = 1;
#pragma unroll
    for (int i = 0; i < AUX_ROWS; i++) { pre[i] = run; run = gl_mul(run, d[i] ? d[i] : 1); }
    uint64_t r = gl_inv(run);
#pragma unroll
    for (int i = AUX_ROWS - 1; i >= 0; i--) {
        inv[i] = d[i] ? gl_mul(r, pre[i]) : 0;
        r = gl_mul(r, d[i] ? d[i] : 1);
    }
}

__global__ void __launch_bounds__(256) helper_kernel(FlatView f, const HelperJob* __restrict__ jobs, const uint64_t* __restrict__ values,
                                                     size_t n, uint64_t* __restrict__ out) {
    const HelperJob j = jobs[blockIdx.y];
    const size_t r0 = (size_t)blockIdx.x * (256 * AUX_ROWS) + threadIdx.x;
    uint64_t num[AUX_ROWS], den[AUX_ROWS], inv[AUX_ROWS];
    if (j.out_col2 != NO_COL) {
        // two challenges: AUX_ROWS / 2 rows x 2 denominators share the inversion; rows r0 + i*256, i < AUX_ROWS/2, and the second
        // half of the block's row range is taken by the threads' upper slots (r0 + (AUX_ROWS/2 + i) * 256) in a second sweep
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
#pragma unroll
            for (int i = 0; i < AUX_ROWS / 2; i++) {
                const size_t r = r0 + (size_t)(half * (AUX_ROWS / 2) + i) * 256;
                num[2 * i] = num[2 * i + 1] = 0; den[2 * i] = den[2 * i + 1] = 1;
                if (r >= n) continue;
                uint64_t fs[2] = {0, 0}, c0[2] = {1, 1}, c1[2] = {1, 1};
                for (uint32_t e = j.entry_begin; e < j.entry_end; e++) {
                    const EntryRec er = f.entries[e];
                    const uint64_t fv = filter_eval_table(f, er.filter, values, n, r);
                    fs[e - j.entry_begin] = fv;
                    if (fv) combine_table2(f, er, j, values, n, r, c0[e - j.entry_begin], c1[e - j.entry_begin]);
                }
                den[2 * i] = gl_mul(c0[0], c0[1]);
                num[2 * i] = gl_add(gl_mul(fs[0], c0[1]), gl_mul(fs[1], c0[0]));
                den[2 * i + 1] = gl_mul(c1[0], c1[1]);
                num[2 * i + 1] = gl_add(gl_mul(fs[0], c1[1]), gl_mul(fs[1], c1[0]));
            }
            batch_inverse(den, inv);
#pragma unroll
            for (int i = 0; i < AUX_ROWS / 2; i++) {
                const size_t r = r0 + (size_t)(half * (AUX_ROWS / 2) + i) * 256;
                if (r < n) {
                    out[(size_t)j.out_col * n + r] = gl_mul(num[2 * i], inv[2 * i]);
                    out[(size_t)j.out_col2 * n + r] = gl_mul(num[2 * i + 1], inv[2 * i + 1]);
                }
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < AUX_ROWS; i++) {
        const size_t r = r0 + (size_t)i * 256;
        num[i] = 0; den[i] = 1;
        if (r >= n) continue;
        uint64_t fs[2] = {0, 0}, cs[2] = {1, 1};
        for (uint32_t e = j.entry_begin; e < j.entry_end; e++) {
            const EntryRec er = f.entries[e];
            uint64_t fv = filter_eval_table(f, er.filter, values, n, r);
            fs[e - j.entry_begin] = fv;
            if (fv) cs[e - j.entry_begin] = combine_table(f, er, j.beta, j.gamma, values, n, r);
        }
        // f0/c0 + f1/c1 = (f0 c1 + f1 c0) / (c0 c1)
        den[i] = gl_mul(cs[0], cs[1]);
        num[i] = gl_add(gl_mul(fs[0], cs[1]), gl_mul(fs[1], cs[0]));
    }
    batch_inverse(den, inv);
#pragma unroll
    for (int i = 0; i < AUX_ROWS; i++) {
        const size_t r = r0 + (size_t)i * 256;
        if (r < n) out[(size_t)j.out_col * n + r] = gl_mul(num[i], inv[i]);
    }
}

struct ZJob { uint32_t kind;   // 0: CTL (sum of helper columns), 1: lookup (shifted increment)
              uint32_t helper_begin, num_helpers, out_col, table_col, freq_col; uint64_t challenge; };

__global__ void __launch_bounds__(256) zsum_kernel(FlatView f, const ZJob* __restrict__ jobs, const uint64_t* __restrict__ values,
                                                   size_t n, uint64_t* __restrict__ out) {
    const ZJob j = jobs[blockIdx.y];
    const size_t r0 = (size_t)blockIdx.x * (256 * AUX_ROWS) + threadIdx.x;
    if (j.kind == 0) {
#pragma unroll 1
        for (int i = 0; i < AUX_ROWS; i++) {
            const size_t r = r0 + (size_t)i * 256;
            if (r >= n) break;
            uint64_t acc = 0;
            for (uint32_t t = 0; t < j.num_helpers; t++) acc = gl_add(acc, out[(size_t)(j.helper_begin + t) * n + r]);
            out[(size_t)j.out_col * n + r] = acc;
        }
        return;
    }
    // Z(r) = Z(r-1) + sum_t h_t(r-1) - freq(r-1) / (table(r-1) + challenge): store the increment at row r
    uint64_t acc[AUX_ROWS], den[AUX_ROWS], inv[AUX_ROWS], fr[AUX_ROWS];
#pragma unroll
    for (int i = 0; i < AUX_ROWS; i++) {
        const size_t r = r0 + (size_t)i * 256;
        acc[i] = 0; den[i] = 1; fr[i] = 0;
        if (r >= n || r == 0) continue;
        const size_t q = r - 1;
        uint64_t a = 0;
        for (uint32_t t = 0; t < j.num_helpers; t++) a = gl_add(a, out[(size_t)(j.helper_begin + t) * n + q]);
        acc[i] = a;
        den[i] = gl_add(col_eval_table(f, j.table_col, values, n, q), j.challenge);
        fr[i] = col_eval_table(f, j.freq_col, values, n, q);
    }
    batch_inverse(den, inv);
#pragma unroll
    for (int i = 0; i < AUX_ROWS; i++) {
        const size_t r = r0 + (size_t)i * 256;
        if (r < n) out[(size_t)j.out_col * n + r] = gl_sub(acc[i], gl_mul(fr[i], inv[i]));
    }
}

// ---- modular scan ------------------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
struct ScanJob { uint32_t col; uint32_t reverse; };   // reverse: suffix sums (scan over n-1-i)

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* sm, uint64_t* total) {
    // Hillis-Steele over SCAN_THREADS values in shared memory
    int t = threadIdx.x;
    sm[t] = v;
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1) {
        uint64_t x = t >= off ? sm[t - off] : 0;
        __syncthreads();
        if (t >= off) sm[t] = gl_add(sm[t], x);
        __syncthreads();
    }
    uint64_t incl = sm[t];
    *total = sm[SCAN_THREADS - 1];
    __syncthreads();
    return gl_sub(incl, v);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const ScanJob* __restrict__ jobs, uint64_t* __restrict__ data, size_t n,
                                                                  uint64_t* __restrict__ totals, size_t ntiles) {
    __shared__ uint64_t sm[SCAN_THREADS];
    const ScanJob j = jobs[blockIdx.y];
    uint64_t* col = data + (size_t)j.col * n;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t run = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + i;
        uint64_t x = 0;
        if (idx < n) x = col[j.reverse ? n - 1 - idx : idx];
        run = gl_add(run, x);
        v[i] = run;
    }
    uint64_t total;
    uint64_t off = block_exclusive_scan(run, sm, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t idx = base + i;
        if (idx < n) col[j.reverse ? n - 1 - idx : idx] = gl_add(v[i], off);
    }
    if (threadIdx.x == 0) totals[(size_t)blockIdx.y * ntiles + blockIdx.x] = total;
}
// exclusive scan of the tile totals of each job (one block per job, sequential over chunks of SCAN_THREADS tiles)
__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(uint64_t* __restrict__ totals, size_t ntiles) {
    __shared__ uint64_t sm[SCAN_THREADS];
    uint64_t* t = totals + (size_t)blockIdx.x * ntiles;
    uint64_t carry = 0;
    for (size_t base = 0; base < ntiles; base += SCAN_THREADS) {
        size_t idx = base + threadIdx.x;
        uint64_t x = idx < ntiles ? t[idx] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_scan(x, sm, &total);
        if (idx < ntiles) t[idx] = gl_add(ex, carry);
        carry = gl_add(carry, total);
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(const ScanJob* __restrict__ jobs, uint64_t* __restrict__ data, size_t n,
                                                                const uint64_t* __restrict__ totals, size_t ntiles) {
    const ScanJob j = jobs[blockIdx.y];
    uint64_t off = totals[(size_t)blockIdx.y * ntiles + blockIdx.x];
    if (off == 0) return;
    uint64_t* col = data + (size_t)j.col * n;
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS) {
        size_t idx = base + i;
        if (idx < n) { size_t p = j.reverse ? n - 1 - idx : idx; col[p] = gl_add(col[p], off); }
    }
}

template <class T> static DevBuf upload_vec(Ctx& c, const std::vector<T>& v) {
    DevBuf b(&c, v.size() * sizeof(T));
    if (!v.empty()) c.h2d(b.get(), v.data(), v.size() * sizeof(T));
    return b;
}

static void run_scans(Ctx& c, const std::vector<ScanJob>& jobs, uint64_t* data, size_t n) {
    if (jobs.empty()) return;
    size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf dj = upload_vec(c, jobs);
    DevBuf totals(&c, jobs.size() * ntiles * 8);
    for (size_t j0 = 0; j0 < jobs.size(); j0 += 65535) {
        unsigned cnt = (unsigned)std::min<size_t>(65535, jobs.size() - j0);
        dim3 grid((unsigned)ntiles, cnt);
        const ScanJob* jp = (const ScanJob*)dj.get() + j0;
        uint64_t* tp = totals.get() + j0 * ntiles;
        scan_tiles_kernel<<<grid, SCAN_THREADS, 0, c.stream>>>(jp, data, n, tp, ntiles);
        if (ntiles > 1) {
            scan_totals_kernel<<<cnt, SCAN_THREADS, 0, c.stream>>>(tp, ntiles);
            scan_add_kernel<<<grid, SCAN_THREADS, 0, c.stream>>>(jp, data, n, tp, ntiles);
            c.count_launch(2);
        }
        c.count_launch();
    }
    c.check_launch("scan kernels");
    c.sync();   // dj / totals are freed stream-ordered, but the host vectors behind the async copies must outlive them
}

static void run_helpers(Ctx& c, const TableDev& t, const std::vector<HelperJob>& jobs, const uint64_t* values, size_t n, uint64_t* out) {
    if (jobs.empty()) return;
    DevBuf dj = upload_vec(c, jobs);
    for (size_t j0 = 0; j0 < jobs.size(); j0 += 65535) {
        unsigned cnt = (unsigned)std::min<size_t>(65535, jobs.size() - j0);
        dim3 grid((unsigned)((n + 256 * AUX_ROWS - 1) / (256 * AUX_ROWS)), cnt);
        helper_kernel<<<grid, 256, 0, c.stream>>>(t.view, (const HelperJob*)dj.get() + j0, values, n, out);
        c.count_launch();
    }
    c.check_launch("helper_kernel");
    c.sync();
}
static void run_zsums(Ctx& c, const TableDev& t, const std::vector<ZJob>& jobs, const uint64_t* values, size_t n, uint64_t* out) {
    if (jobs.empty()) return;
    DevBuf dj = upload_vec(c, jobs);
    dim3 grid((unsigned)((n + 256 * AUX_ROWS - 1) / (256 * AUX_ROWS)), (unsigned)jobs.size());
    zsum_kernel<<<grid, 256, 0, c.stream>>>(t.view, (const ZJob*)dj.get(), values, n, out);
    c.count_launch();
    c.check_launch("zsum_kernel");
    c.sync();
}

void ctl_columns(Ctx& c, const TableDev& t, const uint64_t* values, size_t n, const uint64_t* betas, const uint64_t* gammas,
                 uint64_t* out) {
    KernelScope ks(c, KF_AUX, 8.0 * n * (zkstark::table_num_columns(t.table) + t.flat.num_ctl_helpers + t.flat.num_ctl_zs));
    const zkstark::Flat& f = t.flat;
    const uint32_t base = f.num_lookup_cols;   // `out` starts at the first CTL helper column
    std::vector<HelperJob> hj;
    std::vector<ZJob> zj;
    std::vector<ScanJob> sj;
    for (const zkstark::CtlZRec& z : f.ctl_zs) {
        // with two challenges an item and its twin (same entries, other challenge) share their helper jobs
        const bool paired = f.ctl_paired && z.twin != zkstark::NO_TWIN;
        const zkstark::CtlZRec* tw = paired ? &f.ctl_zs[z.twin] : nullptr;
        const bool second = paired && z.challenge != 0;     // its columns are produced by the twin's jobs
        if (z.num_helpers) {
            for (uint32_t h = 0; h < z.num_helpers && !second; h++) {
                HelperJob j;
                j.entry_begin = z.entry_begin + 2 * h;
                j.entry_end = std::min(z.entry_end, j.entry_begin + 2);
                j.out_col = z.helper_begin - base + h; j.out_col2 = NO_COL;
                j.beta = betas[z.challenge]; j.gamma = gammas[z.challenge]; j.beta2 = j.gamma2 = 0;
                if (paired) { j.out_col2 = tw->helper_begin - base + h; j.beta2 = betas[tw->challenge]; j.gamma2 = gammas[tw->challenge]; }
                hj.push_back(j);
            }
            ZJob q; q.kind = 0; q.helper_begin = z.helper_begin - base; q.num_helpers = z.num_helpers; q.out_col = z.z_col - base;
            q.table_col = q.freq_col = 0; q.challenge = 0;
            zj.push_back(q);
        } else if (!second) {
            HelperJob j;   // single pair: filter/combined goes straight into the Z column
            j.entry_begin = z.entry_begin; j.entry_end = z.entry_end; j.out_col = z.z_col - base; j.out_col2 = NO_COL;
            j.beta = betas[z.challenge]; j.gamma = gammas[z.challenge]; j.beta2 = j.gamma2 = 0;
            if (paired) { j.out_col2 = tw->z_col - base; j.beta2 = betas[tw->challenge]; j.gamma2 = gammas[tw->challenge]; }
            hj.push_back(j);
        }
        sj.push_back({z.z_col - base, 1u});
    }
    run_helpers(c, t, hj, values, n, out);
    run_zsums(c, t, zj, values, n, out);
    run_scans(c, sj, out, n);
}

void lookup_columns(Ctx& c, const TableDev& t, const uint64_t* values, size_t n, const uint64_t* betas, uint64_t* out) {
    KernelScope ks(c, KF_AUX, 8.0 * n * (zkstark::table_num_columns(t.table) + t.flat.num_lookup_cols));
    const zkstark::Flat& f = t.flat;
    std::vector<HelperJob> hj;
    std::vector<ZJob> zj;
    std::vector<ScanJob> sj;
    for (const zkstark::LookupRec& l : f.lookups) {
        for (uint32_t h = 0; h < l.num_helpers; h++) {
            HelperJob j;
            j.entry_begin = l.entry_begin + 2 * h;
            j.entry_end = std::min(l.entry_end, j.entry_begin + 2);
            j.out_col = l.helper_begin + h; j.out_col2 = NO_COL;
            j.beta = 1; j.gamma = betas[l.challenge]; j.beta2 = j.gamma2 = 0;
            hj.push_back(j);
        }
        ZJob q; q.kind = 1; q.helper_begin = l.helper_begin; q.num_helpers = l.num_helpers; q.out_col = l.z_col;
        q.table_col = l.table_col; q.freq_col = l.freq_col; q.challenge = betas[l.challenge];
        zj.push_back(q);
        sj.push_back({l.z_col, 0u});
    }
    run_helpers(c, t, hj, values, n, out);
    run_zsums(c, t, zj, values, n, out);
    run_scans(c, sj, out, n);
}

// ---- descriptor upload ----------------------------------------------------------------------------------------------
const TableDev& get_table_dev(Ctx& c, uint32_t table, unsigned num_challenges) {
    uint32_t key = table * 16 + num_challenges;
    auto it = c.stark_tables.find(key);
    if (it != c.stark_tables.end()) return *it->second;
    ZK_REQUIRE(zkstark::table_supported(table), "table id not supported");
    auto td = std::make_shared<TableDev>();
    td->table = table; td->num_challenges = num_challenges;
    auto ctls = zkstark::all_cross_table_lookups();
    auto items = zkstark::table_ctl_items(table, ctls, num_challenges);
    td->flat = zkstark::build_table_flat(zkstark::table_lookups(table), items, num_challenges, zkstark::CONSTRAINT_DEGREE);
    const zkstark::Flat& f = td->flat;
    // pack every array into one buffer, 16-byte aligned sections
    size_t off = 0;
    auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    size_t o_tc = place(f.term_col.size() * 4), o_tf = place(f.term_coef.size() * 8), o_cols = place(f.cols.size() * sizeof(ColRec)),
           o_ci = place(f.col_ids.size() * 4), o_pi = place(f.prod_ids.size() * 4), o_ki = place(f.const_ids.size() * 4),
           o_fl = place(f.filters.size() * sizeof(FilterRec)), o_en = place(f.entries.size() * sizeof(EntryRec)),
           o_cz = place(f.ctl_zs.size() * sizeof(zkstark::CtlZRec)), o_lk = place(f.lookups.size() * sizeof(zkstark::LookupRec));
    std::vector<uint8_t> host(off + 16, 0);
    auto put = [&](size_t o, const void* p, size_t bytes) { if (bytes) memcpy(host.data() + o, p, bytes); };
    put(o_tc, f.term_col.data(), f.term_col.size() * 4); put(o_tf, f.term_coef.data(), f.term_coef.size() * 8);
    put(o_cols, f.cols.data(), f.cols.size() * sizeof(ColRec)); put(o_ci, f.col_ids.data(), f.col_ids.size() * 4);
    put(o_pi, f.prod_ids.data(), f.prod_ids.size() * 4); put(o_ki, f.const_ids.data(), f.const_ids.size() * 4);
    put(o_fl, f.filters.data(), f.filters.size() * sizeof(FilterRec)); put(o_en, f.entries.data(), f.entries.size() * sizeof(EntryRec));
    put(o_cz, f.ctl_zs.data(), f.ctl_zs.size() * sizeof(zkstark::CtlZRec));
    put(o_lk, f.lookups.data(), f.lookups.size() * sizeof(zkstark::LookupRec));
    td->buf = DevBuf(&c, host.size());
    c.h2d(td->buf.get(), host.data(), host.size());
    c.sync();
    const uint8_t* d = (const uint8_t*)td->buf.get();
    FlatView& v = td->view;
    v.term_col = (const uint32_t*)(d + o_tc); v.term_coef = (const uint64_t*)(d + o_tf); v.cols = (const ColRec*)(d + o_cols);
    v.col_ids = (const uint32_t*)(d + o_ci); v.prod_ids = (const uint32_t*)(d + o_pi); v.const_ids = (const uint32_t*)(d + o_ki);
    v.filters = (const FilterRec*)(d + o_fl); v.entries = (const EntryRec*)(d + o_en);
    v.ctl_zs = (const zkstark::CtlZRec*)(d + o_cz); v.lookups = (const zkstark::LookupRec*)(d + o_lk);
    v.n_ctl_zs = (uint32_t)f.ctl_zs.size(); v.n_lookups = (uint32_t)f.lookups.size();
    v.num_lookup_cols = f.num_lookup_cols; v.num_ctl_helpers = f.num_ctl_helpers; v.num_ctl_zs = f.num_ctl_zs;
    v.ctl_num_constraints = f.ctl_num_constraints; v.ctl_paired = f.ctl_paired;
    c.stark_tables[key] = td;
    return *td;
}

}  // namespace zk
