// TEST INFRASTRUCTURE: runs the product's lookup / cross-table-lookup constraint interpreter (zk_evm_b200/csrc/stark/checks.h over the
// flat descriptors of stark/lookup.h — the code the CUDA quotient kernels execute per point) on the CPU, on arbitrary rows, and
// compares the accumulators with the oracle's eval_vanishing_poly (which walks the Column / Filter objects directly).
// Build: g++ -I zk_evm_b200/csrc -I oracle (tests/test_flat_eval_host.py).  Not part of the product.
#include "oracle_stark.h"
#include "stark/checks.h"
using namespace orc;

static zkstark::FlatView view_of(const zkstark::Flat& f) {
    zkstark::FlatView v;
    v.term_col = f.term_col.data(); v.term_coef = f.term_coef.data(); v.cols = f.cols.data(); v.col_ids = f.col_ids.data();
    v.prod_ids = f.prod_ids.data(); v.const_ids = f.const_ids.data(); v.filters = f.filters.data(); v.entries = f.entries.data();
    v.ctl_zs = f.ctl_zs.data(); v.lookups = f.lookups.data();
    v.n_ctl_zs = (uint32_t)f.ctl_zs.size(); v.n_lookups = (uint32_t)f.lookups.size();
    v.num_lookup_cols = f.num_lookup_cols; v.num_ctl_helpers = f.num_ctl_helpers; v.num_ctl_zs = f.num_ctl_zs;
    v.ctl_num_constraints = f.ctl_num_constraints; v.ctl_paired = f.ctl_paired;
    return v;
}

// rows: lv, nv (ncols each), alv, anv (naux each); challenges: alphas[2], betas[2], gammas[2]; sel: z_last, lagrange_first, lagrange_last.
// Returns 1 when the flat interpreter and the oracle agree on both accumulators, 0 when they differ, -1 on a shape mismatch.
extern "C" int flat_eval_agrees(uint32_t table, uint32_t num_challenges, const uint64_t* lv, const uint64_t* nv, const uint64_t* alv,
                                const uint64_t* anv, const uint64_t* alphas, const uint64_t* betas, const uint64_t* gammas,
                                const uint64_t* sel, const uint64_t labels[4], uint64_t out_oracle[2], uint64_t out_flat[2],
                                uint32_t* num_aux_out) {
    const unsigned cd = zkstark::CONSTRAINT_DEGREE;
    auto ctls = zkstark::all_cross_table_lookups();
    AuxShape sh = aux_shape(table, ctls, num_challenges, cd);
    zkstark::Flat flat = zkstark::build_table_flat(zkstark::table_lookups(table), zkstark::table_ctl_items(table, ctls, num_challenges),
                                                  num_challenges, cd);
    if (num_aux_out) *num_aux_out = flat.num_aux() | (flat.ctl_paired ? 0x80000000u : 0u);
    if (flat.num_aux() != sh.num_aux()) return -1;
    zkstark::TableParams prm = {labels[0], labels[1], labels[2], labels[3]};
    ConsumerT<OF> a, b;
    std::vector<OF> bs, gs;
    for (unsigned j = 0; j < num_challenges; j++) { bs.push_back(OF(betas[j])); gs.push_back(OF(gammas[j])); }
    for (ConsumerT<OF>* y : {&a, &b}) {
        for (unsigned j = 0; j < num_challenges; j++) { y->alphas.push_back(OF(alphas[j])); y->acc.push_back(OF(0)); }
        y->z_last = OF(sel[0]); y->lagrange_first = OF(sel[1]); y->lagrange_last = OF(sel[2]);
    }
    RowOF l{lv}, n{nv}, al{alv}, an{anv};
    eval_vanishing_poly<OF>(table, sh, bs, gs, l, n, al, an, a, prm, cd);
    // the device path: table constraints, then the two interpreters
    zkstark::eval_table<OF>(table, l, n, b, prm);
    zkstark::FlatView v = view_of(flat);
    zkstark::flat_eval_lookups<OF>(v, bs.data(), l, n, al, an, b);
    zkstark::flat_eval_ctls<OF>(v, bs.data(), gs.data(), l, n, al, an, b);
    int ok = 1;
    for (unsigned j = 0; j < num_challenges; j++) {
        out_oracle[j] = a.acc[j].v; out_flat[j] = b.acc[j].v;
        if (a.acc[j].v != b.acc[j].v) ok = 0;
    }
    return ok;
}

// ---- for tests/test_ptx_emulation.py: the flat descriptors of a table as the bytes the device sees ---------------------------------
// Packs the arrays like get_table_dev (aux.cu): each section 16-byte aligned inside `buf`; offs[10] = byte offsets of term_col,
// term_coef, cols, col_ids, prod_ids, const_ids, filters, entries, ctl_zs, lookups; scalars[7] = n_ctl_zs, n_lookups, num_lookup_cols,
// num_ctl_helpers, num_ctl_zs, ctl_num_constraints, ctl_paired.  Returns the number of bytes needed (call with buf = NULL first).
extern "C" size_t flat_pack(uint32_t table, uint32_t num_challenges, uint8_t* buf, size_t cap, uint64_t offs[10], uint32_t scalars[7]) {
    auto ctls = zkstark::all_cross_table_lookups();
    zkstark::Flat f = zkstark::build_table_flat(zkstark::table_lookups(table), zkstark::table_ctl_items(table, ctls, num_challenges),
                                               num_challenges, zkstark::CONSTRAINT_DEGREE);
    size_t off = 0;
    auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    const size_t o[10] = {place(f.term_col.size() * 4), place(f.term_coef.size() * 8), place(f.cols.size() * sizeof(zkstark::ColRec)),
                          place(f.col_ids.size() * 4), place(f.prod_ids.size() * 4), place(f.const_ids.size() * 4),
                          place(f.filters.size() * sizeof(zkstark::FilterRec)), place(f.entries.size() * sizeof(zkstark::EntryRec)),
                          place(f.ctl_zs.size() * sizeof(zkstark::CtlZRec)), place(f.lookups.size() * sizeof(zkstark::LookupRec))};
    const size_t need = off + 16;
    if (!buf || cap < need) return need;
    memset(buf, 0, need);
    auto put = [&](size_t at, const void* p, size_t bytes) { if (bytes) memcpy(buf + at, p, bytes); };
    put(o[0], f.term_col.data(), f.term_col.size() * 4); put(o[1], f.term_coef.data(), f.term_coef.size() * 8);
    put(o[2], f.cols.data(), f.cols.size() * sizeof(zkstark::ColRec)); put(o[3], f.col_ids.data(), f.col_ids.size() * 4);
    put(o[4], f.prod_ids.data(), f.prod_ids.size() * 4); put(o[5], f.const_ids.data(), f.const_ids.size() * 4);
    put(o[6], f.filters.data(), f.filters.size() * sizeof(zkstark::FilterRec)); put(o[7], f.entries.data(), f.entries.size() * sizeof(zkstark::EntryRec));
    put(o[8], f.ctl_zs.data(), f.ctl_zs.size() * sizeof(zkstark::CtlZRec)); put(o[9], f.lookups.data(), f.lookups.size() * sizeof(zkstark::LookupRec));
    for (int i = 0; i < 10; i++) offs[i] = o[i];
    scalars[0] = (uint32_t)f.ctl_zs.size(); scalars[1] = (uint32_t)f.lookups.size(); scalars[2] = f.num_lookup_cols;
    scalars[3] = f.num_ctl_helpers; scalars[4] = f.num_ctl_zs; scalars[5] = f.ctl_num_constraints; scalars[6] = f.ctl_paired;
    return need;
}
