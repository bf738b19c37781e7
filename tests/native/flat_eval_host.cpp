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
