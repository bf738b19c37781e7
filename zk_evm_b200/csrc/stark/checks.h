// Lookup (logUp) and cross-table-lookup constraint checks, interpreted from a FlatView.
// Restates starky 1.0.0 lookup.rs `eval_packed_lookups_generic` / `eval_helper_columns` and cross_table_lookup.rs
// `eval_cross_table_lookup_checks` (called from eval_vanishing_poly after the table's own constraints; in-tree
// spec: /root/reference/book/src/framework/range_check.md:55-120 and ctls.md:17-25).
#pragma once
#include "hd.h"
#include "lookup.h"

namespace zkstark {

// The interpreter functions are NOT inlined on the device: they are called from many sites and the quotient kernels are
// already far larger than the instruction cache.
template <class P, class V>
ZKS_HD P flat_eval_col(const FlatView& f, uint32_t id, const V& lv, const V& nv) {
    if (id & FLAT_CELL) return lv[id & ~FLAT_CELL];
    const ColRec r = f.cols[id];
    P acc = P::from_u64(r.constant);
    for (uint32_t t = r.lin_begin; t < r.lin_end; t++) {
        const uint64_t k = f.term_coef[t];
        P v = lv[f.term_col[t]];
        acc = acc + (k == 1 ? v : v * P::from_u64(k));       // most CTL columns are plain cells
    }
    for (uint32_t t = r.next_begin; t < r.next_end; t++) {
        const uint64_t k = f.term_coef[t];
        P v = nv[f.term_col[t]];
        acc = acc + (k == 1 ? v : v * P::from_u64(k));
    }
    return acc;
}
template <class P, class V>
ZKS_HD_NOINLINE P flat_eval_filter(const FlatView& f, uint32_t id, const V& lv, const V& nv) {
    const FilterRec r = f.filters[id];
    P acc = P::zero();
    for (uint32_t k = r.prod_begin; k < r.prod_end; k += 2)
        acc = acc + flat_eval_col<P>(f, f.prod_ids[k], lv, nv) * flat_eval_col<P>(f, f.prod_ids[k + 1], lv, nv);
    for (uint32_t k = r.const_begin; k < r.const_end; k++) acc = acc + flat_eval_col<P>(f, f.const_ids[k], lv, nv);
    return acc;
}
// GrandProductChallenge::combine: sum_i v_i beta^i + gamma
template <class P, class V>
ZKS_HD_NOINLINE P flat_combine(const FlatView& f, const EntryRec& e, P beta, P gamma, const V& lv, const V& nv) {
    P acc = P::zero();
    for (uint32_t k = e.col_end; k-- > e.col_begin;) acc = acc * beta + flat_eval_col<P>(f, f.col_ids[k], lv, nv);
    return acc + gamma;
}

// eval_helper_columns: chunks of two entries per helper column
template <class P, class V, class A, class CC>
ZKS_HD void flat_eval_helpers(const FlatView& f, uint32_t entry_begin, uint32_t entry_end, uint32_t num_helpers,
                              uint32_t helper_begin, P beta, P gamma, const V& lv, const V& nv, const A& aux_lv, CC& yc) {
    for (uint32_t t = 0; t < num_helpers; t++) {
        uint32_t e0 = entry_begin + 2 * t;
        P h = aux_lv[helper_begin + t];
        P c0 = flat_combine<P>(f, f.entries[e0], beta, gamma, lv, nv);
        P f0 = flat_eval_filter<P>(f, f.entries[e0].filter, lv, nv);
        if (e0 + 1 < entry_end) {
            P c1 = flat_combine<P>(f, f.entries[e0 + 1], beta, gamma, lv, nv);
            P f1 = flat_eval_filter<P>(f, f.entries[e0 + 1].filter, lv, nv);
            yc.constraint(c1 * c0 * h - f0 * c1 - f1 * c0);
        } else {
            yc.constraint(c0 * h - f0);
        }
    }
}

// betas / gammas: the CTL challenges (lookups use beta_k as their challenge with combine(beta=1, gamma=beta_k))
template <class P, class V, class A, class CC>
ZKS_HD void flat_eval_lookups(const FlatView& f, const P* betas, const V& lv, const V& nv, const A& aux_lv, const A& aux_nv, CC& yc) {
    for (uint32_t li = 0; li < f.n_lookups; li++) {
        ZKS_SYNC();   // descriptor-driven loops: every thread of the block runs the same iterations, so the warps can share the fetched code
        const LookupRec l = f.lookups[li];
        P ch = betas[l.challenge];
        flat_eval_helpers<P>(f, l.entry_begin, l.entry_end, l.num_helpers, l.helper_begin, P::one(), ch, lv, nv, aux_lv, yc);
        P z = aux_lv[l.z_col], next_z = aux_nv[l.z_col];
        P table_with_challenge = flat_eval_col<P>(f, l.table_col, lv, nv) + ch;   // table column has no next-row terms
        P hsum = P::zero();
        for (uint32_t t = 0; t < l.num_helpers; t++) hsum = hsum + aux_lv[l.helper_begin + t];
        P y = hsum * table_with_challenge - flat_eval_col<P>(f, l.freq_col, lv, nv);
        yc.constraint_first_row(z);
        yc.constraint((next_z - z) * table_with_challenge - y);
    }
}

// GrandProductChallenge::combine for both challenges at once: the column values are walked once
template <class P, class V>
ZKS_HD_NOINLINE void flat_combine2(const FlatView& f, const EntryRec& e, const P* betas, const P* gammas, const V& lv, const V& nv, P& c0, P& c1) {
    P a0 = P::zero(), a1 = P::zero();
    for (uint32_t k = e.col_end; k-- > e.col_begin;) {
        const P v = flat_eval_col<P>(f, f.col_ids[k], lv, nv);
        a0 = a0 * betas[0] + v;
        a1 = a1 * betas[1] + v;
    }
    c0 = a0 + gammas[0];
    c1 = a1 + gammas[1];
}

// The CTL section with two challenges: an item and its twin (same columns and filters, other challenge) are evaluated together —
// one walk over the columns and filters instead of two — and their constraints go to their own positions of the section through
// the index-addressed block of the consumer (block_put), so the result is the one the sequential emission gives.
template <class P, class V, class A, class CC>
ZKS_HD void flat_eval_ctls_paired(const FlatView& f, const P* betas, const P* gammas, const V& lv, const V& nv, const A& aux_lv,
                                  const A& aux_nv, CC& yc) {
    if (f.n_ctl_zs == 0) return;
    yc.block_begin(f.ctl_num_constraints);
    for (uint32_t zi = 0; zi < f.n_ctl_zs; zi++) {
        ZKS_SYNC();
        const CtlZRec c = f.ctl_zs[zi];
        if (c.challenge != 0) continue;
        const CtlZRec d = f.ctl_zs[c.twin];
        const P lz0 = aux_lv[c.z_col], nz0 = aux_nv[c.z_col], lz1 = aux_lv[d.z_col], nz1 = aux_nv[d.z_col];
        if (c.num_helpers) {
            P hs0 = P::zero(), hs1 = P::zero();
            for (uint32_t t = 0; t < c.num_helpers; t++) {
                const uint32_t e0 = c.entry_begin + 2 * t;
                const P h0 = aux_lv[c.helper_begin + t], h1 = aux_lv[d.helper_begin + t];
                P a0, a1;
                flat_combine2<P>(f, f.entries[e0], betas, gammas, lv, nv, a0, a1);
                const P f0 = flat_eval_filter<P>(f, f.entries[e0].filter, lv, nv);
                if (e0 + 1 < c.entry_end) {
                    P b0, b1;
                    flat_combine2<P>(f, f.entries[e0 + 1], betas, gammas, lv, nv, b0, b1);
                    const P f1 = flat_eval_filter<P>(f, f.entries[e0 + 1].filter, lv, nv);
                    yc.block_put(c.cons_begin + t, b0 * a0 * h0 - f0 * b0 - f1 * a0);
                    yc.block_put(d.cons_begin + t, b1 * a1 * h1 - f0 * b1 - f1 * a1);
                } else {
                    yc.block_put(c.cons_begin + t, a0 * h0 - f0);
                    yc.block_put(d.cons_begin + t, a1 * h1 - f0);
                }
                hs0 = hs0 + h0; hs1 = hs1 + h1;
            }
            yc.block_put(c.cons_begin + c.num_helpers, (lz0 - hs0) * yc.lagrange_last);
            yc.block_put(c.cons_begin + c.num_helpers + 1, (lz0 - nz0 - hs0) * yc.z_last);
            yc.block_put(d.cons_begin + d.num_helpers, (lz1 - hs1) * yc.lagrange_last);
            yc.block_put(d.cons_begin + d.num_helpers + 1, (lz1 - nz1 - hs1) * yc.z_last);
        } else {
            P a0, a1;
            flat_combine2<P>(f, f.entries[c.entry_begin], betas, gammas, lv, nv, a0, a1);
            const P f0 = flat_eval_filter<P>(f, f.entries[c.entry_begin].filter, lv, nv);
            yc.block_put(c.cons_begin, (a0 * lz0 - f0) * yc.lagrange_last);
            yc.block_put(c.cons_begin + 1, (a0 * (lz0 - nz0) - f0) * yc.z_last);
            yc.block_put(d.cons_begin, (a1 * lz1 - f0) * yc.lagrange_last);
            yc.block_put(d.cons_begin + 1, (a1 * (lz1 - nz1) - f0) * yc.z_last);
        }
    }
}

template <class P, class V, class A, class CC>
ZKS_HD void flat_eval_ctls(const FlatView& f, const P* betas, const P* gammas, const V& lv, const V& nv, const A& aux_lv,
                           const A& aux_nv, CC& yc) {
    if (f.ctl_paired) { flat_eval_ctls_paired<P>(f, betas, gammas, lv, nv, aux_lv, aux_nv, yc); return; }
    for (uint32_t zi = 0; zi < f.n_ctl_zs; zi++) {
        ZKS_SYNC();
        const CtlZRec c = f.ctl_zs[zi];
        P beta = betas[c.challenge], gamma = gammas[c.challenge];
        P local_z = aux_lv[c.z_col], next_z = aux_nv[c.z_col];
        if (c.num_helpers) {
            flat_eval_helpers<P>(f, c.entry_begin, c.entry_end, c.num_helpers, c.helper_begin, beta, gamma, lv, nv, aux_lv, yc);
            P hsum = P::zero();
            for (uint32_t t = 0; t < c.num_helpers; t++) hsum = hsum + aux_lv[c.helper_begin + t];
            yc.constraint_last_row(local_z - hsum);
            yc.constraint_transition(local_z - next_z - hsum);
        } else {
            P c0 = flat_combine<P>(f, f.entries[c.entry_begin], beta, gamma, lv, nv);
            P f0 = flat_eval_filter<P>(f, f.entries[c.entry_begin].filter, lv, nv);
            yc.constraint_last_row(c0 * local_z - f0);
            yc.constraint_transition(c0 * (local_z - next_z) - f0);
        }
    }
}

}  // namespace zkstark
