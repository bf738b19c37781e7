// ConstraintConsumer (starky 1.0.0 constraint_consumer.rs): Horner accumulation in alpha over the emission order.
// P is the evaluation type: zk::Fp on the LDE coset (prover, device), an extension element at zeta (verifier).
#pragma once
#include "hd.h"

namespace zkstark {

template <class P, int MAXC = 2>
struct Consumer {
    P alpha[MAXC];
    P acc[MAXC];
    int nc;
    P z_last;          // x - g^-1 (g = trace subgroup generator): vanishes on the last row
    P lagrange_first;  // L_0(x)
    P lagrange_last;   // L_{n-1}(x)

    ZKS_HD void constraint(P c) {
#pragma unroll
        for (int j = 0; j < MAXC; j++)
            if (j < nc) acc[j] = acc[j] * alpha[j] + c;
    }
    // A block of M constraints that may be emitted in ANY order by their index inside the block (the position they have in the
    // reference's emission order): with acc' = acc alpha^M + sum_i c_i alpha^(M-1-i) the result is the one Horner gives for
    // the same M constraints emitted in order.  apow[j][e] = alpha_j^e for e <= M.  This is what lets a wide table (Keccak) walk
    // its columns once instead of once per constraint family.
    const P* apow[MAXC];
    uint32_t blk_top;
    ZKS_HD void block_begin(uint32_t M) {
#pragma unroll
        for (int j = 0; j < MAXC; j++)
            if (j < nc) acc[j] = acc[j] * apow[j][M];
        blk_top = M - 1;
    }
    ZKS_HD void block_put(uint32_t idx, P c) {
#pragma unroll
        for (int j = 0; j < MAXC; j++)
            if (j < nc) acc[j] = acc[j] + c * apow[j][blk_top - idx];
    }
    ZKS_HD void constraint_transition(P c) { constraint(c * z_last); }
    ZKS_HD void constraint_first_row(P c) { constraint(c * lagrange_first); }
    ZKS_HD void constraint_last_row(P c) { constraint(c * lagrange_last); }
};

// The same interface, but nothing is accumulated: constraint k of the emission order (index-addressed blocks: at their index) is
// WRITTEN to column k of a buffer, already multiplied by its row selector (z_last / Lagrange).  The constraint values do not depend on
// the alphas, so a prover that has the trace and auxiliary LDEs but not yet its turn in the transcript (a table-sharded segment: the
// tables are finished one after the other, prover.rs:251-259) can evaluate them ahead of time; the quotient values are then the
// Horner combination sum_k alpha^(T-1-k) column_k — a streaming pass instead of the evaluator.
template <class P>
struct RecordConsumer {
    uint64_t* out;        // &buffer[point]; constraint k lives at out[k * stride]
    size_t stride;
    uint32_t idx = 0, blk_base = 0;
    P z_last, lagrange_first, lagrange_last;
    ZKS_HD void put(uint32_t k, P c) { out[(size_t)k * stride] = c.v; }
    ZKS_HD void constraint(P c) { put(idx++, c); }
    ZKS_HD void block_begin(uint32_t M) { blk_base = idx; idx += M; }
    ZKS_HD void block_put(uint32_t i, P c) { put(blk_base + i, c); }
    ZKS_HD void constraint_transition(P c) { constraint(c * z_last); }
    ZKS_HD void constraint_first_row(P c) { constraint(c * lagrange_first); }
    ZKS_HD void constraint_last_row(P c) { constraint(c * lagrange_last); }
};

// Parameters some tables' constraints need besides the two rows.
struct TableParams {
    // KERNEL.global_labels[...] used by CpuStark (cpu/control_flow.rs:38-44, cpu/syscalls_exceptions.rs:68-73)
    uint64_t halt_final, init, syscall_jumptable, exception_jumptable;
};

}  // namespace zkstark
