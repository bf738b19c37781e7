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
