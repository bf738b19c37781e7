// Integer pipe micro-benchmark (development tool): issue cost of the instruction forms the Goldilocks/Poseidon kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipebench.cu -o tools/pipebench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define ILP 8
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 2654435761u + seed + i; y[i] = x[i] ^ 0x9e3779b9u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0) {          // mul.wide.u32 : 1 per unit
                uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x[i]), "r"(y[i]));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(x[i]), "=r"(y[i]) : "l"(r));
            } else if (KIND == 1) {   // mad.lo.u32
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (KIND == 2) {   // mul.hi.u32
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (KIND == 3) {   // lop3 (xor) : alu only
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (KIND == 4) {   // add.cc + addc : 2 per unit
                asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %1, %1, %0;" : "+r"(x[i]), "+r"(y[i]));
            } else if (KIND == 5) {   // mul.wide + 2 xor : 3 per unit
                uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x[i]), "r"(y[i]));
                uint32_t a, b; asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(r));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
            } else if (KIND == 6) {   // mul.wide + 4 xor : 5 per unit
                uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x[i]), "r"(y[i]));
                uint32_t a, b; asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(r));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
            } else if (KIND == 7) {   // shf (funnel shift)
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(y[i]));
            } else if (KIND == 8) {   // mad.lo + xor : 2 per unit (fma + alu overlap test)
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(x[i]));
            } else if (KIND == 9) {   // add.u32 plain
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (KIND == 10) {  // shift-add: (x << 4) + y  (LEA / IMAD)
                asm volatile("{ .reg .u32 t; shl.b32 t, %0, 4; add.u32 %0, t, %1; }" : "+r"(x[i]) : "r"(y[i]));
            } else if (KIND == 11) {  // mad.wide.u32 with 64-bit accumulate
                uint64_t r; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(x[i]), "r"(y[i]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(x[i]), "r"(y[i]));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(x[i]), "=r"(y[i]) : "l"(r));
            } else if (KIND == 12) {  // sub.cc/subc/subc (borrow mask)
                uint32_t m;
                asm volatile("sub.cc.u32 %0, %0, %1;\n\tsubc.u32 %2, 0, 0;" : "+r"(x[i]), "+r"(y[i]), "=r"(m));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(m));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= x[i] ^ y[i];
    if (r == 0x1234567) out[0] = r;
}

template <int KIND> void run(const char* name, double per_unit, uint32_t* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2048; unsigned blocks = 148 * 8; float ms = 0;
    for (int w = 0; w < 3; w++) { cudaEventRecord(e0); k<KIND><<<blocks, 256>>>(d, iters, w); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    double units = (double)blocks * 256 * iters * ILP;
    double units_per_clk_sm = units / (ms * 1e-3) / 148 / 1.965e9;
    printf("%-28s %8.3f ms  %.1f thread-units/clk/SM  -> %.2f clk per warp-unit per SMSP (%.2f clk/instr at %g instr/unit)\n", name, ms, units_per_clk_sm,
           128.0 / units_per_clk_sm, 128.0 / units_per_clk_sm / per_unit, per_unit);
}

int main() {
    uint32_t* d; cudaMalloc(&d, 4096);
    run<0>("mul.wide.u32", 1, d);
    run<1>("mad.lo.u32", 1, d);
    run<2>("mul.hi.u32", 1, d);
    run<3>("xor (LOP3)", 1, d);
    run<9>("add.u32", 1, d);
    run<7>("shf.l.wrap", 1, d);
    run<10>("shl+add (LEA?)", 1, d);
    run<4>("add.cc+addc", 2, d);
    run<12>("sub.cc+subc(mask)+xor", 3, d);
    run<8>("mad.lo + xor", 2, d);
    run<5>("mul.wide + 2 xor", 3, d);
    run<6>("mul.wide + 4 xor", 5, d);
    run<11>("mad.wide (64-bit acc)", 1, d);
    return 0;
}
