// Standalone micro-benchmark of Poseidon permutation variants + integer pipe throughput (development tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk_evm_b200/csrc -I tools tools/pbench.cu -o tools/pbench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon_fast.cuh"

using namespace zk;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int V>
__global__ void __launch_bounds__(128) perm_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) {
        if (V == 0) poseidon_permute(s);
        else { if (V == 1) pf_permute_unrolled<0>(s); else if (V == 2) pf_permute_unrolled<1>(s); else if (V == 3) pf_permute_r1(s); else pf_permute(s);
#pragma unroll
            for (int k = 0; k < 12; k++) s[k] = pf_canon(s[k]); }
    }
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}

// pipe throughput: ILP independent chains of one instruction kind
template <int KIND>
__global__ void __launch_bounds__(256) pipe_kernel(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (KIND == 0) acc[k] = pf_madwide((uint32_t)acc[(k + 1) & 7], b, acc[k]);                           // IMAD.WIDE.U32
            else if (KIND == 1) { uint32_t lo = (uint32_t)acc[k]; lo = lo * a + b; acc[k] = lo; }   // IMAD (32-bit)
            else if (KIND == 2) { uint32_t lo = (uint32_t)acc[k]; asm volatile("add.u32 %0, %0, %1;" : "+r"(lo) : "r"(b)); lo ^= a; acc[k] = lo; }  // IADD3+LOP3
            else { uint32_t lo = (uint32_t)acc[k], hi = (uint32_t)(acc[k] >> 32);            // 2 IMAD.WIDE + 2 IADD3 mix
                   uint64_t t = pf_madwide((uint32_t)acc[(k + 1) & 7], b, acc[k]); (void)lo; uint32_t x, y; pf_unpack(t, x, y);
                   asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(x), "+r"(y) : "r"(hi)); acc[k] = pf_pack(x, y); }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= acc[k];
    if (r == 0x1234567) out[0] = r;
}

int main(int argc, char** argv) {
    size_t count = 1 << 22;
    int reps = argc > 1 ? atoi(argv[1]) : 4;
    std::vector<uint64_t> h(12 * count);
    uint64_t x = 88172645463325252ULL;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x % GL_P; }
    // edge values in the first states
    for (int k = 0; k < 12; k++) { h[k * count + 0] = 0; h[k * count + 1] = GL_P - 1; h[k * count + 2] = 0xFFFFFFFFull; h[k * count + 3] = 0xFFFFFFFF00000000ull; }
    uint64_t *din, *d0, *d1, *d2, *d3, *d4;
    CK(cudaMalloc(&din, 12 * count * 8)); CK(cudaMalloc(&d0, 12 * count * 8)); CK(cudaMalloc(&d1, 12 * count * 8)); CK(cudaMalloc(&d2, 12 * count * 8)); CK(cudaMalloc(&d3, 12 * count * 8)); CK(cudaMalloc(&d4, 12 * count * 8));
    CK(cudaMemcpy(din, h.data(), 12 * count * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    unsigned blocks = (unsigned)((count + 127) / 128);
    float ms;
    for (int v = 0; v < 5; v++) {
        for (int w = 0; w < 3; w++) {
            cudaEventRecord(e0);
            if (v == 0) perm_kernel<0><<<blocks, 128>>>(din, d0, count, reps); else if (v == 1) perm_kernel<1><<<blocks, 128>>>(din, d1, count, reps);
            else if (v == 2) perm_kernel<2><<<blocks, 128>>>(din, d2, count, reps);
            else if (v == 3) perm_kernel<3><<<blocks, 128>>>(din, d3, count, reps);
            else perm_kernel<4><<<blocks, 128>>>(din, d4, count, reps);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("variant %d: %.3f ms for %zu x %d perms -> %.1f Mperm/s\n", v, ms, count, reps, count * (double)reps / ms / 1e3);
    }
    std::vector<uint64_t> r0(12 * count), r1(12 * count), r2(12 * count);
    CK(cudaMemcpy(r0.data(), d0, 12 * count * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r1.data(), d1, 12 * count * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r2.data(), d2, 12 * count * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    CK(cudaMemcpy(r1.data(), d4, 12 * count * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < 12 * count; i++) if (r0[i] != r1[i]) { if (bad < 5) printf("v4 (limb-form partial rounds, poseidon_fast.cuh pf_permute) mismatch at %zu\n", i); bad++; }
    CK(cudaMemcpy(r1.data(), d3, 12 * count * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < 12 * count; i++) if (r0[i] != r1[i]) { if (bad < 5) printf("v3 mismatch at %zu\n", i); bad++; }
    CK(cudaMemcpy(r1.data(), d1, 12 * count * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < 12 * count; i++) if (r0[i] != r2[i]) { if (bad < 5) printf("v2 mismatch at %zu: %llx vs %llx\n", i, (unsigned long long)r0[i], (unsigned long long)r2[i]); bad++; }
    for (size_t i = 0; i < 12 * count; i++) if (r0[i] != r1[i]) { if (bad < 5) printf("mismatch at %zu: %llx vs %llx\n", i, (unsigned long long)r0[i], (unsigned long long)r1[i]); bad++; }
    // host check of a few states against the host reference
    for (size_t i = 0; i < 8; i++) { uint64_t s[12]; for (int k = 0; k < 12; k++) s[k] = h[k * count + i]; for (int r = 0; r < reps; r++) poseidon_permute(s);
        for (int k = 0; k < 12; k++) if (s[k] != r0[k * count + i]) { printf("host mismatch state %zu word %d\n", i, k); bad++; } }
    printf("mismatches: %zu\n", bad);
    const char* names[] = {"IMAD.WIDE.U32", "IMAD(32)", "IADD3+LOP3", "IMAD.WIDE+2xIADD3"};
    const double per[] = {1, 1, 2, 3};
    for (int kind = 0; kind < 4; kind++) {
        int iters = 4096; unsigned pb = 148 * 8;
        for (int w = 0; w < 2; w++) {
            cudaEventRecord(e0);
            if (kind == 0) pipe_kernel<0><<<pb, 256>>>(d0, iters, w); else if (kind == 1) pipe_kernel<1><<<pb, 256>>>(d0, iters, w);
            else if (kind == 2) pipe_kernel<2><<<pb, 256>>>(d0, iters, w); else pipe_kernel<3><<<pb, 256>>>(d0, iters, w);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        double ops = (double)pb * 256 * iters * 8 * per[kind];
        printf("%s: %.3f ms, %.2f Tinstr/s (thread-level), = %.1f lanes/clk/SM at 1.965 GHz\n", names[kind], ms, ops / ms / 1e9, ops / ms / 1e3 / 148 / 1.965e6);
    }
    return bad != 0;
}
