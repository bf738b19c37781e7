// host/device qualifier shared by the single-source constraint headers (usable from nvcc and from plain g++)
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define ZKS_HD __host__ __device__ __forceinline__
#define ZKS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ZKS_HD inline
#define ZKS_HD_NOINLINE inline
#endif

// Loop-unrolling control for the device build of the constraint templates.  The quotient kernels are straight-line code far
// larger than the 32 KB L1.5 instruction cache; keeping the big constraint loops rolled (columns are read from memory by index,
// so nothing needs static indices) keeps the hot loop bodies resident.  No effect on the host (oracle) build.
#if defined(__CUDA_ARCH__)
#define ZKS_NOUNROLL _Pragma("unroll 1")
#define ZKS_UNROLL _Pragma("unroll")
// keeps the warps of a block walking the (instruction-cache-busting) constraint code together: they then share the
// instruction lines one of them has fetched.  Every thread of the block reaches every ZKS_SYNC (no early exits in the kernel).
#define ZKS_SYNC() __syncthreads()
#else
#define ZKS_NOUNROLL
#define ZKS_UNROLL
#define ZKS_SYNC() ((void)0)
#endif

