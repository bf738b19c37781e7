// host/device qualifier shared by the single-source constraint headers (usable from nvcc and from plain g++)
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define ZKS_HD __host__ __device__ __forceinline__
#else
#define ZKS_HD inline
#endif
