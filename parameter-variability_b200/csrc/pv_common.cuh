// Shared definitions for the generated model headers and the integrator templates.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PV_HD __host__ __device__ __forceinline__
#define PV_D __device__ __forceinline__
#else
// The generated model functions and the per-lane integrator steps are plain C++ so that the
// tests can compile them with g++ in the GPU-less build container.  Nothing in the shipped
// library takes this branch: pv_abi.cu is only ever compiled by nvcc.
#define PV_HD inline
#define PV_D inline
#endif
