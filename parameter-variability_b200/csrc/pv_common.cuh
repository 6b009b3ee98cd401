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

// compile-time integer as a value (generic lambdas instantiated per register set)
template <int I>
struct pv_int_tag {
    static constexpr int value = I;
};

PV_HD float pv_fdiv_fast(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);     // MUFU.RCP + FMUL: the controller needs two digits, not IEEE division
#else
    return a / b;
#endif
}
// max(|a|, |b|) for the error scale: a compare and two selects.  fmax() has no fp64 instruction on this GPU - it compiles to
// DSETP.MAX + moves + selects + a NaN fix-up, ~9 instructions per call, three calls per step and component; for ordered
// operands the result is the same bits, and a NaN operand makes the error norm NaN either way (the step is then handled as
// non-finite).
PV_HD double pv_absmax(double a, double b) {
    const double x = fabs(a), y = fabs(b);
    return x > y ? x : y;
}
// ~20-bit reciprocal (one MUFU.RCP64H): enough for the error *norm*, which only steers the step size
PV_HD double pv_rcp_approx(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return 1.0 / x;
#endif
}
// full-precision reciprocal of a normal, finite x without the slow path of __drcp_rn: MUFU.RCP64H seed (~20 bits) and
// two Newton steps
PV_HD double pv_rcp_newton(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}
