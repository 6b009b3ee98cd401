// Locality schedule (divergence control): before a batch launch the trajectories are bucketed by their parameter values
// so that the 32 lanes of a warp integrate NEIGHBOURING points of parameter space.  Neighbours take nearly the same step
// sequence, so a warp's lanes accept / reject together, cross the output grid in the same iterations (the per-grid-point
// code then runs with most lanes active instead of ~7 of 32) and finish together.  Results cannot change: the order in
// which trajectories are picked up is all that differs (every trajectory is integrated by one lane from its own inputs
// and written to its own rows).
//
// Key: per column the top bits of the IEEE pattern of the value (monotone in the value for positive numbers:
// 2^(52 - PV_SCHED_SHIFT) levels per octave), centred on the batch mean, clamped to 8 bits; the first two columns are
// Morton-interleaved into a 16-bit bucket number.  Counting sort: histogram -> exclusive scan -> scatter.  Four small
// launches, ~0.05 ms for 10^6 trajectories; the order inside a bucket is whatever the atomics produce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PV_SCHED_BINS 65536
#define PV_SCHED_SHIFT 47          // 5 mantissa bits: 32 levels per octave, 8 octaves around the batch mean

struct pv_sched_params {
    const double* P;               // [n_cols][B]
    int64_t B;
    int n_key_cols;                // 1 or 2
    long long* sums;               // [2] sum of the raw keys per column
    unsigned int* bins;            // [PV_SCHED_BINS + 1] histogram, then bucket cursors
    int32_t* index;                // [B] work item -> trajectory
};

__device__ __forceinline__ long long pv_sched_raw_key(double v) {
    return (v > 0.0) ? (__double_as_longlong(v) >> PV_SCHED_SHIFT) : 0ll;      // <= 0 / NaN: all in the lowest bucket
}

__global__ void pv_sched_mean_kernel(const pv_sched_params p) {
    long long s0 = 0, s1 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.B; i += (int64_t)gridDim.x * blockDim.x) {
        s0 += pv_sched_raw_key(p.P[i]);
        if (p.n_key_cols > 1) s1 += pv_sched_raw_key(p.P[p.B + i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&p.sums[0], (unsigned long long)s0);
        if (p.n_key_cols > 1) atomicAdd((unsigned long long*)&p.sums[1], (unsigned long long)s1);
    }
}

__device__ __forceinline__ unsigned pv_sched_bucket(const pv_sched_params& p, int64_t i, long long c0, long long c1) {
    long long k0 = pv_sched_raw_key(p.P[i]) - c0 + 128;
    k0 = k0 < 0 ? 0 : (k0 > 255 ? 255 : k0);
    if (p.n_key_cols == 1) return (unsigned)k0 << 8;
    long long k1 = pv_sched_raw_key(p.P[p.B + i]) - c1 + 128;
    k1 = k1 < 0 ? 0 : (k1 > 255 ? 255 : k1);
    // Morton interleave of two 8-bit numbers
    unsigned x = (unsigned)k0, y = (unsigned)k1;
    x = (x | (x << 4)) & 0x0f0fu; x = (x | (x << 2)) & 0x3333u; x = (x | (x << 1)) & 0x5555u;
    y = (y | (y << 4)) & 0x0f0fu; y = (y | (y << 2)) & 0x3333u; y = (y | (y << 1)) & 0x5555u;
    return x | (y << 1);
}

__global__ void pv_sched_hist_kernel(const pv_sched_params p) {
    const long long c0 = p.sums[0] / p.B, c1 = p.sums[1] / p.B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.B; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&p.bins[pv_sched_bucket(p, i, c0, c1)], 1u);
}

// exclusive scan of the histogram in place, one block of 1024 threads (64 bins each)
__global__ void __launch_bounds__(1024) pv_sched_scan_kernel(unsigned int* bins) {
    constexpr int PER = PV_SCHED_BINS / 1024;
    __shared__ unsigned int part[1024];
    unsigned int local[PER];
    unsigned int sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        local[k] = bins[threadIdx.x * PER + k];
        sum += local[k];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const unsigned int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = part[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        bins[threadIdx.x * PER + k] = run;
        run += local[k];
    }
}

__global__ void pv_sched_scatter_kernel(const pv_sched_params p) {
    const long long c0 = p.sums[0] / p.B, c1 = p.sums[1] / p.B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.B; i += (int64_t)gridDim.x * blockDim.x)
        p.index[atomicAdd(&p.bins[pv_sched_bucket(p, i, c0, c1)], 1u)] = (int32_t)i;
}
