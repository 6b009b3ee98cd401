// Locality schedule (divergence control): before a batch launch the trajectories are ordered by their parameter values
// so that the 32 lanes of a warp integrate NEIGHBOURING points of parameter space.  Neighbours take nearly the same step
// sequence, so a warp's lanes accept / reject together, cross the output grid in the same iterations (the per-grid-point
// code then runs with most lanes active instead of ~7 of 32) and finish together; the lockstep kernel
// (pv_batch_lockstep_kernel) goes one step further and gives them one common step size.
//
// Key: per column the top bits of the IEEE pattern of the value (monotone in the value for positive numbers:
// 2^(52 - PV_SCHED_SHIFT) levels per octave), centred on the batch mean, clamped to 8 bits; the first two columns are
// Morton-interleaved into a 16-bit key.  Sort: hand-written STABLE least-significant-digit radix sort, two passes of 8
// bits (per-warp tile histogram -> two-level exclusive scan -> stable per-warp scatter ranked with __match_any_sync), so the work
// order of a launch is a pure function of its inputs: a given batch always forms the same warps.  ~0.05 ms for 10^6
// trajectories.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PV_SCHED_SHIFT 47          // 5 mantissa bits: 32 levels per octave, 8 octaves around the batch mean
#define PV_SCHED_MAX_TILES 2048    // warps of the sort kernels = contiguous tiles of the input

struct pv_sched_params {
    const double* P;               // [n_cols][B]
    int64_t B;
    int n_key_cols;                // 1 or 2
    long long* sums;               // [2] sum of the raw keys per column
    uint32_t* key_a;               // [B] 16-bit sort keys (ping)
    uint32_t* key_b;               // [B] (pong)
    int32_t* idx_a;                // [B] (ping)
    int32_t* idx_b;                // [B] (pong) ; the final order lands in idx_a again after two passes
    uint32_t* counts;              // [256][n_tiles] digit-major tile histograms, scanned in place (per digit row)
    uint32_t* digit_total;         // [256] row totals
    uint32_t* digit_base;          // [256] exclusive scan of the totals
    int n_tiles;
    int64_t tile;                  // elements per tile
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
    __shared__ long long part[2][8];
    if ((threadIdx.x & 31) == 0) {
        part[0][threadIdx.x >> 5] = s0;
        part[1][threadIdx.x >> 5] = s1;
    }
    __syncthreads();
    if (threadIdx.x < 2 && threadIdx.x < p.n_key_cols) {     // integer sums: the result does not depend on the order of the atomics
        long long tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[threadIdx.x][w];
        atomicAdd((unsigned long long*)&p.sums[threadIdx.x], (unsigned long long)tot);
    }
}

__global__ void pv_sched_key_kernel(const pv_sched_params p) {
    const long long c0 = p.sums[0] / p.B, c1 = p.sums[1] / p.B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.B; i += (int64_t)gridDim.x * blockDim.x) {
        long long k0 = pv_sched_raw_key(p.P[i]) - c0 + 128;
        k0 = k0 < 0 ? 0 : (k0 > 255 ? 255 : k0);
        unsigned key = (unsigned)k0 << 8;
        if (p.n_key_cols > 1) {
            long long k1 = pv_sched_raw_key(p.P[p.B + i]) - c1 + 128;
            k1 = k1 < 0 ? 0 : (k1 > 255 ? 255 : k1);
            unsigned x = (unsigned)k0, y = (unsigned)k1;   // Morton interleave of two 8-bit numbers
            x = (x | (x << 4)) & 0x0f0fu; x = (x | (x << 2)) & 0x3333u; x = (x | (x << 1)) & 0x5555u;
            y = (y | (y << 4)) & 0x0f0fu; y = (y | (y << 2)) & 0x3333u; y = (y | (y << 1)) & 0x5555u;
            key = x | (y << 1);
        }
        p.key_a[i] = key;
        p.idx_a[i] = (int32_t)i;
    }
}

// one warp per tile: histogram of the 8-bit digit at `shift`
__global__ void __launch_bounds__(128) pv_sched_hist_kernel(const pv_sched_params p, const uint32_t* __restrict__ keys, int shift) {
    __shared__ uint32_t cnt[4][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_id = blockIdx.x * 4 + w;
    for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    if (tile_id < p.n_tiles) {
        const int64_t lo = (int64_t)tile_id * p.tile, hi = min(lo + p.tile, p.B);
        for (int64_t i = lo + lane; i < hi; i += 32) atomicAdd(&cnt[w][(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) p.counts[(int64_t)d * p.n_tiles + tile_id] = cnt[w][d];
    }
}

// exclusive scan of the digit-major histogram in two parallel levels: (1) one block per digit scans that digit's row of
// tile counts in place and leaves the row total in digit_total[d]; (2) one warp scans the 256 totals into digit_base[d].
// The scatter adds the two.  (A single-block scan of the whole array was latency bound: 46 us per pass.)
__global__ void __launch_bounds__(256) pv_sched_rowscan_kernel(uint32_t* counts, int n_tiles, uint32_t* digit_total) {
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_sh;
    uint32_t* row = counts + (int64_t)blockIdx.x * n_tiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int k0 = 0; k0 < n_tiles; k0 += 256) {
        const int k = k0 + threadIdx.x;
        const uint32_t c = k < n_tiles ? row[k] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += u;
        }
        if (lane == 31) warp_sum[w] = incl;
        __syncthreads();
        uint32_t base = carry_sh;
        for (int v = 0; v < w; ++v) base += warp_sum[v];
        if (k < n_tiles) row[k] = base + incl - c;
        __syncthreads();
        if (threadIdx.x == 255) carry_sh = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_sh;
}

__global__ void __launch_bounds__(32) pv_sched_digitscan_kernel(const uint32_t* digit_total, uint32_t* digit_base) {
    const int lane = threadIdx.x;
    uint32_t run = 0;
    for (int k0 = 0; k0 < 256; k0 += 32) {
        const uint32_t c = digit_total[k0 + lane];
        uint32_t incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += u;
        }
        digit_base[k0 + lane] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// one warp per tile: stable scatter (elements of a tile in input order; equal digits keep their order)
__global__ void __launch_bounds__(128) pv_sched_scatter_kernel(const pv_sched_params p, const uint32_t* __restrict__ keys_in,
                                                               const int32_t* __restrict__ idx_in, uint32_t* __restrict__ keys_out,
                                                               int32_t* __restrict__ idx_out, int shift) {
    __shared__ uint32_t run[4][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_id = blockIdx.x * 4 + w;
    if (tile_id >= p.n_tiles) return;
    for (int d = lane; d < 256; d += 32) run[w][d] = p.digit_base[d] + p.counts[(int64_t)d * p.n_tiles + tile_id];
    __syncwarp();
    const int64_t lo = (int64_t)tile_id * p.tile, hi = min(lo + p.tile, p.B);
    for (int64_t i0 = lo; i0 < hi; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool have = i < hi;
        const uint32_t key = have ? keys_in[i] : 0u;
        const uint32_t d = have ? ((key >> shift) & 255u) : 256u + lane;     // idle lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t off = 0;
        if (have && lane == leader) {
            off = run[w][d];
            run[w][d] = off + __popc(peers);
        }
        off = __shfl_sync(0xffffffffu, off, leader);
        if (have) {
            keys_out[off + rank] = key;
            idx_out[off + rank] = idx_in[i];
        }
        __syncwarp();
    }
}
