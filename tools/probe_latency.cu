// Dependent-issue latencies on the GPU at hand (cycles per operation of ONE warp's dependent chain): the numbers behind the
// small-batch analysis in DESIGN.md (a one-wave launch lasts as long as its longest trajectory's dependent chain).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_latency tools/probe_latency.cu && build/probe_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, int n, double a, double b) {
    double x = threadIdx.x * 1e-3 + 1.0, y = x + 0.5;
    float f = (float)x;
    __shared__ double sh[64];
    sh[threadIdx.x] = x;
    sh[threadIdx.x + 32] = y;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) x = fma(x, a, b);
            if (OP == 1) x = x * a;
            if (OP == 2) x = x + b;
            if (OP == 3) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
            if (OP == 4) f = __log2f(f) + 3.0f;
            if (OP == 5) f = fmaf(f, (float)a, (float)b);
            if (OP == 6) x = (double)(float)x + b;                       // F2F down + up + DADD
            if (OP == 7) x = sh[(__double2loint(x) & 31)] + b;           // LDS + DADD (address from the value)
            if (OP == 8) f = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(f))) + 1.0f;
            if (OP == 9) { x = fma(x, a, b); y = fma(y, a, b); }         // two independent chains
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + y + f;
}

template <int OP>
static void run(const char* name, int per_iter = 8) {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 32 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int n = 20000;
    chain<OP><<<1, 32>>>(out, cyc, n, 0.999999, 1e-7);
    chain<OP><<<1, 32>>>(out, cyc, n, 0.999999, 1e-7);
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %7.2f cycles per operation\n", name, (double)h / ((double)n * per_iter));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("DFMA dependent chain");
    run<1>("DMUL dependent chain");
    run<2>("DADD dependent chain");
    run<3>("SHFL of a double (2 x SHFL.IDX) chain");
    run<4>("MUFU.LG2 + FADD chain");
    run<5>("FFMA dependent chain");
    run<6>("F2F.F32.F64 + F2F.F64.F32 + DADD chain");
    run<7>("LDS.64 (dependent address) + DADD chain");
    run<8>("REDUX.MAX + FADD chain");
    run<9>("2 independent DFMA chains (per pair)");
    return 0;
}
