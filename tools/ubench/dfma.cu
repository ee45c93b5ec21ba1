// Microbenchmark (run under gpurun): latency and throughput of fp64 DFMA / DADD / DMUL and of a 64-bit shuffle on sm_100a
// -- what bounds the fp64 planner's per-step dependent chain.   nvcc -gencode arch=compute_100a,code=sm_100a -o dfma dfma.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int MODE, int CHAINS>
__global__ void __launch_bounds__(128) k(int iters, double* out, double seed) {
    double a[CHAINS];
    for (int j = 0; j < CHAINS; ++j) a[j] = threadIdx.x * 1e-3 + j + seed;
    const double m = 1.0000001 + seed, c = 0.5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) {
            if (MODE == 0) a[j] = __fma_rn(a[j], m, c);
            else if (MODE == 1) a[j] = __dadd_rn(a[j], c);
            else if (MODE == 2) a[j] = __dmul_rn(a[j], m);
            else if (MODE == 3) a[j] = __shfl_sync(0xffffffffu, a[j], (threadIdx.x + 1) & 31);
            else if (MODE == 4) a[j] = a[j] / m;
            else if (MODE == 5) a[j] = sqrt(a[j] + 2.0);
        }
    }
    double s = 0;
    for (int j = 0; j < CHAINS; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
}
template <int MODE, int CHAINS>
static void run(const char* name, int ctas_per_sm, int iters = 4000) {
    double* out; cudaMalloc(&out, 8);
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, CHAINS><<<grid, 128>>>(100, out, 0.0);
    cudaEventRecord(e0);
    k<MODE, CHAINS><<<grid, 128>>>(iters, out, 0.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-34s %d chain(s) x %d warp(s)/SMSP: %.1f clk per dependent op, %.3f warp-ops/clk/SMSP\n", name, CHAINS, ctas_per_sm,
           clk / iters, (double)ctas_per_sm * CHAINS * iters / clk);
}
int main() {
    run<0, 1>("DFMA latency", 1); run<0, 8>("DFMA throughput", 4);
    run<1, 1>("DADD latency", 1); run<2, 1>("DMUL latency", 1);
    run<3, 1>("SHFL.64 (2 x SHFL) latency", 1);
    run<4, 1>("fp64 division latency", 1, 1000); run<5, 1>("fp64 sqrt latency", 1, 1000);
    return 0;
}
