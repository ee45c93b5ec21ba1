// Microbenchmark (run under gpurun): issue rate of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>   // 0: scalar FFMA, 8 chains; 1: FFMA2, 8 chains (16 FMAs per round); 2: FFMA2 with a broadcast scalar operand
__global__ void __launch_bounds__(128) k(int iters, float* out, float seed) {
    float a[8]; u64 p[8];
    const float m = 1.0001f + seed, c = 0.5f;
    for (int j = 0; j < 8; ++j) { a[j] = threadIdx.x * 1e-3f + j; p[j] = pk(a[j], a[j] + 1.f); }
    const u64 M = pk(m, m + 1e-4f), C = pk(c, c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = fmaf(a[j], m, c);
            else if (MODE == 1) p[j] = fma2(p[j], M, C);
            else p[j] = fma2(p[j], pk(m, m), pk(0.25f, 0.25f));
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[j])); s += a[j] + x + y; }
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
static void run(const char* name, double fma_per_instr) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 20000, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 128>>>(100, out, 0.f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 128>>>(iters, out, 0.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)grid * 4 /*warps*/ * iters * 8;
    const double per_smsp_clk = instr / (148.0 * 4) / (ms * 1e-3 * 1.965e9);
    printf("%-28s %.3f ms  %.3f warp-instr/clk/SMSP (at 1965 MHz)  %.1f TFLOP/s\n", name, ms, per_smsp_clk,
           instr * 32 * fma_per_instr * 2 / (ms * 1e-3) / 1e12);
}

int main() {
    run<0>("FFMA (3-reg... imm c)", 1);
    run<1>("FFMA2 (3 packed regs)", 2);
    run<2>("FFMA2 (broadcast + imm)", 2);
    return 0;
}
