// Microbenchmark (run under gpurun): does the issue rate of packed FFMA2 / FMUL2 depend on how many DISTINCT 64-bit
// register operands an instruction reads?  (Hypothesis behind the ~790 clocks per warp-step of K1p against 548 FMA-pipe
// clocks: three register-pair sources = six 32-bit register reads.)
// nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2_operands ffma2_operands.cu && ./ffma2_operands
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// MODE 0: r[j] = fma2(a[j], b[j], r[j])      three distinct register pairs per instruction, no operand shared with its neighbours
//      1: r[j] = fma2(r[j], b[j], C)         two distinct pairs + one pair shared by all (reuse cache)
//      2: r[j] = fma2(r[j], M, C)            one distinct pair (the round-1 microbenchmark)
//      3: r[j] = fma2(r[j], b[j], imm)       two distinct pairs + immediate
//      4: r[j] = mul2(r[j], b[j])            two distinct pairs
//      5: r[j] = fma2(a[j], bs, r[j])        two distinct pairs + broadcast scalar register
//      6: scalar FFMA r = fmaf(a[j], b[j], r[j]) three distinct scalar registers
template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, float* out, float seed) {
    u64 a[8], b[8], r[8];
    float fa[8], fb[8], fr[8];
    for (int j = 0; j < 8; ++j) {
        const float v = threadIdx.x * 1e-3f + j + seed;
        a[j] = pk(1.0f + 1e-4f * v, 1.0f - 1e-4f * v); b[j] = pk(0.999f + 1e-5f * v, 1.001f - 1e-5f * v); r[j] = pk(v, v + 1.f);
        fa[j] = 1.0f + 1e-4f * v; fb[j] = 0.999f + 1e-5f * v; fr[j] = v;
    }
    const float m = 1.0001f + seed;
    const u64 M = pk(m, m + 1e-4f), C = pk(0.5f + seed, 0.25f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) r[j] = fma2(a[j], b[j], r[j]);
            else if (MODE == 1) r[j] = fma2(r[j], b[j], C);
            else if (MODE == 2) r[j] = fma2(r[j], M, C);
            else if (MODE == 3) r[j] = fma2(r[j], b[j], pk(0.25f, 0.25f));
            else if (MODE == 4) r[j] = mul2(r[j], b[j]);
            else if (MODE == 5) r[j] = fma2(a[j], pk(m, m), r[j]);
            else if (MODE == 6) fr[j] = fmaf(fa[j], fb[j], fr[j]);
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r[j])); s += x + y + fr[j]; }
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
static void run(const char* name, double fma_cycles_per_round, int ctas_per_sm = 4) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 20000, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 128>>>(100, out, 0.f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 128>>>(iters, out, 0.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rounds = (double)grid * 4 * iters / (148.0 * 4);
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-64s %d CTA/SM  %.2f clk per 8 instructions  (%.2f clk each; FMA pipe alone %.0f)\n", name, ctas_per_sm, clk / rounds,
           clk / rounds / 8, fma_cycles_per_round);
}

int main() {
    run<0>("FFMA2 r = a*b + r   (3 distinct register pairs)", 16);
    run<1>("FFMA2 r = r*b + C   (2 distinct pairs + 1 shared pair)", 16);
    run<2>("FFMA2 r = r*M + C   (1 distinct pair + 2 shared)", 16);
    run<3>("FFMA2 r = r*b + imm (2 distinct pairs + immediate)", 16);
    run<4>("FMUL2 r = r*b       (2 distinct pairs)", 16);
    run<5>("FFMA2 r = a*bs + r  (2 distinct pairs + broadcast scalar)", 16);
    run<6>("FFMA  r = a*b + r   (3 distinct scalar registers)", 8);
    run<0>("FFMA2 r = a*b + r   (3 distinct register pairs)", 16, 8);
    return 0;
}
