// Microbenchmark (run under gpurun): issue rates of scalar FFMA, packed FFMA2 / FMUL2 / FADD2 (fma/mul/add.rn.f32x2) and
// mixes with MUFU / ALU work on sm_100a -- what bounds the packed look-back / look-ahead steps (K1p, K2p).
// nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rcpa(float a) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// MODE 0: scalar FFMA x8 | 1: FFMA2 x8 | 2: FFMA2 broadcast+imm x8 | 3: FMUL2 x8 | 4: FADD2 x8
//      5: FFMA2 x8 + FMNMX x4 (ALU pipe beside the FMA pipe) | 6: FFMA2 x8 + MUFU.RCP x2 | 7: FFMA2 x4 + scalar FFMA x4 (+ FMUL x4)
//      8: ONE dependent FFMA2 chain (latency) | 9: two dependent FFMA2 chains
template <int MODE>
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
            else if (MODE == 2) p[j] = fma2(p[j], pk(m, m), pk(0.25f, 0.25f));
            else if (MODE == 3) p[j] = mul2(p[j], M);
            else if (MODE == 4) p[j] = add2(p[j], C);
            else if (MODE == 5) { p[j] = fma2(p[j], M, C); if (j < 4) a[j] = fminf(a[j], fmaxf(a[j + 4], m)); }
            else if (MODE == 6) { p[j] = fma2(p[j], M, C); if (j < 2) a[j] = rcpa(a[j]); }
            else if (MODE == 7) { if (j < 4) p[j] = fma2(p[j], M, C); else { a[j] = fmaf(a[j], m, c); a[j - 4] = a[j - 4] * m; } }
            else if (MODE == 8) { if (j == 0) p[0] = fma2(p[0], M, C); }
            else if (MODE == 9) { if (j < 2) p[j] = fma2(p[j], M, C); }
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[j])); s += a[j] + x + y; }
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
static void run(const char* name, double instr_per_round, double fma_cycles_per_round, int ctas_per_sm = 8) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 20000, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 128>>>(100, out, 0.f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 128>>>(iters, out, 0.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rounds = (double)grid * 4 /*warps*/ * iters / (148.0 * 4);      // warp-rounds per SMSP
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-44s %d CTA/SM %.3f ms  %.3f instr/clk/SMSP  FMA pipe %.0f %% (cycles needed / elapsed)  %.2f clk per round\n", name,
           ctas_per_sm, ms, rounds * instr_per_round / clk, 100.0 * rounds * fma_cycles_per_round / clk, clk / rounds);
}

int main() {
    run<0>("FFMA x8", 8, 8);
    run<1>("FFMA2 x8 (3 packed regs)", 8, 16);
    run<2>("FFMA2 x8 (broadcast + imm)", 8, 16);
    run<3>("FMUL2 x8", 8, 16);
    run<4>("FADD2 x8", 8, 16);
    run<5>("FFMA2 x8 + FMNMX x8 (ALU)", 16, 16);
    run<6>("FFMA2 x8 + MUFU.RCP x2", 10, 16);
    run<7>("FFMA2 x4 + FFMA x4 + FMUL x4", 12, 16);
    run<1>("FFMA2 x8, 3 CTAs/SM (12 warps)", 8, 16, 3);
    run<1>("FFMA2 x8, 2 CTAs/SM (8 warps)", 8, 16, 2);
    run<1>("FFMA2 x8, 1 CTA/SM (4 warps)", 8, 16, 1);
    run<8>("1 dependent FFMA2 chain, 1 CTA/SM (latency)", 1, 2, 1);
    run<9>("2 dependent FFMA2 chains, 1 CTA/SM", 2, 4, 1);
    run<9>("2 dependent FFMA2 chains, 3 CTAs/SM", 2, 4, 3);
    run<9>("2 dependent FFMA2 chains, 4 CTAs/SM", 2, 4, 4);
    return 0;
}
