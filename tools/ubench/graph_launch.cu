// Microbenchmark (run under gpurun): latency of a two-kernel dependent tick, host enqueue to host-visible result, as
// (a) two stream launches and (b) a CUDA graph of two kernel nodes whose parameters are patched every tick.
// Kernel A spins ~40 us (stands in for the scoring launch), kernel B ~8 us (the re-score) and writes a sequence word
// to mapped pinned memory that the host polls -- the structure of LookBack.push.
// nvcc -gencode arch=compute_100a,code=sm_100a -o graph_launch graph_launch.cu && ./graph_launch
#include <cuda_runtime.h>
#include <stdio.h>
#include <time.h>
#include <algorithm>
#include <vector>
struct Blob { float v[24]; int slot; };
struct Blob2 { double v[14]; int slot; };
__global__ void kA(Blob b, long long spin, unsigned long long* out) {
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (unsigned long long)(b.v[3] + b.slot);
}
__global__ void kB(Blob2 b, long long spin, const unsigned long long* in, volatile unsigned long long* host, unsigned long long seq) {
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (blockIdx.x == 0 && threadIdx.x == 0) { host[0] = in[0] + (unsigned long long)b.v[1]; __threadfence_system(); host[1] = seq; __threadfence_system(); }
}
static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    unsigned long long *dev, *host, *hostd;
    cudaMalloc(&dev, 64);
    cudaHostAlloc(&host, 64, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&hostd, host, 0);
    const long long spinA = 40 * 1965, spinB = 8 * 1965;
    Blob a = {}; Blob2 b = {};
    unsigned long long seq = 1;
    // graph
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphNode_t nA, nB;
    void* argsA[] = {&a, (void*)&spinA, &dev};
    const unsigned long long* devc = dev; volatile unsigned long long* hv = hostd;
    void* argsB[] = {&b, (void*)&spinB, &devc, &hv, &seq};
    cudaKernelNodeParams pA = {}; pA.func = (void*)kA; pA.gridDim = dim3(1024); pA.blockDim = dim3(128); pA.kernelParams = argsA;
    cudaKernelNodeParams pB = {}; pB.func = (void*)kB; pB.gridDim = dim3(16); pB.blockDim = dim3(128); pB.kernelParams = argsB;
    cudaGraphAddKernelNode(&nA, g, nullptr, 0, &pA);
    cudaGraphAddKernelNode(&nB, g, &nA, 1, &pB);
    cudaGraphExec_t ex; cudaGraphInstantiate(&ex, g, 0);
    for (int mode = 0; mode < 2; ++mode) {
        std::vector<double> enq, tot;
        for (int it = 0; it < 300; ++it) {
            a.slot = it; b.slot = it; ++seq; host[1] = 0;
            const double t0 = now();
            if (mode == 0) {
                kA<<<1024, 128, 0, st>>>(a, spinA, dev);
                kB<<<16, 128, 0, st>>>(b, spinB, dev, hostd, seq);
            } else {
                cudaGraphExecKernelNodeSetParams(ex, nA, &pA);
                cudaGraphExecKernelNodeSetParams(ex, nB, &pB);
                cudaGraphLaunch(ex, st);
            }
            const double t1 = now();
            while (((volatile unsigned long long*)host)[1] != seq) {}
            const double t2 = now();
            if (it >= 50) { enq.push_back(t1 - t0); tot.push_back(t2 - t0); }
        }
        std::sort(enq.begin(), enq.end()); std::sort(tot.begin(), tot.end());
        printf("%-28s enqueue p50 %.2f us   enqueue -> host-visible result p50 %.2f us (kernels spin 40 + 8 us)\n",
               mode == 0 ? "two stream launches" : "graph, 2 nodes re-parameterised", enq[enq.size() / 2], tot[tot.size() / 2]);
    }
    return 0;
}
