// FP64 pipe micro-benchmark for B200: DFMA throughput per SM as a function of independent chains per thread (ILP)
// and resident warps per SM (TLP).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(double* out, int iters, double seed) {
    double a[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) a[c] = seed + c;
    const double m = 0.999999, b = 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) a[c] = fma(a[c], m, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(int warps_per_sm, int nsm, double* buf, double clk_ghz) {
    int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    int blocks_per_sm = warps_per_sm * 32 / threads;
    int iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<CH><<<nsm * blocks_per_sm, threads>>>(buf, iters, 1.0 + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double inst = (double)CH * iters * warps_per_sm * 32;      // thread-level DFMA per SM
    double per_clk = inst / (best * 1e-3 * clk_ghz * 1e9);
    printf("chains %2d warps/SM %2d : %6.2f DFMA/clk/SM  (%.2f cycles per warp-DFMA per warp)\n", CH, warps_per_sm, per_clk,
           warps_per_sm * 32.0 / per_clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    double clk = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, nsm, clk);
    double* buf; cudaMalloc(&buf, (size_t)nsm * 64 * 32 * 8 * 2);
    int ws[] = {4, 8, 16, 32, 64};
    for (int w : ws) { run<1>(w, nsm, buf, clk); run<2>(w, nsm, buf, clk); run<4>(w, nsm, buf, clk); run<8>(w, nsm, buf, clk); run<16>(w, nsm, buf, clk); }
    return 0;
}
