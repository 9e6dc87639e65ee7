// Does FP64 mma.sync (DMMA m8n8k4) on B200 run beside the vector FP64 pipe, or on it?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_ubench dmma_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// MODE 0: DMMA only (NCH independent accumulators), 1: DFMA only, 2: both interleaved
template <int MODE, int NCH>
__global__ void k(double* out, int iters, double seed) {
    double c0[NCH], c1[NCH], f[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { c0[i] = seed + i; c1[i] = seed - i; f[i] = seed * 0.5 + i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 0.999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE != 1) dmma(c0[i], c1[i], a, b, c0[i], c1[i]);
            if (MODE != 0) f[i] = fma(f[i], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += c0[i] + c1[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NCH>
void run(const char* name, int warps_per_sm, int nsm, double* buf, double clk) {
    int threads = 256, bps = warps_per_sm * 32 / threads, iters = 1 << 13;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE, NCH><<<nsm * bps, threads>>>(buf, iters, 1.0 + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double cyc = best * 1e-3 * clk * 1e9;
    double warp_inst = (double)NCH * iters * warps_per_sm;      // per SM, of each kind
    double fma_dmma = (MODE != 1) ? warp_inst * 256 / cyc : 0, fma_dfma = (MODE != 0) ? warp_inst * 32 / cyc : 0;
    printf("%-10s chains %d warps/SM %2d: DMMA %7.2f FMA/clk/SM, DFMA %6.2f FMA/clk/SM, total %7.2f\n", name, NCH, warps_per_sm,
           fma_dmma, fma_dfma, fma_dmma + fma_dfma);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    double* buf; cudaMalloc(&buf, (size_t)nsm * 64 * 32 * 8 * 2);
    int ws[] = {8, 16, 32};
    for (int w : ws) {
        run<0, 4>("dmma", w, nsm, buf, clk); run<0, 8>("dmma", w, nsm, buf, clk);
        run<1, 8>("dfma", w, nsm, buf, clk);
        run<2, 4>("both", w, nsm, buf, clk); run<2, 8>("both", w, nsm, buf, clk);
    }
    return 0;
}
