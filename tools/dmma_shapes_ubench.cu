// FP64 mma.sync shapes on B200: issue rate of m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 (independent accumulator chains).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_shapes_ubench dmma_shapes_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE> struct Frag;
template <> struct Frag<0> { static constexpr int NA = 1, NB = 1, NC = 2, FMA = 256;  static constexpr const char* name = "m8n8k4"; };
template <> struct Frag<1> { static constexpr int NA = 2, NB = 1, NC = 4, FMA = 512;  static constexpr const char* name = "m16n8k4"; };
template <> struct Frag<2> { static constexpr int NA = 4, NB = 2, NC = 4, FMA = 1024; static constexpr const char* name = "m16n8k8"; };
template <> struct Frag<3> { static constexpr int NA = 8, NB = 4, NC = 4, FMA = 2048; static constexpr const char* name = "m16n8k16"; };

template <int SHAPE>
__device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    if constexpr (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
    else if constexpr (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
    else if constexpr (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                     "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int NCH>
__global__ void k(double* out, int iters, double seed) {
    double c[NCH][4], a[8], b[4];
#pragma unroll
    for (int i = 0; i < NCH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = seed + i + 0.1 * j;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = 0.999999 - 1e-9 * j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) mma<SHAPE>(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NCH>
void run(int warps_per_sm, int nsm, double* buf, double clk) {
    int threads = 256, bps = warps_per_sm * 32 / threads, iters = 1 << 12;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<SHAPE, NCH><<<nsm * bps, threads>>>(buf, iters, 1.0 + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) { printf("%-9s launch failed\n", Frag<SHAPE>::name); return; }
    double cyc = best * 1e-3 * clk * 1e9;
    double warp_inst = (double)NCH * iters * warps_per_sm;
    printf("%-9s chains %d warps/SM %2d: %7.2f FMA/clk/SM  (%.1f cycles per warp instruction per SM sub-partition)\n",
           Frag<SHAPE>::name, NCH, warps_per_sm, warp_inst * Frag<SHAPE>::FMA / cyc, cyc / (warp_inst / 4));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    double* buf; cudaMalloc(&buf, (size_t)nsm * 64 * 32 * 8 * 2);
    int ws[] = {8, 16};
    for (int w : ws) {
        run<0, 8>(w, nsm, buf, clk); run<1, 4>(w, nsm, buf, clk); run<1, 8>(w, nsm, buf, clk);
        run<2, 4>(w, nsm, buf, clk); run<2, 8>(w, nsm, buf, clk); run<3, 4>(w, nsm, buf, clk); run<3, 8>(w, nsm, buf, clk);
    }
    return 0;
}
