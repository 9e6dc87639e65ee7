// Cycles of the warp back-substitution of the implicit solves (lhm_step.cuh::backsub_warp) and of the coefficient pass
// that precedes it, on shared-memory arrays, for the sizes of BASELINE configs 1, 2 and 4.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../lehamoc_b200/csrc -o backsub_ubench backsub_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "lhm_step.cuh"

using namespace lhm;

// candidate: chunk rows staged in registers four at a time (loads no longer sit on the dependent chain)
__device__ __forceinline__ void backsub_warp_v1(const double* __restrict__ cp, const double* __restrict__ dp,
                                                double* __restrict__ x, int n) {
    const int lane = threadIdx.x & 31;
    const int Cn = (n + 31) >> 5;
    const int s = min(lane * Cn, n), e = min(s + Cn, n);
    double D = 0.0, M = 1.0;
    for (int i0 = e - 1; i0 >= s; i0 -= 4) {
        double c[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 - u;
            const bool on = i >= s;
            c[u] = on ? ((i == n - 1) ? 0.0 : cp[i]) : -1.0;     // (c, d) = (-1, 0) is the identity map
            d[u] = on ? dp[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { D = fma(-c[u], D, d[u]); M = -c[u] * M; }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double Dn = __shfl_down_sync(0xffffffffu, D, d), Mn = __shfl_down_sync(0xffffffffu, M, d);
        if (lane + d < 32) { D = fma(M, Dn, D); M *= Mn; }
    }
    double xv = __shfl_down_sync(0xffffffffu, D, 1);
    if (lane == 31) xv = 0.0;
    for (int i0 = e - 1; i0 >= s; i0 -= 4) {
        double c[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 - u;
            const bool on = i >= s;
            c[u] = on ? ((i == n - 1) ? 0.0 : cp[i]) : 0.0;
            d[u] = on ? dp[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 - u;
            if (i >= s) { xv = (i == n - 1) ? d[u] : d[u] - c[u] * xv; x[i] = xv; }
        }
    }
}

template <int V>
__global__ void __launch_bounds__(256, 2) k(int n, int iters, long long* cyc, double* out) {
    __shared__ double cp[1024], dp[1024], x[1024], t1[1024], t2[1024];
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += 256) { cp[i] = -0.3 - 1e-3 * (i % 7); dp[i] = 1.0 + 1e-3 * i; t1[i] = 1.0 + 1e-4 * i; t2[i] = 0.1 + 1e-5 * i; }
    __syncthreads();
    long long tb = 0, tc = 0;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        long long t0 = clock64();
        for (int j = tid; j < n; j += 256) {            // two divisions per element, like phase_electrons
            double V2 = 1.0 + 1e-3 * (t1[j] + t2[j] * x[j]);
            double V3 = -1e-3 * t2[j];
            cp[j] = V3 / V2;
            dp[j] = (x[j] + t1[j]) / V2;
        }
        __syncthreads();
        long long t1c = clock64();
        if (V == 0) { if (tid < 32) backsub_warp(cp, dp, x, n); }
        else if (V == 1) { if (tid < 32) backsub_warp_v1(cp, dp, x, n); }
        __syncthreads();
        long long t2c = clock64();
        tc += t1c - t0; tb += t2c - t1c;
    }
    if (tid == 0 && blockIdx.x == 0) { cyc[0] = tc / iters; cyc[1] = tb / iters; }
    if (tid == 0) out[blockIdx.x] = x[0] + x[n - 1];
}

int main() {
    long long* cyc; double* out;
    cudaMalloc(&cyc, 16); cudaMalloc(&out, 8 * 296);
    for (int n : {98, 158, 199, 258, 398, 798}) {
        long long h[2], h1[2], h2[2];
        k<0><<<296, 256>>>(n, 200, cyc, out);
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        k<1><<<296, 256>>>(n, 200, cyc, out);
        cudaMemcpy(h1, cyc, 16, cudaMemcpyDeviceToHost);
        k<2><<<296, 256>>>(n, 200, cyc, out);
        cudaMemcpy(h2, cyc, 16, cudaMemcpyDeviceToHost);
        printf("n = %4d: coefficient pass %5lld cycles, backsub_warp %5lld, register-staged %5lld, barrier + clock alone %5lld cycles "
               "(296 CTAs of 256 threads, 2 per SM)\n", n, h[0], h[1], h1[1], h2[1]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
