// Micro-benchmark of the IC inner loop body (lhm_step.cuh phase_ic_R, unclamped part) on B200: DFMA/clk/SM as a
// function of resident warps per SM and pairs per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int PT, int UNR>
__global__ void k(const double* __restrict__ coef, double* out, int Ntar, int reps) {
    __shared__ __align__(16) double tri[4 * 128];
    for (int i = threadIdx.x; i < 4 * Ntar; i += blockDim.x) tri[i] = 1.0 / (1.0 + i);
    __syncthreads();
    double a1[PT], a2[PT], a3[PT], G1[PT], acc[PT];
#pragma unroll
    for (int r = 0; r < PT; ++r) {
        const double* c = coef + (size_t)(threadIdx.x * PT + r) * 4;
        a1[r] = c[0]; a2[r] = c[1]; a3[r] = c[2]; G1[r] = c[3]; acc[r] = 0.0;
    }
    const double2* tri2 = reinterpret_cast<const double2*>(tri);
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNR
        for (int i = 0; i < Ntar; ++i) {
            const double2 u = tri2[2 * i];
            const double2 v = tri2[2 * i + 1];
#pragma unroll
            for (int r = 0; r < PT; ++r) {
                double Fv = fma(a1[r], u.x, G1[r]);
                Fv = fma(a2[r], u.y, Fv);
                Fv = fma(a3[r], v.y, Fv);
                acc[r] = fma(v.x, Fv, acc[r]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < PT; ++r) s += acc[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same body with the next nu_i's operands prefetched one iteration ahead and the register budget of the fused kernel
// (256 threads, 2 CTAs/SM -> 128 registers): does a wider thread (8 pairs) keep the per-nu_i operand in the reuse cache
// for 7 of 8 DFMAs?
template <int PT>
__global__ void __launch_bounds__(256, 2) kpf(const double* __restrict__ coef, double* out, int Ntar, int reps) {
    __shared__ __align__(16) double tri[4 * 128];
    for (int i = threadIdx.x; i < 4 * Ntar; i += blockDim.x) tri[i] = 1.0 / (1.0 + i);
    __syncthreads();
    double a1[PT], a2[PT], a3[PT], G1[PT], acc[PT];
#pragma unroll
    for (int r = 0; r < PT; ++r) {
        const double* c = coef + (size_t)(threadIdx.x * PT + r) * 4;
        a1[r] = c[0]; a2[r] = c[1]; a3[r] = c[2]; G1[r] = c[3]; acc[r] = 0.0;
    }
    const double2* tri2 = reinterpret_cast<const double2*>(tri);
    for (int rep = 0; rep < reps; ++rep) {
        double2 u = tri2[0], v = tri2[1];
#pragma unroll 1
        for (int i = 0; i < Ntar; ++i) {
            const double2 un = tri2[2 * (i + 1)], vn = tri2[2 * (i + 1) + 1];      // tri has 128 rows: in bounds
            double Fv[PT];
#pragma unroll
            for (int r = 0; r < PT; ++r) Fv[r] = fma(a1[r], u.x, G1[r]);
#pragma unroll
            for (int r = 0; r < PT; ++r) Fv[r] = fma(a2[r], u.y, Fv[r]);
#pragma unroll
            for (int r = 0; r < PT; ++r) Fv[r] = fma(a3[r], v.y, Fv[r]);
#pragma unroll
            for (int r = 0; r < PT; ++r) acc[r] = fma(v.x, Fv[r], acc[r]);
            u = un; v = vn;
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < PT; ++r) s += acc[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int PT>
void runpf(int nsm, const double* coef, double* buf, double clk) {
    int Ntar = 96, reps = 200;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        kpf<PT><<<nsm * 2, 256>>>(coef, buf, Ntar, reps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double inst = 4.0 * PT * Ntar * reps * 16 * 32;
    printf("prefetch form, PT %d, 2 CTAs x 8 warps per SM: %6.2f DFMA/clk/SM\n", PT, inst / (best * 1e-3 * clk * 1e9));
}

template <int PT, int UNR>
void run(int warps_per_sm, int nsm, const double* coef, double* buf, double clk) {
    int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    int bps = warps_per_sm * 32 / threads;
    int Ntar = 96, reps = 200;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<PT, UNR><<<nsm * bps, threads>>>(coef, buf, Ntar, reps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double inst = 4.0 * PT * Ntar * reps * warps_per_sm * 32;
    double per_clk = inst / (best * 1e-3 * clk * 1e9);
    printf("PT %d unroll %d warps/SM %2d : %6.2f DFMA/clk/SM (%.2f cycles per warp-DFMA per warp)\n", PT, UNR, warps_per_sm,
           per_clk, warps_per_sm * 32.0 / per_clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    double *coef, *buf;
    cudaMalloc(&coef, 256 * 16 * 4 * 8); cudaMemset(coef, 0, 256 * 16 * 4 * 8);
    cudaMalloc(&buf, (size_t)nsm * 64 * 32 * 8 * 2);
    int ws[] = {4, 8, 16};
    for (int w : ws) {
        run<2, 2>(w, nsm, coef, buf, clk); run<4, 1>(w, nsm, coef, buf, clk); run<4, 2>(w, nsm, coef, buf, clk);
        run<4, 4>(w, nsm, coef, buf, clk); run<6, 2>(w, nsm, coef, buf, clk); run<8, 1>(w, nsm, coef, buf, clk); run<8, 2>(w, nsm, coef, buf, clk);
    }
    runpf<4>(nsm, coef, buf, clk); runpf<6>(nsm, coef, buf, clk); runpf<8>(nsm, coef, buf, clk); runpf<10>(nsm, coef, buf, clk);
    return 0;
}
