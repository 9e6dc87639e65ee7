// lhm_icg.cuh -- IC emissivity, path G: the shared-grid dense contraction on the FP64 tensor cores (SURVEY.md 8f-3,
// Q_IC_space_optimized LeHaMoC_f.py:113-124).
//
// When every walker of a batch has the same g_el / nu_ic / nu_tot (fixed-grid parameter sweeps, BASELINE configs 1 and
// 5) the clamped kernel F+[o,gamma,i] = max(F, 0) depends on the grids only, and the emissivity of walker w is
//     Q[w,o] = 3/4 sigma_T c  sum_gamma av[w,gamma]  sum_i F+[o,gamma,i] p[w,i]
// -- a GEMM over walkers for every (o, gamma) pair followed by a weighted sum over gamma.  F+ is still never
// materialised: a warp forms an 8 (nu_out) x 4 (nu_in) fragment of it in registers (3 DFMA + an integer clamp per
// thread, exactly path R's regrouping G1 + a1 x + a2 x 2ln nu + a3 x^2) and feeds it as the A operand of
// mma.sync.m8n8k4.f64 against the photon fields of 8 walkers at a time; one A fragment serves ICG_NB such walker
// blocks, so the FP64 pipe spends (3 + 8 ICG_NB) / (8 ICG_NB) slots per walker-element instead of path R's 4.
// The time loop runs in lock-step for the whole batch: per timestep one launch of the per-walker phases (stepG_kernel
// in lhm.cu, state parked in HBM in between) and one launch of this contraction.
//
// Work decomposition: a CTA owns ICG_NBW walkers (their p[w,i] and av[w,gamma] staged in shared memory) and every S-th
// task of a cost-sorted task list; a task is (group of 8 nu_out) x (chunk of ICG_GC gammas) and is taken by one warp,
// which walks the chunk two gammas at a time (two A fragments share every B fragment).  Each task writes its own
// slot of the partial-sum buffer, so the result does not depend on which warp ran it: deterministic without atomics.
#pragma once
#include "lhm_device.cuh"
#include "lhm_step.cuh"

namespace lhm {

constexpr int ICG_NB = 4;                 // 8-walker blocks per A fragment
constexpr int ICG_NBW = 8 * ICG_NB;       // walkers per CTA
constexpr int ICG_GC = 16;                // gammas per task
constexpr int ICG_NT = 256;               // threads per CTA
constexpr int ICG_STAGE_BYTES = (ICG_NT / 32) * 2 * 32 * 16;   // per warp: two stages of the 2 x 8 x 2 double2 pair constants of a gamma pair
constexpr int ICG_AVS = ICG_NBW + 2;      // row stride of the transposed av tile (16-byte aligned rows, spread banks)

struct IcgDims {
    int Ng, NgP;        // gammas, padded to a multiple of ICG_GC
    int No, NOG, NoPad; // outgoing frequencies nu_ic[0..Ni-2], groups of 8, padded
    int Ntar, KS, Kp;   // target frequencies nu_tot[0..Nt-2], k-steps of 4, padded length 4 KS
    int PST;            // row stride of the p tile in shared memory (PST % 16 == 4: conflict-free B fragments)
    int NGC;            // gamma chunks per group = NgP / ICG_GC
    int ntasks;         // NOG * NGC
    // Large grids: a CTA stages only a SECTION of the target frequencies (nks sections of KSs k-steps) and of the gammas
    // (ngs sections of GCs chunks), so that its tiles fit shared memory; every (task, k-section) has its own slot of
    // the partial-sum buffer.  nks = ngs = 1 up to grid ~200.
    int nks, KSs, ngs, GCs;
    int PSTs;           // row stride of the sectioned p tile
};

__host__ __device__ inline IcgDims icg_dims(int Ng, int Ni, int Nt) {
    IcgDims D;
    D.Ng = Ng; D.NgP = (Ng + ICG_GC - 1) / ICG_GC * ICG_GC;
    D.No = Ni - 1; D.NOG = (D.No + 7) / 8; D.NoPad = D.NOG * 8;
    D.Ntar = Nt - 1; D.KS = (D.Ntar + 3) / 4; D.Kp = 4 * D.KS;
    D.PST = D.Kp + ((4 - D.Kp % 16) + 16) % 16;
    D.NGC = D.NgP / ICG_GC;
    D.ntasks = D.NOG * D.NGC;
    D.nks = 1; D.KSs = D.KS; D.ngs = 1; D.GCs = D.NGC; D.PSTs = D.PST;
    return D;
}
__host__ __device__ inline size_t icg_smem_bytes(const IcgDims& D) {
    return ((size_t)ICG_NBW * D.PSTs + (size_t)D.GCs * ICG_GC * ICG_AVS + (size_t)D.KSs * 16) * sizeof(double) + 16
           + ICG_STAGE_BYTES;
}
// split the tiles until they fit `limit` bytes (k-sections first: they also shorten the p rows a warp strides over)
__host__ inline void icg_choose_sections(IcgDims& D, size_t limit) {
    for (int tries = 0; tries < 16 && icg_smem_bytes(D) > limit; ++tries) {
        const size_t pbytes = (size_t)ICG_NBW * D.PSTs, abytes = (size_t)D.GCs * ICG_GC * ICG_AVS;
        if (pbytes >= abytes && D.KSs > 8) D.nks *= 2; else if (D.GCs > 1) D.ngs *= 2; else D.nks *= 2;
        D.KSs = (D.KS + D.nks - 1) / D.nks;
        D.GCs = (D.NGC + D.ngs - 1) / D.ngs;
        const int kp = 4 * D.KSs;
        D.PSTs = kp + ((4 - kp % 16) + 16) % 16;
    }
    D.nks = (D.KS + D.KSs - 1) / D.KSs;                 // drop empty trailing sections
    D.ngs = (D.NGC + D.GCs - 1) / D.GCs;
}

// ---------------------------------------------------------------------------------------------------------------
// set-up, once per batch (grids only): pair constants in fragment order, the nu_in range every gamma pair has to
// visit, the per-nu_in operands, and the task list sorted by cost.  Reads walker 0's tables.
struct IcgSetupArgs {
    Layout L;
    IcgDims D;
    const double* tabd;     // walker 0's double tables (built by setup_kernel)
    double* ptg;            // [NOG][NgP][8][4]  {G1, a1, a2, a3}
    ushort2* kinfo;         // [NOG][NgP/2]      {first k-step, first k-step without clamp}
    double* xs;             // [Kp][4]           {1/nu_i, (1/nu_i) 2 ln nu_i, 1/nu_i^2, 0}
    int* tasks;             // [ntasks]          (og << 16) | gc, most expensive first
    int* cost;              // [ntasks] scratch
};

// part 1, any number of CTAs (grid-stride): the embarrassingly parallel tables
__global__ void __launch_bounds__(256) icg_setup_kernel(IcgSetupArgs A) {
    const Layout& L = A.L;
    const IcgDims& D = A.D;
    const double* T = A.tabd;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, NTH = gridDim.x * blockDim.x;
    const double* g = T + L.g;
    const double* ni = T + L.nic;
    const double* xi = T + L.invnt;
    for (int i = tid; i < D.Kp; i += NTH) {
        const bool in = i < D.Ntar;
        const double x = in ? xi[i] : 0.0;
        A.xs[4 * i + 0] = x;
        A.xs[4 * i + 1] = in ? x * T[L.l2nt + i] : 0.0;
        A.xs[4 * i + 2] = x * x;
        A.xs[4 * i + 3] = 0.0;
    }
    for (int e = tid; e < D.NOG * D.NgP * 8; e += NTH) {
        const int r = e & 7, j = (e >> 3) % D.NgP, og = (e >> 3) / D.NgP;
        const int o = og * 8 + r;
        const bool valid = o < D.No && j < D.Ng;
        const int oc = min(o, D.No - 1), jc = min(j, D.Ng - 1);
        double G1, a1, a2, a3;
        ic_pair_constants(valid, T[L.eps + oc], ni[oc], T[L.lnu4 + oc], g[jc], T[L.inv4g + jc], T[L.lg + jc], G1, a1, a2, a3);
        double* p = A.ptg + (size_t)e * 4;
        p[0] = G1; p[1] = a1; p[2] = a2; p[3] = a3;
    }
    // nu_in range per (group, gamma pair): same rule as path R's tiles (lhm_setup.cuh) -- F > 0 <=> nu_i > A
    for (int c = tid; c < D.NOG * (D.NgP / 2); c += NTH) {
        const int og = c / (D.NgP / 2), j0 = 2 * (c - og * (D.NgP / 2));
        double Amax = -1.0, Amin = 1e308;
        for (int r = 0; r < 8; ++r) {
            const int o = og * 8 + r;
            if (o >= D.No) break;
            const double eps = T[L.eps + o], nuo = ni[o];
            for (int jj = 0; jj < 2; ++jj) {
                const int j = j0 + jj;
                if (j < D.Ng && g[j] > eps) {
                    const double Av = nuo * (1.0 / (g[j] - eps)) * T[L.inv4g + j];
                    Amax = fmax(Amax, Av);
                    Amin = fmin(Amin, Av);
                }
            }
        }
        int ks0 = D.KS, ksc = D.KS;
        if (Amax >= 0.0) {
            int is = 0;
            while (is < D.Ntar && Amin * xi[is] >= 1.0) ++is;
            const int i0 = max(is - 1, 0);
            while (is < D.Ntar && Amax * xi[is] >= 1.0) ++is;
            const int i1 = min(is + 1, D.Ntar);
            ks0 = i0 / 4;
            ksc = (i1 + 3) / 4;
        }
        A.kinfo[c] = make_ushort2((unsigned short)ks0, (unsigned short)ksc);
    }
}

// part 2, one CTA: task costs from the ranges of part 1, then the rank sort
__global__ void __launch_bounds__(256) icg_tasks_kernel(IcgSetupArgs A) {
    const IcgDims& D = A.D;
    const int tid = threadIdx.x, NTH = blockDim.x;
    for (int t = tid; t < D.ntasks; t += NTH) {
        const int og = t / D.NGC, gc = t - og * D.NGC;
        int cst = 0;
        for (int pr = 0; pr < ICG_GC / 2; ++pr) {
            const ushort2 ki = A.kinfo[og * (D.NgP / 2) + gc * (ICG_GC / 2) + pr];
            cst += D.KS - ki.x;
        }
        A.cost[t] = cst;
    }
    __syncthreads();
    for (int t = tid; t < D.ntasks; t += NTH) {                      // rank sort, most expensive first
        const int c = A.cost[t];
        int rank = 0;
        for (int u = 0; u < D.ntasks; ++u) {
            const int cu = A.cost[u];
            rank += (cu > c) || (cu == c && u < t);
        }
        const int og = t / D.NGC, gc = t - og * D.NGC;
        A.tasks[rank] = (og << 16) | gc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct IcgArgs {
    IcgDims D;
    int W, Wpad, S;             // walkers, padded to ICG_NBW, task splits (CTAs per walker block)
    const double* ptg;
    const ushort2* kinfo;
    const double* xs;
    const int* tasks;
    const double* gp;           // [Wpad][Kp]   p[w,i] = trapezoid weight x photon density (zero for i >= Ntar)
    const double* gav;          // [Wpad][NgP]  av[w,gamma] = trapezoid weight x N_e/gamma  (zero padding)
    double* part;               // [nks][NGC][Wpad][NoPad] partial sums over one gamma chunk and one k-section
};

__global__ void __launch_bounds__(ICG_NT, 2) icg_kernel(IcgArgs A) {
    extern __shared__ __align__(16) double smg[];
    const IcgDims& D = A.D;
    double* ps = smg;                                   // [ICG_NBW][PSTs]
    double* avs = ps + (size_t)ICG_NBW * D.PSTs;        // [GCs * ICG_GC][ICG_AVS]
    double* xss = avs + (size_t)D.GCs * ICG_GC * ICG_AVS;   // [4 KSs][4]
    int* counter = reinterpret_cast<int*>(xss + (size_t)D.KSs * 16);
    const int tid = threadIdx.x, lane = tid & 31;
    // per-warp staging of the pair constants: the 512 bytes of the NEXT gamma pair travel from L2 by cp.async while the
    // current pair is contracted (a warp used to wait out an L2 round trip at the top of every pair)
    double2* stg = reinterpret_cast<double2*>(reinterpret_cast<char*>(counter) + 16) + (tid >> 5) * 64;
    const unsigned stg_s = (unsigned)__cvta_generic_to_shared(stg);
    // block -> (walker block, task split, gamma section, k-section)
    int b = blockIdx.x;
    const int ksec = b % D.nks; b /= D.nks;
    const int gsec = b % D.ngs; b /= D.ngs;
    const int wb = b / A.S, split = b - wb * A.S;
    const int w0 = wb * ICG_NBW;
    const int k0 = ksec * D.KSs, k1 = min(k0 + D.KSs, D.KS);            // k-steps of this section
    const int gc0 = gsec * D.GCs, gc1 = min(gc0 + D.GCs, D.NGC);        // gamma chunks of this section
    const int i0s = 4 * k0, ni = 4 * (k1 - k0), j0s = gc0 * ICG_GC, nj = (gc1 - gc0) * ICG_GC;
    for (int e = tid; e < ICG_NBW * ni; e += ICG_NT) {
        const int w = e / ni, i = e - w * ni;
        ps[w * D.PSTs + i] = __ldg(A.gp + (size_t)(w0 + w) * D.Kp + i0s + i);
    }
    for (int e = tid; e < ICG_NBW * nj; e += ICG_NT) {
        const int w = e / nj, j = e - w * nj;
        avs[j * ICG_AVS + w] = __ldg(A.gav + (size_t)(w0 + w) * D.NgP + j0s + j);
    }
    for (int e = tid; e < ni * 4; e += ICG_NT) xss[e] = __ldg(A.xs + (size_t)i0s * 4 + e);
    if (tid == 0) *counter = 0;
    __syncthreads();
    const int row = lane >> 2, kl = lane & 3;           // fragment coordinates: A[row][kl], B[kl][row], C[row][2 kl + {0,1}]
    const double2* xs2 = reinterpret_cast<const double2*>(xss);
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(counter, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        const int t = split + k * A.S;
        if (t >= D.ntasks) break;
        const int task = __ldg(A.tasks + t);
        const int og = task >> 16, gc = task & 0xffff;
        if (gc < gc0 || gc >= gc1) continue;            // another gamma section's task
        double run[ICG_NB][2];
#pragma unroll
        for (int nb = 0; nb < ICG_NB; ++nb) { run[nb][0] = 0.0; run[nb][1] = 0.0; }
        const ushort2* kin = A.kinfo + (size_t)og * (D.NgP / 2) + gc * (ICG_GC / 2);
        const double2* ptl = reinterpret_cast<const double2*>(A.ptg) + (size_t)(og * D.NgP + gc * ICG_GC) * 16 + lane;
        auto fetch = [&](int pr) {                      // lane l moves double2 l of the pair's 32
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(stg_s + (unsigned)(((pr & 1) * 32 + lane) * 16)),
                         "l"(ptl + (size_t)pr * 32));
        };
        __syncwarp();                                   // the last pair of the previous task has been read by every lane
        fetch(0);
        asm volatile("cp.async.commit_group;");
        ushort2 ki_next = __ldg(kin);
        for (int pr = 0; pr < ICG_GC / 2; ++pr) {
            const ushort2 ki = ki_next;
            __syncwarp();                               // every lane is done with the stage the next copy overwrites
            if (pr + 1 < ICG_GC / 2) { fetch(pr + 1); ki_next = __ldg(kin + pr + 1); }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 1;");
            __syncwarp();
            const int ks0 = max((int)ki.x, k0), ksc = min(max((int)ki.y, k0), k1);
            if (ks0 >= k1) continue;                    // no active pair among the 8 x 2 in this k-section
            const int j0 = (gc - gc0) * ICG_GC + 2 * pr;                      // gamma index inside the staged section
            const double2* cs = stg + (pr & 1) * 32 + row * 2;
            const double2 c0a = cs[0], c0b = cs[1], c1a = cs[16], c1b = cs[17];
            double C0[ICG_NB][2], C1[ICG_NB][2];
#pragma unroll
            for (int nb = 0; nb < ICG_NB; ++nb) { C0[nb][0] = C0[nb][1] = C1[nb][0] = C1[nb][1] = 0.0; }
            const double* pb = ps + row * D.PSTs + kl;
            int ks = ks0;
            for (; ks < ksc; ++ks) {                    // some pair may still have F <= 0 here: clamp
                const int i = 4 * (ks - k0) + kl;
                const double2 xa = xs2[2 * i];          // {1/nu_i, (1/nu_i) 2 ln nu_i}
                const double x2 = xss[4 * i + 2];
                double F0 = fma(c0a.y, xa.x, c0a.x), F1 = fma(c1a.y, xa.x, c1a.x);
                F0 = fma(c0b.x, xa.y, F0); F1 = fma(c1b.x, xa.y, F1);
                F0 = fma(c0b.y, x2, F0); F1 = fma(c1b.y, x2, F1);
                F0 = clamp_neg_to_zero(F0); F1 = clamp_neg_to_zero(F1);      // function_IC[function_IC < 0.] = 0.
#pragma unroll
                for (int nb = 0; nb < ICG_NB; ++nb) {
                    const double bv = pb[(size_t)(8 * nb) * D.PSTs + 4 * (ks - k0)];
                    dmma884(C0[nb][0], C0[nb][1], F0, bv);
                    dmma884(C1[nb][0], C1[nb][1], F1, bv);
                }
            }
#pragma unroll 2
            for (; ks < k1; ++ks) {                     // q < 1 for every active pair: F > 0, the clamp is the identity
                const int i = 4 * (ks - k0) + kl;
                const double2 xa = xs2[2 * i];
                const double x2 = xss[4 * i + 2];
                double F0 = fma(c0a.y, xa.x, c0a.x), F1 = fma(c1a.y, xa.x, c1a.x);
                F0 = fma(c0b.x, xa.y, F0); F1 = fma(c1b.x, xa.y, F1);
                F0 = fma(c0b.y, x2, F0); F1 = fma(c1b.y, x2, F1);
#pragma unroll
                for (int nb = 0; nb < ICG_NB; ++nb) {
                    const double bv = pb[(size_t)(8 * nb) * D.PSTs + 4 * (ks - k0)];
                    dmma884(C0[nb][0], C0[nb][1], F0, bv);
                    dmma884(C1[nb][0], C1[nb][1], F1, bv);
                }
            }
            // weighted sum over gamma: run[o, w] += av[w, gamma] * C[o, w]
            const double2* av0 = reinterpret_cast<const double2*>(avs + (size_t)j0 * ICG_AVS + 2 * kl);
            const double2* av1 = reinterpret_cast<const double2*>(avs + (size_t)(j0 + 1) * ICG_AVS + 2 * kl);
#pragma unroll
            for (int nb = 0; nb < ICG_NB; ++nb) {
                const double2 a0 = av0[4 * nb], a1 = av1[4 * nb];
                run[nb][0] = fma(a0.x, C0[nb][0], run[nb][0]);
                run[nb][1] = fma(a0.y, C0[nb][1], run[nb][1]);
                run[nb][0] = fma(a1.x, C1[nb][0], run[nb][0]);
                run[nb][1] = fma(a1.y, C1[nb][1], run[nb][1]);
            }
        }
        double* out = A.part + (((size_t)ksec * D.NGC + gc) * A.Wpad + w0) * D.NoPad + og * 8 + row;
#pragma unroll
        for (int nb = 0; nb < ICG_NB; ++nb) {
            out[(size_t)(8 * nb + 2 * kl) * D.NoPad] = run[nb][0];
            out[(size_t)(8 * nb + 2 * kl + 1) * D.NoPad] = run[nb][1];
        }
    }
}

// Q_IC[o] = 3/4 sigma_T c  sum over gamma chunks (fixed order) of the partial sums of walker w
template <int NT>
__device__ __forceinline__ void icg_finish(const double* __restrict__ part, const IcgDims& D, int Wpad, int w, double* QIC) {
    for (int o = threadIdx.x; o < D.No; o += NT) {
        double s = 0.0;
        const double* __restrict__ p = part + (size_t)w * D.NoPad + o;
        const size_t st = (size_t)Wpad * D.NoPad;
        const int nslots = D.nks * D.NGC;                   // (k-section, gamma chunk) slots, k-section-major
        int gc = 0;
        for (; gc + 8 <= nslots; gc += 8) {                 // eight loads in flight, added in the fixed slot order
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(gc + u) * st);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; gc < nslots; ++gc) s += __ldg(p + (size_t)gc * st);
        QIC[o] = kPC.ic_pref * s;
    }
}

}  // namespace lhm
