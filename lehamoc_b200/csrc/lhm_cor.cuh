// lhm_cor.cuh -- A5: cor_factor_syn_el (LeMoC/LeHaMoC_f.py:175-237), one WARP per walker.
//
// The reference runs a private synchrotron-only simulation (B = 1e4 G, 200 implicit steps of
// 0.1001 R0/c) and returns L_ph/L_e,inj at the end; its arguments are constants of a model, so it is
// evaluated once per walker instead of once per timestep (bit-identical).
//
// Two exact re-associations make it a latency-bound warp job instead of 5.1 M exponentials:
//  (1) the returned luminosity is a linear functional of the photon vector, so only
//      z_j = sum_k omega_k * 4 pi dt V * dQ_k/dN_j is needed (one 98 x Ng pass of exp) and each step
//      is  Lambda <- (Lambda + z.N) / (1 + dt c/R0)                      (SURVEY.md 7.1b-2);
//  (2) the bidiagonal solve x_j = dp_j - cp_j x_{j+1} has constant cp over the 200 steps, so it is an
//      affine suffix scan: per-lane chunk composition + 5 shuffle levels with precomputed products.
#pragma once
#include "lhm_device.cuh"
#include "../../include/lhm.h"

namespace lhm {


struct CorArgs {
    int W, Ng;
    const double* consts;   // [W, LHM_NCONST]
    const double* g_el;     // [W, Ng]
    double* out;            // [W]
};

constexpr int COR_WARPS = 4;            // warps (walkers) per CTA
constexpr int COR_NNU = 100;            // private frequency grid, LeHaMoC_f.py:188

// per-warp shared arrays, element e of lane l / slot i stored at [i*32 + l] (conflict-free)
__host__ __device__ inline int cor_chunk(int Ng) { return (Ng - 2 + 31) / 32; }
__host__ __device__ inline size_t cor_smem_bytes(int Ng) {
    // m, binv, injdt, Mp, N, z : 6 arrays of C*32 ; g copy (Ng) ; nu, cnu (2 x 100)
    return (size_t)COR_WARPS * ((size_t)6 * cor_chunk(Ng) * 32 + ((Ng + 1) & ~1) + 2 * COR_NNU) * sizeof(double);
}

// The 200-step loop with the lane's chunk of every array in registers (CN = chunk length, compile time): no shared-memory
// round trips inside the recurrences.  The luminosity functional Lambda <- (Lambda + z.N)/bph is linear, so every lane
// carries its own share of it and the lanes are only summed when the reference samples the ratio (every 10th step).
template <int CN>
__device__ __forceinline__ double cor_time_loop(const double* m_, const double* binv, const double* injdt, const double* Mp,
                                                const double* z, const double (&Mlev)[5], double lam_lane, double zb,
                                                double ends, double El, double dt, double R0, int lane) {
    const PhysConst& pc = kPC;
    double rm[CN], rb[CN], ri[CN], rM[CN], rz[CN], rN[CN];
#pragma unroll
    for (int i = 0; i < CN; ++i) {
        const int s = i * 32 + lane;
        rm[i] = m_[s]; rb[i] = binv[s]; ri[i] = injdt[s]; rM[i] = Mp[s]; rz[i] = z[s]; rN[i] = 0.0;
    }
    const double bph = 1.0 + dt * (pc.c / R0);               // photon equation is diagonal (:226)
    const double t_end = 20.0 * R0 / pc.c;
    double t = 0.0, day = 0.0, ratio = 0.0;
    while (t < t_end) {
        t += dt;
        double x0 = 0.0;
#pragma unroll
        for (int i = CN - 1; i >= 0; --i) {
            x0 = fma(rm[i], x0, (rN[i] + ri[i]) * rb[i]);
            rN[i] = x0;
        }
        double Aacc = x0;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
            const double Ao = __shfl_down_sync(0xffffffffu, Aacc, 1 << l);
            if (lane + (1 << l) < 32) Aacc = fma(Mlev[l], Ao, Aacc);
        }
        double xin = __shfl_down_sync(0xffffffffu, Aacc, 1);
        if (lane == 31) xin = 0.0;
        double zn = (lane == 0) ? zb : 0.0;
#pragma unroll
        for (int i = 0; i < CN; ++i) {
            const double xv = fma(rM[i], xin, rN[i]);
            rN[i] = xv;
            zn = fma(rz[i], xv, zn);
        }
        lam_lane = (lam_lane + zn) / bph;
        if (day < t) {                                       // (:231-236)
            day = day + 1.0 * R0 / pc.c;
            ratio = (warp_sum(lam_lane) + ends) / El;
        }
    }
    return ratio;
}

__global__ void __launch_bounds__(COR_WARPS * 32) cor_kernel(CorArgs A) {
    extern __shared__ double smc[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * COR_WARPS + warp;
    if (w >= A.W) return;                                    // whole warp exits together
    const int Ng = A.Ng, n = Ng - 2, Cn = cor_chunk(Ng), CS = Cn * 32;
    double* base = smc + (size_t)warp * (6 * CS + ((Ng + 1) & ~1) + 2 * COR_NNU);
    double* m_ = base;             // -cp
    double* binv = m_ + CS;
    double* injdt = binv + CS;
    double* Mp = injdt + CS;
    double* Nn = Mp + CS;
    double* z = Nn + CS;
    double* g = z + CS;
    double* nu = g + ((Ng + 1) & ~1);
    double* cnu = nu + COR_NNU;
    const PhysConst& pc = kPC;
    const double* cst = A.consts + (size_t)w * LHM_NCONST;
    const double R0 = cst[LHM_C_R0], B0 = cst[LHM_C_COR_B], p_el = cst[LHM_C_COR_P], Q0 = cst[LHM_C_COR_Q0];
    for (int j = lane; j < Ng; j += 32) g[j] = A.g_el[(size_t)w * Ng + j];
    __syncwarp();

    const double dt = 0.1001 * R0 / pc.c;                    // LeHaMoC_f.py:205
    const double Vol = volume_of(R0);
    const double b_syn = (4.0 / 3.0 * pc.sigmaT / (8.0 * M_PI * pc.m_el * pc.c)) * (B0 * B0);   // :184,212
    const double a_cr = 3.0 * pc.q * B0 / (4.0 * M_PI * pc.m_el * pc.c);                      // :210
    // frequency grid np.logspace(7.5, log10(7 nu_c(g[-1],B0))+1.4, 100)  (:188)
    const double la = 7.5, lb = log10(7.0 * (pc.nuc_pref * B0 * (g[Ng - 1] * g[Ng - 1]))) + 1.4;
    const double lstep = (lb - la) / (double)(COR_NNU - 1);
    for (int k = lane; k < COR_NNU; k += 32) nu[k] = exp10((k == COR_NNU - 1) ? lb : la + (double)k * lstep);
    __syncwarp();
    // omega_k = trapz weight over ln(nu) x h nu^2 (4 pi/3) R0^2 c / V   (:234,236)
    // c_k     = omega_k * 4 pi dt V * C_syn B^(2/3) nu^(-2/3)           (:91,228)
    double lam0_part = 0.0, ends = 0.0;
    for (int k = lane; k < COR_NNU; k += 32) {
        double lm = (k > 0) ? log(nu[k - 1]) : 0.0, lc = log(nu[k]), lp = (k < COR_NNU - 1) ? log(nu[k + 1]) : 0.0;
        double wk = (((k > 0) ? (lc - lm) : 0.0) + ((k < COR_NNU - 1) ? (lp - lc) : 0.0)) * 0.5;
        double om = wk * (pc.h * (nu[k] * nu[k])) * 4.0 * M_PI / 3.0 * (R0 * R0) * pc.c / Vol;
        if (k >= 1 && k <= COR_NNU - 2) {
            cnu[k] = om * 4.0 * M_PI * dt * Vol * (pc.C_syn * pow(B0, 2.0 / 3.0) * pow(nu[k], -2.0 / 3.0));
            lam0_part += om * LHM_SENT;                      // photons_syn starts at 1e-260 everywhere (:194)
        } else {
            cnu[k] = 0.0;
            ends += om * LHM_SENT;                           // boundary bins are never updated
        }
    }
    const double lam_lane0 = lam0_part;                      // this lane's share of the initial luminosity functional
    double Lam = warp_sum(lam0_part);
    ends = warp_sum(ends);
    __syncwarp();

    // static solver tables + z
    double El_part = 0.0, zb_part = 0.0;
    for (int j = lane; j < Ng; j += 32) {                    // L_e,inj = trapz(el_inj g^2 m c^2, ln g)  (:235)
        double lm = (j > 0) ? log(g[j - 1]) : 0.0, lc = log(g[j]), lp = (j < Ng - 1) ? log(g[j + 1]) : 0.0;
        double wj = (((j > 0) ? (lc - lm) : 0.0) + ((j < Ng - 1) ? (lp - lc) : 0.0)) * 0.5;
        double inj = Q0 * pow(g[j], -p_el);                  // el_inj on ALL bins (:202)
        El_part += wj * (inj * (g[j] * g[j]) * pc.mc2);
    }
    const double El = warp_sum(El_part);
    for (int i = 0; i < Cn; ++i) {
        const int e = lane * Cn + i;                         // interior unknown e <-> grid index e+1
        const int s = i * 32 + lane;
        double mv = 0.0, bi = 1.0, idt = 0.0, zv = 0.0;
        if (e < n) {
            const int j = e + 1;
            double mp0 = (g[j] + g[(e == 0) ? (Ng - 1) : (e - 1)]) / 2.0;   // g_el_mp[e]  (:182, wraps at e=0)
            double mp1 = (g[j + 1] + g[j - 1]) / 2.0;                         // g_el_mp[e+1]
            double dg = (g[j + 1] - g[j - 1]) / 2.0;                          // dg_el[e]    (:185)
            double V2 = 1.0 + dt * (pc.c / R0 + b_syn * ((mp0 * mp0) / dg));     // (:213,217)
            double V3 = -dt * (b_syn * ((mp1 * mp1) / dg));                       // (:214,218)
            bi = 1.0 / V2;
            mv = (e == n - 1) ? 0.0 : -(V3 / V2);            // last row's c never enters (:259)
            idt = (Q0 * pow(g[j], -p_el)) * dt;              // el_inj[1:-1]*dt (:219)
        }
        m_[s] = mv; binv[s] = bi; injdt[s] = idt; Nn[s] = 0.0; z[s] = zv;
    }
    __syncwarp();
    // z_j for every grid index (interior in z[], the two boundary bins folded into a constant)
    for (int jj = lane; jj < Ng; jj += 32) {
        const double gj = g[jj];
        const double inv = 1.0 / (a_cr * (gj * gj));
        const double thr = pc.mc2 * gj;
        double acc = 0.0;
        for (int k = 1; k <= COR_NNU - 2; ++k) {
            double x = nu[k] * inv;
            if (!(thr < pc.h * nu[k]) && x <= 745.2) acc = fma(cnu[k], exp(-x), acc);   // mask (:89-90)
        }
        double lm = (jj > 0) ? log(g[jj - 1]) : 0.0, lc = log(gj), lp = (jj < Ng - 1) ? log(g[jj + 1]) : 0.0;
        double wj = (((jj > 0) ? (lc - lm) : 0.0) + ((jj < Ng - 1) ? (lp - lc) : 0.0)) * 0.5;
        double zj = wj * pow(gj, 1.0 / 3.0) / Vol * acc;
        if (jj == 0 || jj == Ng - 1) zb_part += zj * 1e-160;      // N_el[0] = N_el[-1] = 1e-160 (:195)
        else { int e = jj - 1; z[(e % Cn) * 32 + (e / Cn)] = zj; }
    }
    const double zb = warp_sum(zb_part);
    __syncwarp();
    // partial products Mp_i = prod_{i'>=i} m_{i'} within the lane chunk, lane product, scan-level products
    double Mlane = 1.0;
    for (int i = Cn - 1; i >= 0; --i) { Mlane *= m_[i * 32 + lane]; Mp[i * 32 + lane] = Mlane; }
    double Mlev[5];
    {
        double M = Mlane;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
            Mlev[l] = M;
            double Mo = __shfl_down_sync(0xffffffffu, M, 1 << l);
            M = (lane + (1 << l) < 32) ? M * Mo : M;
        }
    }
    if (Cn <= 12) {
        __syncwarp();
        double r;
        switch (Cn) {
#define LHM_COR_CASE(n) case n: r = cor_time_loop<n>(m_, binv, injdt, Mp, z, Mlev, lam_lane0, zb, ends, El, dt, R0, lane); break;
            LHM_COR_CASE(1) LHM_COR_CASE(2) LHM_COR_CASE(3) LHM_COR_CASE(4) LHM_COR_CASE(5) LHM_COR_CASE(6)
            LHM_COR_CASE(7) LHM_COR_CASE(8) LHM_COR_CASE(9) LHM_COR_CASE(10) LHM_COR_CASE(11) default:
            LHM_COR_CASE(12)
#undef LHM_COR_CASE
        }
        if (lane == 0) A.out[w] = r;
        return;
    }
    // time loop (:204-236), chunk arrays in shared memory (grids beyond 12 * 32 + 2 points)
    const double bph = 1.0 + dt * (pc.c / R0);               // photon equation is diagonal (:226)
    const double t_end = 20.0 * R0 / pc.c;
    double t = 0.0, day = 0.0, ratio = 0.0;
    while (t < t_end) {
        t += dt;
        // chunk solution for zero inflow, from the top of the chunk down
        double x0 = 0.0;
        for (int i = Cn - 1; i >= 0; --i) {
            const int s = i * 32 + lane;
            double dpv = (Nn[s] + injdt[s]) * binv[s];
            x0 = fma(m_[s], x0, dpv);
            Nn[s] = x0;
        }
        // suffix scan over lanes: S_l = A_l + M_l S_{l+1}
        double Aacc = x0;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
            double Ao = __shfl_down_sync(0xffffffffu, Aacc, 1 << l);
            if (lane + (1 << l) < 32) Aacc = fma(Mlev[l], Ao, Aacc);
        }
        double xin = __shfl_down_sync(0xffffffffu, Aacc, 1);
        if (lane == 31) xin = 0.0;
        double zn = 0.0;
        for (int i = 0; i < Cn; ++i) {
            const int s = i * 32 + lane;
            double xv = fma(Mp[s], xin, Nn[s]);
            Nn[s] = xv;
            zn = fma(z[s], xv, zn);
        }
        zn = warp_sum(zn) + zb;
        Lam = (Lam + zn) / bph;
        if (day < t) {                                       // (:231-236)
            day = day + 1.0 * R0 / pc.c;
            ratio = (Lam + ends) / El;
        }
    }
    if (lane == 0) A.out[w] = ratio;
}

}  // namespace lhm
