// lhm_pack.cuh -- grid and initial-condition set-up of the leptonic drivers ON THE DEVICE (LeMoC.py:107-191,
// Code.py:51-180), one CTA per walker.  The host packer (lehamoc_b200/setup_host.py::_pack) evaluates the same
// expressions with NumPy and stays the bit-for-bit statement of the reference; this kernel exists because at the
// fit's batch sizes the host packer costs twice as much as every GPU kernel of the evaluation together
// (tools/bench_fit.py).  Same expressions in the same order; transcendental calls (10**y, log10, g**-p) may differ
// from NumPy's in the last bit: g_el agrees to <= 2 ulp, nu_syn and el_inj to ~1e-14 relative.
#pragma once
#include "lhm_device.cuh"
#include "../../include/lhm.h"

namespace lhm {

constexpr int PACK_NT = 128;

struct PackArgs {
    int W, Ng, Nh, Nt, variant;             // variant 0: LeMoC.py ("script"), 1: Code.py ("code")
    double time_init, time_end, step_alg, Vexp, m;
    double c, c3, m_el, mc2, sigmaT, four_pi, nuc_pref;   // host-evaluated constants (Python floats, as the reference has them)
    const double* wp;                       // [W, LHM_PACK_NPAR]
    const double *nu_ic_row, *nu_tot_row, *ext_row;       // [Nh], [Nt], [Nt]: identical for every walker of these drivers
    double *consts, *g_el, *nu_syn, *nu_ic, *nu_tot, *el_inj, *ext;
    int Np;
    double m_pr, mpc2;
    double *g_pr, *pr_inj;                  // null: leptonic model
};

// np.linspace(lo, hi, n)[i]: arange * step + start, two roundings, last point exactly hi
__device__ __forceinline__ double linspace_pt(double lo, double hi, int i, int n) {
    const double step = (hi - lo) / (double)(n - 1);
    return (i == n - 1) ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);
}

__device__ __forceinline__ double pack_Q_el_Lum(const PackArgs& A, double L_el, double p_el, double g_min, double g_max) {
    if (p_el == 2.) return L_el / (A.mc2 * log(g_max / g_min));                                     // LeHaMoC_f.py:52-56
    const double e = -p_el + 2.;
    return __dmul_rn(L_el, e) / __dmul_rn(A.mc2, __dsub_rn(pow(g_max, e), pow(g_min, e)));
}

__global__ void __launch_bounds__(PACK_NT) pack_kernel(PackArgs A) {
    const int w = blockIdx.x, tid = threadIdx.x;
    const int Ng = A.Ng, Nh = A.Nh, Nt = A.Nt;
    const double* P = A.wp + (size_t)w * LHM_PACK_NPAR;
    const double R0 = P[LHM_PK_R0], B0 = P[LHM_PK_B0], g_min = P[LHM_PK_GMIN], g_max = P[LHM_PK_GMAX];
    const double p_el = P[LHM_PK_P];
    double* g = A.g_el + (size_t)w * Ng;
    // protons of Code.py::LeHaMoC (Code.py:299-300,325-347,419-422): grid, power-law window, number-rate injection
    if (A.g_pr) {
        const int Np = A.Np;
        double* gp = A.g_pr + (size_t)w * Np;
        const double p_pr = P[LHM_PK_P_PR];
        for (int j = tid; j < Np; j += PACK_NT) gp[j] = exp10(linspace_pt(P[LHM_PK_GMIN_PR], P[LHM_PK_GMAX_PR], j, Np));
        __syncthreads();
        // power-law window: the host's integers (the reference's comparison on the NumPy grid), clamped into the grid
        const int ip_max = min((int)P[LHM_PK_IDXMAX_PR], Np - 1);
        const int ip_min = max(0, min((int)P[LHM_PK_IDXMIN_PR], Np - 1));
        const int stop_p = ip_max < 0 ? Np + ip_max : ip_max;
        const double g_lo = gp[ip_min], g_hi = gp[ip_max < 0 ? Np - 1 : min(ip_max, Np - 1)];
        const double Lp = A.four_pi * R0 * A.m_pr * A.c3 * P[LHM_PK_COMP_PR] / A.sigmaT;
        double Q0;                                                                    // Q_pr_Lum, LeHaMoC_f.py:75-79
        if (p_pr == 2.) Q0 = Lp / (A.mpc2 * log(g_hi / g_lo));
        else {
            const double e = -p_pr + 2.;
            Q0 = __dmul_rn(Lp, e) / __dmul_rn(A.mpc2, __dsub_rn(pow(g_hi, e), pow(g_lo, e)));
        }
        double* pin = A.pr_inj + (size_t)w * Np;
        for (int j = tid; j < Np; j += PACK_NT) pin[j] = (j >= ip_min && j < stop_p) ? Q0 * pow(gp[j], -p_pr) : LHM_SENT;
    }
    for (int j = tid; j < Ng; j += PACK_NT) g[j] = exp10(linspace_pt(g_min, g_max, j, Ng));         // LeMoC.py:115
    __syncthreads();
    // power-law window, LeMoC.py:120-128: first index above 10**g_PL_max (or -1), last index below 10**g_PL_min (or 1) --
    // evaluated by the host with the reference's comparison on the NumPy grid (this kernel's exp10 may differ from
    // np.logspace in the last bits, which must not move the window)
    const int idx_max = min((int)P[LHM_PK_IDXMAX], Ng - 1);
    const int idx_min = max(0, min((int)P[LHM_PK_IDXMIN], Ng - 1));
    const int stop = idx_max < 0 ? Ng + idx_max : idx_max;
    const double g_lo = g[idx_min], g_hi = g[idx_max < 0 ? Ng - 1 : min(idx_max, Ng - 1)];
    // frequency grids, LeMoC.py:131-133,185
    const double nuc = A.nuc_pref * B0 / (A.m_el * A.c) * (g[Ng - 1] * g[Ng - 1]);                   // nu_c(g_el[-1], B0)
    const double syn_hi = log10(7. * nuc) + 1.4;
    double* ns = A.nu_syn + (size_t)w * (Nh + 1);
    for (int k = tid; k < Nh; k += PACK_NT) ns[k] = exp10(linspace_pt(7.5, syn_hi, k, Nh));
    if (tid == 0) ns[Nh] = A.nu_tot_row[Nt - 1];
    for (int k = tid; k < Nh; k += PACK_NT) A.nu_ic[(size_t)w * Nh + k] = A.nu_ic_row[k];
    for (int k = tid; k < Nt; k += PACK_NT) {
        A.nu_tot[(size_t)w * Nt + k] = A.nu_tot_row[k];
        A.ext[(size_t)w * Nt + k] = A.ext_row[k];
    }
    // injection, LeMoC.py:174-175: Q_el_Lum(Lum_e_inj(comp, R0), p, g[lo], g[hi]) * g**-p inside the window
    const double L_inj = A.four_pi * R0 * A.m_el * A.c3 * P[LHM_PK_COMP_INJ] / A.sigmaT;           // LeHaMoC_f.py:66-67
    const double Q0 = pack_Q_el_Lum(A, L_inj, p_el, g_lo, g_hi);
    double* inj = A.el_inj + (size_t)w * Ng;
    for (int j = tid; j < Ng; j += PACK_NT) inj[j] = (j >= idx_min && j < stop) ? Q0 * pow(g[j], -p_el) : LHM_SENT;
    if (tid == 0) {
        const double L_cor = A.four_pi * R0 * A.m_el * A.c3 * P[LHM_PK_COMP_COR] / A.sigmaT;
        const double dt = A.step_alg * R0 / A.c;                                                       // LeMoC.py:108
        int nsteps;
        if (A.variant == 0) nsteps = (int)A.time_end;                                                  // LeMoC.py:196
        else {                                                                                         // Code.py:183
            const double lim = A.time_end * R0 / A.c;
            double t = A.time_init;
            nsteps = 0;
            while (t < lim && nsteps < (1 << 20)) { t += dt; ++nsteps; }
        }
        double* c = A.consts + (size_t)w * LHM_NCONST;
        for (int i = 0; i < LHM_NCONST; ++i) c[i] = 0.0;
        c[LHM_C_R0] = R0; c[LHM_C_B0] = B0; c[LHM_C_M] = A.m; c[LHM_C_VEXP] = A.Vexp; c[LHM_C_DT] = dt;
        c[LHM_C_TINIT] = A.time_init; c[LHM_C_NSTEPS] = (double)nsteps; c[LHM_C_COR_P] = p_el;
        c[LHM_C_COR_Q0] = pack_Q_el_Lum(A, L_cor, p_el, g[1], g[Ng - 2]);                              // LeHaMoC_f.py:202
        c[LHM_C_COR_B] = 1e4;
    }
}

}  // namespace lhm
