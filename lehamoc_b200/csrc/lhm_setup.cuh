// lhm_setup.cuh -- state-independent tables of one walker (one CTA per walker).
// Everything here is a function of the grids only: trapezoid weights, np.interp
// stencils, the gamma-gamma sample kernels.  Built once per batch, read by every timestep.
#pragma once
#include "lhm_device.cuh"
#include "../../include/lhm.h"

namespace lhm {


struct SetupArgs {
    Layout L;
    int W;
    int loss_minus_one, qee_from_ic;
    const double* consts;   // [W, LHM_NCONST]
    const double* g_el;     // [W, Ng]
    const double* nu_syn;   // [W, Ns]
    const double* nu_ic;    // [W, Ni]
    const double* nu_tot;   // [W, Nt]
    const double* el_inj;   // [W, Ng]
    const double* ext;      // [W, Nt]
    double* tabd;           // [W, L.n_doubles]
    double* Ecache;         // [W, (Ns+Ni)*Ng] or null: exp(-nu/nu_c) table for time-constant B
    int* tabi;              // [W, L.n_ints]
};

// log-spaced resample grid of a_gg / Q_ee_f: np.logspace(a, b, NP)[s] = 10**(a + s*step), last = 10**b
__device__ __forceinline__ double logspace_pt(double a, double b, int s, int NP) {
    double step = (b - a) / (double)(NP - 1);
    double y = (s == NP - 1) ? b : (a + (double)s * step);
    return exp10(y);
}

template <int NT>
__global__ void __launch_bounds__(NT) setup_kernel(SetupArgs A) {
    const Layout& L = A.L;
    const int w = blockIdx.x;
    const int tid = threadIdx.x;
    const int Ng = L.Ng, Ns = L.Ns, Ni = L.Ni, Nt = L.Nt, NP = L.NP;
    double* T = A.tabd + (size_t)w * L.n_doubles;
    int* TI = A.tabi + (size_t)w * L.n_ints;
    const double* g = A.g_el + (size_t)w * Ng;
    const double* ns = A.nu_syn + (size_t)w * Ns;
    const double* ni = A.nu_ic + (size_t)w * Ni;
    const double* nt = A.nu_tot + (size_t)w * Nt;
    const double one = A.loss_minus_one ? 1.0 : 0.0;
    const PhysConst& pc = kPC;

    // ---- pass 1: pointwise tables -----------------------------------------------------------
    for (int j = tid; j < Ng; j += NT) {
        double gj = g[j];
        T[L.g + j] = gj;
        T[L.lg + j] = log(gj);
        T[L.g13 + j] = pow(gj, 1.0 / 3.0);                      // g**(1./3.)  LeHaMoC_f.py:88
        T[L.inv4g + j] = 1.0 / (4.0 * gj);
        T[L.elinj + j] = A.el_inj[(size_t)w * Ng + j];
        if (j < Ng - 2) {
            // g_el_mp[k] = (g[k+1]+g[k-1])/2 with Python's negative-index wrap at k=0 (LeMoC.py:116)
            double mp0 = (g[j + 1] + g[(j == 0) ? (Ng - 1) : (j - 1)]) / 2.0;
            double mp1 = (g[j + 2] + g[j]) / 2.0;
            double dg = (g[j + 2] - g[j]) / 2.0;                 // dg_el[j]  LeMoC.py:117
            T[L.sm + j] = (mp0 * mp0 - one) / dg;
            T[L.sp + j] = (mp1 * mp1 - one) / dg;
            T[L.am + j] = mp0 / dg;
            T[L.ap + j] = mp1 / dg;
        } else {
            T[L.sm + j] = T[L.sp + j] = T[L.am + j] = T[L.ap + j] = 0.0;
        }
    }
    for (int k = tid; k < Ns; k += NT) {
        double nu = ns[k];
        T[L.nsyn + k] = nu;
        T[L.lsyn + k] = log10(nu);
        T[L.nm23s + k] = pow(nu, -2.0 / 3.0);
        T[L.nm53s + k] = pow(nu, -5.0 / 3.0);
        if (k < Ns - 2) {                                        // LeMoC.py:188,190,210-211
            double mp0 = (ns[k + 1] + ns[(k == 0) ? (Ns - 1) : (k - 1)]) / 2.0;
            double mp1 = (ns[k + 2] + ns[k]) / 2.0;
            double dn = (ns[k + 2] - ns[k]) / 2.0;
            T[L.asm_ + k] = mp0 / dn;
            T[L.asp_ + k] = mp1 / dn;
        } else {
            T[L.asm_ + k] = T[L.asp_ + k] = 0.0;
        }
        // Q_syn_space mask `m_el*c**2*g < h*nu` is a prefix in gamma: first unmasked index
        int lo = 0, hi = Ng;                                     // first j with !(mc2*g[j] < h*nu)
        double hn = pc.h * nu;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (pc.mc2 * g[mid] < hn) lo = mid + 1; else hi = mid; }
        TI[L.jm_syn + k] = lo;
    }
    for (int k = tid; k < Ni; k += NT) {
        double nu = ni[k];
        T[L.nic + k] = nu;
        T[L.lic + k] = log10(nu);
        T[L.nm53i + k] = pow(nu, -5.0 / 3.0);
        T[L.eps + k] = pc.h * nu / pc.mc2;                       // h*nu/(m_el*c**2.)  LeHaMoC_f.py:116
        T[L.lnu4 + k] = log(nu) - log(4.0);
        if (k < Ni - 2) {                                        // LeMoC.py:189,191,212-213
            double mp0 = (ni[k + 1] + ni[(k == 0) ? (Ni - 1) : (k - 1)]) / 2.0;
            double dn = (ni[k + 2] - ni[k]) / 2.0;
            T[L.aim + k] = (mp0 - ni[k]) / dn;
            T[L.aip + k] = ni[0] / dn;
        } else {
            T[L.aim + k] = T[L.aip + k] = 0.0;
        }
    }
    for (int k = tid; k < Nt; k += NT) {
        double nu = nt[k];
        double x = pc.h * nu / pc.mc2;
        T[L.nt + k] = nu;
        T[L.lt + k] = log10(nu);
        T[L.xt + k] = x;
        T[L.lxt + k] = log10(x);
        T[L.invnt + k] = 1.0 / nu;
        T[L.l2nt + k] = 2.0 * log(nu);
        T[L.hnu2 + k] = pc.h * (nu * nu);                        // h*nu_tot**2.  LeMoC.py:289
        T[L.ext + k] = A.ext[(size_t)w * Nt + k];
    }
    __syncthreads();

    // ---- pass 2: trapezoid weights and interpolation stencils ---------------------------------
    const double* lg = T + L.lg;
    const double* lt = T + L.lt;
    for (int j = tid; j < Ng; j += NT) {
        T[L.wg + j] = trapz_w(lg, j, Ng);                        // np.trapz(., np.log(g))
        T[L.wgi + j] = (j >= 1 && j <= Ng - 2) ? trapz_w(lg + 1, j - 1, Ng - 2) : 0.0;   // log(g[1:-1])
        // U_ph_f: nu_T = 3 m c^2/(4 h gamma); insert position + interpolation bracket (LeHaMoC_f.py:164-166)
        double gj = g[j];
        double nuT = pc.nuT_num / (4.0 * pc.h * gj);
        int idx = -1, code = 0; double dx = 0.0, Dx = 1.0;
        if (nuT < nt[Nt - 1]) {
            double lT = log10(nuT);
            interp_stencil(lT, lt, Nt, code, dx, Dx);
            int lo = 0, hi = Nt;                                 // count of k with lt[k] < lT
            while (lo < hi) { int mid = (lo + hi) >> 1; if (lt[mid] < lT) lo = mid + 1; else hi = mid; }
            idx = lo;
        }
        TI[L.u_idx + j] = idx;
        TI[L.u_j + j] = code;
        T[L.u_dx + j] = dx; T[L.u_Dx + j] = Dx;
        T[L.u_xT + j] = pc.h * nuT / pc.mc2;
        // Q_ee_f: n_g = photon density at nu = 2 gamma m c^2/h (LeHaMoC_f.py:147)
        double lq = log10(2.0 * gj * pc.mc2 / pc.h);
        if (A.qee_from_ic) interp_stencil(lq, T + L.lic, Ni, code, dx, Dx);
        else               interp_stencil(lq, lt, Nt, code, dx, Dx);
        TI[L.q_j + j] = code;
        T[L.q_dx + j] = dx; T[L.q_Dx + j] = Dx;
    }
    {
        // IC target grid is nu_tot[:Nt-1] (LeMoC.py:256 passes index = len(nu_tot)-1)
        // l2nt/2 = ln(nu); weights of np.trapz(., np.log(nu_targ[:index]))
        for (int k = tid; k < Nt; k += NT) {
            double wk = 0.0;
            if (k < Nt - 1) {
                double a = (k > 0) ? (log(nt[k]) - log(nt[k - 1])) : 0.0;
                double b = (k < Nt - 2) ? (log(nt[k + 1]) - log(nt[k])) : 0.0;
                wk = (a + b) * 0.5;
            }
            T[L.wt + k] = wk;
            int code; double dx, Dx;
            interp_stencil(lt[k], T + L.lsyn, Ns, code, dx, Dx);   // photons_tot, LeHaMoC_f.py:156
            TI[L.ps_j + k] = code; T[L.ps_dx + k] = dx; T[L.ps_Dx + k] = Dx;
            interp_stencil(lt[k], T + L.lic, Ni, code, dx, Dx);
            TI[L.pi_j + k] = code; T[L.pi_dx + k] = dx; T[L.pi_Dx + k] = Dx;
        }
    }
    // ---- synchrotron exponentials when B(t) == B0 (m == 0 or Vexp == 0): same expressions as phase_vectors /
    //      phase_syn_rows evaluate per step, tabulated once ------------------------------------------------------
    if (A.Ecache) {
        double* E = A.Ecache + (size_t)w * (Ns + Ni) * Ng;
        const double B0 = A.consts[(size_t)w * LHM_NCONST + LHM_C_B0];
        for (int row = tid; row < Ns + Ni; row += NT) {
            const double nu = (row < Ns) ? ns[row] : ni[row - Ns];
            int lo = 0, hi = Ng;                                 // first j with nu/nu_c(g_j) <= 745.2 (x falls with j)
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                double x = nu * (1.0 / (pc.nuc_pref * B0 * (g[mid] * g[mid])));
                if (x > 745.2) lo = mid + 1; else hi = mid;
            }
            TI[L.jz + row] = lo;
        }
        for (int e = tid; e < (Ns + Ni) * Ng; e += NT) {
            const int row = e / Ng, j = e - row * Ng;
            const double nu = (row < Ns) ? ns[row] : ni[row - Ns];
            const double x = nu * (1.0 / (pc.nuc_pref * B0 * (g[j] * g[j])));
            E[(size_t)j * (Ns + Ni) + row] = (x > 745.2) ? 0.0 : exp(-x);   // transposed: [gamma][row]
        }
    }
    // ---- gamma-gamma sample tables ---------------------------------------------------------------
    const double* lxt = T + L.lxt;
    const double xmax_l = lxt[Nt - 1];                           // log10(max(x)) / log10(h nu_tot[-1]/mc2)
    // a_gg: NP samples per nu_ic (LeHaMoC_f.py:130-136).  K = trapz weight x 0.652 sigmaT (s^2-1)/s^3 ln s x'
    for (int e = tid; e < Ni * NP; e += NT) {
        int o = e / NP, s = e - o * NP;
        double xIC = T[L.eps + o];
        double a = log10(1.3 / xIC);
        double K = 0.0;
        double x0 = logspace_pt(a, xmax_l, 0, NP), x1 = logspace_pt(a, xmax_l, 1, NP);
        if (x0 < x1) {
            double xs = logspace_pt(a, xmax_l, s, NP);
            double lm = (s > 0) ? log(logspace_pt(a, xmax_l, s - 1, NP)) : 0.0;
            double lc = log(xs);
            double lp = (s < NP - 1) ? log(logspace_pt(a, xmax_l, s + 1, NP)) : 0.0;
            double wtr = (((s > 0) ? (lc - lm) : 0.0) + ((s < NP - 1) ? (lp - lc) : 0.0)) * 0.5;
            double sv = xs * xIC;
            K = wtr * (pc.gg_pref * (sv * sv - 1.0) / (sv * sv * sv) * log(sv) * xs);
        }
        T[L.aggK + (size_t)s * Ni + o] = K;                   // transposed: [sample][row]
        if (s == 0) { T[L.agg_a + o] = a; T[L.agg_d + o] = (xmax_l - a) / (double)(NP - 1); }
    }
    // Q_ee_f: NP samples per gamma (LeHaMoC_f.py:143-149); the 2.61 R_gg kernel and trapz weight folded in K
    for (int e = tid; e < Ng * NP; e += NT) {
        int j = e / NP, s = e - j * NP;
        double gj = g[j];
        double K = 0.0;
        double a = log10(1.0 / (2.0 * gj));
        if (2.0 * gj > 1.0) {
            double xs = logspace_pt(a, xmax_l, s, NP);
            double lm = (s > 0) ? log(logspace_pt(a, xmax_l, s - 1, NP)) : 0.0;
            double lc = log(xs);
            double lp = (s < NP - 1) ? log(logspace_pt(a, xmax_l, s + 1, NP)) : 0.0;
            double wtr = (((s > 0) ? (lc - lm) : 0.0) + ((s < NP - 1) ? (lp - lc) : 0.0)) * 0.5;
            double sv = 2.0 * gj * xs;
            K = wtr * (2.61 * (sv * sv - 1.0) / (sv * sv * sv) * log(sv) * xs);
        }
        T[L.qeeK + (size_t)s * Ng + j] = K;
        if (s == 0) { T[L.qee_a + j] = a; T[L.qee_d + j] = (xmax_l - a) / (double)(NP - 1); }
    }
}

}  // namespace lhm
