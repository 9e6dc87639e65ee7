// lhm_had.cuh -- hadronic stage, function-level kernels, batched over walkers.
//   A11  dg_dt_BH  `LeHaMoC Code and Tests/LeHaMoC_f.py`:185-196   Bethe-Heitler proton loss rate (table f(xi).csv)
//        dg_dt_pg  LeHaMoC_f.py:201-220                            photopion proton loss rate (cross_section.csv, kp_pg.txt)
//   A12  Qp_g_opt  LeHaMoC_f.py:358-464 with Phi_g :466-563        photopion secondaries (Kelner & Aharonian 2008 eq. 30)
// The six tables are tiny (< 4 KB each); they are staged in shared memory per CTA.  float64 throughout.
#pragma once
#include "lhm_device.cuh"
#include "../../include/lhm.h"

namespace lhm {

struct HadTables {            // device pointers, row-major
    const double* phi_g;   int n_phi_g;      // [n,4]  eta/eta0, s, d, B                         Phi_g_K&A.txt
    const double* phi_el;  int n_phi_el;     // [n,13] eta/eta0, (s,d,B) x {e+, anti-nu_mu, nu_mu, nu_e}   Phi_g_leptons_K&A.txt
    const double* phi_el1; int n_phi_el1;    // [n,7]  eta/eta0, (s,d,B) x {e-, anti-nu_e}       Phi_e-nu_e_K&A.txt
    const double* fk;      int n_fk;         // [n,2]  log10 kappa, log10 f                       f(xi).csv
    const double* cs;      int n_cs;         // [n,2]  photon energy (GeV), cross section (microbarn)  cross_section.csv
    const double* kp;      int n_kp;         // [n,2]  photon energy (GeV), inelasticity          kp_pg.txt
};

struct HadConst {
    double c, h, m_el, m_pr, sigmaT, eV;
};
__constant__ HadConst kHC;

// np.interp(x, xp, fp) with end clamping (finite ordinates)
__device__ __forceinline__ double interp_clamped(double x, const double* __restrict__ xp, const double* __restrict__ fp,
                                                 int n, int fstride = 1) {
    int code; double dx, Dx;
    interp_stencil(x, xp, n, code, dx, Dx);
    const int j = code >> 1;
    const double f0 = fp[(size_t)j * fstride];
    if (code & 1) return f0;
    const double f1 = fp[(size_t)(j + 1) * fstride];
    const double slope = (f1 - f0) / Dx;
    double r = slope * dx + f0;
    if (isnan(r)) {
        r = slope * (dx - Dx) + f1;
        if (isnan(r) && f0 == f1) r = f0;
    }
    return r;
}

// np.logspace(a, b, n)[i] = 10**(a + i*step), the last point exactly 10**b (np.linspace semantics)
__device__ __forceinline__ double logspace_exp(double a, double b, int i, int n) {
    const double step = (b - a) / (double)(n - 1);
    return (i == n - 1) ? b : fma((double)i, step, a);
}

// ---------------------------------------------------------------------------------------------------------------
// A11: proton loss rates.  One CTA per walker, one warp per proton Lorentz factor, lanes over the outer samples.
// ---------------------------------------------------------------------------------------------------------------
struct LossArgs {
    int W, Ne, Nt, which;          // which: 0 = photopion (dg_dt_pg), 1 = Bethe-Heitler (dg_dt_BH)
    const double *g_eval, *nu, *photons;      // [W,Ne], [W,Nt], [W,Nt]
    HadTables T;
    double* out;                   // [W,Ne]
};

constexpr int HAD_NT = 256;

// tables of the loss rates staged in shared memory: (log10 energy, value) of cross_section.csv[1:] and kp_pg.txt for
// the photopion rate, (log10 kappa, log10 f) of f(xi).csv for the Bethe-Heitler rate
struct LossTab {
    const double *tx, *ty; int ncs;
    const double *ux, *uy; int nkp;
    const double *fx, *fy; int nfk;
};
__host__ __device__ inline size_t loss_tab_doubles(const HadTables& T) { return (size_t)2 * (T.n_cs + T.n_kp + T.n_fk); }

// stage the tables (all threads of the CTA; no barrier inside)
__device__ inline void loss_tab_stage(LossTab& Tb, double* sm, const HadTables& T, int nthreads) {
    const HadConst& K = kHC;
    double* tx = sm; double* ty = tx + T.n_cs; double* ux = ty + T.n_cs; double* uy = ux + T.n_kp;
    double* fx = uy + T.n_kp; double* fy = fx + T.n_fk;
    const int ncs = T.cs ? T.n_cs - 1 : 0;                          // the reference drops the first row ([1:])
    for (int k = threadIdx.x; k < ncs; k += nthreads) {
        tx[k] = log10(T.cs[2 * (k + 1)] * 1e9 * K.eV);
        ty[k] = T.cs[2 * (k + 1) + 1];
    }
    for (int k = threadIdx.x; k < (T.kp ? T.n_kp : 0); k += nthreads) { ux[k] = log10(T.kp[2 * k] * 1e9 * K.eV); uy[k] = T.kp[2 * k + 1]; }
    for (int k = threadIdx.x; k < (T.fk ? T.n_fk : 0); k += nthreads) { fx[k] = T.fk[2 * k]; fy[k] = T.fk[2 * k + 1]; }
    Tb.tx = tx; Tb.ty = ty; Tb.ncs = ncs; Tb.ux = ux; Tb.uy = uy; Tb.nkp = T.n_kp; Tb.fx = fx; Tb.fy = fy; Tb.nfk = T.n_fk;
}

// one proton Lorentz factor by one warp (all 32 lanes take part, all get the result).
// which = 0: dg_dt_pg (lx = log10(h nu), lp = log10(photons/h)); 1: dg_dt_BH (lx = log10(h nu/mc2), lp = log10(photons))
__device__ double had_loss_warp(int which, double g_p, int Nt, double hnu_max, const double* __restrict__ lx,
                                const double* __restrict__ lp, const LossTab& Tb) {
    const int lane = threadIdx.x & 31;
    const HadConst& K = kHC;
    const double mc2 = K.m_el * K.c * K.c;
    double res = 0.0;
    if (which == 0) {
        const double E_th = 145. * 1e6 * K.eV;
        if (g_p * hnu_max > E_th) {
            const double a = log10(E_th), b = log10(2. * g_p * hnu_max) + 3;
            double s = 0.0;
            for (int i = lane; i < 100; i += 32) {
                const double leb = logspace_exp(a, b, i, 100);
                const double eb = exp10(leb);
                double inner = 0.0;
                if (eb / (2. * g_p) < hnu_max / 2.) {
                    // trapz(dN/eps', ln eps') over eps' = logspace(log10(eb/(2 g_p)), log10(h nu_max), 30)
                    const double a2 = log10(eb / (2. * g_p)), b2 = log10(hnu_max);
                    double prev_l = 0.0, prev_y = 0.0;
                    for (int q = 0; q < 30; ++q) {
                        const double le = logspace_exp(a2, b2, q, 30);
                        const double ep = exp10(le);
                        const double y = exp10(interp_clamped(log10(ep), lx, lp, Nt)) / ep;
                        const double ln_e = log(ep);
                        if (q > 0) inner += (ln_e - prev_l) * (y + prev_y) / 2.0;
                        prev_l = ln_e; prev_y = y;
                    }
                }
                const double CS = interp_clamped(log10(eb), Tb.tx, Tb.ty, Tb.ncs);
                const double KP = interp_clamped(log10(eb), Tb.ux, Tb.uy, Tb.nkp);
                // np.trapz(., np.log(eps_bar)) weights from the actual logs
                const double lm = (i > 0) ? log(exp10(logspace_exp(a, b, i - 1, 100))) : 0.0;
                const double lc = log(eb);
                const double lq = (i < 99) ? log(exp10(logspace_exp(a, b, i + 1, 100))) : 0.0;
                const double wgt = (((i > 0) ? (lc - lm) : 0.0) + ((i < 99) ? (lq - lc) : 0.0)) * 0.5;
                s += wgt * (CS * KP * inner * (eb * eb));
            }
            s = warp_sum(s);
            res = 1. / g_p * s * (5e-31 * K.c);
        }
    } else {
        if (g_p * hnu_max > 2.1 * mc2) {
            const double C_BH = 3. * K.sigmaT * K.c * K.m_el / (8 * M_PI * 137. * K.m_pr);
            const double a = log10(2.), b = log10(g_p * hnu_max / mc2);
            double s = 0.0;
            for (int i = lane; i < 50; i += 32) {
                const double lk = logspace_exp(a, b, i, 50);
                const double kap = exp10(lk);
                const double phb = mc2 / K.h * exp10(interp_clamped(log10(kap / (2. * g_p)), lx, lp, Nt));
                const double fk = interp_clamped(log10(kap), Tb.fx, Tb.fy, Tb.nfk);
                // np.trapz(..., np.log10(kappa))
                const double lm = (i > 0) ? log10(exp10(logspace_exp(a, b, i - 1, 50))) : 0.0;
                const double lc = log10(kap);
                const double lq = (i < 49) ? log10(exp10(logspace_exp(a, b, i + 1, 50))) : 0.0;
                const double wgt = (((i > 0) ? (lc - lm) : 0.0) + ((i < 49) ? (lq - lc) : 0.0)) * 0.5;
                s += wgt * (exp10(fk) * phb * kap);
            }
            s = warp_sum(s);
            res = s * C_BH;
        }
    }
    return res;
}

__global__ void __launch_bounds__(HAD_NT) had_loss_kernel(LossArgs A) {
    extern __shared__ __align__(16) double sh[];
    const int w = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = HAD_NT / 32;
    const int Nt = A.Nt;
    const HadConst& K = kHC;
    double* lx = sh;                 // abscissae of the photon interpolation
    double* lp = lx + Nt;            // ordinates
    const double* nu = A.nu + (size_t)w * Nt;
    const double* ph = A.photons + (size_t)w * Nt;
    const double mc2 = K.m_el * K.c * K.c;
    LossTab Tb;
    loss_tab_stage(Tb, lp + Nt, A.T, HAD_NT);
    if (A.which == 0) for (int k = tid; k < Nt; k += HAD_NT) { lx[k] = log10(K.h * nu[k]); lp[k] = log10(ph[k] / K.h); }
    else for (int k = tid; k < Nt; k += HAD_NT) { lx[k] = log10(K.h * nu[k] / mc2); lp[k] = log10(ph[k]); }
    __syncthreads();
    const double hnu_max = K.h * nu[Nt - 1];
    for (int e = warp; e < A.Ne; e += NW) {
        const double res = had_loss_warp(A.which, A.g_eval[(size_t)w * A.Ne + e], Nt, hnu_max, lx, lp, Tb);
        if (lane == 0) A.out[(size_t)w * A.Ne + e] = res;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A12: photopion secondaries.  One CTA per (walker, species); per proton energy the eta grid, the interpolated target
// photons and the integration measure are independent of the outgoing energy and are staged in shared memory; a warp
// then owns one outgoing energy and runs its lanes over the 50 x 25 (gamma_p, eta) points.
// ---------------------------------------------------------------------------------------------------------------
enum { SP_2G = 0, SP_EP, SP_EM, SP_NUMU, SP_ANUMU, SP_NUE, SP_ANUE, SP_COUNT };
constexpr int PG_NG = 50, PG_NE = 25, PG_N = PG_NG * PG_NE;

struct PionArgs {
    int W, species, Ng, Ni, Np, Nt, Nnu, Nout;
    const double *g_el, *nu_ic, *E_nu;        // [W,Ng], [W,Ni], [Nnu] (erg)
    const double *N_pr, *g_pr;                // [W,Np]
    const double *photons, *nu_t;             // [W,Nt]
    HadTables T;
    double* out;                              // [W,Nout]
};

__device__ __forceinline__ double ka_x_plus(double eta, double r) {
    return 1. / (2. * (1. + eta)) * (eta + r * r + sqrt((eta - r * r - 2. * r) * (eta - r * r + 2. * r)));
}
__device__ __forceinline__ double ka_x_minus(double eta, double r) {
    return 1. / (2. * (1. + eta)) * (eta + r * r - sqrt((eta - r * r - 2. * r) * (eta - r * r + 2. * r)));
}
__device__ __forceinline__ double ka_x_plus_e(double eta, double r) {
    return 1. / (2. * (1. + eta)) * (eta - 2 * r + sqrt(eta * (eta - 4. * r * (1. + r))));
}
__device__ __forceinline__ double ka_x_minus_e(double eta, double r) {
    return 1. / (2. * (1. + eta)) * (eta - 2 * r - sqrt(eta * (eta - 4. * r * (1. + r))));
}

// Inputs of one product evaluation.  All array pointers may point to shared or global memory.
struct PionIn {
    int species, Ng, Ni, Np, Nt, Nnu, Nout;
    const double *g_el, *nu_ic, *E_nu;        // output grids (only the one of the product is read; g_el[Ng-1] always for e+-)
    const double *lgp, *lNp;                  // [Np] log10(g_pr), log10(N_pr)
    const double *lnt, *lpt;                  // [Nt] log10(nu_target), log10(photons)
    double gpr0, gprN, nu_max;                // g_pr[0], g_pr[-1], nu_target[-1]
    HadTables T;
};
__host__ __device__ inline size_t pion_work_doubles(int ntab) { return (size_t)4 * ntab + 5 * PG_NG + 10 * PG_N; }
__host__ __device__ inline size_t pion_smem_doubles(int Nt, int Np, int ntab) {
    return (size_t)2 * Nt + 2 * Np + pion_work_doubles(ntab);
}

// Qp_g_opt for one product of one walker, by the whole CTA (HAD_NT threads).  out[k] = scale * Q(E_k) for k in
// [0, Nout); `add` accumulates into out instead of overwriting.  Ends with a barrier.  `nthreads` = blockDim.x.
__device__ void pion_eval(const PionIn& I, double* work, double* out, double scale, bool add, int nthreads = HAD_NT) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = nthreads / 32;
    const HadConst& K = kHC;
    const int Nt = I.Nt, Np = I.Np, sp = I.species;
    const double h_0 = 0.313, r = 0.1458;
    const double mpc2 = K.m_pr * K.c * K.c, mec2 = K.m_el * K.c * K.c;
    // species-dependent pieces (LeHaMoC_f.py:359-461, 466-563)
    const bool is_em = (sp == SP_EM || sp == SP_ANUE);
    const double thr = is_em ? 2.14 * h_0 : h_0;
    const double* tab; int ntab, ncol, c_s;
    if (sp == SP_2G) { tab = I.T.phi_g; ntab = I.T.n_phi_g; ncol = 4; c_s = 1; }
    else if (is_em) { tab = I.T.phi_el1; ntab = I.T.n_phi_el1; ncol = 7; c_s = (sp == SP_EM) ? 1 : 4; }
    else { tab = I.T.phi_el; ntab = I.T.n_phi_el; ncol = 13;
           c_s = (sp == SP_EP) ? 1 : (sp == SP_ANUMU) ? 4 : (sp == SP_NUMU) ? 7 : 10; }
    double g_min_int, pref;
    if (sp == SP_2G) { g_min_int = h_0 * mpc2 / (4. * K.h * I.nu_max); pref = K.h; }
    else if (sp == SP_EP || sp == SP_EM) { g_min_int = thr * mpc2 / (4. * I.g_el[I.Ng - 1] * mec2); pref = mec2; }
    else { g_min_int = thr * mpc2 / (4. * I.E_nu[I.Nnu - 1]); pref = K.h; }
    const double* lnt = I.lnt; const double* lpt = I.lpt; const double* lgp = I.lgp; const double* lNp = I.lNp;

    double* tx = work;                 double* ts = tx + ntab; double* td = ts + ntab; double* tB = td + ntab;
    double* g_t = tB + ntab;           double* coef = g_t + PG_NG;  double* act = coef + PG_NG;   // per gamma_p
    double* Wt = act + PG_NG;          // [PG_N] trapezoid weight x photons x nu
    double* Es = Wt + PG_N;            double* Ed = Es + PG_N;    double* EB = Ed + PG_N;
    double* Epsi = EB + PG_N;          double* Exm = Epsi + PG_N; double* Exp = Exm + PG_N; double* Elow = Exp + PG_N;
    double* Elxm = Elow + PG_N;        double* Erinv = Elxm + PG_N;      // ln(x_-), 1/(x_+ - x_-)
    double* lgm = Erinv + PG_N;        double* igm = lgm + PG_NG;        // per gamma_p: ln(gamma_p m_p c^2) and its inverse

    for (int k = tid; k < ntab; k += nthreads) {
        tx[k] = tab[k * ncol]; ts[k] = tab[k * ncol + c_s]; td[k] = tab[k * ncol + c_s + 1]; tB[k] = tab[k * ncol + c_s + 2];
    }
    __syncthreads();
    const double ga = log10(fmax(g_min_int, I.gpr0)), gb = log10(I.gprN);
    for (int gi = tid; gi < PG_NG; gi += nthreads) {
        const double lg = logspace_exp(ga, gb, gi, PG_NG);
        const double gp = exp10(lg);
        g_t[gi] = gp;
        const double Nt_ = exp10(interp_clamped(log10(gp), lgp, lNp, Np));
        // np.trapz(., np.log(g_pr_temp)) weights
        const double lm = (gi > 0) ? log(exp10(logspace_exp(ga, gb, gi - 1, PG_NG))) : 0.0;
        const double lc = log(gp);
        const double lq = (gi < PG_NG - 1) ? log(exp10(logspace_exp(ga, gb, gi + 1, PG_NG))) : 0.0;
        const double wg = (((gi > 0) ? (lc - lm) : 0.0) + ((gi < PG_NG - 1) ? (lq - lc) : 0.0)) * 0.5;
        coef[gi] = wg * (pref * Nt_ / mpc2);
        const double c_t = mpc2 / (4. * gp);
        act[gi] = (log10(thr * c_t / K.h) < log10(I.nu_max)) ? 1.0 : 0.0;
        lgm[gi] = log(gp * mpc2);
        igm[gi] = 1.0 / (gp * mpc2);
    }
    __syncthreads();
    for (int e = tid; e < PG_N; e += nthreads) {
        const int gi = e / PG_NE, ei = e - gi * PG_NE;
        const double gp = g_t[gi];
        const double c_t = mpc2 / (4. * gp);
        const double eta_max = 4. * K.h * I.nu_max * gp / mpc2;
        const double ea = log10(thr), eb = log10(eta_max);
        auto eta_at = [&](int i) { return exp10(logspace_exp(ea, eb, i, PG_NE)); };
        const double eta = eta_at(ei);
        const double nut = eta * c_t / K.h;
        const double phv = exp10(interp_clamped(log10(nut), lnt, lpt, Nt));
        // np.trapz(., np.log(h*nu_target_temp))
        const double lc = log(K.h * nut);
        const double lm = (ei > 0) ? log(K.h * (eta_at(ei - 1) * c_t / K.h)) : 0.0;
        const double lq = (ei < PG_NE - 1) ? log(K.h * (eta_at(ei + 1) * c_t / K.h)) : 0.0;
        const double wgt = (((ei > 0) ? (lc - lm) : 0.0) + ((ei < PG_NE - 1) ? (lq - lc) : 0.0)) * 0.5;
        Wt[e] = (act[gi] != 0.0) ? wgt * (phv * nut) : 0.0;
        const double hh = eta / h_0;
        const double s = interp_clamped(hh, tx, ts, ntab), d = interp_clamped(hh, tx, td, ntab), B = interp_clamped(hh, tx, tB, ntab);
        double xp, xm, psi;
        if (sp == SP_2G) { xp = ka_x_plus(eta, r); xm = ka_x_minus(eta, r); psi = 2.5 + 0.4 * log(hh); }
        else if (sp == SP_EP || sp == SP_ANUMU || sp == SP_NUE) { xp = ka_x_plus(eta, r); xm = ka_x_minus(eta, r) / 4.; psi = 2.5 + 1.4 * log(hh); }
        else if (sp == SP_NUMU) {
            if (hh < 2.17) xp = 0.427 * ka_x_plus(eta, r);
            else if (hh > 2.17 && hh < 10) xp = (0.427 + 0.0729 * (hh - 2.14)) * ka_x_plus(eta, r);
            else xp = ka_x_plus(eta, r);
            xm = 0.427 * ka_x_minus(eta, r);
            psi = 2.5 + 1.4 * log(hh);
        } else {
            xp = ka_x_plus_e(eta, r); xm = ka_x_minus_e(eta, r) / 2.;
            psi = (hh > 4.) ? 6. * (1 - exp(1.5 * (4. - hh))) : 0.;
        }
        Es[e] = s; Ed[e] = d; EB[e] = B; Epsi[e] = psi; Exm[e] = xm; Exp[e] = xp;
        Elow[e] = B * pow(log(2.), psi);
        Elxm[e] = log(xm); Erinv[e] = 1.0 / (xp - xm);
    }
    __syncthreads();
    const bool nan_to_num = (sp != SP_2G);
    for (int k = warp; k < I.Nout; k += NW) {
        double Eo;
        if (sp == SP_2G) Eo = K.h * I.nu_ic[k];
        else if (sp == SP_EP || sp == SP_EM) Eo = I.g_el[k] * mec2;
        else Eo = I.E_nu[k];
        double acc = 0.0;
        const double lEo = log(Eo);
        for (int e = lane; e < PG_N; e += 32) {
            const int gi = e / PG_NE;
            const double wv = Wt[e];
            if (act[gi] == 0.0) continue;
            const double x = Eo / (g_t[gi] * mpc2);
            const double xm = Exm[e], xp = Exp[e];
            double Phi;
            if (x < xm) Phi = Elow[e];
            else if (xp > x && x > xm) {
                // Phi = B exp(-s ln(x/x_-)^delta) ln(2/(1+y^2))^psi  (LeHaMoC_f.py:466-563) with both powers written as
                // exp(p ln b) of positive bases and folded into one exponential; ln(x/x_-) from tabulated logarithms,
                // y with the tabulated 1/(x_+ - x_-): 4 ln + 2 exp instead of 2 pow + 2 ln + 1 exp + 3 divisions
                // (relative deviation from the pow form < 1e-13, far inside the parity tolerance)
                const double y = (x - xm) * Erinv[e];
                double L = (lEo - lgm[gi]) - Elxm[e];
                if (L < 0.0) L = 0.0;        // last-bit case of the tabulated difference for x barely above x_-; a NaN
                                             // (x_- <= 0 next to the threshold) stays a NaN, as in the reference
                const double t1 = exp(Ed[e] * log(L));
                const double L2 = log(0.693147180559945309417 - log1p(y * y));
                Phi = EB[e] * exp(fma(Epsi[e], L2, -Es[e] * t1));
            } else Phi = 0.0;
            if (nan_to_num) {                                   // np.nan_to_num: nan -> 0, +-inf -> +-max double
                if (isnan(Phi)) Phi = 0.0;
                else if (isinf(Phi)) Phi = copysign(1.7976931348623157e308, Phi);
            }
            acc += coef[gi] * (wv * Phi);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[k] = add ? out[k] + scale * acc : scale * acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(HAD_NT) had_pion_kernel(PionArgs A) {
    extern __shared__ __align__(16) double sh[];
    const int w = blockIdx.x, tid = threadIdx.x;
    const int Nt = A.Nt, Np = A.Np;
    double* lnt = sh;                  double* lpt = lnt + Nt;
    double* lgp = lpt + Nt;            double* lNp = lgp + Np;
    double* work = lNp + Np;
    const double* nu_t = A.nu_t + (size_t)w * Nt;
    const double* g_pr = A.g_pr + (size_t)w * Np;
    for (int k = tid; k < Nt; k += HAD_NT) {
        lnt[k] = log10(nu_t[k]);
        lpt[k] = log10(A.photons[(size_t)w * Nt + k]);
    }
    for (int k = tid; k < Np; k += HAD_NT) { lgp[k] = log10(g_pr[k]); lNp[k] = log10(A.N_pr[(size_t)w * Np + k]); }
    __syncthreads();
    PionIn I;
    I.species = A.species; I.Ng = A.Ng; I.Ni = A.Ni; I.Np = Np; I.Nt = Nt; I.Nnu = A.Nnu; I.Nout = A.Nout;
    I.g_el = A.g_el + (size_t)w * A.Ng; I.nu_ic = A.nu_ic + (size_t)w * A.Ni; I.E_nu = A.E_nu;
    I.lgp = lgp; I.lNp = lNp; I.lnt = lnt; I.lpt = lpt;
    I.gpr0 = g_pr[0]; I.gprN = g_pr[Np - 1]; I.nu_max = nu_t[Nt - 1];
    I.T = A.T;
    pion_eval(I, work, A.out + (size_t)w * A.Nout, 1.0, false);
}

}  // namespace lhm

// =================================================================================================================
// A13: Bethe-Heitler pair injection Q_BH_sol (LeHaMoC_f.py:566-616).  The two smoothing-spline surfaces of the
// energy-integrated cross section come from SciPy's bisplrep on the host (interp_cs_BH_int, LeHaMoC_f.py:276-341); the
// device evaluates them (FITPACK bispev/fpbspl restated: cubic x cubic tensor-product B-splines, arguments clamped to
// the knot range) and does the three nested Simpson integrals with scipy.integrate.simpson's rule for irregular
// spacing, including its last-interval correction for an even number of samples.
// =================================================================================================================
namespace lhm {

struct BhSurf {            // one tck of bisplrep with kx = ky = 3 (pointers to shared or global memory)
    const double *tx, *ty, *c;
    int nx, ny;
    double zmax;           // max_intrp_cs: values above it are replaced by 0 (LeHaMoC_f.py:592-593)
    const double *rx, *ry; // [3][n]: 1/(t[i] - t[i-j]), j = 1..3 -- the denominators of the de Boor recurrence, built by
                           // bh_stage (null: divide)
};
// What bh_pair reads of a staged surface: offsets (in doubles, from the start of the staging area) of tx, ty, c, rx, ry,
// then nx, ny; zmax follows as a double.  Lives in shared memory: a pair picks one of the two surfaces at run time, and
// selecting between two structs by address would push both into local memory (it did: 1.9 G local-load sectors per
// Test-4 launch).
constexpr int BH_DESC_INTS = 8;                        // 7 used
struct BhStaged {
    const double* base;                                // staging area (shared memory)
    const int* desc;                                   // [2][BH_DESC_INTS]: surface "m" (gp < ge), surface "p" (gp > ge)
    const double* zmax;                                // [2]
};

// FITPACK fpbspl: the 4 non-zero cubic B-splines at x in the knot interval l (t[l] <= x < t[l+1])
__device__ __forceinline__ void bh_fpbspl(const double* __restrict__ t, double x, int l, double* h) {
    double hh[3];
    h[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
#pragma unroll
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
#pragma unroll
        for (int i = 0; i < j; ++i) {
            const int li = l + 1 + i, lj = li - j;
            const double f = hh[i] / (t[li] - t[lj]);
            h[i] = h[i] + f * (t[li] - x);
            h[i + 1] = f * (x - t[lj]);
        }
    }
}
// the same with the knot-difference reciprocals tabulated (r = BhSurf::rx / ry): no division in the recurrence
__device__ __forceinline__ void bh_fpbspl_r(const double* __restrict__ t, const double* __restrict__ r, int n, double x,
                                            int l, double* h) {
    double hh[3];
    h[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
#pragma unroll
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
#pragma unroll
        for (int i = 0; i < j; ++i) {
            const int li = l + 1 + i, lj = li - j;
            const double f = hh[i] * r[(j - 1) * n + li];
            h[i] = h[i] + f * (t[li] - x);
            h[i + 1] = f * (x - t[lj]);
        }
    }
}
__device__ __forceinline__ int bh_interval(const double* __restrict__ t, int n, double& arg) {
    const double tb = t[3], te = t[n - 4];
    if (arg < tb) arg = tb;
    if (arg > te) arg = te;
    int l = 3;
    while (!(arg < t[l + 1]) && l != n - 5) ++l;
    return l;
}

// scipy.integrate.simpson(y, x=x) for N samples (N >= 2)
__device__ double simpson_irregular(const double* y, const double* x, int N) {
    auto basic = [&](int stop) {                 // pairs of intervals starting at 0, 2, ... < stop
        double r = 0.0;
        for (int i = 0; i < stop; i += 2) {
            const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
            const double hsum = h0 + h1, hprod = h0 * h1;
            const double h0d = (h1 != 0.0) ? h0 / h1 : 0.0;
            r += hsum / 6.0 * (y[i] * (2.0 - ((h0d != 0.0) ? 1.0 / h0d : 0.0))
                               + y[i + 1] * (hsum * ((hprod != 0.0) ? hsum / hprod : 0.0)) + y[i + 2] * (2.0 - h0d));
        }
        return r;
    };
    if (N % 2 == 1) return basic(N - 2);
    if (N == 2) return 0.5 * (x[1] - x[0]) * (y[1] + y[0]);
    double r = basic(N - 3);
    const double h0 = x[N - 2] - x[N - 3], h1 = x[N - 1] - x[N - 2];
    double num = 2 * h1 * h1 + 3 * h0 * h1, den = 6 * (h1 + h0);
    const double alpha = (den != 0.0) ? num / den : 0.0;
    num = h1 * h1 + 3.0 * h0 * h1; den = 6 * h0;
    const double beta = (den != 0.0) ? num / den : 0.0;
    num = 1 * h1 * h1 * h1; den = 6 * h0 * (h0 + h1);
    const double eta = (den != 0.0) ? num / den : 0.0;
    r += alpha * y[N - 1] + beta * y[N - 2] - eta * y[N - 3];
    return r;
}

// scipy.integrate.simpson weights (in units of the spacing) for 10 equidistant samples: composite Simpson on samples
// 0..8 plus the even-N last-interval correction 5/12 y9 + 2/3 y8 - 1/12 y7
__constant__ double kSimpson10[10] = {1. / 3., 4. / 3., 2. / 3., 4. / 3., 2. / 3., 4. / 3., 2. / 3., 4. / 3. - 1. / 12.,
                                      1. / 3. + 2. / 3., 5. / 12.};

// scipy.integrate.simpson(y, x=x) is linear in y: W[j] with simpson(y, x) == sum_j W[j] y[j], same pair-of-intervals
// formula and even-N last-interval correction as simpson_irregular above.  One thread, once per walker (the abscissae
// are ln(g_pr)); the per-gamma_e integrals then are dot products a warp can do.
__device__ void simpson_weights_irregular(const double* x, int N, double* W) {
    for (int j = 0; j < N; ++j) W[j] = 0.0;
    if (N == 2) { W[0] = W[1] = 0.5 * (x[1] - x[0]); return; }
    const int stop = (N % 2 == 1) ? N - 2 : N - 3;
    for (int i = 0; i < stop; i += 2) {
        const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        const double hsum = h0 + h1, hprod = h0 * h1;
        const double h0d = (h1 != 0.0) ? h0 / h1 : 0.0;
        W[i] += hsum / 6.0 * (2.0 - ((h0d != 0.0) ? 1.0 / h0d : 0.0));
        W[i + 1] += hsum / 6.0 * (hsum * ((hprod != 0.0) ? hsum / hprod : 0.0));
        W[i + 2] += hsum / 6.0 * (2.0 - h0d);
    }
    if (N % 2 == 0) {
        const double h0 = x[N - 2] - x[N - 3], h1 = x[N - 1] - x[N - 2];
        double num = 2 * h1 * h1 + 3 * h0 * h1, den = 6 * (h1 + h0);
        W[N - 1] += (den != 0.0) ? num / den : 0.0;
        num = h1 * h1 + 3.0 * h0 * h1; den = 6 * h0;
        W[N - 2] += (den != 0.0) ? num / den : 0.0;
        num = 1 * h1 * h1 * h1; den = 6 * h0 * (h0 + h1);
        W[N - 3] -= (den != 0.0) ? num / den : 0.0;
    }
}

// g_pr_int of one (gamma_e, gamma_p) pair (LeHaMoC_f.py:571-610), one thread.
// xt [Nt] = h nu/(m_e c^2) ascending, lph [Nt] = log10(photons)
__device__ double bh_pair(double ge, double gp, double n_p, int Nt, const double* __restrict__ xt,
                          const double* __restrict__ lph, const BhStaged& G) {
    const HadConst& K = kHC;
    const double LN10 = 2.302585092994045684;
    const double Emin = (gp + ge) * (gp + ge) / (4. * (gp * gp) * ge);
    const double Emax = K.m_pr / (gp * K.m_el);
    if (!(Emin < Emax)) return 0.0;
    const double Ee_min = (gp * gp + ge * ge) / (2 * gp * ge);
    const double wmin = (gp + ge) * (gp + ge) / (2 * gp * ge);
    const bool have = (gp != ge);                           // gp == ge: neither surface applies (LeHaMoC_f.py:586-591)
    const int which = (gp > ge) ? 1 : 0;
    const int* D = G.desc + which * BH_DESC_INTS;
    const double* __restrict__ ty = G.base + D[1];
    const double* __restrict__ cc = G.base + D[2];
    const double* __restrict__ ry = G.base + D[4];
    const int ny = D[6], nky1 = ny - 4;
    const double zmax = G.zmax[which];
    const double tb = ty[3], te = ty[ny - 4];
    double hx[4]; int lx = 0;
    if (have) {
        const double* __restrict__ tx = G.base + D[0];
        const int nx = D[5];
        double ax = log10(Ee_min);
        lx = bh_interval(tx, nx, ax);
        bh_fpbspl_r(tx, G.base + D[3], nx, ax, lx, hx);
    }
    const double ea = log10(1.001 * Emin), eb = log10(Emax);
    // Both inner integrals are scipy.integrate.simpson over 10 samples of a np.logspace grid, i.e. over abscissae
    // ln(x_q) that are equidistant up to rounding: composite Simpson on the first nine points plus the last-interval
    // correction for an even number of samples (SciPy >= 1.11) collapse to fixed weights x the spacing.  The samples
    // are accumulated as they are produced (no per-thread arrays), the abscissae advance by one multiplication.
    const double* __restrict__ SW = kSimpson10;
    const double dE = (eb - ea) / 9.0, rE = exp10(dE);
    double Eph = exp10(ea), accE = 0.0;
#pragma unroll 1
    for (int e = 0; e < 10; ++e) {
        // 10**np.interp(E_arr, nu_trgt, log10(photons)): interpolation LINEAR in x (LeHaMoC_f.py:577)
        const double f_ph = exp10(interp_clamped(Eph, xt, lph, Nt)) * K.m_el * K.c * K.c / (K.h * (Eph * Eph));
        const double wmax = 2 * gp * Eph;
        double temp = 0.0;
        if (wmin < wmax) {
            const double wa = log10(wmin), wb = log10(wmax);
            const double dw = (wb - wa) / 9.0, rw = exp10(dw);
            double w = exp10(wa), accw = 0.0;
            int ly = 3;                              // log10(Ee_max) rises with q: the knot interval only moves forward
#pragma unroll 1
            for (int q = 0; q < 10; ++q) {
                const double Ee_max = w - 1;
                double res = 0.0;
                if (have && Ee_min < Ee_max && Ee_min < 80. / 100. * Ee_max) {
                    double ay = log10(Ee_max);
                    ay = fmin(fmax(ay, tb), te);     // bh_interval, resumed from the previous sample's interval
                    while (!(ay < ty[ly + 1]) && ly != ny - 5) ++ly;
                    double hy[4];
                    bh_fpbspl_r(ty, ry, ny, ay, ly, hy);
                    double sp = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const double* cr = cc + (lx - 3 + a) * nky1 + (ly - 3);
#pragma unroll
                        double ta = cr[0] * hy[0];                         // row of the tensor product first: 5 DFMA per a
#pragma unroll
                        for (int b = 1; b < 4; ++b) ta = fma(cr[b], hy[b], ta);
                        sp = fma(hx[a], ta, sp);
                    }
                    if (sp > zmax) sp = 0.;
                    res = w * exp10(sp);
                }
                accw = fma(SW[q], res * w, accw);
                w *= rw;
            }
            temp = accw * (dw * LN10);
            if (temp < 0.) temp = 0.;
        }
        accE = fma(SW[e], f_ph * temp * Eph, accE);
        Eph *= rE;
    }
    return n_p * (accE * (dE * LN10)) / (2. * (gp * gp * gp));
}

constexpr int BH_CH = 16;       // gamma_e values per chunk

struct BhArgs {
    int W, Ng, Np, Nt;
    const double *g_el, *g_pr, *N_pr;      // [W,Ng], [W,Np], [W,Np] (dN_pr/dV dgamma)
    const double *xt, *photons;            // [W,Nt]: h nu/(m_e c^2), photons
    BhSurf Sm, Sp;                         // device pointers
    double* out;                           // [W,Ng] = c * dnde
};

__host__ __device__ inline size_t bh_smem_doubles(int Nt, int Np, const BhSurf& a, const BhSurf& b) {
    return (size_t)2 * Nt + 4 * Np + BH_CH * Np + 4 * (a.nx + a.ny) + (size_t)(a.nx - 4) * (a.ny - 4) + 4 * (b.nx + b.ny) +
           (size_t)(b.nx - 4) * (b.ny - 4) + 8 + 2 + BH_DESC_INTS;
}

// stage a surface into shared memory (all threads; no barrier); `area` = start of the staging area, `desc`/`zmax` = this
// surface's slot of the BhStaged descriptor
__device__ inline double* bh_stage(BhSurf& dst, const BhSurf& src, double* sm, int nthreads, const double* area = nullptr,
                                   int* desc = nullptr, double* zmax = nullptr) {
    const int nc = (src.nx - 4) * (src.ny - 4);
    double* tx = sm; double* ty = tx + src.nx; double* cc = ty + src.ny;
    double* rx = cc + nc; double* ry = rx + 3 * src.nx;
    for (int k = threadIdx.x; k < src.nx; k += nthreads) tx[k] = src.tx[k];
    for (int k = threadIdx.x; k < src.ny; k += nthreads) ty[k] = src.ty[k];
    for (int k = threadIdx.x; k < nc; k += nthreads) cc[k] = src.c[k];
    for (int k = threadIdx.x; k < 3 * src.nx; k += nthreads) {
        const int j = k / src.nx + 1, i = k % src.nx;
        rx[k] = (i >= j) ? 1.0 / (src.tx[i] - src.tx[i - j]) : 0.0;          // inf on the repeated end knots: never read
    }
    for (int k = threadIdx.x; k < 3 * src.ny; k += nthreads) {
        const int j = k / src.ny + 1, i = k % src.ny;
        ry[k] = (i >= j) ? 1.0 / (src.ty[i] - src.ty[i - j]) : 0.0;
    }
    dst = src; dst.tx = tx; dst.ty = ty; dst.c = cc; dst.rx = rx; dst.ry = ry;
    if (desc && threadIdx.x == 0) {
        desc[0] = (int)(tx - area); desc[1] = (int)(ty - area); desc[2] = (int)(cc - area); desc[3] = (int)(rx - area);
        desc[4] = (int)(ry - area); desc[5] = src.nx; desc[6] = src.ny; desc[7] = 0;
        *zmax = src.zmax;
    }
    return ry + 3 * src.ny;
}

// Q_BH_sol for the gamma_e range [i0, i1) of one walker by the whole CTA: buf [BH_CH][Np] scratch, gp/np_ [Np] staged,
// Wp [Np] = Simpson weights over ln(g_pr) (simpson_weights_irregular)
__device__ void bh_chunk(int i0, int i1, int Np, int Nt, const double* ge_arr, const double* gp, const double* Wp,
                         const double* np_, const double* xt, const double* lph, const BhStaged& G,
                         double* buf, double* out, int nthreads) {
    const int n = (i1 - i0) * Np;
    for (int e = threadIdx.x; e < n; e += nthreads) {
        const int ii = e / Np, j = e - ii * Np;
        buf[e] = bh_pair(ge_arr[i0 + ii], gp[j], np_[j], Nt, xt, lph, G) * gp[j];            // g_pr_int * g_pr
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int ii = warp; ii < i1 - i0; ii += nthreads / 32) {                                 // c * simpson(., ln g_pr)
        double s = 0.0;
        for (int j = lane; j < Np; j += 32) s = fma(Wp[j], buf[(size_t)ii * Np + j], s);
        s = warp_sum(s);
        if (lane == 0) out[i0 + ii] = kHC.c * s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(HAD_NT) had_bh_kernel(BhArgs A) {
    extern __shared__ __align__(16) double sh[];
    const int w = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
    const int Nt = A.Nt, Np = A.Np;
    double* xt = sh; double* lph = xt + Nt; double* gp = lph + Nt; double* lgp = gp + Np; double* np_ = lgp + Np;
    double* Wp = np_ + Np;
    double* buf = Wp + Np;
    BhSurf Sm, Sp;
    double* area = buf + (size_t)BH_CH * Np;
    double* zm = area; int* desc = reinterpret_cast<int*>(area + 2);           // [2] zmax, [2][BH_DESC_INTS] ints
    double* p = bh_stage(Sm, A.Sm, area + 2 + BH_DESC_INTS, HAD_NT, area, desc, zm);
    bh_stage(Sp, A.Sp, p, HAD_NT, area, desc + BH_DESC_INTS, zm + 1);
    BhStaged G; G.base = area; G.desc = desc; G.zmax = zm;
    for (int k = tid; k < Nt; k += HAD_NT) { xt[k] = A.xt[(size_t)w * Nt + k]; lph[k] = log10(A.photons[(size_t)w * Nt + k]); }
    for (int k = tid; k < Np; k += HAD_NT) {
        const double g = A.g_pr[(size_t)w * Np + k];
        gp[k] = g; lgp[k] = log(g); np_[k] = A.N_pr[(size_t)w * Np + k];
    }
    __syncthreads();
    if (tid == 0) simpson_weights_irregular(lgp, Np, Wp);
    __syncthreads();
    const int i0 = chunk * BH_CH, i1 = min(i0 + BH_CH, A.Ng);
    if (i0 < A.Ng)
        bh_chunk(i0, i1, Np, Nt, A.g_el + (size_t)w * A.Ng, gp, Wp, np_, xt, lph, G, buf, A.out + (size_t)w * A.Ng, HAD_NT);
}

}  // namespace lhm

// =================================================================================================================
// A14: pp channels (LeHaMoC_f.py:681-1014; Kelner, Aharonian & Bugayov 2006).  Cheap (the reference spends ~10 ms per
// call), so one THREAD per outgoing energy; the reference's data-dependent list partitioning becomes index ranges
// because x = E/E_p falls monotonically along the proton grid.
// =================================================================================================================
namespace lhm {

enum { PP_E = 0, PP_G, PP_NUMU, PP_NUE, PP_COUNT };      // Q_e_pp, Q_g_pp, Q_nu_mu_pp, Q_nu_e_pp
enum { PPS_G = 0, PPS_NUMU1, PPS_NUMU2, PPS_E, PPS_NUE };

struct PpConst {
    double E_th_pi, x_m_pr, x_m_el, x_m_pi, x_m_pi0, x_m_mu, K_pi, tev;
};
__device__ __forceinline__ PpConst pp_const() {
    PpConst P;
    P.E_th_pi = 1.22 * 1e-3; P.x_m_pr = 938.272046 * 1e-6; P.x_m_el = 0.511 * 1e-6; P.x_m_pi = 139.57 * 1e-6;
    P.x_m_pi0 = 134.9766 * 1e-6; P.x_m_mu = 105.66 * 1e-6; P.K_pi = 0.17; P.tev = 0.624151;
    return P;
}

__device__ __forceinline__ double pp_cs_inel(double E_p) {                   // LeHaMoC_f.py:687-692
    if (!(E_p > 1.22 * 1e-3)) return 0.0;
    const double L = log(E_p);
    const double a = 1. - pow(1.22 * 1e-3 / E_p, 4.);
    return 1e-27 * (34.3 + 1.88 * L + 0.25 * (L * L)) * (a * a);
}

// eta of q_pi (LeHaMoC_f.py:695-714): the np.interp fall-back is called with swapped arguments and, for any spectral
// index above 0.91 that is not exactly 2, 2.5 or 3, returns 3.0
__device__ __forceinline__ double pp_eta(double p_pr, bool is_g) {
    if (p_pr == 2.) return is_g ? 1.1 : 0.77;
    if (p_pr == 2.5) return is_g ? 0.86 : 0.62;
    if (p_pr == 3.) return is_g ? 0.91 : 0.67;
    // np.interp(p_pr, [1.1, 0.86, 0.91], [2., 2.5, 3.]) on an unsorted xp: above xp[-1] -> fp[-1], below xp[0] -> fp[0]
    if (p_pr > 0.91) return 3.0;
    if (p_pr < 1.1) return 2.0;                                    // (only reachable for p_pr <= 0.91)
    return 3.0;
}

// 10**np.interp(log10(Ep), log10(g_pr*x_m_pr), log10(N_pr_TeV)) with the tables lEg / lN staged by the caller
__device__ __forceinline__ double pp_Np(double Ep, const double* __restrict__ lEg, const double* __restrict__ lN, int Np) {
    return exp10(interp_clamped(log10(Ep), lEg, lN, Np));
}

__device__ __forceinline__ double pp_q_pi(double E_pi, double eta, double n_H, const double* lEg, const double* lN, int Np) {
    const PpConst P = pp_const();
    const double Ep = P.x_m_pr + E_pi / P.K_pi;
    return eta * kHC.c * n_H / P.K_pi * pp_cs_inel(Ep) * pp_Np(Ep, lEg, lN, Np);
}

__device__ __forceinline__ double pp_g_nu_mu(double x, double r) { return (3. - 2. * r) / (9. * ((1. - r) * (1. - r))) * (9. * (x * x) - 6. * log(x) - 4. * (x * x * x) - 5.); }
__device__ __forceinline__ double pp_h_nu_mu_1(double r) { return (3. - 2. * r) / (9. * ((1. - r) * (1. - r))) * (9. * (r * r) - 6. * log(r) - 4. * (r * r * r) - 5.); }
__device__ __forceinline__ double pp_h_nu_mu_2(double x, double r) { return (1. + 2. * r) * (r - x) / (9. * (r * r)) * (9. * (r + x) - 4. * (r * r + r * x + x * x)); }
__device__ __forceinline__ double pp_g_nu_e(double x, double r) { return 2. * (1. - x) / (3. * ((1. - r) * (1. - r))) * (6. * ((1. - x) * (1. - x)) + r * (5. + 5. * x - 4. * (x * x)) + 6. * r * log(x)); }
__device__ __forceinline__ double pp_h_nu_e_1(double r) { return 2. / (3. * ((1. - r) * (1. - r))) * ((1. - r) * (6. - 7. * r + 11. * (r * r) - 4. * (r * r * r)) + 6. * r * log(r)); }
__device__ __forceinline__ double pp_h_nu_e_2(double x, double r) { return 2. * (r - x) / (3. * (r * r)) * (7. * (r * r) - 4. * (r * r * r) + 7. * x * (r * r) - 2. * (x * x) - 4. * (x * x) * r); }

// Q_pp_sub2 (LeHaMoC_f.py:735-778) over the proton indices j = jhi, jhi-1, ..., jlo (x ascending)
__device__ double pp_sub2(int sp, double E, int jlo, int jhi, const double* __restrict__ Ep_grid, double n_H,
                          const double* lEg, const double* lN, int Np) {
    double acc = 0.0, prev_l = 0.0, prev_y = 0.0;
    bool first = true;
    for (int j = jhi; j >= jlo; --j) {
        const double x = E / Ep_grid[j];
        const double Ep = E / x;                                    // E/x_sub2, as the reference recomputes it
        const double L = log(Ep);
        const double src = kHC.c * n_H * pp_cs_inel(Ep) * pp_Np(Ep, lEg, lN, Np);
        double F;
        if (sp == PPS_G || sp == PPS_NUMU1) {
            double y, Bg, bg, kg;
            if (sp == PPS_G) { y = x; Bg = 1.3 + 0.14 * L + 0.011 * (L * L); bg = 1. / (1.79 + 0.11 * L + 0.008 * (L * L)); kg = 1. / (0.801 + 0.049 * L + 0.014 * (L * L)); }
            else { y = x / 0.427; Bg = 1.75 + 0.204 * L + 0.01 * (L * L); bg = 1. / (1.67 + 0.111 * L + 0.0038 * (L * L)); kg = 1.07 - 0.086 * L + 0.002 * (L * L); }
            const double yb = pow(y, bg), ly = log(y);
            const double q = (1. - yb) / (1. + kg * yb * (1. - yb));
            F = Bg * ly / y * ((q * q) * (q * q)) * (1. / ly - 4. * bg * yb / (1. - yb) - (4. * kg * bg * yb * (1. - 2. * yb)) / (1. + kg * yb * (1. - yb)));
        } else {
            const double Be = 1. / (69.5 + 2.65 * L + 0.3 * (L * L));
            const double be = 1. / pow(0.201 + 0.062 * L + 0.00042 * (L * L), 1. / 4.);
            const double ke = (0.279 + 0.141 * L + 0.0172 * (L * L)) / (0.3 + (2.3 + L) * (2.3 + L));
            const double lx = log(x);
            const double t = 1. + ke * (lx * lx);
            F = Be * (t * t * t) / (x * (1. + 0.3 / pow(x, be))) * pow(-lx, 5.);
        }
        const double yv = src * F, lx = log(x);
        if (!first) acc += (lx - prev_l) * (yv + prev_y) / 2.0;
        first = false; prev_l = lx; prev_y = yv;
    }
    return acc;
}

// Q_pp_sub1 (LeHaMoC_f.py:780-858) over the proton indices j = jhi ... jlo (x ascending, E_p = E/x descending)
__device__ double pp_sub1(int sp, double E, int jlo, int jhi, const double* __restrict__ Ep_grid, double p_pr, double n_H,
                          const double* lEg, const double* lN, int Np) {
    const PpConst P = pp_const();
    const double r = (P.x_m_mu / P.x_m_pi) * (P.x_m_mu / P.x_m_pi);
    // Es = x*E_p recomputed element by element like np.multiply(x, E/x)
    double mn_a = 1e308, mx_a = -1e308, mx_b = -1e308, mx_x = -1e308, mx_Ep = -1e308, mnEs = 1e308, mxEs = -1e308;
    const double mass = (sp == PPS_G) ? P.x_m_pi0 : (sp == PPS_NUMU1) ? P.x_m_pi : (sp == PPS_NUMU2) ? P.x_m_mu : P.x_m_el;
    for (int j = jhi; j >= jlo; --j) {
        const double x = E / Ep_grid[j], Ep = E / x, Es = x * Ep;
        const double a = Es + mass * mass / (4. * Es);
        mn_a = fmin(mn_a, a); mx_a = fmax(mx_a, a);
        mx_b = fmax(mx_b, Es / (1. - (P.x_m_mu / P.x_m_pi) * (P.x_m_mu / P.x_m_pi)));
        mx_x = fmax(mx_x, x); mx_Ep = fmax(mx_Ep, Ep); mnEs = fmin(mnEs, Es); mxEs = fmax(mxEs, Es);
    }
    double la, lb, pre = 2.;
    if (sp == PPS_G) { la = log10(mn_a); lb = 3.; }
    else if (sp == PPS_NUMU1) { la = log10(fmax(mx_b, mx_a)); lb = log10(mx_x * (mx_Ep - P.x_m_pr)); pre = 2. / 0.427; }
    else { la = log10(mx_a); lb = log10(((sp == PPS_NUE) ? 100. : 1000.) * mx_a); }
    const bool shaped = (sp == PPS_NUMU2 || sp == PPS_E || sp == PPS_NUE);
    const double top = (sp == PPS_E) ? mnEs : mxEs;
    const double eta = pp_eta(p_pr, sp == PPS_G);
    const double mpi = (sp == PPS_G) ? P.x_m_pi0 : P.x_m_pi;
    auto E_pi_at = [&](int k) { return exp10(logspace_exp(la, lb, k, 40)); };
    auto shape_at = [&](int k) {                       // f(x_new[k]), x_new = sorted(top/E_pi) ascending = top/E_pi[39-k]
        const double xe = top / E_pi_at(39 - k);
        if (sp == PPS_NUE) return (xe > r) ? pp_g_nu_e(xe, r) : pp_h_nu_e_1(r) + pp_h_nu_e_2(xe, r);
        if (sp == PPS_NUMU2) return (xe > r) ? pp_g_nu_mu(xe, r) : pp_h_nu_mu_1(r) + pp_h_nu_mu_2(xe, r);
        if (xe > r) { const double gv = pp_g_nu_mu(xe, r); return (gv < 0.) ? 0. : gv; }
        const double h1 = pp_h_nu_mu_1(r), h2 = pp_h_nu_mu_2(xe, r);
        const double f1 = (h1 < 0. || h2 < 0.) ? 0. : 1.;      // the reference zeroes flag_h_nu_mu_1 in both cases
        return f1 * h1 + 1. * h2;                               // flag_h_nu_mu_2 stays 1 (h2 >= 0 for x <= r)
    };
    double norm = 1.0;
    if (shaped) {                                       // 1/np.trapz(f, x_new)
        double s = 0.0, px = 0.0, pf = 0.0;
        for (int k = 0; k < 40; ++k) {
            const double xe = top / E_pi_at(39 - k), fv = shape_at(k);
            if (k > 0) s += (xe - px) * (fv + pf) / 2.0;
            px = xe; pf = fv;
        }
        norm = 1. / s;
    }
    double acc = 0.0, pl = 0.0, py = 0.0;
    for (int k = 0; k < 40; ++k) {
        const double Epi = E_pi_at(k);
        double yv = pp_q_pi(Epi, eta, n_H, lEg, lN, Np) / sqrt(Epi * Epi - 0. * (mpi * mpi)) * Epi;
        if (shaped) yv = (norm * shape_at(k)) * yv;     // element k of the shape multiplies element k of E_pi (sic)
        const double le = log(Epi);
        if (k > 0) acc += (le - pl) * (yv + py) / 2.0;
        pl = le; py = yv;
    }
    return pre * acc;
}

// One outgoing energy of one product: the un-normalised value and whether the delta-function range contributed
__device__ double pp_value(int product, double E, int Np, const double* __restrict__ Ep_grid, double p_pr, double n_H,
                           const double* lEg, const double* lN, bool& used_sub1) {
    const bool nu = (product == PP_NUMU || product == PP_NUE);
    // x_j = E/E_p[j] falls with j: {x < a} = {j >= first j with x_j < a}
    auto first_below = [&](double a) { int lo = 0, hi = Np; while (lo < hi) { int m = (lo + hi) >> 1; if (E / Ep_grid[m] < a) hi = m; else lo = m + 1; } return lo; };
    double q = 0.0;
    used_sub1 = false;
    if (E > 0.1) {
        const int jlo = first_below(0.427);
        const int jhi = nu ? first_below(1e-3) - 1 : Np - 1;       // neutrinos: x > 1e-3 as well; x == 1e-3 belongs to neither
        int jh = jhi;
        if (nu) { while (jh >= jlo && !(E / Ep_grid[jh] > 1e-3)) --jh; }
        if (jh >= jlo) {
            if (product == PP_E) q += pp_sub2(PPS_E, E, jlo, jh, Ep_grid, n_H, lEg, lN, Np);
            else if (product == PP_G) q += pp_sub2(PPS_G, E, jlo, jh, Ep_grid, n_H, lEg, lN, Np);
            else if (product == PP_NUE) q += pp_sub2(PPS_NUE, E, jlo, jh, Ep_grid, n_H, lEg, lN, Np);
            else q += pp_sub2(PPS_NUMU1, E, jlo, jh, Ep_grid, n_H, lEg, lN, Np) + pp_sub2(PPS_NUMU2, E, jlo, jh, Ep_grid, n_H, lEg, lN, Np);
        }
    } else if (E < 0.1) {
        const int jlo = first_below(nu ? 1e-3 : 1.);
        if (jlo <= Np - 1) {
            used_sub1 = true;
            if (product == PP_E) q += pp_sub1(PPS_E, E, jlo, Np - 1, Ep_grid, p_pr, n_H, lEg, lN, Np);
            else if (product == PP_G) q += pp_sub1(PPS_G, E, jlo, Np - 1, Ep_grid, p_pr, n_H, lEg, lN, Np);
            else if (product == PP_NUE) q += pp_sub1(PPS_NUE, E, jlo, Np - 1, Ep_grid, p_pr, n_H, lEg, lN, Np);
            else q += pp_sub1(PPS_NUMU1, E, jlo, Np - 1, Ep_grid, p_pr, n_H, lEg, lN, Np) + pp_sub1(PPS_NUMU2, E, jlo, Np - 1, Ep_grid, p_pr, n_H, lEg, lN, Np);
        }
    }
    return q;
}

// the whole function (LeHaMoC_f.py:860-1014) for one product by the CTA (all `nthreads` threads must call): out[k],
// k < Nout; Es[k] = outgoing energies in TeV (shared memory).  Ends with a barrier.
__device__ void pp_product(int product, int Nout, const double* Es, int Np, const double* Ep_grid, double p_pr, double n_H,
                           const double* lEg, const double* lN, double* out, int nthreads) {
    // norm_ind = number of energies whose delta-function range contributed; two-point linregress at norm_ind, norm_ind+1
    int ni = 0;
    for (int k0 = 0; k0 < Nout; k0 += nthreads) {
        const int k = k0 + threadIdx.x;
        bool u1 = false;
        if (k < Nout) out[k] = pp_value(product, Es[k], Np, Ep_grid, p_pr, n_H, lEg, lN, u1);
        ni += __syncthreads_count(u1 ? 1 : 0);
    }
    if (ni >= 1) {
        double norm;
        if (ni + 1 < Nout) {
            const double x0 = log10(Es[ni]), x1 = log10(Es[ni + 1]), y0 = log10(out[ni]), y1 = log10(out[ni + 1]);
            // scipy.stats.linregress on two points: slope = ssxym/ssxm with biased second moments about the means
            const double xm = (x0 + x1) / 2., ym = (y0 + y1) / 2.;
            const double ssxm = ((x0 - xm) * (x0 - xm) + (x1 - xm) * (x1 - xm)) / 2.;
            const double ssxym = ((x0 - xm) * (y0 - ym) + (x1 - xm) * (y1 - ym)) / 2.;
            const double slope = ssxym / ssxm, intercept = ym - slope * xm;
            const double y_fit = exp10(slope * log10(Es[ni - 1]) + intercept);
            norm = y_fit / out[ni - 1];
        } else {
            // fewer than two energies above the delta-function range (a grid that does not reach 0.1 TeV): SciPy >= 1.10
            // returns NaN from linregress on a one-point sample, and the reference multiplies the low-energy part by it
            norm = nan("");
        }
        __syncthreads();
        for (int k = threadIdx.x; k < ni; k += nthreads) out[k] = norm * out[k];
    }
    __syncthreads();
}

struct PpArgs {
    int W, product, Nout, Np;
    const double *grid;                 // [W,Nout]: g_el, nu_ic or nu_nu of the product
    const double *g_pr, *N_pr_TeV;      // [W,Np]
    const double *p_pr, *n_H;           // [W]
    double* out;                        // [W,Nout]
};

__global__ void __launch_bounds__(HAD_NT) had_pp_kernel(PpArgs A) {
    extern __shared__ __align__(16) double sh[];
    const int w = blockIdx.x, tid = threadIdx.x, Np = A.Np, Nout = A.Nout;
    const PpConst P = pp_const();
    const HadConst& K = kHC;
    double* Ep = sh; double* lEg = Ep + Np; double* lN = lEg + Np; double* Es = lN + Np; double* out = Es + Nout;
    for (int j = tid; j < Np; j += HAD_NT) {
        const double g = A.g_pr[(size_t)w * Np + j];
        Ep[j] = g * K.m_pr * (K.c * K.c) * P.tev;                        // E_p_space, LeHaMoC_f.py:864
        lEg[j] = log10(g * P.x_m_pr);
        lN[j] = log10(A.N_pr_TeV[(size_t)w * Np + j]);
    }
    for (int k = tid; k < Nout; k += HAD_NT) {
        const double v = A.grid[(size_t)w * Nout + k];
        Es[k] = (A.product == PP_E) ? v * K.m_el * (K.c * K.c) * P.tev : K.h * v * P.tev;
    }
    __syncthreads();
    pp_product(A.product, Nout, Es, Np, Ep, A.p_pr[w], A.n_H[w], lEg, lN, out, HAD_NT);
    for (int k = tid; k < Nout; k += HAD_NT) A.out[(size_t)w * Nout + k] = out[k];
}

}  // namespace lhm
