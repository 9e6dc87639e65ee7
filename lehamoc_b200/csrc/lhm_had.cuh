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
__host__ __device__ inline size_t pion_work_doubles(int ntab) { return (size_t)4 * ntab + 3 * PG_NG + 8 * PG_N; }
__host__ __device__ inline size_t pion_smem_doubles(int Nt, int Np, int ntab) {
    return (size_t)2 * Nt + 2 * Np + pion_work_doubles(ntab);
}

// Qp_g_opt for one product of one walker, by the whole CTA (HAD_NT threads).  out[k] = scale * Q(E_k) for k in
// [0, Nout); `add` accumulates into out instead of overwriting.  Ends with a barrier.
__device__ void pion_eval(const PionIn& I, double* work, double* out, double scale, bool add) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = HAD_NT / 32;
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

    for (int k = tid; k < ntab; k += HAD_NT) {
        tx[k] = tab[k * ncol]; ts[k] = tab[k * ncol + c_s]; td[k] = tab[k * ncol + c_s + 1]; tB[k] = tab[k * ncol + c_s + 2];
    }
    __syncthreads();
    const double ga = log10(fmax(g_min_int, I.gpr0)), gb = log10(I.gprN);
    for (int gi = tid; gi < PG_NG; gi += HAD_NT) {
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
    }
    __syncthreads();
    for (int e = tid; e < PG_N; e += HAD_NT) {
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
    }
    __syncthreads();
    const bool nan_to_num = (sp != SP_2G);
    for (int k = warp; k < I.Nout; k += NW) {
        double Eo;
        if (sp == SP_2G) Eo = K.h * I.nu_ic[k];
        else if (sp == SP_EP || sp == SP_EM) Eo = I.g_el[k] * mec2;
        else Eo = I.E_nu[k];
        double acc = 0.0;
        for (int e = lane; e < PG_N; e += 32) {
            const int gi = e / PG_NE;
            const double wv = Wt[e];
            if (act[gi] == 0.0) continue;
            const double x = Eo / (g_t[gi] * mpc2);
            const double xm = Exm[e], xp = Exp[e];
            double Phi;
            if (x < xm) Phi = Elow[e];
            else if (xp > x && x > xm) {
                const double y = (x - xm) / (xp - xm);
                Phi = EB[e] * exp(-Es[e] * pow(log(x / xm), Ed[e])) * pow(log(2. / (1 + y * y)), Epsi[e]);
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

// g_pr_int of one (gamma_e, gamma_p) pair (LeHaMoC_f.py:571-610), one thread.
// xt [Nt] = h nu/(m_e c^2) ascending, lph [Nt] = log10(photons)
__device__ double bh_pair(double ge, double gp, double n_p, int Nt, const double* __restrict__ xt,
                          const double* __restrict__ lph, const BhSurf& Sm, const BhSurf& Sp) {
    const HadConst& K = kHC;
    const double LN10 = 2.302585092994045684;
    const double Emin = (gp + ge) * (gp + ge) / (4. * (gp * gp) * ge);
    const double Emax = K.m_pr / (gp * K.m_el);
    if (!(Emin < Emax)) return 0.0;
    const double Ee_min = (gp * gp + ge * ge) / (2 * gp * ge);
    const double wmin = (gp + ge) * (gp + ge) / (2 * gp * ge);
    const BhSurf* Sf = (gp > ge) ? &Sp : (gp < ge) ? &Sm : nullptr;
    double hx[4]; int lx = 0;
    if (Sf) {
        double ax = log10(Ee_min);
        lx = bh_interval(Sf->tx, Sf->nx, ax);
        bh_fpbspl(Sf->tx, ax, lx, hx);
    }
    const double ea = log10(1.001 * Emin), eb = log10(Emax);
    double yE[10], xE[10];
    for (int e = 0; e < 10; ++e) {
        const double Eph = exp10(logspace_exp(ea, eb, e, 10));
        // 10**np.interp(E_arr, nu_trgt, log10(photons)): interpolation LINEAR in x (LeHaMoC_f.py:577)
        const double f_ph = exp10(interp_clamped(Eph, xt, lph, Nt)) * K.m_el * K.c * K.c / (K.h * (Eph * Eph));
        const double wmax = 2 * gp * Eph;
        double temp = 0.0;
        if (wmin < wmax) {
            const double wa = log10(wmin), wb = log10(wmax);
            double yw[10], xw[10];
            for (int q = 0; q < 10; ++q) {
                const double w = exp10(logspace_exp(wa, wb, q, 10));
                const double Ee_max = w - 1;
                double res = 0.0;
                if (Sf && Ee_min < Ee_max && Ee_min < 80. / 100. * Ee_max) {
                    double ay = log10(Ee_max);
                    const int ly = bh_interval(Sf->ty, Sf->ny, ay);
                    double hy[4];
                    bh_fpbspl(Sf->ty, ay, ly, hy);
                    const int nky1 = Sf->ny - 4;
                    double sp = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const double* cr = Sf->c + (size_t)(lx - 3 + a) * nky1 + (ly - 3);
#pragma unroll
                        for (int b = 0; b < 4; ++b) sp = sp + cr[b] * hx[a] * hy[b];
                    }
                    if (sp > Sf->zmax) sp = 0.;
                    res = w * exp10(sp);
                }
                yw[q] = res * w;
                xw[q] = log(w);
            }
            temp = simpson_irregular(yw, xw, 10);
            if (temp < 0.) temp = 0.;
        }
        yE[e] = f_ph * temp * Eph;
        xE[e] = log(Eph);
    }
    (void)LN10;
    return n_p * simpson_irregular(yE, xE, 10) / (2. * (gp * gp * gp));
}

constexpr int BH_CH = 8;        // gamma_e values per chunk

struct BhArgs {
    int W, Ng, Np, Nt;
    const double *g_el, *g_pr, *N_pr;      // [W,Ng], [W,Np], [W,Np] (dN_pr/dV dgamma)
    const double *xt, *photons;            // [W,Nt]: h nu/(m_e c^2), photons
    BhSurf Sm, Sp;                         // device pointers
    double* out;                           // [W,Ng] = c * dnde
};

__host__ __device__ inline size_t bh_smem_doubles(int Nt, int Np, const BhSurf& a, const BhSurf& b) {
    return (size_t)2 * Nt + 3 * Np + BH_CH * Np + a.nx + a.ny + (size_t)(a.nx - 4) * (a.ny - 4) + b.nx + b.ny +
           (size_t)(b.nx - 4) * (b.ny - 4) + 8;
}

// stage a surface into shared memory (all threads; no barrier)
__device__ inline double* bh_stage(BhSurf& dst, const BhSurf& src, double* sm, int nthreads) {
    const int nc = (src.nx - 4) * (src.ny - 4);
    double* tx = sm; double* ty = tx + src.nx; double* cc = ty + src.ny;
    for (int k = threadIdx.x; k < src.nx; k += nthreads) tx[k] = src.tx[k];
    for (int k = threadIdx.x; k < src.ny; k += nthreads) ty[k] = src.ty[k];
    for (int k = threadIdx.x; k < nc; k += nthreads) cc[k] = src.c[k];
    dst = src; dst.tx = tx; dst.ty = ty; dst.c = cc;
    return cc + nc;
}

// Q_BH_sol for the gamma_e range [i0, i1) of one walker by the whole CTA: buf [BH_CH][Np] scratch, gp/lgp/np_ [Np] staged
__device__ void bh_chunk(int i0, int i1, int Np, int Nt, const double* ge_arr, const double* gp, const double* lgp,
                         const double* np_, const double* xt, const double* lph, const BhSurf& Sm, const BhSurf& Sp,
                         double* buf, double* out, int nthreads) {
    const int n = (i1 - i0) * Np;
    for (int e = threadIdx.x; e < n; e += nthreads) {
        const int ii = e / Np, j = e - ii * Np;
        buf[e] = bh_pair(ge_arr[i0 + ii], gp[j], np_[j], Nt, xt, lph, Sm, Sp) * gp[j];      // g_pr_int * g_pr
    }
    __syncthreads();
    for (int ii = threadIdx.x; ii < i1 - i0; ii += nthreads)
        out[i0 + ii] = kHC.c * simpson_irregular(buf + (size_t)ii * Np, lgp, Np);               // c * simpson(., ln g_pr)
    __syncthreads();
}

__global__ void __launch_bounds__(HAD_NT) had_bh_kernel(BhArgs A) {
    extern __shared__ __align__(16) double sh[];
    const int w = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
    const int Nt = A.Nt, Np = A.Np;
    double* xt = sh; double* lph = xt + Nt; double* gp = lph + Nt; double* lgp = gp + Np; double* np_ = lgp + Np;
    double* buf = np_ + Np;
    BhSurf Sm, Sp;
    double* p = bh_stage(Sm, A.Sm, buf + (size_t)BH_CH * Np, HAD_NT);
    bh_stage(Sp, A.Sp, p, HAD_NT);
    for (int k = tid; k < Nt; k += HAD_NT) { xt[k] = A.xt[(size_t)w * Nt + k]; lph[k] = log10(A.photons[(size_t)w * Nt + k]); }
    for (int k = tid; k < Np; k += HAD_NT) {
        const double g = A.g_pr[(size_t)w * Np + k];
        gp[k] = g; lgp[k] = log(g); np_[k] = A.N_pr[(size_t)w * Np + k];
    }
    __syncthreads();
    const int i0 = chunk * BH_CH, i1 = min(i0 + BH_CH, A.Ng);
    if (i0 < A.Ng)
        bh_chunk(i0, i1, Np, Nt, A.g_el + (size_t)w * A.Ng, gp, lgp, np_, xt, lph, Sm, Sp, buf, A.out + (size_t)w * A.Ng, HAD_NT);
}

}  // namespace lhm
