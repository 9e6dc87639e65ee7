// lhm_hadtab.cuh -- the state-independent operators of the photo-hadronic phases, built once per batch (or once per
// walker when the grids differ) instead of every timestep (SURVEY.md 7.1b applied to `LeHaMoC Code and Tests/LeHaMoC_f.py`).
//
// In all three functions the state enters at a few interpolated points only; everything else is a function of the grids:
//   Q_BH_sol   (:566-616)  per (gamma_e, gamma_p) pair the inner integral over the invariant energy w -- 100 spline-surface
//                          evaluations -- depends on the pair and on the ten photon energies E_arr of the pair, not on the
//                          photon field, which is only read at those ten energies (:577).  Tabulated: K[e][pair] (Simpson
//                          weights, the w integral, the measure and the gamma_p quadrature weight folded in) and the np.interp
//                          stencil of E_arr[e] on the photon grid.  A step then costs 10 interpolations per pair.
//   Qp_g_opt   (:358-464)  Phi_g(eta, x) depends on (E_out, gamma_p, eta) only; the photon field and the proton
//                          distribution enter through 50 + 1250 interpolated values per product.  Tabulated: Phi[E_out][1250];
//                          a step is a matrix-vector product per product.
//   dg_dt_pg   (:201-220)  the 100 x 30 sample energies per proton energy and all quadrature weights depend on the grids only:
//                          A[k][3000] and the stencils; a step is 3000 interpolations of the photon field per proton energy.
// Same samples, same interpolation rules, same weights as the per-step evaluation in lhm_had.cuh (which stays as the
// fall-back when the tables do not fit in memory); only the association of the products differs (1e-15 relative).
#pragma once
#include "lhm_had.cuh"
#include "lhm_lh.cuh"

namespace lhm {

constexpr int LOSS_NI = 100, LOSS_NQ = 30, LOSS_N = LOSS_NI * LOSS_NQ;

struct HadTabDims {
    int Ng, Ni, Np, Nt;
    int npairs;             // Ng * Np
    int nout;               // rows of the photopion table: e+ (Ng), e- (Ng), 2 gamma (Ni), 4 neutrino flavours (50 each)
    int nloss;              // Np - 1 midpoints
};
__host__ __device__ inline HadTabDims hadtab_dims(const Layout& L) {
    HadTabDims D;
    D.Ng = L.Ng; D.Ni = L.Ni; D.Np = L.Np; D.Nt = L.Nt;
    D.npairs = L.Ng * L.Np;
    D.nout = 2 * L.Ng + L.Ni + 4 * 50;
    D.nloss = L.Np - 1;
    return D;
}
__host__ __device__ inline int pion_row0(const HadTabDims& D, int species) {
    switch (species) {
    case SP_EP: return 0;
    case SP_EM: return D.Ng;
    case SP_2G: return 2 * D.Ng;
    case SP_NUMU: return 2 * D.Ng + D.Ni;
    case SP_ANUMU: return 2 * D.Ng + D.Ni + 50;
    case SP_NUE: return 2 * D.Ng + D.Ni + 100;
    default: return 2 * D.Ng + D.Ni + 150;      // SP_ANUE
    }
}

// device pointers of one walker's (or the batch's) tables
struct HadTabs {
    double* bhK;  double* bhdx;  int* bhcode;       // [10][npairs]
    double* phi;                                    // [nout][PG_N]
    double* lossA; double* lossdx; int* losscode;   // [nloss][LOSS_N]
    int shared;                                     // 1: one set for the whole batch
    size_t bh_stride, phi_stride, loss_stride;      // elements per walker
};
__device__ __forceinline__ HadTabs hadtabs_of(const HadTabs& T, int w) {
    HadTabs R = T;
    if (!T.shared) {
        R.bhK += (size_t)w * T.bh_stride; R.bhdx += (size_t)w * T.bh_stride; R.bhcode += (size_t)w * T.bh_stride;
        R.phi += (size_t)w * T.phi_stride;
        R.lossA += (size_t)w * T.loss_stride; R.lossdx += (size_t)w * T.loss_stride; R.losscode += (size_t)w * T.loss_stride;
    }
    return R;
}

// ---------------------------------------------------------------------------------------------------------------
// Bethe-Heitler: table of one (gamma_e, gamma_p) pair.  Same arithmetic as bh_pair (lhm_had.cuh) up to the photon field.
// wq = c-free quadrature weight of gamma_p in the outer Simpson rule times gamma_p (g_pr_int * g_pr).
__device__ void bh_pair_table(double ge, double gp, double wq, int Nt, const double* __restrict__ xt, const BhStaged& G,
                              double* __restrict__ K, double* __restrict__ dxo, int* __restrict__ codeo, size_t stride) {
    const HadConst& HC = kHC;
    const double LN10 = 2.302585092994045684;
    const double Emin = (gp + ge) * (gp + ge) / (4. * (gp * gp) * ge);
    const double Emax = HC.m_pr / (gp * HC.m_el);
    if (!(Emin < Emax)) {
        for (int e = 0; e < 10; ++e) { K[e * stride] = 0.0; dxo[e * stride] = 0.0; codeo[e * stride] = 1; }
        return;
    }
    const double Ee_min = (gp * gp + ge * ge) / (2 * gp * ge);
    const double wmin = (gp + ge) * (gp + ge) / (2 * gp * ge);
    const bool have = (gp != ge);
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
    const double* __restrict__ SW = kSimpson10;
    const double dE = (eb - ea) / 9.0, rE = exp10(dE);
    const double outer = (dE * LN10) / (2. * (gp * gp * gp)) * wq;
    double Eph = exp10(ea);
#pragma unroll 1
    for (int e = 0; e < 10; ++e) {
        const double fac = HC.m_el * HC.c * HC.c / (HC.h * (Eph * Eph));
        const double wmax = 2 * gp * Eph;
        double temp = 0.0;
        if (wmin < wmax) {
            const double wa = log10(wmin), wb = log10(wmax);
            const double dw = (wb - wa) / 9.0, rw = exp10(dw);
            double w = exp10(wa), accw = 0.0;
            int ly = 3;
#pragma unroll 1
            for (int q = 0; q < 10; ++q) {
                const double Ee_max = w - 1;
                double res = 0.0;
                if (have && Ee_min < Ee_max && Ee_min < 80. / 100. * Ee_max) {
                    double ay = log10(Ee_max);
                    ay = fmin(fmax(ay, tb), te);
                    while (!(ay < ty[ly + 1]) && ly != ny - 5) ++ly;
                    double hy[4];
                    bh_fpbspl_r(ty, ry, ny, ay, ly, hy);
                    double sp = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const double* cr = cc + (lx - 3 + a) * nky1 + (ly - 3);
                        double ta = cr[0] * hy[0];
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
        int code; double dx, Dx;
        interp_stencil(Eph, xt, Nt, code, dx, Dx);          // np.interp(E_arr, nu_trgt, .): linear in x (LeHaMoC_f.py:577)
        K[e * stride] = SW[e] * (fac * temp * Eph) * outer;
        dxo[e * stride] = dx;
        codeo[e * stride] = code;
        Eph *= rE;
    }
}

struct HadTabArgs {
    Layout L;
    HadTabDims D;
    int W;                      // walkers whose tables are built (1 when shared)
    const double* tabd;         // per-walker double tables (setup_kernel)
    HadTabs T;
    HadTables HT;
    BhSurf Sm, Sp;
    int do_bh, do_pion, do_loss, do_nu;
};

// grid (walker, chunk of BH_CH gamma_e); same staging as had_bh_kernel
__global__ void __launch_bounds__(HAD_NT) bh_table_kernel(HadTabArgs A) {
    extern __shared__ __align__(16) double sh[];
    const Layout& L = A.L;
    const int w = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
    const int Nt = L.Nt, Np = L.Np, Ng = L.Ng;
    const double* T = A.tabd + (size_t)w * L.n_doubles;
    double* xt = sh; double* gp = xt + Nt; double* lgp = gp + Np; double* Wp = lgp + Np;
    double* area = Wp + Np;
    double* zm = area; int* desc = reinterpret_cast<int*>(area + 2);
    BhSurf Sm, Sp;
    double* p = bh_stage(Sm, A.Sm, area + 2 + BH_DESC_INTS, HAD_NT, area, desc, zm);
    bh_stage(Sp, A.Sp, p, HAD_NT, area, desc + BH_DESC_INTS, zm + 1);
    BhStaged G; G.base = area; G.desc = desc; G.zmax = zm;
    for (int k = tid; k < Nt; k += HAD_NT) xt[k] = T[L.xt + k];
    for (int k = tid; k < Np; k += HAD_NT) { gp[k] = T[L.gp + k]; lgp[k] = T[L.lgp + k]; }
    __syncthreads();
    if (tid == 0) simpson_weights_irregular(lgp, Np, Wp);
    __syncthreads();
    const HadTabs Tw = hadtabs_of(A.T, w);
    const int i0 = chunk * BH_CH, i1 = min(i0 + BH_CH, Ng);
    for (int e = tid; e < (i1 - i0) * Np; e += HAD_NT) {
        const int ii = e / Np, j = e - ii * Np;
        const size_t pair = (size_t)(i0 + ii) * Np + j;
        bh_pair_table(T[L.g + i0 + ii], gp[j], Wp[j] * gp[j], Nt, xt, G, Tw.bhK + pair, Tw.bhdx + pair, Tw.bhcode + pair,
                      (size_t)A.D.npairs);
    }
}

// The tabulated phases evaluate 10**np.interp(x_s, xp, fp) at ~10^6 fixed abscissae per step.  Per bracket the slope
// (fp[j+1]-fp[j])/(xp[j+1]-xp[j]) is the same for every sample that falls into it, so it is formed once per step (Nt
// divisions instead of one per sample), together with the left ordinate, both in units of log2: a sample is then one FMA
// and the staged 2^x of the gamma-gamma phase (fast_exp2, 1e-15) instead of a division and an exp10.  np.interp's NaN
// fall-backs (-inf ordinates) are taken by the exact routine on the rare sample that needs them.
template <int NT>
__device__ __forceinline__ void tab_interp_stage(const double* __restrict__ xp, const double* __restrict__ fp, int Nt,
                                                 double* __restrict__ isl, double* __restrict__ if0) {
    const double L2_10 = 3.321928094887362347870319429489390175865;
    for (int j = threadIdx.x; j < Nt; j += NT) {
        if0[j] = fp[j] * L2_10;
        isl[j] = (j < Nt - 1) ? ((fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j])) * L2_10 : 0.0;
    }
}
__device__ __forceinline__ double tab_interp_pow10(const double* __restrict__ xp, const double* __restrict__ fp, int Nt,
                                                   const double* __restrict__ isl, const double* __restrict__ if0,
                                                   const double* __restrict__ e2tab, int code, double dx) {
    const int j = code >> 1;
    double r = if0[j];
    if (!(code & 1)) {
        r = fma(isl[j], dx, r);
        if (isnan(r)) {                                       // -inf ordinates: np.interp's fall-backs
            const int jx = min(j, Nt - 2);
            r = interp_eval(fp, code, dx, xp[jx + 1] - xp[jx]) * 3.321928094887362347870319429489390175865;
        }
    }
    return fast_exp2(r, e2tab);
}

// Q_BH_sol with the table: Qbh[i] = c sum_j n_p[j] sum_e K[e][i,j] 10^interp(log10 photons at E_arr[e]); one warp per gamma_e
template <int NT>
__device__ void phase_bh_emis_tab(const Ctx& C, Smem& S, SmemLh& H, const HadTabs& Tw, int npairs) {
    const Layout& L = C.L;
    const int Np = L.Np, Ng = L.Ng, Nt = L.Nt;
    const double* __restrict__ xt = C.T + L.xt;
    for (int j = threadIdx.x; j < Np; j += NT) H.npl[j] = H.Npr[j] / C.V;
    tab_interp_stage<NT>(xt, S.Ln, Nt, H.isl, H.if0);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < Ng; i += NT / 32) {
        double s = 0.0;
        for (int j = lane; j < Np; j += 32) {
            const size_t pair = (size_t)i * Np + j;
            double g = 0.0;
#pragma unroll
            for (int e0 = 0; e0 < 10; e0 += 5) {              // all loads of five samples in flight before the first use
                double K[5], dx[5];
                int code[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const size_t o = (size_t)(e0 + u) * npairs + pair;
                    K[u] = __ldg(Tw.bhK + o); code[u] = __ldg(Tw.bhcode + o); dx[u] = __ldg(Tw.bhdx + o);
                }
#pragma unroll
                for (int u = 0; u < 5; ++u)
                    if (K[u] != 0.0) g = fma(K[u], tab_interp_pow10(xt, S.Ln, Nt, H.isl, H.if0, S.e2tab, code[u], dx[u]), g);
            }
            s = fma(H.npl[j], g, s);
        }
        s = warp_sum(s);
        if (lane == 0) H.Qbh[i] = kHC.c * s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// photopion: the species-dependent staging of pion_eval (lhm_had.cuh) split from the Phi evaluation.
// stage = per-gamma_p and per-(gamma_p, eta) vectors; with_params also the K&A parameters Phi needs.
struct PionStage {
    double *tx, *ts, *td, *tB, *g_t, *coef, *act, *Wt, *Es, *Ed, *EB, *Epsi, *Exm, *Exp, *Elow, *Elxm, *Erinv, *lgm, *igm;
    int ntab;
    bool nan_to_num;
};

__device__ void pion_stage(const PionIn& I, double* work, PionStage& P, bool with_state, bool with_params, int nthreads) {
    const int tid = threadIdx.x;
    const HadConst& K = kHC;
    const int Nt = I.Nt, Np = I.Np, sp = I.species;
    const double h_0 = 0.313, r = 0.1458;
    const double mpc2 = K.m_pr * K.c * K.c, mec2 = K.m_el * K.c * K.c;
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
    P.ntab = ntab; P.nan_to_num = (sp != SP_2G);
    P.tx = work; P.ts = P.tx + ntab; P.td = P.ts + ntab; P.tB = P.td + ntab;
    P.g_t = P.tB + ntab; P.coef = P.g_t + PG_NG; P.act = P.coef + PG_NG;
    P.Wt = P.act + PG_NG;
    P.Es = P.Wt + PG_N; P.Ed = P.Es + PG_N; P.EB = P.Ed + PG_N; P.Epsi = P.EB + PG_N; P.Exm = P.Epsi + PG_N;
    P.Exp = P.Exm + PG_N; P.Elow = P.Exp + PG_N; P.Elxm = P.Elow + PG_N; P.Erinv = P.Elxm + PG_N;
    P.lgm = P.Erinv + PG_N; P.igm = P.lgm + PG_NG;
    if (with_params) {
        for (int k = tid; k < ntab; k += nthreads) {
            P.tx[k] = tab[k * ncol]; P.ts[k] = tab[k * ncol + c_s]; P.td[k] = tab[k * ncol + c_s + 1]; P.tB[k] = tab[k * ncol + c_s + 2];
        }
    }
    __syncthreads();
    const double ga = log10(fmax(g_min_int, I.gpr0)), gb = log10(I.gprN);
    for (int gi = tid; gi < PG_NG; gi += nthreads) {
        const double lg = logspace_exp(ga, gb, gi, PG_NG);
        const double gp = exp10(lg);
        P.g_t[gi] = gp;
        const double lm = (gi > 0) ? log(exp10(logspace_exp(ga, gb, gi - 1, PG_NG))) : 0.0;
        const double lc = log(gp);
        const double lq = (gi < PG_NG - 1) ? log(exp10(logspace_exp(ga, gb, gi + 1, PG_NG))) : 0.0;
        const double wg = (((gi > 0) ? (lc - lm) : 0.0) + ((gi < PG_NG - 1) ? (lq - lc) : 0.0)) * 0.5;
        double Nt_ = 1.0;
        if (with_state) Nt_ = exp10(interp_clamped(log10(gp), I.lgp, I.lNp, Np));
        P.coef[gi] = wg * (pref * Nt_ / mpc2);
        const double c_t = mpc2 / (4. * gp);
        P.act[gi] = (log10(thr * c_t / K.h) < log10(I.nu_max)) ? 1.0 : 0.0;
        P.lgm[gi] = log(gp * mpc2);
        P.igm[gi] = 1.0 / (gp * mpc2);
    }
    __syncthreads();
    for (int e = tid; e < PG_N; e += nthreads) {
        const int gi = e / PG_NE, ei = e - gi * PG_NE;
        const double gp = P.g_t[gi];
        const double c_t = mpc2 / (4. * gp);
        const double eta_max = 4. * K.h * I.nu_max * gp / mpc2;
        const double ea = log10(thr), eb = log10(eta_max);
        auto eta_at = [&](int i) { return exp10(logspace_exp(ea, eb, i, PG_NE)); };
        const double eta = eta_at(ei);
        const double nut = eta * c_t / K.h;
        if (with_state) {
            const double phv = exp10(interp_clamped(log10(nut), I.lnt, I.lpt, Nt));
            const double lc = log(K.h * nut);
            const double lm = (ei > 0) ? log(K.h * (eta_at(ei - 1) * c_t / K.h)) : 0.0;
            const double lq = (ei < PG_NE - 1) ? log(K.h * (eta_at(ei + 1) * c_t / K.h)) : 0.0;
            const double wgt = (((ei > 0) ? (lc - lm) : 0.0) + ((ei < PG_NE - 1) ? (lq - lc) : 0.0)) * 0.5;
            P.Wt[e] = (P.act[gi] != 0.0) ? P.coef[gi] * (wgt * (phv * nut)) : 0.0;     // the whole state vector V[e]
        }
        if (with_params) {
            const double hh = eta / h_0;
            const double s = interp_clamped(hh, P.tx, P.ts, ntab), d = interp_clamped(hh, P.tx, P.td, ntab),
                         B = interp_clamped(hh, P.tx, P.tB, ntab);
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
            P.Es[e] = s; P.Ed[e] = d; P.EB[e] = B; P.Epsi[e] = psi; P.Exm[e] = xm; P.Exp[e] = xp;
            P.Elow[e] = B * pow(log(2.), psi);
            P.Elxm[e] = log(xm); P.Erinv[e] = 1.0 / (xp - xm);
        }
    }
    __syncthreads();
}

// Phi_g(eta_e, x = Eo / (gamma_p m_p c^2)) as pion_eval evaluates it (np.nan_to_num included for the lepton products)
__device__ __forceinline__ double pion_phi(const PionStage& P, int e, double Eo, double lEo) {
    const int gi = e / PG_NE;
    const double mpc2 = kHC.m_pr * kHC.c * kHC.c;
    const double x = Eo / (P.g_t[gi] * mpc2);
    const double xm = P.Exm[e], xp = P.Exp[e];
    double Phi;
    if (x < xm) Phi = P.Elow[e];
    else if (xp > x && x > xm) {
        const double y = (x - xm) * P.Erinv[e];
        double Lg = (lEo - P.lgm[gi]) - P.Elxm[e];
        if (Lg < 0.0) Lg = 0.0;
        const double t1 = exp(P.Ed[e] * log(Lg));
        const double L2 = log(0.693147180559945309417 - log1p(y * y));
        Phi = P.EB[e] * exp(fma(P.Epsi[e], L2, -P.Es[e] * t1));
    } else Phi = 0.0;
    if (P.nan_to_num) {
        if (isnan(Phi)) Phi = 0.0;
        else if (isinf(Phi)) Phi = copysign(1.7976931348623157e308, Phi);
    }
    return Phi;
}

// grid (walker, species): Phi[row0 + k][e] for every outgoing energy k of the species
__global__ void __launch_bounds__(HAD_NT) pion_table_kernel(HadTabArgs A) {
    extern __shared__ __align__(16) double sh[];
    const Layout& L = A.L;
    const int w = blockIdx.x, sp = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (!A.do_nu && sp >= SP_NUMU) return;
    const double* T = A.tabd + (size_t)w * L.n_doubles;
    const HadConst& K = kHC;
    double* Enu = sh;                       // [52]
    double* work = Enu + 52;
    for (int k = tid; k < 52; k += HAD_NT) Enu[k] = (k < 50) ? exp10(logspace_exp(10., 22., k, 50)) * K.eV : 0.0;
    __syncthreads();
    PionIn I;
    I.species = sp; I.Ng = L.Ng; I.Ni = L.Ni; I.Np = L.Np; I.Nt = L.Nt; I.Nnu = 50;
    I.g_el = T + L.g; I.nu_ic = T + L.nic; I.E_nu = Enu;
    I.lgp = nullptr; I.lNp = nullptr; I.lnt = nullptr; I.lpt = nullptr;
    I.gpr0 = T[L.gp]; I.gprN = T[L.gp + L.Np - 1]; I.nu_max = T[L.nt + L.Nt - 1];
    I.T = A.HT;
    I.Nout = (sp == SP_2G) ? L.Ni : (sp == SP_EP || sp == SP_EM) ? L.Ng : 50;
    PionStage P;
    pion_stage(I, work, P, false, true, HAD_NT);
    const HadTabs Tw = hadtabs_of(A.T, w);
    double* out = Tw.phi + (size_t)pion_row0(A.D, sp) * PG_N;
    const double mec2 = K.m_el * K.c * K.c;
    for (int k = warp; k < I.Nout; k += HAD_NT / 32) {
        double Eo;
        if (sp == SP_2G) Eo = K.h * I.nu_ic[k];
        else if (sp == SP_EP || sp == SP_EM) Eo = I.g_el[k] * mec2;
        else Eo = Enu[k];
        const double lEo = log(Eo);
        for (int e = lane; e < PG_N; e += 32)
            out[(size_t)k * PG_N + e] = (P.act[e / PG_NE] != 0.0) ? pion_phi(P, e, Eo, lEo) : 0.0;
    }
}

// one product with the table: out[k] (+)= scale * sum_e Phi[k][e] V[e]; the state vector V is staged by the whole CTA.
// NaN entries of the 2-gamma table propagate like in the per-step evaluation (V == 0 only where act == 0, and there the
// table holds 0).  Ends with a barrier.
__device__ void pion_eval_tab(const PionIn& I, double* work, const double* __restrict__ phi, double* out, double scale,
                              bool add, int nthreads) {
    PionStage P;
    pion_stage(I, work, P, true, false, nthreads);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < I.Nout; k += nthreads / 32) {
        const double* __restrict__ row = phi + (size_t)k * PG_N;
        double a0 = 0.0, a1 = 0.0;
        int e = lane;
        for (; e + 96 < PG_N; e += 128) {                   // four table entries in flight per lane (the rows come from L2)
            const double p0 = __ldg(row + e), p1 = __ldg(row + e + 32), p2 = __ldg(row + e + 64), p3 = __ldg(row + e + 96);
            const double v0 = P.Wt[e], v1 = P.Wt[e + 32], v2 = P.Wt[e + 64], v3 = P.Wt[e + 96];
            if (v0 != 0.0) a0 = fma(p0, v0, a0);            // act == 0 entries are skipped, as in pion_eval
            if (v1 != 0.0) a1 = fma(p1, v1, a1);
            if (v2 != 0.0) a0 = fma(p2, v2, a0);
            if (v3 != 0.0) a1 = fma(p3, v3, a1);
        }
        for (; e + 32 < PG_N; e += 64) {
            const double p0 = __ldg(row + e), p1 = __ldg(row + e + 32);
            const double v0 = P.Wt[e], v1 = P.Wt[e + 32];
            if (v0 != 0.0) a0 = fma(p0, v0, a0);
            if (v1 != 0.0) a1 = fma(p1, v1, a1);
        }
        for (; e < PG_N; e += 32) { const double v0 = P.Wt[e]; if (v0 != 0.0) a0 = fma(__ldg(row + e), v0, a0); }
        const double acc = warp_sum(a0 + a1);
        if (lane == 0) out[k] = add ? out[k] + scale * acc : scale * acc;
    }
    __syncthreads();
}

template <int NT>
__device__ void phase_photopion_tab(const Ctx& C, Smem& S, SmemLh& H, const HadTables& T, const HadTabs& Tw, const HadTabDims& D) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, Np = L.Np;
    PionIn I;
    I.Ng = L.Ng; I.Ni = L.Ni; I.Np = Np; I.Nt = L.Nt; I.Nnu = 50;
    I.g_el = S.gS; I.nu_ic = S.nuoS; I.E_nu = H.Enu;
    I.lgp = C.T + L.l10gp; I.lNp = H.lNp; I.lnt = C.T + L.lt; I.lpt = S.Ln;
    I.gpr0 = C.T[L.gp]; I.gprN = C.T[L.gp + Np - 1]; I.nu_max = C.T[L.nt + L.Nt - 1];
    I.T = T;
    for (int j = tid; j < Np; j += NT) H.lNp[j] = log10(H.Npr[j] / C.V);
    __syncthreads();
    auto rows = [&](int sp) { return Tw.phi + (size_t)pion_row0(D, sp) * PG_N; };
    I.species = SP_EP; I.Nout = L.Ng; pion_eval_tab(I, H.work, rows(SP_EP), H.Qpi, 1.0, false, NT);
    I.species = SP_EM; I.Nout = L.Ng; pion_eval_tab(I, H.work, rows(SP_EM), H.Qpi, 1.0, true, NT);
    I.species = SP_2G; I.Nout = L.Ni; pion_eval_tab(I, H.work, rows(SP_2G), H.Qpg, 1.0, false, NT);
    if (C.F.nu) {
        for (int j = tid; j < Np; j += NT) H.lNp[j] = log10(H.Npr[j]);          // LeHaMoC.py:347 passes N_pr itself
        __syncthreads();
        I.Nout = 50;
        I.species = SP_NUMU;  pion_eval_tab(I, H.work, rows(SP_NUMU), H.Qnu, C.dt, false, NT);
        I.species = SP_ANUMU; pion_eval_tab(I, H.work, rows(SP_ANUMU), H.Qnu, C.dt, true, NT);
        I.species = SP_NUE;   pion_eval_tab(I, H.work, rows(SP_NUE), H.Qnu, C.dt, true, NT);
        I.species = SP_ANUE;  pion_eval_tab(I, H.work, rows(SP_ANUE), H.Qnu, C.dt, true, NT);
    } else {
        for (int k = tid; k < 52; k += NT) H.Qnu[k] = 0.0;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// photopion loss rate dg_dt_pg at the midpoints g_pr_mp[k]: A[k][i*30+q] and the stencil of log10(eps') on log10(h nu).
// res_k = sum_m A[k][m] 10^interp(log10(photons/h) at sample m).  Same samples and weights as had_loss_warp(which = 0).
__global__ void __launch_bounds__(HAD_NT) loss_table_kernel(HadTabArgs A) {
    extern __shared__ __align__(16) double sh[];
    const Layout& L = A.L;
    const int w = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = HAD_NT / 32;
    const int Nt = L.Nt;
    const double* T = A.tabd + (size_t)w * L.n_doubles;
    const HadConst& K = kHC;
    double* lx = sh;
    LossTab Tb;
    loss_tab_stage(Tb, lx + Nt, A.HT, HAD_NT);
    for (int k = tid; k < Nt; k += HAD_NT) lx[k] = log10(K.h * T[L.nt + k]);
    __syncthreads();
    const double hnu_max = K.h * T[L.nt + Nt - 1];
    const double E_th = 145. * 1e6 * K.eV;
    const HadTabs Tw = hadtabs_of(A.T, w);
    for (int k = blockIdx.y * NW + warp; k < A.D.nloss; k += gridDim.y * NW) {
        const double g_p = T[L.gpm + k];
        double* Ao = Tw.lossA + (size_t)k * LOSS_N;
        double* dxo = Tw.lossdx + (size_t)k * LOSS_N;
        int* co = Tw.losscode + (size_t)k * LOSS_N;
        const bool on = g_p * hnu_max > E_th;
        const double a = log10(E_th), b = log10(2. * g_p * hnu_max) + 3;
        for (int i = lane; i < LOSS_NI; i += 32) {
            double outer = 0.0, a2 = 0.0, b2 = 0.0;
            bool inner_on = false;
            if (on) {
                const double eb = exp10(logspace_exp(a, b, i, LOSS_NI));
                inner_on = eb / (2. * g_p) < hnu_max / 2.;
                a2 = log10(eb / (2. * g_p)); b2 = log10(hnu_max);
                const double CS = interp_clamped(log10(eb), Tb.tx, Tb.ty, Tb.ncs);
                const double KP = interp_clamped(log10(eb), Tb.ux, Tb.uy, Tb.nkp);
                const double lm = (i > 0) ? log(exp10(logspace_exp(a, b, i - 1, LOSS_NI))) : 0.0;
                const double lc = log(eb);
                const double lq = (i < LOSS_NI - 1) ? log(exp10(logspace_exp(a, b, i + 1, LOSS_NI))) : 0.0;
                const double wgt = (((i > 0) ? (lc - lm) : 0.0) + ((i < LOSS_NI - 1) ? (lq - lc) : 0.0)) * 0.5;
                outer = wgt * (CS * KP * (eb * eb)) * (1. / g_p) * (5e-31 * K.c);
            }
            double prev_l = 0.0, prev_ep = 1.0;
            for (int q = 0; q < LOSS_NQ; ++q) {
                double Aq = 0.0, dx = 0.0; int code = 1;
                if (inner_on) {
                    const double ep = exp10(logspace_exp(a2, b2, q, LOSS_NQ));
                    const double ln_e = log(ep);
                    // trapz over ln eps': node q gets half of both neighbouring intervals
                    const double ln_next = (q < LOSS_NQ - 1) ? log(exp10(logspace_exp(a2, b2, q + 1, LOSS_NQ))) : 0.0;
                    const double wq = (((q > 0) ? (ln_e - prev_l) : 0.0) + ((q < LOSS_NQ - 1) ? (ln_next - ln_e) : 0.0)) * 0.5;
                    double Dx;
                    interp_stencil(log10(ep), lx, Nt, code, dx, Dx);
                    Aq = outer * wq / ep;
                    prev_l = ln_e; prev_ep = ep;
                }
                Ao[i * LOSS_NQ + q] = Aq; dxo[i * LOSS_NQ + q] = dx; co[i * LOSS_NQ + q] = code;
            }
            (void)prev_ep;
        }
    }
}

// dg_dt_pg at every midpoint with the table (one warp per midpoint); the caller has staged lx = log10(h nu) in H.lxpg and
// lp = log10(photons / h) in H.lppg
template <int NT>
__device__ void phase_loss_pg_tab(const Ctx& C, Smem& S, SmemLh& H, const HadTabs& Tw, int nloss) {
    const double* __restrict__ lx = H.lxpg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, Nt = C.L.Nt;
    tab_interp_stage<NT>(lx, H.lppg, Nt, H.isl, H.if0);
    __syncthreads();
    for (int k = warp; k < nloss; k += NT / 32) {
        const double* __restrict__ Ak = Tw.lossA + (size_t)k * LOSS_N;
        const double* __restrict__ dk = Tw.lossdx + (size_t)k * LOSS_N;
        const int* __restrict__ ck = Tw.losscode + (size_t)k * LOSS_N;
        double s0 = 0.0, s1 = 0.0;                            // s0: samples m = lane (mod 64), s1: lane + 32 (mod 64)
        constexpr int LU = 8;                                 // samples per lane and round, all loads first
        for (int m = lane; m < LOSS_N; m += 32 * LU) {
            double A[LU], dx[LU];
            int code[LU];
#pragma unroll
            for (int u = 0; u < LU; ++u) {
                const int mm = m + 32 * u;
                const bool in = mm < LOSS_N;
                A[u] = in ? __ldg(Ak + mm) : 0.0; code[u] = in ? __ldg(ck + mm) : 0; dx[u] = in ? __ldg(dk + mm) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < LU; ++u) {
                if (A[u] != 0.0) {
                    const double v = tab_interp_pow10(lx, H.lppg, Nt, H.isl, H.if0, S.e2tab, code[u], dx[u]);
                    if (u & 1) s1 = fma(A[u], v, s1); else s0 = fma(A[u], v, s0);
                }
            }
        }
        const double s = warp_sum(s0 + s1);
        if (lane == 0) H.lossA[k] = s;
    }
}

}  // namespace lhm
