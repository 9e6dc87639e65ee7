// lhm_lh.cuh -- the `LeHaMoC Code and Tests/LeHaMoC.py` driver (BASELINE configs 3 and 5) on top of the leptonic phases:
// electron AND proton equations, proton synchrotron, the f_had variants of U_ph_f (50-point resample, LeHaMoC_f.py:171-182),
// a_gg on the nu_ic, nu_syn and nu_tot grids (100 samples, :145-155; LeHaMoC.py:414-416), the gamma-gamma pre-attenuation
// of the first two steps (LeHaMoC.py:418-426) and Q_ee_f (100 samples, full length, :158-168).  One CTA per walker, all
// timesteps in one launch, state in shared memory.  The photo-hadronic source and loss terms (lhm_had.cuh) are not wired
// in yet: this kernel covers runs with the hadronic flags off (leptonic processes + proton synchrotron).
#pragma once
#include "lhm_step.cuh"
#include "lhm_had.cuh"

namespace lhm {

// ---- set-up additions (called from setup_kernel when the layout has a proton grid) ------------------------------------
struct LhSetup {
    const double* g_pr;     // [W, Np]
    const double* pr_inj;   // [W, Np]  dN/(dV dgamma dt)
};

template <int NT>
__device__ void setup_lh_tables(const Layout& L, double* T, int* TI, const double* gp, const double* prinj,
                                const double* ns, const double* nt, double one) {
    const int tid = threadIdx.x, Np = L.Np, Ns = L.Ns, Nt = L.Nt, Ng = L.Ng, NP = L.NP;
    const PhysConst& pc = kPC;
    const double m_pr = 1.67262192369e-24;
    for (int j = tid; j < Np; j += NT) {
        const double g = gp[j];
        T[L.gp + j] = g;
        T[L.lgp + j] = log(g);
        T[L.g13p + j] = pow(g, 1.0 / 3.0);
        T[L.prinj + j] = prinj[j];
        T[L.l10gp + j] = log10(g);
        T[L.gpm + j] = (j < Np - 1) ? (gp[j + 1] + gp[(j == 0) ? (Np - 1) : (j - 1)]) / 2.0 : 0.0;     // g_pr_mp[j]
        T[L.idgp + j] = (j < Np - 2) ? 1.0 / ((gp[j + 2] - gp[j]) / 2.0) : 0.0;
        if (j < Np - 2) {                                   // g_pr_mp / dg_pr, LeHaMoC.py:149-150 (negative-index wrap at 0)
            const double mp0 = (gp[j + 1] + gp[(j == 0) ? (Np - 1) : (j - 1)]) / 2.0;
            const double mp1 = (gp[j + 2] + gp[j]) / 2.0;
            const double dg = (gp[j + 2] - gp[j]) / 2.0;
            T[L.spm + j] = (mp0 * mp0 - one) / dg;          // LeHaMoC.py:296-297 (gamma^2-1) / Code.py:452-453 (gamma^2)
            T[L.spp + j] = (mp1 * mp1 - one) / dg;
        } else {
            T[L.spm + j] = T[L.spp + j] = 0.0;
        }
    }
    __syncthreads();
    for (int j = tid; j < Np; j += NT) T[L.wgp + j] = trapz_w(T + L.lgp, j, Np);
    for (int k = tid; k < Ns; k += NT) {                    // Q_syn_space mask m_pr*c**2*g < h*nu: a prefix in gamma
        int lo = 0, hi = Np;
        const double hn = pc.h * ns[k];
        const double mpc2 = m_pr * pc.c * pc.c;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (mpc2 * gp[mid] < hn) lo = mid + 1; else hi = mid; }
        TI[L.jm_p + k] = lo;
    }
    // a_gg resample grids on nu_syn and nu_tot (same four numbers per row as for nu_ic, see setup_kernel)
    const double xmax_l = T[L.lxt + Nt - 1];
    for (int e = tid; e < Ns + Nt; e += NT) {
        const bool on_syn = e < Ns;
        const int k = on_syn ? e : e - Ns;
        const double x = pc.h * (on_syn ? ns[k] : nt[k]) / pc.mc2;
        const double a = log10(1.3 / x);
        const double step = (xmax_l - a) / (double)(NP - 1);
        const bool on = logspace_pt(a, xmax_l, 0, NP) < logspace_pt(a, xmax_l, 1, NP);
        const int oa = on_syn ? L.ags_a : L.agt_a, od = on_syn ? L.ags_d : L.agt_d, oc = on_syn ? L.ags_c : L.agt_c,
                  orr = on_syn ? L.ags_r2 : L.agt_r2;
        T[oa + k] = a; T[od + k] = on ? step : 0.0; T[oc + k] = pc.gg_pref / x; T[orr + k] = exp10(-2.0 * step);
    }
    // U_ph_f: nu_temp = np.logspace(log10(nu_target[0]), log10(nu_T), 50)  (LeHaMoC_f.py:175-177)
    for (int j = tid; j < Ng; j += NT) {
        const double nuT = pc.nuT_num / (4.0 * pc.h * T[L.g + j]);
        const bool on = nuT < nt[Nt - 1];
        TI[L.uv_on + j] = on ? 1 : 0;
        T[L.uv_d + j] = (log10(nuT) - log10(nt[0])) / 49.0;
    }
}

// ---- shared memory of the LH kernel: the leptonic carve-up + proton state and the extra gamma-gamma vectors -------------
struct SmemLh {
    double *Npr, *up, *inucp;          // [Np] proton number, weight x n_p x g^(1/3), 1/(a_cr_pr g^2)
    double *QsynP;                     // [Ns] proton synchrotron emissivity
    double *aggS;                      // [Ns] a_gg on nu_syn (computed at the end of a step, used by the next solve)
    double *atmp;                      // [Nt] a_gg on nu_tot
    double *Lm2, *Ln2;                 // [Nt] log2 / log10 of the field Q_ee_f integrates (attenuated in the first two steps);
                                       //      alias Smem::y / Smem::cum, which only the leptonic U_ph_f uses
    // photo-hadronic terms (only carved when a hadronic flag is set)
    double *lossA, *lossB, *lNp;       // [Np] dg_dt_pg / dg_dt_BH at g_pr_mp, log10 of the proton distribution
    double *Qpi, *Qpg;                 // [Ng] photopion pairs (e+ + e-), [Ni] photopion photons
    double *Qnu, *Nnu, *Enu;           // [52] neutrino source x dt, neutrino state, E_nu_space (50 used)
    double *lxpg, *lppg;               // [Nt] log10(h nu), log10(photons/h)
    double *isl, *if0;                 // [Nt] per-bracket slope and left ordinate (x log2 10) of the field a tabulated phase samples
    double *ltab;                      // loss tables (LossTab)
    double *work;                      // pion_eval work area
    double *bhbuf, *npl, *Qbh, *bhsurf; // [BH_CH*Np] scratch, [Np] dN_pr/dV, [Ng] Q_BH_sol, staged spline surfaces
    double *bhw;                        // [Np] Simpson weights over ln(g_pr) (simpson_weights_irregular)
};
constexpr int HAD_MAXTAB = 32, HAD_MAXLOSS = 2 * (128 + 64 + 128), BH_MAXSURF = 2048;
__host__ __device__ inline size_t smem_lh_had_doubles(const Layout& L) {
    return (size_t)3 * pad2(L.Np) + pad2(L.Ng) + pad2(L.Ni) + 3 * 52 + 4 * pad2(L.Nt) + HAD_MAXLOSS + pion_work_doubles(HAD_MAXTAB)
           + (size_t)BH_CH * pad2(L.Np) + 2 * pad2(L.Np) + pad2(L.Ng) + BH_MAXSURF;
}
__host__ __device__ inline size_t smem_lh_doubles(const Layout& L) {
    return (size_t)3 * pad2(L.Np) + 2 * pad2(L.Ns) + pad2(L.Nt);
}
__device__ inline void carve_lh(SmemLh& H, double* p, const Layout& L, const Smem& S) {
    auto take = [&](int n) { double* r = p; p += n; return r; };
    H.Npr = take(pad2(L.Np)); H.up = take(pad2(L.Np)); H.inucp = take(pad2(L.Np));
    H.QsynP = take(pad2(L.Ns)); H.aggS = take(pad2(L.Ns));
    H.atmp = take(pad2(L.Nt)); H.Lm2 = S.y; H.Ln2 = S.cum;
    H.lossA = take(pad2(L.Np)); H.lossB = take(pad2(L.Np)); H.lNp = take(pad2(L.Np));
    H.Qpi = take(pad2(L.Ng)); H.Qpg = take(pad2(L.Ni));
    H.Qnu = take(52); H.Nnu = take(52); H.Enu = take(52);
    H.lxpg = take(pad2(L.Nt)); H.lppg = take(pad2(L.Nt));
    H.isl = take(pad2(L.Nt)); H.if0 = take(pad2(L.Nt));
    H.ltab = take(HAD_MAXLOSS); H.work = take((int)pion_work_doubles(HAD_MAXTAB));     // only valid when the handle
    H.bhbuf = take(BH_CH * pad2(L.Np)); H.npl = take(pad2(L.Np)); H.Qbh = take(pad2(L.Ng)); H.bhw = take(pad2(L.Np));
    H.bhsurf = p;                                                                        // reserved smem_lh_had_doubles
}

// U_ph_f, LeHaMoC_f.py:171-182: per gamma a 50-point log resample of the photon field up to nu_T, linear-x trapezoid.
// pre*K = (mc^2)^2/h, so with x = h nu/mc^2:  U = mc^2 * trapz(x * (n mc^2/h), x);  n mc^2/h = 2^Lm.
template <int NT>
__device__ void phase_uph_lh(const Ctx& C, Smem& S) {
    const Layout& L = C.L;
    const int Nt = L.Nt, Ng = L.Ng;
    const double* __restrict__ lxt = S.lxtS;
    const double* __restrict__ ilx = S.ilxS;
    const double lx0 = lxt[0], lxN = lxt[Nt - 1];
    const double invD = (double)(Nt - 1) / (lxN - lx0);
    const double K = kPC.sigmaT * C.Rad * kPC.mc2 / kPC.h;
    const double pre = kPC.mc2 / (kPC.sigmaT * C.Rad);
    for (int j = threadIdx.x; j < Ng; j += NT) {
        double U;
        if (C.TI[L.uv_on + j]) {
            const double step = __ldg(C.T + L.uv_d + j);
            const double rr = exp10(step);                          // ratio of consecutive abscissae
            double acc[GG_U];
#pragma unroll
            for (int u = 0; u < GG_U; ++u) acc[u] = 0.0;
            double xs[GG_U];                                        // x of sample e0+u
            const double x0 = exp10(lx0);
            xs[0] = x0;
#pragma unroll
            for (int u = 1; u < GG_U; ++u) xs[u] = xs[u - 1] * rr;
            double rU = rr;
#pragma unroll
            for (int u = 1; u < GG_U; ++u) rU *= rr;
            const double irr = 1.0 / rr;
            for (int e0 = 0; e0 < 50; e0 += GG_U) {
                double y[GG_U], wv[GG_U], t[GG_U], Lv[GG_U];
                int jb[GG_U];
#pragma unroll
                for (int u = 0; u < GG_U; ++u) {
                    const int e = e0 + u;
                    y[u] = fma((double)min(e, 49), step, lx0);
                    // np.trapz(f, x): weight of node e = (x_{e+1} - x_{e-1})/2 (one-sided at the ends)
                    const double xm = (e > 0) ? xs[u] * irr : xs[u], xp = (e < 49) ? xs[u] * rr : xs[u];
                    wv[u] = (e < 50) ? 0.5 * (xp - xm) * xs[u] : 0.0;
                    xs[u] *= rU;
                }
#pragma unroll
                for (int u = 0; u < GG_U; ++u) {
                    int jj = (int)((y[u] - lx0) * invD);
                    jj = max(0, min(jj, Nt - 2));
                    while (jj > 0 && y[u] < lxt[jj]) --jj;
                    while (jj < Nt - 2 && y[u] >= lxt[jj + 1]) ++jj;
                    jb[u] = jj;
                }
#pragma unroll
                for (int u = 0; u < GG_U; ++u) {
                    t[u] = (y[u] - lxt[jb[u]]) * ilx[jb[u]];
                    if (y[u] <= lx0) t[u] = 0.0;
                    if (y[u] >= lxN) t[u] = 1.0;
                }
#pragma unroll
                for (int u = 0; u < GG_U; ++u) {
                    const double f0 = S.Lm[jb[u]], f1 = S.Lm[jb[u] + 1];
                    Lv[u] = fma(t[u], f1 - f0, f0);
                }
#pragma unroll
                for (int u = 0; u < GG_U; ++u) acc[u] = fma(wv[u], fast_exp2(Lv[u], S.e2tab), acc[u]);
            }
            double s = acc[0];
#pragma unroll
            for (int u = 1; u < GG_U; ++u) s += acc[u];
            U = kPC.mc2 * s;
        } else {                                                    // U_ph_tot, LeHaMoC_f.py:173
            double s = 0.0;
            for (int i = 0; i < Nt - 2; ++i) {
                const double y0 = C.T[L.nt + i] * S.n[i] * K, y1 = C.T[L.nt + i + 1] * S.n[i + 1] * K;
                s += (C.T[L.nt + i + 1] - C.T[L.nt + i]) * (y1 + y0) / 2.0;
            }
            U = pre * s;
        }
        S.bcom[j] = 4.0 / 3.0 * kPC.sigmaT * (kPC.c * U) / kPC.mc2;      // LeHaMoC.py:304
    }
    __syncthreads();
}

// proton equation (LeHaMoC.py:330-335 with the hadronic loss terms off): synchrotron cooling + escape + injection
template <int NT>
__device__ void phase_protons(const Ctx& C, Smem& S, SmemLh& H, int esc_pr, const LossTab* Tb, bool per_volume = true,
                              bool inject = true, double n_H = 0.0, bool pg_tabulated = false) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, n = L.Np - 2;
    const bool had = C.F.pg_l || C.F.bh_l;
    if (had) {                                       // LeHaMoC.py:311-323: loss rates at the midpoints g_pr_mp[0..Np-2]
        const int Nt = L.Nt;
        const double hK = kHC.h;
        for (int k = tid; k < Nt; k += NT) {
            H.lxpg[k] = log10(hK * C.T[L.nt + k]);
            H.lppg[k] = log10(S.n[k] / hK);
        }
        __syncthreads();
        const double hnu_max = hK * C.T[L.nt + Nt - 1];
        for (int k = tid >> 5; k < L.Np - 1; k += NT / 32) {
            const double gm = C.T[L.gpm + k];
            // pg_tabulated: H.lossA already holds dg_dt_pg (phase_loss_pg_tab, lhm_hadtab.cuh)
            const double a = (C.F.pg_l && !pg_tabulated) ? had_loss_warp(0, gm, Nt, hnu_max, H.lxpg, H.lppg, *Tb) : 0.0;
            const double b = C.F.bh_l ? had_loss_warp(1, gm, Nt, hnu_max, S.lxtS, S.Ln, *Tb) : 0.0;
            if ((tid & 31) == 0) { if (!(C.F.pg_l && pg_tabulated)) H.lossA[k] = a; H.lossB[k] = b; }
        }
        __syncthreads();
    }
    const double m_el = 9.1093837015e-28, m_pr = 1.67262192369e-24;
    const double rm = m_el / m_pr;
    const double b_syn = C.F.syn_l ? kPC.bsyn_pref * (rm * rm * rm) * (C.B * C.B) : 0.0;      // const_pr * M_F**2
    const double esc = kPC.c / C.Rad * (double)esc_pr;
    for (int j = tid; j < n; j += NT) {
        double lm = 0.0, lp = 0.0;
        if (had) {
            const double idg = C.T[L.idgp + j];
            lm = (H.lossA[j] + H.lossB[j]) * idg;
            lp = (H.lossA[j + 1] + H.lossB[j + 1]) * idg;
        }
        if (C.F.pp_l) {                              // LeHaMoC.py:322-326: 0.65 c n_H sigma_inel(E_kin [TeV])/(m_p c^2)/dg_pr
            const double idg = C.T[L.idgp + j];
            const double mpc2 = kHC.m_pr * (kHC.c * kHC.c);
            const double rest = mpc2 * 0.624151;
            const double gm0 = C.T[L.gpm + j], gm1 = C.T[L.gpm + j + 1];
            lm += 0.65 * kHC.c * n_H * pp_cs_inel(gm0 * kHC.m_pr * (kHC.c * kHC.c) * 0.624151 - rest) / mpc2 * idg;
            lp += 0.65 * kHC.c * n_H * pp_cs_inel(gm1 * kHC.m_pr * (kHC.c * kHC.c) * 0.624151 - rest) / mpc2 * idg;
        }
        const double V2 = 1.0 + C.dt * (esc + b_syn * C.T[L.spm + j] + lm);
        const double V3 = -C.dt * (b_syn * C.T[L.spp + j] + lp);
        double d = H.Npr[j + 1];
        if (inject) d += per_volume ? (C.T[L.prinj + j + 1] * C.dt) * C.V : C.T[L.prinj + j + 1] * C.dt;
        S.cp[j] = V3 / V2;
        S.dp[j] = d / V2;
    }
    __syncthreads();
    if (tid < 32) backsub_warp(S.cp, S.dp, H.Npr + 1, n);
    __syncthreads();
}

// electron equation, LeHaMoC.py:359-367 (injection per volume; Q_ee aligned with the interior points)
template <int NT>
__device__ void phase_electrons_lh(const Ctx& C, Smem& S, const double* Qpi, const double* Qbh) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, n = L.Ng - 2;
    const double dt = C.dt;
    const double b_ad = C.F.ad ? C.Vexp / C.Rad : 0.0;
    const double b_syn = C.F.syn_l ? kPC.bsyn_pref * (C.B * C.B) : 0.0;
    const double esc = kPC.c / C.Rad * (double)C.F.esc;
    for (int j = tid; j < n; j += NT) {
        const double smj = C.T[L.sm + j], spj = C.T[L.sp + j];
        double ic_m = 0.0, ic_p = 0.0;
        if (C.F.ic_l) { ic_m = S.bcom[j + 1] * smj; ic_p = S.bcom[j + 2] * spj; }
        const double V2 = 1.0 + dt * (esc + b_syn * smj + ic_m + b_ad * C.T[L.am + j]);
        const double V3 = -dt * (b_syn * spj + ic_p + b_ad * C.T[L.ap + j]);
        double src = S.qee[j + 1];
        if (Qpi) src += Qpi[j + 1];
        if (Qbh) src += Qbh[j + 1];
        if (C.F.inj) src += C.T[L.elinj + j + 1];
        const double d = S.N[j + 1] + (src * dt) * C.V;
        S.cp[j] = V3 / V2;
        S.dp[j] = d / V2;
    }
    __syncthreads();
    if (tid < 32) backsub_warp(S.cp, S.dp, S.N + 1, n);
    __syncthreads();
}

// proton synchrotron emissivity Q_syn_space(dN_pr, M_F, nu, a_cr_pr, C_syn_pr, g_pr, m_pr), LeHaMoC.py:372
template <int NT>
__device__ void phase_syn_protons(const Ctx& C, Smem& S, SmemLh& H, bool use_mask = true) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, Np = L.Np, Ns = L.Ns;
    const double m_el = 9.1093837015e-28, m_pr = 1.67262192369e-24;
    const double a_cr_pr = kPC.acr_pref * (m_el / m_pr) * C.B;                 // 3 q B/(4 pi m_pr c)
    for (int j = tid; j < Np; j += NT) {
        const double g = C.T[L.gp + j];
        H.inucp[j] = 1.0 / (a_cr_pr * (g * g));
        H.up[j] = C.T[L.wgp + j] * ((H.Npr[j] / C.V) * C.T[L.g13p + j]);
    }
    __syncthreads();
    const double C_syn_pr = kPC.C_syn * pow(m_pr / m_el, 4.0 / 3.0) * ((m_el / m_pr) * (m_el / m_pr));
    const double qpref = C_syn_pr * pow(C.B, 2.0 / 3.0);
    for (int k = tid; k < Ns; k += NT) {
        double s0 = 0.0, s1 = 0.0;
        if (C.F.syn_e && k >= 1 && k <= Ns - 2) {
            const double nu = C.T[L.nsyn + k];
            const int jm = use_mask ? C.TI[L.jm_p + k] : 0;     // Q_syn_space_pr of Fit_emcee/LeHaMoC_f.py:95-98 has no mask
            for (int j0 = 0; j0 < Np; j0 += 4) {
                double x[4], Ev[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = nu * H.inucp[min(j0 + u, Np - 1)];
                fast_exp_neg<4>(x, Ev, S.e2tab);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u;
                    if (j < Np && j >= jm) { if (u & 1) s1 = fma(Ev[u], H.up[j], s1); else s0 = fma(Ev[u], H.up[j], s0); }
                }
            }
        }
        H.QsynP[k] = qpref * C.T[L.nm23s + k] * (s0 + s1);
    }
    __syncthreads();
}

// photopion secondaries, LeHaMoC.py:343-352: pairs and photons from dN_pr/dV, neutrinos (x dt) from N_pr
template <int NT>
__device__ void phase_photopion(const Ctx& C, Smem& S, SmemLh& H, const HadTables& T) {
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
    I.species = SP_EP; I.Nout = L.Ng; pion_eval(I, H.work, H.Qpi, 1.0, false, NT);
    I.species = SP_EM; I.Nout = L.Ng; pion_eval(I, H.work, H.Qpi, 1.0, true, NT);
    I.species = SP_2G; I.Nout = L.Ni; pion_eval(I, H.work, H.Qpg, 1.0, false, NT);
    if (C.F.nu) {
        for (int j = tid; j < Np; j += NT) H.lNp[j] = log10(H.Npr[j]);          // LeHaMoC.py:347 passes N_pr itself
        __syncthreads();
        I.Nout = 50;
        I.species = SP_NUMU;  pion_eval(I, H.work, H.Qnu, C.dt, false, NT);
        I.species = SP_ANUMU; pion_eval(I, H.work, H.Qnu, C.dt, true, NT);
        I.species = SP_NUE;   pion_eval(I, H.work, H.Qnu, C.dt, true, NT);
        I.species = SP_ANUE;  pion_eval(I, H.work, H.Qnu, C.dt, true, NT);
    } else {
        for (int k = tid; k < 52; k += NT) H.Qnu[k] = 0.0;
        __syncthreads();
    }
}

// pp secondaries, LeHaMoC.py:354-357, 387-395.  Adds m_e c^2 Q_e_pp onto the photopion pair source H.Qpi, h Q_g_pp onto
// H.Qpg and 2 (Q_nu_e_pp + Q_nu_mu_pp) dt h V onto H.Qnu (each zeroed first when photopion production is off).
// Scratch: the Bethe-Heitler chunk buffer (free after phase_bh_emis) and the solver scratch S.cp / S.dp.
template <int NT>
__device__ void phase_pp(const Ctx& C, Smem& S, SmemLh& H, double n_H, double p_pr) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, Np = L.Np;
    const PpConst P = pp_const();
    double* Ep = H.bhbuf; double* lEg = Ep + pad2(Np); double* lN = lEg + pad2(Np);
    double* Es = S.cp; double* out = S.dp;
    const double mc2e = kHC.m_el * (kHC.c * kHC.c);
    for (int j = tid; j < Np; j += NT) {
        const double g = C.T[L.gp + j];
        Ep[j] = g * kHC.m_pr * (kHC.c * kHC.c) * P.tev;
        lEg[j] = log10(g * P.x_m_pr);
        lN[j] = log10((H.Npr[j] / C.V) / (kHC.m_pr * (kHC.c * kHC.c) * 0.624151));
    }
    if (!C.F.pg_e) {
        for (int j = tid; j < L.Ng; j += NT) H.Qpi[j] = 0.0;
        for (int k = tid; k < L.Ni; k += NT) H.Qpg[k] = 0.0;
    }
    __syncthreads();
    if (C.F.pp_ee) {
        for (int j = tid; j < L.Ng; j += NT) Es[j] = S.gS[j] * kHC.m_el * (kHC.c * kHC.c) * P.tev;
        __syncthreads();
        pp_product(PP_E, L.Ng, Es, Np, Ep, p_pr, n_H, lEg, lN, out, NT);
        for (int j = tid; j < L.Ng; j += NT) H.Qpi[j] += out[j] * mc2e;
        __syncthreads();
    }
    if (C.F.pp_g) {
        for (int k = tid; k < L.Ni; k += NT) Es[k] = kHC.h * S.nuoS[k] * P.tev;
        __syncthreads();
        pp_product(PP_G, L.Ni, Es, Np, Ep, p_pr, n_H, lEg, lN, out, NT);
        for (int k = tid; k < L.Ni; k += NT) H.Qpg[k] += out[k] * kHC.h;
        __syncthreads();
    }
    if (C.F.pp_nu) {
        for (int k = tid; k < 50; k += NT) Es[k] = kHC.h * (H.Enu[k] / kHC.h) * P.tev;     // nu_nu = E_nu/h, LeHaMoC.py:173
        __syncthreads();
        double q = 0.0;
        pp_product(PP_NUE, 50, Es, Np, Ep, p_pr, n_H, lEg, lN, out, NT);
        if (tid < 50) q = 2. * (out[tid] * C.dt) * kHC.h * C.V;
        __syncthreads();
        pp_product(PP_NUMU, 50, Es, Np, Ep, p_pr, n_H, lEg, lN, out, NT);
        if (tid < 50) H.Qnu[tid] += q + 2. * (out[tid] * C.dt) * kHC.h * C.V;
        __syncthreads();
    }
}

// Bethe-Heitler pair injection Q_BH_sol(g_el, g_pr, dN_pr/dV, h nu_tot/mc2, photons, ...)[1:-1], LeHaMoC.py:337-340
template <int NT>
__device__ void phase_bh_emis(const Ctx& C, Smem& S, SmemLh& H, const BhStaged& G) {
    const Layout& L = C.L;
    for (int j = threadIdx.x; j < L.Np; j += NT) H.npl[j] = H.Npr[j] / C.V;
    __syncthreads();
    for (int i0 = 0; i0 < L.Ng; i0 += BH_CH)
        bh_chunk(i0, min(i0 + BH_CH, L.Ng), L.Np, L.Nt, S.gS, C.T + L.gp, H.bhw, H.npl, C.T + L.xt, S.Ln, G,
                 H.bhbuf, H.Qbh, NT);
}

// photon equations, LeHaMoC.py:397-407: as the leptonic ones plus gamma-gamma absorption on the synchrotron grid and
// the proton synchrotron source
template <int NT>
__device__ void phase_photon_solves_lh(const Ctx& C, Smem& S, SmemLh& H, double cor, const double* Qpg) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    const double dt = C.dt, cc = kPC.c;
    const double b_ad = C.F.ad ? C.Vexp / C.Rad : 0.0;
    const double esc = cc / C.Rad;
    const int ns = L.Ns - 2, ni = L.Ni - 2;
    for (int j = tid; j < ns; j += NT) {
        const double aS = C.F.ssa ? -fabs(S.aS[j + 1]) : 0.0;
        const double V2 = 1.0 + dt * (esc + b_ad * C.T[L.asm_ + j] - aS * cc + H.aggS[j + 1] * cc);
        const double V3 = -dt * (b_ad * C.T[L.asp_ + j]);
        const double Q = C.F.syn_e ? (S.Qsyn[j + 1] / cor + H.QsynP[j + 1]) : 0.0;
        const double d = S.psyn[j + 1] + 4.0 * M_PI * Q * dt * C.V;
        S.cp[j] = V3 / V2;
        S.dp[j] = d / V2;
    }
    __syncthreads();
    if (C.F.ad) { if (tid < 32) backsub_warp(S.cp, S.dp, S.psyn + 1, ns); }
    else { for (int j = tid; j < ns; j += NT) S.psyn[j + 1] = S.dp[j]; }
    __syncthreads();
    for (int j = tid; j < ni; j += NT) {
        const double aI = C.F.ssa ? -fabs(S.aI[j + 1]) : 0.0;
        const double V2 = 1.0 + dt * (esc + b_ad * C.T[L.aim + j] - aI * cc + S.agg[j + 1] * cc);
        const double V3 = -dt * (b_ad * C.T[L.aip + j]);
        double Q = C.F.ic_e ? S.QIC[j + 1] : 0.0;
        if (Qpg) Q += Qpg[j + 1];
        const double d = S.pic[j + 1] + (Q * dt) * C.V;
        S.cp[j] = V3 / V2;
        S.dp[j] = d / V2;
    }
    __syncthreads();
    if (C.F.ad) { if (tid < 32) backsub_warp(S.cp, S.dp, S.pic + 1, ni); }
    else { for (int j = tid; j < ni; j += NT) S.pic[j + 1] = S.dp[j]; }
    __syncthreads();
}

// gamma-gamma block, LeHaMoC.py:409-428: a_gg on three grids from the photon field of the top of the step; in the first
// two steps the field Q_ee_f integrates is attenuated by one implicit absorption step first.
template <int NT>
__device__ void phase_gg_lh(const Ctx& C, Smem& S, SmemLh& H, int interv) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, NP = L.NP, Ni = L.Ni, Ns = L.Ns, Nt = L.Nt, Ng = L.Ng;
    const int nrows = Ni + Ns + Nt;
    for (int row = tid; row < nrows; row += NT) {
        int r, oa, od, oc, orr; double* out;
        if (row < Ni) { r = row; oa = L.agg_a; od = L.agg_d; oc = L.agg_c; orr = L.agg_r2; out = S.agg; }
        else if (row < Ni + Ns) { r = row - Ni; oa = L.ags_a; od = L.ags_d; oc = L.ags_c; orr = L.ags_r2; out = H.aggS; }
        else { r = row - Ni - Ns; oa = L.agt_a; od = L.agt_d; oc = L.agt_c; orr = L.agt_r2; out = H.atmp; }
        // the nu_syn grid follows the walker's B0 (LeHaMoC.py:160): its rows are not in the batch-wide sample table
        const bool tab = C.ggK && !(row >= Ni && row < Ni + Ns);
        out[r] = tab ? gg_row_sum_tab(S, S.Lm, NP, C.ggK, C.ggT, C.ggJ, C.gg_rows, row)
                       : gg_row_sum(S, S.Lm, Nt, NP, __ldg(C.T + oa + r), __ldg(C.T + od + r), __ldg(C.T + oc + r),
                                    __ldg(C.T + orr + r), true);
    }
    __syncthreads();
    for (int k = tid; k < Nt; k += NT) {
        double nk = S.n[k];
        if (interv < 2 && k >= 1 && k <= Nt - 2) {              // thomas() of the diagonal system LeHaMoC.py:419-425
            const double pt = nk * C.V;
            nk = (pt / (1.0 + C.dt * (H.atmp[k] * kPC.c))) / C.V;
        }
        H.Lm2[k] = log2(nk * kPC.mc2_h);
        H.Ln2[k] = log10(nk);
    }
    __syncthreads();
    for (int r = tid; r < Ng; r += NT) {                            // Q_ee_f, full length; the driver uses [1:-1]
        const double s = C.ggK ? gg_row_sum_tab(S, H.Lm2, NP, C.ggK, C.ggT, C.ggJ, C.gg_rows, nrows + r)
                               : gg_row_sum(S, H.Lm2, Nt, NP, __ldg(C.T + L.qee_a + r), __ldg(C.T + L.qee_d + r),
                                            __ldg(C.T + L.qee_c + r), __ldg(C.T + L.qee_r2 + r), false);
        double ng = 0.0;
        if (2.0 * S.gS[r] > 1.0)
            ng = exp10(interp_eval(H.Ln2, C.TI[L.q_j + r], C.T[L.q_dx + r], C.T[L.q_Dx + r]));
        S.qee[r] = kPC.qee_pref * (ng * s);
    }
    __syncthreads();
}

}  // namespace lhm
