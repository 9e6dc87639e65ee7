// lhm_step.cuh -- the per-timestep kinetic update of one walker, executed by one CTA with the whole
// state resident in shared memory (LeMoC/LeMoC.py:196-296, Fit_emcee/Code.py:183-277).
//
// Phases of a step (Appendix A of SURVEY.md gives the index-exact statement):
//   A  log10 of the photon state                      -> Lsyn, Lic
//   B  photons_tot (np.interp merge)                  -> n, log tables           LeHaMoC_f.py:154
//   C  U_ph_f prefix integral -> Compton loss rate    -> bcom                    LeHaMoC_f.py:160
//   D  electron coefficients + bidiagonal solve       -> N                       LeMoC.py:239-247
//   E  per-gamma vectors for the contractions         -> ne, av, u, v, inuc
//   F  synchrotron/SSA rows + IC pairs                -> Qsyn, aS, aI, QIC       LeHaMoC_f.py:86,101,113
//   G  photon solves                                  -> psyn, pic               LeMoC.py:267-277
//   H  gamma-gamma opacity + pair injection           -> agg, qee                LeHaMoC_f.py:127,140
#pragma once
#include "lhm_device.cuh"

namespace lhm {


struct StepFlags {
    int inj, ad, syn_l, syn_e, ic_l, ic_e, ssa, gg, esc, qee_from_ic, ic_path;
    int pp_l, pp_ee, pp_g, pp_nu;   // LeHaMoC.py driver: pp loss term, pp pairs / pi0 photons / neutrinos
    int pg_l, bh_l, pg_e, nu, bh_e; // LeHaMoC.py driver: photopion / Bethe-Heitler loss, photopion emission, neutrinos,
                                    // Bethe-Heitler pair injection
};

__host__ __device__ inline bool any_hadronic(const StepFlags& F) {
    return F.pg_l || F.bh_l || F.pg_e || F.bh_e || F.pp_l || F.pp_ee || F.pp_g || F.pp_nu;
}

// A walker evolved by a thread-block CLUSTER (run_cluster_kernel): every CTA keeps the whole state and runs the cheap
// serial phases redundantly; the three expensive phases deal their rows / tiles to the CTAs (rank of count) and the
// results are exchanged through distributed shared memory.  {0, 1} is the single-CTA case.
struct Part { int rank, count; int tpr = 1; };      // tpr: lanes that share one row in the *_split phases (power of two <= 8)

// shared-memory carve-up; all pointers are into dynamic shared memory
struct Smem {
    double *N, *psyn, *pic, *agg, *qee;          // state
    double *Lsyn, *Lic;                          // log10 of photon state
    double *n, *Ln, *Lm, *y, *cum;               // total photon field + derived
    double *bcom;                                // Compton loss coefficient per gamma
    double *cp, *dp;                             // solver scratch (size Nmax)
    double *ne, *av, *u, *v, *inuc;              // per-gamma vectors of phase E
    double *Qsyn, *aS, *aI, *QIC;                // contraction results
    double *tri;                                 // [Nt][4] = {1/nu_i, (1/nu_i) 2 ln nu_i, w_i n_i, 1/nu_i^2}
    double *gS, *lgS, *inv4gS, *epsS, *lnu4S, *nuoS;   // IC pair-constant inputs
    double *lxtS;                                // log10(h nu_tot/mc2): abscissae of the gamma-gamma interpolation
    double *ilxS;                                // 1/(lxtS[j+1]-lxtS[j])
    double *e2tab;                               // [64] 2^(i/64): table of fast_exp2
    double *sfx;                                 // [4][Nt] suffix sums of IC path S
    double *partial;                             // [NW][NiPad] IC partial sums
    unsigned short* tasks;                       // IC (o-group, gamma-block) task list, active tasks only
    unsigned short* ti0;                         // first target-frequency index a task has to visit
    unsigned short* ti1;                         // first target-frequency index from which no pair of the task clamps
    int* tbound;                                 // [NW+1] task range of every warp (cost-balanced, static)
};

__host__ __device__ inline int imax3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
__host__ __device__ inline int pad2(int n) { return (n + 1) & ~1; }

__host__ __device__ inline size_t smem_doubles(const Layout& L, int NW) {
    int Ng = L.Ng, Ns = L.Ns, Ni = L.Ni, Nt = L.Nt;
    int Nmax = imax3(imax3(Ng, Ns, Ni), Nt, L.Np > 0 ? imax3(L.Np, 52, 0) : 0);   // LeHaMoC.py driver: proton and neutrino (50) solves too
    size_t d = 0;
    d += pad2(Ng) * 2 + pad2(Ns) + pad2(Ni) * 2;            // state
    d += pad2(Ns) + pad2(Ni);                               // logs
    d += pad2(Nt) * 5;                                      // n Ln Lm y cum
    d += pad2(Ng);                                          // bcom
    d += pad2(Nmax) * 2;                                    // cp dp
    d += pad2(Ng) * 5;                                      // ne av u v inuc
    d += pad2(Ns) * 2 + pad2(Ni) * 2;                       // Qsyn aS aI QIC
    d += (size_t)Nt * 4;                                    // tri
    d += pad2(Ng) * 3 + pad2(Ni) * 3;                       // gS lgS inv4gS epsS lnu4S nuoS
    d += pad2(Nt) * 2 + 64;                                 // lxtS ilxS e2tab
    // `partial` (IC path R) and `sfx` (IC path S) alias the scratch block {Lsyn Lic y cum bcom cp dp}, which
    // is dead while the IC phase runs; pad if that block is too small for them
    size_t scratch = pad2(Ns) + pad2(Ni) + 2 * pad2(Nt) + pad2(Ng) + 2 * pad2(Nmax);
    int NiPad = ((Ni + IC_OG - 1) / IC_OG) * IC_OG;
    size_t need = (size_t)NW * NiPad;
    if ((size_t)4 * pad2(Nt) > need) need = (size_t)4 * pad2(Nt);
    if (need > scratch) d += need - scratch;
    return d;
}
__host__ __device__ inline int ic_max_tasks(const Layout& L) {
    return ((L.Ni + IC_OG - 1) / IC_OG) * ((L.Ng + IC_GB - 1) / IC_GB);
}
__host__ __device__ inline size_t smem_bytes(const Layout& L, int NW) {
    return smem_doubles(L, NW) * sizeof(double) + (size_t)pad2(ic_max_tasks(L)) * sizeof(unsigned short) * 3
           + (size_t)(NW + 1 + 3) / 4 * 4 * sizeof(int);
}

__device__ inline void carve(Smem& S, double* base, const Layout& L, int NW) {
    int Ng = L.Ng, Ns = L.Ns, Ni = L.Ni, Nt = L.Nt;
    int Nmax = imax3(imax3(Ng, Ns, Ni), Nt, L.Np > 0 ? imax3(L.Np, 52, 0) : 0);   // LeHaMoC.py driver: proton and neutrino (50) solves too
    double* p = base;
    auto take = [&](int n) { double* r = p; p += n; return r; };
    S.N = take(pad2(Ng)); S.qee = take(pad2(Ng)); S.psyn = take(pad2(Ns)); S.pic = take(pad2(Ni)); S.agg = take(pad2(Ni));
    S.n = take(pad2(Nt)); S.Ln = take(pad2(Nt)); S.Lm = take(pad2(Nt));
    S.ne = take(pad2(Ng)); S.av = take(pad2(Ng)); S.u = take(pad2(Ng)); S.v = take(pad2(Ng)); S.inuc = take(pad2(Ng));
    S.Qsyn = take(pad2(Ns)); S.aS = take(pad2(Ns)); S.aI = take(pad2(Ni)); S.QIC = take(pad2(Ni));
    S.tri = take(Nt * 4);
    S.gS = take(pad2(Ng)); S.lgS = take(pad2(Ng)); S.inv4gS = take(pad2(Ng));
    S.epsS = take(pad2(Ni)); S.lnu4S = take(pad2(Ni)); S.nuoS = take(pad2(Ni));
    S.lxtS = take(pad2(Nt)); S.ilxS = take(pad2(Nt)); S.e2tab = take(64);
    // scratch block, last: dead during the IC phase, so the IC partial sums / suffix sums live on top of it
    double* scratch0 = p;
    S.Lsyn = take(pad2(Ns)); S.Lic = take(pad2(Ni));
    S.y = take(pad2(Nt)); S.cum = take(pad2(Nt));
    S.bcom = take(pad2(Ng));
    S.cp = take(pad2(Nmax)); S.dp = take(pad2(Nmax));
    S.partial = scratch0;
    S.sfx = scratch0;
    p = base + smem_doubles(L, NW);
    const int mt = pad2(ic_max_tasks(L));
    S.tbound = reinterpret_cast<int*>(p);
    S.tasks = reinterpret_cast<unsigned short*>(S.tbound + (NW + 1 + 3) / 4 * 4);
    S.ti0 = S.tasks + mt;
    S.ti1 = S.ti0 + mt;
}

// Shared memory of run_light_kernel (runs without row, IC and gamma-gamma phases): the three populations and the scratch of
// their three solves; the output's scratch (Lsyn, Lic, n) lives on solver scratch that is dead by then.
__host__ __device__ inline size_t smem_light_bytes(const Layout& L) {
    return (size_t)(pad2(L.Ng) + pad2(L.Ns) + pad2(L.Ni)) * 3 * sizeof(double);
}

// Per-walker context: table pointers + scalars of the current step
struct Ctx {
    Layout L;
    const double* T;      // double tables of this walker
    const int* TI;        // int tables of this walker
    // The same tables as far as they depend on g_el / nu_ic / nu_tot only.  Normally == T / TI; the lock-step kernels of a
    // shared-grid batch (path G) point them at walker 0's copy, so that 1184 walkers read ONE set of cache lines
    // instead of 1184 cold ones.  Entries on the nu_syn grid (it follows the walker's B0), el_inj and the external photon
    // field stay per walker.
    const double* Ts; const int* TIs;
    const double* E;      // [(Ns+Ni), Ng] exp(-nu/nu_c(gamma,B)) of this walker when B is constant in time, else null
    const double2* P;     // IC pair table of this walker (ic_fill_pair_table), or null: recompute per step
    // shared-grid batches (path G): the per-sample geometry of the gamma-gamma rows, tabulated once per batch, or null
    const double* ggK; const double* ggT; const int* ggJ; int gg_rows;      // [gg_batches*GG_U][gg_rows] each
    StepFlags F;
    double R0, B0, m, Vexp, dt, tinit;
    double Rad, B, V;     // current step
};

// ---------------------------------------------------------------------------------------------
// load the static IC inputs and the IC tile schedule built at set-up (once per walker)
template <int NT>
__device__ void ic_prepare(const Ctx& C, Smem& S) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    static_assert(NT / 32 >= IC_NW, "the tile schedule is built for IC_NW warps (further warps sit the IC phase out)");
    for (int j = tid; j < L.Ng; j += NT) {
        S.gS[j] = C.Ts[L.g + j]; S.lgS[j] = C.Ts[L.lg + j]; S.inv4gS[j] = C.Ts[L.inv4g + j];
    }
    for (int o = tid; o < L.Ni; o += NT) {
        S.epsS[o] = C.Ts[L.eps + o]; S.lnu4S[o] = C.Ts[L.lnu4 + o]; S.nuoS[o] = C.Ts[L.nic + o];
    }
    for (int i = tid; i < 64; i += NT) S.e2tab[i] = exp2((double)i / 64.0);
    for (int i = tid; i < L.Nt; i += NT) {
        const double x = C.Ts[L.invnt + i];
        S.lxtS[i] = C.Ts[L.lxt + i];
        S.ilxS[i] = (i < L.Nt - 1) ? 1.0 / (C.Ts[L.lxt + i + 1] - C.Ts[L.lxt + i]) : 0.0;
        S.tri[4 * i + 0] = x;                            // 1/nu_i
        S.tri[4 * i + 1] = x * C.Ts[L.l2nt + i];          // (1/nu_i) 2 ln nu_i
        S.tri[4 * i + 2] = 0.0;                          // p_i = trapz weight x photons (every step)
        S.tri[4 * i + 3] = x * x;                        // 1/nu_i^2
    }
    const int nt = C.TI[L.ic_nt];
    for (int t = tid; t < nt; t += NT) {
        const int tk = C.TI[L.ic_task + t];
        S.tasks[t] = (unsigned short)(((tk >> 16) << 8) | (tk & 255));
        S.ti0[t] = (unsigned short)C.TI[L.ic_i0 + t];
        S.ti1[t] = (unsigned short)C.TI[L.ic_i1 + t];
    }
    for (int k = tid; k <= IC_NW; k += NT) S.tbound[k] = C.TI[L.ic_tb + k];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Phase B: photons_tot (LeHaMoC_f.py:154-156) with the /Volume of LeMoC.py:204.  The external
// fields enter as densities already interpolated on nu_tot (host), because
// 10**interp(log10(bb*V))/V == 10**interp(log10(bb)) up to rounding.
template <int NT>
__device__ void phase_photons(const Ctx& C, Smem& S, bool aux) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    for (int k = tid; k < L.Ns; k += NT) S.Lsyn[k] = log10(S.psyn[k]);
    for (int k = tid; k < L.Ni; k += NT) S.Lic[k] = log10(S.pic[k]);
    __syncthreads();
    const double K = kPC.sigmaT * C.Rad * kPC.mc2 / kPC.h;          // sigmaT*R*m_el*c**2./h  LeHaMoC_f.py:170
    for (int k = tid; k < L.Nt; k += NT) {
        double a = interp_eval(S.Lsyn, C.TI[L.ps_j + k], C.T[L.ps_dx + k], C.T[L.ps_Dx + k]);
        double b = interp_eval(S.Lic, C.TIs[L.pi_j + k], C.Ts[L.pi_dx + k], C.Ts[L.pi_Dx + k]);
        double nk = (exp10(a) + exp10(b)) / C.V + C.T[L.ext + k];
        S.n[k] = nk;
        if (aux) {
            S.Ln[k] = log10(nk);
            S.Lm[k] = log2(nk * kPC.mc2_h);                      // gamma-gamma phases interpolate in log2 (fast_exp2)
            S.y[k] = C.Ts[L.xt + k] * nk * K;
            S.tri[4 * k + 2] = C.Ts[L.wt + k] * nk;               // p_i = trapz weight x photons
        }
    }
    __syncthreads();
}

// The same back-substitution by one warp: the recurrence x_i = dp_i - cp_i x_{i+1} is an affine map per row, so each
// lane composes the maps of a contiguous chunk, a 5-level shuffle scan composes the chunks, and each lane then replays
// its chunk with the reference's own recurrence from the incoming value (rounding differs from the serial order by a
// few ulp only through that incoming value).  All 32 lanes of the calling warp must take part.
__device__ __forceinline__ void backsub_warp(const double* __restrict__ cp, const double* __restrict__ dp,
                                             double* __restrict__ x, int n) {
    const int lane = threadIdx.x & 31;
    const int Cn = (n + 31) >> 5;
    const int s = min(lane * Cn, n), e = min(s + Cn, n);
    double D = 0.0, M = 1.0;                                   // x_s = D + M * x_e
    for (int i = e - 1; i >= s; --i) {
        const double c = (i == n - 1) ? 0.0 : cp[i];           // the last row's c never enters (LeHaMoC_f.py:258)
        D = fma(-c, D, dp[i]);
        M = -c * M;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double Dn = __shfl_down_sync(0xffffffffu, D, d), Mn = __shfl_down_sync(0xffffffffu, M, d);
        if (lane + d < 32) { D = fma(M, Dn, D); M *= Mn; }
    }
    double xv = __shfl_down_sync(0xffffffffu, D, 1);           // x at the start of the next chunk
    if (lane == 31) xv = 0.0;
    for (int i = e - 1; i >= s; --i) {
        xv = (i == n - 1) ? dp[i] : dp[i] - cp[i] * xv;
        x[i] = xv;
    }
}

// exclusive prefix sum out[m] = sum_{i<m} v[i] by one warp (chunk per lane + shuffle scan); out may not alias v
__device__ __forceinline__ void exclusive_prefix_warp(const double* __restrict__ v, double* __restrict__ out, int n) {
    const int lane = threadIdx.x & 31;
    const int Cn = (n + 31) >> 5;
    const int s = min(lane * Cn, n), e = min(s + Cn, n);
    double tot = 0.0;
    for (int i = s; i < e; ++i) tot += v[i];
    double incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    double run = incl - tot;
    for (int i = s; i < e; ++i) { out[i] = run; run += v[i]; }
}

// Phase C: U_ph_f (LeHaMoC_f.py:160-172) -> b_Com (LeMoC.py:232)
template <int NT>
__device__ void phase_uph(const Ctx& C, Smem& S, double* U_out /*smem or null*/) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    const int Nt = L.Nt;
    // exclusive prefix of the trapezoids of np.trapz(y, x): cum[m] = sum_{i<m} (x[i+1]-x[i])*(y[i+1]+y[i])/2
    for (int m = tid; m < Nt; m += NT)
        S.dp[m] = (m < Nt - 1) ? (C.Ts[L.xt + m + 1] - C.Ts[L.xt + m]) * (S.y[m + 1] + S.y[m]) / 2.0 : 0.0;
    __syncthreads();
    if (tid < 32) exclusive_prefix_warp(S.dp, S.cum, Nt);
    __syncthreads();
    const double K = kPC.sigmaT * C.Rad * kPC.mc2 / kPC.h;
    const double pre = kPC.mc2 / (kPC.sigmaT * C.Rad);
    // the `else` branch value U_ph_tot integrates nu*n*K over nu (not x) on nu[:-1]  LeHaMoC_f.py:162
    for (int j = tid; j < L.Ng; j += NT) {
        int idx = C.TIs[L.u_idx + j];
        double U;
        if (idx >= 1) {
            double v = interp_eval(S.Ln, C.TIs[L.u_j + j], C.Ts[L.u_dx + j], C.Ts[L.u_Dx + j]);
            double xT = C.Ts[L.u_xT + j];
            double yT = xT * exp10(v) * K;
            double xm = C.Ts[L.xt + idx - 1];
            U = pre * (S.cum[idx - 1] + (xT - xm) * (yT + S.y[idx - 1]) / 2.0);
        } else if (idx == 0) {
            U = 0.0;       // nu_T below the grid: the reference would raise (max of empty); unreachable
        } else {
            double s = 0.0;
            for (int i = 0; i < Nt - 2; ++i) {
                double y0 = C.Ts[L.nt + i] * S.n[i] * K, y1 = C.Ts[L.nt + i + 1] * S.n[i + 1] * K;
                s += (C.Ts[L.nt + i + 1] - C.Ts[L.nt + i]) * (y1 + y0) / 2.0;
            }
            U = pre * s;
        }
        if (U_out) U_out[j] = U;
        S.bcom[j] = 4.0 / 3.0 * kPC.sigmaT * (kPC.c * U) / kPC.mc2;   // LeMoC.py:232
    }
    __syncthreads();
}

// serial upper-bidiagonal back-substitution, exactly thomas() with a == 0 (LeHaMoC_f.py:250-262)
__device__ __forceinline__ void backsub(const double* __restrict__ cp, const double* __restrict__ dp,
                                        double* __restrict__ x, int n) {
    double xv = dp[n - 1];
    x[n - 1] = xv;
    for (int i = n - 2; i >= 0; --i) {
        xv = dp[i] - cp[i] * xv;
        x[i] = xv;
    }
}

// Phase D: electron equation (LeMoC.py:206-247)
template <int NT>
__device__ void phase_electrons(const Ctx& C, Smem& S) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    const int n = L.Ng - 2;
    const double dt = C.dt;
    const double b_ad = C.F.ad ? C.Vexp / C.Rad : 0.0;
    const double b_syn = C.F.syn_l ? kPC.bsyn_pref * (C.B * C.B) : 0.0;
    const double esc = kPC.c / C.Rad * (double)C.F.esc;
    for (int j = tid; j < n; j += NT) {
        double smj = C.Ts[L.sm + j], spj = C.Ts[L.sp + j];
        double syn_m = b_syn * smj, syn_p = b_syn * spj;
        double ic_m = 0.0, ic_p = 0.0;
        if (C.F.ic_l) { ic_m = S.bcom[j + 1] * smj; ic_p = S.bcom[j + 2] * spj; }
        double ad_m = b_ad * C.Ts[L.am + j], ad_p = b_ad * C.Ts[L.ap + j];
        double V2 = 1.0 + dt * (esc + syn_m + ic_m + ad_m);
        double V3 = -dt * (syn_p + ic_p + ad_p);
        double d = S.N[j + 1];
        if (C.F.inj) d += C.T[L.elinj + j + 1] * dt;
        d += (S.qee[j + 1] * dt) * C.V;
        S.cp[j] = V3 / V2;
        S.dp[j] = d / V2;
    }
    __syncthreads();
    if (tid < 32) backsub_warp(S.cp, S.dp, S.N + 1, n);
    __syncthreads();
}

// Phase E: per-gamma vectors consumed by the contractions
template <int NT>
__device__ void phase_vectors(const Ctx& C, Smem& S) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    const int Ng = L.Ng;
    const double dgl = C.Ts[L.lg + 1] - C.Ts[L.lg + 0];            // dg_l_el  LeMoC.py:118
    for (int j = tid; j < Ng; j += NT) S.ne[j] = S.N[j] / C.V;
    __syncthreads();
    for (int j = tid; j < Ng; j += NT) {
        double gj = S.gS[j];
        double nej = S.ne[j];
        double nuc = kPC.nuc_pref * C.B * (gj * gj);                // nu_c(g, B)  LeHaMoC_f.py:78-79
        S.inuc[j] = 1.0 / nuc;
        S.av[j] = C.Ts[L.wg + j] * (nej * (1.0 / gj));                 // trapz weight x Np*g**(-1.)
        S.u[j] = C.Ts[L.wg + j] * (nej * C.Ts[L.g13 + j]);
        double vj = 0.0;
        if (j >= 1 && j <= Ng - 2) {
            double D = (S.ne[j + 1] - nej) / dgl - 2.0 * nej;        // np.diff(N_el[1:])/dg_l_el-2.*N_el[1:-1]
            vj = C.Ts[L.wgi + j] * (pow(2.0 * nuc, -1.0 / 3.0) * D);
        }
        S.v[j] = vj;
    }
    __syncthreads();
}

// exp(-x) for U values x >= 0 at once, written stage by stage so that the U dependency chains interleave.
// -x = n ln2/64 + r, |r| <= ln2/128 (two-term Cody-Waite reduction), exp(r) by a degree-5 polynomial, 2^(n mod 64 / 64)
// from the 64-entry table, 2^(n div 64) by two exponent-field multiplies (gradual underflow kept); ~1e-15 relative.
// x > 745.2 gives exactly 0, like np.exp(-x) in IEEE double.
template <int U>
__device__ __forceinline__ void fast_exp_neg(const double (&x)[U], double (&out)[U], const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;                      // 2^52 + 2^51
    const double L2E64 = -92.332482616893658071;                  // -64/ln 2
    const double LN2_64_HI = 0.01083042469326756, LN2_64_LO = 2.9815858269852933e-12;   // ln2/64 = HI + LO, HI has 21 trailing zero bits (n*HI exact)
    double zm[U], r[U], p[U];
    int n[U];
#pragma unroll
    for (int u = 0; u < U; ++u) zm[u] = fma(x[u], L2E64, MAGIC);
#pragma unroll
    for (int u = 0; u < U; ++u) { n[u] = __double2loint(zm[u]); zm[u] -= MAGIC; }
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(zm[u], -LN2_64_HI, -x[u]);      // -x - n ln2/64
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(zm[u], -LN2_64_LO, r[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(1.0 / 120.0, r[u], 1.0 / 24.0);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], 1.0 / 6.0);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], 0.5);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] *= tab[n[u] & 63];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int k = n[u] >> 6, k1 = k >> 1, k2 = k - k1;
        p[u] *= __hiloint2double((k1 + 1023) << 20, 0);
        p[u] *= __hiloint2double((k2 + 1023) << 20, 0);
        out[u] = (x[u] > 745.2) ? 0.0 : p[u];
    }
}

// exp(-x) for 0 <= x < SYN_XSMALL by its series: the truncation x^5/120 < 1e-17 is below the rounding of the result, so
// the value agrees with exp() to the last bit or two.  In the constant-B mode the part of a row where nu/nu_c has become
// this small (a third of the non-zero table at Test 1) is not tabulated at all: four FMAs replace an 8-byte stream from
// HBM, which is what bounds the phase when a whole batch reaches it in lock-step (path G).
constexpr double SYN_XSMALL = 1e-3;
__device__ __forceinline__ double exp_neg_small(double x) {
    double p = fma(x, 1.0 / 24.0, -1.0 / 6.0);
    p = fma(p, x, 0.5);
    p = fma(p, x, -1.0);
    return fma(p, x, 1.0);
}

// Phase F1: synchrotron emissivity + SSA, one THREAD per frequency row, serial over gamma with two independent
// accumulator chains.  exp(-nu/nu_c) is shared between Q_syn_space (nu/(a_cr g^2)) and aSSA (nu/nu_c(g,B)) on the
// nu_syn grid.  When B is constant in time the exponentials come from the per-walker table built at set-up
// (transposed [gamma][row] so that the 32 rows of a warp read one coalesced line per gamma); leading gammas
// whose exponential underflowed to exactly 0 are skipped.  Both modes use the same expressions and summation order.
template <int NT>
__device__ void phase_syn_rows(const Ctx& C, Smem& S, bool all_rows) {
    const Layout& L = C.L;
    const int Ng = L.Ng, Ns = L.Ns, Ni = L.Ni;
    const bool need_q = C.F.syn_e, need_a = C.F.ssa;
    if (!need_q && !need_a) return;
    constexpr int NR = 32;                                   // gamma stride inside a 32-row block of the table (ecache_index)
    const int nrows = Ns + (need_a ? Ni : 0);
    const double qpref = kPC.C_syn * pow(C.B, 2.0 / 3.0);
    const double apref = kPC.ssa_pref * C.B;
    for (int row = threadIdx.x; row < nrows; row += NT) {
        const bool is_syn = row < Ns;
        const int k = is_syn ? row : row - Ns;
        if (!all_rows) {        // the drivers only consume [1:] / [1:-1]  (LeMoC.py:268-276)
            if (k == 0) continue;
            if (is_syn ? (k > Ns - 2) : (k > Ni - 2)) continue;
        }
        const double nu = is_syn ? C.T[L.nsyn + k] : C.Ts[L.nic + k];
        const bool do_q = is_syn && need_q && (k < Ns - 1);
        const int jm = do_q ? C.TI[L.jm_syn + k] : Ng;       // Q_syn_space mask: gammas below jm emit nothing
        double sq0 = 0.0, sq1 = 0.0, sa0 = 0.0, sa1 = 0.0;
        if (C.E) {
            const double* __restrict__ Et = C.E + ecache_index(row, 0, Ng);
            int j = C.TI[L.jz + row] & ~1;
            const int jx = C.TI[L.jx + row];                 // even; from here on nu/nu_c < SYN_XSMALL: series, no table
#pragma unroll 4
            for (; j < jx; j += 2) {
                const double E0 = __ldg(Et + (size_t)j * NR), E1 = __ldg(Et + (size_t)(j + 1) * NR);
                if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                if (j + 1 >= jm) sq1 = fma(E1, S.u[j + 1], sq1);
                sa0 = fma(E0, S.v[j], sa0);
                sa1 = fma(E1, S.v[j + 1], sa1);
            }
#pragma unroll 2
            for (; j + 1 < Ng; j += 2) {
                const double E0 = exp_neg_small(nu * S.inuc[j]), E1 = exp_neg_small(nu * S.inuc[j + 1]);
                if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                if (j + 1 >= jm) sq1 = fma(E1, S.u[j + 1], sq1);
                sa0 = fma(E0, S.v[j], sa0);
                sa1 = fma(E1, S.v[j + 1], sa1);
            }
            if (j < Ng) {                                    // odd Ng: the last gamma, tabulated whatever its x
                const double x = nu * S.inuc[j];
                const double E0 = (x < SYN_XSMALL) ? exp_neg_small(x) : __ldg(Et + (size_t)j * NR);
                if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                sa0 = fma(E0, S.v[j], sa0);
            }
        } else {
            // recompute the exponentials: SYN_U gammas side by side, stage by stage (independent chains interleave).
            // x = nu/nu_c falls along the gamma grid, so the row splits like the table does: [0, jz) exp(-x) == 0 exactly
            // (x > 745.2), skipped; [jz, jx) the staged exponential; [jx, Ng) x < SYN_XSMALL, the four-term series.
            constexpr int SYN_U = 4;
            int lo = 0, hi = Ng;                                // first j with x <= 745.2
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (nu * S.inuc[mid] > 745.2) lo = mid + 1; else hi = mid; }
            const int jz = lo & ~(SYN_U - 1);
            hi = Ng;                                            // first j with x < SYN_XSMALL, rounded up to a multiple of SYN_U
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (!(nu * S.inuc[mid] < SYN_XSMALL)) lo = mid + 1; else hi = mid; }
            const int jx = min((lo + SYN_U - 1) & ~(SYN_U - 1), Ng);
            int j0 = jz;
            for (; j0 < jx; j0 += SYN_U) {
                double x[SYN_U], Ev[SYN_U];
#pragma unroll
                for (int u = 0; u < SYN_U; ++u) x[u] = nu * S.inuc[min(j0 + u, Ng - 1)];
                fast_exp_neg<SYN_U>(x, Ev, S.e2tab);
#pragma unroll
                for (int u = 0; u < SYN_U; ++u) {
                    const int j = j0 + u;
                    if (j < Ng) {
                        if (u & 1) {
                            if (j >= jm) sq1 = fma(Ev[u], S.u[j], sq1);
                            sa1 = fma(Ev[u], S.v[j], sa1);
                        } else {
                            if (j >= jm) sq0 = fma(Ev[u], S.u[j], sq0);
                            sa0 = fma(Ev[u], S.v[j], sa0);
                        }
                    }
                }
            }
#pragma unroll 2
            for (int j = j0; j < Ng; j += 2) {                  // j0 is even: the two chains keep their parity
                const double E0 = exp_neg_small(nu * S.inuc[j]);
                if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                sa0 = fma(E0, S.v[j], sa0);
                if (j + 1 < Ng) {
                    const double E1 = exp_neg_small(nu * S.inuc[j + 1]);
                    if (j + 1 >= jm) sq1 = fma(E1, S.u[j + 1], sq1);
                    sa1 = fma(E1, S.v[j + 1], sa1);
                }
            }
        }
        const double sq = sq0 + sq1, sa = sa0 + sa1;
        if (is_syn) {
            if (k < Ns - 1) S.Qsyn[k] = do_q ? qpref * C.T[L.nm23s + k] * sq : 0.0;
            S.aS[k] = need_a ? apref * C.T[L.nm53s + k] * sa : 0.0;
        } else {
            S.aI[k] = apref * C.Ts[L.nm53i + k] * sa;
        }
    }
}

// Phase F2, path R: Blumenthal-Gould/Jones triple sum with the kernel recomputed in registers.
//   F = 2q ln q + (1+2q)(1-q) + G(1-q),  q = A/nu_i,  A = nu_o/(4 g (g-eps)),  G = w^2/(2(1+w)), w = eps/(g-eps)
//     = G1 + q*(c1 - 2 ln nu_i - 2q),    G1 = 1+G,    c1 = 1 - G + 2 ln A                    (exact regrouping)
//     = G1 + a1*x_i + a2*(x_i 2 ln nu_i) + a3*x_i^2,   x_i = 1/nu_i,  a1 = A c1, a2 = -A, a3 = -2 A^2
// per (o, gamma) pair: one reciprocal + one log; per (o, gamma, nu_i) element: 3 DFMA for F and one DFMA for the
// accumulation, the latter predicated on the sign bit of F (that is `function_IC[function_IC < 0.] = 0.`,
// LeHaMoC_f.py:121: an integer compare on the high word instead of a DSETP/DMNMX on the FP64 pipe).
// the four per-(nu_o, gamma) constants of F; an inactive pair (gamma <= eps_o, LeHaMoC_f.py:117, or padding) gets
// zeros, i.e. F == 0 for every nu_i
__device__ __forceinline__ void ic_pair_constants(bool valid, double eps, double nuo, double lnu4, double g, double inv4g,
                                                  double lg, double& G1, double& a1, double& a2, double& a3) {
    const bool act = valid && (g > eps);
    const double d = act ? (g - eps) : 1.0;
    const double rr = 1.0 / d;
    const double Gv = 2.0 * eps * eps * rr * inv4g;
    const double Av = nuo * rr * inv4g;
    const double c1 = (1.0 - Gv) + 2.0 * (lnu4 - lg - log(d));
    G1 = act ? 1.0 + Gv : 0.0;
    a1 = act ? Av * c1 : 0.0;
    a2 = act ? -Av : 0.0;
    a3 = act ? -2.0 * Av * Av : 0.0;
}

__device__ __forceinline__ void ic_acc_if_nonneg(double& acc, double p, double F) {
    asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %1, 0;\n\t@q fma.rn.f64 %0, %2, %3, %0;\n\t}"
        : "+d"(acc) : "r"(__double2hiint(F)), "d"(p), "d"(F));
}

template <int NT>
__device__ void phase_ic_R(const Ctx& C, Smem& S, Part cta = Part{0, 1}) {
    const Layout& L = C.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ng = L.Ng, No = L.Ni - 1, Ntar = L.Nt - 1;
    const int NiPad = ((L.Ni + IC_OG - 1) / IC_OG) * IC_OG;
    if (warp >= IC_NW) return;                      // 512-thread callers: the static schedule feeds IC_NW warps
    double* part = S.partial + warp * NiPad;
    for (int o = lane; o < NiPad; o += 32) part[o] = 0.0;
    __syncwarp();
    int t0 = S.tbound[warp], t1 = S.tbound[warp + 1];
    if (cta.count > 1) {                            // this CTA's contiguous share of the warp's static tile range
        const int n = t1 - t0;
        t1 = t0 + (int)(((long long)n * (cta.rank + 1)) / cta.count);
        t0 = t0 + (int)(((long long)n * cta.rank) / cta.count);
    }
    const int osub = lane >> 2, gsub = lane & 3;
    int cur_og = -1;
    double run = 0.0;
    const double2* tri2 = reinterpret_cast<const double2*>(S.tri);
    for (int t = t0; t < t1; ++t) {
        const int og = S.tasks[t] >> 8, gb = S.tasks[t] & 255;
        const int i0 = S.ti0[t], i1 = S.ti1[t];
        if (og != cur_og) {
            if (cur_og >= 0) {
                run += __shfl_xor_sync(0xffffffffu, run, 1);
                run += __shfl_xor_sync(0xffffffffu, run, 2);
                if (gsub == 0) part[cur_og * IC_OG + osub] = run;
            }
            cur_og = og; run = 0.0;
        }
        double a1[IC_PT], a2[IC_PT], a3[IC_PT], G1[IC_PT], acc[IC_PT];
        if (C.P) {
            // per-pair constants from the table built at set-up (tile-major, one coalesced 512-byte line per load)
            const double2* __restrict__ Pt = C.P + (size_t)t * (IC_PT * 2 * 32) + lane;
#pragma unroll
            for (int r = 0; r < IC_PT; ++r) {
                const double2 q01 = __ldg(Pt + (r * 2) * 32), q23 = __ldg(Pt + (r * 2 + 1) * 32);
                G1[r] = q01.x; a1[r] = q01.y; a2[r] = q23.x; a3[r] = q23.y;
                acc[r] = 0.0;
            }
        } else {
            const int o = og * IC_OG + osub;
            const bool ovalid = o < No;
            const double eps = ovalid ? S.epsS[o] : 0.0;
            const double nuo = ovalid ? S.nuoS[o] : 0.0;
            const double lnu4 = ovalid ? S.lnu4S[o] : 0.0;
#pragma unroll
            for (int r = 0; r < IC_PT; ++r) {
                const int j = gb * IC_GB + r * IC_GL + gsub;
                ic_pair_constants(ovalid && (j < Ng), eps, nuo, lnu4, S.gS[min(j, Ng - 1)], S.inv4gS[min(j, Ng - 1)],
                                  S.lgS[min(j, Ng - 1)], G1[r], a1[r], a2[r], a3[r]);
                acc[r] = 0.0;
            }
        }
        // [i0, i1): some pair of the tile may still have q >= 1, i.e. F <= 0 -> clamp
        for (int i = i0; i < i1; ++i) {
            const double2 u = tri2[2 * i];           // {1/nu_i, (1/nu_i) 2 ln nu_i}
            const double2 v = tri2[2 * i + 1];       // {w_i n_i, 1/nu_i^2}
#pragma unroll
            for (int r = 0; r < IC_PT; ++r) {
                double Fv = fma(a1[r], u.x, G1[r]);
                Fv = fma(a2[r], u.y, Fv);
                Fv = fma(a3[r], v.y, Fv);
                ic_acc_if_nonneg(acc[r], v.x, Fv);   // function_IC[function_IC < 0.] = 0.
            }
        }
        // [i1, Ntar): q < 1 for every active pair of the tile, so F > 0 and the clamp is the identity
#pragma unroll 4
        for (int i = i1; i < Ntar; ++i) {
            const double2 u = tri2[2 * i];
            const double2 v = tri2[2 * i + 1];
#pragma unroll
            for (int r = 0; r < IC_PT; ++r) {
                double Fv = fma(a1[r], u.x, G1[r]);
                Fv = fma(a2[r], u.y, Fv);
                Fv = fma(a3[r], v.y, Fv);
                acc[r] = fma(v.x, Fv, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < IC_PT; ++r)      // an inactive pair has F == 0 throughout, so acc == 0 whatever the weight
            run = fma(S.av[min(gb * IC_GB + r * IC_GL + gsub, Ng - 1)], acc[r], run);
    }
    if (cur_og >= 0) {
        run += __shfl_xor_sync(0xffffffffu, run, 1);
        run += __shfl_xor_sync(0xffffffffu, run, 2);
        if (gsub == 0) part[cur_og * IC_OG + osub] = run;
    }
}

// path S: first target index with nu_i > A (q < 1) -- arithmetic guess on the log-uniform nu_tot grid, then the exact
// fix-up with the comparison the clamp makes.  x[i * stride] = 1/nu_i.  Used by the time loop and by the set-up table.
__device__ __forceinline__ int ic_s_cut_index(double lnA, double Av, double lnu0, double inv_dl, const double* x, int stride,
                                              int Ntar) {
    int is = (int)floor((lnA - lnu0) * inv_dl) + 1;
    is = max(0, min(is, Ntar));
    while (is > 0 && !(Av * x[stride * (is - 1)] >= 1.0)) --is;          // nu_{is-1} > A  -> move left
    while (is < Ntar && (Av * x[stride * is] >= 1.0)) ++is;              // nu_is <= A      -> move right
    return is;
}

// The pair loop of path S from the set-up table.  What a pair needs beyond a reciprocal and a dozen multiply-adds depends
// on the grids only: ln(gamma - eps) and the cut index i* were tabulated at set-up by the expressions of phase_ic_S's own
// loop (setup_kernel; a path-S handle keeps them where path R keeps its pair constants), so the pairs of a row carry no
// logarithm, no search and no data-dependent branch.  U blocks of 32 gammas side by side -- all their table loads are
// issued before the first is used, nothing between them is conditional -- with the arithmetic and the summation order of
// the loop it replaces.
template <int U>
__device__ __forceinline__ double ic_s_group(const Smem& S, const double* __restrict__ lnd_o,
                                             const unsigned short* __restrict__ is_o, const double* P0, const double* P1,
                                             const double* P2, const double* P3, double eps, double nuo, double lnu4,
                                             int j0, int Ng, double sum) {
    double lnd[U]; int is[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = min(j0 + 32 * u, Ng - 1);
        lnd[u] = __ldg(lnd_o + j);
        is[u] = __ldg(is_o + j);
    }
    double inner[U]; bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = min(j0 + 32 * u, Ng - 1);
        const double gj = S.gS[j];
        act[u] = (j0 + 32 * u < Ng) && (gj > eps);
        const double d = gj - eps, rr = 1.0 / d, i4 = S.inv4gS[j];
        const double Gv = 2.0 * eps * eps * rr * i4;
        const double Av = nuo * rr * i4;
        const double lnA = lnu4 - S.lgS[j] - lnd[u];
        const double G1 = 1.0 + Gv, c1 = (1.0 - Gv) + 2.0 * lnA;
        inner[u] = G1 * P0[is[u]] + Av * (c1 * P1[is[u]] - P2[is[u]] - 2.0 * Av * P3[is[u]]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (act[u]) sum = fma(S.av[min(j0 + 32 * u, Ng - 1)], inner[u], sum);
    return sum;
}

template <int NT>
__device__ __forceinline__ void phase_ic_S_tab(const Ctx& C, Smem& S) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = NT / 32;
    const int Ng = L.Ng, No = L.Ni - 1;
    const int NtP = pad2(L.Nt);
    const double* P0 = S.sfx, *P1 = S.sfx + NtP, *P2 = S.sfx + 2 * NtP, *P3 = S.sfx + 3 * NtP;
    const double* __restrict__ lndT = reinterpret_cast<const double*>(C.P);
    const unsigned short* __restrict__ isT = reinterpret_cast<const unsigned short*>(lndT + (size_t)No * Ng);
    constexpr int SU = 8;
    const int nb = (Ng + 31) >> 5;                               // blocks of 32 gammas in a row
    const int nfull = (nb / SU) * SU * 32;                       // gammas covered by whole groups of SU blocks
    const int tail = nb - (nb / SU) * SU;                        // the rest in groups of 4, 2 and 1 blocks
    for (int o = warp; o < No; o += NW) {
        const double eps = S.epsS[o], nuo = S.nuoS[o], lnu4 = S.lnu4S[o];
        const double* __restrict__ lnd_o = lndT + (size_t)o * Ng;
        const unsigned short* __restrict__ is_o = isT + (size_t)o * Ng;
        double sum = 0.0;
        int j0 = lane;
        for (; j0 < nfull; j0 += 32 * SU)
            sum = ic_s_group<SU>(S, lnd_o, is_o, P0, P1, P2, P3, eps, nuo, lnu4, j0, Ng, sum);
        if (tail & 4) { sum = ic_s_group<4>(S, lnd_o, is_o, P0, P1, P2, P3, eps, nuo, lnu4, j0, Ng, sum); j0 += 128; }
        if (tail & 2) { sum = ic_s_group<2>(S, lnd_o, is_o, P0, P1, P2, P3, eps, nuo, lnu4, j0, Ng, sum); j0 += 64; }
        if (tail & 1) sum = ic_s_group<1>(S, lnd_o, is_o, P0, P1, P2, P3, eps, nuo, lnu4, j0, Ng, sum);
        sum = warp_sum(sum);
        if (lane == 0) S.QIC[o] = kPC.ic_pref * sum;
    }
}

// Phase F2, path S: the nu_i sum collapses to four suffix sums because max(F,0) only cuts the sum at the
// first nu_i > A (F > 0 <=> q < 1) and F is rank-4 separable in (o,gamma) x (i)  (SURVEY.md 7.1b-1).
template <int NT, bool TAB = false>
__device__ void phase_ic_S(const Ctx& C, Smem& S) {
    const Layout& L = C.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = NT / 32;
    const int Ng = L.Ng, No = L.Ni - 1, Ntar = L.Nt - 1;
    const int NtP = pad2(L.Nt);
    // suffix sums P_k[j] = sum_{i>=j} p_i * {1, 1/nu_i, (1/nu_i)*2ln nu_i, 1/nu_i^2};  P_k[Ntar] = 0
    if (tid < 4) {
        double s = 0.0;
        double* P = S.sfx + tid * NtP;
        P[Ntar] = 0.0;
#pragma unroll 4
        for (int i = Ntar - 1; i >= 0; --i) {
            const double p = S.tri[4 * i + 2];
            const double f = (tid == 0) ? 1.0 : (tid == 1) ? S.tri[4 * i] : (tid == 2) ? S.tri[4 * i + 1] : S.tri[4 * i + 3];
            s = fma(p, f, s);
            P[i] = s;
        }
    }
    __syncthreads();
    const double* P0 = S.sfx, *P1 = S.sfx + NtP, *P2 = S.sfx + 2 * NtP, *P3 = S.sfx + 3 * NtP;
    if (TAB && C.P) { phase_ic_S_tab<NT>(C, S); return; }      // only the kernel compiled for path-S handles (run_kernel<S>)
    const double lnu0 = 0.5 * C.Ts[L.l2nt];                              // ln nu_0
    const double inv_dl = (double)(Ntar - 1) / (0.5 * C.Ts[L.l2nt + Ntar - 1] - lnu0);
    for (int o = warp; o < No; o += NW) {
        const double eps = S.epsS[o], nuo = S.nuoS[o], lnu4 = S.lnu4S[o];
        double sum = 0.0;
        for (int j = lane; j < Ng; j += 32) {
            const double gj = S.gS[j];
            if (!(gj > eps)) continue;
            const double d = gj - eps, rr = 1.0 / d, i4 = S.inv4gS[j];
            const double Gv = 2.0 * eps * eps * rr * i4;
            const double Av = nuo * rr * i4;
            const double lnA = lnu4 - S.lgS[j] - log(d);
            const int is = ic_s_cut_index(lnA, Av, lnu0, inv_dl, S.tri, 4, Ntar);
            const double G1 = 1.0 + Gv, c1 = (1.0 - Gv) + 2.0 * lnA;
            // sum_i p_i [G1 + q(c1 - 2 ln nu_i - 2 q)],  q = A/nu_i
            double inner = G1 * P0[is] + Av * (c1 * P1[is] - P2[is] - 2.0 * Av * P3[is]);
            sum = fma(S.av[j], inner, sum);
        }
        sum = warp_sum(sum);
        if (lane == 0) S.QIC[o] = kPC.ic_pref * sum;
    }
}

template <int NT>
__device__ void phase_ic_finish_R(const Ctx& C, Smem& S) {
    const int NW = IC_NW;
    const int No = C.L.Ni - 1;
    const int NiPad = ((C.L.Ni + IC_OG - 1) / IC_OG) * IC_OG;
    for (int o = threadIdx.x; o < No; o += NT) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += S.partial[w * NiPad + o];
        S.QIC[o] = kPC.ic_pref * s;                                         // sigmaT*c*0.75*trapz(...)
    }
}

// Phase G: photon equations (LeMoC.py:267-277)
template <int NT>
__device__ void phase_photon_solves(const Ctx& C, Smem& S, double cor, const double* QsynP = nullptr) {
    const Layout& L = C.L;
    const int tid = threadIdx.x;
    const double dt = C.dt, cc = kPC.c;
    const double b_ad = C.F.ad ? C.Vexp / C.Rad : 0.0;
    const double esc = cc / C.Rad;
    const int ns = L.Ns - 2, ni = L.Ni - 2;
    // The two photon equations are independent: both coefficient sets are formed in one pass (the IC set in bcom / Lic,
    // which are dead between the electron solve and the next photon field) and, when the adiabatic term makes the
    // systems bidiagonal, warps 0 and 1 back-substitute side by side.  Same arithmetic per element as one after the other.
    double* cp2 = S.bcom;
    double* dp2 = S.Lic;
    for (int j = tid; j < ns; j += NT) {                                 // synchrotron photons -> cp/dp[0..ns)
        double aS = C.F.ssa ? -fabs(S.aS[j + 1]) : 0.0;                  // -np.absolute(aSSA)  LeMoC.py:261
        double V2 = 1.0 + dt * (esc + b_ad * C.T[L.asm_ + j] - aS * cc);
        double V3 = -dt * (b_ad * C.T[L.asp_ + j]);
        double Q = C.F.syn_e ? S.Qsyn[j + 1] / cor : 0.0;                // np.divide(Q_syn, cor_factor)
        double d = S.psyn[j + 1] + 4.0 * M_PI * (Q * dt) * C.V;
        if (QsynP) d += 4.0 * M_PI * (QsynP[j + 1] * dt) * C.V;          // proton synchrotron, Code.py:515
        S.cp[j] = V3 / V2;
        S.dp[j] = d / V2;
    }
    for (int j = tid; j < ni; j += NT) {                                 // IC photons -> cp2/dp2[0..ni)
        double aI = C.F.ssa ? -fabs(S.aI[j + 1]) : 0.0;
        double V2 = 1.0 + dt * (esc + b_ad * C.Ts[L.aim + j] + S.agg[j + 1] * cc - aI * cc);
        double V3 = -dt * (b_ad * C.Ts[L.aip + j]);
        double Q = C.F.ic_e ? S.QIC[j + 1] : 0.0;
        double d = S.pic[j + 1] + (Q * dt) * C.V;
        cp2[j] = V3 / V2;
        dp2[j] = d / V2;
    }
    __syncthreads();
    if (C.F.ad) {
        if (tid < 32) backsub_warp(S.cp, S.dp, S.psyn + 1, ns);
        else if (tid < 64) backsub_warp(cp2, dp2, S.pic + 1, ni);
    } else {                                                             // V3 == 0: diagonal systems
        for (int j = tid; j < ns; j += NT) S.psyn[j + 1] = S.dp[j];
        for (int j = tid; j < ni; j += NT) S.pic[j + 1] = dp2[j];
    }
    __syncthreads();
}

// 2^x for x finite, -inf or NaN (the latter two give 0, which is what 10**np.interp yields for the -inf ordinates
// of an empty photon bin); x <= ~1000.  x = n/64 + f, |f| <= 1/128: 2^f by a degree-5 polynomial (truncation 4e-17),
// 2^(n mod 64 / 64) from a 64-entry table, 2^(n div 64) by two exponent-field multiplies (gradual underflow kept).
__device__ __forceinline__ double fast_exp2(double x, const double* __restrict__ tab) {
    x = fmax(x, -1100.0);                                         // NaN and -inf -> -1100 (fmax drops the NaN)
    const double MAGIC = 6755399441055744.0;                      // 2^52 + 2^51: round-to-nearest-integer by addition
    const double zm = fma(x, 64.0, MAGIC);
    const int n = __double2loint(zm);
    const double f = fma(zm - MAGIC, -1.0 / 64.0, x);             // exact
    const double r = f * 0.693147180559945309417;
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= tab[n & 63];
    const int k = n >> 6;
    if (k >= -1020) return p * __hiloint2double((k + 1023) << 20, 0);     // normal range: one exact scaling
    const int k1 = k >> 1, k2 = k - k1;                                       // gradual underflow: two steps
    p *= __hiloint2double((k1 + 1023) << 20, 0);
    p *= __hiloint2double((k2 + 1023) << 20, 0);
    return p;
}

// Phase H: a_gg (LeHaMoC_f.py:127-137) and Q_ee_f (LeHaMoC_f.py:140-151), one THREAD per row, GG_U samples of the
// row evaluated side by side and written stage by stage so that GG_U independent dependency chains interleave (a
// single chain of FP64 operations advances one instruction per ~10 cycles on this pipe).
// The NP resample abscissae of a row are log10-uniform, y_s = a + s*step (np.logspace), and so is the nu_tot grid
// they are interpolated on: the np.interp bracket is found arithmetically and then verified against the actual
// grid doubles.  The static factor of a sample (trapezoid weight x pair-production kernel x x') is rebuilt in
// registers from four per-row numbers (see lhm_setup.cuh): no table is streamed at all.  S.Lm holds log2 of the
// photon field (x mc^2/h), so 10**np.interp(log10 ...) becomes fast_exp2 of the same linear interpolant.
// `Lng` gives n_g: log10 of the field Q_ee_f's n_g is interpolated from.
constexpr int GG_U = 4;
__host__ __device__ inline int gg_batches(int NP) { return (NP + GG_U - 1) / GG_U; }

// The grid-only part of a row: for every batch of GG_U samples e0..e0+GG_U-1 the np.interp bracket j, the position t in
// it and the static factor Kv = w_e * cf * (1 - s_e^-2) * ln s_e, handed to emit(e0, j, t, Kv).  y_e = a + e*step.
// bstart / bstride: the batches bstart, bstart + bstride, ... only (several lanes sharing a row); 0 / 1 is the whole row.
template <class Emit>
__device__ __forceinline__ void gg_row_samples(const Smem& S, int Nt, int NP, double a, double step, double cf, double r2,
                                               bool is_a, Emit&& emit, int bstart = 0, int bstride = 1) {
    const double* __restrict__ lxt = S.lxtS;
    const double* __restrict__ ilx = S.ilxS;
    const double lx0 = lxt[0], lxN = lxt[Nt - 1];
    const double invD = (double)(Nt - 1) / (lxN - lx0);
    const double LN10 = 2.302585092994045684, LN13 = 0.262364264467491052;     // ln 10, ln 1.3
    if (!(step > 0.0)) return;              // x_space[0] < x_space[1]  (LeHaMoC_f.py:133); else the row is 0
    const double dl = step * LN10;                      // spacing of ln x' = trapezoid weight of inner samples
    const double lns0 = is_a ? LN13 : 0.0;              // ln s_0
    double inv[GG_U];                                   // s_e^-2 = s_0^-2 * r2^e
    inv[0] = is_a ? (1.0 / 1.69) : 1.0;
#pragma unroll
    for (int u = 1; u < GG_U; ++u) inv[u] = inv[u - 1] * r2;
    double rU = r2;
#pragma unroll
    for (int u = 1; u < GG_U; ++u) rU *= r2;            // r2^GG_U
    const double dlcf = dl * cf;                        // weight of an interior sample x the row's prefactor
    double ed[GG_U];
#pragma unroll
    for (int u = 0; u < GG_U; ++u) ed[u] = (double)u;
    double rUs = rU;                                    // r2^(GG_U bstride): what a lane's s^-2 advances by per batch
    for (int q = 1; q < bstride; ++q) rUs *= rU;
    for (int q = 0; q < bstart; ++q) {
#pragma unroll
        for (int u = 0; u < GG_U; ++u) { inv[u] *= rU; ed[u] += (double)GG_U; }
    }
    const double edstep = (double)(GG_U * bstride);
    for (int e0 = bstart * GG_U; e0 < NP; e0 += GG_U * bstride) {
        double y[GG_U], Kv[GG_U], t[GG_U];
        int j[GG_U];
#pragma unroll
        for (int u = 0; u < GG_U; ++u) {
            y[u] = fma(ed[u], step, a);
            Kv[u] = dlcf * (1.0 - inv[u]) * fma(ed[u], dl, lns0);
            inv[u] *= rUs;
            ed[u] += edstep;
        }
        // trapezoid end weights and the exact last abscissa: only the first and the last batch (uniform branches)
        if (e0 == 0) Kv[0] *= 0.5;
        if (e0 + GG_U >= NP) {
#pragma unroll
            for (int u = 0; u < GG_U; ++u) {
                const int e = e0 + u;
                if (e == NP - 1) { y[u] = lxN; Kv[u] *= 0.5; }
                else if (e >= NP) { y[u] = lxN; Kv[u] = 0.0; }
            }
        }
#pragma unroll
        for (int u = 0; u < GG_U; ++u) {
            y[u] = fmax(y[u], lx0);                         // np.interp clamps below the grid (t = 0 in segment 0)
            int jj = (int)((y[u] - lx0) * invD);            // both grids are log-uniform: the guess is right
            jj = max(0, min(jj, Nt - 2));                   // except next to a node ...
            double lo = lxt[jj];
            if (!(y[u] >= lo && y[u] < lxt[jj + 1])) {      // ... where it is corrected against the actual doubles
                while (jj > 0 && y[u] < lxt[jj]) --jj;
                while (jj < Nt - 2 && y[u] >= lxt[jj + 1]) ++jj;
                lo = lxt[jj];
            }
            j[u] = jj;
            t[u] = (y[u] - lo) * ilx[jj];
        }
        emit(e0, j, t, Kv);
    }
}

// the state-dependent part of a batch: acc[u] += Kv[u] * 2^(interp of Lm at sample u)
__device__ __forceinline__ void gg_batch_acc(const double* __restrict__ Lm, const double* __restrict__ tab,
                                             const int (&j)[GG_U], const double (&t)[GG_U], const double (&Kv)[GG_U],
                                             double (&acc)[GG_U]) {
    double Lv[GG_U];
#pragma unroll
    for (int u = 0; u < GG_U; ++u) {
        const double f0 = Lm[j[u]], f1 = Lm[j[u] + 1];
        Lv[u] = fma(t[u], f1 - f0, f0);                     // NaN (from -inf ordinates) -> 0 inside fast_exp2
    }
#pragma unroll
    for (int u = 0; u < GG_U; ++u) acc[u] = fma(Kv[u], fast_exp2(Lv[u], tab), acc[u]);
}

// one row: sum_k w_k * cf * (1 - s_k^-2) * ln s_k * 2^(interp of Lm at y_k),  y_k = a + k*step, k = 0..NP-1
__device__ __forceinline__ double gg_row_sum(const Smem& S, const double* __restrict__ Lm, int Nt, int NP, double a,
                                             double step, double cf, double r2, bool is_a) {
    double acc[GG_U];
#pragma unroll
    for (int u = 0; u < GG_U; ++u) acc[u] = 0.0;
    gg_row_samples(S, Nt, NP, a, step, cf, r2, is_a,
                   [&](int, const int (&j)[GG_U], const double (&t)[GG_U], const double (&Kv)[GG_U]) {
                       gg_batch_acc(Lm, S.e2tab, j, t, Kv, acc);
                   });
    double s = acc[0];
#pragma unroll
    for (int u = 1; u < GG_U; ++u) s += acc[u];
    return s;
}

// the same row from the batch-wide table of its sample geometry (shared-grid batches: what gg_row_samples emits does not
// depend on the walker, so it is stored once per batch -- [sample][row], coalesced over the rows of a warp -- and a
// sample costs two shared-memory reads, one FMA and the staged 2^x instead of ~20 further FP64 instructions).  Same
// numbers in the same order as gg_row_sum: bit-identical.
__device__ __forceinline__ double gg_row_sum_tab(const Smem& S, const double* __restrict__ Lm, int NP, const double* __restrict__ gK,
                                                 const double* __restrict__ gT, const int* __restrict__ gJ, int stride, int row) {
    double acc[GG_U];
#pragma unroll
    for (int u = 0; u < GG_U; ++u) acc[u] = 0.0;
    const int nb = gg_batches(NP);
    for (int b = 0; b < nb; ++b) {
        int j[GG_U]; double t[GG_U], Kv[GG_U];
#pragma unroll
        for (int u = 0; u < GG_U; ++u) {
            const size_t o = (size_t)(b * GG_U + u) * stride + row;
            Kv[u] = __ldg(gK + o); t[u] = __ldg(gT + o); j[u] = __ldg(gJ + o);
        }
        gg_batch_acc(Lm, S.e2tab, j, t, Kv, acc);
    }
    double s = acc[0];
#pragma unroll
    for (int u = 1; u < GG_U; ++u) s += acc[u];
    return s;
}

template <int NT>
__device__ void phase_gg(const Ctx& C, Smem& S, const double* Lng, double* agg_out, double* qee_out) {
    const Layout& L = C.L;
    const int NP = L.NP, Ni = L.Ni, Ng = L.Ng, Nt = L.Nt;
    const int nrows = Ni + (Ng - 1);
    for (int row = threadIdx.x; row < nrows; row += NT) {
        const bool is_a = row < Ni;
        const int r = is_a ? row : row - Ni;
        const double a = __ldg(C.Ts + (is_a ? L.agg_a : L.qee_a) + r);
        const double step = __ldg(C.Ts + (is_a ? L.agg_d : L.qee_d) + r);
        const double cf = __ldg(C.Ts + (is_a ? L.agg_c : L.qee_c) + r);
        const double r2 = __ldg(C.Ts + (is_a ? L.agg_r2 : L.qee_r2) + r);
        const double s = C.ggK ? gg_row_sum_tab(S, S.Lm, NP, C.ggK, C.ggT, C.ggJ, C.gg_rows, row)
                               : gg_row_sum(S, S.Lm, Nt, NP, a, step, cf, r2, is_a);
        if (is_a) agg_out[r] = s;
        else {
            double ng = 0.0;
            if (2.0 * S.gS[r] > 1.0)
                ng = exp10(interp_eval(Lng, C.TIs[L.q_j + r], C.Ts[L.q_dx + r], C.Ts[L.q_Dx + r]));
            qee_out[r] = kPC.qee_pref * (ng * s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The two row phases for a walker evolved by a CLUSTER (run_cluster_kernel): the rows are dealt round-robin to the CTAs,
// which leaves a CTA with fewer rows than threads -- so `tpr` lanes share a row (interleaved gamma pairs / sample
// batches, joined by an xor-shuffle tree): the phase shrinks with the cluster size instead of staying one row long.
// serp: odd passes deal the rows to the lane groups in reverse order.  Within either family (nu_syn rows, nu_ic rows) the
// non-zero length of a row falls with its frequency, so a warp that holds the longest rows of one pass holds the shortest of
// the next -- the warps of a CTA reach the barrier behind the phase together (stepG_kernel: two lanes per row, two passes).
// Which warp evaluates a row does not enter its arithmetic.
template <int NT>
__device__ void phase_syn_rows_split(const Ctx& C, Smem& S, Part part, bool serp = false) {
    const Layout& L = C.L;
    const int Ng = L.Ng, Ns = L.Ns, Ni = L.Ni;
    const bool need_q = C.F.syn_e, need_a = C.F.ssa;
    if (!need_q && !need_a) return;
    constexpr int NR = 32;
    const int nrows = Ns + (need_a ? Ni : 0);
    const double qpref = kPC.C_syn * pow(C.B, 2.0 / 3.0);
    const double apref = kPC.ssa_pref * C.B;
    const int tpr = part.tpr, grp = threadIdx.x / tpr, sub = threadIdx.x % tpr, ngrp = NT / tpr;
    for (int base = 0; base < nrows; base += part.count * ngrp) {        // the same trip count for every lane of a warp
        const bool rev = serp && (((base / (part.count * ngrp)) & 1) != 0);
        const int row = base + part.rank + part.count * (rev ? ngrp - 1 - grp : grp);
        const bool is_syn = row < Ns;
        const int k = is_syn ? row : row - Ns;
        bool act = row < nrows && k != 0 && !(is_syn ? (k > Ns - 2) : (k > Ni - 2));   // [1:] / [1:-1], LeMoC.py:268-276
        const bool do_q = act && is_syn && need_q && (k < Ns - 1);
        double sq0 = 0.0, sq1 = 0.0, sa0 = 0.0, sa1 = 0.0;
        if (act) {
            const double nu = is_syn ? C.T[L.nsyn + k] : C.Ts[L.nic + k];
            const int jm = do_q ? C.TI[L.jm_syn + k] : Ng;
            int jz, jx;
            const double* __restrict__ Et = nullptr;
            if (C.E) {
                Et = C.E + ecache_index(row, 0, Ng);
                jz = C.TI[L.jz + row] & ~1; jx = C.TI[L.jx + row];
            } else {
                int lo = 0, hi = Ng;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (nu * S.inuc[mid] > 745.2) lo = mid + 1; else hi = mid; }
                jz = lo & ~1;
                hi = Ng;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (!(nu * S.inuc[mid] < SYN_XSMALL)) lo = mid + 1; else hi = mid; }
                jx = min((lo + 1) & ~1, Ng & ~1);
            }
            // this lane's gamma pairs (j, j + 1): j = jz + 2 sub, + 2 tpr, ...; first the part that needs the exponential
            // (table or staged evaluation, loads batched), then the part where the four-term series does
            const int st = 2 * tpr;
            int j = jz + 2 * sub;
            if (Et) {
#pragma unroll 4
                for (; j < jx; j += st) {
                    const double E0 = __ldg(Et + (size_t)j * NR), E1 = __ldg(Et + (size_t)(j + 1) * NR);
                    if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                    if (j + 1 >= jm) sq1 = fma(E1, S.u[j + 1], sq1);
                    sa0 = fma(E0, S.v[j], sa0);
                    sa1 = fma(E1, S.v[j + 1], sa1);
                }
            } else {
                for (; j < jx; j += 2 * st) {                             // two pairs = four staged exponentials at once
                    const bool more = j + st < jx;
                    double x[4], Ev[4];
                    x[0] = nu * S.inuc[j]; x[1] = nu * S.inuc[j + 1];
                    x[2] = nu * S.inuc[more ? j + st : j]; x[3] = nu * S.inuc[more ? j + st + 1 : j + 1];
                    fast_exp_neg<4>(x, Ev, S.e2tab);
                    if (j >= jm) sq0 = fma(Ev[0], S.u[j], sq0);
                    if (j + 1 >= jm) sq1 = fma(Ev[1], S.u[j + 1], sq1);
                    sa0 = fma(Ev[0], S.v[j], sa0);
                    sa1 = fma(Ev[1], S.v[j + 1], sa1);
                    if (more) {
                        const int j2 = j + st;
                        if (j2 >= jm) sq0 = fma(Ev[2], S.u[j2], sq0);
                        if (j2 + 1 >= jm) sq1 = fma(Ev[3], S.u[j2 + 1], sq1);
                        sa0 = fma(Ev[2], S.v[j2], sa0);
                        sa1 = fma(Ev[3], S.v[j2 + 1], sa1);
                    }
                }
                if (j >= jx + st) j -= st;                                // the last round took one pair only
            }
            // j is now this lane's first pair at or beyond jx (jx - jz is even, the lane's pairs keep their residue)
#pragma unroll 2
            for (; j + 1 < Ng; j += st) {
                const double E0 = exp_neg_small(nu * S.inuc[j]), E1 = exp_neg_small(nu * S.inuc[j + 1]);
                if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                if (j + 1 >= jm) sq1 = fma(E1, S.u[j + 1], sq1);
                sa0 = fma(E0, S.v[j], sa0);
                sa1 = fma(E1, S.v[j + 1], sa1);
            }
            if (j == Ng - 1) {                                           // odd Ng: the last gamma, whatever its x
                const double x = nu * S.inuc[j];
                double E0;
                if (x < SYN_XSMALL) E0 = exp_neg_small(x);
                else if (Et) E0 = __ldg(Et + (size_t)j * NR);
                else { double xx[1] = {x}, ee[1]; fast_exp_neg<1>(xx, ee, S.e2tab); E0 = ee[0]; }
                if (j >= jm) sq0 = fma(E0, S.u[j], sq0);
                sa0 = fma(E0, S.v[j], sa0);
            }
        }
        double sq = sq0 + sq1, sa = sa0 + sa1;
        for (int d = tpr >> 1; d > 0; d >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, d);
            sa += __shfl_xor_sync(0xffffffffu, sa, d);
        }
        if (act && sub == 0) {
            if (is_syn) {
                if (k < Ns - 1) S.Qsyn[k] = do_q ? qpref * C.T[L.nm23s + k] * sq : 0.0;
                S.aS[k] = need_a ? apref * C.T[L.nm53s + k] * sa : 0.0;
            } else {
                S.aI[k] = apref * C.Ts[L.nm53i + k] * sa;
            }
        }
    }
}

template <int NT>
__device__ void phase_gg_split(const Ctx& C, Smem& S, const double* Lng, double* agg_out, double* qee_out, Part part) {
    const Layout& L = C.L;
    const int NP = L.NP, Ni = L.Ni, Ng = L.Ng, Nt = L.Nt;
    const int nrows = Ni + (Ng - 1);
    const int tpr = part.tpr, grp = threadIdx.x / tpr, sub = threadIdx.x % tpr, ngrp = NT / tpr;
    for (int base = 0; base < nrows; base += part.count * ngrp) {
        const int row = base + part.rank + part.count * grp;
        const bool act = row < nrows;
        const bool is_a = row < Ni;
        const int r = act ? (is_a ? row : row - Ni) : 0;
        double acc[GG_U];
#pragma unroll
        for (int u = 0; u < GG_U; ++u) acc[u] = 0.0;
        if (act) {
            const double a = __ldg(C.Ts + (is_a ? L.agg_a : L.qee_a) + r);
            const double step = __ldg(C.Ts + (is_a ? L.agg_d : L.qee_d) + r);
            const double cf = __ldg(C.Ts + (is_a ? L.agg_c : L.qee_c) + r);
            const double r2 = __ldg(C.Ts + (is_a ? L.agg_r2 : L.qee_r2) + r);
            gg_row_samples(S, Nt, NP, a, step, cf, r2, is_a,
                           [&](int, const int (&j)[GG_U], const double (&t)[GG_U], const double (&Kv)[GG_U]) {
                               gg_batch_acc(S.Lm, S.e2tab, j, t, Kv, acc);
                           }, sub, tpr);
        }
        double s = acc[0];
#pragma unroll
        for (int u = 1; u < GG_U; ++u) s += acc[u];
        for (int d = tpr >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (act && sub == 0) {
            if (is_a) agg_out[r] = s;
            else {
                double ng = 0.0;
                if (2.0 * S.gS[r] > 1.0)
                    ng = exp10(interp_eval(Lng, C.TIs[L.q_j + r], C.Ts[L.q_dx + r], C.Ts[L.q_Dx + r]));
                qee_out[r] = kPC.qee_pref * (ng * s);
            }
        }
    }
}

}  // namespace lhm
