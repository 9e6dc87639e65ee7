// lhm_setup.cuh -- state-independent tables of one walker (one CTA per walker).
// Everything here is a function of the grids only: trapezoid weights, np.interp
// stencils, the gamma-gamma sample kernels.  Built once per batch, read by every timestep.
#pragma once
#include "lhm_device.cuh"
#include "lhm_step.cuh"
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
    double2* ptab;          // IC pair table [W or 1, ic_maxT*IC_PT*2*32] or null
    int ptab_shared;        // 1: every walker has the same g_el / nu_ic / nu_tot, one table (walker 0 builds it)
    int ptab_path_S;        // 1: the handle runs IC path S -- the same memory holds ITS pair table (ln(gamma - eps) [No, Ng]
                            //    doubles, then the cut index i* [No, Ng] as unsigned short; phase_ic_S_tab)
    int skip_ic_schedule;   // 1: the handle runs IC path G, nothing reads the per-walker tile schedule of paths R / S
    const double* g_pr;     // [W, Np] or null: proton grid / injection of the LeHaMoC.py driver
    const double* pr_inj;   // [W, Np]
};

// log-spaced resample grid of a_gg / Q_ee_f: np.logspace(a, b, NP)[s] = 10**(a + s*step), last = 10**b
__device__ __forceinline__ double logspace_pt(double a, double b, int s, int NP) {
    double step = (b - a) / (double)(NP - 1);
    double y = (s == NP - 1) ? b : (a + (double)s * step);
    return exp10(y);
}

template <int NT>
__device__ void setup_lh_tables(const Layout& L, double* T, int* TI, const double* gp, const double* prinj,
                                const double* ns, const double* nt, double one);

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
        double* E = A.Ecache + (size_t)w * ecache_doubles(Ns, Ni, Ng);
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
            int lo2 = lo, hi2 = Ng;                              // first j with x < SYN_XSMALL, rounded up to an even index
            while (lo2 < hi2) {
                int mid = (lo2 + hi2) >> 1;
                double x = nu * (1.0 / (pc.nuc_pref * B0 * (g[mid] * g[mid])));
                if (!(x < SYN_XSMALL)) lo2 = mid + 1; else hi2 = mid;
            }
            lo2 = max((lo2 + 1) & ~1, lo & ~1);
            TI[L.jx + row] = min(lo2, Ng & ~1);
        }
        // 1/nu_c(gamma, B0) once per gamma (the expression of phase_vectors), 2^(i/64) for the staged exponential
        __shared__ double e2tab_s[64];
        double* inuc = T + L.inuc0;
        for (int j = tid; j < Ng; j += NT) inuc[j] = 1.0 / (pc.nuc_pref * B0 * (g[j] * g[j]));
        for (int i = tid; i < 64; i += NT) e2tab_s[i] = exp2((double)i / 64.0);
        __syncthreads();                                         // jz of every row is visible to the whole CTA
        // Only what phase_syn_rows reads is written: a row's entries below (jz & ~1) are exact zeros it skips (56 % of
        // the Test-1 table), so they cost neither an exponential nor a store.  Four elements per thread and round, so
        // that the four dependency chains of fast_exp_neg interleave; (j, row) advance without a division.
        const int NR = Ns + Ni;
        const int total = NR * Ng;
        int jq[4], rq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int e = tid + u * NT; jq[u] = e / NR; rq[u] = e - jq[u] * NR; }
        const int dj = (4 * NT) / NR, dr = 4 * NT - dj * NR;
        for (int e0 = tid; e0 < total; e0 += 4 * NT) {
            double x[4], Ev[4];
            bool on[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jq[u], row = rq[u];
                const int rr = min(row, NR - 1);
                const int jxr = TI[L.jx + rr];                  // [jz & ~1, jx) and the last gamma of an odd grid are what is read
                on[u] = (e0 + u * NT < total) && j >= (TI[L.jz + rr] & ~1) && (j < jxr || ((Ng & 1) && j == Ng - 1));
                const double nu = (row < Ns) ? ns[row] : ni[min(row - Ns, Ni - 1)];
                x[u] = on[u] ? nu * inuc[min(j, Ng - 1)] : 0.0;
            }
            fast_exp_neg<4>(x, Ev, e2tab_s);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (on[u]) E[ecache_index(rq[u], jq[u], Ng)] = Ev[u];   // x > 745.2 gives exactly 0
                jq[u] += dj; rq[u] += dr;
                if (rq[u] >= NR) { rq[u] -= NR; ++jq[u]; }
            }
        }
    }
    // ---- IC tile schedule (LeHaMoC_f.py:113-124; see phase_ic_R) ----------------------------------------
    // A task is a tile of IC_OG outgoing frequencies x IC_GB gammas.  Everything that decides which tiles matter is a
    // function of the grids only: a pair is active iff gamma > eps_o (the reference's `q_IC[0] > 0.` test,
    // LeHaMoC_f.py:117), and within a pair F <= 0 (clamped to 0 by LeHaMoC_f.py:121) for every target frequency
    // nu_i <= A = nu_o/(4 gamma (gamma-eps_o)), because F > 0 <=> q = A/nu_i < 1.  So per tile the smallest A over
    // its active pairs gives the first target index i0 the tile has to visit (one guard element kept), the largest A
    // the index i1 from which no pair clamps any more, and the tile costs (Ntar - i0) inner iterations.  Tiles are
    // dealt to the warps as contiguous, cost-balanced ranges -- a static schedule, so the summation order (and the
    // result bits) do not depend on timing.
    if (A.skip_ic_schedule) {                             // path G contracts over walkers with its own task list (lhm_icg.cuh):
        if (tid == 0) TI[L.ic_nt] = 0;                    // an empty schedule for whoever loads it (ic_prepare)
        for (int k = tid; k <= IC_NW; k += NT) TI[L.ic_tb + k] = 0;
    } else {
        const int No = Ni - 1, Ntar = Nt - 1;             // outgoing nu_ic[0..Ni-2], targets nu_tot[0..Nt-2]
        const int nog = (No + IC_OG - 1) / IC_OG, ngb = (Ng + IC_GB - 1) / IC_GB;
        const double* xi = T + L.invnt;
        for (int c = tid; c < nog * ngb; c += NT) {
            const int og = c / ngb, gb = c - og * ngb;
            double Amax = -1.0, Amin = 1e308;
            for (int oo = 0; oo < IC_OG; ++oo) {
                const int o = og * IC_OG + oo;
                if (o >= No) break;
                const double eps = T[L.eps + o], nuo = ni[o];
                for (int jj = 0; jj < IC_GB; ++jj) {
                    const int j = gb * IC_GB + jj;
                    if (j >= Ng) break;
                    if (g[j] > eps) {
                        const double rr = 1.0 / (g[j] - eps);
                        const double Av = nuo * rr * T[L.inv4g + j];
                        Amax = fmax(Amax, Av);
                        Amin = fmin(Amin, Av);
                    }
                }
            }
            int i0 = -1, i1 = -1;
            if (Amax >= 0.0) {
                int is = 0;
                while (is < Ntar && Amin * xi[is] >= 1.0) ++is;     // nu_i <= Amin: q >= 1, F <= 0 for EVERY pair
                i0 = max(is - 1, 0);                                // (one guard element kept)
                while (is < Ntar && Amax * xi[is] >= 1.0) ++is;     // first i with q < 1 for every pair of the tile
                i1 = min(is + 1, Ntar);                             // one guard element: beyond, q <= nu_i/nu_{i+1} < 1
            }
            TI[L.ic_cand + 2 * c] = i0;
            TI[L.ic_cand + 2 * c + 1] = i1;
        }
        __syncthreads();
        // compaction of the active tiles and the warp boundaries, by warp 0 (32 candidates per pass, coalesced; the
        // same result as the obvious serial loop, which cost ~0.3 ms of global-memory latency per batch)
        if (tid < 32) {
            const int lane = tid, ncand = nog * ngb;
            // per-tile preamble in units of inner iterations.  The same figure with and without the pair table (whose
            // preamble is ~6x shorter): the schedule -- and with it the summation order and the result bits -- must
            // not depend on how the pair constants are obtained (test_ic_pair_table_modes_are_bit_identical)
            const int FIXED = 2;
            const unsigned lt_mask = (1u << lane) - 1u;
            int nt = 0;
            long long tot = 0;
            for (int base = 0; base < ncand; base += 32) {
                const int c = base + lane;
                const int i0 = (c < ncand) ? TI[L.ic_cand + 2 * c] : -1;
                const int i1 = (c < ncand) ? TI[L.ic_cand + 2 * c + 1] : -1;
                const bool act = i0 >= 0;
                const unsigned m = __ballot_sync(0xffffffffu, act);
                if (act) {
                    const int pos = nt + __popc(m & lt_mask);
                    const int og = c / ngb, gb = c - og * ngb;
                    TI[L.ic_task + pos] = (og << 16) | gb;
                    TI[L.ic_i0 + pos] = i0;
                    TI[L.ic_i1 + pos] = i1;
                }
                long long cst = act ? (long long)((Ntar - i0) + FIXED) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cst += __shfl_xor_sync(0xffffffffu, cst, o);
                tot += cst;
                nt += __popc(m);
            }
            if (lane == 0) { TI[L.ic_nt] = nt; TI[L.ic_tb] = 0; }
            __syncwarp();
            // warp b takes the tiles whose cumulative-cost midpoint lies in [b, b+1) * tot / IC_NW:
            // bnd(t) = min(IC_NW-1, floor(mid2(t) * IC_NW / (2 tot))), tb[b] = first t with bnd(t) >= b
            long long cum = 0;
            int bprev = 0;
            for (int base = 0; base < nt; base += 32) {
                const int t = base + lane;
                const long long cst = (t < nt) ? (long long)((Ntar - TI[L.ic_i0 + t]) + FIXED) : 0;
                long long inc = cst;                            // inclusive prefix of the costs within the pass
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const long long mid2 = 2 * (cum + inc - cst) + cst;
                int bnd = (t < nt) ? (int)min((long long)(IC_NW - 1), (mid2 * IC_NW) / (2 * tot)) : IC_NW - 1;
                int bp = __shfl_up_sync(0xffffffffu, bnd, 1);
                if (lane == 0) bp = bprev;
                if (t < nt) for (int q = bp + 1; q <= bnd; ++q) TI[L.ic_tb + q] = t;
                bprev = __shfl_sync(0xffffffffu, bnd, min(31, nt - 1 - base));
                cum += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) for (int q = (nt > 0 ? bprev : 0) + 1; q <= IC_NW; ++q) TI[L.ic_tb + q] = nt;
        }
        __syncthreads();
        // per-pair constants {G1, a1}, {a2, a3} in tile order: [tile][pair r][half][lane]
        if (A.ptab && A.ptab_path_S) {
            // 10 bytes per pair in the room of path R's 32 (a tile of IC_OG x IC_GB pairs holds IC_PT*2*32 double2)
            static_assert(IC_PT * 2 * 32 * sizeof(double2) >= (size_t)IC_OG * IC_GB * 10, "path-S pair table must fit path R's");
            if (!A.ptab_shared || w == 0) {
                double* lndT = reinterpret_cast<double*>(A.ptab + (A.ptab_shared ? 0 : (size_t)w * L.ic_maxT * (IC_PT * 2 * 32)));
                unsigned short* isT = reinterpret_cast<unsigned short*>(lndT + (size_t)No * Ng);
                const double lnu0 = 0.5 * T[L.l2nt];
                const double inv_dl = (double)(Ntar - 1) / (0.5 * T[L.l2nt + Ntar - 1] - lnu0);
                for (int e = tid; e < No * Ng; e += NT) {
                    const int o = e / Ng, j = e - o * Ng;
                    const double eps = T[L.eps + o], nuo = T[L.nic + o], lnu4 = T[L.lnu4 + o];
                    const double gj = T[L.g + j];
                    double lnd = 0.0;
                    int is = 0;
                    if (gj > eps) {                                     // the expressions of phase_ic_S, to the letter
                        const double d = gj - eps, rr = 1.0 / d, i4 = T[L.inv4g + j];
                        const double Av = nuo * rr * i4;
                        lnd = log(d);
                        const double lnA = lnu4 - T[L.lg + j] - lnd;
                        is = ic_s_cut_index(lnA, Av, lnu0, inv_dl, xi, 1, Ntar);
                    }
                    lndT[e] = lnd;
                    isT[e] = (unsigned short)is;
                }
            }
        } else if (A.ptab && (!A.ptab_shared || w == 0)) {
            double2* P = A.ptab + (A.ptab_shared ? 0 : (size_t)w * L.ic_maxT * (IC_PT * 2 * 32));
            const int nt = TI[L.ic_nt];
            for (int e = tid; e < nt * IC_PT * 32; e += NT) {
                const int t = e / (IC_PT * 32), r = (e >> 5) % IC_PT, lane = e & 31;
                const int tk = TI[L.ic_task + t];
                const int o = (tk >> 16) * IC_OG + (lane >> 2);
                const int j = (tk & 0xffff) * IC_GB + r * IC_GL + (lane & 3);
                const int oc = min(o, No - 1), jc = min(j, Ng - 1);
                double G1, a1, a2, a3;
                ic_pair_constants(o < No && j < Ng, T[L.eps + oc], ni[oc], T[L.lnu4 + oc], g[jc], T[L.inv4g + jc],
                                  T[L.lg + jc], G1, a1, a2, a3);
                P[((size_t)t * IC_PT + r) * 64 + lane] = make_double2(G1, a1);
                P[((size_t)t * IC_PT + r) * 64 + 32 + lane] = make_double2(a2, a3);
            }
        }
    }
    // ---- gamma-gamma sample grids -----------------------------------------------------------------
    // a_gg (LeHaMoC_f.py:130-136) resamples the photon field on x' = np.logspace(log10(1.3/x_IC), log10(max(x)), NP)
    // and integrates 0.652 sigmaT (s^2-1)/s^3 ln(s) x' n(x') over ln x', s = x' x_IC.  With x'_k = 10^(a + k step):
    //   s_k = 1.3 * 10^(k step),  ln s_k = ln 1.3 + k step ln 10,  (s^2-1)/s^3 * x' = (1 - s^-2)/x_IC,
    // and the trapezoid weights in ln x' are step*ln10 (half at the ends), so the whole static factor of sample k is
    //   w_k * (0.652 sigmaT / x_IC) * (1 - s_k^-2) * ln s_k,    s_k^-2 = s_0^-2 * (10^(-2 step))^k
    // which phase_gg rebuilds in registers from (a, step, c, r2); nothing per sample is tabulated.
    // Q_ee_f (LeHaMoC_f.py:143-149) is the same with s = 2 gamma x', s_0 = 1 and prefactor 2.61/(2 gamma).
    const double* lxt = T + L.lxt;
    const double xmax_l = lxt[Nt - 1];                           // log10(max(x)) / log10(h nu_tot[-1]/mc2)
    for (int o = tid; o < Ni; o += NT) {
        const double xIC = T[L.eps + o];
        const double a = log10(1.3 / xIC);
        const double step = (xmax_l - a) / (double)(NP - 1);
        // the reference's guard `x_space[0] < x_space[1]`
        const bool on = logspace_pt(a, xmax_l, 0, NP) < logspace_pt(a, xmax_l, 1, NP);
        T[L.agg_a + o] = a;
        T[L.agg_d + o] = on ? step : 0.0;
        T[L.agg_c + o] = pc.gg_pref / xIC;
        T[L.agg_r2 + o] = exp10(-2.0 * step);
    }
    for (int j = tid; j < Ng; j += NT) {
        const double gj = g[j];
        const double a = log10(1.0 / (2.0 * gj));
        const double step = (xmax_l - a) / (double)(NP - 1);
        T[L.qee_a + j] = a;
        T[L.qee_d + j] = (2.0 * gj > 1.0) ? step : 0.0;         // `if (2.*g_e) > 1.` else 0
        T[L.qee_c + j] = 2.61 / (2.0 * gj);
        T[L.qee_r2 + j] = exp10(-2.0 * step);
    }
    if (L.Np > 0 && A.g_pr) {
        __syncthreads();
        setup_lh_tables<NT>(L, T, TI, A.g_pr + (size_t)w * L.Np, A.pr_inj + (size_t)w * L.Np, ns, nt, one);
    }
}

}  // namespace lhm
