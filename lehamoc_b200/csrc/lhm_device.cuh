// lhm_device.cuh -- layout of the per-walker tables and small device helpers.
// sm_100a only; float64 everywhere (values span 1e-306 .. 1e+60).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace lhm {

// ---- physical constants: astropy 5.0.4 CGS values the reference pins (LeHaMoC_f.py:17-31) -------
struct PhysConst {
    double c, h, q, sigmaT, m_el, mc2;  // mc2 = m_el*c**2.
    double C_syn;       // LeHaMoC_f.py:87
    double ssa_pref;    // q**3./(2.*m_el**2.*c**2.*0.8975)            LeHaMoC_f.py:102
    double bsyn_pref;   // (4./3.)*sigmaT/(8.*pi*m_el*c)               LeMoC.py:223
    double acr_pref;    // 3.*q/(4.*pi*m_el*c)                          LeMoC.py:201
    double nuc_pref;    // 3./(4.*pi)*q/(m_el*c)                        LeHaMoC_f.py:79
    double ic_pref;     // sigmaT*c*0.75                                LeHaMoC_f.py:123
    double gg_pref;     // 0.652*sigmaT                                 LeHaMoC_f.py:135
    double qee_pref;    // c*sigmaT*m_el*c**2./h                        LeHaMoC_f.py:151
    double mc2_h;       // m_el*c**2./h
    double nuT_num;     // 3.*m_el*c**2.                                LeHaMoC_f.py:164
};

// single translation unit (lhm.cu): the constant block lives here
__constant__ PhysConst kPC;

#define LHM_SENT 1e-260   // 10**(-260.)  LeMoC.py:174-183

// ---- IC tiling: a warp task covers IC_OG outgoing frequencies x IC_GB gammas; a thread owns 1 frequency x IC_PT
//      gammas (lane>>2 = frequency, lane&3 = gamma lane) -------------------------------------------------------
constexpr int IC_OG = 8;
constexpr int IC_GL = 4;
constexpr int IC_PT = 4;
constexpr int IC_GB = IC_GL * IC_PT;   // 16 gammas per task
constexpr int IC_NW = 8;               // warps of the walker CTA the static tile schedule is built for

// ---- per-walker table layout (offsets in doubles / ints from the walker's base) ------------------
struct Layout {
    int Ng, Ns, Ni, Nt, NP;
    int Np;                         // proton grid of the LeHaMoC.py driver (0: leptonic drivers)
    // double tables
    int g, lg, wg, wgi, g13, inv4g, sm, sp, am, ap, elinj, u_xT, u_dx, u_Dx, q_dx, q_Dx;   // Ng each
    int nsyn, lsyn, nm23s, nm53s, asm_, asp_;                                              // Ns each
    int nic, lic, nm53i, aim, aip, eps, lnu4;                                              // Ni each
    int nt, lt, lxt, xt, wt, invnt, l2nt, hnu2, ext, ps_dx, ps_Dx, pi_dx, pi_Dx;           // Nt each
    int agg_a, agg_d;               // Ni: first abscissa / step (log10) of the a_gg resample grid
    int qee_a, qee_d;               // Ng: same for Q_ee_f
    int agg_c, agg_r2;              // Ni: 0.652 sigmaT/x_IC and 10**(-2 step): the sample kernel is rebuilt from these
    int qee_c, qee_r2;              // Ng: 2.61/(2 gamma) and 10**(-2 step)
    // LeHaMoC.py driver only (Np > 0)
    int gp, lgp, wgp, g13p, spm, spp, prinj;        // Np: proton grid, ln, trapz weights, g^(1/3), loss stencils, injection
    int gpm, idgp, l10gp;                           // Np: g_pr_mp[k], 1/dg_pr[j], log10(g_pr)
    int ags_a, ags_d, ags_c, ags_r2;                // Ns: a_gg resample grid on nu_syn (LeHaMoC.py:415)
    int agt_a, agt_d, agt_c, agt_r2;                // Nt: a_gg resample grid on nu_tot (LeHaMoC.py:416)
    int uv_d;                                       // Ng: step (log10) of U_ph_f's 50-point resample (LeHaMoC_f.py:177)
    int inuc0;                                      // Ng: 1/nu_c(gamma, B0), set-up scratch of the exponential cache
    int n_doubles;
    // int tables
    int ps_j, pi_j;                 // Nt
    int u_idx, u_j, q_j;            // Ng
    int jm_syn;                     // Ns
    int jz;                         // Ns+Ni: first gamma index whose exp(-nu/nu_c) is non-zero (const-B cache)
    int jx;                         // Ns+Ni: first EVEN gamma index from which nu/nu_c < SYN_XSMALL (series instead of table)
    // IC tile schedule (lhm_setup.cuh: ic_build_schedule)
    int ic_maxT;                    // number of candidate tiles = ceil((Ni-1)/IC_OG) * ceil(Ng/IC_GB)
    int ic_nt;                      // 1: number of active tiles
    int ic_task, ic_i0, ic_i1;      // ic_maxT each: (og << 16) | gb, first visited / first unclamped target index
    int ic_tb;                      // IC_NW+1: tile range of every warp
    int ic_cand;                    // 2*ic_maxT scratch

    int jm_p;                       // Ns: first proton index that may emit at nu_syn[k] (Q_syn_space mask, m = m_pr)
    int uv_on;                      // Ng: 1 if nu_T < nu_target[-1] (LeHaMoC_f.py:176)
    int n_ints;
};

inline Layout make_layout(int Ng, int Ns, int Ni, int Nt, int NP, int Np = 0) {
    Layout L;
    L.Ng = Ng; L.Ns = Ns; L.Ni = Ni; L.Nt = Nt; L.NP = NP; L.Np = Np;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };   // keep 16-byte alignment
    L.g = take(Ng); L.lg = take(Ng); L.wg = take(Ng); L.wgi = take(Ng); L.g13 = take(Ng);
    L.inv4g = take(Ng); L.sm = take(Ng); L.sp = take(Ng); L.am = take(Ng); L.ap = take(Ng);
    L.elinj = take(Ng); L.u_xT = take(Ng); L.u_dx = take(Ng); L.u_Dx = take(Ng); L.q_dx = take(Ng);
    L.q_Dx = take(Ng);
    L.nsyn = take(Ns); L.lsyn = take(Ns); L.nm23s = take(Ns); L.nm53s = take(Ns); L.asm_ = take(Ns);
    L.asp_ = take(Ns);
    L.nic = take(Ni); L.lic = take(Ni); L.nm53i = take(Ni); L.aim = take(Ni); L.aip = take(Ni);
    L.eps = take(Ni); L.lnu4 = take(Ni);
    L.nt = take(Nt); L.lt = take(Nt); L.lxt = take(Nt); L.xt = take(Nt); L.wt = take(Nt);
    L.invnt = take(Nt); L.l2nt = take(Nt); L.hnu2 = take(Nt); L.ext = take(Nt); L.ps_dx = take(Nt);
    L.ps_Dx = take(Nt); L.pi_dx = take(Nt); L.pi_Dx = take(Nt);
    L.agg_a = take(Ni); L.agg_d = take(Ni); L.qee_a = take(Ng); L.qee_d = take(Ng);
    L.agg_c = take(Ni); L.agg_r2 = take(Ni); L.qee_c = take(Ng); L.qee_r2 = take(Ng);
    const int np_ = Np > 0 ? Np : 0, nsl = Np > 0 ? Ns : 0, ntl = Np > 0 ? Nt : 0, ngl = Np > 0 ? Ng : 0;
    L.gp = take(np_); L.lgp = take(np_); L.wgp = take(np_); L.g13p = take(np_); L.spm = take(np_); L.spp = take(np_);
    L.prinj = take(np_); L.gpm = take(np_); L.idgp = take(np_); L.l10gp = take(np_);
    L.ags_a = take(nsl); L.ags_d = take(nsl); L.ags_c = take(nsl); L.ags_r2 = take(nsl);
    L.agt_a = take(ntl); L.agt_d = take(ntl); L.agt_c = take(ntl); L.agt_r2 = take(ntl);
    L.uv_d = take(ngl);
    L.inuc0 = take(Ng);
    L.n_doubles = o;
    o = 0;
    L.ps_j = take(Nt); L.pi_j = take(Nt); L.u_idx = take(Ng); L.u_j = take(Ng); L.q_j = take(Ng);
    L.jm_syn = take(Ns); L.jz = take(Ns + Ni); L.jx = take(Ns + Ni);
    L.ic_maxT = ((Ni - 1 + IC_OG - 1) / IC_OG) * ((Ng + IC_GB - 1) / IC_GB);
    L.ic_nt = take(2); L.ic_task = take(L.ic_maxT); L.ic_i0 = take(L.ic_maxT); L.ic_i1 = take(L.ic_maxT);
    L.ic_tb = take(IC_NW + 1); L.ic_cand = take(2 * L.ic_maxT);
    L.jm_p = take(Np > 0 ? Ns : 0); L.uv_on = take(Np > 0 ? Ng : 0);
    L.n_ints = o;
    return L;
}

// constant-B exponential cache of one walker: blocks of 32 frequency rows, gamma-major inside a block, so that the 32 rows
// a warp of phase_syn_rows owns are ONE contiguous stream of 256-byte lines over its gamma loop (a [gamma][row] layout
// made every warp hop 2 KB per gamma and the eight warps of a CTA interleave their lines in DRAM)
__host__ __device__ inline size_t ecache_doubles(int Ns, int Ni, int Ng) { return (size_t)((Ns + Ni + 31) / 32) * Ng * 32; }
__host__ __device__ inline size_t ecache_index(int row, int j, int Ng) { return ((size_t)(row >> 5) * Ng + j) * 32 + (row & 31); }

// ---- np.interp semantics (NumPy C `arr_interp`) ---------------------------------
// A stencil is (code, dx, Dx): code = (j << 1) | exact.  exact -> result fp[j] (clamped ends or
// x == xp[j]); otherwise slope form with NumPy's NaN fall-backs, which decide what happens to the
// -inf ordinates log10(0) produces.
__device__ __forceinline__ void interp_stencil(double x, const double* __restrict__ xp, int n,
                                               int& code, double& dx, double& Dx) {
    dx = 0.0; Dx = 1.0;
    if (x > xp[n - 1]) { code = ((n - 1) << 1) | 1; return; }
    if (x < xp[0])     { code = 1; return; }
    if (x == xp[n - 1]) { code = ((n - 1) << 1) | 1; return; }
    int lo = 0, hi = n - 1;                      // invariant xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (xp[lo] == x) { code = (lo << 1) | 1; return; }
    code = lo << 1;
    dx = x - xp[lo];
    Dx = xp[lo + 1] - xp[lo];
}

__device__ __forceinline__ double interp_eval(const double* __restrict__ fp, int code, double dx, double Dx) {
    int j = code >> 1;
    double f0 = fp[j];
    if (code & 1) return f0;
    double f1 = fp[j + 1];
    double slope = (f1 - f0) / Dx;
    double r = slope * dx + f0;
    if (isnan(r)) {
        r = slope * (dx - Dx) + f1;
        if (isnan(r) && f0 == f1) r = f0;
    }
    return r;
}

// lerp form used for the (many) gamma-gamma samples: fp[j] + t*(fp[j+1]-fp[j]); a NaN can only come
// from -inf ordinates, for which np.interp yields -inf (see DESIGN.md "np.interp corner cases").
__device__ __forceinline__ double lerp_log(const double* __restrict__ fp, int j, double t) {
    double f0 = fp[j], f1 = fp[j + 1];
    double r = fma(t, f1 - f0, f0);
    return isnan(r) ? -INFINITY : r;
}

// max(F, 0) for finite F without touching the FP64 pipe: clear all bits when the sign bit is set
__device__ __forceinline__ double clamp_neg_to_zero(double F) {
    int hi = __double2hiint(F), lo = __double2loint(F);
    int mask = ~(hi >> 31);
    return __hiloint2double(hi & mask, lo & mask);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// trapezoid weight of node j on abscissae x[0..n-1] (np.trapz(y, x) == sum_j w_j y_j)
__device__ __forceinline__ double trapz_w(const double* __restrict__ x, int j, int n) {
    double a = (j > 0) ? (x[j] - x[j - 1]) : 0.0;
    double b = (j < n - 1) ? (x[j + 1] - x[j]) : 0.0;
    return (a + b) * 0.5;
}

__device__ __forceinline__ double volume_of(double R) {      // LeHaMoC_f.py:47-48
    return 4.0 * M_PI / 3.0 * (R * R * R);
}

}  // namespace lhm
