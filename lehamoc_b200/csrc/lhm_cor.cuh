// lhm_cor.cuh -- A5: cor_factor_syn_el (LeMoC/LeHaMoC_f.py:175-237): cor_kernel (tables, one 128-thread CTA per walker)
// + cor_loop_kernel (the 200 timesteps, one warp per walker).
//
// The reference runs a private synchrotron-only simulation (B = 1e4 G, 200 implicit steps of
// 0.1001 R0/c) and returns L_ph/L_e,inj at the end; its arguments are constants of a model, so it is
// evaluated once per walker instead of once per timestep (bit-identical).
//
// Two exact re-associations make it a short latency-bound job instead of 5.1 M exponentials:
//  (1) the returned luminosity is a linear functional of the photon vector, so only
//      z_j = sum_k omega_k * 4 pi dt V * dQ_k/dN_j is needed (one 98 x Ng pass of exp, spread over the four warps) and
//      each step is  Lambda <- (Lambda + z.N) / (1 + dt c/R0)            (SURVEY.md 7.1b-2);
//  (2) the 200 bidiagonal solves x_j^(k+1) = dp_j(x_j^(k)) - cp_j x_{j+1}^(k+1) form a two-dimensional recurrence whose
//      anti-diagonals are independent: lane l owns a contiguous chunk of the unknowns and works on timestep
//      k = s - (31 - l) in macro-step s, taking the value at the top of its chunk from lane l+1's previous macro-step
//      (one shuffle).  200 + 31 macro-steps of (chunk length) dependent FMAs replace 200 x (chunk pass + 5-level scan +
//      replay); every unknown is computed by the reference's own serial recurrence (thomas(), LeHaMoC_f.py:250-262), no
//      scan re-association.  Lambda is linear in N, so every lane carries the share of its chunk through its own
//      timesteps and the shares are summed once, at the last step where the reference samples the ratio (:231-236).
#pragma once
#include "lhm_device.cuh"
#include "lhm_step.cuh"
#include "../../include/lhm.h"

namespace lhm {


struct CorArgs {
    int W, Ng;
    const double* consts;   // [W, LHM_NCONST]
    const double* g_el;     // [W, Ng]
    double* out;            // [W]
    double* scratch;        // [W, cor_scratch_doubles(Ng)]: what cor_kernel hands to cor_loop_kernel
};

constexpr int COR_NT = 128;             // threads per walker (set-up by all four warps, time loop by warp 0)
constexpr int COR_NQ = COR_NT / 32;     // k-quarters of the z pass
constexpr int COR_NNU = 100;            // private frequency grid, LeHaMoC_f.py:188

// chunk arrays: element i of lane l stored at [i*32 + l] (conflict-free)
__host__ __device__ inline int cor_chunk(int Ng) { return (Ng - 2 + 31) / 32; }
// per-walker hand-over: m, binv, injdt, z, N (chunk arrays, [i*32 + lane]) then lam0[32], El, ends, zb
__host__ __device__ inline size_t cor_scratch_doubles(int Ng) { return (size_t)5 * cor_chunk(Ng) * 32 + 32 + 4; }
constexpr int COR_LOOP_WARPS = 4;       // walkers per CTA of cor_loop_kernel
__host__ __device__ inline size_t cor_smem_bytes(int Ng) {
    const size_t NgP = (size_t)((Ng + 1) & ~1);
    // m, binv, injdt, z, N: 5 chunk arrays; g, lg, inj, zfull: 4 x Ng; zq: NQ x Ng; nu, lnu, cnu, lam0c: 4 x 100; e2tab
    return ((size_t)5 * cor_chunk(Ng) * 32 + (4 + COR_NQ) * NgP + 4 * COR_NNU + 64) * sizeof(double);
}

// the sampling rule of the reference's loop (:204-236) for this lane's next timestep
struct CorClock {
    double t, day, t_end, dt, period;
    __device__ __forceinline__ bool pending() const { return t < t_end; }
    __device__ __forceinline__ bool advance() {            // returns whether the reference samples the ratio at this step
        t += dt;
        if (day < t) { day = day + period; return true; }
        return false;
    }
};

// Wavefront time loop with the lane's chunk of every array in registers (CN = chunk length, compile time).
template <int CN>
__device__ __forceinline__ double cor_time_loop(const double* m_, const double* binv, const double* injdt, const double* z,
                                                double lam_lane, double zb, double bph, CorClock ck, int lane) {
    double rm[CN], rb[CN], ri[CN], rz[CN], rN[CN];
#pragma unroll
    for (int i = 0; i < CN; ++i) {
        const int s = i * 32 + lane;
        rm[i] = m_[s]; rb[i] = binv[s]; ri[i] = injdt[s]; rz[i] = z[s]; rN[i] = 0.0;
    }
    const int delay = 31 - lane;
    const double zn0 = (lane == 0) ? zb : 0.0;
    double xbot = 0.0, lam_s = 0.0;
    for (int s = 0;; ++s) {
        double xin = __shfl_down_sync(0xffffffffu, xbot, 1);        // lane l+1's bottom value of the same timestep
        if (lane == 31) xin = 0.0;
        if (s >= delay && ck.pending()) {
            double dpv[CN];
#pragma unroll
            for (int i = 0; i < CN; ++i) dpv[i] = (rN[i] + ri[i]) * rb[i];        // d'_i of thomas(): independent of xin
            double x = xin, zn = zn0;
#pragma unroll
            for (int i = CN - 1; i >= 0; --i) {
                x = fma(rm[i], x, dpv[i]);                                        // x_i = d'_i - c'_i x_{i+1}
                rN[i] = x;
            }
#pragma unroll
            for (int i = 0; i < CN; ++i) zn = fma(rz[i], rN[i], zn);
            xbot = x;
            lam_lane = (lam_lane + zn) / bph;
            if (ck.advance()) lam_s = lam_lane;
        }
        if (!__any_sync(0xffffffffu, s < delay || ck.pending())) break;
    }
    return warp_sum(lam_s);
}

// the same with the chunk arrays in shared memory (grids beyond 12 * 32 + 2 points)
__device__ __forceinline__ double cor_time_loop_smem(int Cn, const double* m_, const double* binv, const double* injdt,
                                                     const double* z, double* Nn, double lam_lane, double zb, double bph,
                                                     CorClock ck, int lane) {
    const int delay = 31 - lane;
    const double zn0 = (lane == 0) ? zb : 0.0;
    double xbot = 0.0, lam_s = 0.0;
    for (int s = 0;; ++s) {
        double xin = __shfl_down_sync(0xffffffffu, xbot, 1);
        if (lane == 31) xin = 0.0;
        if (s >= delay && ck.pending()) {
            double x = xin, zn = zn0;
            for (int i = Cn - 1; i >= 0; --i) {
                const int e = i * 32 + lane;
                x = fma(m_[e], x, (Nn[e] + injdt[e]) * binv[e]);
                Nn[e] = x;
                zn = fma(z[e], x, zn);
            }
            xbot = x;
            lam_lane = (lam_lane + zn) / bph;
            if (ck.advance()) lam_s = lam_lane;
        }
        if (!__any_sync(0xffffffffu, s < delay || ck.pending())) break;
    }
    return warp_sum(lam_s);
}

__global__ void __launch_bounds__(COR_NT) cor_kernel(CorArgs A) {
    extern __shared__ double smc[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int w = blockIdx.x;
    const int Ng = A.Ng, n = Ng - 2, Cn = cor_chunk(Ng), CS = Cn * 32, NgP = (Ng + 1) & ~1;
    double* m_ = smc;              // -c'_i
    double* binv = m_ + CS;
    double* injdt = binv + CS;
    double* z = injdt + CS;
    double* Nn = z + CS;
    double* g = Nn + CS;
    double* lg = g + NgP;
    double* inj = lg + NgP;        // Q0 g^-p on all bins (:202)
    double* zfull = inj + NgP;
    double* zq = zfull + NgP;      // [COR_NQ][NgP] partial k-sums of the z pass
    double* nu = zq + COR_NQ * NgP;
    double* lnu = nu + COR_NNU;
    double* cnu = lnu + COR_NNU;
    double* lam0c = cnu + COR_NNU; // omega_k x 1e-260 (initial photon vector, :194)
    double* e2tab = lam0c + COR_NNU;
    const PhysConst& pc = kPC;
    const double* cst = A.consts + (size_t)w * LHM_NCONST;
    const double R0 = cst[LHM_C_R0], B0 = cst[LHM_C_COR_B], p_el = cst[LHM_C_COR_P], Q0 = cst[LHM_C_COR_Q0];
    for (int j = tid; j < Ng; j += COR_NT) {
        const double gj = A.g_el[(size_t)w * Ng + j];
        g[j] = gj;
        lg[j] = log(gj);
        inj[j] = Q0 * pow(gj, -p_el);
    }
    for (int i = tid; i < 64; i += COR_NT) e2tab[i] = exp2((double)i / 64.0);
    __syncthreads();

    const double dt = 0.1001 * R0 / pc.c;                    // LeHaMoC_f.py:205
    const double Vol = volume_of(R0);
    const double b_syn = (4.0 / 3.0 * pc.sigmaT / (8.0 * M_PI * pc.m_el * pc.c)) * (B0 * B0);   // :184,212
    const double a_cr = 3.0 * pc.q * B0 / (4.0 * M_PI * pc.m_el * pc.c);                      // :210
    // frequency grid np.logspace(7.5, log10(7 nu_c(g[-1],B0))+1.4, 100)  (:188)
    const double la = 7.5, lb = log10(7.0 * (pc.nuc_pref * B0 * (g[Ng - 1] * g[Ng - 1]))) + 1.4;
    const double lstep = (lb - la) / (double)(COR_NNU - 1);
    for (int k = tid; k < COR_NNU; k += COR_NT) {
        const double v = exp10((k == COR_NNU - 1) ? lb : la + (double)k * lstep);
        nu[k] = v;
        lnu[k] = log(v);
    }
    __syncthreads();
    // omega_k = trapz weight over ln(nu) x h nu^2 (4 pi/3) R0^2 c / V   (:234,236)
    // c_k     = omega_k * 4 pi dt V * C_syn B^(2/3) nu^(-2/3)           (:91,228)
    for (int k = tid; k < COR_NNU; k += COR_NT) {
        const double lm = (k > 0) ? lnu[k - 1] : 0.0, lc = lnu[k], lp = (k < COR_NNU - 1) ? lnu[k + 1] : 0.0;
        const double wk = (((k > 0) ? (lc - lm) : 0.0) + ((k < COR_NNU - 1) ? (lp - lc) : 0.0)) * 0.5;
        const double om = wk * (pc.h * (nu[k] * nu[k])) * 4.0 * M_PI / 3.0 * (R0 * R0) * pc.c / Vol;
        const bool interior = k >= 1 && k <= COR_NNU - 2;    // boundary bins are never updated
        cnu[k] = interior ? om * 4.0 * M_PI * dt * Vol * (pc.C_syn * pow(B0, 2.0 / 3.0) * pow(nu[k], -2.0 / 3.0)) : 0.0;
        lam0c[k] = om * LHM_SENT;                            // photons_syn starts at 1e-260 everywhere (:194)
    }
    // static solver tables of the interior unknowns e <-> grid index e+1
    for (int s = tid; s < CS; s += COR_NT) {
        const int l = s & 31, i = s >> 5;
        const int e = l * Cn + i;
        double mv = 0.0, bi = 1.0, idt = 0.0;
        if (e < n) {
            const int j = e + 1;
            const double mp0 = (g[j] + g[(e == 0) ? (Ng - 1) : (e - 1)]) / 2.0;   // g_el_mp[e]  (:182, wraps at e=0)
            const double mp1 = (g[j + 1] + g[j - 1]) / 2.0;                         // g_el_mp[e+1]
            const double dg = (g[j + 1] - g[j - 1]) / 2.0;                          // dg_el[e]    (:185)
            const double V2 = 1.0 + dt * (pc.c / R0 + b_syn * ((mp0 * mp0) / dg));     // (:213,217)
            const double V3 = -dt * (b_syn * ((mp1 * mp1) / dg));                       // (:214,218)
            bi = 1.0 / V2;
            mv = (e == n - 1) ? 0.0 : -(V3 / V2);            // last row's c never enters (:259)
            idt = inj[j] * dt;                               // el_inj[1:-1]*dt (:219)
        }
        m_[s] = mv; binv[s] = bi; injdt[s] = idt; Nn[s] = 0.0; z[s] = 0.0;
    }
    __syncthreads();
    // z pass: warp q sums its quarter of the interior frequencies for every gamma, four exponentials side by side
    {
        const int per = (COR_NNU - 2 + COR_NQ - 1) / COR_NQ;
        const int k0 = 1 + warp * per, k1 = min(k0 + per, COR_NNU - 1);
        for (int jj = lane; jj < Ng; jj += 32) {
            const double gj = g[jj];
            const double inv = 1.0 / (a_cr * (gj * gj));
            const double thr = pc.mc2 * gj;
            double acc = 0.0;
            for (int k = k0; k < k1; k += 4) {
                double x[4], Ev[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = nu[min(k + u, k1 - 1)] * inv;
                if (x[0] > 745.2) break;                    // x grows with k: exp(-x) == 0 from here on
                if (x[3] < SYN_XSMALL) {                    // the whole batch in the range of the four-term series
#pragma unroll
                    for (int u = 0; u < 4; ++u) Ev[u] = exp_neg_small(x[u]);
                } else
                fast_exp_neg<4>(x, Ev, e2tab);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kk = k + u;
                    if (kk < k1 && !(thr < pc.h * nu[kk])) acc = fma(cnu[kk], Ev[u], acc);   // mask (:89-90)
                }
            }
            zq[warp * NgP + jj] = acc;
        }
    }
    __syncthreads();
    for (int jj = tid; jj < Ng; jj += COR_NT) {
        double acc = zq[jj];
#pragma unroll
        for (int q = 1; q < COR_NQ; ++q) acc += zq[q * NgP + jj];
        const double lm = (jj > 0) ? lg[jj - 1] : 0.0, lc = lg[jj], lp = (jj < Ng - 1) ? lg[jj + 1] : 0.0;
        const double wj = (((jj > 0) ? (lc - lm) : 0.0) + ((jj < Ng - 1) ? (lp - lc) : 0.0)) * 0.5;
        const double zj = wj * pow(g[jj], 1.0 / 3.0) / Vol * acc;
        zfull[jj] = zj;
        if (jj >= 1 && jj <= Ng - 2) { const int e = jj - 1; z[(e % Cn) * 32 + (e / Cn)] = zj; }
        // L_e,inj = trapz(el_inj g^2 m c^2, ln g)  (:235): the summand parked where the z partials were
        zq[jj] = wj * (inj[jj] * (g[jj] * g[jj]) * pc.mc2);
    }
    __syncthreads();
    if (warp != 0) return;

    double El_part = 0.0, lam0_part = 0.0, ends = 0.0;
    for (int j = lane; j < Ng; j += 32) El_part += zq[j];
    for (int k = lane; k < COR_NNU; k += 32) {
        if (k >= 1 && k <= COR_NNU - 2) lam0_part += lam0c[k];
        else ends += lam0c[k];
    }
    const double El = warp_sum(El_part);
    ends = warp_sum(ends);
    const double zb = zfull[0] * 1e-160 + zfull[Ng - 1] * 1e-160;   // N_el[0] = N_el[-1] = 1e-160 (:195)
    // hand-over to cor_loop_kernel (the time loop needs one warp and 160 registers per walker: kept out of this kernel,
    // whose other three warps would hold their registers for nothing while it runs)
    double* sc = A.scratch + (size_t)w * cor_scratch_doubles(Ng);
    for (int s = lane; s < CS; s += 32) {
        sc[s] = m_[s]; sc[CS + s] = binv[s]; sc[2 * CS + s] = injdt[s]; sc[3 * CS + s] = z[s]; sc[4 * CS + s] = 0.0;
    }
    double* tail = sc + 5 * CS;
    tail[lane] = lam0_part;
    if (lane == 0) { tail[32] = El; tail[33] = ends; tail[34] = zb; }
}

// the 200 implicit timesteps (:204-236) as a time-skewed wavefront, one warp per walker
__global__ void __launch_bounds__(COR_LOOP_WARPS * 32) cor_loop_kernel(CorArgs A) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * COR_LOOP_WARPS + warp;
    if (w >= A.W) return;                                    // whole warp exits together
    const int Ng = A.Ng, Cn = cor_chunk(Ng), CS = Cn * 32;
    const PhysConst& pc = kPC;
    const double R0 = A.consts[(size_t)w * LHM_NCONST + LHM_C_R0];
    double* sc = A.scratch + (size_t)w * cor_scratch_doubles(Ng);
    const double* m_ = sc; const double* binv = sc + CS; const double* injdt = sc + 2 * CS; const double* z = sc + 3 * CS;
    double* Nn = sc + 4 * CS;
    const double* tail = sc + 5 * CS;
    const double lam0_part = tail[lane], El = tail[32], ends = tail[33], zb = tail[34];
    const double dt = 0.1001 * R0 / pc.c;                    // LeHaMoC_f.py:205
    const double bph = 1.0 + dt * (pc.c / R0);               // photon equation is diagonal (:226)
    CorClock ck;
    ck.t = 0.0; ck.day = 0.0; ck.t_end = 20.0 * R0 / pc.c; ck.dt = dt; ck.period = 1.0 * R0 / pc.c;
    double Lam;
    switch (Cn) {
#define LHM_COR_CASE(c) case c: Lam = cor_time_loop<c>(m_, binv, injdt, z, lam0_part, zb, bph, ck, lane); break;
        LHM_COR_CASE(1) LHM_COR_CASE(2) LHM_COR_CASE(3) LHM_COR_CASE(4) LHM_COR_CASE(5) LHM_COR_CASE(6)
        LHM_COR_CASE(7) LHM_COR_CASE(8) LHM_COR_CASE(9) LHM_COR_CASE(10) LHM_COR_CASE(11) LHM_COR_CASE(12)
#undef LHM_COR_CASE
        default: Lam = cor_time_loop_smem(Cn, m_, binv, injdt, z, Nn, lam0_part, zb, bph, ck, lane); break;
    }
    if (lane == 0) A.out[w] = (Lam + ends) / El;
}

}  // namespace lhm
