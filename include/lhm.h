/* lhm.h -- C ABI of liblhm.so: B200 (sm_100a) implementation of LeHaMoC's per-timestep kinetic update (leptonic
 * drivers LeMoC.py / Code.py, leptohadronic driver LeHaMoC.py with photopion, Bethe-Heitler and pp channels), batched
 * over independent parameter sets ("walkers").
 *
 * The reference (mariapetro/LeHaMoC) is pure Python and has no FFI; its boundary is the Python
 * surface  `python LeMoC.py <Parameters.txt> <suffix>`  (LeMoC/LeMoC.py:45-58,290-296),
 * `Code.LeMoC(params, fileName) -> (nu_tot, nuLnu)`  (Fit_emcee/Code.py:51,277) and the function
 * library LeHaMoC_f.py.  This header is what a ctypes binding of that surface loads; each entry
 * point cites the reference code it replaces.  `lehamoc_b200/_lib.py` is that binding.
 *
 * Conventions
 *  - all arithmetic is float64; `double*` arguments are DEVICE pointers unless the name ends in
 *    `_host`; 2-D arrays are row-major [W, n];
 *  - the caller owns every input/output buffer; the handle owns only its per-batch table workspace;
 *  - every call is asynchronous on the `stream` given to lhm_create (no hidden synchronisation),
 *    except the *_host convenience calls, which copy and synchronise;
 *  - one handle per GPU, not thread-safe per handle; errors are negative lhm_status codes, never exit().
 */
#ifndef LHM_H
#define LHM_H

#ifdef __cplusplus
extern "C" {
#endif

#define LHM_VERSION 102

typedef enum {
    LHM_OK = 0,
    LHM_ERR_ARG = -1,          /* bad argument / size */
    LHM_ERR_CUDA = -2,         /* CUDA runtime error (see lhm_last_cuda_error) */
    LHM_ERR_NO_DEVICE = -3,    /* no sm_100 device: there is NO CPU fallback */
    LHM_ERR_STATE = -4,        /* call order (e.g. run before setup) */
    LHM_ERR_CAPACITY = -5      /* W larger than the handle's capacity / shared memory exceeded */
} lhm_status;

/* IC emissivity formulation */
#define LHM_IC_PATH_R 0        /* triple sum nu_out x gamma x nu_in, kernel recomputed in registers */
#define LHM_IC_PATH_S 1        /* exact suffix-sum collapse of the nu_in sum */
#define LHM_IC_PATH_G 2        /* shared-grid batches (lhm_config::shared_ic_grids, driver 0): the triple sum as a dense
                                  contraction over walkers on the FP64 tensor cores (mma.sync m8n8k4 f64), the clamped
                                  kernel formed in registers once per 32 walkers; the batch advances in lock-step, two
                                  launches per timestep (lhm_set_lockstep gives the common number of timesteps) */

/* Static description of a batch: grid sizes, process flags, driver variant.
 * Flags are the reference's float flags compared with ==1. / ==0. (LeMoC/LeMoC.py:206-279). */
typedef struct {
    int Ng;                 /* len(g_el)                 = int(grid_g_el)          LeMoC.py:115 */
    int Ns;                 /* len(nu_syn) AFTER the appended point = Ng/2+1        LeMoC.py:131,185 */
    int Ni;                 /* len(nu_ic)                = Ng/2                     LeMoC.py:132 */
    int Nt;                 /* len(nu_tot)               = int(grid_nu)             LeMoC.py:133 */
    int inj_flag, Ad_l_flag, Syn_l_flag, Syn_emis_flag, IC_l_flag, IC_emis_flag, SSA_l_flag,
        gg_flag, esc_flag;
    int loss_minus_one;     /* 1: (gamma^2-1) loss shape LeMoC.py:224-234; 0: gamma^2 Code.py:210-220 */
    int qee_from_ic;        /* 0: Q_ee_f(nu_tot,n,nu_tot,n) LeMoC.py:283; 1: (...,nu_ic,photons_IC/V) Code.py:269 */
    int gg_npts;            /* resample points of a_gg / Q_ee_f: 50 in LeMoC/LeHaMoC_f.py:132,144 */
    int ic_path;            /* LHM_IC_PATH_R, LHM_IC_PATH_S or LHM_IC_PATH_G */
    int const_B;            /* 1: B(t) == B0 for every walker (m == 0 or Vexp == 0, LeMoC.py:200): the synchrotron
                               exponentials are tabulated once per walker instead of every step */
    int ic_pair_table;      /* 1: the four per-(nu_out, gamma) constants of the IC kernel are tabulated at set-up (32 B per
                               pair, read once per timestep) instead of recomputed (a reciprocal and a log per pair) */
    int shared_ic_grids;    /* 1: the caller guarantees that all walkers of every batch have identical g_el, nu_ic and
                               nu_tot rows (fixed-grid parameter sweeps): one pair table for the whole batch.  Checked
                               by the *_host entry points; needs ic_pair_table */
    int driver;             /* 0: LeMoC/LeMoC.py and Fit_emcee/Code.py (leptonic); 1: `LeHaMoC Code and Tests/LeHaMoC.py`
                               (electron + proton equations, proton synchrotron, f_had variants of U_ph_f / a_gg /
                               Q_ee_f with gg_npts = 100, gamma-gamma pre-attenuation of the first two steps);
                               2: Fit_emcee/Code.py::LeHaMoC (Code.py:280-535: the leptonic Code.py step + proton equation with
                               synchrotron cooling + proton synchrotron source; lhm_set_protons / lhm_run_batch_lh as for 1,
                               pr_inj then is a number rate, Code.py:422) */
    int Np;                 /* len(g_pr) = int(grid_g_pr), driver 1 only                   LeHaMoC.py:147 */
    int esc_flag_pr;        /* proton escape flag, driver 1 only                            LeHaMoC.py:331 */
    int pg_pi_l_flag, pg_BH_l_flag;     /* driver 1: photopion / Bethe-Heitler proton loss terms   LeHaMoC.py:311-323 */
    int pg_pi_emis_flag, neutrino_flag; /* driver 1: photopion pairs + photons, neutrinos           LeHaMoC.py:343-352 */
    int pg_BH_emis_flag;                /* driver 1: Bethe-Heitler pair injection (lhm_set_bh_surfaces) LeHaMoC.py:337-340 */
    int pp_l_flag;                      /* driver 1: pp loss term of the proton equation            LeHaMoC.py:322-328 */
    int pp_ee_emis_flag, pp_g_emis_flag, pp_nu_emis_flag; /* driver 1: pp pairs, pi0 photons, neutrinos (n_H, p_pr in
                                           the per-walker scalar block)                               LeHaMoC.py:354-395 */
    int shared_had_grids;   /* driver 1 with photo-hadronic flags: 1 = all walkers of every batch share g_el, g_pr, nu_ic and nu_tot,
                               so the state-independent operators of Q_BH_sol / Qp_g_opt / dg_dt_pg (lhm_hadtab.cuh) are
                               tabulated once per batch instead of once per walker */
    int max_walkers;        /* capacity W of the handle's workspace */
} lhm_config;

/* per-walker scalar block, [W, LHM_NCONST] doubles, host-packed by lehamoc_b200/setup_host.py */
enum {
    LHM_C_R0 = 0,           /* cm                                        LeMoC.py:79  */
    LHM_C_B0,               /* G                                         LeMoC.py:80  */
    LHM_C_M,                /* B = B0 (R0/R)^m                           LeMoC.py:81  */
    LHM_C_VEXP,             /* cm/s (already multiplied by c)            LeMoC.py:78  */
    LHM_C_DT,               /* s, = step_alg*R0/c                        LeMoC.py:108 */
    LHM_C_TINIT,            /* time_init exactly as the driver uses it   LeMoC.py:107 */
    LHM_C_NSTEPS,           /* number of timesteps of this walker        LeMoC.py:196 / Code.py:183 */
    LHM_C_COR_P,            /* p_el passed to cor_factor_syn_el          LeMoC.py:251 */
    LHM_C_COR_Q0,           /* Q_el_Lum(L, p, g[1], g[-2])               LeHaMoC_f.py:202 */
    LHM_C_COR_B,            /* 1e4 G                                     LeMoC.py:251 */
    LHM_C_NH,               /* cold proton density n_H, cm^-3 (pp)       LeHaMoC.py:114 */
    LHM_C_PPR,              /* p_pr as the pp functions receive it        LeHaMoC.py:95  */
    LHM_NCONST = 12
};

typedef struct lhm_handle lhm_handle;

int lhm_version(void);
const char* lhm_strerror(int status);
const char* lhm_last_cuda_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long lhm_launch_count(void);

int lhm_create(const lhm_config* cfg, int device, void* cuda_stream, lhm_handle** out);
void lhm_destroy(lhm_handle*);
/* bytes of dynamic shared memory per CTA the fused step kernel uses for this config */
int lhm_step_smem_bytes(const lhm_handle*);
/* bytes of table workspace (HBM) per walker */
long long lhm_table_bytes_per_walker(const lhm_handle*);

/* Build every state-independent table of W walkers (interpolation stencils, trapezoid weights,
 * gamma-gamma sample kernels ...).  Replaces the per-call grid work of LeMoC.py:113-191 that does not
 * depend on N(gamma), n(nu).  consts [W,LHM_NCONST]; g_el [W,Ng]; nu_syn [W,Ns]; nu_ic [W,Ni];
 * nu_tot [W,Nt]; el_inj [W,Ng] (LeMoC.py:174-175); ext [W,Nt] = BB+PL+user photon densities already
 * interpolated on nu_tot (LeMoC.py:140-169 through LeHaMoC_f.py:154-156). */
int lhm_setup_batch(lhm_handle*, int W, const double* consts, const double* g_el, const double* nu_syn,
                    const double* nu_ic, const double* nu_tot, const double* el_inj, const double* ext);

/* ic_path G only: the number of timesteps every walker of the following batches runs (the lock-step loop is driven by
 * the host: nsteps + 1 launches of the per-walker half-steps around nsteps contractions; LeMoC.py:196 range(int(time_end))). */
int lhm_set_lockstep(lhm_handle*, int nsteps);
/* Ensembles smaller than the machine (the reference's fit: 24 walkers per emcee half-ensemble, Fit_emcee/fit_emcee.py:
 * 225-242): evolve every walker on a thread-block CLUSTER of 2, 4 or 8 CTAs (1 restores the one-CTA kernel).  All CTAs keep
 * the state and run the cheap phases redundantly; IC tiles, gamma-gamma rows and synchrotron / SSA rows are dealt to the
 * CTAs and exchanged through distributed shared memory.  Leptonic drivers, IC path R.  Results equal the one-CTA kernel's up
 * to the association of the IC partial sums (1e-15); the setting, not the batch size, decides the bits, so sharded and
 * unsharded evaluations of the same walkers still agree bit for bit when they use the same setting. */
int lhm_set_cluster(lhm_handle*, int ctas_per_walker);
/* ic_path G only: mean device time of the contraction launches of the last run (CUDA events on the stream), how many
 * were timed, the grid size and the dynamic shared memory per CTA */
int lhm_last_icg_timing(lhm_handle*, float* mean_ms, int* launches, int* ctas, int* smem_bytes);

/* driver 1 with photo-hadronic flags: device time of the last build of the state-independent operator tables of Q_BH_sol /
 * Qp_g_opt / dg_dt_pg (queued by lhm_setup_batch), the number of table sets (1 = shared by the batch) and their size;
 * ms = -1 when the per-step evaluation is in use (tables switched off with LHM_HAD_TABLES=0 or too large) */
int lhm_last_hadtab_timing(lhm_handle*, float* ms, int* table_sets, long long* bytes);

/* A5: cor_factor_syn_el (LeHaMoC_f.py:175-237) for every walker of the current batch -> out[W].
 * One warp per walker; the 200-step inner simulation is advanced with a warp-level affine scan. */
int lhm_cor_factor_batch(lhm_handle*, double* out);

/* A0: the whole time integration (LeMoC.py:196-296 / Code.py:183-277), one CTA per walker, state
 * resident in shared memory for all steps.  snapshots=0: only the state after the last step is
 * written (N_el_out [W,Ng] = N_el/Volume, nuLnu_out [W,Nt]); snapshots=S>0: one row per step for the
 * first S steps ([W,S,Ng], [W,S,Nt]).  Either output may be NULL.  lhm_cor_factor_batch is run
 * internally when Syn_emis_flag is set. */
int lhm_run_batch(lhm_handle*, double* N_el_out, double* nuLnu_out, int snapshots);

/* Driver 1 (LeHaMoC.py): proton grid and injection rate per volume of the next batch, g_pr [W,Np], pr_inj [W,Np]
 * (LeHaMoC.py:147,258); call before lhm_setup_batch.  el_inj of lhm_setup_batch is per volume too (LeHaMoC.py:257). */
int lhm_set_protons(lhm_handle*, const double* g_pr, const double* pr_inj);
/* Driver 1: the whole time integration LeHaMoC.py:266-434; as lhm_run_batch plus N_pr/Volume [W,(S),Np] and the
 * neutrino distribution N_nu [W,(S),50] on E_nu_space (either may be NULL).  With a photo-hadronic flag set the tables
 * must have been given with lhm_set_had_tables (struct declared below). */
struct lhm_had_tables_s;
struct lhm_bh_surface_s;
int lhm_set_had_tables(lhm_handle*, const struct lhm_had_tables_s* tables);
/* pg_BH_emis_flag: the two bisplrep surfaces of interp_cs_BH_int (gamma_p < gamma_e, gamma_p > gamma_e), declared below */
int lhm_set_bh_surfaces(lhm_handle*, const struct lhm_bh_surface_s* surf_m, const struct lhm_bh_surface_s* surf_p);
int lhm_run_batch_lh(lhm_handle*, double* N_el_out, double* nuLnu_out, double* N_pr_out, double* N_nu_out,
                     int snapshots);

/* Device time (ms, CUDA events on the handle's stream) of the most recent set-up, cor-factor and fused-step
 * kernels; -1 if never launched.  Synchronises on those events only.  bench.py's roofline uses run_ms. */
int lhm_last_timings(lhm_handle*, float* setup_ms, float* cor_ms, float* run_ms);

/* Diagnostics: when set (device pointer to 16 zeroed uint64), thread 0 of every walker CTA adds the clock64
 * cycles it spent in each phase of the timestep loop; NULL (default) switches it off. */
int lhm_set_phase_profile(lhm_handle*, unsigned long long* dev_counters16);

/* Host-buffer convenience used by Code.LeMoC / LeMoC_batch: H2D of the packed inputs, setup, run,
 * D2H of the outputs, stream synchronise.  All pointers are HOST pointers (pinned or pageable). */
int lhm_run_batch_host(lhm_handle*, int W, const double* consts_host, const double* g_el_host,
                       const double* nu_syn_host, const double* nu_ic_host, const double* nu_tot_host,
                       const double* el_inj_host, const double* ext_host,
                       double* N_el_out_host, double* nuLnu_out_host, int snapshots);

/* ---- function-level entry points (unit-test twins of LeHaMoC_f.py; same device code as the fused
 *      kernel).  All act on the walkers of the current batch (after lhm_setup_batch). ---- */
/* A1 photons_tot LeHaMoC_f.py:154-156 incl. the /Volume of LeMoC.py:204: n[W,Nt] */
int lhm_photons_tot_batch(lhm_handle*, const double* photons_syn, const double* photons_IC,
                          const double* radius /*[W]*/, double* n_out);
/* A2 U_ph_f LeHaMoC_f.py:160-172: U_ph[W,Ng] */
int lhm_uph_batch(lhm_handle*, const double* n, const double* radius, double* U_out);
/* A4+A6 Q_syn_space (raw, before the cor-factor division) LeHaMoC_f.py:86-92 -> Q[W,Ns-1];
 * aSSA LeHaMoC_f.py:101-102 on nu_syn -> aS[W,Ns] and on nu_ic -> aI[W,Ni] (signed, as the function returns) */
int lhm_syn_ssa_batch(lhm_handle*, const double* n_e /*[W,Ng] = N_el/V*/, const double* Bfield /*[W]*/,
                      double* Q_syn_out, double* aS_out, double* aI_out);
/* A7 Q_IC_space_optimized LeHaMoC_f.py:113-124 for nu_ic[0..Ni-2]: Q[W,Ni-1]; ic_path overrides cfg */
int lhm_ic_emissivity_batch(lhm_handle*, const double* n_e, const double* n, double* Q_out, int ic_path);
/* A8 a_gg LeHaMoC_f.py:127-137 -> a_gg[W,Ni];  A9 Q_ee_f LeHaMoC_f.py:140-151 -> Q_ee[W,Ng-1].
 * photons_IC_dens [W,Ni] is read only when cfg.qee_from_ic (Code.py:269); may be NULL otherwise. */
int lhm_gg_batch(lhm_handle*, const double* n, const double* photons_IC_dens, double* a_gg_out,
                 double* Q_ee_out);
/* A3 thomas() with a == 0 (every call site: LeMoC.py:239,267,273): x = upper-bidiagonal solve, [W,n] */
int lhm_bidiag_solve_batch(int W, int n, const double* b, const double* c, const double* d, double* x,
                           void* cuda_stream);
/* thomas(a,b,c,d) with a non-zero sub-diagonal (LeHaMoC_f.py:241-264), [W,n] systems; work2: caller-owned scratch of
 * 2*W*n doubles.  Unit-test twin: every call site of the drivers has a == 0 (lhm_bidiag_solve_batch). */
int lhm_tridiag_solve_batch(int W, int n, const double* a, const double* b, const double* c, const double* d,
                            double* work2, double* x, void* cuda_stream);

/* A10 fit epilogue for W walkers, one warp per walker (handle-free, like lhm_bidiag_solve_batch):
 *   Model.__call__ (Fit_emcee/fit_emcee.py:103-115): xp = log10(nu_tot) + doppler - log10(1+z),
 *     yp = log10(nuLnu) - distance_norm + 4*doppler, y_model = np.interp(x, xp, yp);
 *   log_likelihood_LeMoC (fit_emcee.py:121-145): Gaussian chi^2 on the detections (flags == 0) with jitter
 *     exp(2 log_f), the lower error bar on the last two detections when the model lies below them, and the
 *     erf upper-limit term (flags == 1, rms = 0.05).
 * nu_tot, nuLnu [W,Nt]; doppler, log_f [W]; x, y, yerr_hi, yerr_lo, flags [nd] (flags as the reference's floats);
 * logl_out [W]; ymodel_out [W,nd] or NULL. */
int lhm_loglike_batch(int W, int Nt, const double* nu_tot, const double* nuLnu, const double* doppler,
                      const double* log_f, double log10_1pz, double distance_norm, int nd, const double* x,
                      const double* y, const double* yerr_hi, const double* yerr_lo, const double* flags,
                      double* logl_out, double* ymodel_out, void* cuda_stream);

/* ---- hadronic stage, function-level entry points (handle-free) -------------------------
 * The six tables the reference reads at import (`LeHaMoC Code and Tests/LeHaMoC_f.py`:14-19), as DEVICE pointers to
 * row-major copies of the files. */
typedef struct lhm_had_tables_s {
    const double* phi_g;   int n_phi_g;     /* Phi_g_K&A.txt          [n,4]  eta/eta0, s, d, B                       */
    const double* phi_el;  int n_phi_el;    /* Phi_g_leptons_K&A.txt  [n,13] eta/eta0, (s,d,B) x e+, anti-nu_mu, nu_mu, nu_e */
    const double* phi_el1; int n_phi_el1;   /* Phi_e-nu_e_K&A.txt     [n,7]  eta/eta0, (s,d,B) x e-, anti-nu_e       */
    const double* fk;      int n_fk;        /* f(xi).csv              [n,2]  log10 kappa, log10 f                    */
    const double* cs;      int n_cs;        /* cross_section.csv      [n,2]  photon energy (GeV), sigma (microbarn)  */
    const double* kp;      int n_kp;        /* kp_pg.txt              [n,2]  photon energy (GeV), inelasticity       */
} lhm_had_tables;

/* A11 proton loss rates for W walkers at Ne proton Lorentz factors each: which = 0 dg_dt_pg (LeHaMoC_f.py:201-220,
 * photopion), which = 1 dg_dt_BH (LeHaMoC_f.py:185-196, Bethe-Heitler).  g_eval [W,Ne]; nu, photons [W,Nt];
 * out [W,Ne]. */
int lhm_had_loss_batch(int W, int Ne, int Nt, int which, const double* g_eval, const double* nu, const double* photons,
                       const lhm_had_tables* tables, double* out, void* cuda_stream);

/* A12 photopion secondaries Qp_g_opt (LeHaMoC_f.py:358-464, Phi_g :466-563).  species: 0 "2_g" (out [W,Ni] on h nu_ic),
 * 1 "e+", 2 "e-" (out [W,Ng] on g_el m_e c^2), 3 "nu_mu", 4 anti-nu_mu, 5 "nu_e", 6 anti-nu_e (out [W,Nnu] on E_nu).
 * g_el [W,Ng]; nu_ic [W,Ni]; E_nu [Nnu] (erg, shared); N_pr, g_pr [W,Np]; photons, nu_target [W,Nt]. */
int lhm_photopion_batch(int W, int species, int Ng, int Ni, int Np, int Nt, int Nnu, const double* g_el,
                        const double* nu_ic, const double* E_nu, const double* N_pr, const double* g_pr,
                        const double* photons, const double* nu_target, const lhm_had_tables* tables, double* out,
                        void* cuda_stream);

/* A13 Bethe-Heitler pair injection Q_BH_sol (LeHaMoC_f.py:566-616).  The two smoothing-spline surfaces (tck of
 * scipy.interpolate.bisplrep, kx = ky = 3; host set-up interp_cs_BH_int, LeHaMoC_f.py:276-341) are given as DEVICE
 * pointers: knots tx [nx], ty [ny], coefficients c [(nx-4)*(ny-4)]; zmax = max_intrp_cs.  g_el [W,Ng]; g_pr, N_pr [W,Np]
 * (dN_pr/dV dgamma); x_target [W,Nt] = h nu/(m_e c^2); photons [W,Nt]; out [W,Ng] (the driver uses [1:-1]). */
typedef struct lhm_bh_surface_s {
    const double *tx, *ty, *c;
    int nx, ny;
    double zmax;
} lhm_bh_surface;
int lhm_bh_emis_batch(int W, int Ng, int Np, int Nt, const double* g_el, const double* g_pr, const double* N_pr,
                      const double* x_target, const double* photons, const lhm_bh_surface* surf_m,
                      const lhm_bh_surface* surf_p, double* out, void* cuda_stream);

/* Set-up of the leptonic drivers on the device: what LeMoC.py:107-191 / Code.py:51-180 build with NumPy per model --
 * g_el, nu_syn (+ appended point), el_inj, the per-walker scalar block -- from one row of parameters per walker.
 * Same expressions in the same order; 10**y, log10 and g**-p come from the CUDA math library, so g_el agrees with the
 * NumPy grid to <= 2 ulp and nu_syn / el_inj to ~1e-14 relative (one ulp of a log10 that ends up in an exponent); the
 * host packer in lehamoc_b200/setup_host.py stays the bit-for-bit statement of the reference.
 * The outputs are the seven arrays lhm_setup_batch takes (device pointers, [W, n] row-major). */
enum {
    LHM_PK_R0 = 0,          /* cm (10**R0 evaluated by the host like the reference, LeMoC.py:79 / Code.py:53) */
    LHM_PK_B0,              /* G */
    LHM_PK_GMIN, LHM_PK_GMAX,       /* log10 limits of the gamma grid                       LeMoC.py:115 / Code.py:77 */
    LHM_PK_GPLMIN, LHM_PK_GPLMAX,   /* log10 limits of the injection power law              LeMoC.py:120-128 */
    LHM_PK_P,               /* p_el */
    LHM_PK_COMP_INJ,        /* compactness used for the injection                           LeMoC.py:110,174 / Code.py:163 */
    LHM_PK_COMP_COR,        /* value handed to cor_factor_syn_el (Code.py:237 passes the log)                  */
    LHM_PK_IDXMIN, LHM_PK_IDXMAX,   /* index_PL_min, index_PL_max (-1: g_PL_max == g_max_el) as the reference's own
                                       comparison `g_el > 10**g_PL_max` on the NumPy grid gives them: integers computed
                                       by the host (LeMoC.py:120-128), so that a grid point that coincides with a
                                       threshold cannot move the window by a bin                                  */
    /* protons of Code.py::LeHaMoC (Code.py:299-300,325-347,419-422); used when lhm_pack_args.g_pr != NULL */
    LHM_PK_GMIN_PR, LHM_PK_GMAX_PR,         /* log10 limits of the proton grid                                  */
    LHM_PK_GPLMIN_PR, LHM_PK_GPLMAX_PR,     /* log10 limits of the proton power law                             */
    LHM_PK_IDXMIN_PR, LHM_PK_IDXMAX_PR,     /* window indices on the proton grid, host-computed like the two above */
    LHM_PK_P_PR, LHM_PK_COMP_PR,            /* p_pr, proton compactness 10**P9                                  */
    LHM_PACK_NPAR = 20
};
typedef struct lhm_pack_args {
    int W, Ng, Nh, Nt;      /* Nh = int(grid_g_el/2): len(nu_ic); len(nu_syn) = Nh + 1 */
    int variant;            /* 0: LeMoC.py (nsteps = int(time_end)), 1: Code.py (while time_real < time_end R0/c) */
    double time_init, time_end, step_alg, Vexp_cm_s, m;
    double c, c3, m_el, mc2, sigmaT, four_pi, nuc_pref;   /* c, c**3., m_el, m_el*c**2., sigmaT, 4.*pi, 3./(4.*pi)*q as
                                                             Python evaluates them */
    const double* walker_params;                          /* device [W, LHM_PACK_NPAR] */
    const double *nu_ic_row, *nu_tot_row, *ext_row;       /* device [Nh], [Nt], [Nt]: the same for every walker    */
    double *consts, *g_el, *nu_syn, *nu_ic, *nu_tot, *el_inj, *ext;   /* device outputs */
    int Np;                                 /* len(g_pr) when the proton outputs are wanted */
    double m_pr, mpc2;                      /* m_pr, m_pr*c**2. */
    double *g_pr, *pr_inj;                  /* device outputs [W, Np] for lhm_set_protons, or NULL */
} lhm_pack_args;
int lhm_pack_batch(const lhm_pack_args* args, void* cuda_stream);

/* Host-side helper for the two integer entries of a walker_params row: index_PL_min and index_PL_max of
 * LeMoC.py:120-128 / Code.py:82-90 (`g_el[g_el > 10**g_PL_max][0]`, `g_el[g_el < 10**g_PL_min][-1]` on the np.logspace
 * grid of n points between 10**g_lo and 10**g_hi), for W walkers, without building the grids.  No device work.
 * idx_min / idx_max [W] receive the integers as doubles (the layout of walker_params; -1 = g_PL_max == g_max_el);
 * status [W]: 0 decided; 1 a grid value lies within a few ulp of a threshold, where NumPy's array pow (the reference's
 * grid) and libm's may disagree -- the caller decides that walker with the reference's own NumPy expressions
 * (lehamoc_b200/setup_host.py does); 2 the window lies outside the grid (the reference raises IndexError there). */
int lhm_window_indices_host(int W, int n, const double* g_lo, const double* g_hi, const double* pl_min, const double* pl_max,
                            double* idx_min, double* idx_max, int* status);

/* ---- one call per fit evaluation (replaces `pool.map(log_probability, thetas)` of Fit_emcee/fit_emcee.py:190-199,239-242
 * for flag_c = 'LeMoC', without the prior, which stays on the host): host in, host out.
 * obs: the observation table of fit_emcee.py:61-93 as DEVICE pointers [nd] (flags as the reference's floats) and the two
 * scalars of Model.__call__ (fit_emcee.py:108-115).
 * pack: as for lhm_pack_batch, variant 1 (Code.py); its POINTER members are ignored -- the library packs into buffers the
 *       handle owns.
 * stage_host: [W*LHM_PACK_NPAR | Nh | Nt | Nt | W | W] doubles in HOST memory = walker_params, nu_ic_row, nu_tot_row,
 *       ext_row, log10 of the Doppler factor per walker, log_f per walker.
 * logl_host [W]: log-likelihood of every walker (fit_emcee.py:121-145).
 * Inside: pinned staging + one H2D copy, pack_kernel, the table set-up next to the cor-factor, the fused time loop, the
 * observation transform + likelihood, one D2H copy, one stream synchronisation.  Leptonic driver 0 handles only. */
typedef struct lhm_fit_obs {
    int nd;
    const double *x, *y, *yerr_hi, *yerr_lo, *flags;
    double log10_1pz, distance_norm;
} lhm_fit_obs;
int lhm_fit_logl_host(lhm_handle*, const lhm_pack_args* pack, const double* stage_host, const lhm_fit_obs* obs,
                      double* logl_host);
/* Same evaluation, result left on the device: logl_dev [W] (device memory) is written on the handle's stream and the call
 * returns without waiting -- for a caller that goes on with device work on that stream, e.g. the all-gather of the
 * per-rank blocks of a sharded ensemble (lehamoc_b200/fit.py::log_probability_sharded, fit_emcee.py:239-242).  stage_host may
 * be reused as soon as the call returns (it is copied into the handle's pinned buffer first). */
int lhm_fit_logl_device(lhm_handle*, const lhm_pack_args* pack, const double* stage_host, const lhm_fit_obs* obs,
                        double* logl_dev);

/* A14 pp channels (LeHaMoC_f.py:860-1014): product 0 Q_e_pp (grid = g_el), 1 Q_g_pp (grid = nu_ic), 2 Q_nu_mu_pp,
 * 3 Q_nu_e_pp (grid = nu_nu = E_nu/h).  grid [W,Nout]; g_pr, N_pr_TeV [W,Np] (dN_pr/dV dgamma divided by
 * m_p c^2 * 0.624151, LeHaMoC.py:355); p_pr, n_H [W]; out [W,Nout]. */
int lhm_pp_batch(int W, int product, int Nout, int Np, const double* grid, const double* g_pr, const double* N_pr_TeV,
                 const double* p_pr, const double* n_H, double* out, void* cuda_stream);

/* measured FP64 FMA throughput of this device (TFLOP/s, FMA = 2 flop): the roofline denominator */
int lhm_fp64_peak_probe(int device, void* cuda_stream, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* LHM_H */
