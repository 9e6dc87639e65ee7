"""ctypes binding of liblhm.so (include/lhm.h).  There is no CPU fallback: if the library or a B200 is
missing, calls raise."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LHM_LIB", os.path.join(HERE, "liblhm.so"))     # LHM_LIB: experimental builds

LHM_OK = 0
LHM_IC_PATH_R = 0
LHM_IC_PATH_S = 1
LHM_IC_PATH_G = 2
LHM_NCONST = 12
# indices of the per-walker scalar block (include/lhm.h)
C_R0, C_B0, C_M, C_VEXP, C_DT, C_TINIT, C_NSTEPS, C_COR_P, C_COR_Q0, C_COR_B, C_NH, C_PPR = range(12)


class LhmError(RuntimeError):
    pass


class lhm_config(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "Ng", "Ns", "Ni", "Nt",
        "inj_flag", "Ad_l_flag", "Syn_l_flag", "Syn_emis_flag", "IC_l_flag", "IC_emis_flag", "SSA_l_flag",
        "gg_flag", "esc_flag", "loss_minus_one", "qee_from_ic", "gg_npts", "ic_path", "const_B", "ic_pair_table",
        "shared_ic_grids", "driver", "Np", "esc_flag_pr", "pg_pi_l_flag", "pg_BH_l_flag", "pg_pi_emis_flag",
        "neutrino_flag", "pg_BH_emis_flag", "pp_l_flag", "pp_ee_emis_flag", "pp_g_emis_flag", "pp_nu_emis_flag",
        "shared_had_grids", "max_walkers")]


LHM_PACK_NPAR = 20
(PK_R0, PK_B0, PK_GMIN, PK_GMAX, PK_GPLMIN, PK_GPLMAX, PK_P, PK_COMP_INJ, PK_COMP_COR, PK_IDXMIN, PK_IDXMAX,
 PK_GMIN_PR, PK_GMAX_PR, PK_GPLMIN_PR, PK_GPLMAX_PR, PK_IDXMIN_PR, PK_IDXMAX_PR, PK_P_PR, PK_COMP_PR) = range(19)


class lhm_pack_args(C.Structure):                    # include/lhm.h: lhm_pack_args
    _fields_ = ([(n, C.c_int) for n in ("W", "Ng", "Nh", "Nt", "variant")]
                + [(n, C.c_double) for n in ("time_init", "time_end", "step_alg", "Vexp_cm_s", "m", "c", "c3", "m_el",
                                             "mc2", "sigmaT", "four_pi", "nuc_pref")]
                + [(n, C.c_void_p) for n in ("walker_params", "nu_ic_row", "nu_tot_row", "ext_row", "consts", "g_el",
                                             "nu_syn", "nu_ic", "nu_tot", "el_inj", "ext")]
                + [("Np", C.c_int), ("m_pr", C.c_double), ("mpc2", C.c_double), ("g_pr", C.c_void_p),
                   ("pr_inj", C.c_void_p)])


class lhm_fit_obs(C.Structure):                      # include/lhm.h: lhm_fit_obs
    _fields_ = [("nd", C.c_int)] + [(n, C.c_void_p) for n in ("x", "y", "yerr_hi", "yerr_lo", "flags")] + [
        ("log10_1pz", C.c_double), ("distance_norm", C.c_double)]


# every symbol include/lhm.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "lhm_version": (C.c_int, []),
    "lhm_strerror": (C.c_char_p, [C.c_int]),
    "lhm_last_cuda_error": (C.c_char_p, []),
    "lhm_launch_count": (C.c_longlong, []),
    "lhm_create": (C.c_int, [C.POINTER(lhm_config), C.c_int, _P, C.POINTER(_P)]),
    "lhm_destroy": (None, [_P]),
    "lhm_step_smem_bytes": (C.c_int, [_P]),
    "lhm_table_bytes_per_walker": (C.c_longlong, [_P]),
    "lhm_setup_batch": (C.c_int, [_P, C.c_int] + [_P] * 7),
    "lhm_cor_factor_batch": (C.c_int, [_P, _P]),
    "lhm_run_batch": (C.c_int, [_P, _P, _P, C.c_int]),
    "lhm_set_protons": (C.c_int, [_P, _P, _P]),
    "lhm_set_had_tables": (C.c_int, [_P, _P]),
    "lhm_set_bh_surfaces": (C.c_int, [_P, _P, _P]),
    "lhm_run_batch_lh": (C.c_int, [_P, _P, _P, _P, _P, C.c_int]),
    "lhm_last_timings": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "lhm_set_phase_profile": (C.c_int, [_P, _P]),
    "lhm_set_lockstep": (C.c_int, [_P, C.c_int]),
    "lhm_set_cluster": (C.c_int, [_P, C.c_int]),
    "lhm_fit_logl_host": (C.c_int, [_P, C.POINTER(lhm_pack_args), _P, C.POINTER(lhm_fit_obs), _P]),
    "lhm_fit_logl_device": (C.c_int, [_P, C.POINTER(lhm_pack_args), _P, C.POINTER(lhm_fit_obs), _P]),
    "lhm_last_hadtab_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "lhm_last_icg_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "lhm_run_batch_host": (C.c_int, [_P, C.c_int] + [_P] * 9 + [C.c_int]),
    "lhm_photons_tot_batch": (C.c_int, [_P] * 5),
    "lhm_uph_batch": (C.c_int, [_P] * 4),
    "lhm_syn_ssa_batch": (C.c_int, [_P] * 6),
    "lhm_ic_emissivity_batch": (C.c_int, [_P] * 4 + [C.c_int]),
    "lhm_gg_batch": (C.c_int, [_P] * 5),
    "lhm_tridiag_solve_batch": (C.c_int, [C.c_int, C.c_int] + [_P] * 7),
    "lhm_bidiag_solve_batch": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "lhm_loglike_batch": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, C.c_double, C.c_double, C.c_int] + [_P] * 8),
    "lhm_had_loss_batch": (C.c_int, [C.c_int] * 4 + [_P] * 5 + [_P]),
    "lhm_photopion_batch": (C.c_int, [C.c_int] * 7 + [_P] * 9 + [_P]),
    "lhm_bh_emis_batch": (C.c_int, [C.c_int] * 4 + [_P] * 5 + [_P, _P, _P, _P]),
    "lhm_pack_batch": (C.c_int, [_P, _P]),
    "lhm_window_indices_host": (C.c_int, [C.c_int, C.c_int] + [_P] * 7),
    "lhm_pp_batch": (C.c_int, [C.c_int] * 4 + [_P] * 6 + [_P]),
    "lhm_fp64_peak_probe": (C.c_int, [C.c_int, _P, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """dlopen liblhm.so and attach the prototypes.  Raises LhmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LhmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or `python -m lehamoc_b200.build`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != LHM_OK:
        lib = load()
        msg = lib.lhm_strerror(rc).decode()
        if rc == -2:
            msg += ": " + lib.lhm_last_cuda_error().decode()
        raise LhmError(f"{what or 'liblhm'} failed ({rc}): {msg}")
