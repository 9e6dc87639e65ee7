"""GPU twins of the reference's function library LeMoC/LeHaMoC_f.py (same names, arguments and return
shapes), evaluated by the SAME device code the fused timestep kernel uses, through liblhm.so's
function-level entry points.  They exist so that parity tests read like the reference's own notebooks;
production runs go through `Code.LeMoC` / `LeMoC.run` (one launch for the whole time integration).

Scalar helpers (R, B, Volume, Q_el_Lum, ...) are plain host arithmetic, as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .constants import G, c, Ro, Mo, yr, kpc, pc, m_pr, m_el, kb, h, q, sigmaT, eV, B_cr  # noqa: F401
from .engine import Engine, bidiag_solve, tridiag_solve
from .setup_host import Batch, _Lum_e_inj, _Q_el_Lum, _nu_c


# --- host helpers, LeHaMoC_f.py:39-83 ---------------------------------------------------------------
def R(R0, t, t_i, Vexp):
    return R0 + Vexp * (t - t_i)


def B(B0, R0, R, m):
    return B0 * (R0 / R) ** m


def Volume(R):
    return 4. * np.pi / 3. * R ** 3.


Q_el_Lum = _Q_el_Lum
Lum_e_inj = _Lum_e_inj
nu_c = _nu_c


def nuF_nu_obs(nu_L_nu, Dist_in_pc, delta):
    return np.multiply(nu_L_nu, delta ** 4.) / (4. * np.pi * (Dist_in_pc * pc) ** 2.)


# --- ad-hoc single-walker batches ---------------------------------------------------------------------
_R_UNIT_VOLUME = (3. / (4. * np.pi)) ** (1. / 3.)


def _grid(x, fallback, nmin):
    if x is None:
        return fallback
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    while len(x) < nmin:                    # pad with a larger dummy point; callers slice the result
        x = np.append(x, x[-1] * 10.)
    return x


def raw_batch(g=None, nu_syn=None, nu_ic=None, nu_tot=None, R0=1e15, B0=1., ext=None, cor=None, qee_from_ic=0,
              gg_npts=50, loss_minus_one=1):
    g = _grid(g, np.logspace(0., 4., 8), 4)
    nu_syn = _grid(nu_syn, np.logspace(8., 18., 4), 3)
    nu_ic = _grid(nu_ic, np.logspace(10., 30., 4), 3)
    nu_tot = _grid(nu_tot, np.logspace(7.5, 30., 8), 4)
    consts = np.zeros((1, _lib.LHM_NCONST))
    consts[0, _lib.C_R0] = R0
    consts[0, _lib.C_B0] = B0
    consts[0, _lib.C_DT] = R0 / c
    if cor is not None:
        consts[0, _lib.C_COR_P], consts[0, _lib.C_COR_Q0], consts[0, _lib.C_COR_B] = cor
    cfg = dict(Ng=len(g), Ns=len(nu_syn), Ni=len(nu_ic), Nt=len(nu_tot), gg_npts=gg_npts,
               loss_minus_one=loss_minus_one, qee_from_ic=qee_from_ic, const_B=0, ic_pair_table=1, shared_ic_grids=0, shared_had_grids=0, driver=0, Np=0, esc_flag_pr=0, pg_pi_l_flag=0, pg_BH_l_flag=0, pg_pi_emis_flag=0, neutrino_flag=0, pg_BH_emis_flag=0,
               inj_flag=1, Ad_l_flag=0, Syn_l_flag=1, Syn_emis_flag=1, IC_l_flag=1, IC_emis_flag=1, SSA_l_flag=1,
               gg_flag=1, esc_flag=1)
    ext = np.zeros(len(nu_tot)) if ext is None else np.asarray(ext, dtype=np.float64)
    row = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64)[None, :])
    return Batch("raw", cfg, consts, row(g), row(nu_syn), row(nu_ic), row(nu_tot), row(np.full(len(g), 1e-260)),
                 row(ext), np.ones(1, dtype=np.int64))


def _engine(batch, ic_path="R"):
    eng = Engine(batch.cfg, 1, ic_path=ic_path)
    eng.setup(batch)
    return eng


# --- A4  LeHaMoC_f.py:86-92 ---------------------------------------------------------------------------
def Q_syn_space(Np, B, nu, a_cr, g):
    """Synchrotron emissivity at frequency `nu` (scalar or array)."""
    if abs(a_cr / (3. * q * B / (4. * np.pi * m_el * c)) - 1.) > 1e-12:
        raise ValueError("a_cr must equal 3 q B/(4 pi m_e c) (LeMoC.py:201); the kernel derives it from B")
    nus = np.atleast_1d(np.asarray(nu, dtype=np.float64))
    eng = _engine(raw_batch(g=g, nu_syn=np.append(nus, nus[-1] * 10.)))
    Q, _, _ = eng.syn_ssa(np.asarray(Np, dtype=np.float64)[None, :], np.array([float(B)]))
    out = Q[0, :len(nus)].cpu().numpy()
    return float(out[0]) if np.ndim(nu) == 0 else out


# --- A6  LeHaMoC_f.py:101-102 --------------------------------------------------------------------------
def aSSA(N_el, B, nu, g, dg_l_el):
    g = np.asarray(g, dtype=np.float64)
    if abs(dg_l_el / (np.log(g[1]) - np.log(g[0])) - 1.) > 1e-9:
        raise ValueError("dg_l_el must be log(g[1])-log(g[0]) (LeMoC.py:118)")
    nus = np.atleast_1d(np.asarray(nu, dtype=np.float64))
    eng = _engine(raw_batch(g=g, nu_syn=nus))
    _, aS, _ = eng.syn_ssa(np.asarray(N_el, dtype=np.float64)[None, :], np.array([float(B)]))
    out = aS[0, :len(nus)].cpu().numpy()
    return float(out[0]) if np.ndim(nu) == 0 else out


# --- A7  LeHaMoC_f.py:113-124 --------------------------------------------------------------------------
def Q_IC_space_optimized(Np, g_el, nu_ic_temp, photons, nu_targ, index, ic_path="R"):
    nus = np.atleast_1d(np.asarray(nu_ic_temp, dtype=np.float64))
    nu_targ = np.asarray(nu_targ, dtype=np.float64)[:index + 1]        # the kernel integrates over nu_tot[:Nt-1]
    photons = np.asarray(photons, dtype=np.float64)[:index + 1]
    if len(nu_targ) != index + 1:
        nu_targ = np.append(nu_targ, nu_targ[-1] * 10.)                # index == len(nu_targ): keep every point
        photons = np.append(photons, photons[-1])
    eng = _engine(raw_batch(g=g_el, nu_ic=np.append(nus, nus[-1] * 10.), nu_tot=nu_targ), ic_path)
    Q = eng.ic_emissivity(np.asarray(Np, dtype=np.float64)[None, :], photons[None, :], ic_path)
    out = Q[0, :len(nus)].cpu().numpy()
    return float(out[0]) if np.ndim(nu_ic_temp) == 0 else out


# --- A8  LeHaMoC_f.py:127-137 --------------------------------------------------------------------------
def a_gg(nu_ic, nu_target, photons_target):
    nu_ic = np.asarray(nu_ic, dtype=np.float64)
    eng = _engine(raw_batch(nu_ic=nu_ic, nu_tot=nu_target))
    a, _ = eng.gg(np.asarray(photons_target, dtype=np.float64)[None, :])
    return list(a[0, :len(nu_ic)].cpu().numpy())


# --- A9  LeHaMoC_f.py:140-151 --------------------------------------------------------------------------
def Q_ee_f(nu_target, photons_target, nu_ic, photons_IC, g, R0):
    g = np.asarray(g, dtype=np.float64)
    eng = _engine(raw_batch(g=g, nu_ic=nu_ic, nu_tot=nu_target, qee_from_ic=1))
    _, Qee = eng.gg(np.asarray(photons_target, dtype=np.float64)[None, :],
                    np.asarray(photons_IC, dtype=np.float64)[None, :])
    return Qee[0, :len(g) - 1].cpu().numpy()


# --- A1  LeHaMoC_f.py:154-156 --------------------------------------------------------------------------
def photons_tot(nu_syn, nu_bb, photons_syn, nu_ic, photons_IC, nu_tot, photons_bb, photons_pl, photons_user):
    nu_tot = np.asarray(nu_tot, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):        # constant external terms stay on the host
        ext = 10 ** (np.interp(np.log10(nu_tot), np.log10(nu_bb), np.log10(photons_bb))) + photons_pl + photons_user
    eng = _engine(raw_batch(nu_syn=nu_syn, nu_ic=nu_ic, nu_tot=nu_tot, ext=ext))
    n = eng.photons_tot(np.asarray(photons_syn, dtype=np.float64)[None, :],
                        np.asarray(photons_IC, dtype=np.float64)[None, :], np.array([_R_UNIT_VOLUME]))
    return n[0].cpu().numpy()


# --- A2  LeHaMoC_f.py:160-172 --------------------------------------------------------------------------
def U_ph_f(g, nu_target, photons_target, R):
    g = np.asarray(g, dtype=np.float64)
    eng = _engine(raw_batch(g=g, nu_tot=nu_target))
    U = eng.U_ph(np.asarray(photons_target, dtype=np.float64)[None, :], np.array([float(R)]))
    return list(U[0, :len(g)].cpu().numpy())


# --- A5  LeHaMoC_f.py:175-237 --------------------------------------------------------------------------
def cor_factor_syn_el(g_el_space, R0, B0, p_el, Lum_e_injected):
    g = np.asarray(g_el_space, dtype=np.float64)
    Q0 = Q_el_Lum(Lum_e_injected, p_el, g[1], g[-2])
    eng = _engine(raw_batch(g=g, R0=R0, cor=(p_el, Q0, B0)))
    return float(eng.cor_factor()[0].item())


# --- A3  LeHaMoC_f.py:241-264 --------------------------------------------------------------------------
def thomas(a, b, c, d):
    """Every call site of the drivers passes a == 0 (LeMoC.py:239,267,273), i.e. an upper-bidiagonal system; a
    non-zero sub-diagonal takes the general forward-sweep kernel."""
    if np.any(np.asarray(a) != 0.):
        return tridiag_solve(a, b, c, d)[0].cpu().numpy()
    return bidiag_solve(b, c, d)[0].cpu().numpy()


# --- A11-A14: the hadronic functions of `LeHaMoC Code and Tests/LeHaMoC_f.py` live in lehamoc_b200.hadronic ----------
from .hadronic import (dg_dt_BH, dg_dt_pg, Qp_g_opt, interp_cs_BH_int, Q_BH_sol,            # noqa: E402,F401
                       Q_e_pp, Q_g_pp, Q_nu_mu_pp, Q_nu_e_pp)
