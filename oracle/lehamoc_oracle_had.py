"""CPU oracle, hadronic stage: NumPy restatement of the photo-hadronic functions of
`LeHaMoC Code and Tests/LeHaMoC_f.py`.

TEST INFRASTRUCTURE ONLY (same rules as lehamoc_oracle.py: only tests/, smoke() and bench.py's CPU legs may import
it; the product never falls back to it).

Restated (citations relative to `LeHaMoC Code and Tests/LeHaMoC_f.py`):
  * 185-196  dg_dt_BH   proton loss rate by Bethe-Heitler pair production (Blumenthal 1970, table f(xi).csv)
  * 201-220  dg_dt_pg   proton loss rate by photopion production (Begelman+ 1990; cross_section.csv, kp_pg.txt)
  * 343-356  x_plus / x_minus / x_plus_e / x_minus_e          (Kelner & Aharonian 2008, eqs 21, 40)
  * 358-464  Qp_g_opt   photopion secondaries e+, e-, 2 gamma, nu_mu, anti-nu_mu, nu_e, anti-nu_e (K&A 2008 eq. 30)
  * 466-563  Phi_g      the K&A parametrisation read from Phi_g_K&A.txt, Phi_g_leptons_K&A.txt, Phi_e-nu_e_K&A.txt

Parity pinning: `oracle/make_golden_had.py` imports the UNMODIFIED reference module (build container, astropy/shapely
shim) and stores its outputs for seeded proton and photon fields in tests/golden/had_functions.npz together with the six
tables; tests/test_hadronic.py checks this restatement (and, on the GPU, the CUDA kernels) against those vectors.
The reference's own tests for these functions are notebook overlays against another code (Tests.ipynb cells 5-6,
"within a factor"), i.e. no numeric fixture: the live-reference vectors are the pin.
"""
from __future__ import annotations

import numpy as np

if not hasattr(np, "trapz"):              # numpy >= 2.4
    np.trapz = np.trapezoid               # type: ignore[attr-defined]

from lehamoc_oracle import c, h, m_el, m_pr, sigmaT, eV   # noqa: E402  (same astropy 5.0.4 constants)

H_0 = 0.313          # LeHaMoC_f.py:38
R_KA = 0.1458        # LeHaMoC_f.py:39  m_pi/m_p
E_NU_SPACE = np.logspace(10., 22., 50) * eV          # LeHaMoC_f.py:41

SPECIES = ("2_g", "e+", "e-", "nu_mu", "\bar_nu_mu", "nu_e", "\bar_nu_e")     # flag_product strings of the reference
TABLE_FILES = {"Phi_g": "Phi_g_K&A.txt", "Phi_el": "Phi_g_leptons_K&A.txt", "Phi_el_1": "Phi_e-nu_e_K&A.txt",
               "f_k_i": "f(xi).csv", "cross_section": "cross_section.csv", "kp_pg": "kp_pg.txt"}


def load_tables(dirname):
    """The six tables the reference reads at import (LeHaMoC_f.py:14-19) as plain arrays."""
    import os
    T = {}
    for key, fn in TABLE_FILES.items():
        sep = "," if fn.endswith(".csv") else None
        T[key] = np.loadtxt(os.path.join(dirname, fn), delimiter=sep)
    return T


# ---- A11 ------------------------------------------------------------------------------------------------------
def dg_dt_BH(g_pr, nu, photons, T):
    C_BH = 3. * sigmaT * c * m_el / (8 * np.pi * 137. * m_pr)
    lx = np.log10(h * nu / (m_el * c ** 2))
    with np.errstate(divide="ignore"):
        lp = np.log10(photons)
    fk = T["f_k_i"]
    out = []
    for g_p in g_pr:
        if g_p * h * nu[-1] > 2.1 * m_el * c ** 2.:
            kappa = np.logspace(np.log10(2.), np.log10(g_p * h * nu[-1] / (m_el * c ** 2.)), 50)
            ph = m_el * c ** 2. / h * 10 ** np.interp(np.log10(kappa / (2. * g_p)), lx, lp)
            f_k = np.interp(np.log10(kappa), fk[:, 0], fk[:, 1])
            out.append(np.trapz(10 ** f_k * ph * kappa, np.log10(kappa)))
        else:
            out.append(0.)
    return np.multiply(out, C_BH)


def dg_dt_pg(g_pr, nu, photons, T, E_th=145. * 10 ** 6. * eV):
    C_pion = 5. * 10 ** (-31.) * c
    cs, kp = T["cross_section"], T["kp_pg"]
    lcs_x = np.log10(cs[1:, 0] * 10 ** 9. * eV)
    lkp_x = np.log10(kp[:, 0] * 10 ** 9. * eV)
    lhn = np.log10(h * np.array(nu))
    with np.errstate(divide="ignore"):
        lph = np.log10(np.array(photons / h))
    out = []
    for g_p in g_pr:
        if g_p * h * nu[-1] > E_th:
            eb = np.logspace(np.log10(E_th), np.log10(2. * g_p * h * nu[-1]) + 3, 100)
            CS = np.interp(np.log10(eb), lcs_x, cs[1:, 1])
            KP = np.interp(np.log10(eb), lkp_x, kp[:, 1])
            inner = np.zeros(100)
            on = eb / (2. * g_p) < h * nu[-1] / 2.
            if on.any():
                ep = np.logspace(np.log10(eb[on] / (2. * g_p)), np.log10(h * nu[-1]), 30, axis=-1)     # [n_on, 30]
                dN = 10 ** np.interp(np.log10(ep), lhn, lph)
                inner[on] = np.trapz(dN / ep, np.log(ep), axis=-1)
            out.append(1. / g_p * np.trapz(CS * KP * inner * eb ** 2., np.log(eb)))
        else:
            out.append(0.)
    return np.multiply(out, C_pion)


# ---- A12 ------------------------------------------------------------------------------------------------------
def x_plus(eta, r):
    return 1. / (2. * (1. + eta)) * (eta + r ** 2. + np.sqrt((eta - r ** 2. - 2. * r) * (eta - r ** 2. + 2. * r)))


def x_minus(eta, r):
    return 1. / (2. * (1. + eta)) * (eta + r ** 2. - np.sqrt((eta - r ** 2. - 2. * r) * (eta - r ** 2. + 2. * r)))


def x_plus_e(eta, r):
    return 1. / (2. * (1. + eta)) * (eta - 2 * r + np.sqrt(eta * (eta - 4. * r * (1. + r))))


def x_minus_e(eta, r):
    return 1. / (2. * (1. + eta)) * (eta - 2 * r - np.sqrt(eta * (eta - 4. * r * (1. + r))))


# (table, columns of s/d/B) per species, LeHaMoC_f.py:475-548
_COLS = {"2_g": ("Phi_g", 1, 2, 3), "e+": ("Phi_el", 1, 2, 3), "\bar_nu_mu": ("Phi_el", 4, 5, 6),
         "nu_mu": ("Phi_el", 7, 8, 9), "nu_e": ("Phi_el", 10, 11, 12), "e-": ("Phi_el_1", 1, 2, 3),
         "\bar_nu_e": ("Phi_el_1", 4, 5, 6)}


def Phi_g(eta_space, x, flag, T, h_0=H_0, r=R_KA):
    """Phi(eta, x) for an array of eta and a scalar (or broadcastable) x; same branches as LeHaMoC_f.py:466-563."""
    eta = np.asarray(eta_space, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    hh = eta / h_0
    tab, cs_, cd_, cB_ = _COLS[flag]
    tb = T[tab]
    s = np.interp(hh, tb[:, 0], tb[:, cs_])
    d = np.interp(hh, tb[:, 0], tb[:, cd_])
    B = np.interp(hh, tb[:, 0], tb[:, cB_])
    with np.errstate(all="ignore"):
        if flag == "2_g":
            xp, xm = x_plus(eta, r), x_minus(eta, r)
            psi = 2.5 + 0.4 * np.log(hh)
        elif flag in ("e+", "\bar_nu_mu", "nu_e"):
            xp, xm = x_plus(eta, r), x_minus(eta, r) / 4.
            psi = 2.5 + 1.4 * np.log(hh)
        elif flag == "nu_mu":
            xp = np.where(hh < 2.17, 0.427 * x_plus(eta, r),
                          np.where((hh > 2.17) & (hh < 10), (0.427 + 0.0729 * (hh - 2.14)) * x_plus(eta, r), x_plus(eta, r)))
            xm = 0.427 * x_minus(eta, r)
            psi = 2.5 + 1.4 * np.log(hh)
        else:
            xp, xm = x_plus_e(eta, r), x_minus_e(eta, r) / 2.
            psi = np.where(hh > 4., 6. * (1 - np.exp(1.5 * (4. - hh))), 0.)
        y = (x - xm) / (xp - xm)
        low = B * (np.log(2.)) ** psi
        mid = B * np.exp(-s * np.log(x / xm) ** d) * np.log(2. / (1 + y ** 2.)) ** psi
        out = np.where(x < xm, low, np.where((xp > x) & (x > xm), mid, 0.))
    return out


def Qp_g_opt(g_el, nu_ic, N_pr, g_pr, photons_targ, nu_target, flag, T, h_0=H_0, r=R_KA):
    """LeHaMoC_f.py:358-464; returns one value per element of the species' output grid."""
    thr = 2.14 * h_0 if flag in ("e-", "\bar_nu_e") else h_0
    if flag == "2_g":
        Eout = h * np.asarray(nu_ic)
        g_min_int = h_0 * m_pr * c ** 2. / (4. * h * nu_target[-1])
        pref = h
    elif flag in ("e+", "e-"):
        Eout = np.asarray(g_el) * m_el * c ** 2.
        g_min_int = thr * m_pr * c ** 2. / (4. * g_el[-1] * m_el * c ** 2.)
        pref = m_el * c ** 2.
    else:
        Eout = E_NU_SPACE
        g_min_int = thr * m_pr * c ** 2. / (4. * E_NU_SPACE[-1])
        pref = h
    g_t = np.logspace(np.log10(max(g_min_int, g_pr[0])), np.log10(g_pr[-1]), 50)
    with np.errstate(divide="ignore"):
        N_t = 10 ** np.interp(np.log10(g_t), np.log10(g_pr), np.log10(N_pr))
        lnt, lpt = np.log10(nu_target), np.log10(photons_targ)
    nan_to_num = flag != "2_g"
    out = np.zeros(len(Eout))
    # per proton energy: the eta grid, the interpolated target photons and the integration measure are independent of
    # the outgoing energy
    pre = []
    for g_p in g_t:
        c_t = m_pr * c ** 2. / (4. * g_p)
        eta_max = 4. * h * nu_target[-1] * g_p / (m_pr * c ** 2.)
        if np.log10(thr * c_t / h) < np.log10(nu_target[-1]):
            eta = np.logspace(np.log10(thr), np.log10(eta_max), 25)
            nu_t = eta * c_t / h
            ph = 10 ** np.interp(np.log10(nu_t), lnt, lpt)
            pre.append((eta, nu_t, ph))
        else:
            pre.append(None)
    for k, E in enumerate(Eout):
        inner = np.zeros(50)
        for gi, g_p in enumerate(g_t):
            if pre[gi] is None:
                continue
            eta, nu_t, ph = pre[gi]
            x = E / (g_p * m_pr * c ** 2)
            Phi = Phi_g(eta, x, flag, T, h_0, r)
            if nan_to_num:
                Phi = np.nan_to_num(Phi)
            inner[gi] = np.trapz(ph * Phi * nu_t, np.log(h * nu_t))
        out[k] = np.trapz(pref * N_t * inner / (m_pr * c ** 2.), np.log(g_t))
    return out


# ---- A13 Bethe-Heitler pair injection, LeHaMoC_f.py:566-616 ------------------------------------------------------------
def Q_BH_sol(g_el, g_pr, N_pr, nu_trgt, photons_trgt, intrp_cs_m, intrp_cs_p, max_m, max_p):
    """nu_trgt is the dimensionless photon energy grid h nu/(m_e c^2) (LeHaMoC.py:338).  The spline surfaces come from
    SciPy's bisplrep (host set-up interp_cs_BH_int); evaluation by scipy.interpolate.bisplev, integrals by
    scipy.integrate.simpson of the installed SciPy (its treatment of an even number of samples is part of the pin)."""
    from scipy import integrate, interpolate
    dnde = np.zeros(len(g_el))
    with np.errstate(divide="ignore"):
        lph = np.log10(photons_trgt)
    for i, ge in enumerate(g_el):
        g_pr_int = []
        for j, gp in enumerate(g_pr):
            Emin = (gp + ge) ** 2 / (4. * gp ** 2 * ge)
            Emax = m_pr / (gp * m_el)
            if Emin < Emax:
                E_arr = np.logspace(np.log10(1.001 * Emin), np.log10(Emax), 10)
                f_ph = 10 ** np.interp(E_arr, nu_trgt, lph) * m_el * c ** 2. / (h * E_arr ** 2)      # linear-x interpolation
                temp_ph = []
                Ee_min = (gp ** 2 + ge ** 2) / (2 * gp * ge)
                wmin = (gp + ge) ** 2 / (2 * gp * ge)
                for Eph in E_arr:
                    wmax = 2 * gp * Eph
                    if wmin < wmax:
                        w_arr = np.logspace(np.log10(wmin), np.log10(wmax), 10)
                        tw = []
                        for w in w_arr:
                            Ee_max = w - 1
                            ok = Ee_min < Ee_max and Ee_min < 80. / 100. * Ee_max
                            if ok and gp > ge:
                                zn = interpolate.bisplev(np.log10(Ee_min), np.log10(Ee_max), intrp_cs_p)
                                if zn > max_p:
                                    zn = 0.
                                tw.append(w * 10 ** zn)
                            elif ok and gp < ge:
                                zn = interpolate.bisplev(np.log10(Ee_min), np.log10(Ee_max), intrp_cs_m)
                                if zn > max_m:
                                    zn = 0.
                                tw.append(w * 10 ** zn)
                            else:
                                tw.append(0)
                        temp_ph.append(integrate.simpson(tw * w_arr, x=np.log(w_arr)))
                    else:
                        temp_ph.append(0)
                    temp_ph = list(np.where(np.array(temp_ph) < 0., 0., np.array(temp_ph, dtype=np.float64)))
                g_pr_int.append(N_pr[j] * integrate.simpson(f_ph * temp_ph * E_arr, x=np.log(E_arr)) / (2. * gp ** 3.))
            else:
                g_pr_int.append(0.)
        dnde[i] = integrate.simpson(g_pr_int * g_pr, x=np.log(g_pr)) if len(g_pr_int) else 0.
    return c * dnde
