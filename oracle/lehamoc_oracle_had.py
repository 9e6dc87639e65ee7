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


# ---- A14 pp channels, LeHaMoC_f.py:681-1014 (Kelner, Aharonian & Bugayov 2006 parametrisations) ------------------------
E_TH_PI = 1.22 * 10 ** (-3.)          # TeV, LeHaMoC_f.py:40
X_M_PR = 938.272046 * 10 ** (-6.)     # rest masses in TeV, LeHaMoC_f.py:44-48
X_M_EL = 0.511 * 10 ** (-6.)
X_M_PI = 139.57 * 10 ** (-6.)
X_M_PI0 = 134.9766 * 10 ** (-6.)
X_M_MU = 105.66 * 10 ** (-6.)
K_PI = 0.17
TEV = 0.624151                        # erg -> TeV as the reference writes it


def cs_pp_inel(E_p):                  # LeHaMoC_f.py:687-692
    E_p = np.asarray(E_p, dtype=np.float64)
    out = np.zeros(len(E_p))
    m = E_p > E_TH_PI
    L = np.log(E_p[m])
    out[m] = 10 ** (-27.) * (34.3 + 1.88 * L + 0.25 * L ** 2.) * (1. - (np.divide(E_TH_PI, E_p[m])) ** 4.) ** 2
    return out


def _eta_pp(p_pr, species):           # LeHaMoC_f.py:695-714, incl. the np.interp call with swapped arguments
    tab = {2.: 1.1, 2.5: 0.86, 3.: 0.91} if species == "g" else {2.: 0.77, 2.5: 0.62, 3.: 0.67}
    if p_pr in tab:
        return tab[p_pr]
    return np.interp(p_pr, [1.1, 0.86, 0.91], [2., 2.5, 3.])


def q_pi(E_pi, p_pr, species, N_pr_TeV, n_H, g_pr):          # LeHaMoC_f.py:694-716
    eta = _eta_pp(p_pr, species)
    Ep = X_M_PR + E_pi / K_PI
    return eta * c * n_H / K_PI * cs_pp_inel(Ep) * 10. ** (np.interp(np.log10(Ep), np.log10(g_pr * X_M_PR), np.log10(N_pr_TeV)))


def _g_nu_mu(x, r): return (3. - 2. * r) / (9. * (1. - r) ** 2.) * (9. * x ** 2. - 6. * np.log(x) - 4. * x ** 3. - 5.)
def _h_nu_mu_1(x, r): return (3. - 2. * r) / (9. * (1. - r) ** 2.) * (9. * r ** 2. - 6. * np.log(r) - 4. * r ** 3. - 5.)
def _h_nu_mu_2(x, r): return (1. + 2. * r) * (r - x) / (9. * r ** 2.) * (9. * (r + x) - 4. * (r ** 2. + r * x + x ** 2.))
def _g_nu_e(x, r): return 2. * (1. - x) / (3. * (1. - r) ** 2.) * (6. * (1. - x) ** 2. + r * (5. + 5. * x - 4. * x ** 2.) + 6. * r * np.log(x))
def _h_nu_e_1(x, r): return 2. / (3. * (1. - r) ** 2.) * ((1. - r) * (6. - 7. * r + 11. * r ** 2. - 4. * r ** 3.) + 6. * r * np.log(r))
def _h_nu_e_2(x, r): return 2. * (r - x) / (3. * r ** 2.) * (7. * r ** 2. - 4. * r ** 3. + 7. * x * r ** 2. - 2. * x ** 2. - 4. * x ** 2. * r)


def Q_pp_sub2(x_space, species, E_p, E, p_pr, N_pr_TeV, n_H, g_pr):      # LeHaMoC_f.py:735-778
    x = np.asarray(x_space, dtype=np.float64)
    L = np.log(E_p)
    src = c * n_H * cs_pp_inel(E / x) * 10. ** np.interp(np.log10(E / x), np.log10(g_pr * X_M_PR), np.log10(N_pr_TeV))
    if species in ("g", "nu_mu_1"):
        if species == "g":
            y = x
            Bg, bg, kg = 1.3 + 0.14 * L + 0.011 * L ** 2., 1. / (1.79 + 0.11 * L + 0.008 * L ** 2.), 1. / (0.801 + 0.049 * L + 0.014 * L ** 2.)
        else:
            y = np.divide(x, 0.427)
            Bg, bg, kg = 1.75 + 0.204 * L + 0.01 * L ** 2., 1. / (1.67 + 0.111 * L + 0.0038 * L ** 2.), 1.07 - 0.086 * L + 0.002 * L ** 2.
        yb = y ** bg
        F = Bg * np.log(y) / y * ((1. - yb) / (1. + kg * yb * (1. - yb))) ** 4. * (
            1. / np.log(y) - 4. * bg * yb / (1. - yb) - (4. * kg * bg * yb * (1. - 2. * yb)) / (1. + kg * yb * (1. - yb)))
    elif species in ("nu_mu_2", "e", "nu_e"):
        Be = 1. / (69.5 + 2.65 * L + 0.3 * L ** 2.)
        be = 1. / (0.201 + 0.062 * L + 0.00042 * L ** 2.) ** (1. / 4.)
        ke = (0.279 + 0.141 * L + 0.0172 * L ** 2.) / (0.3 + (2.3 + L) ** 2.)
        F = Be * (1. + ke * (np.log(x)) ** 2.) ** 3. / (x * (1. + 0.3 / x ** be)) * (-np.log(x)) ** 5.
    else:
        raise ValueError("Unkown particles species")
    return np.trapz(src * F, np.log(x))


def Q_pp_sub1(x, species, E_p, E, p_pr, N_pr_TeV, n_H, g_pr):            # LeHaMoC_f.py:780-858
    x = np.asarray(x, dtype=np.float64)
    Es = np.multiply(x, E_p)
    r = (X_M_MU / X_M_PI) ** 2.
    shape = None
    if species == "g":
        E_min = Es + X_M_PI0 ** 2. / (4. * Es)
        E_pi = np.logspace(np.log10(min(E_min)), 3., 40)
        pre = 2.
    elif species == "nu_mu_1":
        E_max = max(x) * (max(E_p) - X_M_PR)
        E_min = max(max(Es / (1. - (X_M_MU / X_M_PI) ** 2.)), max(Es + X_M_PI ** 2. / (4. * Es)))
        E_pi = np.logspace(np.log10(E_min), np.log10(E_max), 40)
        pre = 2. / 0.427
    elif species in ("nu_mu_2", "e", "nu_e"):
        mass = X_M_MU if species == "nu_mu_2" else X_M_EL
        E_min = Es + mass ** 2. / (4. * Es)
        E_pi = np.logspace(np.log10(max(E_min)), np.log10((100. if species == "nu_e" else 1000.) * max(E_min)), 40)
        top = min(np.unique(Es)) if species == "e" else max(Es)
        x_new = sorted(top / E_pi)
        shape = []
        for xe in x_new:
            if species == "nu_e":
                shape.append(_g_nu_e(xe, r) if xe > r else _h_nu_e_1(xe, r) + _h_nu_e_2(xe, r))
            elif species == "nu_mu_2":
                shape.append(_g_nu_mu(xe, r) if xe > r else _h_nu_mu_1(xe, r) + _h_nu_mu_2(xe, r))
            else:                                            # "e": negative pieces are dropped (LeHaMoC_f.py:826-841)
                if xe > r:
                    shape.append(0. if _g_nu_mu(xe, r) < 0. else _g_nu_mu(xe, r))
                else:
                    f1 = 0. if _h_nu_mu_1(xe, r) < 0. else 1.
                    if _h_nu_mu_2(xe, r) < 0.:
                        f1, f2 = 0., None                    # the reference zeroes flag_h_nu_mu_1 here and leaves ..._2 as it was
                    else:
                        f2 = 1.
                    if f2 is None:
                        f2 = Q_pp_sub1._flag2                # value left over from an earlier element (NameError if none)
                    Q_pp_sub1._flag2 = f2
                    shape.append(f1 * _h_nu_mu_1(xe, r) + f2 * _h_nu_mu_2(xe, r))
        shape = np.multiply(1. / np.trapz(shape, x_new), shape)
        pre = 2.
    else:
        raise ValueError("Unkown particles species")
    mpi = X_M_PI0 if species == "g" else X_M_PI
    integrand = q_pi(E_pi, p_pr, species, N_pr_TeV, n_H, g_pr) / np.sqrt(E_pi ** 2. - 0. * mpi ** 2.) * E_pi
    if shape is not None:
        integrand = shape * integrand
    return pre * np.trapz(integrand, np.log(E_pi))


def _Q_pp(E_species, g_pr, N_pr_TeV, p_pr, n_H, parts, cut_lo):
    """Common body of Q_e_pp / Q_g_pp / Q_nu_mu_pp / Q_nu_e_pp (LeHaMoC_f.py:860-1014): split the proton energies into
    the K&A range (x < 0.427 [and x > 1e-3 for neutrinos], E > 0.1 TeV) and the delta-function range (E < 0.1 TeV),
    then rescale the low-energy part so that it meets the two-point power-law extrapolation of the high-energy part."""
    E_p_space = g_pr * m_pr * c ** 2. * TEV
    out, norm_ind = [], 0
    for E in E_species:
        xs = E / E_p_space
        if cut_lo:
            s2 = sorted(xs[(E > 0.1) & (xs < 0.427) & (xs > 10 ** (-3.))])
            s1 = sorted(xs[(E < 0.1) & (xs < 10 ** (-3.)) & ~((E > 0.1) & (xs < 0.427) & (xs > 10 ** (-3.)))])
        else:
            s2 = sorted(xs[(E > 0.1) & (xs < 0.427)])
            s1 = sorted(xs[(E < 0.1) & (xs < 1.) & ~((E > 0.1) & (xs < 0.427))])
        q = 0.
        if len(s2) > 0:
            q += sum(Q_pp_sub2(s2, sp, E / np.asarray(s2), E, p_pr, N_pr_TeV, n_H, g_pr) for sp in parts)
        if len(s1) > 0:
            norm_ind += 1
            q += sum(Q_pp_sub1(s1, sp, E / np.asarray(s1), E, p_pr, N_pr_TeV, n_H, g_pr) for sp in parts)
        out.append(q)
    from scipy import stats
    with np.errstate(all="ignore"):
        slope, intercept = stats.linregress(np.log10(E_species[norm_ind:norm_ind + 2]), np.log10(out[norm_ind:norm_ind + 2]))[:2]
    y_fit = 10 ** (slope * np.log10(E_species[norm_ind - 1]) + intercept)
    norm = y_fit / out[norm_ind - 1]
    out[:norm_ind] = np.multiply(norm, out[:norm_ind])
    return out


def Q_e_pp(g_el, g_pr, N_pr_TeV, p_pr, n_H): return _Q_pp(g_el * m_el * c ** 2. * TEV, g_pr, N_pr_TeV, p_pr, n_H, ("e",), False)
def Q_g_pp(nu_ic, g_pr, N_pr_TeV, p_pr, n_H): return _Q_pp(h * nu_ic * TEV, g_pr, N_pr_TeV, p_pr, n_H, ("g",), False)
def Q_nu_mu_pp(nu_nu, g_pr, N_pr_TeV, p_pr, n_H): return _Q_pp(h * nu_nu * TEV, g_pr, N_pr_TeV, p_pr, n_H, ("nu_mu_1", "nu_mu_2"), True)
def Q_nu_e_pp(nu_nu, g_pr, N_pr_TeV, p_pr, n_H): return _Q_pp(h * nu_nu * TEV, g_pr, N_pr_TeV, p_pr, n_H, ("nu_e",), True)
