"""CPU oracle of the leptohadronic driver `LeHaMoC Code and Tests/LeHaMoC.py` (BASELINE configs 3 and 5).

TEST INFRASTRUCTURE ONLY (same rules as lehamoc_oracle.py).

Restates the driver LeHaMoC.py:132-453 with the `f_had` variants of the function library
(`LeHaMoC Code and Tests/LeHaMoC_f.py`): Q_syn_space :98-103 (C_syn and the particle mass passed in), a_gg :145-155
(100 samples), Q_ee_f :158-168 (100 samples, full length), U_ph_f :171-182 (50-point resample), photons_tot :1025.
Electron AND proton equations, proton synchrotron, the gamma-gamma pre-attenuation of the first two steps
(LeHaMoC.py:418-426) and the neutrino equation are included; of the hadronic channels the photopion and Bethe-Heitler
LOSS terms, the photopion emissivities and the Bethe-Heitler pair injection (spline surfaces handed in as `bh_surf`)
and the pp loss term and secondaries (LeHaMoC.py:322-328, 354-357, 387-395) are wired to lehamoc_oracle_had.py.

Parity pinning: oracle/make_golden_lh.py runs the unmodified script (oracle/run_reference_lh.py) and stores every
step's four output snapshots in tests/golden/lh_*.npz; tests/test_lehamoc_driver.py checks this restatement against them.
"""
from __future__ import annotations

import numpy as np

if not hasattr(np, "trapz"):
    np.trapz = np.trapezoid               # type: ignore[attr-defined]

import lehamoc_oracle as lo
from lehamoc_oracle import (B, R, Volume, Q_el_Lum, Lum_e_inj, nu_c, aSSA_all, Q_IC_all, bidiag_solve,  # noqa: F401
                            c, h, m_el, m_pr, q, sigmaT, eV, kb, SENT)

LH_KEYS = ("time_init", "time_end", "step_alg", "PL_inj", "g_min_el", "g_max_el", "g_el_PL_min", "g_el_PL_max",
           "grid_g_el", "g_min_pr", "g_max_pr", "g_pr_PL_min", "g_pr_PL_max", "grid_g_pr", "grid_nu", "p_el", "L_el",
           "p_pr", "L_pr", "Vexp", "R0", "B0", "m", "delta", "inj_flag", "Ad_l_flag", "Syn_l_flag", "Syn_emis_flag",
           "IC_l_flag", "IC_emis_flag", "SSA_l_flag", "gg_flag", "pg_pi_l_flag", "pg_pi_emis_flag", "pg_BH_l_flag",
           "pg_BH_emis_flag", "n_H", "pp_l_flag", "pp_ee_emis_flag", "pp_g_emis_flag", "pp_nu_emis_flag",
           "neutrino_flag", "esc_flag_el", "esc_flag_pr", "BB_flag", "temperature", "GB_ext", "PL_flag", "dE_dV_ph",
           "nu_min_ph", "nu_max_ph", "s_ph", "User_ph")


def Q_pr_Lum(L_pr, p_pr, g_min_pr, g_max_pr):            # LeHaMoC_f.py:75-79
    if p_pr == 2.:
        return L_pr / (m_pr * c ** 2. * np.log(g_max_pr / g_min_pr))
    return L_pr * (-p_pr + 2.) / (m_pr * c ** 2. * (g_max_pr ** (-p_pr + 2.) - g_min_pr ** (-p_pr + 2.)))


def Lum_p_inj(comp, R0):                                 # LeHaMoC_f.py:89-90
    return 4. * np.pi * R0 * m_pr * c ** 3. * comp / sigmaT


def Q_norm_exp_c(R0, comp, g, g_max_ind, p_i):           # LeHaMoC_f.py:81-82
    return 4. * np.pi * R0 * c * comp / (sigmaT * np.trapz(g[:g_max_ind] ** (-p_i + 2.) * np.exp(-g[:g_max_ind] / g[g_max_ind]),
                                                            np.log(g[:g_max_ind])))


def Q_syn_all_m(Np, B_, nus, a_cr, C_syn, g, m):          # LeHaMoC_f.py:98-103 for all nu at once
    nus = np.asarray(nus)[:, None]
    with np.errstate(under="ignore"):
        integrand = Np[None, :] * g[None, :] ** (1. / 3.) * np.exp(-nus / (a_cr * g[None, :] ** 2.))
    integrand[(m * c ** 2 * g)[None, :] < h * nus] = 0
    return C_syn * B_ ** (2. / 3.) * nus[:, 0] ** (-2. / 3.) * np.trapz(integrand, np.log(g), axis=1)


def a_gg(nu_ic, nu_target, photons_target):                # LeHaMoC_f.py:145-155
    return lo.a_gg(nu_ic, nu_target, photons_target, npts=100)


def Q_ee_f(nu_target, photons_target, nu_ic, photons_IC, g, R0=None):     # LeHaMoC_f.py:158-168
    out = []
    lxt = np.log10(h * nu_target / (m_el * c ** 2.))
    with np.errstate(divide="ignore"):
        lyt = np.log10(photons_target * m_el * c ** 2. / h)
        lxi, lyi = np.log10(nu_ic), np.log10(photons_IC)
    xhi = np.log10(h * nu_target[-1] / (m_el * c ** 2.))
    for g_e in g:
        if (2. * g_e) > 1.:
            xp = np.logspace(np.log10((2. * g_e) ** (-1.)), xhi, 100)
            nph = 10 ** np.interp(np.log10(xp), lxt, lyt)
            s = 2. * g_e * xp
            n_g = 10. ** np.interp(np.log10(2. * g_e * m_el * c ** 2. / h), lxi, lyi)
            R_gg = 2.61 * (s ** 2. - 1) / s ** 3. * np.log(s)
            out.append(n_g * np.trapz(nph * R_gg * xp, np.log(xp)))
        else:
            out.append(0.)
    return np.multiply(c * sigmaT * m_el * c ** 2. / h, out)


def U_ph_f(g, nu_target, photons_target, R_):              # LeHaMoC_f.py:171-182
    out = []
    K = sigmaT * R_ * m_el * c ** 2. / h
    pre = m_el * c ** 2. / (sigmaT * R_)
    U_tot = pre * np.trapz(nu_target[:-1] * photons_target[:-1] * K, nu_target[:-1])
    lnu = np.log10(nu_target)
    with np.errstate(divide="ignore"):
        lph = np.log10(photons_target)
    for l_f in g:
        nu_T = 3. * m_el * c ** 2. / (4. * h * l_f)
        if nu_T < nu_target[-1]:
            nu_temp = np.logspace(np.log10(nu_target[0]), np.log10(nu_T), 50)
            ph = 10. ** np.interp(np.log10(nu_temp), lnu, lph)
            nd = h * nu_temp / (m_el * c ** 2.)
            out.append(pre * np.trapz(nd * ph * K, nd))
        else:
            out.append(U_tot)
    return np.array(out)


def photons_tot(nu_syn, nu_bb, photons_syn, nu_ic, photons_IC, nu_tot, photons_bb, photons_pl, photons_user):
    with np.errstate(divide="ignore", invalid="ignore"):    # LeHaMoC_f.py:1025-1026 (bare in the reference)
        lt = np.log10(nu_tot)
        return (10 ** np.interp(lt, np.log10(nu_bb), np.log10(photons_bb))
                + 10 ** np.interp(lt, np.log10(nu_syn), np.log10(photons_syn))
                + 10 ** np.interp(lt, np.log10(nu_ic), np.log10(photons_IC)) + photons_pl + photons_user)


class LeptoHadronicRun:
    def __init__(self, P: dict, tables=None):
        missing = [k for k in LH_KEYS if k not in P]
        if missing:
            raise KeyError(missing)
        for k in LH_KEYS:
            setattr(self, k, float(P[k]))
        self.Vexp = self.Vexp * c
        self.R0 = 10 ** float(P["R0"])
        self.temperature = 10 ** float(P["temperature"])
        self.tables = tables
        self.user_spectrum = P.get("_user_spectrum")
        self._setup()

    def _setup(self):                                        # LeHaMoC.py:132-264
        R0, B0 = self.R0, self.B0
        self.time_real = self.time_init
        self.dt = self.step_alg * R0 / c
        self.comp_el = sigmaT * 10 ** self.L_el / (4. * np.pi * R0 * m_el * c ** 3)
        self.comp_pr = sigmaT * 10 ** self.L_pr / (4. * np.pi * R0 * m_pr * c ** 3)
        g_el = np.logspace(self.g_min_el, self.g_max_el, int(self.grid_g_el))
        self.g_el = g_el
        self.g_el_mp = np.array([(g_el[i + 1] + g_el[i - 1]) / 2. for i in range(0, len(g_el) - 1)])
        self.dg_el = np.array([(g_el[i + 1] - g_el[i - 1]) / 2. for i in range(1, len(g_el) - 1)])
        self.dg_l_el = np.log(g_el[1]) - np.log(g_el[0])
        g_pr = np.logspace(self.g_min_pr, self.g_max_pr, int(self.grid_g_pr))
        self.g_pr = g_pr
        self.g_pr_mp = np.array([(g_pr[i + 1] + g_pr[i - 1]) / 2. for i in range(0, len(g_pr) - 1)])
        self.dg_pr = np.array([(g_pr[i + 1] - g_pr[i - 1]) / 2. for i in range(1, len(g_pr) - 1)])
        ie_max = -1 if self.g_el_PL_max == self.g_max_el else int(np.min(np.where(g_el > 10 ** self.g_el_PL_max)[0]))
        ie_min = 1 if self.g_el_PL_min == 0. else int(np.max(np.where(g_el < 10 ** self.g_el_PL_min)[0]))
        ip_max = -1 if self.g_pr_PL_max == self.g_max_pr else int(np.min(np.where(g_pr > 10 ** self.g_pr_PL_max)[0]) + 1)
        ip_min = 1 if self.g_pr_PL_min == 0. else int(np.max(np.where(g_pr < 10 ** self.g_pr_PL_min)[0]) + 1)
        nh = int(self.grid_g_el / 2)
        nu_syn = np.logspace(7.5, np.log10(7. * nu_c(g_el[-1], B0)) + 1.4, nh)
        nu_ic = np.logspace(15., 32., nh)
        self.nu_nu = np.logspace(10., 22., 50) * eV / h
        nu_tot = np.logspace(np.log10(nu_syn[0]), np.log10(max(nu_ic[-1], nu_syn[-1])), int(self.grid_nu))
        if self.BB_flag == 0.:
            bb = np.zeros(2)
            nu_bb = np.array([nu_syn[0], nu_syn[-1]])
        else:
            T = self.temperature
            nu_bb = np.logspace(np.log10(5.879 * 10 ** 10 * T) - 6., np.log10(5.879 * 10 ** 10 * T) + 1.5, 60)
            with np.errstate(over="ignore"):
                Bnu = 2.0 * h * nu_bb ** 3 / c ** 2 / np.expm1(h * nu_bb / (kb * T))
            ph_bb = 4. * np.pi / c * Bnu / (h * nu_bb)
            bb = ph_bb / (np.trapz(ph_bb * h * nu_bb ** 2., np.log(nu_bb)) / self.GB_ext)
        if self.PL_flag == 0.:
            pl = np.zeros(len(nu_tot))
        else:                                                # LeHaMoC.py:195-199
            sp = np.logspace(self.nu_min_ph, self.nu_max_ph, 100)
            k_ph = self.dE_dV_ph / np.trapz(sp ** (-self.s_ph + 1.), sp) / h
            tmp = k_ph * sp ** (-self.s_ph)
            tmp[-1] = 10 ** (-260.)
            pl = 10 ** np.interp(np.log10(nu_tot), np.log10(sp), np.log10(tmp))
        self.user = lo.user_field(self.User_ph, self.user_spectrum, nu_tot)
        self.pl = pl
        el_inj = np.ones(len(g_el)) * SENT
        pr_inj = np.ones(len(g_pr)) * SENT
        V0 = Volume(R0)
        if self.PL_inj == 1.:
            el_inj[ie_min:ie_max] = Q_el_Lum(Lum_e_inj(self.comp_el, R0), self.p_el, g_el[ie_min], g_el[ie_max]) \
                * g_el[ie_min:ie_max] ** (-self.p_el) / V0
            pr_inj[ip_min:ip_max] = Q_pr_Lum(Lum_p_inj(self.comp_pr, R0), self.p_pr, g_pr[ip_min], g_pr[ip_max]) \
                * g_pr[ip_min:ip_max] ** (-self.p_pr) / V0
        else:
            pr_inj = Q_norm_exp_c(R0, self.comp_pr, g_pr, ip_max, self.p_pr) * g_pr ** (-self.p_pr) * np.exp(-g_pr / g_pr[ip_max]) / V0
            el_inj = Q_norm_exp_c(R0, self.comp_el, g_el, ie_max, self.p_el) * g_el ** (-self.p_el) * np.exp(-g_el / g_el[ie_max]) / V0
        self.el_inj, self.pr_inj = el_inj, pr_inj
        self.idx = (ie_min, ie_max, ip_min, ip_max)
        self.N_pr = np.zeros(len(g_pr))
        self.N_nu = np.zeros(50)
        self.Q_ee = np.zeros(len(g_el) - 2)
        self.photons_syn = np.ones(len(nu_syn) + 1) * SENT
        self.photons_IC = np.ones(len(nu_ic)) * SENT
        self.bb = np.append(bb, SENT)
        self.nu_syn = np.append(nu_syn, nu_tot[-1])
        self.nu_bb = np.append(nu_bb, nu_tot[-1])
        self.nu_ic, self.nu_tot = nu_ic, nu_tot
        ns, ni = self.nu_syn, nu_ic
        self.nu_syn_mp = np.array([(ns[i + 1] + ns[i - 1]) / 2. for i in range(0, len(ns) - 1)])
        self.nu_ic_mp = np.array([(ni[i + 1] + ni[i - 1]) / 2. for i in range(0, len(ni) - 1)])
        self.dnu = np.array([(ns[i + 1] - ns[i - 1]) / 2. for i in range(1, len(ns) - 1)])
        self.dnu_ic = np.array([(ni[i + 1] - ni[i - 1]) / 2. for i in range(1, len(ni) - 1)])
        self.a_gg_syn = np.zeros(len(ns) - 2)
        self.a_gg_ic = np.zeros(len(ni) - 2)
        self.cor = lo.cor_factor_syn_el(g_el, R0, 10 ** 4., self.p_el, Lum_e_inj(self.comp_el, R0))
        self.C_syn_el = sigmaT * c / (h * 24. * np.pi ** 2. * 0.8975) * (4. * np.pi * m_el * c / (3. * q)) ** (4. / 3.) / self.cor
        self.C_syn_pr = sigmaT * c / (h * 24. * np.pi ** 2. * 0.8975) * (4. * np.pi * m_pr * c / (3. * q)) ** (4. / 3.) * (m_el / m_pr) ** 2.
        self.const_el = 4. / 3. * sigmaT / (8. * np.pi * m_el * c)
        self.const_pr = self.const_el * (m_el / m_pr) ** 3.
        self.N_el = el_inj.copy() * V0
        self.N_el[0] = self.N_el[-1] = SENT
        self.interv = 0.
        self.Radius = R0
        self.bh_surf = None                                   # (tck_m, tck_p, max_m, max_p) of interp_cs_BH_int

    def n_steps(self):
        return int(self.time_end / self.step_alg)            # LeHaMoC.py:267

    def _photons(self):
        V = Volume(self.Radius)
        return photons_tot(self.nu_syn, self.nu_bb, self.photons_syn, self.nu_ic, self.photons_IC, self.nu_tot,
                           self.bb * V, self.pl * V, self.user * V) / V

    def step(self):                                          # LeHaMoC.py:268-434
        import lehamoc_oracle_had as oh
        g_el, g_pr, nu_syn, nu_ic, nu_tot, dt = self.g_el, self.g_pr, self.nu_syn, self.nu_ic, self.nu_tot, self.dt
        self.time_real += dt
        Radius = self.Radius = R(self.R0, self.time_real, self.time_init, self.Vexp)
        M_F = B(self.B0, self.R0, Radius, self.m)
        a_cr_el = 3. * q * M_F / (4. * np.pi * m_el * c)
        a_cr_pr = 3. * q * M_F / (4. * np.pi * m_pr * c)
        V = Volume(Radius)
        photons = self._photons()
        emp, dge, pmp, dgp = self.g_el_mp, self.dg_el, self.g_pr_mp, self.dg_pr
        ne_, np_ = len(g_el) - 2, len(g_pr) - 2
        if self.Ad_l_flag == 1.:
            b_ad = self.Vexp / Radius
            ad_el_m, ad_el_p = b_ad * emp[0:-1] / dge, b_ad * emp[1:] / dge
            ads_m, ads_p = b_ad * self.nu_syn_mp[0:-1] / self.dnu, b_ad * self.nu_syn_mp[1:] / self.dnu
            adi_m = b_ad * np.divide(self.nu_ic_mp[0:-1] - nu_ic[:-2], self.dnu_ic)
            adi_p = b_ad * np.divide(nu_ic[:1], self.dnu_ic)
        else:
            ad_el_m = ad_el_p = np.zeros(ne_)
            ads_m = ads_p = np.zeros(len(nu_syn) - 2)
            adi_m = adi_p = np.zeros(len(nu_ic) - 2)
        if self.Syn_l_flag == 1.:
            b_e, b_p = self.const_el * M_F ** 2., self.const_pr * M_F ** 2.
            syn_el_m, syn_el_p = b_e * (emp[0:-1] ** 2. - 1) / dge, b_e * (emp[1:] ** 2. - 1) / dge
            syn_pr_m, syn_pr_p = b_p * (pmp[0:-1] ** 2. - 1) / dgp, b_p * (pmp[1:] ** 2. - 1) / dgp
        else:
            syn_el_m = syn_el_p = np.zeros(ne_)
            syn_pr_m = syn_pr_p = np.zeros(np_)
        if self.IC_l_flag == 1.:
            U_ph = U_ph_f(g_el, nu_tot, photons, Radius)
            b_Com = 4. / 3. * sigmaT * np.multiply(c, U_ph) / (m_el * c ** 2.)
            ic_m, ic_p = b_Com[1:-1] * (emp[0:-1] ** 2. - 1) / dge, b_Com[2:] * (emp[1:] ** 2. - 1) / dge
        else:
            ic_m = ic_p = np.zeros(ne_)
        if self.pg_pi_l_flag == 1.:
            pg_m = oh.dg_dt_pg(pmp[0:-1], nu_tot, photons, self.tables) / dgp
            pg_p = oh.dg_dt_pg(pmp[1:], nu_tot, photons, self.tables) / dgp
        else:
            pg_m = pg_p = np.zeros(np_)
        if self.pg_BH_l_flag == 1.:
            bh_m = oh.dg_dt_BH(pmp[0:-1], nu_tot, photons, self.tables) / dgp
            bh_p = oh.dg_dt_BH(pmp[1:], nu_tot, photons, self.tables) / dgp
        else:
            bh_m = bh_p = np.zeros(np_)
        if self.pp_l_flag == 1.:                               # LeHaMoC.py:322-328
            rest = m_pr * c ** 2. * 0.624151
            pp_m = np.divide(0.65 * c * self.n_H * oh.cs_pp_inel(pmp[0:-1] * m_pr * c ** 2. * 0.624151 - rest) / (m_pr * c ** 2.), dgp)
            pp_p = np.divide(0.65 * c * self.n_H * oh.cs_pp_inel(pmp[1:] * m_pr * c ** 2. * 0.624151 - rest) / (m_pr * c ** 2.), dgp)
        else:
            pp_m = pp_p = np.zeros(np_)
        V2 = 1. + dt * (c / Radius * self.esc_flag_pr + syn_pr_m + pg_m + bh_m + pp_m)
        V3 = -dt * (syn_pr_p + pg_p + bh_p + pp_p)
        S = self.N_pr[1:-1] + np.multiply(self.pr_inj[1:-1], dt) * V
        self.N_pr[1:-1] = bidiag_solve(V2, V3, S)
        n_p = np.array(self.N_pr / V)
        if self.pg_BH_emis_flag == 1.:                          # LeHaMoC.py:337-340
            Q_pg_BH = oh.Q_BH_sol(g_el, g_pr, n_p, nu_tot * h / (m_el * c ** 2.), np.array(photons), *self.bh_surf)[1:-1]
        else:
            Q_pg_BH = np.zeros(ne_)
        if self.pg_pi_emis_flag == 1.:
            T = self.tables
            Q_pg_pi = oh.Qp_g_opt(g_el, nu_ic, n_p, g_pr, photons, nu_tot, "e+", T)[1:-1] \
                + oh.Qp_g_opt(g_el, nu_ic, n_p, g_pr, photons, nu_tot, "e-", T)[1:-1]
            Q_pg_g = oh.Qp_g_opt(g_el, nu_ic, n_p, g_pr, photons, nu_tot, "2_g", T)[1:-1]
            if self.neutrino_flag == 1.:
                Q_pg_nu = np.multiply(sum(oh.Qp_g_opt(g_el, nu_ic, self.N_pr, g_pr, photons, nu_tot, sp, T)
                                          for sp in ("nu_mu", "\bar_nu_mu", "nu_e", "\bar_nu_e")), dt)[1:-1]
            else:
                Q_pg_nu = np.zeros(48)
        else:
            Q_pg_pi, Q_pg_g, Q_pg_nu = np.zeros(ne_), np.zeros(len(nu_ic) - 2), np.zeros(48)
        N_TeV = n_p / (m_pr * c ** 2. * 0.624151)
        nu_nu = np.logspace(10., 22., 50) * eV / h               # LeHaMoC.py:173
        with np.errstate(all="ignore"):
            if self.pp_ee_emis_flag == 1.:                       # LeHaMoC.py:354-357
                Q_pp_ee = np.multiply(oh.Q_e_pp(g_el, g_pr, N_TeV, self.p_pr, self.n_H)[1:-1], m_el * c ** 2.)
            else:
                Q_pp_ee = np.zeros(ne_)
            if self.pp_g_emis_flag == 1.:                        # LeHaMoC.py:387-390
                Q_pp_g = np.multiply(oh.Q_g_pp(nu_ic, g_pr, N_TeV, self.p_pr, self.n_H)[1:-1], h)
            else:
                Q_pp_g = np.zeros(len(nu_ic) - 2)
            if self.pp_nu_emis_flag == 1.:                       # LeHaMoC.py:392-395
                Q_pp_nu = 2. * np.multiply(oh.Q_nu_e_pp(nu_nu, g_pr, N_TeV, self.p_pr, self.n_H)[1:-1], dt) * h * V \
                    + 2. * np.multiply(oh.Q_nu_mu_pp(nu_nu, g_pr, N_TeV, self.p_pr, self.n_H)[1:-1], dt) * h * V
            else:
                Q_pp_nu = np.zeros(48)
        V2 = 1. + dt * (c / Radius * self.esc_flag_el + syn_el_m + ic_m + ad_el_m)
        V3 = -dt * (syn_el_p + ic_p + ad_el_p)
        if self.inj_flag == 1.:
            S = self.N_el[1:-1] + np.multiply(self.el_inj[1:-1] + self.Q_ee + Q_pg_pi + Q_pg_BH + Q_pp_ee, dt) * V
        elif self.inj_flag == 0.:
            S = self.N_el[1:-1] + (self.Q_ee + Q_pg_pi + Q_pg_BH + Q_pp_ee) * dt * V
        else:
            raise ValueError("inj_flag must be 0 or 1")
        self.N_el[1:-1] = bidiag_solve(V2, V3, S)
        n_e = np.array(self.N_el / V)
        if self.Syn_emis_flag == 1.:
            Q_Syn_el = Q_syn_all_m(n_e, M_F, nu_syn[1:-1], a_cr_el, self.C_syn_el, g_el, m_el)
            Q_Syn_pr = Q_syn_all_m(n_p, M_F, nu_syn[1:-1], a_cr_pr, self.C_syn_pr, g_pr, m_pr)
        else:
            Q_Syn_el = Q_Syn_pr = np.zeros(len(nu_syn) - 2)
        if self.IC_emis_flag == 1.:
            Q_IC = Q_IC_all(n_e, g_el, nu_ic[:-1], photons, nu_tot, len(nu_tot) - 1)[1:]
        else:
            Q_IC = np.zeros(len(nu_ic) - 2)
        if self.SSA_l_flag == 1.:
            aS = -np.absolute(aSSA_all(n_e, M_F, nu_syn[1:-1], g_el, self.dg_l_el))
            aI = -np.absolute(aSSA_all(n_e, M_F, nu_ic[1:-1], g_el, self.dg_l_el))
        else:
            aS, aI = np.zeros(len(nu_syn) - 2), np.zeros(len(nu_ic) - 2)
        V2 = 1. + dt * (c / Radius + ads_m - aS * c + self.a_gg_syn * c)
        V3 = -dt * ads_p
        S = self.photons_syn[1:-1] + 4. * np.pi * (Q_Syn_el + Q_Syn_pr) * dt * V
        self.photons_syn[1:-1] = bidiag_solve(V2, V3, S)
        V2 = 1. + dt * (c / Radius + adi_m - aI * c + self.a_gg_ic * c)
        V3 = -dt * adi_p
        S = self.photons_IC[1:-1] + (Q_IC + Q_pg_g + Q_pp_g) * dt * V
        self.photons_IC[1:-1] = bidiag_solve(V2, V3, S)
        if self.gg_flag == 0.:
            self.a_gg_ic = np.zeros(len(nu_ic) - 2)
            self.a_gg_syn = np.zeros(len(nu_syn) - 2)
            self.Q_ee = np.zeros(ne_)
        else:
            self.a_gg_ic = np.array(a_gg(nu_ic, nu_tot, photons)[1:-1])
            self.a_gg_syn = np.array(a_gg(nu_syn, nu_tot, photons)[1:-1])
            a_tmp = np.array(a_gg(nu_tot, nu_tot, photons)[1:-1])
            if self.interv < 2.:
                pt = photons * V
                pt[1:-1] = pt[1:-1] / (1. + dt * (a_tmp * c))      # thomas() of a diagonal system, LeHaMoC.py:419-425
                pt[1:-1] = pt[1:-1] / V
                self.Q_ee = Q_ee_f(nu_tot, pt, nu_tot, pt, g_el, Radius)[1:-1]
            else:
                self.Q_ee = Q_ee_f(nu_tot, photons, nu_tot, photons, g_el, Radius)[1:-1]
        V2 = 1. + dt * (c / Radius * np.ones(48))
        self.N_nu[1:-1] = (self.N_nu[1:-1] + Q_pp_nu + Q_pg_nu) / V2
        self.interv += 1
        self.n_e, self.n_p = n_e, n_p

    def output(self):                                        # LeHaMoC.py:436-439
        with np.errstate(divide="ignore"):
            photons = self._photons()
            spec = np.multiply(photons, h * self.nu_tot ** 2.) * 4. * np.pi / 3. * self.Radius ** 2. * c
        return self.n_e, spec, self.n_p, self.N_nu.copy()

    def run(self):
        out = [[], [], [], []]
        for _ in range(self.n_steps()):
            self.step()
            for lst, v in zip(out, self.output()):
                lst.append(np.array(v, copy=True))
        return tuple(np.array(x) for x in out)
