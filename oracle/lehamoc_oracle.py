"""CPU oracle: NumPy restatement of LeHaMoC's leptonic per-timestep kinetic update.

TEST INFRASTRUCTURE ONLY.  This module is the *checker* for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.
Nothing under `lehamoc_b200/` imports it and the product never falls back to it.

What it restates (all citations relative to the upstream repository mariapetro/LeHaMoC):
  * the function library `LeMoC/LeHaMoC_f.py` (byte-identical to `Fit_emcee/LeHaMoC_f.py`):
    lines 39-83 helpers, 86 Q_syn_space, 101 aSSA, 113 Q_IC_space_optimized, 127 a_gg,
    140 Q_ee_f, 154 photons_tot, 160 U_ph_f, 175 cor_factor_syn_el, 241 thomas;
  * the timestep drivers `LeMoC/LeMoC.py:107-296` (variant "script") and
    `Fit_emcee/Code.py:51-277` (variant "code");
  * the observation transform and likelihood of `Fit_emcee/fit_emcee.py:95-145,173-199`.

Parity pinning: `oracle/make_golden.py` runs the unmodified reference from /root/reference (build
container only) and stores its per-step state in `tests/golden/*.npz`;
`tests/test_oracle_vs_golden.py` checks this restatement against those vectors and against the
reference's own shipped outputs (`LeMoC/*_Distribution_Test1_LeMoC.txt`, which pin only ~1e-3).
The reference ships no fixture or test for the likelihood; it is pinned with the reference's OWN
definitions instead: `oracle/make_golden_fit.py` exec's `read_data`, `Model`, `log_likelihood_LeMoC`,
`log_prior_LeMoC`, `log_probability` unmodified out of `Fit_emcee/fit_emcee.py` (the module itself
cannot be imported: it runs the MCMC at import) on the reference's data files and stores the results
in `tests/golden/fit_loglike.npz`; `tests/test_fit.py` checks this restatement against them.

Two evaluation styles are provided for the two expensive functions:
  * `literal=True`  -- the same Python-loop structure as the reference (one call per frequency,
    inner loop over gamma): this is what `bench.py` times as the CPU baseline, because its cost
    is the reference's cost;
  * `literal=False` -- the same arithmetic broadcast over whole arrays (about 10x faster), used
    by the tests.  Both are checked against each other and against the live reference.
"""
from __future__ import annotations

import math
import numpy as np

if not hasattr(np, "trapz"):              # numpy >= 2.4
    np.trapz = np.trapezoid               # type: ignore[attr-defined]

# --- constants: astropy 5.0.4 (CODATA 2018) CGS values, LeHaMoC_f.py:17-31 -------------------
G = 6.6743e-8
c = 2.99792458e10
kb = 1.380649e-16
h = 6.62607015e-27
q = 4.803204712570263e-10
sigmaT = 6.6524587321e-25
m_el = 9.1093837015e-28
m_pr = 1.67262192369e-24
eV = 1.602176634e-12
pc = 3.0856775814913674e18
SENT = 10 ** (-260.)                      # boundary / empty-bin sentinel, LeMoC.py:174-183


# --- helpers, LeHaMoC_f.py:39-83 ---------------------------------------------------------------
def R(R0, t, t_i, Vexp):
    return R0 + Vexp * (t - t_i)


def B(B0, R0, R_, m):
    return B0 * (R0 / R_) ** m


def Volume(R_):
    return 4. * np.pi / 3. * R_ ** 3.


def Q_el_Lum(L_el, p_el, g_min_el, g_max_el):
    if p_el == 2.:
        return L_el / (m_el * c ** 2. * np.log(g_max_el / g_min_el))
    return L_el * (-p_el + 2.) / (m_el * c ** 2. * (g_max_el ** (-p_el + 2.) - g_min_el ** (-p_el + 2.)))


def Lum_e_inj(comp, R_):
    return 4. * np.pi * R_ * m_el * c ** 3. * comp / sigmaT


def nu_c(gamma, B_):
    return 3. / (4. * np.pi) * q * B_ / (m_el * c) * gamma ** 2.


C_SYN_EL = sigmaT * c / (h * 24. * np.pi ** 2. * 0.8975) * (4. * np.pi * m_el * c / (3. * q)) ** (4. / 3.)


# --- A4 synchrotron emissivity, LeHaMoC_f.py:86-92 ----------------------------------------------
def Q_syn_space(Np, B_, nu, a_cr, g):
    integrand = Np * g ** (1. / 3.) * np.exp(-nu / (a_cr * g ** 2.))
    integrand[m_el * c ** 2 * g < h * nu] = 0
    return C_SYN_EL * B_ ** (2. / 3.) * nu ** (-2. / 3.) * np.trapz(integrand, np.log(g))


def Q_syn_all(Np, B_, nus, a_cr, g):
    """Q_syn_space for every nu in `nus` at once (same arithmetic, broadcast over nu)."""
    nus = np.asarray(nus)[:, None]
    with np.errstate(under="ignore"):
        integrand = Np[None, :] * g[None, :] ** (1. / 3.) * np.exp(-nus / (a_cr * g[None, :] ** 2.))
    integrand[(m_el * c ** 2 * g)[None, :] < h * nus] = 0
    return C_SYN_EL * B_ ** (2. / 3.) * nus[:, 0] ** (-2. / 3.) * np.trapz(integrand, np.log(g), axis=1)


# --- A6 synchrotron self-absorption, LeHaMoC_f.py:101-102 ----------------------------------------
def aSSA(N_el, B_, nu, g, dg_l_el):
    gi = g[1:-1]
    return q ** 3. * B_ * nu ** (-5. / 3.) / (2. * m_el ** 2. * c ** 2. * 0.8975) * np.trapz(
        (2. * nu_c(gi, B_)) ** (-1. / 3.) * np.exp(-nu / nu_c(gi, B_))
        * (np.diff(N_el[1:]) / dg_l_el - 2. * N_el[1:-1]), np.log(gi))


def aSSA_all(N_el, B_, nus, g, dg_l_el):
    gi = g[1:-1]
    nus = np.asarray(nus)[:, None]
    nc = nu_c(gi, B_)[None, :]
    with np.errstate(under="ignore"):
        y = (2. * nc) ** (-1. / 3.) * np.exp(-nus / nc) * (np.diff(N_el[1:]) / dg_l_el - 2. * N_el[1:-1])[None, :]
    return q ** 3. * B_ * nus[:, 0] ** (-5. / 3.) / (2. * m_el ** 2. * c ** 2. * 0.8975) * np.trapz(y, np.log(gi), axis=1)


# --- A7 inverse-Compton emissivity, LeHaMoC_f.py:113-124 -----------------------------------------
def Q_IC_space_optimized(Np, g_el, nu_ic_temp, photons, nu_targ, index):
    """One outgoing frequency; Python loop over gamma exactly as the reference does."""
    inner = []
    lx = np.log(nu_targ[:index])
    for gam in g_el:
        q_IC = nu_ic_temp / (4. * nu_targ * gam * (gam - h * nu_ic_temp / (m_el * c ** 2.)))
        if q_IC[0] > 0.:
            w = h * nu_ic_temp / (gam * m_el * c ** 2. - h * nu_ic_temp)
            F = 2. * q_IC * np.log(q_IC) + (1. + 2. * q_IC) * (1. - q_IC) + w ** 2. / (1. + w) * (1. - q_IC) / 2.
            F[F < 0.] = 0.
            inner.append(np.trapz(photons[:index] * F[:index], lx))
        else:
            inner.append(0.)
    return sigmaT * c * 0.75 * np.trapz(Np * g_el ** (-1.) * inner, np.log(g_el))


def Q_IC_all(Np, g_el, nus_out, photons, nu_targ, index):
    """All outgoing frequencies: per nu_out the (gamma x nu_in) plane is evaluated at once."""
    out = np.empty(len(nus_out))
    lx = np.log(nu_targ[:index])
    lg = np.log(g_el)
    nt = nu_targ[None, :index]
    gam = g_el[:, None]
    for k, nu_o in enumerate(nus_out):
        e_o = h * nu_o / (m_el * c ** 2.)
        with np.errstate(all="ignore"):
            q_IC = nu_o / (4. * nt * gam * (gam - e_o))
            active = nu_o / (4. * nu_targ[0] * g_el * (g_el - e_o)) > 0.
            w = h * nu_o / (gam * m_el * c ** 2. - h * nu_o)
            F = 2. * q_IC * np.log(q_IC) + (1. + 2. * q_IC) * (1. - q_IC) + w ** 2. / (1. + w) * (1. - q_IC) / 2.
            F[F < 0.] = 0.
            inner = np.trapz(photons[None, :index] * F, lx, axis=1)
        inner[~active] = 0.
        out[k] = sigmaT * c * 0.75 * np.trapz(Np * g_el ** (-1.) * inner, lg)
    return out


# --- A8 gamma-gamma opacity, LeHaMoC_f.py:127-137 ------------------------------------------------
def a_gg(nu_ic, nu_target, photons_target, npts=50):
    out = []
    x = h * nu_target / (m_el * c ** 2.)
    lx = np.log10(x)
    with np.errstate(divide="ignore"):
        ly = np.log10(photons_target / h * m_el * c ** 2.)
    xmax = max(x)
    for nu in nu_ic:
        x_IC = h * nu / (m_el * c ** 2.)
        xs = np.logspace(np.log10(1.3 / x_IC), np.log10(xmax), npts)
        if xs[0] < xs[1]:
            ph = 10 ** np.interp(np.log10(xs), lx, ly)
            s = xs * x_IC
            out.append(np.trapz(0.652 * sigmaT * (s ** 2. - 1.) / s ** 3. * np.log(s) * xs * ph, np.log(xs)))
        else:
            out.append(0.)
    return np.array(out)


# --- A9 pair injection, LeHaMoC_f.py:140-151 -----------------------------------------------------
def Q_ee_f(nu_target, photons_target, nu_ic, photons_IC, g, R0=None):
    out = []
    lxt = np.log10(h * nu_target / (m_el * c ** 2.))
    with np.errstate(divide="ignore"):
        lyt = np.log10(photons_target * m_el * c ** 2. / h)
        lxi, lyi = np.log10(nu_ic), np.log10(photons_IC)
    xhi = np.log10(h * nu_target[-1] / (m_el * c ** 2.))
    for g_e in g:
        if (2. * g_e) > 1.:
            xp = np.logspace(np.log10((2. * g_e) ** (-1.)), xhi)          # default num=50
            nph = 10 ** np.interp(np.log10(xp), lxt, lyt)
            s = 2. * g_e * xp
            n_g = 10. ** np.interp(np.log10(2. * g_e * m_el * c ** 2. / h), lxi, lyi)
            R_gg = 2.61 * (s ** 2. - 1) / s ** 3. * np.log(s)
            out.append(n_g * np.trapz(nph * R_gg * xp, np.log(xp)))
        else:
            out.append(0.)
    return np.multiply(c * sigmaT * m_el * c ** 2. / h, out[:-1])


# --- A1 photon-field merge, LeHaMoC_f.py:154-156 -------------------------------------------------
def user_field(User_ph, user_spectrum, nu_tot):
    """LeMoC.py:160-169 / LeHaMoC.py:201-210: the user table (log10 nu, log10 dN/dVdnu), as read from
    Photons_spec_user.txt, on nu_tot."""
    if User_ph == 0.:
        return np.zeros(len(nu_tot))
    if user_spectrum is None:
        raise ValueError("User_ph != 0 needs the (logx, logy) table of Photons_spec_user.txt")
    nu_user = 10 ** np.array(user_spectrum[0], dtype=np.float64)
    tmp = 10 ** np.array(user_spectrum[1], dtype=np.float64)
    tmp[-1] = 10 ** (-160.)
    with np.errstate(divide="ignore", invalid="ignore"):
        return 10 ** np.interp(np.log10(nu_tot), np.log10(nu_user), np.log10(tmp))


def photons_tot(nu_syn, nu_bb, photons_syn, nu_ic, photons_IC, nu_tot, photons_bb, photons_pl, photons_user):
    with np.errstate(divide="ignore", invalid="ignore"):
        lt = np.log10(nu_tot)
        return (10 ** np.interp(lt, np.log10(nu_bb), np.log10(photons_bb))
                + 10 ** np.interp(lt, np.log10(nu_syn), np.log10(photons_syn))
                + 10 ** np.interp(lt, np.log10(nu_ic), np.log10(photons_IC))
                + photons_pl + photons_user)


# --- A2 Thomson-limited photon energy density, LeHaMoC_f.py:160-172 ------------------------------
def U_ph_f(g, nu_target, photons_target, R_):
    out = []
    K = sigmaT * R_ * m_el * c ** 2. / h
    pre = m_el * c ** 2. / (sigmaT * R_)
    U_tot = pre * np.trapz(nu_target[:-1] * photons_target[:-1] * K, nu_target[:-1])
    lnu = np.log10(nu_target)
    lph = np.log10(photons_target)
    for l_f in g:
        nuT = (3. * m_el * c ** 2.) / (4. * h * l_f)
        if nuT < nu_target[-1]:
            v = np.interp(np.log10(nuT), lnu, lph)
            idx = int(np.max(np.where(lnu < np.log10(nuT))[0])) + 1
            ph = np.insert(photons_target, idx, 10 ** v)
            nu = np.insert(nu_target, idx, nuT)
            x = h * nu / (m_el * c ** 2.)
            out.append(pre * np.trapz(x[:idx + 1] * ph[:idx + 1] * K, x[:idx + 1]))
        else:
            out.append(U_tot)
    return np.array(out)


# --- A3 Thomas algorithm, LeHaMoC_f.py:241-264 ---------------------------------------------------
def thomas(a, b, c_, d):
    n = len(a)
    cp = np.zeros(n)
    dp = np.zeros(n)
    x = np.zeros(n)
    cp[0] = c_[0] / b[0]
    dp[0] = d[0] / b[0]
    for i in range(1, n):
        den = b[i] - a[i] * cp[i - 1]
        cp[i] = c_[i] / den
        dp[i] = (d[i] - a[i] * dp[i - 1]) / den
    x[n - 1] = dp[n - 1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


def bidiag_solve(b, c_, d):
    """thomas() specialised to a == 0 (every call site of the drivers passes zeros for `a`)."""
    cp = c_ / b
    dp = d / b
    x = np.empty_like(dp)
    x[-1] = dp[-1]
    for i in range(len(b) - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


# --- A5 synchrotron energy-balance correction, LeHaMoC_f.py:175-237 ------------------------------
def cor_factor_syn_el(g_el, R0, B0, p_el, Lum_e_injected, literal=False):
    time_end = 20.
    step_alg = 1.
    n = len(g_el)
    g_mp = np.array([(g_el[i + 1] + g_el[i - 1]) / 2. for i in range(0, n - 1)])
    const_el = 4. / 3. * sigmaT / (8. * np.pi * m_el * c)
    dg = np.array([(g_el[i + 1] - g_el[i - 1]) / 2. for i in range(1, n - 1)])
    nu_syn = np.logspace(7.5, np.log10(7. * nu_c(g_el[-1], B0)) + 1.4, 100)
    day_counter = 0.
    N_el = np.zeros(n)
    photons_syn = np.ones(len(nu_syn)) * 10 ** (-260.)
    N_el[0] = N_el[-1] = 10 ** (-160.)
    el_inj = Q_el_Lum(Lum_e_injected, p_el, g_el[1], g_el[-2]) * g_el ** (-p_el)
    time_real = 0.
    ratio = None
    Vol = Volume(R0)
    a_cr = 3. * q * B0 / (4. * np.pi * m_el * c)
    b_syn = const_el * B0 ** 2.
    while time_real < time_end * R0 / c:
        dt = 0.1001 * R0 / c
        time_real += dt
        syn_m = b_syn * np.divide(np.power(g_mp[0:-1], 2.), dg)
        syn_p = b_syn * np.divide(np.power(g_mp[1:], 2.), dg)
        V2 = 1. + dt * (c / R0 + syn_m)
        V3 = -dt * syn_p
        S = N_el[1:-1] + el_inj[1:-1].copy() * dt
        N_el[1:-1] = bidiag_solve(V2, V3, S)
        if literal:
            Q = np.array([Q_syn_space(N_el / Vol, B0, nu_syn[k], a_cr, g_el) for k in range(len(nu_syn) - 1)])
        else:
            Q = Q_syn_all(N_el / Vol, B0, nu_syn[:-1], a_cr, g_el)
        V2p = 1. + dt * (c / R0 * np.ones(len(nu_syn) - 2))
        S = photons_syn[1:-1] + 4. * np.pi * np.multiply(Q, dt)[1:] * Vol
        photons_syn[1:-1] = S / V2p
        if day_counter < time_real:
            day_counter = day_counter + step_alg * R0 / c
            spec = np.multiply(photons_syn / Vol, h * nu_syn ** 2.) * 4. * np.pi / 3. * R0 ** 2. * c
            El = np.trapz(el_inj * g_el ** 2. * m_el * c ** 2., np.log(g_el))
            ratio = np.trapz(spec, np.log(nu_syn)) / El
    return ratio


# --- parameter file, LeMoC.py:60-100 / Code.py:63-94 ---------------------------------------------
def read_parameter_file(path):
    out = {}
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            kv = line.split("=")
            out[kv[0].strip()] = float(kv[1].strip())
    return out


class LeptonicRun:
    """State + stepping of one model.  variant="script": LeMoC/LeMoC.py; variant="code":
    Fit_emcee/Code.py::LeMoC (free parameters `theta` = [logR0, logB0, log g_min, log g_max,
    log l_e, p], frozen ones from the parameter file)."""

    def __init__(self, P: dict, variant="script", theta=None, literal=False, hoist_cor=True):
        self.variant = variant
        self.literal = literal
        self.hoist_cor = hoist_cor
        self._cor = None
        g = lambda k: float(P[k])
        self.time_init, self.time_end, self.step_alg = g('time_init'), g('time_end'), g('step_alg')
        if variant == "script":                                   # LeMoC.py:67-100
            self.g_min_el, self.g_max_el = g('g_min_el'), g('g_max_el')
            self.g_PL_min, self.g_PL_max = g('g_PL_min'), g('g_PL_max')
            self.p_el = g('p_el')
            self.R0 = 10 ** g('R0')
            self.B0 = g('B0')
            self.comp_el = sigmaT * 10 ** g('L_el') / (4. * np.pi * self.R0 * m_el * c ** 3)   # :110
            self.comp_inj = self.comp_el
            self.comp_cor = self.comp_el
        elif variant == "code":                                   # Code.py:52-61
            self.R0 = 10 ** theta[0]
            self.B0 = 10 ** theta[1]
            self.g_PL_min, self.g_PL_max = theta[2], theta[3]
            self.p_el = theta[5]
            self.g_min_el = self.g_PL_min - 1.
            self.g_max_el = self.g_PL_max + 1.
            self.comp_inj = 10 ** theta[4]                        # Code.py:163
            self.comp_cor = theta[4]                              # Code.py:237 (raw log value)
        elif variant == "code_lh":                                # Code.py:280-300 (proton-synchrotron model)
            self.R0 = 10 ** theta[0]
            self.B0 = 10 ** theta[1]
            self.g_PL_min, self.g_PL_max = theta[2], theta[3]
            self.p_el = theta[5]
            self.g_min_el = self.g_PL_min - 0.5
            self.g_max_el = self.g_PL_max + 0.5
            self.comp_inj = 10 ** theta[4]
            self.comp_cor = theta[4]
            self.g_PL_min_pr, self.g_PL_max_pr, self.comp_pr, self.p_pr = theta[6], theta[7], theta[8], theta[9]
            self.g_min_pr = self.g_PL_min_pr - 0.1
            self.g_max_pr = self.g_PL_max_pr + 0.5
            self.grid_g_pr = g('grid_g_pr')
        else:
            raise ValueError(variant)
        self.grid_g_el, self.grid_nu = g('grid_g_el'), g('grid_nu')
        self.Vexp = g('Vexp') * c
        self.m = g('m')
        if variant == "code_lh":
            P = dict(P, Ad_l_flag=0.)                             # Code.py:281: no adiabatic losses in this model
            g = lambda k: float(P[k])
        for k in ('inj_flag', 'Ad_l_flag', 'Syn_l_flag', 'Syn_emis_flag', 'IC_l_flag', 'IC_emis_flag',
                  'SSA_l_flag', 'gg_flag', 'esc_flag', 'BB_flag', 'GB_ext', 'PL_flag', 'dE_dV_ph',
                  'nu_min_ph', 'nu_max_ph', 's_ph', 'User_ph'):
            setattr(self, k, g(k))
        self.temperature = 10 ** g('BB_temperature')
        self.user_spectrum = P.get("_user_spectrum")
        self._setup()

    # LeMoC.py:107-191 (== Code.py:96-180)
    def _setup(self):
        R0, B0 = self.R0, self.B0
        self.time_real = self.time_init
        self.dt = self.step_alg * R0 / c
        self.day_counter = 0.
        n = int(self.grid_g_el)
        g_el = np.logspace(self.g_min_el, self.g_max_el, n)
        self.g_el = g_el
        self.g_el_mp = np.array([(g_el[i + 1] + g_el[i - 1]) / 2. for i in range(0, n - 1)])
        self.dg_el = np.array([(g_el[i + 1] - g_el[i - 1]) / 2. for i in range(1, n - 1)])
        self.dg_l_el = np.log(g_el[1]) - np.log(g_el[0])
        if self.g_PL_max == self.g_max_el:
            self.index_PL_max = -1
        else:
            self.index_PL_max = int(np.min(np.where(g_el > 10 ** self.g_PL_max)[0]))
        if self.g_PL_min == 0.:
            self.index_PL_min = 1
        else:
            self.index_PL_min = int(np.max(np.where(g_el < 10 ** self.g_PL_min)[0]))
        nu_syn = np.logspace(7.5, np.log10(7. * nu_c(g_el[-1], B0)) + 1.4, int(self.grid_g_el / 2))
        nu_ic = np.logspace(10., 30., int(self.grid_g_el / 2))
        nu_tot = np.logspace(np.log10(nu_syn[0]), np.log10(nu_ic[-1]), int(self.grid_nu))
        self.a_gg_f = np.zeros(len(nu_ic))
        if self.BB_flag == 0.:
            bb = np.zeros(2)
            nu_bb = np.array([nu_syn[0], nu_syn[-1]])
        else:                                                     # LeMoC.py:144-148
            T = self.temperature
            nu_bb = np.logspace(np.log10(5.879 * 10 ** 10 * T) - 6., np.log10(5.879 * 10 ** 10 * T) + 1.5, 60)
            with np.errstate(over="ignore"):
                Bnu = 2.0 * h * nu_bb ** 3 / c ** 2 / np.expm1(h * nu_bb / (kb * T))   # astropy BlackBody
            ph_bb = 4. * np.pi / c * Bnu / (h * nu_bb)
            GB_norm = np.trapz(ph_bb * h * nu_bb ** 2., np.log(nu_bb)) / self.GB_ext
            bb = ph_bb / GB_norm
        if self.PL_flag == 0.:
            pl = np.zeros(len(nu_tot))
        else:                                                     # LeMoC.py:155-158 (literal, incl. its defects)
            sp = np.logspace(self.nu_min_ph, self.nu_max_ph, 100)
            k_ph = (np.trapz(self.dE_dV_ph * sp ** (-self.s_ph + 1.))) ** (-1.)
            sp[-1] = 0.
            with np.errstate(divide="ignore", invalid="ignore"):
                pl = 10 ** np.interp(np.log10(nu_tot), np.log10(sp), np.log10(k_ph * sp ** (-self.s_ph)))
        user = user_field(self.User_ph, self.user_spectrum, nu_tot)
        self.Q_ee = np.zeros(n - 1)
        el_inj = np.ones(n) * SENT
        a, b_ = self.index_PL_min, self.index_PL_max
        el_inj[a:b_] = Q_el_Lum(Lum_e_inj(self.comp_inj, R0), self.p_el, g_el[a], g_el[b_]) * g_el[a:b_] ** (-self.p_el)
        self.el_inj = el_inj
        self.N_el = el_inj.copy()
        self.N_el[0] = self.N_el[-1] = SENT
        self.photons_syn = np.ones(len(nu_syn) + 1) * SENT
        self.photons_IC = np.ones(len(nu_ic)) * SENT
        self.bb = np.append(bb, SENT)
        self.nu_syn = np.append(nu_syn, nu_tot[-1])
        self.nu_bb = np.append(nu_bb, nu_tot[-1])
        self.nu_ic, self.nu_tot, self.pl, self.user = nu_ic, nu_tot, pl, user
        ns, ni = self.nu_syn, nu_ic
        self.nu_syn_mp = np.array([(ns[i + 1] + ns[i - 1]) / 2. for i in range(0, len(ns) - 1)])
        self.nu_ic_mp = np.array([(ni[i + 1] + ni[i - 1]) / 2. for i in range(0, len(ni) - 1)])
        self.dnu = np.array([(ns[i + 1] - ns[i - 1]) / 2. for i in range(1, len(ns) - 1)])
        self.dnu_ic = np.array([(ni[i + 1] - ni[i - 1]) / 2. for i in range(1, len(ni) - 1)])
        self.Radius = R0
        self.nstep = 0
        self.trace = {}
        if self.variant == "code_lh":                             # Code.py:325-331,341-347,419-425
            gp = np.logspace(self.g_min_pr, self.g_max_pr, int(self.grid_g_pr))
            self.g_pr = gp
            self.g_pr_mp = np.array([(gp[i + 1] + gp[i - 1]) / 2. for i in range(0, len(gp) - 1)])
            self.dg_pr = np.array([(gp[i + 1] - gp[i - 1]) / 2. for i in range(1, len(gp) - 1)])
            ipmax = -1 if self.g_PL_max_pr == self.g_max_pr else int(np.min(np.where(gp > 10 ** self.g_PL_max_pr)[0]))
            ipmin = 1 if self.g_PL_min_pr == 0. else int(np.max(np.where(gp < 10 ** self.g_PL_min_pr)[0]))
            pr_inj = np.ones(len(gp)) * SENT
            Lp = 4. * np.pi * R0 * m_pr * c ** 3. * (10 ** self.comp_pr) / sigmaT
            if self.p_pr == 2.:
                Q0 = Lp / (m_pr * c ** 2. * np.log(gp[ipmax] / gp[ipmin]))
            else:
                Q0 = Lp * (-self.p_pr + 2.) / (m_pr * c ** 2. * (gp[ipmax] ** (-self.p_pr + 2.) - gp[ipmin] ** (-self.p_pr + 2.)))
            pr_inj[ipmin:ipmax] = Q0 * gp[ipmin:ipmax] ** (-self.p_pr)
            self.pr_inj = pr_inj
            self.N_pr = pr_inj.copy()
            self.N_pr[0] = self.N_pr[-1] = SENT

    def n_steps(self):
        if self.variant == "script":
            return int(self.time_end)                             # LeMoC.py:196
        # Code.py:183 / :447
        n, t = 0, self.time_init                                  # Code.py:183
        while t < self.time_end * self.R0 / c:
            t += self.dt
            n += 1
        return n

    def _photons(self):
        V = Volume(self.Radius)
        return photons_tot(self.nu_syn, self.nu_bb, self.photons_syn, self.nu_ic, self.photons_IC,
                           self.nu_tot, self.bb * V, self.pl * V, self.user * V) / V

    def cor_factor(self):
        if self._cor is None or not self.hoist_cor:
            self._cor = cor_factor_syn_el(self.g_el, self.R0, 10 ** 4., self.p_el,
                                          Lum_e_inj(self.comp_cor, self.R0), literal=self.literal)
        return self._cor

    # LeMoC.py:198-284 / Code.py:184-269
    def step(self, record=False):
        g_el, nu_syn, nu_ic, nu_tot = self.g_el, self.nu_syn, self.nu_ic, self.nu_tot
        dt = self.dt
        self.time_real += dt
        Radius = self.Radius = R(self.R0, self.time_real, self.time_init, self.Vexp)
        M_F = B(self.B0, self.R0, Radius, self.m)
        a_cr = 3. * q * M_F / (4. * np.pi * m_el * c)
        V = Volume(Radius)
        photons = self._photons()
        one = 1. if self.variant == "script" else 0.              # gamma^2-1 (LeMoC.py:224) vs gamma^2 (Code.py:210)
        mp, dg = self.g_el_mp, self.dg_el
        n = len(g_el)
        if self.Ad_l_flag == 1.:
            b_ad = self.Vexp / Radius
            ad_m = b_ad * np.divide(np.power(mp[0:-1], 1.), dg)
            ad_p = b_ad * np.divide(np.power(mp[1:], 1.), dg)
            ads_m = b_ad * np.divide(self.nu_syn_mp[0:-1], self.dnu)
            ads_p = b_ad * np.divide(self.nu_syn_mp[1:], self.dnu)
            adi_m = b_ad * np.divide(self.nu_ic_mp[0:-1] - nu_ic[:-2], self.dnu_ic)
            adi_p = b_ad * np.divide(nu_ic[:1], self.dnu_ic)
        else:
            ad_m = ad_p = np.zeros(n - 2)
            ads_m = ads_p = np.zeros(len(nu_syn) - 2)
            adi_m = adi_p = np.zeros(len(nu_ic) - 2)
        if self.Syn_l_flag == 1.:
            b_syn = (4. / 3.) * sigmaT / (8. * np.pi * m_el * c) * M_F ** 2.
            syn_m = b_syn * np.divide(np.power(mp[0:-1], 2.) - one, dg)
            syn_p = b_syn * np.divide(np.power(mp[1:], 2.) - one, dg)
        else:
            syn_m = syn_p = np.zeros(n - 2)
        if self.IC_l_flag == 1.:
            U_ph = U_ph_f(g_el, nu_tot, photons, Radius)
            b_Com = 4. / 3. * sigmaT * np.multiply(c, U_ph) / (m_el * c ** 2.)
            ic_m = b_Com[1:-1] * np.divide(np.power(mp[0:-1], 2.) - one, dg)
            ic_p = b_Com[2:] * np.divide(np.power(mp[1:], 2.) - one, dg)
        else:
            U_ph = None
            ic_m = ic_p = np.zeros(n - 2)
        V2 = 1. + dt * (c / Radius * self.esc_flag + syn_m + ic_m + ad_m)
        V3 = -dt * (syn_p + ic_p + ad_p)
        if self.inj_flag == 1.:
            S = self.N_el[1:-1] + np.multiply(self.el_inj[1:-1], dt) + np.multiply(self.Q_ee[1:], dt) * V
        elif self.inj_flag == 0.:
            S = self.N_el[1:-1] + np.multiply(self.Q_ee[1:], dt) * V
        else:
            raise ValueError("inj_flag must be 0 or 1 (the reference leaves S_ij undefined otherwise)")
        self.N_el[1:-1] = bidiag_solve(V2, V3, S)
        ne = np.array(self.N_el / V)
        Q_syn_pr = None
        if self.variant == "code_lh":                             # Code.py:480-490,494
            pm, dgp = self.g_pr_mp, self.dg_pr
            if self.Syn_l_flag == 1.:
                b_p = (4. / 3.) * sigmaT / (8. * np.pi * m_el * c) * M_F ** 2. * (m_el / m_pr) ** 3.
                sp_m, sp_p = b_p * np.power(pm[0:-1], 2.) / dgp, b_p * np.power(pm[1:], 2.) / dgp
            else:
                sp_m = sp_p = np.zeros(len(self.g_pr) - 2)
            V2p = 1. + dt * (c / Radius * self.esc_flag + sp_m)
            V3p = -dt * sp_p
            Sp = self.N_pr[1:-1] + (np.multiply(self.pr_inj[1:-1], dt) if self.inj_flag == 1. else 0.)
            self.N_pr[1:-1] = bidiag_solve(V2p, V3p, Sp)
            if self.Syn_emis_flag == 1.:
                n_p = np.array(self.N_pr / V)
                a_cr_pr = 3. * q * M_F / (4. * np.pi * m_pr * c)
                C_pr = sigmaT * c / (h * 24. * np.pi ** 2. * 0.8975) * (4. * np.pi * m_pr * c / (3. * q)) ** (4. / 3.) * (m_el / m_pr) ** 2.
                nus = np.asarray(nu_syn[:-1])[:, None]
                with np.errstate(under="ignore"):
                    integ = n_p[None, :] * self.g_pr[None, :] ** (1. / 3.) * np.exp(-nus / (a_cr_pr * self.g_pr[None, :] ** 2.))
                Q_syn_pr = C_pr * M_F ** (2. / 3.) * nus[:, 0] ** (-2. / 3.) * np.trapz(integ, np.log(self.g_pr), axis=1)
            else:
                Q_syn_pr = np.zeros(len(nu_syn) - 1)
        if self.Syn_emis_flag == 1.:
            if self.literal:
                Qs = np.array([Q_syn_space(ne, M_F, nu_syn[k], a_cr, g_el) for k in range(len(nu_syn) - 1)])
            else:
                Qs = Q_syn_all(ne, M_F, nu_syn[:-1], a_cr, g_el)
            Q_syn_raw = Qs
            Q_syn = np.divide(Qs, self.cor_factor())
        else:
            Q_syn_raw = Q_syn = np.zeros(len(nu_syn) - 1)
        idx = len(nu_tot) - 1
        if self.IC_emis_flag == 1.:
            if self.literal:
                Q_IC = np.array([Q_IC_space_optimized(ne, g_el, nu_ic[k], photons, nu_tot, idx)
                                 for k in range(0, len(nu_ic) - 1)])
            else:
                Q_IC = Q_IC_all(ne, g_el, nu_ic[:-1], photons, nu_tot, idx)
        else:
            Q_IC = np.zeros(len(nu_ic) - 1)
        if self.SSA_l_flag == 1.:
            if self.literal:
                aS = np.array([-np.absolute(aSSA(ne, M_F, nu_syn[k], g_el, self.dg_l_el)) for k in range(len(nu_syn))])
                aI = np.array([-np.absolute(aSSA(ne, M_F, nu_ic[k], g_el, self.dg_l_el)) for k in range(len(nu_ic))])
            else:
                aS = -np.absolute(aSSA_all(ne, M_F, nu_syn, g_el, self.dg_l_el))
                aI = -np.absolute(aSSA_all(ne, M_F, nu_ic, g_el, self.dg_l_el))
        else:
            aS = np.zeros(len(nu_syn))
            aI = np.zeros(len(nu_ic))
        V2 = 1. + dt * (c / Radius + ads_m - np.multiply(aS[1:-1], 1) * c)
        V3 = -dt * ads_p
        S = self.photons_syn[1:-1] + 4. * np.pi * np.multiply(Q_syn, dt)[1:] * V
        if Q_syn_pr is not None:
            S = S + 4. * np.pi * np.multiply(Q_syn_pr, dt)[1:] * V  # Code.py:515
        self.photons_syn[1:-1] = bidiag_solve(V2, V3, S)
        V2 = 1. + dt * (c / Radius + adi_m + np.multiply(self.a_gg_f[1:-1], c) - np.multiply(aI[1:-1], 1) * c)
        V3 = -dt * adi_p
        S = self.photons_IC[1:-1] + np.multiply(Q_IC, dt)[1:] * V
        self.photons_IC[1:-1] = bidiag_solve(V2, V3, S)
        if self.gg_flag == 0.:
            self.a_gg_f = np.zeros(len(nu_ic))
        else:
            self.a_gg_f = a_gg(nu_ic, nu_tot, photons)
            if self.variant == "script":                           # LeMoC.py:283
                self.Q_ee = Q_ee_f(nu_tot, photons, nu_tot, photons, g_el, Radius)
            else:                                                  # Code.py:269
                self.Q_ee = Q_ee_f(nu_tot, photons, nu_ic, self.photons_IC / V, g_el, Radius)
        self.nstep += 1
        if record:
            tr = self.trace
            for k, v in (("photons_top", photons), ("U_ph", U_ph), ("N_el", self.N_el), ("Q_syn_raw", Q_syn_raw),
                         ("Q_IC", Q_IC), ("aSSA_syn", aS), ("aSSA_ic", aI), ("photons_syn", self.photons_syn),
                         ("photons_IC", self.photons_IC), ("a_gg", self.a_gg_f), ("Q_ee", self.Q_ee)):
                if v is not None:
                    tr.setdefault(k, []).append(np.array(v, dtype=np.float64, copy=True))

    # LeMoC.py:286-291 / Code.py:271-277
    def output(self):
        with np.errstate(divide="ignore"):
            photons = self._photons()
            spec = np.multiply(photons, h * self.nu_tot ** 2.) * 4. * np.pi / 3. * self.Radius ** 2. * c
            return self.N_el / Volume(self.Radius), spec

    def run(self, record=False):
        """Returns (n_e[steps, Ng], nuLnu[steps, Nt]) -- one snapshot per step, linear units."""
        ne_all, sed_all = [], []
        for _ in range(self.n_steps()):
            self.step(record=record)
            ne, sed = self.output()
            ne_all.append(ne)
            sed_all.append(sed)
        return np.array(ne_all), np.array(sed_all)


def run_script(param_file_or_dict, literal=False, record=False, hoist_cor=True):
    """`python LeMoC.py <Parameters.txt> <suffix>` (LeMoC.py:45-296) without the text I/O."""
    P = param_file_or_dict if isinstance(param_file_or_dict, dict) else read_parameter_file(param_file_or_dict)
    run = LeptonicRun(P, "script", literal=literal, hoist_cor=hoist_cor)
    ne, sed = run.run(record=record)
    return run, ne, sed


def code_LeHaMoC(theta, param_file_or_dict, literal=False, hoist_cor=True):
    """Fit_emcee/Code.py::LeHaMoC(params[10], fileName) -> (nu_tot, nuLnu of the last step)."""
    P = param_file_or_dict if isinstance(param_file_or_dict, dict) else read_parameter_file(param_file_or_dict)
    run = LeptonicRun(P, "code_lh", theta=list(theta), literal=literal, hoist_cor=hoist_cor)
    _, sed = run.run()
    return run.nu_tot, sed[-1]


def code_LeMoC(theta, param_file_or_dict, literal=False, hoist_cor=True):
    """Fit_emcee/Code.py::LeMoC(params, fileName) -> (nu_tot, nuLnu of the last step)."""
    P = param_file_or_dict if isinstance(param_file_or_dict, dict) else read_parameter_file(param_file_or_dict)
    run = LeptonicRun(P, "code", theta=list(theta), literal=literal, hoist_cor=hoist_cor)
    _, sed = run.run()
    return run.nu_tot, sed[-1]


# --- observation transform + likelihood, fit_emcee.py:95-145,173-199 ------------------------------
class Model:
    def __init__(self, numcode, D=3262, z=0.557, param_file='Parameters.txt'):
        self.D = D * 1e6 * pc
        self.z = z
        self.distance_norm = np.log10(4 * np.pi * self.D ** 2)
        self.numcode = numcode
        self.param_file = param_file

    def transform(self, x, nu, sed, doppler):
        flux_norm = -self.distance_norm + 4 * doppler
        xp = np.log10(nu) + doppler - np.log10(1. + self.z)
        with np.errstate(divide="ignore"):
            yp = np.log10(sed) + flux_norm
        return np.interp(x, xp, yp)

    def __call__(self, x, theta):
        nu, sed = self.numcode(theta[:-1], self.param_file)
        return self.transform(x, nu, sed, theta[-1])


def log_likelihood_from_model(y_model, log_f, y, yerr, yerr1, flags):
    """fit_emcee.py:121-145 given the model already interpolated onto the data abscissae."""
    from scipy.special import erf
    mask0 = (flags == 0)
    mask1 = (flags == 1)
    y_det, y_ul = y[mask0], y[mask1]
    yerr_det = yerr[mask0].copy()      # NB the reference mutates the caller's array through this
    yerr1_det = yerr1[mask0]           # boolean-mask copy only, so a copy is what it acts on too
    ym_ul, ym_det = y_model[mask1], y_model[mask0]
    if ym_det[-2] < y_det[-2]:
        yerr_det[-2] = yerr1_det[-2]
    if ym_det[-1] < y_det[-1]:
        yerr_det[-1] = yerr1_det[-1]
    sigma2 = yerr_det ** 2 + np.exp(2 * log_f)
    rms = 0.05
    with np.errstate(divide="ignore"):
        temp = erf((y_ul - ym_ul) / ((2 ** 0.5) * rms))
        lim = 1. * np.sum(np.log((np.pi / 2.) ** 0.5 * rms * (1. + temp)))
    det = -0.5 * np.sum((y_det - ym_det) ** 2 / sigma2 + np.log(sigma2))
    return det + lim


def log_prior_LeMoC(theta):
    P1, P2, P3, P4, P5, P6, P7, log_f = theta
    if (14 < P1 < 17) and (-2 < P2 < 2) and (0. < P3 < 4.5) and (5.0 < P4 < 7) and (-6 < P5 < -2) \
            and (1.5 < P6 < 3) and (0 < P7 < 3) and (-6.0 < log_f < 1.0):
        return 0.0
    return -np.inf


def read_data(filename1, filename2):
    """fit_emcee.py:63-88 without pandas."""
    df = np.loadtxt(filename2, usecols=range(0, 4))
    E = np.log10(df[:, 0])
    vFv = np.log10(df[:, 1] * 1e6 * eV)
    hi = np.log10(df[:, 1] + df[:, 2]) - np.log10(df[:, 1])
    lo = -np.log10(df[:, 1] - df[:, 2]) + np.log10(df[:, 1])
    d = np.loadtxt(filename1, usecols=range(0, 4))
    xd = np.append(d[:, 0], E)
    yd = np.append(d[:, 2], vFv)
    return xd, yd, np.append(d[:, 3], hi), np.append(d[:, 3], lo), np.append(np.zeros(len(d)), df[:, 3])
