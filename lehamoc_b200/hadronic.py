"""GPU twins of the photo-hadronic functions of `LeHaMoC Code and Tests/LeHaMoC_f.py`,
same names and argument meaning, batched over walkers underneath (lhm_had_loss_batch, lhm_photopion_batch):

    dg_dt_BH(g_pr, nu, photons, f_k_i=None)                                   LeHaMoC_f.py:185-196
    dg_dt_pg(g_pr, nu, photons)                                               LeHaMoC_f.py:201-220
    Qp_g_opt(g_el, nu_ic, N_pr, g_pr, photons_targ, nu_target, flag_product)  LeHaMoC_f.py:358-464

The reference reads its six physics tables from the working directory at import (LeHaMoC_f.py:14-19);
here `load_tables(dirname)` reads the same files (or `set_tables(dict of arrays)`), once.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import LhmError, check
from .constants import eV
from .engine import _require_cuda

TABLE_FILES = {"phi_g": "Phi_g_K&A.txt", "phi_el": "Phi_g_leptons_K&A.txt", "phi_el1": "Phi_e-nu_e_K&A.txt",
               "fk": "f(xi).csv", "cs": "cross_section.csv", "kp": "kp_pg.txt"}
_NCOL = {"phi_g": 4, "phi_el": 13, "phi_el1": 7, "fk": 2, "cs": 2, "kp": 2}
# flag_product strings of the reference (LeHaMoC_f.py:359-432; "\bar" really is backspace + "ar" there)
SPECIES = {"2_g": 0, "e+": 1, "e-": 2, "nu_mu": 3, "\bar_nu_mu": 4, "nu_e": 5, "\bar_nu_e": 6}
E_nu_space = np.logspace(10., 22., 50) * eV          # LeHaMoC_f.py:41


class lhm_had_tables(C.Structure):
    _fields_ = [(n, t) for k in ("phi_g", "phi_el", "phi_el1", "fk", "cs", "kp")
                for n, t in ((k, C.c_void_p), ("n_" + k, C.c_int))]


_tables = None        # (struct, device tensors kept alive)


def set_tables(arrays: dict, device=None):
    """arrays: the six tables as 2-D arrays keyed phi_g, phi_el, phi_el1, fk, cs, kp (rows as in the files)."""
    global _tables
    _require_cuda()
    dev = f"cuda:{torch.cuda.current_device() if device is None else device}"
    st, keep = lhm_had_tables(), []
    for k, ncol in _NCOL.items():
        a = np.ascontiguousarray(arrays[k], dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != ncol:
            raise ValueError(f"table {k}: expected [n,{ncol}], got {a.shape}")
        t = torch.from_numpy(a).to(dev)
        keep.append(t)
        setattr(st, k, t.data_ptr())
        setattr(st, "n_" + k, a.shape[0])
    _tables = (st, keep)


def load_tables(dirname=".", device=None):
    """Read the reference's table files from `dirname` (default: working directory, like the reference does)."""
    arrays = {}
    for k, fn in TABLE_FILES.items():
        arrays[k] = np.loadtxt(os.path.join(dirname, fn), delimiter="," if fn.endswith(".csv") else None)
    set_tables(arrays, device)


def _need_tables():
    if _tables is None:
        try:
            load_tables(".")
        except OSError as e:
            raise LhmError("hadronic tables not loaded: call lehamoc_b200.hadronic.load_tables(<directory of "
                           "Phi_g_K&A.txt ...>) first") from e
    return _tables[0]


def _dev2(a):
    a = np.atleast_2d(np.ascontiguousarray(a, dtype=np.float64))
    return torch.from_numpy(a).cuda()


def _p(t):
    return C.c_void_p(t.data_ptr())


def _loss(which, g_pr, nu, photons):
    st = _need_tables()
    lib = _lib.load()
    g, n, p = _dev2(g_pr), _dev2(nu), _dev2(photons)
    W = g.shape[0]
    if n.shape[0] != W or p.shape != n.shape:
        raise ValueError("g_pr [W,Ne], nu and photons [W,Nt] must agree on W")
    out = torch.empty_like(g)
    check(lib.lhm_had_loss_batch(W, g.shape[1], n.shape[1], which, _p(g), _p(n), _p(p), C.byref(st), _p(out),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lhm_had_loss_batch")
    r = out.cpu().numpy()
    return r[0] if np.ndim(g_pr) == 1 else r


def dg_dt_BH(g_pr, nu, photons, f_k_i=None):
    """Proton loss rate by Bethe-Heitler pair production at the Lorentz factors g_pr ([Ne] or [W,Ne])."""
    return _loss(1, g_pr, nu, photons)


def dg_dt_pg(g_pr, nu, photons):
    """Proton loss rate by photopion production."""
    return _loss(0, g_pr, nu, photons)


def Qp_g_opt(g_el, nu_ic, N_pr, g_pr, photons_targ, nu_target, flag_product):
    """Photopion secondaries; returns one value per point of the product's own grid (h nu_ic for "2_g", g_el m_e c^2
    for e+-, E_nu_space for the neutrinos).  1-D inputs: one walker; 2-D [W, n]: a batch."""
    st = _need_tables()
    lib = _lib.load()
    sp = SPECIES[flag_product]
    ge, ni, Np_, gp, ph, nt = (_dev2(x) for x in (g_el, nu_ic, N_pr, g_pr, photons_targ, nu_target))
    W = ge.shape[0]
    for t in (ni, Np_, gp, ph, nt):
        if t.shape[0] != W:
            raise ValueError("all inputs must have the same number of walkers")
    En = torch.from_numpy(E_nu_space).cuda()
    Nout = ni.shape[1] if sp == 0 else ge.shape[1] if sp in (1, 2) else len(E_nu_space)
    out = torch.empty((W, Nout), dtype=torch.float64, device="cuda")
    check(lib.lhm_photopion_batch(W, sp, ge.shape[1], ni.shape[1], gp.shape[1], nt.shape[1], len(E_nu_space), _p(ge),
                                  _p(ni), _p(En), _p(Np_), _p(gp), _p(ph), _p(nt), C.byref(st), _p(out),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lhm_photopion_batch")
    r = out.cpu().numpy()
    return r[0] if np.ndim(g_el) == 1 else r


# ---------------------------------------------------------------------------------------------------------------------
# A14 pp channels (LeHaMoC_f.py:687-1014).  Same signatures as the reference; 1-D inputs: one walker, 2-D: a batch.
# ---------------------------------------------------------------------------------------------------------------------
def _pp(product, grid, g_pr, N_pr_TeV, p_pr, n_H):
    lib = _lib.load()
    one = np.ndim(grid) == 1
    gr, gp, NT = _dev2(grid), _dev2(g_pr), _dev2(N_pr_TeV)
    W = gr.shape[0]
    if gp.shape[0] != W or NT.shape != gp.shape:
        raise ValueError("grid [W,Nout], g_pr and N_pr_TeV [W,Np] must agree on W")
    pp = torch.from_numpy(np.broadcast_to(np.asarray(p_pr, dtype=np.float64), (W,)).copy()).cuda()
    nh = torch.from_numpy(np.broadcast_to(np.asarray(n_H, dtype=np.float64), (W,)).copy()).cuda()
    out = torch.empty_like(gr)
    check(lib.lhm_pp_batch(W, product, gr.shape[1], gp.shape[1], _p(gr), _p(gp), _p(NT), _p(pp), _p(nh), _p(out),
                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lhm_pp_batch")
    r = out.cpu().numpy()
    return r[0] if one else r


def Q_e_pp(g_el, g_pr, N_pr_TeV, p_pr, n_H):
    """Secondary e+- from pp collisions per electron energy in TeV (LeHaMoC_f.py:860-895)."""
    return _pp(0, g_el, g_pr, N_pr_TeV, p_pr, n_H)


def Q_g_pp(nu_ic, g_pr, N_pr_TeV, p_pr, n_H):
    """pi0-decay photons from pp collisions (LeHaMoC_f.py:897-932)."""
    return _pp(1, nu_ic, g_pr, N_pr_TeV, p_pr, n_H)


def Q_nu_mu_pp(nu_nu, g_pr, N_pr_TeV, p_pr, n_H):
    """Muon neutrinos from pp collisions, both decay stages (LeHaMoC_f.py:934-975)."""
    return _pp(2, nu_nu, g_pr, N_pr_TeV, p_pr, n_H)


def Q_nu_e_pp(nu_nu, g_pr, N_pr_TeV, p_pr, n_H):
    """Electron neutrinos from pp collisions (LeHaMoC_f.py:977-1014)."""
    return _pp(3, nu_nu, g_pr, N_pr_TeV, p_pr, n_H)


# ---------------------------------------------------------------------------------------------------------------------
# A13 Bethe-Heitler pair injection: host set-up (LeHaMoC_f.py:238-341).  The reference tabulates the energy-integrated
# differential cross section on a 30 x 30 x 30 lattice (27 000 Simpson integrals) and fits two smoothing bivariate
# splines with scipy.interpolate.bisplrep; that fit stays on the host (SciPy/FITPACK, as in the reference), the spline
# EVALUATION (4.8 M bisplev calls per step in the reference) runs on the GPU.
# ---------------------------------------------------------------------------------------------------------------------
def _nonneg(a):
    """NaN -> 0, negative -> 0 (the reference scrubs its intermediates this way, LeHaMoC_f.py:239-265)."""
    a = np.where(np.isnan(a), 0., a)
    return np.where(a < 0., 0., a)


def _bh_lattice_integrand(E_el, gam_p, gam_e, kap):
    """Blumenthal's (1970) Bethe-Heitler differential cross section d sigma / dE_- on a whole lattice at once.

    E_el [M, n]: electron energies (m_e c^2) of M lattice points; gam_p, gam_e, kap [M, 1]: proton / electron Lorentz
    factor and photon energy in the proton rest frame.  This is the formula `cs_BH_diff` of LeHaMoC_f.py:238-273
    evaluates one lattice point at a time; every elementwise operation is kept in the reference's order (terms summed
    left to right) so that the doubles -- and with them the knots SciPy's bisplrep picks -- come out identical."""
    from .constants import sigmaT
    # kap is a scalar per lattice point in the reference, and NumPy raises SCALARS to a power with libm's pow() but
    # ARRAYS with its own SIMD loop (1 ulp apart on ~3 % of the lattice): take the scalar route for kap**2, kap**3.
    kap_sq = np.array([v ** 2. for v in kap[:, 0]]).reshape(-1, 1)
    kap_cu = np.array([v ** 3. for v in kap[:, 0]]).reshape(-1, 1)
    E_pos = np.where(np.isnan(kap - E_el), 0., kap - E_el)              # positron energy, scrubbed
    E_pos = np.where(E_pos < 0., 0., E_pos)
    mom_el = np.sqrt(E_el ** 2 - 1.)
    mom_el = np.where(mom_el < 0., 0., mom_el)
    mom_pos = _nonneg(np.sqrt(E_pos ** 2. - 1.))
    mu = (gam_p * E_el - gam_e) / (gam_p * mom_el)                      # cosine of the electron's polar angle
    Tm = np.sqrt(kap_sq + mom_el ** 2. - 2. * kap * mom_el * mu)
    Tm = np.where(Tm < 0., 0., Tm)
    Dm = _nonneg(E_el - mom_el * mu)
    Yl = _nonneg((2. / mom_el ** 2.) * np.log((E_el * E_pos + mom_el * mom_pos + 1.) / kap))
    yp = _nonneg(mom_pos ** (-1.) * np.log((E_pos + mom_pos) / (E_pos - mom_pos)))
    dT = _nonneg(np.log((Tm + mom_pos) / (Tm - mom_pos)))
    sin2 = (np.sin(np.arccos(mu))) ** 2.
    pref = 3. / (8. * np.pi * 137.) * sigmaT * mom_pos * mom_el / kap_cu
    t1 = -4. * sin2 * (2. * E_el ** 2. + 1.) / (mom_el ** 2. * Dm ** 4.)
    t2 = (5. * E_el ** 2. - 2. * E_el * E_pos + 3.) / (mom_el ** 2. * Dm ** 2.)
    t3 = (mom_el ** 2. - kap_sq) / (Tm ** 2. * Dm ** 2.)
    t4 = 2. * E_pos / (mom_el ** 2. * Dm)
    t5 = Yl / (mom_el * mom_pos) * (2. * E_el * sin2 * (3. * kap + mom_el ** 2. * E_pos) / Dm ** 4.
                                    + (2. * E_el ** 2 * (E_el ** 2. + E_pos ** 2.) - 7. * E_el ** 2. - 3. * E_el * E_pos
                                       - E_pos ** 2. + 1.) / Dm ** 2.
                                    + kap * (E_el ** 2. - E_el * E_pos - 1.) / Dm)
    t6 = dT / (mom_pos * Tm) * (2. / Dm ** 2. - 3. * kap / Dm - kap * (mom_el ** 2. - kap_sq) / (Tm ** 2. * Dm))
    t7 = 2. * yp / Dm
    return pref * (t1 + t2 + t3 + t4 + t5 - t6 - t7) / mom_el


_BH_TCK_VERSION = 1


def _bh_cache_file(key):
    import hashlib
    import scipy
    d = os.environ.get("LHM_CACHE_DIR") or os.path.join(os.path.expanduser("~"), ".cache", "lehamoc_b200")
    tag = hashlib.sha1(repr((key, scipy.__version__, np.__version__, _BH_TCK_VERSION)).encode()).hexdigest()[:20]
    return os.path.join(d, f"bh_surfaces_{tag}.npz")


def interp_cs_BH_int(g_pr, g_el, nu_int_min, nu_int_max, cache=True):
    """LeHaMoC_f.py:276-341: returns (tck_m, tck_p, max_m, max_p), the two smoothing-spline surfaces of
    log10 int sigma_BH E dlnE over (log10 Ee_min, log10 Ee_max) for gamma_p < gamma_e and gamma_p > gamma_e.

    The reference walks its 30 x 30 x 30 lattice in three Python loops (27 000 Simpson integrals over 50 electron
    energies each, ~7 s); here the whole lattice is ONE array expression [27 000, 50] in the reference's loop order
    (0.2 s), then SciPy's bisplrep exactly like the reference.  The result is persisted per (grid ends, frequency
    range, SciPy / NumPy version) under ~/.cache/lehamoc_b200 (LHM_CACHE_DIR overrides; cache=False recomputes)."""
    from scipy import integrate, interpolate
    from .constants import h, m_el, c
    key = tuple(float(v) for v in (g_pr[0], g_pr[-1], g_el[0], g_el[-1], nu_int_min, nu_int_max))
    path = _bh_cache_file(key)
    if cache and os.path.exists(path):
        try:
            z = np.load(path)
            return ([z["m_tx"], z["m_ty"], z["m_c"], 3, 3], [z["p_tx"], z["p_ty"], z["p_c"], 3, 3],
                    float(z["max_m"]), float(z["max_p"]))
        except Exception:
            pass
    gp_i = np.logspace(np.log10(g_pr[0]), np.log10(g_pr[-1]), 30)
    ge_i = np.logspace(np.log10(g_el[0]), np.log10(g_el[-1]), 30)
    x = np.logspace(np.log10(h * nu_int_min / (m_el * c ** 2.)), np.log10(h * nu_int_max / (m_el * c ** 2.)), 30)
    with np.errstate(all="ignore"):
        # lattice in the reference's loop order: proton, electron, photon
        GP, GE, X = (a.reshape(-1, 1) for a in np.meshgrid(gp_i, ge_i, x, indexing="ij"))
        kap = 2. * GP * X
        E_lo = (GP ** 2. + GE ** 2.) / (2. * GP * GE)
        E_el = np.logspace(np.log10(E_lo[:, 0]), np.log10(kap[:, 0] - 1.), 50, axis=-1)          # [27000, 50]
        ok = (E_el[:, 0] < E_el[:, -1]) & (GP[:, 0] != GE[:, 0])
        # filter BEFORE integrating: on a lattice that still holds its NaN rows SciPy's masked divisions change the
        # memory order of the summands and np.sum then adds each row in another order than the reference's 1-D calls
        E_el, GPk, GEk, kk = E_el[ok], GP[ok], GE[ok], kap[ok]
        sBH = _bh_lattice_integrand(E_el, GPk, GEk, kk)
        sBH = np.where(np.isnan(sBH), 0., sBH)
        tot = integrate.simpson(sBH * E_el, x=np.log(E_el), axis=-1)
    above = GPk[:, 0] > GEk[:, 0]
    out = {}
    for name, sel in (("m", ~above), ("p", above)):
        with np.errstate(all="ignore"):
            zz = np.where(np.isnan(np.log10(tot[sel])), 0.1, np.log10(tot[sel]))
        zz[zz == -np.inf] = 0.
        keep = np.where(zz < -15)[0]
        xs, ys, zs = np.log10(E_el[sel, 0])[keep], np.log10(E_el[sel, -1])[keep], zz[keep]
        out[name] = (interpolate.bisplrep(list(xs), list(ys), list(zs), s=2), max(zs))
    res = (out["m"][0], out["p"][0], out["m"][1], out["p"][1])
    if cache:
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = path + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, m_tx=res[0][0], m_ty=res[0][1], m_c=res[0][2], p_tx=res[1][0], p_ty=res[1][1], p_c=res[1][2],
                     max_m=res[2], max_p=res[3])
            os.replace(tmp, path)
        except OSError:
            pass
    return res


class lhm_bh_surface(C.Structure):
    _fields_ = [("tx", C.c_void_p), ("ty", C.c_void_p), ("c", C.c_void_p), ("nx", C.c_int), ("ny", C.c_int),
                ("zmax", C.c_double)]


def bh_surface(tck, zmax, device="cuda"):
    """(ctypes lhm_bh_surface, tensors to keep alive) from a bisplrep tck = [tx, ty, c, 3, 3] and max_intrp_cs."""
    if int(tck[3]) != 3 or int(tck[4]) != 3:
        raise ValueError("the Bethe-Heitler surfaces are bicubic (bisplrep default kx = ky = 3)")
    ts = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device) for a in tck[:3]]
    if ts[2].numel() != (ts[0].numel() - 4) * (ts[1].numel() - 4):
        raise ValueError("tck: len(c) != (nx-4)*(ny-4)")
    s = lhm_bh_surface(ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), ts[0].numel(), ts[1].numel(), float(zmax))
    return s, ts


def Q_BH_sol(g_el, g_pr, N_pr, nu_trgt, photons_trgt, intrp_cs_m, intrp_cs_p, max_intrp_cs_m, max_intrp_cs_p):
    """LeHaMoC_f.py:566-616, same arguments (nu_trgt is h nu/(m_e c^2), N_pr is dN/dV dgamma); 1-D inputs: one
    walker, 2-D [W, n]: a batch sharing the two spline surfaces."""
    _require_cuda()
    lib = _lib.load()
    ge, gp, Np_, xt, ph = (_dev2(x) for x in (g_el, g_pr, N_pr, nu_trgt, photons_trgt))
    W = ge.shape[0]
    sm, keep_m = bh_surface(intrp_cs_m, max_intrp_cs_m)
    sp, keep_p = bh_surface(intrp_cs_p, max_intrp_cs_p)
    out = torch.empty_like(ge)
    check(lib.lhm_bh_emis_batch(W, ge.shape[1], gp.shape[1], xt.shape[1], _p(ge), _p(gp), _p(Np_), _p(xt), _p(ph),
                                C.byref(sm), C.byref(sp), _p(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "lhm_bh_emis_batch")
    r = out.cpu().numpy()
    del keep_m, keep_p
    return r[0] if np.ndim(g_el) == 1 else r
