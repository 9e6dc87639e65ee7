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
