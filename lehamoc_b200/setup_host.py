"""Host side of a batch: parameter files, grids and initial conditions for W walkers at once.

Mirrors the set-up part of the reference drivers with the reference's own NumPy expressions, so
grid doubles and integer indices are identical by construction:
  * LeMoC/LeMoC.py:60-191   (variant "script": every parameter comes from the parameter file),
  * Fit_emcee/Code.py:51-180 (variant "code": six free parameters + frozen parameter file).
Everything state-dependent happens on the GPU; nothing here depends on N(gamma) or n(nu).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from operator import itemgetter

import numpy as np

from .constants import c, h, kb, m_el, q, sigmaT, SENTINEL

if not hasattr(np, "trapz"):                       # numpy >= 2.4
    np.trapz = np.trapezoid                        # type: ignore[attr-defined]

FLAG_KEYS = ("inj_flag", "Ad_l_flag", "Syn_l_flag", "Syn_emis_flag", "IC_l_flag", "IC_emis_flag",
             "SSA_l_flag", "gg_flag", "esc_flag")
# keys of LeMoC/Parameters*.txt that may differ between the walkers of one batch
WALKER_KEYS = ("g_min_el", "g_max_el", "g_PL_min", "g_PL_max", "p_el", "L_el", "R0", "B0")
SCRIPT_KEYS = ("time_init", "time_end", "step_alg", "g_min_el", "g_max_el", "g_PL_min", "g_PL_max", "grid_g_el",
               "grid_nu", "p_el", "L_el", "Vexp", "R0", "B0", "m", "delta") + FLAG_KEYS + (
               "BB_flag", "BB_temperature", "GB_ext", "PL_flag", "dE_dV_ph", "nu_min_ph", "nu_max_ph", "s_ph", "User_ph")
CODE_KEYS = tuple(k for k in SCRIPT_KEYS if k not in WALKER_KEYS + ("delta",))


def read_parameter_file(fileName) -> dict:
    """`key = value` per line, every value a float (LeMoC.py:60-65, Code.py:63-68)."""
    params = {}
    with open(fileName) as fileObj:
        for line in fileObj:
            line = line.strip()
            if not line:
                continue
            key_value = line.split("=")
            params[key_value[0].strip()] = float(key_value[1].strip())
    return params


def _flag01(P, key):
    v = float(P[key])
    if v not in (0.0, 1.0):
        raise ValueError(f"{key} = {v}: the drivers compare flags with ==1./==0.; only 0 or 1 is supported")
    return int(v)


@dataclass
class Batch:
    """Packed inputs of lhm_setup_batch for W walkers (all float64, C-contiguous)."""
    variant: str
    cfg: dict                      # fields of lhm_config except ic_path/max_walkers
    consts: np.ndarray             # [W, LHM_NCONST]
    g_el: np.ndarray               # [W, Ng]
    nu_syn: np.ndarray             # [W, Ns]  (appended point included)
    nu_ic: np.ndarray              # [W, Ni]
    nu_tot: np.ndarray             # [W, Nt]
    el_inj: np.ndarray             # [W, Ng]
    ext: np.ndarray                # [W, Nt]
    nsteps: np.ndarray             # [W] int
    index_PL_min: np.ndarray = field(default=None)
    index_PL_max: np.ndarray = field(default=None)
    g_pr: np.ndarray = field(default=None)         # [W, Np]   LeHaMoC.py driver only
    pr_inj: np.ndarray = field(default=None)       # [W, Np]

    def __post_init__(self):
        # zero timesteps: LeMoC.py:196 / LeHaMoC.py loop over range(0) and write no row, Code.py:277 fails on the unbound
        # Spec_temp_tot -- there is no output to reproduce, so the request is refused here rather than answered
        if self.nsteps is not None and len(self.nsteps) and int(np.min(self.nsteps)) < 1:
            raise ValueError("time_init/time_end/step_alg give a walker zero timesteps: the reference produces no output")

    @property
    def W(self):
        return self.consts.shape[0]

    def arrays(self):
        return (self.consts, self.g_el, self.nu_syn, self.nu_ic, self.nu_tot, self.el_inj, self.ext)

    def pin(self):
        """Move the arrays into page-locked memory (torch pinned tensors, kept alive by this object): Engine.setup's H2D
        copies then are asynchronous DMA transfers instead of staged copies out of pageable memory, and broadcast views
        of shared grids are materialised once instead of on every set-up."""
        import torch
        keep = self.__dict__.setdefault("_pins", {})
        for name in ("consts", "g_el", "nu_syn", "nu_ic", "nu_tot", "el_inj", "ext", "g_pr", "pr_inj"):
            a = getattr(self, name)
            if a is None or name in keep:
                continue
            t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = a
            keep[name] = t
            setattr(self, name, t.numpy())
        return self


def _nu_c(gamma, B):                               # LeHaMoC_f.py:78-79
    return 3. / (4. * np.pi) * q * B / (m_el * c) * gamma ** 2.


def _Q_el_Lum(L_el, p_el, g_min_el, g_max_el):     # LeHaMoC_f.py:52-56
    if p_el == 2.:
        return L_el / (m_el * c ** 2. * np.log(g_max_el / g_min_el))
    return L_el * (-p_el + 2.) / (m_el * c ** 2. * (g_max_el ** (-p_el + 2.) - g_min_el ** (-p_el + 2.)))


def _Lum_e_inj(comp, R):                           # LeHaMoC_f.py:66-67
    return 4. * np.pi * R * m_el * c ** 3. * comp / sigmaT


USER_SPECTRUM_FILE = "Photons_spec_user.txt"      # read from the working directory, LeMoC.py:165 / LeHaMoC.py:206
USER_SPECTRUM_KEY = "_user_spectrum"              # optional (log10 nu, log10 dN/dVdnu) pair in a parameter dict instead


def user_photon_field(F: dict, nu_tot: np.ndarray) -> np.ndarray:
    """External user-defined photon field on nu_tot (LeMoC.py:160-169, LeHaMoC.py:201-210): a two-column CSV of
    log10(nu), log10(dN/dV dnu); last ordinate replaced by 1e-160, log-log np.interp with its end clamping."""
    if float(F["User_ph"]) == 0.:
        return np.zeros(len(nu_tot))
    if USER_SPECTRUM_KEY in F:
        logx, logy = (np.asarray(a, dtype=np.float64) for a in F[USER_SPECTRUM_KEY])
    else:
        try:
            tab = np.loadtxt(USER_SPECTRUM_FILE, delimiter=",", ndmin=2)
        except OSError as e:
            raise FileNotFoundError(f"User_ph = {F['User_ph']}: {USER_SPECTRUM_FILE} not found in the working "
                                    "directory (the reference reads it from there)") from e
        logx, logy = tab[:, 0], tab[:, 1]
    nu_user = 10 ** np.array(logx)
    tmp = 10 ** np.array(logy)
    tmp[-1] = 10 ** (-160.)
    with np.errstate(divide="ignore", invalid="ignore"):
        return 10 ** np.interp(np.log10(nu_tot), np.log10(nu_user), np.log10(tmp))


def external_fields(F: dict, nu_tot: np.ndarray) -> np.ndarray:
    """BB + PL + user photon densities on nu_tot as they come out of photons_tot (LeMoC.py:140-169 through
    LeHaMoC_f.py:154-156), for one nu_tot grid.  With BB_flag = 0 the BB term is log10(0) = -inf everywhere
    (np.interp's NaN fall-backs give 0) except the last bin, which hits the appended 1e-260 point."""
    Nt = len(nu_tot)
    if F["BB_flag"] == 0.:
        bb = np.zeros(Nt)
        bb[-1] = SENTINEL
    else:
        T = 10 ** float(F["BB_temperature"])
        nu_bb = np.logspace(np.log10(5.879 * 10 ** 10 * T) - 6., np.log10(5.879 * 10 ** 10 * T) + 1.5, 60)
        with np.errstate(over="ignore"):
            Bnu = 2.0 * h * nu_bb ** 3 / c ** 2 / np.expm1(h * nu_bb / (kb * T))    # astropy BlackBody(T)(nu)
        photons_bb = 4. * np.pi / c * Bnu / (h * nu_bb)
        GB_norm = np.trapz(photons_bb * h * nu_bb ** 2., np.log(nu_bb)) / float(F["GB_ext"])
        dN = np.append(photons_bb / GB_norm, SENTINEL)
        nu_bb = np.append(nu_bb, nu_tot[-1])
        with np.errstate(divide="ignore", invalid="ignore"):
            bb = 10 ** np.interp(np.log10(nu_tot), np.log10(nu_bb), np.log10(dN))
    if F["PL_flag"] == 0.:
        pl = np.zeros(Nt)
    else:                                          # literal LeMoC.py:155-158 (known to yield inf)
        sp = np.logspace(float(F["nu_min_ph"]), float(F["nu_max_ph"]), 100)
        k_ph = (np.trapz(float(F["dE_dV_ph"]) * sp ** (-float(F["s_ph"]) + 1.))) ** (-1.)
        sp[-1] = 0.
        with np.errstate(divide="ignore", invalid="ignore"):
            pl = 10 ** np.interp(np.log10(nu_tot), np.log10(sp), np.log10(k_ph * sp ** (-float(F["s_ph"]))))
    return bb + pl + user_photon_field(F, nu_tot)                 # order of photons_tot, LeHaMoC_f.py:156


def _leptonic_cfg(variant, F, Ng, Nh, Nt, Vexp, shared):
    cfg = dict(Ng=Ng, Ns=Nh + 1, Ni=Nh, Nt=Nt, gg_npts=50,
               loss_minus_one=1 if variant == "script" else 0,       # LeMoC.py:224 vs Code.py:210
               qee_from_ic=0 if variant == "script" else 1,          # LeMoC.py:283 vs Code.py:269
               const_B=int(float(F["m"]) == 0. or Vexp == 0.),      # B = B0*(R0/R)**m never changes (LeMoC.py:199-200)
               ic_pair_table=1, driver=0, Np=0, esc_flag_pr=0, pg_pi_l_flag=0, pg_BH_l_flag=0, pg_pi_emis_flag=0,
               neutrino_flag=0, pg_BH_emis_flag=0, shared_ic_grids=int(shared), shared_had_grids=0)
    for k in FLAG_KEYS:
        cfg[k] = _flag01(F, k)
    # gg_flag is tested with `== 0.` (LeMoC.py:279): anything else switches it on; inj_flag is tested both ways
    return cfg


@dataclass
class PackRequest:
    """What the device-side packer (lhm_pack_batch) needs instead of the seven [W, n] arrays of a Batch: one row of
    scalars per walker and the three grids every walker shares."""
    variant: str
    cfg: dict
    walker_params: np.ndarray      # [W, LHM_PACK_NPAR]
    nu_ic_row: np.ndarray          # [Nh]
    nu_tot_row: np.ndarray         # [Nt]
    ext_row: np.ndarray            # [Nt]
    scalars: dict                  # time_init, time_end, step_alg, Vexp_cm_s, m

    @property
    def W(self):
        return self.walker_params.shape[0]

    # the axes a caller needs next to the spectra, as the host packer would have built them (NumPy)
    @property
    def nu_tot(self):
        return np.broadcast_to(self.nu_tot_row, (self.W, len(self.nu_tot_row)))

    @property
    def nu_ic(self):
        return np.broadcast_to(self.nu_ic_row, (self.W, len(self.nu_ic_row)))

    @property
    def g_el(self):
        from . import _lib
        lo, hi = self.walker_params[:, _lib.PK_GMIN], self.walker_params[:, _lib.PK_GMAX]
        if np.all(lo == lo[0]) and np.all(hi == hi[0]):
            return np.broadcast_to(np.logspace(lo[0], hi[0], self.cfg["Ng"]), (self.W, self.cfg["Ng"]))
        return np.logspace(lo, hi, self.cfg["Ng"], axis=-1)


def _pow10(x):
    """10**x for the device-packed path: one vectorised np.power instead of a Python-float pow per walker (which is what
    the reference's scalars go through and what the host packer keeps).  The two agree to 1 ulp (they differ in ~5 % of
    the arguments); the device-side set-up is specified to that accuracy anyway."""
    return np.power(10., np.asarray(x, dtype=np.float64))


def _logspace_points(lo, hi, n, idx):
    """np.logspace(lo[w], hi[w], n)[idx[w, k]] without building the rows: np.linspace is `arange*step + start` with
    the last point set to `stop`, np.logspace raises 10 to it with np.power -- the same two expressions here, on the
    requested points only (same doubles as the full rows; tests/test_host_and_abi.py checks that)."""
    step = (hi - lo) / (n - 1)
    y = idx * step[:, None] + lo[:, None]
    y = np.where(idx == n - 1, hi[:, None], y)
    return np.power(10., y)


def _window_indices(lo, hi, n, pl_min, pl_max):
    """index_PL_min, index_PL_max of LeMoC.py:120-128 / Code.py:82-90 for every walker, with the reference's own
    comparisons (`g_el > 10**g_PL_max`, `g_el < 10**g_PL_min`; thresholds through Python's float pow) on the NumPy
    grid values -- evaluated only on a few candidates around the arithmetic position of the threshold, O(W) instead
    of O(W*Ng).  Walkers whose candidates do not bracket the threshold fall back to the full row."""
    W = len(lo)
    # both thresholds in one stacked pass (rows 0..W-1: `g > 10**g_PL_max`, rows W..2W-1: `not g < 10**g_PL_min`): the
    # call sits on the latency path of every fit evaluation, and at 24 walkers it is all NumPy call overhead
    thr = np.array([10 ** x for x in pl_max.tolist()] + [10 ** x for x in pl_min.tolist()])
    target = np.concatenate([pl_max, pl_min])
    lo2, hi2 = np.concatenate([lo, lo]), np.concatenate([hi, hi])
    step = (hi2 - lo2) / (n - 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        guess = np.floor((target - lo2) / step)
        guess = np.where(np.isfinite(guess), guess, np.where(guess > 0, n - 1., 0.))     # nan, -inf -> 0; +inf -> n-1
    cand = np.clip(guess[:, None].astype(np.int64) + np.arange(-2, 4)[None, :], 0, n - 1)
    g = _logspace_points(lo2, hi2, n, cand)
    ok = np.empty(cand.shape, dtype=bool)
    ok[:W] = g[:W] > thr[:W, None]
    ok[W:] = ~(g[W:] < thr[W:, None])
    # first index j with the predicate true (it is monotone along the grid); n if none
    out = np.where(ok.any(axis=1), cand[np.arange(2 * W), ok.argmax(axis=1)], n)
    # bracket check: predicate false on the first candidate (or it is grid index 0 and true) and true on the last
    # (or it is the last grid index and false); walkers whose candidates do not bracket fall back to the full row
    good = (~ok[:, 0] | (cand[:, 0] == 0)) & (ok[:, -1] | (cand[:, -1] == n - 1))
    for r in np.flatnonzero(~good):
        row = np.logspace(lo2[r], hi2[r], n)
        row = (row > thr[r]) if r < W else ~(row < thr[r])
        out[r] = int(row.argmax()) if row.any() else n
    first_gt, first_ge = out[:W], out[W:]                                  # first_ge = count of g < thr
    same = pl_max == hi
    if np.any(~same & (first_gt >= n)) or np.any((pl_min != 0.) & (first_ge < 1)):
        raise ValueError("power-law window outside the gamma grid (the reference raises here too)")
    return np.where(pl_min == 0., 1, first_ge - 1), np.where(same, -1, first_gt)


def _window_indices_lib(lo, hi, n, pl_min, pl_max):
    """_window_indices through liblhm's host routine (lhm_window_indices_host: the same candidates and comparisons in C,
    ~3 us instead of ~70 us of NumPy call overhead on the latency path of every fit evaluation).  Walkers with a grid value
    within a few ulp of a threshold -- the only place where libm's pow and NumPy's array pow could tip the comparison
    differently -- are decided by _window_indices, i.e. by the reference's own NumPy expressions."""
    from . import _lib
    W = len(lo)
    ws = _WINDOW_WS.get(W)
    if ws is None:                           # workspace and its addresses per batch size (`.ctypes.data` costs ~2 us a go)
        buf, status = np.empty((6, W)), np.empty(W, dtype=np.int32)
        ws = (buf, status, [buf.ctypes.data + 8 * W * k for k in range(6)] + [status.ctypes.data], _lib.load())
        if len(_WINDOW_WS) > 64:
            _WINDOW_WS.clear()
        _WINDOW_WS[W] = ws
    buf, status, ptrs, lib = ws
    buf[0], buf[1], buf[2], buf[3] = lo, hi, pl_min, pl_max
    _lib.check(lib.lhm_window_indices_host(W, n, *ptrs), "lhm_window_indices_host")
    idx = buf[4:6].copy()
    if status.any():
        if np.any(status == 2):
            raise ValueError("power-law window outside the gamma grid (the reference raises here too)")
        r = np.flatnonzero(status == 1)
        idx[0, r], idx[1, r] = _window_indices(buf[0, r], buf[1, r], n, buf[2, r], buf[3, r])
    return idx[0], idx[1]


_WINDOW_WS = {}


def _pack_request(variant, F, R0, B0, g_min_el, g_max_el, g_PL_min, g_PL_max, p_el, comp_inj, comp_cor) -> PackRequest:
    """Host part of the device-side set-up: the scalars the reference evaluates with Python floats (10**x thresholds)
    and the grids that do not depend on the walker (nu_ic, nu_tot, external photon fields), with the reference's NumPy
    expressions; everything that is [W, n] is left to lhm_pack_batch."""
    from . import _lib
    W = len(R0)
    Ng, Nh, Nt = int(float(F["grid_g_el"])), int(float(F["grid_g_el"]) / 2), int(float(F["grid_nu"]))
    wp = np.zeros((W, _lib.LHM_PACK_NPAR))
    wp[:, _lib.PK_R0], wp[:, _lib.PK_B0] = R0, B0
    wp[:, _lib.PK_GMIN], wp[:, _lib.PK_GMAX] = g_min_el, g_max_el
    wp[:, _lib.PK_GPLMIN], wp[:, _lib.PK_GPLMAX] = g_PL_min, g_PL_max
    wp[:, _lib.PK_P], wp[:, _lib.PK_COMP_INJ], wp[:, _lib.PK_COMP_COR] = p_el, comp_inj, comp_cor
    wp[:, _lib.PK_IDXMIN], wp[:, _lib.PK_IDXMAX] = _window_indices_lib(g_min_el, g_max_el, Ng, g_PL_min, g_PL_max)
    # zero timesteps (LeMoC.py:196 loops over range(0), Code.py:183 never enters its while): the reference produces no
    # output, so the request is refused like Batch.__post_init__ does for the host-packed path
    if variant == "script":
        zero = int(float(F["time_end"])) < 1
    else:
        zero = bool(np.any(~(float(F["time_init"]) < float(F["time_end"]) * R0 / c)))
    if zero:
        raise ValueError("time_init/time_end/step_alg give a walker zero timesteps: the reference produces no output")
    shared = W > 1 and bool(np.all(g_min_el == g_min_el[0]) and np.all(g_max_el == g_max_el[0]))
    # what does not depend on the walkers is a function of the parameter file alone: memoised per file contents (a
    # sampler calls this thousands of times with the same frozen parameters)
    try:
        key = (variant, shared, tuple(F.items()))
        hit = _REQUEST_ROWS.get(key)
    except TypeError:                                   # unhashable entry (a user-supplied photon spectrum array)
        key = hit = None
    if hit is None:
        nu_ic = np.logspace(10., 30., Nh)                                                      # LeMoC.py:132
        nu_tot = np.logspace(np.log10(np.logspace(7.5, 8.5, 2)[0]), np.log10(nu_ic[-1]), Nt)   # LeMoC.py:131,133
        Vexp = float(F["Vexp"]) * c
        cfg = _leptonic_cfg(variant, F, Ng, Nh, Nt, Vexp, shared=shared)
        scal = dict(time_init=float(F["time_init"]), time_end=float(F["time_end"]), step_alg=float(F["step_alg"]),
                    Vexp_cm_s=Vexp, m=float(F["m"]))
        hit = (cfg, nu_ic, nu_tot, external_fields(F, nu_tot), scal)
        for a in hit[1:4]:
            a.flags.writeable = False
        if key is not None:
            if len(_REQUEST_ROWS) > 64:
                _REQUEST_ROWS.clear()
            _REQUEST_ROWS[key] = hit
    cfg, nu_ic, nu_tot, ext, scal = hit
    return PackRequest(variant, dict(cfg), wp, nu_ic, nu_tot, ext, dict(scal))


_REQUEST_ROWS = {}


def _pack(variant, F, R0, B0, g_min_el, g_max_el, g_PL_min, g_PL_max, p_el, comp_inj, comp_cor) -> Batch:
    from ._lib import LHM_NCONST, C_R0, C_B0, C_M, C_VEXP, C_DT, C_TINIT, C_NSTEPS, C_COR_P, C_COR_Q0, C_COR_B
    W = len(R0)
    Ng = int(float(F["grid_g_el"]))
    Nh = int(float(F["grid_g_el"]) / 2)
    Nt = int(float(F["grid_nu"]))
    time_init, time_end, step_alg = float(F["time_init"]), float(F["time_end"]), float(F["step_alg"])
    Vexp = float(F["Vexp"]) * c
    dt = step_alg * R0 / c                                                    # LeMoC.py:108

    def rows_logspace(lo, hi, n):
        """np.logspace(lo[w], hi[w], n) per walker; one evaluation broadcast when all walkers agree (same doubles)."""
        lo, hi = np.broadcast_to(np.asarray(lo, dtype=np.float64), (W,)), np.broadcast_to(np.asarray(hi, dtype=np.float64), (W,))
        if W > 1 and np.all(lo == lo[0]) and np.all(hi == hi[0]):
            return np.broadcast_to(np.logspace(lo[0], hi[0], n), (W, n))
        return np.logspace(lo, hi, n, axis=-1)

    g_el = rows_logspace(g_min_el, g_max_el, Ng)                              # LeMoC.py:115  [W,Ng]
    # PL window (LeMoC.py:120-128); thresholds with Python float pow like the reference's `10**g_PL_max`
    thr_max = np.array([10 ** x for x in g_PL_max.tolist()])
    thr_min = np.array([10 ** x for x in g_PL_min.tolist()])
    gt = g_el > thr_max[:, None]
    lt = g_el < thr_min[:, None]
    same = g_PL_max == g_max_el
    if np.any(~same & ~gt.any(axis=1)) or np.any((g_PL_min != 0.) & ~lt.any(axis=1)):
        raise ValueError("power-law window outside the gamma grid (the reference raises here too)")
    index_PL_max = np.where(same, -1, gt.argmax(axis=1))
    index_PL_min = np.where(g_PL_min == 0., 1, lt.sum(axis=1) - 1)

    nu_syn = rows_logspace(7.5, np.log10(7. * _nu_c(g_el[:, -1], B0)) + 1.4, Nh)         # LeMoC.py:131
    nu_ic = np.logspace(10., 30., Nh)                                                      # LeMoC.py:132
    nu_tot = rows_logspace(np.log10(nu_syn[:, 0]), np.log10(nu_ic[-1]), Nt)                # LeMoC.py:133
    nu_syn = np.concatenate([nu_syn, nu_tot[:, -1:]], axis=1)                              # LeMoC.py:185
    nu_ic = np.broadcast_to(nu_ic, (W, Nh))

    # injection (LeMoC.py:174-175).  The reference evaluates the normalisation with scalars: the two scalar powers
    # and the log go through Python floats (libm), everything else is elementwise IEEE arithmetic in the same order.
    rows = np.arange(W)
    col = np.arange(Ng)[None, :]
    g_lo = g_el[rows, index_PL_min]
    g_hi = g_el[rows, index_PL_max]                                  # -1 -> last point, as g_el[index_PL_max]

    def Q_el_Lum_rows(L_el, g_min, g_max):                           # LeHaMoC_f.py:52-56
        e = -p_el + 2.
        pw_hi = np.array([x ** y for x, y in zip(g_max.tolist(), e.tolist())])
        pw_lo = np.array([x ** y for x, y in zip(g_min.tolist(), e.tolist())])
        with np.errstate(divide="ignore", invalid="ignore"):
            out = L_el * e / (m_el * c ** 2. * (pw_hi - pw_lo))
        for w in np.flatnonzero(p_el == 2.):
            out[w] = L_el[w] / (m_el * c ** 2. * np.log(g_max[w] / g_min[w]))
        return out

    Q0 = Q_el_Lum_rows(_Lum_e_inj(comp_inj, R0), g_lo, g_hi)
    stop = np.where(index_PL_max < 0, Ng + index_PL_max, index_PL_max)
    inside = (col >= index_PL_min[:, None]) & (col < stop[:, None])
    el_inj = np.where(inside, Q0[:, None] * g_el ** (-p_el[:, None]), SENTINEL)
    cor_Q0 = Q_el_Lum_rows(_Lum_e_inj(comp_cor, R0), g_el[:, 1], g_el[:, -2])              # LeHaMoC_f.py:202

    # number of steps: range(int(time_end)) (LeMoC.py:196) / while time_real < time_end*R0/c (Code.py:183)
    if variant == "script":
        nsteps = np.full(W, int(time_end), dtype=np.int64)
    else:
        nsteps = np.zeros(W, dtype=np.int64)
        t = np.full(W, time_init)
        lim = time_end * R0 / c
        live = t < lim
        while live.any():
            t = np.where(live, t + dt, t)
            nsteps += live
            live = t < lim

    # external fields: identical for walkers that share nu_tot (nu_syn[0] = 10**7.5 and nu_ic[-1] = 1e30 always)
    ext1 = external_fields(F, nu_tot[0])
    ext = np.broadcast_to(ext1, (W, Nt))
    if not np.all(nu_tot == nu_tot[0]):
        ext = ext.copy()
        for w in range(1, W):
            ext[w] = external_fields(F, nu_tot[w])

    consts = np.zeros((W, LHM_NCONST))
    consts[:, C_R0] = R0
    consts[:, C_B0] = B0
    consts[:, C_M] = float(F["m"])
    consts[:, C_VEXP] = Vexp
    consts[:, C_DT] = dt
    consts[:, C_TINIT] = time_init
    consts[:, C_NSTEPS] = nsteps
    consts[:, C_COR_P] = p_el
    consts[:, C_COR_Q0] = cor_Q0
    consts[:, C_COR_B] = 10 ** 4.
    # fixed-grid sweeps (every walker has the same g_el / nu_ic / nu_tot): one IC pair table per batch
    cfg = _leptonic_cfg(variant, F, Ng, Nh, Nt, Vexp,
                        shared=W > 1 and bool(np.all(g_el == g_el[0]) and np.all(nu_tot == nu_tot[0])))
    cont = np.ascontiguousarray
    return Batch(variant, cfg, cont(consts), cont(g_el), cont(nu_syn), cont(nu_ic), cont(nu_tot), cont(el_inj),
                 cont(ext), nsteps, index_PL_min, index_PL_max)


LH_KEYS = ("time_init", "time_end", "step_alg", "PL_inj", "g_min_el", "g_max_el", "g_el_PL_min", "g_el_PL_max",
           "grid_g_el", "g_min_pr", "g_max_pr", "g_pr_PL_min", "g_pr_PL_max", "grid_g_pr", "grid_nu", "p_el", "L_el",
           "p_pr", "L_pr", "Vexp", "R0", "B0", "m", "delta", "inj_flag", "Ad_l_flag", "Syn_l_flag", "Syn_emis_flag",
           "IC_l_flag", "IC_emis_flag", "SSA_l_flag", "gg_flag", "pg_pi_l_flag", "pg_pi_emis_flag", "pg_BH_l_flag",
           "pg_BH_emis_flag", "n_H", "pp_l_flag", "pp_ee_emis_flag", "pp_g_emis_flag", "pp_nu_emis_flag",
           "neutrino_flag", "esc_flag_el", "esc_flag_pr", "BB_flag", "temperature", "GB_ext", "PL_flag", "dE_dV_ph",
           "nu_min_ph", "nu_max_ph", "s_ph", "User_ph")
PHOTOHADRONIC_FLAGS = ("pg_pi_l_flag", "pg_BH_l_flag", "pg_pi_emis_flag", "neutrino_flag", "pg_BH_emis_flag")
PP_FLAGS = ("pp_l_flag", "pp_ee_emis_flag", "pp_g_emis_flag", "pp_nu_emis_flag")
m_pr_ = 1.67262192369e-24


def _Q_norm_exp_c(R0, comp, g, g_max_ind, p_i):          # LeHaMoC_f.py:81-82
    return 4. * np.pi * R0 * c * comp / (sigmaT * np.trapz(g[:g_max_ind] ** (-p_i + 2.) * np.exp(-g[:g_max_ind] / g[g_max_ind]),
                                                            np.log(g[:g_max_ind])))


def _Q_pr_Lum(L_pr, p_pr, g_min_pr, g_max_pr):           # LeHaMoC_f.py:75-79
    if p_pr == 2.:
        return L_pr / (m_pr_ * c ** 2. * np.log(g_max_pr / g_min_pr))
    return L_pr * (-p_pr + 2.) / (m_pr_ * c ** 2. * (g_max_pr ** (-p_pr + 2.) - g_min_pr ** (-p_pr + 2.)))


def _pack_lehamoc_loop(param_sets) -> Batch:
    """pack_lehamoc one walker at a time with the reference's own expressions (LeHaMoC.py:132-259): the statement the
    vectorised packer below is tested against (~50 us per walker)."""
    from ._lib import (LHM_NCONST, C_R0, C_B0, C_M, C_VEXP, C_DT, C_TINIT, C_NSTEPS, C_COR_P, C_COR_Q0, C_COR_B, C_NH,
                       C_PPR)
    if isinstance(param_sets, dict):
        param_sets = [param_sets]
    F = param_sets[0]
    missing = [k for k in LH_KEYS if k not in F]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    per_walker = ("g_min_el", "g_max_el", "g_el_PL_min", "g_el_PL_max", "g_min_pr", "g_max_pr", "g_pr_PL_min",
                  "g_pr_PL_max", "p_el", "L_el", "p_pr", "L_pr", "R0", "B0", "n_H")
    for P in param_sets:
        for k in LH_KEYS:
            if k not in per_walker and float(P[k]) != float(F[k]):
                raise ValueError(f"walkers of one batch must share '{k}'")
    W = len(param_sets)
    Ng, Nh, Np, Nt = int(float(F["grid_g_el"])), int(float(F["grid_g_el"]) / 2), int(float(F["grid_g_pr"])), int(float(F["grid_nu"]))
    time_init, time_end, step_alg = float(F["time_init"]), float(F["time_end"]), float(F["step_alg"])
    Vexp = float(F["Vexp"]) * c
    g_el = np.empty((W, Ng)); g_prs = np.empty((W, Np)); nu_syn = np.empty((W, Nh + 1)); nu_ic = np.empty((W, Nh))
    nu_tot = np.empty((W, Nt)); el_inj = np.empty((W, Ng)); pr_inj = np.empty((W, Np)); ext = np.empty((W, Nt))
    consts = np.zeros((W, LHM_NCONST))
    for w, P in enumerate(param_sets):
        f = lambda k: float(P[k])
        R0 = 10 ** f("R0")
        B0 = f("B0")
        comp_el = sigmaT * 10 ** f("L_el") / (4. * np.pi * R0 * m_el * c ** 3)                 # LeHaMoC.py:136
        comp_pr = sigmaT * 10 ** f("L_pr") / (4. * np.pi * R0 * m_pr_ * c ** 3)
        ge = np.logspace(f("g_min_el"), f("g_max_el"), Ng)
        gp = np.logspace(f("g_min_pr"), f("g_max_pr"), Np)
        ie_max = -1 if f("g_el_PL_max") == f("g_max_el") else int(np.min(np.where(ge > 10 ** f("g_el_PL_max"))[0]))
        ie_min = 1 if f("g_el_PL_min") == 0. else int(np.max(np.where(ge < 10 ** f("g_el_PL_min"))[0]))
        ip_max = -1 if f("g_pr_PL_max") == f("g_max_pr") else int(np.min(np.where(gp > 10 ** f("g_pr_PL_max"))[0]) + 1)
        ip_min = 1 if f("g_pr_PL_min") == 0. else int(np.max(np.where(gp < 10 ** f("g_pr_PL_min"))[0]) + 1)
        ns = np.logspace(7.5, np.log10(7. * _nu_c(ge[-1], B0)) + 1.4, Nh)                      # LeHaMoC.py:171
        ni = np.logspace(15., 32., Nh)
        nt = np.logspace(np.log10(ns[0]), np.log10(max(ni[-1], ns[-1])), Nt)
        V0 = 4. * np.pi / 3. * R0 ** 3.
        ei = np.ones(Ng) * SENTINEL
        pi_ = np.ones(Np) * SENTINEL
        if f("PL_inj") == 1.:                                                                  # LeHaMoC.py:256-261
            ei[ie_min:ie_max] = _Q_el_Lum(_Lum_e_inj(comp_el, R0), f("p_el"), ge[ie_min], ge[ie_max]) \
                * ge[ie_min:ie_max] ** (-f("p_el")) / V0
            Lp = 4. * np.pi * R0 * m_pr_ * c ** 3. * comp_pr / sigmaT
            pi_[ip_min:ip_max] = _Q_pr_Lum(Lp, f("p_pr"), gp[ip_min], gp[ip_max]) * gp[ip_min:ip_max] ** (-f("p_pr")) / V0
        else:
            pi_ = _Q_norm_exp_c(R0, comp_pr, gp, ip_max, f("p_pr")) * gp ** (-f("p_pr")) * np.exp(-gp / gp[ip_max]) / V0
            ei = _Q_norm_exp_c(R0, comp_el, ge, ie_max, f("p_el")) * ge ** (-f("p_el")) * np.exp(-ge / ge[ie_max]) / V0
        g_el[w], g_prs[w], nu_ic[w], nu_tot[w], el_inj[w], pr_inj[w] = ge, gp, ni, nt, ei, pi_
        nu_syn[w] = np.append(ns, nt[-1])                                                     # LeHaMoC.py:234
        ext[w] = external_fields_lh(P, nt)
        dt = step_alg * R0 / c
        consts[w, [C_R0, C_B0, C_M, C_VEXP, C_DT, C_TINIT, C_NSTEPS]] = [R0, B0, float(F["m"]), Vexp, dt, time_init,
                                                                         int(time_end / step_alg)]
        consts[w, C_COR_P] = f("p_el")
        consts[w, C_COR_Q0] = _Q_el_Lum(_Lum_e_inj(comp_el, R0), f("p_el"), ge[1], ge[-2])     # LeHaMoC_f.py:646
        consts[w, C_COR_B] = 10 ** 4.
        consts[w, C_NH], consts[w, C_PPR] = f("n_H"), f("p_pr")                                # LeHaMoC.py:95,114
    cfg = dict(Ng=Ng, Ns=Nh + 1, Ni=Nh, Nt=Nt, gg_npts=100, loss_minus_one=1, qee_from_ic=0,
               const_B=int(float(F["m"]) == 0. or Vexp == 0.), ic_pair_table=1,
               shared_ic_grids=int(W > 1 and bool(np.all(g_el == g_el[0]) and np.all(nu_tot == nu_tot[0]))),
               # one set of photo-hadronic operator tables for the batch: g_el, g_pr, nu_ic and nu_tot shared
               shared_had_grids=int(bool(np.all(g_el == g_el[0]) and np.all(nu_tot == nu_tot[0])
                                         and np.all(g_prs == g_prs[0]) and np.all(nu_ic == nu_ic[0]))),
               driver=1, Np=Np, esc_flag_pr=_flag01(F, "esc_flag_pr"))
    for k in PHOTOHADRONIC_FLAGS + PP_FLAGS:
        cfg[k] = _flag01(F, k)
    if cfg["neutrino_flag"] and not cfg["pg_pi_emis_flag"]:
        cfg["neutrino_flag"] = 0                 # LeHaMoC.py:346 only looks at it inside the pg_pi_emis branch
    for k in FLAG_KEYS:
        cfg[k] = _flag01(F, "esc_flag_el" if k == "esc_flag" else k)
    nsteps = np.full(W, int(time_end / step_alg), dtype=np.int64)
    b = Batch("lehamoc", cfg, consts, g_el, nu_syn, nu_ic, nu_tot, el_inj, ext, nsteps)
    b.g_pr, b.pr_inj = g_prs, pr_inj
    return b


def pack_lehamoc(param_sets) -> Batch:
    """Batch for the `LeHaMoC Code and Tests/LeHaMoC.py` driver (one dict or a list of dicts with the 53 keys of its
    Parameters.txt).  Set-up with the reference's own expressions, LeHaMoC.py:132-259, for all walkers at once (scalar
    powers through Python floats like the reference's scalars, elementwise NumPy arithmetic in the same order: the
    doubles are those of `_pack_lehamoc_loop`).  Supported: leptonic processes, proton equation with synchrotron /
    photopion / Bethe-Heitler losses, proton synchrotron, photopion pairs, photons and neutrinos, Bethe-Heitler pair
    injection, pp loss term and pp secondaries, user photon field (n_H and p_pr may differ per walker)."""
    from ._lib import (LHM_NCONST, C_R0, C_B0, C_M, C_VEXP, C_DT, C_TINIT, C_NSTEPS, C_COR_P, C_COR_Q0, C_COR_B, C_NH,
                       C_PPR)
    if isinstance(param_sets, dict):
        param_sets = [param_sets]
    F = param_sets[0]
    missing = [k for k in LH_KEYS if k not in F]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    per_walker = ("g_min_el", "g_max_el", "g_el_PL_min", "g_el_PL_max", "g_min_pr", "g_max_pr", "g_pr_PL_min",
                  "g_pr_PL_max", "p_el", "L_el", "p_pr", "L_pr", "R0", "B0", "n_H")
    shared_keys = [k for k in LH_KEYS if k not in per_walker]
    get_shared, get_walker = itemgetter(*shared_keys), itemgetter(*per_walker)
    ref = tuple(float(v) for v in get_shared(F))
    for P in param_sets[1:]:
        if get_shared(P) != ref and tuple(float(v) for v in get_shared(P)) != ref:
            bad = [k for k in shared_keys if float(P[k]) != float(F[k])]
            raise ValueError(f"walkers of one batch must share '{bad[0]}'")
    table = np.array([get_walker(P) for P in param_sets], dtype=np.float64)
    col = lambda k: np.ascontiguousarray(table[:, per_walker.index(k)])
    pw10 = lambda a: np.array([10 ** x for x in a.tolist()])
    W = len(param_sets)
    Ng, Nh, Np, Nt = int(float(F["grid_g_el"])), int(float(F["grid_g_el"]) / 2), int(float(F["grid_g_pr"])), int(float(F["grid_nu"]))
    time_init, time_end, step_alg = float(F["time_init"]), float(F["time_end"]), float(F["step_alg"])
    Vexp = float(F["Vexp"]) * c
    R0, B0, p_el, p_pr = pw10(col("R0")), col("B0"), col("p_el"), col("p_pr")
    comp_el = sigmaT * pw10(col("L_el")) / (4. * np.pi * R0 * m_el * c ** 3)                     # LeHaMoC.py:136
    comp_pr = sigmaT * pw10(col("L_pr")) / (4. * np.pi * R0 * m_pr_ * c ** 3)

    def rows_logspace(lo, hi, n):
        lo, hi = np.broadcast_to(np.asarray(lo, dtype=np.float64), (W,)), np.broadcast_to(np.asarray(hi, dtype=np.float64), (W,))
        if W > 1 and np.all(lo == lo[0]) and np.all(hi == hi[0]):
            return np.broadcast_to(np.logspace(lo[0], hi[0], n), (W, n))
        return np.logspace(lo, hi, n, axis=-1)

    def window(g, lo_t, hi_t, gmax_t, shift):
        """index_PL_min / index_PL_max of LeHaMoC.py:150-169 (the proton ones are shifted by one)."""
        gt, lt = g > pw10(hi_t)[:, None], g < pw10(lo_t)[:, None]
        same = hi_t == gmax_t
        if np.any(~same & ~gt.any(axis=1)) or np.any((lo_t != 0.) & ~lt.any(axis=1)):
            raise ValueError("power-law window outside the gamma grid (the reference raises here too)")
        i_max = np.where(same, -1, gt.argmax(axis=1) + shift)
        i_min = np.where(lo_t == 0., 1, lt.sum(axis=1) - 1 + shift)
        return i_min, i_max

    g_el = rows_logspace(col("g_min_el"), col("g_max_el"), Ng)
    g_prs = rows_logspace(col("g_min_pr"), col("g_max_pr"), Np)
    ie_min, ie_max = window(g_el, col("g_el_PL_min"), col("g_el_PL_max"), col("g_max_el"), 0)
    ip_min, ip_max = window(g_prs, col("g_pr_PL_min"), col("g_pr_PL_max"), col("g_max_pr"), 1)
    ns = rows_logspace(7.5, np.log10(7. * _nu_c(g_el[:, -1], B0)) + 1.4, Nh)                   # LeHaMoC.py:171
    ni1 = np.logspace(15., 32., Nh)
    nu_tot = rows_logspace(np.log10(ns[:, 0]), np.log10(np.maximum(ni1[-1], ns[:, -1])), Nt)
    nu_syn = np.concatenate([ns, nu_tot[:, -1:]], axis=1)                                       # LeHaMoC.py:234
    nu_ic = np.broadcast_to(ni1, (W, Nh))
    V0 = np.array([4. * np.pi / 3. * x ** 3. for x in R0.tolist()])       # scalar pow, like the reference's Volume(R0)
    rows = np.arange(W)

    def Q_Lum_rows(L, p, g_lo, g_hi, mass):                                                     # LeHaMoC_f.py:52-56,75-79
        e = -p + 2.
        pw_hi = np.array([x ** y for x, y in zip(g_hi.tolist(), e.tolist())])
        pw_lo = np.array([x ** y for x, y in zip(g_lo.tolist(), e.tolist())])
        with np.errstate(divide="ignore", invalid="ignore"):
            out = L * e / (mass * c ** 2. * (pw_hi - pw_lo))
        for w in np.flatnonzero(p == 2.):
            out[w] = L[w] / (mass * c ** 2. * np.log(g_hi[w] / g_lo[w]))
        return out

    def inside(n, i_min, i_max):
        stop = np.where(i_max < 0, n + i_max, i_max)
        k = np.arange(n)[None, :]
        return (k >= i_min[:, None]) & (k < stop[:, None])

    if float(F["PL_inj"]) == 1.:                                                                # LeHaMoC.py:256-261
        Qe = Q_Lum_rows(_Lum_e_inj(comp_el, R0), p_el, g_el[rows, ie_min], g_el[rows, ie_max], m_el)
        el_inj = np.where(inside(Ng, ie_min, ie_max), Qe[:, None] * g_el ** (-p_el[:, None]) / V0[:, None], SENTINEL)
        Lp = 4. * np.pi * R0 * m_pr_ * c ** 3. * comp_pr / sigmaT
        Qp = Q_Lum_rows(Lp, p_pr, g_prs[rows, ip_min], g_prs[rows, ip_max], m_pr_)
        pr_inj = np.where(inside(Np, ip_min, ip_max), Qp[:, None] * g_prs ** (-p_pr[:, None]) / V0[:, None], SENTINEL)
    else:
        # the normalisation integrates over g[:index_PL_max] (np.trapz, pairwise summation: length-dependent rounding),
        # so that scalar stays a per-walker call; the arrays are built at once
        Np_n = np.array([_Q_norm_exp_c(R0[w], comp_pr[w], g_prs[w], int(ip_max[w]), p_pr[w]) for w in range(W)])
        Ne_n = np.array([_Q_norm_exp_c(R0[w], comp_el[w], g_el[w], int(ie_max[w]), p_el[w]) for w in range(W)])
        pr_inj = Np_n[:, None] * g_prs ** (-p_pr[:, None]) * np.exp(-g_prs / g_prs[rows, ip_max][:, None]) / V0[:, None]
        el_inj = Ne_n[:, None] * g_el ** (-p_el[:, None]) * np.exp(-g_el / g_el[rows, ie_max][:, None]) / V0[:, None]
    ext1 = external_fields_lh(F, nu_tot[0])
    ext = np.broadcast_to(ext1, (W, Nt))
    if not np.all(nu_tot == nu_tot[0]):
        ext = ext.copy()
        for w in range(1, W):
            ext[w] = external_fields_lh(F, nu_tot[w])
    consts = np.zeros((W, LHM_NCONST))
    consts[:, C_R0], consts[:, C_B0], consts[:, C_M], consts[:, C_VEXP] = R0, B0, float(F["m"]), Vexp
    consts[:, C_DT] = step_alg * R0 / c
    consts[:, C_TINIT], consts[:, C_NSTEPS] = time_init, int(time_end / step_alg)
    consts[:, C_COR_P] = p_el
    consts[:, C_COR_Q0] = Q_Lum_rows(_Lum_e_inj(comp_el, R0), p_el, g_el[:, 1], g_el[:, -2], m_el)       # LeHaMoC_f.py:646
    consts[:, C_COR_B] = 10 ** 4.
    consts[:, C_NH], consts[:, C_PPR] = col("n_H"), p_pr                                        # LeHaMoC.py:95,114
    cfg = dict(Ng=Ng, Ns=Nh + 1, Ni=Nh, Nt=Nt, gg_npts=100, loss_minus_one=1, qee_from_ic=0,
               const_B=int(float(F["m"]) == 0. or Vexp == 0.), ic_pair_table=1,
               shared_ic_grids=int(W > 1 and bool(np.all(g_el == g_el[0]) and np.all(nu_tot == nu_tot[0]))),
               # one set of photo-hadronic operator tables for the batch: g_el, g_pr, nu_ic and nu_tot shared
               shared_had_grids=int(bool(np.all(g_el == g_el[0]) and np.all(nu_tot == nu_tot[0])
                                         and np.all(g_prs == g_prs[0]) and np.all(nu_ic == nu_ic[0]))),
               driver=1, Np=Np, esc_flag_pr=_flag01(F, "esc_flag_pr"))
    for k in PHOTOHADRONIC_FLAGS + PP_FLAGS:
        cfg[k] = _flag01(F, k)
    if cfg["neutrino_flag"] and not cfg["pg_pi_emis_flag"]:
        cfg["neutrino_flag"] = 0                 # LeHaMoC.py:346 only looks at it inside the pg_pi_emis branch
    for k in FLAG_KEYS:
        cfg[k] = _flag01(F, "esc_flag_el" if k == "esc_flag" else k)
    cont = np.ascontiguousarray
    nsteps = np.full(W, int(time_end / step_alg), dtype=np.int64)
    b = Batch("lehamoc", cfg, consts, cont(g_el), cont(nu_syn), cont(nu_ic), cont(nu_tot), cont(el_inj), cont(ext), nsteps)
    b.g_pr, b.pr_inj = cont(g_prs), cont(pr_inj)
    return b


def external_fields_lh(F: dict, nu_tot: np.ndarray) -> np.ndarray:
    """BB + PL photon densities on nu_tot for the LeHaMoC.py driver (LeHaMoC.py:177-212 through photons_tot)."""
    Nt = len(nu_tot)
    if float(F["BB_flag"]) == 0.:
        bb = np.zeros(Nt)
        bb[-1] = SENTINEL
    else:
        T = 10 ** float(F["temperature"])
        nu_bb = np.logspace(np.log10(5.879 * 10 ** 10 * T) - 6., np.log10(5.879 * 10 ** 10 * T) + 1.5, 60)
        with np.errstate(over="ignore"):
            Bnu = 2.0 * h * nu_bb ** 3 / c ** 2 / np.expm1(h * nu_bb / (kb * T))
        photons_bb = 4. * np.pi / c * Bnu / (h * nu_bb)
        GB_norm = np.trapz(photons_bb * h * nu_bb ** 2., np.log(nu_bb)) / float(F["GB_ext"])
        dN = np.append(photons_bb / GB_norm, SENTINEL)
        nu_bb = np.append(nu_bb, nu_tot[-1])
        with np.errstate(divide="ignore", invalid="ignore"):
            bb = 10 ** np.interp(np.log10(nu_tot), np.log10(nu_bb), np.log10(dN))
    if float(F["PL_flag"]) == 0.:
        pl = np.zeros(Nt)
    else:                                                    # LeHaMoC.py:195-199
        sp = np.logspace(float(F["nu_min_ph"]), float(F["nu_max_ph"]), 100)
        k_ph = float(F["dE_dV_ph"]) / np.trapz(sp ** (-float(F["s_ph"]) + 1.), sp) / h
        tmp = k_ph * sp ** (-float(F["s_ph"]))
        tmp[-1] = SENTINEL
        pl = 10 ** np.interp(np.log10(nu_tot), np.log10(sp), np.log10(tmp))
    return bb + pl + user_photon_field(F, nu_tot)


def pack_script(param_sets) -> Batch:
    """Batch for the LeMoC.py driver.  `param_sets`: one dict (or a list of dicts) with the keys of
    LeMoC/Parameters*.txt; all walkers must agree on everything except WALKER_KEYS."""
    if isinstance(param_sets, dict):
        param_sets = [param_sets]
    F = param_sets[0]
    missing = [k for k in SCRIPT_KEYS if k not in F]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")           # the reference raises KeyError likewise
    shared_keys = [k for k in SCRIPT_KEYS if k not in WALKER_KEYS]
    get_shared, get_walker = itemgetter(*shared_keys), itemgetter(*WALKER_KEYS)
    ref = tuple(float(v) for v in get_shared(F))
    for P in param_sets[1:]:
        if get_shared(P) != ref and tuple(float(v) for v in get_shared(P)) != ref:
            bad = [k for k in shared_keys if float(P[k]) != float(F[k])]
            raise ValueError(f"walkers of one batch must share '{bad[0]}'")
    table = np.array([get_walker(P) for P in param_sets], dtype=np.float64)           # [W, len(WALKER_KEYS)]
    col = lambda k: np.ascontiguousarray(table[:, WALKER_KEYS.index(k)])
    R0 = np.array([10 ** x for x in col("R0").tolist()])                      # LeMoC.py:79
    B0 = col("B0")
    L_el = np.array([10 ** x for x in col("L_el").tolist()])                  # LeMoC.py:76
    comp_el = sigmaT * L_el / (4. * np.pi * R0 * m_el * c ** 3)               # LeMoC.py:110
    return _pack("script", F, R0, B0, col("g_min_el"), col("g_max_el"), col("g_PL_min"), col("g_PL_max"),
                 col("p_el"), comp_el, comp_el)


def pack_code(thetas, frozen: dict) -> Batch:
    """Batch for Fit_emcee/Code.py::LeMoC.  thetas [W,6] = [log R0, log B0, log g_min, log g_max, log l_e, p]."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    missing = [k for k in CODE_KEYS if k not in frozen]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    R0 = np.array([10 ** x for x in thetas[:, 0].tolist()])                    # Code.py:53
    B0 = np.array([10 ** x for x in thetas[:, 1].tolist()])                    # Code.py:54
    g_PL_min, g_PL_max = thetas[:, 2].copy(), thetas[:, 3].copy()
    comp_inj = np.array([10 ** x for x in thetas[:, 4].tolist()])              # Code.py:163
    comp_cor = thetas[:, 4].copy()                                             # Code.py:237 passes the log value
    return _pack("code", frozen, R0, B0, g_PL_min - 1., g_PL_max + 1., g_PL_min, g_PL_max, thetas[:, 5].copy(),
                 comp_inj, comp_cor)


def pack_code_request(thetas, frozen: dict) -> PackRequest:
    """Device-packed twin of pack_code (same arguments)."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    missing = [k for k in CODE_KEYS if k not in frozen]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    R0 = _pow10(thetas[:, 0])
    B0 = _pow10(thetas[:, 1])
    g_PL_min, g_PL_max = thetas[:, 2].copy(), thetas[:, 3].copy()
    comp_inj = _pow10(thetas[:, 4])
    return _pack_request("code", frozen, R0, B0, g_PL_min - 1., g_PL_max + 1., g_PL_min, g_PL_max, thetas[:, 5].copy(),
                         comp_inj, thetas[:, 4].copy())


def pack_code_lehamoc_request(thetas, frozen: dict) -> PackRequest:
    """Device-packed twin of pack_code_lehamoc (same arguments): the proton grid and injection are built by
    lhm_pack_batch as well."""
    from . import _lib
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    if thetas.shape[1] != 10:
        raise ValueError("Code.LeHaMoC takes 10 free parameters (Code.py:285-295)")
    missing = [k for k in CODE_KEYS + ("grid_g_pr",) if k not in frozen and k != "Ad_l_flag"]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    F = dict(frozen, Ad_l_flag=0.)
    R0 = _pow10(thetas[:, 0])
    B0 = _pow10(thetas[:, 1])
    g_PL_min, g_PL_max = thetas[:, 2].copy(), thetas[:, 3].copy()
    comp_inj = _pow10(thetas[:, 4])
    req = _pack_request("code", F, R0, B0, g_PL_min - 0.5, g_PL_max + 0.5, g_PL_min, g_PL_max, thetas[:, 5].copy(),
                        comp_inj, thetas[:, 4].copy())
    lo_t, hi_t = thetas[:, 6], thetas[:, 7]
    gmin, gmax = lo_t - 0.1, hi_t + 0.5
    if np.any((hi_t != gmax) & ~(hi_t < gmax)) or np.any((lo_t != 0.) & ~(lo_t > gmin)):
        raise ValueError("proton power-law window outside the gamma grid (the reference raises here too)")
    wp = req.walker_params
    wp[:, _lib.PK_GMIN_PR], wp[:, _lib.PK_GMAX_PR] = gmin, gmax
    wp[:, _lib.PK_GPLMIN_PR], wp[:, _lib.PK_GPLMAX_PR] = lo_t, hi_t
    wp[:, _lib.PK_IDXMIN_PR], wp[:, _lib.PK_IDXMAX_PR] = _window_indices_lib(gmin, gmax, int(float(F["grid_g_pr"])), lo_t, hi_t)
    wp[:, _lib.PK_P_PR] = thetas[:, 9]
    wp[:, _lib.PK_COMP_PR] = _pow10(thetas[:, 8])
    req.cfg.update(driver=2, Np=int(float(F["grid_g_pr"])), esc_flag_pr=req.cfg["esc_flag"], shared_ic_grids=0, shared_had_grids=0)
    req.variant = "code_lh"
    return req


def pack_script_request(param_sets) -> PackRequest:
    """Device-packed twin of pack_script (same arguments)."""
    if isinstance(param_sets, dict):
        param_sets = [param_sets]
    F = param_sets[0]
    missing = [k for k in SCRIPT_KEYS if k not in F]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    shared_keys = [k for k in SCRIPT_KEYS if k not in WALKER_KEYS]
    get_shared, get_walker = itemgetter(*shared_keys), itemgetter(*WALKER_KEYS)
    ref = get_shared(F)
    # the walkers of a batch differ in WALKER_KEYS only: one C-level pass over the dicts; sets whose shared values are
    # equal but not identical objects (1 vs 1.0) take the slow comparison
    if not all(map(ref.__eq__, map(get_shared, param_sets))):
        reff = tuple(float(v) for v in ref)
        for P in param_sets[1:]:
            if tuple(float(v) for v in get_shared(P)) != reff:
                bad = [k for k in shared_keys if float(P[k]) != float(F[k])]
                raise ValueError(f"walkers of one batch must share '{bad[0]}'")
    table = np.array(list(map(get_walker, param_sets)), dtype=np.float64)
    col = lambda k: np.ascontiguousarray(table[:, WALKER_KEYS.index(k)])
    R0 = _pow10(col("R0"))
    L_el = _pow10(col("L_el"))
    comp_el = sigmaT * L_el / (4. * np.pi * R0 * m_el * c ** 3)
    return _pack_request("script", F, R0, col("B0"), col("g_min_el"), col("g_max_el"), col("g_PL_min"),
                         col("g_PL_max"), col("p_el"), comp_el, comp_el)


def _protons_code_lh_loop(thetas, R0, Np):
    """Proton grid and injection of Code.py::LeHaMoC (Code.py:299-300,325-347,419-422), one walker at a time with the
    reference's own expressions.  Kept as the statement the vectorised form below is tested against."""
    W = thetas.shape[0]
    g_pr = np.empty((W, Np))
    pr_inj = np.full((W, Np), SENTINEL)
    for w in range(W):
        gmin_pr, gmax_pr = float(thetas[w, 6]) - 0.1, float(thetas[w, 7]) + 0.5
        gp = np.logspace(gmin_pr, gmax_pr, Np)
        ip_max = -1 if float(thetas[w, 7]) == gmax_pr else int(np.min(np.where(gp > 10 ** float(thetas[w, 7]))[0]))
        ip_min = 1 if float(thetas[w, 6]) == 0. else int(np.max(np.where(gp < 10 ** float(thetas[w, 6]))[0]))
        Lp = 4. * np.pi * R0[w] * m_pr_ * c ** 3. * (10 ** float(thetas[w, 8])) / sigmaT
        pr_inj[w, ip_min:ip_max] = _Q_pr_Lum(Lp, float(thetas[w, 9]), gp[ip_min], gp[ip_max]) * gp[ip_min:ip_max] ** (-float(thetas[w, 9]))
        g_pr[w] = gp
    return g_pr, pr_inj


def _protons_code_lh(thetas, R0, Np):
    """The same for all walkers at once (bit-identical: scalar powers through Python floats like the reference's
    scalars, elementwise NumPy arithmetic in the same order)."""
    W = thetas.shape[0]
    lo_t, hi_t, lcomp, p_pr = thetas[:, 6], thetas[:, 7], thetas[:, 8], thetas[:, 9]
    gmin, gmax = lo_t - 0.1, hi_t + 0.5
    g_pr = np.logspace(gmin, gmax, Np, axis=-1)                                   # [W, Np]
    thr_hi = np.array([10 ** x for x in hi_t.tolist()])
    thr_lo = np.array([10 ** x for x in lo_t.tolist()])
    gt, lt = g_pr > thr_hi[:, None], g_pr < thr_lo[:, None]
    same = hi_t == gmax
    if np.any(~same & ~gt.any(axis=1)) or np.any((lo_t != 0.) & ~lt.any(axis=1)):
        raise ValueError("proton power-law window outside the gamma grid (the reference raises here too)")
    ip_max = np.where(same, -1, gt.argmax(axis=1))
    ip_min = np.where(lo_t == 0., 1, lt.sum(axis=1) - 1)
    rows = np.arange(W)
    g_lo, g_hi = g_pr[rows, ip_min], g_pr[rows, ip_max]
    comp = np.array([10 ** x for x in lcomp.tolist()])
    Lp = 4. * np.pi * R0 * m_pr_ * c ** 3. * comp / sigmaT
    e = -p_pr + 2.
    pw_hi = np.array([x ** y for x, y in zip(g_hi.tolist(), e.tolist())])
    pw_lo = np.array([x ** y for x, y in zip(g_lo.tolist(), e.tolist())])
    with np.errstate(divide="ignore", invalid="ignore"):
        Q0 = Lp * e / (m_pr_ * c ** 2. * (pw_hi - pw_lo))                           # LeHaMoC_f.py:75-79
    for w in np.flatnonzero(p_pr == 2.):
        Q0[w] = Lp[w] / (m_pr_ * c ** 2. * np.log(g_hi[w] / g_lo[w]))
    stop = np.where(ip_max < 0, Np + ip_max, ip_max)
    col = np.arange(Np)[None, :]
    inside = (col >= ip_min[:, None]) & (col < stop[:, None])
    pr_inj = np.where(inside, Q0[:, None] * g_pr ** (-p_pr[:, None]), SENTINEL)
    return g_pr, pr_inj


def pack_code_lehamoc(thetas, frozen: dict) -> Batch:
    """Batch for Fit_emcee/Code.py::LeHaMoC (Code.py:280-535, leptonic + proton synchrotron).  thetas [W,10] =
    [log R0, log B0, log g_min, log g_max, log l_e, p_e, log g_min_pr, log g_max_pr, log l_p, p_p]."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    if thetas.shape[1] != 10:
        raise ValueError("Code.LeHaMoC takes 10 free parameters (Code.py:285-295)")
    missing = [k for k in CODE_KEYS + ("grid_g_pr",) if k not in frozen and k != "Ad_l_flag"]
    if missing:
        raise KeyError(f"parameter file lacks {missing}")
    F = dict(frozen, Ad_l_flag=0.)                                              # Code.py:281: no adiabatic losses
    R0 = np.array([10 ** x for x in thetas[:, 0].tolist()])
    B0 = np.array([10 ** x for x in thetas[:, 1].tolist()])
    g_PL_min, g_PL_max = thetas[:, 2].copy(), thetas[:, 3].copy()
    comp_inj = np.array([10 ** x for x in thetas[:, 4].tolist()])
    b = _pack("code", F, R0, B0, g_PL_min - 0.5, g_PL_max + 0.5, g_PL_min, g_PL_max, thetas[:, 5].copy(),
              comp_inj, thetas[:, 4].copy())                                    # Code.py:297-298
    Np = int(float(F["grid_g_pr"]))
    g_pr, pr_inj = _protons_code_lh(thetas, R0, Np)
    b.cfg.update(driver=2, Np=Np, esc_flag_pr=b.cfg["esc_flag"], shared_ic_grids=0, shared_had_grids=0)
    b.variant = "code_lh"
    b.g_pr, b.pr_inj = g_pr, pr_inj
    return b
