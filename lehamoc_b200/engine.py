"""Thin Python owner of one lhm_handle: PyTorch supplies device memory and the CUDA stream, liblhm.so does
the work.  One Engine per (grid sizes, flags, variant) per GPU; not thread-safe (one process per GPU)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import LhmError, lhm_config, check
from .setup_host import Batch

IC_PATHS = {"R": _lib.LHM_IC_PATH_R, "S": _lib.LHM_IC_PATH_S, "G": _lib.LHM_IC_PATH_G}


def _require_cuda():
    if not torch.cuda.is_available():
        raise LhmError("no CUDA device: lehamoc_b200 has no CPU fallback (the CPU oracle lives in oracle/ "
                       "and is test infrastructure only)")


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, cfg: dict, max_walkers: int, ic_path: str = "R", device: int | None = None):
        _require_cuda()
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.cfg = dict(cfg)
        if "LHM_CONST_B" in os.environ:                      # experiments: force the exponential cache on/off
            self.cfg["const_B"] = int(os.environ["LHM_CONST_B"])
        cfg = self.cfg
        hadronic = any(cfg.get(k) for k in ("pg_pi_l_flag", "pg_BH_l_flag", "pg_pi_emis_flag", "neutrino_flag",
                                            "pg_BH_emis_flag", "pp_l_flag", "pp_ee_emis_flag", "pp_g_emis_flag",
                                            "pp_nu_emis_flag"))
        if ic_path == "G" and not (cfg.get("shared_ic_grids") and (cfg.get("driver", 0) == 0
                                                                   or (cfg.get("driver") == 1 and not hadronic))):
            raise LhmError("ic_path 'G' (dense contraction over walkers) needs a batch whose walkers share g_el / nu_ic / "
                           "nu_tot (cfg.shared_ic_grids) and a leptonic run: LeMoC.py / Code.py, or LeHaMoC.py with "
                           "every hadronic flag off")
        self.ic_path = ic_path
        c = lhm_config()
        for k, v in cfg.items():
            setattr(c, k, int(v))
        c.ic_path = IC_PATHS[ic_path]
        c.max_walkers = int(max_walkers)
        self._c = c
        self.max_walkers = int(max_walkers)
        self.stream = torch.cuda.current_stream(self.device)
        h = C.c_void_p()
        check(self.lib.lhm_create(C.byref(c), self.device, C.c_void_p(self.stream.cuda_stream), C.byref(h)), "lhm_create")
        self.h = h
        self.W = 0
        self._keep = None           # device inputs of the current batch (the library reads them again)

    def set_cluster(self, ctas_per_walker: int):
        """Evolve every walker on a thread-block cluster of 2, 4 or 8 CTAs (lhm_set_cluster): for ensembles smaller than
        the GPU (148 SMs), where one CTA per walker leaves most of the machine idle.  Leptonic drivers, ic_path 'R'."""
        check(self.lib.lhm_set_cluster(self.h, int(ctas_per_walker)), "lhm_set_cluster")
        self.cluster = int(ctas_per_walker)

    def close(self):
        if getattr(self, "h", None):
            self.lib.lhm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- sizes ------------------------------------------------------------------------------------
    @property
    def Ng(self): return self.cfg["Ng"]
    @property
    def Ns(self): return self.cfg["Ns"]
    @property
    def Ni(self): return self.cfg["Ni"]
    @property
    def Nt(self): return self.cfg["Nt"]

    def smem_bytes(self):
        return self.lib.lhm_step_smem_bytes(self.h)

    def table_bytes_per_walker(self):
        return int(self.lib.lhm_table_bytes_per_walker(self.h))

    def _dev(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(device=f"cuda:{self.device}", dtype=torch.float64).contiguous()
        a = np.ascontiguousarray(a, dtype=np.float64)
        if not a.flags.writeable:                  # broadcast views of shared grids: torch wants a writable buffer
            a = a.copy()
        return torch.from_numpy(a).to(f"cuda:{self.device}", non_blocking=True)

    def _new(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=f"cuda:{self.device}")

    # -- batch set-up -----------------------------------------------------------------------------
    def setup(self, batch: Batch):
        """Copy the packed inputs to the device and build the per-walker tables (lhm_setup_batch)."""
        for k in ("Ng", "Ns", "Ni", "Nt"):
            if batch.cfg[k] != self.cfg[k]:
                raise LhmError(f"batch {k}={batch.cfg[k]} does not match the engine ({self.cfg[k]})")
        W = batch.W
        self._check_shared(batch)
        if self.ic_path == "G":
            self.set_lockstep(batch.nsteps)
        dev = [self._dev(a) for a in batch.arrays()]
        if self.cfg.get("driver", 0) >= 1:
            self._keep_pr = (self._dev(batch.g_pr), self._dev(batch.pr_inj))
            check(self.lib.lhm_set_protons(self.h, _ptr(self._keep_pr[0]), _ptr(self._keep_pr[1])), "lhm_set_protons")
        self.setup_device(W, *dev)

    def set_lockstep(self, nsteps):
        """ic_path G advances the whole batch in lock-step: every walker must run the same number of timesteps."""
        nsteps = np.atleast_1d(np.asarray(nsteps))
        if not np.all(nsteps == nsteps[0]):
            raise LhmError("ic_path 'G': the walkers of a batch must run the same number of timesteps")
        check(self.lib.lhm_set_lockstep(self.h, int(nsteps[0])), "lhm_set_lockstep")

    def _check_shared(self, batch):
        if self.cfg.get("shared_ic_grids") and not (np.all(batch.g_el == batch.g_el[0])
                                                    and np.all(batch.nu_ic == batch.nu_ic[0])
                                                    and np.all(batch.nu_tot == batch.nu_tot[0])):
            raise LhmError("this engine was created for batches whose walkers share g_el/nu_ic/nu_tot "
                           "(cfg.shared_ic_grids); the batch does not")

    def setup_device(self, W, consts, g_el, nu_syn, nu_ic, nu_tot, el_inj, ext):
        """Device tensors in, tables built on the stream.  With cfg.shared_ic_grids the caller guarantees that every
        walker has the same g_el / nu_ic / nu_tot rows (not re-checked here: that would need a synchronisation)."""
        self._keep = (consts, g_el, nu_syn, nu_ic, nu_tot, el_inj, ext)
        check(self.lib.lhm_setup_batch(self.h, int(W), *[_ptr(t) for t in self._keep]), "lhm_setup_batch")
        self.W = int(W)

    def setup_packed(self, req):
        """Set-up on the device from a setup_host.PackRequest: one small H2D copy (a row of scalars per walker and the
        three shared grids), lhm_pack_batch builds the [W, n] arrays in device buffers this engine owns, then the tables
        (lhm_setup_batch).  The arrays stay available as self._keep (device tensors, order of Batch.arrays())."""
        from .constants import c, m_el, sigmaT, q
        cfg = req.cfg
        for k in ("Ng", "Ns", "Ni", "Nt"):
            if cfg[k] != self.cfg[k]:
                raise LhmError(f"batch {k}={cfg[k]} does not match the engine ({self.cfg[k]})")
        if bool(self.cfg.get("shared_ic_grids")) and not cfg["shared_ic_grids"]:
            raise LhmError("this engine was created for batches whose walkers share g_el/nu_ic/nu_tot "
                           "(cfg.shared_ic_grids); the batch does not")
        W, Ng, Nh, Nt = req.W, cfg["Ng"], cfg["Ni"], cfg["Nt"]
        if self.ic_path == "G":
            if req.variant != "script":
                raise LhmError("ic_path 'G': the step count of Code.py depends on the walker (Code.py:183)")
            self.set_lockstep(int(req.scalars["time_end"]))                 # LeMoC.py:196 range(int(time_end))
        if W > self.max_walkers:
            raise LhmError(f"batch of {W} walkers exceeds the engine capacity {self.max_walkers}")
        # one pinned staging buffer -> one H2D copy
        n_in = W * _lib.LHM_PACK_NPAR + Nh + 2 * Nt
        sv = self._pinned("pack_in", (self.max_walkers * _lib.LHM_PACK_NPAR + Nh + 2 * Nt,))
        stage = self._pin["pack_in"]
        if getattr(self, "_pack_ev", None) is not None:
            self._pack_ev.synchronize()              # the previous copy out of the staging buffer has completed
        o = W * _lib.LHM_PACK_NPAR
        sv[:o] = req.walker_params.ravel()
        sv[o:o + Nh] = req.nu_ic_row
        sv[o + Nh:o + Nh + Nt] = req.nu_tot_row
        sv[o + Nh + Nt:n_in] = req.ext_row
        if getattr(self, "_pack_dev", None) is None:
            Wm = self.max_walkers
            self._pack_dev = (self._new(len(sv)), self._new(Wm, _lib.LHM_NCONST), self._new(Wm, Ng),
                              self._new(Wm, Nh + 1), self._new(Wm, Nh), self._new(Wm, Nt), self._new(Wm, Ng),
                              self._new(Wm, Nt))
        din, consts, g_el, nu_syn, nu_ic, nu_tot, el_inj, ext = self._pack_dev
        with torch.cuda.stream(self.stream):
            din[:n_in].copy_(stage[:n_in], non_blocking=True)
            self._pack_ev = torch.cuda.Event()
            self._pack_ev.record(self.stream)
        base = din.data_ptr()
        a = _lib.lhm_pack_args(W=W, Ng=Ng, Nh=Nh, Nt=Nt, variant=0 if req.variant == "script" else 1,
                               c=c, c3=c ** 3., m_el=m_el, mc2=m_el * c ** 2., sigmaT=sigmaT, four_pi=4. * np.pi,
                               nuc_pref=3. / (4. * np.pi) * q, **req.scalars)
        if self.cfg.get("driver", 0) == 2:          # Code.py::LeHaMoC: proton grid and injection as well
            from .constants import m_pr
            Np = self.cfg["Np"]
            if getattr(self, "_pack_pr", None) is None:
                self._pack_pr = (self._new(self.max_walkers, Np), self._new(self.max_walkers, Np))
            a.Np, a.m_pr, a.mpc2 = Np, m_pr, m_pr * c ** 2.
            a.g_pr, a.pr_inj = self._pack_pr[0].data_ptr(), self._pack_pr[1].data_ptr()
            self._keep_pr = (self._pack_pr[0][:W], self._pack_pr[1][:W])
            check(self.lib.lhm_set_protons(self.h, _ptr(self._keep_pr[0]), _ptr(self._keep_pr[1])), "lhm_set_protons")
        elif self.cfg.get("driver", 0) != 0:
            raise LhmError("device-side set-up covers the leptonic drivers and Code.LeHaMoC only")
        a.walker_params = base
        a.nu_ic_row, a.nu_tot_row, a.ext_row = base + 8 * o, base + 8 * (o + Nh), base + 8 * (o + Nh + Nt)
        outs = (consts, g_el, nu_syn, nu_ic, nu_tot, el_inj, ext)
        for name, t in zip(("consts", "g_el", "nu_syn", "nu_ic", "nu_tot", "el_inj", "ext"), outs):
            setattr(a, name, t.data_ptr())
        check(self.lib.lhm_pack_batch(C.byref(a), C.c_void_p(self.stream.cuda_stream)), "lhm_pack_batch")
        self.setup_device(W, *[t[:W] for t in outs])

    def fit_logl_host(self, req, doppler, log_f, obs, out_dev=None):
        """One fit evaluation in ONE library call (lhm_fit_logl_host): PackRequest of the Code.py variant + per-walker
        log10 Doppler factor and log_f + the observation table (a _lib.lhm_fit_obs of device pointers) -> log-likelihoods
        [W] as a NumPy array.  Packing, set-up, cor-factor, time loop, observation transform, likelihood and both copies
        run inside the call.
        out_dev (a float64 CUDA tensor with >= W elements): lhm_fit_logl_device instead -- the likelihoods are written there
        on the engine's stream, nothing is copied back and the call does not wait (returns None)."""
        from .constants import c, m_el, sigmaT, q
        cfg = req.cfg
        for k in ("Ng", "Ns", "Ni", "Nt"):
            if cfg[k] != self.cfg[k]:
                raise LhmError(f"batch {k}={cfg[k]} does not match the engine ({self.cfg[k]})")
        if req.variant != "code":
            raise LhmError("fit_logl_host evaluates the Code.py variant (Fit_emcee)")
        W, Nh, Nt = req.W, cfg["Ni"], cfg["Nt"]
        if W > self.max_walkers:
            raise LhmError(f"batch of {W} walkers exceeds the engine capacity {self.max_walkers}")
        stage = np.empty(W * (_lib.LHM_PACK_NPAR + 2) + Nh + 2 * Nt)
        o = W * _lib.LHM_PACK_NPAR
        stage[:o] = req.walker_params.ravel()
        stage[o:o + Nh] = req.nu_ic_row
        stage[o + Nh:o + Nh + Nt] = req.nu_tot_row
        stage[o + Nh + Nt:o + Nh + 2 * Nt] = req.ext_row
        stage[o + Nh + 2 * Nt:o + Nh + 2 * Nt + W] = doppler
        stage[o + Nh + 2 * Nt + W:] = log_f
        a = _lib.lhm_pack_args(W=W, Ng=cfg["Ng"], Nh=Nh, Nt=Nt, variant=1, c=c, c3=c ** 3., m_el=m_el, mc2=m_el * c ** 2.,
                               sigmaT=sigmaT, four_pi=4. * np.pi, nuc_pref=3. / (4. * np.pi) * q, **req.scalars)
        self.W = W
        if out_dev is not None:
            if out_dev.dtype != torch.float64 or not out_dev.is_cuda or out_dev.numel() < W or not out_dev.is_contiguous():
                raise LhmError("out_dev must be a contiguous float64 CUDA tensor with at least W elements")
            check(self.lib.lhm_fit_logl_device(self.h, C.byref(a), stage.ctypes.data_as(C.c_void_p), C.byref(obs),
                                               C.c_void_p(out_dev.data_ptr())), "lhm_fit_logl_device")
            return None
        out = np.empty(W)
        check(self.lib.lhm_fit_logl_host(self.h, C.byref(a), stage.ctypes.data_as(C.c_void_p), C.byref(obs),
                                         out.ctypes.data_as(C.c_void_p)), "lhm_fit_logl_host")
        return out

    # -- the whole time integration ---------------------------------------------------------------
    def run(self, snapshots: int = 0, want_N: bool = True, want_sed: bool = True):
        S = snapshots if snapshots > 0 else 1
        N = self._new(self.W, S, self.Ng) if want_N else None
        sed = self._new(self.W, S, self.Nt) if want_sed else None
        check(self.lib.lhm_run_batch(self.h, _ptr(N), _ptr(sed), int(snapshots)), "lhm_run_batch")
        if snapshots == 0:
            N = N[:, 0] if N is not None else None
            sed = sed[:, 0] if sed is not None else None
        return N, sed

    def run_lh(self, snapshots: int = 0):
        """Driver 1 (LeHaMoC.py): returns (n_e [W,S,Ng], nuLnu [W,S,Nt], n_p [W,S,Np], N_nu [W,S,50]) device tensors."""
        S = snapshots if snapshots > 0 else 1
        N = self._new(self.W, S, self.Ng)
        sed = self._new(self.W, S, self.Nt)
        Npr = self._new(self.W, S, self.cfg["Np"])
        Nnu = self._new(self.W, S, 50)
        check(self.lib.lhm_run_batch_lh(self.h, _ptr(N), _ptr(sed), _ptr(Npr), _ptr(Nnu), int(snapshots)),
              "lhm_run_batch_lh")
        return N, sed, Npr, Nnu

    def set_bh_surfaces(self, surf_m, surf_p, keep):
        """Driver 1 with pg_BH_emis_flag: ctypes lhm_bh_surface pair (lehamoc_b200.hadronic.bh_surface)."""
        self._bh = (surf_m, surf_p, keep)
        check(self.lib.lhm_set_bh_surfaces(self.h, C.byref(surf_m), C.byref(surf_p)), "lhm_set_bh_surfaces")

    def set_had_tables(self, tables_struct):
        """Driver 1 with photo-hadronic flags: ctypes lhm_had_tables of device pointers (lehamoc_b200.hadronic)."""
        self._had_tables = tables_struct
        check(self.lib.lhm_set_had_tables(self.h, C.byref(tables_struct)), "lhm_set_had_tables")

    def run_into(self, N, sed, snapshots: int = 0):
        check(self.lib.lhm_run_batch(self.h, _ptr(N), _ptr(sed), int(snapshots)), "lhm_run_batch")

    def _pinned(self, key, shape):
        """NumPy view of a cached page-locked buffer (grown on demand)."""
        n = int(np.prod(shape))
        pool = self.__dict__.setdefault("_pin", {})
        t = pool.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True)
            pool[key] = t
        return t.numpy()[:n].reshape(shape)

    def run_host(self, batch: Batch, snapshots: int = 0, N_out=None, sed_out=None, pinned: bool = False, out_slot: int = 0):
        """lhm_run_batch_host: host buffers in, host buffers out (H2D + set-up + run + D2H + sync).
        With pinned=True the returned arrays are views of reusable page-locked buffers: the pair named by `out_slot`,
        overwritten by the next call with the same slot."""
        W = batch.W
        S = snapshots if snapshots > 0 else 1
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.ic_path == "G":
            self.set_lockstep(batch.nsteps)
        ins = batch.arrays()
        if pinned:
            # stage through page-locked buffers so the copies run at full PCIe/NVLink-C2C rate
            ins = [self._pinned(f"in{i}", a.shape) for i, a in enumerate(batch.arrays())]
            for dst, src in zip(ins, batch.arrays()):
                np.copyto(dst, src)
            if N_out is None:
                N_out = self._pinned(f"N{out_slot}", (W, S, self.Ng))
            if sed_out is None:
                sed_out = self._pinned(f"sed{out_slot}", (W, S, self.Nt))
        if N_out is None:
            N_out = np.empty((W, S, self.Ng))
        if sed_out is None:
            sed_out = np.empty((W, S, self.Nt))
        check(self.lib.lhm_run_batch_host(self.h, W, *[hp(a) for a in ins], hp(N_out), hp(sed_out),
                                          int(snapshots)), "lhm_run_batch_host")
        self.W = W
        self._keep = None
        if snapshots == 0:
            return N_out.reshape(W, S, self.Ng)[:, 0], sed_out.reshape(W, S, self.Nt)[:, 0]
        return N_out, sed_out

    def run_packed_host(self, req, out_slot: int = 0):
        """Device-packed twin of run_host(pinned=True): PackRequest in, (n_e [W,Ng], nuLnu [W,Nt]) out as views of
        page-locked host buffers (the pair named by `out_slot`, overwritten by the next call with the same slot);
        final state only.  H2D: one row of scalars per walker + three shared grids."""
        self.setup_packed(req)
        W = req.W
        if getattr(self, "_out_dev", None) is None:
            self._out_dev = (self._new(self.max_walkers, 1, self.Ng), self._new(self.max_walkers, 1, self.Nt))
        Nd, sd = self._out_dev
        self.run_into(Nd, sd, 0)
        kN, kS = f"N{out_slot}", f"sed{out_slot}"
        Nh_, sh_ = self._pinned(kN, (W, self.Ng)), self._pinned(kS, (W, self.Nt))
        # The copies back run on a stream of their own behind an event: other engines may have queued later batches on
        # the compute stream (LeMoC.run_many: one FIFO stream for all worker threads), and their kernels need not wait
        # for this batch's 3.4 MB to cross PCIe.  The device buffers are this engine's; it is not launched again before
        # `done` has been waited for.
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        ready = torch.cuda.Event()
        ready.record(self.stream)
        self._copy_stream.wait_event(ready)
        with torch.cuda.stream(self._copy_stream):
            self._pin[kN][:W * self.Ng].copy_(Nd.view(-1)[:W * self.Ng], non_blocking=True)
            self._pin[kS][:W * self.Nt].copy_(sd.view(-1)[:W * self.Nt], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        done.synchronize()          # this batch only
        return Nh_, sh_

    def last_icg_timing(self):
        """ic_path G: (mean ms per contraction launch, launches timed, CTAs, dynamic smem bytes per CTA) of the last run."""
        ms, n, ctas, smem = C.c_float(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.lhm_last_icg_timing(self.h, C.byref(ms), C.byref(n), C.byref(ctas), C.byref(smem)),
              "lhm_last_icg_timing")
        return ms.value, n.value, ctas.value, smem.value

    def last_hadtab_timing(self):
        """Driver 1 with photo-hadronic flags: (ms of the last operator-table build or -1, table sets, bytes)."""
        ms, n, nb = C.c_float(), C.c_int(), C.c_longlong()
        check(self.lib.lhm_last_hadtab_timing(self.h, C.byref(ms), C.byref(n), C.byref(nb)), "lhm_last_hadtab_timing")
        return ms.value, n.value, nb.value

    def last_timings(self):
        """(setup_ms, cor_ms, run_ms) of the most recent launches, measured with CUDA events on the stream."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        check(self.lib.lhm_last_timings(self.h, C.byref(a), C.byref(b), C.byref(c)), "lhm_last_timings")
        return a.value, b.value, c.value

    # -- function-level entry points (unit-test twins of LeHaMoC_f.py) ------------------------------
    def cor_factor(self):
        out = self._new(self.W)
        check(self.lib.lhm_cor_factor_batch(self.h, _ptr(out)), "lhm_cor_factor_batch")
        return out

    def photons_tot(self, photons_syn, photons_IC, radius):
        out = self._new(self.W, self.Nt)
        a, b, r = self._dev(photons_syn), self._dev(photons_IC), self._dev(radius)   # keep alive across the launch
        check(self.lib.lhm_photons_tot_batch(self.h, _ptr(a), _ptr(b), _ptr(r), _ptr(out)), "lhm_photons_tot_batch")
        return out

    def U_ph(self, n, radius):
        out = self._new(self.W, self.Ng)
        a, r = self._dev(n), self._dev(radius)
        check(self.lib.lhm_uph_batch(self.h, _ptr(a), _ptr(r), _ptr(out)), "lhm_uph_batch")
        return out

    def syn_ssa(self, n_e, Bfield):
        Q = self._new(self.W, self.Ns - 1)
        aS = self._new(self.W, self.Ns)
        aI = self._new(self.W, self.Ni)
        a, b = self._dev(n_e), self._dev(Bfield)
        check(self.lib.lhm_syn_ssa_batch(self.h, _ptr(a), _ptr(b), _ptr(Q), _ptr(aS), _ptr(aI)), "lhm_syn_ssa_batch")
        return Q, aS, aI

    def ic_emissivity(self, n_e, n, ic_path: str | None = None, out=None):
        Q = self._new(self.W, self.Ni - 1) if out is None else out
        path = IC_PATHS[ic_path or self.ic_path]
        a, b = self._dev(n_e), self._dev(n)
        check(self.lib.lhm_ic_emissivity_batch(self.h, _ptr(a), _ptr(b), _ptr(Q), path), "lhm_ic_emissivity_batch")
        return Q

    def gg(self, n, photons_IC_dens=None):
        a = self._new(self.W, self.Ni)
        Q = self._new(self.W, self.Ng - 1)
        nn = self._dev(n)
        pic = self._dev(photons_IC_dens) if photons_IC_dens is not None else None
        check(self.lib.lhm_gg_batch(self.h, _ptr(nn), _ptr(pic), _ptr(a), _ptr(Q)), "lhm_gg_batch")
        return a, Q


def bidiag_solve(b, c, d, device=None):
    """thomas() with a == 0 for a batch of systems [W, n] (lhm_bidiag_solve_batch)."""
    _require_cuda()
    lib = _lib.load()
    dev = f"cuda:{torch.cuda.current_device() if device is None else device}"
    tb, tc, td = (torch.as_tensor(np.atleast_2d(x), dtype=torch.float64).to(dev).contiguous() for x in (b, c, d))
    x = torch.empty_like(td)
    W, n = td.shape
    check(lib.lhm_bidiag_solve_batch(W, n, _ptr(tb), _ptr(tc), _ptr(td), _ptr(x),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lhm_bidiag_solve_batch")
    return x


def tridiag_solve(a, b, c, d, device=None):
    """thomas(a, b, c, d) for a batch of systems [W, n] (lhm_tridiag_solve_batch)."""
    _require_cuda()
    lib = _lib.load()
    dev = f"cuda:{torch.cuda.current_device() if device is None else device}"
    ta, tb, tc, td = (torch.as_tensor(np.atleast_2d(x), dtype=torch.float64).to(dev).contiguous() for x in (a, b, c, d))
    x = torch.empty_like(td)
    W, n = td.shape
    work = torch.empty((2, W, n), dtype=torch.float64, device=dev)
    check(lib.lhm_tridiag_solve_batch(W, n, _ptr(ta), _ptr(tb), _ptr(tc), _ptr(td), _ptr(work), _ptr(x),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lhm_tridiag_solve_batch")
    return x


def fp64_peak_tflops(device: int = 0) -> float:
    _require_cuda()
    lib = _lib.load()
    out = C.c_double()
    check(lib.lhm_fp64_peak_probe(device, C.c_void_p(torch.cuda.current_stream(device).cuda_stream), C.byref(out)),
          "lhm_fp64_peak_probe")
    return out.value
