"""Drop-in for Fit_emcee/Code.py: `LeMoC(params, fileName) -> (nu_tot, nuLnu)` plus a batched twin.

Same call, same meaning of `params = [log R0, log B0, log gamma_min, log gamma_max, log l_e, p]`
(Fit_emcee/Code.py:51-61), same frozen parameter file (Code.py:63-94), same return value (Code.py:277):
the comoving-frame spectrum after the last timestep.  The time integration runs on the GPU
(lhm_run_batch_host); there is no CPU path.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .setup_host import (pack_code, pack_code_lehamoc, pack_code_lehamoc_request, pack_code_request,
                         read_parameter_file)

_engines = {}


def _engine_for(batch, ic_path, capacity):
    key = (tuple(sorted(batch.cfg.items())), ic_path)
    eng = _engines.get(key)
    if eng is None or eng.max_walkers < capacity:
        if eng is not None:
            eng.close()
        eng = Engine(batch.cfg, max(capacity, 64), ic_path=ic_path)
        _engines[key] = eng
    return eng


def LeMoC_batch(params, fileName, ic_path="R", return_particles=False, device_pack=False):
    """`params` [W,6]; returns (nu_tot [W,Nt], nuLnu [W,Nt]) (and n_e = N_el/V [W,Ng] if asked).
    device_pack=True builds grids and injection on the device (lhm_pack_batch: 2-3x less host time per call, grids
    equal to the NumPy ones to ~1e-14 instead of bit for bit); the returned arrays are then views of page-locked
    buffers that the next call reuses."""
    frozen = fileName if isinstance(fileName, dict) else read_parameter_file(fileName)
    if device_pack:
        req = pack_code_request(params, frozen)
        eng = _engine_for(req, ic_path, req.W)
        N, sed = eng.run_packed_host(req)
        return (req.nu_tot, sed, N) if return_particles else (req.nu_tot, sed)
    batch = pack_code(params, frozen)
    eng = _engine_for(batch, ic_path, batch.W)
    N, sed = eng.run_host(batch, snapshots=0)
    if return_particles:
        return batch.nu_tot, sed, N
    return batch.nu_tot, sed


def LeMoC(params, fileName):
    """Fit_emcee/Code.py:51: one model evaluation."""
    nu_tot, sed = LeMoC_batch(np.asarray(params, dtype=np.float64)[None, :], fileName)
    return (nu_tot[0], sed[0])


def LeHaMoC_batch(params, fileName, ic_path="R", device_pack=False):
    """`params` [W,10] (Code.py:285-295); returns (nu_tot [W,Nt], nuLnu [W,Nt]) of Fit_emcee/Code.py::LeHaMoC, the
    leptonic model plus a proton population that cools and radiates by synchrotron only.  device_pack: as LeMoC_batch."""
    frozen = fileName if isinstance(fileName, dict) else read_parameter_file(fileName)
    if device_pack:
        req = pack_code_lehamoc_request(params, frozen)
        eng = _engine_for(req, ic_path, req.W)
        eng.setup_packed(req)
        _, sed, _, _ = eng.run_lh(snapshots=0)
        return req.nu_tot, sed[:, 0].cpu().numpy()
    batch = pack_code_lehamoc(params, frozen)
    eng = _engine_for(batch, ic_path, batch.W)
    eng.setup(batch)
    _, sed, _, _ = eng.run_lh(snapshots=0)
    return batch.nu_tot, sed[:, 0].cpu().numpy()


def LeHaMoC(params, fileName):
    """Fit_emcee/Code.py:280: one evaluation of the proton-synchrotron model."""
    nu_tot, sed = LeHaMoC_batch(np.asarray(params, dtype=np.float64)[None, :], fileName)
    return (nu_tot[0], sed[0])
