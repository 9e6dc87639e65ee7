"""Drop-in for the model / likelihood layer of Fit_emcee/fit_emcee.py plus a vectorised ensemble sampler.

Same names and arguments as the reference:
    read_data(filename1, filename2)                                       fit_emcee.py:63-88
    Model(D, z)(x, theta)                                                 fit_emcee.py:95-115
    log_likelihood_LeMoC(theta, x, y, yerr, yerr1, flags)                 fit_emcee.py:121-145
    log_prior_LeMoC(theta)                                                fit_emcee.py:173-179
    log_probability(theta, x, y, yerr, yerr1, flags, flag_c)              fit_emcee.py:190-199
and the batched twins `log_prior_batch`, `log_probability_batch(thetas[W,8], ...)`: all walkers of an
ensemble half-step are evolved in ONE fused-kernel launch and their log-likelihoods come out of the
`lhm_loglike_batch` epilogue -- this replaces `Pool().map(log_probability, thetas)` (fit_emcee.py:239-242).
With `torch.distributed` initialised the walkers are cut into per-GPU blocks and the log-probabilities are
all-gathered (lehamoc_b200/shard.py).

`run_mcmc` is an affine-invariant stretch-move sampler (Goodman & Weare 2010) with emcee's red/blue split and
call pattern (`log_prob_fn(thetas[W/2, ndim]) -> [W/2]`, emcee's `vectorize=True`); chains are stored as `.npz`
(emcee's HDF5 backend, corner and matplotlib are outside the scope of this path).

The model evaluation runs on the GPU only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np
import torch

from . import _lib, shard
from ._lib import check
from .constants import eV, pc
from .engine import Engine, _ptr, _require_cuda
from .setup_host import (pack_code, pack_code_lehamoc, pack_code_lehamoc_request, pack_code_request,
                         read_parameter_file)

# fit_emcee.py:31-32
D_DEFAULT = 3262    # luminosity distance (Mpc)
Z_DEFAULT = 0.557   # redshift


# ---------------------------------------------------------------------------------------------------------
# data reader, fit_emcee.py:63-88 (pandas-free: same arithmetic on plain arrays)
# ---------------------------------------------------------------------------------------------------------
def read_data(filename1, filename2):
    data_fermi = np.loadtxt(filename2, usecols=range(0, 4))
    E = np.log10(data_fermi[:, 0])
    vFv = np.log10(data_fermi[:, 1] * 1e6 * eV)                                   # MeV/cm2/s to erg/cm2/s
    vFve_hi_log = np.log10(data_fermi[:, 1] + data_fermi[:, 2]) - np.log10(data_fermi[:, 1])
    vFve_lo_log = -np.log10(data_fermi[:, 1] - data_fermi[:, 2]) + np.log10(data_fermi[:, 1])
    data = np.loadtxt(filename1, usecols=range(0, 4))
    xd = np.append(data[:, 0], E)
    yd = np.append(data[:, 2], vFv)
    yderr_hi = np.append(data[:, 3], vFve_hi_log)
    yderr_lo = np.append(data[:, 3], vFve_lo_log)
    flags = np.append(np.zeros(len(data)), data_fermi[:, 3])
    return xd, yd, yderr_hi, yderr_lo, flags


# ---------------------------------------------------------------------------------------------------------
# priors, fit_emcee.py:173-179 (host arithmetic, vectorised)
# ---------------------------------------------------------------------------------------------------------
_PRIOR_LO = np.array([14., -2., 0., 5.0, -6., 1.5, 0., -6.0])
_PRIOR_HI = np.array([17., 2., 4.5, 7., -2., 3., 3., 1.0])


# log_prior_LeHaMoC, fit_emcee.py:181-188: [P1..P11, log_f]
_PRIOR_LH_LO = np.array([14., -1., 0., 4., -5., 1.1, 0., 6., -5., 1.1, 0., -7.0])
_PRIOR_LH_HI = np.array([17., 3., 3., 6., 0., 3., 5., 10., 0., 3., 3., 1.0])


def log_prior_batch(thetas, flag_c="LeMoC"):
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    lo, hi = (_PRIOR_LO, _PRIOR_HI) if flag_c == "LeMoC" else (_PRIOR_LH_LO, _PRIOR_LH_HI)
    if thetas.shape[1] != len(lo):
        raise ValueError(f"{flag_c} fit: theta must have {len(lo)} values (fit_emcee.py:174,182)")
    ok = np.all((thetas > lo) & (thetas < hi), axis=1)
    return np.where(ok, 0.0, -np.inf)


def log_prior_LeHaMoC(theta):
    return float(log_prior_batch(np.asarray(theta, dtype=np.float64)[None, :], "LeHaMoC")[0])


def log_prior_LeMoC(theta):
    return float(log_prior_batch(np.asarray(theta, dtype=np.float64)[None, :])[0])


# ---------------------------------------------------------------------------------------------------------
# batched model + likelihood on the GPU
# ---------------------------------------------------------------------------------------------------------
class FitEvaluator:
    """Owns the engine, the frozen parameter file and the device copy of the observed SED."""

    def __init__(self, x, y, yerr, yerr1, flags, param_file="Parameters.txt", D=D_DEFAULT, z=Z_DEFAULT,
                 ic_path="R", device=None, flag_c="LeMoC", device_pack=True, cluster=1):
        """cluster = 2 | 4 | 8 (flag_c 'LeMoC', ic_path 'R'): every walker on a thread-block cluster of that many CTAs
        (Engine.set_cluster) -- for ensembles smaller than the GPU, like the reference's own 24 walkers per half-ensemble.
        The setting decides the last bits of the result, not the ensemble size."""
        if flag_c not in ("LeMoC", "LeHaMoC"):
            raise ValueError("flag_c must be 'LeMoC' or 'LeHaMoC' (fit_emcee.py:23)")
        if int(cluster) != 1 and (flag_c != "LeMoC" or ic_path != "R"):
            raise ValueError("cluster > 1 needs flag_c 'LeMoC' and ic_path 'R'")
        self.cluster = int(cluster)
        self.flag_c = flag_c
        self.device_pack = bool(device_pack)    # False: NumPy set-up on the host (bit-identical grids, 2-3x slower)
        self.nfree = 6 if flag_c == "LeMoC" else 10                 # free parameters of the numerical model
        _require_cuda()
        self.lib = _lib.load()
        self.frozen = param_file if isinstance(param_file, dict) else read_parameter_file(param_file)
        self.D = D * 1e6 * pc                                       # fit_emcee.py:99, D in Mpc
        self.z = z
        self.distance_norm = float(np.log10(4 * np.pi * self.D ** 2))
        self.log1pz = float(np.log10(1. + z))
        self.ic_path = ic_path
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.eng = None
        self.set_data(x, y, yerr, yerr1, flags)

    def set_data(self, x, y, yerr, yerr1, flags):
        dev = f"cuda:{self.device}"
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, yerr, yerr1, flags)]
        self.nd = len(arrs[0])
        if any(len(a) != self.nd for a in arrs):
            raise ValueError("x, y, yerr, yerr1, flags must have the same length")
        self._data = [torch.from_numpy(a).to(dev) for a in arrs]
        self._obs = _lib.lhm_fit_obs(nd=self.nd, log10_1pz=self.log1pz, distance_norm=self.distance_norm,
                                     **{n: t.data_ptr() for n, t in zip(("x", "y", "yerr_hi", "yerr_lo", "flags"), self._data)})

    def _engine(self, batch):
        if self.eng is None or self.eng.max_walkers < batch.W or any(
                self.eng.cfg[k] != batch.cfg[k] for k in batch.cfg):
            if self.eng is not None:
                self.eng.close()
            self.eng = Engine(batch.cfg, max(batch.W, 64), ic_path=self.ic_path, device=self.device)
            if self.cluster > 1:
                self.eng.set_cluster(self.cluster)
        return self.eng

    def model_and_loglike(self, thetas, want_model=False):
        """thetas [W, nfree+2] = [model parameters, log10 delta, log_f] -> (logl [W] device tensor, y_model [W, nd] or
        None).  No prior here (the reference evaluates the likelihood whatever the prior says, fit_emcee.py:192-193)."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        W, nf = thetas.shape[0], self.nfree
        if self.flag_c == "LeMoC" and self.device_pack:
            req = pack_code_request(thetas[:, :nf], self.frozen)     # grids and injection are built on the device
            eng = self._engine(req)
            eng.setup_packed(req)
            _, sed = eng.run(want_N=False)                           # [W, Nt] on the device
        elif self.flag_c == "LeMoC":
            batch = pack_code(thetas[:, :nf], self.frozen)
            eng = self._engine(batch)
            eng.setup(batch)
            _, sed = eng.run(want_N=False)
        elif self.device_pack:
            req = pack_code_lehamoc_request(thetas[:, :nf], self.frozen)
            eng = self._engine(req)
            eng.setup_packed(req)
            sed = eng.run_lh(snapshots=0)[1][:, 0]
        else:
            batch = pack_code_lehamoc(thetas[:, :nf], self.frozen)
            eng = self._engine(batch)
            eng.setup(batch)
            sed = eng.run_lh(snapshots=0)[1][:, 0]
        nu_tot = eng._keep[4]
        dev = sed.device
        dop = torch.from_numpy(np.ascontiguousarray(thetas[:, nf])).to(dev)
        logf = torch.from_numpy(np.ascontiguousarray(thetas[:, nf + 1])).to(dev)
        logl = torch.empty(W, dtype=torch.float64, device=dev)
        ym = torch.empty((W, self.nd), dtype=torch.float64, device=dev) if want_model else None
        sed = sed.contiguous()
        check(self.lib.lhm_loglike_batch(W, eng.Nt, _ptr(nu_tot), _ptr(sed), _ptr(dop), _ptr(logf),
                                         self.log1pz, self.distance_norm, self.nd, *[_ptr(t) for t in self._data],
                                         _ptr(logl), _ptr(ym), C.c_void_p(eng.stream.cuda_stream)),
              "lhm_loglike_batch")
        return logl, ym

    def log_probability(self, thetas):
        """[W] NumPy array; -inf outside the prior box (those walkers are not evolved at all)."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        lp = log_prior_batch(thetas, self.flag_c)
        out = np.full(thetas.shape[0], -np.inf)
        inside = np.isfinite(lp)
        if inside.any():
            th = thetas[inside]
            if self.flag_c == "LeMoC" and self.device_pack and self.ic_path != "G":
                # the whole evaluation in one library call (lhm_fit_logl_host): no torch op, no second synchronisation
                req = pack_code_request(th[:, :self.nfree], self.frozen)
                logl = self._engine(req).fit_logl_host(req, th[:, self.nfree], th[:, self.nfree + 1], self._obs)
            else:
                logl = self.model_and_loglike(th)[0].cpu().numpy()
            out[inside] = lp[inside] + logl
        return out

    def log_probability_sharded(self, thetas):
        """Same, with the walkers cut into per-rank blocks and the result all-gathered (shard.py)."""
        if (shard.world()[1] > 1 and self.flag_c == "LeMoC" and self.device_pack and self.ic_path != "G"
                and torch.distributed.get_backend() == "nccl"):
            return self._log_probability_nccl(thetas)

        def local(block):
            if len(block) == 0:
                return np.empty(0)
            return self.log_probability(block)
        return shard.evaluate_sharded(local, thetas)

    def _log_probability_nccl(self, thetas):
        """The sharded evaluation without a host round trip per rank: this rank's block goes through
        lhm_fit_logl_device (likelihoods stay on the GPU, no synchronisation), ONE all-gather on the same stream carries
        the blocks and a per-rank status word, one copy brings the ensemble's likelihoods to the host.  The prior is a
        function of theta alone, so every rank evaluates it for the whole ensemble on the host.  A rank whose block fails
        (a packer ValueError, say) still takes part in the gather and all ranks raise together."""
        import torch.distributed as dist
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        W, nf = thetas.shape[0], self.nfree
        rank, ws = shard.world()
        lp = log_prior_batch(thetas, self.flag_c)
        a, b = shard.block_range(W, ws, rank)
        m = -(-W // ws)                                     # the largest block; slot m of a rank's row is its status word
        bufs = getattr(self, "_nccl_bufs", None)
        if bufs is None or bufs[0].numel() != m + 1 or bufs[1].numel() != ws * (m + 1):
            dev = torch.device("cuda", self.device)
            bufs = (torch.zeros(m + 1, dtype=torch.float64, device=dev), torch.empty(ws * (m + 1), dtype=torch.float64, device=dev),
                    torch.empty(m + 1, dtype=torch.float64, device=dev), torch.empty(ws * (m + 1), dtype=torch.float64).pin_memory())
            self._nccl_bufs = bufs
        loc, full, tmp, pin = bufs
        err = None
        try:
            inside = np.isfinite(lp[a:b])
            n_in = int(inside.sum())
            if n_in:
                th = thetas[a:b] if n_in == b - a else thetas[a:b][inside]
                req = pack_code_request(th[:, :nf], self.frozen)
                self._engine(req).fit_logl_host(req, th[:, nf], th[:, nf + 1], self._obs,
                                                out_dev=loc if n_in == b - a else tmp)
                if n_in != b - a:                           # walkers outside the prior box are not evolved: scatter the rest
                    with torch.cuda.stream(self.eng.stream):
                        idx = torch.from_numpy(np.flatnonzero(inside)).to(loc.device)
                        loc[:b - a].index_copy_(0, idx, tmp[:n_in])
        except Exception as e:                              # noqa: BLE001 -- raised below, on every rank
            err = e
            with torch.cuda.stream(self.eng.stream if self.eng is not None else torch.cuda.current_stream(self.device)):
                loc[m:].fill_(1.0)
        # gather and copy on the stream the evaluation was queued on (the engine's), whatever stream is current now
        stream = self.eng.stream if self.eng is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(full, loc)
            pin.copy_(full, non_blocking=True)
        stream.synchronize()
        host = pin.numpy().reshape(ws, m + 1)
        if err is not None:
            loc[m:].zero_()
            raise err
        if host[:, m].any():
            raise RuntimeError("log_probability_sharded: another rank failed to evaluate its block of walkers")
        sizes = shard.block_sizes(W, ws)
        logl = host[0, :W].copy() if ws == 1 else np.concatenate([host[r, :sizes[r]] for r in range(ws)])
        out = np.full(W, -np.inf)
        ins = np.isfinite(lp)
        out[ins] = lp[ins] + logl[ins]
        return out


_evaluators = {}


def _evaluator(x, y, yerr, yerr1, flags, param_file, D, z, ic_path, flag_c="LeMoC", cluster=1):
    key = (param_file if isinstance(param_file, str) else id(param_file), D, z, ic_path, flag_c, cluster)
    ev = _evaluators.get(key)
    if ev is None:
        ev = FitEvaluator(x, y, yerr, yerr1, flags, param_file, D, z, ic_path, flag_c=flag_c, cluster=cluster)
        _evaluators[key] = ev
    else:
        ev.set_data(x, y, yerr, yerr1, flags)
    return ev


class Model:
    """fit_emcee.py:95-115: comoving SED -> observed log10 nuF_nu on the abscissae x (log10 Hz)."""

    def __init__(self, D=D_DEFAULT, z=Z_DEFAULT, param_file="Parameters.txt"):
        self.D = D * 1e6 * pc
        self.z = z
        self.distance_norm = np.log10(4 * np.pi * self.D ** 2)
        self._D_Mpc = D
        self.param_file = param_file

    def __call__(self, x, theta):
        x = np.asarray(x, dtype=np.float64)
        th = np.append(np.asarray(theta, dtype=np.float64), 0.0)[None, :]     # log_f is not used by the model
        ev = _evaluator(x, x, x, x, np.zeros_like(x), self.param_file, self._D_Mpc, self.z, "R")
        _, ym = ev.model_and_loglike(th, want_model=True)
        return ym[0].cpu().numpy()


def log_likelihood_LeMoC(theta, x, y, yerr, yerr1, flags, param_file="Parameters.txt", D=D_DEFAULT, z=Z_DEFAULT):
    ev = _evaluator(x, y, yerr, yerr1, flags, param_file, D, z, "R")
    logl, _ = ev.model_and_loglike(np.asarray(theta, dtype=np.float64)[None, :])
    return float(logl[0].item())


def log_likelihood_LeHaMoC(theta, x, y, yerr, yerr1, flags, param_file="Parameters.txt", D=D_DEFAULT, z=Z_DEFAULT):
    """fit_emcee.py:147-171: theta = [P1..P11, log_f], model Code.LeHaMoC."""
    ev = _evaluator(x, y, yerr, yerr1, flags, param_file, D, z, "R", "LeHaMoC")
    logl, _ = ev.model_and_loglike(np.asarray(theta, dtype=np.float64)[None, :])
    return float(logl[0].item())


def log_probability_batch(thetas, x, y, yerr, yerr1, flags, flag_c="LeMoC", param_file="Parameters.txt",
                          D=D_DEFAULT, z=Z_DEFAULT, ic_path="R", cluster=1):
    """fit_emcee.py:190-199 for thetas [W, 8] (LeMoC) or [W, 12] (LeHaMoC) at once -> [W].  cluster: CTAs per walker
    (FitEvaluator), worth 2-8 for ensembles of a few dozen walkers."""
    return _evaluator(x, y, yerr, yerr1, flags, param_file, D, z, ic_path, flag_c, cluster).log_probability_sharded(thetas)


def log_probability(theta, x, y, yerr, yerr1, flags, flag_c):
    return float(log_probability_batch(np.asarray(theta, dtype=np.float64)[None, :], x, y, yerr, yerr1, flags,
                                       flag_c)[0])


# ---------------------------------------------------------------------------------------------------------
# ensemble sampler: stretch move with emcee's red/blue split (emcee.moves.RedBlueMove / StretchMove, a = 2)
# ---------------------------------------------------------------------------------------------------------
def run_mcmc(log_prob_fn, pos, nsteps, seed=0, a=2.0, progress=False, chain_file=None, checkpoint_every=0):
    """Advance an ensemble `pos` [nwalkers, ndim] for `nsteps` iterations.  `log_prob_fn(thetas[n, ndim]) -> [n]`
    is called with half an ensemble at a time (emcee's vectorize=True call pattern).  Every rank of a
    `torch.distributed` job must call this with the same arguments: the proposal RNG is seeded identically, so all
    ranks hold the same chain and only `log_prob_fn` is sharded.
    A NaN log-probability of a proposal raises, like emcee does ("Probability function returned NaN").  With
    `chain_file` and `checkpoint_every = N` rank 0 rewrites the file every N iterations (the iterations done so far), so a
    crash does not lose the run -- the role of the reference's HDF backend (fit_emcee.py:229-233).
    Returns dict(chain [nsteps, nwalkers, ndim], log_prob [nsteps, nwalkers], accepted [nwalkers])."""
    rng = np.random.default_rng(seed)
    p = np.array(pos, dtype=np.float64)
    nwalkers, ndim = p.shape
    if nwalkers < 2 * ndim or nwalkers % 2:
        raise ValueError("the stretch move needs an even number of walkers >= 2*ndim (emcee's rule)")
    lp = np.asarray(log_prob_fn(p), dtype=np.float64)
    if np.any(np.isnan(lp)):
        raise ValueError("log_prob_fn returned NaN for the initial ensemble")
    chain = np.empty((nsteps, nwalkers, ndim))
    lps = np.empty((nsteps, nwalkers))
    accepted = np.zeros(nwalkers, dtype=np.int64)
    t0 = time.time()
    for it in range(nsteps):
        inds = np.arange(nwalkers) % 2                          # red/blue split, reshuffled every iteration
        rng.shuffle(inds)
        for split in range(2):
            S1 = inds == split
            s, c = p[S1], p[~S1]
            Ns, Nc = len(s), len(c)
            zz = ((a - 1.) * rng.random(Ns) + 1.) ** 2. / a
            factors = (ndim - 1.) * np.log(zz)
            rint = rng.integers(Nc, size=Ns)
            q = c[rint] - (c[rint] - s) * zz[:, None]
            new_lp = np.asarray(log_prob_fn(q), dtype=np.float64)
            if np.any(np.isnan(new_lp)):
                raise ValueError("Probability function returned NaN")          # emcee's behaviour and message
            lnpdiff = factors + new_lp - lp[S1]
            acc = lnpdiff > np.log(rng.random(Ns))
            idx = np.where(S1)[0][acc]
            p[idx] = q[acc]
            lp[idx] = new_lp[acc]
            accepted[idx] += 1
        chain[it] = p
        lps[it] = lp
        if chain_file and checkpoint_every and (it + 1) % checkpoint_every == 0 and it + 1 < nsteps \
                and shard.world()[0] == 0:
            _save_chain(chain_file, chain[:it + 1], lps[:it + 1], accepted, it + 1)
        if progress and shard.world()[0] == 0:
            print(f"\r{it + 1}/{nsteps}  max logp {lp.max():.3f}  {time.time() - t0:.1f} s", end="", flush=True)
    if progress and shard.world()[0] == 0:
        print()
    out = dict(chain=chain, log_prob=lps, accepted=accepted, acceptance_fraction=accepted / max(nsteps, 1))
    if chain_file and shard.world()[0] == 0:
        _save_chain(chain_file, chain, lps, accepted, nsteps)
    return out


def _save_chain(chain_file, chain, lps, accepted, done):
    """Atomic rewrite (temporary file + rename): a reader or a crash never sees a half-written chain."""
    tmp = chain_file + ".tmp.npz"
    np.savez_compressed(tmp, chain=chain, log_prob=lps, accepted=accepted, acceptance_fraction=accepted / max(done, 1),
                        iterations=done)
    os.replace(tmp, chain_file)


def chain_directory(base):
    """fit_emcee.py:38-58: create `base`; if it exists create `base_1`, `base_2`, ... so that the chains of earlier
    runs are never overwritten.  Returns the directory created."""
    try:
        os.makedirs(base)
        return base
    except FileExistsError:
        suffix = 1
        while True:
            cand = f"{base}_{suffix}"
            try:
                os.makedirs(cand, exist_ok=False)
                return cand
            except OSError:
                suffix += 1


def main(argv=None):
    """`python -m lehamoc_b200.fit LeMoC <steps> <_folderName>` -- fit_emcee.py's CLI (fit_emcee.py:20-27,205-245):
    48 walkers x 8 dims around the reference's start point, data and Parameters.txt read from the working
    directory, chain written to chains<_folderName>/LeMoC/out.npz."""
    argv = sys.argv if argv is None else argv
    if len(argv) != 4:
        print('missing code module')
        print('try something like this')
        print('python -m lehamoc_b200.fit LeMoC 1000 _folderName')
        return 1
    flag_c, steps = argv[1], int(argv[2])
    # one process per GPU under torchrun: the walkers of every half-ensemble are sharded over the ranks (shard.py), all
    # ranks hold the same chain, rank 0 writes it
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank, world_size = shard.world()
    directory = 'chains' + argv[3] + '/' + flag_c
    if rank == 0:
        directory = chain_directory(directory)
        print("Directory '%s' created successfully" % directory)
    if world_size > 1:
        box = [directory]
        torch.distributed.broadcast_object_list(box, src=0)
        directory = box[0]
    if flag_c == 'LeMoC':
        params = [15.25808, 0.038818, 3., 6.081445, -4., 2., 1.3]               # fit_emcee.py:208
    else:
        params = [15.5, 1., 0.5, 5., -3., 1.2, 0.5, 7.3, -4.5, 2., 1.]          # fit_emcee.py:212
    xd, yd, yderr_hi, yderr_lo, flags = read_data('OUsed_5BZBJ09553551-11jan2020.txt', 'SED_fermi_data.txt')
    init = [*params, -2]
    pos = init + 1e-2 * np.random.default_rng(0).standard_normal((48, len(init)))   # fit_emcee.py:224-225
    start = time.perf_counter()
    # 24 walkers per half-ensemble on a 148-SM GPU: four CTAs per walker (thread-block clusters) for the SSC model
    cl = 4 if flag_c == 'LeMoC' else 1
    fn = lambda th: log_probability_batch(th, xd, yd, yderr_hi, yderr_lo, flags, flag_c, cluster=cl)
    res = run_mcmc(fn, pos, steps, progress=True, chain_file=os.path.join(directory, "out.npz"),
                   checkpoint_every=max(1, min(100, steps // 10)))
    if rank == 0:
        print("GPU batch evaluation took {0:.1f} seconds".format(time.perf_counter() - start))
        print("mean acceptance fraction {0:.3f}".format(float(np.mean(res["acceptance_fraction"]))))
    if world_size > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
