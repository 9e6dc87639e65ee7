"""Fit layer: observation transform, likelihood, priors, sampler.

The golden vectors in tests/golden/fit_loglike.npz were produced by the reference's OWN definitions of
read_data / Model / log_likelihood_LeMoC / log_prior_LeMoC / log_probability (oracle/make_golden_fit.py exec's
them unmodified from Fit_emcee/fit_emcee.py) on the reference's data files and the reference's Code.LeMoC spectra.
"""
import os

import numpy as np
import pytest

import lehamoc_oracle as orc
from util import GOLD, assert_spectra_close, load_golden


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "fit_loglike.npz"))


# ---- CPU: the oracle restatement against the reference's own functions ---------------------------------------
def test_oracle_likelihood_matches_reference(gold):
    cf, P = load_golden("code_fit")
    m = orc.Model(lambda th, f: None, D=float(gold["D_Mpc"]), z=float(gold["z"]))
    assert m.distance_norm == float(gold["distance_norm"])
    for k, th in enumerate(gold["theta8"]):
        ym = m.transform(gold["xd"], cf["nu_tot"][k], cf["nuLnu"][k], th[6])
        np.testing.assert_allclose(ym, gold["ymodel"][k], rtol=1e-14, atol=0)
        ll = orc.log_likelihood_from_model(ym, th[7], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"])
        assert abs(ll - gold["loglike"][k]) <= 1e-12 * abs(gold["loglike"][k])
        assert orc.log_prior_LeMoC(th) == gold["logprior"][k]


def test_host_prior_and_reader(gold, tmp_path):
    from lehamoc_b200 import fit
    np.testing.assert_array_equal(fit.log_prior_batch(gold["theta8"]), gold["logprior"])
    bad = gold["theta8"].copy()
    bad[0, 0] = 13.0; bad[1, 7] = 1.5; bad[2, 5] = 3.0          # outside / on the open boundary
    assert np.all(np.isneginf(fit.log_prior_batch(bad)[:3]))
    assert fit.log_prior_LeMoC(list(gold["theta8"][0])) == 0.0
    # reader: same file format as the reference's two data files (4 columns each)
    f1, f2 = tmp_path / "mw.txt", tmp_path / "fermi.txt"
    f1.write_text("10.5 0.1 -12.3 0.05\n14.2 0.1 -11.9 0.07\n")
    f2.write_text("1e23 2.0e-6 5.0e-7 0\n1e24 1.0e-6 9.9e-7 1\n")
    x, y, hi, lo, fl = fit.read_data(str(f1), str(f2))
    xo, yo, hio, loo, flo = orc.read_data(str(f1), str(f2))
    for a, b in zip((x, y, hi, lo, fl), (xo, yo, hio, loo, flo)):
        np.testing.assert_array_equal(a, b)
    assert list(fl) == [0., 0., 0., 1.]


def test_sampler_recovers_gaussian():
    """Stretch move on a 3-d Gaussian with a NumPy log-prob: mean and variance come out right, calls are batched."""
    from lehamoc_b200 import fit
    mu, sig = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    calls = []

    def logp(th):
        calls.append(th.shape)
        return -0.5 * np.sum(((th - mu) / sig) ** 2, axis=1)
    rng = np.random.default_rng(1)
    pos = mu + 0.1 * rng.standard_normal((32, 3))
    res = fit.run_mcmc(logp, pos, 1500, seed=3)
    flat = res["chain"][500:].reshape(-1, 3)
    assert np.all(np.abs(flat.mean(axis=0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(axis=0) / sig - 1.0) < 0.15)
    assert calls[0] == (32, 3) and all(c == (16, 3) for c in calls[1:])
    assert 0.2 < res["acceptance_fraction"].mean() < 0.9
    res2 = fit.run_mcmc(logp, pos, 50, seed=3)
    np.testing.assert_array_equal(res2["chain"], res["chain"][:50])      # reproducible from the seed
    with pytest.raises(ValueError):
        fit.run_mcmc(logp, pos[:5], 10)


def test_sampler_nan_checkpoints_and_chain_directories(tmp_path):
    """emcee's behaviours the CLI relies on: a NaN log-probability raises; the chain file is rewritten every N iterations
    (a crash does not lose the run); chain directories get a numeric suffix instead of being overwritten
    (fit_emcee.py:38-58)."""
    from lehamoc_b200 import fit
    rng = np.random.default_rng(2)
    pos = rng.standard_normal((8, 2))
    calls = [0]

    def logp(th):
        calls[0] += 1
        out = -0.5 * np.sum(th ** 2, axis=1)
        if calls[0] == 12:
            out[0] = np.nan
        return out
    f = str(tmp_path / "out.npz")
    with pytest.raises(ValueError, match="Probability function returned NaN"):
        fit.run_mcmc(logp, pos, 50, seed=1, chain_file=f, checkpoint_every=2)
    z = np.load(f)                                             # the checkpoint written before the failure
    assert int(z["iterations"]) == 4 and z["chain"].shape == (4, 8, 2)
    ok = fit.run_mcmc(lambda th: -0.5 * np.sum(th ** 2, axis=1), pos, 4, seed=1)
    np.testing.assert_array_equal(z["chain"], ok["chain"])
    with pytest.raises(ValueError, match="initial ensemble"):
        fit.run_mcmc(lambda th: np.full(len(th), np.nan), pos, 2)
    base = str(tmp_path / "chains_x" / "LeMoC")
    assert fit.chain_directory(base) == base
    assert fit.chain_directory(base) == base + "_1"
    assert fit.chain_directory(base) == base + "_2"
    assert os.path.isdir(base + "_2")


# ---- GPU: the CUDA epilogue against the reference's numbers --------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("device_pack", [False, True])
def test_gpu_loglike_matches_reference(gold, device_pack):
    """device_pack=False: grids and injection packed by the host with NumPy (bit-identical to the reference's);
    True (the default): built on the device by lhm_pack_batch (<= 2 ulp on the grids)."""
    from lehamoc_b200 import fit
    _, P = load_golden("code_fit")
    ev = fit.FitEvaluator(gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"], param_file=P,
                          D=float(gold["D_Mpc"]), z=float(gold["z"]), device_pack=device_pack)
    logl, ym = ev.model_and_loglike(gold["theta8"], want_model=True)
    logl, ym = logl.cpu().numpy(), ym.cpu().numpy()
    # y_model is log10 of the spectrum: 1e-8 relative on the spectrum <=> 4.3e-9 absolute here
    assert np.max(np.abs(ym - gold["ymodel"])) <= 4.3e-9
    # chi^2 amplifies a model shift dy by (y-ym)/sigma^2 ~ up to 1e5 here: compare at the propagated tolerance
    np.testing.assert_allclose(logl, gold["loglike"], rtol=1e-6)
    lp = ev.log_probability(gold["theta8"])
    np.testing.assert_allclose(lp, gold["logprob"], rtol=1e-6)
    bad = gold["theta8"].copy()
    bad[3, 0] = 17.5
    lpb = ev.log_probability(bad)
    assert lpb[3] == -np.inf and np.allclose(np.delete(lpb, 3), np.delete(lp, 3), rtol=0, atol=0)
    # single-call surface of the reference
    m = fit.Model(float(gold["D_Mpc"]), float(gold["z"]), param_file=P)
    y1 = m(gold["xd"], list(gold["theta8"][0, :7]))           # the module-level evaluators are device-packed
    ll0 = fit.log_likelihood_LeMoC(list(gold["theta8"][0]), gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"],
                                   gold["flags"], param_file=P)
    if device_pack:
        np.testing.assert_array_equal(y1, ym[0])
        assert ll0 == logl[0]
    else:
        assert np.max(np.abs(y1 - ym[0])) <= 4.3e-9
        np.testing.assert_allclose(ll0, logl[0], rtol=1e-6)


@pytest.mark.gpu
def test_gpu_one_call_fit_evaluation_equals_the_staged_path(gold):
    """lhm_fit_logl_host (packing, set-up, cor-factor, time loop, likelihood and both copies in ONE library call -- what
    FitEvaluator.log_probability uses) gives bit for bit what the call-by-call path gives, for the golden walkers, for
    random walkers over the prior box, for a batch that grows the engine, and with walkers outside the prior."""
    from lehamoc_b200 import fit
    from lehamoc_b200.presets import fit_walkers
    _, P = load_golden("code_fit")
    ev = fit.FitEvaluator(gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"], param_file=P,
                          D=float(gold["D_Mpc"]), z=float(gold["z"]))
    rng = np.random.default_rng(11)
    for W in (8, 40, 130):
        th = gold["theta8"] if W == 8 else np.column_stack([fit_walkers(W, seed=W), rng.uniform(0.5, 2.5, W),
                                                            rng.uniform(-5., 0., W)])
        staged = fit.log_prior_batch(th, "LeMoC") + ev.model_and_loglike(th)[0].cpu().numpy()
        one = ev.log_probability(th)
        np.testing.assert_array_equal(one, staged)
        assert np.isfinite(one).mean() > 0.5
    np.testing.assert_allclose(ev.log_probability(gold["theta8"]), gold["logprob"], rtol=1e-6)
    bad = gold["theta8"].copy()
    bad[[1, 6], 0] = 17.5
    lpb = ev.log_probability(bad)
    assert np.isneginf(lpb[[1, 6]]).all()
    np.testing.assert_array_equal(np.delete(lpb, [1, 6]), np.delete(ev.log_probability(gold["theta8"]), [1, 6]))


@pytest.mark.gpu
@pytest.mark.parametrize("cluster", [2, 4, 8])
def test_gpu_cluster_per_walker_matches_one_cta(gold, cluster):
    """One walker per thread-block cluster (run_cluster_kernel: IC tiles, gamma-gamma rows and synchrotron rows dealt to
    the CTAs, exchanged through distributed shared memory) against the one-CTA kernel: same log-probabilities up to the
    association of the IC partial sums, the reference's numbers at the usual tolerance, spectra at 1e-12, and the result
    of a walker does not depend on the ensemble it is evaluated in (bit for bit)."""
    from lehamoc_b200 import Code, fit
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.presets import fit_walkers
    from lehamoc_b200.setup_host import pack_code
    _, P = load_golden("code_fit")
    kw = dict(param_file=P, D=float(gold["D_Mpc"]), z=float(gold["z"]))
    data = (gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"])
    ev1 = fit.FitEvaluator(*data, **kw)
    evc = fit.FitEvaluator(*data, cluster=cluster, **kw)
    rng = np.random.default_rng(cluster)
    th = np.column_stack([fit_walkers(37, seed=5), rng.uniform(0.8, 2.2, 37), rng.uniform(-4., -1., 37)])
    lp1, lpc = ev1.log_probability(th), evc.log_probability(th)
    ok = np.isfinite(lp1)
    assert np.array_equal(ok, np.isfinite(lpc)) and ok.sum() > 20
    np.testing.assert_allclose(lpc[ok], lp1[ok], rtol=1e-9)
    np.testing.assert_allclose(evc.log_probability(gold["theta8"]), gold["logprob"], rtol=1e-6)
    np.testing.assert_array_equal(evc.log_probability(th[5:19]), lpc[5:19])          # batch independence, bitwise
    # spectra and electron distributions, every snapshot
    b = pack_code(th[:9, :6], P)
    outs = []
    for n in (1, cluster):
        eng = Engine(b.cfg, b.W)
        if n > 1:
            eng.set_cluster(n)
        eng.setup(b)
        S = int(b.nsteps.max())
        N, sed = eng.run(snapshots=S)
        outs.append((N.cpu().numpy(), sed.cpu().numpy(), S))
        eng.close()
    for w in range(b.W):
        ns = int(b.nsteps[w])
        assert_spectra_close(outs[1][1][w, :ns], outs[0][1][w, :ns], f"cluster {cluster} nuLnu walker {w}", tol=1e-12)
        assert_spectra_close(outs[1][0][w, :ns], outs[0][0][w, :ns], f"cluster {cluster} n_e walker {w}", tol=1e-12)


@pytest.mark.gpu
def test_gpu_loglike_epilogue_alone(gold):
    """lhm_loglike_batch on the REFERENCE's spectra (isolates the epilogue from the time integration)."""
    import ctypes as C
    import torch
    from lehamoc_b200 import _lib
    from lehamoc_b200._lib import check
    cf, _ = load_golden("code_fit")
    lib = _lib.load()
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    W, Nt, nd = len(gold["theta8"]), cf["nu_tot"].shape[1], len(gold["xd"])
    args = [t(cf["nu_tot"]), t(cf["nuLnu"]), t(gold["theta8"][:, 6]), t(gold["theta8"][:, 7])]
    data = [t(gold[k]) for k in ("xd", "yd", "yerr_hi", "yerr_lo", "flags")]
    logl = torch.empty(W, dtype=torch.float64, device=dev)
    ym = torch.empty((W, nd), dtype=torch.float64, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    check(lib.lhm_loglike_batch(W, Nt, *[p(a) for a in args], float(np.log10(1. + gold["z"])),
                                float(gold["distance_norm"]), nd, *[p(a) for a in data], p(logl), p(ym),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lhm_loglike_batch")
    torch.cuda.synchronize()
    np.testing.assert_allclose(ym.cpu().numpy(), gold["ymodel"], rtol=1e-14, atol=1e-13)
    np.testing.assert_allclose(logl.cpu().numpy(), gold["loglike"], rtol=1e-11)


@pytest.mark.gpu
def test_gpu_mcmc_smoke(gold):
    """A short ensemble run end to end on the GPU: 16 walkers x 8 dims, 3 iterations, finite log-probs."""
    from lehamoc_b200 import fit
    _, P = load_golden("code_fit")
    rng = np.random.default_rng(0)
    init = np.array([*gold["theta8"][0, :7], -2.0])
    pos = init + 1e-2 * rng.standard_normal((16, 8))
    fn = lambda th: fit.log_probability_batch(th, gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"],
                                              gold["flags"], "LeMoC", param_file=P)
    res = fit.run_mcmc(fn, pos, 3, seed=1)
    assert res["chain"].shape == (3, 16, 8) and np.all(np.isfinite(res["log_prob"]))


def test_lehamoc_prior():
    from lehamoc_b200 import fit
    th = [15.5, 1., 0.5, 5., -3., 1.2, 0.5, 7.3, -4.5, 2., 1., -2.]          # fit_emcee.py:212 + log_f
    assert fit.log_prior_LeHaMoC(th) == 0.0
    bad = list(th); bad[7] = 10.5
    assert fit.log_prior_LeHaMoC(bad) == -np.inf
    with pytest.raises(ValueError):
        fit.log_prior_batch(np.zeros((2, 8)), "LeHaMoC")


@pytest.mark.gpu
def test_gpu_loglike_lehamoc_model(gold):
    """flag_c = 'LeHaMoC': Code.LeHaMoC spectra through the same epilogue; checked against the restated likelihood
    evaluated on the REFERENCE's Code.LeHaMoC spectra."""
    from lehamoc_b200 import fit
    z, P = load_golden("code_lh_fit")
    W = len(z["thetas"])
    th12 = np.column_stack([z["thetas"], np.linspace(0.8, 1.6, W), np.linspace(-3., -1., W)])
    ev = fit.FitEvaluator(gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"], param_file=P,
                          D=float(gold["D_Mpc"]), z=float(gold["z"]), flag_c="LeHaMoC")
    logl, ym = ev.model_and_loglike(th12, want_model=True)
    m = orc.Model(lambda th, f: None, D=float(gold["D_Mpc"]), z=float(gold["z"]))
    for k in range(W):
        ym_ref = m.transform(gold["xd"], z["nu_tot"][k], z["nuLnu"][k], th12[k, 10])
        assert np.max(np.abs(ym[k].cpu().numpy() - ym_ref)) <= 4.3e-9
        ll_ref = orc.log_likelihood_from_model(ym_ref, th12[k, 11], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"])
        if np.isinf(ll_ref):                      # an upper limit violated by many sigma: log(1 + erf(-x)) = -inf in both
            assert float(logl[k]) == ll_ref
        else:
            assert abs(float(logl[k]) - ll_ref) <= 1e-6 * abs(ll_ref)
    lp = ev.log_probability(th12)
    assert lp.shape == (W,)
