"""Hadronic stage: photopion / Bethe-Heitler proton loss rates and the Kelner & Aharonian
photopion secondaries.  tests/golden/had_functions.npz holds the outputs of the UNMODIFIED reference functions
(oracle/make_golden_had.py) and the six tables they read."""
import os

import numpy as np
import pytest

import lehamoc_oracle_had as oh
from util import GOLD, assert_spectra_close


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "had_functions.npz"))


def tables_of(gold):
    return {k[len("table_"):]: gold[k] for k in gold.files if k.startswith("table_")}


def rel_close(got, ref, tol, what):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, what
    zero = ref == 0
    assert np.array_equal(got[zero], ref[zero]), f"{what}: exact zeros differ"
    nz = ~zero
    if nz.any():
        scale = np.abs(ref[nz]).max()
        big = nz & (np.abs(ref) > 1e-200 * scale)
        err = np.abs(got[big] - ref[big]) / np.abs(ref[big])
        assert err.max() <= tol, f"{what}: max rel err {err.max():.3e} > {tol:.1e}"
        small = nz & ~big
        assert np.all(np.isfinite(got[small]))


# ---- CPU: the restatement against the live reference ---------------------------------------------------------
@pytest.mark.parametrize("case", [0, 1])
def test_oracle_loss_rates(gold, case):
    T = tables_of(gold)
    p = f"c{case}_"
    rel_close(oh.dg_dt_BH(gold[p + "g_pr"], gold[p + "nu"], gold[p + "photons"], T), gold[p + "dg_dt_BH"], 1e-12, "dg_dt_BH")
    rel_close(oh.dg_dt_pg(gold[p + "g_pr"], gold[p + "nu"], gold[p + "photons"], T), gold[p + "dg_dt_pg"], 1e-12, "dg_dt_pg")


@pytest.mark.parametrize("sp", oh.SPECIES)
def test_oracle_photopion_secondaries(gold, sp):
    T = tables_of(gold)
    p = "c0_"
    got = oh.Qp_g_opt(gold[p + "g_el"], gold[p + "nu_ic"], gold[p + "N_pr"], gold[p + "g_pr"], gold[p + "photons"],
                      gold[p + "nu"], sp, T)
    rel_close(got, gold[p + "Q_" + sp.replace("\b", "b")], 1e-11, f"Qp_g_opt {sp!r}")
    assert np.any(got > 0)


def test_neutrino_grid(gold):
    np.testing.assert_array_equal(oh.E_NU_SPACE, gold["E_nu_space"])


# ---- GPU: the CUDA kernels against the live reference ---------------------------------------------------------
def _load(gold):
    from lehamoc_b200 import hadronic
    T = tables_of(gold)
    hadronic.set_tables({"phi_g": T["Phi_g"], "phi_el": T["Phi_el"], "phi_el1": T["Phi_el_1"], "fk": T["f_k_i"],
                         "cs": T["cross_section"], "kp": T["kp_pg"]})
    return hadronic


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1])
def test_gpu_loss_rates(gold, case):
    had = _load(gold)
    p = f"c{case}_"
    rel_close(had.dg_dt_BH(gold[p + "g_pr"], gold[p + "nu"], gold[p + "photons"]), gold[p + "dg_dt_BH"], 1e-8, "dg_dt_BH")
    rel_close(had.dg_dt_pg(gold[p + "g_pr"], gold[p + "nu"], gold[p + "photons"]), gold[p + "dg_dt_pg"], 1e-8, "dg_dt_pg")


@pytest.mark.gpu
@pytest.mark.parametrize("sp", oh.SPECIES)
def test_gpu_photopion_secondaries(gold, sp):
    had = _load(gold)
    for case in (0, 1):
        p = f"c{case}_"
        got = had.Qp_g_opt(gold[p + "g_el"], gold[p + "nu_ic"], gold[p + "N_pr"], gold[p + "g_pr"], gold[p + "photons"],
                           gold[p + "nu"], sp)
        rel_close(got, gold[p + "Q_" + sp.replace("\b", "b")], 1e-8, f"Qp_g_opt {sp!r} case {case}")


@pytest.mark.gpu
def test_gpu_hadronic_batch_equals_single(gold):
    """Two different walkers in one launch give what each gives alone."""
    had = _load(gold)
    keys = ("g_el", "nu_ic", "N_pr", "g_pr", "photons", "nu")
    both = [np.stack([gold[f"c{c}_" + k] for c in (0, 1)]) for k in keys]
    got = had.Qp_g_opt(*both, "e+")
    for c in (0, 1):
        one = had.Qp_g_opt(*[gold[f"c{c}_" + k] for k in keys], "e+")
        np.testing.assert_array_equal(got[c], one)
    lb = had.dg_dt_pg(np.stack([gold["c0_g_pr"], gold["c1_g_pr"]]), np.stack([gold["c0_nu"], gold["c1_nu"]]),
                      np.stack([gold["c0_photons"], gold["c1_photons"]]))
    np.testing.assert_array_equal(lb[1], had.dg_dt_pg(gold["c1_g_pr"], gold["c1_nu"], gold["c1_photons"]))


# ---- A13 Bethe-Heitler pair injection -------------------------------------------------------------------------
def bh_tck(gold, which):
    return [gold[f"bh_t{which}_tx"], gold[f"bh_t{which}_ty"], gold[f"bh_t{which}_c"], 3, 3]


@pytest.mark.parametrize("case", [0, 1])
def test_oracle_Q_BH_sol(gold, case):
    from lehamoc_oracle import c, h, m_el
    p = f"c{case}_"
    x = gold[p + "nu"] * h / (m_el * c ** 2.)
    got = oh.Q_BH_sol(gold[p + "bh_g_el"], gold[p + "bh_g_pr"], gold[p + "bh_N_pr"], x, gold[p + "photons"],
                      bh_tck(gold, "m"), bh_tck(gold, "p"), float(gold["bh_max_m"]), float(gold["bh_max_p"]))
    rel_close(got, gold[p + "Q_BH"], 1e-12, "Q_BH_sol")
    assert np.any(got > 0)


def test_host_bh_surface_setup_is_the_references(gold):
    """lehamoc_b200.hadronic.interp_cs_BH_int (27 000 Simpson integrals + two bisplrep fits, host) reproduces the
    reference's spline surfaces bit for bit (same SciPy)."""
    from lehamoc_b200 import hadronic
    nu0, nu1 = gold["bh_nu_range"]
    tm, tp, mm, mp_ = hadronic.interp_cs_BH_int(np.logspace(0., 8., 50), np.logspace(0., 10., 50), nu0, nu1,
                                                cache=False)
    for ours, which in ((tm, "m"), (tp, "p")):
        for a, b in zip(ours[:3], bh_tck(gold, which)[:3]):
            np.testing.assert_array_equal(np.asarray(a), b)
    assert mm == float(gold["bh_max_m"]) and mp_ == float(gold["bh_max_p"])


def test_host_bh_surface_disk_cache_round_trip(gold, tmp_path, monkeypatch):
    """The persisted tck (one .npz per grid ends / frequency range / SciPy + NumPy version) reloads bit for bit."""
    from lehamoc_b200 import hadronic
    monkeypatch.setenv("LHM_CACHE_DIR", str(tmp_path))
    nu0, nu1 = gold["bh_nu_range"]
    args = (np.logspace(0., 8., 50), np.logspace(0., 10., 50), nu0, nu1)
    first = hadronic.interp_cs_BH_int(*args)
    assert len(list(tmp_path.glob("bh_surfaces_*.npz"))) == 1
    again = hadronic.interp_cs_BH_int(*args)
    for k in (0, 1):
        for a, b in zip(first[k][:3], again[k][:3]):
            np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
        assert first[k][3:] == again[k][3:] == [3, 3]
    assert first[2:] == again[2:]


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1])
def test_gpu_Q_BH_sol(gold, case):
    """Device spline evaluation + nested Simpson integrals against the reference's Q_BH_sol (SciPy bisplev/simpson)."""
    from lehamoc_b200 import hadronic
    from lehamoc_oracle import c, h, m_el
    p = f"c{case}_"
    x = gold[p + "nu"] * h / (m_el * c ** 2.)
    got = hadronic.Q_BH_sol(gold[p + "bh_g_el"], gold[p + "bh_g_pr"], gold[p + "bh_N_pr"], x, gold[p + "photons"],
                            bh_tck(gold, "m"), bh_tck(gold, "p"), float(gold["bh_max_m"]), float(gold["bh_max_p"]))
    rel_close(got, gold[p + "Q_BH"], 1e-8, "Q_BH_sol")


# ---- A14 pp channels --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gold_pp():
    return np.load(os.path.join(GOLD, "pp_functions.npz"))


PP_FUNCS = (("Q_e_pp", "g_el"), ("Q_g_pp", "nu_ic"), ("Q_nu_mu_pp", "nu_nu"), ("Q_nu_e_pp", "nu_nu"))


def test_oracle_pp_channels(gold_pp):
    g = gold_pp
    np.testing.assert_allclose(oh.cs_pp_inel(g["cs_E"]), g["cs_pp_inel"], rtol=1e-15, atol=0)
    for case in (0, 1):
        for p in (2.01, 2.5):
            tag = f"c{case}_p{int(p * 100)}_"
            for name, grid in PP_FUNCS:
                with np.errstate(all="ignore"):
                    got = np.array(getattr(oh, name)(g[grid], g[f"c{case}_g_pr"], g[f"c{case}_N_pr_TeV"], p, 1.5), dtype=np.float64)
                ref = g[tag + name]
                nz = ref != 0
                assert np.array_equal(got[~nz], ref[~nz]) and np.max(np.abs(got[nz] / ref[nz] - 1)) < 1e-13, (tag, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name,grid", PP_FUNCS)
def test_gpu_pp_channels(gold_pp, name, grid):
    """lhm_pp_batch against the live reference's Q_*_pp (LeHaMoC_f.py:860-1014): 1e-8 relative, zeros exact."""
    from lehamoc_b200 import hadronic as had
    g = gold_pp
    for case in (0, 1):
        for p in (2.01, 2.5):
            tag = f"c{case}_p{int(p * 100)}_"
            got = getattr(had, name)(g[grid], g[f"c{case}_g_pr"], g[f"c{case}_N_pr_TeV"], p, 1.5)
            ref = g[tag + name]
            nz = ref != 0
            assert np.array_equal(got[~nz], ref[~nz]), (tag, name, got[~nz])
            err = np.max(np.abs(got[nz] / ref[nz] - 1))
            assert err < 1e-8, (tag, name, err, np.argmax(np.abs(got[nz] / ref[nz] - 1)))


@pytest.mark.gpu
def test_gpu_pp_batch_equals_single(gold_pp):
    from lehamoc_b200 import hadronic as had
    g = gold_pp
    gp = np.stack([g["c0_g_pr"], g["c1_g_pr"]]); NT = np.stack([g["c0_N_pr_TeV"], g["c1_N_pr_TeV"]])
    both = had.Q_nu_mu_pp(np.stack([g["nu_nu"]] * 2), gp, NT, np.array([2.01, 2.5]), np.array([1.5, 1.5]))
    for c, p in ((0, 2.01), (1, 2.5)):
        np.testing.assert_array_equal(both[c], had.Q_nu_mu_pp(g["nu_nu"], gp[c], NT[c], p, 1.5))


def _random_fields(seed):
    """Same construction as oracle/make_golden_had.py::fields (photon field with a synchrotron bump and a sentinel-level
    tail, cut-off power-law protons), other seeds, plus pp inputs."""
    rng = np.random.default_rng(seed)
    nu = np.logspace(7.5, 32., int(rng.integers(50, 90)))
    l = np.log10(nu)
    a1, a2 = rng.uniform(0.4, 0.9), rng.uniform(0.3, 0.8)
    logn = rng.uniform(2., 6.) - 1.5 * (l - 12.) - a1 * np.maximum(l - rng.uniform(14., 17.), 0.) ** 1.5 + 0.3 * np.sin(a2 * l)
    photons = 10 ** logn
    photons[-3:] = 10 ** (-260.)
    g_pr = np.logspace(0., rng.uniform(7., 8.5), int(rng.integers(25, 45)))
    p = rng.uniform(1.7, 2.4)
    N_pr = 10 ** rng.uniform(-2., 2.) * g_pr ** (-p) * np.exp(-g_pr / g_pr[-4])
    N_pr[0] = N_pr[-1] = 10 ** (-260.)
    g_el = np.logspace(0., rng.uniform(9., 11.), int(rng.integers(16, 30)))
    nu_ic = np.logspace(15., 32., int(rng.integers(12, 24)))
    return nu, photons, g_pr, N_pr, g_el, nu_ic, float(p)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [301, 302, 303])
def test_gpu_hadronic_functions_random_inputs_vs_oracle(gold, seed):
    """Other photon fields, proton distributions and grid sizes than the fixtures: the device functions against the
    oracle restatement (pinned on the live reference functions by the tests above)."""
    had = _load(gold)
    T = tables_of(gold)
    nu, photons, g_pr, N_pr, g_el, nu_ic, p = _random_fields(seed)
    rel_close(had.dg_dt_BH(g_pr, nu, photons), oh.dg_dt_BH(g_pr, nu, photons, T), 1e-8, f"dg_dt_BH seed {seed}")
    rel_close(had.dg_dt_pg(g_pr, nu, photons), oh.dg_dt_pg(g_pr, nu, photons, T), 1e-8, f"dg_dt_pg seed {seed}")
    for sp in ("2_g", "e+", "\bar_nu_e"):
        rel_close(had.Qp_g_opt(g_el, nu_ic, N_pr, g_pr, photons, nu, sp), oh.Qp_g_opt(g_el, nu_ic, N_pr, g_pr, photons, nu, sp, T),
                  1e-8, f"Qp_g_opt {sp!r} seed {seed}")
    NT = N_pr / (oh.m_pr * oh.c ** 2. * 0.624151)
    nu_nu = oh.E_NU_SPACE / oh.h
    with np.errstate(all="ignore"):
        for name, grid in (("Q_e_pp", g_el), ("Q_g_pp", nu_ic), ("Q_nu_mu_pp", nu_nu), ("Q_nu_e_pp", nu_nu)):
            ref = np.array(getattr(oh, name)(grid, g_pr, NT, p, 7.5), dtype=np.float64)
            got = getattr(had, name)(grid, g_pr, NT, p, 7.5)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (name, seed)
            ok = ~np.isnan(ref)
            rel_close(got[ok], ref[ok], 1e-8, f"{name} seed {seed}")
