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
