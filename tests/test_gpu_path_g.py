"""IC path G (lhm_icg.cuh): shared-grid batches advance in lock-step and the IC emissivity is a dense contraction over
walkers on the FP64 tensor cores.  Same triple sum as path R (LeHaMoC_f.py:113-124), different summation order: parity
against the live-reference fixtures at the north_star tolerance, and against path R at 1e-12 in log10."""
import numpy as np
import pytest

from lehamoc_b200._lib import LhmError

from util import TOL_LOG10, assert_spectra_close, load_golden, quiet

pytestmark = pytest.mark.gpu
quiet()


def _V(R):
    return 4. * np.pi / 3. * R ** 3.


def test_ic_emissivity_steps_path_G():
    """Function level: Q_IC_space_optimized on the captured Test-1 state of every step, three walkers that share the
    grids (the captured state and two rescaled copies: the contraction is linear in both N_e and the photon field)."""
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    z, P = load_golden("test1")
    b = pack_script([P, P, P])
    assert b.cfg["shared_ic_grids"] == 1
    eng = Engine(b.cfg, 3, ic_path="G")
    eng.setup(b)
    engR = Engine(b.cfg, 3, ic_path="R")        # a path-G handle builds no per-walker tile schedule: R needs its own
    engR.setup(b)
    with pytest.raises(LhmError):
        eng.ic_emissivity(np.ones((3, b.cfg["Ng"])), np.ones((3, b.cfg["Nt"])), "R")
    R0 = 10 ** P["R0"]
    sent = 10 ** (-260.)
    for s in range(1, 10):
        Rad = R0 + P["Vexp"] * 2.99792458e10 * ((s + 1) * float(z["dt"]))
        N = np.concatenate([[sent], z["step_N_el_int"][s], [sent]]) / _V(Rad)
        n = z["step_photons_top_raw"][s] / _V(Rad)
        Ns = np.stack([N, 2. * N, N])
        ns = np.stack([n, n, 4. * n])
        Q = eng.ic_emissivity(Ns, ns, "G").cpu().numpy()
        QR = engR.ic_emissivity(Ns, ns, "R").cpu().numpy()
        ref = z["step_Q_IC"][s]
        assert_spectra_close(Q[0], ref, f"Q_IC path G step {s + 1}", tol=1e-10)
        assert_spectra_close(Q[1], 2. * ref, f"Q_IC path G step {s + 1} (2 N_e)", tol=1e-10)
        assert_spectra_close(Q[2], 4. * ref, f"Q_IC path G step {s + 1} (4 n)", tol=1e-10)
        assert_spectra_close(Q, QR, f"Q_IC path G vs path R step {s + 1}", tol=1e-12)
    eng.close()
    engR.close()


@pytest.mark.parametrize("name", ["test1", "var_bb", "var_expand", "var_nogg", "var_user"])
def test_script_runs_match_reference_path_G(name):
    """Whole runs, every step, against the unmodified LeMoC.py: walkers 0 and 2 of a shared-grid batch are the fixture's
    parameter set, walker 1 a perturbed one (checked against path R)."""
    from lehamoc_b200 import LeMoC
    z, P = load_golden(name)
    sets = [dict(P), dict(P, B0=P["B0"] * 1.7, p_el=P["p_el"] + 0.2, L_el=P["L_el"] - 0.3), dict(P)]
    batch, N, sed = LeMoC.run(sets, ic_path="G")
    _, NR, sedR = LeMoC.run(sets, ic_path="R")
    ns = int(z["nsteps"])
    assert N.shape == (3, ns, len(z["g_el"])) and sed.shape == (3, ns, len(z["nu_tot"]))
    for w in (0, 2):
        for s in range(ns):
            assert_spectra_close(N[w, s], 10 ** z["out_log_ne"][s], f"{name}/G n_e walker {w} step {s + 1}", tol=TOL_LOG10)
            assert_spectra_close(sed[w, s], 10 ** z["out_log_nuLnu"][s], f"{name}/G nuLnu walker {w} step {s + 1}", tol=TOL_LOG10)
    np.testing.assert_array_equal(sed[0], sed[2])
    assert_spectra_close(sed, sedR, f"{name}: path G vs path R nuLnu", tol=1e-11)
    assert_spectra_close(N, NR, f"{name}: path G vs path R n_e", tol=1e-11)


@pytest.mark.parametrize("W", [2, 31, 32, 33, 100])
def test_path_G_batch_sizes_and_determinism(W):
    """Walker counts around the 32-walker blocks of the contraction: every walker equals path R within 1e-11, the result
    does not depend on the batch it is evolved in (bit for bit), and a second run gives the same bits."""
    from lehamoc_b200 import LeMoC
    from lehamoc_b200.presets import test1_walkers
    sets = [dict(P, grid_g_el=100., grid_nu=70., time_end=4.) for P in test1_walkers(W, seed=3)]
    _, N, sed = LeMoC.run(sets, ic_path="G", snapshots=False)
    _, NR, sedR = LeMoC.run(sets, ic_path="R", snapshots=False)
    assert_spectra_close(sed, sedR, f"W={W}: path G vs path R nuLnu", tol=1e-11)
    assert_spectra_close(N, NR, f"W={W}: path G vs path R n_e", tol=1e-11)
    _, N2, sed2 = LeMoC.run(sets, ic_path="G", snapshots=False)
    np.testing.assert_array_equal(sed2, sed)
    np.testing.assert_array_equal(N2, N)
    if W >= 4:
        _, N3, sed3 = LeMoC.run(sets[1:W // 2 + 1], ic_path="G", snapshots=False)
        np.testing.assert_array_equal(sed3, sed[1:W // 2 + 1])
        np.testing.assert_array_equal(N3, N[1:W // 2 + 1])


@pytest.mark.parametrize("grid", [61, 87, 101, 122])
def test_odd_gamma_grids_all_paths(grid):
    """Grid sizes that are not multiples of two / four (the shipped LeHaMoC Parameters.txt has 183 gammas): the row phases
    deal gamma pairs to two lanes per row and end on a partial pair.  Every step against the CPU oracle, all three IC paths,
    walker 0 of a shared-grid batch; the other walkers path G against path R."""
    import lehamoc_oracle
    from lehamoc_b200 import LeMoC
    from lehamoc_b200.presets import test1_walkers
    sets = [dict(P, grid_g_el=float(grid), grid_nu=57., time_end=3.) for P in test1_walkers(5, seed=11)]
    ge, ne_ref, sed_ref = lehamoc_oracle.run_script(sets[0])
    out = {path: LeMoC.run(sets, ic_path=path) for path in ("G", "R", "S")}
    for path, (_, N, sed) in out.items():
        assert N.shape[2] == grid == len(ne_ref[0]) and N.shape[1] == len(ne_ref)
        for s in range(len(ne_ref)):
            assert_spectra_close(N[0, s], ne_ref[s], f"grid {grid} path {path} n_e step {s + 1}")
            assert_spectra_close(sed[0, s], sed_ref[s], f"grid {grid} path {path} nuLnu step {s + 1}")
    assert_spectra_close(out["G"][2], out["R"][2], f"grid {grid}: path G vs path R nuLnu", tol=1e-11)
    assert_spectra_close(out["G"][1], out["R"][1], f"grid {grid}: path G vs path R n_e", tol=1e-11)


def test_path_G_full_size_batch_and_device_packing():
    """The bench's workload: 1184 perturbed Test-1 walkers; walker 0 is Parameters_Test1.txt (live-reference fixture),
    all walkers against path R; the pipelined device-packed API (LeMoC.run_many) gives the same spectra within the
    parity tolerance."""
    from lehamoc_b200 import LeMoC
    from lehamoc_b200.presets import test1_walkers
    z, P = load_golden("test1")
    sets = test1_walkers(1184)
    _, N, sed = LeMoC.run(sets, ic_path="G", snapshots=False)
    _, NR, sedR = LeMoC.run(sets, ic_path="R", snapshots=False)
    assert_spectra_close(sed[0, 0], 10 ** z["out_log_nuLnu"][-1], "path G walker 0 vs LeMoC.py", tol=TOL_LOG10)
    assert_spectra_close(sed, sedR, "1184 walkers: path G vs path R nuLnu", tol=1e-11)
    assert_spectra_close(N, NR, "1184 walkers: path G vs path R n_e", tol=1e-11)
    got = [(Nh.copy(), sh.copy()) for _, Nh, sh in LeMoC.run_many(iter([sets, sets[:500]]), ic_path="G")]
    assert_spectra_close(got[0][1], sed[:, 0], "run_many path G nuLnu")
    assert_spectra_close(got[1][1], sed[:500, 0], "run_many path G nuLnu (second batch)")
    assert_spectra_close(got[0][0], N[:, 0], "run_many path G n_e")


@pytest.mark.parametrize("name", ["lh_shipped", "lh_bbpl", "lh_expand", "lh_nogg", "lh_user", "lh_bbpl200", "lh_bbpl400",
                                  "lh_bbpl800"])
def test_lehamoc_driver_path_G(name):
    """Path G under the LeHaMoC.py driver (hadronic flags 0; BASELINE config 5 is such a fixed-grid sweep): walkers 0 and 2
    of a shared-grid batch are the fixture's parameter set (every step against the unmodified script), walker 1 is
    perturbed and checked against path R."""
    from lehamoc_b200 import LeHaMoC
    from test_lehamoc_driver import log_close
    z, P = load_golden(name)
    sets = [dict(P), dict(P, B0=P["B0"] * 1.7, p_el=P["p_el"] + 0.2, L_el=P["L_el"] - 0.3, L_pr=P["L_pr"] + 0.4), dict(P)]
    batch, N, sed, Npr, Nnu = LeHaMoC.run(sets, ic_path="G")
    _, NR, sedR, NprR, _ = LeHaMoC.run(sets, ic_path="R")
    assert batch.cfg["shared_ic_grids"] == 1 and not Nnu.any()
    for w in (0, 2):
        log_close(N[w], z["n_e"], f"{name}/G n_e walker {w}")
        log_close(sed[w], z["nuLnu"], f"{name}/G nuLnu walker {w}")
        log_close(Npr[w], z["n_p"], f"{name}/G n_p walker {w}")
    np.testing.assert_array_equal(sed[0], sed[2])
    with np.errstate(divide="ignore"):
        log_close(sed, np.log10(sedR), f"{name}: path G vs path R nuLnu", tol=1e-11)
        log_close(N, np.log10(NR), f"{name}: path G vs path R n_e", tol=1e-11)
        log_close(Npr, np.log10(NprR), f"{name}: path G vs path R n_p", tol=1e-11)


def test_lehamoc_driver_path_G_limits(monkeypatch):
    """A hadronic flag is refused outright.  The contraction's tiles (32 walkers x target frequencies + gammas x 32) are cut
    into k- and gamma-sections when they outgrow shared memory (grid 800 at the full budget, lh_bbpl800 above); a small
    budget forces sections at every size and must not change a bit of the result (every (task, section) has its own
    slot of the partial-sum buffer, added in a fixed order -- but the order differs from the unsectioned one: 1e-12)."""
    from lehamoc_b200 import LeHaMoC, _lib
    from test_lehamoc_driver import log_close
    z, P = load_golden("lh_bbpl200")
    with pytest.raises(_lib.LhmError):
        LeHaMoC.run([dict(P, pg_pi_l_flag=1.)] * 2, ic_path="G")
    sets = [dict(P), dict(P, B0=P["B0"] * 1.3, p_el=P["p_el"] + 0.1)]
    _, N0, sed0, _, _ = LeHaMoC.run(sets, ic_path="G")
    monkeypatch.setenv("LHM_ICG_SMEM_KB", "48")
    _, N1, sed1, _, _ = LeHaMoC.run(sets, ic_path="G")
    log_close(sed1[0], z["nuLnu"], "sectioned contraction vs LeHaMoC.py")
    with np.errstate(divide="ignore"):
        log_close(sed1, np.log10(sed0), "sectioned vs whole-grid contraction nuLnu", tol=1e-12)
        log_close(N1, np.log10(N0), "sectioned vs whole-grid contraction n_e", tol=1e-12)


def test_path_G_refuses_what_it_cannot_do():
    from lehamoc_b200 import LeMoC, _lib
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    _, P = load_golden("test1")
    P.update(grid_g_el=60., grid_nu=40., time_end=2.)
    with pytest.raises(_lib.LhmError):                   # one walker alone has no shared grids to exploit
        LeMoC.run(P, ic_path="G")
    ragged = pack_script([dict(P), dict(P, g_max_el=5.9, g_PL_max=5.5)])
    with pytest.raises(_lib.LhmError):
        Engine(ragged.cfg, 2, ic_path="G")
    # IC emission switched off: path G has nothing to contract and runs the fused kernel
    sets = [dict(P, IC_emis_flag=0.), dict(P, IC_emis_flag=0., B0=2.)]
    _, N, sed = LeMoC.run(sets, ic_path="G")
    _, NR, sedR = LeMoC.run(sets, ic_path="R")
    np.testing.assert_array_equal(sed, sedR)
