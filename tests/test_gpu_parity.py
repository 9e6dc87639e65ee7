"""GPU parity tests (run on a B200: `pytest -m gpu`).  Every comparison goes through the C ABI of liblhm.so;
the checker is the NumPy oracle and the golden vectors captured from the unmodified reference."""
import numpy as np
import pytest

from util import (TOL_LOG10, assert_signed_close, assert_spectra_close, load_golden, quiet)

pytestmark = pytest.mark.gpu
quiet()


@pytest.fixture(scope="module")
def orc():
    import lehamoc_oracle
    return lehamoc_oracle


def _run_script(P, ic_path="R", snapshots=True):
    from lehamoc_b200 import LeMoC
    return LeMoC.run(P, ic_path=ic_path, snapshots=snapshots)


# ---------------------------------------------------------------------------------------------------
# function-level parity on the captured Test-1 state (every step of the reference run)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def test1_engine():
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    z, P = load_golden("test1")
    b = pack_script(P)
    eng = Engine(b.cfg, 1, ic_path="R")
    eng.setup(b)
    return eng, z, P, b


def _V(R):
    return 4. * np.pi / 3. * R ** 3.


def test_photons_tot_steps(test1_engine):
    eng, z, P, b = test1_engine
    R0 = 10 ** P["R0"]
    sent = 10 ** (-260.)
    for s in range(1, 10):
        ps = np.concatenate([[sent], z["step_photons_syn_int"][s - 1], [sent]])
        pi = np.concatenate([[sent], z["step_photons_IC_int"][s - 1], [sent]])
        Rad = R0 + P["Vexp"] * 2.99792458e10 * ((s + 1) * float(z["dt"]))
        n = eng.photons_tot(ps[None], pi[None], np.array([Rad])).cpu().numpy()[0]
        ref = z["step_photons_top_raw"][s] / _V(Rad)
        assert_spectra_close(n, ref, f"photons_tot step {s + 1}")


def test_U_ph_steps(test1_engine):
    eng, z, P, b = test1_engine
    R0 = 10 ** P["R0"]
    for s in (0, 1, 4, 9):
        Rad = R0 + P["Vexp"] * 2.99792458e10 * ((s + 1) * float(z["dt"]))
        n = z["step_photons_top_raw"][s] / _V(Rad)
        U = eng.U_ph(n[None], np.array([Rad])).cpu().numpy()[0]
        assert_spectra_close(U, z["step_U_ph"][s], f"U_ph step {s + 1}")


def test_syn_ssa_steps(test1_engine):
    eng, z, P, b = test1_engine
    R0 = 10 ** P["R0"]
    sent = 10 ** (-260.)
    for s in (0, 1, 4, 9):
        Rad = R0 + P["Vexp"] * 2.99792458e10 * ((s + 1) * float(z["dt"]))
        N = np.concatenate([[sent], z["step_N_el_int"][s], [sent]])
        Q, aS, aI = [t.cpu().numpy()[0] for t in eng.syn_ssa((N / _V(Rad))[None], np.array([P["B0"]]))]
        assert_spectra_close(Q, z["step_Q_syn_raw"][s], f"Q_syn step {s + 1}")
        assert_signed_close(-np.abs(aS), z["step_aSSA_syn"][s], f"aSSA(nu_syn) step {s + 1}")
        assert_signed_close(-np.abs(aI), z["step_aSSA_ic"][s], f"aSSA(nu_ic) step {s + 1}")


@pytest.mark.parametrize("path", ["R", "S"])
def test_ic_emissivity_steps(test1_engine, path):
    eng, z, P, b = test1_engine
    R0 = 10 ** P["R0"]
    sent = 10 ** (-260.)
    for s in range(0, 10):
        Rad = R0 + P["Vexp"] * 2.99792458e10 * ((s + 1) * float(z["dt"]))
        N = np.concatenate([[sent], z["step_N_el_int"][s], [sent]])
        n = z["step_photons_top_raw"][s] / _V(Rad)
        Q = eng.ic_emissivity((N / _V(Rad))[None], n[None], path).cpu().numpy()[0]
        ref = z["step_Q_IC"][s]
        # step 1 has only sentinel photons: values ~1e-320 are denormal noise in both codes
        if s == 0:
            assert np.all(np.isfinite(Q)) and np.all(Q[ref == 0] == 0)
            continue
        assert_spectra_close(Q, ref, f"Q_IC path {path} step {s + 1}", tol=1e-10)


def test_gg_steps(test1_engine):
    eng, z, P, b = test1_engine
    R0 = 10 ** P["R0"]
    for s in (1, 4, 9):
        Rad = R0 + P["Vexp"] * 2.99792458e10 * ((s + 1) * float(z["dt"]))
        n = z["step_photons_top_raw"][s] / _V(Rad)
        a, Qee = [t.cpu().numpy()[0] for t in eng.gg(n[None])]
        assert_spectra_close(a, z["step_a_gg"][s], f"a_gg step {s + 1}", tol=1e-10)
        assert_spectra_close(Qee, z["step_Q_ee"][s], f"Q_ee step {s + 1}", tol=1e-10)


def test_cor_factor(test1_engine, golden_dir):
    eng, z, P, b = test1_engine
    cor = float(eng.cor_factor()[0].item())
    assert abs(cor / float(z["cor_factor"]) - 1.) < 1e-11


def test_function_twins(orc):
    """lehamoc_b200.LeHaMoC_f mirrors the reference's function library call for call."""
    from lehamoc_b200 import LeHaMoC_f as f
    fz = np.load(__import__("os").path.join(__import__("util").GOLD, "functions.npz"))
    z, P = load_golden("test1")
    g = z["g_el"]
    got = f.cor_factor_syn_el(g, 1e15, 1e4, 1.9, f.Lum_e_inj(1e-3, 1e15))
    assert abs(got / float(fz["cor_factor_test1"]) - 1.) < 1e-11
    gfit = np.logspace(2., 7.081445, 160)
    got = f.cor_factor_syn_el(gfit, 10 ** 15.25808, 1e4, 2.0, f.Lum_e_inj(-4., 10 ** 15.25808))
    assert abs(got / float(fz["cor_factor_fitgrid"]) - 1.) < 1e-11
    x = f.thomas(np.zeros(17), fz["thomas_in_b"], fz["thomas_in_c"], fz["thomas_in_d"])
    np.testing.assert_allclose(x, fz["thomas_out_bidiag"], rtol=1e-14)
    x = f.thomas(fz["thomas_in_a"], fz["thomas_in_b"], fz["thomas_in_c"], fz["thomas_in_d"])      # general tridiagonal
    np.testing.assert_array_equal(x, fz["thomas_out_tridiag"])         # same operations in the same order, no FMA
    # per-frequency functions against the oracle's literal restatement on a synthetic state
    rng = np.random.default_rng(3)
    gs = np.logspace(0., 5., 40)
    Np = 1e3 * gs ** -2.2 * np.exp(rng.normal(0, 0.1, 40))
    Bf = 0.7
    a_cr = 3. * f.q * Bf / (4. * np.pi * f.m_el * f.c)
    for nu in (1e9, 3e12, 1e16):
        assert abs(f.Q_syn_space(Np, Bf, nu, a_cr, gs) / orc.Q_syn_space(Np.copy(), Bf, nu, a_cr, gs) - 1.) < 1e-11
        dgl = np.log(gs[1]) - np.log(gs[0])
        assert abs(f.aSSA(Np, Bf, nu, gs, dgl) / orc.aSSA(Np, Bf, nu, gs, dgl) - 1.) < 1e-10
    nu_t = np.logspace(8., 22., 30)
    ph = 1e5 * (nu_t / 1e12) ** -1.5
    for nu in (1e15, 1e20, 3e23):
        ref = orc.Q_IC_space_optimized(Np, gs, nu, ph, nu_t, len(nu_t) - 1)
        for path in ("R", "S"):
            got = f.Q_IC_space_optimized(Np, gs, nu, ph, nu_t, len(nu_t) - 1, ic_path=path)
            assert abs(got / ref - 1.) < 1e-11, (nu, path)
    nu_ic = np.logspace(18., 26., 12)
    np.testing.assert_allclose(f.a_gg(nu_ic, nu_t, ph), orc.a_gg(nu_ic, nu_t, ph), rtol=1e-10)
    pic = 1e-3 * (nu_ic / 1e20) ** -2.
    np.testing.assert_allclose(f.Q_ee_f(nu_t, ph, nu_ic, pic, gs, 1e15), orc.Q_ee_f(nu_t, ph, nu_ic, pic, gs, 1e15),
                               rtol=1e-10)
    np.testing.assert_allclose(f.U_ph_f(gs, nu_t, ph, 1e15), orc.U_ph_f(gs, nu_t, ph, 1e15), rtol=1e-11)


# ---------------------------------------------------------------------------------------------------
# whole time integrations against the golden vectors of the unmodified reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["test1", "var_bb", "var_expand", "var_nogg", "var_synonly", "var_user"])
@pytest.mark.parametrize("path", ["R", "S"])
def test_script_runs_match_reference(name, path):
    z, P = load_golden(name)
    batch, N, sed = _run_script(P, ic_path=path)
    ns = int(z["nsteps"])
    assert N.shape == (1, ns, len(z["g_el"])) and sed.shape == (1, ns, len(z["nu_tot"]))
    for s in range(ns):
        assert_spectra_close(N[0, s], 10 ** z["out_log_ne"][s], f"{name}/{path} n_e step {s + 1}", tol=TOL_LOG10)
        assert_spectra_close(sed[0, s], 10 ** z["out_log_nuLnu"][s], f"{name}/{path} nuLnu step {s + 1}", tol=TOL_LOG10)
    # the text files only carry ~16 digits of log10: compare the raw state of the final step tightly
    sent = 10 ** (-260.)


def test_test1_final_state_tight():
    """Test 1 (BASELINE config 1) final state vs the reference's in-memory arrays at <= 1e-8 relative."""
    z, P = load_golden("test1")
    for path in ("R", "S"):
        batch, N, sed = _run_script(P, ic_path=path, snapshots=False)
        V = 4. * np.pi / 3. * (10 ** P["R0"] + P["Vexp"] * 2.99792458e10 * 10 * float(z["dt"])) ** 3
        assert_spectra_close(N[0, 0], z["final_N_el"] / V, f"Test1 n_e path {path}", tol=TOL_LOG10)
        assert_spectra_close(sed[0, 0], 10 ** z["out_log_nuLnu"][-1], f"Test1 nuLnu path {path}", tol=TOL_LOG10)


@pytest.mark.parametrize("name", ["test3_synadi", "test3_adiabatic"])
def test_long_runs_match_reference(name):
    """LeMoC/Parameters.txt (100 steps, syn+SSA+adiabatic: the run behind the shipped *_Test3_LeMoC.txt) and
    Parameters_Test3.txt (600 steps, adiabatic only)."""
    z, P = load_golden(name)
    batch, N, sed = _run_script(P)
    keep = z["kept_steps"]
    for row, s in enumerate(keep):
        assert_spectra_close(N[0, s], 10 ** z["out_log_ne"][row], f"{name} n_e step {s + 1}", tol=TOL_LOG10)
        ref = 10 ** z["out_log_nuLnu"][row]
        if np.all(ref < 1e-150):
            # adiabatic-only run: the photon field is the 1e-260 sentinels decaying through the denormal
            # range to exact zeros (the reference prints -inf there).  Normal-range bins must agree to 1e-6
            # in log10; bins the reference has below 1e-300 (few or no significant bits) must be as negligible
            lref = z["out_log_nuLnu"][row]
            normal = lref > -300.
            got = sed[0, s]
            assert np.all(got >= 0.) and np.all(np.isfinite(got))
            np.testing.assert_allclose(np.log10(got[normal]), lref[normal], atol=1e-6)
            assert np.all(got[~normal] < 1e-295)
        else:
            assert_spectra_close(sed[0, s], ref, f"{name} nuLnu step {s + 1}", tol=TOL_LOG10)


def test_code_lemoc_matches_reference():
    """Fit_emcee/Code.py::LeMoC at the fit start point and 7 perturbed walkers (BASELINE config 4)."""
    from lehamoc_b200 import Code
    z, P = load_golden("code_fit")
    nu, sed = Code.LeMoC_batch(z["thetas"], P)
    for i in range(len(z["thetas"])):
        np.testing.assert_array_equal(nu[i], z["nu_tot"][i])
        assert_spectra_close(sed[i], z["nuLnu"][i], f"Code.LeMoC walker {i}")
    nu1, sed1 = Code.LeMoC(list(z["thetas"][0]), P)
    np.testing.assert_array_equal(sed1, sed[0])              # batch == single evaluation, bit for bit
    nuS, sedS = Code.LeMoC_batch(z["thetas"], P, ic_path="S")
    for i in range(len(z["thetas"])):
        assert_spectra_close(sedS[i], z["nuLnu"][i], f"Code.LeMoC walker {i} path S")


def test_batch_of_walkers_vs_oracle(orc):
    """Heterogeneous walkers in one launch: each must equal its own single-model oracle run."""
    from lehamoc_b200 import LeMoC
    _, P = load_golden("test1")
    P.update(grid_g_el=80., grid_nu=50., time_end=4.)
    rng = np.random.default_rng(20231)
    sets = []
    for i in range(12):
        Q = dict(P)
        Q.update(R0=rng.uniform(14.5, 16), B0=10 ** rng.uniform(-1, 1), L_el=40.4857 + rng.uniform(-1, 1),
                 p_el=rng.uniform(1.6, 2.6), g_PL_min=rng.uniform(0.5, 3), g_PL_max=rng.uniform(5, 6.0))
        sets.append(Q)
    batch, N, sed = LeMoC.run(sets, snapshots=False)
    for i, Q in enumerate(sets):
        _, ne_ref, sed_ref = orc.run_script(Q)
        assert_spectra_close(N[i, 0], ne_ref[-1], f"walker {i} n_e")
        assert_spectra_close(sed[i, 0], sed_ref[-1], f"walker {i} nuLnu")


def test_full_size_properties():
    """BASELINE-size batch (Test-1 grid, 296 walkers): independence of walkers and determinism --
    identical walkers give bit-identical spectra wherever they sit in the batch, and a repeat run is
    bit-identical."""
    from lehamoc_b200 import LeMoC
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    _, P = load_golden("test1")
    sets = []
    for i in range(296):
        Q = dict(P)
        if i % 4:
            Q.update(B0=1.0 + 0.01 * (i % 7), p_el=1.9 + 0.001 * (i % 5))
        sets.append(Q)
    b = pack_script(sets)
    eng = Engine(b.cfg, b.W)
    N1, s1 = eng.run_host(b)
    N1, s1 = N1.copy(), s1.copy()
    N2, s2 = eng.run_host(b)
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(N1, N2)
    base = s1[0]
    for i in range(0, 296, 4):
        np.testing.assert_array_equal(s1[i], base)
    z, _ = load_golden("test1")
    assert_spectra_close(base, 10 ** z["out_log_nuLnu"][-1], "walker 0 of the big batch", tol=TOL_LOG10)


def test_cli_text_format(tmp_path, monkeypatch):
    """python -m lehamoc_b200.LeMoC writes the reference's text format (LeMoC.py:290-296)."""
    from lehamoc_b200 import LeMoC
    z, P = load_golden("test1")
    pf = tmp_path / "Parameters_Test1.txt"
    pf.write_text("\n".join(f"{k} = {v!r}" for k, v in P.items()))
    monkeypatch.chdir(tmp_path)
    assert LeMoC.main(["LeMoC.py", str(pf), "t1"]) == 0
    part = np.loadtxt(tmp_path / "Particles_Distribution_t1.txt")
    phot = np.loadtxt(tmp_path / "Photons_Distribution_t1.txt")
    assert part.shape == (10 * 260, 2) and phot.shape == (10 * 100, 2)
    ref1 = np.array([[float(t) for t in ln.split()] for ln in str(z["photons_txt_step1"]).strip().split("\n")])
    np.testing.assert_array_equal(phot[:100, 0], ref1[:, 0])
    np.testing.assert_allclose(phot[:100, 1], ref1[:, 1], atol=1e-7)
    assert LeMoC.main(["LeMoC.py"]) == 1


def test_error_paths():
    from lehamoc_b200 import _lib
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    _, P = load_golden("test1")
    P.update(grid_g_el=40., grid_nu=30.)
    b = pack_script(P)
    eng = Engine(b.cfg, 1)
    with pytest.raises(_lib.LhmError):
        eng.run()                                   # run before setup -> LHM_ERR_STATE
    b2 = pack_script([P, P])
    with pytest.raises(_lib.LhmError):
        eng.setup(b2)                               # over capacity
    with pytest.raises(_lib.LhmError):
        Engine(dict(b.cfg, Ng=2), 1)                # invalid size


def test_ic_pair_table_modes_are_bit_identical():
    """The IC pair table (per walker, or one per batch when the grids are shared) holds exactly what the kernel
    would recompute: all three modes must give the same bits.  A shared-grid engine refuses a ragged batch."""
    from lehamoc_b200 import _lib
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    _, P = load_golden("test1")
    P.update(grid_g_el=90., grid_nu=60., time_end=3.)
    sets = [dict(P, B0=1.0 + 0.1 * i, p_el=1.8 + 0.02 * i) for i in range(5)]
    b = pack_script(sets)
    assert b.cfg["shared_ic_grids"] == 1 and b.cfg["ic_pair_table"] == 1
    outs = []
    for cfg in (b.cfg, dict(b.cfg, shared_ic_grids=0), dict(b.cfg, shared_ic_grids=0, ic_pair_table=0)):
        eng = Engine(cfg, b.W)
        N, sed = eng.run_host(b)
        outs.append((N.copy(), sed.copy()))
        eng.close()
    for N, sed in outs[1:]:
        np.testing.assert_array_equal(sed, outs[0][1])
        np.testing.assert_array_equal(N, outs[0][0])
    ragged = pack_script([dict(P), dict(P, g_max_el=5.9, g_PL_max=5.5)])
    assert ragged.cfg["shared_ic_grids"] == 0
    eng = Engine(dict(ragged.cfg, shared_ic_grids=1), 2)
    with pytest.raises(_lib.LhmError):
        eng.run_host(ragged)
    with pytest.raises(_lib.LhmError):
        eng.setup(ragged)


@pytest.mark.parametrize("device_pack", [False, True])
def test_run_many_matches_run(device_pack):
    """The pipelined throughput API gives, batch by batch and in order, what the synchronous call gives: exactly with
    host packing, within the parity tolerance when the grids are built on the device."""
    from lehamoc_b200 import LeMoC
    _, P = load_golden("test1")
    P.update(grid_g_el=70., grid_nu=40., time_end=3.)
    batches = [[dict(P, B0=0.5 + 0.1 * (i + 3 * k), p_el=1.7 + 0.05 * i) for i in range(3 + k)] for k in range(5)]
    ref = [LeMoC.run(b, snapshots=False) for b in batches]
    got = [(b.W, N.copy(), sed.copy(), np.array(b.nu_tot), np.array(b.g_el))
           for b, N, sed in LeMoC.run_many(iter(batches), device_pack=device_pack)]
    assert [g[0] for g in got] == [len(b) for b in batches]
    for (_, N, sed, nu, g), (br, Nr, sedr) in zip(got, ref):
        np.testing.assert_array_equal(nu, br.nu_tot)
        np.testing.assert_array_equal(g, br.g_el)
        if device_pack:
            assert_spectra_close(sed, sedr[:, 0], "run_many(device_pack) nuLnu")
            assert_spectra_close(N, Nr[:, 0], "run_many(device_pack) n_e")
        else:
            np.testing.assert_array_equal(sed, sedr[:, 0])
            np.testing.assert_array_equal(N, Nr[:, 0])


@pytest.mark.parametrize("device_pack", [False, True])
def test_run_many_results_outlive_the_next_draws(device_pack):
    """The yielded arrays are views of page-locked buffers: held WITHOUT copying, result k must still be intact after
    `depth` further batches have been drawn (the docstring's promise; each engine alternates two buffer pairs)."""
    from lehamoc_b200 import LeMoC
    _, P = load_golden("test1")
    P.update(grid_g_el=70., grid_nu=40., time_end=3.)
    depth = 2
    batches = [[dict(P, B0=0.5 + 0.07 * (i + 3 * k), p_el=1.7 + 0.05 * i) for i in range(4)] for k in range(9)]
    ref = [LeMoC.run(b, snapshots=False) for b in batches]
    held = []
    for k, (b, N, sed) in enumerate(LeMoC.run_many(iter(batches), depth=depth, device_pack=device_pack)):
        held.append((N, sed))                      # no copy
        j = k - depth                              # `depth` more batches have been drawn since result j
        if j >= 0:
            Nj, sedj = held[j]
            if device_pack:
                assert_spectra_close(sedj, ref[j][2][:, 0], f"held result {j} nuLnu")
                assert_spectra_close(Nj, ref[j][1][:, 0], f"held result {j} n_e")
            else:
                np.testing.assert_array_equal(sedj, ref[j][2][:, 0])
                np.testing.assert_array_equal(Nj, ref[j][1][:, 0])


@pytest.mark.parametrize("variant", ["script", "code"])
def test_device_packed_path_refuses_zero_timesteps(variant):
    """time_end <= 0 (script) / time_init >= time_end*R0/c (Code.py): the reference produces no output; the
    device-packed default path refuses the request like the host packer does, and a caller that bypasses the packers
    gets NaN rows, never stale or uninitialised memory."""
    import torch
    from lehamoc_b200 import setup_host as sh
    from lehamoc_b200.engine import Engine
    if variant == "script":
        _, P = load_golden("test1")
        P.update(grid_g_el=60., grid_nu=40.)
        with pytest.raises(ValueError):
            sh.pack_script_request([dict(P, time_end=0.5)])
        with pytest.raises(ValueError):
            sh.pack_script([dict(P, time_end=0.5)])
        b = sh.pack_script([dict(P, time_end=2.), dict(P, time_end=2., B0=2.)])
    else:
        cf, P = load_golden("code_fit")
        th = np.array(cf["thetas"])[:2]
        with pytest.raises(ValueError):
            sh.pack_code_request(th, dict(P, time_init=1e30))
        with pytest.raises(ValueError):
            sh.pack_code(th, dict(P, time_init=1e30))
        b = sh.pack_code(th, P)
    from lehamoc_b200 import _lib
    b.consts[1, _lib.C_NSTEPS] = 0.                 # bypass the packers
    eng = Engine(b.cfg, b.W)
    eng.setup(b)
    N = torch.full((b.W, 1, eng.Ng), 7.0, dtype=torch.float64, device="cuda")
    sed = torch.full((b.W, 1, eng.Nt), 7.0, dtype=torch.float64, device="cuda")
    eng.run_into(N, sed, 0)
    torch.cuda.synchronize()
    assert torch.isfinite(sed[0]).all() and torch.isfinite(N[0]).all() and not (sed[0] == 7.0).any()
    assert torch.isnan(sed[1]).all() and torch.isnan(N[1]).all()
    eng.close()


def test_code_lehamoc_matches_reference():
    """Fit_emcee/Code.py::LeHaMoC (leptonic + proton synchrotron) at the reference's hadronic start point and four
    perturbed walkers, one launch."""
    from lehamoc_b200 import Code
    z, P = load_golden("code_lh_fit")
    nu, sed = Code.LeHaMoC_batch(z["thetas"], P)
    for i in range(len(z["thetas"])):
        np.testing.assert_array_equal(nu[i], z["nu_tot"][i])
        assert_spectra_close(sed[i], z["nuLnu"][i], f"Code.LeHaMoC walker {i}")
    nu1, sed1 = Code.LeHaMoC(list(z["thetas"][2]), P)
    np.testing.assert_array_equal(sed1, sed[2])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["code", "script", "code_lh"])
def test_gpu_device_packer_matches_host_packer(variant):
    """lhm_pack_batch (grids, injection and scalar block built on the device) against setup_host._pack (NumPy, the
    reference's expressions): identical window indices, sentinels and step counts, <= 2 ulp on g_el, 1e-13 relative on
    what carries a rounded log10 in an exponent (nu_syn) or a g**-p, and the same spectra within the parity tolerance."""
    import torch
    from lehamoc_b200 import setup_host as sh
    from lehamoc_b200.engine import Engine
    if variant == "code":
        cf, P = load_golden("code_fit")
        thetas = np.array(cf["thetas"])
        thetas = np.vstack([thetas, thetas[0] + [0.3, -0.2, 0.4, -0.3, 0.2, 0.]])       # p_el == 2 row included below
        thetas[-1, 5] = 2.0
        hb, req = sh.pack_code(thetas, P), sh.pack_code_request(thetas, P)
    elif variant == "code_lh":
        cf, P = load_golden("code_lh_fit")
        thetas = np.array(cf["thetas"])
        thetas = np.vstack([thetas, thetas[0] + [0.2, -0.1, 0.3, -0.2, 0.1, 0.3, -0.5, 0.2, 0.3, 0.4]])
        thetas[-1, 6] = 0.0                                                           # g_PL_min_pr == 0 -> index 1
        hb, req = sh.pack_code_lehamoc(thetas, P), sh.pack_code_lehamoc_request(thetas, P)
    else:
        from lehamoc_b200.presets import test1_walkers
        sets = test1_walkers(12)
        sets[3] = dict(sets[3], g_PL_min=1.5, g_PL_max=5.0, p_el=2.0)
        hb, req = sh.pack_script(sets), sh.pack_script_request(sets)
    assert {k: v for k, v in req.cfg.items()} == {k: hb.cfg[k] for k in req.cfg}
    eng = Engine(req.cfg, req.W)
    eng.setup_packed(req)
    torch.cuda.synchronize()
    dev = [t.cpu().numpy() for t in eng._keep]
    names = ("consts", "g_el", "nu_syn", "nu_ic", "nu_tot", "el_inj", "ext")
    for name, d, h in zip(names, dev, hb.arrays()):
        assert d.shape == h.shape, name
        if name in ("nu_ic", "nu_tot", "ext"):
            np.testing.assert_array_equal(d, h, err_msg=name)          # host rows, replicated
            continue
        sent = h == sh.SENTINEL
        assert np.array_equal(d == sh.SENTINEL, sent), name + ": injection window differs"
        nz = (h != 0) & ~sent
        assert np.array_equal(d[~nz & ~sent], h[~nz & ~sent]), name
        if name == "g_el":                                           # same exponents, 10**y from two math libraries
            ulp = np.abs(d[nz] - h[nz]) / np.spacing(np.abs(h[nz]))
            assert ulp.max() <= 2, (name, ulp.max())
        else:       # nu_syn: its upper exponent carries one ulp of a log10 -> ln(10)*|y|*eps ~ 1e-14 relative
            rel = np.abs(d[nz] / h[nz] - 1.0)
            assert rel.max() <= 1e-13, (name, rel.max())
    from lehamoc_b200 import _lib
    np.testing.assert_array_equal(dev[0][:, _lib.C_NSTEPS], hb.consts[:, _lib.C_NSTEPS])
    eng2 = Engine(hb.cfg, hb.W)
    eng2.setup(hb)
    if variant == "code_lh":
        for name, d, h in zip(("g_pr", "pr_inj"), [t.cpu().numpy() for t in eng._keep_pr], (hb.g_pr, hb.pr_inj)):
            sent = h == sh.SENTINEL
            assert np.array_equal(d == sh.SENTINEL, sent), name + ": proton injection window differs"
            rel = np.abs(d[~sent] / h[~sent] - 1.0)
            assert rel.max() <= 1e-13, (name, rel.max())
        r1, r2 = eng.run_lh(), eng2.run_lh()
        assert_spectra_close(r1[1][:, 0].cpu().numpy(), r2[1][:, 0].cpu().numpy(), "device-packed vs host-packed nuLnu")
        assert_spectra_close(r1[2][:, 0].cpu().numpy(), r2[2][:, 0].cpu().numpy(), "device-packed vs host-packed n_p")
        return
    N1, sed1 = eng.run()
    N2, sed2 = eng2.run()
    assert_spectra_close(sed1.cpu().numpy(), sed2.cpu().numpy(), "device-packed vs host-packed nuLnu")
    assert_spectra_close(N1.cpu().numpy(), N2.cpu().numpy(), "device-packed vs host-packed n_e")


@pytest.mark.gpu
def test_engines_of_different_size_coexist():
    """The dynamic shared-memory opt-in is a property of the kernel, not of a handle: an engine for small grids created
    after one for large grids must not make the large one's launches fail."""
    from lehamoc_b200 import LeMoC
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    _, P = load_golden("test1")
    big = pack_script(dict(P, grid_g_el=400., grid_nu=200., time_end=2.))
    small = pack_script(dict(P, grid_g_el=60., grid_nu=40., time_end=2.))
    e_big = Engine(big.cfg, 1)
    e_small = Engine(small.cfg, 1)
    assert e_big.smem_bytes() > e_small.smem_bytes()
    e_small.setup(small)
    Ns, _ = e_small.run()
    e_big.setup(big)
    Nb, sedb = e_big.run()
    _, Nb2, sedb2 = LeMoC.run(dict(P, grid_g_el=400., grid_nu=200., time_end=2.), snapshots=False)
    np.testing.assert_array_equal(sedb.cpu().numpy(), sedb2[:, 0])
    assert np.isfinite(Ns.cpu().numpy()[:, 1:-1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_random_flag_combinations_vs_oracle(orc, seed):
    """Randomly drawn switches, grids, step sizes and physical parameters on small grids: every step of the CUDA path
    against the oracle (itself pinned on the live reference for each of these branches)."""
    from lehamoc_b200 import LeMoC
    _, P = load_golden("test1")
    rng = np.random.default_rng(1000 + seed)
    flag = lambda p=0.5: float(rng.random() < p)
    gmax = float(rng.uniform(4.5, 6.5))
    Q = dict(P, grid_g_el=float(rng.integers(24, 45) * 2), grid_nu=float(rng.integers(30, 70)),
             time_end=float(rng.integers(2, 5)), step_alg=float(rng.choice([0.5, 1., 2., 3.])),
             g_min_el=0., g_max_el=gmax, g_PL_min=float(rng.choice([0., rng.uniform(0.5, 2.5)])),
             g_PL_max=float(rng.choice([gmax, rng.uniform(3.5, gmax - 0.2)])),
             p_el=float(rng.choice([2., rng.uniform(1.5, 3.)])), L_el=float(rng.uniform(39., 44.)),
             R0=float(rng.uniform(14., 16.5)), B0=float(10 ** rng.uniform(-2, 1.5)),
             Vexp=float(rng.choice([1e-13, rng.uniform(0.01, 0.3)])), m=float(rng.choice([0., 1., 2.])),
             inj_flag=flag(0.7), Ad_l_flag=flag(), Syn_l_flag=flag(0.8), Syn_emis_flag=flag(0.8), IC_l_flag=flag(0.7),
             IC_emis_flag=flag(0.7), SSA_l_flag=flag(), gg_flag=flag(), esc_flag=flag(0.7), BB_flag=flag(0.3),
             BB_temperature=float(rng.uniform(3., 6.)), GB_ext=float(10 ** rng.uniform(-4, 0)))
    ge, ne_ref, sed_ref = orc.run_script(Q)
    for path in ("R", "S"):
        batch, N, sed = LeMoC.run(Q, ic_path=path)
        assert N.shape[1] == len(ne_ref), (seed, path)
        # path S: the suffix sums cancel in bins that only hold the 1e-260 start-up photons, so bins more than 200 decades
        # below the peak agree to ~1e-4 there instead of 1e-6 (every other bin to the same 1e-8 as path R)
        tf = 1e-6 if path == "R" else 1e-3
        for s in range(len(ne_ref)):
            assert_spectra_close(N[0, s], ne_ref[s], f"seed {seed} path {path} n_e step {s + 1}: {Q}", tol_floor=tf)
            assert_spectra_close(sed[0, s], sed_ref[s], f"seed {seed} path {path} nuLnu step {s + 1}: {Q}", tol_floor=tf)


@pytest.mark.gpu
def test_full_size_linearity_without_feedback():
    """Size-independent property at the Test-1 grid size: with inverse-Compton, gamma-gamma and self-absorption off
    nothing feeds back on the electrons, so doubling the injected luminosity doubles every bin of both spectra."""
    from lehamoc_b200 import LeMoC
    _, P = load_golden("test1")
    base = dict(P, IC_l_flag=0., IC_emis_flag=0., gg_flag=0., SSA_l_flag=0.)
    two = dict(base, L_el=P["L_el"] + np.log10(2.))
    _, N, sed = LeMoC.run([base, two], snapshots=False)
    inner = slice(1, -1)                                  # the boundary bins hold the 1e-260 sentinels
    np.testing.assert_allclose(N[1, 0][inner] / N[0, 0][inner], 2.0, rtol=1e-12)
    big = sed[0, 0] > sed[0, 0].max() * 1e-60            # below: bins that still hold the 1e-260 initial photons
    np.testing.assert_allclose(sed[1, 0][big] / sed[0, 0][big], 2.0, rtol=1e-12)


@pytest.mark.gpu
def test_random_fit_walkers_vs_oracle(orc):
    """Walkers drawn over the whole prior box of the fit (fit_emcee.py:173-188) in one launch, both numerical models
    (Code.LeMoC: per-walker gamma grids; Code.LeHaMoC: + protons), each against its own oracle run."""
    from lehamoc_b200 import Code
    _, P = load_golden("code_fit")
    rng = np.random.default_rng(4242)
    lo = np.array([14.2, -1.8, 0.2, 5.1, -5.8, 1.6])
    hi = np.array([16.8, 1.8, 4.3, 6.9, -2.2, 2.9])
    th = lo + (hi - lo) * rng.random((12, 6))
    th[3, 5] = 2.0
    nu, sed = Code.LeMoC_batch(th, P)
    nuS, sedS = Code.LeMoC_batch(th, P, ic_path="S")
    for i in range(len(th)):
        nu_ref, sed_ref = orc.code_LeMoC(th[i], P)
        np.testing.assert_array_equal(nu[i], nu_ref)
        assert_spectra_close(sed[i], sed_ref, f"Code.LeMoC random walker {i} {th[i]}")
        assert_spectra_close(sedS[i], sed_ref, f"Code.LeMoC random walker {i} path S {th[i]}", tol_floor=1e-3)
    _, P2 = load_golden("code_lh_fit")
    lo2 = np.array([14.2, -0.8, 0.2, 4.1, -4.8, 1.2, 0.2, 6.2, -4.8, 1.2])
    hi2 = np.array([16.8, 2.8, 2.8, 5.9, -0.5, 2.9, 4.5, 9.5, -0.5, 2.9])
    th2 = lo2 + (hi2 - lo2) * rng.random((8, 10))
    th2[2, 9] = 2.0
    nu2, sed2 = Code.LeHaMoC_batch(th2, P2)
    nu2d, sed2d = Code.LeHaMoC_batch(th2, P2, device_pack=True)
    nud, sedd = Code.LeMoC_batch(th, P, device_pack=True)
    sedd = sedd.copy()
    for i in range(len(th2)):
        nu_ref, sed_ref = orc.code_LeHaMoC(th2[i], P2)
        np.testing.assert_array_equal(nu2[i], nu_ref)
        np.testing.assert_array_equal(nu2d[i], nu_ref)
        assert_spectra_close(sed2[i], sed_ref, f"Code.LeHaMoC random walker {i} {th2[i]}")
        assert_spectra_close(sed2d[i], sed_ref, f"Code.LeHaMoC random walker {i}, device-side set-up")
    for i in range(len(th)):
        np.testing.assert_array_equal(nud[i], nu[i])
        assert_spectra_close(sedd[i], sed[i], f"Code.LeMoC random walker {i}, device-side set-up")


@pytest.mark.gpu
def test_ragged_step_counts_in_one_batch():
    """Walkers of one launch that stop after different numbers of timesteps (the Code.py drivers count steps per walker,
    Code.py:183): every walker's final state is bit-identical to the same walker in a batch where all share its count.
    A walker with zero steps produces no output, like the reference's loop (LeMoC.py:196, range(0) writes no row): its
    rows are unspecified, and its presence must not disturb the other walkers of the launch."""
    from lehamoc_b200._lib import C_NSTEPS
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    _, P = load_golden("test1")
    P.update(grid_g_el=60., grid_nu=40., time_end=7.)
    sets = [dict(P, B0=0.3 + 0.2 * i, p_el=1.8 + 0.1 * i) for i in range(5)]
    counts = [7, 1, 4, 0, 2]
    b = pack_script(sets)
    b.nsteps[:] = counts
    b.consts[:, C_NSTEPS] = counts
    eng = Engine(b.cfg, b.W)
    N = np.full((b.W, 1, b.cfg["Ng"]), -1.0)
    sed = np.full((b.W, 1, b.cfg["Nt"]), -1.0)
    eng.run_host(b, N_out=N, sed_out=sed)
    for i, k in enumerate(counts):
        if k == 0:
            continue
        u = pack_script(sets)
        u.nsteps[:] = k
        u.consts[:, C_NSTEPS] = k
        Nu, su = eng.run_host(u)
        np.testing.assert_array_equal(N[i, 0], Nu[i])
        np.testing.assert_array_equal(sed[i, 0], su[i])
        assert np.all(np.isfinite(N[i])) and np.all(sed[i] >= 0)
