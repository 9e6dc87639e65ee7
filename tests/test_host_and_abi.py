"""CPU: host-side logic (parameter files, batched grid construction) and the C-ABI library surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import lehamoc_oracle as orc
from lehamoc_b200 import _lib, setup_host
from util import ROOT, load_golden, quiet

quiet()


def test_parameter_file_roundtrip(tmp_path):
    _, P = load_golden("test1")
    pf = tmp_path / "Parameters.txt"
    pf.write_text("\n".join(f"{k} = {v!r}" for k, v in P.items()))      # no trailing newline, like the shipped files
    assert setup_host.read_parameter_file(pf) == P == orc.read_parameter_file(pf)
    with pytest.raises(KeyError):
        bad = dict(P); bad.pop("gg_flag")
        setup_host.pack_script(bad)
    with pytest.raises(ValueError):
        setup_host.pack_script(dict(P, SSA_l_flag=0.5))


@pytest.mark.parametrize("name", ["test1", "var_bb", "var_expand", "var_nogg", "test3_synadi"])
def test_pack_script_is_bit_identical_to_reference_setup(name):
    z, P = load_golden(name)
    b = setup_host.pack_script(P)
    for k, a in (("g_el", b.g_el), ("nu_syn", b.nu_syn), ("nu_ic", b.nu_ic), ("nu_tot", b.nu_tot), ("el_inj", b.el_inj)):
        np.testing.assert_array_equal(a[0], z[k], err_msg=k)          # vectors captured from the live reference
    assert int(b.index_PL_min[0]) == int(z["index_PL_min"]) and int(b.index_PL_max[0]) == int(z["index_PL_max"])
    assert int(b.nsteps[0]) == int(z["nsteps"]) and b.consts[0, _lib.C_DT] == float(z["dt"])
    # external fields as they come out of photons_tot
    r = orc.LeptonicRun(P, "script")
    V = orc.Volume(r.R0)
    with np.errstate(all="ignore"):
        ext = 10 ** np.interp(np.log10(r.nu_tot), np.log10(r.nu_bb), np.log10(r.bb * V)) / V + r.pl
    np.testing.assert_allclose(b.ext[0], ext, rtol=1e-12)


def test_pack_code_matches_oracle_and_batches():
    z, P = load_golden("code_fit")
    b = setup_host.pack_code(z["thetas"], P)
    assert b.cfg["loss_minus_one"] == 0 and b.cfg["qee_from_ic"] == 1
    for i, th in enumerate(z["thetas"]):
        r = orc.LeptonicRun(P, "code", theta=list(th))
        np.testing.assert_array_equal(b.g_el[i], r.g_el)
        np.testing.assert_array_equal(b.nu_syn[i], r.nu_syn)
        np.testing.assert_array_equal(b.nu_tot[i], z["nu_tot"][i])
        np.testing.assert_array_equal(b.el_inj[i], r.el_inj)
        assert int(b.nsteps[i]) == r.n_steps() == 4
    one = setup_host.pack_code(z["thetas"][3], P)
    np.testing.assert_array_equal(one.g_el[0], b.g_el[3])


def test_heterogeneous_script_batch_rules():
    _, P = load_golden("test1")
    b = setup_host.pack_script([P, dict(P, B0=3.0, p_el=2.0, g_PL_min=2.0, g_PL_max=5.0)])
    assert b.W == 2 and b.index_PL_max[0] == -1 and b.index_PL_max[1] > 0
    assert b.el_inj[1, 0] == 10 ** (-260.) and b.el_inj[1, b.index_PL_min[1]] > 1.
    with pytest.raises(ValueError):
        setup_host.pack_script([P, dict(P, grid_nu=50.)])            # grid sizes are shared by a batch
    with pytest.raises(ValueError):
        setup_host.pack_script(dict(P, g_PL_max=7.5))                # window outside the grid


def test_library_exports_every_declared_symbol():
    """liblhm.so loads on a GPU-less host and exports exactly what include/lhm.h declares."""
    header = open(os.path.join(ROOT, "include", "lhm.h")).read()
    declared = set(re.findall(r"\b(lhm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lhm_version() == 102
    assert b"no CPU fallback" in lib.lhm_strerror(-3)
    # struct layout agreed between header and binding
    body = re.sub(r"/\*.*?\*/", " ", re.search(r"typedef struct \{(.*?)\} lhm_config;", header, re.S).group(1), flags=re.S)
    header_fields = [n.strip() for decl in body.split(";") if decl.strip()
                     for n in re.sub(r"^\s*int\b", "", decl.strip()).split(",")]
    assert ctypes.sizeof(_lib.lhm_config) == 4 * len(_lib.lhm_config._fields_)
    assert [n for n, _ in _lib.lhm_config._fields_] == header_fields          # same names, same order


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lehamoc_b200 import Code
    z, P = load_golden("code_fit")
    with pytest.raises(_lib.LhmError):
        Code.LeMoC(list(z["thetas"][0]), P)
    cfg = _lib.lhm_config(Ng=16, Ns=9, Ni=8, Nt=10, gg_npts=50, max_walkers=1)
    h = ctypes.c_void_p()
    rc = _lib.load().lhm_create(ctypes.byref(cfg), 0, None, ctypes.byref(h))
    assert rc == -3                                                   # LHM_ERR_NO_DEVICE, never a silent CPU path


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lehamoc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "lehamoc_oracle" not in src and "run_reference" not in src, fn


def test_integration_md_stub_matches_binding():
    """The ctypes stub printed in INTEGRATION.md declares the same lhm_config as the binding and the header."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"class lhm_config\(C.Structure\):.*?\)\]", text, re.S).group(0)
    names = re.findall(r'"(\w+)"', block)
    assert names == [n for n, _ in _lib.lhm_config._fields_]
    header = open(os.path.join(ROOT, "include", "lhm.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} lhm_config;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    hdr_names = [n.strip() for decl in body.split(";") for n in decl.replace("int", "", 1).split(",") if n.strip()]
    assert hdr_names == names


def test_vectorised_proton_packing_is_the_loop():
    """pack_code_lehamoc builds the proton grid / injection of all walkers at once; the per-walker loop with the
    reference's own expressions (Code.py:299-300,325-347,419-422) gives the same doubles."""
    from lehamoc_b200 import setup_host as sh
    rng = np.random.default_rng(3)
    start = np.array([15.5, 1., 0.5, 5., -3., 1.2, 0.5, 7.3, -4.5, 2.])
    th = start + 0.3 * rng.standard_normal((64, 10))
    th[5, 9] = 2.0                      # p_pr == 2 branch of Q_pr_Lum
    th[7, 6] = 0.0                      # g_PL_min_pr == 0 -> index 1
    R0 = np.array([10 ** x for x in th[:, 0].tolist()])
    a = sh._protons_code_lh_loop(th, R0, 160)
    b = sh._protons_code_lh(th, R0, 160)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_zero_timesteps_are_refused():
    """time_end = 0: the reference's loop body never runs (LeMoC.py:196 writes no row, Code.py:277 raises on the unbound
    spectrum), so there is nothing to reproduce and the packers say so."""
    _, P = load_golden("test1")
    with pytest.raises(ValueError):
        setup_host.pack_script(dict(P, time_end=0.))


def test_window_indices_are_the_references():
    """The device-packed path gets index_PL_min / index_PL_max from the host (setup_host._window_indices evaluates the
    reference's comparison on a few NumPy grid points around the threshold): identical to LeMoC.py:120-128 on the full
    np.logspace row, also when a threshold coincides with a grid point to the last bit."""
    from lehamoc_b200 import setup_host as sh
    rng = np.random.default_rng(1)
    W = 4000
    for n in (160, 260, 37):
        lo = rng.uniform(-1, 2, W)
        hi = lo + rng.uniform(3, 7, W)
        pl_min = np.where(rng.random(W) < 0.2, 0., lo + rng.uniform(0.01, 1.5, W))
        pl_max = np.where(rng.random(W) < 0.2, hi, hi - rng.uniform(0.01, 1.5, W))
        for w in range(W // 2):                     # adversarial half: thresholds sitting on a grid point
            g = np.logspace(lo[w], hi[w], n)
            j = rng.integers(2, n - 2)
            if w % 2:
                pl_max[w] = np.log10(g[j])
            else:
                pl_min[w] = np.log10(g[j])
        imin, imax = sh._window_indices(lo, hi, n, pl_min, pl_max)
        # liblhm's host routine (the same candidates with libm's pow; near-ties handed back to the NumPy evaluation)
        lmin, lmax = sh._window_indices_lib(lo, hi, n, pl_min, pl_max)
        assert np.array_equal(lmin, imin) and np.array_equal(lmax, imax), n
        for w in range(W):
            g = np.logspace(lo[w], hi[w], n)
            rmax = -1 if pl_max[w] == hi[w] else int(np.min(np.where(g > 10 ** pl_max[w])[0]))
            rmin = 1 if pl_min[w] == 0. else int(np.max(np.where(g < 10 ** pl_min[w])[0]))
            assert (rmin, rmax) == (imin[w], imax[w]), (n, w)
    with pytest.raises(ValueError):                 # window above the grid: the reference raises (min of empty)
        sh._window_indices(np.array([0.]), np.array([4.]), 50, np.array([1.]), np.array([5.]))
    with pytest.raises(ValueError):
        sh._window_indices_lib(np.array([0.]), np.array([4.]), 50, np.array([1.]), np.array([5.]))
