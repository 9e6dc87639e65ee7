"""Shared helpers of the parity tests."""
import os
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: <= 1e-8 relative per grid bin  <=>  |delta log10| <= 4.3e-9
TOL_LOG10 = 4.3e-9
# bins more than 200 decades below the spectrum's peak carry only the 1e-260 sentinels:
# they must be finite and agree to 1e-6 in log10
FLOOR_DECADES = 200.0
TOL_LOG10_FLOOR = 1e-6


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    P = dict(zip([str(k) for k in z["param_keys"]], [float(v) for v in z["param_vals"]]))
    if "user_logx" in z.files:          # User_ph = 1 fixtures: the table the reference read from Photons_spec_user.txt
        P["_user_spectrum"] = (np.array(z["user_logx"]), np.array(z["user_logy"]))
    return z, P


def assert_spectra_close(got, ref, what="", tol=TOL_LOG10, tol_floor=TOL_LOG10_FLOOR):
    """Compare two positive spectra bin by bin in log10 with the north_star tolerance."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    with np.errstate(divide="ignore", invalid="ignore"):
        lg, lr = np.log10(got), np.log10(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{what}: NaN pattern differs from the reference's"
    zero = ref == 0
    assert np.array_equal(got[zero], ref[zero]), f"{what}: exact zeros of the reference not reproduced"
    pos = ref > 0
    assert np.all(np.isfinite(lg[pos])), f"{what}: non-finite / non-positive bins where the reference is positive"
    peak = np.max(lr[pos], axis=-1, keepdims=True) if ref.ndim > 1 else np.max(lr[pos])
    d = np.abs(lg - lr)
    main = pos & (lr > (peak - FLOOR_DECADES))
    floor = pos & ~main
    worst = float(d[main].max()) if main.any() else 0.0
    assert worst <= tol, f"{what}: max |dlog10| = {worst:.3e} > {tol:.1e} on {int(main.sum())} bins"
    if floor.any():
        wf = float(d[floor].max())
        assert wf <= tol_floor, f"{what}: sentinel-level bins differ by {wf:.3e} in log10"
    return worst


def assert_signed_close(got, ref, what="", rtol=1e-8):
    """Relative comparison for signed quantities (aSSA); exact zeros must match."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    nz = ref != 0
    assert np.array_equal(got[~nz], ref[~nz]), f"{what}: zeros differ"
    scale = np.abs(ref[nz])
    err = np.abs(got[nz] - ref[nz]) / scale
    big = scale > np.max(scale) * 1e-200
    worst = float(err[big].max()) if big.any() else 0.0
    assert worst <= rtol, f"{what}: max rel err {worst:.3e} > {rtol:.1e}"
    return worst


def quiet():
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
