"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE ONLY.  Usage (build container, /root/reference present):

    python oracle/make_golden.py            # writes tests/golden/*.npz (a few minutes)

Every fixture stores the parameter set it was generated from, so the tests rebuild the inputs
themselves; the reference is only needed here.  The fixtures are the pin for both the NumPy oracle
(oracle/lehamoc_oracle.py) and, on the GPU box, the CUDA path.
"""
from __future__ import annotations

import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as rr  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

TEST1 = dict(time_init=0., time_end=10., step_alg=1., g_min_el=0., g_max_el=6.1, g_PL_min=0.,
             g_PL_max=6.1, grid_g_el=260., grid_nu=100., p_el=1.9, L_el=40.48572142648158,
             Vexp=0.0000000000001, R0=15., B0=1., m=0., delta=1., inj_flag=1., Ad_l_flag=0.,
             Syn_l_flag=1., Syn_emis_flag=1., IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1., gg_flag=1.,
             esc_flag=1., BB_flag=0., BB_temperature=6., GB_ext=1., PL_flag=0., dE_dV_ph=0.,
             nu_min_ph=0., nu_max_ph=0., s_ph=0., User_ph=0.)


def variant(**kw):
    p = dict(TEST1)
    p.update(kw)
    return p


# small derived cases that exercise the branches the three shipped files do not reach together
VARIANTS = {
    # grey-body external field (LeMoC.py:143-148) on a coarse grid
    "bb": variant(grid_g_el=100., grid_nu=60., BB_flag=1., BB_temperature=4., GB_ext=1e-3, time_end=5.),
    # adiabatic expansion + decaying field B ~ R^-1 with every radiative process on
    "expand": variant(grid_g_el=120., grid_nu=80., Ad_l_flag=1., Vexp=0.05, m=1., time_end=6., step_alg=2.,
                      g_PL_min=1., g_PL_max=5., g_max_el=6.),
    # no pair production, no escape of electrons, no injection after t=0, p == 2 normalisation branch
    "nogg": variant(grid_g_el=90., grid_nu=50., gg_flag=0., esc_flag=0., inj_flag=0., p_el=2., time_end=4.,
                    g_PL_min=2., g_PL_max=4.5, g_max_el=5.),
    # synchrotron only (IC/SSA/gg off)
    "synonly": variant(grid_g_el=80., grid_nu=40., IC_l_flag=0., IC_emis_flag=0., SSA_l_flag=0., gg_flag=0.,
                       time_end=3., B0=10., R0=14.5),
}

FIT_FROZEN = dict(time_init=0., time_end=10., step_alg=3., grid_g_el=160., grid_g_pr=160., grid_nu=100.,
                  Vexp=0.0000000000001, m=0., inj_flag=1., Ad_l_flag=0., Syn_l_flag=1., Syn_emis_flag=1.,
                  IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1., gg_flag=1., esc_flag=1., BB_flag=0.,
                  BB_temperature=6., GB_ext=1., PL_flag=0., dE_dV_ph=0., nu_min_ph=0., nu_max_ph=0., s_ph=0.,
                  User_ph=0.)
FIT_START = [15.25808, 0.038818, 3., 6.081445, -4., 2.]        # fit_emcee.py:208 (without log delta)


def pack_params(p: dict):
    keys = sorted(p)
    return {"param_keys": np.array(keys), "param_vals": np.array([p[k] for k in keys], dtype=np.float64)}


def save(name, cap, params, extra=None):
    out = dict(cap)
    out.update(pack_params(params))
    if extra:
        out.update(extra)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def thin(cap, every):
    """Keep every `every`-th per-step row (and the last) of the step_*/out_* arrays."""
    n = int(cap["nsteps"])
    keep = sorted(set(list(range(every - 1, n, every)) + [n - 1]))
    out = {}
    for k, v in cap.items():
        if (k.startswith("step_") or k.startswith("out_")) and np.ndim(v) == 2 and v.shape[0] == n:
            out[k] = v[keep]
        else:
            out[k] = v
    out["kept_steps"] = np.array(keep, dtype=np.int64)      # 0-based step numbers of the kept rows
    return out


USER_LOGX = np.linspace(11., 16., 26)
USER_LOGY = -3. - 0.4 * (USER_LOGX - 13.2) ** 2                 # a log-parabola photon field, cm^-3 Hz^-1
USER_TXT = "".join(f"{x!r},{y!r}\n" for x, y in zip(USER_LOGX.tolist(), USER_LOGY.tolist()))


def make_user():
    """var_user / lh_user: User_ph = 1 with a Photons_spec_user.txt in the working directory (LeMoC.py:160-169,
    LeHaMoC.py:201-210), unmodified scripts."""
    warnings.filterwarnings("ignore")
    tmp = tempfile.mkdtemp(prefix="lhm_gold_")
    p = variant(grid_g_el=100., grid_nu=60., User_ph=1., time_end=4.)
    pf = os.path.join(tmp, "Parameters_user.txt")
    rr.write_params(pf, p)
    cap, wall, _ = rr.run_lemoc_script(pf, extra_files={"Photons_spec_user.txt": USER_TXT})
    print(f"variant user: {wall:.1f} s")
    save("var_user", cap, p, {"wall_s": np.float64(wall), "user_logx": USER_LOGX, "user_logy": USER_LOGY})
    import run_reference_lh as rl
    import make_golden_lh as mgl
    P = mgl.variant(grid_g_el=80., grid_nu=60., time_end=4., User_ph=1.)
    r = rl.run_lehamoc_script(P, extra_files={"Photons_spec_user.txt": USER_TXT})
    keys = sorted(P)
    out = {"param_keys": np.array(keys), "param_vals": np.array([P[k] for k in keys], dtype=np.float64),
           "user_logx": USER_LOGX, "user_logy": USER_LOGY}
    out.update({k: v for k, v in r.items() if k != "wall_s"})
    out["wall_s"] = np.float64(r["wall_s"])
    np.savez_compressed(os.path.join(GOLD, "lh_user.npz"), **out)
    print(f"lh_user: {r['wall_s']:.1f} s")


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="lhm_gold_")

    # 1. Test 1 exactly as shipped (BASELINE config 1), every step, all intermediates
    cap, wall, (ptxt, ftxt) = rr.run_lemoc_script("Parameters_Test1.txt")
    print(f"Test1: {wall:.1f} s (cor_factor memoised)")
    n_rows_g, n_rows_t = len(cap["g_el"]), len(cap["nu_tot"])
    save("test1", cap, rr.read_params(os.path.join(rr.REF_ROOT, "LeMoC", "Parameters_Test1.txt")),
         {"wall_s": np.float64(wall),
          # first snapshot of both output files verbatim: pins the text format (LeMoC.py:290-296)
          "particles_txt_step1": np.array("\n".join(ptxt.split("\n")[:n_rows_g]) + "\n"),
          "photons_txt_step1": np.array("\n".join(ftxt.split("\n")[:n_rows_t]) + "\n")})

    # 2. LeMoC/Parameters.txt: the file that produced the shipped *_Test3_LeMoC.txt (syn+SSA+adiabatic)
    cap, wall, _ = rr.run_lemoc_script("Parameters.txt")
    print(f"Parameters.txt: {wall:.1f} s")
    save("test3_synadi", thin(cap, 10), rr.read_params(os.path.join(rr.REF_ROOT, "LeMoC", "Parameters.txt")),
         {"wall_s": np.float64(wall)})

    # 3. Parameters_Test3.txt as shipped: 600 steps, adiabatic losses only
    cap, wall, _ = rr.run_lemoc_script("Parameters_Test3.txt")
    print(f"Parameters_Test3.txt: {wall:.1f} s")
    save("test3_adiabatic", thin(cap, 100),
         rr.read_params(os.path.join(rr.REF_ROOT, "LeMoC", "Parameters_Test3.txt")), {"wall_s": np.float64(wall)})

    # 4. derived parameter files
    for name, p in VARIANTS.items():
        pf = os.path.join(tmp, f"Parameters_{name}.txt")
        rr.write_params(pf, p)
        cap, wall, _ = rr.run_lemoc_script(pf)
        print(f"variant {name}: {wall:.1f} s")
        save("var_" + name, cap, p, {"wall_s": np.float64(wall)})

    # 5. Fit_emcee/Code.py::LeMoC at the fit start point and at perturbed walkers (BASELINE config 4)
    rng = np.random.default_rng(20231)
    lo = np.array([14.5, -1.0, 1.0, 5.2, -5.0, 1.6])
    hi = np.array([16.5, 1.0, 3.5, 6.5, -3.0, 2.8])
    thetas = [np.array(FIT_START)] + [lo + (hi - lo) * rng.random(6) for _ in range(7)]
    thetas[1][5] = 2.0                                          # p == 2 branch of Q_el_Lum
    nus, seds, walls = [], [], []
    for th in thetas:
        nu, sed, wall = rr.run_code_lemoc(th, "Parameters.txt")
        nus.append(nu), seds.append(sed), walls.append(wall)
        print(f"Code.LeMoC {np.round(th, 3)}: {wall:.1f} s")
    out = {"thetas": np.array(thetas), "nu_tot": np.array(nus), "nuLnu": np.array(seds), "wall_s": np.array(walls)}
    out.update(pack_params(FIT_FROZEN))
    path = os.path.join(GOLD, "code_fit.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)

    # 6. function-level vectors: the reference library called directly on the Test-1 state of step 5
    t1 = np.load(os.path.join(GOLD, "test1.npz"))
    with rr.reference_functions("LeMoC") as f:
        g = t1["g_el"]
        fl = {"cor_factor_test1": f.cor_factor_syn_el(g, 1e15, 1e4, 1.9, f.Lum_e_inj(1e-3, 1e15)),
              "cor_factor_fitgrid": f.cor_factor_syn_el(np.logspace(2., 7.081445, 160), 10 ** 15.25808, 1e4, 2.0,
                                                        f.Lum_e_inj(-4., 10 ** 15.25808)),
              "thomas_in_b": np.linspace(1.5, 2.5, 17), "thomas_in_c": -np.linspace(0.1, 0.9, 17),
              "thomas_in_a": np.linspace(0.05, 0.3, 17), "thomas_in_d": np.cos(np.arange(17.)) + 2.}
        fl["thomas_out_tridiag"] = f.thomas(fl["thomas_in_a"], fl["thomas_in_b"], fl["thomas_in_c"], fl["thomas_in_d"])
        fl["thomas_out_bidiag"] = f.thomas(np.zeros(17), fl["thomas_in_b"], fl["thomas_in_c"], fl["thomas_in_d"])
    np.savez_compressed(os.path.join(GOLD, "functions.npz"), **fl)
    print("wrote functions.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "user":
        make_user()
    else:
        main()
