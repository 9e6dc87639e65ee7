"""Generate tests/golden/pp_functions.npz: outputs of the UNMODIFIED pp-channel functions of
`LeHaMoC Code and Tests/LeHaMoC_f.py` (cs_pp_inel, Q_e_pp, Q_g_pp, Q_nu_mu_pp, Q_nu_e_pp; :687-1014) for the two seeded
proton distributions of make_golden_had.py.   TEST INFRASTRUCTURE ONLY; build container only."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_had as mg  # noqa: E402


def main():
    out = {}
    with mg.reference_had_functions() as f:
        out["cs_E"] = np.logspace(-4., 4., 33)
        out["cs_pp_inel"] = np.array(f.cs_pp_inel(out["cs_E"]))
        out["g_el"] = np.logspace(0., 11., 60)
        out["nu_ic"] = np.logspace(15., 32., 30)
        out["nu_nu"] = np.logspace(10., 22., 50) * f.eV / f.h
        for case in range(2):
            nu, photons, g_pr, N_pr, g_el, nu_ic = mg.fields(100 + case)
            NT = N_pr / (f.m_pr * f.c ** 2. * 0.624151)
            out[f"c{case}_g_pr"], out[f"c{case}_N_pr_TeV"] = g_pr, NT
            for p_pr in (2.01, 2.5):
                tag = f"c{case}_p{int(p_pr * 100)}_"
                with np.errstate(all="ignore"):
                    out[tag + "Q_e_pp"] = np.array(f.Q_e_pp(out["g_el"], g_pr, NT, p_pr, 1.5), dtype=np.float64)
                    out[tag + "Q_g_pp"] = np.array(f.Q_g_pp(out["nu_ic"], g_pr, NT, p_pr, 1.5), dtype=np.float64)
                    out[tag + "Q_nu_mu_pp"] = np.array(f.Q_nu_mu_pp(out["nu_nu"], g_pr, NT, p_pr, 1.5), dtype=np.float64)
                    out[tag + "Q_nu_e_pp"] = np.array(f.Q_nu_e_pp(out["nu_nu"], g_pr, NT, p_pr, 1.5), dtype=np.float64)
    path = os.path.join(mg.GOLD, "pp_functions.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
