"""Drop-in for `python LeHaMoC.py <Parameters.txt> <_suffix>` (`LeHaMoC Code and Tests/LeHaMoC.py`:56-71,436-453):

    python -m lehamoc_b200.LeHaMoC Parameters.txt _out

writes Pairs_Distribution<suffix>.txt, Photons_Distribution<suffix>.txt, Protons_Distribution<suffix>.txt and
Neutrinos_Distribution<suffix>.txt in the reference's text format (one block per timestep).  `run()` is the importable
form and takes several parameter sets at once.  The photopion, Bethe-Heitler and pp proton loss terms, the photopion
pairs / photons / neutrinos, the Bethe-Heitler pair injection and the pp pairs / pi0 photons / neutrinos are all part
of the fused timestep.
"""
from __future__ import annotations

import sys
import time

import numpy as np

from .constants import eV, h
from .engine import Engine
from .setup_host import pack_lehamoc, read_parameter_file


def run(param_sets, ic_path="R", snapshots=True, engine=None):
    """Returns (batch, n_e [W,S,Ng], nuLnu [W,S,Nt], n_p [W,S,Np], N_nu [W,S,50]) as NumPy arrays; S = number of
    timesteps when `snapshots` (the reference writes every step), else 1 (final state).  With a photo-hadronic flag
    set the six tables are read from the working directory like the reference does (or set them beforehand with
    lehamoc_b200.hadronic.load_tables / set_tables)."""
    batch = pack_lehamoc(param_sets)
    eng = engine or Engine(batch.cfg, max(batch.W, 1), ic_path=ic_path)
    if any(batch.cfg[k] for k in ("pg_pi_l_flag", "pg_BH_l_flag", "pg_pi_emis_flag")):
        from . import hadronic
        eng.set_had_tables(hadronic._need_tables())
    if batch.cfg["pg_BH_emis_flag"]:
        from . import hadronic
        if not (np.all(batch.nu_tot[:, 0] == batch.nu_tot[0, 0]) and np.all(batch.nu_tot[:, -1] == batch.nu_tot[0, -1])):
            raise ValueError("pg_BH_emis_flag: the walkers of one batch must share the nu_tot range (one pair of "
                             "Bethe-Heitler surfaces per batch, LeHaMoC.py:263)")
        tm, tp, mm, mp_ = bh_surfaces(batch.nu_tot[0, 0], batch.nu_tot[0, -1])
        sm, km = hadronic.bh_surface(tm, mm)
        sp, kp = hadronic.bh_surface(tp, mp_)
        eng.set_bh_surfaces(sm, sp, (km, kp))
    eng.setup(batch)
    S = int(batch.nsteps.max()) if snapshots else 0
    N, sed, Npr, Nnu = eng.run_lh(snapshots=S)
    return batch, N.cpu().numpy(), sed.cpu().numpy(), Npr.cpu().numpy(), Nnu.cpu().numpy()


_bh_cache = {}


def bh_surfaces(nu_min, nu_max, disk_cache=True):
    """interp_cs_BH_int(np.logspace(0,8,50), np.logspace(0,10,50), nu_tot[0], nu_tot[-1]) of LeHaMoC.py:263 (host: 27 000
    Simpson integrals as one array expression + two SciPy bisplrep fits, ~0.5 s), kept per frequency range in the
    process and, unless disk_cache is False, on disk (hadronic.interp_cs_BH_int)."""
    from . import hadronic
    key = (float(nu_min), float(nu_max))
    if key not in _bh_cache:
        _bh_cache[key] = hadronic.interp_cs_BH_int(np.logspace(0., 8., 50), np.logspace(0., 10., 50), nu_min, nu_max,
                                                   cache=disk_cache)
    return _bh_cache[key]


def write_outputs(batch, N, sed, Npr, Nnu, out_s, w=0):
    """LeHaMoC.py:440-453: str(float) of log10 values, space separated, four files."""
    nu_nu = np.logspace(10., 22., 50) * eV / h                                   # LeHaMoC.py:173
    names = ["Pairs_Distribution" + out_s + ".txt", "Photons_Distribution" + out_s + ".txt",
             "Protons_Distribution" + out_s + ".txt", "Neutrinos_Distribution" + out_s + ".txt"]
    with np.errstate(divide="ignore"):
        xs = [np.log10(batch.g_el[w]), np.log10(batch.nu_tot[w]), np.log10(batch.g_pr[w]), np.log10(nu_nu)]
        files = [open(n, "w") for n in names]
        try:
            for s in range(int(batch.nsteps[w])):
                ys = [np.log10(N[w, s]), np.log10(sed[w, s]), np.log10(Npr[w, s]), np.log10(Nnu[w, s])]
                for f, x, y in zip(files, xs, ys):
                    for a, b in zip(x, y):
                        f.write(str(a) + " " + str(b) + "\n")
        finally:
            for f in files:
                f.close()
    return names


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) != 3:
        print('incorrect parameters passed')
        print('try something like this')
        print('python -m lehamoc_b200.LeHaMoC Parameters.txt _fileName')
        return 1
    start_time = time.time()
    batch, N, sed, Npr, Nnu = run(read_parameter_file(argv[1]))
    write_outputs(batch, N, sed, Npr, Nnu, argv[2])
    print("--- %s seconds ---" % "{:.2f}".format((time.time() - start_time)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
