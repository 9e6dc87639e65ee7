"""Drop-in for `python LeMoC.py <Parameters.txt> <suffix>` (LeMoC/LeMoC.py:45-58,290-296):

    python -m lehamoc_b200.LeMoC Parameters_Test1.txt out1

writes Particles_Distribution_<suffix>.txt and Photons_Distribution_<suffix>.txt in the reference's text
format (one `log10 x log10 y` row per bin, one block per timestep, Python float repr).  `run()` is the
importable form and also takes several parameter sets at once."""
from __future__ import annotations

import sys
import time

import numpy as np

from .engine import Engine
from .setup_host import pack_script, read_parameter_file


def run(param_sets, ic_path="R", snapshots=True, engine=None, pinned=False):
    """Evolve one or more LeMoC.py parameter sets.  Returns (batch, n_e [W,S,Ng], nuLnu [W,S,Nt]); S = number
    of timesteps when `snapshots` (the reference writes every step), else 1."""
    batch = pack_script(param_sets)
    eng = engine or Engine(batch.cfg, max(batch.W, 1), ic_path=ic_path)
    S = int(batch.nsteps.max()) if snapshots else 0
    N, sed = eng.run_host(batch, snapshots=S, pinned=pinned)
    if not snapshots:
        N, sed = N[:, None], sed[:, None]
    return batch, N, sed


def write_outputs(batch, N, sed, out_s, w=0):
    """LeMoC.py:290-296: str(float) of log10 values, space separated."""
    out1 = "Particles_Distribution_" + out_s + ".txt"
    out2 = "Photons_Distribution_" + out_s + ".txt"
    with np.errstate(divide="ignore"):
        lg, lnu = np.log10(batch.g_el[w]), np.log10(batch.nu_tot[w])
        with open(out1, "w") as f1, open(out2, "w") as f2:
            for s in range(int(batch.nsteps[w])):
                for a, b in zip(lg, np.log10(N[w, s])):
                    f1.write(str(a) + " " + str(b) + "\n")
                for a, b in zip(lnu, np.log10(sed[w, s])):
                    f2.write(str(a) + " " + str(b) + "\n")
    return out1, out2


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) != 3:
        print('incorrect parameters passed')
        print('try something like this')
        print('python -m lehamoc_b200.LeMoC Parameters.txt out1')
        return 1
    start_time = time.time()
    batch, N, sed = run(read_parameter_file(argv[1]))
    write_outputs(batch, N, sed, argv[2])
    print("--- %s seconds ---" % "{:.2f}".format((time.time() - start_time)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
