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
from .setup_host import pack_script, pack_script_request, read_parameter_file


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


_PIPE_STREAMS = {}


def _pipe_stream(device):
    """The one CUDA stream per device that run_many's engines share."""
    import torch
    if device not in _PIPE_STREAMS:
        _PIPE_STREAMS[device] = torch.cuda.Stream(device=device)
    return _PIPE_STREAMS[device]


def run_many(batches, ic_path="R", engines=None, depth=3, device_pack=True):
    """Evolve a stream of walker batches (an iterable of lists of parameter dicts) and yield
    (batch, n_e [W,Ng], nuLnu [W,Nt]) per batch in order, final state only.

    This is the throughput path for sweeps and samplers: `depth` worker threads each own an engine and do packing
    (NumPy) -> pinned H2D -> set-up + cor-factor + time loop -> D2H for their batch, so the host work of one batch
    overlaps the kernels of the others (NumPy and the ctypes calls release the GIL).  All engines queue their work on ONE
    CUDA stream: the GPU finishes the batches first in, first out.  (With a stream per worker the GPU shares itself
    evenly between the batches in flight, they all complete at the same moment and the next round of host packing finds
    the GPU idle -- measured on the 21-launch time loop of ic_path "G": 8.0 ms per batch on separate streams against 6.1
    of GPU work.)  The yielded spectra are views of page-locked buffers of the engine that produced them; every engine
    alternates between two such buffer pairs, so the arrays of batch k stay valid until `depth` more batches have been
    DRAWN from the generator (batch k + 2*depth, the next user of the same pair, is submitted when batch k + depth + 1
    is requested); copy them if they must live longer.

    device_pack=True (default): the grids, the injection and the scalar block of every walker are built on the device
    (lhm_pack_batch) from one row of scalars per walker; the yielded `batch` is then a setup_host.PackRequest (cfg,
    walker_params, nu_tot / nu_ic / g_el axes).  False: NumPy packing on the host, bit-identical to the reference's
    set-up, and a setup_host.Batch."""
    from concurrent.futures import ThreadPoolExecutor
    import torch
    depth = max(1, int(depth))
    engines = engines if engines is not None else []        # a caller-owned list is filled in place and can be reused
    while len(engines) < depth:
        engines.append(None)
    dev = torch.cuda.current_device()
    stream = next((e.stream for e in engines if e is not None and e.device == dev), None) or _pipe_stream(dev)

    def work(k, sets):
        slot, pair = k % depth, (k // depth) % 2
        batch = pack_script_request(sets) if device_pack else pack_script(sets)
        with torch.cuda.stream(stream):
            eng = engines[slot]
            if (eng is None or eng.max_walkers < batch.W or eng.stream != stream
                    or any(eng.cfg.get(k_) != v for k_, v in batch.cfg.items())):
                eng = engines[slot] = Engine(batch.cfg, max(batch.W, 1), ic_path=ic_path)
            if device_pack:
                N, sed = eng.run_packed_host(batch, out_slot=pair)
            else:
                N, sed = eng.run_host(batch, snapshots=0, pinned=True, out_slot=pair)
        return batch, N, sed

    it = iter(batches)
    with ThreadPoolExecutor(max_workers=depth) as pool:
        window = []
        k = 0
        for sets in it:
            window.append(pool.submit(work, k, sets))
            k += 1
            if len(window) == depth:
                yield window.pop(0).result()
        while window:
            yield window.pop(0).result()


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
