"""The multi-GPU path on real GPUs (needs >= 2 of them: `gpurun --gpus 2 -- python -m pytest tests/test_nccl_gpu.py -m gpu`).

What it replaces is `Pool().map(log_probability, thetas)` of Fit_emcee/fit_emcee.py:239-242: the walkers of an ensemble
are cut into per-rank blocks (lehamoc_b200/shard.py), every rank evolves its block with the fused kernel and ONE NCCL
all-gather gives every rank the rows of all walkers.  The gathered rows must equal the single-GPU evaluation of the
same walkers bit for bit, on every rank, for even and ragged block sizes.
"""
import os
import socket

import numpy as np
import pytest

from util import GOLD, load_golden

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _thetas(W):
    from lehamoc_b200.presets import fit_walkers
    rng = np.random.default_rng(5)
    th = np.column_stack([fit_walkers(W, seed=77), rng.uniform(0.5, 2.5, W), rng.uniform(-5., 0., W)])
    th[3, 0] = 17.5                                 # one walker outside the prior box: -inf, never evolved
    return th


def _evaluator(gold, P, device):
    from lehamoc_b200 import fit
    return fit.FitEvaluator(gold["xd"], gold["yd"], gold["yerr_hi"], gold["yerr_lo"], gold["flags"], param_file=P,
                            D=float(gold["D_Mpc"]), z=float(gold["z"]), device=device)


def _rows(ev, block):
    """[w, Nt + 1] device rows: the SED and the log-likelihood of every walker of the block."""
    import torch
    logl, _ = ev.model_and_loglike(block)
    # the SED of the block: evolve it again on the evaluator's engine (same kernels, same bits)
    from lehamoc_b200.setup_host import pack_code_request
    req = pack_code_request(block[:, :6], ev.frozen)
    ev.eng.setup_packed(req)
    _, s = ev.eng.run(want_N=False)
    return torch.cat([s, logl[:, None]], dim=1)


def _worker(rank, ws, port, W, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from lehamoc_b200 import shard
        gold = np.load(os.path.join(GOLD, "fit_loglike.npz"))
        _, P = load_golden("code_fit")
        ev = _evaluator(gold, P, rank)
        th = _thetas(W)
        from lehamoc_b200 import fit
        lp = ev.log_probability_sharded(th)                                     # fit_emcee.py:239-242 semantics
        inside = np.isfinite(fit.log_prior_batch(th))
        rows = shard.evaluate_sharded(lambda blk: _rows(ev, blk), th[inside])
        np.save(os.path.join(out_dir, f"lp_{W}_{rank}.npy"), lp)
        np.save(os.path.join(out_dir, f"rows_{W}_{rank}.npy"), rows)
        # a rank whose block cannot be packed: every rank raises (nobody is left waiting in the gather), and the
        # evaluator is usable afterwards
        real = fit.pack_code_request
        if rank == 1:
            def broken(*a, **k):
                raise ValueError("this rank's block cannot be packed")
            fit.pack_code_request = broken
        try:
            ev.log_probability_sharded(th)
            raised = "nothing"
        except Exception as e:                                                  # noqa: BLE001
            raised = type(e).__name__
        fit.pack_code_request = real
        lp2 = ev.log_probability_sharded(th)
        with open(os.path.join(out_dir, f"raised_{W}_{rank}.txt"), "w") as f:
            f.write(raised)
        np.save(os.path.join(out_dir, f"lp2_{W}_{rank}.npy"), lp2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W", [64, 37])
def test_nccl_gathered_rows_equal_single_gpu(tmp_path, W):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    ws = 2
    mp.spawn(_worker, args=(ws, _free_port(), W, str(tmp_path)), nprocs=ws, join=True)
    gold = np.load(os.path.join(GOLD, "fit_loglike.npz"))
    _, P = load_golden("code_fit")
    ev = _evaluator(gold, P, 0)
    th = _thetas(W)
    from lehamoc_b200 import fit
    lp = ev.log_probability(th)                                                 # unsharded, one GPU
    inside = np.isfinite(fit.log_prior_batch(th))
    # -inf also where an upper limit is violated by many sigma (log(1 + erf(-x)) underflows, in the reference as well)
    assert lp[3] == -np.inf and not inside[3] and inside.sum() == W - 1 and np.isfinite(lp).sum() > W // 4
    assert not np.any(np.isnan(lp))
    rows = _rows(ev, th[inside]).cpu().numpy()
    for r in range(ws):
        np.testing.assert_array_equal(np.load(tmp_path / f"lp_{W}_{r}.npy"), lp)
        np.testing.assert_array_equal(np.load(tmp_path / f"rows_{W}_{r}.npy"), rows)
        np.testing.assert_array_equal(np.load(tmp_path / f"lp2_{W}_{r}.npy"), lp)
        assert (tmp_path / f"raised_{W}_{r}.txt").read_text() == ("ValueError" if r == 1 else "RuntimeError")
