"""World-size-2 tests of the walker sharding on CPU (gloo): block partition, ragged all-gather, and a sharded
ensemble run that must reproduce the single-process chain bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lehamoc_b200 import shard


def test_block_partition():
    for W in (1, 2, 7, 48, 1024, 1025):
        for ws in (1, 2, 3, 8):
            r = [shard.block_range(W, ws, k) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == W
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            sizes = shard.block_sizes(W, ws)
            assert sum(sizes) == W and max(sizes) - min(sizes) <= 1


def _fake_model(thetas):
    """Deterministic stand-in for the GPU model: one 'SED' row + log-prob per walker."""
    th = np.asarray(thetas, dtype=np.float64)
    x = np.linspace(0., 1., 5)
    sed = np.sin(th[:, :1] * x[None, :]) + th[:, 1:2]
    logp = -0.5 * np.sum((th - 0.3) ** 2, axis=1)
    return np.column_stack([sed, logp])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, W, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from lehamoc_b200 import fit, shard as sh
        rng = np.random.default_rng(7)
        thetas = rng.standard_normal((W, 3))
        # 1. gathered rows == unsharded evaluation, every rank holds the full result
        full = sh.evaluate_sharded(_fake_model, thetas)
        np.save(os.path.join(out_dir, f"full_{W}_{rank}.npy"), full)
        # 2. gather_rows on tensors with a trailing shape
        a, b = sh.block_range(W, ws, rank)
        loc = torch.arange(a, b, dtype=torch.float64)[:, None, None].expand(b - a, 2, 3).contiguous()
        g = sh.gather_rows(loc, W)
        assert g.shape == (W, 2, 3) and torch.equal(g[:, 0, 0], torch.arange(W, dtype=torch.float64))
        # 3. sharded ensemble run: only log_prob_fn is sharded, the chain is replicated
        if W % 2 == 0 and W >= 6:
            fn = lambda th: sh.evaluate_sharded(lambda blk: _fake_model(blk)[:, -1], th)
            res = fit.run_mcmc(fn, thetas, 20, seed=11)
            np.save(os.path.join(out_dir, f"chain_{W}_{rank}.npy"), res["chain"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W", [8, 7])
def test_world2_gather_and_mcmc(tmp_path, W):
    ws = 2
    port = _free_port()
    mp.spawn(_worker, args=(ws, port, W, str(tmp_path)), nprocs=ws, join=True)
    rng = np.random.default_rng(7)
    thetas = rng.standard_normal((W, 3))
    ref = _fake_model(thetas)
    for r in range(ws):
        np.testing.assert_array_equal(np.load(tmp_path / f"full_{W}_{r}.npy"), ref)
    if W % 2 == 0:
        from lehamoc_b200 import fit
        single = fit.run_mcmc(lambda th: _fake_model(th)[:, -1], thetas, 20, seed=11)["chain"]
        for r in range(ws):
            np.testing.assert_array_equal(np.load(tmp_path / f"chain_{W}_{r}.npy"), single)


def _failing_worker(rank, ws, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from lehamoc_b200 import shard as sh

        def fn(block):
            if rank == 1:
                raise ValueError("walker outside the gamma grid")
            return _fake_model(block)
        try:
            sh.evaluate_sharded(fn, np.zeros((6, 3)))
            what = "returned"
        except ValueError as e:
            what = f"ValueError: {e}"
        except RuntimeError as e:
            what = f"RuntimeError: {e}"
        with open(os.path.join(out_dir, f"outcome_{rank}.txt"), "w") as f:
            f.write(what)
        # the group is still usable afterwards
        np.save(os.path.join(out_dir, f"after_{rank}.npy"), sh.evaluate_sharded(_fake_model, np.ones((4, 3))))
    finally:
        dist.destroy_process_group()


def test_world2_failure_on_one_rank_raises_on_all(tmp_path):
    """A rank whose block cannot be evaluated must not leave the others hanging in the all-gather: the failing rank
    re-raises its own exception, the others raise a RuntimeError, and the process group stays usable."""
    mp.spawn(_failing_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "outcome_0.txt").read_text().startswith("RuntimeError: evaluate_sharded: another rank failed")
    assert (tmp_path / "outcome_1.txt").read_text() == "ValueError: walker outside the gamma grid"
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"after_{r}.npy"), _fake_model(np.ones((4, 3))))
