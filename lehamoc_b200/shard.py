"""Walker sharding over the GPUs of one box.

Walkers (independent parameter sets: emcee walkers, sweep points) never exchange data during the time
integration -- the reference's only parallel construct is `multiprocessing.Pool().map(log_probability, thetas)`
(Fit_emcee/fit_emcee.py:239-242).  So the batch is cut into contiguous blocks, one per rank (one process per GPU),
each rank runs the fused kernel on its block, and ONE all-gather of the per-walker result rows (SED and/or
log-likelihood) gives every rank the full result, so that all ranks can run the identical ensemble update.
There is no other collective on the path.

Works with any initialised `torch.distributed` backend: `nccl` on GPUs (device tensors, NVLink), `gloo` on CPU
(the world-size-2 tests in tests/test_shard_gloo.py).  Without an initialised process group everything
degenerates to world size 1.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) if there is none."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def block_range(W: int, world_size: int, rank: int):
    """Contiguous block [start, stop) of rank `rank`; the first W % world_size ranks get one extra walker."""
    base, rem = divmod(int(W), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def block_sizes(W: int, world_size: int):
    return [block_range(W, world_size, r)[1] - block_range(W, world_size, r)[0] for r in range(world_size)]


def gather_rows(local: torch.Tensor, W: int, group=None) -> torch.Tensor:
    """All-gather the row blocks of every rank into the full [W, ...] tensor (same on all ranks).

    `local` holds this rank's rows in block order.  Equal blocks use one `all_gather_into_tensor` straight into
    the output; ragged blocks (W not divisible by the world size) are padded to the largest block first."""
    rank, ws = world()
    if ws == 1:
        assert local.shape[0] == W
        return local
    sizes = block_sizes(W, ws)
    assert local.shape[0] == sizes[rank], f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}"
    local = local.contiguous()
    tail = tuple(local.shape[1:])
    if min(sizes) == max(sizes):
        out = torch.empty((W,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    m = max(sizes)
    padded = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
    padded[:sizes[rank]] = local
    buf = torch.empty((ws * m,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * m:r * m + sizes[r]] for r in range(ws)], dim=0)


def evaluate_sharded(fn, thetas, device=None, group=None):
    """Evaluate `fn(thetas_block) -> rows [w_local, ...]` on this rank's block of `thetas` [W, ndim] and return the
    gathered rows of the whole batch as a NumPy array [W, ...] (identical on every rank).

    `fn` may return a NumPy array or a torch tensor (CPU or CUDA); with NCCL the gather runs on `device`."""
    thetas = np.asarray(thetas)
    W = thetas.shape[0]
    rank, ws = world()
    a, b = block_range(W, ws, rank)
    if ws == 1:
        rows = fn(thetas[a:b])
        return rows.detach().cpu().numpy() if isinstance(rows, torch.Tensor) else np.asarray(rows)
    # a rank whose block fails (e.g. a packer ValueError for one of ITS walkers) must not leave the others waiting in
    # the all-gather: every rank first learns whether all blocks were evaluated, then all raise together
    err = None
    try:
        rows = fn(thetas[a:b])
    except Exception as e:                                  # noqa: BLE001 -- re-raised below, on every rank
        err, rows = e, None
    nccl = dist.get_backend(group) == "nccl"
    flag_dev = (device if device is not None else torch.device("cuda", torch.cuda.current_device())) if nccl else "cpu"
    flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=flag_dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if int(flag.item()):
        if err is not None:
            raise err
        raise RuntimeError("evaluate_sharded: another rank failed to evaluate its block of walkers")
    t = rows if isinstance(rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rows))
    if t.dim() == 0 or t.shape[0] != b - a:
        raise ValueError("fn must return one row per walker of its block")
    if dist.get_backend(group) == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    else:
        t = t.cpu()
    return gather_rows(t, W, group=group).cpu().numpy()
