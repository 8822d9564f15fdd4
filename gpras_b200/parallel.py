"""Multi-GPU sharding of the GPR hot path: one process per GPU (torchrun), NCCL only for small all-gathers.

The path shards without any data-path collective: independent optimiser restarts /
hyperparameter candidates go round-robin to ranks, test-event batches go to ranks in contiguous blocks, and
the only exchanges are an all-gather of per-restart ``[loss, theta]`` rows and of mode-space prediction
shards.  The reference itself is single-process (``gpras/gpr.py:273-274,336-339`` loop sequentially).
On CPU test runs the same code runs over ``gloo``.
"""

from __future__ import annotations

import numpy as np


def dist_info():
    """(rank, world_size, device_for_collectives) -- (0, 1, None) when torch.distributed is not initialised."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return 0, 1, None
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1, None
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    return dist.get_rank(), dist.get_world_size(), dev


def shard_indices(n_items: int, rank: int, world: int) -> np.ndarray:
    """Round-robin assignment used for restarts / candidates: item r -> rank r mod world."""
    return np.arange(rank, n_items, world)


def shard_rows(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, stop) block of test rows for this rank (first ranks take the remainder)."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_rows(local: np.ndarray, n_cols: int) -> np.ndarray:
    """All-gather variable-length (k_r, n_cols) float64 blocks; returns the concatenation in rank order."""
    rank, world, dev = dist_info()
    local = np.asarray(local, np.float64).reshape(-1, n_cols)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    kmax = int(counts.max().item())
    buf = torch.zeros((kmax, n_cols), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    out = torch.empty((world * kmax, n_cols), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, kmax, n_cols)
    return np.concatenate([out[r, : int(counts[r].item())] for r in range(world)], axis=0)


def run_restarts(model, opt, starts: np.ndarray, opt_kwargs: dict) -> np.ndarray:
    """Run recipe ``opt`` from every start (rows of constrained [variance, noise, lengthscale(s)]), sharded
    round-robin over ranks; every rank ends with the parameters of the lowest final loss (ties -> lowest
    restart index).  Returns the gathered table with rows [restart, loss, theta...] sorted by restart."""
    from .gpr import _assign_theta

    rank, world, _ = dist_info()
    n_theta = starts.shape[1]
    rows = []
    for r in shard_indices(starts.shape[0], rank, world):
        _assign_theta(model, starts[r])
        opt(model, **opt_kwargs)
        loss = model.training_loss()
        th = model.theta()
        nls = n_theta - 2
        rows.append(np.concatenate([[float(r), float(loss)], th[: 2 + nls]]))
    table = all_gather_rows(np.array(rows).reshape(-1, 2 + n_theta), 2 + n_theta)
    table = table[np.argsort(table[:, 0], kind="stable")]
    finite = np.where(np.isfinite(table[:, 1]), table[:, 1], np.inf)
    best = int(np.argmin(finite))
    _assign_theta(model, table[best, 2:])
    model.restart_table = table
    return table


def predict_sharded(gpras, x: np.ndarray):
    """Shard test events across ranks in contiguous blocks, predict locally, all-gather the mode-space
    (T, P) means and variances (cell-space output stays sharded by construction)."""
    rank, world, _ = dist_info()
    x = np.asarray(x, np.float64)
    lo, hi = shard_rows(x.shape[0], rank, world)
    mean, var = gpras.predict(x[lo:hi]) if hi > lo else (np.zeros((0, gpras.y.shape[1])), np.zeros((0, gpras.y.shape[1])))
    p = mean.shape[1]
    both = all_gather_rows(np.concatenate([mean, var], axis=1), 2 * p)
    return both[:, :p], both[:, p:]
