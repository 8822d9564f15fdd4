"""Multi-GPU sharding of the GPR hot path: one process per GPU (torchrun), NCCL only for small all-gathers.

The path shards without any data-path collective: independent optimiser restarts /
hyperparameter candidates go round-robin to ranks, test-event batches go to ranks in contiguous blocks, and
the only exchanges are an all-gather of per-restart ``[loss, theta]`` rows and of mode-space prediction
shards.  Two more axes of BASELINE.json's north_star are covered here: target-column blocks (``shard_columns`` /
``lml_grad_column_sharded``: with one theta shared by all columns the log marginal likelihood and its gradient are sums
over columns, so ranks holding disjoint column blocks exchange ``3 + D`` doubles per evaluation) and events for the
streaming metrics (``metrics_sharded``: whole events go round-robin to ranks, the per-event scalar rows are gathered).  The reference itself is single-process (``gpras/gpr.py:273-274,336-339`` loop sequentially).
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


_TICKET_CALLS = [0]


class _Tickets:
    """Work queue over ranks without a data-path collective: an atomic counter in torch.distributed's key-value store (the
    rendezvous TCPStore on rank 0) hands out restart indices, so a rank whose restarts converge early takes more of them.
    ``dynamic=False`` (or no store) gives the static round-robin ``r -> rank r mod world``."""

    def __init__(self, n_items: int, rank: int, world: int, dynamic: bool):
        import threading

        self.n, self.rank, self.world = n_items, rank, world
        self._lock = threading.Lock()
        self._static = iter(shard_indices(n_items, rank, world).tolist())
        self.store, self.key = None, None
        _TICKET_CALLS[0] += 1  # every rank makes the same sequence of calls, so the key names agree
        if dynamic and world > 1:
            try:
                import torch.distributed as dist

                self.store = dist.distributed_c10d._get_default_store()
                self.key = f"gpras_b200/tickets/{_TICKET_CALLS[0]}"
            except Exception:  # pragma: no cover
                self.store = None

    def next(self):
        with self._lock:
            if self.store is None:
                return next(self._static, None)
            r = int(self.store.add(self.key, 1)) - 1
            return r if r < self.n else None


def run_restarts(model, opt, starts: np.ndarray, opt_kwargs: dict, lanes: int = 1, dynamic: bool = True) -> np.ndarray:
    """Run recipe ``opt`` from every start (rows of constrained [variance, noise, lengthscale(s)]), sharded over ranks;
    every rank ends with the parameters of the lowest final loss (ties -> lowest restart index).  Returns the gathered
    table with rows [restart, loss, theta...] sorted by restart.

    Restarts are handed out by a ticket counter (``_Tickets``): independent restarts differ in how many L-BFGS iterations
    they need, and a static split leaves ranks idle at the end.  ``lanes > 1`` runs that many restarts concurrently per rank
    (host threads, one clone of the model and one device handle each): at N = 8192 a single evaluation leaves the GPU idle
    behind the Cholesky's serial chain, two in flight fill it.

    Every restart begins from the same state: the start's hyperparameters, the model's INITIAL inducing inputs and
    trainable flags (the Z-training recipes move Z, and "diffential_evolution" leaves the hyperparameters frozen), and the
    winner's inducing inputs travel with its hyperparameters, so the model every rank ends with is the one whose loss is
    reported.  A start whose covariance matrix is not positive definite (the reference's ranges reach noise 1e-3 with
    lengthscales of 10, ``gpr.py:88-90``) scores ``+inf`` instead of aborting the whole fit."""
    import time

    from .gpr import _assign_theta

    rank, world, _ = dist_info()
    n_theta = starts.shape[1]
    nls = n_theta - 2
    has_z = bool(getattr(model, "supports_z_training", False))
    z0 = np.array(model.inducing_variable.Z, np.float64) if has_z else np.zeros((0, 0))
    z_trainable0 = bool(model.inducing_variable.trainable) if has_z else False
    flags0 = [p.trainable for p in model.parameters]
    width = 2 + n_theta + z0.size
    tickets = _Tickets(starts.shape[0], rank, world, dynamic)
    lanes = max(1, min(int(lanes), starts.shape[0]))
    if lanes > 1 and not hasattr(model, "lane_models"):
        lanes = 1

    def reset(mdl, theta):
        _assign_theta(mdl, theta)
        for p, f in zip(mdl.parameters, flags0):
            p.trainable = f
        if has_z:
            mdl.inducing_variable.Z = z0.copy()
            mdl.inducing_variable.trainable = z_trainable0

    def lane(mdl):
        rows, busy = [], 0.0
        while True:
            r = tickets.next()
            if r is None:
                break
            t0 = time.perf_counter()
            reset(mdl, starts[r])
            try:
                opt(mdl, **opt_kwargs)
                loss = float(mdl.training_loss())
            except np.linalg.LinAlgError:  # NotPositiveDefiniteError is one
                loss = float("inf")
            z = np.asarray(mdl.inducing_variable.Z, np.float64).ravel() if has_z else np.zeros(0)
            rows.append(np.concatenate([[float(r), loss], mdl.theta()[: 2 + nls], z]))
            busy += time.perf_counter() - t0
        return rows, busy, getattr(mdl, "n_evals", 0)

    if lanes == 1:
        results = [lane(model)]
    else:
        from concurrent.futures import ThreadPoolExecutor

        clones = model.lane_models(lanes)  # one device handle per lane, kept by the model's slot between calls
        try:
            with ThreadPoolExecutor(max_workers=lanes) as ex:
                results = list(ex.map(lane, clones))
        finally:
            for c in clones:  # the lanes' handles go back to the spinning (lowest-latency) host wait
                gp = getattr(getattr(c, "_slot", None), "gp", None)
                if gp is not None:
                    gp.set_blocking_wait(False)
        model.n_evals = getattr(model, "n_evals", 0) + sum(r[2] for r in results)
    rows = [row for res in results for row in res[0]]
    model.restart_busy_s = max((res[1] for res in results), default=0.0)
    table = all_gather_rows(np.array(rows).reshape(-1, width), width)
    table = table[np.argsort(table[:, 0], kind="stable")]
    finite = np.where(np.isfinite(table[:, 1]), table[:, 1], np.inf)
    best = int(np.argmin(finite))
    reset(model, table[best, 2 : 2 + n_theta])
    if has_z:
        model.inducing_variable.Z = table[best, 2 + n_theta :].reshape(z0.shape).copy()
    model.restart_table = table[:, : 2 + n_theta]
    if not np.isfinite(finite[best]):
        from ._lib import NotPositiveDefiniteError

        raise NotPositiveDefiniteError("every restart ended at a hyperparameter vector whose covariance matrix is not positive definite")
    return model.restart_table


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


def shard_columns(n_cols: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, stop) block of target columns for this rank (BASELINE config 4: targets column-sharded)."""
    return shard_rows(n_cols, rank, world)


def all_reduce_sum(vec: np.ndarray) -> np.ndarray:
    """Sum a small float64 vector over ranks (identity when not distributed)."""
    rank, world, dev = dist_info()
    vec = np.asarray(vec, np.float64)
    if world == 1:
        return vec
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(vec.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def lml_grad_column_sharded(evaluate, theta, want_grad: bool = True):
    """Log marginal likelihood and gradient of a shared-theta model whose target columns are sharded over ranks.

    ``evaluate(theta, want_grad) -> (lml, grad)`` is this rank's evaluation on ITS column block (an ``ExactGP`` bound to
    ``y[:, shard_columns(P, rank, world)]``).  With one theta for all columns
    ``LML = sum_p [-1/2 y_p^T K^-1 y_p - 1/2 log|K| - N/2 log 2 pi]`` and its gradient are sums over columns, so the
    exchange is one all-reduce of ``1 + (2 + D)`` doubles.  Every rank factorises K itself: at P << N that is redundant
    work (replicas), at P >> N (raw-cell targets) the 3 N^2 P term that dominates is what gets divided."""
    lml, grad = evaluate(np.asarray(theta, np.float64), want_grad)
    packed = np.concatenate([[lml], grad if want_grad else []])
    total = all_reduce_sum(packed)
    return float(total[0]), (total[1:] if want_grad else None)


def metrics_sharded(event_ids, summarise_event, keys) -> np.ndarray:
    """Scalar metrics of many events, whole events sharded round-robin over ranks.

    ``summarise_event(event_id) -> dict`` runs on this rank's GPU (``gpras_b200.metrics.summarise`` or a
    ``MetricsAccumulator``); returns the gathered table with rows ``[event index, *keys]`` in event order on every rank."""
    rank, world, _ = dist_info()
    rows = []
    for i in shard_indices(len(event_ids), rank, world):
        s = summarise_event(event_ids[i])
        rows.append([float(i)] + [float(s[k]) for k in keys])
    table = all_gather_rows(np.array(rows).reshape(-1, 1 + len(keys)), 1 + len(keys))
    return table[np.argsort(table[:, 0], kind="stable")]


def run_models_sharded(models, run_one, run_many=None) -> None:
    """Per-column models (the reference's default: one independent model per spatial mode, ``gpr.py:273-274``) sharded
    round-robin over ranks: rank r optimises models r, r + world, ...; the optimised parameters (variance, likelihood
    variance, lengthscales, inducing inputs) are all-gathered so that every rank ends with every model.  One exchange of
    ``P x (2 + n_ls + M D)`` doubles at the end, nothing during optimisation.  ``run_many(list_of_models) -> bool`` (optional)
    trains a rank's whole shard at once (the device-resident batched trainer); when it declines, ``run_one`` runs per model."""
    rank, world, _ = dist_info()
    mine = shard_indices(len(models), rank, world)
    if not (run_many is not None and len(mine) > 0 and run_many([models[i] for i in mine])):
        for i in mine:
            run_one(models[i])

    def pack(i):
        d = models[i].parameter_dict()
        return np.concatenate([[float(i)], np.atleast_1d(d[".kernel.variance"]).ravel(), np.atleast_1d(d[".likelihood.variance"]).ravel(),
                               np.atleast_1d(d[".kernel.lengthscales"]).ravel(), np.asarray(d[".inducing_variable.Z"], np.float64).ravel()])

    if world == 1 or not models:
        return
    width = pack(0).size
    table = all_gather_rows(np.array([pack(i) for i in mine]).reshape(-1, width), width)
    nls = np.atleast_1d(models[0].parameter_dict()[".kernel.lengthscales"]).size
    zshape = np.asarray(models[0].parameter_dict()[".inducing_variable.Z"]).shape
    for row in table:
        i = int(row[0])
        if i % world == rank:
            continue
        ls = row[3 : 3 + nls]
        models[i].assign_parameters({".kernel.variance": row[1], ".likelihood.variance": row[2],
                                     ".kernel.lengthscales": ls if nls > 1 else ls[0],
                                     ".inducing_variable.Z": row[3 + nls :].reshape(zshape)})
