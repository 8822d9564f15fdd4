"""Sparse (inducing-point) model objects: the reference's own model family.

``gpras/gpr.py:293-308`` builds one ``gpflow.models.SGPR`` per target column with shared initial inducing inputs,
LogNormal(0, 1) priors on the three constrained hyperparameters and trainable ``Z``.  ``SparseModel`` exposes the same
attribute names the reference's recipes touch (``kernel.variance / .lengthscales``, ``likelihood.variance``,
``inducing_variable.Z``, ``data``, ``trainable_variables``, ``training_loss()``) and evaluates
``training_loss = -(ELBO + log prior)`` with its gradient on the GPU (``gpras_sgpr_elbo_grad``).
"""

from __future__ import annotations

import threading

import numpy as np

from .engine import SparseGP
from .gpr import _Inducing, _Kernel, _Likelihood

JITTER = 1e-6  # gpflow.config.default_jitter()

_HANDLES: dict = {}


def _handle(kernel: str, n: int, d: int, m: int, r: int, device: int) -> dict:
    """One device handle per calling thread and problem shape; per-column models take turns on it (they run
    sequentially in the reference, gpr.py:273-274; ``fit(n_jobs > 1)`` uses one thread and handle per model in flight)."""
    key = (threading.get_ident(), kernel, n, d, m, r, device)
    if key not in _HANDLES:
        _HANDLES[key] = {"gp": SparseGP(kernel, n, d, m, r, device=device), "owner": None}
    return _HANDLES[key]


def release_other_threads() -> None:
    me = threading.get_ident()
    for key in [k for k in _HANDLES if k[0] != me]:
        _HANDLES.pop(key)["gp"].close()


class SparseModel:
    supports_z_training = True
    device_batch = True  # may join a device batch (``SparseBatch``); test doubles that replace the device evaluation say False

    def __init__(self, kernel_name, x, y, z, lengthscales, device: int = 0, priors: bool = True,
                 parameterisation: str = "softplus"):
        self.x, self.y = x, y
        self.data = (x, y)
        self.device = device
        self.kernel = _Kernel(kernel_name, 1.0, lengthscales, parameterisation)
        self.likelihood = _Likelihood(1.0, parameterisation)
        self.inducing_variable = _Inducing(z, trainable=True)
        if not priors:
            self.kernel.variance.prior = self.kernel.lengthscales.prior = self.likelihood.variance.prior = None
        self.n_evals = 0

    # -- device plumbing --
    def _gp(self) -> SparseGP:
        n, d = self.x.shape
        slot = _handle(self.kernel.name, n, d, self.inducing_variable.Z.shape[0], self.y.shape[1], self.device)
        if slot["owner"] is not self:
            slot["gp"].set_data(self.x, self.y)
            slot["owner"] = self
        return slot["gp"]

    # -- parameter plumbing (same protocol as ExactModel) --
    @property
    def parameters(self):
        return [self.kernel.variance, self.likelihood.variance, self.kernel.lengthscales]

    @property
    def trainable_variables(self):
        out = [p.unconstrained for p in self.parameters if p.trainable]
        if self.inducing_variable.trainable:
            out.append(self.inducing_variable.Z)
        return out

    def set_trainable(self, flag: bool, hypers: bool = True) -> None:
        if hypers:
            for p in self.parameters:
                p.trainable = flag

    def get_u(self):
        parts = [p.unconstrained for p in self.parameters if p.trainable]
        if self.inducing_variable.trainable:
            parts.append(np.asarray(self.inducing_variable.Z, np.float64).ravel())
        return np.concatenate(parts) if parts else np.zeros(0)

    def set_u(self, u) -> None:
        u = np.asarray(u, np.float64)
        o = 0
        for p in self.parameters:
            if p.trainable:
                p.unconstrained = u[o : o + p.size].copy()
                o += p.size
        if self.inducing_variable.trainable:
            z = self.inducing_variable.Z
            self.inducing_variable.Z = u[o : o + z.size].reshape(z.shape).copy()

    def theta(self):
        d = self.x.shape[1]
        ls = np.atleast_1d(self.kernel.lengthscales.numpy())
        if ls.size == 1:
            ls = np.full(d, ls[0])
        return np.concatenate([[self.kernel.variance.numpy(), self.likelihood.variance.numpy()], ls])

    def _log_prior(self) -> float:
        return sum(p.log_prior() for p in self.parameters if p.trainable)

    # -- objective --
    def training_loss(self) -> float:
        elbo, _, _ = self._gp().elbo_grad(self.theta(), np.asarray(self.inducing_variable.Z, np.float64), JITTER, want_grad=False)
        self.n_evals += 1
        return -(elbo + self._log_prior())

    def loss_and_grad(self, u=None):
        if u is not None:
            self.set_u(u)
        elbo, gt, gz = self._gp().elbo_grad(self.theta(), np.asarray(self.inducing_variable.Z, np.float64), JITTER)
        return self._assemble(elbo, gt, gz)

    # -- asynchronous form on a handle of the model's own (lock-stepped training of the per-column models) --
    def bind(self, gp: SparseGP) -> None:
        gp.set_data(self.x, self.y)
        self._own = gp

    def enqueue_loss_and_grad(self, u) -> None:
        self.set_u(u)
        self._own.enqueue(self.theta(), np.asarray(self.inducing_variable.Z, np.float64), JITTER, True)

    def fetch_loss_and_grad(self):
        return self._assemble(*self._own.fetch())

    def _assemble(self, elbo, gt, gz):
        """Loss and gradient w.r.t. the trainable unconstrained variables from the device's ELBO and its gradients
        (log-prior terms and the softplus chain rule are host logic)."""
        self.n_evals += 1
        g_ls = gt[2:] if self.kernel.lengthscales.size > 1 else np.array([gt[2:].sum()])
        parts = []
        for p, gl in ((self.kernel.variance, np.array([gt[0]])), (self.likelihood.variance, np.array([gt[1]])),
                      (self.kernel.lengthscales, g_ls)):
            if p.trainable:
                v = p.value()
                parts.append(-((gl / v + p.dlog_prior_dvalue()) * p.dvalue_du()))
        if self.inducing_variable.trainable:
            parts.append(-gz.ravel())
        return -(elbo + self._log_prior()), (np.concatenate(parts) if parts else np.zeros(0))

    # -- prediction --
    def predict_y(self, xs, keep_handle: bool = False):
        """``keep_handle``: a device handle of the model's own for prediction, whose conditioned state survives while the other
        per-column models are predicted and across calls (``gpras/preprocess.py:601-606`` predicts once per plan)."""
        if keep_handle:
            if getattr(self, "_pred_gp", None) is None:
                n, d = self.x.shape
                self._pred_gp = SparseGP(self.kernel.name, n, d, self.inducing_variable.Z.shape[0], self.y.shape[1], device=self.device)
                self._pred_gp.set_data(self.x, self.y)
            gp = self._pred_gp
        else:
            gp = self._gp()
        gp.condition(self.theta(), np.asarray(self.inducing_variable.Z, np.float64), JITTER)
        return gp.predict(np.asarray(xs, np.float64))

    def release(self) -> None:
        gp = getattr(self, "_pred_gp", None)
        if gp is not None:
            gp.close()
            self._pred_gp = None

    def parameter_dict(self) -> dict:
        return {
            ".kernel.variance": np.asarray(self.kernel.variance.numpy()),
            ".kernel.lengthscales": np.asarray(self.kernel.lengthscales.numpy()),
            ".likelihood.variance": np.asarray(self.likelihood.variance.numpy()),
            ".inducing_variable.Z": np.asarray(self.inducing_variable.Z),
        }

    def assign_parameters(self, d: dict) -> None:
        self.kernel.variance.assign(d[".kernel.variance"])
        self.kernel.lengthscales.assign(d[".kernel.lengthscales"])
        self.likelihood.variance.assign(d[".likelihood.variance"])
        self.inducing_variable.Z = np.asarray(d[".inducing_variable.Z"], np.float64)


# ------------------------------------------------------------------------------------------------
# lock-stepped training of the per-column models
# ------------------------------------------------------------------------------------------------
_MODEL_POOLS: dict = {}


def _model_pool(kernel: str, n: int, d: int, m: int, r: int, device: int, count: int) -> list:
    """``count`` sparse handles of one problem shape, kept for the life of the process (creating one costs several ms)."""
    pool = _MODEL_POOLS.setdefault((threading.get_ident(), kernel, n, d, m, r, device), [])
    while len(pool) < count:
        pool.append(SparseGP(kernel, n, d, m, r, device=device))
    return pool[:count]


def _exact_pool(kernel: str, n: int, d: int, p: int, device: int, count: int) -> list:
    from .engine import ExactGP

    pool = _MODEL_POOLS.setdefault((threading.get_ident(), "exact", kernel, n, d, p, device), [])
    while len(pool) < count:
        pool.append(ExactGP(kernel, n, d, p, device=device))
    return pool[:count]


def adam_lockstep(models, max_iter: int, learning_rate: float = 0.001) -> None:
    """``_optimize_adam`` (Keras Adam + the reference's early-stopping rule, ``gpr.py:147-173``) for SEVERAL independent models
    at once: every round enqueues one loss+gradient evaluation per still-active model on that model's own handle, then
    fetches them, so the evaluations of the P per-column models overlap on the GPU instead of running one after the
    other (``gpr.py:273-274``).  Each model follows exactly the trajectory of the one-model loop."""
    u = [mdl.get_u() for mdl in models]
    active = [i for i, v in enumerate(u) if v.size]
    mom = [np.zeros_like(v) for v in u]
    vel = [np.zeros_like(v) for v in u]
    best, count = [np.inf] * len(models), [0] * len(models)
    b1, b2, eps, tol, patience = 0.9, 0.999, 1e-7, 10e-6, 50
    for t in range(1, int(max_iter) + 1):
        if not active:
            break
        for i in active:
            models[i].enqueue_loss_and_grad(u[i])
        still = []
        for k, i in enumerate(active):
            try:
                loss, g = models[i].fetch_loss_and_grad()
            except Exception:
                for j in active[k + 1:]:  # leave no evaluation pending on the other handles
                    try:
                        models[j].fetch_loss_and_grad()
                    except Exception:
                        pass
                raise
            mom[i] = b1 * mom[i] + (1.0 - b1) * g
            vel[i] = b2 * vel[i] + (1.0 - b2) * g * g
            alpha = learning_rate * np.sqrt(1.0 - b2**t) / (1.0 - b1**t)
            u[i] = u[i] - alpha * mom[i] / (np.sqrt(vel[i]) + eps)
            models[i].set_u(u[i])
            if ((best[i] - loss) / abs(loss)) > tol:
                best[i], count[i] = loss, 0
                still.append(i)
            else:
                count[i] += 1
                if count[i] <= patience:
                    still.append(i)
        active = still


_BATCH_POOL: dict = {}


def _lib_errors():
    from ._lib import GprasError

    return (GprasError,)


def release_batches(models=None) -> None:
    """Close the pooled device batches of every thread (device memory goes back to the driver); with ``models`` only the
    batches that currently hold those models."""
    ids = None if models is None else {id(m) for m in models}
    for key in list(_BATCH_POOL):
        batch = _BATCH_POOL[key]
        if ids is None or ids & set(getattr(batch, "_owners", None) or ()):
            _BATCH_POOL.pop(key).close()


def _device_batch(models):
    """A ``SparseBatch`` holding all of ``models`` when they qualify for the device-resident trainer (one target column each over
    the same inputs, the same kernel / transform / prior configuration, at most 128 inducing points), else None."""
    from .engine import SparseBatch

    m0 = models[0]
    if not all(isinstance(mdl, SparseModel) and mdl.device_batch for mdl in models):
        return None
    m, d = m0.inducing_variable.Z.shape
    cfg0 = _trainer_config(m0)
    for mdl in models:
        if (mdl.y.shape[1] != 1 or mdl.inducing_variable.Z.shape != (m, d) or mdl.device != m0.device or _trainer_config(mdl) != cfg0
                or not (mdl.x is m0.x or np.array_equal(mdl.x, m0.x))):
            return None
    if m > SparseBatch.MAX_INDUCING or cfg0 is None:
        return None
    n = m0.x.shape[0]
    key = (threading.get_ident(), m0.kernel.name, n, d, m, len(models), m0.device)
    if key not in _BATCH_POOL:
        for old in [k for k in _BATCH_POOL if k[0] == key[0]]:  # one batch per thread: a new shape replaces the old arena
            _BATCH_POOL.pop(old).close()
        try:
            _BATCH_POOL[key] = SparseBatch(m0.kernel.name, n, d, m, len(models), device=m0.device)
        except _lib_errors() as exc:  # e.g. the arena of all models does not fit: one handle per model still works
            if "cudaMalloc" not in str(exc):
                raise
            return None
    batch = _BATCH_POOL[key]
    owners = tuple(id(mdl) for mdl in models)
    if getattr(batch, "_owners", None) != owners:  # (a model's data never change after construction)
        batch.set_data(m0.x, np.concatenate([mdl.y for mdl in models], axis=1))
        batch._owners = owners
        batch._keep = list(models)  # keeps the ids alive
    return batch


def predict_batched(models, xs):
    """``predict_y`` of all per-column sparse models in one device pass per tile of test inputs (``gpras_sgpr_batch_predict``),
    or None when they do not qualify (then the caller predicts model by model)."""
    if not models or len(models) < 2:
        return None
    batch = _device_batch(models)
    if batch is None or not batch.fused:
        return None
    theta = np.stack([mdl.theta() for mdl in models])
    z = np.stack([np.asarray(mdl.inducing_variable.Z, np.float64) for mdl in models])
    batch.condition(theta, z, JITTER)
    return batch.predict(xs)


def _trainer_config(mdl):
    """(kernel, transform, priors, noise floor, number of lengthscales) of a model, or None if its three hyperparameters are not
    configured alike (the device trainer applies one transform and one prior setting to all of them)."""
    ps = mdl.parameters
    tr = {p.transform for p in ps}
    pr = {p.prior for p in ps}
    if len(tr) != 1 or len(pr) != 1 or mdl.kernel.variance.lower != 0.0 or mdl.kernel.lengthscales.lower != 0.0:
        return None
    if mdl.kernel.variance.size != 1 or mdl.likelihood.variance.size != 1:
        return None
    return (mdl.kernel.name, tr.pop(), pr.pop(), mdl.likelihood.variance.lower, mdl.kernel.lengthscales.size)


def adam_device(models, batch, max_iter: int, learning_rate: float = 0.001, rule: str = "adam") -> bool:
    """``_optimize_adam`` (``gpr.py:147-173``; ``rule="adadelta"``: ``_optimize_adadelta``, ``gpr.py:176-192``) of all ``models``
    as ONE device-resident loop (``gpras_sgpr_batch_train``): no host round trip per step.  Returns False when the models'
    trainable flags are not one of the recipes' stages."""
    m0 = models[0]
    flags = [[p.trainable for p in mdl.parameters] + [mdl.inducing_variable.trainable] for mdl in models]
    if any(f != flags[0] for f in flags) or len(set(flags[0][:3])) != 1:
        return False
    train_h, train_z = flags[0][0], flags[0][3]
    if not (train_h or train_z):
        return True  # nothing trainable: _optimize_adam returns at once
    _, transform, prior, floor, n_ls = _trainer_config(m0)
    u0 = np.stack([np.concatenate([mdl.kernel.variance.unconstrained, mdl.likelihood.variance.unconstrained,
                                   mdl.kernel.lengthscales.unconstrained, np.asarray(mdl.inducing_variable.Z, np.float64).ravel()])
                   for mdl in models])
    u, losses, iters = batch.adam(u0, n_ls, train_h, train_z, int(max_iter), learning_rate, JITTER, transform, prior is not None, floor,
                                  rule=rule)
    for b, mdl in enumerate(models):
        if train_h:
            mdl.kernel.variance.unconstrained = u[b, 0:1].copy()
            mdl.likelihood.variance.unconstrained = u[b, 1:2].copy()
            mdl.kernel.lengthscales.unconstrained = u[b, 2:2 + n_ls].copy()
        if train_z:
            z = mdl.inducing_variable.Z
            mdl.inducing_variable.Z = u[b, 2 + n_ls:].reshape(z.shape).copy()
        mdl.n_evals += int(iters[b])
        mdl.adam_losses = losses[: int(iters[b]), b].copy()
    return True


def multi_start_device(model, rng, n_starts: int, iter_initial: int, starts, pick_best: bool, train_z: bool):
    """Coarse stage of the "stochastic" recipe (``gpr.py:73-101``) for ONE sparse model with all starts as the models of a
    device batch: per start the reference's draws in the reference's order (variance, lengthscale, noise, then the inducing
    inputs uniformly in the data's bounding box), ``iter_initial`` Adam steps on the device, the final loss of every start,
    and the reference's pick (the last start; the lowest loss with ``pick_best``).  Returns (variance, lengthscales, noise, Z)
    of the selected start, or None if the model does not qualify for the device batch."""
    from .engine import SparseBatch

    cfg = _trainer_config(model)
    m, d = model.inducing_variable.Z.shape
    if (cfg is None or not model.device_batch or m > SparseBatch.MAX_INDUCING or model.y.shape[1] != 1 or n_starts < 1
            or len({p.trainable for p in model.parameters}) != 1):
        return None
    _, transform, prior, floor, n_ls = cfg
    x = model.data[0]
    mins, maxs = x.min(axis=0), x.max(axis=0)
    u0 = []
    for r in range(n_starts):
        if starts is not None:
            var0, ls0, noise0 = starts[r]
        else:
            var0, ls0, noise0 = 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-3, 0)
        model.kernel.variance.assign(var0)
        model.kernel.lengthscales.assign(np.full(n_ls, ls0) if n_ls > 1 else ls0)
        model.likelihood.variance.assign(noise0)
        z = rng.uniform(mins, maxs, size=(m, d))
        u0.append(np.concatenate([model.kernel.variance.unconstrained, model.likelihood.variance.unconstrained,
                                  model.kernel.lengthscales.unconstrained, z.ravel()]))
    model.inducing_variable.trainable = bool(train_z)
    n = x.shape[0]
    key = (threading.get_ident(), model.kernel.name, n, d, m, n_starts, model.device)
    if key not in _BATCH_POOL:
        for old in [k for k in _BATCH_POOL if k[0] == key[0]]:
            _BATCH_POOL.pop(old).close()
        _BATCH_POOL[key] = SparseBatch(model.kernel.name, n, d, m, n_starts, device=model.device)
    batch = _BATCH_POOL[key]
    batch._owners = None
    batch.set_data(x, np.repeat(model.y, n_starts, axis=1))
    hyp = all(p.trainable for p in model.parameters)
    u, _, iters = batch.adam(np.stack(u0), n_ls, hyp, bool(train_z), int(iter_initial), 0.001, JITTER, transform, prior is not None, floor)
    model.n_evals += int(iters.sum())

    def assign(row):
        model.kernel.variance.unconstrained = row[0:1].copy()
        model.likelihood.variance.unconstrained = row[1:2].copy()
        model.kernel.lengthscales.unconstrained = row[2:2 + n_ls].copy()
        model.inducing_variable.Z = row[2 + n_ls:].reshape(m, d).copy()

    pick = n_starts - 1
    if pick_best:  # final loss of every start (one batched evaluation; the log prior is host logic), first minimum
        thetas, zs, lps = [], [], []
        for r in range(n_starts):
            assign(u[r])
            thetas.append(model.theta())
            zs.append(np.asarray(model.inducing_variable.Z))
            lps.append(model._log_prior())
        elbo, _, _, info = batch.elbo_grad(np.stack(thetas), np.stack(zs), JITTER)
        final = [np.inf if info[r] else -(elbo[r] + lps[r]) for r in range(n_starts)]
        model.n_evals += n_starts
        pick = min(range(n_starts), key=lambda r: (final[r], r))
    assign(u[pick])
    return (model.kernel.variance.numpy(), model.kernel.lengthscales.numpy(), model.likelihood.variance.numpy(),
            np.array(model.inducing_variable.Z))


def fit_lockstep(models, method: str, max_iter: int = 100, device_trainer: bool = True) -> bool:
    """Train all per-column models together with the first-order recipes.  Returns False (nothing done) when a recipe cannot
    run this way.  ``device_trainer`` (default): sparse models that qualify (``_device_batch``) are evaluated in one batched pass
    and their update steps run on the device -- ``"adam"``, ``"two-stage"``, ``"adadelta"``, and the Adam stage of
    ``"three-stage"`` (its two L-BFGS stages then run model by model).  Otherwise (``"adam"`` / ``"two-stage"`` only) every round
    enqueues one evaluation per model and the update rule runs on the host (``adam_lockstep``)."""
    from .gpr import _optimize_bfgs, _set_stage

    if method not in ("adam", "two-stage", "adadelta", "three-stage") or not models:
        return False
    m0 = models[0]
    n, d = m0.x.shape
    batch = _device_batch(models) if device_trainer else None
    if method in ("adadelta", "three-stage"):
        if batch is None:
            return False
        if method == "adadelta":
            return adam_device(models, batch, max_iter, rule="adadelta")
        for mdl in models:  # gpr.py:130-144: Adam on Z (all models together), then L-BFGS on the hyperparameters, then on everything
            _set_stage(mdl, hypers=False, z=True)
        adam_device(models, batch, max_iter)
        for mdl in models:
            _set_stage(mdl, hypers=True, z=False)
            _optimize_bfgs(mdl, max_iter)
            _set_stage(mdl, hypers=True, z=True)
            _optimize_bfgs(mdl, max_iter)
        return True
    if batch is not None:
        if method == "adam":
            if adam_device(models, batch, max_iter):
                return True
        else:
            for mdl in models:
                _set_stage(mdl, hypers=False, z=True)
            adam_device(models, batch, max_iter)
            for mdl in models:
                _set_stage(mdl, hypers=True, z=False)
            adam_device(models, batch, max_iter)
            for mdl in models:
                _set_stage(mdl, hypers=True, z=True)
            return True
    if isinstance(m0, SparseModel):
        pool = _model_pool(m0.kernel.name, n, d, m0.inducing_variable.Z.shape[0], m0.y.shape[1], m0.device, len(models))
    else:  # per-column exact models: one exact-GP handle each
        pool = _exact_pool(m0.kernel.name, n, d, m0.y.shape[1], m0.device, len(models))
    for mdl, gp in zip(models, pool):
        mdl.bind(gp)
    try:
        if method == "adam":
            adam_lockstep(models, max_iter)
            return True
        for mdl in models:
            _set_stage(mdl, hypers=False, z=True)
        adam_lockstep(models, max_iter)
        for mdl in models:
            _set_stage(mdl, hypers=True, z=False)
        adam_lockstep(models, max_iter)
        for mdl in models:
            _set_stage(mdl, hypers=True, z=True)
        return True
    finally:
        for mdl in models:
            mdl._own = None
        if not isinstance(m0, SparseModel) and n > 2048:
            # large exact handles (3 n^2 doubles each) are not worth keeping around: give them back
            for key in [k for k in _MODEL_POOLS if k[0] == threading.get_ident() and k[1] == "exact" and k[3] == n]:
                for gp in _MODEL_POOLS.pop(key):
                    gp.close()
