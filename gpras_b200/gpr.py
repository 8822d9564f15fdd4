"""B200-native drop-in for ``gpras.gpr`` -- the GPR surrogate of fema-ffrd/gpras.

Mirrors the reference module's public surface (``gpras/gpr.py:21-41,206-384``): ``KERNEL_FACTORY``,
``KernelType``, ``OptimizerType``, ``InductionInitializerType``, ``OPTIMIZERS`` and the ``GPRAS`` class with
``fit / predict / to_file / from_file / models``.  The numerics that the reference delegates to
GPflow/TensorFlow run in hand-written sm_100a kernels behind ``libgpras_b200.so``; the host side stays
Python (SciPy L-BFGS-B / differential evolution drive the GPU objective exactly as
``gpflow.optimizers.Scipy`` does at ``gpr.py:197-203``).  There is no CPU fallback.

Two model families sit behind the same API:

* ``n_inducing=None`` (or ``exact=True``): the exact GP named by BASELINE.json's north_star, i.e. the
  ``Z == X`` limit of the reference's SGPR.  ``shared_kernel=True`` shares one hyperparameter set across all
  target columns (one factorisation, multi-RHS solves); the default keeps the reference's one model per
  column (``gpr.py:293-308``).
* ``n_inducing=M``: the reference's sparse model (see ``sparse.py``).

Parameterisation, initial values and priors follow the reference: variance 1, scalar lengthscale
``mean(|x|)`` (``gpr.py:289,298``; ``ard=True`` is an extension), likelihood variance 1 (GPflow default),
softplus transforms with a 1e-6 floor on the likelihood variance, LogNormal(0, 1) priors on the three
constrained hyperparameters (``gpr.py:303-305``) counted only while the parameter is trainable.
"""

from __future__ import annotations

import collections
import pickle
import threading
from pathlib import Path
from typing import Any, Literal

import numpy as np
from numpy.typing import NDArray

from .engine import ExactGP

LOG_2PI = float(np.log(2.0 * np.pi))
NOISE_FLOOR = 1e-6  # GPflow Gaussian likelihood DEFAULT_VARIANCE_LOWER_BOUND


class KernelSpec:
    """Stand-in for the GPflow kernel classes held by the reference's ``KERNEL_FACTORY``."""

    def __init__(self, name: str, supported: bool = True):
        self.name, self.supported = name, supported

    def __repr__(self) -> str:
        return f"KernelSpec({self.name!r})"


KERNEL_FACTORY = {
    "Matern12": KernelSpec("Matern12"),
    "Matern32": KernelSpec("Matern32"),
    "Matern52": KernelSpec("Matern52"),
    "RBF": KernelSpec("RBF"),
    # listed by the reference but not constructible there either (``lengthscales=`` is rejected, gpr.py:26-28,298)
    "Linear": KernelSpec("Linear", supported=False),
    "Polynomial": KernelSpec("Polynomial", supported=False),
    "Periodic": KernelSpec("Periodic", supported=False),
    "Exponential": KernelSpec("Exponential"),
}

KernelType = Literal["Matern12", "Matern32", "Matern52", "RBF", "Linear", "Polynomial", "Periodic", "Exponential"]
OptimizerType = Literal["two-stage", "adam", "L-BFGS-B", "stochastic", "diffential_evolution"]
InductionInitializerType = Literal["kmeans", "grid"]


# ------------------------------------------------------------------------------------------------
# parameters (GPflow ``Parameter`` look-alikes: ``.numpy()``, ``.assign()``, ``.prior``, trainable flag)
# ------------------------------------------------------------------------------------------------
def _softplus(u):
    return np.logaddexp(0.0, u)


def _softplus_inv(v):
    return v + np.log(-np.expm1(-v))


def _sigmoid(u):
    return 0.5 * (1.0 + np.tanh(0.5 * u))


class Parameter:
    """Positive parameter stored unconstrained.  ``transform="softplus"`` (GPflow ``positive()``, the reference's
    parameterisation): value = softplus(u) + lower; ``transform="log"`` (scikit-learn's): value = exp(u) + lower."""

    def __init__(self, value, lower: float = 0.0, prior: str | None = None, trainable: bool = True, transform: str = "softplus"):
        if transform not in ("softplus", "log"):
            raise ValueError(f"unknown transform {transform!r}")
        self.lower = float(lower)
        self.prior = prior  # "LogNormal(0,1)" or None
        self.trainable = trainable
        self.transform = transform
        self.unconstrained = np.atleast_1d(self._inverse(np.asarray(value, np.float64) - self.lower)).astype(np.float64)
        self._scalar = np.ndim(value) == 0

    def _forward(self, u):
        return np.exp(u) if self.transform == "log" else _softplus(u)

    def _inverse(self, v):
        return np.log(v) if self.transform == "log" else _softplus_inv(v)

    def value(self):
        """Constrained value as an array (always 1-D)."""
        return self._forward(self.unconstrained) + self.lower

    def numpy(self):
        v = self.value()
        return float(v[0]) if self._scalar else v

    def assign(self, value) -> None:
        value = np.asarray(value, np.float64)
        self._scalar = value.ndim == 0
        self.unconstrained = np.atleast_1d(self._inverse(value - self.lower)).astype(np.float64)

    def set_transform(self, transform: str) -> None:
        """Switch the parameterisation, keeping the constrained value."""
        v = self.value()
        self.transform = transform
        self.unconstrained = np.atleast_1d(self._inverse(v - self.lower)).astype(np.float64)

    @property
    def size(self) -> int:
        return self.unconstrained.size

    def dvalue_du(self):
        return np.exp(self.unconstrained) if self.transform == "log" else _sigmoid(self.unconstrained)

    def log_prior(self) -> float:
        if self.prior is None:
            return 0.0
        with np.errstate(all="ignore"):  # a line-search trial point may underflow a value to 0: the loss is then non-finite, handled by the caller
            lv = np.log(self.value())
            return float(np.sum(-lv - 0.5 * LOG_2PI - 0.5 * lv * lv))

    def dlog_prior_dvalue(self):
        if self.prior is None:
            return np.zeros_like(self.unconstrained)
        v = self.value()
        with np.errstate(all="ignore"):
            return -(1.0 + np.log(v)) / v

    def __repr__(self) -> str:
        return f"Parameter({self.numpy()!r}, trainable={self.trainable})"


class _Kernel:
    def __init__(self, name: str, variance, lengthscales, transform: str = "softplus"):
        self.name = name
        self.variance = Parameter(variance, prior="LogNormal(0,1)", transform=transform)
        self.lengthscales = Parameter(lengthscales, prior="LogNormal(0,1)", transform=transform)


class _Likelihood:
    def __init__(self, variance=1.0, transform: str = "softplus"):
        # the 1e-6 floor belongs to GPflow's Gaussian likelihood; scikit-learn's log-space WhiteKernel has none
        self.variance = Parameter(variance, lower=NOISE_FLOOR if transform == "softplus" else 0.0, prior="LogNormal(0,1)",
                                  transform=transform)


class _Inducing:
    """``model.inducing_variable.Z`` as read by ``production/analysis/pipeline.py:115``."""

    def __init__(self, z: NDArray[Any], trainable: bool = True):
        self.Z = np.asarray(z, np.float64)
        self.trainable = trainable


# Pools of small-problem handles outlive a GPRAS instance (creating a handle costs ~1 ms and a pool of 16 serves the batched
# restarts), but not without bound: the cache keeps the most recently used shapes and closes the handles of the rest.
_POOL_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_POOL_CACHE_MAX_SHAPES = 4


def release_pools() -> None:
    """Close every pooled device handle of this process (device memory goes back to the driver)."""
    while _POOL_CACHE:
        _, ent = _POOL_CACHE.popitem(last=False)
        for g in ent["gps"]:
            g.close()
    from .sparse import release_batches

    release_batches()


class _DeviceSlot:
    """Device handles of one GPRAS instance: one per calling thread and problem shape.  Per-column models take turns
    on their thread's handle; with ``fit(n_jobs > 1)`` several threads keep several evaluations in flight on one GPU
    (a handle is not thread-safe, ``include/gpras_b200.h``)."""

    def __init__(self):
        self._per_thread: dict = {}

    def acquire(self, model: "ExactModel") -> ExactGP:
        st = self._per_thread.setdefault(threading.get_ident(), {"gp": None, "key": None, "owner": None})
        key = (model.kernel.name, model.x.shape[0], model.x.shape[1], model.y.shape[1], model.device)
        if st["gp"] is None or st["key"] != key:
            if st["gp"] is not None:
                st["gp"].close()
            st["gp"] = ExactGP(model.kernel.name, key[1], key[2], key[3], device=model.device)
            st["key"], st["owner"] = key, None
        if st["owner"] is not model:
            st["gp"].set_data(model.x, model.y)
            st["owner"] = model
        return st["gp"]

    def pool(self, model: "ExactModel", count: int) -> list:
        """``count`` handles bound to the model's data (the calling thread's own handle first): independent evaluations
        enqueued on them run concurrently on the GPU, which is what fills the machine at small N (measured at N = 256:
        3.7e3 evals/s with one evaluation in flight, 7.3e4 with 32)."""
        first = self.acquire(model)
        st = self._per_thread[threading.get_ident()]
        small = model.x.shape[0] <= 2048
        # creating a handle costs ~10 ms (pinned + device allocations), so pools for small problems outlive this GPRAS
        # instance in a process-wide cache (<= 100 MB per handle); large ones belong to the slot and are closed with it
        cache = _POOL_CACHE if small else st.setdefault("own_pools", {})
        ckey = (threading.get_ident(),) + st["key"]
        ent = cache.setdefault(ckey, {"gps": [], "owner": None})
        if small:
            _POOL_CACHE.move_to_end(ckey)
            while len(_POOL_CACHE) > _POOL_CACHE_MAX_SHAPES:
                _, old = _POOL_CACHE.popitem(last=False)
                for g in old["gps"]:
                    g.close()
        extra = ent["gps"]
        if not small:
            st["extra"] = extra
        while len(extra) < count - 1:
            extra.append(ExactGP(model.kernel.name, model.x.shape[0], model.x.shape[1], model.y.shape[1], device=model.device))
            ent["owner"] = None
        if ent["owner"] is not model:  # models take turns on the pool: rebinding data is a ~20 us copy per handle
            for g in extra:
                g.set_data(model.x, model.y)
            ent["owner"] = model
        return [first] + extra[: count - 1]

    def release_other_threads(self) -> None:
        me = threading.get_ident()
        for tid in [t for t in self._per_thread if t != me]:
            st = self._per_thread.pop(tid)
            if st["gp"] is not None:
                st["gp"].close()
            for g in st.get("extra", []):
                g.close()


class _FixedSlot:
    """Device slot of a restart lane: always the same handle, already bound to the model's data."""

    def __init__(self, gp: ExactGP):
        self.gp = gp

    def acquire(self, model: "ExactModel") -> ExactGP:
        return self.gp


class ExactModel:
    """Exact GP with one hyperparameter set for all columns of ``y`` (the ``Z == X`` limit of the
    reference's per-column ``SGPR``).  Exposes the attribute names the reference's recipes touch:
    ``kernel.variance / kernel.lengthscales / likelihood.variance / inducing_variable.Z / data /
    trainable_variables / training_loss()`` (``gpr.py:57-62,80-81,88-91,155``)."""

    def __init__(self, kernel_name, x, y, lengthscales, slot: _DeviceSlot, device: int = 0, priors: bool = True,
                 parameterisation: str = "softplus"):
        self.x, self.y = x, y
        self.data = (x, y)
        self.device = device
        self.kernel = _Kernel(kernel_name, 1.0, lengthscales, parameterisation)
        self.likelihood = _Likelihood(1.0, parameterisation)
        self.inducing_variable = _Inducing(x, trainable=False)
        if not priors:
            self.kernel.variance.prior = self.kernel.lengthscales.prior = self.likelihood.variance.prior = None
        self._slot = slot
        self.n_evals = 0

    def clone(self, slot=None) -> "ExactModel":
        """A second model object on the same data with its own parameters, on ``slot`` (default: this model's device slot,
        which hands every calling thread its own device handle)."""
        ls = self.kernel.lengthscales.numpy()
        m = ExactModel(self.kernel.name, self.x, self.y, np.array(ls) if np.ndim(ls) else float(ls), slot or self._slot, self.device,
                       True, self.kernel.variance.transform)
        for src, dst in zip(self.parameters, m.parameters):
            dst.prior, dst.trainable, dst.lower = src.prior, src.trainable, src.lower
            dst.unconstrained = src.unconstrained.copy()
        return m

    def lane_models(self, count: int) -> list:
        """``count`` clones for concurrent restart lanes, each on a device handle of its own taken from the slot's pool: the
        handles are created once (in the calling thread) and kept for the next ``fit`` -- creating and freeing a 1.6 GB
        workspace per lane and call costs ~0.3 s each next to other live allocations and stalls the evaluations in flight."""
        if not hasattr(self._slot, "pool"):  # e.g. the oracle-backed test double: evaluations are pure functions
            return [self.clone() for _ in range(count)]
        gps = self._slot.pool(self, count)
        for gp in gps:  # one host thread per lane: let them sleep while they wait
            gp.set_blocking_wait(count > 1)
        return [self.clone(_FixedSlot(gp)) for gp in gps]

    def release_other_threads(self) -> None:
        release = getattr(self._slot, "release_other_threads", None)
        if release is not None:
            release()

    # -- parameter plumbing --
    @property
    def parameters(self):
        return [self.kernel.variance, self.likelihood.variance, self.kernel.lengthscales]

    @property
    def trainable_variables(self):
        return [p.unconstrained for p in self.parameters if p.trainable]

    def set_trainable(self, flag: bool, hypers: bool = True) -> None:
        if hypers:
            for p in self.parameters:
                p.trainable = flag

    def get_u(self) -> NDArray[Any]:
        ps = [p for p in self.parameters if p.trainable]
        return np.concatenate([p.unconstrained for p in ps]) if ps else np.zeros(0)

    def set_u(self, u) -> None:
        u = np.asarray(u, np.float64)
        o = 0
        for p in self.parameters:
            if p.trainable:
                p.unconstrained = u[o : o + p.size].copy()
                o += p.size

    def theta(self) -> NDArray[Any]:
        d = self.x.shape[1]
        ls = np.atleast_1d(self.kernel.lengthscales.numpy())
        if ls.size == 1:
            ls = np.full(d, ls[0])
        return np.concatenate([[self.kernel.variance.numpy(), self.likelihood.variance.numpy()], ls])

    # -- objective --
    def _log_prior(self) -> float:
        return sum(p.log_prior() for p in self.parameters if p.trainable)

    def training_loss(self) -> float:
        """-(LML + log prior over trainable hyperparameters), loss only."""
        gp = self._slot.acquire(self)
        lml, _ = gp.lml_grad(self.theta(), want_grad=False)
        self.n_evals += 1
        return -(lml + self._log_prior())

    def loss_and_grad(self, u=None):
        """Loss and its gradient w.r.t. the trainable unconstrained variables (what ``GradientTape`` returns)."""
        if u is not None:
            self.set_u(u)
        gp = self._slot.acquire(self)
        lml, glog = gp.lml_grad(self.theta(), want_grad=True)
        return self._assemble(lml, glog)

    # -- asynchronous form on a handle of the model's own (lock-stepped training of per-column models) --
    def bind(self, gp: ExactGP) -> None:
        gp.set_data(self.x, self.y)
        self._own = gp

    def enqueue_loss_and_grad(self, u) -> None:
        self.set_u(u)
        self._own.enqueue(self.theta(), True)

    def fetch_loss_and_grad(self):
        return self._assemble(*self._own.fetch())

    def _assemble(self, lml, glog):
        self.n_evals += 1
        g_var, g_noise, g_ls = glog[0], glog[1], glog[2:]
        if self.kernel.lengthscales.size == 1:
            g_ls = np.array([g_ls.sum()])
        parts = []
        for p, gl in ((self.kernel.variance, np.array([g_var])), (self.likelihood.variance, np.array([g_noise])),
                      (self.kernel.lengthscales, g_ls)):
            if p.trainable:
                v = p.value()
                dv = gl / v + p.dlog_prior_dvalue()
                parts.append(-(dv * p.dvalue_du()))
        grad = np.concatenate(parts) if parts else np.zeros(0)
        return -(lml + self._log_prior()), grad

    def loss_and_grad_many(self, us, want_grad: bool = True):
        """``loss_and_grad`` at several unconstrained vectors with the device evaluations in flight TOGETHER (independent
        restarts / candidates are the batch axis of the path).  Returns (losses, grads); the
        model is left at the last vector.  Each value is bitwise what the one-at-a-time call returns."""
        us = [np.asarray(u, np.float64) for u in us]
        if not hasattr(self._slot, "pool"):  # e.g. the oracle-backed test double: one at a time
            outs = [self.loss_and_grad(u) if want_grad else (self._loss_at(u), None) for u in us]
            return [o[0] for o in outs], [o[1] for o in outs]
        n = self.x.shape[0]
        width = 16 if n <= 2048 else (4 if n <= 4096 else 2)
        gps = self._slot.pool(self, min(width, len(us)))
        losses, grads = [None] * len(us), [None] * len(us)
        for lo in range(0, len(us), len(gps)):
            chunk = list(range(lo, min(lo + len(gps), len(us))))
            host = []
            for k, i in enumerate(chunk):
                self.set_u(us[i])
                gps[k].enqueue(self.theta(), want_grad)
                host.append((self._log_prior(), [(p, p.value(), p.dlog_prior_dvalue(), p.dvalue_du())
                                                 for p in self.parameters]))
            for k, i in enumerate(chunk):
                try:
                    lml, glog = gps[k].fetch()
                except Exception:
                    for k2 in range(k + 1, len(chunk)):  # leave no evaluation pending on the other handles
                        try:
                            gps[k2].fetch()
                        except Exception:
                            pass
                    raise
                self.n_evals += 1
                log_prior, per_param = host[k]
                losses[i] = -(lml + log_prior)
                if want_grad:
                    g_ls = glog[2:] if self.kernel.lengthscales.size > 1 else np.array([glog[2:].sum()])
                    parts = []
                    for (p, v, dlp, dvdu), gl in zip(per_param, (np.array([glog[0]]), np.array([glog[1]]), g_ls)):
                        if p.trainable:
                            parts.append(-((gl / v + dlp) * dvdu))
                    grads[i] = np.concatenate(parts) if parts else np.zeros(0)
        return losses, grads

    def _loss_at(self, u) -> float:
        self.set_u(u)
        return self.training_loss()

    # -- prediction --
    def predict_y(self, xs, keep_handle: bool = False):
        """``predict_y`` of the reference's models (``gpr.py:337``).  ``keep_handle``: give this model a device handle of its
        own for prediction, so that its factor stays resident while the other per-column models are predicted and across
        calls (the reference's second caller predicts once per plan with unchanged models, ``gpras/preprocess.py:601-606``)."""
        if keep_handle:
            if getattr(self, "_pred_gp", None) is None:
                self._pred_gp = ExactGP(self.kernel.name, self.x.shape[0], self.x.shape[1], self.y.shape[1], device=self.device)
                self._pred_gp.set_data(self.x, self.y)
            gp = self._pred_gp
        else:
            gp = self._slot.acquire(self)
        gp.condition(self.theta())  # free when the handle already holds this model's factor at this theta
        return gp.predict(np.asarray(xs, np.float64))

    def release(self) -> None:
        gp = getattr(self, "_pred_gp", None)
        if gp is not None:
            gp.close()
            self._pred_gp = None

    def parameter_dict(self) -> dict:
        """Plain-ndarray version of ``gpflow.utilities.parameter_dict`` (same keys)."""
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


# ------------------------------------------------------------------------------------------------
# optimiser recipes (gpras/gpr.py:44-214) acting on any model exposing get_u / set_u / loss_and_grad
# ------------------------------------------------------------------------------------------------
def _has_z(model) -> bool:
    return getattr(model, "supports_z_training", False)


def _set_stage(model, hypers: bool, z: bool) -> None:
    """``gpflow.set_trainable`` choreography of the recipes."""
    model.set_trainable(hypers, hypers=True)
    if _has_z(model):
        model.inducing_variable.trainable = z


def _optimize_adam(model, max_iter: int, learning_rate: float = 0.001) -> None:
    """Keras Adam (lr 1e-3, beta 0.9 / 0.999, eps 1e-7) with the reference's early-stopping rule
    (relative improvement <= 10e-6 for more than 50 steps, ``gpr.py:159-173``)."""
    u = model.get_u()
    if u.size == 0:
        return
    m = np.zeros_like(u)
    v = np.zeros_like(u)
    b1, b2, eps = 0.9, 0.999, 1e-7
    best, count, tol, patience = np.inf, 0, 10e-6, 50
    for t in range(1, int(max_iter) + 1):
        loss, g = model.loss_and_grad(u)
        m = b1 * m + (1.0 - b1) * g
        v = b2 * v + (1.0 - b2) * g * g
        alpha = learning_rate * np.sqrt(1.0 - b2**t) / (1.0 - b1**t)
        u = u - alpha * m / (np.sqrt(v) + eps)
        model.set_u(u)
        if ((best - loss) / abs(loss)) > tol:
            best, count = loss, 0
        else:
            count += 1
            if count > patience:
                break


def _optimize_adadelta(model, max_iter: int, learning_rate: float = 0.001) -> float:
    """Keras Adadelta (lr 1e-3, rho 0.95, eps 1e-7), fixed ``max_iter`` steps (``gpr.py:176-192``)."""
    u = model.get_u()
    acc_g = np.zeros_like(u)
    acc_d = np.zeros_like(u)
    rho, eps = 0.95, 1e-7
    loss = float("nan")
    for _ in range(int(max_iter)):
        loss, g = model.loss_and_grad(u)
        acc_g = rho * acc_g + (1.0 - rho) * g * g
        upd = g * np.sqrt(acc_d + eps) / np.sqrt(acc_g + eps)
        acc_d = rho * acc_d + (1.0 - rho) * upd * upd
        u = u - learning_rate * upd
        model.set_u(u)
    return loss


def _optimize_bfgs(model, max_iter: int, *, bounds=None, **scipy_options: Any) -> Any:
    """SciPy L-BFGS-B over the trainable unconstrained variables, as ``gpflow.optimizers.Scipy`` (``gpr.py:195-203``):
    ``maxiter`` is the only option the reference sets, so SciPy's defaults (``ftol`` 2.2e-9, ``gtol`` 1e-5) decide when it
    stops.  Extras: ``bounds`` (on the unconstrained variables, e.g. scikit-learn's log-space box) and any further SciPy
    option (``ftol``, ``gtol``, ``maxfun`` ...).  A line-search trial point whose covariance matrix is not positive
    definite makes the search back off instead of aborting the fit (the reference would die in TensorFlow's Cholesky)."""
    from scipy.optimize import minimize

    from ._lib import NotPositiveDefiniteError

    u0 = model.get_u()
    if u0.size == 0:
        return None

    good = {}

    def fun(u):
        try:
            f, g = model.loss_and_grad(u) if np.all(np.isfinite(u)) else (np.nan, None)
            bad = not (np.isfinite(f) and np.all(np.isfinite(g)))
        except np.linalg.LinAlgError:  # NotPositiveDefiniteError is one
            bad = True
        if bad:
            if not good:
                raise np.linalg.LinAlgError("the start itself has no finite objective (covariance matrix not positive definite)")
            # A line-search trial point stepped outside the region where the covariance matrix factorises.  Answer with the
            # mirror image of the last good point's descent (value above it by half the predicted decrease, slope reversed):
            # the line search's quadratic interpolation then retries at a third of the step instead of aborting the fit.
            d = np.nan_to_num(u - good["u"], nan=0.0, posinf=1e6, neginf=-1e6)
            slope = float(good["g"] @ d)
            nd = float(d @ d)
            if nd == 0.0 or not slope < 0.0:
                return good["f"] + 1.0, np.zeros_like(good["g"])
            return good["f"] - 0.5 * slope, (-slope / nd) * d
        if not good or f <= good["f"]:
            good.update(u=np.array(u, np.float64), f=float(f), g=np.array(g, np.float64))
        return f, g

    res = minimize(fun, u0, jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": int(max_iter), **scipy_options})
    model.set_u(res.x)
    return res


def _optimize_two_stage(model, max_iter: int = 100) -> float:
    """Adam on the inducing inputs, then Adam on the hyperparameters (``gpr.py:112-127``)."""
    _set_stage(model, hypers=False, z=True)
    _optimize_adam(model, max_iter)
    _set_stage(model, hypers=True, z=False)
    _optimize_adam(model, max_iter)
    _set_stage(model, hypers=True, z=True)
    return model.training_loss()


def _optimize_three_stage(model, max_iter: int = 100) -> None:
    """Adam on Z, L-BFGS on hyperparameters, L-BFGS on everything (``gpr.py:130-144``)."""
    _set_stage(model, hypers=False, z=True)
    _optimize_adam(model, max_iter)
    _set_stage(model, hypers=True, z=False)
    _optimize_bfgs(model, max_iter)
    _set_stage(model, hypers=True, z=True)
    _optimize_bfgs(model, max_iter)


def _optimize_multi_start(model, n_starts: int = 40, iter_initial: int = 20, iter_final: int = 1000, seed=None,
                          starts=None, pick_best: bool = False, lockstep: bool | None = None, train_z: bool = False) -> None:
    """Random restarts with a short Adam run each, then L-BFGS from the selected start (``gpr.py:73-109``).

    Faithful to the reference by default: its ``best_loss`` is never assigned (``gpr.py:86,96``), so the LAST
    start is the one that gets polished; ``pick_best=True`` selects the lowest coarse loss instead.  The
    reference's generator is unseeded (``gpr.py:76-77``); ``seed`` / ``starts`` ((R, 3) constrained
    [variance, lengthscale, noise]) make runs reproducible.  ``lockstep`` (default on) advances all starts together: exact
    models evaluate the starts' objectives in flight together (identical result); sparse models with at most 128 inducing
    points put the starts into one device batch whose coarse Adam stage runs device resident (``gpras_sgpr_batch_adam``; same
    draws, same trajectories up to the rounding of the device's transcendental functions).
    The reference overwrites ``model.inducing_variable.Z`` with a raw ndarray (``gpr.py:91,108``), which replaces the GPflow
    ``Parameter``: from the first redraw on Z is no longer a trainable variable, so Adam and the final L-BFGS move only the
    three hyperparameters.  That is the default here too; ``train_z=True`` keeps the redrawn Z trainable instead.
    """
    rng = np.random.default_rng(seed)
    x = model.data[0]
    mins, maxs = x.min(axis=0), x.max(axis=0)
    best_loss, best = None, None
    if lockstep is None:
        lockstep = (hasattr(model, "loss_and_grad_many") and not _has_z(model)) or _has_z(model)
    if lockstep and _has_z(model):
        # sparse model: every start becomes one model of a device batch (same draws in the same order), the coarse Adam stage
        # runs device resident for all starts at once; None when the model does not qualify (then the loop below runs)
        from .sparse import multi_start_device

        best = multi_start_device(model, rng, int(n_starts), int(iter_initial), starts, pick_best, bool(train_z))
        lockstep = best is not None
    elif lockstep:
        best = _multi_start_lockstep(model, rng, int(n_starts), int(iter_initial), starts, pick_best)
    for r in range(0 if lockstep else int(n_starts)):
        if starts is not None:
            var0, ls0, noise0 = starts[r]
        else:
            var0, ls0, noise0 = 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-3, 0)
        model.kernel.variance.assign(var0)
        model.kernel.lengthscales.assign(np.full(model.kernel.lengthscales.size, ls0) if model.kernel.lengthscales.size > 1 else ls0)
        model.likelihood.variance.assign(noise0)
        if _has_z(model):
            model.inducing_variable.Z = rng.uniform(mins, maxs, size=model.inducing_variable.Z.shape)
            model.inducing_variable.trainable = bool(train_z)
        _optimize_adam(model, iter_initial)
        loss = model.training_loss()
        if not pick_best or best_loss is None or loss < best_loss:
            best = (model.kernel.variance.numpy(), model.kernel.lengthscales.numpy(), model.likelihood.variance.numpy(),
                    np.array(model.inducing_variable.Z))
            best_loss = loss
    model.kernel.variance.assign(best[0])
    model.kernel.lengthscales.assign(best[1])
    model.likelihood.variance.assign(best[2])
    if _has_z(model):
        model.inducing_variable.Z = best[3]
    _optimize_bfgs(model, iter_final)


def _multi_start_lockstep(model, rng, n_starts: int, iter_initial: int, starts, pick_best: bool):
    """The coarse stage of ``_optimize_multi_start`` with all starts advancing one Adam step per round, their objective
    evaluations in flight together on the GPU.  Starts are independent (no inducing inputs to draw, Adam consumes no random
    numbers), so every start follows exactly the trajectory of the one-after-the-other loop; only the order of the device
    work changes.  Returns the selected (variance, lengthscales, noise, Z)."""
    nls = model.kernel.lengthscales.size
    u0 = []
    for r in range(n_starts):
        if starts is not None:
            var0, ls0, noise0 = starts[r]
        else:
            var0, ls0, noise0 = 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-1, 1), 10 ** rng.uniform(-3, 0)
        model.kernel.variance.assign(var0)
        model.kernel.lengthscales.assign(np.full(nls, ls0) if nls > 1 else ls0)
        model.likelihood.variance.assign(noise0)
        u0.append(model.get_u())
    u = [v.copy() for v in u0]
    if u and u[0].size:
        m = [np.zeros_like(v) for v in u]
        vv = [np.zeros_like(v) for v in u]
        best, count = [np.inf] * n_starts, [0] * n_starts
        active = list(range(n_starts))
        b1, b2, eps, lr, tol, patience = 0.9, 0.999, 1e-7, 0.001, 10e-6, 50
        for t in range(1, iter_initial + 1):
            if not active:
                break
            losses, grads = model.loss_and_grad_many([u[r] for r in active])
            still = []
            for loss, g, r in zip(losses, grads, active):
                m[r] = b1 * m[r] + (1.0 - b1) * g
                vv[r] = b2 * vv[r] + (1.0 - b2) * g * g
                alpha = lr * np.sqrt(1.0 - b2**t) / (1.0 - b1**t)
                u[r] = u[r] - alpha * m[r] / (np.sqrt(vv[r]) + eps)
                if ((best[r] - loss) / abs(loss)) > tol:
                    best[r], count[r] = loss, 0
                    still.append(r)
                else:
                    count[r] += 1
                    if count[r] <= patience:
                        still.append(r)
            active = still
    final, _ = model.loss_and_grad_many(u, want_grad=False)
    pick = n_starts - 1
    if pick_best:  # first minimum, like the strict `<` of the sequential loop
        pick = min(range(n_starts), key=lambda r: (final[r], r))
    model.set_u(u[pick])
    return (model.kernel.variance.numpy(), model.kernel.lengthscales.numpy(), model.likelihood.variance.numpy(),
            np.array(model.inducing_variable.Z))


def _optimize_differential_evolutions(model, popsize: int = 15, max_iter: int = 500, seed=None, verbose: bool = False,
                                      **de_options: Any) -> None:
    """Adam on Z (3000 its), then SciPy differential evolution over (log10 variance, log10 lengthscale,
    log10 noise) in [-1, 1]^2 x [-3, 0] with the loss as objective (``gpr.py:44-70``).  The reference prints
    every evaluation (``gpr.py:61``); pass ``verbose=True`` for that.  Further SciPy options (``polish``, ``tol`` ...) pass
    through; the defaults are the reference's (it sets none)."""
    from scipy.optimize import differential_evolution

    _set_stage(model, hypers=False, z=True)
    _optimize_adam(model, max_iter=3000)
    bounds = [(-1, 1), (-1, 1), (-3, 0)]
    nls = model.kernel.lengthscales.size

    def assign(p):
        model.kernel.variance.assign(10 ** p[0])
        model.kernel.lengthscales.assign(np.full(nls, 10 ** p[1]) if nls > 1 else 10 ** p[1])
        model.likelihood.variance.assign(10 ** p[2])

    def objective(p):
        assign(p)
        loss = model.training_loss()
        if verbose:
            print(loss)
        return loss

    res = differential_evolution(objective, bounds, popsize=popsize, maxiter=max_iter, seed=seed, **de_options)
    assign(res.x)


OPTIMIZERS: dict[str, Any] = {
    "two-stage": _optimize_two_stage,
    "three-stage": _optimize_three_stage,
    "adam": _optimize_adam,
    "adadelta": _optimize_adadelta,
    "L-BFGS-B": _optimize_bfgs,
    "stochastic": _optimize_multi_start,
    "diffential_evolution": _optimize_differential_evolutions,
}


# ------------------------------------------------------------------------------------------------
# GPRAS
# ------------------------------------------------------------------------------------------------
class GPRAS:
    """Gaussian Process Regression for HEC-RAS model upskilling and emulation (``gpras/gpr.py:217``)."""

    def __init__(self, kernel: KernelType) -> None:
        self.kernel_str = kernel
        self.kernel = KERNEL_FACTORY[kernel]  # KeyError on an unknown name, as the reference (gpr.py:230)
        self.models: list[Any] = []
        self.x: NDArray[Any] | None = None
        self.y: NDArray[Any] | None = None
        self._slot = _DeviceSlot()
        self._opts: dict[str, Any] = {}

    # -- fit ------------------------------------------------------------------------------------
    def fit(
        self,
        x: NDArray[Any],
        y: NDArray[Any],
        n_inducing: int | None,
        inducing_initializer: InductionInitializerType = "kmeans",
        optimization_method: OptimizerType = "two-stage",
        *,
        exact: bool | None = None,
        ard: bool = False,
        shared_kernel: bool = False,
        priors: bool = True,
        parameterisation: Literal["softplus", "log"] = "softplus",
        device: int | None = None,
        initial_theta: NDArray[Any] | None = None,
        restarts: NDArray[Any] | None = None,
        restart_lanes: int | None = None,
        kmeans_on_device: bool = True,
        n_jobs: int = 1,
        lockstep_models: bool = True,
        device_trainer: bool = True,
        **opt_kwargs: Any,
    ) -> None:
        """Fit the surrogate (``gpr.py:237-275``).  Positional arguments and ``**opt_kwargs`` are the reference's.

        Keyword-only extensions (defaults reproduce the reference): ``exact`` / ``n_inducing=None`` selects the
        exact GP; ``ard`` one lengthscale per feature; ``shared_kernel`` one hyperparameter set for all columns;
        ``priors=False`` drops the LogNormal priors (scikit-learn's objective); ``parameterisation="log"`` optimises
        log(theta) like scikit-learn instead of GPflow's softplus-unconstrained variables (same optimum, different path and
        stopping point under SciPy's default tolerances); ``device`` defaults to the current CUDA device (``LOCAL_RANK``
        under ``torch.distributed``); ``initial_theta``
        ([variance, noise, lengthscale(s)]) overrides the initial values; ``restarts`` ((R, 2 + n_ls) constrained
        start points, column order [variance, noise, lengthscale(s)]) runs the recipe from every start and keeps
        the lowest final loss (handed out to ranks by a ticket counter when ``torch.distributed`` is initialised;
        ``restart_lanes`` restarts in flight per GPU, default 2 for exact models above 4096 rows, else 1);
        ``kmeans_on_device`` (default on) runs the Lloyd iterations of the "kmeans" initialiser on the GPU from scikit-learn's own
        k-means++ seeds (same iteration count, labels and inertia as ``KMeans(random_state=0, n_init="auto")``, centres to 1e-15;
        ``False`` calls scikit-learn's ``KMeans`` itself, as the reference does); ``n_jobs > 1``
        optimises that many per-column models concurrently on the GPU (host threads, one device handle each; the
        reference loops sequentially, ``gpr.py:273-274``, and so does the default).  Under ``torch.distributed`` (one
        process per GPU) per-column models are sharded round-robin over ranks and their parameters all-gathered.
        ``lockstep_models`` (default on) trains the per-column sparse models of the Adam-based recipes together, one round of
        evaluations in flight at a time; the result is the one of the sequential loop.  ``device_trainer`` (default on): sparse
        models with at most 128 inducing points are evaluated in ONE batched pass (model index in the grid) and their Adam
        steps -- transforms, priors, chain rule, update and early-stopping rule -- run on the device as a replayed CUDA graph
        (``gpras_sgpr_batch_adam``), the host reading the variables back once per stage; same trajectories as the host loop up
        to the rounding of ``exp`` / ``log`` / ``pow`` (measured 1e-13 relative on the fitted parameters).
        """
        self.x = np.asarray(x).astype(np.float64)
        self.y = np.asarray(y).astype(np.float64)
        if exact is None:
            exact = n_inducing is None
        if device is None:
            device = _default_device()
        self._opts = dict(exact=exact, ard=ard, shared_kernel=shared_kernel, priors=priors, device=device,
                          parameterisation=parameterisation, kmeans_on_device=bool(kmeans_on_device))
        self._init_models(self.x, self.y, n_inducing, inducing_initializer)
        opt = OPTIMIZERS[optimization_method]  # KeyError on an unknown method, as the reference (gpr.py:272)
        unique = self.models[:1] if (shared_kernel and exact) else self.models

        def run_one(model) -> None:
            if initial_theta is not None:
                _assign_theta(model, initial_theta)
            if restarts is None:
                opt(model, **opt_kwargs)
            else:
                from .parallel import run_restarts

                lanes = restart_lanes if restart_lanes is not None else (2 if exact and self.x.shape[0] > 4096 else 1)
                run_restarts(model, opt, np.asarray(restarts, np.float64), opt_kwargs, lanes=lanes)

        from .parallel import dist_info

        done = False
        lockstep_methods = ("adam", "two-stage") if (exact or not device_trainer) else ("adam", "two-stage", "adadelta", "three-stage")
        if (lockstep_models and n_jobs == 1 and restarts is None and dist_info()[1] == 1
                and (len(unique) > 1 or (not exact and device_trainer))
                and optimization_method in lockstep_methods and set(opt_kwargs) <= {"max_iter"}
                and ("max_iter" in opt_kwargs or optimization_method in ("two-stage", "three-stage"))
                and (not exact or self.x.shape[0] <= 4096)):
            # the reference's default path: independent per-column models trained by first-order recipes -- all of them advance
            # together: sparse models in one device batch with the update steps on the device, otherwise their evaluations
            # overlapping on the GPU (same trajectories as the sequential loop)
            from .sparse import fit_lockstep

            if initial_theta is not None:
                for model in unique:
                    _assign_theta(model, initial_theta)
            done = fit_lockstep(unique, optimization_method, device_trainer=device_trainer, **opt_kwargs)
        if done:
            pass
        elif dist_info()[1] > 1 and len(unique) > 1 and restarts is None:
            # one process per GPU: per-column models go round-robin to ranks, parameters are all-gathered at the end
            from .parallel import run_models_sharded

            run_many = None
            if (lockstep_models and n_jobs == 1 and initial_theta is None and optimization_method in lockstep_methods
                    and set(opt_kwargs) <= {"max_iter"}
                    and ("max_iter" in opt_kwargs or optimization_method in ("two-stage", "three-stage"))
                    and (not exact or self.x.shape[0] <= 4096)):
                from .sparse import fit_lockstep

                def run_many(shard):  # this rank's models advance together (device-resident trainer where they qualify)
                    return fit_lockstep(shard, optimization_method, device_trainer=device_trainer, **opt_kwargs)

            run_models_sharded(unique, run_one, run_many)
        elif n_jobs > 1 and len(unique) > 1:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=int(n_jobs)) as ex:
                list(ex.map(run_one, unique))
            self._slot.release_other_threads()
            if not exact:
                from .sparse import release_other_threads

                release_other_threads()
        else:
            for model in unique:
                run_one(model)

    def _init_models(self, x, y, n_inducing, inducing_initializer: InductionInitializerType = "kmeans") -> None:
        """One model per spatial mode with the reference's initial values (``gpr.py:277-308``)."""
        o = dict(exact=n_inducing is None, ard=False, shared_kernel=False, priors=True, device=0, parameterisation="softplus")
        o.update(self._opts or {})
        if not self.kernel.supported:
            raise NotImplementedError(
                f"kernel {self.kernel_str!r} cannot be constructed with lengthscales= in the reference either (gpr.py:26-28,298)"
            )
        ini_length = float(np.mean(np.abs(x)))
        ls0 = np.full(x.shape[1], ini_length) if o["ard"] else ini_length
        self.release()
        self.models = []
        if o["exact"]:
            if o["shared_kernel"]:
                m = ExactModel(self.kernel_str, x, y, ls0, self._slot, o["device"], o["priors"], o["parameterisation"])
                self.models = [m] * y.shape[1]
            else:
                for i in range(y.shape[1]):
                    self.models.append(ExactModel(self.kernel_str, x, np.ascontiguousarray(y[:, i : i + 1]), ls0, self._slot,
                                                  o["device"], o["priors"], o["parameterisation"]))
            return
        from .sparse import SparseModel

        inducing = self._create_inducing(x, int(n_inducing), inducing_initializer)
        for i in range(y.shape[1]):
            self.models.append(SparseModel(self.kernel_str, x, np.ascontiguousarray(y[:, i : i + 1]), inducing.copy(), ls0,
                                           o["device"], o["priors"], o["parameterisation"]))

    def _create_inducing(self, x, n_inducing: int, method: InductionInitializerType) -> NDArray[Any]:
        """Inducing-input initialisation (``gpr.py:310-320``): KMeans centres or a per-feature linspace diagonal.

        ``"kmeans"`` is the reference's own call, ``KMeans(n_clusters=M, random_state=0, n_init="auto")``.  With
        ``fit(..., kmeans_on_device=True)`` the same k-means++ seeding (scikit-learn's ``kmeans_plusplus`` with the same random
        state, on the mean-centred inputs like ``KMeans.fit``) is followed by Lloyd iterations on the GPU
        (``gpras_kmeans_lloyd``: same stopping rules, fixed-order sums)."""
        if method == "kmeans" and self._opts.get("kmeans_on_device"):
            from sklearn.cluster import kmeans_plusplus

            from .engine import kmeans_lloyd

            x = np.asarray(x, np.float64)
            mu = x.mean(axis=0)
            xc = x - mu
            seeds, _ = kmeans_plusplus(xc, int(n_inducing), random_state=0)
            centres, _, inertia, n_iter = kmeans_lloyd(xc, seeds, device=self._opts.get("device", 0))
            self.kmeans_info = {"inertia": inertia, "n_iter": n_iter}
            return centres + mu
        if method == "kmeans":
            from sklearn.cluster import KMeans

            km = KMeans(n_clusters=n_inducing, random_state=0, n_init="auto")
            km.fit(x)
            return km.cluster_centers_.astype(np.float64)
        elif method == "grid":
            cols = [np.linspace(x[:, j].min(), x[:, j].max(), n_inducing) for j in range(x.shape[1])]
            return np.stack(cols, axis=1).astype(np.float64)
        return None  # the reference falls through the same way for an unknown initializer

    # -- predict --------------------------------------------------------------------------------
    def predict(self, x: NDArray[Any]) -> tuple[NDArray[Any], NDArray[Any]]:
        """Predicted means and VARIANCES, both (n_samples, n_outputs), likelihood noise included (``gpr.py:322-342``)."""
        x = np.asarray(x).astype(np.float64)
        if self._opts.get("shared_kernel") and self._opts.get("exact") and self.models:
            return self.models[0].predict_y(x)
        if self.models and not self._opts.get("exact") and self._opts.get("batched_predict", True):
            # the reference's model family: all per-column sparse models conditioned and predicted in one batched device pass
            from .sparse import predict_batched

            out = predict_batched(self.models, x)
            if out is not None:
                return out
        # Per-column models: each keeps its conditioned factor on a handle of its own while they fit in the budget, so a
        # second predict() costs only the predictor (the reference loops over the models, gpr.py:336-339).
        keep = self._predict_handles_fit()
        means, variances = [], []
        for m in self.models:
            mu, var = m.predict_y(x, keep_handle=keep)
            means.append(mu)
            variances.append(var)
        return np.concatenate(means, axis=1), np.concatenate(variances, axis=1)

    PREDICT_HANDLE_BUDGET_BYTES = 32 << 30

    def _predict_handles_fit(self) -> bool:
        if not self.models:
            return False
        n = self.x.shape[0]
        n_pad = (n + 127) // 128 * 128
        if self._opts.get("exact"):
            per = 3 * 8 * n_pad * n_pad
        else:
            m_ind = int(np.asarray(self.models[0].inducing_variable.Z).shape[0])
            per = 8 * (4 * n_pad * ((m_ind + 127) // 128 * 128) + 8 * 512 * 512)
        return len(self.models) * per <= self.PREDICT_HANDLE_BUDGET_BYTES

    def release(self) -> None:
        """Give the per-model prediction handles and the device batch of these models back (device memory)."""
        for m in self.models:
            rel = getattr(m, "release", None)
            if rel is not None:
                rel()
        if self.models and not (self._opts or {}).get("exact", False):
            from .sparse import release_batches

            release_batches(self.models)

    def predict_std(self, x: NDArray[Any]) -> tuple[NDArray[Any], NDArray[Any]]:
        """Explicit extra: (mean, std); callers of the reference take ``np.sqrt`` themselves (pipeline.py:262-263)."""
        mean, var = self.predict(x)
        return mean, np.sqrt(var)

    # -- persistence ----------------------------------------------------------------------------
    def to_file(self, json_path: str | Path, model_dir: str | Path | None = None) -> None:
        """Pickle with the reference's keys (``gpr.py:344-366``); values are plain ndarrays (GPflow ``Parameter``
        objects cannot be unpickled without GPflow)."""
        d = {
            "kernel": self.kernel_str,
            "data": {"x": self.x, "y": self.y},
            "n_inducing": self.models[0].inducing_variable.Z.shape[0],
            "models": [m.parameter_dict() for m in self.models],
            "gpras_b200": dict(self._opts),
        }
        with open(json_path, mode="wb") as f:
            pickle.dump(d, f)

    @classmethod
    def from_file(cls, json_path: str | Path):
        """Rebuild from :meth:`to_file` output (``gpr.py:368-384``): re-initialise with the "grid" initializer,
        then assign the saved parameters."""
        with open(json_path, mode="rb") as f:
            d = pickle.load(f)
        inst = cls(d["kernel"])
        inst.x, inst.y = d["data"]["x"], d["data"]["y"]
        inst._opts = dict(d.get("gpras_b200", dict(exact=False, ard=False, shared_kernel=False, priors=True)))
        inst._opts["device"] = _default_device()  # the saving process's device index means nothing here
        inst._init_models(inst.x, inst.y, d["n_inducing"], "grid")
        seen = set()
        for ind, params in enumerate(d["models"]):
            m = inst.models[ind]
            if id(m) in seen:
                continue
            seen.add(id(m))
            m.assign_parameters(params)
        return inst


def _default_device() -> int:
    """The CUDA device a fit lands on when the caller names none: torch's current device when torch is loaded and
    initialised for CUDA (``torch.cuda.set_device(local_rank)`` under torchrun), else ``LOCAL_RANK``, else 0."""
    import os
    import sys

    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return int(torch.cuda.current_device())
    return int(os.environ.get("LOCAL_RANK", "0"))


def _assign_theta(model, theta) -> None:
    theta = np.asarray(theta, np.float64)
    model.kernel.variance.assign(theta[0])
    model.likelihood.variance.assign(theta[1])
    nls = model.kernel.lengthscales.size
    model.kernel.lengthscales.assign(theta[2 : 2 + nls] if nls > 1 else theta[2])
