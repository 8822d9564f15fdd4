"""GPflow ``SGPR`` (Titsias collapsed bound) restated in torch-CPU float64 with autograd.

Test infrastructure only (see ``oracle/__init__.py``).  PARITY UNPINNED: GPflow 2.x is an
un-vendored, un-pinned dependency of the reference (``pyproject.toml:17``) and cannot be installed
here; this file restates its published algorithm (``gpflow/models/sgpr.py``: ``_common_calculation``,
``logdet_term``, ``quad_term``, ``elbo``, ``predict_f``) as summarised in SURVEY.md section 3.4, for
exactly the model the reference builds at ``gpras/gpr.py:293-308``.  What it IS pinned to
(``tests/test_oracle.py``): the published definition of the bound and of the predictive distribution
(Titsias 2009, eqs. 6 and 9) evaluated densely -- N x N matrices, explicit inverses, SciPy's
multivariate normal, scikit-learn's kernel implementations, sharing no code or algebra with this
file -- to 1e-9, for every kernel; the exact GP's LML (itself pinned to scikit-learn) at Z = X; and
finite differences for the gradients.  The model:

    one output column, zero mean function, Gaussian likelihood (variance 1.0 initially, softplus
    + 1e-6 shift), stationary kernel with softplus-constrained variance / lengthscales,
    LogNormal(0, 1) priors on the three constrained hyperparameters, trainable inducing inputs Z,
    jitter 1e-6 on Kuu.

``training_loss = -(ELBO + log prior)`` is what every optimiser recipe in ``gpr.py:44-214`` minimises;
gradients come from ``torch.autograd`` (mirrors ``tf.GradientTape`` at ``gpr.py:153-155``).
Distances use direct differences (see ``oracle/kernels.py``).
"""

from __future__ import annotations

import math

import numpy as np
import torch

JITTER = 1e-6
NOISE_SHIFT = 1e-6
LOG_2PI = math.log(2.0 * math.pi)
SQRT3 = math.sqrt(3.0)
SQRT5 = math.sqrt(5.0)


def _k_of_r2(name: str, r2: torch.Tensor, variance: torch.Tensor) -> torch.Tensor:
    if name == "RBF":
        return variance * torch.exp(-0.5 * r2)
    r = torch.sqrt(torch.clamp(r2, min=1e-36))  # GPflow scaled_euclid_dist
    if name == "Matern12":
        return variance * torch.exp(-r)
    if name == "Exponential":
        return variance * torch.exp(-0.5 * r)
    if name == "Matern32":
        return variance * (1.0 + SQRT3 * r) * torch.exp(-SQRT3 * r)
    if name == "Matern52":
        return variance * (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r) * torch.exp(-SQRT5 * r)
    raise KeyError(name)


def _cov(name, a, b, variance, ls):
    diff = (a[:, None, :] - b[None, :, :]) / ls
    return _k_of_r2(name, (diff * diff).sum(-1), variance)


def elbo(name, x, y, z, variance, ls, noise, jitter=JITTER):
    """Collapsed bound for y of shape (N, R) (R = 1 in the reference)."""
    n, r = y.shape
    m = z.shape[0]
    sigma = torch.sqrt(noise)
    kuf = _cov(name, z, x, variance, ls)
    kuu = _cov(name, z, z, variance, ls) + jitter * torch.eye(m, dtype=x.dtype)
    low = torch.linalg.cholesky(kuu)
    a = torch.linalg.solve_triangular(low, kuf, upper=False) / sigma
    aat = a @ a.T
    b = aat + torch.eye(m, dtype=x.dtype)
    lb = torch.linalg.cholesky(b)
    err = y / sigma
    c = torch.linalg.solve_triangular(lb, a @ err, upper=False)
    trace = n * variance / noise - torch.trace(aat)
    logdet = -r * (torch.log(torch.diagonal(lb)).sum() + 0.5 * n * torch.log(noise) + 0.5 * trace)
    quad = -0.5 * ((err * err).sum() - (c * c).sum())
    return -0.5 * n * r * LOG_2PI + logdet + quad


def log_prior(v: torch.Tensor) -> torch.Tensor:
    """LogNormal(0, 1) log-density on constrained values (``gpr.py:303-305``), summed."""
    lv = torch.log(v)
    return (-lv - 0.5 * LOG_2PI - 0.5 * lv * lv).sum()


def constrain(u_var, u_ls, u_noise):
    sp = torch.nn.functional.softplus
    return sp(u_var), sp(u_ls), sp(u_noise) + NOISE_SHIFT


def training_loss_and_grads(name, x, y, z, u_var, u_ls, u_noise, train_hypers=True, train_z=True, jitter=JITTER):
    """loss = -(ELBO + log prior over *trainable* hyperparameters) and gradients w.r.t. the
    unconstrained variables (``u_*``) and Z.  All inputs are array-likes; returns numpy."""
    t = lambda a, g: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=g)  # noqa: E731
    xt, yt = t(x, False), t(y, False)
    zt = t(z, train_z)
    uv, ul, un = t(u_var, train_hypers), t(np.atleast_1d(u_ls), train_hypers), t(u_noise, train_hypers)
    var, ls, noise = constrain(uv, ul, un)
    obj = elbo(name, xt, yt, zt, var, ls, noise, jitter)
    if train_hypers:
        obj = obj + log_prior(var) + log_prior(ls) + log_prior(noise)
    loss = -obj
    wrt = [v for v in (uv, ul, un, zt) if v.requires_grad]
    grads = torch.autograd.grad(loss, wrt) if wrt else ()
    out = {"loss": float(loss.detach())}
    it = iter(grads)
    if train_hypers:
        out["u_var"], out["u_ls"], out["u_noise"] = (next(it).numpy() for _ in range(3))
    if train_z:
        out["z"] = next(it).numpy()
    return out


def predict_y(name, x, y, z, variance, ls, noise, xs, jitter=JITTER):
    """GPflow ``SGPR.predict_f`` + Gaussian-likelihood noise (``predict_y``; ``gpr.py:337``)."""
    with torch.no_grad():
        t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64)  # noqa: E731
        x, y, z, xs = t(x), t(y), t(z), t(xs)
        variance, ls, noise = t(variance), t(np.atleast_1d(ls)), t(noise)
        m = z.shape[0]
        sigma = torch.sqrt(noise)
        kuf = _cov(name, z, x, variance, ls)
        kuu = _cov(name, z, z, variance, ls) + jitter * torch.eye(m, dtype=x.dtype)
        kus = _cov(name, z, xs, variance, ls)
        low = torch.linalg.cholesky(kuu)
        a = torch.linalg.solve_triangular(low, kuf, upper=False) / sigma
        b = a @ a.T + torch.eye(m, dtype=x.dtype)
        lb = torch.linalg.cholesky(b)
        c = torch.linalg.solve_triangular(lb, a @ (y / sigma), upper=False)
        tmp1 = torch.linalg.solve_triangular(low, kus, upper=False)
        tmp2 = torch.linalg.solve_triangular(lb, tmp1, upper=False)
        mean = tmp2.T @ c
        var = variance + (tmp2 * tmp2).sum(0) - (tmp1 * tmp1).sum(0) + noise
        return mean.numpy(), var[:, None].repeat(1, y.shape[1]).numpy()
