"""Exact Gaussian-process regression on the CPU (NumPy / SciPy-LAPACK, FP64).

Test infrastructure only (see ``oracle/__init__.py``).  This is oracle O1 of
SURVEY.md section 8c: the exact-GP form BASELINE.json's north_star names, i.e. the
``Z == X`` limit of the GPflow ``SGPR`` models built at ``gpras/gpr.py:293-308``,
and the formulation scikit-learn's ``GaussianProcessRegressor`` evaluates
(``sklearn/gaussian_process/_gpr.py:541-656``) which pins it (``tests/golden``).

One hyperparameter set theta = (variance s_f^2, noise s^2, lengthscales l_1..l_D)
is shared by all P columns of Y:

    Kt    = k(X, X) + s^2 I                 L = chol(Kt)         alpha = Kt^{-1} Y
    LML   = -1/2 tr(Y^T alpha) - P sum_i log L_ii - N P / 2 log(2 pi)
    dLML/dtheta_j = 1/2 tr( (alpha alpha^T - P Kt^{-1}) dKt/dtheta_j )

Gradients are returned with respect to log(theta_j) (``dlog``); the softplus /
prior chain rule of the reference's parameterisation (``gpr.py:303-305``,
GPflow ``positive()`` transforms) lives in :class:`Objective`.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.linalg.lapack import dpotri

from .kernels import _ls_vector, cov, dk_dlogl_factor, k_of_r2, scaled_sqdist

LOG_2PI = float(np.log(2.0 * np.pi))
NOISE_SHIFT = 1e-6  # GPflow Gaussian likelihood: variance = 1e-6 + softplus(u)


@dataclass
class Theta:
    """Constrained hyperparameters. ``lengthscales`` is scalar (isotropic) or (D,) (ARD)."""

    variance: float
    noise: float
    lengthscales: np.ndarray

    def ls_vec(self, d: int) -> np.ndarray:
        return _ls_vector(self.lengthscales, d)


def lml_and_grad(name: str, x: np.ndarray, y: np.ndarray, th: Theta, want_grad: bool = True):
    """LML and d LML / d log(theta): returns (lml, g_variance, g_noise, g_ls[D])."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    n, d = x.shape
    p = y.shape[1]
    ls = th.ls_vec(d)
    r2 = scaled_sqdist(x, x, ls)
    k = k_of_r2(name, r2, th.variance)
    kt = k + th.noise * np.eye(n)
    low = cholesky(kt, lower=True)
    alpha = cho_solve((low, True), y)
    lml = -0.5 * float(np.sum(y * alpha)) - p * float(np.sum(np.log(np.diag(low)))) - 0.5 * n * p * LOG_2PI
    if not want_grad:
        return lml, None, None, None
    kinv, info = dpotri(low, lower=1)
    if info != 0:
        raise np.linalg.LinAlgError(f"dpotri info={info}")
    kinv = np.tril(kinv) + np.tril(kinv, -1).T
    w = alpha @ alpha.T - p * kinv
    g_var = 0.5 * float(np.sum(w * k))
    g_noise = 0.5 * th.noise * float(np.trace(w))
    wf = w * dk_dlogl_factor(name, r2, th.variance)
    g_ls = np.empty(d)
    for j in range(d):
        diff = (x[:, j, None] - x[None, :, j]) / ls[j]
        g_ls[j] = 0.5 * float(np.sum(wf * diff * diff))
    return lml, g_var, g_noise, g_ls


def lml_and_grad_fast(name: str, x: np.ndarray, y: np.ndarray, th: Theta):
    """Same quantities as :func:`lml_and_grad`, arranged for CPU speed (the timed baseline).

    The per-dimension traces use  sum_ij Wf_ij (a_i - b_j)^2 = 2 (sum_i a_i^2 rowsum_i - a^T Wf a)
    so the D passes over an N x N array become one GEMM.  Loses a few digits to
    cancellation; used for timing and checked against the direct form at 1e-6.
    """
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    n, d = x.shape
    p = y.shape[1]
    ls = th.ls_vec(d)
    xs = x / ls
    sq = np.sum(xs * xs, axis=1)
    r2 = sq[:, None] + sq[None, :] - 2.0 * (xs @ xs.T)
    np.maximum(r2, 0.0, out=r2)
    np.fill_diagonal(r2, 0.0)
    k = k_of_r2(name, r2, th.variance)
    kt = k.copy()
    kt[np.diag_indices(n)] += th.noise
    low = cholesky(kt, lower=True, overwrite_a=True, check_finite=False)
    alpha = cho_solve((low, True), y, check_finite=False)
    lml = -0.5 * float(np.sum(y * alpha)) - p * float(np.sum(np.log(np.diag(low)))) - 0.5 * n * p * LOG_2PI
    kinv, info = dpotri(low, lower=1, overwrite_c=1)
    if info != 0:
        raise np.linalg.LinAlgError(f"dpotri info={info}")
    il = np.tril_indices(n, -1)
    kinv.T[il] = kinv[il]
    w = alpha @ alpha.T
    w -= p * kinv
    g_var = 0.5 * float(np.vdot(w, k))
    g_noise = 0.5 * th.noise * float(np.trace(w))
    w *= dk_dlogl_factor(name, r2, th.variance)
    rows = w.sum(axis=1)
    g_ls = rows @ (xs * xs) - np.einsum("id,id->d", xs, w @ xs)
    return lml, g_var, g_noise, g_ls


def predict(name: str, x: np.ndarray, y: np.ndarray, th: Theta, xs: np.ndarray):
    """Posterior of y* (noise included, GPflow ``predict_y`` semantics, ``gpr.py:337``).

    Returns (mean (T, P), var (T, P)); the variance is identical across columns because
    theta is shared.
    """
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    xs = np.asarray(xs, np.float64)
    n, d = x.shape
    ls = th.ls_vec(d)
    kt = cov(name, x, x, th.variance, ls) + th.noise * np.eye(n)
    low = cholesky(kt, lower=True)
    alpha = cho_solve((low, True), y)
    ks = cov(name, xs, x, th.variance, ls)  # (T, N)
    mean = ks @ alpha
    v = solve_triangular(low, ks.T, lower=True)
    var = th.variance - np.sum(v * v, axis=0) + th.noise
    return mean, np.repeat(var[:, None], y.shape[1], axis=1)


# ----------------------------------------------------------------------------------------------
# Parameterisation + priors of the reference (gpras/gpr.py:298-305; GPflow positive() = softplus,
# Gaussian-likelihood variance additionally shifted by 1e-6; LogNormal(0, 1) priors evaluated on the
# constrained value, no Jacobian term -- GPflow PriorOn.CONSTRAINED).
# ----------------------------------------------------------------------------------------------
def softplus(u):
    u = np.asarray(u, np.float64)
    return np.logaddexp(0.0, u)


def softplus_inv(v):
    v = np.asarray(v, np.float64)
    return v + np.log(-np.expm1(-v))


def sigmoid(u):
    u = np.asarray(u, np.float64)
    return 0.5 * (1.0 + np.tanh(0.5 * u))


def lognormal01_logpdf(v):
    v = np.asarray(v, np.float64)
    lv = np.log(v)
    return -lv - 0.5 * LOG_2PI - 0.5 * lv * lv


def lognormal01_dlogpdf(v):
    v = np.asarray(v, np.float64)
    return -(1.0 + np.log(v)) / v


class Objective:
    """loss(u) = -(LML + sum log prior) as a function of unconstrained u (what the optimisers see).

    ``space='softplus'`` is the reference's GPflow parameterisation; ``space='log'`` is
    scikit-learn's (theta = exp(u)); ``priors=False`` drops the LogNormal terms.
    Layout of u: [variance, noise, lengthscale(s)] with 1 (isotropic) or D (ARD) lengthscales.
    """

    def __init__(self, name, x, y, ard=False, space="softplus", priors=True, fast=False):
        self.name, self.x, self.y = name, np.asarray(x, np.float64), np.asarray(y, np.float64)
        self.ard, self.space, self.priors, self.fast = ard, space, priors, fast
        self.n_ls = self.x.shape[1] if ard else 1
        self.n_evals = 0

    def constrain(self, u):
        u = np.asarray(u, np.float64)
        if self.space == "log":
            v = np.exp(u)
        else:
            v = softplus(u)
            v[1] = v[1] + NOISE_SHIFT
        return v

    def unconstrain(self, v):
        v = np.asarray(v, np.float64).copy()
        if self.space == "log":
            return np.log(v)
        v[1] = v[1] - NOISE_SHIFT
        return softplus_inv(v)

    def dconstrain(self, u):
        u = np.asarray(u, np.float64)
        return np.exp(u) if self.space == "log" else sigmoid(u)

    def __call__(self, u):
        v = self.constrain(u)
        th = Theta(v[0], v[1], v[2:] if self.ard else v[2])
        f = lml_and_grad_fast if self.fast else lml_and_grad
        lml, g_var, g_noise, g_ls = f(self.name, self.x, self.y, th)
        self.n_evals += 1
        glog = np.concatenate([[g_var, g_noise], g_ls if self.ard else [np.sum(g_ls)]])
        dv = glog / v  # d LML / d theta
        obj = lml
        if self.priors:
            obj = obj + float(np.sum(lognormal01_logpdf(v)))
            dv = dv + lognormal01_dlogpdf(v)
        return -obj, -(dv * self.dconstrain(u))


def fit_lbfgs(obj: Objective, u0, max_iter=1000):
    """SciPy L-BFGS-B exactly as ``gpflow.optimizers.Scipy`` calls it (``gpr.py:197-203``)."""
    from scipy.optimize import minimize

    return minimize(obj, np.asarray(u0, np.float64), jac=True, method="L-BFGS-B", options={"maxiter": max_iter})
