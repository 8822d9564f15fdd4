"""Stationary covariance functions and their log-lengthscale derivative factors (NumPy, FP64).

Test infrastructure only (see ``oracle/__init__.py``).

Follows the kernel classes the reference selects in ``gpras/gpr.py:21-29`` (GPflow
``IsotropicStationary`` family; forms recalled in SURVEY.md section 3.4):

    RBF          s2 * exp(-r^2 / 2)
    Matern12     s2 * exp(-r)
    Exponential  s2 * exp(-r / 2)          (GPflow's Exponential, != Matern12)
    Matern32     s2 * (1 + sqrt3 r) exp(-sqrt3 r)
    Matern52     s2 * (1 + sqrt5 r + 5 r^2 / 3) exp(-sqrt5 r)

with r^2 = sum_d ((x_d - x'_d) / l_d)^2 evaluated by direct differences (the
scikit-learn ``cdist`` form; GPflow's ``|a|^2 + |b|^2 - 2ab`` differs by rounding
only).  ``l`` may be a scalar (what gpras passes, ``gpr.py:289,298``) or a
length-D vector (ARD).
"""

from __future__ import annotations

import numpy as np

KERNEL_NAMES = ("RBF", "Matern12", "Matern32", "Matern52", "Exponential")
SQRT3 = np.sqrt(3.0)
SQRT5 = np.sqrt(5.0)


def _ls_vector(lengthscales, d: int) -> np.ndarray:
    ls = np.asarray(lengthscales, dtype=np.float64).reshape(-1)
    if ls.size == 1:
        ls = np.full(d, ls[0])
    if ls.size != d:
        raise ValueError(f"lengthscales has {ls.size} entries for {d} features")
    return ls


def scaled_sqdist(x1: np.ndarray, x2: np.ndarray, lengthscales) -> np.ndarray:
    """r^2[i, j] = sum_d ((x1[i, d] - x2[j, d]) / l_d)^2 by direct differences."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    ls = _ls_vector(lengthscales, x1.shape[1])
    r2 = np.zeros((x1.shape[0], x2.shape[0]))
    for d in range(x1.shape[1]):
        diff = (x1[:, d, None] - x2[None, :, d]) / ls[d]
        r2 += diff * diff
    return r2


def k_of_r2(name: str, r2: np.ndarray, variance: float) -> np.ndarray:
    """Covariance as a function of the scaled squared distance."""
    if name == "RBF":
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    if name == "Matern12":
        return variance * np.exp(-r)
    if name == "Exponential":
        return variance * np.exp(-0.5 * r)
    if name == "Matern32":
        return variance * (1.0 + SQRT3 * r) * np.exp(-SQRT3 * r)
    if name == "Matern52":
        return variance * (1.0 + SQRT5 * r + (5.0 / 3.0) * r2) * np.exp(-SQRT5 * r)
    raise KeyError(name)


def dk_dlogl_factor(name: str, r2: np.ndarray, variance: float) -> np.ndarray:
    """F with  d k / d log l_d = F * s_d,  s_d = ((x_d - x'_d) / l_d)^2  (SURVEY.md 3.6)."""
    if name == "RBF":
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    with np.errstate(divide="ignore", invalid="ignore"):
        if name == "Matern12":
            return np.where(r > 0.0, variance * np.exp(-r) / r, 0.0)
        if name == "Exponential":
            return np.where(r > 0.0, variance * np.exp(-0.5 * r) / (2.0 * r), 0.0)
    if name == "Matern32":
        return 3.0 * variance * np.exp(-SQRT3 * r)
    if name == "Matern52":
        return (5.0 / 3.0) * variance * (1.0 + SQRT5 * r) * np.exp(-SQRT5 * r)
    raise KeyError(name)


def cov(name: str, x1: np.ndarray, x2: np.ndarray, variance: float, lengthscales) -> np.ndarray:
    """k(x1, x2) without any noise / jitter term."""
    return k_of_r2(name, scaled_sqdist(x1, x2, lengthscales), variance)
