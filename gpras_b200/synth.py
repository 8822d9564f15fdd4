"""Seeded synthetic storm-event x mesh-cell data for tests and ``bench.py``.

The reference repository ships no HEC-RAS results (``data/`` holds run-creation inputs only), so
every parity / benchmark input is generated here with ``numpy.random.default_rng(seed)``:

* ``X`` standard-normal features (gpras standardises them, ``gpras/preprocess.py:1037,1280``),
* mode-space targets ``Y = F R + 0.1 eps`` with a smooth latent ``F = sum_k a_k sin(X w_k + b_k)``,
  standardised per column,
* an orthonormal-row EOF basis and per-mode / per-cell scales for the modes -> cells map of
  ``gpras/preprocess.py:1052-1094``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# name -> (N, D, P, C, T, kernel, ard)   BASELINE.json "configs" with the reference convention P = D
CONFIGS = {
    "cfg1": dict(n=256, d=8, p=8, c=5_000, t=1_000, kernel="RBF", ard=False),
    "cfg2": dict(n=2_048, d=16, p=16, c=50_000, t=10_000, kernel="Matern52", ard=True),
    "cfg3": dict(n=8_192, d=32, p=32, c=200_000, t=10_000, kernel="Matern52", ard=True),
    "cfg4": dict(n=16_384, d=64, p=64, c=200_000, t=10_000, kernel="Matern52", ard=True),
    "cfg5": dict(n=8_192, d=32, p=32, c=200_000, t=1_000_000, kernel="Matern52", ard=True),
}


@dataclass
class GPData:
    x: np.ndarray  # (N, D)
    y: np.ndarray  # (N, P)
    x_test: np.ndarray  # (T, D)


def make_gp_data(n: int, d: int, p: int, t: int = 0, seed: int = 0, n_latent: int = 8) -> GPData:
    """Training inputs / mode-space targets / test inputs."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    w = rng.standard_normal((d, n_latent)) / np.sqrt(d)
    b = rng.uniform(0.0, 2.0 * np.pi, n_latent)
    a = rng.uniform(0.5, 1.5, n_latent)
    f = np.sin(x @ w + b) * a
    r = rng.standard_normal((n_latent, p))
    y = f @ r + 0.1 * rng.standard_normal((n, p))
    y = (y - y.mean(axis=0)) / y.std(axis=0)
    x_test = rng.standard_normal((t, d))
    return GPData(x=x, y=y, x_test=x_test)


def fixed_theta(d: int, ard: bool, seed: int = 1):
    """Fixed-theta parity point: variance 1.3, noise 0.05, l_d = U(1.5, 4) sqrt(D / 8)."""
    rng = np.random.default_rng(seed)
    scale = np.sqrt(d / 8.0)
    ls = rng.uniform(1.5, 4.0, d) * scale if ard else np.array([2.1 * scale])
    return 1.3, 0.05, ls


def random_starts(n_starts: int, n_ls: int, seed: int = 2) -> np.ndarray:
    """(R, 2 + n_ls) constrained starts, log-uniform in the reference's ranges (``gpr.py:88-90``):
    variance, lengthscales in 10^[-1, 1]; noise in 10^[-3, 0].  Column order [variance, noise, ls...]."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_starts, 2 + n_ls))
    out[:, 0] = 10.0 ** rng.uniform(-1, 1, n_starts)
    out[:, 1] = 10.0 ** rng.uniform(-3, 0, n_starts)
    out[:, 2:] = 10.0 ** rng.uniform(-1, 1, (n_starts, n_ls))
    return out


@dataclass
class CellMap:
    """State of a fitted ``PreProcessor`` (``gpras/preprocess.py:866-1162``) needed by the
    modes -> cells reverse transform."""

    eofs: np.ndarray  # (P, C_wet), orthonormal rows
    x_mean: np.ndarray  # (P,)
    x_std: np.ndarray  # (P,)
    weights: np.ndarray  # (C_wet,)
    input_mean: np.ndarray  # (C_wet,)
    dry_indices: np.ndarray  # (C,) bool
    elevations: np.ndarray  # (C,)


def make_cell_map(p: int, c: int, seed: int = 0, dry_frac: float = 0.10) -> CellMap:
    rng = np.random.default_rng(seed + 1000)
    dry = rng.uniform(size=c) < dry_frac
    c_wet = int((~dry).sum())
    q, _ = np.linalg.qr(rng.standard_normal((c_wet, p)))
    return CellMap(
        eofs=np.ascontiguousarray(q.T),
        x_mean=np.zeros(p),
        x_std=rng.uniform(0.5, 2.0, p),
        weights=rng.uniform(0.5, 2.0, c_wet),
        input_mean=rng.standard_normal(c_wet),
        dry_indices=dry,
        elevations=rng.uniform(0.0, 10.0, c),
    )
