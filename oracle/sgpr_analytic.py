"""Analytic gradient of the GPflow ``SGPR`` collapsed bound (NumPy, FP64) -- the formulas the CUDA path implements.

Test infrastructure only (see ``oracle/__init__.py``).  ``oracle/sgpr.py`` obtains gradients by autograd (as the
reference does with ``tf.GradientTape``, ``gpras/gpr.py:153-155``); this file restates them in closed form in terms
of the well-conditioned intermediates of ``gpflow/models/sgpr.py`` (A, B = I + A A^T, LB, c) and is checked against
the autograd version in ``tests/test_oracle.py``.

With  W = L^-1 (L = chol(Kuu)),  A = W Kuf / sigma,  err = y / sigma,  c = LB^-1 A err,  chat = LB^-T c,
R = I - B^-1 - chat chat^T:

    dF/dKuu = 1/2 W^T (R - A A^T) W                          (symmetric)
    dF/dKuf = (W^T R A + (W^T chat) err^T) / sigma
    dF/dlog s2 = -N/2 + (M - tr B^-1)/2 + |err|^2/2 - |c|^2 + chat^T A A^T chat / 2 + N sf2/(2 s2) - tr(A A^T)/2
and the kernel chain rule   dk/dlog l_d = Fk s_d,   dk/dz_d = -Fk (z_d - x_d) / l_d^2,   dk/dlog sf2 = k.
"""

from __future__ import annotations

import numpy as np
from scipy.linalg import cholesky, solve_triangular

from .kernels import _ls_vector, dk_dlogl_factor, k_of_r2, scaled_sqdist

LOG_2PI = float(np.log(2.0 * np.pi))


def elbo_and_grads(name, x, y, z, variance, lengthscales, noise, jitter=1e-6):
    """Returns (elbo, dict of gradients w.r.t. log variance, log noise, log lengthscale(s) [D], z [M, D])."""
    x, y, z = (np.asarray(a, np.float64) for a in (x, y, z))
    n, d = x.shape
    m = z.shape[0]
    r = y.shape[1]
    ls = _ls_vector(lengthscales, d)
    sigma = np.sqrt(noise)
    r2_uf = scaled_sqdist(z, x, ls)
    r2_uu = scaled_sqdist(z, z, ls)
    kuf = k_of_r2(name, r2_uf, variance)
    kuu0 = k_of_r2(name, r2_uu, variance)
    kuu = kuu0 + jitter * np.eye(m)
    low = cholesky(kuu, lower=True)
    w = solve_triangular(low, np.eye(m), lower=True)
    a = w @ kuf / sigma
    aat = a @ a.T
    b = np.eye(m) + aat
    lb = cholesky(b, lower=True)
    wb = solve_triangular(lb, np.eye(m), lower=True)
    binv = wb.T @ wb
    err = y / sigma
    ae = a @ err  # (M, R)
    c = wb @ ae
    chat = wb.T @ c
    elbo = (-0.5 * n * r * LOG_2PI - r * (np.log(np.diag(lb)).sum() + 0.5 * n * np.log(noise)
                                          + 0.5 * (n * variance / noise - np.trace(aat)))
            - 0.5 * ((err * err).sum() - (c * c).sum()))
    # gradients w.r.t. the covariance blocks (R output columns share the model: sum over columns)
    rr = r * (np.eye(m) - binv) - chat @ chat.T
    g_uu = 0.5 * w.T @ (rr - r * aat) @ w
    g_uf = (w.T @ rr @ a + (w.T @ chat) @ err.T) / sigma
    g_lognoise = (-0.5 * n * r + 0.5 * r * (m - np.trace(binv)) + 0.5 * (err * err).sum() - (c * c).sum()
                  + 0.5 * np.trace(chat.T @ aat @ chat) + 0.5 * r * n * variance / noise - 0.5 * r * np.trace(aat))
    g_logvar = float((g_uu * kuu0).sum() + (g_uf * kuf).sum() - 0.5 * r * n * variance / noise)
    f_uf = dk_dlogl_factor(name, r2_uf, variance)
    f_uu = dk_dlogl_factor(name, r2_uu, variance)
    gf_uf = g_uf * f_uf
    gf_uu = g_uu * f_uu
    g_logls = np.empty(d)
    g_z = np.empty((m, d))
    for j in range(d):
        duf = (z[:, j, None] - x[None, :, j]) / ls[j]
        duu = (z[:, j, None] - z[None, :, j]) / ls[j]
        g_logls[j] = (gf_uf * duf * duf).sum() + (gf_uu * duu * duu).sum()
        g_z[:, j] = -((gf_uf * duf).sum(axis=1) + 2.0 * (gf_uu * duu).sum(axis=1)) / ls[j]
    return float(elbo), {"log_variance": g_logvar, "log_noise": float(g_lognoise), "log_lengthscales": g_logls, "z": g_z}
