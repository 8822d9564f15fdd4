"""Thin object wrapper over the exact-GP handle of the C ABI (``include/gpras_b200.h``).

``ExactGP`` owns one ``gpras_gp`` handle: one GPU, one stream, one (N, D, P) problem with a single
hyperparameter set shared by all P target columns.  It replaces, for the exact (``Z == X``) model,
what the reference obtains from ``gpflow.models.SGPR`` at ``gpras/gpr.py:293-308``:
``training_loss`` + gradient (``lml_grad``) and ``predict_y`` (``condition`` + ``predict``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KERNEL_IDS, check, ptr


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class ExactGP:
    def __init__(self, kernel: str, n: int, d: int, p: int, device: int = 0):
        self.lib = _lib.load()
        if self.lib.gpras_device_count() <= 0:
            raise _lib.GprasError("no CUDA device visible: gpras_b200 has no CPU fallback")
        self.kernel, self.n, self.d, self.p, self.device = kernel, int(n), int(d), int(p), int(device)
        h = C.c_void_p()
        check(self.lib.gpras_gp_create(C.byref(h), device, KERNEL_IDS[kernel], self.n, self.d, self.p))
        self._h = h
        self._keep = []
        self._cond_theta = None  # theta the handle is conditioned at (factor + alpha resident), or None

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.gpras_gp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_blocking_wait(self, enabled: bool) -> None:
        """Sleep instead of spinning while waiting for an evaluation (several host threads per GPU)."""
        check(self.lib.gpras_gp_set_blocking_wait(self._h, int(enabled)))

    # ---- data -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None) -> None:
        check(self.lib.gpras_gp_set_stream(self._h, cuda_stream or None))

    def set_data(self, x, y) -> None:
        """x (N, D), y (N, P): numpy (host) arrays or CUDA float64 torch tensors."""
        on_device = not isinstance(x, np.ndarray) and hasattr(x, "data_ptr")
        if not on_device:
            x, y = _f64(x), _f64(y)
        if tuple(x.shape) != (self.n, self.d) or tuple(y.shape) != (self.n, self.p):
            raise ValueError(f"expected x {(self.n, self.d)} and y {(self.n, self.p)}, got {tuple(x.shape)}, {tuple(y.shape)}")
        self._keep = [x, y]
        self._cond_theta = None
        check(self.lib.gpras_gp_set_data(self._h, ptr(x), ptr(y), int(on_device)))

    def theta_vector(self, variance: float, noise: float, lengthscales) -> np.ndarray:
        ls = np.asarray(lengthscales, np.float64).reshape(-1)
        if ls.size == 1:
            ls = np.full(self.d, ls[0])
        if ls.size != self.d:
            raise ValueError(f"{ls.size} lengthscales for {self.d} features")
        return np.concatenate([[float(variance), float(noise)], ls])

    # ---- objective --------------------------------------------------------------------------
    def lml_grad(self, theta, want_grad: bool = True):
        """(LML, d LML / d log theta [2 + D]) at theta = [variance, noise, l_0..l_{D-1}]."""
        theta = _f64(theta)
        lml = C.c_double()
        grad = np.empty(2 + self.d) if want_grad else None
        self._cond_theta = None
        check(self.lib.gpras_gp_lml_grad(self._h, ptr(theta), C.byref(lml), ptr(grad) if want_grad else None))
        return lml.value, grad

    def lml_grad_host(self, x, y, theta):
        """End-to-end call from host buffers (uploads x, y, theta; returns host scalars)."""
        x, y, theta = _f64(x), _f64(y), _f64(theta)
        lml = C.c_double()
        grad = np.empty(2 + self.d)
        self._cond_theta = None
        check(self.lib.gpras_gp_lml_grad_host(self._h, ptr(x), ptr(y), ptr(theta), C.byref(lml), ptr(grad)))
        return lml.value, grad

    def enqueue(self, theta, want_grad: bool = True) -> None:
        theta = _f64(theta)
        self._cond_theta = None
        check(self.lib.gpras_gp_lml_grad_enqueue(self._h, ptr(theta), int(want_grad)))

    def fetch(self):
        lml = C.c_double()
        grad = np.empty(2 + self.d)
        check(self.lib.gpras_gp_lml_grad_fetch(self._h, C.byref(lml), ptr(grad)))
        return lml.value, grad

    # ---- prediction -------------------------------------------------------------------------
    def condition(self, theta, force: bool = False) -> None:
        """Factorise at ``theta`` and keep the factor and alpha resident for ``predict``.  Conditioning again at the theta
        the handle already holds is free (the reference's second caller predicts once per plan with an unchanged model,
        ``gpras/preprocess.py:601-606``: that must not cost N^3 per call)."""
        theta = _f64(theta)
        if not force and self._cond_theta is not None and np.array_equal(theta, self._cond_theta):
            return
        self._cond_theta = None
        check(self.lib.gpras_gp_condition(self._h, ptr(theta)))
        self._cond_theta = theta.copy()

    def conditioned_at(self, theta) -> bool:
        return self._cond_theta is not None and np.array_equal(_f64(theta), self._cond_theta)

    def predict(self, xs):
        """(mean (T, P), var (T, P)) with likelihood noise included (``predict_y``)."""
        xs = _f64(xs)
        if xs.ndim != 2 or xs.shape[1] != self.d:
            raise ValueError(f"expected (T, {self.d}) test inputs, got {xs.shape}")
        t = xs.shape[0]
        mean = np.empty((t, self.p))
        var = np.empty((t, self.p))
        check(self.lib.gpras_gp_predict(self._h, ptr(xs), t, ptr(mean), ptr(var), 0))
        return mean, var

    def set_cell_map(self, e_mean, bias) -> None:
        e_mean, bias = _f64(e_mean), _f64(bias)
        if e_mean.shape != (self.p, bias.shape[0]):
            raise ValueError("e_mean must be (P, C) and bias (C,)")
        check(self.lib.gpras_gp_set_cell_map(self._h, ptr(e_mean), ptr(bias), bias.shape[0]))

    def cell_pitch(self) -> int:
        return int(self.lib.gpras_gp_cell_pitch(self._h))

    def predict_cells(self, xs, cell_mean=None, cell_var=None, want_modes: bool = True, modes_out=None):
        """Predict and expand to mesh cells on the device.  ``cell_mean`` / ``cell_var`` are CUDA float64
        torch tensors of shape (T, cell_pitch()) or None (tiles go to an internal ring buffer).  ``modes_out``: a pair of
        contiguous (T, P) CUDA tensors that receive the mode-space mean / variance instead of host arrays (sharded sweeps
        all-gather them over NCCL without a host round trip)."""
        on_device = not isinstance(xs, np.ndarray) and hasattr(xs, "data_ptr")
        if not on_device:
            xs = _f64(xs)
        t = int(xs.shape[0])
        if modes_out is not None:
            mm, mv = modes_out
            if tuple(mm.shape) != (t, self.p) or tuple(mv.shape) != (t, self.p) or not (mm.is_contiguous() and mv.is_contiguous()):
                raise ValueError(f"modes_out must be two contiguous ({t}, {self.p}) tensors")
            want_modes = True
        else:
            mm = np.empty((t, self.p)) if want_modes else None
            mv = np.empty((t, self.p)) if want_modes else None
        ldc = int(cell_mean.shape[1]) if cell_mean is not None else (int(cell_var.shape[1]) if cell_var is not None else 0)
        check(
            self.lib.gpras_gp_predict_cells(
                self._h, ptr(xs), t, int(on_device), ptr(mm) if want_modes else None, ptr(mv) if want_modes else None,
                ptr(cell_mean) if cell_mean is not None else None, ptr(cell_var) if cell_var is not None else None, ldc,
            )
        )
        return mm, mv

    # ---- introspection ----------------------------------------------------------------------
    def get_matrix(self, which: int) -> np.ndarray:
        out = np.empty((self.n, self.p if which == 4 else self.n))
        check(self.lib.gpras_gp_get_matrix(self._h, which, ptr(out)))
        return out

    def last_launches(self) -> int:
        return int(self.lib.gpras_gp_last_launches(self._h))

    def set_stage_timing(self, enabled: bool) -> None:
        check(self.lib.gpras_gp_set_stage_timing(self._h, int(enabled)))

    def last_stage_ms(self) -> dict:
        ms = np.zeros(7)
        check(self.lib.gpras_gp_last_stage_ms(self._h, ptr(ms)))
        return dict(zip(["cov", "potrf", "trtri", "lauum", "alpha", "grad", "total"], ms.tolist()))


class SparseGP:
    """One sparse (inducing-point) model on one GPU: GPflow ``SGPR`` semantics (``gpras/gpr.py:293-308``) behind
    ``gpras_sgpr_*``: collapsed bound + analytic gradient w.r.t. hyperparameters and inducing inputs, ``predict_y``."""

    def __init__(self, kernel: str, n: int, d: int, m: int, r: int = 1, device: int = 0):
        self.lib = _lib.load()
        if self.lib.gpras_device_count() <= 0:
            raise _lib.GprasError("no CUDA device visible: gpras_b200 has no CPU fallback")
        self.kernel, self.n, self.d, self.m, self.r, self.device = kernel, int(n), int(d), int(m), int(r), int(device)
        h = C.c_void_p()
        check(self.lib.gpras_sgpr_create(C.byref(h), device, KERNEL_IDS[kernel], self.n, self.d, self.m, self.r))
        self._h = h
        self._cond = None  # (theta, z, jitter) the handle is conditioned at

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.gpras_sgpr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_data(self, x, y) -> None:
        x, y = _f64(x), _f64(y)
        if x.shape != (self.n, self.d) or y.shape != (self.n, self.r):
            raise ValueError(f"expected x {(self.n, self.d)} and y {(self.n, self.r)}, got {x.shape}, {y.shape}")
        self._cond = None
        check(self.lib.gpras_sgpr_set_data(self._h, ptr(x), ptr(y), 0))

    def theta_vector(self, variance: float, noise: float, lengthscales) -> np.ndarray:
        ls = np.asarray(lengthscales, np.float64).reshape(-1)
        if ls.size == 1:
            ls = np.full(self.d, ls[0])
        return np.concatenate([[float(variance), float(noise)], ls])

    def elbo_grad(self, theta, z, jitter: float = 1e-6, want_grad: bool = True):
        """(ELBO, d/dlog theta [2 + D], d/dZ [M, D]); gradients are None when ``want_grad`` is False."""
        theta, z = _f64(theta), _f64(z)
        if z.shape != (self.m, self.d):
            raise ValueError(f"expected inducing inputs {(self.m, self.d)}, got {z.shape}")
        elbo = C.c_double()
        gt = np.empty(2 + self.d) if want_grad else None
        gz = np.empty((self.m, self.d)) if want_grad else None
        self._cond = None
        check(self.lib.gpras_sgpr_elbo_grad(self._h, ptr(theta), ptr(z), float(jitter), C.byref(elbo),
                                            ptr(gt) if want_grad else None, ptr(gz) if want_grad else None))
        return elbo.value, gt, gz

    def enqueue(self, theta, z, jitter: float = 1e-6, want_grad: bool = True) -> None:
        """Start one evaluation on the handle's stream (replayed from a CUDA graph after the first two calls)."""
        theta, z = _f64(theta), _f64(z)
        if z.shape != (self.m, self.d):
            raise ValueError(f"expected inducing inputs {(self.m, self.d)}, got {z.shape}")
        self._cond = None
        check(self.lib.gpras_sgpr_elbo_grad_enqueue(self._h, ptr(theta), ptr(z), float(jitter), int(want_grad)))
        self._want_grad = bool(want_grad)

    def fetch(self):
        elbo = C.c_double()
        gt = np.empty(2 + self.d) if self._want_grad else None
        gz = np.empty((self.m, self.d)) if self._want_grad else None
        check(self.lib.gpras_sgpr_elbo_grad_fetch(self._h, C.byref(elbo), ptr(gt) if self._want_grad else None,
                                                  ptr(gz) if self._want_grad else None))
        return elbo.value, gt, gz

    def condition(self, theta, z, jitter: float = 1e-6, force: bool = False) -> None:
        theta, z = _f64(theta), _f64(z)
        if theta.shape != (2 + self.d,):
            raise ValueError(f"expected theta of length {2 + self.d}, got {theta.shape}")
        if z.shape != (self.m, self.d):
            raise ValueError(f"expected inducing inputs {(self.m, self.d)}, got {z.shape}")
        c = self._cond
        if not force and c is not None and c[2] == float(jitter) and np.array_equal(c[0], theta) and np.array_equal(c[1], z):
            return
        self._cond = None
        check(self.lib.gpras_sgpr_condition(self._h, ptr(theta), ptr(z), float(jitter)))
        self._cond = (theta.copy(), z.copy(), float(jitter))

    def predict(self, xs):
        xs = _f64(xs)
        if xs.ndim != 2 or xs.shape[1] != self.d:
            raise ValueError(f"expected (T, {self.d}) test inputs, got {xs.shape}")
        t = xs.shape[0]
        mean, var = np.empty((t, self.r)), np.empty((t, self.r))
        check(self.lib.gpras_sgpr_predict(self._h, ptr(xs), t, ptr(mean), ptr(var)))
        return mean, var

    def last_launches(self) -> int:
        return int(self.lib.gpras_sgpr_last_launches(self._h))


class SparseBatch:
    """The per-column sparse models of one fit on one GPU, evaluated together and trained on the device
    (``gpras_sgpr_batch_*``): ``p`` independent SGPR models over the same inputs with ``m <= 128`` inducing points each."""

    MAX_INDUCING = 128

    def __init__(self, kernel: str, n: int, d: int, m: int, p: int, device: int = 0):
        self.lib = _lib.load()
        if self.lib.gpras_device_count() <= 0:
            raise _lib.GprasError("no CUDA device visible: gpras_b200 has no CPU fallback")
        self.kernel, self.n, self.d, self.m, self.p, self.device = kernel, int(n), int(d), int(m), int(p), int(device)
        h = C.c_void_p()
        check(self.lib.gpras_sgpr_batch_create(C.byref(h), device, KERNEL_IDS[kernel], self.n, self.d, self.m, self.p))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.gpras_sgpr_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_data(self, x, y) -> None:
        x, y = _f64(x), _f64(y)
        if x.shape != (self.n, self.d) or y.shape != (self.n, self.p):
            raise ValueError(f"expected x {(self.n, self.d)} and y {(self.n, self.p)}, got {x.shape}, {y.shape}")
        self._cond = None
        check(self.lib.gpras_sgpr_batch_set_data(self._h, ptr(x), ptr(y)))

    def elbo_grad(self, theta, z, jitter: float = 1e-6):
        """(ELBO [p], d/dlog theta [p, 2 + D], d/dZ [p, M, D], info [p]) of all models in one pass."""
        theta, z = _f64(theta), _f64(z)
        if theta.shape != (self.p, 2 + self.d) or z.shape != (self.p, self.m, self.d):
            raise ValueError(f"expected theta {(self.p, 2 + self.d)} and z {(self.p, self.m, self.d)}, got {theta.shape}, {z.shape}")
        elbo, gt, gz = np.empty(self.p), np.empty((self.p, 2 + self.d)), np.empty((self.p, self.m, self.d))
        info = np.zeros(self.p, np.int32)
        self._cond = None
        check(self.lib.gpras_sgpr_batch_elbo_grad(self._h, ptr(theta), ptr(z), float(jitter), ptr(elbo), ptr(gt), ptr(gz),
                                                  info.ctypes.data))
        return elbo, gt, gz, info

    def adam(self, u, n_ls: int, train_hypers: bool, train_z: bool, max_iter: int, learning_rate: float = 0.001,
             jitter: float = 1e-6, transform: str = "softplus", priors: bool = True, noise_floor: float = 1e-6, rule: str = "adam"):
        """One Adam stage of all models on the device (``gpras/gpr.py:147-173``), or with ``rule="adadelta"`` exactly
        ``max_iter`` Keras-Adadelta steps (``gpr.py:176-192``).  ``u`` (p, 2 + n_ls + M D): unconstrained
        [variance, noise, lengthscale(s), Z] per model.  Returns (u, losses [max_iter, p], steps [p])."""
        u = _f64(u).copy()
        nu = 2 + int(n_ls) + self.m * self.d
        if u.shape != (self.p, nu):
            raise ValueError(f"expected u of shape {(self.p, nu)}, got {u.shape}")
        losses = np.empty((int(max_iter), self.p))
        iters, info = np.zeros(self.p, np.int32), np.zeros(self.p, np.int32)
        self._cond = None
        if rule not in ("adam", "adadelta"):
            raise ValueError(f"unknown update rule {rule!r}")
        check(self.lib.gpras_sgpr_batch_train(self._h, ptr(u), int(n_ls), int(train_hypers), int(train_z), int(max_iter),
                                              float(learning_rate), float(jitter), 1 if transform == "log" else 0, int(priors),
                                              float(noise_floor), 1 if rule == "adadelta" else 0,
                                              ptr(losses) if max_iter > 0 else None, iters.ctypes.data, info.ctypes.data))
        bad = np.flatnonzero(info)
        if bad.size:
            raise _lib.NotPositiveDefiniteError(
                f"Kuu or B lost positive definiteness in model {int(bad[0])} (first failing pivot {int(info[bad[0]])})")
        return u, losses, iters

    @property
    def fused(self) -> bool:
        """The fused evaluation (and with it batched prediction) covers m <= 64 inducing points and d <= 32 features."""
        return self.m <= 64 and self.d <= 32

    def condition(self, theta, z, jitter: float = 1e-6) -> None:
        """Condition all models for ``predict`` (skipped when they are already conditioned at the same values)."""
        theta, z = _f64(theta), _f64(z)
        if theta.shape != (self.p, 2 + self.d) or z.shape != (self.p, self.m, self.d):
            raise ValueError(f"expected theta {(self.p, 2 + self.d)} and z {(self.p, self.m, self.d)}, got {theta.shape}, {z.shape}")
        c = getattr(self, "_cond", None)
        if c is not None and c[2] == float(jitter) and np.array_equal(c[0], theta) and np.array_equal(c[1], z):
            return
        self._cond = None
        info = np.zeros(self.p, np.int32)
        check(self.lib.gpras_sgpr_batch_condition(self._h, ptr(theta), ptr(z), float(jitter), info.ctypes.data))
        bad = np.flatnonzero(info)
        if bad.size:
            raise _lib.NotPositiveDefiniteError(
                f"Kuu or B lost positive definiteness in model {int(bad[0])} (first failing pivot {int(info[bad[0]])})")
        self._cond = (theta.copy(), z.copy(), float(jitter))

    def predict(self, xs):
        """``predict_y`` of all models: (mean, variance), both (T, p), likelihood noise included."""
        xs = _f64(xs)
        if xs.ndim != 2 or xs.shape[1] != self.d:
            raise ValueError(f"expected (T, {self.d}) test inputs, got {xs.shape}")
        t = xs.shape[0]
        mean, var = np.empty((t, self.p)), np.empty((t, self.p))
        check(self.lib.gpras_sgpr_batch_predict(self._h, ptr(xs), t, ptr(mean), ptr(var)))
        return mean, var

    def last_launches(self) -> int:
        return int(self.lib.gpras_sgpr_batch_last_launches(self._h))


def kmeans_lloyd(x, centers0, max_iter: int = 300, tol: float = 1e-4, device: int = 0):
    """Lloyd iterations of ``sklearn.cluster.KMeans`` from the initial centres ``centers0`` on the GPU
    (``gpras_kmeans_lloyd``; the reference's inducing-input initialiser, ``gpras/gpr.py:313-315``).  ``tol`` is scikit-learn's
    RELATIVE tolerance (scaled by the mean feature variance, ``_tolerance``).  Returns (centres, labels, inertia, n_iter)."""
    lib = _lib.load()
    if lib.gpras_device_count() <= 0:
        raise _lib.GprasError("no CUDA device visible: gpras_b200 has no CPU fallback")
    x = _f64(x)
    c = _f64(centers0).copy()
    if x.ndim != 2 or c.ndim != 2 or c.shape[1] != x.shape[1]:
        raise ValueError("x must be (N, D) and centers0 (M, D)")
    n, d = x.shape
    labels = np.empty(n, np.int32)
    inertia, n_iter = C.c_double(), C.c_int()
    tol_abs = float(np.mean(np.var(x, axis=0)) * tol)
    check(lib.gpras_kmeans_lloyd(int(device), ptr(x), n, d, ptr(c), c.shape[0], int(max_iter), tol_abs, labels.ctypes.data, C.byref(inertia),
                                 C.byref(n_iter)))
    return c, labels, inertia.value, n_iter.value
