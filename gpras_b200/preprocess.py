"""Cells <-> modes transforms of the surrogate pipeline -- device mirror of ``gpras.preprocess.PreProcessor``.

Same constructor, attributes and methods as the reference class (``gpras/preprocess.py:866-1162``): ``fit``,
``transform``, ``reverse_transform``, ``wse_2_depth``, ``dry_indices``, ``eof``, ``classify_wetness_*``, ``to_dict`` /
``to_file`` / ``from_file``, and the module-level ``compute_norths_rule``.  The arithmetic runs on the GPU through the C ABI
(``gpras_pre_*``, ``include/gpras_b200.h``):

* ``fit``: one pass for the column max / min / mean (wetness classes, ``input_mean``), the centred + weighted samples, their
  Gram matrix on the FP64 tensor pipe, its leading eigenpairs by blocked subspace iteration (scikit-learn's
  ``IncrementalPCA`` in the reference computes the same principal axes by SVD), EOFs, score statistics;
* ``transform``: a long-k skinny GEMM with centring / weighting / depth clamp fused into the operand load;
* ``reverse_transform``: the folded modes -> cells map as two GEMMs with the offset fused into the epilogue.

Differences from the reference, all documented limits rather than approximations: at most 64 spatial modes; ``eigenvalues``
holds the leading ``min(128, samples, cells)`` explained variances (the reference keeps all ``min(samples, cells)``), of which
the retained ones (all 64 candidates when North's rule chooses) are converged to the residual tolerance and the rest are the
Ritz values of the last subspace iteration (lower bounds, typically within a percent); North's rule is evaluated on the
converged ones.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import pickle
from typing import Any

import numpy as np

from . import _lib
from ._lib import check, ptr

_HP = {"wse": 0, "depth": 1, "velocity": 2}
_CLASS_NAMES = np.array(["", "AD", "TF", "AF"])


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def compute_norths_rule(eigenvalues, n_samples: int | None = None) -> int:
    """North's rule of thumb on explained variances (``gpras/preprocess.py:1323-1353``).  Accepts either a fitted object
    exposing ``explained_variance_`` and ``n_samples_`` / ``n_samples_seen_`` (like the reference) or the two values."""
    if n_samples is None:
        pca = eigenvalues
        n_samples = getattr(pca, "n_samples_", None) or getattr(pca, "n_samples_seen_", None)
        if n_samples is None:
            return 0
        eigenvalues = pca.explained_variance_
    ev = np.asarray(eigenvalues, np.float64)
    ev = ev[ev > 1]
    if len(ev) == 0:
        return 0
    d_eigen = np.abs(np.diff(ev))
    d_error = np.sqrt(2 / n_samples) * ev[:-1]
    ind = np.argmax(d_eigen <= d_error) if d_eigen.size else 0
    return int(len(ev)) if ind == 0 else int(ind)


class PreProcessor:
    """Transform HEC-RAS cell data to / from a few spatial modes (``gpras/preprocess.py:866``)."""

    def __init__(self, spatial_mode_count: int = 0, input_mean=None, wet_threshold: float = 0.03, elevations=None,
                 hydraulic_parameter: str = "wse", wetness_classes=None, weights=None, eofs=None, eigenvalues=None,
                 n_samples_fit: float = 0, x_mean=None, x_std=None, device: int = 0):
        self.spatial_mode_count = spatial_mode_count
        self.input_mean = input_mean if input_mean is not None else np.empty(0, dtype=float)
        self.wet_threshold = wet_threshold
        self.elevations = elevations if elevations is not None else np.empty(0, dtype=float)
        self.hydraulic_parameter = hydraulic_parameter
        self.wetness_classes = wetness_classes if wetness_classes is not None else np.empty(0, dtype=np.str_)
        self.weights = weights if weights is not None else np.empty(0, dtype=float)
        self.eofs = eofs if eofs is not None else np.empty(0, dtype=float)
        self.eigenvalues = eigenvalues if eigenvalues is not None else np.empty(0, dtype=float)
        self.n_samples_fit = n_samples_fit
        self.x_mean = x_mean if x_mean is not None else np.empty(0, dtype=float)
        self.x_std = x_std if x_std is not None else np.empty(0, dtype=float)
        self.device = int(device)
        self._h = None
        self.fit_info: dict[str, Any] = {}

    # ---- reference properties ---------------------------------------------------------------------------------------
    @property
    def dry_indices(self):
        if self.wetness_classes is None:
            raise ValueError("wetness_classes must be numpy array to access dry_indices")
        return np.equal(self.wetness_classes, "AD")

    @property
    def eof(self):
        if self.eofs is None:
            raise ValueError("EOFs have not been computed")
        return self.eofs

    # ---- device handle ----------------------------------------------------------------------------------------------
    def _handle(self, cells: int):
        lib = _lib.load()
        if lib.gpras_device_count() <= 0:
            raise _lib.GprasError("no CUDA device visible: gpras_b200 has no CPU fallback")
        if self._h is None or self._cells != cells:
            self.close()
            h = C.c_void_p()
            check(lib.gpras_pre_create(C.byref(h), self.device, int(cells), _HP[self.hydraulic_parameter], float(self.wet_threshold)))
            self._h, self._cells, self._state_loaded = h, int(cells), False
        return lib

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.load().gpras_pre_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ensure_state(self):
        """Push host-side state (constructor arguments / ``from_file``) to the device handle."""
        cells = int(np.asarray(self.wetness_classes).shape[0])
        if cells == 0:
            raise ValueError("PreProcessor has not been fitted")
        lib = self._handle(cells)
        if self._state_loaded:
            return lib
        dry = np.ascontiguousarray(self.dry_indices.astype(np.uint8))
        wet = ~self.dry_indices
        p = int(self.spatial_mode_count)
        w = _scatter(np.asarray(self.weights, np.float64), wet, cells, 0.0) if np.asarray(self.weights).size else wet.astype(np.float64)
        eofs = np.zeros((p, cells))
        eofs[:, wet] = np.asarray(self.eofs, np.float64)[:p]
        # keep every temporary alive across the call: ptr() hands out raw addresses
        bufs = [_scatter(np.asarray(self.input_mean, np.float64), wet, cells, 0.0), _f64(w), _f64(eofs), _f64(self.x_mean),
                _f64(self.x_std), _f64(self.elevations)]
        check(lib.gpras_pre_set_state(self._h, dry.ctypes.data, *[ptr(b) for b in bufs], p))
        self._state_loaded = True
        return lib

    # ---- fit / transform / reverse ----------------------------------------------------------------------------------
    def fit(self, x, elevations, weights=None, spatial_mode_count: int | None = None, tol: float = 1e-12, max_iter: int = 200) -> None:
        """Fit the preprocessor (``gpras/preprocess.py:947-1007``).  ``x``: (samples, cells) NumPy array or CUDA float64
        torch tensor."""
        dev = not isinstance(x, np.ndarray) and hasattr(x, "data_ptr")
        if not dev:
            x = _f64(x)
        n, cells = int(x.shape[0]), int(x.shape[1])
        self.elevations = elevations
        elev = _f64(elevations)
        if weights is None:
            # the reference leaves self.weights empty here and then fails in transform(); unit weights are the evident intent
            w_in = np.ones(cells)
        else:
            w_in = _f64(weights)
        lib = self._handle(cells)
        ldx = int(x.stride(0)) if dev else cells
        modes = -1 if spatial_mode_count is None else int(spatial_mode_count)
        check(lib.gpras_pre_fit(self._h, ptr(x), ldx, n, int(dev), ptr(elev), ptr(w_in), modes, float(tol), int(max_iter)))
        cls = np.empty(cells)
        check(lib.gpras_pre_get(self._h, 0, ptr(cls)))
        self.wetness_classes = _CLASS_NAMES[cls.astype(int)]
        wet = ~self.dry_indices
        mean = np.empty(cells)
        check(lib.gpras_pre_get(self._h, 1, ptr(mean)))
        self.input_mean = mean[wet]
        if weights is not None:
            self.weights = np.asarray(weights)[wet]
        n_eig = lib.gpras_pre_eigen_count(self._h)
        ev = np.empty(n_eig)
        check(lib.gpras_pre_get(self._h, 4, ptr(ev)))
        self.eigenvalues = ev
        self.n_samples_fit = n
        if spatial_mode_count is None:
            self.spatial_mode_count = min(compute_norths_rule(ev, n), lib.gpras_pre_modes(self._h))
        else:
            self.spatial_mode_count = int(spatial_mode_count)
        check(lib.gpras_pre_set_modes(self._h, int(self.spatial_mode_count)))
        p = int(self.spatial_mode_count)
        eofs = np.empty((p, cells))
        xm, xs, res = np.empty(p), np.empty(p), np.empty(n_eig)
        if p > 0:
            check(lib.gpras_pre_get(self._h, 3, ptr(eofs)))
            check(lib.gpras_pre_get(self._h, 5, ptr(xm)))
            check(lib.gpras_pre_get(self._h, 6, ptr(xs)))
        check(lib.gpras_pre_get(self._h, 7, ptr(res)))
        self.eofs = np.ascontiguousarray(eofs[:, wet])
        self.x_mean, self.x_std = xm, xs
        ms = np.zeros(7)
        check(lib.gpras_pre_last_stage_ms(self._h, ptr(ms)))
        self.fit_info = {
            "iterations": int(lib.gpras_pre_iterations(self._h)), "residuals": res, "launches": int(lib.gpras_pre_last_launches(self._h)),
            "stage_ms": dict(zip(["colstats", "centre", "gram", "eigen", "eofs", "scores", "total"], ms.tolist())),
        }
        self._state_loaded = True
        if p > 0 and not np.all(res[:p] <= 100 * tol):
            raise RuntimeError("PCA fit: retained eigenpairs did not converge")

    def transform(self, x):
        """Project onto the retained EOFs and standardise (``gpras/preprocess.py:1009-1039``)."""
        lib = self._ensure_state()
        dev = not isinstance(x, np.ndarray) and hasattr(x, "data_ptr")
        if not dev:
            x = _f64(x)
        n, cells = int(x.shape[0]), int(x.shape[1])
        if cells != self._cells:
            raise ValueError(f"expected {self._cells} cells, got {cells}")
        p = int(self.spatial_mode_count)
        if dev:
            import torch

            z = torch.empty((n, p), dtype=torch.float64, device=x.device)
        else:
            z = np.empty((n, p))
        check(lib.gpras_pre_transform(self._h, ptr(x), int(x.stride(0)) if dev else cells, n, int(dev), ptr(z)))
        return z

    def wse_2_depth(self, x):
        """Convert water surface elevation data to depths (``gpras/preprocess.py:1041-1045``): host-side helper on small
        arrays; the device paths apply the same clamp inside their kernels."""
        d = x - self.elevations
        d[d < 0] = 0
        return d

    def reverse_transform(self, mean, var=None):
        """Back to cell space (``gpras/preprocess.py:1052-1084``): ``x_full`` or ``(x_full, var_prop_full)``."""
        lib = self._ensure_state()
        mean = _f64(mean)
        t, p = mean.shape
        if p != int(self.spatial_mode_count):
            raise ValueError(f"expected {self.spatial_mode_count} modes, got {p}")
        out_m = np.empty((t, self._cells))
        out_v = None
        if var is not None:
            var = _f64(var)
            out_v = np.empty((t, self._cells))
        check(lib.gpras_pre_reverse(self._h, ptr(mean), ptr(var) if var is not None else None, int(t), ptr(out_m),
                                    ptr(out_v) if var is not None else None))
        return out_m if var is None else (out_m, out_v)

    def cell_pitch(self) -> int:
        lib = self._ensure_state()
        return int(lib.gpras_pre_cell_pitch(self._h))

    def reverse_transform_device(self, mean, var, cell_mean=None, cell_var=None) -> None:
        """``reverse_transform`` with the (T x cells) mean and variance left on the GPU: ``cell_mean`` / ``cell_var`` are CUDA
        float64 torch tensors of shape (>= round_up(T, 64), cell_pitch()), or both None (the tiles stream through an internal
        ring buffer -- throughput runs of sweeps whose output cannot be kept).  ``mean`` / ``var`` (T x modes): NumPy arrays or
        CUDA tensors; ``var`` holds one variance PER MODE (the reference's per-column models, ``gpras/gpr.py:293-308``), mapped
        by ``var @ (diag(x_std) eofs / w)^2`` (``gpras/preprocess.py:1081-1094``) in the same fused kernel as the mean."""
        lib = self._ensure_state()
        dev = not isinstance(mean, np.ndarray) and hasattr(mean, "data_ptr")
        if not dev:
            mean, var = _f64(mean), _f64(var)
        t, p = int(mean.shape[0]), int(mean.shape[1])
        if p != int(self.spatial_mode_count) or tuple(var.shape) != (t, p):
            raise ValueError(f"expected (T, {self.spatial_mode_count}) means and variances")
        if dev and (not mean.is_contiguous() or not var.is_contiguous()):
            raise ValueError("device inputs must be contiguous")
        ldc = 0
        if cell_mean is not None:
            rows = (t + 63) // 64 * 64
            if cell_var is None or tuple(cell_mean.shape) != tuple(cell_var.shape) or cell_mean.shape[0] < rows:
                raise ValueError(f"cell_mean / cell_var must both be given with at least {rows} rows")
            ldc = int(cell_mean.stride(0))
        check(lib.gpras_pre_reverse_device(self._h, ptr(mean), ptr(var), t, int(dev), ptr(cell_mean) if cell_mean is not None else None,
                                           ptr(cell_var) if cell_var is not None else None, ldc))

    # ---- wetness helpers (host-side restatements used by callers on small arrays) ------------------------------------
    def classify_wetness_wse(self, x, elevations):
        return self._classify_depths(x.max(axis=0) - elevations, x.min(axis=0) - elevations)

    def classify_wetness_depth(self, x):
        return self._classify_depths(x.max(axis=0), x.min(axis=0))

    def _classify_depths(self, max_depth, min_depth):
        classes = np.empty(max_depth.shape, dtype="<U2")
        classes[max_depth < self.wet_threshold] = "AD"
        classes[max_depth > self.wet_threshold] = "TF"
        classes[min_depth > self.wet_threshold] = "AF"
        return classes

    # ---- persistence (same keys as the reference, gpras/preprocess.py:1134-1161) -------------------------------------
    def to_dict(self) -> dict[str, Any]:
        return {
            "spatial_mode_count": self.spatial_mode_count, "wet_threshold": self.wet_threshold,
            "hydraulic_parameter": self.hydraulic_parameter, "elevations": self.elevations,
            "wetness_classes": self.wetness_classes, "input_mean": self.input_mean, "weights": self.weights, "eofs": self.eofs,
            "eigenvalues": self.eigenvalues, "n_samples_fit": self.n_samples_fit, "x_mean": self.x_mean, "x_std": self.x_std,
        }

    def to_file(self, out_path) -> None:
        with open(out_path, mode="wb") as f:
            pickle.dump(self.to_dict(), f)

    @classmethod
    def from_file(cls, in_path):
        with open(in_path, mode="rb") as f:
            d = pickle.load(f)
        return cls(**d)


def _scatter(v, wet, cells, fill):
    out = np.full(cells, fill, np.float64)
    out[wet] = v
    return np.ascontiguousarray(out)
