"""Accuracy metrics of a surrogate against the high-fidelity model -- device mirror of ``gpras/metrics.py``.

Same module-level names and signatures as the reference (``gpras/metrics.py:11-324``): ``rmse_aoi_toi``, ``mae_aoi_toi``,
``conf_aoi_toi``, ``rmse_aoi_ts``, ``rmse_cell_toi``, ``rmse_aoi_mts``, ``err_cell_mts``, ``nse_aoi_mts``,
``err_aoi_toi``, ``err_aoi_mts``, ``err_aoi_ts``, ``conf_aoi_ts``, ``err_cell_toi``, ``conf_cell_toi``, ``fi_aoi_toi``,
``pod_mts``, ``rfa_mts``, ``csi_mts``, ``f2_mts``, ``f3_mts`` and ``export_metric_summary``.

The reference evaluates every metric as its own whole-array NumPy expression; each is a closed form of a handful of
running reductions (per cell: sum, sum of squares, sum of confidence, max over time of truth and prediction; per
timestep: sums over cells), so ``MetricsAccumulator`` makes ONE pass on the GPU (``csrc/metrics_kernel.cuh``) and every
function below reads its value off that summary.  ``MetricsAccumulator.predict_update`` goes one step further and
consumes a conditioned model's predictions tile by tile, so the (timesteps x cells) prediction is never written to memory
(the 3.2 TB output of BASELINE config 5).

There is no CPU fallback: without the CUDA library / a device these functions raise.
"""

from __future__ import annotations

import ctypes as C
import sqlite3
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import check, ptr

_SCALARS = 15


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _is_device(a) -> bool:
    return a is not None and not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")


class Summary(dict):
    """Result of one accumulated event: raw reductions plus every metric of ``gpras/metrics.py`` (``t_tol == 0``)."""


class MetricsAccumulator:
    """Running reductions of one event (``timesteps x cells``) on one GPU."""

    def __init__(self, cells: int, capacity: int, device: int = 0):
        self.lib = _lib.load()
        if self.lib.gpras_device_count() <= 0:
            raise _lib.GprasError("no CUDA device visible: gpras_b200 has no CPU fallback")
        self.c, self.capacity, self.device = int(cells), int(capacity), int(device)
        h = C.c_void_p()
        check(self.lib.gpras_metrics_create(C.byref(h), self.device, self.c, self.capacity))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.gpras_metrics_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_elevations(self, elev_truth=None, elev_pred=None) -> None:
        """Convert water-surface elevations to depths on the fly (``PreProcessor.wse_2_depth``,
        ``gpras/preprocess.py:1041-1045``): ``truth <- max(truth - elev_truth, 0)``, same for the prediction."""
        et = None if elev_truth is None else _f64(elev_truth)
        ep = None if elev_pred is None else _f64(elev_pred)
        for e in (et, ep):
            if e is not None and e.shape != (self.c,):
                raise ValueError(f"expected {self.c} elevations, got {e.shape}")
        check(self.lib.gpras_metrics_set_elevations(self._h, ptr(et) if et is not None else None, ptr(ep) if ep is not None else None))

    def reset(self, v_tol: float = 0.0) -> None:
        check(self.lib.gpras_metrics_reset(self._h, float(v_tol)))

    def update(self, x, y, conf=None) -> None:
        """Accumulate timesteps given as ``(t, cells)`` arrays: NumPy (host) or CUDA float64 torch tensors (all alike)."""
        dev = _is_device(y)
        if not dev:
            x = None if x is None else _f64(x)
            y = _f64(y)
            conf = None if conf is None else _f64(conf)
        for a in (x, y, conf):
            if a is not None and (a.ndim != 2 or a.shape[1] != self.c or a.shape[0] != y.shape[0]):
                raise ValueError(f"expected (t, {self.c}) arrays of equal length")
        ld = (lambda a: int(a.stride(0)) if dev else a.shape[1])
        check(self.lib.gpras_metrics_update(
            self._h, ptr(x) if x is not None else None, ld(x) if x is not None else 0, ptr(y), ld(y),
            ptr(conf) if conf is not None else None, ld(conf) if conf is not None else 0, int(y.shape[0]), int(dev)))

    def predict_update(self, gp, xs, truth=None, want_modes: bool = False):
        """Predict the events ``xs`` with the conditioned ``ExactGP`` ``gp`` (cell map bound) and accumulate them against
        ``truth`` (``(t, cells)``, host or CUDA tensor, or None) without materialising the cell-space prediction."""
        xs_dev = _is_device(xs)
        if not xs_dev:
            xs = _f64(xs)
        t = int(xs.shape[0])
        tr_dev = _is_device(truth)
        if truth is not None and not tr_dev:
            truth = _f64(truth)
        if truth is not None and (truth.shape[0] != t or truth.shape[1] != self.c):
            raise ValueError(f"expected truth of shape ({t}, {self.c})")
        ldx = 0 if truth is None else (int(truth.stride(0)) if tr_dev else truth.shape[1])
        mm = np.empty((t, gp.p)) if want_modes else None
        mv = np.empty((t, gp.p)) if want_modes else None
        check(self.lib.gpras_gp_predict_metrics(
            gp._h, self._h, ptr(xs), t, int(xs_dev), ptr(truth) if truth is not None else None, ldx, int(tr_dev),
            ptr(mm) if want_modes else None, ptr(mv) if want_modes else None))
        return mm, mv

    def reverse_update(self, pp, mean, var, truth=None) -> None:
        """Accumulate events predicted in MODE space by per-column models (one variance per mode, ``gpras/gpr.py:293-308``):
        ``pp`` is the fitted / loaded ``PreProcessor`` mirror whose ``reverse_transform`` maps modes to cells; the cell-space
        prediction ``max(mean E + bias - elev, 0)`` and its confidence ``sqrt(var E^2)`` are formed tile by tile and consumed
        against ``truth`` (``(t, cells)``, host array, CUDA tensor or None) without ever being written
        (``pipeline.py:260-286`` fused; ``gpras_pre_reverse_metrics``)."""
        lib = pp._ensure_state()
        dev = _is_device(mean)
        if not dev:
            mean, var = _f64(mean), _f64(var)
        t, p = int(mean.shape[0]), int(mean.shape[1])
        if tuple(var.shape) != (t, p) or p != int(pp.spatial_mode_count):
            raise ValueError(f"expected (T, {pp.spatial_mode_count}) means and variances")
        tr_dev = _is_device(truth)
        if truth is not None and not tr_dev:
            truth = _f64(truth)
        if truth is not None and (truth.shape[0] != t or truth.shape[1] != self.c):
            raise ValueError(f"expected truth of shape ({t}, {self.c})")
        ldx = 0 if truth is None else (int(truth.stride(0)) if tr_dev else truth.shape[1])
        check(lib.gpras_pre_reverse_metrics(pp._h, self._h, ptr(mean), ptr(var), t, int(dev), ptr(truth) if truth is not None else None,
                                            ldx, int(tr_dev)))

    def timesteps(self) -> int:
        return int(self.lib.gpras_metrics_timesteps(self._h))

    def last_launches(self) -> int:
        return int(self.lib.gpras_metrics_last_launches(self._h))

    def finalize(self, depth_threshold: float = 0.5) -> Summary:
        t = self.timesteps()
        scal = np.zeros(_SCALARS)
        cells = np.zeros((5, self.c))
        rows = np.zeros((3, t))
        check(self.lib.gpras_metrics_finalize(self._h, float(depth_threshold), ptr(scal), ptr(cells), ptr(rows)))
        return _summary(scal, cells, rows, t, self.c)


def _summary(scal, cells, rows, t, c) -> Summary:
    n = float(t) * float(c)
    s = Summary()
    s["cell_sum_err"], s["cell_sum_sq"], s["cell_sum_conf"], s["cell_max_x"], s["cell_max_y"] = cells
    s["rmse_cell_toi"] = np.sqrt(cells[1] / t)
    s["err_cell_toi"] = cells[0] / t
    s["conf_cell_toi"] = cells[2] / t
    s["err_cell_mts"] = cells[3] - cells[4]
    s["rmse_aoi_ts"] = np.sqrt(rows[1] / c)
    s["err_aoi_ts"] = rows[0] / c
    s["conf_aoi_ts"] = rows[2] / c
    s["rmse_aoi_toi"] = float(np.sqrt(scal[1] / n))
    s["err_aoi_toi"] = float(scal[0] / n)
    s["conf_aoi_toi"] = float(scal[2] / n)
    s["mae_aoi_toi"] = float(scal[3] / n)
    s["fi_aoi_toi"] = float(scal[4] / n)
    s["rmse_aoi_mts"] = float(np.sqrt(scal[6] / c))
    s["err_aoi_mts"] = float(scal[5] / c)
    with np.errstate(divide="ignore", invalid="ignore"):
        s["nse_aoi_mts"] = float(1.0 - np.float64(scal[6]) / np.float64(scal[8]))
        a, miss, fa = (np.float64(v) for v in scal[9:12])
        pod, rfa = a / (a + miss), fa / (a + fa)
        s["pod_mts"], s["rfa_mts"] = float(pod), float(rfa)
        s["csi_mts"] = float(1.0 / ((1.0 / pod) + (1.0 / (1.0 - rfa)) - 1.0))
    a0, c0, b0 = scal[12], scal[13], scal[14]  # hits, misses (x >= 0 > y), false alarms at threshold 0
    den = a0 + b0 + c0
    s["f2_mts"] = 1.0 if den == 0 else float((a0 - c0) / den)
    s["f3_mts"] = 1.0 if den == 0 else float((a0 - b0) / den)
    return s


def summarise(x, y, conf=None, depth_threshold: float = 0.5, v_tol: float = 0.0, device: int = 0) -> Summary:
    """All metrics of one event in a single device pass over ``x`` (truth), ``y`` (prediction), ``conf``."""
    t, c = int(y.shape[0]), int(y.shape[1])
    acc = MetricsAccumulator(c, t, device)
    try:
        acc.reset(v_tol)
        acc.update(x, y, conf)
        return acc.finalize(depth_threshold)
    finally:
        acc.close()


def _gather_rows(a, rows) -> np.ndarray:
    """``a[rows, arange(cells)]`` on the host (``a``: NumPy array or CUDA tensor) -- O(cells) work."""
    rows = np.asarray(rows).astype(np.int64)
    if _is_device(a):
        import torch

        r = torch.as_tensor(rows, device=a.device)
        return a[r, torch.arange(a.shape[1], device=a.device)].double().cpu().numpy()
    a = np.asarray(a)
    return np.asarray(a[rows, np.arange(a.shape[1])], np.float64)


def _peak_metrics(xp, yp, depth_threshold) -> dict:
    """Every ``*_mts`` metric of ``gpras/metrics.py:111-318`` as a closed form of the two per-cell peak vectors
    ``xp = x[x_mts, cells]`` and ``yp = y[y_mts, cells]``; ``depth_threshold`` is a scalar or, in the reference's own
    call of f2 / f3 (``metrics.py:53-54``), a per-cell vector."""
    xp, yp = np.asarray(xp, np.float64), np.asarray(yp, np.float64)
    thr = np.asarray(depth_threshold, np.float64)
    diff = xp - yp
    out = {"err_cell_mts": diff, "rmse_aoi_mts": float(np.sqrt(np.mean(diff**2))), "err_aoi_mts": float(np.mean(diff))}
    a = np.float64(np.sum((xp >= thr) & (yp >= thr)))
    b = np.float64(np.sum((xp < thr) & (yp >= thr)))   # false alarms
    c = np.float64(np.sum((xp >= thr) & (yp < thr)))   # misses
    with np.errstate(divide="ignore", invalid="ignore"):
        out["nse_aoi_mts"] = float(1.0 - np.float64(np.sum(diff**2)) / np.float64(np.sum((xp - xp.mean()) ** 2)))
        pod, rfa = a / (a + c), b / (a + b)
        out["pod_mts"], out["rfa_mts"] = float(pod), float(rfa)
        out["csi_mts"] = float(1.0 / ((1.0 / pod) + (1.0 / (1.0 - rfa)) - 1.0))
    den = a + b + c
    out["f2_mts"] = 1 if den == 0 else float((a - c) / den)
    out["f3_mts"] = 1 if den == 0 else float((a - b) / den)
    return out


def _peaks(x, y, x_mts, y_mts, depth_threshold=0.5) -> dict:
    """Peak metrics.  With no rows supplied the peaks are the per-cell maxima of the device pass (``x[argmax x] == max x``);
    caller-supplied ``x_mts`` / ``y_mts`` (``metrics.py:35-36`` caches the arg-max rows, but any rows are legal) are honoured
    by gathering those rows."""
    if x_mts is None and y_mts is None and np.ndim(depth_threshold) == 0:
        return summarise(x, y, None, float(depth_threshold))
    s = summarise(x, y) if (x_mts is None or y_mts is None) else None
    xp = s["cell_max_x"] if x_mts is None else _gather_rows(x, x_mts)
    yp = s["cell_max_y"] if y_mts is None else _gather_rows(y, y_mts)
    return _peak_metrics(xp, yp, depth_threshold)


# ---- the reference's function set (gpras/metrics.py:85-324) -----------------------------------------------------------
def rmse_aoi_toi(x, y) -> float:
    return summarise(x, y)["rmse_aoi_toi"]


def mae_aoi_toi(x, y) -> float:
    return summarise(x, y)["mae_aoi_toi"]


def conf_aoi_toi(x) -> float:
    return summarise(None, x, x)["conf_aoi_toi"]


def rmse_aoi_ts(x, y):
    return summarise(x, y)["rmse_aoi_ts"]


def rmse_cell_toi(x, y):
    return summarise(x, y)["rmse_cell_toi"]


def rmse_aoi_mts(x, y, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts)["rmse_aoi_mts"]


def err_cell_mts(x, y, x_mts=None, y_mts=None):
    return _peaks(x, y, x_mts, y_mts)["err_cell_mts"]


def nse_aoi_mts(x, y, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts)["nse_aoi_mts"]


def err_aoi_toi(x, y) -> float:
    return summarise(x, y)["err_aoi_toi"]


def err_aoi_mts(x, y, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts)["err_aoi_mts"]


def err_aoi_ts(x, y):
    return summarise(x, y)["err_aoi_ts"]


def conf_aoi_ts(x):
    return summarise(None, x, x)["conf_aoi_ts"]


def err_cell_toi(x, y):
    return summarise(x, y)["err_cell_toi"]


def conf_cell_toi(x):
    return summarise(None, x, x)["conf_cell_toi"]


def fi_aoi_toi(x, y, t_tol: int, v_tol: float) -> float:
    """Fidelity index (``metrics.py:203-212``); the shifted comparisons of ``t_tol > 0`` run in their own kernel."""
    if int(t_tol) == 0:
        return summarise(x, y, None, 0.5, v_tol)["fi_aoi_toi"]
    lib = _lib.load()
    dev = _is_device(x)
    if not dev:
        x, y = _f64(x), _f64(y)
    t, c = int(x.shape[0]), int(x.shape[1])
    ld = (lambda a: int(a.stride(0)) if dev else a.shape[1])
    out = C.c_double()
    check(lib.gpras_metrics_fidelity(ptr(x), ld(x), ptr(y), ld(y), t, c, int(t_tol), float(v_tol), 0, C.byref(out)))
    return float(out.value / (t * c))


def pod_mts(x, y, depth_threshold: float = 0, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts, depth_threshold)["pod_mts"]


def rfa_mts(x, y, depth_threshold: float = 0, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts, depth_threshold)["rfa_mts"]


def csi_mts(x, y, depth_threshold: float = 0, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts, depth_threshold)["csi_mts"]


def f2_mts(x, y, depth_threshold=0, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts, depth_threshold)["f2_mts"]


def f3_mts(x, y, depth_threshold=0, x_mts=None, y_mts=None) -> float:
    return _peaks(x, y, x_mts, y_mts, depth_threshold)["f3_mts"]


def export_metric_summary(x_all, y_all, conf_all, out_path, depth_threshold: float = 0.5, t_tol: int = 0, v_tol: float = 0,
                          hydraulic_parameter: str = "depth") -> None:
    """Export all metrics to a sqlite database (``metrics.py:11-82``): same tables and columns, one device pass per event."""
    import pandas as pd

    all_scalar, all_timeseries, all_cells = [], [], []
    for event in x_all.index.unique(level=0):
        x = x_all.loc[event].values
        y = y_all.loc[event].values
        conf = conf_all.loc[event].values
        tsteps = x_all.loc[event].index.values
        s = summarise(x, y, conf, depth_threshold, v_tol)
        fi = s["fi_aoi_toi"] if t_tol == 0 else fi_aoi_toi(x, y, t_tol, v_tol)
        vel = hydraulic_parameter == "velocity"
        # The reference calls f2_mts(x, y, x_mts, y_mts) POSITIONALLY (metrics.py:53-54): the arg-max rows of x land in
        # `depth_threshold` and those of y in the `x_mts` parameter, so the "truth peak" is x at the row of y's maximum and the
        # threshold is the row index of x's maximum.  Reproduced as is (two O(T C) host arg-max passes, like the reference).
        x_rows, y_rows = np.argmax(x, axis=0), np.argmax(y, axis=0)
        f23 = _peak_metrics(_gather_rows(x, y_rows), s["cell_max_y"], x_rows)
        all_scalar.append(pd.DataFrame.from_dict({
            "event": event, "rmse_aoi_toi": [s["rmse_aoi_toi"]], "mae_aoi_toi": [s["mae_aoi_toi"]],
            "conf_aoi_toi": [s["conf_aoi_toi"]], "rmse_aoi_mts": [s["rmse_aoi_mts"]], "nse_aoi_mts": [s["nse_aoi_mts"]],
            "err_aoi_toi": [s["err_aoi_toi"]], "err_aoi_mts": [s["err_aoi_mts"]], "fi_aoi_toi": [fi],
            "pod_mts": [np.nan if vel else s["pod_mts"]], "rfa_mts": [np.nan if vel else s["rfa_mts"]],
            "csi_mts": [np.nan if vel else s["csi_mts"]],
            "f2_mts": [f23["f2_mts"]], "f3_mts": [f23["f3_mts"]],
        }))
        all_timeseries.append(pd.DataFrame.from_dict({
            "event": np.repeat(event, x.shape[0]), "timestep": tsteps, "rmse_aoi_ts": s["rmse_aoi_ts"],
            "err_aoi_ts": s["err_aoi_ts"], "conf_aoi_ts": s["conf_aoi_ts"],
        }))
        all_cells.append(pd.DataFrame.from_dict({
            "event": np.repeat(event, x.shape[1]), "cell_id": x_all.columns, "rmse_cell_toi": s["rmse_cell_toi"],
            "err_cell_mts": s["err_cell_mts"], "err_cell_toi": s["err_cell_toi"], "conf_cell_toi": s["conf_cell_toi"],
        }))
    with sqlite3.connect(str(Path(out_path))) as con:
        pd.concat(all_scalar).to_sql("scalar_metrics", con, index=False, if_exists="replace")
        pd.concat(all_timeseries).to_sql("timeseries_metrics", con, index=False, if_exists="replace")
        pd.concat(all_cells).to_sql("cell_metrics", con, index=False, if_exists="replace")
