"""Golden vectors produced by the UNMODIFIED reference code (fema-ffrd/gpras) in the build container.

Run once, here (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden_reference.py

Two reference modules carry NumPy-only arithmetic and can be executed offline:

* ``gpras/metrics.py`` imports cleanly;
* ``gpras/preprocess.py`` imports once the I/O libraries it pulls in at module level (geopandas, rasterio,
  hecdss, gpflow, tensorflow, matplotlib ...; none are touched by ``PreProcessor.fit / transform /
  reverse_transform / wse_2_depth`` or ``compute_norths_rule``) are replaced by inert stub modules.  The class itself
  is the reference's own code: ``PreProcessor`` (``gpras/preprocess.py:866-1162``) on NumPy + scikit-learn
  ``IncrementalPCA``.

Outputs (committed): ``preprocess_reference.npz`` and ``metrics_reference.npz`` -- inputs and the reference's outputs for
``fit`` (wetness classes, input mean, EOFs, eigenvalues, mode statistics, North's rule), ``transform``,
``reverse_transform`` (mean, and mean + variance) and every function of ``metrics.py``.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE = "/root/reference"


def import_reference(module: str):
    """Import a reference module, stubbing absent third-party (never ``gpras.*``) imports."""
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    stubbed = []
    for _ in range(200):
        try:
            return importlib.import_module(module), stubbed
        except ModuleNotFoundError as e:
            if e.name is None or e.name.startswith("gpras"):
                raise
            stub = MagicMock(name=e.name)
            stub.__path__, stub.__name__, stub.__spec__ = [], e.name, None
            sys.modules[e.name] = stub
            stubbed.append(e.name)
    raise RuntimeError("too many missing modules")


def flood_samples(n, c, k, seed, noise=0.02):
    """Synthetic water-surface samples (n x c): terrain + a few smooth spatial modes with decaying amplitudes."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0.0, 1.0, c)
    elevations = 5.0 + 3.0 * np.sin(7.0 * s) + 2.0 * s + 0.3 * rng.standard_normal(c)
    modes = np.stack([np.cos((j + 1) * np.pi * s + rng.uniform(0, 1)) for j in range(k)])
    amp = 2.0 * 0.55 ** np.arange(k)
    coef = rng.standard_normal((n, k)) * amp
    stage = 6.5 + coef @ modes + noise * rng.standard_normal((n, c))
    wse = np.maximum(stage, elevations)  # dry cells report the terrain elevation
    weights = rng.uniform(0.5, 2.0, c)
    return wse, elevations, weights


PRE_CASES = [
    # name, hydraulic_parameter, n, c, latent k, spatial_mode_count (None = North's rule), seed, T
    ("wse_fixed", "wse", 40, 300, 6, 5, 21, 25),
    ("depth_fixed", "depth", 64, 257, 5, 4, 22, 19),
    ("wse_north", "wse", 96, 200, 4, None, 23, 11),
    ("velocity", "velocity", 33, 129, 5, 3, 24, 7),
]


def make_preprocess():
    pre, stubbed = import_reference("gpras.preprocess")
    out = {"stubbed_modules": np.array(",".join(stubbed))}
    for name, hp, n, c, k, modes, seed, t in PRE_CASES:
        wse, elev, weights = flood_samples(n, c, k, seed)
        x_in = wse if hp != "velocity" else np.abs(wse - elev) + 0.1
        pp = pre.PreProcessor(wet_threshold=0.03, hydraulic_parameter=hp)
        pp.fit(x_in.copy(), elev.copy(), weights.copy(), modes)
        rng = np.random.default_rng(seed + 100)
        x_new = x_in[rng.permutation(n)[: max(3, n // 4)]] + 0.01 * rng.standard_normal((max(3, n // 4), c))
        z = pp.transform(x_new.copy())
        p = pp.spatial_mode_count
        m = rng.standard_normal((t, p))
        v = rng.uniform(0.05, 0.5, (t, 1)) * np.ones((1, p))  # every mode shares one variance per event (shared theta)
        v_free = rng.uniform(0.05, 0.5, (t, p))                # per-mode variances (per-column models)
        back_only = pp.reverse_transform(m.copy())
        back_m, back_v = pp.reverse_transform(m.copy(), v.copy())
        _, back_v_free = pp.reverse_transform(m.copy(), v_free.copy())
        depth_of_back = pp.wse_2_depth(back_m.copy())
        vals = dict(
            x=x_in, elevations=elev, weights=weights, modes_requested=-1 if modes is None else modes, hydraulic_parameter=np.array(hp),
            wetness_classes=pp.wetness_classes, dry_indices=pp.dry_indices, input_mean=pp.input_mean, fit_weights=pp.weights,
            eofs=pp.eofs, eigenvalues=pp.eigenvalues, n_samples_fit=pp.n_samples_fit, x_mean=pp.x_mean, x_std=pp.x_std,
            spatial_mode_count=pp.spatial_mode_count, x_new=x_new, transformed=z, mode_mean=m, mode_var=v, mode_var_free=v_free,
            reverse_mean_only=back_only, reverse_mean=back_m, reverse_var=back_v, reverse_var_free=back_v_free, reverse_depth=depth_of_back,
        )
        for key, val in vals.items():
            out[f"{name}.{key}"] = np.asarray(val)
        print(name, "modes", pp.spatial_mode_count, "dry", int(pp.dry_indices.sum()), "of", c)
    np.savez_compressed(HERE / "preprocess_reference.npz", **out)
    print("wrote preprocess_reference.npz; stubbed:", stubbed)


METRIC_CASES = [("small", 30, 50, 31), ("ragged", 17, 131, 32), ("one_step", 1, 40, 33), ("peaks_differ", 9, 23, 34)]


def make_metrics():
    met, _ = import_reference("gpras.metrics")
    out = {}
    for name, t, c, seed in METRIC_CASES:
        rng = np.random.default_rng(seed)
        x = np.maximum(rng.standard_normal((t, c)) + 0.4, 0.0)               # "truth" depths, many exact zeros
        y = np.maximum(x + 0.3 * rng.standard_normal((t, c)), 0.0)           # predicted depths
        if name == "peaks_differ":
            # prediction lags the truth by a few timesteps, so the arg-max rows of x and y differ in most cells, and depths of
            # 0..8 straddle the per-cell thresholds argmax(x) in 0..8 that export_metric_summary hands to f2 / f3
            x = np.maximum(4.0 * rng.random((t, c)) ** 2 + 4.0 * (np.arange(t)[:, None] == rng.integers(0, t, c)[None, :]), 0.0)
            y = np.maximum(np.roll(x, 2, axis=0) * rng.uniform(0.6, 1.4, (t, c)), 0.0)
        conf = rng.uniform(0.01, 0.4, (t, c))
        thr, t_tol, v_tol = 0.5, min(2, t - 1), 0.1
        x_mts, y_mts = np.argmax(x, axis=0), np.argmax(y, axis=0)
        vals = dict(
            x=x, y=y, conf=conf, depth_threshold=thr, t_tol=t_tol, v_tol=v_tol,
            rmse_aoi_toi=met.rmse_aoi_toi(x, y), mae_aoi_toi=met.mae_aoi_toi(x, y), conf_aoi_toi=met.conf_aoi_toi(conf),
            rmse_aoi_mts=met.rmse_aoi_mts(x, y, x_mts, y_mts), nse_aoi_mts=met.nse_aoi_mts(x, y, x_mts, y_mts),
            err_aoi_toi=met.err_aoi_toi(x, y), err_aoi_mts=met.err_aoi_mts(x, y, x_mts, y_mts),
            fi_aoi_toi=met.fi_aoi_toi(x, y, t_tol, v_tol), fi_aoi_toi_0=met.fi_aoi_toi(x, y, 0, 0),
            pod_mts=met.pod_mts(x, y, thr, x_mts, y_mts), rfa_mts=met.rfa_mts(x, y, thr, x_mts, y_mts),
            csi_mts=met.csi_mts(x, y, thr, x_mts, y_mts),
            # export_metric_summary passes x_mts POSITIONALLY into depth_threshold for f2 / f3 (metrics.py:53-54): both forms
            f2_mts=met.f2_mts(x, y, 0, x_mts, y_mts), f3_mts=met.f3_mts(x, y, 0, x_mts, y_mts),
            f2_mts_as_called=met.f2_mts(x, y, x_mts, y_mts), f3_mts_as_called=met.f3_mts(x, y, x_mts, y_mts),
            rmse_aoi_ts=met.rmse_aoi_ts(x, y), err_aoi_ts=met.err_aoi_ts(x, y), conf_aoi_ts=met.conf_aoi_ts(conf),
            rmse_cell_toi=met.rmse_cell_toi(x, y), err_cell_mts=met.err_cell_mts(x, y, x_mts, y_mts),
            err_cell_toi=met.err_cell_toi(x, y), conf_cell_toi=met.conf_cell_toi(conf), x_mts=x_mts, y_mts=y_mts,
        )
        # the *_mts functions with caller-supplied rows that are NOT the arg-max rows (drawn after everything above, so the
        # earlier cases keep their values)
        xr, yr = rng.integers(0, t, c), rng.integers(0, t, c)
        with np.errstate(all="ignore"):
            vals.update(
                rows_x=xr, rows_y=yr, rows_rmse_aoi_mts=met.rmse_aoi_mts(x, y, xr, yr), rows_nse_aoi_mts=met.nse_aoi_mts(x, y, xr, yr),
                rows_err_aoi_mts=met.err_aoi_mts(x, y, xr, yr), rows_err_cell_mts=met.err_cell_mts(x, y, xr, yr),
                rows_pod_mts=met.pod_mts(x, y, thr, xr, yr), rows_rfa_mts=met.rfa_mts(x, y, thr, xr, yr),
                rows_csi_mts=met.csi_mts(x, y, thr, xr, yr), rows_f2_mts=met.f2_mts(x, y, thr, xr, yr),
                rows_f3_mts=met.f3_mts(x, y, thr, xr, yr),
            )
        for key, val in vals.items():
            out[f"{name}.{key}"] = np.asarray(val)
    # the reference's export_metric_summary itself, on two events whose truth / prediction peaks fall on different rows
    import sqlite3
    import tempfile

    import pandas as pd

    rng = np.random.default_rng(35)
    t, c = 12, 37
    idx = pd.MultiIndex.from_product([["e1", "e2"], range(t)], names=["event", "t"])
    cols = [f"c{i}" for i in range(c)]
    xv = 3.0 * rng.random((2 * t, c)) ** 2
    yv = np.maximum(np.roll(xv, 1, axis=0) * rng.uniform(0.5, 1.5, (2 * t, c)), 0.0)
    cv = rng.uniform(0.1, 0.3, (2 * t, c))
    with tempfile.TemporaryDirectory() as tmp:
        db = Path(tmp) / "m.db"
        met.export_metric_summary(pd.DataFrame(xv, index=idx, columns=cols), pd.DataFrame(yv, index=idx, columns=cols),
                                  pd.DataFrame(cv, index=idx, columns=cols), db, depth_threshold=0.5, t_tol=1, v_tol=0.05)
        with sqlite3.connect(db) as con:
            sc = pd.read_sql("select * from scalar_metrics", con)
            ts = pd.read_sql("select * from timeseries_metrics", con)
            ce = pd.read_sql("select * from cell_metrics", con)
    out["export.x"], out["export.y"], out["export.conf"] = xv, yv, cv
    out["export.scalar_columns"] = np.array(list(sc.columns))
    out["export.scalar"] = sc.drop(columns=["event"]).to_numpy(np.float64)
    out["export.timeseries"] = ts.drop(columns=["event", "timestep"]).to_numpy(np.float64)
    out["export.cells"] = ce.drop(columns=["event", "cell_id"]).to_numpy(np.float64)
    np.savez_compressed(HERE / "metrics_reference.npz", **out)
    print("wrote metrics_reference.npz")


def make_edge_cases():
    """CPU-only extras: North's rule on many spectra, and metrics on degenerate events (all dry, ties, a single cell)."""
    pre, _ = import_reference("gpras.preprocess")
    met, _ = import_reference("gpras.metrics")
    from sklearn.decomposition import PCA

    rng = np.random.default_rng(99)
    out = {}
    evs, ns, expect = [], [], []
    for i in range(120):
        k = int(rng.integers(1, 40))
        ev = np.sort(rng.gamma(0.7, 6.0, k) + (rng.random() < 0.3) * 1.0)[::-1]
        n = int(rng.integers(5, 400))
        fake = PCA()
        fake.n_samples_, fake.explained_variance_ = n, ev
        pad = np.full(40, np.nan)
        pad[:k] = ev
        try:
            r = pre.compute_norths_rule(fake)
        except ValueError:  # the reference takes argmax of an empty array when exactly one eigenvalue exceeds 1
            r = -1
        evs.append(pad), ns.append(n), expect.append(r)
    out["north.eigenvalues"], out["north.n"], out["north.expected"] = np.array(evs), np.array(ns), np.array(expect)
    cases = {
        "all_dry": (np.zeros((7, 9)), np.zeros((7, 9)), np.zeros((7, 9))),
        "ties": (np.tile(np.array([[0.0, 1.0, 1.0, 2.0]]), (5, 1)), np.tile(np.array([[0.5, 1.0, 0.0, 2.0]]), (5, 1)), np.full((5, 4), 0.1)),
        "one_cell": (rng.random((11, 1)) * 2, rng.random((11, 1)) * 2, rng.random((11, 1))),
        "never_detected": (np.full((4, 6), 0.1), np.full((4, 6), 0.2), np.full((4, 6), 0.05)),
    }
    for name, (x, y, conf) in cases.items():
        x_mts, y_mts = np.argmax(x, axis=0), np.argmax(y, axis=0)
        with np.errstate(all="ignore"):
            vals = dict(
                x=x, y=y, conf=conf, rmse_aoi_toi=met.rmse_aoi_toi(x, y), mae_aoi_toi=met.mae_aoi_toi(x, y),
                conf_aoi_toi=met.conf_aoi_toi(conf), rmse_aoi_mts=met.rmse_aoi_mts(x, y, x_mts, y_mts),
                nse_aoi_mts=met.nse_aoi_mts(x, y, x_mts, y_mts), err_aoi_toi=met.err_aoi_toi(x, y),
                err_aoi_mts=met.err_aoi_mts(x, y, x_mts, y_mts), fi_aoi_toi_0=met.fi_aoi_toi(x, y, 0, 0),
                pod_mts=met.pod_mts(x, y, 0.5, x_mts, y_mts), rfa_mts=met.rfa_mts(x, y, 0.5, x_mts, y_mts),
                csi_mts=met.csi_mts(x, y, 0.5, x_mts, y_mts), f2_mts=met.f2_mts(x, y, 0, x_mts, y_mts),
                f3_mts=met.f3_mts(x, y, 0, x_mts, y_mts), rmse_aoi_ts=met.rmse_aoi_ts(x, y), err_aoi_ts=met.err_aoi_ts(x, y),
                conf_aoi_ts=met.conf_aoi_ts(conf), rmse_cell_toi=met.rmse_cell_toi(x, y), err_cell_mts=met.err_cell_mts(x, y, x_mts, y_mts),
                err_cell_toi=met.err_cell_toi(x, y), conf_cell_toi=met.conf_cell_toi(conf),
            )
        for key, val in vals.items():
            out[f"{name}.{key}"] = np.asarray(val, dtype=np.float64)
    np.savez_compressed(HERE / "edge_cases_reference.npz", **out)
    print("wrote edge_cases_reference.npz")


if __name__ == "__main__":
    import sys

    which = sys.argv[1:] or ["preprocess", "metrics", "edge"]
    if "preprocess" in which:
        make_preprocess()
    if "metrics" in which:
        make_metrics()
    if "edge" in which:
        make_edge_cases()
