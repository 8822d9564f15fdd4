"""CPU restatement of ``gpras/metrics.py`` as ONE pass of running reductions (NumPy, FP64).

Test infrastructure only (see ``oracle/__init__.py``).  PINNED: ``tests/golden/metrics_reference.npz`` holds the
outputs of the reference's own, unmodified ``gpras/metrics.py`` functions (generated in the build container by
``tests/golden/make_golden_reference.py``); ``tests/test_oracle.py`` checks ``summarise`` against every one of them.

The reference evaluates each metric as a separate whole-array expression over the (timesteps x cells) truth ``x``,
prediction ``y`` and confidence ``conf`` of one event (``metrics.py:85-324``).  Every one of them is a function of a few
running sums / maxima, which is what the streaming device kernel accumulates:

    per cell      sum(x-y), sum((x-y)^2), sum(conf), max_t x, max_t y           (x[argmax x] == max x)
    per timestep  sum_c(x-y), sum_c((x-y)^2), sum_c(conf)
    scalars       sum|x-y|, count(|x-y| <= v_tol)

``fi_aoi_toi`` with ``t_tol > 0`` compares shifted rows and is not a pure running reduction; ``summarise`` supports
``t_tol == 0`` (what ``export_metric_summary`` uses by default, ``metrics.py:17-18``).
"""

from __future__ import annotations

import numpy as np


def summarise(x, y, conf, depth_threshold=0.5, v_tol=0.0):
    x, y, conf = (np.asarray(a, np.float64) for a in (x, y, conf))
    t, c = x.shape
    e = x - y
    cell_e, cell_e2, cell_conf = e.sum(axis=0), (e * e).sum(axis=0), conf.sum(axis=0)
    xm, ym = x.max(axis=0), y.max(axis=0)
    ts_e, ts_e2, ts_conf = e.sum(axis=1), (e * e).sum(axis=1), conf.sum(axis=1)
    out = {
        "rmse_cell_toi": np.sqrt(cell_e2 / t), "err_cell_toi": cell_e / t, "conf_cell_toi": cell_conf / t,
        "err_cell_mts": xm - ym,
        "rmse_aoi_ts": np.sqrt(ts_e2 / c), "err_aoi_ts": ts_e / c, "conf_aoi_ts": ts_conf / c,
        "rmse_aoi_toi": float(np.sqrt(cell_e2.sum() / (t * c))), "mae_aoi_toi": float(np.abs(e).sum() / (t * c)),
        "conf_aoi_toi": float(cell_conf.sum() / (t * c)), "err_aoi_toi": float(cell_e.sum() / (t * c)),
        "fi_aoi_toi": float((np.abs(e) <= v_tol).sum() / (t * c)),
    }
    out.update(peak_scores(xm, ym, depth_threshold))
    return out


def peak_scores(xm, ym, depth_threshold=0.5):
    """Scalars of the per-cell peaks (``*_mts`` metrics, ``metrics.py:104-324``)."""
    d = xm - ym
    hx, hy = xm >= depth_threshold, ym >= depth_threshold
    a, miss, fa = float((hx & hy).sum()), float((hx & ~hy).sum()), float((~hx & hy).sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        pod = np.float64(a) / (a + miss)
        rfa = np.float64(fa) / (a + fa)
        csi = 1.0 / ((1.0 / pod) + (1.0 / (1.0 - rfa)) - 1.0)
    # f2 / f3 at threshold 0 (their default, metrics.py:274,301)
    h0x, h0y = xm >= 0.0, ym >= 0.0
    a0, b0, c0 = float((h0x & h0y).sum()), float((~h0x & h0y).sum()), float((h0x & ~h0y).sum())
    den = a0 + b0 + c0
    return {
        "rmse_aoi_mts": float(np.sqrt((d * d).mean())), "err_aoi_mts": float(d.mean()),
        "nse_aoi_mts": float(1.0 - (d * d).sum() / ((xm - xm.mean()) ** 2).sum()),
        "pod_mts": float(pod), "rfa_mts": float(rfa), "csi_mts": float(csi),
        "f2_mts": 1.0 if den == 0 else (a0 - c0) / den, "f3_mts": 1.0 if den == 0 else (a0 - b0) / den,
    }
