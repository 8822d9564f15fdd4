"""The reference's analysis pipeline, steps ``production/analysis/pipeline.py:235-286``, run end to end on the drop-in classes:

    reducers.fit / transform  ->  GPRAS.fit (default: per-column sparse models, "two-stage")  ->  to_file / from_file
    ->  GPRAS.predict  ->  reverse_transform  ->  wse_2_depth  ->  export_metric_summary

and checked stage by stage against the CPU oracles on the same inputs (each stage's oracle is fed the DEVICE result of the stage
before it, so a tolerance never has to absorb the conditioning of an earlier stage).  Sizes are small; the sizes of BASELINE.json
are covered per component in the other ``-m gpu`` files.
"""

import sqlite3

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda(lib):
    import torch

    assert torch.cuda.is_available() and lib.gpras_device_count() > 0, "GPU tests need a CUDA device"
    return torch


def _flood(n_events, cells, seed, coarse=False):
    """Water-surface elevations of n_events events over `cells` cells from 5 smooth spatial patterns driven by 3 latent forcings;
    the low-fidelity model sees a biased, damped version of the same forcings."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0.0, 1.0, cells)
    elev = 5.0 + 3.0 * np.sin(7.0 * s) + 2.0 * s
    pats = np.stack([np.cos((j + 1) * np.pi * s + 0.3 * j) for j in range(5)])
    f = rng.standard_normal((n_events, 3))
    coef = np.stack([f[:, 0], f[:, 1], np.tanh(f[:, 0] * f[:, 2]), np.sin(f[:, 1]), 0.5 * f[:, 2] ** 2], axis=1) * (2.0 * 0.6 ** np.arange(5))
    hf = np.maximum(6.5 + coef @ pats + 0.01 * rng.standard_normal((n_events, cells)), elev)
    lf = np.maximum(6.3 + 0.85 * (coef[:, :3] @ pats[:3]) + 0.01 * rng.standard_normal((n_events, cells)), elev)
    w = rng.uniform(0.5, 2.0, cells)
    return hf, lf, elev, w


@pytest.mark.parametrize("hydraulic_parameter", ["wse", "depth"])
def test_reference_pipeline_on_the_drop_in_classes(cuda, tmp_path, hydraulic_parameter):
    import pandas as pd
    import torch

    from gpras_b200 import GPRAS
    from gpras_b200 import metrics as gm
    from gpras_b200.preprocess import PreProcessor
    from oracle import cells as ocells
    from oracle import metrics as ometrics
    from oracle import preprocess as opre
    from oracle import sgpr

    cells, modes, m_ind = 700, 4, 10
    hf, lf, elev, w = _flood(160, cells, seed=1)
    hf_test, lf_test, _, _ = _flood(24, cells, seed=2)
    # (the reducers are always fed water-surface elevations; with hydraulic_parameter="depth" they convert internally,
    #  gpras/preprocess.py:969-971,1021-1022)

    # ---- preprocess (pipeline.py:233-239) ----
    hf_red = PreProcessor(hydraulic_parameter=hydraulic_parameter)
    lf_red = PreProcessor(hydraulic_parameter=hydraulic_parameter)
    hf_red.fit(hf, elev, w, modes)
    lf_red.fit(lf, elev, w, modes)
    y, x = hf_red.transform(hf), lf_red.transform(lf)
    x_test = lf_red.transform(lf_test)
    o_hf = opre.fit(hf, elev, w, modes, hydraulic_parameter=hydraulic_parameter)
    o_lf = opre.fit(lf, elev, w, modes, hydraulic_parameter=hydraulic_parameter)
    oy, ox = opre.transform(o_hf, hf, elev, hydraulic_parameter), opre.transform(o_lf, lf, elev, hydraulic_parameter)
    sign = np.sign(np.sum(oy * y, axis=0))  # an EOF's sign is a convention; both sides follow scikit-learn's, check it holds
    assert np.all(sign == 1.0)
    np.testing.assert_allclose(y, oy, rtol=1e-8, atol=1e-9 * np.abs(oy).max())
    np.testing.assert_allclose(x, ox, rtol=1e-8, atol=1e-9 * np.abs(ox).max())

    # ---- fit (pipeline.py:241-256): the reference's defaults, persistence round trip ----
    gpr = GPRAS("Matern32")
    gpr.fit(x, y, m_ind, "kmeans", "two-stage", max_iter=30)
    path = tmp_path / "model.json"
    gpr.to_file(path)
    gpr = GPRAS.from_file(path)
    assert len(gpr.models) == modes and gpr.models[0].inducing_variable.Z.shape == (m_ind, modes)

    # ---- predict (pipeline.py:258-261) against the SGPR oracle at the fitted parameters ----
    mean_pred, var_pred = gpr.predict(x_test)
    assert mean_pred.shape == var_pred.shape == (24, modes)
    for j, mdl in enumerate(gpr.models):
        th = mdl.theta()
        om, ov = sgpr.predict_y("Matern32", x, y[:, j : j + 1], np.asarray(mdl.inducing_variable.Z), th[0], th[2:], th[1], x_test)
        np.testing.assert_allclose(mean_pred[:, j], om[:, 0], rtol=1e-8, atol=1e-9 * np.abs(om).max())
        np.testing.assert_allclose(var_pred[:, j], ov[:, 0], rtol=1e-8)

    # ---- modes -> cells (pipeline.py:261) ----
    y_pred, y_var = hf_red.reverse_transform(mean_pred, var_pred)
    o_pred, o_var = ocells.reverse_transform(mean_pred, var_pred, hf_red.eofs, hf_red.x_mean, hf_red.x_std, hf_red.weights,
                                             hf_red.input_mean, hf_red.dry_indices, np.asarray(elev),
                                             depth=hydraulic_parameter == "depth")
    np.testing.assert_allclose(y_pred, o_pred, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(y_var, o_var, rtol=1e-10, atol=1e-14)
    assert np.all(y_var >= 0.0)
    # the same numbers from the device-resident epilogue (results stay in HBM)
    c_pad = hf_red.cell_pitch()
    d_mean = torch.empty((64, c_pad), dtype=torch.float64, device="cuda")  # rows rounded up to the kernel's 64-event tile
    d_var = torch.empty((64, c_pad), dtype=torch.float64, device="cuda")
    hf_red.reverse_transform_device(torch.from_numpy(mean_pred).cuda(), torch.from_numpy(var_pred).cuda(), d_mean, d_var)
    np.testing.assert_allclose(d_mean[:24, :cells].cpu().numpy(), y_pred, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(d_var[:24, :cells].cpu().numpy(), y_var, rtol=1e-12, atol=1e-15)

    # ---- depths and metrics (pipeline.py:265-286) ----
    if hydraulic_parameter == "depth":
        y_pred = y_pred + hf_red.elevations
    truth_depth = hf_red.wse_2_depth(hf_test)
    pred_depth = hf_red.wse_2_depth(y_pred)
    np.testing.assert_array_equal(pred_depth, opre.wse_to_depth(y_pred, np.asarray(elev)))
    conf = np.sqrt(y_var)
    idx = pd.MultiIndex.from_product([["e1", "e2"], range(12)], names=["event", "t"])
    cols = [f"c{i}" for i in range(cells)]
    out = tmp_path / "metrics.db"
    gm.export_metric_summary(pd.DataFrame(truth_depth, index=idx, columns=cols), pd.DataFrame(pred_depth, index=idx, columns=cols),
                             pd.DataFrame(conf, index=idx, columns=cols), out)
    with sqlite3.connect(out) as con:
        sc = pd.read_sql("select * from scalar_metrics", con)
        ce = pd.read_sql("select * from cell_metrics", con)
    assert len(sc) == 2 and len(ce) == 2 * cells
    for k, ev in enumerate(("e1", "e2")):
        rows = slice(12 * k, 12 * k + 12)
        o = ometrics.summarise(truth_depth[rows], pred_depth[rows], conf[rows])
        row = sc[sc["event"] == ev].iloc[0]
        for name in ("rmse_aoi_toi", "mae_aoi_toi", "err_aoi_toi", "conf_aoi_toi"):
            assert abs(row[name] - o[name]) <= 1e-10 * max(1.0, abs(o[name])), name
        np.testing.assert_allclose(ce[ce["event"] == ev]["rmse_cell_toi"].to_numpy(), o["rmse_cell_toi"], rtol=1e-10, atol=1e-12)
    assert np.isfinite(gm.rmse_aoi_toi(truth_depth, pred_depth))
