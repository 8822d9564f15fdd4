"""GPU (-m gpu): the streaming metrics kernels (``gpras_b200/metrics.py`` -> ``gpras_metrics_*`` C ABI) against golden vectors
produced by the reference's own ``gpras/metrics.py`` (``tests/golden/metrics_reference.npz``), against the oracle at larger /
ragged sizes, and -- fused with the predictor -- against metrics of the materialised cell-space prediction."""
import numpy as np
import pytest

from conftest import MET_CASES, sub

pytestmark = pytest.mark.gpu

VEC = ("rmse_cell_toi", "err_cell_toi", "conf_cell_toi", "err_cell_mts", "rmse_aoi_ts", "err_aoi_ts", "conf_aoi_ts")
SCA = ("rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "err_aoi_toi", "rmse_aoi_mts", "err_aoi_mts", "nse_aoi_mts", "pod_mts",
       "rfa_mts", "csi_mts", "f2_mts", "f3_mts")


@pytest.fixture(scope="module")
def cuda(lib):
    import torch

    assert torch.cuda.is_available() and lib.gpras_device_count() > 0, "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("name", MET_CASES)
def test_summary_matches_reference_golden(cuda, met_golden, name):
    from gpras_b200 import metrics as gm

    c = sub(met_golden, name)
    s = gm.summarise(c["x"], c["y"], c["conf"], float(c["depth_threshold"]), 0.0)
    for k in VEC:
        np.testing.assert_allclose(s[k], c[k], rtol=1e-12, atol=1e-14, err_msg=k)
    for k in SCA:
        np.testing.assert_allclose(s[k], float(c[k]), rtol=1e-12, atol=1e-14, err_msg=k)
    assert s["fi_aoi_toi"] == float(c["fi_aoi_toi_0"])


@pytest.mark.parametrize("name", MET_CASES)
def test_reference_function_set(cuda, met_golden, name):
    """Every function of gpras/metrics.py by name, with the reference's signatures."""
    from gpras_b200 import metrics as gm

    c = sub(met_golden, name)
    x, y, conf = c["x"], c["y"], c["conf"]
    thr, t_tol, v_tol = float(c["depth_threshold"]), int(c["t_tol"]), float(c["v_tol"])
    xm, ym = c["x_mts"], c["y_mts"]
    tol = dict(rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gm.rmse_aoi_toi(x, y), float(c["rmse_aoi_toi"]), **tol)
    np.testing.assert_allclose(gm.mae_aoi_toi(x, y), float(c["mae_aoi_toi"]), **tol)
    np.testing.assert_allclose(gm.conf_aoi_toi(conf), float(c["conf_aoi_toi"]), **tol)
    np.testing.assert_allclose(gm.rmse_aoi_ts(x, y), c["rmse_aoi_ts"], **tol)
    np.testing.assert_allclose(gm.rmse_cell_toi(x, y), c["rmse_cell_toi"], **tol)
    np.testing.assert_allclose(gm.rmse_aoi_mts(x, y, xm, ym), float(c["rmse_aoi_mts"]), **tol)
    np.testing.assert_allclose(gm.err_cell_mts(x, y), c["err_cell_mts"], **tol)
    np.testing.assert_allclose(gm.nse_aoi_mts(x, y), float(c["nse_aoi_mts"]), **tol)
    np.testing.assert_allclose(gm.err_aoi_toi(x, y), float(c["err_aoi_toi"]), **tol)
    np.testing.assert_allclose(gm.err_aoi_mts(x, y), float(c["err_aoi_mts"]), **tol)
    np.testing.assert_allclose(gm.err_aoi_ts(x, y), c["err_aoi_ts"], **tol)
    np.testing.assert_allclose(gm.conf_aoi_ts(conf), c["conf_aoi_ts"], **tol)
    np.testing.assert_allclose(gm.err_cell_toi(x, y), c["err_cell_toi"], **tol)
    np.testing.assert_allclose(gm.conf_cell_toi(conf), c["conf_cell_toi"], **tol)
    assert gm.fi_aoi_toi(x, y, t_tol, v_tol) == float(c["fi_aoi_toi"])
    assert gm.fi_aoi_toi(x, y, 0, 0) == float(c["fi_aoi_toi_0"])
    np.testing.assert_allclose(gm.pod_mts(x, y, thr, xm, ym), float(c["pod_mts"]), **tol)
    np.testing.assert_allclose(gm.rfa_mts(x, y, thr, xm, ym), float(c["rfa_mts"]), **tol)
    np.testing.assert_allclose(gm.csi_mts(x, y, thr, xm, ym), float(c["csi_mts"]), **tol)
    np.testing.assert_allclose(gm.f2_mts(x, y, 0, xm, ym), float(c["f2_mts"]), **tol)
    np.testing.assert_allclose(gm.f3_mts(x, y, 0, xm, ym), float(c["f3_mts"]), **tol)
    np.testing.assert_allclose(gm.f2_mts(x, y, xm, ym), float(c["f2_mts_as_called"]), **tol)
    np.testing.assert_allclose(gm.f3_mts(x, y, xm, ym), float(c["f3_mts_as_called"]), **tol)
    # caller-supplied rows that are not the arg-max rows are honoured (gpras/metrics.py:111-318 index with them)
    xr, yr = c["rows_x"], c["rows_y"]
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(gm.rmse_aoi_mts(x, y, xr, yr), float(c["rows_rmse_aoi_mts"]), **tol)
        np.testing.assert_allclose(gm.nse_aoi_mts(x, y, xr, yr), float(c["rows_nse_aoi_mts"]), **tol)
        np.testing.assert_allclose(gm.err_aoi_mts(x, y, xr, yr), float(c["rows_err_aoi_mts"]), **tol)
        np.testing.assert_allclose(gm.err_cell_mts(x, y, xr, yr), c["rows_err_cell_mts"], **tol)
        np.testing.assert_allclose(gm.pod_mts(x, y, thr, xr, yr), float(c["rows_pod_mts"]), **tol)
        np.testing.assert_allclose(gm.rfa_mts(x, y, thr, xr, yr), float(c["rows_rfa_mts"]), **tol)
        np.testing.assert_allclose(gm.csi_mts(x, y, thr, xr, yr), float(c["rows_csi_mts"]), **tol)
        np.testing.assert_allclose(gm.f2_mts(x, y, thr, xr, yr), float(c["rows_f2_mts"]), **tol)
        np.testing.assert_allclose(gm.f3_mts(x, y, thr, xr, yr), float(c["rows_f3_mts"]), **tol)
        # only one side supplied: the other comes from the device pass
        np.testing.assert_allclose(gm.err_cell_mts(x, y, xr, None), gm._gather_rows(x, xr) - y.max(axis=0), **tol)


@pytest.mark.parametrize("t,c,blocks", [(333, 1000, 1), (2500, 777, 3), (70, 40000, 2)])
def test_streaming_updates_against_oracle(cuda, t, c, blocks):
    """Several update() calls (host and device inputs), row / column counts that are not tile multiples, depth conversion."""
    torch = cuda
    from gpras_b200.metrics import MetricsAccumulator
    from oracle import metrics as om

    rng = np.random.default_rng(t + c)
    elev = rng.uniform(0, 3, c)
    x = elev + np.maximum(rng.standard_normal((t, c)) + 0.3, -0.5)
    y = x + 0.3 * rng.standard_normal((t, c))
    conf = rng.uniform(0.01, 0.4, (t, c))
    acc = MetricsAccumulator(c, t)
    acc.set_elevations(elev, elev)
    acc.reset(0.05)
    cuts = np.linspace(0, t, blocks + 1).astype(int)
    for i in range(blocks):
        sl = slice(cuts[i], cuts[i + 1])
        if i % 2 == 0:
            acc.update(x[sl], y[sl], conf[sl])
        else:
            acc.update(torch.from_numpy(x[sl]).cuda(), torch.from_numpy(y[sl]).cuda(), torch.from_numpy(conf[sl]).cuda())
    s = acc.finalize(0.5)
    ref = om.summarise(np.maximum(x - elev, 0), np.maximum(y - elev, 0), conf, 0.5, 0.05)
    for k in VEC:
        np.testing.assert_allclose(s[k], ref[k], rtol=1e-11, atol=1e-13, err_msg=k)
    for k in SCA + ("fi_aoi_toi",):
        np.testing.assert_allclose(s[k], ref[k], rtol=1e-11, atol=1e-13, err_msg=k)
    # bitwise repeatable (fixed-shape reductions, no atomics)
    acc.reset(0.05)
    acc.update(x, y, conf)
    s2 = acc.finalize(0.5)
    acc.reset(0.05)
    acc.update(x, y, conf)
    s3 = acc.finalize(0.5)
    for k in VEC + SCA:
        np.testing.assert_array_equal(s2[k], s3[k])
    acc.close()


@pytest.mark.parametrize("p,with_truth", [(8, True), (32, True), (40, False)])
def test_fused_predict_metrics_match_materialised_path(cuda, p, with_truth):
    """predict -> cells -> metrics without materialising the cell-space prediction == metrics of the materialised one."""
    torch = cuda
    from gpras_b200.cells import fold_cell_map
    from gpras_b200.engine import ExactGP
    from gpras_b200.metrics import MetricsAccumulator
    from gpras_b200.synth import fixed_theta, make_cell_map, make_gp_data
    from oracle import metrics as om
    from oracle.cells import reverse_transform

    n, d, c, t = 300, 6, 3001, 1111
    data = make_gp_data(n, d, p, t, seed=3)
    cm = make_cell_map(p, c, seed=4)
    e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    gp = ExactGP("Matern52", n, d, p)
    gp.set_data(data.x, data.y)
    v, s, ls = fixed_theta(d, True)
    gp.condition(gp.theta_vector(v, s, ls))
    gp.set_cell_map(e_mean, bias)
    mean, var = gp.predict(data.x_test)
    cell_m, cell_v = reverse_transform(mean, var, cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices,
                                       cm.elevations)
    rng = np.random.default_rng(9)
    truth = cell_m + 0.2 * rng.standard_normal(cell_m.shape) if with_truth else None
    acc = MetricsAccumulator(c, t)
    acc.set_elevations(cm.elevations if with_truth else None, cm.elevations)
    acc.reset(0.1)
    mm, mv = acc.predict_update(gp, data.x_test[:600], None if truth is None else truth[:600], want_modes=True)
    acc.predict_update(gp, torch.from_numpy(data.x_test[600:]).cuda(), None if truth is None else torch.from_numpy(truth[600:]).cuda())
    got = acc.finalize(0.5)
    np.testing.assert_allclose(mm, mean[:600], rtol=1e-12, atol=1e-12)
    y = np.maximum(cell_m - cm.elevations, 0)
    x = np.maximum(truth - cm.elevations, 0) if with_truth else np.zeros_like(y)
    ref = om.summarise(x, y, np.sqrt(cell_v), 0.5, 0.1)
    for k in VEC:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-11, err_msg=k)
    for k in SCA + ("fi_aoi_toi",):
        if np.isfinite(ref[k]):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-11, err_msg=k)
    assert acc.timesteps() == t
    acc.close()
    gp.close()


def test_export_metric_summary_tables(cuda, tmp_path):
    """Same sqlite tables / columns as gpras/metrics.py:75-82, values equal to the per-event summaries."""
    import sqlite3

    import pandas as pd

    from gpras_b200 import metrics as gm

    rng = np.random.default_rng(0)
    idx = pd.MultiIndex.from_product([["e1", "e2"], range(12)], names=["event", "t"])
    cols = [f"c{i}" for i in range(37)]
    x = pd.DataFrame(np.maximum(rng.standard_normal((24, 37)) + 0.5, 0), index=idx, columns=cols)
    y = pd.DataFrame(np.maximum(x.values + 0.2 * rng.standard_normal((24, 37)), 0), index=idx, columns=cols)
    conf = pd.DataFrame(rng.uniform(0.1, 0.3, (24, 37)), index=idx, columns=cols)
    out = tmp_path / "m.db"
    gm.export_metric_summary(x, y, conf, out)
    with sqlite3.connect(out) as con:
        sc = pd.read_sql("select * from scalar_metrics", con)
        ts = pd.read_sql("select * from timeseries_metrics", con)
        ce = pd.read_sql("select * from cell_metrics", con)
    assert list(sc.columns) == ["event", "rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "rmse_aoi_mts", "nse_aoi_mts", "err_aoi_toi",
                                "err_aoi_mts", "fi_aoi_toi", "pod_mts", "rfa_mts", "csi_mts", "f2_mts", "f3_mts"]
    assert len(sc) == 2 and len(ts) == 24 and len(ce) == 74
    e1 = x.loc["e1"].values - y.loc["e1"].values
    np.testing.assert_allclose(sc.rmse_aoi_toi[0], np.sqrt((e1 ** 2).mean()), rtol=1e-12)
    np.testing.assert_allclose(ts.err_aoi_ts[:12], e1.mean(axis=1), rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(ce.rmse_cell_toi[:37], np.sqrt((e1 ** 2).mean(axis=0)), rtol=1e-12)


def test_export_metric_summary_matches_the_reference_export(cuda, met_golden, tmp_path):
    """All three sqlite tables against the reference's own export_metric_summary run on the same frames (golden), including its
    positional f2 / f3 call (gpras/metrics.py:53-54) on events whose truth and prediction peak on different rows, t_tol > 0."""
    import sqlite3

    import pandas as pd

    from gpras_b200 import metrics as gm

    g = sub(met_golden, "export")
    t, c = 12, 37
    idx = pd.MultiIndex.from_product([["e1", "e2"], range(t)], names=["event", "t"])
    cols = [f"c{i}" for i in range(c)]
    frames = [pd.DataFrame(g[k], index=idx, columns=cols) for k in ("x", "y", "conf")]
    out = tmp_path / "m.db"
    gm.export_metric_summary(*frames, out, depth_threshold=0.5, t_tol=1, v_tol=0.05)
    with sqlite3.connect(out) as con:
        sc = pd.read_sql("select * from scalar_metrics", con)
        ts = pd.read_sql("select * from timeseries_metrics", con)
        ce = pd.read_sql("select * from cell_metrics", con)
    assert list(sc.columns) == [str(v) for v in g["scalar_columns"]]
    np.testing.assert_allclose(sc.drop(columns=["event"]).to_numpy(np.float64), g["scalar"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ts.drop(columns=["event", "timestep"]).to_numpy(np.float64), g["timeseries"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ce.drop(columns=["event", "cell_id"]).to_numpy(np.float64), g["cells"], rtol=1e-12, atol=1e-14)


def test_full_size_properties_cfg5_block(cuda):
    """BASELINE sizes (C = 200 000 cells, one 2 048-timestep block): properties that do not need an oracle pass.
    y = x + delta  =>  rmse = mae = |delta|, err = -delta, every per-cell / per-timestep value equals the closed form;
    swapping x and y flips the sign of the errors; results are bitwise repeatable."""
    torch = cuda
    from gpras_b200.metrics import MetricsAccumulator

    t, c, delta = 2048, 200_000, 0.125
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(t, c, dtype=torch.float64, device="cuda", generator=g) * 4
    y = x + delta
    conf = torch.full((t, c), 0.25, dtype=torch.float64, device="cuda")
    acc = MetricsAccumulator(c, t)
    acc.reset(0.2)
    acc.update(x, y, conf)
    s = acc.finalize(0.5)
    assert abs(s["rmse_aoi_toi"] - delta) < 1e-12 and abs(s["mae_aoi_toi"] - delta) < 1e-12 and abs(s["err_aoi_toi"] + delta) < 1e-12
    np.testing.assert_allclose(s["rmse_cell_toi"], delta, rtol=1e-12)
    np.testing.assert_allclose(s["err_aoi_ts"], -delta, rtol=1e-12)
    np.testing.assert_allclose(s["err_cell_mts"], -delta, rtol=1e-12)
    np.testing.assert_allclose(s["conf_aoi_ts"], 0.25, rtol=1e-13)
    assert s["conf_aoi_toi"] == 0.25 and s["fi_aoi_toi"] == 1.0  # |e| = 0.125 <= v_tol everywhere
    acc.reset(0.2)
    acc.update(y, x, conf)
    s2 = acc.finalize(0.5)
    np.testing.assert_allclose(s2["err_cell_toi"], -s["err_cell_toi"], rtol=1e-13)
    np.testing.assert_array_equal(s2["rmse_cell_toi"], s["rmse_cell_toi"])
    acc.reset(0.2)
    acc.update(x, y, conf)
    s3 = acc.finalize(0.5)
    for k in VEC + SCA:
        np.testing.assert_array_equal(s3[k], s[k])
    acc.close()


def test_metrics_misuse_raises(cuda):
    from gpras_b200 import _lib
    from gpras_b200.metrics import MetricsAccumulator

    acc = MetricsAccumulator(40, 10)
    acc.reset(0.0)
    with pytest.raises(ValueError):
        acc.update(np.zeros((3, 41)), np.zeros((3, 41)))  # wrong cell count
    with pytest.raises(_lib.GprasError):
        acc.finalize()  # nothing accumulated
    acc.update(np.zeros((6, 40)), np.ones((6, 40)))
    with pytest.raises(_lib.GprasError):
        acc.update(np.zeros((6, 40)), np.ones((6, 40)))  # beyond the accumulator's capacity of 10 timesteps
    s = acc.finalize(0.5)
    assert s["rmse_aoi_toi"] == 1.0 and s["err_aoi_toi"] == -1.0 and acc.timesteps() == 6
    acc.close()


@pytest.mark.parametrize("p,cells,t,with_truth", [(7, 333, 70, True), (40, 1500, 2300, True), (16, 900, 300, False)])
def test_reverse_metrics_with_per_mode_variances_match_materialised_path(cuda, p, cells, t, with_truth):
    """Per-column models give one variance per mode; the fused reverse-transform + metrics consumer (nothing written per
    cell-depth) must equal the metrics of the materialised prediction: oracle reverse transform -> depth -> oracle metrics."""
    torch = cuda
    from gpras_b200.metrics import MetricsAccumulator
    from gpras_b200.preprocess import PreProcessor
    from gpras_b200.synth import make_cell_map
    from oracle import metrics as om
    from oracle.cells import reverse_transform

    cm = make_cell_map(p, cells, seed=p)
    rng = np.random.default_rng(p + t)
    mean, var = rng.standard_normal((t, p)), rng.uniform(0.01, 2.0, (t, p))
    pp = PreProcessor(spatial_mode_count=p, input_mean=cm.input_mean, elevations=cm.elevations, hydraulic_parameter="wse",
                      wetness_classes=np.where(cm.dry_indices, "AD", "TF"), weights=cm.weights, eofs=cm.eofs, eigenvalues=np.ones(p),
                      n_samples_fit=100, x_mean=cm.x_mean, x_std=cm.x_std)
    rm, rv = reverse_transform(mean, var, cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    truth = rm + 0.3 * rng.standard_normal(rm.shape) if with_truth else None
    acc = MetricsAccumulator(cells, t)
    acc.set_elevations(cm.elevations if with_truth else None, cm.elevations)
    acc.reset(0.1)
    half = t // 2
    acc.reverse_update(pp, mean[:half], var[:half], None if truth is None else truth[:half])
    acc.reverse_update(pp, torch.from_numpy(mean[half:]).cuda(), torch.from_numpy(var[half:]).cuda(),
                       None if truth is None else torch.from_numpy(truth[half:]).cuda())
    got = acc.finalize(0.5)
    y = np.maximum(rm - cm.elevations, 0)
    x = np.maximum(truth - cm.elevations, 0) if with_truth else np.zeros_like(y)
    ref = om.summarise(x, y, np.sqrt(rv), 0.5, 0.1)
    for k in VEC:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-11, err_msg=k)
    for k in SCA + ("fi_aoi_toi",):
        if np.isfinite(ref[k]):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-11, err_msg=k)
    assert acc.timesteps() == t
    acc.close()
    pp.close()
