"""CPU: host-side logic of the PreProcessor / metrics mirrors that needs no device -- the closed forms that turn the kernel's
raw reductions into the metrics of ``gpras/metrics.py``, North's rule, persistence keys -- against the reference-generated
golden vectors."""
import pickle

import numpy as np
import pytest

from conftest import MET_CASES, PRE_CASES, sub
from gpras_b200 import metrics as gm
from gpras_b200.preprocess import PreProcessor, compute_norths_rule


def _raw_reductions(x, y, conf, thr, v_tol):
    """What metrics_stream_kernel + metrics_finalize_kernel hand to the host (restated with NumPy)."""
    e = x - y
    xm, ym = x.max(axis=0), y.max(axis=0)
    d = xm - ym
    cells = np.stack([e.sum(axis=0), (e * e).sum(axis=0), conf.sum(axis=0), xm, ym])
    rows = np.stack([e.sum(axis=1), (e * e).sum(axis=1), conf.sum(axis=1)])
    hx, hy, zx, zy = xm >= thr, ym >= thr, xm >= 0, ym >= 0
    scal = np.array([cells[0].sum(), cells[1].sum(), cells[2].sum(), np.abs(e).sum(), (np.abs(e) <= v_tol).sum(), d.sum(), (d * d).sum(),
                     xm.sum(), ((xm - xm.mean()) ** 2).sum(), (hx & hy).sum(), (hx & ~hy).sum(), (~hx & hy).sum(),
                     (zx & zy).sum(), (zx & ~zy).sum(), (~zx & zy).sum()], float)
    return scal, cells, rows


@pytest.mark.parametrize("name", MET_CASES)
def test_summary_closed_forms_match_reference(met_golden, name):
    c = sub(met_golden, name)
    x, y, conf = c["x"], c["y"], c["conf"]
    scal, cells, rows = _raw_reductions(x, y, conf, float(c["depth_threshold"]), 0.0)
    s = gm._summary(scal, cells, rows, x.shape[0], x.shape[1])
    for k in ("rmse_cell_toi", "err_cell_toi", "conf_cell_toi", "err_cell_mts", "rmse_aoi_ts", "err_aoi_ts", "conf_aoi_ts"):
        np.testing.assert_allclose(s[k], c[k], rtol=1e-12, atol=1e-14, err_msg=k)
    for k in ("rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "err_aoi_toi", "rmse_aoi_mts", "err_aoi_mts", "nse_aoi_mts", "pod_mts",
              "rfa_mts", "csi_mts", "f2_mts", "f3_mts"):
        np.testing.assert_allclose(s[k], float(c[k]), rtol=1e-12, atol=1e-14, err_msg=k)
    assert s["fi_aoi_toi"] == float(c["fi_aoi_toi_0"])


@pytest.mark.parametrize("name", MET_CASES)
def test_peak_metrics_honour_caller_rows_and_the_reference_call(met_golden, name):
    """The O(cells) host closed form behind every ``*_mts`` function: caller-supplied rows that are not the arg-max rows, and the
    reference's own positional call of f2 / f3 (``gpras/metrics.py:53-54``: threshold = arg-max rows of x, truth peak read at the
    arg-max rows of y)."""
    c = sub(met_golden, name)
    x, y, thr = c["x"], c["y"], float(c["depth_threshold"])
    xr, yr = c["rows_x"], c["rows_y"]
    m = gm._peak_metrics(gm._gather_rows(x, xr), gm._gather_rows(y, yr), thr)
    np.testing.assert_allclose(m["err_cell_mts"], c["rows_err_cell_mts"], rtol=1e-12, atol=1e-14)
    for k in ("rmse_aoi_mts", "nse_aoi_mts", "err_aoi_mts", "pod_mts", "rfa_mts", "csi_mts", "f2_mts", "f3_mts"):
        np.testing.assert_allclose(m[k], float(c["rows_" + k]), rtol=1e-12, atol=1e-14, err_msg=k)
    called = gm._peak_metrics(gm._gather_rows(x, c["y_mts"]), y.max(axis=0), c["x_mts"])
    np.testing.assert_allclose(called["f2_mts"], float(c["f2_mts_as_called"]), rtol=1e-12)
    np.testing.assert_allclose(called["f3_mts"], float(c["f3_mts_as_called"]), rtol=1e-12)


def test_advisor_counter_example_for_f2_f3():
    # x and y peak on different rows and the thresholds (row indices 1, 1) are small: the reference gives 0.5 and 0.0
    x = np.array([[0, 0], [3, 3], [0.5, 0.5]], float)
    y = np.array([[0, 0], [1, 2.5], [2, 1]], float)
    called = gm._peak_metrics(gm._gather_rows(x, np.argmax(y, 0)), y.max(axis=0), np.argmax(x, 0))
    assert called["f2_mts"] == 0.5 and called["f3_mts"] == 0.0


def test_metrics_module_surface_matches_reference():
    # gpras/metrics.py:11-324
    names = ["export_metric_summary", "rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "rmse_aoi_ts", "rmse_cell_toi", "rmse_aoi_mts",
             "err_cell_mts", "nse_aoi_mts", "err_aoi_toi", "err_aoi_mts", "err_aoi_ts", "conf_aoi_ts", "err_cell_toi", "conf_cell_toi",
             "fi_aoi_toi", "pod_mts", "rfa_mts", "csi_mts", "f2_mts", "f3_mts"]
    for n in names:
        assert callable(getattr(gm, n)), n


@pytest.mark.parametrize("name", PRE_CASES)
def test_norths_rule_and_persistence_keys(pre_golden, name, tmp_path):
    c = sub(pre_golden, name)
    if int(c["modes_requested"]) < 0:
        assert compute_norths_rule(c["eigenvalues"], int(c["n_samples_fit"])) == int(c["spatial_mode_count"])
    pp = PreProcessor(spatial_mode_count=int(c["spatial_mode_count"]), input_mean=c["input_mean"], elevations=c["elevations"],
                      hydraulic_parameter=str(c["hydraulic_parameter"]), wetness_classes=c["wetness_classes"], weights=c["fit_weights"],
                      eofs=c["eofs"], eigenvalues=c["eigenvalues"], n_samples_fit=int(c["n_samples_fit"]), x_mean=c["x_mean"], x_std=c["x_std"])
    np.testing.assert_array_equal(pp.dry_indices, c["dry_indices"])
    pp.to_file(tmp_path / "pp.pkl")
    with open(tmp_path / "pp.pkl", "rb") as f:
        d = pickle.load(f)
    # the reference's to_dict keys (gpras/preprocess.py:1134-1148)
    assert set(d) == {"spatial_mode_count", "wet_threshold", "hydraulic_parameter", "elevations", "wetness_classes", "input_mean",
                      "weights", "eofs", "eigenvalues", "n_samples_fit", "x_mean", "x_std"}
    q = PreProcessor.from_file(tmp_path / "pp.pkl")
    np.testing.assert_array_equal(q.eofs, c["eofs"])
    np.testing.assert_array_equal(q.wse_2_depth(c["reverse_mean"].copy()), c["reverse_depth"])


def test_norths_rule_object_form():
    class Fitted:
        explained_variance_ = np.array([9.0, 4.0, 3.9, 1.2, 0.5])
        n_samples_seen_ = 50

    assert compute_norths_rule(Fitted()) == 1
    assert compute_norths_rule(np.array([0.5, 0.2]), 10) == 0


# ---- edge cases generated from the reference itself (tests/golden/edge_cases_reference.npz) -----------------------------
from pathlib import Path

EDGE = np.load(Path(__file__).resolve().parent / "golden" / "edge_cases_reference.npz")
EDGE_EVENTS = ["all_dry", "ties", "one_cell", "never_detected"]


def test_norths_rule_matches_reference_on_random_spectra():
    evs, ns, expected = EDGE["north.eigenvalues"], EDGE["north.n"], EDGE["north.expected"]
    from oracle import preprocess as opre

    checked = 0
    for ev, n, want in zip(evs, ns, expected):
        ev = ev[~np.isnan(ev)]
        if want < 0:
            # the reference raises here (argmax of an empty array when exactly one eigenvalue exceeds 1); the mirror keeps that mode
            assert compute_norths_rule(ev, int(n)) == 1 and opre.norths_rule(ev, int(n)) == 1
            continue
        assert compute_norths_rule(ev, int(n)) == int(want)
        assert opre.norths_rule(ev, int(n)) == int(want)
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("name", EDGE_EVENTS)
def test_degenerate_events_match_reference(name):
    """All-dry events, ties in the arg-max, a single cell, never-detected peaks: NaNs and the `== 1` branches of f2 / f3 come out
    exactly as the reference produces them -- for the oracle and for the host closed forms of the device summary."""
    from oracle import metrics as ometrics

    c = {k[len(name) + 1:]: EDGE[k] for k in EDGE.files if k.startswith(name + ".")}
    x, y, conf = c["x"], c["y"], c["conf"]
    with np.errstate(all="ignore"):
        so = ometrics.summarise(x, y, conf, 0.5, 0.0)
        sh = gm._summary(*_raw_reductions(x, y, conf, 0.5, 0.0), x.shape[0], x.shape[1])
    for s in (so, sh):
        for k in ("rmse_cell_toi", "err_cell_toi", "conf_cell_toi", "err_cell_mts", "rmse_aoi_ts", "err_aoi_ts", "conf_aoi_ts"):
            np.testing.assert_allclose(s[k], c[k], rtol=1e-12, atol=1e-14, err_msg=k)
        for k in ("rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "err_aoi_toi", "rmse_aoi_mts", "err_aoi_mts", "nse_aoi_mts", "pod_mts",
                  "rfa_mts", "csi_mts", "f2_mts", "f3_mts"):
            np.testing.assert_allclose(s[k], float(c[k]), rtol=1e-12, atol=1e-14, equal_nan=True, err_msg=k)
        assert s["fi_aoi_toi"] == float(c["fi_aoi_toi_0"])


def test_new_mirrors_fail_loudly_without_a_gpu(lib):
    """No CPU fallback: on a box without a CUDA device every compute entry of the PreProcessor / metrics mirrors raises."""
    if lib.gpras_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from gpras_b200 import _lib

    with pytest.raises(_lib.GprasError):
        gm.MetricsAccumulator(10, 5)
    with pytest.raises(_lib.GprasError):
        gm.summarise(np.zeros((2, 3)), np.ones((2, 3)))
    with pytest.raises(_lib.GprasError):
        PreProcessor().fit(np.random.default_rng(0).random((5, 7)), np.zeros(7), np.ones(7), 2)
    import ctypes as C

    h = C.c_void_p()
    assert lib.gpras_pre_create(C.byref(h), 0, 16, 0, 0.03) == -2 and b"no CPU fallback" in lib.gpras_last_error()
    assert lib.gpras_metrics_create(C.byref(h), 0, 16, 4) == -2 and b"no CPU fallback" in lib.gpras_last_error()


def test_host_closed_forms_agree_with_oracle_on_random_events():
    """Property test (hypothesis): for random event shapes, thresholds and tolerances the host closed forms applied to the raw
    reductions equal the oracle's whole-array evaluation -- two independent restatements of gpras/metrics.py."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from oracle import metrics as ometrics

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 23), st.integers(1, 37), st.integers(0, 2**31 - 1), st.floats(0.0, 1.5), st.floats(0.0, 0.5))
    def check(t, c, seed, thr, v_tol):
        rng = np.random.default_rng(seed)
        x = np.maximum(rng.standard_normal((t, c)) + 0.3, 0.0)
        y = np.maximum(x + 0.4 * rng.standard_normal((t, c)), 0.0)
        conf = rng.random((t, c))
        with np.errstate(all="ignore"):
            so = ometrics.summarise(x, y, conf, thr, v_tol)
            sh = gm._summary(*_raw_reductions(x, y, conf, thr, v_tol), t, c)
        for k, vo in so.items():
            np.testing.assert_allclose(sh[k], vo, rtol=1e-11, atol=1e-13, equal_nan=True, err_msg=k)

    check()
