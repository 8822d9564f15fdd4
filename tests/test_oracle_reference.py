"""CPU: the preprocessing / cell-expansion / metrics oracles against golden vectors produced by the reference's OWN code
(``tests/golden/make_golden_reference.py`` ran the unmodified ``gpras.preprocess.PreProcessor`` and
``gpras.metrics`` in the build container).  These pins make rows a14 / next #2 / next #3 of SURVEY.md section 8
"parity pinned"."""
import numpy as np
import pytest

from conftest import MET_CASES, PRE_CASES, sub
from oracle import metrics as ometrics
from oracle import preprocess as opre
from oracle.cells import reverse_transform

CLASS_NAMES = {0: "", 1: "AD", 2: "TF", 3: "AF"}


@pytest.mark.parametrize("name", PRE_CASES)
def test_fit_matches_reference(pre_golden, name):
    c = sub(pre_golden, name)
    hp = str(c["hydraulic_parameter"])
    modes = None if int(c["modes_requested"]) < 0 else int(c["modes_requested"])
    f = opre.fit(c["x"], c["elevations"], c["weights"], modes, 0.03, hp)
    assert [CLASS_NAMES[int(k)] for k in f.wet_class] == list(c["wetness_classes"])
    np.testing.assert_array_equal(f.dry, c["dry_indices"])
    assert f.modes == int(c["spatial_mode_count"])
    np.testing.assert_allclose(f.input_mean, c["input_mean"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(f.weights, c["fit_weights"], rtol=0, atol=0)
    np.testing.assert_allclose(f.eigenvalues, c["eigenvalues"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(f.eofs, c["eofs"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(f.x_mean, c["x_mean"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(f.x_std, c["x_std"], rtol=1e-10)


@pytest.mark.parametrize("name", PRE_CASES)
def test_transform_matches_reference(pre_golden, name):
    c = sub(pre_golden, name)
    hp = str(c["hydraulic_parameter"])
    f = opre.Fitted(None, c["dry_indices"], c["input_mean"], c["fit_weights"], c["eofs"], c["eigenvalues"], c["x_mean"],
                    c["x_std"], int(c["spatial_mode_count"]))
    z = opre.transform(f, c["x_new"], c["elevations"], hp)
    np.testing.assert_allclose(z, c["transformed"], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("name", PRE_CASES)
def test_reverse_transform_matches_reference(pre_golden, name):
    c = sub(pre_golden, name)
    depth = str(c["hydraulic_parameter"]) == "depth"
    args = (c["eofs"], c["x_mean"], c["x_std"], c["fit_weights"], c["input_mean"], c["dry_indices"], c["elevations"], depth)
    m = reverse_transform(c["mode_mean"], None, *args)
    np.testing.assert_allclose(m, c["reverse_mean_only"], rtol=1e-13, atol=1e-13)
    for vk, rk in (("mode_var", "reverse_var"), ("mode_var_free", "reverse_var_free")):
        m, v = reverse_transform(c["mode_mean"], c[vk], *args)
        np.testing.assert_allclose(m, c["reverse_mean"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(v, c[rk], rtol=1e-13, atol=1e-15)
    np.testing.assert_array_equal(opre.wse_to_depth(c["reverse_mean"], c["elevations"]), c["reverse_depth"])


@pytest.mark.parametrize("name", MET_CASES)
def test_streaming_metrics_match_reference(met_golden, name):
    c = sub(met_golden, name)
    s = ometrics.summarise(c["x"], c["y"], c["conf"], float(c["depth_threshold"]), 0.0)
    for k in ("rmse_cell_toi", "err_cell_toi", "conf_cell_toi", "err_cell_mts", "rmse_aoi_ts", "err_aoi_ts", "conf_aoi_ts"):
        np.testing.assert_allclose(s[k], c[k], rtol=1e-12, atol=1e-14, err_msg=k)
    for k in ("rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "err_aoi_toi", "rmse_aoi_mts", "err_aoi_mts", "nse_aoi_mts",
              "pod_mts", "rfa_mts", "csi_mts", "f2_mts", "f3_mts"):
        np.testing.assert_allclose(s[k], float(c[k]), rtol=1e-12, atol=1e-14, err_msg=k)
    assert s["fi_aoi_toi"] == float(c["fi_aoi_toi_0"])


def test_norths_rule_edge_cases():
    assert opre.norths_rule([0.5, 0.2], 10) == 0            # Kaiser filter leaves nothing
    assert opre.norths_rule([9.0, 8.9, 1.5], 10) == 3       # first gap already within the error bar -> keep all
    assert opre.norths_rule([9.0, 4.0, 3.9, 1.2], 50) == 1  # second gap fails first
    assert opre.norths_rule([9.0], 50) == 1
