"""GPU (-m gpu): the device PreProcessor mirror (``gpras_b200/preprocess.py`` -> ``gpras_pre_*`` C ABI) against golden vectors
produced by the reference's own ``PreProcessor`` (``tests/golden/preprocess_reference.npz``) and against the oracle at sizes
the goldens do not cover.  FP64 throughout; tolerances: classes exact, means 1e-12, EOFs / scores 1e-8 (the PCA goes
through the Gram matrix, whose eigenvectors carry ~eps * lambda_1 / gap), transform / reverse transform 1e-10."""
import numpy as np
import pytest

from conftest import PRE_CASES, sub

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda(lib):
    import torch

    assert torch.cuda.is_available() and lib.gpras_device_count() > 0, "GPU tests need a CUDA device"
    return torch


def _fit(c):
    from gpras_b200.preprocess import PreProcessor

    hp = str(c["hydraulic_parameter"])
    modes = None if int(c["modes_requested"]) < 0 else int(c["modes_requested"])
    pp = PreProcessor(wet_threshold=0.03, hydraulic_parameter=hp)
    pp.fit(c["x"].copy(), c["elevations"].copy(), c["weights"].copy(), modes)
    return pp


@pytest.mark.parametrize("name", PRE_CASES)
def test_fit_matches_reference_golden(cuda, pre_golden, name):
    c = sub(pre_golden, name)
    pp = _fit(c)
    assert list(pp.wetness_classes) == list(c["wetness_classes"])
    assert pp.spatial_mode_count == int(c["spatial_mode_count"])
    np.testing.assert_allclose(pp.input_mean, c["input_mean"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(pp.weights, c["fit_weights"])
    k = len(pp.eigenvalues)
    np.testing.assert_allclose(pp.eigenvalues, c["eigenvalues"][:k], rtol=1e-9, atol=1e-9 * c["eigenvalues"][0])
    np.testing.assert_allclose(pp.eofs, c["eofs"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(pp.x_mean, c["x_mean"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(pp.x_std, c["x_std"], rtol=1e-9)
    assert pp.n_samples_fit == int(c["n_samples_fit"])
    pp.close()


@pytest.mark.parametrize("name", PRE_CASES)
def test_transform_and_reverse_match_reference_golden(cuda, pre_golden, name, tmp_path):
    from gpras_b200.preprocess import PreProcessor

    c = sub(pre_golden, name)
    # state taken from the reference's fit: isolates transform / reverse_transform from the PCA
    pp = PreProcessor(
        spatial_mode_count=int(c["spatial_mode_count"]), input_mean=c["input_mean"], wet_threshold=0.03, elevations=c["elevations"],
        hydraulic_parameter=str(c["hydraulic_parameter"]), wetness_classes=c["wetness_classes"], weights=c["fit_weights"],
        eofs=c["eofs"], eigenvalues=c["eigenvalues"], n_samples_fit=int(c["n_samples_fit"]), x_mean=c["x_mean"], x_std=c["x_std"])
    # persistence round trip with the reference's pickle keys
    pp.to_file(tmp_path / "pp.pkl")
    pp = PreProcessor.from_file(tmp_path / "pp.pkl")
    z = pp.transform(c["x_new"].copy())
    np.testing.assert_allclose(z, c["transformed"], rtol=1e-10, atol=1e-10)
    m = pp.reverse_transform(c["mode_mean"])
    np.testing.assert_allclose(m, c["reverse_mean_only"], rtol=1e-12, atol=1e-12)
    for vk, rk in (("mode_var", "reverse_var"), ("mode_var_free", "reverse_var_free")):
        m, v = pp.reverse_transform(c["mode_mean"], c[vk])
        np.testing.assert_allclose(m, c["reverse_mean"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(v, c[rk], rtol=1e-12, atol=1e-14)
    np.testing.assert_array_equal(pp.wse_2_depth(c["reverse_mean"].copy()), c["reverse_depth"])
    pp.close()


def test_fit_then_transform_device_tensor(cuda, pre_golden):
    """CUDA-tensor inputs (no host staging) give the same result as host arrays."""
    torch = cuda
    c = sub(pre_golden, "wse_fixed")
    pp = _fit(c)
    z_host = pp.transform(c["x_new"].copy())
    z_dev = pp.transform(torch.from_numpy(c["x_new"]).cuda())
    np.testing.assert_array_equal(z_dev.cpu().numpy(), z_host)
    np.testing.assert_allclose(z_host, c["transformed"], rtol=1e-7, atol=1e-7)
    pp.close()


@pytest.mark.parametrize("n,cells,k,modes,hp", [(700, 3000, 6, 6, "wse"), (1500, 1100, 5, 5, "depth"), (300, 5000, 8, None, "wse"), (2500, 1500, 6, 6, "wse")])
def test_fit_subspace_iteration_against_oracle(cuda, n, cells, k, modes, hp):
    """More samples than one 128-block: the subspace iteration has to converge (not just diagonalise the whole Gram).
    The last case has more than 2048 samples and takes the Gram-free (operator) form of the iteration."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    from make_golden_reference import flood_samples
    from gpras_b200.preprocess import PreProcessor
    from oracle import preprocess as opre

    wse, elev, w = flood_samples(n, cells, k, seed=n + cells, noise=0.02)
    f = opre.fit(wse, elev, w, modes, 0.03, hp)
    pp = PreProcessor(wet_threshold=0.03, hydraulic_parameter=hp)
    pp.fit(wse.copy(), elev, w, modes)
    assert pp.spatial_mode_count == f.modes
    np.testing.assert_array_equal(pp.dry_indices, f.dry)
    np.testing.assert_allclose(pp.input_mean, f.input_mean, rtol=1e-12, atol=1e-12)
    p = f.modes
    np.testing.assert_allclose(pp.eigenvalues[:p], f.eigenvalues[:p], rtol=1e-9)
    np.testing.assert_allclose(pp.eofs, f.eofs, rtol=0, atol=1e-8)
    np.testing.assert_allclose(pp.x_std, f.x_std, rtol=1e-9)
    np.testing.assert_allclose(pp.transform(wse[:50].copy()), opre.transform(f, wse[:50], elev, hp), rtol=1e-8, atol=1e-8)
    assert pp.fit_info["iterations"] >= 1
    # round trip: reverse_transform(transform(x)) is the rank-P projector of the centred, weighted field
    if hp == "wse":  # ("depth" clamps negative reconstructed depths on the way back in, so it is not a projector)
        z = pp.transform(wse[:20].copy())
        back = pp.reverse_transform(z)
        z2 = pp.transform(back)
        np.testing.assert_allclose(z2, z, rtol=1e-8, atol=1e-8)
    pp.close()


def test_dsyev128_against_lapack(cuda, lib):
    torch = cuda
    from gpras_b200 import _lib

    rng = np.random.default_rng(5)
    a = rng.standard_normal((128, 90))
    h = a @ np.diag(np.logspace(0, -6, 90)) @ a.T
    H = torch.from_numpy(h).cuda()
    lam = torch.zeros(128, dtype=torch.float64, device="cuda")
    V = torch.zeros(128, 128, dtype=torch.float64, device="cuda")
    _lib.check(lib.gpras_dsyev128(torch.cuda.current_stream().cuda_stream, H.data_ptr(), lam.data_ptr(), V.data_ptr()))
    w = np.linalg.eigvalsh(h)[::-1]
    np.testing.assert_allclose(lam.cpu().numpy(), np.maximum(w, 0), rtol=0, atol=1e-13 * w[0])
    v = V.cpu().numpy()[:, :40]
    np.testing.assert_allclose(v.T @ v, np.eye(40), atol=1e-9)
    np.testing.assert_allclose(h @ v, v * lam.cpu().numpy()[:40], atol=1e-12 * w[0])


def test_full_size_properties_cfg3(cuda):
    """BASELINE config 3 sizes (8 192 samples x 200 000 cells, 32 modes, generated on the device): EOF rows orthonormal,
    eigenvalues descending and consistent with the score variances, retained residuals converged,
    transform(reverse_transform(z)) == z, scores of the training samples standardised."""
    import sys
    from pathlib import Path

    torch = cuda
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
    from bench_pre_metrics import flood_tensor
    from gpras_b200.preprocess import PreProcessor

    n, c, p = 8192, 200_000, 32
    x, elev, w = flood_tensor(torch, n, c, k=40)
    pp = PreProcessor(hydraulic_parameter="wse")
    pp.fit(x, elev, w, p)
    e = pp.eofs
    np.testing.assert_allclose(e @ e.T, np.eye(p), atol=1e-10)
    ev = pp.eigenvalues
    assert np.all(np.diff(ev[: p + 8]) <= 0)
    assert np.max(pp.fit_info["residuals"][:p]) <= 1e-11
    np.testing.assert_allclose(pp.x_std**2 * n / (n - 1), ev[:p], rtol=1e-9)   # population variance of the scores vs s^2 / (n - 1)
    z = pp.transform(x[:256]).cpu().numpy()
    z_all = pp.transform(x).cpu().numpy()
    np.testing.assert_allclose(z_all.mean(axis=0), 0.0, atol=1e-9)
    np.testing.assert_allclose(z_all.std(axis=0), 1.0, rtol=1e-9)
    back = pp.reverse_transform(z)
    np.testing.assert_allclose(pp.transform(back), z, rtol=1e-8, atol=1e-8)
    pp.close()


def test_edge_shapes_and_errors(cuda):
    """One mode, one sample to transform, one event to expand, fewer cells than one tile; misuse raises."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    from make_golden_reference import flood_samples
    from gpras_b200.preprocess import PreProcessor
    from oracle import preprocess as opre
    from oracle.cells import reverse_transform

    wse, elev, w = flood_samples(37, 50, 3, seed=77)
    f = opre.fit(wse, elev, w, 1, 0.03, "wse")
    pp = PreProcessor(hydraulic_parameter="wse")
    with pytest.raises(ValueError):
        pp.transform(wse[:1].copy())  # not fitted
    pp.fit(wse.copy(), elev, w, 1)
    assert pp.eofs.shape == f.eofs.shape == (1, int((~f.dry).sum()))
    np.testing.assert_allclose(pp.eofs, f.eofs, atol=1e-9)
    z = pp.transform(wse[:1].copy())
    assert z.shape == (1, 1)
    np.testing.assert_allclose(z, opre.transform(f, wse[:1]), rtol=1e-9, atol=1e-9)
    m, v = pp.reverse_transform(z, np.abs(z))
    rm, rv = reverse_transform(z, np.abs(z), pp.eofs, pp.x_mean, pp.x_std, pp.weights, pp.input_mean, pp.dry_indices, elev)
    np.testing.assert_allclose(m, rm, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(v, rv, rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        pp.transform(np.zeros((3, 49)))  # wrong number of cells
    with pytest.raises(ValueError):
        pp.reverse_transform(np.zeros((2, 3)))  # wrong number of modes
    pp.close()


@pytest.mark.parametrize("name", PRE_CASES)
def test_reverse_transform_device_general_variance_matches_reference(cuda, pre_golden, name):
    """The fused device-resident reverse transform (one variance per mode -- the reference's per-column model family) against
    the reference's own ``reverse_transform`` outputs for free per-mode variances (golden ``reverse_var_free``)."""
    torch = cuda
    from gpras_b200.preprocess import PreProcessor

    c = sub(pre_golden, name)
    pp = PreProcessor(spatial_mode_count=int(c["spatial_mode_count"]), input_mean=c["input_mean"], elevations=c["elevations"],
                      hydraulic_parameter=str(c["hydraulic_parameter"]), wetness_classes=c["wetness_classes"], weights=c["fit_weights"],
                      eofs=c["eofs"], eigenvalues=c["eigenvalues"], n_samples_fit=int(c["n_samples_fit"]), x_mean=c["x_mean"], x_std=c["x_std"])
    t, cells = c["reverse_mean"].shape
    pitch = pp.cell_pitch()
    rows = (t + 63) // 64 * 64
    cm = torch.full((rows, pitch), float("nan"), dtype=torch.float64, device="cuda")
    cv = torch.full((rows, pitch), float("nan"), dtype=torch.float64, device="cuda")
    pp.reverse_transform_device(c["mode_mean"], c["mode_var_free"], cm, cv)
    torch.cuda.synchronize()
    np.testing.assert_allclose(cm[:t, :cells].cpu().numpy(), c["reverse_mean"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(cv[:t, :cells].cpu().numpy(), c["reverse_var_free"], rtol=1e-12, atol=1e-14)
    # device inputs, and the ring-buffer mode (nothing kept) runs the same kernel
    cm2, cv2 = torch.empty_like(cm), torch.empty_like(cv)
    pp.reverse_transform_device(torch.from_numpy(c["mode_mean"]).cuda(), torch.from_numpy(c["mode_var_free"]).cuda(), cm2, cv2)
    torch.cuda.synchronize()
    assert torch.equal(cm2[:t, :cells], cm[:t, :cells]) and torch.equal(cv2[:t, :cells], cv[:t, :cells])
    pp.reverse_transform_device(c["mode_mean"], c["mode_var_free"])
    # and it agrees with the host-output path
    hm, hv = pp.reverse_transform(c["mode_mean"], c["mode_var_free"])
    np.testing.assert_allclose(cm[:t, :cells].cpu().numpy(), hm, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(cv[:t, :cells].cpu().numpy(), hv, rtol=1e-13, atol=1e-15)
    pp.close()


def test_reverse_transform_device_many_events_and_64_modes(cuda):
    """Several 1024-event blocks, a row count that is no tile multiple, 40 modes (the 64-wide kernel) against the oracle."""
    torch = cuda
    from gpras_b200.preprocess import PreProcessor
    from gpras_b200.synth import make_cell_map
    from oracle.cells import reverse_transform

    for p, cells, t in ((40, 1500, 2300), (7, 333, 70)):
        cm = make_cell_map(p, cells, seed=p)
        rng = np.random.default_rng(p)
        mean, var = rng.standard_normal((t, p)), rng.uniform(0.01, 2.0, (t, p))
        classes = np.where(cm.dry_indices, "AD", "TF")
        eofs_wet = cm.eofs
        pp = PreProcessor(spatial_mode_count=p, input_mean=cm.input_mean, elevations=cm.elevations, hydraulic_parameter="wse",
                          wetness_classes=classes, weights=cm.weights, eofs=eofs_wet, eigenvalues=np.ones(p), n_samples_fit=100,
                          x_mean=cm.x_mean, x_std=cm.x_std)
        pitch = pp.cell_pitch()
        rows = (t + 63) // 64 * 64
        om = torch.empty((rows, pitch), dtype=torch.float64, device="cuda")
        ov = torch.empty((rows, pitch), dtype=torch.float64, device="cuda")
        pp.reverse_transform_device(mean, var, om, ov)
        torch.cuda.synchronize()
        rm, rv = reverse_transform(mean, var, cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
        np.testing.assert_allclose(om[:t, :cells].cpu().numpy(), rm, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(ov[:t, :cells].cpu().numpy(), rv, rtol=1e-11, atol=1e-13)
        pp.close()
