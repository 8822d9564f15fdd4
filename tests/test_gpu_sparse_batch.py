"""The per-column sparse models of one fit, batched and trained on the device (``gpras_sgpr_batch_*``).

The batched evaluation must be bitwise the single-model one (same kernels, model index in the grid) and agree with the torch
oracle; the device-resident Adam stage must follow the reference's ``_optimize_adam`` (``gpras/gpr.py:147-173``) as restated
in ``gpras_b200/gpr.py`` -- on the oracle-backed double and on the host-driven device path -- including the early-stopping rule.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda(lib):
    import torch

    assert torch.cuda.is_available() and lib.gpras_device_count() > 0, "GPU tests need a CUDA device"
    return torch


def _models_setup(n, d, m, p, seed, ard):
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(n, d, p, 0, seed=seed)
    rng = np.random.default_rng(seed + 1)
    theta = np.empty((p, 2 + d))
    z = np.empty((p, m, d))
    for b in range(p):
        ls = rng.uniform(1.0, 3.0, d) if ard else np.full(d, rng.uniform(1.0, 3.0))
        theta[b] = np.concatenate([[rng.uniform(0.5, 2.0), rng.uniform(0.05, 0.5)], ls])
        z[b] = data.x[rng.choice(n, m, replace=False)] + 0.05 * rng.standard_normal((m, d))
    return data, theta, z


@pytest.mark.parametrize("kernel,ard,n,d,m,p", [("Matern52", False, 700, 5, 24, 6), ("RBF", True, 1000, 10, 50, 10),
                                                ("Matern32", True, 257, 3, 128, 3), ("Matern12", False, 300, 20, 7, 2),
                                                ("Exponential", False, 130, 2, 1, 4), ("Matern52", True, 5000, 32, 64, 3),
                                                ("RBF", False, 400, 33, 40, 2), ("Matern32", False, 640, 6, 65, 2)])
def test_batched_evaluation_matches_the_single_model_one(cuda, kernel, ard, n, d, m, p):
    """General path (m > 64 or d > 32): the same kernels with the model index in the grid, bitwise the single-model results.
    Fused path (sgpr_fused.cuh): same formulas, other summation orders -- compared at 1e-10 / 1e-8; both against the oracle."""
    import torch

    from gpras_b200.engine import SparseBatch, SparseGP
    from oracle import sgpr

    data, theta, z = _models_setup(n, d, m, p, seed=n + m, ard=ard)
    batch = SparseBatch(kernel, n, d, m, p)
    batch.set_data(data.x, data.y)
    elbo, gt, gz, info = batch.elbo_grad(theta, z)
    assert not info.any()
    one = SparseGP(kernel, n, d, m, 1)
    fused = m <= 64 and d <= 32
    worst = [0.0, 0.0, 0.0]
    for b in range(p):
        one.set_data(data.x, np.ascontiguousarray(data.y[:, b : b + 1]))
        e1, gt1, gz1 = one.elbo_grad(theta[b], z[b])
        if not fused:
            assert e1 == elbo[b]
            np.testing.assert_array_equal(gt1, gt[b])
            np.testing.assert_array_equal(gz1, gz[b])
            continue
        worst[0] = max(worst[0], abs(e1 - elbo[b]) / abs(e1))
        worst[1] = max(worst[1], float(np.max(np.abs(gt1 - gt[b])) / max(1.0, np.abs(gt1).max())))
        worst[2] = max(worst[2], float(np.max(np.abs(gz1 - gz[b])) / max(1.0, np.abs(gz1).max())))
    one.close()
    if fused:
        print(f"fused vs general path ({kernel}, n={n}, d={d}, m={m}): elbo {worst[0]:.1e}, grad theta {worst[1]:.1e}, grad Z {worst[2]:.1e}")
        assert worst[0] < 1e-10 and worst[1] < 1e-8 and worst[2] < 1e-8
    # and against the oracle (first and last model)
    t = lambda a, g=False: torch.tensor(np.asarray(a, np.float64), requires_grad=g)  # noqa: E731
    for b in (0, p - 1):
        tv, tn, tl, tz = t(theta[b, 0], True), t(theta[b, 1], True), t(theta[b, 2:], True), t(z[b], True)
        e = sgpr.elbo(kernel, t(data.x), t(data.y[:, b : b + 1]), tz, tv, tl, tn)
        gv, gn, gl, gzz = torch.autograd.grad(e, [tv, tn, tl, tz])
        assert abs(elbo[b] - float(e.detach())) <= 1e-8 * abs(float(e.detach()))
        ref = np.concatenate([[float(gv) * theta[b, 0], float(gn) * theta[b, 1]], gl.numpy() * theta[b, 2:]])
        np.testing.assert_allclose(gt[b], ref, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(ref).max()))
        np.testing.assert_allclose(gz[b], gzz.numpy(), rtol=1e-6, atol=1e-7 * max(1.0, np.abs(gzz.numpy()).max()))
    # a second evaluation with other values on the same handle, and repeatability
    elbo2, gt2, _, _ = batch.elbo_grad(theta[::-1].copy(), z[::-1].copy())
    elbo3, gt3, _, _ = batch.elbo_grad(theta, z)
    assert np.array_equal(elbo3, elbo) and np.array_equal(gt3, gt) and not np.array_equal(elbo2, elbo)
    batch.close()


@pytest.mark.parametrize("kernel,ard,n,d,m,p,t", [("Matern52", False, 700, 5, 24, 6, 300), ("RBF", True, 1000, 10, 50, 10, 1),
                                                  ("Matern12", True, 300, 32, 64, 2, 129), ("Exponential", False, 130, 2, 1, 3, 70000)])
def test_batched_prediction_matches_the_single_model_one_and_the_oracle(cuda, kernel, ard, n, d, m, p, t):
    """``predict_y`` (``gpr.py:337``) of all models in one pass per tile of test inputs against one handle per model (the general
    kernels) and the torch oracle; T = 1, T not a tile multiple, and T beyond one staging chunk."""
    from gpras_b200.engine import SparseBatch, SparseGP
    from oracle import sgpr

    data, theta, z = _models_setup(n, d, m, p, seed=n + m + 1, ard=ard)
    xs = np.random.default_rng(t).standard_normal((t, d))
    batch = SparseBatch(kernel, n, d, m, p)
    batch.set_data(data.x, data.y)
    batch.condition(theta, z)
    mean, var = batch.predict(xs)
    assert mean.shape == var.shape == (t, p)
    one = SparseGP(kernel, n, d, m, 1)
    tt = min(t, 500)
    for b in range(p):
        one.set_data(data.x, np.ascontiguousarray(data.y[:, b : b + 1]))
        one.condition(theta[b], z[b])
        m1, v1 = one.predict(xs[:tt])
        np.testing.assert_allclose(mean[:tt, b], m1[:, 0], rtol=1e-10, atol=1e-11 * max(1.0, np.abs(m1).max()))
        np.testing.assert_allclose(var[:tt, b], v1[:, 0], rtol=1e-10)
    one.close()
    for b in (0, p - 1):
        om, ov = sgpr.predict_y(kernel, data.x, data.y[:, b : b + 1], z[b], theta[b, 0], theta[b, 2:], theta[b, 1], xs[-tt:])
        np.testing.assert_allclose(mean[-tt:, b], om[:, 0], rtol=1e-8, atol=1e-9 * max(1.0, np.abs(om).max()))
        np.testing.assert_allclose(np.sqrt(var[-tt:, b]), np.sqrt(ov[:, 0]), rtol=1e-6)
    # conditioning is remembered; an evaluation in between invalidates it
    calls = []
    orig = batch.lib.gpras_sgpr_batch_condition
    batch.lib.gpras_sgpr_batch_condition = lambda *a: (calls.append(1), orig(*a))[1]
    try:
        batch.condition(theta, z)
        assert not calls
        batch.elbo_grad(theta, z)
        batch.condition(theta, z)
        assert len(calls) == 1
    finally:
        batch.lib.gpras_sgpr_batch_condition = orig
    m2, v2 = batch.predict(xs[:tt])
    np.testing.assert_array_equal(m2, mean[:tt])
    np.testing.assert_array_equal(v2, var[:tt])
    batch.close()


def test_gpras_predict_uses_the_batch_for_the_reference_family(cuda):
    from gpras_b200 import GPRAS
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(600, 4, 5, 200, seed=13)
    g = GPRAS("Matern32")
    g.fit(data.x, data.y, 20, "grid", "adam", max_iter=10)
    mean, var = g.predict(data.x_test)
    g._opts["batched_predict"] = False
    mean1, var1 = g.predict(data.x_test)
    assert mean.shape == (200, 5)
    np.testing.assert_allclose(mean, mean1, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(var, var1, rtol=1e-10)
    # M > 64 falls back to one handle per model
    g2 = GPRAS("RBF")
    g2.fit(data.x, data.y, 70, "grid", "adam", max_iter=3)
    m2, v2 = g2.predict(data.x_test[:10])
    assert m2.shape == (10, 5) and np.all(v2 > 0)


def test_batch_rejects_what_it_cannot_hold(cuda):
    from gpras_b200._lib import GprasError
    from gpras_b200.engine import SparseBatch

    with pytest.raises(GprasError):
        SparseBatch("RBF", 300, 4, 129, 2)
    with pytest.raises(GprasError):
        SparseBatch("RBF", 300, 65, 10, 2)
    b = SparseBatch("RBF", 300, 4, 10, 2)
    with pytest.raises(GprasError):  # no data yet
        b.elbo_grad(np.ones((2, 6)), np.zeros((2, 10, 4)))
    with pytest.raises(ValueError):
        b.set_data(np.zeros((300, 4)), np.zeros((300, 3)))
    with pytest.raises(ValueError):
        b.adam(np.zeros((2, 5)), 1, True, True, 3)
    b.close()


@pytest.mark.parametrize("method,ard,param", [("two-stage", False, "softplus"), ("adam", False, "softplus"), ("adam", True, "softplus"),
                                              ("two-stage", True, "log"), ("adadelta", False, "softplus"), ("adadelta", True, "log"),
                                              ("three-stage", False, "softplus")])
def test_device_trainer_matches_oracle_backed_recipe(cuda, method, ard, param):
    """The reference's first-order recipes on its default model family, all columns at once on the device, against the recipe
    run model by model on the torch-oracle-backed double: hyperparameters and inducing inputs to 1e-6 (three-stage, whose last
    two stages are L-BFGS runs on the host: the north_star's 1e-4 for optima)."""
    from gpras_b200 import GPRAS, gpr
    from gpras_b200.synth import make_gp_data
    from test_host_cpu import OracleBackedSparseModel

    data = make_gp_data(220, 3, 3, 0, seed=23)
    g = GPRAS("Matern32")
    g.fit(data.x, data.y, 9, "grid", method, max_iter=25, ard=ard, parameterisation=param)
    rtol = 1e-4 if method == "three-stage" else 1e-6
    if method != "three-stage":
        assert all(mdl.n_evals == (50 if method == "two-stage" else 25) for mdl in g.models)
    assert all(hasattr(mdl, "adam_losses") for mdl in g.models)  # the device-resident loop ran
    z0 = g._create_inducing(data.x, 9, "grid")
    ls0 = np.full(3, np.mean(np.abs(data.x))) if ard else float(np.mean(np.abs(data.x)))
    for b, dev in enumerate(g.models):
        ref = OracleBackedSparseModel("Matern32", data.x, np.ascontiguousarray(data.y[:, b : b + 1]), z0.copy(), ls0, parameterisation=param)
        gpr.OPTIMIZERS[method](ref, max_iter=25)
        np.testing.assert_allclose(dev.theta(), ref.theta(), rtol=rtol)
        zr = np.asarray(ref.inducing_variable.Z)
        np.testing.assert_allclose(np.asarray(dev.inducing_variable.Z), zr, rtol=rtol, atol=rtol * np.abs(zr).max())
        assert dev.inducing_variable.trainable and all(p.trainable for p in dev.parameters)
    mean, var = g.predict(data.x[:10])
    assert mean.shape == (10, 3) and np.all(var > 0)


@pytest.mark.parametrize("pick_best,train_z", [(False, False), (True, False), (True, True)])
def test_multi_start_on_the_device_batch_equals_the_sequential_loop(cuda, pick_best, train_z):
    """"stochastic" recipe (``gpr.py:73-109``) on a sparse model: all starts as one device batch (same draws in the same order,
    coarse Adam stage device resident) against one start after the other on the host-driven path; then the same final L-BFGS."""
    from gpras_b200 import GPRAS
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(400, 4, 1, 0, seed=31)
    out = []
    for lock in (False, True):
        g = GPRAS("Matern52")
        g.fit(data.x, data.y, 16, "grid", "stochastic", lockstep_models=False, n_starts=7, iter_initial=12, iter_final=0, seed=6,
              pick_best=pick_best, train_z=train_z, lockstep=lock)
        mdl = g.models[0]
        out.append(np.concatenate([mdl.theta(), np.asarray(mdl.inducing_variable.Z).ravel()]))
    np.testing.assert_allclose(out[1], out[0], rtol=1e-9, atol=1e-12)


def test_device_trainer_follows_the_host_loop_step_by_step(cuda):
    """Loss history of the device-resident stage == the losses the host-driven ``_optimize_adam`` sees on the same device
    evaluation (priors included), for every model and step; without priors too."""
    from gpras_b200 import gpr
    from gpras_b200.sparse import SparseModel, _device_batch, adam_device
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(500, 4, 4, 0, seed=5)
    rng = np.random.default_rng(0)
    z0 = data.x[rng.choice(500, 30, replace=False)]
    for priors in (True, False):
        mk = lambda b: SparseModel("Matern52", data.x, np.ascontiguousarray(data.y[:, b : b + 1]), z0.copy(), 1.5, priors=priors)  # noqa: E731
        dev_models = [mk(b) for b in range(4)]
        batch = _device_batch(dev_models)
        assert batch is not None
        assert adam_device(dev_models, batch, 40)
        for b in range(4):
            host = mk(b)
            seen = []

            def recording(u, _orig=host.loss_and_grad, _seen=seen):
                loss, grad = _orig(u)
                _seen.append(loss)
                return loss, grad

            host.loss_and_grad = recording
            gpr._optimize_adam(host, 40)
            np.testing.assert_allclose(dev_models[b].adam_losses, np.array(seen), rtol=1e-12)
            np.testing.assert_allclose(dev_models[b].get_u(), host.get_u(), rtol=1e-9, atol=1e-12)


def test_device_trainer_early_stopping_rule(cuda):
    """``gpr.py:159-173``: a step is an improvement when (best - loss) / |loss| > 10e-6; more than 50 steps without one stop the
    loop -- per model.  With a vanishing learning rate only the first step improves: 52 steps, on the device as on the host."""
    from gpras_b200 import gpr
    from gpras_b200.sparse import SparseModel, _device_batch, _trainer_config
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(300, 3, 2, 0, seed=8)
    z0 = data.x[:12].copy()
    models = [SparseModel("RBF", data.x, np.ascontiguousarray(data.y[:, b : b + 1]), z0.copy(), 1.2) for b in range(2)]
    batch = _device_batch(models)
    _, transform, prior, floor, n_ls = _trainer_config(models[0])
    u0 = np.stack([m.get_u() for m in models])
    u, losses, iters = batch.adam(u0, n_ls, True, True, 80, learning_rate=1e-12, transform=transform, priors=prior is not None,
                                  noise_floor=floor)
    assert list(iters) == [52, 52]
    assert np.all(np.isfinite(losses[:52])) and np.all(np.isnan(losses[52:]))
    host = SparseModel("RBF", data.x, np.ascontiguousarray(data.y[:, :1]), z0.copy(), 1.2)
    gpr._optimize_adam(host, 80, learning_rate=1e-12)
    assert host.n_evals == 52
    # fewer steps than the patience: nobody stops
    u, losses, iters = batch.adam(u0, n_ls, True, True, 20, learning_rate=1e-12, transform=transform, priors=prior is not None,
                                  noise_floor=floor)
    assert list(iters) == [20, 20]
    # max_iter = 0 is a no-op
    u, losses, iters = batch.adam(u0, n_ls, True, True, 0)
    assert np.array_equal(u, u0) and list(iters) == [0, 0]


def test_device_trainer_reports_lost_positive_definiteness(cuda):
    from gpras_b200._lib import NotPositiveDefiniteError
    from gpras_b200.engine import SparseBatch
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(200, 2, 2, 0, seed=1)
    batch = SparseBatch("RBF", 200, 2, 8, 2)
    batch.set_data(data.x, data.y)
    u0 = np.zeros((2, 2 + 1 + 16))
    u0[:, 3:] = data.x[:8].ravel()
    u0[1, 0] = 800.0  # exp(800) = inf: Kuu of model 1 is not finite
    with pytest.raises(NotPositiveDefiniteError, match="model 1"):
        batch.adam(u0, 1, True, True, 5, transform="log", noise_floor=0.0)
    # the handle stays usable
    u0[1, 0] = 0.0
    u, losses, iters = batch.adam(u0, 1, True, True, 5, transform="log", noise_floor=0.0)
    assert list(iters) == [5, 5] and np.all(np.isfinite(losses))
    batch.close()


def test_reference_default_fit_time(cuda):
    """Reference-scale default call (N = 5000, D = 10, M = 50, 10 per-column models, two-stage Adam 100 + 100): the device trainer
    against the host-driven lock-step loop -- same parameters, and the time of each (printed; asserted only loosely)."""
    import time

    from gpras_b200 import GPRAS
    from gpras_b200.synth import make_gp_data

    data = make_gp_data(5000, 10, 10, 0, seed=3)
    out, secs = [], []
    for dev in (False, True, True):
        g = GPRAS("Matern52")
        t0 = time.perf_counter()
        g.fit(data.x, data.y, 50, "grid", "two-stage", device_trainer=dev)
        secs.append(time.perf_counter() - t0)
        out.append(np.concatenate([np.concatenate([m.theta(), np.asarray(m.inducing_variable.Z).ravel()]) for m in g.models]))
    err = float(np.max(np.abs(out[2] - out[0]) / np.maximum(np.abs(out[0]), 1e-3)))
    print(f"reference-default fit: host lock-step {secs[0]:.3f} s, device trainer {secs[1]:.3f} s (first call), {secs[2]:.3f} s "
          f"(handles warm); max rel. difference of the fitted parameters {err:.2e}")
    assert err < 1e-8
    assert secs[2] < secs[0]
