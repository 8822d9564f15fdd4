"""CPU: host logic of the drop-in (parameters, recipes, persistence keys) and the C-ABI surface.
No compute call reaches the library here: without a GPU the ABI must refuse loudly, never fall back."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from gpras_b200 import _lib, gpr, sparse
from gpras_b200.synth import make_gp_data
from oracle.exact_gp import Objective, Theta, lml_and_grad

ROOT = Path(__file__).resolve().parents[1]


class OracleBackedModel(gpr.ExactModel):
    """ExactModel whose device evaluation is replaced by the CPU oracle -- lets the recipes, chain rule and
    restart sharding be tested on a box without a GPU.  Test infrastructure only."""

    class _Slot:
        def __init__(self, outer):
            self.outer = outer

        def acquire(self, model):
            return self

        def lml_grad(self, theta, want_grad=True):
            m = self.outer
            lml, gv, gn, gl = lml_and_grad(m.kernel.name, m.x, m.y, Theta(theta[0], theta[1], theta[2:]), want_grad=want_grad)
            return lml, (np.concatenate([[gv, gn], gl]) if want_grad else None)

    def __init__(self, kernel_name, x, y, lengthscales, priors=True, parameterisation="softplus"):
        super().__init__(kernel_name, x, y, lengthscales, slot=None, priors=priors, parameterisation=parameterisation)
        self._slot = OracleBackedModel._Slot(self)


class OracleBackedSparseModel(sparse.SparseModel):
    """SparseModel whose device evaluation is replaced by the torch-CPU SGPR oracle (autograd).  Test infrastructure only."""

    device_batch = False  # never joins a device batch: every evaluation must come from the oracle

    class _GP:
        def __init__(self, outer):
            self.outer = outer

        def elbo_grad(self, theta, z, jitter=1e-6, want_grad=True):
            import torch

            from oracle import sgpr

            m = self.outer
            t = lambda a, g=False: torch.tensor(np.asarray(a, np.float64), requires_grad=g)  # noqa: E731
            tv, tn, tl, tz = t(theta[0], want_grad), t(theta[1], want_grad), t(theta[2:], want_grad), t(z, want_grad)
            e = sgpr.elbo(m.kernel.name, t(m.x), t(m.y), tz, tv, tl, tn, jitter)
            if not want_grad:
                return float(e), None, None
            gv, gn, gl, gz = torch.autograd.grad(e, [tv, tn, tl, tz])
            return float(e.detach()), np.concatenate([[float(gv) * theta[0], float(gn) * theta[1]], gl.numpy() * theta[2:]]), gz.numpy()

    def _gp(self):
        return OracleBackedSparseModel._GP(self)


def test_module_surface_matches_reference():
    # gpras/gpr.py:21-41,206-214
    assert set(gpr.KERNEL_FACTORY) == {"Matern12", "Matern32", "Matern52", "RBF", "Linear", "Polynomial", "Periodic", "Exponential"}
    assert set(gpr.OPTIMIZERS) == {"two-stage", "three-stage", "adam", "adadelta", "L-BFGS-B", "stochastic", "diffential_evolution"}
    for name in ("GPRAS", "KernelType", "OptimizerType", "InductionInitializerType"):
        assert hasattr(gpr, name)
    with pytest.raises(KeyError):
        gpr.GPRAS("NoSuchKernel")
    g = gpr.GPRAS("RBF")
    assert g.kernel_str == "RBF" and g.models == [] and g.x is None and g.y is None


def test_parameter_transform_roundtrip_and_floor():
    p = gpr.Parameter(0.37)
    assert abs(p.numpy() - 0.37) < 1e-15
    q = gpr.Parameter(1.0, lower=gpr.NOISE_FLOOR)
    assert abs(q.numpy() - 1.0) < 1e-15
    q.unconstrained[:] = -50.0
    assert q.numpy() >= gpr.NOISE_FLOOR
    r = gpr.Parameter(np.array([0.5, 2.0, 7.0]))
    np.testing.assert_allclose(r.numpy(), [0.5, 2.0, 7.0], rtol=1e-15)
    r.assign(np.array([1.0, 1.0, 1.0]))
    np.testing.assert_allclose(r.numpy(), 1.0)


def test_loss_and_grad_matches_oracle_objective():
    d = make_gp_data(50, 3, 1, seed=1)
    for ard in (False, True):
        ls0 = np.full(3, 1.4) if ard else 1.4
        m = OracleBackedModel("Matern52", d.x, d.y, ls0)
        m.kernel.variance.assign(1.3)
        m.likelihood.variance.assign(0.07)
        loss, g = m.loss_and_grad()
        obj = Objective("Matern52", d.x, d.y, ard=ard, space="softplus", priors=True)
        u = obj.unconstrain(np.concatenate([[1.3, 0.07], np.atleast_1d(ls0)]))
        f, go = obj(u)
        assert abs(loss - f) <= 1e-12 * abs(f)
        np.testing.assert_allclose(g, go, rtol=1e-10, atol=1e-12)
        assert abs(m.training_loss() - f) <= 1e-12 * abs(f)


def test_recipes_decrease_the_loss_and_respect_trainable_sets():
    d = make_gp_data(60, 2, 1, seed=2)
    for method, kw in [("adam", dict(max_iter=30)), ("adadelta", dict(max_iter=5)), ("L-BFGS-B", dict(max_iter=50)),
                       ("two-stage", dict(max_iter=20)), ("three-stage", dict(max_iter=30)),
                       ("stochastic", dict(n_starts=2, iter_initial=3, iter_final=20, seed=0)),
                       ("diffential_evolution", dict(popsize=3, max_iter=2, seed=0))]:
        m = OracleBackedModel("RBF", d.x, d.y, float(np.mean(np.abs(d.x))))
        before = m.training_loss()
        gpr.OPTIMIZERS[method](m, **kw)
        after = m.training_loss()
        assert np.isfinite(after)
        if method not in ("adadelta",):
            assert after < before, method
        if method != "diffential_evolution":  # the reference leaves the hyperparameters non-trainable after DE (gpr.py:48-49)
            assert all(p.trainable for p in m.parameters), method


def test_adam_early_stopping_rule():
    # a flat objective stops after patience + 1 = 52 evaluations (gpr.py:159-173)
    class Flat:
        n = 0

        def get_u(self):
            return np.zeros(2)

        def set_u(self, u):
            pass

        def loss_and_grad(self, u=None):
            Flat.n += 1
            return 1.0, np.zeros(2)

    gpr._optimize_adam(Flat(), 1000)
    assert Flat.n == 52


def test_multi_start_last_start_wins_like_the_reference(monkeypatch):
    monkeypatch.setattr(gpr, "_optimize_bfgs", lambda model, max_iter: None)  # isolate the start selection
    d = make_gp_data(40, 2, 1, seed=3)
    starts = np.array([[1.0, 1.0, 0.1], [0.2, 3.0, 0.5]])
    m = OracleBackedModel("RBF", d.x, d.y, 1.0)
    gpr._optimize_multi_start(m, n_starts=2, iter_initial=0, iter_final=0, starts=starts)
    assert abs(m.kernel.variance.numpy() - 0.2) < 1e-12  # last start (reference quirk, gpr.py:86,96)
    m2 = OracleBackedModel("RBF", d.x, d.y, 1.0)
    gpr._optimize_multi_start(m2, n_starts=2, iter_initial=0, iter_final=0, starts=starts, pick_best=True)
    l0 = lml_and_grad("RBF", d.x, d.y, Theta(1.0, 0.1, 1.0), want_grad=False)[0]
    l1 = lml_and_grad("RBF", d.x, d.y, Theta(0.2, 0.5, 3.0), want_grad=False)[0]
    assert abs(m2.kernel.variance.numpy() - (1.0 if l0 > l1 else 0.2)) < 1e-12


def test_create_inducing_grid_and_kmeans():
    g = gpr.GPRAS("RBF")
    x = np.random.default_rng(0).standard_normal((100, 3))
    z = g._create_inducing(x, 7, "grid")
    assert z.shape == (7, 3)
    np.testing.assert_allclose(z[0], x.min(axis=0))
    np.testing.assert_allclose(z[-1], x.max(axis=0))
    zk = g._create_inducing(x, 5, "kmeans")
    assert zk.shape == (5, 3) and zk.dtype == np.float64


def test_unsupported_kernels_fail_like_the_reference():
    g = gpr.GPRAS("Linear")
    with pytest.raises(NotImplementedError):
        g._init_models(np.zeros((4, 2)), np.zeros((4, 1)), None)


# ---- C ABI surface ---------------------------------------------------------------------------
def test_header_symbols_are_exported_and_bound(lib):
    header = (ROOT / "include" / "gpras_b200.h").read_text()
    declared = set(re.findall(r"\b(gpras_[a-z0-9_]+)\s*\(", header))
    declared.discard("gpras_gp")
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert declared <= exported, declared - exported
    assert lib.gpras_abi_version() == _lib.ABI_VERSION == 4


def test_library_contains_dmma_and_no_cpu_fallback(lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass and "sm_100a" in sass
    if lib.gpras_device_count() == 0:
        h = C.c_void_p()
        assert lib.gpras_gp_create(C.byref(h), 0, 0, 8, 2, 1) == -2  # GPRAS_E_CUDA, loudly
        assert b"no CPU fallback" in lib.gpras_last_error()
        from gpras_b200.engine import ExactGP

        with pytest.raises(_lib.GprasError):
            ExactGP("RBF", 8, 2, 1)


def test_abi_argument_validation(lib):
    h = C.c_void_p()
    assert lib.gpras_gp_create(C.byref(h), 0, 9, 8, 2, 1) == -1
    assert lib.gpras_gp_create(C.byref(h), 0, 0, 0, 2, 1) == -1
    assert lib.gpras_gp_create(C.byref(h), 0, 0, 8, 65, 1) == -1
    assert lib.gpras_dgemm_tiles(None, 0, 0, 0, None, 0, None, 0, None, 0, 100, 128, 16, 1.0, 0.0) == -1
    assert lib.gpras_dpotrf(None, None, 0, None, 0, 100, None, None) == -1


@pytest.mark.parametrize("pick_best", [False, True])
def test_multi_start_lockstep_equals_sequential(pick_best):
    """All starts advancing together (batched evaluations) must follow the one-after-the-other trajectories exactly."""
    d = make_gp_data(40, 2, 2, seed=9)
    res = []
    for lock in (False, True):
        m = OracleBackedModel("Matern32", d.x, d.y, 1.0)
        gpr._optimize_multi_start(m, n_starts=5, iter_initial=6, iter_final=8, seed=3, pick_best=pick_best, lockstep=lock)
        res.append(np.concatenate([m.theta(), [m.training_loss()]]))
    np.testing.assert_array_equal(res[0], res[1])


# ---- round-2 host logic ----------------------------------------------------------------------
def test_log_parameterisation_matches_sklearn_style_objective():
    """parameterisation="log" is scikit-learn's variable: theta = exp(u), no floor on the noise, no priors."""
    d = make_gp_data(50, 3, 2, seed=4)
    for ard in (False, True):
        ls0 = np.full(3, 1.4) if ard else 1.4
        m = OracleBackedModel("Matern32", d.x, d.y, ls0, priors=False, parameterisation="log")
        m.kernel.variance.assign(0.8)
        m.likelihood.variance.assign(0.03)
        np.testing.assert_allclose(m.get_u(), np.log(np.concatenate([[0.8, 0.03], np.atleast_1d(ls0)])), rtol=1e-15)
        loss, g = m.loss_and_grad()
        obj = Objective("Matern32", d.x, d.y, ard=ard, space="log", priors=False)
        f, go = obj(np.log(np.concatenate([[0.8, 0.03], np.atleast_1d(ls0)])))
        assert abs(loss - f) <= 1e-13 * abs(f)
        np.testing.assert_allclose(g, go, rtol=1e-12, atol=1e-13)
    # same optimum from the same start as the oracle-driven SciPy run, bounds passed through
    m = OracleBackedModel("RBF", d.x, d.y, 1.0, priors=False, parameterisation="log")
    bounds = [(-11.5, 11.5)] * 3
    gpr._optimize_bfgs(m, 200, bounds=bounds)
    from scipy.optimize import minimize

    obj = Objective("RBF", d.x, d.y, ard=False, space="log", priors=False)
    r = minimize(obj, np.zeros(3), jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": 200})
    np.testing.assert_allclose(m.get_u(), r.x, rtol=1e-10)


def test_bfgs_survives_a_non_positive_definite_trial_point():
    class Wall(OracleBackedModel):
        calls = 0

        def loss_and_grad(self, u=None):
            Wall.calls += 1
            if Wall.calls == 2:  # the first line-search trial point
                raise _lib.NotPositiveDefiniteError("pivot 3")
            return super().loss_and_grad(u)

    d = make_gp_data(40, 2, 1, seed=2)
    m = Wall("RBF", d.x, d.y, 1.0)
    before = m.training_loss()
    res = gpr._optimize_bfgs(m, 30)
    assert np.isfinite(res.fun) and m.training_loss() < before


def test_run_restarts_scores_bad_starts_inf_and_keeps_z_with_theta():
    from gpras_b200 import parallel

    d = make_gp_data(60, 2, 1, seed=5)
    z0 = d.x[:6].copy()
    m = OracleBackedSparseModel("RBF", d.x, d.y, z0.copy(), 1.0)
    starts = np.array([[1.0, 0.1, 1.0], [0.5, 0.3, 2.0], [2.0, 0.05, 0.6]])
    seen_z, seen_flags = [], []

    def recipe(model, max_iter):
        seen_z.append(np.array(model.inducing_variable.Z))
        seen_flags.append((model.inducing_variable.trainable, [p.trainable for p in model.parameters]))
        if len(seen_z) == 2:
            raise _lib.NotPositiveDefiniteError("pivot 1")
        gpr._optimize_two_stage(model, max_iter)
        model.set_trainable(False)  # like the reference's DE recipe, which leaves the hyperparameters frozen (gpr.py:48-49)

    table = parallel.run_restarts(m, recipe, starts, dict(max_iter=4))
    for z, fl in zip(seen_z, seen_flags):  # every restart begins from the initial Z and trainable flags
        np.testing.assert_array_equal(z, z0)
        assert fl == (True, [True, True, True])
    assert np.isinf(table[1, 1]) and np.all(np.isfinite(table[[0, 2], 1]))
    best = int(np.argmin(table[:, 1]))
    assert best != 1
    # the model ends with the winner's theta AND the winner's Z: its loss is the reported one
    np.testing.assert_allclose(m.theta()[:3], table[best, 2:5], rtol=1e-14)
    assert not np.array_equal(m.inducing_variable.Z, z0)
    assert all(p.trainable for p in m.parameters) and m.inducing_variable.trainable  # flags restored after the last restart
    m.set_trainable(False)  # the table's losses were taken in the state the recipe left (priors count on trainable parameters only)
    np.testing.assert_allclose(m.training_loss(), table[best, 1], rtol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        def always_bad(model, **kw):
            raise _lib.NotPositiveDefiniteError("pivot 1")
        parallel.run_restarts(m, always_bad, starts, {})


def test_stochastic_freezes_redrawn_z_like_the_reference():
    """gpr.py:91 replaces the GPflow Parameter by a raw array: after the redraw Z is no longer trained."""
    d = make_gp_data(50, 2, 1, seed=6)
    m = OracleBackedSparseModel("Matern52", d.x, d.y, d.x[:5].copy(), 1.0)
    drawn = []
    orig = gpr._optimize_adam

    def spy(model, max_iter, *a, **k):
        drawn.append(np.array(model.inducing_variable.Z))
        return orig(model, max_iter, *a, **k)

    gpr._optimize_adam, keep = spy, gpr._optimize_adam
    try:
        gpr._optimize_multi_start(m, n_starts=2, iter_initial=2, iter_final=3, seed=1)
    finally:
        gpr._optimize_adam = keep
    np.testing.assert_array_equal(m.inducing_variable.Z, drawn[-1])  # neither Adam nor L-BFGS moved the last draw
    m2 = OracleBackedSparseModel("Matern52", d.x, d.y, d.x[:5].copy(), 1.0)
    gpr._optimize_multi_start(m2, n_starts=2, iter_initial=2, iter_final=3, seed=1, train_z=True)
    assert not np.array_equal(m2.inducing_variable.Z, drawn[-1])


def test_pool_cache_is_bounded():
    assert gpr._POOL_CACHE_MAX_SHAPES <= 8
    gpr.release_pools()
    assert len(gpr._POOL_CACHE) == 0


def test_run_restarts_lanes_match_single_lane():
    """Two restarts in flight per rank (host threads on clones of the model) give the table of the one-at-a-time loop."""
    from gpras_b200 import parallel

    d = make_gp_data(50, 2, 2, seed=7)
    starts = np.array([[1.0, 0.1, 1.0], [0.3, 0.5, 2.5], [2.0, 0.05, 0.7], [0.7, 0.2, 1.3], [1.5, 0.4, 0.9]])
    tabs = []
    for lanes in (1, 2, 3):
        m = OracleBackedModel("Matern52", d.x, d.y, 1.0)
        tabs.append(parallel.run_restarts(m, gpr.OPTIMIZERS["L-BFGS-B"], starts, dict(max_iter=25), lanes=lanes))
        best = int(np.argmin(tabs[-1][:, 1]))
        np.testing.assert_allclose(m.theta()[:3], tabs[-1][best, 2:5], rtol=1e-14)
        assert m.n_evals > 0
    np.testing.assert_array_equal(tabs[0], tabs[1])
    np.testing.assert_array_equal(tabs[0], tabs[2])


def test_fit_routes_first_order_recipes_of_sparse_models_to_the_batched_trainer(monkeypatch):
    """Host logic of ``GPRAS.fit`` (no device needed): which calls go to ``sparse.fit_lockstep`` (all per-column models together),
    with which options, and what happens when it declines."""
    from gpras_b200 import sparse as sp

    data = make_gp_data(60, 3, 4, seed=1)
    calls, sequential = [], []

    def fake_lockstep(models, method, max_iter=100, device_trainer=True):
        calls.append((len(models), method, max_iter, device_trainer))
        return method != "adadelta"  # decline one recipe: fit must then run it model by model

    monkeypatch.setattr(sp, "fit_lockstep", fake_lockstep)
    for name in list(gpr.OPTIMIZERS):
        monkeypatch.setitem(gpr.OPTIMIZERS, name, lambda model, _n=name, **kw: sequential.append((_n, kw)))

    def run(method, **kw):
        calls.clear()
        sequential.clear()
        g = gpr.GPRAS("Matern32")
        g.fit(data.x, data.y, 7, "grid", method, **kw)
        return list(calls), list(sequential)

    c, s_ = run("two-stage")
    assert c == [(4, "two-stage", 100, True)] and not s_
    c, s_ = run("adam", max_iter=12)
    assert c == [(4, "adam", 12, True)] and not s_
    c, s_ = run("three-stage", max_iter=9)
    assert c == [(4, "three-stage", 9, True)] and not s_
    c, s_ = run("adadelta", max_iter=5)  # declined -> sequential loop over the four models
    assert c == [(4, "adadelta", 5, True)] and s_ == [("adadelta", {"max_iter": 5})] * 4
    c, s_ = run("adam")  # no max_iter: the reference's _optimize_adam has no default either -> the recipe itself is called
    assert not c and len(s_) == 4
    c, s_ = run("L-BFGS-B", max_iter=10)
    assert not c and s_ == [("L-BFGS-B", {"max_iter": 10})] * 4
    c, s_ = run("two-stage", lockstep_models=False)
    assert not c and len(s_) == 4
    c, s_ = run("adadelta", max_iter=5, device_trainer=False)  # host lock-step knows Adam only
    assert not c and len(s_) == 4
    c, s_ = run("two-stage", device_trainer=False)
    assert c == [(4, "two-stage", 100, False)] and not s_
    c, s_ = run("two-stage", n_jobs=2)
    assert not c and len(s_) == 4


def test_sparse_models_qualify_for_a_device_batch_only_when_configured_alike():
    from gpras_b200 import sparse as sp

    data = make_gp_data(40, 2, 3, seed=2)
    z = data.x[:6].copy()
    mk = lambda j, **kw: sp.SparseModel("RBF", data.x, np.ascontiguousarray(data.y[:, j : j + 1]), z.copy(), 1.0, **kw)  # noqa: E731
    a, b = mk(0), mk(1)
    assert sp._trainer_config(a) == ("RBF", "softplus", "LogNormal(0,1)", gpr.NOISE_FLOOR, 1) == sp._trainer_config(b)
    assert sp._trainer_config(mk(0, priors=False))[2] is None
    assert sp._trainer_config(mk(0, parameterisation="log"))[1:4] == ("log", "LogNormal(0,1)", 0.0)
    a.kernel.variance.prior = None  # mixed prior settings: the device trainer applies one setting to all three hyperparameters
    assert sp._trainer_config(a) is None
    # test doubles never join a device batch; neither do models that differ in configuration (checked before any device call)
    assert sp._device_batch([OracleBackedSparseModel("RBF", data.x, data.y[:, :1], z.copy(), 1.0)]) is None
    assert sp._device_batch([mk(0), mk(1, parameterisation="log")]) is None
    assert sp._device_batch([mk(0), sp.SparseModel("RBF", data.x, np.ascontiguousarray(data.y[:, 1:2]), z[:5].copy(), 1.0)]) is None
