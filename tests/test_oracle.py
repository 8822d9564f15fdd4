"""CPU: the oracle against the scikit-learn golden vectors, finite differences and algebraic identities
(SURVEY.md section 8c known-answer list).  No CUDA involved."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_case
from gpras_b200.synth import fixed_theta, make_cell_map, make_gp_data
from oracle import sgpr
from oracle.cells import reverse_transform
from oracle.exact_gp import Objective, Theta, fit_lbfgs, lml_and_grad, lml_and_grad_fast, predict
from oracle.kernels import KERNEL_NAMES, cov


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_sklearn_golden(golden, name):
    c = golden_case(golden, name)
    th = Theta(c["variance"], c["noise"], c["ls"])
    lml, gv, gn, gl = lml_and_grad(c["kernel"], c["x"], c["y"], th)
    g = np.concatenate([[gv, gn], gl if c["ls"].size > 1 else [gl.sum()]])
    assert abs(lml - c["lml"]) <= 1e-10 * abs(c["lml"])
    np.testing.assert_allclose(g, c["grad_log"], rtol=1e-8, atol=1e-9)
    mean, var = predict(c["kernel"], c["x"], c["y"], th, c["xs"])
    np.testing.assert_allclose(mean, c["mean"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.sqrt(var), c["std"].reshape(var.shape), rtol=1e-8)


@pytest.mark.parametrize("name", KERNEL_NAMES)
@pytest.mark.parametrize("ard", [False, True])
def test_gradient_matches_central_differences(name, ard):
    d = make_gp_data(60, 3, 2, seed=5)
    v, s, ls = fixed_theta(3, ard)
    th = Theta(v, s, ls)
    _, gv, gn, gl = lml_and_grad(name, d.x, d.y, th)
    eps = 1e-6

    def f(v_, s_, ls_):
        return lml_and_grad(name, d.x, d.y, Theta(v_, s_, ls_), want_grad=False)[0]

    fd_v = (f(v * np.exp(eps), s, ls) - f(v * np.exp(-eps), s, ls)) / (2 * eps)
    fd_s = (f(v, s * np.exp(eps), ls) - f(v, s * np.exp(-eps), ls)) / (2 * eps)
    assert abs(fd_v - gv) <= 1e-6 * max(1.0, abs(gv))
    assert abs(fd_s - gn) <= 1e-6 * max(1.0, abs(gn))
    full = np.full(3, ls[0]) if not ard else ls
    for j in range(3):
        up, dn = full.copy(), full.copy()
        up[j] *= np.exp(eps)
        dn[j] *= np.exp(-eps)
        fd = (f(v, s, up) - f(v, s, dn)) / (2 * eps)
        assert abs(fd - gl[j]) <= 1e-6 * max(1.0, abs(gl[j]))


def test_fast_form_matches_direct_form():
    d = make_gp_data(200, 6, 4, seed=3)
    v, s, ls = fixed_theta(6, True)
    a = lml_and_grad("Matern52", d.x, d.y, Theta(v, s, ls))
    b = lml_and_grad_fast("Matern52", d.x, d.y, Theta(v, s, ls))
    assert abs(a[0] - b[0]) <= 1e-12 * abs(a[0])
    np.testing.assert_allclose(b[3], a[3], rtol=1e-8)
    np.testing.assert_allclose([b[1], b[2]], [a[1], a[2]], rtol=1e-10)


def test_exponential_is_matern12_with_doubled_lengthscale():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((30, 4))
    np.testing.assert_allclose(cov("Exponential", x, x, 1.7, 0.9), cov("Matern12", x, x, 1.7, 1.8), rtol=1e-14)


def test_interpolation_limit():
    d = make_gp_data(50, 2, 2, seed=9)
    th = Theta(1.0, 1e-8, np.array([1.5]))
    mean, var = predict("Matern12", d.x, d.y, th, d.x)
    np.testing.assert_allclose(mean, d.y, atol=1e-5)
    assert np.all(np.abs(var) < 1e-6)


def test_objective_softplus_and_log_reach_same_optimum():
    d = make_gp_data(80, 2, 2, seed=4)
    res = {}
    for space in ("softplus", "log"):
        obj = Objective("RBF", d.x, d.y, ard=False, space=space, priors=False)
        r = fit_lbfgs(obj, obj.unconstrain(np.array([1.0, 0.5, 1.0])), max_iter=500)
        res[space] = obj.constrain(r.x)
    np.testing.assert_allclose(res["softplus"], res["log"], rtol=2e-4)


def test_objective_gradient_chain_rule():
    d = make_gp_data(40, 3, 1, seed=8)
    obj = Objective("Matern32", d.x, d.y, ard=True, space="softplus", priors=True)
    u = obj.unconstrain(np.array([1.3, 0.05, 1.0, 2.0, 0.7]))
    f0, g = obj(u)
    for j in range(u.size):
        e = np.zeros_like(u)
        e[j] = 1e-6
        fd = (obj(u + e)[0] - obj(u - e)[0]) / 2e-6
        assert abs(fd - g[j]) <= 1e-5 * max(1.0, abs(g[j]))


# ---- SGPR restatement (GPflow semantics): algebraic pins -------------------------------------
def _sgpr_setup(n=70, d=3, m=12, seed=2):
    data = make_gp_data(n, d, 1, seed=seed)
    rng = np.random.default_rng(seed)
    z = data.x[rng.choice(n, m, replace=False)]
    return data, z


def test_sgpr_collapses_to_exact_lml_when_z_is_x():
    import torch

    data, _ = _sgpr_setup()
    t = lambda a: torch.tensor(np.asarray(a, np.float64))  # noqa: E731
    var, ls, noise = 1.3, np.array([1.7]), 0.05
    e = float(sgpr.elbo("Matern52", t(data.x), t(data.y), t(data.x), t(var), t(ls), t(noise), jitter=0.0))
    lml = lml_and_grad("Matern52", data.x, data.y, Theta(var, noise, ls), want_grad=False)[0]
    assert abs(e - lml) <= 1e-9 * abs(lml)


def test_sgpr_bound_below_lml_and_monotone_under_nested_z():
    import torch

    data, _ = _sgpr_setup(n=90)
    t = lambda a: torch.tensor(np.asarray(a, np.float64))  # noqa: E731
    var, ls, noise = 1.1, np.array([1.4]), 0.1
    lml = lml_and_grad("RBF", data.x, data.y, Theta(var, noise, ls), want_grad=False)[0]
    prev = -np.inf
    for m in (5, 15, 45, 90):
        e = float(sgpr.elbo("RBF", t(data.x), t(data.y), t(data.x[:m]), t(var), t(ls), t(noise)))
        assert e <= lml + 1e-6 * abs(lml)
        assert e >= prev - 1e-8 * abs(e)
        prev = e


def _sklearn_kernel(name, variance, ls):
    """scikit-learn's implementation of the same covariance functions: independent of oracle/kernels.py and oracle/sgpr.py."""
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

    base = {"RBF": lambda: RBF(ls), "Matern12": lambda: Matern(ls, nu=0.5), "Matern32": lambda: Matern(ls, nu=1.5),
            "Matern52": lambda: Matern(ls, nu=2.5), "Exponential": lambda: Matern(2.0 * np.asarray(ls), nu=0.5)}[name]()
    return ConstantKernel(variance) * base


@pytest.mark.parametrize("name,ard", [("RBF", False), ("Matern12", True), ("Matern32", False), ("Matern52", True), ("Exponential", False)])
def test_sgpr_bound_and_prediction_match_the_published_dense_form(name, ard):
    """Pin of the restatement to the published definition (Titsias 2009, eqs. 6 and 9; GPflow's SGPR docstring):

        ELBO = log N(y | 0, Qff + s2 I) - tr(Kff - Qff) / (2 s2),      Qff = Kfu Kuu^-1 Kuf,
        q(f*) : mean = K*u S Kuf y / s2,  cov = K** - K*u Kuu^-1 Ku* + K*u S Ku*,   S = (Kuu + Kuf Kfu / s2)^-1,

    evaluated densely (N x N matrices, explicit inverses, SciPy's multivariate normal) with scikit-learn's kernels, i.e.
    sharing no code and no algebra with ``oracle/sgpr.py``'s Cholesky form.  Kuu carries GPflow's jitter in both."""
    import torch
    from scipy.stats import multivariate_normal

    data, z = _sgpr_setup(n=120, d=3, m=15, seed=5)
    xs = make_gp_data(40, 3, 1, seed=9).x
    var, noise = 1.3, 0.07
    ls = np.array([1.1, 1.9, 0.8]) if ard else np.array([1.4])
    k = _sklearn_kernel(name, var, ls if ard else float(ls[0]))
    n, m = data.x.shape[0], z.shape[0]
    kff, kuf, kuu = k(data.x), k(z, data.x), k(z) + sgpr.JITTER * np.eye(m)
    qff = kuf.T @ np.linalg.solve(kuu, kuf)
    dense = multivariate_normal(mean=np.zeros(n), cov=qff + noise * np.eye(n)).logpdf(data.y[:, 0]) - np.trace(kff - qff) / (2 * noise)
    t = lambda a: torch.tensor(np.asarray(a, np.float64))  # noqa: E731
    e = float(sgpr.elbo(name, t(data.x), t(data.y), t(z), t(var), t(ls), t(noise)))
    assert abs(e - dense) <= 1e-9 * abs(dense)
    sig = np.linalg.inv(kuu + kuf @ kuf.T / noise)
    ksu = k(xs, z)
    mean = ksu @ sig @ kuf @ data.y / noise
    cov = k(xs) - ksu @ np.linalg.solve(kuu, ksu.T) + ksu @ sig @ ksu.T
    om, ov = sgpr.predict_y(name, data.x, data.y, z, var, ls, noise, xs)
    np.testing.assert_allclose(om, mean, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(ov[:, 0], np.diag(cov) + noise, rtol=1e-7)


def test_sgpr_autograd_matches_finite_differences():
    data, z = _sgpr_setup()
    u = dict(u_var=0.3, u_ls=np.array([0.9]), u_noise=-1.0)
    out = sgpr.training_loss_and_grads("Matern32", data.x, data.y, z, **u)
    eps = 1e-6
    for key in ("u_var", "u_noise"):
        up, dn = dict(u), dict(u)
        up[key] = u[key] + eps
        dn[key] = u[key] - eps
        fd = (sgpr.training_loss_and_grads("Matern32", data.x, data.y, z, **up)["loss"]
              - sgpr.training_loss_and_grads("Matern32", data.x, data.y, z, **dn)["loss"]) / (2 * eps)
        assert abs(fd - float(out[key])) <= 1e-5 * max(1.0, abs(fd))
    zp, zm = z.copy(), z.copy()
    zp[3, 1] += eps
    zm[3, 1] -= eps
    fd = (sgpr.training_loss_and_grads("Matern32", data.x, data.y, zp, **u)["loss"]
          - sgpr.training_loss_and_grads("Matern32", data.x, data.y, zm, **u)["loss"]) / (2 * eps)
    assert abs(fd - out["z"][3, 1]) <= 1e-5 * max(1.0, abs(fd))


def test_sgpr_predict_matches_exact_predict_when_z_is_x():
    data, _ = _sgpr_setup()
    var, ls, noise = 0.9, np.array([2.0]), 0.2
    xs = np.random.default_rng(1).standard_normal((11, 3))
    m1, v1 = sgpr.predict_y("RBF", data.x, data.y, data.x, var, ls, noise, xs, jitter=0.0)
    m2, v2 = predict("RBF", data.x, data.y, Theta(var, noise, ls), xs)
    np.testing.assert_allclose(m1, m2, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(v1, v2, rtol=1e-7)


# ---- modes -> cells --------------------------------------------------------------------------
def test_reverse_transform_roundtrip_is_rank_p_projector():
    p, c = 5, 300
    cm = make_cell_map(p, c, seed=0)
    rng = np.random.default_rng(0)
    modes = rng.standard_normal((7, p))
    cells = reverse_transform(modes, None, cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    wet = ~cm.dry_indices
    # forward transform of the reference (preprocess.py:1009-1039) recovers the modes exactly (orthonormal EOF rows)
    back = (((cells[:, wet] - cm.input_mean) * cm.weights) @ cm.eofs.T - cm.x_mean) / cm.x_std
    np.testing.assert_allclose(back, modes, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cells[:, cm.dry_indices], np.broadcast_to(cm.elevations[cm.dry_indices], (7, int(cm.dry_indices.sum()))))
    var = rng.uniform(0.1, 1.0, (7, p))
    _, vc = reverse_transform(modes, var, cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    assert np.all(vc[:, cm.dry_indices] == 0.0) and np.all(vc[:, wet] > 0.0)
