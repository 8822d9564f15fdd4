"""Generate the committed golden vectors for the exact-GP path from scikit-learn (independent implementation).

Run once, in the build container:   python tests/golden/make_golden.py

The reference (fema-ffrd/gpras) ships no tests or expected outputs and its GPflow/TensorFlow stack cannot be
installed here, so the pin is scikit-learn 1.9.0's ``GaussianProcessRegressor`` -- an independent
implementation of the exact-GP formulation named by BASELINE.json's north_star
(``sklearn/gaussian_process/_gpr.py:541-656``).  For each case we store inputs, fixed hyperparameters,
``log_marginal_likelihood(theta, eval_gradient=True)`` (gradient w.r.t. log theta, re-ordered to
[variance, noise, lengthscales...]) and ``predict(return_std=True)``.
"""
from pathlib import Path

import numpy as np
import sklearn
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

HERE = Path(__file__).resolve().parent
NU = {"Matern12": 0.5, "Matern32": 1.5, "Matern52": 2.5}
CASES = [
    # name, kernel, ard, N, D, P, T, seed
    ("rbf_iso", "RBF", False, 96, 4, 3, 17, 11),
    ("rbf_ard", "RBF", True, 150, 6, 2, 9, 12),
    ("m12_iso", "Matern12", False, 80, 3, 1, 13, 13),
    ("m32_ard", "Matern32", True, 130, 5, 4, 21, 14),
    ("m52_ard", "Matern52", True, 200, 8, 8, 33, 15),
    ("m52_iso_dup", "Matern52", False, 64, 2, 2, 5, 16),  # contains duplicated training rows
]


def main():
    out = {"sklearn_version": np.array(sklearn.__version__)}
    for name, kern, ard, n, d, p, t, seed in CASES:
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((n, d))
        if name.endswith("dup"):
            x[n // 2 :] = x[: n - n // 2]
            x[-3:] += 1e-3 * rng.standard_normal((3, d))
        w = rng.standard_normal((d, p)) / np.sqrt(d)
        y = np.sin(x @ w) + 0.1 * rng.standard_normal((n, p))
        xs = rng.standard_normal((t, d))
        variance, noise = float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.02, 0.3))
        ls = rng.uniform(0.8, 3.0, d) if ard else np.array([rng.uniform(0.8, 3.0)])
        base = RBF(length_scale=ls if ard else ls[0]) if kern == "RBF" else Matern(length_scale=ls if ard else ls[0], nu=NU[kern])
        gp = GaussianProcessRegressor(kernel=ConstantKernel(variance) * base + WhiteKernel(noise), alpha=0.0, optimizer=None)
        gp.fit(x, y)
        lml, g = gp.log_marginal_likelihood(gp.kernel_.theta, eval_gradient=True)
        # sklearn theta order: [constant, lengthscale(s), noise] -> ours [variance, noise, lengthscale(s)]
        g = np.concatenate([[g[0], g[-1]], g[1:-1]])
        mean, std = gp.predict(xs, return_std=True)
        mean = mean.reshape(t, p)
        for k, v in dict(x=x, y=y, xs=xs, variance=variance, noise=noise, ls=ls, lml=lml, grad_log=g, mean=mean, std=std).items():
            out[f"{name}.{k}"] = np.asarray(v)
        out[f"{name}.kernel"] = np.array(kern)
    np.savez_compressed(HERE / "exact_gp_sklearn.npz", **out)
    print("wrote", HERE / "exact_gp_sklearn.npz", "cases", [c[0] for c in CASES])


if __name__ == "__main__":
    main()
