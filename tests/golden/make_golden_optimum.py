"""Golden OPTIMA for the exact-GP path from scikit-learn's own fit (independent implementation, same start).

Run once, in the build container:   python tests/golden/make_golden_optimum.py [cfg1 cfg2]

BASELINE.json's north_star asks that "optimized hyperparameters land within 1e-4 relative of the reference's
optimum from the same starts", the reference CPU GPR being scikit-learn-style.  For BASELINE configs 1 and 2
(``gpras_b200.synth.CONFIGS``; the data is regenerated from the seed, its checksum is stored) this script
runs ``GaussianProcessRegressor.fit`` -- SciPy L-BFGS-B in LOG-theta space with scikit-learn's default
tolerances, no priors -- from the reference's initial values (variance 1, lengthscale mean|x|, noise 1;
``gpras/gpr.py:289,298``) and stores the start, the fitted theta*, LML* and the number of objective
evaluations.  ``tests/test_gpu_parity.py::test_optimum_matches_sklearn_fit`` drives the CUDA objective through
the same optimiser call (``fit(..., parameterisation="log", priors=False, bounds=...)``) and compares.

It also records what happens when BOTH parameterisations (log = scikit-learn, softplus = GPflow / the reference)
are run to convergence with the oracle objective: the default SciPy ``ftol`` stops either run on a plateau, which is
why round 1 saw a 9-nat gap between a softplus-space GPU run and scikit-learn's log-space run.
"""
import sys
import time
import zlib
from pathlib import Path

import numpy as np
import sklearn
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from gpras_b200.synth import CONFIGS, make_gp_data  # noqa: E402
from oracle.exact_gp import Objective  # noqa: E402

NU = {"Matern12": 0.5, "Matern32": 1.5, "Matern52": 2.5}
BOUNDS = dict(variance=(1e-5, 1e5), ls=(1e-5, 1e5), noise=(1e-8, 1e5))


def checksum(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a, np.float64).tobytes())


def run_case(name: str) -> dict:
    c = CONFIGS[name]
    n, d, p, kern, ard = c["n"], c["d"], c["p"], c["kernel"], c["ard"]
    data = make_gp_data(n, d, p, 64, seed=0)
    l0 = float(np.mean(np.abs(data.x)))
    ls0 = np.full(d, l0) if ard else l0
    base = RBF(ls0, BOUNDS["ls"]) if kern == "RBF" else Matern(ls0, BOUNDS["ls"], nu=NU[kern])
    k = ConstantKernel(1.0, BOUNDS["variance"]) * base + WhiteKernel(1.0, BOUNDS["noise"])
    sk = GaussianProcessRegressor(kernel=k, alpha=0.0, n_restarts_optimizer=0)
    nev = [0]
    orig = sk.log_marginal_likelihood

    def counted(*a, **kw):
        nev[0] += 1
        return orig(*a, **kw)

    sk.log_marginal_likelihood = counted
    t0 = time.perf_counter()
    sk.fit(data.x, data.y)
    fit_s = time.perf_counter() - t0
    th = np.exp(sk.kernel_.theta)  # [constant, lengthscale(s), noise]
    theta_star = np.concatenate([[th[0], th[-1]], th[1:-1]])  # ours: [variance, noise, lengthscale(s)]
    mean, std = sk.predict(data.x_test, return_std=True)
    out = {
        "n": n, "d": d, "p": p, "kernel": np.array(kern), "ard": ard, "seed": 0,
        "x_crc32": checksum(data.x), "y_crc32": checksum(data.y),
        "theta0": np.concatenate([[1.0, 1.0], np.atleast_1d(ls0)]),
        "theta_star": theta_star, "lml_star": float(sk.log_marginal_likelihood_value_), "n_evals": nev[0] - 1,
        "fit_seconds": fit_s, "pred_mean": mean.reshape(64, p), "pred_std": std.reshape(64, -1),
        "bounds_log": np.log(np.array([BOUNDS["variance"], BOUNDS["noise"]] + [BOUNDS["ls"]] * (d if ard else 1))),
    }
    # the same objective run to convergence (oracle port, tight tolerances) in both parameterisations
    from scipy.optimize import minimize

    for space in ("log", "softplus"):
        obj = Objective(kern, data.x, data.y, ard=ard, space=space, priors=False, fast=True)
        u0 = obj.unconstrain(out["theta0"])
        r_def = minimize(obj, u0, jac=True, method="L-BFGS-B", options={"maxiter": 1000})
        r_tight = minimize(obj, r_def.x, jac=True, method="L-BFGS-B",
                           options={"maxiter": 5000, "maxfun": 20000, "ftol": 1e-15, "gtol": 1e-9})
        out[f"oracle_{space}_default_theta"] = obj.constrain(r_def.x)
        out[f"oracle_{space}_default_lml"] = -float(r_def.fun)
        out[f"oracle_{space}_default_nfev"] = r_def.nfev
        out[f"oracle_{space}_converged_theta"] = obj.constrain(r_tight.x)
        out[f"oracle_{space}_converged_lml"] = -float(r_tight.fun)
        out[f"oracle_{space}_converged_nfev"] = r_def.nfev + r_tight.nfev
    return out


def main():
    which = sys.argv[1:] or ["cfg1", "cfg2"]
    path = HERE / "exact_gp_sklearn_optimum.npz"
    store = dict(np.load(path)) if path.exists() else {}
    store["sklearn_version"] = np.array(sklearn.__version__)
    for name in which:
        t0 = time.perf_counter()
        res = run_case(name)
        for k, v in res.items():
            store[f"{name}.{k}"] = np.asarray(v)
        print(name, "sklearn lml*", res["lml_star"], "evals", res["n_evals"], f"{time.perf_counter() - t0:.1f} s", flush=True)
        for space in ("log", "softplus"):
            print("   oracle", space, "default", res[f"oracle_{space}_default_lml"], res[f"oracle_{space}_default_nfev"], "converged",
                  res[f"oracle_{space}_converged_lml"], res[f"oracle_{space}_converged_nfev"], flush=True)
        np.savez_compressed(path, **store)
    print("wrote", path)


if __name__ == "__main__":
    main()
