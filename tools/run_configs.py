"""BASELINE.json configs 1, 2, 4 end to end on one GPU next to the CPU baselines (scikit-learn / oracle port).
Writes one JSON line per config.  Not part of the test-suite; results are copied into profiles/."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gpras_b200 import GPRAS
from gpras_b200.engine import ExactGP
from gpras_b200.synth import CONFIGS, make_gp_data, fixed_theta

which = sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]
for name in which:
    c = CONFIGS[name]
    n, d, p, t, kern, ard = c["n"], c["d"], c["p"], c["t"], c["kernel"], c["ard"]
    data = make_gp_data(n, d, p, min(t, 10000), seed=0)
    out = {"config": name, "n": n, "d": d, "p": p, "t": min(t, 10000), "kernel": kern, "ard": ard}
    # single LML+grad evaluation
    gp = ExactGP(kern, n, d, p)
    gp.set_data(data.x, data.y)
    v, s, ls = fixed_theta(d, ard)
    th = gp.theta_vector(v, s, ls)
    for _ in range(3):
        gp.lml_grad(th)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        gp.lml_grad(th)
    out["gpu_eval_ms"] = (time.perf_counter() - t0) / reps * 1e3
    gp.set_stage_timing(True); gp.lml_grad(th); out["gpu_stage_ms"] = gp.last_stage_ms(); gp.set_stage_timing(False)
    out["launches_per_eval"] = gp.last_launches()
    gp.close()
    if name in ("cfg1", "cfg2"):
        # GPU: shared-kernel exact GP, L-BFGS-B from the reference's initial values, then predict
        g = GPRAS(kern)
        t0 = time.perf_counter()
        g.fit(data.x, data.y, None, "kmeans", "L-BFGS-B", ard=ard, shared_kernel=True, priors=False, max_iter=200)
        out["gpu_fit_s"] = time.perf_counter() - t0
        out["gpu_fit_evals"] = g.models[0].n_evals
        t0 = time.perf_counter()
        mean, var = g.predict(data.x_test)
        out["gpu_predict_s"] = time.perf_counter() - t0
        m = g.models[0]
        # CPU: scikit-learn GaussianProcessRegressor, same objective (no priors, log-space L-BFGS-B), same start
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel
        l0 = float(np.mean(np.abs(data.x)))
        ls0 = np.full(d, l0) if ard else l0
        base = RBF(ls0, (1e-5, 1e5)) if kern == "RBF" else Matern(ls0, (1e-5, 1e5), nu={"Matern12": 0.5, "Matern32": 1.5, "Matern52": 2.5}[kern])
        k = ConstantKernel(1.0, (1e-5, 1e5)) * base + WhiteKernel(1.0, (1e-8, 1e5))
        sk = GaussianProcessRegressor(kernel=k, alpha=0.0, n_restarts_optimizer=0)
        t0 = time.perf_counter(); sk.fit(data.x, data.y); out["sklearn_fit_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); sm, ss = sk.predict(data.x_test, return_std=True); out["sklearn_predict_s"] = time.perf_counter() - t0
        out["sklearn_lml"] = float(sk.log_marginal_likelihood_value_)
        gpx = ExactGP(kern, n, d, p); gpx.set_data(data.x, data.y)
        out["gpu_lml_at_gpu_optimum"] = gpx.lml_grad(m.theta(), want_grad=False)[0]; gpx.close()
        out["pred_mean_max_abs_diff"] = float(np.max(np.abs(mean - sm.reshape(mean.shape))))
        out["cpu_cores"] = os.cpu_count()
    print(json.dumps(out), flush=True)
