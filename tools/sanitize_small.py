"""Small instances of every device path, for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from gpras_b200.cells import fold_cell_map
from gpras_b200.engine import ExactGP, SparseGP
from gpras_b200.metrics import MetricsAccumulator, fi_aoi_toi, summarise
from gpras_b200.preprocess import PreProcessor
from gpras_b200.synth import fixed_theta, make_cell_map, make_gp_data

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n, d, p, c, t = 384, 5, 6, 700, 150
data = make_gp_data(n, d, p, t, seed=0)
v, s, ls = fixed_theta(d, True)
if which in ("all", "gp"):
    gp = ExactGP("Matern52", n, d, p)
    gp.set_data(data.x, data.y)
    th = gp.theta_vector(v, s, ls)
    print("lml", gp.lml_grad(th)[0], gp.lml_grad(th)[0])
    gp.condition(th)
    m, vv = gp.predict(data.x_test)
    cm = make_cell_map(p, c, seed=0)
    e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    gp.set_cell_map(e_mean, bias)
    gp.predict_cells(data.x_test)
    acc = MetricsAccumulator(c, t)
    acc.set_elevations(cm.elevations, cm.elevations)
    acc.reset(0.1)
    rng = np.random.default_rng(0)
    acc.predict_update(gp, data.x_test, rng.uniform(2, 8, (t, c)))
    print("fused rmse", acc.finalize(0.5)["rmse_aoi_toi"])
    acc.close()
    gp.close()
if which in ("all", "metrics"):
    rng = np.random.default_rng(1)
    x, y, cf = rng.random((70, 333)), rng.random((70, 333)), rng.random((70, 333))
    print("plain rmse", summarise(x, y, cf)["rmse_aoi_toi"], fi_aoi_toi(x, y, 2, 0.1))
if which in ("all", "pre"):
    rng = np.random.default_rng(2)
    cells, ns = 500, 300
    s_ = np.linspace(0, 1, cells)
    elev = 5 + 3 * np.sin(7 * s_)
    modes = np.stack([np.cos((j + 1) * np.pi * s_) for j in range(4)])
    wse = np.maximum(6.5 + (rng.standard_normal((ns, 4)) * [2, 1, 0.5, 0.25]) @ modes + 0.02 * rng.standard_normal((ns, cells)), elev)
    pp = PreProcessor(hydraulic_parameter="depth")
    pp.fit(wse.copy(), elev, rng.uniform(0.5, 2, cells), 4)
    z = pp.transform(wse[:40].copy())
    back, bv = pp.reverse_transform(z, np.abs(z))
    print("pre", pp.fit_info["iterations"], float(np.abs(z).max()))
    pp.close()
if which in ("all", "sgpr"):
    sp = SparseGP("Matern52", n, d, 40, 1)
    sp.set_data(data.x, data.y[:, :1])
    e, gt, gz = sp.elbo_grad(gp.theta_vector(v, s, ls) if which == "all" else np.concatenate([[v, s], ls]), data.x[:40].copy())
    sp.condition(np.concatenate([[v, s], ls]), data.x[:40].copy())
    print("elbo", e, sp.predict(data.x_test)[0][:2].ravel())
    sp.close()
if which in ("all", "batch"):
    from gpras_b200.engine import SparseBatch

    for mm_, dd_ in ((40, d), (70, d)):  # fused path (M <= 64) and the general kernels batched over the models
        sb = SparseBatch("Matern52", n, d, mm_, 3)
        sb.set_data(data.x, data.y[:, :3])
        th3 = np.tile(np.concatenate([[v, s], ls]), (3, 1))
        z3 = np.stack([data.x[k : k + mm_] for k in (0, 50, 100)])
        e3, _, _, info3 = sb.elbo_grad(th3, z3)
        u0 = np.concatenate([np.full((3, 2 + d), 0.5), z3.reshape(3, -1)], axis=1)
        u1, losses, iters = sb.adam(u0, d, True, True, 3)
        print("batch elbo", e3, info3, "adam", losses[-1], iters)
        if sb.fused:
            sb.condition(th3, z3)
            print("batch predict", sb.predict(data.x_test)[0][0])
        sb.close()
