"""Thread-level stress: synchronous entry points (predict, predict_cells+metrics, sparse ELBO) on many handles at once."""
import sys, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.engine import ExactGP, SparseGP
from gpras_b200.cells import fold_cell_map
from gpras_b200.metrics import MetricsAccumulator
from gpras_b200.synth import make_gp_data, fixed_theta, make_cell_map

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, d, p, c, t = 1024, 8, 8, 6000, 700
data = make_gp_data(n, d, p, t, seed=0)
v, s, ls = fixed_theta(d, True)
cm = make_cell_map(p, c, seed=0)
e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
rng = np.random.default_rng(1)
truth = rng.uniform(3, 7, (t, c))
z0 = data.x[:40].copy()
out = [None] * T

def work(i):
    res = []
    gp = ExactGP("Matern52", n, d, p)
    gp.set_data(data.x, data.y)
    th = gp.theta_vector(v, s, ls)
    sp = SparseGP("Matern52", n, d, 40, 1)
    sp.set_data(data.x, data.y[:, :1])
    acc = MetricsAccumulator(c, t)
    acc.set_elevations(cm.elevations, cm.elevations)
    for rep in range(6):
        gp.condition(th)
        gp.set_cell_map(e_mean, bias)
        m, vv = gp.predict(data.x_test)
        acc.reset(0.0)
        acc.predict_update(gp, data.x_test, truth)
        sm = acc.finalize(0.5)
        e, gt, gz = sp.elbo_grad(th, z0)
        res.append(np.concatenate([m.ravel(), vv.ravel(), sm["rmse_cell_toi"], sm["err_aoi_ts"], [sm["rmse_aoi_toi"], sm["mae_aoi_toi"], e], gt, gz.ravel()]))
    out[i] = res
    gp.close(); sp.close(); acc.close()

ths = [threading.Thread(target=work, args=(i,)) for i in range(T)]
[x.start() for x in ths]; [x.join() for x in ths]
ref = out[0][0]
bad = 0
for i in range(T):
    for r, a in enumerate(out[i]):
        if not np.array_equal(a, ref):
            bad += 1
            dd = np.abs(a - ref)
            print("thread", i, "rep", r, "max diff", dd.max(), "first idx", int(np.argmax(dd > 0)), "of", a.size)
print(f"threads={T} bad={bad}")
