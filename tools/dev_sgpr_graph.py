import sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.engine import SparseGP
from gpras_b200.synth import make_gp_data
n, d, m = 5000, 10, 50
data = make_gp_data(n, d, 2, 0, seed=3)
th = np.concatenate([[1.0, 1.0], np.full(d, 0.8)])
z = data.x[:m].copy()
sp = SparseGP("Matern52", n, d, m, 1); sp.set_data(data.x, data.y[:, :1])
res = []
for i in range(6):
    t0 = time.perf_counter(); e, gt, gz = sp.elbo_grad(th, z); dt = time.perf_counter() - t0
    res.append(np.concatenate([[e], gt, gz.ravel()]))
    print(i, f"{dt*1e3:.3f} ms", e, sp.last_launches())
print("all equal:", all(np.array_equal(r, res[0]) for r in res))
t0 = time.perf_counter()
for i in range(200): sp.elbo_grad(th * (1 + 1e-4 * i), z)
print("per eval", (time.perf_counter() - t0) / 200 * 1e3, "ms")
sps = [SparseGP("Matern52", n, d, m, 1) for _ in range(10)]
for s_ in sps: s_.set_data(data.x, data.y[:, :1])
for rep in range(3):
    for s_ in sps: s_.enqueue(th, z)
    out = [s_.fetch() for s_ in sps]
t0 = time.perf_counter()
for rep in range(100):
    for s_ in sps: s_.enqueue(th, z)
    out = [s_.fetch() for s_ in sps]
print("10 in flight: per eval", (time.perf_counter() - t0) / 1000 * 1e3, "ms; equal to single:", np.array_equal(np.concatenate([[out[3][0]], out[3][1], out[3][2].ravel()]), res[0]))
