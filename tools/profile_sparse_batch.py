"""Where the time of the reference-default sparse fit goes (N = 5000, D = 10, M = 50, 10 per-column models, two-stage Adam
100 + 100): wall-clock of the fit's host steps and of the device-resident Adam stages, per-iteration device time from CUDA
events.  Run under ncu (--metrics gpu__time_duration.sum) for the per-kernel list."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from gpras_b200 import GPRAS
from gpras_b200.engine import SparseBatch
from gpras_b200.synth import make_gp_data

n, d, m, p = 5000, 10, 50, 10
if len(sys.argv) > 4:
    n, d, m, p = map(int, sys.argv[1:5])
data = make_gp_data(n, d, p, 0, seed=3)
for rep in range(3):
    g = GPRAS("Matern52")
    t0 = time.perf_counter()
    g.fit(data.x, data.y, m, "grid", "two-stage")
    print(f"fit #{rep}: {time.perf_counter() - t0:.4f} s", flush=True)

xt = np.random.default_rng(1).standard_normal((10000, d))
for label, flag in (("batched", True), ("one handle per model", False)):
    g._opts["batched_predict"] = flag
    g.predict(xt[:10])
    t0 = time.perf_counter()
    g.predict(xt)
    t1 = time.perf_counter()
    g.predict(xt)
    print(f"predict 10000 events x {p} models, {label}: first {1e3 * (t1 - t0):.2f} ms (conditions), again {1e3 * (time.perf_counter() - t1):.2f} ms")

batch = SparseBatch("Matern52", n, d, m, p)
t0 = time.perf_counter()
batch.set_data(data.x, data.y)
print(f"set_data: {time.perf_counter() - t0:.4f} s")
z0 = g._create_inducing(data.x, m, "grid")
u0 = np.stack([np.concatenate([[0.5413, 0.5413, 0.5413], z0.ravel()]) for _ in range(p)])
for iters in (1, 10, 100, 100):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    u, losses, it = batch.adam(u0, 1, True, True, iters)
    dt = time.perf_counter() - t0
    print(f"adam stage, {iters} iterations: {dt * 1e3:.3f} ms  ({dt / iters * 1e6:.1f} us / iteration), launches per iteration {batch.last_launches()}")
th = np.tile(np.concatenate([[1.0, 1.0], np.full(d, 2.0)]), (p, 1))
zz = np.tile(z0, (p, 1, 1))
for _ in range(3):
    t0 = time.perf_counter()
    batch.elbo_grad(th, zz)
    print(f"one batched evaluation incl. copies: {(time.perf_counter() - t0) * 1e6:.1f} us")
