"""Developer timing: the reference's default recipe (sparse models, two-stage Adam) at reference scale."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpras_b200 import GPRAS
from gpras_b200.synth import make_gp_data
n, d, p, m = 5000, 10, 10, 50
data = make_gp_data(n, d, p, 1000, seed=3)
for label, kw in [("lock-step (default), first call", {}), ("lock-step (default)", {}), ("sequential", dict(lockstep_models=False)),
                  ("n_jobs=10 threads", dict(n_jobs=10))]:
    g = GPRAS("Matern52")
    t0 = time.perf_counter()
    g.fit(data.x, data.y, m, "kmeans", "two-stage", max_iter=100, **kw)
    dt = time.perf_counter() - t0
    evals = sum(mm.n_evals for mm in g.models)
    t0 = time.perf_counter(); mean, var = g.predict(data.x_test); dp = time.perf_counter() - t0
    print(f"{label}: fit {dt:.3f} s, {evals} loss+grad evals -> {dt/evals*1e3:.3f} ms/eval aggregate; predict 1000 events x {p} models {dp*1e3:.1f} ms", flush=True)
