"""Developer timing of the sparse model at reference scale (not a test)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpras_b200.engine import SparseGP
from gpras_b200.synth import make_gp_data
from oracle import sgpr

for (n, d, m) in [(5000, 10, 50), (5000, 10, 300), (10000, 20, 128)]:
    data = make_gp_data(n, d, 1, 1000, seed=1)
    rng = np.random.default_rng(0)
    z = data.x[rng.choice(n, m, replace=False)].copy()
    gp = SparseGP("Matern52", n, d, m, 1)
    gp.set_data(data.x, data.y)
    th = gp.theta_vector(1.0, 0.5, 2.0)
    for _ in range(3):
        gp.elbo_grad(th, z)
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        e, gt, gz = gp.elbo_grad(th, z)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        gp.elbo_grad(th, z, want_grad=False)
    dt0 = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    o = sgpr.training_loss_and_grads("Matern52", data.x, data.y, z, 0.5413, np.array([1.8546]), 0.0)
    dtc = time.perf_counter() - t0
    t0 = time.perf_counter()
    o = sgpr.training_loss_and_grads("Matern52", data.x, data.y, z, 0.5413, np.array([1.8546]), 0.0)
    dtc = min(dtc, time.perf_counter() - t0)
    print(f"N={n} D={d} M={m}: gpu loss+grad {dt*1e3:.3f} ms, loss only {dt0*1e3:.3f} ms, launches {gp.last_launches()}; torch-CPU oracle loss+grad {dtc*1e3:.1f} ms -> x{dtc/dt:.0f}")
    gp.close()
