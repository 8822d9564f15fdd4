"""Developer timing: wall-clock per LML+grad evaluation at several N (not a test)."""
import sys, time
sys.path.insert(0, ".")
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta
for (n, d, p) in [(256, 8, 8), (1024, 8, 8), (2048, 16, 16), (4096, 16, 16)]:
    data = make_gp_data(n, d, p, 0, seed=0)
    gp = ExactGP("Matern52", n, d, p); gp.set_data(data.x, data.y)
    v, s, ls = fixed_theta(d, True); th = gp.theta_vector(v, s, ls)
    for _ in range(4): gp.lml_grad(th)
    t0 = time.perf_counter(); reps = 20
    for _ in range(reps): gp.lml_grad(th)
    print(n, d, p, f"{(time.perf_counter()-t0)/reps*1e3:.3f} ms/eval", gp.last_launches(), "launches")
    gp.close()
