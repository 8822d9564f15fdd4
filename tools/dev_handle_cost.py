import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta
n, d, p = 256, 8, 8
data = make_gp_data(n, d, p, 0, seed=0)
v, s, ls = fixed_theta(d, False)
g0 = ExactGP("RBF", n, d, p); g0.set_data(data.x, data.y); th = g0.theta_vector(v, s, ls)
for _ in range(3): g0.lml_grad(th)
t0 = time.perf_counter(); gs = [ExactGP("RBF", n, d, p) for _ in range(16)]; t1 = time.perf_counter()
for g in gs: g.set_data(data.x, data.y)
t2 = time.perf_counter()
for g in gs: g.lml_grad(th)
t3 = time.perf_counter()
for g in gs: g.lml_grad(th)
t4 = time.perf_counter()
for g in gs: g.lml_grad(th)
t5 = time.perf_counter()
print(f"per handle: create {1e3*(t1-t0)/16:.2f} ms, set_data {1e3*(t2-t1)/16:.2f} ms, eval#1 (eager) {1e3*(t3-t2)/16:.2f} ms, eval#2 (capture) {1e3*(t4-t3)/16:.2f} ms, eval#3 (replay) {1e3*(t5-t4)/16:.2f} ms")
