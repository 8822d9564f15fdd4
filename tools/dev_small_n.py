"""Throughput of concurrent small-N evaluations (handles in flight on one GPU, one host thread, enqueue / fetch)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta

for n, d, p, kern, ard in [(256, 8, 8, "RBF", False), (2048, 16, 16, "Matern52", True)]:
    data = make_gp_data(n, d, p, 0, seed=0)
    v, s, ls = fixed_theta(d, ard)
    for C in (1, 2, 4, 8, 16, 32):
        gps = []
        for c in range(C):
            g = ExactGP(kern, n, d, p)
            g.set_data(data.x, data.y)
            gps.append(g)
        th = gps[0].theta_vector(v, s, ls)
        evals = 64 * max(1, C // 2)
        def run(count):
            pend = [False] * C
            for i in range(count):
                c = i % C
                if pend[c]:
                    gps[c].fetch()
                gps[c].enqueue(th * (1 + 1e-3 * (i % 7)))
                pend[c] = True
            for c in range(C):
                if pend[c]:
                    gps[c].fetch()
        run(3 * C)
        t0 = time.perf_counter()
        run(evals)
        dt = time.perf_counter() - t0
        print(f"N={n} handles={C:2d}: {evals / dt:9.1f} evals/s  ({1e3 * dt / evals:.3f} ms/eval)", flush=True)
        for g in gps:
            g.close()
