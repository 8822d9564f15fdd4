import sys, time, cProfile, pstats
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gpras_b200 import GPRAS
from gpras_b200.synth import make_gp_data
d = make_gp_data(256, 8, 8, seed=0)
def run(lock):
    g = GPRAS("RBF")
    g.fit(d.x, d.y, None, "kmeans", "stochastic", shared_kernel=True, n_starts=40, iter_initial=20, iter_final=50, seed=0, lockstep=lock)
    return g
run(False); 
for lock in (False, True, True):
    t0 = time.perf_counter(); run(lock); print("lock", lock, time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); run(True); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
