"""Wall-clock of the reference's default call GPRAS.fit(x, y, 50, "kmeans", "two-stage") at reference scale, with the k-means
initialiser's Lloyd iterations on the device (default) and with scikit-learn's KMeans."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gpras_b200 import GPRAS
from gpras_b200.synth import make_gp_data

d = make_gp_data(5000, 10, 10, 0, seed=3)
for flag in (True, False, True, False):
    g = GPRAS("Matern52")
    t0 = time.perf_counter()
    g.fit(d.x, d.y, 50, "kmeans", "two-stage", kmeans_on_device=flag)
    print("kmeans_on_device", flag, "fit", round(time.perf_counter() - t0, 4), "s", getattr(g, "kmeans_info", ""), flush=True)
