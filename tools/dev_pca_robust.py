"""Odd inputs for the device PCA: must not hang or crash, and must stay consistent with the oracle where that is defined."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.preprocess import PreProcessor
from oracle import preprocess as opre

rng = np.random.default_rng(0)
def run(name, x, elev, w, modes, hp="velocity"):
    t0 = time.perf_counter()
    pp = PreProcessor(hydraulic_parameter=hp)
    try:
        pp.fit(x.copy(), elev, w, modes)
        f = opre.fit(x, elev, w, modes, 0.03, hp)
        k = min(len(pp.eigenvalues), len(f.eigenvalues))
        ev_err = float(np.max(np.abs(pp.eigenvalues[:k] - f.eigenvalues[:k])) / max(f.eigenvalues[0], 1e-300)) if k else 0.0
        # compare the projector E^T E (well defined even for degenerate eigenvalues)
        proj_err = float(np.max(np.abs(pp.eofs.T @ pp.eofs - f.eofs.T @ f.eofs))) if pp.eofs.size else 0.0
        print(f"{name:28s} ok  modes {pp.spatial_mode_count}  iters {pp.fit_info['iterations']}  eigenvalue err {ev_err:.1e}  projector err {proj_err:.1e}  {time.perf_counter()-t0:.2f}s")
    except Exception as e:
        print(f"{name:28s} raised {type(e).__name__}: {str(e)[:90]}  {time.perf_counter()-t0:.2f}s")
    pp.close()

c = 300
elev, w = np.zeros(c), np.ones(c)
run("constant data", np.full((50, c), 3.0), elev, w, 2)
q, _ = np.linalg.qr(rng.standard_normal((c, 4)))
coef = rng.standard_normal((400, 4)) * [3.0, 3.0, 1.0, 0.5]   # two (statistically) close leading eigenvalues
run("close eigenvalues", coef @ q.T + 5, elev, w, 4)
coef2 = rng.standard_normal((400, 4)); coef2 -= coef2.mean(0); u, _, _ = np.linalg.svd(coef2, full_matrices=False)
run("exactly equal eigenvalues", (u * [2.0, 2.0, 1.0, 0.5]) @ q.T + 5, elev, w, 2)   # modes 1-2 span a degenerate plane
run("two samples", rng.standard_normal((2, c)) + 5, elev, w, 1)
run("one cell", rng.standard_normal((40, 1)) + 5, np.zeros(1), np.ones(1), 1)
run("flat noise spectrum", rng.standard_normal((600, 700)), np.zeros(700), np.ones(700), 3)
run("North's rule, pure noise", rng.standard_normal((200, 500)) * 3, np.zeros(500), np.ones(500), None)
