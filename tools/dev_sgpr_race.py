"""Is the sparse objective bitwise independent of (a) which handle evaluates it, (b) what else is in flight, (c) what the
handle evaluated before?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.engine import SparseGP
from gpras_b200.synth import make_gp_data
n, d, m, P = 700, 5, 24, 6
data = make_gp_data(n, d, P, 0, seed=12)
rng = np.random.default_rng(0)
thetas = [np.concatenate([[rng.uniform(0.5, 2), rng.uniform(0.3, 1.5)], np.full(d, rng.uniform(0.5, 2))]) for _ in range(12)]
zs = [data.x[rng.permutation(n)[:m]].copy() for _ in range(12)]
def pack(o): return np.concatenate([[o[0]], o[1], o[2].ravel()])
# reference: one fresh handle per (column), sequential
ref = {}
for c in range(P):
    sp = SparseGP("Matern52", n, d, m, 1); sp.set_data(data.x, data.y[:, c:c+1])
    for k in range(12): ref[(c, k)] = pack(sp.elbo_grad(thetas[k], zs[k]))
    sp.close()
# (a)+(b): P handles in flight
sps = [SparseGP("Matern52", n, d, m, 1) for _ in range(P)]
for c, sp in enumerate(sps): sp.set_data(data.x, data.y[:, c:c+1])
bad_ab = 0
for rep in range(5):
    for k in range(12):
        for sp in sps: sp.enqueue(thetas[k], zs[k])
        for c, sp in enumerate(sps):
            bad_ab += not np.array_equal(pack(sp.fetch()), ref[(c, k)])
# (c): one handle switching data between columns
sp = SparseGP("Matern52", n, d, m, 1)
bad_c = 0
for k in range(12):
    for c in range(P):
        sp.set_data(data.x, data.y[:, c:c+1])
        o = pack(sp.elbo_grad(thetas[k], zs[k]))
        if not np.array_equal(o, ref[(c, k)]):
            bad_c += 1
            if bad_c <= 3: print("switching handle differs: col", c, "k", k, "max diff", np.abs(o - ref[(c, k)]).max(), "elbo diff", o[0] - ref[(c, k)][0])
print("in-flight mismatches", bad_ab, "; data-switching mismatches", bad_c)
