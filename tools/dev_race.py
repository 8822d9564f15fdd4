"""Stress: many handles in flight, same theta; every result must be bitwise identical."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta

n, C, rounds = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = p = 16
data = make_gp_data(n, d, p, 0, seed=0)
v, s, ls = fixed_theta(d, True)
gps = []
for c in range(C):
    g = ExactGP("Matern52", n, d, p)
    g.set_data(data.x, data.y)
    gps.append(g)
th = gps[0].theta_vector(v, s, ls)
ref = None
bad = 0
for r in range(rounds):
    for g in gps:
        g.enqueue(th)
    for c, g in enumerate(gps):
        try:
            lml, grad = g.fetch()
        except Exception as e:
            print("round", r, "handle", c, "ERR", str(e)[:80]); bad += 1; continue
        cur = np.concatenate([[lml], grad])
        if ref is None:
            ref = cur
        elif not np.array_equal(cur, ref):
            bad += 1
            print("round", r, "handle", c, "MISMATCH", float(np.max(np.abs(cur - ref))), "lml diff", cur[0] - ref[0])
            for which, name in [(1, "L"), (2, "W"), (3, "Wt(Kinv)"), (4, "alpha")]:
                a, b = g.get_matrix(which), gps[0].get_matrix(which)
                if which < 4:
                    a, b = np.tril(a), np.tril(b)
                dd = np.abs(a - b)
                if dd.max() > 0:
                    idx = np.argwhere(dd > 0)
                    print("    ", name, "differs: max", dd.max(), "count", len(idx), "rows", idx[:, 0].min(), idx[:, 0].max(), "cols", idx[:, 1].min(), idx[:, 1].max())
                else:
                    print("    ", name, "identical")
print(f"N={n} handles={C} rounds={rounds} graphs={'off' if os.environ.get('GPRAS_B200_NO_GRAPHS') else 'on'} bad={bad}")
