"""Thread-level stress of the PreProcessor paths: concurrent fits / transforms must be bitwise identical."""
import sys, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests" / "golden"))
import numpy as np
from make_golden_reference import flood_samples
from gpras_b200.preprocess import PreProcessor

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
wse, elev, w = flood_samples(2600, 3000, 6, seed=5)
out = [None] * T

def work(i):
    res = []
    for rep in range(3):
        pp = PreProcessor(hydraulic_parameter="wse")
        pp.fit(wse.copy(), elev, w, 6)
        z = pp.transform(wse[:300].copy())
        back, bv = pp.reverse_transform(z, np.abs(z))
        res.append(np.concatenate([pp.eofs.ravel(), pp.eigenvalues, pp.x_std, z.ravel(), back.ravel()[::7], bv.ravel()[::7]]))
        pp.close()
    out[i] = res

ths = [threading.Thread(target=work, args=(i,)) for i in range(T)]
[x.start() for x in ths]; [x.join() for x in ths]
ref = out[0][0]
bad = sum(not np.array_equal(a, ref) for r in out for a in r)
print(f"threads={T} bad={bad}")
