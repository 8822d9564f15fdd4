"""One warm-up + one LML+grad evaluation at BASELINE config 3 (for `ncu --set full -k regex:...`)."""
import sys
sys.path.insert(0, ".")
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta

n, d, p = 8192, 32, 32
data = make_gp_data(n, d, p, 0, seed=0)
v, s, ls = fixed_theta(d, True)
gp = ExactGP("Matern52", n, d, p)
gp.set_data(data.x, data.y)
th = gp.theta_vector(v, s, ls)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print(gp.lml_grad(th)[0])
gp.close()
