"""predict_cells keeping the cell-space output (no ring buffer), 2048 events (for ncu)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta, make_cell_map
from gpras_b200.cells import fold_cell_map
n, d, p, c = 8192, 32, 32, 200_000
data = make_gp_data(n, d, p, 2048, seed=0)
v, s, ls = fixed_theta(d, True)
gp = ExactGP("Matern52", n, d, p)
gp.set_data(data.x, data.y)
th = gp.theta_vector(v, s, ls)
gp.condition(th)
cm = make_cell_map(p, c, seed=0)
e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
gp.set_cell_map(e_mean, bias)
xt = torch.from_numpy(data.x_test).cuda()
pitch = gp.cell_pitch()
om = torch.empty((2048, pitch), dtype=torch.float64, device="cuda")
ov = torch.empty((2048, pitch), dtype=torch.float64, device="cuda")
for _ in range(2):
    gp.predict_cells(xt, om, ov, want_modes=False)
torch.cuda.synchronize()
print("ok")
gp.close()
