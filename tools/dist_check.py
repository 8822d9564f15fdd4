"""Multi-GPU host paths over NCCL with real device handles (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py

Checks, against a single-rank evaluation on rank 0's GPU: restart sharding (`fit(restarts=...)`), per-column model sharding,
target-column sharding of the shared-theta objective, event sharding of prediction and of the streaming metrics.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from gpras_b200 import GPRAS, parallel
from gpras_b200.engine import ExactGP
from gpras_b200.metrics import summarise
from gpras_b200.synth import fixed_theta, make_gp_data, random_starts

ok = True
data = make_gp_data(600, 6, 6, 333, seed=1)

# 1. restarts sharded over ranks == the best of all restarts
starts = random_starts(6, 6, seed=3)
starts[:, 1] = np.clip(starts[:, 1], 0.05, 1.0)
g = GPRAS("Matern52")
g.fit(data.x, data.y, None, "kmeans", "L-BFGS-B", ard=True, shared_kernel=True, device=local, restarts=starts, max_iter=20)
table = g.models[0].restart_table
th = g.models[0].theta()
allth = [None] * world
dist.all_gather_object(allth, th)
ok &= all(np.array_equal(allth[0], t) for t in allth) and table.shape[0] == 6
best = int(np.argmin(table[:, 1]))

# 2. per-column models sharded over ranks, parameters gathered
g2 = GPRAS("RBF")
g2.fit(data.x, data.y, 16, "grid", "adam", device=local, max_iter=8)
p2 = np.concatenate([np.concatenate([m.theta(), np.asarray(m.inducing_variable.Z).ravel()]) for m in g2.models])
allp = [None] * world
dist.all_gather_object(allp, p2)
ok &= all(np.allclose(allp[0], t, rtol=1e-13, atol=0) for t in allp)

# 3. target columns sharded: sum over ranks == all columns on one GPU
v, s, ls = fixed_theta(6, True)
lo, hi = parallel.shard_columns(6, rank, world)
gp_loc = ExactGP("Matern52", 600, 6, hi - lo, device=local)
gp_loc.set_data(data.x, np.ascontiguousarray(data.y[:, lo:hi]))
theta = gp_loc.theta_vector(v, s, ls)
lml, grad = parallel.lml_grad_column_sharded(lambda t, wg: gp_loc.lml_grad(t, wg), theta)
gp_all = ExactGP("Matern52", 600, 6, 6, device=local)
gp_all.set_data(data.x, data.y)
lml1, grad1 = gp_all.lml_grad(theta)
ok &= abs(lml - lml1) <= 1e-10 * abs(lml1) and np.allclose(grad, grad1, rtol=1e-9, atol=1e-9)

# 4. prediction events sharded, mode-space results gathered
mean, var = parallel.predict_sharded(g, data.x_test)
if rank == 0:
    m1, v1 = g.predict(data.x_test)
    ok &= np.array_equal(mean, m1) and np.array_equal(var, v1)

# 5. whole events of the metrics sharded
rng = np.random.default_rng(5)
events = [(rng.random((20, 700)), rng.random((20, 700)), rng.random((20, 700))) for _ in range(5)]
keys = ["rmse_aoi_toi", "mae_aoi_toi", "pod_mts"]
tab = parallel.metrics_sharded(list(range(5)), lambda i: summarise(*events[i], device=local), keys)
ref = np.array([[i] + [summarise(*events[i], device=local)[k] for k in keys] for i in range(5)])
ok &= np.array_equal(tab, ref)

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world}: {'OK' if flag.item() == 1 else 'FAILED'}; best restart {best}, lml(all columns) {lml1:.6f} vs sharded {lml:.6f}")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
