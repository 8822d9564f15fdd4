"""CPU, world_size 2 over gloo: restart sharding and prediction-shard gathering (gpras_b200/parallel.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpras_b200 import gpr, parallel
    from gpras_b200.synth import make_gp_data
    from test_host_cpu import OracleBackedModel

    # all_gather_rows with ragged blocks
    local = np.full((rank + 1, 3), float(rank))
    got = parallel.all_gather_rows(local, 3)
    assert got.shape == (3, 3) and np.all(got[0] == 0) and np.all(got[1:] == 1)
    assert parallel.shard_rows(10, 0, 2) == (0, 5) and parallel.shard_rows(11, 1, 2) == (6, 11)

    d = make_gp_data(40, 2, 1, seed=6)
    starts = np.array([[1.0, 0.1, 1.0], [0.3, 0.5, 2.5], [2.0, 0.05, 0.7]])
    m = OracleBackedModel("RBF", d.x, d.y, 1.0)
    table = parallel.run_restarts(m, gpr.OPTIMIZERS["L-BFGS-B"], starts, dict(max_iter=30))
    np.save(os.path.join(out_dir, f"table{rank}.npy"), table)
    np.save(os.path.join(out_dir, f"theta{rank}.npy"), m.theta())
    dist.destroy_process_group()


def test_restart_sharding_world2_matches_serial(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "table0.npy"), np.load(tmp_path / "table1.npy")
    np.testing.assert_array_equal(t0, t1)
    np.testing.assert_array_equal(np.load(tmp_path / "theta0.npy"), np.load(tmp_path / "theta1.npy"))
    assert list(t0[:, 0]) == [0.0, 1.0, 2.0]
    # serial run of the same restarts gives the same table (no cross-restart state)
    from gpras_b200 import gpr, parallel
    from gpras_b200.synth import make_gp_data
    from test_host_cpu import OracleBackedModel

    d = make_gp_data(40, 2, 1, seed=6)
    starts = np.array([[1.0, 0.1, 1.0], [0.3, 0.5, 2.5], [2.0, 0.05, 0.7]])
    m = OracleBackedModel("RBF", d.x, d.y, 1.0)
    ts = parallel.run_restarts(m, gpr.OPTIMIZERS["L-BFGS-B"], starts, dict(max_iter=30))
    np.testing.assert_allclose(ts, t0, rtol=1e-12)
    best = int(np.argmin(ts[:, 1]))
    np.testing.assert_allclose(m.theta()[:3], ts[best, 2:5], rtol=1e-12)
