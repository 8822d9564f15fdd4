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

    # sparse model whose recipe trains Z: the winner's inducing inputs travel with its hyperparameters, on every rank
    from test_host_cpu import OracleBackedSparseModel

    ds = make_gp_data(50, 2, 1, seed=4)
    sm = OracleBackedSparseModel("RBF", ds.x, ds.y, ds.x[:5].copy(), 1.0)
    stab = parallel.run_restarts(sm, gpr.OPTIMIZERS["two-stage"], np.array([[1.0, 0.1, 1.0], [0.4, 0.3, 2.0], [2.0, 0.05, 0.6]]), dict(max_iter=3))
    np.save(os.path.join(out_dir, f"sparse_table{rank}.npy"), stab)
    np.save(os.path.join(out_dir, f"sparse_state{rank}.npy"), np.concatenate([sm.theta(), np.asarray(sm.inducing_variable.Z).ravel(),
                                                                              [sm.training_loss()]]))

    # target columns sharded over ranks: LML and gradient are sums over column blocks (oracle as the local evaluator)
    from oracle import metrics as om
    from oracle.exact_gp import Theta, lml_and_grad

    d5 = make_gp_data(50, 3, 5, seed=8)
    lo, hi = parallel.shard_columns(5, rank, world)
    th = np.array([1.3, 0.07, 1.5, 2.0, 0.8])

    def local_eval(theta, want_grad):
        lml, gv, gn, gl = lml_and_grad("Matern52", d5.x, d5.y[:, lo:hi], Theta(theta[0], theta[1], theta[2:]), want_grad=want_grad)
        return lml, np.concatenate([[gv, gn], gl])

    lml, grad = parallel.lml_grad_column_sharded(local_eval, th)
    np.save(os.path.join(out_dir, f"colshard{rank}.npy"), np.concatenate([[lml], grad]))

    # per-column models sharded over ranks, parameters gathered
    d3 = make_gp_data(30, 2, 3, seed=11)
    models = [OracleBackedModel("RBF", d3.x, d3.y[:, j : j + 1], 1.0) for j in range(3)]
    parallel.run_models_sharded(models, lambda m_: gpr.OPTIMIZERS["L-BFGS-B"](m_, max_iter=15))
    np.save(os.path.join(out_dir, f"models{rank}.npy"), np.array([m_.theta() for m_ in models]))
    # the same through the whole-shard hook (what the device-resident trainer plugs into): once accepting, once declining
    for accept in (True, False):
        models2 = [OracleBackedModel("RBF", d3.x, d3.y[:, j : j + 1], 1.0) for j in range(3)]
        seen = []

        def run_many(shard, _accept=accept, _seen=seen):
            _seen.append(len(shard))
            if _accept:
                for m_ in shard:
                    gpr.OPTIMIZERS["L-BFGS-B"](m_, max_iter=15)
            return _accept

        parallel.run_models_sharded(models2, lambda m_: gpr.OPTIMIZERS["L-BFGS-B"](m_, max_iter=15), run_many)
        assert seen == [len(parallel.shard_indices(3, rank, world))]
        assert np.array_equal(np.array([m_.theta() for m_ in models2]), np.array([m_.theta() for m_ in models]))

    # events sharded over ranks for the metrics
    rng = np.random.default_rng(3)
    events = [(rng.random((6, 9)), rng.random((6, 9)), rng.random((6, 9))) for _ in range(5)]
    keys = ["rmse_aoi_toi", "mae_aoi_toi", "conf_aoi_toi", "err_aoi_mts"]
    tab = parallel.metrics_sharded(list(range(5)), lambda i: om.summarise(*events[i]), keys)
    np.save(os.path.join(out_dir, f"metrics{rank}.npy"), tab)
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


def test_sparse_restarts_world2_keep_z_with_theta(tmp_path):
    """ADVICE (round 1): with Z-training recipes every rank must end with the winner's theta AND its inducing inputs."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "sparse_table0.npy"), np.load(tmp_path / "sparse_table1.npy")
    np.testing.assert_array_equal(t0, t1)
    s0, s1 = np.load(tmp_path / "sparse_state0.npy"), np.load(tmp_path / "sparse_state1.npy")
    np.testing.assert_array_equal(s0, s1)  # same theta, same Z, same loss on both ranks
    best = int(np.argmin(t0[:, 1]))
    np.testing.assert_allclose(s0[:3], t0[best, 2:5], rtol=1e-14)
    np.testing.assert_allclose(s0[-1], t0[best, 1], rtol=1e-10)  # the model every rank holds is the one whose loss is reported


def test_column_and_event_sharding_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from gpras_b200.synth import make_gp_data
    from oracle import metrics as om
    from oracle.exact_gp import Theta, lml_and_grad

    c0, c1 = np.load(tmp_path / "colshard0.npy"), np.load(tmp_path / "colshard1.npy")
    np.testing.assert_array_equal(c0, c1)
    d5 = make_gp_data(50, 3, 5, seed=8)
    lml, gv, gn, gl = lml_and_grad("Matern52", d5.x, d5.y, Theta(1.3, 0.07, np.array([1.5, 2.0, 0.8])))
    np.testing.assert_allclose(c0, np.concatenate([[lml, gv, gn], gl]), rtol=1e-11)
    a0, a1 = np.load(tmp_path / "models0.npy"), np.load(tmp_path / "models1.npy")
    np.testing.assert_allclose(a0, a1, rtol=1e-14)  # gathered values pass through the inverse softplus once more
    from gpras_b200 import gpr
    from test_host_cpu import OracleBackedModel

    d3 = make_gp_data(30, 2, 3, seed=11)
    for j in range(3):
        mj = OracleBackedModel("RBF", d3.x, d3.y[:, j : j + 1], 1.0)
        gpr.OPTIMIZERS["L-BFGS-B"](mj, max_iter=15)
        np.testing.assert_allclose(a0[j], mj.theta(), rtol=1e-12)
    m0, m1 = np.load(tmp_path / "metrics0.npy"), np.load(tmp_path / "metrics1.npy")
    np.testing.assert_array_equal(m0, m1)
    rng = np.random.default_rng(3)
    events = [(rng.random((6, 9)), rng.random((6, 9)), rng.random((6, 9))) for _ in range(5)]
    for i, (x, y, cf) in enumerate(events):
        s = om.summarise(x, y, cf)
        np.testing.assert_allclose(m0[i], [i, s["rmse_aoi_toi"], s["mae_aoi_toi"], s["conf_aoi_toi"], s["err_aoi_mts"]], rtol=1e-13)
