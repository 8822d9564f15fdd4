"""Timings of the cells<->modes transforms and the streaming metrics at BASELINE config sizes (one GPU).

    python tools/bench_pre_metrics.py [--quick]

Prints one JSON line per measurement (CUDA-event / wall-clock time after warm-up; inputs resident in HBM unless stated).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
HBM_GBS = 6543.7


def flood_tensor(torch, n, c, k=8, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    c_pad = (c + 127) // 128 * 128
    s = torch.linspace(0, 1, c, dtype=torch.float64, device="cuda")
    elev = 5 + 3 * torch.sin(7 * s) + 2 * s
    modes = torch.stack([torch.cos((j + 1) * np.pi * s + 0.3 * j) for j in range(k)])
    amp = 2.0 * 0.6 ** torch.arange(k, dtype=torch.float64, device="cuda")
    coef = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g) * amp
    buf = torch.zeros(n, c_pad, dtype=torch.float64, device="cuda")
    x = buf[:, :c]
    rows = 1024
    for r0 in range(0, n, rows):
        st = 6.5 + coef[r0:r0 + rows] @ modes + 0.02 * torch.randn(min(rows, n - r0), c, dtype=torch.float64, device="cuda", generator=g)
        x[r0:r0 + rows] = torch.maximum(st, elev)
    w = torch.rand(c, dtype=torch.float64, device="cuda", generator=g) * 1.5 + 0.5
    return x, elev.cpu().numpy(), w.cpu().numpy()


def bench_preprocess(torch, n, c, p, name):
    from gpras_b200.preprocess import PreProcessor

    x, elev, w = flood_tensor(torch, n, c)
    pp = PreProcessor(hydraulic_parameter="wse")
    pp.fit(x, elev, w, p)  # warm-up (kernel attributes, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pp.fit(x, elev, w, p)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = pp.fit_info
    print(json.dumps({"what": "PreProcessor.fit", "config": name, "samples": n, "cells": c, "modes": p, "wall_s": dt,
                      "stage_ms": info["stage_ms"], "iterations": info["iterations"], "launches": info["launches"],
                      "gram_tflops": n * float(n) * c / (info["stage_ms"]["gram"] * 1e-3) * 1e-12,
                      "colstats_gbs": 8.0 * n * c / (info["stage_ms"]["colstats"] * 1e-3) * 1e-9,
                      "max_residual": float(np.max(info["residuals"][:p]))}), flush=True)
    # transform: device in, device out
    for _ in range(2):
        z = pp.transform(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        z = pp.transform(x)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    gbs = 8.0 * n * c / dt * 1e-9
    print(json.dumps({"what": "PreProcessor.transform", "config": name, "samples": n, "cells": c, "modes": p, "ms": dt * 1e3,
                      "algorithmic_GBps": gbs, "hbm_frac": gbs / HBM_GBS, "tflops": 2.0 * n * c * p / dt * 1e-12,
                      "cell_samples_per_s": n * c / dt}), flush=True)
    pp.close()
    del x
    torch.cuda.empty_cache()


def bench_metrics(torch, n, d, p, c, t, name):
    from gpras_b200.cells import fold_cell_map
    from gpras_b200.engine import ExactGP
    from gpras_b200.metrics import MetricsAccumulator
    from gpras_b200.synth import fixed_theta, make_cell_map, make_gp_data

    data = make_gp_data(n, d, p, t, seed=0)
    cm = make_cell_map(p, c, seed=0)
    e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    gp = ExactGP("Matern52", n, d, p)
    gp.set_data(data.x, data.y)
    v, s, ls = fixed_theta(d, True)
    gp.condition(gp.theta_vector(v, s, ls))
    gp.set_cell_map(e_mean, bias)
    xt = torch.from_numpy(data.x_test).cuda()
    acc = MetricsAccumulator(c, 4 * t)
    acc.set_elevations(cm.elevations, cm.elevations)
    g = torch.Generator(device="cuda").manual_seed(1)
    truth = torch.rand(t, c, dtype=torch.float64, device="cuda", generator=g) * 4 + 3

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def run_cells():
        gp.predict_cells(xt, want_modes=False)

    def run_fused(tr):
        acc.reset(0.0)
        acc.predict_update(gp, xt, tr)

    def run_predict_only():
        gp.predict(data.x_test[:1])  # keeps buffers alive; the device-only predictor is timed through predict_cells / fused

    dt_cells = timed(run_cells)
    dt_fused = timed(lambda: run_fused(truth))
    dt_fused_nt = timed(lambda: run_fused(None))
    print(json.dumps({"what": "predict -> cells (T x C written to ring buffer)", "config": name, "events": t, "cells": c, "ms": dt_cells * 1e3,
                      "cell_depths_per_s": t * c / dt_cells}), flush=True)
    print(json.dumps({"what": "predict -> cells -> metrics fused (truth read, nothing written)", "config": name, "events": t, "cells": c,
                      "ms": dt_fused * 1e3, "cell_depths_per_s": t * c / dt_fused}), flush=True)
    print(json.dumps({"what": "predict -> cells -> per-cell reductions fused (no truth)", "config": name, "events": t, "cells": c,
                      "ms": dt_fused_nt * 1e3, "cell_depths_per_s": t * c / dt_fused_nt}), flush=True)
    # stand-alone metrics over resident arrays (24 B per cell-timestep)
    tt = min(t, 2048)
    y = truth[:tt] + 0.1
    conf = torch.rand(tt, c, dtype=torch.float64, device="cuda", generator=g)

    def run_plain():
        acc.reset(0.0)
        acc.update(truth[:tt], y, conf)

    dt_plain = timed(run_plain)
    gbs = 24.0 * tt * c / dt_plain * 1e-9
    print(json.dumps({"what": "metrics over resident (T x C) truth / prediction / confidence", "config": name, "timesteps": tt, "cells": c,
                      "ms": dt_plain * 1e3, "algorithmic_GBps": gbs, "hbm_frac": gbs / HBM_GBS, "cell_depths_per_s": tt * c / dt_plain}),
          flush=True)
    acc.close()
    gp.close()


def main():
    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--transform-only", action="store_true", help="PreProcessor legs only, plus 16 modes at config-3 size")
    args = ap.parse_args()
    if args.transform_only:
        bench_preprocess(torch, 2048, 50_000, 16, "cfg2")
        bench_preprocess(torch, 8192, 200_000, 16, "cfg3 size, 16 modes")
        bench_preprocess(torch, 8192, 200_000, 32, "cfg3")
        return
    bench_preprocess(torch, 2048, 50_000, 16, "cfg2")
    if not args.quick:
        bench_preprocess(torch, 8192, 200_000, 32, "cfg3")
    bench_metrics(torch, 2048, 16, 16, 50_000, 4096, "cfg2")
    if not args.quick:
        bench_metrics(torch, 8192, 32, 32, 200_000, 4096, "cfg3/cfg5")


if __name__ == "__main__":
    main()
