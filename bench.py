#!/usr/bin/env python
"""Benchmark of the GPR hot path: LML+grad evaluations/s at N=8192 (BASELINE.json metric, config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one exact-GP log-marginal-likelihood + analytic-gradient evaluation at a fresh hyperparameter
vector (covariance build, Cholesky, triangular inverse, K^-1, alpha, fused trace pass) on synthetic storm-event
data of BASELINE config 3: N=8192 training rows x 32 features, 32 target columns sharing theta, Matern-5/2 ARD.
With N GPUs (torchrun, one rank per GPU) every rank evaluates its own shard of optimiser restarts -- K steps per
rank, no data-path collective, one tiny all-gather of the [LML, grad] rows at the end ("weak" scaling).

Printed JSON line (rank 0): `value` = evaluations/s with inputs resident in HBM, timed with CUDA events on the
launching stream; `e2e` = the same through the host-buffer C-ABI call (X, Y, theta uploaded from pinned host memory
and LML+grad read back every step); `roofline` = the DMMA tile-GEMM engine's algorithmic FP64 FLOP/s against the
measured cuBLAS DGEMM rate; `cpu_baseline` = the oracle port timed on this box's host cores.
`--impl reference` times that CPU port (the reference's GPflow/TensorFlow stack cannot be installed offline).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(n=8192, d=32, p=32, kernel="Matern52", ard=True)
METRIC = "LML+grad evals/s at N=8192"
UNIT = "evals/s"
# measured on this pool's B200 (profiles/r01/lib_bars_cublas_cusolver.json): cuBLAS DGEMM 8192^3, sustained
FP64_DGEMM_TFLOPS = 35.4
# one `ncu --set full` capture per kernel (cold cache, one launch each), summarised in profiles/
NCU = {"source": "profiles/r01/ncu_full_final_r01b.json, ncu_eval_traffic_final_summary.json (one launch each, cold cache)",
       "lauum_dram_bytes": 1.634e9, "lauum_dmma_pct": 96.7, "syrk_dram_bytes": 4.668e8, "syrk_dmma_pct": 75.8,
       "eval_dram_bytes": 11.2e9, "eval_launches": 262}
FP64_PEAK_SOURCE = "measured cuBLAS Dgemm 8192^3 on this pool (profiles/r01/lib_bars_cublas_cusolver.json); MEASURED_PEAKS.json has no FP64 entry"


def f_eval(n: int, p: int) -> float:
    """Algorithmic FP64 FLOPs of one evaluation's dense stages: N^3 + 3 N^2 P."""
    return float(n) ** 3 + 3.0 * float(n) ** 2 * p


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def cpu_port_eval_seconds(data, thetas, n_evals: int, warm: int):
    """Time the oracle's CPU restatement (LAPACK potrf/potri + GEMM traces) on all host cores."""
    from threadpoolctl import threadpool_limits

    from oracle.exact_gp import Theta, lml_and_grad_fast

    cores = os.cpu_count() or 1
    with threadpool_limits(limits=cores):
        for i in range(warm):
            lml_and_grad_fast(WORKLOAD["kernel"], data.x, data.y, Theta(thetas[i, 0], thetas[i, 1], thetas[i, 2:]))
        t0 = time.perf_counter()
        for i in range(n_evals):
            th = thetas[(warm + i) % len(thetas)]
            lml_and_grad_fast(WORKLOAD["kernel"], data.x, data.y, Theta(th[0], th[1], th[2:]))
        dt = time.perf_counter() - t0
    return dt, cores


def make_thetas(count: int, d: int, seed: int = 2) -> np.ndarray:
    """Well-conditioned hyperparameter candidates around the fixed parity point."""
    from gpras_b200.synth import fixed_theta

    rng = np.random.default_rng(seed)
    v, s, ls = fixed_theta(d, True)
    out = np.empty((count, 2 + d))
    out[:, 0] = v * np.exp(rng.uniform(-0.3, 0.3, count))
    out[:, 1] = s * np.exp(rng.uniform(-0.3, 0.3, count))
    out[:, 2:] = ls[None, :] * np.exp(rng.uniform(-0.2, 0.2, (count, d)))
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gpras_b200.synth import make_gp_data

    w = WORKLOAD
    data = make_gp_data(w["n"], w["d"], w["p"], 0, seed=0)
    thetas = make_thetas(8, w["d"])
    n_evals = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    dt, cores = cpu_port_eval_seconds(data, thetas, n_evals, warm)
    value = n_evals / dt
    sample = f"{n_evals} full LML+grad evals at N={w['n']} (of {args.steps} requested) after {warm} warm-up, NumPy/SciPy-OpenBLAS oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / n_evals, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg3: exact GP LML+grad, N={w['n']}, D={w['d']}, P={w['p']}, {w['kernel']} ARD + noise, shared theta"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from gpras_b200.engine import ExactGP
    from gpras_b200.synth import make_cell_map, make_gp_data
    from gpras_b200.cells import fold_cell_map

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOAD
    n, d, p = w["n"], w["d"], w["p"]
    K, W = args.steps, max(args.warmup, 3)
    data = make_gp_data(n, d, p, 4096, seed=0)
    thetas = make_thetas(W + K, d, seed=2 + rank)  # every rank = its own shard of restarts / candidates
    stream = torch.cuda.current_stream()
    # C independent evaluations in flight per GPU (independent restarts / candidates): each handle owns its workspace
    # and stream, so one evaluation's latency-bound Cholesky tail overlaps another's dense products.
    C = max(1, args.concurrent)
    streams = [torch.cuda.Stream() for _ in range(C)]
    gps = []
    for c in range(C):
        g_ = ExactGP(w["kernel"], n, d, p, device=local)
        g_.set_stream(streams[c].cuda_stream)
        g_.set_data(data.x, data.y)
        gps.append(g_)
    gp = gps[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_evals(idx0, count, out=None, host_xy=None):
        """Round-robin `count` evaluations over the C handles; returns when all results are on the host.
        host_xy = (x, y) pinned host arrays: upload them before every evaluation (the end-to-end path)."""
        pending = [None] * C
        for i in range(count):
            c = i % C
            if pending[c] is not None:
                lml, g = gps[c].fetch()
                if out is not None:
                    out[pending[c], 0], out[pending[c], 1:] = lml, g
            if host_xy is not None:
                gps[c].set_data(*host_xy)
            gps[c].enqueue(thetas[idx0 + i])
            pending[c] = i
        for c in range(C):
            if pending[c] is not None:
                lml, g = gps[c].fetch()
                if out is not None:
                    out[pending[c], 0], out[pending[c], 1:] = lml, g

    results = np.zeros((K, 3 + d))
    # ---- resident-input throughput -------------------------------------------------------------
    run_evals(0, W)
    launches_per_eval = gp.last_launches()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for st in streams:
        st.wait_stream(stream)
    run_evals(W, K, results)
    for st in streams:
        stream.wait_stream(st)
    if world > 1:  # the path's only exchange: all-gather of per-restart [LML, grad] rows
        loc = torch.from_numpy(results).cuda()
        allr = torch.empty((world * K, 3 + d), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, loc)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * K / (ms_total * 1e-3)

    # ---- the timed results must not depend on what else was in flight: recompute two of them alone ----
    verified = True
    for i in (0, K - 1):
        lml1, g1 = gp.lml_grad(thetas[W + i])
        verified = verified and lml1 == results[i, 0] and bool(np.array_equal(g1, results[i, 1:]))

    # ---- stage timing for the roofline (separate pass; event records perturb nothing measurable) --
    gp.set_stage_timing(True)
    stage = []
    for i in range(min(K, 5)):
        gp.lml_grad(thetas[W + i])
        stage.append(gp.last_stage_ms())
    gp.set_stage_timing(False)
    dense_ms = float(np.mean([s["potrf"] + s["trtri"] + s["lauum"] + s["alpha"] for s in stage]))
    stage_mean = {k: float(np.mean([s[k] for s in stage])) for k in stage[0]}
    achieved = f_eval(n, p) / (dense_ms * 1e-3) * 1e-12
    third = float(n) ** 3 / 3.0
    per_stage = {k: third / (stage_mean[k] * 1e-3) * 1e-12 for k in ("potrf", "trtri", "lauum")}

    # ---- end to end through the host-buffer C-ABI entry point ------------------------------------
    xp = torch.from_numpy(data.x).pin_memory().numpy()
    yp = torch.from_numpy(data.y).pin_memory().numpy()
    run_evals(0, 2, host_xy=(xp, yp))
    barrier()
    t0 = time.perf_counter()
    run_evals(W, K, host_xy=(xp, yp))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * K / float(dt.item())
    h2d = 8 * (n * d + n * p + 2 + d)
    d2h = 8 * (3 + d) + 4

    # ---- second headline: predicted cell-depths/s (predict + modes->cells on device, ring buffer) --
    predict = None
    if rank == 0 or world > 1:
        c = 200_000
        cm = make_cell_map(p, c, seed=0)
        e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
        gp.condition(thetas[0])
        gp.set_cell_map(e_mean, bias)
        xt = torch.from_numpy(data.x_test).cuda()
        gp.predict_cells(xt, want_modes=False)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        p0.record(stream)
        for _ in range(reps):
            gp.predict_cells(xt, want_modes=False)
        p1.record(stream)
        barrier()
        pms = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        t_events = world * reps * xt.shape[0]
        predict = {
            "metric": "predicted cell-depths/s", "value": t_events * c / (float(pms.item()) * 1e-3), "unit": "cell-depths/s",
            "events_per_s": t_events / (float(pms.item()) * 1e-3), "cells": c, "events_per_rank": int(xt.shape[0]),
            "output": "mean+variance per cell written to a device ring buffer (T*C*16 B cannot be kept)",
        }
        # the same sweep with the consumer fused: depth conversion + every metric of gpras/metrics.py accumulated on the fly
        # against a resident truth block, the T x C prediction never written
        from gpras_b200.metrics import MetricsAccumulator

        acc = MetricsAccumulator(c, reps * int(xt.shape[0]) + 1, device=local)
        acc.set_elevations(cm.elevations, cm.elevations)
        gtr = torch.Generator(device="cuda").manual_seed(7 + rank)
        truth = torch.rand(int(xt.shape[0]), c, dtype=torch.float64, device="cuda", generator=gtr) * 4 + 3
        acc.reset(0.0)
        acc.predict_update(gp, xt, truth)
        barrier()
        p0.record(stream)
        acc.reset(0.0)
        for _ in range(reps):
            acc.predict_update(gp, xt, truth)
        summary = acc.finalize(0.5)
        p1.record(stream)
        barrier()
        fms = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(fms, op=dist.ReduceOp.MAX)
        predict["fused_metrics"] = {
            "value": t_events * c / (float(fms.item()) * 1e-3), "unit": "cell-depths/s",
            "what": "predict -> cells -> depth -> all gpras/metrics.py reductions vs a resident truth block; nothing written per cell-depth",
            "rmse_aoi_toi": summary["rmse_aoi_toi"],
        }
        acc.close()
        del truth

    # ---- the cells -> modes side at BASELINE config 2 sizes: PCA fit and the forward transform ----
    preprocess = None
    if rank == 0:
        from gpras_b200.preprocess import PreProcessor

        sys.path.insert(0, str(ROOT / "tools"))
        from bench_pre_metrics import flood_tensor

        ns, cells, modes = 2048, 50_000, 16
        xs_, elev_, w_ = flood_tensor(torch, ns, cells)
        pp = PreProcessor(hydraulic_parameter="wse", device=local)
        pp.fit(xs_, elev_, w_, modes)
        pp.fit(xs_, elev_, w_, modes)
        fit_ms = pp.fit_info["stage_ms"]["total"]
        pp.transform(xs_)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            pp.transform(xs_)
        torch.cuda.synchronize()
        tr_s = (time.perf_counter() - t0) / 5
        preprocess = {"workload": f"PreProcessor: {ns} samples x {cells} cells, {modes} modes, inputs resident in HBM",
                      "fit_ms": fit_ms, "fit_iterations": pp.fit_info["iterations"], "transform_ms": tr_s * 1e3,
                      "transform_GBps": 8.0 * ns * cells / tr_s * 1e-9, "transform_hbm_frac": 8.0 * ns * cells / tr_s * 1e-9 / 6543.7}
        pp.close()
        del xs_

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dtc, cores = cpu_port_eval_seconds(data, thetas, 1, 0)
        cpu = {"value": 1.0 / dtc, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 full LML+grad eval at N={n} (NumPy/SciPy-OpenBLAS oracle port, all host threads), {dtc:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"cfg3: exact GP LML+grad, N={n}, D={d}, P={p}, {w['kernel']} ARD + noise, shared theta; "
                                   f"{K} restarts' evaluations per GPU, {C} in flight", "l2": "working set 1.5 GiB per eval >> 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_eval * K,
            "verified": {"concurrent_equals_serial_bitwise": verified},
            # Dominant kernel = the tile-GEMM engine (gemm_tile_kernel): ~90 % of the step.  Its largest single launch is
            # K^-1 = W^T W (N^3/3 FLOP in ONE launch, timed live between CUDA events on its stream); `aggregate` is the same
            # ratio over ALL dense stages of the evaluation (Cholesky chain and leaf kernels included), the harsher number.
            "roofline": {"bound": "tensor", "achieved": per_stage["lauum"], "peak": FP64_DGEMM_TFLOPS, "unit": "TFLOP/s",
                         "frac": per_stage["lauum"] / FP64_DGEMM_TFLOPS, "traffic": NCU["lauum_dram_bytes"],
                         "kernel": "gemm_tile_kernel<128x128, k-major A, k-major B> (W^T W launch, N^3/3 FLOP)",
                         "flops_per_launch": third, "ms_per_launch": stage_mean["lauum"],
                         "algorithmic_bytes_per_launch": 8.0 * n * n, "peak_source": FP64_PEAK_SOURCE,
                         "aggregate": {"what": "all dense stages of one evaluation: potrf + inverse + W^T W + alpha (F_eval = N^3 + 3 N^2 P)",
                                       "achieved": achieved, "frac": achieved / FP64_DGEMM_TFLOPS, "flops_per_eval": f_eval(n, p),
                                       "dense_ms_per_eval": dense_ms},
                         "job": {"what": "F_eval x measured evals/s of the timed region (two evaluations in flight per GPU)",
                                 "achieved": value / world * f_eval(n, p) * 1e-12, "frac": value / world * f_eval(n, p) * 1e-12 / FP64_DGEMM_TFLOPS},
                         "stage_tflops": per_stage,
                         "ncu": {"source": NCU["source"], "whole_eval_dram_bytes": NCU["eval_dram_bytes"],
                                 "whole_eval_algorithmic_bytes": 16.0 * n * n + 16.0 * n * d + 8.0 * n * p,
                                 "lauum_launch": {"dram_bytes": NCU["lauum_dram_bytes"], "algorithmic_bytes": 8.0 * n * n,
                                                  "dmma_pipe_active_pct": NCU["lauum_dmma_pct"]},
                                 "syrk_launch": {"dram_bytes": NCU["syrk_dram_bytes"], "algorithmic_bytes": 2 * 1953 * 128 * 128 * 8.0,
                                                 "dmma_pipe_active_pct": NCU["syrk_dmma_pct"]}}},
            "stage_ms": stage_mean,
            "cpu_baseline": cpu,
            "predict": predict,
            "preprocess": preprocess,
        }
        print(json.dumps(line), flush=True)
    for g_ in gps:
        g_.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--concurrent", type=int, default=2, help="independent evaluations in flight per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
