#!/usr/bin/env python
"""Benchmark of the GPR hot path: LML+grad evaluations/s at N=8192 (BASELINE.json metric, config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--concurrent C] [--quick]

One "step" is one exact-GP log-marginal-likelihood + analytic-gradient evaluation at a fresh hyperparameter
vector (covariance build, Cholesky, triangular inverse, K^-1, alpha, fused trace pass) on synthetic storm-event
data of BASELINE config 3: N=8192 training rows x 32 features, 32 target columns sharing theta, Matern-5/2 ARD.
With N GPUs (torchrun, one rank per GPU) every rank evaluates its own shard of optimiser restarts -- K steps per
rank, no data-path collective, one tiny all-gather of the [LML, grad] rows at the end ("weak" scaling).

Printed JSON line (rank 0):
  value      evaluations/s with inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same work through the host-buffer C-ABI call (X, Y, theta uploaded from pinned host memory, LML +
             gradient read back, every evaluation) including the same all-gather
  roofline   job level: F_eval x evals/s per GPU against the cuBLAS DGEMM rate MEASURED IN THIS RUN; per-stage and
             per-launch figures in `roofline.detail`
  sustained  the same loop kept up for >= 3 s
  predict    predicted cell-depths/s (BASELINE's second headline) with its own roofline (FP64 N^2 T and HBM 16 B/cell-depth)
  cfg4, cfg5_sweep, strong   BASELINE configs 4 and 5 and config 3's fixed 64 restarts (strong scaling: fixed total work)
  cpu_baseline   the oracle port on this box's host cores (N = 1 only)
`--impl reference` times that CPU port alone (the reference's GPflow/TensorFlow stack cannot be installed offline).
"""
from __future__ import annotations

import argparse
import glob
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
HBM_PEAK_GBS = 6543.7  # MEASURED_PEAKS.json (driver-written copy bandwidth on this pool's B200s)


def workload_string() -> str:
    w = WORKLOAD
    return f"cfg3: exact GP LML+grad, N={w['n']}, D={w['d']}, P={w['p']}, {w['kernel']} ARD + noise, shared theta"


def config_block() -> dict:
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_string(), "l2": "working set 1.5 GiB per evaluation >> 126 MB L2 (no flush needed)"}


def f_eval(n: int, p: int) -> float:
    """Algorithmic FP64 FLOPs of one evaluation's dense stages: N^3 + 3 N^2 P."""
    return float(n) ** 3 + 3.0 * float(n) ** 2 * p


def hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return HBM_PEAK_GBS, "MEASURED_PEAKS.json absent: value recorded by the driver for this pool in round 1"


def ncu_inputs() -> dict:
    """Figures taken from committed ncu captures (never typed into this file): the newest profiles/r*/bench_ncu_inputs.json."""
    files = sorted(glob.glob(str(ROOT / "profiles" / "r*" / "bench_ncu_inputs.json")))
    if not files:
        return {"source": None}
    d = json.loads(Path(files[-1]).read_text())
    d["file"] = str(Path(files[-1]).relative_to(ROOT))
    return d


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
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
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
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
            "power_w_max": max(self.power) if self.power else None,
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
    n_evals = max(1, min(args.steps, 3))  # one evaluation takes ~10 s on 16 cores: a bounded sample of the K requested
    warm = 1 if args.warmup > 0 else 0
    dt, cores = cpu_port_eval_seconds(data, thetas, n_evals, warm)
    value = n_evals / dt
    sample = (f"{n_evals} full LML+grad evaluations at N={w['n']} timed (min(steps, 3) of the {args.steps} requested; ms_per_step is the mean "
              f"of those) after {warm} warm-up, NumPy/SciPy-OpenBLAS oracle port on all host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n_evals,
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * dt / n_evals, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_block(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_fp64_peak(torch, seconds: float = 1.5) -> dict:
    """cuBLAS DGEMM 8192^3 on this GPU, now: best of 5 (burst) and back to back for `seconds` (sustained).  The library call is
    the roofline DENOMINATOR only (MEASURED_PEAKS.json records no FP64 figure); nothing on the product path calls cuBLAS."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(4):
            torch.matmul(a, b, out=c)
        reps += 4
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = reps * 2.0 * n**3 / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    del a, b, c
    torch.cuda.empty_cache()
    return {"burst_tflops": 2.0 * n**3 / (best * 1e-3) * 1e-12, "sustained_tflops": sustained,
            "how": f"torch.matmul float64 {n}^3 (cuBLAS Dgemm) in this process: best of 5, and back to back for {seconds} s"}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from gpras_b200.cells import fold_cell_map
    from gpras_b200.engine import ExactGP
    from gpras_b200.synth import make_cell_map, make_gp_data

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
    hbm_gbs, hbm_src = hbm_peak()
    legs = None if args.legs == "all" else set(args.legs.split(","))

    def leg(name: str) -> bool:
        return not args.quick and (legs is None or name in legs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- FP64 roofline denominator, measured here and now ----------------------------------------
    fp64 = measure_fp64_peak(torch)
    barrier()

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

    def run_evals(th, count, out=None, host_xy=None):
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
            gps[c].enqueue(th[i % len(th)])
            pending[c] = i
        for c in range(C):
            if pending[c] is not None:
                lml, g = gps[c].fetch()
                if out is not None:
                    out[pending[c], 0], out[pending[c], 1:] = lml, g

    def gather_results(res):
        """The path's only exchange: all-gather of the per-restart [LML, grad] rows."""
        if world > 1:
            loc = torch.from_numpy(res).cuda()
            allr = torch.empty((world * res.shape[0], res.shape[1]), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allr, loc)
            return allr
        return None

    def timed_region(th, count, host_xy=None):
        """`count` evaluations per rank + the all-gather, bracketed by barrier + synchronize; device time, max over ranks."""
        res = np.zeros((count, 3 + d))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for st in streams:
            st.wait_stream(stream)
        run_evals(th, count, res, host_xy)
        for st in streams:
            stream.wait_stream(st)
        gather_results(res)
        e1.record(stream)
        barrier()
        return reduce_max(e0.elapsed_time(e1)), res

    # ---- resident-input throughput (the headline `value`) ------------------------------------------
    run_evals(thetas[:W], W)
    launches_per_eval = gp.last_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, results = timed_region(thetas[W:], K)
    clocks = sampler.stop()
    value = world * K / (ms_total * 1e-3)

    # ---- the timed results must not depend on what else was in flight: recompute two of them alone ----
    verified = True
    for i in (0, K - 1):
        lml1, g1 = gp.lml_grad(thetas[W + i])
        verified = verified and lml1 == results[i, 0] and bool(np.array_equal(g1, results[i, 1:]))

    # ---- end to end through the host-buffer C-ABI entry point: same work, same all-gather --------------
    xp = torch.from_numpy(data.x).pin_memory().numpy()
    yp = torch.from_numpy(data.y).pin_memory().numpy()
    run_evals(thetas[:2], 2, host_xy=(xp, yp))
    ms_e2e, _ = timed_region(thetas[W:], K, host_xy=(xp, yp))
    e2e_value = world * K / (ms_e2e * 1e-3)
    h2d = 8 * (n * d + n * p + 2 + d)
    d2h = 8 * (3 + d) + 4

    # ---- sustained: the same loop for >= 3 s (power / clock behaviour of a long run) -------------------
    sustained = None
    if leg("sustained"):
        per = max(K, 10)
        sampler = ClockSampler(local)
        sampler.start()
        t_ms, count = 0.0, 0
        while t_ms < 3000.0:
            ms_i, _ = timed_region(thetas[W:], per)
            t_ms += ms_i
            count += per
        sc = sampler.stop()
        sustained = {"value": world * count / (t_ms * 1e-3), "unit": UNIT, "seconds": t_ms * 1e-3, "evals_per_gpu": count, "clocks": sc}

    # ---- stage timing for the roofline detail (separate pass; event records perturb nothing measurable) --
    gp.set_stage_timing(True)
    stage = []
    for i in range(min(K, 5)):
        gp.lml_grad(thetas[W + i])
        stage.append(gp.last_stage_ms())
    gp.set_stage_timing(False)
    stage_mean = {k: float(np.mean([s[k] for s in stage])) for k in stage[0]}
    third = float(n) ** 3 / 3.0
    per_stage = {k: third / (stage_mean[k] * 1e-3) * 1e-12 for k in ("potrf", "trtri", "lauum")}
    peak = fp64["sustained_tflops"]
    job_tflops = value / world * f_eval(n, p) * 1e-12
    ncu = ncu_inputs()
    roofline = {
        "bound": "tensor", "achieved": job_tflops, "peak": peak, "unit": "TFLOP/s", "frac": job_tflops / peak,
        "traffic": ncu.get("eval_dram_bytes"),
        "what": f"job level: F_eval = N^3 + 3 N^2 P = {f_eval(n, p):.4g} FLOP per evaluation x measured evaluations/s per GPU ({C} in flight)",
        "peak_source": "cuBLAS Dgemm 8192^3 sustained, measured in this run (MEASURED_PEAKS.json has no FP64 entry); burst "
                       f"{fp64['burst_tflops']:.2f} TFLOP/s",
        "algorithmic_bytes_per_eval": 16.0 * n * n + 16.0 * n * d + 8.0 * n * p,
        "traffic_source": ncu.get("file"),
        "detail": {
            "one_evaluation_alone": {"ms": stage_mean["total"], "tflops": f_eval(n, p) / (stage_mean["total"] * 1e-3) * 1e-12,
                                     "frac": f_eval(n, p) / (stage_mean["total"] * 1e-3) * 1e-12 / peak},
            "stage_ms": stage_mean,
            "stage_tflops_n3_over_3": per_stage,
            "stage_frac": {k: v / peak for k, v in per_stage.items()},
            "ncu": ncu,
        },
    }

    # ---- second headline: predicted cell-depths/s (predict + modes->cells on device, ring buffer) --
    predict = None
    c_cells = 200_000
    cm = make_cell_map(p, c_cells, seed=0)
    e_mean, bias = fold_cell_map(cm.eofs, cm.x_mean, cm.x_std, cm.weights, cm.input_mean, cm.dry_indices, cm.elevations)
    gp.condition(thetas[0])
    gp.set_cell_map(e_mean, bias)
    xt = torch.from_numpy(data.x_test).cuda()
    gp.predict_cells(xt, want_modes=False)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if args.quick else 12
    p0.record(stream)
    for _ in range(reps):
        gp.predict_cells(xt, want_modes=False)
    p1.record(stream)
    barrier()
    pms = reduce_max(p0.elapsed_time(p1))
    t_events = world * reps * xt.shape[0]
    ev_s = t_events / (pms * 1e-3)
    f_evt = float(n) ** 2 + 2.0 * n * p + 4.0 * p * c_cells  # SURVEY.md section 8d (general variance map: two P x C products)
    f_exec = float(n) ** 2 + 2.0 * n * p + 2.0 * p * c_cells  # what runs: with one theta for all modes the cell variance is rank one
    predict = {
        "metric": "predicted cell-depths/s", "value": ev_s * c_cells, "unit": "cell-depths/s",
        "events_per_s": ev_s, "cells": c_cells, "events_per_rank": int(xt.shape[0]) * reps, "seconds": pms * 1e-3,
        "output": "mean+variance per cell written to a device ring buffer (T*C*16 B cannot be kept)",
        "roofline": {
            "bound": "tensor", "achieved": ev_s / world * f_exec * 1e-12, "peak": peak, "unit": "TFLOP/s",
            "frac": ev_s / world * f_exec * 1e-12 / peak,
            "what": "executed FP64 FLOP per event, N^2 + 2 N P + 2 P C (variance product, mean, modes->cells mean; the cell variance is "
                    "rank one under a shared theta) x events/s per GPU",
            "with_survey_f_evt": {"flops_per_event": f_evt, "achieved": ev_s / world * f_evt * 1e-12, "frac": ev_s / world * f_evt * 1e-12 / peak,
                                  "what": "SURVEY.md 8d's F_evt = N^2 + 2 N P + 4 P C counts a second P x C product for the variance"},
            "hbm": {"achieved_GBps": ev_s / world * 16.0 * c_cells * 1e-9, "peak_GBps": hbm_gbs,
                    "frac": ev_s / world * 16.0 * c_cells * 1e-9 / hbm_gbs, "what": "16 B written per cell-depth (mean + variance)",
                    "peak_source": hbm_src},
        },
    }
    if leg("predict"):
        # the same sweep with the consumer fused: depth conversion + every metric of gpras/metrics.py accumulated on the fly
        # against a resident truth block, the T x C prediction never written
        from gpras_b200.metrics import MetricsAccumulator

        mreps = 3
        acc = MetricsAccumulator(c_cells, mreps * int(xt.shape[0]) + 1, device=local)
        acc.set_elevations(cm.elevations, cm.elevations)
        gtr = torch.Generator(device="cuda").manual_seed(7 + rank)
        truth = torch.rand(int(xt.shape[0]), c_cells, dtype=torch.float64, device="cuda", generator=gtr) * 4 + 3
        acc.reset(0.0)
        acc.predict_update(gp, xt, truth)
        barrier()
        p0.record(stream)
        acc.reset(0.0)
        for _ in range(mreps):
            acc.predict_update(gp, xt, truth)
        summary = acc.finalize(0.5)
        p1.record(stream)
        barrier()
        fms = reduce_max(p0.elapsed_time(p1))
        predict["fused_metrics"] = {
            "value": world * mreps * xt.shape[0] * c_cells / (fms * 1e-3), "unit": "cell-depths/s",
            "what": "predict -> cells -> depth -> all gpras/metrics.py reductions vs a resident truth block; nothing written per cell-depth",
            "rmse_aoi_toi": summary["rmse_aoi_toi"],
        }
        acc.close()
        del truth

    # ---- BASELINE config 5: the stochastic prediction sweep, 1,000,000 events x 200,000 cells, events sharded over ranks ----
    cfg5 = None
    if leg("cfg5"):
        from gpras_b200.parallel import all_gather_rows, shard_rows

        t_all = 1_000_000
        lo, hi = shard_rows(t_all, rank, world)
        gen = torch.Generator(device="cuda").manual_seed(11)
        x_sweep = torch.randn((t_all, d), dtype=torch.float64, device="cuda", generator=gen)[lo:hi].contiguous()
        # mode-space results stay on the device, are all-gathered over NCCL (the path's one exchange) and read back once
        mm = torch.empty((hi - lo, p), dtype=torch.float64, device="cuda")
        mv = torch.empty((hi - lo, p), dtype=torch.float64, device="cuda")
        even = t_all % world == 0
        host_out = torch.empty((2, t_all, p), dtype=torch.float64).pin_memory() if rank == 0 else None
        barrier()
        t0 = time.perf_counter()
        gp.predict_cells(x_sweep, modes_out=(mm, mv))
        if world > 1 and even:
            full = torch.empty((2, t_all, p), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(full[0], mm)
            dist.all_gather_into_tensor(full[1], mv)
        elif world > 1:
            parts = all_gather_rows(torch.cat([mm, mv], dim=1).cpu().numpy(), 2 * p)
            full = torch.from_numpy(np.stack([parts[:, :p], parts[:, p:]])).cuda()
        else:
            full = torch.stack([mm, mv])
        if rank == 0:
            host_out.copy_(full, non_blocking=True)
        torch.cuda.synchronize()
        dt5 = reduce_max(time.perf_counter() - t0)
        both = full[0]
        cfg5 = {"workload": f"cfg5: trained N={n} surrogate predicting {t_all} events x {c_cells} cells (mean + variance), events sharded over "
                            f"{world} rank(s); cell-space output to a device ring buffer, mode-space (T x P) means and variances all-gathered "
                            "over NCCL and read back by rank 0",
                "scaling": "strong", "wall_s": dt5, "events_per_s": t_all / dt5, "cell_depths_per_s": t_all * c_cells / dt5,
                "mode_rows_gathered": int(both.shape[0]), "events_per_rank": hi - lo,
                "mode_var_min": float(full[1].min().item())}
        del full, host_out
        del x_sweep, mm, mv, both

    # ---- BASELINE config 3 as stated: 64 optimiser restarts sharded over the GPUs (fixed total work) ----
    strong = None
    if leg("strong"):
        from gpras_b200 import GPRAS
        from gpras_b200.parallel import all_gather_rows
        from gpras_b200.synth import random_starts

        for g_ in gps[1:]:
            g_.close()
        gps = gps[:1]
        r_total, maxiter = 64, 15
        st3 = random_starts(r_total, 1, seed=2)  # the reference's own ranges (gpras/gpr.py:88-90), one scalar lengthscale per start
        starts = np.concatenate([st3[:, :2], np.repeat(st3[:, 2:3], d, axis=1)], axis=1)
        model = GPRAS(w["kernel"])
        lanes = int(os.environ.get("BENCH_RESTART_LANES", min(C, 2)))  # 64 restarts over 8 GPUs x 2 lanes = 4 rounds
        # warm-up: a one-iteration restart per lane creates the lanes' device handles (1.6 GB each) before the timed region,
        # like the handles of the weak leg
        model.fit(data.x, data.y, None, "kmeans", "L-BFGS-B", ard=True, shared_kernel=True, device=local, restarts=starts[: lanes * world],
                  restart_lanes=lanes, max_iter=1)
        barrier()
        t0 = time.perf_counter()
        model.fit(data.x, data.y, None, "kmeans", "L-BFGS-B", ard=True, shared_kernel=True, device=local, restarts=starts,
                  restart_lanes=lanes, max_iter=maxiter)
        torch.cuda.synchronize()
        dts = reduce_max(time.perf_counter() - t0)
        m0 = model.models[0]
        stats = all_gather_rows(np.array([[float(m0.n_evals), float(getattr(m0, "restart_busy_s", 0.0))]]), 2)  # (evaluations of the timed fit only: fit() builds new models)
        tab = m0.restart_table
        strong = {
            "workload": f"cfg3 as stated: {r_total} optimiser restarts (starts log-uniform in the reference's ranges) x L-BFGS-B(maxiter {maxiter}) "
                        f"through GPRAS.fit(restarts=...), handed out to {world} rank(s) by a ticket counter, {C} restarts in flight per GPU",
            "scaling": "strong", "wall_s": dts, "evals": int(stats[:, 0].sum()), "evals_per_s": float(stats[:, 0].sum()) / dts,
            "restarts_per_s": r_total / dts, "best_loss": float(np.min(tab[:, 1])), "restarts_not_positive_definite": int(np.sum(~np.isfinite(tab[:, 1]))),
            "rank_busy_s": [float(v) for v in stats[:, 1]], "host_cores": os.cpu_count(), "host_threads_per_rank": lanes,
            "imbalance": float(stats[:, 1].max() / max(stats[:, 1].mean(), 1e-12) - 1.0),
        }
        model._slot.release_other_threads()

    # ---- BASELINE config 4: N = 16384, D = P = 64 (one GPU's worth of FP64 Cholesky near the HBM footprint of 3 N^2 doubles) ----
    cfg4 = None
    if leg("cfg4") and rank == 0:
        n4, d4 = 16384, 64
        data4 = make_gp_data(n4, d4, d4, 0, seed=0)
        g4 = ExactGP(w["kernel"], n4, d4, d4, device=local)
        g4.set_data(data4.x, data4.y)
        from gpras_b200.synth import fixed_theta

        v4, s4, ls4 = fixed_theta(d4, True)
        th4 = g4.theta_vector(v4, s4, ls4)
        g4.lml_grad(th4)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        reps4 = 3
        for _ in range(reps4):
            g4.lml_grad(th4)
        e1.record(stream)
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / reps4
        cfg4 = {"workload": f"cfg4: exact GP LML+grad, N={n4}, D=P={d4}, Matern52 ARD, one evaluation at a time on one GPU",
                "ms_per_eval": ms4, "evals_per_s": 1e3 / ms4, "tflops": f_eval(n4, d4) / (ms4 * 1e-3) * 1e-12,
                "frac_of_dgemm": f_eval(n4, d4) / (ms4 * 1e-3) * 1e-12 / peak, "device_bytes": 3 * 8 * n4 * n4}
        g4.close()
        del data4

    # ---- the cells -> modes side at BASELINE config 2 sizes: PCA fit and the forward transform ----
    preprocess = None
    if rank == 0 and leg("preprocess"):
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
                      "transform_GBps": 8.0 * ns * cells / tr_s * 1e-9, "transform_hbm_frac": 8.0 * ns * cells / tr_s * 1e-9 / hbm_gbs}
        pp.close()
        del xs_
        # the same at the cell count of configs 3-5 (8 192 samples x 200 000 cells, 13 GB resident): the stream is long enough for
        # launch and call overhead not to matter
        try:
            ns3, cells3 = 8192, 200_000
            xs3, elev3, w3 = flood_tensor(torch, ns3, cells3)
            big = {}
            for modes3 in (16, 32):
                pp3 = PreProcessor(hydraulic_parameter="wse", device=local)
                pp3.fit(xs3, elev3, w3, modes3)
                pp3.transform(xs3)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    pp3.transform(xs3)
                torch.cuda.synchronize()
                tr3 = (time.perf_counter() - t0) / 3
                big[f"modes_{modes3}"] = {"fit_ms": pp3.fit_info["stage_ms"]["total"], "transform_ms": tr3 * 1e3,
                                          "transform_GBps": 8.0 * ns3 * cells3 / tr3 * 1e-9,
                                          "transform_hbm_frac": 8.0 * ns3 * cells3 / tr3 * 1e-9 / hbm_gbs}
                pp3.close()
            preprocess["cfg3_cells"] = {"workload": f"{ns3} samples x {cells3} cells resident in HBM", **big}
            del xs3
        except Exception as exc:  # (memory on a shared box): the config-2 numbers above stand
            preprocess["cfg3_cells"] = {"skipped": str(exc)[:200]}
        torch.cuda.empty_cache()

    sparse_fit = None
    if rank == 0 and world == 1 and leg("sparse"):  # (GPRAS.fit shards per-column models over ranks: a one-rank call would hang)
        # The reference's DEFAULT call at the reference's own scale (gpras/gpr.py:237-275: one SGPR per column, "two-stage" =
        # Adam 100 steps on the inducing inputs + Adam 100 steps on the hyperparameters): device-resident batched trainer vs
        # the host-driven loops of the same device evaluation.
        from gpras_b200 import GPRAS
        from gpras_b200.synth import make_gp_data as _mk

        ns_, ds_, ms_, ps_ = 5000, 10, 50, 10
        sd = _mk(ns_, ds_, ps_, 0, seed=3)
        secs, pars = {}, {}
        for name, kw in (("device_trainer", {}), ("host_lockstep", {"device_trainer": False}), ("sequential", {"lockstep_models": False})):
            for rep in range(2):  # first call creates handles / graphs
                gs = GPRAS("Matern52")
                t0 = time.perf_counter()
                gs.fit(sd.x, sd.y, ms_, "grid", "two-stage", **kw)
                secs[name] = time.perf_counter() - t0
            pars[name] = np.concatenate([np.concatenate([m_.theta(), np.asarray(m_.inducing_variable.Z).ravel()]) for m_ in gs.models])
            steps_ = sum(m_.n_evals for m_ in gs.models)
            gs.release()
        err = float(np.max(np.abs(pars["device_trainer"] - pars["sequential"]) / np.maximum(np.abs(pars["sequential"]), 1e-3)))
        sparse_fit = {"workload": f"reference default fit: N={ns_}, D={ds_}, M={ms_}, {ps_} per-column SGPR models, two-stage Adam 100 + 100, "
                                  "grid inducing inputs",
                      "seconds": secs, "model_steps": steps_, "device_trainer_us_per_lockstep_iteration": secs["device_trainer"] / 200 * 1e6,
                      "speedup_vs_sequential": secs["sequential"] / secs["device_trainer"],
                      "max_rel_diff_of_fitted_parameters_vs_sequential": err, "launches_per_iteration": "6 per lane of models (two staggered lanes when there are four or more models)"}
        if world == 1 and not args.no_cpu_baseline:
            import torch as _t

            from oracle import sgpr as _sg

            z_ = GPRAS("Matern52")._create_inducing(sd.x, ms_, "grid")
            t0 = time.perf_counter()
            for _ in range(3):
                _sg.training_loss_and_grads("Matern52", sd.x, sd.y[:, :1], z_, 0.54, np.array([0.54]), 0.54)
            sparse_fit["cpu_port_ms_per_loss_and_grad"] = (time.perf_counter() - t0) / 3 * 1e3
            sparse_fit["cpu_port_fit_seconds_extrapolated"] = sparse_fit["cpu_port_ms_per_loss_and_grad"] * 1e-3 * steps_
            del _t

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dtc, cores = cpu_port_eval_seconds(data, thetas, 1, 0)
        cpu = {"value": 1.0 / dtc, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 full LML+grad eval at N={n} (NumPy/SciPy-OpenBLAS oracle port, all host threads), {dtc:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_block(),
            "run": {"evaluations_per_gpu": K, "in_flight_per_gpu": C, "timed_region_s": ms_total * 1e-3, "host_cores": os.cpu_count(),
                    "exchange": "all-gather of the per-restart [LML, grad] rows inside both timed regions"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_eval * K,
            "verified": {"concurrent_equals_serial_bitwise": verified},
            "roofline": roofline,
            "fp64_peak": fp64,
            "sustained": sustained,
            "stage_ms": stage_mean,
            "cpu_baseline": cpu,
            "predict": predict,
            "cfg4": cfg4,
            "cfg5_sweep": cfg5,
            "strong": strong,
            "preprocess": preprocess,
            "sparse_fit": sparse_fit,
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
    ap.add_argument("--quick", action="store_true", help="headline + predict legs only (profiling runs)")
    ap.add_argument("--concurrent", type=int, default=4, help="independent evaluations in flight per GPU (throughput saturates at 4)")
    ap.add_argument("--legs", default="all", help="comma list of extra legs to run (sustained,predict,cfg5,strong,cfg4,preprocess,sparse) or 'all'")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
