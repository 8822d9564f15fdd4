"""CPU: the committed bench lines (profiles/r01/bench_*_final.json, produced by `bench.py` on B200s) carry every key of the
driver's contract, and bench.py's argument surface is the contracted one.  Guards the JSON schema against accidental edits."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


@pytest.mark.parametrize("name", ["r01/bench_1gpu_final.json", "r01/bench_2gpu_final.json", "r01/bench_8gpu_final.json",
                                  "r02/bench_1gpu_final.json", "r02/bench_2gpu_dev.json", "r02/bench_2gpu_final.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == "LML+grad evals/s at N=8192" and d["unit"] == "evals/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] is not None
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert d["verified"]["concurrent_equals_serial_bitwise"] is True
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert set(c) >= {"value", "unit", "cores", "kind", "sample"} and c["kind"] in ("reference", "port")


@pytest.mark.parametrize("rnd", ["r01", "r02"])
def test_reference_arm_line_follows_the_contract(rnd):
    d = json.loads((ROOT / "profiles" / rnd / "bench_reference_final.json").read_text().strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "LML+grad evals/s at N=8192" and d["unit"] == "evals/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"


def test_round2_bench_line_is_job_level_and_both_arms_share_the_config():
    ours = json.loads((ROOT / "profiles" / "r02" / "bench_1gpu_final.json").read_text().strip().splitlines()[-1])
    ref = json.loads((ROOT / "profiles" / "r02" / "bench_reference_final.json").read_text().strip().splitlines()[-1])
    assert ours["config"] == ref["config"]  # the driver compares them
    r = ours["roofline"]
    # top-level fraction = F_eval x evals/s per GPU over the DGEMM rate measured in the run, not a single launch
    f_eval = 8192.0**3 + 3 * 8192.0**2 * 32
    assert abs(r["achieved"] - ours["value"] / ours["n_gpus"] * f_eval * 1e-12) < 1e-9 * r["achieved"]
    assert abs(r["peak"] - ours["fp64_peak"]["sustained_tflops"]) < 1e-12 and "detail" in r
    assert ours["sustained"]["seconds"] >= 3.0 and ours["predict"]["roofline"]["frac"] < 1.0
    for k in ("cfg4", "cfg5_sweep", "strong"):
        assert ours[k] is not None, k
    assert ours["strong"]["scaling"] == "strong" and ours["cfg5_sweep"]["scaling"] == "strong"
    assert ref["steps"] <= 3 and ref["steps_requested"] >= ref["steps"]


def test_round2_bench_line_carries_the_sparse_family_and_the_large_transform():
    """The reference's default model family (per-column sparse models) is timed through ``GPRAS.fit`` itself, against the
    host-driven loops of the same device evaluation and with the fitted parameters compared; one-rank runs only (under
    torch.distributed ``fit`` shards the models over ranks)."""
    ours = json.loads((ROOT / "profiles" / "r02" / "bench_1gpu_final.json").read_text().strip().splitlines()[-1])
    sf = ours["sparse_fit"]
    assert set(sf["seconds"]) == {"device_trainer", "host_lockstep", "sequential"}
    assert sf["seconds"]["device_trainer"] < sf["seconds"]["host_lockstep"] < sf["seconds"]["sequential"]
    assert sf["max_rel_diff_of_fitted_parameters_vs_sequential"] < 1e-8
    big = ours["preprocess"]["cfg3_cells"]["modes_16"]
    assert 0.0 < big["transform_hbm_frac"] < 1.0 and abs(big["transform_GBps"] / 6543.7 - big["transform_hbm_frac"]) < 1e-9
    two = json.loads((ROOT / "profiles" / "r02" / "bench_2gpu_final.json").read_text().strip().splitlines()[-1])
    assert two["n_gpus"] == 2 and two["sparse_fit"] is None and two["strong"]["scaling"] == "strong"


def test_bench_ncu_figures_come_from_profiles():
    src = (ROOT / "bench.py").read_text()
    assert "bench_ncu_inputs.json" in src and "lauum_dram_bytes" not in src  # no ncu constants typed into bench.py
    d = json.loads(sorted((ROOT / "profiles").glob("r*/bench_ncu_inputs.json"))[-1].read_text())
    assert d["eval_dram_bytes"] > 0 and d["eval_launches"] > 0


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True, check=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out
