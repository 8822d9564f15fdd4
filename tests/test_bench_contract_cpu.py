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


@pytest.mark.parametrize("name", ["bench_1gpu_final.json", "bench_2gpu_final.json", "bench_8gpu_final.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = json.loads((ROOT / "profiles" / "r01" / name).read_text())
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


def test_reference_arm_line_follows_the_contract():
    d = json.loads((ROOT / "profiles" / "r01" / "bench_reference_final.json").read_text())
    assert d["impl"] == "reference" and d["metric"] == "LML+grad evals/s at N=8192" and d["unit"] == "evals/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True, check=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out
