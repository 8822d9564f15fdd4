"""Turn the round-end captures in gpurun_out/final/ into the small, tracked summaries under profiles/<round>/.

    python tools/summarise_profiles.py [round_dir]

* ncu_launches_bench_summary.csv : per-kernel totals of the `gpu__time_duration` launch list of `bench.py --steps 2 --warmup 1`
* ncu_eval_traffic_summary.json  : per-kernel time / DRAM bytes / DMMA-pipe activity over ONE whole LML+grad evaluation
* ncu_full_<kernel>.json         : selected metrics of the `--set full` capture of each main kernel
"""
import collections, csv, json, re, subprocess, sys
from pathlib import Path

SRC = Path("gpurun_out/final")
DST = Path(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02")
DST.mkdir(parents=True, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("gpras::", "").replace("<unnamed>::", "")
    return re.sub(r"\(.*$", "", name)[:90]


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit.startswith("us") else (v if unit.startswith("ms") else v * 1e3))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def read_ncu_csv(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


# ---- 1. launch list of the bench command ----
rows = read_ncu_csv(SRC / "ncu_launches_bench.csv")
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0, 0.0])
    ms = to_ms(r["Metric Value"], r["Metric Unit"])
    a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
total = sum(a[1] for a in agg.values())
with open(DST / "ncu_launches_bench_summary.csv", "w") as f:
    f.write("kernel,launches,total_ms,share_of_gpu_time,max_ms\n")
    for k, (n, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{t:.3f},{t / total:.4f},{m:.3f}\n")
print("bench launch list:", len(rows), "launches,", round(total, 1), "ms of kernels")

# ---- 2. one whole evaluation: time, traffic, DMMA activity per kernel ----
rows = read_ncu_csv(SRC / "ncu_eval_traffic.csv")
per = collections.OrderedDict()
for r in rows:
    key = (r["ID"], short(r["Kernel Name"]))
    per.setdefault(key, {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
ev = collections.OrderedDict()
for (_, name), m in per.items():
    e = ev.setdefault(name, {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "dmma_weighted": 0.0})
    ms = to_ms(*m["gpu__time_duration.sum"])
    e["launches"] += 1; e["ms"] += ms
    e["dram_read_bytes"] += to_bytes(*m["dram__bytes_read.sum"]); e["dram_write_bytes"] += to_bytes(*m["dram__bytes_write.sum"])
    e["dmma_weighted"] += ms * float(m["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"][0].replace(",", ""))
tot_ms = sum(e["ms"] for e in ev.values())
out = {"what": "every launch of ONE LML+grad evaluation at N=8192, D=P=32 (ncu, cold cache, serialised)", "launches": len(per), "kernel_ms_total": tot_ms,
       "dram_bytes_total": sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in ev.values()), "kernels": {}}
for k, e in sorted(ev.items(), key=lambda kv: -kv[1]["ms"]):
    out["kernels"][k] = {"launches": e["launches"], "ms": round(e["ms"], 4), "share": round(e["ms"] / tot_ms, 4),
                         "dram_read_GB": round(e["dram_read_bytes"] / 1e9, 4), "dram_write_GB": round(e["dram_write_bytes"] / 1e9, 4),
                         "dmma_pipe_active_pct_time_weighted": round(e["dmma_weighted"] / e["ms"], 2) if e["ms"] else 0.0}
json.dump(out, open(DST / "ncu_eval_traffic_final_summary.json", "w"), indent=1)
print("eval:", len(per), "launches,", round(tot_ms, 2), "ms,", round(out["dram_bytes_total"] / 1e9, 2), "GB DRAM")

# ---- 3. --set full captures ----
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
full = {}
for rep in sorted(SRC.glob("full_*.csv")):
    raw = rep.read_text()
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    full[rep.stem.replace("full_", "")] = [{w: (d[idx[w]] + (" " + units[idx[w]] if units[idx[w]] else "")) for w in WANT if w in idx} for d in rr[2:]]
if full:
    json.dump(full, open(DST / "ncu_full_final.json", "w"), indent=1)
print("full captures:", {k: len(v) for k, v in full.items()})

# ---- 4. the TMA-streamed kernels: time, DRAM bytes and throughput per launch ----
tma_csv = SRC / "ncu_tma_kernels.csv"
if tma_csv.exists():
    per = collections.OrderedDict()
    for r in read_ncu_csv(tma_csv):
        per.setdefault((r["ID"], short(r["Kernel Name"]), r["Grid Size"]), {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    rows = []
    for (_, name, grid), m in per.items():
        ms = to_ms(*m["gpu__time_duration.sum"])
        byts = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
        rows.append({"kernel": name, "grid": grid, "ms": round(ms, 4), "dram_GB": round(byts / 1e9, 4), "dram_GBps": round(byts / ms * 1e-6, 1),
                     "dram_throughput_pct_of_peak": float(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0].replace(",", "")),
                     "dmma_pipe_pct": float(m["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"][0].replace(",", ""))})
    json.dump({"what": "one launch each at 8192 samples x 200 000 cells (tools/profile_tma.py), ncu per-launch metrics", "launches": rows},
              open(DST / "ncu_tma_kernels_summary.json", "w"), indent=1)
    print("tma kernels:", len(rows))

# ---- 5. the figures bench.py quotes from ncu ----
ev_k = out["kernels"]
def pick(sub):
    return [v for k, v in ev_k.items() if sub in k]
lau = full.get("lauum", [{}])[0]
inputs = {
    "source": f"{DST}/ncu_eval_traffic_final_summary.json, {DST}/ncu_full_final.json (one launch each, cold cache, serialised by ncu)",
    "eval_dram_bytes": out["dram_bytes_total"], "eval_launches": out["launches"], "eval_kernel_ms_serialised": out["kernel_ms_total"],
    "lauum_launch": {"dram_bytes": (pick("128, 128, 64, 32, 32, 3, 1>, 1, 1")[0]["dram_read_GB"] + pick("128, 128, 64, 32, 32, 3, 1>, 1, 1")[0]["dram_write_GB"]) * 1e9,
                     "dmma_pipe_active_pct": pick("128, 128, 64, 32, 32, 3, 1>, 1, 1")[0]["dmma_pipe_active_pct_time_weighted"]},
    "trailing_update_launches": {"dmma_pipe_active_pct_time_weighted": pick("128, 64, 32, 32, 16, 3, 2>")[0]["dmma_pipe_active_pct_time_weighted"],
                                 "dram_bytes": (pick("128, 64, 32, 32, 16, 3, 2>")[0]["dram_read_GB"] + pick("128, 64, 32, 32, 16, 3, 2>")[0]["dram_write_GB"]) * 1e9},
}
json.dump(inputs, open(DST / "bench_ncu_inputs.json", "w"), indent=1)
print("bench inputs written")
