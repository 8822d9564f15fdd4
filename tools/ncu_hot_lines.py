"""Summarise an ncu report's warp-stall samples per CUDA source line: python tools/ncu_hot_lines.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""; data = []; tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "") : continue
    try: line = int(r[0]); samples = int(r[4]); inst = int(r[7])
    except ValueError: continue
    data.append((samples, cur_file, line, inst, r[1].strip()[:100])); tot += samples
print("total samples", tot)
for s, f, l, i, src in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {100*s/max(tot,1):5.1f}%  {f}:{l:<4} inst={i:>8}  {src}")
