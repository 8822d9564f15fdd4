"""Per-kernel SASS instruction counts of libgpras_b200.so (cuobjdump -sass): the evidence that the FP64 tensor pipe (DMMA),
asynchronous copies (LDGSTS = cp.async) and the TMA unit (UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier) are what the
kernels execute.  Runs without a GPU.

    python tools/sass_counts.py [out.json]
"""
import collections
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "gpras_b200" / "libgpras_b200.so"
WANT = ["DMMA", "DFMA", "LDGSTS", "UTMALDG", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "MUFU", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w):
                    counts[cur][w] += 1
                    break
    names = demangle(list(counts))
    out = {}
    for k, c in counts.items():
        short = re.sub(r"^void ", "", names.get(k, k)).replace("gpras::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*$", "", short)
        out[short] = {w: c[w] for w in WANT if c[w]}
    dst = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r02" / "sass_counts.json"
    dst.parent.mkdir(parents=True, exist_ok=True)
    summary = {
        "what": "SASS instruction counts per kernel of gpras_b200/libgpras_b200.so (cuobjdump -sass, sm_100a)",
        "kernels_with_DMMA": sorted(k for k, v in out.items() if v.get("DMMA")),
        "kernels_with_TMA_UTMALDG": sorted(k for k, v in out.items() if v.get("UTMALDG")),
        "kernels_with_LDGSTS": sorted(k for k, v in out.items() if v.get("LDGSTS")),
        "kernels": out,
    }
    dst.write_text(json.dumps(summary, indent=1))
    print("wrote", dst, "| DMMA kernels:", len(summary["kernels_with_DMMA"]), "| TMA kernels:", summary["kernels_with_TMA_UTMALDG"])


if __name__ == "__main__":
    main()
