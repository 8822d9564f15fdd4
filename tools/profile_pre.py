"""One PreProcessor.fit + transform + fused predict->cells->metrics pass at cfg3 sizes (for ncu launch lists)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch

from bench_pre_metrics import bench_metrics, flood_tensor
from gpras_b200.preprocess import PreProcessor

n, c, p = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 200000, 32)))
x, elev, w = flood_tensor(torch, n, c)
pp = PreProcessor(hydraulic_parameter="wse")
pp.fit(x, elev, w, p)
print(pp.fit_info["stage_ms"], pp.fit_info["iterations"])
z = pp.transform(x)
torch.cuda.synchronize()
pp.close()
del x
if len(sys.argv) <= 4:
    bench_metrics(torch, n, p, p, c, 2048, "profile")
