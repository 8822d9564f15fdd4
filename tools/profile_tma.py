"""One pass of each TMA-streamed kernel at config-3 size (for ncu): PreProcessor.fit (colstats), transform with 16 and 32 modes
(project), metrics over resident arrays (plain metrics), and the general-variance modes -> cells expansion."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
import torch

from bench_pre_metrics import flood_tensor
from gpras_b200.metrics import MetricsAccumulator
from gpras_b200.preprocess import PreProcessor

n, c = 8192, 200000
x, elev, w = flood_tensor(torch, n, c)
for p in (16, 32):
    pp = PreProcessor(hydraulic_parameter="wse")
    pp.fit(x, elev, w, p)
    z = pp.transform(x)
    torch.cuda.synchronize()
    if p == 32:
        t = 4096
        rng = np.random.default_rng(0)
        mean, var = torch.from_numpy(rng.standard_normal((t, p))).cuda(), torch.from_numpy(rng.uniform(0.1, 1, (t, p))).cuda()
        pp.reverse_transform_device(mean, var)  # ring-buffer mode
        torch.cuda.synchronize()
    pp.close()
tt = 2048
acc = MetricsAccumulator(c, tt)
truth = x[:tt]
acc.reset(0.0)
acc.update(truth, truth + 0.1, truth * 0.01)
print(acc.finalize(0.5)["rmse_aoi_toi"])
acc.close()
