"""Microbenchmark (not product): pure-write and copy bandwidth, to calibrate the cells kernel."""
import torch
def t(f, n=5):
    f(); torch.cuda.synchronize(); best=1e9
    for _ in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
n = 2048*200064
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty(n, dtype=torch.float64, device="cuda")
ms = t(lambda: a.fill_(1.5)); print(f"fill 3.28 GB: {ms:.3f} ms -> {n*8/ms*1e-6:.0f} GB/s")
ms = t(lambda: (a.fill_(1.5), b.fill_(2.5))); print(f"fill 2 x 3.28 GB: {ms:.3f} ms -> {2*n*8/ms*1e-6:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy 3.28 GB: {ms:.3f} ms -> {2*n*8/ms*1e-6:.0f} GB/s (r+w)")
ring = torch.empty(256*200064, dtype=torch.float64, device="cuda")
ms = t(lambda: [ring.fill_(1.0) for _ in range(8)]); print(f"8 x fill 410 MB ring: {ms:.3f} ms -> {8*ring.numel()*8/ms*1e-6:.0f} GB/s")
