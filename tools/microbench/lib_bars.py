"""Library bars on the box (NOT product code): cuBLAS DGEMM, cuSOLVER potrf / potri via torch float64."""
import json, time, torch
def t(f, n=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for N in (2048, 4096, 8192, 16384):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
    ms = t(lambda: a @ b, 3 if N > 8192 else 5)
    out[f"dgemm_{N}"] = {"ms": ms, "tflops": 2 * N**3 / ms * 1e-9}
    if N <= 8192:
        spd = a @ a.T + N * torch.eye(N, dtype=torch.float64, device="cuda")
        ms = t(lambda: torch.linalg.cholesky(spd))
        out[f"potrf_{N}"] = {"ms": ms, "tflops": N**3 / 3 / ms * 1e-9}
        L = torch.linalg.cholesky(spd)
        ms = t(lambda: torch.cholesky_inverse(L))
        out[f"potri_{N}"] = {"ms": ms, "tflops": 2 * N**3 / 3 / ms * 1e-9}
        y = torch.randn(N, 32, dtype=torch.float64, device="cuda", generator=g)
        ms = t(lambda: torch.cholesky_solve(y, L))
        out[f"potrs32_{N}"] = {"ms": ms}
        del spd, L
# sustained DGEMM 8192 for ~3 s
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda"); b = torch.randn(N, N, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); n = 0; t0 = time.time()
while time.time() - t0 < 3.0:
    for _ in range(5): a @ b
    n += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained"] = {"tflops": n * 2 * N**3 / e0.elapsed_time(e1) * 1e-9, "n": n}
print(json.dumps(out, indent=1))
