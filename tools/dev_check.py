"""Developer smoke check on a GPU box (not a test): engine GEMMs, potrf/trtri/lauum, LML+grad, predict."""
import sys, time
import numpy as np
import torch

sys.path.insert(0, ".")
from gpras_b200 import _lib
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta
from oracle.exact_gp import Theta, lml_and_grad, predict

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


# ---- engine GEMM, four layouts ----
m, n, k = 256, 384, 160
for akm in (0, 1):
    for bkm in (0, 1):
        A = torch.randn(k, m, dtype=torch.float64, device=dev) if akm else torch.randn(m, k, dtype=torch.float64, device=dev)
        B = torch.randn(k, n, dtype=torch.float64, device=dev) if bkm else torch.randn(n, k, dtype=torch.float64, device=dev)
        Cc = torch.randn(m, n, dtype=torch.float64, device=dev)
        ref = 0.7 * ((A.T if akm else A) @ (B if bkm else B.T)) - 0.3 * Cc
        _lib.check(lib.gpras_dgemm_tiles(st, 0, akm, bkm, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cc.data_ptr(), Cc.stride(0), m, n, k, 0.7, -0.3))
        torch.cuda.synchronize()
        print("gemm", akm, bkm, rel(Cc, ref))

# ---- potrf / trtri / lauum ----
for N in (128, 256, 384, 1024):
    a = torch.randn(N, N, dtype=torch.float64, device=dev)
    spd = a @ a.T / N + torch.eye(N, dtype=torch.float64, device=dev)
    Lref = torch.linalg.cholesky(spd)
    A = spd.clone()
    W = torch.zeros(N, N, dtype=torch.float64, device=dev)
    T = torch.zeros(N, N, dtype=torch.float64, device=dev)
    Kinv = torch.zeros(N, N, dtype=torch.float64, device=dev)
    ld = torch.zeros(N // 128, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.gpras_dpotrf(st, A.data_ptr(), N, W.data_ptr(), N, N, ld.data_ptr(), info.data_ptr()))
    _lib.check(lib.gpras_dtrtri(st, A.data_ptr(), N, W.data_ptr(), N, T.data_ptr(), N, N))
    _lib.check(lib.gpras_dlauum(st, W.data_ptr(), N, Kinv.data_ptr(), N, N))
    torch.cuda.synchronize()
    Wref = torch.linalg.inv(Lref)
    print("N", N, "info", int(info), "L", rel(torch.tril(A), Lref), "W", rel(torch.tril(W), Wref), "Kinv", rel(torch.tril(Kinv), torch.tril(torch.linalg.inv(spd))),
          "logdet", float(ld.sum()), float(torch.log(torch.diagonal(Lref)).sum()))

# ---- LML + grad + predict vs oracle ----
for (name, ard, N, D, P) in [("RBF", False, 256, 8, 8), ("Matern52", True, 300, 5, 3), ("Matern32", True, 200, 16, 16), ("Matern12", False, 130, 4, 2), ("Exponential", False, 257, 3, 1)]:
    data = make_gp_data(N, D, P, 77)
    v, s, ls = fixed_theta(D, ard)
    gp = ExactGP(name, N, D, P)
    gp.set_data(data.x, data.y)
    th = gp.theta_vector(v, s, ls)
    lml, g = gp.lml_grad(th)
    o = lml_and_grad(name, data.x, data.y, Theta(v, s, ls))
    og = np.concatenate([[o[1], o[2]], o[3]])
    gp.condition(th)
    mean, var = gp.predict(data.x_test)
    om, ov = predict(name, data.x, data.y, Theta(v, s, ls), data.x_test)
    print(name, N, D, P, "lml rel", abs(lml - o[0]) / abs(o[0]), "grad rel", np.max(np.abs(g - og) / np.abs(og)),
          "mean", np.max(np.abs(mean - om)) / np.max(np.abs(om)), "std", np.max(np.abs(np.sqrt(var) - np.sqrt(ov)) / np.sqrt(ov)), "launches", gp.last_launches())
    gp.close()

# ---- timing at N=8192 ----
for (N, D, P) in [(2048, 16, 16), (8192, 32, 32)]:
    data = make_gp_data(N, D, P, 128)
    v, s, ls = fixed_theta(D, True)
    gp = ExactGP("Matern52", N, D, P)
    gp.set_data(data.x, data.y)
    th = gp.theta_vector(v, s, ls)
    gp.set_stage_timing(True)
    for _ in range(3):
        t0 = time.time(); lml, g = gp.lml_grad(th); dt = time.time() - t0
        print(N, "lml", lml, "wall ms", dt * 1e3, gp.last_stage_ms(), "launches", gp.last_launches())
    gp.close()
