"""Localise the concurrency-dependent mismatch: building blocks on many streams at once vs alone."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from gpras_b200 import _lib
lib = _lib.load()
n, C = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
a = rng.standard_normal((n, n)); spd = a @ a.T / n + np.eye(n)
A0 = torch.from_numpy(spd).cuda()
streams = [torch.cuda.Stream() for _ in range(C)]
def run_all(which):
    outs = []
    As = [A0.clone() for _ in range(C)]
    Ws = [torch.zeros(n, n, dtype=torch.float64, device="cuda") for _ in range(C)]
    Ts = [torch.zeros_like(Ws[0]) for _ in range(C)]
    Ks = [torch.zeros_like(Ws[0]) for _ in range(C)]
    lds = [torch.zeros(n // 128, dtype=torch.float64, device="cuda") for _ in range(C)]
    infos = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(C)]
    torch.cuda.synchronize()
    for c in range(C):
        st = streams[c].cuda_stream
        _lib.check(lib.gpras_dpotrf(st, As[c].data_ptr(), n, Ws[c].data_ptr(), n, n, lds[c].data_ptr(), infos[c].data_ptr()))
        if which >= 1:
            _lib.check(lib.gpras_dtrtri(st, As[c].data_ptr(), n, Ws[c].data_ptr(), n, Ts[c].data_ptr(), n, n))
        if which >= 2:
            _lib.check(lib.gpras_dlauum(st, Ws[c].data_ptr(), n, Ks[c].data_ptr(), n, n))
    torch.cuda.synchronize()
    return [torch.tril(x).cpu().numpy() for x in As], [torch.tril(x).cpu().numpy() for x in Ws], [torch.tril(x).cpu().numpy() for x in Ks]
for which, name in [(0, "potrf"), (1, "potrf+trtri"), (2, "potrf+trtri+lauum")]:
    for rep in range(3):
        L, W, K = run_all(which)
        badL = sum(not np.array_equal(L[c], L[0]) for c in range(C))
        badW = sum(not np.array_equal(W[c], W[0]) for c in range(C))
        badK = sum(not np.array_equal(K[c], K[0]) for c in range(C))
        print(f"{name:20s} rep {rep}: L mismatches {badL}, W {badW}, Kinv {badK}", flush=True)
