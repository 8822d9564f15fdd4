"""CPU error model for an FP64-accurate GEMM on the int8 tensor pipe (Ozaki scheme I) -- groundwork for DESIGN.md section 8 item 3.

Emulates, in NumPy integer arithmetic, what a tcgen05 `kind::i8` kernel would compute for the path's long-k products
(K^-1 = W^T W, V = W K*^T): every operand row (column) is scaled by a power of two, cut into `s` signed slices (6 bits for the first,
7 for the others: each fits int8), all slice pairs with p + q <= s + 1 are multiplied exactly (int32 accumulation is exact for
k <= 2^17) and the diagonals are summed in FP64.  Reports the error against the FP64 product for matrices taken from a real
GP evaluation, and what it does to the quantities the parity tests look at.

    python tools/ozaki_error_model.py [N]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import scipy.linalg as sla

from gpras_b200.synth import fixed_theta, make_gp_data
from oracle.kernels import cov


def slices(a, s, axis):
    """Power-of-two scaling along `axis` (the non-contracted index) and `s` signed int slices: a ~ 2^e * sum_p d_p 2^-sh_p."""
    amax = np.max(np.abs(a), axis=axis, keepdims=True)
    e = np.ceil(np.log2(np.where(amax > 0, amax, 1.0)))
    r = a / np.exp2(e)                      # |r| <= 1
    out, shifts, sh = [], [], 0
    for p in range(s):
        sh += 6 if p == 0 else 7
        d = np.rint(r * np.exp2(sh))
        r = r - d / np.exp2(sh)
        assert np.max(np.abs(d)) <= 64
        out.append(d)                      # integer-valued float64: d_p @ d_q is exact in FP64 BLAS (|sum| < 2^53)
        shifts.append(sh)
    return e, out, shifts


def ozaki_matmul(a, b, s):
    """a (m x k) @ b (k x n) with s slices per operand; pairs with p + q <= s + 1 (1-based)."""
    ea, da, sha = slices(a, s, axis=1)
    eb, db, shb = slices(b, s, axis=0)
    c = np.zeros((a.shape[0], b.shape[1]))
    n_prod = 0
    for diag in range(2 * s - 1, -1, -1):       # least significant diagonals first
        acc = None
        for p in range(s):
            q = diag - p
            if 0 <= q < s and p + q <= s - 1:
                t = da[p] @ db[q]               # exact integer product (fits int32 on the device for k <= 2^17)
                assert np.max(np.abs(t)) < 2**31
                acc = t if acc is None else acc + t
                n_prod += 1
        if acc is not None:
            # all pairs of one diagonal share the scale 2^-(sha[p] + shb[q]) only if the slice widths are equal beyond the
            # first; with 6 + 7 (p) bits: sha[p] + shb[q] = 12 + 7 (p + q) -- constant on a diagonal
            c += acc / np.exp2(12 + 7 * diag)
    return c * np.exp2(ea) * np.exp2(eb), n_prod


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
    d, p = 16, 8
    data = make_gp_data(n, d, p, 64, seed=0)
    v, s2, ls = fixed_theta(d, True)
    if len(sys.argv) > 2:
        s2 = float(sys.argv[2])             # smaller noise -> worse conditioning
    k = cov("Matern52", data.x, data.x, v, ls) + s2 * np.eye(n)
    L = np.linalg.cholesky(k)
    W = sla.solve_triangular(L, np.eye(n), lower=True)
    kinv = W.T @ W
    ks = cov("Matern52", data.x_test, data.x, v, ls)
    V = W @ ks.T
    var = v + s2 - np.sum(V * V, axis=0)
    print(f"N = {n}: cond(K) = {np.linalg.cond(k):.2e}, |W|max = {np.abs(W).max():.2e}")
    print("slices  int8 products  rel.err(W^T W)  rel.err(tr(K^-1 dK))  rel.err(pred. std)   [FP64 reference: LAPACK-accurate product]")
    dk = k - s2 * np.eye(n)                      # dK / dlog variance
    tr_ref = np.sum(kinv * dk)
    for s in (5, 6, 7, 8, 9):
        kin2, n_prod = ozaki_matmul(W.T.copy(), W, s)
        V2, _ = ozaki_matmul(W, ks.T.copy(), s)
        var2 = v + s2 - np.sum(V2 * V2, axis=0)
        e1 = np.max(np.abs(kin2 - kinv)) / np.max(np.abs(kinv))
        e2 = abs(np.sum(kin2 * dk) - tr_ref) / abs(tr_ref)
        e3 = np.max(np.abs(np.sqrt(var2) - np.sqrt(var)) / np.sqrt(var))
        print(f"{s:6d}  {n_prod:13d}  {e1:14.2e}  {e2:20.2e}  {e3:18.2e}")
    print("tolerances of the north_star: LML / mean 1e-8, std 1e-6 (gradients 1e-7 in the tests)")


if __name__ == "__main__":
    main()
