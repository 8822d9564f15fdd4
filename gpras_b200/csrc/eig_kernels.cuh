// Small-matrix pieces of the PCA fit (PreProcessor.fit, gpras/preprocess.py:988-1007; scikit-learn IncrementalPCA there):
// the leading eigenpairs of the N x N Gram matrix G = Xw Xw^T come from a blocked subspace iteration whose dense
// products run on the DMMA engine (gemm_engine.cuh); the 128 x 128 Rayleigh-Ritz problem of every iteration is solved
// here by a one-CTA, one-sided (Hestenes) Jacobi sweep in shared memory.
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int EIG_B = 128;            // block size of the subspace iteration == engine tile
constexpr int EIG_THREADS = 1024;     // 64 column pairs x 16 threads: a half-warp reads 16 consecutive rows of one column
constexpr int EIG_LD = EIG_B + 2;     // column pitch (doubles)
constexpr int EIG_SMEM_BYTES = (EIG_B * EIG_LD + 2 * EIG_B) * (int)sizeof(double);

// Eigen-decomposition of the symmetric positive semi-definite 128 x 128 matrix H (symmetrised on load).
// One-sided Jacobi on M = H: plane rotations of column pairs until all columns are mutually orthogonal; then
// M = H V = V diag(lambda), so lambda_j = |m_j| and v_j = m_j / lambda_j.
//   lambda[128]   eigenvalues, descending
//   V  (128x128, pitch 128)  eigenvectors as columns, same order; columns with lambda_j <= cut * lambda_0 are zeroed
//   Vs (128x128, pitch 128)  V diag(1 / lambda) (zero for dropped columns)
static __global__ void __launch_bounds__(EIG_THREADS, 1)
jacobi_eig128_kernel(const double* __restrict__ H, long ldh, double cut, double* __restrict__ lambda, double* __restrict__ V,
                     double* __restrict__ Vs, int* __restrict__ sweeps_out) {
  extern __shared__ __align__(16) double smem[];
  double* M = smem;                    // column-major: M[col * EIG_LD + row]
  double* nrm = M + EIG_B * EIG_LD;    // [128]
  double* rnk = nrm + EIG_B;           // [128] (rank as double)
  const int tid = threadIdx.x;
  for (int e = tid; e < EIG_B * EIG_B; e += EIG_THREADS) {
    const int r = e >> 7, c = e & 127;
    M[c * EIG_LD + r] = 0.5 * (H[(long)r * ldh + c] + H[(long)c * ldh + r]);
  }
  __syncthreads();
  const int pair = tid >> 4, sub = tid & 15;
  int sweeps = 0;
  for (; sweeps < 40; sweeps++) {
    int rotated = 0;
    for (int round = 0; round < EIG_B - 1; round++) {
      int ca, cb;
      if (pair == 0) {
        ca = EIG_B - 1, cb = round;
      } else {
        ca = (round + pair) % (EIG_B - 1);
        cb = (round - pair + (EIG_B - 1)) % (EIG_B - 1);
      }
      double* pa = M + ca * EIG_LD + sub;
      double* pb = M + cb * EIG_LD + sub;
      double xa[8], xb[8];
      double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        xa[k] = pa[16 * k], xb[k] = pb[16 * k];
        al += xa[k] * xa[k], be += xb[k] * xb[k], ga += xa[k] * xb[k];
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        al += __shfl_xor_sync(0xffffffffu, al, o);
        be += __shfl_xor_sync(0xffffffffu, be, o);
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
      }
      if (fabs(ga) > 2e-15 * sqrt(al * be) && al > 0.0 && be > 0.0) {
        rotated = 1;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = rsqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          pa[16 * k] = cs * xa[k] - sn * xb[k];
          pb[16 * k] = sn * xa[k] + cs * xb[k];
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) break;
  }
  // norms, descending rank (ties by index), outputs
  if (tid < EIG_B) {
    double s = 0.0;
    for (int r = 0; r < EIG_B; r++) s += M[tid * EIG_LD + r] * M[tid * EIG_LD + r];
    nrm[tid] = sqrt(s);
  }
  __syncthreads();
  if (tid < EIG_B) {
    int rk = 0;
    const double mine = nrm[tid];
    for (int k = 0; k < EIG_B; k++) rk += (nrm[k] > mine) || (nrm[k] == mine && k < tid);
    rnk[tid] = (double)rk;
    lambda[rk] = mine;
  }
  __syncthreads();
  double top = 0.0;
  for (int k = 0; k < EIG_B; k++) top = fmax(top, nrm[k]);
  for (int e = tid; e < EIG_B * EIG_B; e += EIG_THREADS) {
    const int r = e >> 7, c = e & 127;
    const int dst = (int)rnk[c];
    const double l = nrm[c];
    const bool live = l > cut * top && l > 0.0;
    const double v = live ? M[c * EIG_LD + r] / l : 0.0;
    V[(long)r * EIG_B + dst] = v;
    Vs[(long)r * EIG_B + dst] = live ? v / l : 0.0;
  }
  if (tid == 0 && sweeps_out) *sweeps_out = sweeps;
}

// Deterministic pseudo-random start block Q0 (n_pad x 128), rows >= n zero.
static __global__ void subspace_init_kernel(double* __restrict__ Q, int n, int n_pad) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)n_pad * EIG_B) return;
  const int i = (int)(e >> 7);
  unsigned long long z = (unsigned long long)e * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  Q[e] = i < n ? (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5 : 0.0;
}

// Before the Cholesky of B = Z^T Z (columns of Z are unit vectors or exactly zero): a dead column (B_jj < 1/2) gets the
// identity row / column, so the factor exists and the corresponding column of Q = Z L^-T stays zero.
static __global__ void gram_guard_kernel(double* __restrict__ B, long ldb) {
  __shared__ int dead[EIG_B];
  const int tid = threadIdx.x;
  if (tid < EIG_B) dead[tid] = !(B[(long)tid * ldb + tid] >= 0.5);
  __syncthreads();
  for (int e = tid; e < EIG_B * EIG_B; e += blockDim.x) {
    const int r = e >> 7, c = e & 127;
    if (dead[r] || dead[c]) B[(long)r * ldb + c] = r == c ? 1.0 : 0.0;
  }
}

// A[i][j] for j > i  <-  A[j][i]   (the engine's triangular mode fills lower tiles only)
static __global__ void mirror_lower_full_kernel(double* __restrict__ A, int n, long ld) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < n && j > i) A[(long)i * ld + j] = A[(long)j * ld + i];
}

// Column norms of R (rows x 128): out[j] = sqrt(sum_i R[i][j]^2); one CTA per column, fixed order.
static __global__ void __launch_bounds__(256) colnorm128_kernel(const double* __restrict__ R, int rows, double* __restrict__ out) {
  __shared__ double red[8];
  const int j = blockIdx.x, tid = threadIdx.x;
  double s = 0.0;
  for (int i = tid; i < rows; i += 256) {
    const double v = R[(long)i * EIG_B + j];
    s += v * v;
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; i++) t += red[i];
    out[j] = sqrt(t);
  }
}

// EOF rows from Ft (c_pad x pn) = Xw^T U: E[j][c] = sign_j * Ft[c][j] / s_j with s_j = sqrt(lambda_j) and sign_j chosen
// so that the largest-magnitude entry of every component is positive (scikit-learn svd_flip, v-based).
// One CTA per mode; the arg-max scan runs in a fixed order (first maximum wins, like numpy.argmax).
static __global__ void __launch_bounds__(256) eof_finish_kernel(const double* __restrict__ Ft, long ldf, int c, long c_pad,
                                                                const double* __restrict__ lambda, double* __restrict__ E) {
  __shared__ double bestv[256];
  __shared__ int besti[256];
  const int j = blockIdx.x, tid = threadIdx.x;
  double bv = -1.0;
  int bi = 0;
  for (int k = tid; k < c; k += 256) {
    const double v = fabs(Ft[(long)k * ldf + j]);
    if (v > bv) bv = v, bi = k;
  }
  bestv[tid] = bv, besti[tid] = bi;
  __syncthreads();
  if (tid == 0) {
    for (int t = 1; t < 256; t++)
      if (bestv[t] > bestv[0] || (bestv[t] == bestv[0] && besti[t] < besti[0])) bestv[0] = bestv[t], besti[0] = besti[t];
  }
  __syncthreads();
  const double s = sqrt(lambda[j]);
  const double sg = Ft[(long)besti[0] * ldf + j] < 0.0 ? -1.0 : 1.0;
  const double f = s > 0.0 ? sg / s : 0.0;
  for (long k = tid; k < c_pad; k += 256) E[(long)j * c_pad + k] = k < c ? f * Ft[k * ldf + j] : 0.0;
}

}  // namespace gpras
