// Single-CTA factorisation of one 128 x 128 diagonal block: L = chol(A_jj) and W = L^-1.
//
// The block lives in shared memory as S[128][132] doubles (row pitch 132 == 4 mod 16 makes every
// DMMA fragment LDS.64 conflict free).  L occupies the lower triangle (incl. diagonal); W is built
// transposed in the strictly-upper triangle (W[i][k] at S[k][i], i > k) with its diagonal 1/L_ii in
// dvec[], so L, W and the temporaries of the inverse all fit in one 135 KB tile.
//
//   1. potrf, blocked by 8 columns: warp 0 factors the 8-wide panel in registers with shuffles
//      (one rsqrt per column on the critical path), all 8 warps then apply the rank-8 update to the
//      trailing triangle with DMMA.8x8x4 on 8 x 8 tiles.
//   2. inverse by recursive doubling: 8 x 8 diagonal blocks by forward substitution (one thread per
//      column), then for b = 8, 16, 32, 64 every pair of adjacent blocks fills its off-diagonal block
//      W21 = -W22 (L21 W11) with two DMMA products (T = L21 W11 is parked in the pair's mirrored
//      upper block, which W21^T then overwrites).
//   3. L (upper zeroed) goes back over A_jj; W (upper zeroed) goes to the W buffer; sum(log L_ii) to
//      logdet_part[jb]; a non-positive pivot records info = global column + 1 (LAPACK convention).
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int LEAF_N = 128, LEAF_LD = 132, LEAF_THREADS = 256;
constexpr int LEAF_SMEM_BYTES = (LEAF_N * LEAF_LD + LEAF_N) * (int)sizeof(double);

__device__ __forceinline__ double leaf_getW(const double* S, const double* dvec, int i, int k) {
  return i == k ? dvec[i] : (i > k ? S[k * LEAF_LD + i] : 0.0);
}

__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_inv_kernel(double* __restrict__ A, long lda, double* __restrict__ W, long ldw,
                      double* __restrict__ logdet_part, int* __restrict__ info, int jb) {
  extern __shared__ __align__(16) double smem[];
  double* S = smem;
  double* dvec = smem + LEAF_N * LEAF_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  A += (long)jb * LEAF_N * lda + (long)jb * LEAF_N;
  W += (long)jb * LEAF_N * ldw + (long)jb * LEAF_N;

  // ---- load (lower triangle; upper zeroed) ----
  for (int e = tid; e < LEAF_N * (LEAF_N / 2); e += LEAF_THREADS) {
    int r = e >> 6, c2 = (e & 63) * 2;
    double2 v = *reinterpret_cast<const double2*>(A + (long)r * lda + c2);
    S[r * LEAF_LD + c2] = c2 <= r ? v.x : 0.0;
    S[r * LEAF_LD + c2 + 1] = c2 + 1 <= r ? v.y : 0.0;
  }
  __syncthreads();

  // ---- 1. blocked Cholesky ----
  double logsum = 0.0;  // warp 0, lane 0..7 partials
  for (int J = 0; J < 16; J++) {
    const int j0 = 8 * J;
    if (warp == 0) {
      double v[4][8];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        int r = j0 + lane + 32 * m;
#pragma unroll
        for (int c = 0; c < 8; c++) v[m][c] = r < LEAF_N ? S[r * LEAF_LD + j0 + c] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double dpiv = __shfl_sync(0xffffffffu, v[0][c], c);
        if (!(dpiv > 0.0)) {
          if (lane == 0) atomicCAS(info, 0, jb * LEAF_N + j0 + c + 1);
          dpiv = 1.0;
        }
        double rs = rsqrt(dpiv);
        rs = rs * (1.5 - 0.5 * dpiv * rs * rs);  // one Newton step: full double accuracy
        double sq = dpiv * rs;
        if (lane == c) {
          logsum += log(sq);
          dvec[j0 + c] = rs;
        }
#pragma unroll
        for (int m = 0; m < 4; m++) v[m][c] = (m == 0 && lane == c) ? sq : v[m][c] * rs;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; c2++) {
          double l = __shfl_sync(0xffffffffu, v[0][c], c2);
#pragma unroll
          for (int m = 0; m < 4; m++) v[m][c2] -= v[m][c] * l;
        }
      }
#pragma unroll
      for (int m = 0; m < 4; m++) {
        int r = j0 + lane + 32 * m;
        if (r < LEAF_N) {
#pragma unroll
          for (int c = 0; c < 8; c++) S[r * LEAF_LD + j0 + c] = (r >= j0 + c) ? v[m][c] : 0.0;
        }
      }
    }
    __syncthreads();
    // rank-8 update of the trailing lower triangle, 8x8 tiles (I >= Cc > J)
    const int nt = 15 - J;
    const int ntile = nt * (nt + 1) / 2;
    for (int t = warp; t < ntile; t += 8) {
      int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
      while (ti * (ti + 1) / 2 > t) ti--;
      int tj = t - ti * (ti + 1) / 2;
      const int I = J + 1 + ti, Cc = J + 1 + tj;
      double* cp = S + (8 * I + g) * LEAF_LD + 8 * Cc + 2 * q;
      double2 cv = *reinterpret_cast<double2*>(cp);
      double c0 = -cv.x, c1 = -cv.y;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        double a = S[(8 * I + g) * LEAF_LD + j0 + 4 * s + q];
        double b = S[(8 * Cc + g) * LEAF_LD + j0 + 4 * s + q];
        dmma(c0, c1, a, b);
      }
      cv.x = -c0;
      cv.y = -c1;
      *reinterpret_cast<double2*>(cp) = cv;
    }
    __syncthreads();
  }
  if (warp == 0) {
    double tot = warp_sum(logsum);
    if (lane == 0) logdet_part[jb] = tot;
  }

  // ---- 2. inverse: 8x8 diagonal blocks (thread = one column of one block) ----
  if (tid < 128) {
    const int blk = tid >> 3, c = tid & 7, o = 8 * blk;
    double w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k < i && k >= c) s -= S[(o + i) * LEAF_LD + o + k] * w[k];
      w[i] = (i >= c) ? s * dvec[o + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (i > c) S[(o + c) * LEAF_LD + o + i] = w[i];  // W[o+i][o+c] stored transposed
  }
  __syncthreads();

  // ---- recursive doubling ----
  for (int b = 8; b <= 64; b <<= 1) {
    const int tb = b >> 3;              // 8x8 tiles per block edge
    const int tpp = tb * tb;            // tiles per pair
    const int ntile = (64 / b) * tpp;   // <= 64 -> <= 8 per warp
    // GEMM1: T[i][j] = sum_{k >= j} L21[i][k] W11[k][j]; stored transposed in the mirrored block
    for (int t = warp; t < ntile; t += 8) {
      const int p = t / tpp, rem = t - p * tpp;
      const int i0 = 8 * (rem / tb), jj0 = 8 * (rem % tb);
      const int r0 = 2 * b * p;
      double c0 = 0.0, c1 = 0.0;
      for (int k0 = jj0; k0 < b; k0 += 4) {
        double a = S[(r0 + b + i0 + g) * LEAF_LD + r0 + k0 + q];
        double bb = leaf_getW(S, dvec, r0 + k0 + q, r0 + jj0 + g);
        dmma(c0, c1, a, bb);
      }
      S[(r0 + jj0 + 2 * q) * LEAF_LD + r0 + b + i0 + g] = c0;
      S[(r0 + jj0 + 2 * q + 1) * LEAF_LD + r0 + b + i0 + g] = c1;
    }
    __syncthreads();
    // GEMM2: W21[i][j] = -sum_{k <= i} W22[i][k] T[k][j]; overwrites T in place after a barrier
    double r0c[8], r1c[8];
#pragma unroll
    for (int n = 0; n < 8; n++) {
      const int t = warp + 8 * n;
      r0c[n] = r1c[n] = 0.0;
      if (t < ntile) {
        const int p = t / tpp, rem = t - p * tpp;
        const int i0 = 8 * (rem / tb), jj0 = 8 * (rem % tb);
        const int r0 = 2 * b * p;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < i0 + 8; k0 += 4) {
          double a = leaf_getW(S, dvec, r0 + b + i0 + g, r0 + b + k0 + q);
          double bb = S[(r0 + jj0 + g) * LEAF_LD + r0 + b + k0 + q];
          dmma(c0, c1, a, bb);
        }
        r0c[n] = -c0;
        r1c[n] = -c1;
      }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < 8; n++) {
      const int t = warp + 8 * n;
      if (t < ntile) {
        const int p = t / tpp, rem = t - p * tpp;
        const int i0 = 8 * (rem / tb), jj0 = 8 * (rem % tb);
        const int r0 = 2 * b * p;
        S[(r0 + jj0 + 2 * q) * LEAF_LD + r0 + b + i0 + g] = r0c[n];
        S[(r0 + jj0 + 2 * q + 1) * LEAF_LD + r0 + b + i0 + g] = r1c[n];
      }
    }
    __syncthreads();
  }

  // ---- 3. write back ----
  for (int e = tid; e < LEAF_N * (LEAF_N / 2); e += LEAF_THREADS) {
    int r = e >> 6, c2 = (e & 63) * 2;
    double2 v;
    v.x = c2 <= r ? S[r * LEAF_LD + c2] : 0.0;
    v.y = c2 + 1 <= r ? S[r * LEAF_LD + c2 + 1] : 0.0;
    *reinterpret_cast<double2*>(A + (long)r * lda + c2) = v;
    double2 w;
    w.x = leaf_getW(S, dvec, r, c2);
    w.y = leaf_getW(S, dvec, r, c2 + 1);
    *reinterpret_cast<double2*>(W + (long)r * ldw + c2) = w;
  }
}

}  // namespace gpras
