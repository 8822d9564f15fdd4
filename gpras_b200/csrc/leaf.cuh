// Single-CTA factorisation of one 128 x 128 diagonal block: L = chol(A_jj) and W = L^-1.
//
// The block lives in shared memory as S[128][132] doubles (row pitch 132 == 4 mod 16 makes every
// DMMA fragment LDS.64 conflict free).  L occupies the lower triangle (incl. diagonal); W is built
// transposed in the strictly-upper triangle (W[i][k] at S[k][i], i > k) with its diagonal 1/L_ii in
// dvec[], so L, W and the temporaries of the inverse all fit in one 135 KB tile.
//
//   1. potrf, right-looking over 16 panels of 8 columns with one panel of look-ahead: warp 0 factors panel
//      J+1 in registers with shuffles (one rsqrt per column on the critical path) WHILE warps 1..7 apply
//      panel J's rank-8 update to the rest of the trailing triangle with DMMA.8x8x4 on 8 x 8 tiles; only
//      the update of block column J+1 itself sits between two panel factorisations.
//   2. inverse by recursive doubling: 8 x 8 diagonal blocks by forward substitution (one thread per
//      column), then for b = 8, 16, 32, 64 every pair of adjacent blocks fills its off-diagonal block
//      W21 = -W22 (L21 W11) with two DMMA products (T = L21 W11 is parked in the pair's mirrored
//      upper block, which W21^T then overwrites).  Each warp advances its b/8 output tiles together so the
//      dependent DMMA chains of different tiles interleave.
//   3. L (upper zeroed) goes back over A_jj; W (upper zeroed) goes to the W buffer; sum(log L_ii) to
//      logdet_part[jb]; a non-positive pivot records info = global column + 1 (LAPACK convention).
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int LEAF_N = 128, LEAF_LD = 132, LEAF_THREADS = 256;
constexpr int LEAF_SMEM_BYTES = (LEAF_N * LEAF_LD + LEAF_N) * (int)sizeof(double);

// 8x8 tiles of the 16 x 16 lower block triangle in column-major order: tiles of block column c start at
// LEAF_COL_OFF[c]; entry t is (row LEAF_TILE_I[t], column LEAF_TILE_C[t]).
struct LeafTileTable {
  unsigned char I[136], C[136];
  unsigned char off[17];
};
__host__ __device__ constexpr LeafTileTable make_leaf_table() {
  LeafTileTable t{};
  int n = 0;
  for (int c = 0; c < 16; c++) {
    t.off[c] = (unsigned char)n;
    for (int i = c; i < 16; i++) {
      t.I[n] = (unsigned char)i;
      t.C[n] = (unsigned char)c;
      n++;
    }
  }
  t.off[16] = (unsigned char)n;
  return t;
}
__constant__ LeafTileTable c_leaf_tiles = make_leaf_table();

__device__ __forceinline__ double leaf_getW(const double* S, const double* dvec, int i, int k) {
  const double s = S[k * LEAF_LD + i];
  const double dg = dvec[i];
  return i == k ? dg : (i > k ? s : 0.0);
}

// rank-8 update of one 8x8 tile (I, Cc) with panel columns [j0, j0+8)
__device__ __forceinline__ void leaf_update_tile(double* S, int I, int Cc, int j0, int g, int q) {
  double* cp = S + (8 * I + g) * LEAF_LD + 8 * Cc + 2 * q;
  double2 cv = *reinterpret_cast<double2*>(cp);
  double c0 = -cv.x, c1 = -cv.y;
  const double a0 = S[(8 * I + g) * LEAF_LD + j0 + q], a1 = S[(8 * I + g) * LEAF_LD + j0 + 4 + q];
  const double b0 = S[(8 * Cc + g) * LEAF_LD + j0 + q], b1 = S[(8 * Cc + g) * LEAF_LD + j0 + 4 + q];
  dmma(c0, c1, a0, b0);
  dmma(c0, c1, a1, b1);
  cv.x = -c0;
  cv.y = -c1;
  *reinterpret_cast<double2*>(cp) = cv;
}

// warp 0: factor the 8-column panel starting at column j0 (rows j0..127), in registers.
__device__ __forceinline__ void leaf_factor_panel(double* S, double* dvec, int j0, int lane, double& logsum, int* info,
                                                  int col_base) {
  const int nm = (LEAF_N - j0 + 31) >> 5;  // row groups of 32 that hold real rows (warp-uniform)
  double v[4][8];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int r = j0 + lane + 32 * m;
#pragma unroll
    for (int c = 0; c < 8; c++) v[m][c] = (m < nm && r < LEAF_N) ? S[r * LEAF_LD + j0 + c] : 0.0;
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    double dpiv = __shfl_sync(0xffffffffu, v[0][c], c);
    if (!(dpiv > 0.0)) {
      if (lane == 0) atomicCAS(info, 0, col_base + j0 + c + 1);
      dpiv = 1.0;
    }
    double rs = rsqrt(dpiv);
    rs = rs * (1.5 - 0.5 * dpiv * rs * rs);  // one Newton step: full double accuracy
    const double sq = dpiv * rs;
    if (lane == c) {
      logsum += log(sq);
      dvec[j0 + c] = rs;
    }
#pragma unroll
    for (int m = 0; m < 4; m++) v[m][c] = (m == 0 && lane == c) ? sq : v[m][c] * rs;
#pragma unroll
    for (int c2 = c + 1; c2 < 8; c2++) {
      const double l = __shfl_sync(0xffffffffu, v[0][c], c2);
#pragma unroll
      for (int m = 0; m < 4; m++) v[m][c2] -= v[m][c] * l;
    }
  }
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int r = j0 + lane + 32 * m;
    if (m < nm && r < LEAF_N) {
#pragma unroll
      for (int c = 0; c < 8; c++) S[r * LEAF_LD + j0 + c] = (r >= j0 + c) ? v[m][c] : 0.0;
    }
  }
}

// One recursive-doubling level of the inverse at block size B (8x8-tile units TB = B/8): every warp owns
// TPW = B/8 output tiles and advances them together.
template <int B>
__device__ __forceinline__ void leaf_inverse_level(double* S, const double* dvec, int warp, int g, int q) {
  constexpr int TB = B / 8, TPP = TB * TB, TPW = B / 8;
  int i0[TPW], jj0[TPW], r0[TPW];
#pragma unroll
  for (int n = 0; n < TPW; n++) {
    const int t = warp + 8 * n;
    const int p = t / TPP, rem = t - p * TPP;
    i0[n] = 8 * (rem / TB);
    jj0[n] = 8 * (rem % TB);
    r0[n] = 2 * B * p;
  }
  double c0[TPW], c1[TPW];
  // GEMM1: T[i][j] = sum_{k >= j} L21[i][k] W11[k][j]; stored transposed in the pair's mirrored upper block
#pragma unroll
  for (int n = 0; n < TPW; n++) c0[n] = c1[n] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < B; k0 += 4) {
#pragma unroll
    for (int n = 0; n < TPW; n++) {
      if (k0 >= jj0[n]) {
        const double a = S[(r0[n] + B + i0[n] + g) * LEAF_LD + r0[n] + k0 + q];
        const double bb = leaf_getW(S, dvec, r0[n] + k0 + q, r0[n] + jj0[n] + g);
        dmma(c0[n], c1[n], a, bb);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < TPW; n++) {
    S[(r0[n] + jj0[n] + 2 * q) * LEAF_LD + r0[n] + B + i0[n] + g] = c0[n];
    S[(r0[n] + jj0[n] + 2 * q + 1) * LEAF_LD + r0[n] + B + i0[n] + g] = c1[n];
  }
  __syncthreads();
  // GEMM2: W21[i][j] = -sum_{k <= i} W22[i][k] T[k][j]; overwrites T in place after a barrier
#pragma unroll
  for (int n = 0; n < TPW; n++) c0[n] = c1[n] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < B; k0 += 4) {
#pragma unroll
    for (int n = 0; n < TPW; n++) {
      if (k0 < i0[n] + 8) {
        const double a = leaf_getW(S, dvec, r0[n] + B + i0[n] + g, r0[n] + B + k0 + q);
        const double bb = S[(r0[n] + jj0[n] + g) * LEAF_LD + r0[n] + B + k0 + q];
        dmma(c0[n], c1[n], a, bb);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < TPW; n++) {
    S[(r0[n] + jj0[n] + 2 * q) * LEAF_LD + r0[n] + B + i0[n] + g] = -c0[n];
    S[(r0[n] + jj0[n] + 2 * q + 1) * LEAF_LD + r0[n] + B + i0[n] + g] = -c1[n];
  }
  __syncthreads();
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
    const int r = e >> 6, c2 = (e & 63) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (c2 <= r) v = *reinterpret_cast<const double2*>(A + (long)r * lda + c2);
    S[r * LEAF_LD + c2] = v.x;
    S[r * LEAF_LD + c2 + 1] = c2 + 1 <= r ? v.y : 0.0;
  }
  __syncthreads();

  // ---- 1. blocked Cholesky with one panel of look-ahead ----
  double logsum = 0.0;  // warp 0, lanes 0..7: partial sums of log L_ii
  if (warp == 0) leaf_factor_panel(S, dvec, 0, lane, logsum, info, jb * LEAF_N);
  __syncthreads();
  for (int J = 0; J < 15; J++) {
    const int j0 = 8 * J;
    // phase A: block column J+1 (tiles (I, J+1), I = J+1..15) gets panel J's update
    for (int I = J + 1 + warp; I < 16; I += 8) leaf_update_tile(S, I, J + 1, j0, g, q);
    __syncthreads();
    // phase B: warp 0 factors panel J+1; warps 1..7 update the rest of the trailing triangle with panel J
    if (warp == 0) {
      leaf_factor_panel(S, dvec, j0 + 8, lane, logsum, info, jb * LEAF_N);
    } else {
      const int t_end = c_leaf_tiles.off[16];
#pragma unroll 2
      for (int t = c_leaf_tiles.off[J + 2 > 16 ? 16 : J + 2] + (warp - 1); t < t_end; t += 7)
        leaf_update_tile(S, c_leaf_tiles.I[t], c_leaf_tiles.C[t], j0, g, q);
    }
    __syncthreads();
  }
  if (warp == 0) {
    const double tot = warp_sum(logsum);
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
  leaf_inverse_level<8>(S, dvec, warp, g, q);
  leaf_inverse_level<16>(S, dvec, warp, g, q);
  leaf_inverse_level<32>(S, dvec, warp, g, q);
  leaf_inverse_level<64>(S, dvec, warp, g, q);

  // ---- 3. write back ----
  for (int e = tid; e < LEAF_N * (LEAF_N / 2); e += LEAF_THREADS) {
    const int r = e >> 6, c2 = (e & 63) * 2;
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
