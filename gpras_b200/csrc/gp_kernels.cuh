// Pointwise / reduction kernels of the exact-GP path (sm_100a): feature scaling, the fused
// covariance builder (north_star (a)), the fused trace/gradient pass (north_star (d)), and the
// small deterministic reductions that assemble LML and its gradient.
//
// Layout conventions (all FP64, row-major):
//   theta_dev = [variance, noise, l_0 .. l_{D-1}]   (constrained values, device memory so that a
//               captured CUDA graph can be replayed with new hyperparameters)
//   Xs        = X / l            (N_pad x D; rows >= n are zero)
//   K, W, Kinv = N_pad x N_pad with pitch ld, lower triangle meaningful; padding rows/cols carry the
//               identity so that chol / inverse of the padded matrix embed those of the real one.
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int CT = 128;           // pointwise CTA tile edge (cov / grad kernels) == engine tile
constexpr int CT_LD = CT + 2;     // shared pitch of the transposed feature tiles (16-byte aligned rows)
constexpr int PT_THREADS = 256;   // 16 x 16 threads; each owns 4 x 4 entries in each 64 x 64 quadrant

// Xs[i][d] = X[i][d] / l_d for i < n, 0 for padding rows.
// (bs: element stride between the models of a batch, blockIdx.y = model; 0 for a single model.  Same convention in every
// kernel of the sparse model's evaluation: all per-model buffers live at one fixed offset from each other.)
static __global__ void scale_features_kernel(const double* __restrict__ X, double* __restrict__ Xs, int n, int n_pad, int D,
                                      const double* __restrict__ theta, long bs = 0) {
  X += blockIdx.y * bs, Xs += blockIdx.y * bs, theta += blockIdx.y * bs;
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)n_pad * D) return;
  int i = (int)(e / D), dd = (int)(e - (long)i * D);
  Xs[e] = i < n ? X[e] / theta[2 + dd] : 0.0;
}

// Stage rows [r0, r0+128) of Xs (pitch D) into shared memory transposed: s[d][r].
__device__ __forceinline__ void stage_features(double* __restrict__ s, const double* __restrict__ Xs, int r0, int D,
                                               int tid) {
  for (int e = tid; e < CT * D; e += PT_THREADS) {
    int r = e / D, dd = e - r * D;
    s[dd * CT_LD + r] = Xs[(long)(r0 + r) * D + dd];
  }
}

__device__ __forceinline__ void tri_tile(int t, int& ti, int& tj) {
  ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= t) ti++;
  while ((long)ti * (ti + 1) / 2 > t) ti--;
  tj = t - (int)((long)ti * (ti + 1) / 2);
}

__device__ __forceinline__ void load4(const double* __restrict__ s, double (&x)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(s);
  const double2 b = *reinterpret_cast<const double2*>(s + 2);
  x[0] = a.x, x[1] = a.y, x[2] = b.x, x[3] = b.y;
}

// ---------------------------------------------------------------------------------------------
// (a) fused covariance builder, 128 x 128 tile per CTA.
//   out[i][j] = variance * k(|Xs1_i - Xs2_j|^2) (+ noise on the diagonal when square),  i < n1, j < n2
//   padding:   square mode -> identity;  rectangular mode -> 0
// tri = 1 launches only the tiles with tj <= ti (grid.x = nt(nt+1)/2); used for K(X, X).
// ---------------------------------------------------------------------------------------------
template <int KID>
__global__ void __launch_bounds__(PT_THREADS) cov_kernel(const double* __restrict__ Xs1, int n1, const double* __restrict__ Xs2,
                                                         int n2, int D, const double* __restrict__ theta,
                                                         double* __restrict__ out, long ldo, int n_tiles_x, int tri,
                                                         int square, double jitter, long bs = 0) {
  extern __shared__ __align__(16) double smem[];
  Xs1 += blockIdx.y * bs, Xs2 += blockIdx.y * bs, theta += blockIdx.y * bs, out += blockIdx.y * bs;
  double* s1 = smem;
  double* s2 = smem + (long)D * CT_LD;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int ti, tj;
  if (tri) {
    tri_tile(blockIdx.x, ti, tj);
  } else {
    ti = blockIdx.x / n_tiles_x;
    tj = blockIdx.x - ti * n_tiles_x;
  }
  stage_features(s1, Xs1, ti * CT, D, tid);
  stage_features(s2, Xs2, tj * CT, D, tid);
  __syncthreads();
  // square mode adds the likelihood noise (jitter < 0) or a fixed jitter (sparse model's Kuu) on the diagonal
  const double variance = theta[0], noise = jitter < 0.0 ? theta[1] : jitter;
#pragma unroll 1
  for (int quad = 0; quad < 4; quad++) {
    const int ro = 64 * (quad >> 1) + 4 * ty, co = 64 * (quad & 1) + 4 * tx;
    double r2[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) r2[a][b] = 0.0;
#pragma unroll 4
    for (int dd = 0; dd < D; dd++) {
      double xa[4], xb[4];
      load4(s1 + dd * CT_LD + ro, xa);
      load4(s2 + dd * CT_LD + co, xb);
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          double df = xa[a] - xb[b];
          r2[a][b] = fma(df, df, r2[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int i = ti * CT + ro + a;
      double v[4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int j = tj * CT + co + b;
        double k = variance * kernel_value<KID>(r2[a][b]);
        if (square && i == j) k += noise;
        if (i >= n1 || j >= n2) k = (square && i == j) ? 1.0 : 0.0;
        v[b] = k;
      }
      double2* p = reinterpret_cast<double2*>(out + (long)i * ldo + tj * CT + co);
      p[0] = make_double2(v[0], v[1]);
      p[1] = make_double2(v[2], v[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (d) fused trace pass over  Wt = alpha alpha^T - P * Kinv  (lower 128-tiles, diagonal tiles full):
//   part[tile][0]      = sum wt * Wt_ij * k_ij / variance
//   part[tile][1]      = sum_i Wt_ii                         (diagonal tiles only)
//   part[tile][2 + d]  = sum wt * Wt_ij * F_ij * s_d,ij      (F = dk/dlog l factor / variance)
// wt = 2 off the block diagonal, 1 on it.  dK/dtheta is recomputed from the staged feature tiles and
// never materialised (8 N^2 n_theta bytes otherwise).  DC = lengthscale accumulators per thread (D <= DC).
// ---------------------------------------------------------------------------------------------
template <int KID, int DC>
__global__ void __launch_bounds__(PT_THREADS, DC <= 32 ? 2 : 1) grad_kernel(const double* __restrict__ Xs, int n, int D,
                                                          const double* __restrict__ Wt, long ldw,
                                                          double* __restrict__ part, int npart_cols) {
  extern __shared__ __align__(16) double smem[];
  double* s1 = smem;                           // [D][CT_LD]
  double* s2 = s1 + (long)D * CT_LD;           // [D][CT_LD]
  __shared__ double red[PT_THREADS / 32][DC + 2];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.x;
  int ti, tj;
  tri_tile(t, ti, tj);
  stage_features(s1, Xs, ti * CT, D, tid);
  stage_features(s2, Xs, tj * CT, D, tid);
  __syncthreads();

  const double wt = ti == tj ? 1.0 : 2.0;
  double g_var = 0.0, g_tr = 0.0;
  double gl[DC];
#pragma unroll
  for (int dd = 0; dd < DC; dd++) gl[dd] = 0.0;

#pragma unroll 1
  for (int quad = 0; quad < 4; quad++) {
    const int ro = 64 * (quad >> 1) + 4 * ty, co = 64 * (quad & 1) + 4 * tx;
    double w[4][4], r2[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) r2[a][b] = 0.0;
#pragma unroll 4
    for (int dd = 0; dd < D; dd++) {
      double xa[4], xb[4];
      load4(s1 + dd * CT_LD + ro, xa);
      load4(s2 + dd * CT_LD + co, xb);
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          double df = xa[a] - xb[b];
          r2[a][b] = fma(df, df, r2[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int i = ti * CT + ro + a;
      double kv[4];
      load4(Wt + (long)i * ldw + tj * CT + co, kv);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int j = tj * CT + co + b;
        const double wv = (i < n && j < n) ? wt * kv[b] : 0.0;
        if (i == j) g_tr += wv;
        double kval, fval;
        kernel_eval<KID>(r2[a][b], kval, fval);
        g_var = fma(wv, kval, g_var);
        w[a][b] = wv * fval;
      }
    }
#pragma unroll
    for (int dd = 0; dd < DC; dd++) {
      if (dd < D) {
        double xa[4], xb[4];
        load4(s1 + dd * CT_LD + ro, xa);
        load4(s2 + dd * CT_LD + co, xb);
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) {
            double df = xa[a] - xb[b];
            s = fma(w[a][b], df * df, s);
          }
        gl[dd] += s;
      }
    }
  }
  // deterministic CTA reduction: warp shuffle tree, then warps in fixed order
  g_var = warp_sum(g_var);
  g_tr = warp_sum(g_tr);
  if (lane == 0) {
    red[warp][0] = g_var;
    red[warp][1] = g_tr;
  }
#pragma unroll
  for (int dd = 0; dd < DC; dd++) {
    double v = warp_sum(gl[dd]);
    if (lane == 0) red[warp][2 + dd] = v;
  }
  __syncthreads();
  if (tid < 2 + D) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < PT_THREADS / 32; wv++) s += red[wv][tid];
    part[(long)t * npart_cols + tid] = s;
  }
}

// out[c] = sum_r part[r][c]: one 256-thread CTA per column; strided partial sums then a fixed-shape tree, so the
// result is bitwise repeatable.
static __global__ void colsum_kernel(const double* __restrict__ part, int rows, int cols, long ld, double* __restrict__ out) {
  __shared__ double red[256];
  const int c = blockIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int r = threadIdx.x; r < rows; r += 256) s += part[(long)r * ld + c];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = red[0];
}

// Per-CTA sum of squares of a (rows x cols) block with pitch ld -> part[blockIdx.x].
static __global__ void sumsq_partial_kernel(const double* __restrict__ U, int rows, int cols, long ld, double* __restrict__ part) {
  __shared__ double red[8];
  double s = 0.0;
  const long total = (long)rows * cols;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    long r = e / cols;
    int c = (int)(e - r * cols);
    double v = U[r * ld + c];
    s = fma(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += red[w];
    part[blockIdx.x] = tot;
  }
}

// result = [lml, dLML/dlog variance, dLML/dlog noise, dLML/dlog l_0 .. l_{D-1}]
static __global__ void finalize_kernel(const double* __restrict__ usq_part, int n_usq, const double* __restrict__ logdet_part,
                                int n_logdet, const double* __restrict__ gsum, const double* __restrict__ theta, int n,
                                int P, int D, int want_grad, double* __restrict__ result) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double usq = 0.0, ld = 0.0;
  for (int i = 0; i < n_usq; i++) usq += usq_part[i];
  for (int i = 0; i < n_logdet; i++) ld += logdet_part[i];
  result[0] = -0.5 * usq - (double)P * ld - 0.5 * (double)n * (double)P * 1.8378770664093453;
  if (want_grad) {
    const double variance = theta[0], noise = theta[1];
    result[1] = 0.5 * variance * gsum[0];
    result[2] = 0.5 * noise * gsum[1];
    for (int dd = 0; dd < D; dd++) result[3 + dd] = 0.5 * variance * gsum[2 + dd];
  }
}

// var[t] = variance + noise - sum_r part[r][t]   (predict_y semantics, gpras/gpr.py:337)
static __global__ void predict_var_kernel(const double* __restrict__ part, int rows, int T, long ld,
                                   const double* __restrict__ theta, double* __restrict__ var) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double s = 0.0;
  for (int r = 0; r < rows; r++) s += part[(long)r * ld + t];
  var[t] = theta[0] + theta[1] - s;
}

}  // namespace gpras
