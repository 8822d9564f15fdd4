// Small kernels of the sparse (inducing-point) model: the collapsed bound of GPflow's SGPR as called from
// gpras/gpr.py:293-308, with its gradient in closed form (derivation and CPU restatement: oracle/sgpr_analytic.py).
//
// Notation (M inducing points, N rows, R = output columns sharing the model, 1 in the reference):
//   L = chol(Kuu + jitter I), W = L^-1, A' = W Kuf, AATs = A' A'^T / s2, B = I + AATs, LB = chol(B), WB = LB^-1,
//   ae = A' Y / s2, c = WB ae, chat = WB^T c, u = W^T chat,
//   Rm = R (I - B^-1) - chat chat^T,  dF/dKuu = 1/2 W^T (Rm - R AATs) W,  dF/dKuf = (W^T Rm A' + u Y^T) / s2.
// All matrices are padded to multiples of 128 with zeros (identity on the diagonals of Kuu, B), so the dense
// steps run on the shared DMMA engine without bounds checks.
#pragma once
#include "gp_kernels.cuh"

namespace gpras {

// AATs = sum_z part[z] / s2 ;  B = AATs + I   (m_pad x m_pad, full)
static __global__ void sgpr_finish_b_kernel(const double* __restrict__ part, long slab, int nz, const double* __restrict__ theta,
                                     int m_pad, double* __restrict__ AATs, double* __restrict__ B, long bs = 0) {
  part += blockIdx.y * bs, theta += blockIdx.y * bs, AATs += blockIdx.y * bs, B += blockIdx.y * bs;
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)m_pad * m_pad) return;
  double s = 0.0;
  for (int z = 0; z < nz; z++) s += part[(long)z * slab + e];
  s /= theta[1];
  AATs[e] = s;
  const int i = (int)(e / m_pad), j = (int)(e - (long)i * m_pad);
  B[e] = s + (i == j ? 1.0 : 0.0);
}

// mirror the lower triangle into the upper one (n x n, pitch ld)
static __global__ void mirror_lower_kernel(double* __restrict__ A, int n, long ld, long bs = 0) {
  A += blockIdx.y * bs;
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)n * n) return;
  const int i = (int)(e / n), j = (int)(e - (long)i * n);
  if (j > i) A[(long)i * ld + j] = A[(long)j * ld + i];
}

// v[i][q] *= 1 / s2
static __global__ void sgpr_scale_noise_kernel(double* __restrict__ v, long count, const double* __restrict__ theta,
                                        long bs = 0) {
  v += blockIdx.y * bs, theta += blockIdx.y * bs;
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < count) v[e] /= theta[1];
}

// Rm = R (I - Binv) - chat chat^T ;  RA = Rm - R AATs     (full, m_pad x m_pad; chat is m_pad x ldc, R columns)
static __global__ void sgpr_build_r_kernel(const double* __restrict__ Binv, const double* __restrict__ AATs,
                                    const double* __restrict__ chat, long ldc, int R, int m_pad, double* __restrict__ Rm,
                                    double* __restrict__ RA, long bs = 0) {
  Binv += blockIdx.y * bs, AATs += blockIdx.y * bs, chat += blockIdx.y * bs, Rm += blockIdx.y * bs, RA += blockIdx.y * bs;
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)m_pad * m_pad) return;
  const int i = (int)(e / m_pad), j = (int)(e - (long)i * m_pad);
  double cc = 0.0;
  for (int q = 0; q < R; q++) cc = fma(chat[(long)i * ldc + q], chat[(long)j * ldc + q], cc);
  const double r = (double)R * ((i == j ? 1.0 : 0.0) - Binv[e]) - cc;
  Rm[e] = r;
  RA[e] = r - (double)R * AATs[e];
}

// scal[0] = tr(AATs)  [1] = tr(Binv) over the real m rows  [2] = |c|^2  [3] = |Y|^2  [4] = sum_q chat_q^T AATs chat_q
// One CTA, fixed-shape reductions.
static __global__ void sgpr_scalars_kernel(const double* __restrict__ AATs, const double* __restrict__ Binv,
                                    const double* __restrict__ c, const double* __restrict__ chat, long ldc, int R,
                                    const double* __restrict__ Y, long ldy, int n, int m, int m_pad,
                                    double* __restrict__ scal, long bs = 0) {
  AATs += blockIdx.y * bs, Binv += blockIdx.y * bs, c += blockIdx.y * bs, chat += blockIdx.y * bs, Y += blockIdx.y * bs,
      scal += blockIdx.y * bs;
  __shared__ double red[5][256];
  const int tid = threadIdx.x;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int i = tid; i < m; i += 256) {
    s0 += AATs[(long)i * m_pad + i];
    s1 += Binv[(long)i * m_pad + i];
    for (int q = 0; q < R; q++) s2 = fma(c[(long)i * ldc + q], c[(long)i * ldc + q], s2);
  }
  for (long e = tid; e < (long)n * R; e += 256) {
    const long r = e / R;
    const int q = (int)(e - r * R);
    const double v = Y[r * ldy + q];
    s3 = fma(v, v, s3);
  }
  for (long e = tid; e < (long)m * m; e += 256) {
    const int i = (int)(e / m), j = (int)(e - (long)i * m);
    double cc = 0.0;
    for (int q = 0; q < R; q++) cc = fma(chat[(long)i * ldc + q], chat[(long)j * ldc + q], cc);
    s4 = fma(cc, AATs[(long)i * m_pad + j], s4);
  }
  red[0][tid] = s0, red[1][tid] = s1, red[2][tid] = s2, red[3][tid] = s3, red[4][tid] = s4;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o)
#pragma unroll
      for (int k = 0; k < 5; k++) red[k][tid] += red[k][tid + o];
    __syncthreads();
  }
  if (tid < 5) scal[tid] = red[tid][0];
}

// ---------------------------------------------------------------------------------------------
// Fused chain-rule pass over one covariance block (128 x 128 tile per CTA, as grad_kernel):
//   UF mode: rows = inducing points (Zs), cols = training rows (Xs),  g = (G1[i][n] + sum_q u[i][q] Y[n][q]) / s2
//   UU mode: rows = cols = inducing points,                              g = G1[i][j]   (dF/dKuu, full symmetric)
// accumulating per CTA
//   part[cta][0]     = sum g k / variance
//   part[cta][1 + d] = sum g F s_d                         (F = dk/dlog l factor / variance)
//   zpart[tile_col][row][d] = sum_cols g F (zs_row,d - xs_col,d)   (x 2 in UU mode: k(z_i, z_j) depends on z_i twice)
// ---------------------------------------------------------------------------------------------
template <int KID, int DC>
__global__ void __launch_bounds__(PT_THREADS) sgpr_chain_kernel(const bool UU, const double* __restrict__ Zs, int m, const double* __restrict__ Xs,
                                                                int n, int D, const double* __restrict__ G1, long ldg,
                                                                const double* __restrict__ u, long ldu,
                                                                const double* __restrict__ Y, long ldy, int R,
                                                                const double* __restrict__ theta, int n_tiles_x,
                                                                double* __restrict__ part, int npart_cols,
                                                                double* __restrict__ zpart, int m_pad, long bs = 0) {
  extern __shared__ __align__(16) double smem[];
  Zs += blockIdx.y * bs, Xs += blockIdx.y * bs, G1 += blockIdx.y * bs, u += blockIdx.y * bs, Y += blockIdx.y * bs,
      theta += blockIdx.y * bs, part += blockIdx.y * bs, zpart += blockIdx.y * bs;
  double* s1 = smem;                    // [D][CT_LD]  rows (Zs)
  double* s2 = s1 + (long)D * CT_LD;    // [D][CT_LD]  cols (Xs or Zs)
  double* zacc = s2 + (long)D * CT_LD;  // [128][D]
  double* su = zacc + (long)CT * D;     // [R][CT_LD]  u rows      (UF only)
  double* sy = su + (long)R * CT_LD;    // [R][CT_LD]  Y rows      (UF only)
  __shared__ double red[PT_THREADS / 32][DC + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int ti = blockIdx.x / n_tiles_x, tj = blockIdx.x - ti * n_tiles_x;
  stage_features(s1, Zs, ti * CT, D, tid);
  stage_features(s2, Xs, tj * CT, D, tid);
  for (int e = tid; e < CT * D; e += PT_THREADS) zacc[e] = 0.0;
  if (!UU) {
    for (int e = tid; e < CT * R; e += PT_THREADS) {
      const int r = e / R, qq = e - r * R;
      su[qq * CT_LD + r] = u[(long)(ti * CT + r) * ldu + qq];
      sy[qq * CT_LD + r] = Y[(long)(tj * CT + r) * ldy + qq];
    }
  }
  __syncthreads();
  const double inv_s2 = UU ? 1.0 : 1.0 / theta[1];
  const double zw = UU ? 2.0 : 1.0;
  double g_var = 0.0;
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
      double gv[4];
      load4(G1 + (long)i * ldg + tj * CT + co, gv);
      if (!UU) {
        for (int qq = 0; qq < R; qq++) {
          const double ui = su[qq * CT_LD + ro + a];
          double yb[4];
          load4(sy + qq * CT_LD + co, yb);
#pragma unroll
          for (int b = 0; b < 4; b++) gv[b] = fma(ui, yb[b], gv[b]);
        }
      }
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int j = tj * CT + co + b;
        const double g = (i < m && j < n) ? gv[b] * inv_s2 : 0.0;
        double kval, fval;
        kernel_eval<KID>(r2[a][b], kval, fval);
        g_var = fma(g, kval, g_var);
        w[a][b] = g * fval;
      }
    }
#pragma unroll
    for (int dd = 0; dd < DC; dd++) {
      if (dd < D) {
        double xa[4], xb[4];
        load4(s1 + dd * CT_LD + ro, xa);
        load4(s2 + dd * CT_LD + co, xb);
        double sl = 0.0, zr[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
          zr[a] = 0.0;
#pragma unroll
          for (int b = 0; b < 4; b++) {
            const double df = xa[a] - xb[b];
            const double wd = w[a][b] * df;
            zr[a] += wd;
            sl = fma(wd, df, sl);
          }
        }
        gl[dd] += sl;
        // rows are shared by the 16 threads with equal ty (one half-warp): fixed-shape xor tree, then one owner adds
#pragma unroll
        for (int a = 0; a < 4; a++) {
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) zr[a] += __shfl_xor_sync(0xffffffffu, zr[a], o);
          if (tx == 0) zacc[(ro + a) * D + dd] += zw * zr[a];
        }
      }
    }
  }
  g_var = warp_sum(g_var);
  if (lane == 0) red[warp][0] = g_var;
#pragma unroll
  for (int dd = 0; dd < DC; dd++) {
    double v = warp_sum(gl[dd]);
    if (lane == 0) red[warp][1 + dd] = v;
  }
  __syncthreads();
  if (tid < 1 + D) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < PT_THREADS / 32; wv++) s += red[wv][tid];
    part[(long)blockIdx.x * npart_cols + tid] = s;
  }
  for (int e = tid; e < CT * D; e += PT_THREADS) {
    const int r = e / D, dd = e - r * D;
    zpart[((long)tj * m_pad + ti * CT + r) * D + dd] = zacc[e];
  }
}

// result = [elbo, dF/dlog variance, dF/dlog noise, dF/dlog l_0.., dF/dZ (m x D)]
// pa/pb: per-CTA partials of the UF / UU chain passes; za/zb: their Z partials over column tiles.
static __global__ void sgpr_finalize_kernel(const double* __restrict__ scal, const double* __restrict__ logdet_lb, int nlb,
                                     const double* __restrict__ pa, int na, const double* __restrict__ pb, int nb,
                                     int npart_cols, const double* __restrict__ za, int nza, const double* __restrict__ zb,
                                     int nzb, const double* __restrict__ theta, int n, int m, int m_pad, int D, int R,
                                     double* __restrict__ result, long bs = 0) {
  scal += blockIdx.y * bs, logdet_lb += blockIdx.y * bs, pa += blockIdx.y * bs, pb += blockIdx.y * bs, za += blockIdx.y * bs,
      zb += blockIdx.y * bs, theta += blockIdx.y * bs, result += blockIdx.y * bs;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const double variance = theta[0], s2 = theta[1];
  if (tid == 0) {
    double ldb = 0.0;
    for (int i = 0; i < nlb; i++) ldb += logdet_lb[i];
    const double tr_aat = scal[0], tr_binv = scal[1], cc = scal[2], yy = scal[3] / s2, caac = scal[4];
    result[0] = -0.5 * (double)n * R * 1.8378770664093453 -
                (double)R * (ldb + 0.5 * (double)n * log(s2) + 0.5 * ((double)n * variance / s2 - tr_aat)) - 0.5 * (yy - cc);
    result[2] = -0.5 * (double)n * R + 0.5 * R * ((double)m - tr_binv) + 0.5 * yy - cc + 0.5 * caac +
                0.5 * R * (double)n * variance / s2 - 0.5 * R * tr_aat;
  }
  if (tid < 1 + D) {
    double s = 0.0;
    for (int i = 0; i < na; i++) s += pa[(long)i * npart_cols + tid];
    for (int i = 0; i < nb; i++) s += pb[(long)i * npart_cols + tid];
    s *= variance;
    if (tid == 0)
      result[1] = s - 0.5 * R * (double)n * variance / s2;
    else
      result[2 + tid] = s;
  }
  for (int e = tid; e < m * D; e += gridDim.x * blockDim.x) {
    const int i = e / D, dd = e - i * D;
    double s = 0.0;
    for (int t = 0; t < nza; t++) s += za[((long)t * m_pad + i) * D + dd];
    for (int t = 0; t < nzb; t++) s += zb[((long)t * m_pad + i) * D + dd];
    result[3 + D + e] = -variance * s / theta[2 + dd];
  }
}

// colsq[t] = sum_i V[i][t]^2  over rows < rows (pitch ld): one thread per column, fixed order
static __global__ void colsumsq_kernel(const double* __restrict__ V, int rows, int cols, long ld, double* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cols) return;
  double s = 0.0;
  for (int i = 0; i < rows; i++) {
    const double v = V[(long)i * ld + t];
    s = fma(v, v, s);
  }
  out[t] = s;
}

// var[t] = variance + noise + sum_r p2[r][t] - q1[t]      (GPflow SGPR predict_f + likelihood noise)
static __global__ void sgpr_predict_var_kernel(const double* __restrict__ p2, int rows, long ld, const double* __restrict__ q1,
                                        int T, const double* __restrict__ theta, double* __restrict__ var) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double s = 0.0;
  for (int r = 0; r < rows; r++) s += p2[(long)r * ld + t];
  var[t] = theta[0] + theta[1] + s - q1[t];
}

}  // namespace gpras
