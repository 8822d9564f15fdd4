// Fused evaluation of the sparse model for M <= 64 inducing points and D <= 32 features: FOUR kernels per evaluation
// instead of ~45 ("one persistent launch sequence per SGPR evaluation").
//
// At reference scale (gpras/gpr.py:293-308 with M = 50, N = 5 000, D = 10) the general path (sgpr_abi.cuh) pads every M x M
// matrix to one 128-tile and runs each dense step as its own launch on the tile engine: an evaluation is a chain of ~45
// dependent microsecond kernels, most of them one CTA, and takes ~0.5 ms however many models are batched.  Here every
// M x M step of one model runs inside ONE CTA with the matrices in shared memory (padded to a multiple of 8 only), and the
// N-wide steps are two passes over 128-row tiles of the training inputs:
//   sf_prep      (1 CTA / model)   Zs = Z / l, Kuu + jitter I, L = chol(Kuu), W = L^-1
//   sf_forward   (tiles x models)  Kuf tile from the features, A' = W Kuf (stored), partial A' A'^T and A' y per CTA
//   sf_mid       (1 CTA / model)   B = I + A'A'^T / s2, LB, WB, c, chat, u, B^-1, the bound's scalars, Rm, RA,
//                                  RW = W^T Rm, Guu = 1/2 W^T RA W and the chain rule through Kuu
//   sf_backward  (tiles x models)  G = (RW A' + u y^T) / s2 per tile and the chain rule through Kuf to theta and Z
// followed by sgpr_finalize_kernel (shared with the general path).  Same formulas and notation as sgpr_kernels.cuh; the
// arithmetic is plain FP64 FMA on register tiles (on this part DFMA and DMMA share one datapath, DESIGN.md section 4, and
// the matrices are far too small for the tile engine's 128-wide shapes).
#pragma once
#include "gp_kernels.cuh"
#include "leaf.cuh"

namespace gpras {

constexpr int SF_MP = 64;            // largest padded number of inducing points
constexpr int SF_LD = SF_MP + 1;     // shared-memory pitch of the M x M matrices (odd: rows and columns conflict free)
constexpr int SF_THREADS = 256;
constexpr int SF_TN = 128;           // training rows per tile
constexpr int SF_MAX_D = 32;
constexpr int SF_MAT = SF_MP * SF_LD;

struct SfArgs {
  const double* X;    // [n_pad][D]   shared by the models
  const double* yv;   // [P][n_pad]   targets, zero padded
  const double* yy;   // [P]          |y|^2
  // per model; consecutive models are bs doubles apart
  const double* theta;  // [2 + D]
  const double* Z;      // [m][D]
  double* Zs;           // [SF_MP][D]
  double* W;            // [SF_MP][SF_MP]   L^-1, zeros above the diagonal, identity on the padding
  double* Ap;           // [mp][n_pad]
  double* Kv;           // [mp][n_pad]   k(z_i, x_n) / variance      (forward pass -> backward pass)
  double* Fv;           // [mp][n_pad]   derivative factor F / variance
  double* slabs;        // [ntn][SF_MP * SF_MP]   per tile: partial A' A'^T (compact mp x mp)
  double* aep;          // [ntn][SF_MP]           per tile: partial A' y
  double* aats;         // [SF_MP * SF_MP]        their sums over the tiles (sf_reduce_kernel)
  double* aes;          // [SF_MP]
  double* RW;           // [SF_MP][SF_MP]
  double* WBg;          // [SF_MP][SF_MP]   LB^-1, kept for prediction
  double* uvec;         // [SF_MP]
  double* scal;         // [8]
  double* logdetB;      // [1]
  double* partA;        // [ntn][1 + D]
  double* zpA;          // [ntn][SF_MP][D]
  double* partB;        // [1 + D]
  double* zpB;          // [SF_MP][D]
  int* info;
  long bs;
  int n, n_pad, D, m, mp, ntn;
  double jitter;
};

// ---- in-CTA dense helpers on mp x mp shared-memory matrices (pitch SF_LD), 256 threads --------------------------------

// Cholesky factor AND its inverse in one sweep: A = L L^T (lower triangle of A consumed), Wm = L^-1 (lower triangular, zeros
// above the diagonal), rs[j] = 1 / L_jj.  Right-looking with 8-column panels; Wm starts as the identity and is carried along as
// a right-hand side (block forward substitution L X = I), so the inverse costs no extra pass and no extra barrier:
//   per panel J   every thread factors the 8 x 8 diagonal block redundantly in registers (no broadcast, no barrier);
//                 one thread per row below solves the panel  L_iJ = A_iJ L_JJ^-T  (kept in the panel buffer Lp only: later
//                 panels never read earlier columns of L), one thread per column solves  X_J = L_JJ^-1 R_J;      -- barrier --
//                 rank-8 updates of the trailing triangle of A and of the remaining right-hand-side rows
//                 R_i -= L_iJ X_J  on a 16 x 16 thread grid.                                                        -- barrier --
// A non-positive pivot records info = column + 1 (LAPACK convention) and continues with a unit pivot.  Ends with a barrier.
constexpr int SF_LP = 9;  // pitch of the panel buffer [SF_MP][8]
__device__ __forceinline__ void sf_chol_inv(double* __restrict__ A, double* __restrict__ Lp, double* __restrict__ Wm,
                                            double* __restrict__ rs, int mp, int* __restrict__ info) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    Wm[i * SF_LD + j] = i == j ? 1.0 : 0.0;
  }
  for (int j0 = 0; j0 < mp; j0 += 8) {
    __syncthreads();  // the updates of the previous panel are complete (first pass: A and the identity are in place)
    double Dg[8][8], rsv[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c <= r; c++) Dg[r][c] = A[(j0 + r) * SF_LD + j0 + c];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      double dp = Dg[c][c];
      if (!(dp > 0.0)) {  // identical on every thread
        if (tid == 0) atomicCAS(info, 0, j0 + c + 1);
        dp = 1.0;
      }
      const double r_ = leaf_rsqrt(dp);
      rsv[c] = r_;
      Dg[c][c] = dp * r_;
#pragma unroll
      for (int r = c + 1; r < 8; r++) Dg[r][c] *= r_;
#pragma unroll
      for (int c2 = c + 1; c2 < 8; c2++)
#pragma unroll
        for (int r = c2; r < 8; r++) Dg[r][c2] -= Dg[r][c] * Dg[c2][c];
    }
    const int i = j0 + tid, cx = tid - 128;
    if (tid < 8) {
#pragma unroll
      for (int r = 0; r < 8; r++)
        if (tid == r) rs[i] = rsv[r];
    } else if (i < mp) {  // L_iJ = a L_JJ^-T
      double v[8];
#pragma unroll
      for (int c = 0; c < 8; c++) v[c] = A[i * SF_LD + j0 + c];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double sacc = v[c];
#pragma unroll
        for (int c1 = 0; c1 < c; c1++) sacc = fma(-v[c1], Dg[c][c1], sacc);
        v[c] = sacc * rsv[c];
      }
#pragma unroll
      for (int c = 0; c < 8; c++) Lp[i * SF_LP + c] = v[c];
    } else if (cx >= 0 && cx < j0 + 8) {  // column cx of X_J = L_JJ^-1 R_J
      double x[8];
#pragma unroll
      for (int r = 0; r < 8; r++) x[r] = Wm[(j0 + r) * SF_LD + cx];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        double sacc = x[r];
#pragma unroll
        for (int r1 = 0; r1 < r; r1++) sacc = fma(-Dg[r][r1], x[r1], sacc);
        x[r] = sacc * rsv[r];
      }
#pragma unroll
      for (int r = 0; r < 8; r++) Wm[(j0 + r) * SF_LD + cx] = x[r];
    }
    __syncthreads();
    for (int ii = j0 + 8 + ty; ii < mp; ii += 16) {
      double li[8];
#pragma unroll
      for (int c = 0; c < 8; c++) li[c] = Lp[ii * SF_LP + c];
      for (int k = j0 + 8 + tx; k <= ii; k += 16) {
        double sacc = A[ii * SF_LD + k];
#pragma unroll
        for (int c = 0; c < 8; c++) sacc = fma(-li[c], Lp[k * SF_LP + c], sacc);
        A[ii * SF_LD + k] = sacc;
      }
      for (int cc = tx; cc < j0 + 8; cc += 16) {
        double sacc = Wm[ii * SF_LD + cc];
#pragma unroll
        for (int r = 0; r < 8; r++) sacc = fma(-li[r], Wm[(j0 + r) * SF_LD + cc], sacc);
        Wm[ii * SF_LD + cc] = sacc;
      }
    }
  }
  __syncthreads();
}

// C = alpha * op(A) op(B), all mp x mp; thread (ty, tx) owns rows ty + 16 r, columns tx + 16 s.  C aliases neither operand.
// No barrier inside.
template <bool TA, bool TB>
__device__ __forceinline__ void sf_mm(double* __restrict__ C, const double* __restrict__ A, const double* __restrict__ B, int mp,
                                      double alpha) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int s = 0; s < 4; s++) acc[r][s] = 0.0;
  for (int k = 0; k < mp; k++) {
    double a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; r++) a[r] = TA ? A[k * SF_LD + ty + 16 * r] : A[(ty + 16 * r) * SF_LD + k];
#pragma unroll
    for (int s = 0; s < 4; s++) b[s] = TB ? B[(tx + 16 * s) * SF_LD + k] : B[k * SF_LD + tx + 16 * s];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int s = 0; s < 4; s++) acc[r][s] = fma(a[r], b[s], acc[r][s]);
  }
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int i = ty + 16 * r, j = tx + 16 * s;
      if (i < mp && j < mp) C[i * SF_LD + j] = alpha * acc[r][s];
    }
}

// fixed-shape block sum of NV values per thread (red: [NV][SF_THREADS] shared); result valid on every thread after return
template <int NV>
__device__ __forceinline__ void sf_block_sum(double (&v)[NV], double* __restrict__ red) {
  const int tid = threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) red[k * SF_THREADS + tid] = v[k];
  __syncthreads();
  for (int o = SF_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o)
#pragma unroll
      for (int k = 0; k < NV; k++) red[k * SF_THREADS + tid] += red[k * SF_THREADS + tid + o];
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = red[k * SF_THREADS];
  __syncthreads();
}

// ---- 1. per model: scaled inducing inputs, Kuu, its Cholesky factor and the factor's inverse ---------------------------
constexpr int sf_prep_smem(int D) { return (3 * SF_MAT + SF_MP * D + SF_MP + SF_MAX_D) * (int)sizeof(double); }

template <int KID>
__device__ __forceinline__ void sf_prep_body(const SfArgs& a) {
  extern __shared__ __align__(16) double smem[];
  double* SA = smem;             // Kuu
  double* SL = SA + SF_MAT;      // panel buffer of the factorisation
  double* SWm = SL + SF_MAT;     // W
  double* zs = SWm + SF_MAT;     // [mp][D]
  double* rs = zs + SF_MP * a.D;
  double* ls = rs + SF_MP;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long off = (long)blockIdx.y * a.bs;
  const int D = a.D, m = a.m, mp = a.mp;
  const double* theta = a.theta + off;
  const double* Z = a.Z + off;
  int* info = a.info + off * 2;
  if (tid == 0) *info = 0;
  if (tid < D) ls[tid] = theta[2 + tid];
  __syncthreads();
  for (int e = tid; e < mp * D; e += SF_THREADS) {
    const int i = e / D, dd = e - i * D;
    const double v = i < m ? Z[e] / ls[dd] : 0.0;
    zs[e] = v;
    (a.Zs + off)[e] = v;
  }
  __syncthreads();
  const double variance = theta[0];
  for (int i = ty; i < mp; i += 16)
    for (int j = tx; j <= i; j += 16) {
      double r2 = 0.0;
      for (int dd = 0; dd < D; dd++) {
        const double df = zs[i * D + dd] - zs[j * D + dd];
        r2 = fma(df, df, r2);
      }
      double k = variance * kernel_value<KID>(r2);
      if (i == j) k += a.jitter;
      if (i >= m || j >= m) k = i == j ? 1.0 : 0.0;
      SA[i * SF_LD + j] = k;
    }
  sf_chol_inv(SA, SL, SWm, rs, mp, info);  // (SL serves as the panel buffer)
  double* Wg = a.W + off;
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    Wg[i * SF_MP + j] = SWm[i * SF_LD + j];
  }
}

template <int KID>
__global__ void __launch_bounds__(SF_THREADS, 1) sf_prep_kernel(const SfArgs a) {
  sf_prep_body<KID>(a);
}

// ---- 2. per (tile of 128 training rows, model): A' = W Kuf, partial A' A'^T and A' y -----------------------------------
// Tile kernels: 8 warps, one tile per CTA; the three products (A' = W Kuf, A' A'^T here, RW A' in the backward pass) run on
// the FP64 tensor pipe (DMMA.8x8x4, fragment layout of common.cuh).  Warp w owns rows [8w, 8w + 8) of the M-row tile and all
// 128 columns; its A fragments (8 rows of W or RW) sit in registers, the B operand (Kuf / A' tile) in shared memory with a
// pitch == 4 (mod 16) doubles, which makes the fragment loads conflict free.
constexpr int SF_LDK = SF_TN + 4;   // pitch of the Kuf / A' tile  [mp][128]

struct SfTileSmem {
  double *xsT;   // [D][128]        scaled features of the tile's training rows, transposed
  double *KA;    // [mp][SF_LDK]    Kuf tile, then A' tile
  double *zs;    // [mp][D]         scaled inducing inputs
  double *ysm;   // [128]           targets of the tile
  double *us;    // [SF_MP]         u (backward)
};
__host__ __device__ constexpr int sf_tile_doubles(int D, int mp) { return D * SF_TN + mp * SF_LDK + mp * D + SF_TN + SF_MP; }
__device__ __forceinline__ SfTileSmem sf_tile_layout(double* smem, int D, int mp) {
  SfTileSmem t;
  t.xsT = smem;
  t.KA = t.xsT + D * SF_TN;   // (16-byte aligned: D * 128 doubles)
  t.zs = t.KA + mp * SF_LDK;
  t.ysm = t.zs + mp * D;
  t.us = t.ysm + SF_TN;
  return t;
}

// features of rows [n0, n0 + 128) scaled by the lengthscales, transposed: xsT[d][c]; targets of the tile
__device__ __forceinline__ void sf_stage_rows(const SfArgs& a, int n0, const double* __restrict__ theta, double* __restrict__ xsT,
                                              double* __restrict__ ysm, int model) {
  const int tid = threadIdx.x, D = a.D;
  for (int e = tid; e < SF_TN * D; e += SF_THREADS) {
    const int c = e / D, dd = e - c * D;
    xsT[dd * SF_TN + c] = n0 + c < a.n ? a.X[(long)(n0 + c) * D + dd] / theta[2 + dd] : 0.0;
  }
  if (tid < SF_TN) ysm[tid] = a.yv[(long)model * a.n_pad + n0 + tid];
}

constexpr int SF_AAT_SLOTS = 5;  // lower 8 x 8 tiles of A' A'^T per warp: ceil(36 / 8)

template <int KID>
__global__ void __launch_bounds__(SF_THREADS, 2) sf_forward_kernel(const SfArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int D = a.D, m = a.m, mp = a.mp, na = mp >> 3, mt = mp >> 3;
  const SfTileSmem sm = sf_tile_layout(smem, D, mp);
  double *xsT = sm.xsT, *zs = sm.zs, *KA = sm.KA, *ysm = sm.ysm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int model = blockIdx.y, n0 = blockIdx.x * SF_TN;
  const long off = (long)model * a.bs;
  const double* theta = a.theta + off;
  const double variance = theta[0];
  const bool row_warp = warp < mt;
  const int irow = 8 * warp + g;  // this lane's row of the M-row tile in the DMMA accumulator layout
  sf_stage_rows(a, n0, theta, xsT, ysm, model);
  for (int e = tid; e < mp * D; e += SF_THREADS) zs[e] = (a.Zs + off)[e];
  // A fragments of W for this warp's rows: W is lower triangular, so rows [8w, 8w + 8) need k < 8w + 8 only
  double wfrag[2 * (SF_MP / 8)];
#pragma unroll
  for (int k4 = 0; k4 < 2 * (SF_MP / 8); k4++)
    wfrag[k4] = (row_warp && k4 < 2 * (warp + 1)) ? (a.W + off)[irow * SF_MP + 4 * k4 + q] : 0.0;
  __syncthreads();
  // Kuf entries of this thread: rows warp + 8 r, columns lane + 32 s; k / variance and the derivative factor are kept for
  // the backward pass
  double* Kvg = a.Kv + off;
  double* Fvg = a.Fv + off;
#pragma unroll 1
  for (int r = 0; r < na; r++) {
    const int i = warp + 8 * r;
    double r2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int dd = 0; dd < D; dd++) {
      const double zi = zs[i * D + dd];
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const double df = zi - xsT[dd * SF_TN + lane + 32 * s];
        r2[s] = fma(df, df, r2[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int c = lane + 32 * s;
      double kval, fval;
      kernel_eval<KID>(r2[s], kval, fval);
      Kvg[(long)i * a.n_pad + n0 + c] = kval;
      Fvg[(long)i * a.n_pad + n0 + c] = fval;
      KA[i * SF_LDK + c] = (i >= m || n0 + c >= a.n) ? 0.0 : variance * kval;
    }
  }
  __syncthreads();
  double acc[16][2];
#pragma unroll
  for (int ct = 0; ct < 16; ct++) acc[ct][0] = acc[ct][1] = 0.0;
  if (row_warp) {
#pragma unroll
    for (int k4 = 0; k4 < 2 * (SF_MP / 8); k4++) {
      if (k4 < 2 * (warp + 1)) {
        const double* bp = KA + (4 * k4 + q) * SF_LDK + g;
#pragma unroll
        for (int ct = 0; ct < 16; ct++) dmma(acc[ct][0], acc[ct][1], wfrag[k4], bp[8 * ct]);
      }
    }
    double* Apg = a.Ap + off + (long)irow * a.n_pad + n0;
    double sy = 0.0;
#pragma unroll
    for (int ct = 0; ct < 16; ct++) {
      const int c = 8 * ct + 2 * q;
      *reinterpret_cast<double2*>(Apg + c) = make_double2(acc[ct][0], acc[ct][1]);
      sy = fma(acc[ct][0], ysm[c], sy);
      sy = fma(acc[ct][1], ysm[c + 1], sy);
    }
    sy += __shfl_xor_sync(0xffffffffu, sy, 1);
    sy += __shfl_xor_sync(0xffffffffu, sy, 2);
    if (q == 0) (a.aep + off)[(long)blockIdx.x * SF_MP + irow] = sy;
  }
  __syncthreads();  // every read of the Kuf tile is done: the buffer becomes the A' tile
  if (row_warp) {
#pragma unroll
    for (int ct = 0; ct < 16; ct++)
      *reinterpret_cast<double2*>(KA + irow * SF_LDK + 8 * ct + 2 * q) = make_double2(acc[ct][0], acc[ct][1]);
  }
  __syncthreads();
  // partial A' A'^T of this tile: lower 8 x 8 tiles t = warp, warp + 8, ...; both triangles are written (compact mp x mp)
  double* slab = a.slabs + off + (long)blockIdx.x * SF_MP * SF_MP;
  for (int t = warp; t < mt * (mt + 1) / 2; t += 8) {
    int it = 0;
    while ((it + 1) * (it + 2) / 2 <= t) it++;
    const int jt = t - it * (it + 1) / 2;
    const double* ap = KA + (8 * it + g) * SF_LDK + q;
    const double* bp = KA + (8 * jt + g) * SF_LDK + q;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator pairs: even / odd k4 (shorter dependency chain)
#pragma unroll 4
    for (int k4 = 0; k4 < SF_TN / 4; k4 += 2) {
      dmma(c0, c1, ap[4 * k4], bp[4 * k4]);
      dmma(e0, e1, ap[4 * k4 + 4], bp[4 * k4 + 4]);
    }
    c0 += e0, c1 += e1;
    const int i = 8 * it + g, j = 8 * jt + 2 * q;
    slab[i * mp + j] = c0;
    slab[i * mp + j + 1] = c1;
    if (it != jt) {
      slab[j * mp + i] = c0;
      slab[(j + 1) * mp + i] = c1;
    }
  }
}

// sum of the tiles' partial matrices and vectors, in tile order: slabs[0] <- sum_t slabs[t], aep[0] <- sum_t aep[t]
__global__ void __launch_bounds__(SF_THREADS) sf_reduce_kernel(const SfArgs a) {
  const long off = (long)blockIdx.y * a.bs;
  const int e = blockIdx.x * SF_THREADS + threadIdx.x, mm = a.mp * a.mp;
  if (e < mm) {
    double* sp = a.slabs + off + e;
    double s = 0.0;
#pragma unroll 8
    for (int t = 0; t < a.ntn; t++) s += sp[(long)t * SF_MP * SF_MP];
    (a.aats + off)[e] = s;
  } else if (e < mm + a.mp) {
    double* sp = a.aep + off + (e - mm);
    double s = 0.0;
#pragma unroll 8
    for (int t = 0; t < a.ntn; t++) s += sp[(long)t * SF_MP];
    (a.aes + off)[e - mm] = s;
  }
}

// ---- 3. per model: everything M x M between the two passes -------------------------------------------------------------
constexpr int sf_mid_smem(int D) {
  return (4 * SF_MAT + SF_MP * D + 6 * SF_MP + 5 * SF_THREADS + 8 * (SF_MAX_D + 1)) * (int)sizeof(double);
}

template <int KID>
__global__ void __launch_bounds__(SF_THREADS, 1) sf_mid_kernel(const SfArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* SW = smem;            // W = L^-1
  double* SA = SW + SF_MAT;     // AATs -> RA -> Guu
  double* SB = SA + SF_MAT;     // B -> WB -> Rm -> T1
  double* SL = SB + SF_MAT;     // WB (swapped into SB after the factorisation) -> Binv -> RW
  double* zs = SL + SF_MAT;     // [mp][D]
  double* ae = zs + SF_MP * a.D;
  double* cv = ae + SF_MP;
  double* chat = cv + SF_MP;
  double* uv = chat + SF_MP;
  double* rs = uv + SF_MP;
  double* lgs = rs + SF_MP;                 // [SF_MP] scratch
  double* red = lgs + SF_MP;                // [5][256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.D, m = a.m, mp = a.mp;
  const long off = (long)blockIdx.y * a.bs;
  const double* theta = a.theta + off;
  const double s2 = theta[1];
  int* info = a.info + off * 2;

  for (int e = tid; e < mp * D; e += SF_THREADS) zs[e] = (a.Zs + off)[e];
  // AATs = (sum over the tiles of A' A'^T, from sf_reduce_kernel) / s2 ;  B = I + AATs ;  ae = A' y / s2
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    const double sv = (a.aats + off)[e] / s2;
    SW[i * SF_LD + j] = (a.W + off)[i * SF_MP + j];
    SA[i * SF_LD + j] = sv;
    SB[i * SF_LD + j] = sv + (i == j ? 1.0 : 0.0);
  }
  if (tid < mp) ae[tid] = (a.aes + off)[tid] / s2;
  // LB = chol(B) and WB = LB^-1 in one sweep (B in SB is consumed, WB lands in SL; `red` is the panel buffer), log det.
  // The two buffers then swap names: below, SB holds WB and SL is free, as the comments at their declarations say.
  sf_chol_inv(SB, red, SL, rs, mp, info);
  {
    double* t_ = SB;
    SB = SL;
    SL = t_;
  }
  if (tid < 64) {
    double lg = tid < mp ? -log(rs[tid]) : 0.0;
    lg = warp_sum(lg);
    if (lane == 0) lgs[warp] = lg;
  }
  __syncthreads();
  if (tid == 0) (a.logdetB + off)[0] = lgs[0] + lgs[1];
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    (a.WBg + off)[i * SF_MP + j] = SB[i * SF_LD + j];
  }
  // c = WB ae ; chat = WB^T c ; u = W^T chat
  if (tid < mp) {
    double s = 0.0;
    for (int k = 0; k <= tid; k++) s = fma(SB[tid * SF_LD + k], ae[k], s);
    cv[tid] = s;
  }
  __syncthreads();
  if (tid < mp) {
    double s = 0.0;
    for (int k = tid; k < mp; k++) s = fma(SB[k * SF_LD + tid], cv[k], s);
    chat[tid] = s;
  }
  __syncthreads();
  if (tid < mp) {
    double s = 0.0;
    for (int k = tid; k < mp; k++) s = fma(SW[k * SF_LD + tid], chat[k], s);
    uv[tid] = s;
    (a.uvec + off)[tid] = s;
  }
  // Binv = WB^T WB -> SL
  sf_mm<true, false>(SL, SB, SB, mp, 1.0);
  __syncthreads();
  // the bound's scalars: tr(AATs), tr(Binv), |c|^2, chat^T AATs chat (real rows only)
  {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (tid < m) {
      v[0] = SA[tid * SF_LD + tid];
      v[1] = SL[tid * SF_LD + tid];
      v[2] = cv[tid] * cv[tid];
    }
    for (int e = tid; e < m * m; e += SF_THREADS) {
      const int i = e / m, j = e - i * m;
      v[3] = fma(chat[i] * chat[j], SA[i * SF_LD + j], v[3]);
    }
    sf_block_sum<4>(v, red);
    if (tid == 0) {
      double* sc = a.scal + off;
      sc[0] = v[0], sc[1] = v[1], sc[2] = v[2], sc[3] = a.yy[blockIdx.y], sc[4] = v[3];
    }
  }
  // Rm = (I - Binv) - chat chat^T -> SB ;  RA = Rm - AATs -> SA
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    const double r = ((i == j ? 1.0 : 0.0) - SL[i * SF_LD + j]) - chat[i] * chat[j];
    SB[i * SF_LD + j] = r;
    SA[i * SF_LD + j] = r - SA[i * SF_LD + j];
  }
  __syncthreads();
  // RW = W^T Rm -> SL -> global
  sf_mm<true, false>(SL, SW, SB, mp, 1.0);
  __syncthreads();
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    (a.RW + off)[i * SF_MP + j] = SL[i * SF_LD + j];
  }
  // T1 = W^T RA -> SB ;  Guu = 1/2 T1 W -> SA
  sf_mm<true, false>(SB, SW, SA, mp, 1.0);
  __syncthreads();
  sf_mm<false, false>(SA, SB, SW, mp, 0.5);
  __syncthreads();
  // chain rule through Kuu (g = Guu; k(z_i, z_j) depends on z_i twice).  First w = g F and the variance term, one entry
  // per thread at a time (w -> SB, T1 is dead); then one thread per (row, feature) pair sums over the columns.
  double gvar = 0.0;
  for (int e = tid; e < mp * mp; e += SF_THREADS) {
    const int i = e / mp, j = e - i * mp;
    double r2 = 0.0;
    for (int dd = 0; dd < D; dd++) {
      const double df = zs[i * D + dd] - zs[j * D + dd];
      r2 = fma(df, df, r2);
    }
    const double g = (i < m && j < m) ? SA[i * SF_LD + j] : 0.0;
    double kval, fval;
    kernel_eval<KID>(r2, kval, fval);
    gvar = fma(g, kval, gvar);
    SB[i * SF_LD + j] = g * fval;
  }
  {
    double v[1] = {gvar};
    sf_block_sum<1>(v, red);  // (its leading barrier also publishes w)
    gvar = v[0];
  }
  double* slp = SL;  // [mp][D] partial lengthscale terms per row (RW has been stored)
  for (int pr = tid; pr < mp * D; pr += SF_THREADS) {
    const int i = pr / D, dd = pr - i * D;
    const double zi = zs[pr];
    double zr = 0.0, sl = 0.0;
    for (int j = 0; j < mp; j++) {
      const double df = zi - zs[j * D + dd];
      const double wd = SB[i * SF_LD + j] * df;
      zr += wd;
      sl = fma(wd, df, sl);
    }
    (a.zpB + off)[pr] = 2.0 * zr;
    slp[pr] = sl;
  }
  __syncthreads();
  if (tid == 0) (a.partB + off)[0] = gvar;
  if (tid >= 1 && tid < 1 + D) {
    double sacc = 0.0;
    for (int i = 0; i < mp; i++) sacc += slp[i * D + tid - 1];
    (a.partB + off)[tid] = sacc;
  }
}

// ---- 4. per (tile, model): chain rule through Kuf ----------------------------------------------------------------------
template <int KID>
__global__ void __launch_bounds__(SF_THREADS, 2) sf_backward_kernel(const SfArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int D = a.D, m = a.m, mp = a.mp, mt = mp >> 3;
  const SfTileSmem sm = sf_tile_layout(smem, D, mp);
  double *xsT = sm.xsT, *zs = sm.zs, *KA = sm.KA, *ysm = sm.ysm, *us = sm.us;
  __shared__ double glw[8][SF_MAX_D + 1];  // per warp: variance term, lengthscale terms
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int model = blockIdx.y, n0 = blockIdx.x * SF_TN;
  const long off = (long)model * a.bs;
  const double* theta = a.theta + off;
  const double inv_s2 = 1.0 / theta[1];
  const bool row_warp = warp < mt;
  const int irow = 8 * warp + g;
  // A fragments of RW (full) for this warp's rows
  double rfrag[2 * (SF_MP / 8)];
#pragma unroll
  for (int k4 = 0; k4 < 2 * (SF_MP / 8); k4++)
    rfrag[k4] = (row_warp && 4 * k4 < mp) ? (a.RW + off)[irow * SF_MP + 4 * k4 + q] : 0.0;
  {
    const double* Apg = a.Ap + off + n0;
    for (int e = tid; e < mp * (SF_TN / 2); e += SF_THREADS) {
      const int k = e >> 6, c = 2 * (e & 63);
      *reinterpret_cast<double2*>(KA + k * SF_LDK + c) = *reinterpret_cast<const double2*>(Apg + (long)k * a.n_pad + c);
    }
  }
  sf_stage_rows(a, n0, theta, xsT, ysm, model);
  if (tid < mp) us[tid] = (a.uvec + off)[tid];
  for (int e = tid; e < mp * D; e += SF_THREADS) zs[e] = (a.Zs + off)[e];
  for (int e = tid; e < 8 * (SF_MAX_D + 1); e += SF_THREADS) (&glw[0][0])[e] = 0.0;
  double* zp = a.zpA + off + (long)blockIdx.x * SF_MP * D;
  __syncthreads();
  if (row_warp) {
    // G1 = RW A'
    double acc[16][2];
#pragma unroll
    for (int ct = 0; ct < 16; ct++) acc[ct][0] = acc[ct][1] = 0.0;
#pragma unroll
    for (int k4 = 0; k4 < 2 * (SF_MP / 8); k4++) {
      if (4 * k4 < mp) {
        const double* bp = KA + (4 * k4 + q) * SF_LDK + g;
#pragma unroll
        for (int ct = 0; ct < 16; ct++) dmma(acc[ct][0], acc[ct][1], rfrag[k4], bp[8 * ct]);
      }
    }
    // g = (G1 + u y^T) / s2 ; acc <- g * F ; variance term   (k / variance and F were stored by the forward pass)
    const double ui = us[irow];
    const double* kp = a.Kv + off + (long)irow * a.n_pad + n0 + 2 * q;
    const double* fp = a.Fv + off + (long)irow * a.n_pad + n0 + 2 * q;
    double gvar = 0.0;
#pragma unroll
    for (int ct = 0; ct < 16; ct++) {
      const int c = 8 * ct + 2 * q;
      const double2 kv = *reinterpret_cast<const double2*>(kp + 8 * ct);
      const double2 fv = *reinterpret_cast<const double2*>(fp + 8 * ct);
      const double g0 = (irow < m && n0 + c < a.n) ? fma(ui, ysm[c], acc[ct][0]) * inv_s2 : 0.0;
      const double g1 = (irow < m && n0 + c + 1 < a.n) ? fma(ui, ysm[c + 1], acc[ct][1]) * inv_s2 : 0.0;
      gvar = fma(g0, kv.x, gvar);
      gvar = fma(g1, kv.y, gvar);
      acc[ct][0] = g0 * fv.x;
      acc[ct][1] = g1 * fv.y;
    }
    gvar = warp_sum(gvar);
    if (lane == 0) glw[warp][0] = gvar;
    // lengthscale and Z terms
    for (int dd = 0; dd < D; dd++) {
      const double zi = zs[irow * D + dd];
      const double* xp = xsT + dd * SF_TN + 2 * q;
      double zr = 0.0, sl = 0.0;
#pragma unroll
      for (int ct = 0; ct < 16; ct++) {
        const double2 xv = *reinterpret_cast<const double2*>(xp + 8 * ct);
        const double d0 = zi - xv.x, d1 = zi - xv.y;
        const double w0 = acc[ct][0] * d0, w1 = acc[ct][1] * d1;
        zr += w0;
        zr += w1;
        sl = fma(w0, d0, sl);
        sl = fma(w1, d1, sl);
      }
      zr += __shfl_xor_sync(0xffffffffu, zr, 1);
      zr += __shfl_xor_sync(0xffffffffu, zr, 2);
      if (q == 0) zp[irow * D + dd] = zr;  // row irow belongs to this warp alone
      sl = warp_sum(sl);
      if (lane == 0) glw[warp][1 + dd] = sl;
    }
  }
  __syncthreads();
  if (tid < 1 + D) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) s += glw[wv][tid];
    (a.partA + off)[(long)blockIdx.x * (1 + D) + tid] = s;
  }
}

// ---- 5. per model: bound and gradient from the pieces (one CTA; same formulas as sgpr_finalize_kernel) -----------------
//   result = [elbo, dF/dlog variance, dF/dlog noise, dF/dlog l_0.., dF/dZ (m x D)]
__device__ __forceinline__ void sf_finalize_body(const SfArgs& a, double* __restrict__ result) {
  const int tid = threadIdx.x, D = a.D, m = a.m, n = a.n;
  const long off = (long)blockIdx.y * a.bs;
  const double* theta = a.theta + off;
  const double* scal = a.scal + off;
  const double variance = theta[0], s2 = theta[1];
  if (tid == 0) {
    const double ldb = (a.logdetB + off)[0];
    const double tr_aat = scal[0], tr_binv = scal[1], cc = scal[2], yy = scal[3] / s2, caac = scal[4];
    result[0] = -0.5 * (double)n * 1.8378770664093453 -
                (ldb + 0.5 * (double)n * log(s2) + 0.5 * ((double)n * variance / s2 - tr_aat)) - 0.5 * (yy - cc);
    result[2] = -0.5 * (double)n + 0.5 * ((double)m - tr_binv) + 0.5 * yy - cc + 0.5 * caac + 0.5 * (double)n * variance / s2 -
                0.5 * tr_aat;
  }
  if (tid < 1 + D) {
    const double* pa = a.partA + off + tid;
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < a.ntn; i++) s += pa[(long)i * (1 + D)];
    s += (a.partB + off)[tid];
    s *= variance;
    if (tid == 0)
      result[1] = s - 0.5 * (double)n * variance / s2;
    else
      result[2 + tid] = s;
  }
  for (int e = tid; e < m * D; e += SF_THREADS) {
    const int dd = e % D;
    const double* za = a.zpA + off + e;
    double s = 0.0;
#pragma unroll 8
    for (int t = 0; t < a.ntn; t++) s += za[(long)t * SF_MP * D];
    s += (a.zpB + off)[e];
    result[3 + D + e] = -variance * s / theta[2 + dd];
  }
}

static __global__ void __launch_bounds__(SF_THREADS) sf_finalize_kernel(const SfArgs a, double* __restrict__ result) {
  sf_finalize_body(a, result + (long)blockIdx.y * a.bs);
}

// ---- 6. prediction (predict_y, gpr.py:337) for a conditioned batch: per (tile of 128 test rows, model) ----------------
//   mean = Kus^T u,  var = variance + s2 + |WB W Kus|^2 - |W Kus|^2   (columnwise; GPflow SGPR.predict_f + likelihood noise)
// xs: [t_pad][D] test inputs (device), mean / var: [t][P] row-major.
template <int KID>
__global__ void __launch_bounds__(SF_THREADS, 2) sf_predict_kernel(const SfArgs a, const double* __restrict__ xs, int t_total,
                                                                   double* __restrict__ mean, double* __restrict__ var, int P) {
  extern __shared__ __align__(16) double smem[];
  const int D = a.D, m = a.m, mp = a.mp, na = mp >> 3, mt = mp >> 3;
  const SfTileSmem sm = sf_tile_layout(smem, D, mp);
  double *xsT = sm.xsT, *zs = sm.zs, *KA = sm.KA, *us = sm.us;
  __shared__ double red[3][8][SF_TN];  // per warp and column: mean, |W Kus|^2, |WB W Kus|^2 partial sums
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int model = blockIdx.y, t0 = blockIdx.x * SF_TN;
  const long off = (long)model * a.bs;
  const double* theta = a.theta + off;
  const double variance = theta[0];
  const bool row_warp = warp < mt;
  const int irow = 8 * warp + g;
  for (int e = tid; e < SF_TN * D; e += SF_THREADS) {
    const int c = e / D, dd = e - c * D;
    xsT[dd * SF_TN + c] = t0 + c < t_total ? xs[(long)(t0 + c) * D + dd] / theta[2 + dd] : 0.0;
  }
  for (int e = tid; e < mp * D; e += SF_THREADS) zs[e] = (a.Zs + off)[e];
  if (tid < mp) us[tid] = (a.uvec + off)[tid];
  double wfrag[2 * (SF_MP / 8)], bfrag[2 * (SF_MP / 8)];  // rows [8w, 8w + 8) of W and of WB (both lower triangular)
#pragma unroll
  for (int k4 = 0; k4 < 2 * (SF_MP / 8); k4++) {
    const bool on = row_warp && k4 < 2 * (warp + 1);
    wfrag[k4] = on ? (a.W + off)[irow * SF_MP + 4 * k4 + q] : 0.0;
    bfrag[k4] = on ? (a.WBg + off)[irow * SF_MP + 4 * k4 + q] : 0.0;
  }
  for (int e = tid; e < 3 * 8 * SF_TN; e += SF_THREADS) (&red[0][0][0])[e] = 0.0;
  __syncthreads();
  // Kus entries (rows warp + 8 r, columns lane + 32 s) and this warp's part of the mean
  double msum[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int r = 0; r < na; r++) {
    const int i = warp + 8 * r;
    double r2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int dd = 0; dd < D; dd++) {
      const double zi = zs[i * D + dd];
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const double df = zi - xsT[dd * SF_TN + lane + 32 * s];
        r2[s] = fma(df, df, r2[s]);
      }
    }
    const double ui = us[i];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const double k = i < m ? variance * kernel_value<KID>(r2[s]) : 0.0;
      KA[i * SF_LDK + lane + 32 * s] = k;
      msum[s] = fma(k, ui, msum[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < 4; s++) red[0][warp][lane + 32 * s] = msum[s];
  __syncthreads();
  double acc[16][2];
  auto product = [&](const double (&frag)[2 * (SF_MP / 8)], int which) {
#pragma unroll
    for (int ct = 0; ct < 16; ct++) acc[ct][0] = acc[ct][1] = 0.0;
    if (row_warp) {
#pragma unroll
      for (int k4 = 0; k4 < 2 * (SF_MP / 8); k4++) {
        if (k4 < 2 * (warp + 1)) {
          const double* bp = KA + (4 * k4 + q) * SF_LDK + g;
#pragma unroll
          for (int ct = 0; ct < 16; ct++) dmma(acc[ct][0], acc[ct][1], frag[k4], bp[8 * ct]);
        }
      }
      // column sums of squares over this warp's 8 rows (lanes with equal q hold the same columns)
#pragma unroll
      for (int ct = 0; ct < 16; ct++) {
        double s0 = acc[ct][0] * acc[ct][0], s1 = acc[ct][1] * acc[ct][1];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (g == 0) {
          red[which][warp][8 * ct + 2 * q] = s0;
          red[which][warp][8 * ct + 2 * q + 1] = s1;
        }
      }
    }
  };
  product(wfrag, 1);  // tmp1 = W Kus
  __syncthreads();    // every read of the Kus tile is done: the buffer becomes tmp1
  if (row_warp) {
#pragma unroll
    for (int ct = 0; ct < 16; ct++)
      *reinterpret_cast<double2*>(KA + irow * SF_LDK + 8 * ct + 2 * q) = make_double2(acc[ct][0], acc[ct][1]);
  }
  __syncthreads();
  product(bfrag, 2);  // tmp2 = WB tmp1
  __syncthreads();
  if (tid < SF_TN && t0 + tid < t_total) {
    double mu = 0.0, q1 = 0.0, p2 = 0.0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) mu += red[0][wv][tid], q1 += red[1][wv][tid], p2 += red[2][wv][tid];
    mean[(long)(t0 + tid) * P + model] = mu;
    var[(long)(t0 + tid) * P + model] = variance + theta[1] + p2 - q1;
  }
}

}  // namespace gpras
