// Device side of PreProcessor.fit / transform (gpras/preprocess.py:947-1039): the cells -> modes direction.
//
//   column statistics  max / min / sum over samples of the (depth-converted) input: one HBM pass over the N x C matrix,
//                      feeding the wetness classification (preprocess.py:1096-1132) and input_mean (:976);
//   centre + weight    Xw = (val(x) - mean) * weight, zero on dry / padded cells (:976-982), materialised once for the
//                      Gram matrix and the EOF back-projection of the PCA fit;
//   projection         z = ((val(x) - mean) * weight) @ eofs^T, standardised (:1028-1037): a long-k skinny GEMM on the
//                      DMMA pipe with the centring / weighting / depth clamp applied to the A fragments on the fly, so x
//                      is read exactly once (8 B per cell-sample) and never rewritten.
//
// All cell-indexed vectors live in FULL cell space (length c_pad); always-dry cells simply carry weight 0, so no
// compaction is needed on the device (the host compacts when it hands eofs / input_mean back to Python).
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace gpras {

enum HydraulicParameter { HP_WSE = 0, HP_DEPTH = 1, HP_VELOCITY = 2 };

// ---- column statistics ------------------------------------------------------------------------
// grid (ceil(c / 256), row_splits), 128 threads, each thread owns two adjacent columns of a row range.
// part[split][3][c_pad]: max, min, sum.  clamp != 0: values are max(x - elev, 0) (hydraulic_parameter == "depth").
static __global__ void __launch_bounds__(128) colstats_kernel(const double* __restrict__ X, long ldx, int n, int c,
                                                              const double* __restrict__ elev, int clamp, int rows_per_split,
                                                              double* __restrict__ part, long c_pad) {
  const int c0 = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c0 >= c) return;
  const bool two = c0 + 1 < c;
  const int r0 = blockIdx.y * rows_per_split;
  int r1 = r0 + rows_per_split;
  if (r1 > n) r1 = n;
  const double e0 = clamp ? elev[c0] : 0.0, e1 = (clamp && two) ? elev[c0 + 1] : 0.0;
  double mx0 = -INFINITY, mx1 = -INFINITY, mn0 = INFINITY, mn1 = INFINITY, s0 = 0.0, s1 = 0.0;
  const double* p = X + (long)r0 * ldx + c0;
#pragma unroll 8
  for (int r = r0; r < r1; r++, p += ldx) {
    double v0 = __ldg(p), v1 = two ? __ldg(p + 1) : 0.0;
    if (clamp) v0 = fmax(v0 - e0, 0.0), v1 = fmax(v1 - e1, 0.0);
    mx0 = fmax(mx0, v0), mn0 = fmin(mn0, v0), s0 += v0;
    mx1 = fmax(mx1, v1), mn1 = fmin(mn1, v1), s1 += v1;
  }
  double* o = part + (long)blockIdx.y * 3 * c_pad;
  o[c0] = mx0, o[c_pad + c0] = mn0, o[2 * c_pad + c0] = s0;
  if (two) o[c0 + 1] = mx1, o[c_pad + c0 + 1] = mn1, o[2 * c_pad + c0 + 1] = s1;
}

// The same statistics with the matrix streamed by the TMA unit: a CTA owns a slab of 256 columns and a row range, one thread
// keeps CS_STAGES boxes of CS_ROWS x 256 values in flight (cp.async.bulk.tensor -> shared memory, completion on an mbarrier),
// all 128 threads reduce their two columns out of shared memory.  Measured on B200 (tools/microbench/read_bw.cu, 8192 x
// 200 000): LDG slabs 4.2-5.3 TB/s, TMA boxes 7.4-7.5 TB/s (contiguous reads: 7.15 TB/s).  Out-of-range rows / columns of a
// box read as zero and are masked here (rows by the loop bound, columns by `c`).
constexpr int CS_ROWS = 16, CS_COLS = 256, CS_STAGES = 4;
constexpr int CS_SMEM_BYTES = CS_STAGES * CS_ROWS * CS_COLS * (int)sizeof(double) + 128;
static __global__ void __launch_bounds__(128) colstats_tma_kernel(const __grid_constant__ CUtensorMap map, int n, int c,
                                                                  const double* __restrict__ elev, int clamp, int rows_per_split,
                                                                  double* __restrict__ part, long c_pad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[CS_STAGES];
  double* tile = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int STAGE_DOUBLES = CS_ROWS * CS_COLS;
  constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * sizeof(double);
  const int tid = threadIdx.x;
  const int slab0 = blockIdx.x * CS_COLS;
  const int c0 = slab0 + 2 * tid;
  const int r0 = blockIdx.y * rows_per_split;
  int r1 = r0 + rows_per_split;
  if (r1 > n) r1 = n;
  const int n_it = r1 > r0 ? (r1 - r0 + CS_ROWS - 1) / CS_ROWS : 0;
  if (tid == 0) {
    for (int s = 0; s < CS_STAGES; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < CS_STAGES && s < n_it; s++) {
      mbar_expect_tx(&full[s], STAGE_BYTES);
      tma_load_2d(tile + s * STAGE_DOUBLES, &map, slab0, r0 + s * CS_ROWS, &full[s]);
    }
  const bool live0 = c0 < c, two = c0 + 1 < c;
  const double e0 = (clamp && live0) ? elev[c0] : 0.0, e1 = (clamp && two) ? elev[c0 + 1] : 0.0;
  double mx0 = -INFINITY, mx1 = -INFINITY, mn0 = INFINITY, mn1 = INFINITY, s0 = 0.0, s1 = 0.0;
  for (int it = 0; it < n_it; it++) {
    const int s = it % CS_STAGES;
    mbar_wait(&full[s], (it / CS_STAGES) & 1);
    const double* t = tile + s * STAGE_DOUBLES + 2 * tid;
    const int rows = r1 - (r0 + it * CS_ROWS) < CS_ROWS ? r1 - (r0 + it * CS_ROWS) : CS_ROWS;
#pragma unroll 4
    for (int r = 0; r < rows; r++) {
      const double2 v = *reinterpret_cast<const double2*>(t + r * CS_COLS);
      double v0 = v.x, v1 = v.y;
      if (clamp) v0 = fmax(v0 - e0, 0.0), v1 = fmax(v1 - e1, 0.0);
      mx0 = fmax(mx0, v0), mn0 = fmin(mn0, v0), s0 += v0;
      mx1 = fmax(mx1, v1), mn1 = fmin(mn1, v1), s1 += v1;
    }
    __syncthreads();  // every thread is done with this stage: refill it
    if (tid == 0 && it + CS_STAGES < n_it) {
      mbar_expect_tx(&full[s], STAGE_BYTES);
      tma_load_2d(tile + s * STAGE_DOUBLES, &map, slab0, r0 + (it + CS_STAGES) * CS_ROWS, &full[s]);
    }
  }
  if (!live0) return;
  double* o = part + (long)blockIdx.y * 3 * c_pad;
  o[c0] = mx0, o[c_pad + c0] = mn0, o[2 * c_pad + c0] = s0;
  if (two) o[c0 + 1] = mx1, o[c_pad + c0 + 1] = mn1, o[2 * c_pad + c0 + 1] = s1;
}

// Fold the row-range partials (fixed order) and classify: cls 1 = always dry, 2 = transitional, 3 = always flooded,
// 0 = exactly on the threshold (the reference leaves those unset).  mean[c] = sum / n; wfull[c] = dry ? 0 : weights[c].
static __global__ void colstats_finish_kernel(const double* __restrict__ part, int nsplit, long c_pad, int c, int n, int hp,
                                              const double* __restrict__ elev, const double* __restrict__ weights,
                                              double wet_threshold, double* __restrict__ mean, double* __restrict__ wfull,
                                              int* __restrict__ cls) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c_pad) return;
  if (j >= c) {
    mean[j] = 0.0, wfull[j] = 0.0, cls[j] = 1;
    return;
  }
  double mx = -INFINITY, mn = INFINITY, s = 0.0;
  for (int z = 0; z < nsplit; z++) {
    const double* o = part + (long)z * 3 * c_pad;
    mx = fmax(mx, o[j]), mn = fmin(mn, o[c_pad + j]), s += o[2 * c_pad + j];
  }
  int k = 2;
  if (hp != HP_VELOCITY) {
    const double off = hp == HP_WSE ? elev[j] : 0.0;  // depth mode: the values were already converted
    const double dmax = mx - off, dmin = mn - off;
    k = 0;
    if (dmax < wet_threshold) k = 1;
    if (dmax > wet_threshold) k = 2;
    if (dmin > wet_threshold) k = 3;
  }
  cls[j] = k;
  mean[j] = s / (double)n;
  wfull[j] = k == 1 ? 0.0 : weights[j];
}

// Xw[i][j] = (val(x[i][j]) - mean[j]) * wfull[j] for i < n, j < c; 0 elsewhere (n_pad x c_pad, pitch c_pad).
static __global__ void center_weight_kernel(const double* __restrict__ X, long ldx, int n, int c, const double* __restrict__ elev,
                                            int clamp, const double* __restrict__ mean, const double* __restrict__ wfull,
                                            double* __restrict__ Xw, long c_pad) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= c_pad) return;
  double v = 0.0;
  if (i < n && j < c) {
    v = X[(long)i * ldx + j];
    if (clamp) v = fmax(v - elev[j], 0.0);
    v = (v - mean[j]) * wfull[j];
  }
  Xw[(long)i * c_pad + j] = v;
}

// ---- projection -------------------------------------------------------------------------------
// part[z][i][p] = sum_{c in chunk z} (val(x[i][c]) - mean[c]) * wfull[c] * E[p][c]
// CTA: 128 rows x PN modes, 8 warps of 16 rows; BK = 16 cells per stage, 4-stage cp.async ring carrying the x tile,
// the E tile and the mean / weight / elevation slices (24 KB per stage at PN = 32, so two CTAs share an SM).  X rows must be 16-byte aligned with c_pad readable columns
// (the host stages the input into a padded buffer when that does not hold).
constexpr int PROJ_THREADS = 256;
constexpr int PROJ_BK = 16;
constexpr int PROJ_STAGES_MAX = 4;

template <int PN>
struct ProjCfg {
  static constexpr int LD = PROJ_BK + 4;
  static constexpr int A_DOUBLES = 128 * LD, B_DOUBLES = PN * LD, V_DOUBLES = 3 * PROJ_BK;
  static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES + V_DOUBLES;
  static constexpr int STAGES = PN <= 32 ? PROJ_STAGES_MAX : 3;  // 64 modes: three 31 KB stages, so that two CTAs still share an SM
  static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);
};

template <int PN>
__global__ void __launch_bounds__(PROJ_THREADS, 2)
project_kernel(const double* __restrict__ X, long ldx, int n, int c, const double* __restrict__ elev, int clamp,
               const double* __restrict__ mean, const double* __restrict__ wfull, const double* __restrict__ E, long lde,
               int k_stages_total, int stages_per_split, double* __restrict__ part, long n_pad) {
  using Cfg = ProjCfg<PN>;
  constexpr int NF = PN / 8, LD = Cfg::LD, PROJ_STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp * 16;
  const int ti = blockIdx.x;
  const int s_begin = blockIdx.y * stages_per_split;
  int s_end = s_begin + stages_per_split;
  if (s_end > k_stages_total) s_end = k_stages_total;
  const int nk = s_end - s_begin;

  auto load_stage = [&](int slot, int ks) {
    double* sA = smem + slot * Cfg::STAGE_DOUBLES;
    double* sB = sA + Cfg::A_DOUBLES;
    double* sV = sB + Cfg::B_DOUBLES;
    const long c0 = (long)ks * PROJ_BK;
    constexpr int CPK = PROJ_BK / 2;
    for (int ch = tid; ch < 128 * CPK; ch += PROJ_THREADS) {
      const int row = ch / CPK, kc = ch - row * CPK;
      int gr = ti * 128 + row;
      if (gr >= n) gr = n - 1;  // padding rows re-read the last valid row (their results are never used)
      cp_async16(sA + row * LD + 2 * kc, X + (long)gr * ldx + c0 + 2 * kc);
    }
    for (int ch = tid; ch < PN * CPK; ch += PROJ_THREADS) {
      const int row = ch / CPK, kc = ch - row * CPK;
      cp_async16(sB + row * LD + 2 * kc, E + (long)row * lde + c0 + 2 * kc);
    }
    if (tid < 3 * CPK) {
      const int which = tid / CPK, kc = tid - which * CPK;
      const double* src = which == 0 ? mean : (which == 1 ? wfull : elev);
      if (which < 2 || clamp) cp_async16(sV + which * PROJ_BK + 2 * kc, src + c0 + 2 * kc);
    }
  };
#pragma unroll
  for (int s = 0; s < PROJ_STAGES - 1; s++) {
    if (s < nk) load_stage(s, s_begin + s);
    cp_async_commit();
  }
  double acc[2][NF][2];
#pragma unroll
  for (int f = 0; f < 2; f++)
#pragma unroll
    for (int h = 0; h < NF; h++) acc[f][h][0] = acc[f][h][1] = 0.0;

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<PROJ_STAGES - 2>();
    __syncthreads();
    const int nx = kt + PROJ_STAGES - 1;
    if (nx < nk) load_stage(nx % PROJ_STAGES, s_begin + nx);
    cp_async_commit();
    const double* sA = smem + (kt % PROJ_STAGES) * Cfg::STAGE_DOUBLES;
    const double* sB = sA + Cfg::A_DOUBLES;
    const double* sV = sB + Cfg::B_DOUBLES;
    const long cbase = (long)(s_begin + kt) * PROJ_BK;
#pragma unroll
    for (int ks = 0; ks < PROJ_BK / 4; ks++) {
      const int kk = 4 * ks + q;
      const double mu = sV[kk], wt = sV[PROJ_BK + kk];
      const bool live = cbase + kk < c;
      double av[2], bv[NF];
#pragma unroll
      for (int f = 0; f < 2; f++) {
        double v = sA[(wm + 8 * f + g) * LD + kk];
        if (clamp) v = fmax(v - sV[2 * PROJ_BK + kk], 0.0);
        av[f] = live ? (v - mu) * wt : 0.0;
      }
#pragma unroll
      for (int h = 0; h < NF; h++) bv[h] = sB[(8 * h + g) * LD + kk];
#pragma unroll
      for (int f = 0; f < 2; f++)
#pragma unroll
        for (int h = 0; h < NF; h++) dmma(acc[f][h][0], acc[f][h][1], av[f], bv[h]);
    }
  }
  cp_async_wait<0>();
  double* o = part + ((long)blockIdx.y * n_pad + (long)ti * 128 + wm) * PN;
#pragma unroll
  for (int f = 0; f < 2; f++)
#pragma unroll
    for (int h = 0; h < NF; h++)
      *reinterpret_cast<double2*>(o + (long)(8 * f + g) * PN + 8 * h + 2 * q) = make_double2(acc[f][h][0], acc[f][h][1]);
}

// The projection with its HBM stream (the 128 x 16 tile of x per stage) moved by the TMA unit: one cp.async.bulk.tensor box per
// stage issued by one thread, landing with the 128-byte swizzle (chunk ^= row & 7) in an unpadded 1 KB-aligned tile; the DMMA
// fragments take their four k values in the order k(q, ks) = 2 ks + (q & 1) + 8 (q >> 1), for which the eight rows x four
// lanes of a fragment read hit sixteen different banks per half-warp (any k order is valid as long as A, B and the per-cell
// vectors use the same one).  The small operands (E tile, mean / weight / elevation slices) stay on cp.async in their padded
// layout.  Rows past n and cells past c read as zero.
template <int PN>
struct ProjTmaCfg {
  // pitch of the padded E tile: == 2 mod 16, so that with the k order above (k in {2ks, 2ks+1, 2ks+8, 2ks+9}) the four rows x
  // four k values of a half-warp's B-fragment read fall in sixteen different banks (+4, right for the plain k order, gave a
  // two-way conflict here: 1.0e8 conflict cycles in the first ncu capture)
  static constexpr int LD = PROJ_BK + 2;
  static constexpr int A_DOUBLES = 128 * PROJ_BK, B_DOUBLES = PN * LD, V_DOUBLES = 3 * PROJ_BK;
  static constexpr int SMALL_DOUBLES = B_DOUBLES + V_DOUBLES;
  static constexpr int STAGES = 4;
  static constexpr int SMEM_BYTES = STAGES * (A_DOUBLES + SMALL_DOUBLES) * (int)sizeof(double) + 1024;
};

template <int PN>
__global__ void __launch_bounds__(PROJ_THREADS, 2)
project_tma_kernel(const __grid_constant__ CUtensorMap xmap, int n, int c, const double* __restrict__ elev, int clamp,
                   const double* __restrict__ mean, const double* __restrict__ wfull, const double* __restrict__ E, long lde,
                   int k_stages_total, int stages_per_split, double* __restrict__ part, long n_pad) {
  using Cfg = ProjTmaCfg<PN>;
  constexpr int NF = PN / 8, LD = Cfg::LD, STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[STAGES];
  double* sAall = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* sSmall = sAall + STAGES * Cfg::A_DOUBLES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp * 16;
  const int ti = blockIdx.x;
  const int s_begin = blockIdx.y * stages_per_split;
  int s_end = s_begin + stages_per_split;
  if (s_end > k_stages_total) s_end = k_stages_total;
  const int nk = s_end - s_begin;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto load_stage = [&](int slot, int ks) {
    const long c0 = (long)ks * PROJ_BK;
    if (tid == 0) {
      mbar_expect_tx(&full[slot], Cfg::A_DOUBLES * (uint32_t)sizeof(double));
      tma_load_2d(sAall + slot * Cfg::A_DOUBLES, &xmap, (int)c0, ti * 128, &full[slot]);
    }
    double* sB = sSmall + slot * Cfg::SMALL_DOUBLES;
    double* sV = sB + Cfg::B_DOUBLES;
    constexpr int CPK = PROJ_BK / 2;
    for (int ch = tid; ch < PN * CPK; ch += PROJ_THREADS) {
      const int row = ch / CPK, kc = ch - row * CPK;
      cp_async16(sB + row * LD + 2 * kc, E + (long)row * lde + c0 + 2 * kc);
    }
    if (tid < 3 * CPK) {
      const int which = tid / CPK, kc = tid - which * CPK;
      const double* src = which == 0 ? mean : (which == 1 ? wfull : elev);
      if (which < 2 || clamp) cp_async16(sV + which * PROJ_BK + 2 * kc, src + c0 + 2 * kc);
    }
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s, s_begin + s);
    cp_async_commit();
  }
  double acc[2][NF][2];
#pragma unroll
  for (int f = 0; f < 2; f++)
#pragma unroll
    for (int h = 0; h < NF; h++) acc[f][h][0] = acc[f][h][1] = 0.0;

  for (int kt = 0; kt < nk; kt++) {
    const int slot = kt % STAGES;
    cp_async_wait<STAGES - 2>();
    mbar_wait(&full[slot], (kt / STAGES) & 1);
    __syncthreads();
    const int nx = kt + STAGES - 1;
    if (nx < nk) load_stage(nx % STAGES, s_begin + nx);
    cp_async_commit();
    const double* sA = sAall + slot * Cfg::A_DOUBLES;
    const double* sB = sSmall + slot * Cfg::SMALL_DOUBLES;
    const double* sV = sB + Cfg::B_DOUBLES;
    const long cbase = (long)(s_begin + kt) * PROJ_BK;
#pragma unroll
    for (int ks = 0; ks < PROJ_BK / 4; ks++) {
      const int kk = 2 * ks + (q & 1) + 8 * (q >> 1);                 // this lane's k of the step (see above)
      const int sw = (((ks + 4 * (q >> 1)) ^ g) << 1) + (q & 1);       // its swizzled position in a row whose index is g mod 8
      const double mu = sV[kk], wt = sV[PROJ_BK + kk];
      const bool live = cbase + kk < c;
      double av[2], bv[NF];
#pragma unroll
      for (int f = 0; f < 2; f++) {
        double v = sA[(wm + 8 * f + g) * PROJ_BK + sw];
        if (clamp) v = fmax(v - sV[2 * PROJ_BK + kk], 0.0);
        av[f] = live ? (v - mu) * wt : 0.0;
      }
#pragma unroll
      for (int h = 0; h < NF; h++) bv[h] = sB[(8 * h + g) * LD + kk];
#pragma unroll
      for (int f = 0; f < 2; f++)
#pragma unroll
        for (int h = 0; h < NF; h++) dmma(acc[f][h][0], acc[f][h][1], av[f], bv[h]);
    }
  }
  cp_async_wait<0>();
  double* o = part + ((long)blockIdx.y * n_pad + (long)ti * 128 + wm) * PN;
#pragma unroll
  for (int f = 0; f < 2; f++)
#pragma unroll
    for (int h = 0; h < NF; h++)
      *reinterpret_cast<double2*>(o + (long)(8 * f + g) * PN + 8 * h + 2 * q) = make_double2(acc[f][h][0], acc[f][h][1]);
}

// z[i][p] = (sum_z part[z][i][p] - x_mean[p]) / x_std[p]   (fixed order; standardise == 0 leaves the raw scores)
static __global__ void project_finish_kernel(const double* __restrict__ part, int nz, long n_pad, int pn, int n, int p,
                                             const double* __restrict__ x_mean, const double* __restrict__ x_std, int standardise,
                                             double* __restrict__ Z, long ldz) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)n * pn) return;
  const long i = e / pn;
  const int j = (int)(e - i * pn);
  if (j >= p) return;
  double s = 0.0;
  for (int z = 0; z < nz; z++) s += part[((long)z * n_pad + i) * pn + j];
  Z[i * ldz + j] = standardise ? (s - x_mean[j]) / x_std[j] : s;
}

// Column mean and population standard deviation of the (n x p) scores (preprocess.py:1005-1007): one CTA per column,
// two passes, fixed order.
static __global__ void __launch_bounds__(256) score_stats_kernel(const double* __restrict__ Z, long ldz, int n,
                                                                 double* __restrict__ mean, double* __restrict__ std) {
  __shared__ double red[8];
  __shared__ double bc;
  const int j = blockIdx.x, tid = threadIdx.x;
  auto block_sum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < 8; i++) s += red[i];
      bc = s;
    }
    __syncthreads();
    return bc;
  };
  double s = 0.0;
  for (int i = tid; i < n; i += 256) s += Z[(long)i * ldz + j];
  const double m = block_sum(s) / (double)n;
  double v = 0.0;
  for (int i = tid; i < n; i += 256) {
    const double d = Z[(long)i * ldz + j] - m;
    v += d * d;
  }
  v = block_sum(v);
  if (tid == 0) mean[j] = m, std[j] = sqrt(v / (double)n);
}

}  // namespace gpras
