// Streaming accuracy metrics (gpras/metrics.py) fused with the modes -> cells expansion: the consumer that makes the
// T x C cell-space prediction unnecessary to materialise (the 3.2 TB output of config 5).
//
// Every metric of gpras/metrics.py:85-324 is a function of a few running reductions over the (timesteps x cells)
// truth x, prediction y and confidence conf of one event:
//     per cell      sum(x-y)  sum((x-y)^2)  sum(conf)  max_t x  max_t y      (x[argmax_t x] == max_t x)
//     per timestep  sum_c(x-y)  sum_c((x-y)^2)  sum_c(conf)  sum_c|x-y|  count_c(|x-y| <= v_tol)
// One kernel accumulates all of them in a single pass.  A CTA owns a 128-cell column tile and streams 64-row tiles:
//   FUSED  : y = max(M[t,:] E[:,c] + bias[c] - elev_y[c], 0)  on the DMMA pipe (E tile resident in shared memory),
//            conf = sqrt(var[t]) * sqrt(S[c])                 (shared theta: the cell variance is rank one),
//            so the only HBM stream is the truth x (8 B per cell-timestep; nothing at all when there is no truth);
//   !FUSED : y and conf are read from (T x C) arrays (24 B per cell-timestep), for callers that hold them.
// Depth conversion (PreProcessor.wse_2_depth, preprocess.py:1041-1045; pipeline.py:265-274) is applied on the fly when
// elevation vectors are supplied.  Reductions have a fixed shape (no atomics): results are bitwise repeatable.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace gpras {

constexpr int MET_THREADS = 256;
constexpr int MET_ROWS = 32;   // timesteps per row tile
constexpr int MET_CELLQ = 5;   // per cell: sum e, sum e^2, sum conf, max x, max y
constexpr int MET_ROWQ = 3;    // per row : sum e, sum e^2, sum conf
constexpr int MET_CTAQ = 2;    // per CTA : sum |e|, count(|e| <= v_tol)
constexpr int MET_RED_LD = MET_ROWS + 4;  // row-sum scratch [quantity][32 slots][36]: conflict-free writes (8q + 2g banks) and reads

struct MetricsArgs {
  // FUSED prediction source
  const double* M;      // (t_pad x ldm) mode-space means
  long ldm;
  const double* var;    // (t_pad) mode-space variance (noise included), shared by all modes
  const double* Vm;     // general-variance kernel: (t_pad x ldm) mode-space variances, one per mode
  const double* E;      // (P16 x lde) folded EOF map, zero on dry / padded cells
  long lde;
  const double* bias;   // (c_pad)
  const double* rootS;  // (c_pad) sqrt(sum_p E[p][c]^2)
  // !FUSED prediction source
  const double* Y;
  long ldy;
  const double* CONF;
  long ldconf;
  // truth (NULL: x == 0)
  const double* X;
  long ldx;
  int x_vec;             // truth rows are 16-byte aligned (ldx even, base aligned): 128-bit loads
  const double* elev_x;  // NULL: truth used as is; else x = max(X - elev_x, 0)
  const double* elev_y;  // NULL: prediction used as is; else y = max(y - elev_y, 0)
  int t_rows;            // valid rows of this block
  int c;                 // valid cells
  int t_tiles, tiles_per_cta;
  double v_tol;
  double* cell_part;     // [gridDim.y][MET_CELLQ][c_pad]
  long c_pad;
  double* row_part;      // [MET_ROWQ][t_tiles * MET_ROWS][n_ctile]
  double* cta_part;      // [gridDim.y][n_ctile][MET_CTAQ]
  int n_ctile;
};

template <int P16>
struct MetCfg {
  static constexpr int LDE = 128 + 4;
  static constexpr int LDA = P16 + 4;
  // E tile, two row tiles of modes, two variance vectors, reduction scratch
  static constexpr int RED_DOUBLES = 2 * 32 * MET_RED_LD + 16;
  static constexpr int SMEM_DOUBLES = P16 * LDE + 2 * MET_ROWS * LDA + 2 * MET_ROWS + RED_DOUBLES;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * (int)sizeof(double);
};
constexpr int MET_PLAIN_SMEM_BYTES = (3 * 32 * MET_RED_LD + 16) * (int)sizeof(double);

// Thread layout: warp w owns the 16 columns [16w, 16w+16) of the CTA's 128-cell tile for all 32 rows of a row tile
// (4 x 2 DMMA accumulator tiles); a thread holds rows 8f+g (f < 4) and columns 16w + 8h + 2q + {0,1} (h < 2).
// Row sums go through shared memory (one slot per thread, then 32-way fixed-order sums); |e| and the match count are
// only ever needed as event totals, so they stay in two per-thread scalars until the CTA ends.  In FUSED mode the
// confidence is separable, conf[t][c] = sqrt(var[t]) * rootS[c]: its row sums are sqrt(var[t]) * sum_c rootS[c] and its
// per-cell sums rootS[c] * sum_t sqrt(var[t]), so it costs no per-element work at all.
template <int P16, bool FUSED>
__global__ void __launch_bounds__(MET_THREADS, FUSED ? 2 : 1) metrics_stream_kernel(const MetricsArgs a) {
  using Cfg = MetCfg<P16>;
  constexpr int NQ = FUSED ? 2 : 3;  // row quantities reduced through shared memory
  extern __shared__ __align__(16) double smem[];
  double* sE = smem;
  double* sA = sE + (FUSED ? P16 * Cfg::LDE : 0);
  double* sV = sA + (FUSED ? 2 * MET_ROWS * Cfg::LDA : 0);
  double* sRed = sV + (FUSED ? 2 * MET_ROWS : 0);  // [NQ][32 slots][MET_RED_LD]
  double* sMisc = sRed + NQ * 32 * MET_RED_LD;     // [16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wn = warp * 16;
  const int tj = blockIdx.x;
  const int t_begin = blockIdx.y * a.tiles_per_cta;
  int t_end = t_begin + a.tiles_per_cta;
  if (t_end > a.t_tiles) t_end = a.t_tiles;
  const long col0 = (long)tj * 128 + wn + 2 * q;  // column of (h = 0, w = 0); (h, w) adds 8h + w
  const bool cols_full = (long)tj * 128 + 128 <= a.c;
  const bool has_ex = a.elev_x != nullptr, has_ey = a.elev_y != nullptr;

  double ex[2][2], rs[2][2];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const long c = col0 + 8 * h + w;
      const bool ok = c < a.c;
      ex[h][w] = (has_ex && ok) ? a.elev_x[c] : 0.0;
      rs[h][w] = (FUSED && ok) ? a.rootS[c] : 0.0;
    }
  double c_e[2][2], c_e2[2][2], c_cf[2][2], c_mx[2][2], c_my[2][2];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      c_e[h][w] = c_e2[h][w] = c_cf[h][w] = 0.0;
      c_mx[h][w] = c_my[h][w] = -INFINITY;
    }
  double s_ab = 0.0, s_ct = 0.0, s_sv = 0.0, rs_cta = 0.0;

  auto load_rows = [&](int buf, int tt) {
    constexpr int CPR = P16 / 2;
    for (int c = tid; c < MET_ROWS * CPR; c += MET_THREADS) {
      const int row = c / CPR, kc = c - row * CPR;
      cp_async16(sA + (buf * MET_ROWS + row) * Cfg::LDA + 2 * kc, a.M + (long)(tt * MET_ROWS + row) * a.ldm + 2 * kc);
    }
    if (tid < MET_ROWS / 2) cp_async16(sV + buf * MET_ROWS + 2 * tid, a.var + (long)tt * MET_ROWS + 2 * tid);
  };
  if (FUSED && t_begin < t_end) {
    for (int c = tid; c < P16 * 64; c += MET_THREADS) {
      const int kr = c >> 6, mc = c & 63;
      cp_async16(sE + kr * Cfg::LDE + 2 * mc, a.E + (long)kr * a.lde + (long)tj * 128 + 2 * mc);
    }
    load_rows(0, t_begin);
    cp_async_commit();
    // sum of rootS over this CTA's valid columns (fixed order): warp partial, then 8 warps
    double v = (rs[0][0] + rs[0][1]) + (rs[1][0] + rs[1][1]);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (lane == 0) sMisc[warp] = v;
    __syncthreads();
    rs_cta = ((sMisc[0] + sMisc[1]) + (sMisc[2] + sMisc[3])) + ((sMisc[4] + sMisc[5]) + (sMisc[6] + sMisc[7]));
  }
  // prediction offset: the bias with the depth conversion's elevation folded in
  double b0[2][2];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const long c = col0 + 8 * h + w;
      b0[h][w] = FUSED ? a.bias[c] - ((has_ey && c < a.c) ? a.elev_y[c] : 0.0) : ((has_ey && c < a.c) ? a.elev_y[c] : 0.0);
    }

  for (int tt = t_begin; tt < t_end; tt++) {
    const int buf = (tt - t_begin) & 1;
    const int row_base = tt * MET_ROWS;
    const bool full = cols_full && row_base + MET_ROWS <= a.t_rows;
    // truth for this tile: issued first, consumed after the DMMA loop
    double xv[4][2][2];
    if (a.X == nullptr) {
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 2; h++) xv[f][h][0] = xv[f][h][1] = 0.0;
    } else if (full && a.x_vec) {
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double2 v = __ldg(reinterpret_cast<const double2*>(a.X + (long)(row_base + 8 * f + g) * a.ldx + col0 + 8 * h));
          xv[f][h][0] = v.x, xv[f][h][1] = v.y;
        }
    } else {
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const int row = row_base + 8 * f + g;
        const bool rok = row < a.t_rows;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int w = 0; w < 2; w++) {
            const long c = col0 + 8 * h + w;
            xv[f][h][w] = (rok && c < a.c) ? __ldg(a.X + (long)row * a.ldx + c) : 0.0;
          }
      }
    }
    double acc[4][2][2], cf[FUSED ? 1 : 4][2][2];
    if (FUSED) {
      cp_async_wait<0>();
      __syncthreads();
      if (tt + 1 < t_end) load_rows(buf ^ 1, tt + 1);
      cp_async_commit();
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int f = 0; f < 4; f++) acc[f][h][0] = b0[h][0], acc[f][h][1] = b0[h][1];
      const double* a0 = sA + buf * MET_ROWS * Cfg::LDA;
#pragma unroll
      for (int ks = 0; ks < P16 / 4; ks++) {
        double av[4], bv[2];
#pragma unroll
        for (int f = 0; f < 4; f++) av[f] = a0[(8 * f + g) * Cfg::LDA + 4 * ks + q];
#pragma unroll
        for (int h = 0; h < 2; h++) bv[h] = sE[(4 * ks + q) * Cfg::LDE + wn + 8 * h + g];
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
          for (int h = 0; h < 2; h++) dmma(acc[f][h][0], acc[f][h][1], av[f], bv[h]);
      }
    } else {
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const int row = row_base + 8 * f + g;
        const bool rok = row < a.t_rows;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int w = 0; w < 2; w++) {
            const long c = col0 + 8 * h + w;
            const bool ok = rok && c < a.c;
            acc[f][h][w] = (ok ? __ldg(a.Y + (long)row * a.ldy + c) : 0.0) - b0[h][w];
            cf[FUSED ? 0 : f][h][w] = (ok && a.CONF) ? __ldg(a.CONF + (long)row * a.ldconf + c) : 0.0;
          }
      }
    }
    // ---- elementwise + reductions, one row at a time ----
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int rl = 8 * f + g;
      const bool rok = full || row_base + rl < a.t_rows;
      double r_e = 0.0, r_e2 = 0.0, r_cf = 0.0;
      if (FUSED) {
        const double sv = rok ? sqrt(sV[buf * MET_ROWS + rl]) : 0.0;
        s_sv += sv;
        if (warp == 0 && q == 0) a.row_part[((long)2 * a.t_tiles * MET_ROWS + row_base + rl) * a.n_ctile + tj] = sv * rs_cta;
      }
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int w = 0; w < 2; w++) {
          double x = xv[f][h][w], y = acc[f][h][w];
          if (has_ex) x = fmax(x - ex[h][w], 0.0);
          if (has_ey) y = fmax(y, 0.0);
          double e = x - y;
          if (full) {
            c_mx[h][w] = fmax(c_mx[h][w], x);
            c_my[h][w] = fmax(c_my[h][w], y);
            s_ct += fabs(e) <= a.v_tol ? 1.0 : 0.0;
          } else {
            const bool ok = rok && (col0 + 8 * h + w < a.c);
            e = ok ? e : 0.0;
            if (ok) {
              c_mx[h][w] = fmax(c_mx[h][w], x);
              c_my[h][w] = fmax(c_my[h][w], y);
              s_ct += fabs(e) <= a.v_tol ? 1.0 : 0.0;
            }
          }
          c_e[h][w] += e;
          c_e2[h][w] = fma(e, e, c_e2[h][w]);
          r_e += e;
          r_e2 = fma(e, e, r_e2);
          s_ab += fabs(e);
          if (!FUSED) {
            const double cfv = cf[FUSED ? 0 : f][h][w];  // already zero where masked
            c_cf[h][w] += cfv;
            r_cf += cfv;
          }
        }
      const int slot = warp * 4 + q;
      sRed[(0 * 32 + slot) * MET_RED_LD + rl] = r_e;
      sRed[(1 * 32 + slot) * MET_RED_LD + rl] = r_e2;
      if (!FUSED) sRed[(2 * 32 + slot) * MET_RED_LD + rl] = r_cf;
    }
    __syncthreads();
    if (tid < NQ * MET_ROWS) {
      const int qq = tid / MET_ROWS, row = tid - qq * MET_ROWS;
      const double* r = sRed + (long)qq * 32 * MET_RED_LD + row;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k++) s += r[k * MET_RED_LD];
      a.row_part[((long)qq * a.t_tiles * MET_ROWS + row_base + row) * a.n_ctile + tj] = s;
    }
    __syncthreads();
  }

  // ---- per-cell partials of this CTA's row range: reduce over g (8 lanes); every warp owns its columns ----
  if (FUSED) {
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int w = 0; w < 2; w++) c_cf[h][w] = rs[h][w] * s_sv;
  }
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      double v0 = c_e[h][w], v1 = c_e2[h][w], v2 = c_cf[h][w], v3 = c_mx[h][w], v4 = c_my[h][w];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        v3 = fmax(v3, __shfl_xor_sync(0xffffffffu, v3, o));
        v4 = fmax(v4, __shfl_xor_sync(0xffffffffu, v4, o));
      }
      if (g == 0) {
        double* o = a.cell_part + (long)blockIdx.y * MET_CELLQ * a.c_pad + col0 + 8 * h + w;
        o[0] = v0, o[a.c_pad] = v1, o[2 * a.c_pad] = v2, o[3 * a.c_pad] = v3, o[4 * a.c_pad] = v4;
      }
    }
  // ---- CTA scalars: sum |e| and the match count ----
  s_ab = warp_sum(s_ab);
  s_ct = warp_sum(s_ct);
  __syncthreads();
  if (lane == 0) sMisc[warp] = s_ab, sMisc[8 + warp] = s_ct;
  __syncthreads();
  if (tid < 2) {
    const double* r = sMisc + 8 * tid;
    a.cta_part[((long)blockIdx.y * a.n_ctile + tj) * MET_CTAQ + tid] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  }
}

// ---- fused consumer with ONE VARIANCE PER MODE (the reference's default per-column models) ---------------------------------
// As metrics_stream_kernel<P16, true>, but the confidence is conf[t][c] = sqrt(sum_p V[t][p] E[p][c]^2)
// (gpras/preprocess.py:1081-1094 followed by pipeline.py:262-263's sqrt): a second DMMA product over the same E tile, squared
// as the B fragments are read.  It is no longer separable in (t, c), so its row and cell sums are accumulated per element
// like the error's.  Same outputs and fixed-shape reductions as the other variants.
template <int P16>
struct MetGenCfg {
  static constexpr int LDE = 128 + 4;
  static constexpr int LDA = P16 + 4;
  static constexpr int RED_DOUBLES = 3 * 32 * MET_RED_LD + 16;
  static constexpr int SMEM_DOUBLES = P16 * LDE + 4 * MET_ROWS * LDA + RED_DOUBLES;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * (int)sizeof(double);
};

template <int P16>
__global__ void __launch_bounds__(MET_THREADS, 1) metrics_general_kernel(const MetricsArgs a) {
  using Cfg = MetGenCfg<P16>;
  extern __shared__ __align__(16) double smem[];
  double* sE = smem;
  double* sA = sE + P16 * Cfg::LDE;                 // [2][MET_ROWS][LDA] means
  double* sW = sA + 2 * MET_ROWS * Cfg::LDA;        // [2][MET_ROWS][LDA] variances
  double* sRed = sW + 2 * MET_ROWS * Cfg::LDA;      // [3][32 slots][MET_RED_LD]
  double* sMisc = sRed + 3 * 32 * MET_RED_LD;       // [16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wn = warp * 16;
  const int tj = blockIdx.x;
  const int t_begin = blockIdx.y * a.tiles_per_cta;
  int t_end = t_begin + a.tiles_per_cta;
  if (t_end > a.t_tiles) t_end = a.t_tiles;
  const long col0 = (long)tj * 128 + wn + 2 * q;
  const bool cols_full = (long)tj * 128 + 128 <= a.c;
  const bool has_ex = a.elev_x != nullptr, has_ey = a.elev_y != nullptr;

  double ex[2][2], b0[2][2];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const long c = col0 + 8 * h + w;
      const bool ok = c < a.c;
      ex[h][w] = (has_ex && ok) ? a.elev_x[c] : 0.0;
      b0[h][w] = a.bias[c] - ((has_ey && ok) ? a.elev_y[c] : 0.0);
    }
  double c_e[2][2], c_e2[2][2], c_cf[2][2], c_mx[2][2], c_my[2][2];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      c_e[h][w] = c_e2[h][w] = c_cf[h][w] = 0.0;
      c_mx[h][w] = c_my[h][w] = -INFINITY;
    }
  double s_ab = 0.0, s_ct = 0.0;

  auto load_rows = [&](int buf, int tt) {
    constexpr int CPR = P16 / 2;
    for (int c = tid; c < MET_ROWS * CPR; c += MET_THREADS) {
      const int row = c / CPR, kc = c - row * CPR;
      const long src = (long)(tt * MET_ROWS + row) * a.ldm + 2 * kc;
      cp_async16(sA + (buf * MET_ROWS + row) * Cfg::LDA + 2 * kc, a.M + src);
      cp_async16(sW + (buf * MET_ROWS + row) * Cfg::LDA + 2 * kc, a.Vm + src);
    }
  };
  if (t_begin < t_end) {
    for (int c = tid; c < P16 * 64; c += MET_THREADS) {
      const int kr = c >> 6, mc = c & 63;
      cp_async16(sE + kr * Cfg::LDE + 2 * mc, a.E + (long)kr * a.lde + (long)tj * 128 + 2 * mc);
    }
    load_rows(0, t_begin);
    cp_async_commit();
  }

  for (int tt = t_begin; tt < t_end; tt++) {
    const int buf = (tt - t_begin) & 1;
    const int row_base = tt * MET_ROWS;
    const bool full = cols_full && row_base + MET_ROWS <= a.t_rows;
    double xv[4][2][2];
    if (a.X == nullptr) {
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 2; h++) xv[f][h][0] = xv[f][h][1] = 0.0;
    } else if (full && a.x_vec) {
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double2 v = __ldg(reinterpret_cast<const double2*>(a.X + (long)(row_base + 8 * f + g) * a.ldx + col0 + 8 * h));
          xv[f][h][0] = v.x, xv[f][h][1] = v.y;
        }
    } else {
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const int row = row_base + 8 * f + g;
        const bool rok = row < a.t_rows;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int w = 0; w < 2; w++) {
            const long c = col0 + 8 * h + w;
            xv[f][h][w] = (rok && c < a.c) ? __ldg(a.X + (long)row * a.ldx + c) : 0.0;
          }
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (tt + 1 < t_end) load_rows(buf ^ 1, tt + 1);
    cp_async_commit();
    double acc[4][2][2], vac[4][2][2];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int f = 0; f < 4; f++) {
        acc[f][h][0] = b0[h][0], acc[f][h][1] = b0[h][1];
        vac[f][h][0] = vac[f][h][1] = 0.0;
      }
    const double* a0 = sA + buf * MET_ROWS * Cfg::LDA;
    const double* w0 = sW + buf * MET_ROWS * Cfg::LDA;
#pragma unroll
    for (int ks = 0; ks < P16 / 4; ks++) {
      double av[4], wv[4], bv[2];
#pragma unroll
      for (int f = 0; f < 4; f++) {
        av[f] = a0[(8 * f + g) * Cfg::LDA + 4 * ks + q];
        wv[f] = w0[(8 * f + g) * Cfg::LDA + 4 * ks + q];
      }
#pragma unroll
      for (int h = 0; h < 2; h++) bv[h] = sE[(4 * ks + q) * Cfg::LDE + wn + 8 * h + g];
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          dmma(acc[f][h][0], acc[f][h][1], av[f], bv[h]);
          dmma(vac[f][h][0], vac[f][h][1], wv[f], bv[h] * bv[h]);
        }
    }
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int rl = 8 * f + g;
      const bool rok = full || row_base + rl < a.t_rows;
      double r_e = 0.0, r_e2 = 0.0, r_cf = 0.0;
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int w = 0; w < 2; w++) {
          double x = xv[f][h][w], y = acc[f][h][w];
          if (has_ex) x = fmax(x - ex[h][w], 0.0);
          if (has_ey) y = fmax(y, 0.0);
          const bool ok = full || (rok && (col0 + 8 * h + w < a.c));
          const double e = ok ? x - y : 0.0;
          const double cf = ok ? sqrt(fmax(vac[f][h][w], 0.0)) : 0.0;
          if (ok) {
            c_mx[h][w] = fmax(c_mx[h][w], x);
            c_my[h][w] = fmax(c_my[h][w], y);
            s_ct += fabs(e) <= a.v_tol ? 1.0 : 0.0;
          }
          c_e[h][w] += e;
          c_e2[h][w] = fma(e, e, c_e2[h][w]);
          c_cf[h][w] += cf;
          r_e += e;
          r_e2 = fma(e, e, r_e2);
          r_cf += cf;
          s_ab += fabs(e);
        }
      const int slot = warp * 4 + q;
      sRed[(0 * 32 + slot) * MET_RED_LD + rl] = r_e;
      sRed[(1 * 32 + slot) * MET_RED_LD + rl] = r_e2;
      sRed[(2 * 32 + slot) * MET_RED_LD + rl] = r_cf;
    }
    __syncthreads();
    if (tid < 3 * MET_ROWS) {
      const int qq = tid / MET_ROWS, row = tid - qq * MET_ROWS;
      const double* r = sRed + (long)qq * 32 * MET_RED_LD + row;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k++) s += r[k * MET_RED_LD];
      a.row_part[((long)qq * a.t_tiles * MET_ROWS + row_base + row) * a.n_ctile + tj] = s;
    }
    __syncthreads();
  }

#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int w = 0; w < 2; w++) {
      double v0 = c_e[h][w], v1 = c_e2[h][w], v2 = c_cf[h][w], v3 = c_mx[h][w], v4 = c_my[h][w];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        v3 = fmax(v3, __shfl_xor_sync(0xffffffffu, v3, o));
        v4 = fmax(v4, __shfl_xor_sync(0xffffffffu, v4, o));
      }
      if (g == 0) {
        double* o = a.cell_part + (long)blockIdx.y * MET_CELLQ * a.c_pad + col0 + 8 * h + w;
        o[0] = v0, o[a.c_pad] = v1, o[2 * a.c_pad] = v2, o[3 * a.c_pad] = v3, o[4 * a.c_pad] = v4;
      }
    }
  s_ab = warp_sum(s_ab);
  s_ct = warp_sum(s_ct);
  __syncthreads();
  if (lane == 0) sMisc[warp] = s_ab, sMisc[8 + warp] = s_ct;
  __syncthreads();
  if (tid < 2) {
    const double* r = sMisc + 8 * tid;
    a.cta_part[((long)blockIdx.y * a.n_ctile + tj) * MET_CTAQ + tid] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  }
}

// ---- resident (T x C) arrays, streamed by the TMA unit ------------------------------------------------------------------
// The !FUSED case above reads truth, prediction and confidence with per-thread loads and no prefetch across row tiles
// (measured: 3.3 TB/s, half of HBM).  Here one thread keeps MPT_STAGES stages of three 32 x 128 boxes in flight
// (cp.async.bulk.tensor -> shared memory, completion on an mbarrier) and the arithmetic runs out of shared memory.  Warp w owns
// rows w, w + 8, w + 16, w + 24 of a row tile and all 128 columns of the CTA's slab (lane l: columns 2l, 2l+1, 64+2l, 65+2l), so
// a row sum is one warp reduction and the per-cell partials stay in registers until the CTA ends.  Same outputs, same
// fixed-shape reductions (bitwise repeatable) as metrics_stream_kernel<.., false>.
constexpr int MPT_STAGES = 2;
constexpr int MPT_BOX_DOUBLES = MET_ROWS * 128;
constexpr int MPT_SMEM_BYTES = MPT_STAGES * 3 * MPT_BOX_DOUBLES * (int)sizeof(double) + 128;

static __global__ void __launch_bounds__(MET_THREADS, 1)
metrics_plain_tma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap,
                         const __grid_constant__ CUtensorMap cmap, const MetricsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[MPT_STAGES];
  double* tiles = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tj = blockIdx.x;
  const int t_begin = blockIdx.y * a.tiles_per_cta;
  int t_end = t_begin + a.tiles_per_cta;
  if (t_end > a.t_tiles) t_end = a.t_tiles;
  const int n_it = t_end > t_begin ? t_end - t_begin : 0;
  const bool has_x = a.X != nullptr, has_cf = a.CONF != nullptr;
  const bool has_ex = a.elev_x != nullptr, has_ey = a.elev_y != nullptr;
  const uint32_t stage_bytes = (uint32_t)((1 + (has_x ? 1 : 0) + (has_cf ? 1 : 0)) * MPT_BOX_DOUBLES * sizeof(double));
  if (tid == 0) {
    for (int s = 0; s < MPT_STAGES; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int s, int tt) {
    double* base = tiles + s * 3 * MPT_BOX_DOUBLES;
    mbar_expect_tx(&full[s], stage_bytes);
    tma_load_2d(base, &ymap, tj * 128, tt * MET_ROWS, &full[s]);
    if (has_x) tma_load_2d(base + MPT_BOX_DOUBLES, &xmap, tj * 128, tt * MET_ROWS, &full[s]);
    if (has_cf) tma_load_2d(base + 2 * MPT_BOX_DOUBLES, &cmap, tj * 128, tt * MET_ROWS, &full[s]);
  };
  if (tid == 0)
    for (int s = 0; s < MPT_STAGES && s < n_it; s++) issue(s, t_begin + s);

  // this lane's four columns: 2l, 2l+1, 64+2l, 65+2l of the slab
  long col[4];
  bool cok[4];
  double ex[4], ey[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    col[j] = (long)tj * 128 + 64 * (j >> 1) + 2 * lane + (j & 1);
    cok[j] = col[j] < a.c;
    ex[j] = (has_ex && cok[j]) ? a.elev_x[col[j]] : 0.0;
    ey[j] = (has_ey && cok[j]) ? a.elev_y[col[j]] : 0.0;
  }
  double c_e[4], c_e2[4], c_cf[4], c_mx[4], c_my[4];
#pragma unroll
  for (int j = 0; j < 4; j++) c_e[j] = c_e2[j] = c_cf[j] = 0.0, c_mx[j] = c_my[j] = -INFINITY;
  double s_ab = 0.0, s_ct = 0.0;

  for (int it = 0; it < n_it; it++) {
    const int s = it % MPT_STAGES, tt = t_begin + it;
    const int row_base = tt * MET_ROWS;
    mbar_wait(&full[s], (it / MPT_STAGES) & 1);
    const double* ty = tiles + s * 3 * MPT_BOX_DOUBLES;
    const double* tx = ty + MPT_BOX_DOUBLES;
    const double* tc = ty + 2 * MPT_BOX_DOUBLES;
#pragma unroll
    for (int rr = 0; rr < MET_ROWS / 8; rr++) {
      const int rl = warp + 8 * rr;
      const bool rok = row_base + rl < a.t_rows;
      double xv[4], yv[4], cv[4];
#pragma unroll
      for (int hlf = 0; hlf < 2; hlf++) {
        const int o = rl * 128 + 64 * hlf + 2 * lane;
        const double2 y2 = *reinterpret_cast<const double2*>(ty + o);
        yv[2 * hlf] = y2.x, yv[2 * hlf + 1] = y2.y;
        if (has_x) {
          const double2 x2 = *reinterpret_cast<const double2*>(tx + o);
          xv[2 * hlf] = x2.x, xv[2 * hlf + 1] = x2.y;
        } else {
          xv[2 * hlf] = xv[2 * hlf + 1] = 0.0;
        }
        if (has_cf) {
          const double2 c2 = *reinterpret_cast<const double2*>(tc + o);
          cv[2 * hlf] = c2.x, cv[2 * hlf + 1] = c2.y;
        } else {
          cv[2 * hlf] = cv[2 * hlf + 1] = 0.0;
        }
      }
      double r_e = 0.0, r_e2 = 0.0, r_cf = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const bool ok = rok && cok[j];
        double x = xv[j], y = yv[j] - ey[j];
        if (has_ex) x = fmax(x - ex[j], 0.0);
        if (has_ey) y = fmax(y, 0.0);
        const double e = ok ? x - y : 0.0;
        const double cf = ok ? cv[j] : 0.0;
        if (ok) {
          c_mx[j] = fmax(c_mx[j], x);
          c_my[j] = fmax(c_my[j], y);
          s_ct += fabs(e) <= a.v_tol ? 1.0 : 0.0;
        }
        c_e[j] += e;
        c_e2[j] = fma(e, e, c_e2[j]);
        c_cf[j] += cf;
        r_e += e;
        r_e2 = fma(e, e, r_e2);
        r_cf += cf;
        s_ab += fabs(e);
      }
      r_e = warp_sum(r_e), r_e2 = warp_sum(r_e2), r_cf = warp_sum(r_cf);
      if (lane == 0) {
        const long rix = (long)row_base + rl;
        a.row_part[((long)0 * a.t_tiles * MET_ROWS + rix) * a.n_ctile + tj] = r_e;
        a.row_part[((long)1 * a.t_tiles * MET_ROWS + rix) * a.n_ctile + tj] = r_e2;
        a.row_part[((long)2 * a.t_tiles * MET_ROWS + rix) * a.n_ctile + tj] = r_cf;
      }
    }
    __syncthreads();  // every warp is done with this stage: refill it
    if (tid == 0 && it + MPT_STAGES < n_it) issue(s, tt + MPT_STAGES);
  }

  // ---- per-cell partials: combine the eight warps (fixed order) through shared memory ----
  double* red = tiles;  // [8 warps][MET_CELLQ][128]
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int cl = 64 * (j >> 1) + 2 * lane + (j & 1);
    red[(warp * MET_CELLQ + 0) * 128 + cl] = c_e[j];
    red[(warp * MET_CELLQ + 1) * 128 + cl] = c_e2[j];
    red[(warp * MET_CELLQ + 2) * 128 + cl] = c_cf[j];
    red[(warp * MET_CELLQ + 3) * 128 + cl] = c_mx[j];
    red[(warp * MET_CELLQ + 4) * 128 + cl] = c_my[j];
  }
  s_ab = warp_sum(s_ab);
  s_ct = warp_sum(s_ct);
  double* misc = tiles + 8 * MET_CELLQ * 128;
  if (lane == 0) misc[warp] = s_ab, misc[8 + warp] = s_ct;
  __syncthreads();
  if (tid < 128) {
    double v[MET_CELLQ] = {0.0, 0.0, 0.0, -INFINITY, -INFINITY};
#pragma unroll
    for (int w = 0; w < 8; w++) {
      v[0] += red[(w * MET_CELLQ + 0) * 128 + tid];
      v[1] += red[(w * MET_CELLQ + 1) * 128 + tid];
      v[2] += red[(w * MET_CELLQ + 2) * 128 + tid];
      v[3] = fmax(v[3], red[(w * MET_CELLQ + 3) * 128 + tid]);
      v[4] = fmax(v[4], red[(w * MET_CELLQ + 4) * 128 + tid]);
    }
    double* o = a.cell_part + (long)blockIdx.y * MET_CELLQ * a.c_pad + (long)tj * 128 + tid;
#pragma unroll
    for (int qq = 0; qq < MET_CELLQ; qq++) o[(long)qq * a.c_pad] = v[qq];
  }
  if (tid < 2) {
    const double* r = misc + 8 * tid;
    a.cta_part[((long)blockIdx.y * a.n_ctile + tj) * MET_CTAQ + tid] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  }
}

// state[q][c] (+)= fold over the row-range partials, fixed order.  first != 0: the state is (re)initialised.
static __global__ void metrics_fold_cells_kernel(const double* __restrict__ part, int nsplit, long c_pad, double* __restrict__ state,
                                                 int first) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_pad) return;
#pragma unroll
  for (int qq = 0; qq < MET_CELLQ; qq++) {
    double s = first ? (qq < 3 ? 0.0 : -INFINITY) : state[(long)qq * c_pad + c];
    for (int z = 0; z < nsplit; z++) {
      const double v = part[((long)z * MET_CELLQ + qq) * c_pad + c];
      s = qq < 3 ? s + v : fmax(s, v);
    }
    state[(long)qq * c_pad + c] = s;
  }
}

// rows_out[q][t0 + t] = sum over column tiles of row_part[q][t][:]: one warp per (q, t), fixed order.
static __global__ void metrics_fold_rows_kernel(const double* __restrict__ part, int t_rows, long t_pitch, int n_ctile,
                                                double* __restrict__ rows_out, long t_cap, long t0) {
  const int lane = threadIdx.x & 31;
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long)MET_ROWQ * t_rows) return;
  const int qq = (int)(w / t_rows);
  const long t = w - (long)qq * t_rows;
  const double* p = part + ((long)qq * t_pitch + t) * n_ctile;
  double s = 0.0;
  for (int j = lane; j < n_ctile; j += 32) s += p[j];
  s = warp_sum(s);
  if (lane == 0) rows_out[(long)qq * t_cap + t0 + t] = s;
}

// scal2[0..1] (+)= sum over the CTA partials of one block (fixed order, single warp).
static __global__ void metrics_fold_cta_kernel(const double* __restrict__ part, int count, double* __restrict__ scal2, int first) {
  const int lane = threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  for (int i = lane; i < count; i += 32) s0 += part[(long)i * MET_CTAQ], s1 += part[(long)i * MET_CTAQ + 1];
  s0 = warp_sum(s0), s1 = warp_sum(s1);
  if (lane == 0) {
    scal2[0] = (first ? 0.0 : scal2[0]) + s0;
    scal2[1] = (first ? 0.0 : scal2[1]) + s1;
  }
}

// Scalars of one event from the folded state (single CTA, fixed order).
//   out[0..4]  = sum_c cell{e, e^2, conf},  sum |e|, count(|e| <= v_tol) (running scalars scal2)
//   out[5..8]  = peaks: sum d, sum d^2, sum xm, sum (xm - mean xm)^2      with d = max_t x - max_t y
//   out[9..11] = counts at depth_threshold: hits (x>=thr & y>=thr), misses (x>=thr & y<thr), false alarms
//   out[12..14]= the same three counts at threshold 0 (f2 / f3 defaults)
constexpr int MET_SCALARS = 15;
static __global__ void __launch_bounds__(1024) metrics_finalize_kernel(const double* __restrict__ state, long c_pad, int c,
                                                                       const double* __restrict__ scal2, double thr,
                                                                       double* __restrict__ out) {
  __shared__ double red[32];
  __shared__ double bc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_sum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double s = red[lane];
      s = warp_sum(s);
      if (lane == 0) bc = s;
    }
    __syncthreads();
    return bc;
  };
  double acc[MET_SCALARS];
#pragma unroll
  for (int i = 0; i < MET_SCALARS; i++) acc[i] = 0.0;
  for (long j = tid; j < c; j += 1024) {
    acc[0] += state[j], acc[1] += state[c_pad + j], acc[2] += state[2 * c_pad + j];
    const double xm = state[3 * c_pad + j], ym = state[4 * c_pad + j], d = xm - ym;
    acc[5] += d, acc[6] += d * d, acc[7] += xm;
    const bool hx = xm >= thr, hy = ym >= thr, zx = xm >= 0.0, zy = ym >= 0.0;
    acc[9] += (hx && hy), acc[10] += (hx && !hy), acc[11] += (!hx && hy);
    acc[12] += (zx && zy), acc[13] += (zx && !zy), acc[14] += (!zx && zy);
  }
#pragma unroll
  for (int i = 0; i < MET_SCALARS; i++)
    if (i != 8 && i != 3 && i != 4) acc[i] = block_sum(acc[i]);
  acc[3] = scal2[0], acc[4] = scal2[1];
  const double mean_xm = acc[7] / (double)c;
  double v = 0.0;
  for (long j = tid; j < c; j += 1024) {
    const double u = state[3 * c_pad + j] - mean_xm;
    v += u * u;
  }
  acc[8] = block_sum(v);
  if (tid == 0)
    for (int i = 0; i < MET_SCALARS; i++) out[i] = acc[i];
}

// fi_aoi_toi with a time tolerance (metrics.py:203-212) on device-resident (t x c) arrays: count of matching entries.
// match[t][c] = OR_{i=0..t_tol, t+i<T} ( |y[t]-x[t+i]| <= v  or  |x[t]-y[t+i]| <= v )
static __global__ void metrics_fidelity_kernel(const double* __restrict__ X, long ldx, const double* __restrict__ Y, long ldy, int t,
                                               int c, int t_tol, double v_tol, double* __restrict__ part) {
  __shared__ double red[8];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  double m = 0.0;
  if (col < c) {
    const double x0 = X[(long)row * ldx + col], y0 = Y[(long)row * ldy + col];
    bool ok = fabs(y0 - x0) <= v_tol;
    for (int i = 1; i <= t_tol && row + i < t && !ok; i++)
      ok = fabs(y0 - X[(long)(row + i) * ldx + col]) <= v_tol || fabs(x0 - Y[(long)(row + i) * ldy + col]) <= v_tol;
    m = ok ? 1.0 : 0.0;
  }
  m = warp_sum(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
    part[(long)row * gridDim.x + blockIdx.x] = s;
  }
}

}  // namespace gpras
