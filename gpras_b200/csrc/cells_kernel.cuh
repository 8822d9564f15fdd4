// Modes -> mesh-cells expansion (PreProcessor.reverse_transform, gpras/preprocess.py:1052-1094) as one streaming
// kernel: the step that turns (T x P) predictions into "cell-depths".
//
//   cell_mean[t][c] = sum_p M[t][p] E[p][c] + bias[c]          (DMMA, k = P <= 64)
//   cell_var [t][c] = var[t] * S[c],   S[c] = sum_p E[p][c]^2  (all P columns share theta, so the variance is rank one)
//
// The output is 16 bytes per cell-depth against ~2P flops: HBM-write bound.  A CTA owns one 128-cell column tile,
// keeps its E tile in shared memory and its bias / S fragments in registers, and streams row tiles of 64 events
// through a double-buffered cp.async ring, so E is read once per CTA and the only steady-state traffic is the
// (tiny) mode-space input and the cell-space output.  Two CTAs per SM overlap one tile's stores with the next tile's
// DMMAs (row tiles of 64 keep the accumulators at 64 registers, so two CTAs fit without spills).  Rows wrap modulo
// `ring_rows` when the caller does not keep the T x C result.
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int CELLS_THREADS = 256;
constexpr int CELLS_ROWS = 64;  // events per row tile

template <int P16>
struct CellsCfg {
  static constexpr int LDE = 128 + 4;   // k-major E tile [P16][132]
  static constexpr int LDA = P16 + 4;   // row-major mode tile [128][P16 + 4]
  static constexpr int SMEM_DOUBLES = P16 * LDE + 2 * CELLS_ROWS * LDA + 2 * CELLS_ROWS;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * (int)sizeof(double);
};

template <int P16>
__global__ void __launch_bounds__(CELLS_THREADS, 2)
cells_kernel(const double* __restrict__ M, long ldm, const double* __restrict__ var, const double* __restrict__ E, long lde,
             const double* __restrict__ bias, const double* __restrict__ S, double* __restrict__ out_m,
             double* __restrict__ out_v, long ldo, int t_tiles, int tiles_per_cta, int ring_rows) {
  using Cfg = CellsCfg<P16>;
  extern __shared__ __align__(16) double smem[];
  double* sE = smem;                        // [P16][LDE]
  double* sA = sE + P16 * Cfg::LDE;               // [2][CELLS_ROWS][LDA]
  double* sV = sA + 2 * CELLS_ROWS * Cfg::LDA;    // [2][CELLS_ROWS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
  const int tj = blockIdx.x;
  const int t_begin = blockIdx.y * tiles_per_cta;
  int t_end = t_begin + tiles_per_cta;
  if (t_end > t_tiles) t_end = t_tiles;
  if (t_begin >= t_end) return;

  auto load_rows = [&](int buf, int tt) {
    // CELLS_ROWS rows x P16 doubles, 16-byte chunks
    constexpr int CPR = P16 / 2;
    for (int c = tid; c < CELLS_ROWS * CPR; c += CELLS_THREADS) {
      const int row = c / CPR, kc = c - row * CPR;
      cp_async16(sA + (buf * CELLS_ROWS + row) * Cfg::LDA + 2 * kc, M + (long)(tt * CELLS_ROWS + row) * ldm + 2 * kc);
    }
    if (tid < CELLS_ROWS / 2) cp_async16(sV + buf * CELLS_ROWS + 2 * tid, var + (long)tt * CELLS_ROWS + 2 * tid);
  };
  // E tile + first row tile
  for (int c = tid; c < P16 * 64; c += CELLS_THREADS) {
    const int kr = c >> 6, mc = c & 63;
    cp_async16(sE + kr * Cfg::LDE + 2 * mc, E + (long)kr * lde + (long)tj * 128 + 2 * mc);
  }
  load_rows(0, t_begin);
  cp_async_commit();
  double bs[4][2], ss[4][2];
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const long c = (long)tj * 128 + wn + 8 * h + 2 * q;
    bs[h][0] = bias[c], bs[h][1] = bias[c + 1];
    ss[h][0] = S[c], ss[h][1] = S[c + 1];
  }
  for (int tt = t_begin; tt < t_end; tt++) {
    const int buf = (tt - t_begin) & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (tt + 1 < t_end) load_rows(buf ^ 1, tt + 1);
    cp_async_commit();
    double acc[4][4][2];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int h = 0; h < 4; h++) acc[f][h][0] = bs[h][0], acc[f][h][1] = bs[h][1];
    const double* a0 = sA + buf * CELLS_ROWS * Cfg::LDA;
#pragma unroll
    for (int ks = 0; ks < P16 / 4; ks++) {
      double a[4], b[4];
#pragma unroll
      for (int f = 0; f < 4; f++) a[f] = a0[(wm + 8 * f + g) * Cfg::LDA + 4 * ks + q];
#pragma unroll
      for (int h = 0; h < 4; h++) b[h] = sE[(4 * ks + q) * Cfg::LDE + wn + 8 * h + g];
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 4; h++) dmma(acc[f][h][0], acc[f][h][1], a[f], b[h]);
    }
    // Stores are arranged so that every warp instruction writes whole 128-byte lines (4 rows x 128 B): lanes g and
    // g^1 swap one 16-byte chunk by shuffle, then the 8 lanes of a row pair cover one row's 128-byte half-row.
    // (Writing each line as two 64-byte halves from separate instructions doubled the DRAM write traffic.)
    const long row0 = ((long)tt * CELLS_ROWS) % ring_rows;
    const int godd = g & 1;
#pragma unroll
    for (int f = 0; f < 4; f++) {
#pragma unroll
      for (int w = 0; w < 2; w++) {
        double2 mine0 = make_double2(acc[f][2 * w][0], acc[f][2 * w][1]);
        double2 mine1 = make_double2(acc[f][2 * w + 1][0], acc[f][2 * w + 1][1]);
        double2 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, mine1.x, 4);
        recv.y = __shfl_xor_sync(0xffffffffu, mine1.y, 4);
#pragma unroll
        for (int sft = 0; sft < 2; sft++) {
          // sft = 0: the even row of the pair is written; sft = 1: the odd row
          const int rl = wm + 8 * f + (sft ? (g | 1) : (g & ~1));
          const bool own = (godd == sft);
          const int hh = 2 * w + (own ? 0 : 1);
          const double2 mv = own ? mine0 : recv;
          const double vr = sV[buf * CELLS_ROWS + rl];
          const long off = (row0 + rl) * ldo + (long)tj * 128 + wn + 8 * hh + 2 * q;
          *reinterpret_cast<double2*>(out_m + off) = mv;
          *reinterpret_cast<double2*>(out_v + off) = make_double2(vr * ss[hh][0], vr * ss[hh][1]);
        }
      }
    }
  }
}

// ---- general variance map: one hyperparameter set PER mode (the reference's default model family) ---------------------
//   cell_mean[t][c] = sum_p M[t][p] E[p][c] + bias[c]
//   cell_var [t][c] = sum_p V[t][p] E[p][c]^2                 (gpras/preprocess.py:1081-1094: var @ (diag(x_std) eofs / w)^2)
// Same streaming structure as cells_kernel; the variance is a second DMMA product over the same E tile, squared on the fly
// as the B fragments are read (no second map in memory), with the mode-space variances V (T x P) double-buffered next to the
// means.  The accumulators are reused: the mean tile is stored before the variance tile is formed, so the register budget
// (and two CTAs per SM at P <= 32) is that of the shared-theta kernel.
template <int P16>
struct CellsGenCfg {
  static constexpr int LDE = 128 + 4;
  static constexpr int LDA = P16 + 4;
  static constexpr int SMEM_DOUBLES = P16 * LDE + 4 * CELLS_ROWS * LDA;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * (int)sizeof(double);
};

template <int P16>
__global__ void __launch_bounds__(CELLS_THREADS, P16 <= 32 ? 2 : 1)
cells_general_kernel(const double* __restrict__ M, const double* __restrict__ V, long ldm, const double* __restrict__ E, long lde,
                     const double* __restrict__ bias, double* __restrict__ out_m, double* __restrict__ out_v, long ldo, int t_tiles,
                     int tiles_per_cta, int ring_rows) {
  using Cfg = CellsGenCfg<P16>;
  extern __shared__ __align__(16) double smem[];
  double* sE = smem;                               // [P16][LDE]
  double* sA = sE + P16 * Cfg::LDE;                // [2][CELLS_ROWS][LDA] means
  double* sW = sA + 2 * CELLS_ROWS * Cfg::LDA;     // [2][CELLS_ROWS][LDA] variances
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
  const int tj = blockIdx.x;
  const int t_begin = blockIdx.y * tiles_per_cta;
  int t_end = t_begin + tiles_per_cta;
  if (t_end > t_tiles) t_end = t_tiles;
  if (t_begin >= t_end) return;

  auto load_rows = [&](int buf, int tt) {
    constexpr int CPR = P16 / 2;
    for (int c = tid; c < CELLS_ROWS * CPR; c += CELLS_THREADS) {
      const int row = c / CPR, kc = c - row * CPR;
      const long src = (long)(tt * CELLS_ROWS + row) * ldm + 2 * kc;
      cp_async16(sA + (buf * CELLS_ROWS + row) * Cfg::LDA + 2 * kc, M + src);
      cp_async16(sW + (buf * CELLS_ROWS + row) * Cfg::LDA + 2 * kc, V + src);
    }
  };
  for (int c = tid; c < P16 * 64; c += CELLS_THREADS) {
    const int kr = c >> 6, mc = c & 63;
    cp_async16(sE + kr * Cfg::LDE + 2 * mc, E + (long)kr * lde + (long)tj * 128 + 2 * mc);
  }
  load_rows(0, t_begin);
  cp_async_commit();
  double bs[4][2];
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const long c = (long)tj * 128 + wn + 8 * h + 2 * q;
    bs[h][0] = bias[c], bs[h][1] = bias[c + 1];
  }
  for (int tt = t_begin; tt < t_end; tt++) {
    const int buf = (tt - t_begin) & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (tt + 1 < t_end) load_rows(buf ^ 1, tt + 1);
    cp_async_commit();
    const long row0 = ((long)tt * CELLS_ROWS) % ring_rows;
    const int godd = g & 1;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {  // 0: means -> out_m, 1: variances -> out_v
      double acc[4][4][2];
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 4; h++) acc[f][h][0] = pass ? 0.0 : bs[h][0], acc[f][h][1] = pass ? 0.0 : bs[h][1];
      const double* a0 = (pass ? sW : sA) + buf * CELLS_ROWS * Cfg::LDA;
#pragma unroll
      for (int ks = 0; ks < P16 / 4; ks++) {
        double a[4], b[4];
#pragma unroll
        for (int f = 0; f < 4; f++) a[f] = a0[(wm + 8 * f + g) * Cfg::LDA + 4 * ks + q];
#pragma unroll
        for (int h = 0; h < 4; h++) {
          const double e = sE[(4 * ks + q) * Cfg::LDE + wn + 8 * h + g];
          b[h] = pass ? e * e : e;
        }
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
          for (int h = 0; h < 4; h++) dmma(acc[f][h][0], acc[f][h][1], a[f], b[h]);
      }
      // whole 128-byte lines per warp instruction, as in cells_kernel: lanes g and g^1 swap one 16-byte chunk by shuffle
      double* out = pass ? out_v : out_m;
#pragma unroll
      for (int f = 0; f < 4; f++) {
#pragma unroll
        for (int w = 0; w < 2; w++) {
          const double2 mine0 = make_double2(acc[f][2 * w][0], acc[f][2 * w][1]);
          const double2 mine1 = make_double2(acc[f][2 * w + 1][0], acc[f][2 * w + 1][1]);
          double2 recv;
          recv.x = __shfl_xor_sync(0xffffffffu, mine1.x, 4);
          recv.y = __shfl_xor_sync(0xffffffffu, mine1.y, 4);
#pragma unroll
          for (int sft = 0; sft < 2; sft++) {
            const int rl = wm + 8 * f + (sft ? (g | 1) : (g & ~1));
            const bool own = (godd == sft);
            const int hh = 2 * w + (own ? 0 : 1);
            const double2 mv = own ? mine0 : recv;
            *reinterpret_cast<double2*>(out + (row0 + rl) * ldo + (long)tj * 128 + wn + 8 * hh + 2 * q) = mv;
          }
        }
      }
    }
  }
}

}  // namespace gpras
