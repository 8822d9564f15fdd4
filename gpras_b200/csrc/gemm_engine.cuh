// FP64 tile-GEMM engine on the DMMA pipe (sm_100a).
//
// One kernel template covers every dense contraction of the exact-GP path: the Cholesky panel and
// trailing updates, the triangular inverse (recursive doubling), K^-1 = W^T W, alpha = W^T (W Y), the
// predictive V = W K*^T with a fused column sum-of-squares, and the modes -> cells expansion.
//
//   C[i, j] (op)= alpha * sum_{k in [k_begin, k_end)} A(i, k) * B(k, j)
//
// * CTA tile 128 x 128, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA.8x8x4 accumulators.
// * Operand tiles stream global -> shared with 16-byte cp.async in a 4-stage ring, BK = 16.
// * A is either row-major A[i][k] or k-major A[k][i]; B either n-major B[j][k] or k-major B[k][j];
//   shared tiles are padded (+4 doubles) so every fragment LDS.64 is bank-conflict free.
// * Triangular structure is exploited at tile granularity by clipping the k range per tile
//   (k_begin / k_end modes) and by launching only the lower-triangular tiles of C (tri mode).
//   Operand tiles that straddle the diagonal must hold explicit zeros in their dead half.
// * All extents are multiples of the tile sizes: the host pads (identity on the diagonal).
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, GEMM_THREADS = 256;
constexpr int LD_RM = BK + 4;             // [128][20] row / n-major tile (k contiguous)
constexpr int LD_KM = BM + 4;             // [16][132] k-major tile (m or n contiguous)
constexpr int TILE_DOUBLES = 128 * LD_RM;  // 2560 >= 16 * 132
constexpr int STAGE_DOUBLES = 2 * TILE_DOUBLES;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);  // 160 KiB

enum KBegin { KB_ZERO = 0, KB_TI = 1, KB_TJ = 2 };   // k_begin = 0 | ti*BM | tj*BN
enum KEnd { KE_FULL = 0, KE_TI = 1, KE_TJ = 2 };     // k_end   = K | (ti+1)*BM | (tj+1)*BN
enum Epilogue { EPI_STORE = 0, EPI_COLSUMSQ = 1, EPI_BIAS = 2 };

struct GemmDesc {
  const double* A;
  const double* B;
  double* C;            // EPI_STORE / EPI_BIAS: output; EPI_COLSUMSQ: partials [m_tiles][ldc]
  const double* bias;   // EPI_BIAS: per-column bias (length n_tiles*BN)
  long lda, ldb, ldc;
  long batchA, batchB, batchC;  // element strides between batch entries (grid.y)
  int m_tiles, n_tiles;
  int K;                // multiple of BK
  int kb_mode, ke_mode;
  int tri;              // 1: only tiles with tj <= ti (m_tiles == n_tiles)
  int reverse;          // 1: heaviest-last orders are reversed (LPT scheduling)
  int epilogue;
  double alpha, beta;
};

template <bool KMAJOR>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g, long ld, int tid) {
  // g points at the tile origin: row-major -> (row0, k0); k-major -> (k0, col0).
  if (!KMAJOR) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int c = tid + GEMM_THREADS * q;
      int row = c >> 3, kc = c & 7;
      cp_async16(s + row * LD_RM + 2 * kc, g + (long)row * ld + 2 * kc);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int c = tid + GEMM_THREADS * q;
      int kr = c >> 6, mc = c & 63;
      cp_async16(s + kr * LD_KM + 2 * mc, g + (long)kr * ld + 2 * mc);
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tile_kernel(const GemmDesc d) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;

  // ---- tile coordinates ----
  int t = blockIdx.x;
  if (d.reverse) t = gridDim.x - 1 - t;
  int ti, tj;
  if (d.tri) {
    ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(ti + 1) * (ti + 2) / 2 <= t) ti++;
    while ((long)ti * (ti + 1) / 2 > t) ti--;
    tj = t - (int)((long)ti * (ti + 1) / 2);
  } else {
    ti = t / d.n_tiles;
    tj = t - ti * d.n_tiles;
  }
  const long by = blockIdx.y;
  const double* __restrict__ A = d.A + by * d.batchA;
  const double* __restrict__ B = d.B + by * d.batchB;
  double* __restrict__ C = d.C + by * d.batchC;

  int k_begin = d.kb_mode == KB_TI ? ti * BM : (d.kb_mode == KB_TJ ? tj * BN : 0);
  int k_end = d.ke_mode == KE_TI ? (ti + 1) * BM : (d.ke_mode == KE_TJ ? (tj + 1) * BN : d.K);
  if (k_end > d.K) k_end = d.K;
  const int nk = k_end > k_begin ? (k_end - k_begin) / BK : 0;

  const double* gA = A_KMAJOR ? A + (long)k_begin * d.lda + (long)ti * BM : A + (long)ti * BM * d.lda + k_begin;
  const double* gB = B_KMAJOR ? B + (long)k_begin * d.ldb + (long)tj * BN : B + (long)tj * BN * d.ldb + k_begin;
  const long stepA = A_KMAJOR ? (long)BK * d.lda : BK;
  const long stepB = B_KMAJOR ? (long)BK * d.ldb : BK;

  double acc[8][4][2];
#pragma unroll
  for (int f = 0; f < 8; f++)
#pragma unroll
    for (int h = 0; h < 4; h++) acc[f][h][0] = acc[f][h][1] = 0.0;

  // ---- prologue: fill STAGES-1 slots ----
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) {
      double* sA = smem + s * STAGE_DOUBLES;
      load_tile<A_KMAJOR>(sA, gA + s * stepA, d.lda, tid);
      load_tile<B_KMAJOR>(sA + TILE_DOUBLES, gB + s * stepB, d.ldb, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nx = kt + STAGES - 1;
      if (nx < nk) {
        double* sA = smem + (nx % STAGES) * STAGE_DOUBLES;
        load_tile<A_KMAJOR>(sA, gA + nx * stepA, d.lda, tid);
        load_tile<B_KMAJOR>(sA + TILE_DOUBLES, gB + nx * stepB, d.ldb, tid);
      }
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double* sB = sA + TILE_DOUBLES;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double a[8], b[4];
#pragma unroll
      for (int f = 0; f < 8; f++)
        a[f] = A_KMAJOR ? sA[(4 * ks + q) * LD_KM + wm + 8 * f + g] : sA[(wm + 8 * f + g) * LD_RM + 4 * ks + q];
#pragma unroll
      for (int h = 0; h < 4; h++)
        b[h] = B_KMAJOR ? sB[(4 * ks + q) * LD_KM + wn + 8 * h + g] : sB[(wn + 8 * h + g) * LD_RM + 4 * ks + q];
#pragma unroll
      for (int f = 0; f < 8; f++)
#pragma unroll
        for (int h = 0; h < 4; h++) dmma(acc[f][h][0], acc[f][h][1], a[f], b[h]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----
  if (d.epilogue == EPI_COLSUMSQ) {
    // column sums of squares of this 128 x 128 tile -> C[ti * ldc + tj*BN + col]
    __syncthreads();
    double* red = smem;  // [2][128]
#pragma unroll
    for (int h = 0; h < 4; h++) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int f = 0; f < 8; f++) {
        s0 += acc[f][h][0] * acc[f][h][0];
        s1 += acc[f][h][1] * acc[f][h][1];
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (g == 0) {
        red[(warp >> 2) * 128 + wn + 8 * h + 2 * q] = s0;
        red[(warp >> 2) * 128 + wn + 8 * h + 2 * q + 1] = s1;
      }
    }
    __syncthreads();
    if (tid < 128) C[(long)ti * d.ldc + (long)tj * BN + tid] = d.alpha * d.alpha * (red[tid] + red[128 + tid]);
    return;
  }
  double* Ct = C + (long)(ti * BM + wm) * d.ldc + (long)tj * BN + wn;
  const double alpha = d.alpha, beta = d.beta;
#pragma unroll
  for (int f = 0; f < 8; f++) {
#pragma unroll
    for (int h = 0; h < 4; h++) {
      double2* p = reinterpret_cast<double2*>(Ct + (long)(8 * f + g) * d.ldc + 8 * h + 2 * q);
      double2 v;
      v.x = alpha * acc[f][h][0];
      v.y = alpha * acc[f][h][1];
      if (d.epilogue == EPI_BIAS) {
        const double2 bb = *reinterpret_cast<const double2*>(d.bias + (long)tj * BN + wn + 8 * h + 2 * q);
        v.x += bb.x;
        v.y += bb.y;
      } else if (beta != 0.0) {
        const double2 o = *p;
        v.x += beta * o.x;
        v.y += beta * o.y;
      }
      *p = v;
    }
  }
}

}  // namespace gpras
