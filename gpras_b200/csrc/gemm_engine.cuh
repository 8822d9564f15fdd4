// FP64 tile-GEMM engine on the DMMA pipe (sm_100a).
//
// One kernel template covers every dense contraction of the exact-GP path: the Cholesky panel and
// trailing updates, the triangular inverse (recursive doubling), K^-1 = W^T W, alpha = W^T (W Y), the
// predictive V = W K*^T with a fused column sum-of-squares, and the modes -> cells expansion.
//
//   C[i, j] (op)= alpha * sum_{k in [k_begin, k_end)} A(i, k) * B(k, j)
//
// * 8 or 16 warps arranged WARPS_M x WARPS_N over a BM x BN CTA tile; every warp owns
//   MF x NF DMMA.8x8x4 accumulators.  Three tile shapes are instantiated:
//       CfgL 128x128, warp 64x32, BK 32, 3 stages, 1 CTA/SM  -- long-k products (LAUUM, TRTRI, predictor)
//       CfgS 128x64,  warp 32x32, 3 stages, 2 CTA/SM  -- short-k rank-128 updates of the Cholesky, where a
//                                                        second resident CTA hides prologue / epilogue
//       CfgN 128x32,  warp 16x32, 4 stages, 2 CTA/SM  -- skinny right-hand sides (P <= 32 columns), split-k
//       CfgP 64x128,  warp 32x32, 3 stages, 2 CTA/SM  -- the in-place Cholesky panel L21 = A21 W_jj^T (first panel)
//       CfgT 64x32,   warp 16x16, 3 stages, 4 CTA/SM  -- the chain: next block column update and out-of-place panel;
//                                                        16x more CTAs than CfgL so a 128-wide product finishes in a few us
// * Operand tiles stream global -> shared with 16-byte cp.async in a multi-stage ring (BK = 16, or 32 for CfgL).
// * A is either row-major A[i][k] or k-major A[k][i]; B either n-major B[j][k] or k-major B[k][j];
//   shared tiles are padded (+4 doubles) so every fragment LDS.64 is bank-conflict free.
// * Triangular structure is exploited at tile granularity by clipping the k range per tile
//   (k_begin / k_end modes) and by launching only the lower-triangular tiles of C (tri mode).
//   Operand tiles that straddle the diagonal must hold explicit zeros in their dead half.
// * beta != 0 initialises the accumulators from C before the k loop, so the read of C overlaps the
//   pipeline prologue instead of sitting in the epilogue.
// * k_split > 0 cuts the k range into chunks handled by blockIdx.z, each writing its own partial C
//   (stride splitC); a fixed-order reduction kernel adds them, keeping results bitwise repeatable.
// * All extents are multiples of the tile sizes: the host pads (identity on the diagonal).
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int K_ALIGN = 32;  // every k extent / clip point handed to the engine is a multiple of this (>= any BK)

template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_, int MINB_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int MF = WM / 8, NF = WN / 8;
  static constexpr int LD_RM = BK + 4;  // row / n-major tile [rows][BK + 4] (k contiguous); == 4 mod 16
  static constexpr int A_DOUBLES = BM * LD_RM > BK * (BM + 4) ? BM * LD_RM : BK * (BM + 4);
  static constexpr int B_DOUBLES = BN * LD_RM > BK * (BN + 4) ? BN * LD_RM : BK * (BN + 4);
  static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);
  static_assert(THREADS == 256 || THREADS == 512, "8 or 16 warps per CTA");
  static_assert(BK % 16 == 0 && BK <= K_ALIGN, "BK");
};
using CfgL = TileCfg<128, 128, 64, 32, 32, 3, 1>;  // 216 KiB: three 72 KiB stages, one barrier per 256 DMMAs / warp
// (16 warps of 32x32 on the same tile measured 2% slower on LAUUM: the loop is not barrier-skew bound)
using CfgS = TileCfg<128, 64, 32, 32, 16, 3, 2>;
using CfgN = TileCfg<128, 32, 16, 32, 16, 4, 2>;
using CfgP = TileCfg<64, 128, 32, 32, 16, 3, 2>;   // in-place Cholesky panel: full 128-column width per CTA (no tri mode)
using CfgT = TileCfg<64, 32, 16, 16, 16, 3, 4>;    // latency-critical rank-128 products of the Cholesky chain: many small CTAs

enum KBegin { KB_ZERO = 0, KB_TI = 1, KB_TJ = 2 };   // k_begin = 0 | ti*BM | tj*BN
enum KEnd { KE_FULL = 0, KE_TI = 1, KE_TJ = 2 };     // k_end   = K | (ti+1)*BM | (tj+1)*BN
enum Epilogue { EPI_STORE = 0, EPI_COLSUMSQ = 1, EPI_BIAS = 2 };

struct GemmDesc {
  const double* A;
  const double* B;
  double* C;            // EPI_STORE / EPI_BIAS: output; EPI_COLSUMSQ: partials [m_tiles][ldc]
  const double* bias;   // EPI_BIAS: per-column bias (length n_tiles*BN)
  const double* Cin;    // beta != 0: source of the beta * C term (pitch ldcin); NULL = C itself
  long ldcin;
  long lda, ldb, ldc;
  long batchA, batchB, batchC;  // element strides between batch entries (grid.y)
  long splitC;          // element stride between k-split partial outputs (grid.z)
  int m_tiles, n_tiles;
  int K;                // multiple of K_ALIGN
  int k_split;          // 0, or chunk length (multiple of K_ALIGN) per blockIdx.z
  int kb_mode, ke_mode;
  int tri;              // 1: only tiles whose columns start at or below the row tile's last row
  int reverse;          // 1: launch order reversed (heaviest tiles first for LPT scheduling)
  int colmajor;         // 1: non-tri tiles enumerate rows fastest (use when the work depends on the column tile)
  int epilogue;
  double alpha, beta;
};

// ROWS x BK tile, global -> shared.  Row-major: ROWS rows of BK doubles; k-major: BK rows of ROWS doubles.
// PART / NPARTS issues only every NPARTS-th chunk of the thread, so the main loop can trickle the next stage's
// cp.async between its DMMA groups instead of stalling the tensor pipe behind a burst of LDGSTS after the barrier.
template <bool KMAJOR, int ROWS, int BK, int GEMM_THREADS, int PART = 0, int NPARTS = 1>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g, long ld, int tid) {
  constexpr int CHUNKS = ROWS * BK / 2;  // 16-byte chunks in the tile
  constexpr int LD_RM = BK + 4;
  constexpr int PER_THREAD = (CHUNKS + GEMM_THREADS - 1) / GEMM_THREADS;
#pragma unroll
  for (int q = PART; q < PER_THREAD; q += NPARTS) {
    const int c = tid + GEMM_THREADS * q;
    if (CHUNKS % GEMM_THREADS == 0 || c < CHUNKS) {
      if (!KMAJOR) {
        constexpr int CPK = BK / 2;  // chunks per row
        const int row = c / CPK, kc = c - row * CPK;
        cp_async16(s + row * LD_RM + 2 * kc, g + (long)row * ld + 2 * kc);
      } else {
        constexpr int CPR = ROWS / 2;  // chunks per k row
        const int kr = c / CPR, mc = c - kr * CPR;
        cp_async16(s + kr * (ROWS + 4) + 2 * mc, g + (long)kr * ld + 2 * mc);
      }
    }
  }
}

// compile-time loop over the parts (PART is a template argument of load_tile)
template <bool AK, bool BK_, typename Cfg, int KS>
__device__ __forceinline__ void load_stage_part(double* sA, const double* gA, long lda, const double* gB, long ldb, int tid) {
  constexpr int NP = Cfg::BK / 4;
  load_tile<AK, Cfg::BM, Cfg::BK, Cfg::THREADS, KS, NP>(sA, gA, lda, tid);
  load_tile<BK_, Cfg::BN, Cfg::BK, Cfg::THREADS, KS, NP>(sA + Cfg::A_DOUBLES, gB, ldb, tid);
}

template <typename Cfg, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_tile_kernel(const GemmDesc d) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, MF = Cfg::MF, NF = Cfg::NF, STAGES = Cfg::STAGES, BK = Cfg::BK;
  constexpr int LD_RM = Cfg::LD_RM;
  constexpr int LDA_KM = BM + 4, LDB_KM = BN + 4, RATIO = BM >= BN ? BM / BN : 1;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = (warp / Cfg::WARPS_N) * Cfg::WM, wn = (warp % Cfg::WARPS_N) * Cfg::WN;

  // ---- tile coordinates ----
  int t = blockIdx.x;
  if (d.reverse) t = gridDim.x - 1 - t;
  int ti, tj;
  if (d.tri) {
    // row ti holds RATIO*(ti+1) column tiles; prefix = RATIO*ti*(ti+1)/2
    ti = (int)((sqrt(8.0 * (double)t / RATIO + 1.0) - 1.0) * 0.5);
    while ((long)RATIO * (ti + 1) * (ti + 2) / 2 <= t) ti++;
    while ((long)RATIO * ti * (ti + 1) / 2 > t) ti--;
    tj = t - (int)((long)RATIO * ti * (ti + 1) / 2);
  } else if (d.colmajor) {
    tj = t / d.m_tiles;
    ti = t - tj * d.m_tiles;
  } else {
    ti = t / d.n_tiles;
    tj = t - ti * d.n_tiles;
  }
  const long by = blockIdx.y;
  const double* __restrict__ A = d.A + by * d.batchA;
  const double* __restrict__ B = d.B + by * d.batchB;
  double* __restrict__ C = d.C + by * d.batchC + (long)blockIdx.z * d.splitC;

  int k_begin = d.kb_mode == KB_TI ? ti * BM : (d.kb_mode == KB_TJ ? tj * BN : 0);
  int k_end = d.ke_mode == KE_TI ? (ti + 1) * BM : (d.ke_mode == KE_TJ ? (tj + 1) * BN : d.K);
  if (k_end > d.K) k_end = d.K;
  k_begin = k_begin / BK * BK;
  if (d.k_split > 0) {
    const int lo = (int)blockIdx.z * d.k_split, hi = lo + d.k_split;
    if (k_begin < lo) k_begin = lo;
    if (k_end > hi) k_end = hi;
  }
  const int nk = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

  const double* gA = A_KMAJOR ? A + (long)k_begin * d.lda + (long)ti * BM : A + (long)ti * BM * d.lda + k_begin;
  const double* gB = B_KMAJOR ? B + (long)k_begin * d.ldb + (long)tj * BN : B + (long)tj * BN * d.ldb + k_begin;
  const long stepA = A_KMAJOR ? (long)BK * d.lda : BK;
  const long stepB = B_KMAJOR ? (long)BK * d.ldb : BK;

  // ---- prologue: fill STAGES-1 slots ----
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) {
      double* sA = smem + s * Cfg::STAGE_DOUBLES;
      load_tile<A_KMAJOR, BM, BK, Cfg::THREADS>(sA, gA + s * stepA, d.lda, tid);
      load_tile<B_KMAJOR, BN, BK, Cfg::THREADS>(sA + Cfg::A_DOUBLES, gB + s * stepB, d.ldb, tid);
    }
    cp_async_commit();
  }

  // ---- accumulators; beta != 0 folds C in up front (overlaps the prologue loads) ----
  double acc[MF][NF][2];
  double* Ct = C + (long)(ti * BM + wm) * d.ldc + (long)tj * BN + wn;
  const double alpha = d.alpha;
  if (d.epilogue == EPI_STORE && d.beta != 0.0) {
    const double scale = d.beta / alpha;
    const double* Cs = d.Cin ? d.Cin + by * d.batchC + (long)(ti * BM + wm) * d.ldcin + (long)tj * BN + wn : Ct;
    const long ldcs = d.Cin ? d.ldcin : d.ldc;
#pragma unroll
    for (int f = 0; f < MF; f++)
#pragma unroll
      for (int h = 0; h < NF; h++) {
        const double2 o = *reinterpret_cast<const double2*>(Cs + (long)(8 * f + g) * ldcs + 8 * h + 2 * q);
        acc[f][h][0] = scale * o.x;
        acc[f][h][1] = scale * o.y;
      }
  } else {
#pragma unroll
    for (int f = 0; f < MF; f++)
#pragma unroll
      for (int h = 0; h < NF; h++) acc[f][h][0] = acc[f][h][1] = 0.0;
  }

  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // stage (kt + STAGES - 1) % STAGES was consumed in iteration kt - 1: refill it while computing, one slice of
    // the cp.async burst after each k4 group of DMMAs
    const int nx = kt + STAGES - 1;
    const bool refill = nx < nk;
    double* nA = smem + (nx % STAGES) * Cfg::STAGE_DOUBLES;
    const double* ngA = gA + nx * stepA;
    const double* ngB = gB + nx * stepB;
    const double* sA = smem + (kt % STAGES) * Cfg::STAGE_DOUBLES;
    const double* sB = sA + Cfg::A_DOUBLES;
#define GPRAS_K4_GROUP(KS)                                                                                           \
    {                                                                                                                \
      constexpr int ks = KS;                                                                                         \
      double a[MF], b[NF];                                                                                           \
      _Pragma("unroll") for (int f = 0; f < MF; f++)                                                                 \
        a[f] = A_KMAJOR ? sA[(4 * ks + q) * LDA_KM + wm + 8 * f + g] : sA[(wm + 8 * f + g) * LD_RM + 4 * ks + q];    \
      _Pragma("unroll") for (int h = 0; h < NF; h++)                                                                 \
        b[h] = B_KMAJOR ? sB[(4 * ks + q) * LDB_KM + wn + 8 * h + g] : sB[(wn + 8 * h + g) * LD_RM + 4 * ks + q];    \
      _Pragma("unroll") for (int f = 0; f < MF; f++)                                                                 \
        _Pragma("unroll") for (int h = 0; h < NF; h++) dmma(acc[f][h][0], acc[f][h][1], a[f], b[h]);                 \
      if (refill) load_stage_part<A_KMAJOR, B_KMAJOR, Cfg, KS>(nA, ngA, d.lda, ngB, d.ldb, tid);                     \
    }
    GPRAS_K4_GROUP(0)
    GPRAS_K4_GROUP(1)
    GPRAS_K4_GROUP(2)
    GPRAS_K4_GROUP(3)
    if (BK > 16) {
      GPRAS_K4_GROUP(4)
      GPRAS_K4_GROUP(5)
      GPRAS_K4_GROUP(6)
      GPRAS_K4_GROUP(7)
    }
#undef GPRAS_K4_GROUP
    cp_async_commit();
  }
  cp_async_wait<0>();

  // ---- epilogue ----
  if (d.epilogue == EPI_COLSUMSQ) {
    // column sums of squares of this BM x BN tile -> C[ti * ldc + tj*BN + col]
    __syncthreads();
    double* red = smem;  // [WARPS_M][BN]
#pragma unroll
    for (int h = 0; h < NF; h++) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int f = 0; f < MF; f++) {
        s0 += acc[f][h][0] * acc[f][h][0];
        s1 += acc[f][h][1] * acc[f][h][1];
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (g == 0) {
        red[(warp / Cfg::WARPS_N) * BN + wn + 8 * h + 2 * q] = s0;
        red[(warp / Cfg::WARPS_N) * BN + wn + 8 * h + 2 * q + 1] = s1;
      }
    }
    __syncthreads();
    if (tid < BN) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < Cfg::WARPS_M; r++) s += red[r * BN + tid];
      C[(long)ti * d.ldc + (long)tj * BN + tid] = alpha * alpha * s;
    }
    return;
  }
#pragma unroll
  for (int f = 0; f < MF; f++) {
#pragma unroll
    for (int h = 0; h < NF; h++) {
      double2* p = reinterpret_cast<double2*>(Ct + (long)(8 * f + g) * d.ldc + 8 * h + 2 * q);
      double2 v;
      v.x = alpha * acc[f][h][0];
      v.y = alpha * acc[f][h][1];
      if (d.epilogue == EPI_BIAS) {
        const double2 bb = *reinterpret_cast<const double2*>(d.bias + (long)tj * BN + wn + 8 * h + 2 * q);
        v.x += bb.x;
        v.y += bb.y;
      }
      *p = v;
    }
  }
}

// out[e] = sum_z part[z * stride + e]  (fixed order -> deterministic), e < count
static __global__ void splitk_reduce_kernel(const double* __restrict__ part, long stride, int nz, long count,
                                     double* __restrict__ out, long bs = 0) {
  part += blockIdx.y * bs, out += blockIdx.y * bs;
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  double s = 0.0;
  for (int z = 0; z < nz; z++) s += part[(long)z * stride + e];
  out[e] = s;
}

}  // namespace gpras
