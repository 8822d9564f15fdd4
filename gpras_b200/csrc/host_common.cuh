// Host-side building blocks shared by the exact-GP and sparse-GP translation units: error plumbing, launchers of
// the DMMA tile engine, the look-ahead Cholesky, the recursive-doubling triangular inverse, W^T W, the covariance
// builder.  Everything here has internal linkage (one copy per translation unit); only g_err is shared.
#pragma once
#include "../../include/gpras_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gemm_engine.cuh"
#include "gp_kernels.cuh"
#include "leaf.cuh"

using namespace gpras;

inline thread_local std::string g_err;  // last error message of the calling thread (gpras_last_error)

namespace {



int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess)
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  else
    snprintf(buf, sizeof buf, "%s", what);
  g_err = buf;
  return code;
}

#define CU(x)                                         \
  do {                                                \
    cudaError_t e__ = (x);                            \
    if (e__ != cudaSuccess) return fail(GPRAS_E_CUDA, #x, e__); \
  } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <typename K>
int opt_in_smem(K kernel, int bytes) {
  // (Asking for the maximum shared-memory carveout as well was measured and rejected: the Cholesky got 2 % slower.)
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

// cudaFuncSetAttribute is per device; remember which devices were prepared.
std::atomic<bool> g_prepared[64] = {};

template <int KID>
int prepare_kid() {
  return opt_in_smem(cov_kernel<KID>, 200 * 1024);
}

int prepare_device() {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 64 && g_prepared[dev]) return 0;
  int r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgL, false, false>, CfgL::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgL, false, true>, CfgL::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgL, true, false>, CfgL::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgL, true, true>, CfgL::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgS, false, false>, CfgS::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgP, false, false>, CfgP::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgT, false, false>, CfgT::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgN, false, true>, CfgN::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(gemm_tile_kernel<CfgN, true, true>, CfgN::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(leaf_potrf_inv_kernel, LEAF_SMEM_BYTES))) return r;
  if ((r = opt_in_smem(leaf_potrf_kernel, LEAF_SMEM_BYTES))) return r;
  if ((r = opt_in_smem(leaf_inv_kernel, LEAF_SMEM_BYTES))) return r;
  if ((r = opt_in_smem(trsm_panel_kernel, TRSM_SMEM_BYTES))) return r;
  if ((r = prepare_kid<K_RBF>())) return r;
  if ((r = prepare_kid<K_MATERN12>())) return r;
  if ((r = prepare_kid<K_MATERN32>())) return r;
  if ((r = prepare_kid<K_MATERN52>())) return r;
  if ((r = prepare_kid<K_EXPONENTIAL>())) return r;
  if (dev < 64) g_prepared[dev] = true;
  return 0;
}

GemmDesc make_desc(const double* A, long lda, const double* B, long ldb, double* C, long ldc, int m_tiles, int n_tiles,
                   int K) {
  GemmDesc d;
  memset(&d, 0, sizeof d);
  d.A = A, d.B = B, d.C = C, d.lda = lda, d.ldb = ldb, d.ldc = ldc;
  d.m_tiles = m_tiles, d.n_tiles = n_tiles, d.K = K;
  d.alpha = 1.0, d.beta = 0.0;
  return d;
}

enum TileShape { SHAPE_L = 0, SHAPE_S = 1, SHAPE_N = 2, SHAPE_P = 3, SHAPE_T = 4 };

template <typename Cfg, bool AKM, bool BKM>
int launch_cfg(cudaStream_t s, const GemmDesc& d, int batch, int nz) {
  constexpr int ratio = Cfg::BM / Cfg::BN;
  long tiles = d.tri ? (long)ratio * d.m_tiles * (d.m_tiles + 1) / 2 : (long)d.m_tiles * d.n_tiles;
  dim3 grid((unsigned)tiles, (unsigned)batch, (unsigned)nz);
  gemm_tile_kernel<Cfg, AKM, BKM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(d);
  CU(cudaGetLastError());
  return 0;
}

// m_tiles counts BM-row tiles (128; 64 for SHAPE_P / SHAPE_T); n_tiles counts BN-column tiles of the chosen shape.
int launch_gemm(cudaStream_t s, bool akm, bool bkm, const GemmDesc& d, int batch, int* launches, int shape = SHAPE_L,
                int nz = 1) {
  if (d.m_tiles <= 0 || d.n_tiles <= 0 || batch <= 0) return 0;
  if (launches) ++*launches;
  if (shape == SHAPE_L) {
    if (!akm && !bkm) return launch_cfg<CfgL, false, false>(s, d, batch, nz);
    if (!akm && bkm) return launch_cfg<CfgL, false, true>(s, d, batch, nz);
    if (akm && !bkm) return launch_cfg<CfgL, true, false>(s, d, batch, nz);
    return launch_cfg<CfgL, true, true>(s, d, batch, nz);
  }
  if (shape == SHAPE_S) {
    if (!akm && !bkm) return launch_cfg<CfgS, false, false>(s, d, batch, nz);
    return fail(GPRAS_E_ARG, "CfgS is instantiated for the NT layout only");
  }
  if (shape == SHAPE_T) {
    if (!akm && !bkm && !d.tri) return launch_cfg<CfgT, false, false>(s, d, batch, nz);
    return fail(GPRAS_E_ARG, "CfgT is instantiated for the NT layout, non-triangular, only");
  }
  if (shape == SHAPE_P) {
    if (!akm && !bkm && !d.tri) return launch_cfg<CfgP, false, false>(s, d, batch, nz);
    return fail(GPRAS_E_ARG, "CfgP is instantiated for the NT layout, non-triangular, only");
  }
  if (!akm && bkm) return launch_cfg<CfgN, false, true>(s, d, batch, nz);
  if (akm && bkm) return launch_cfg<CfgN, true, true>(s, d, batch, nz);
  return fail(GPRAS_E_ARG, "CfgN is instantiated for k-major B only");
}

// Skinny product with split-k: partial results in `part` (nz slabs of rows x ldc), reduced in fixed order into C.
int launch_skinny(cudaStream_t s, bool akm, GemmDesc d, double* part, int rows, int* launches) {
  int ks = d.K / 16;
  ks = (ks + 127) / 128 * 128;
  if (ks < 512) ks = 512;
  const int nz = (d.K + ks - 1) / ks;
  double* out = d.C;
  const long slab = (long)rows * d.ldc;
  d.k_split = ks;
  d.splitC = slab;
  d.C = part;
  int r = launch_gemm(s, akm, true, d, 1, launches, SHAPE_N, nz);
  if (r) return r;
  splitk_reduce_kernel<<<(unsigned)((slab + 255) / 256), 256, 0, s>>>(part, slab, nz, slab, out);
  if (launches) ++*launches;
  CU(cudaGetLastError());
  return 0;
}
constexpr int SKINNY_MAX_SLABS = 16;
constexpr int PANEL_GROUP = 4;    // Cholesky: trailing updates for groups of this many panels ...
constexpr int PAIR_MIN_REM = 24;  // ... while more than this many block rows remain ...
constexpr int TAIL_GROUP = 1;     // ... and of this many afterwards
constexpr int BULK_L_RANK = 0;
constexpr int WIDE_COL_REM = 32;  // the side stream's block-column update uses the 128x64 shape above this many remaining rows

// ---- dense building blocks -------------------------------------------------------------------
// Tunables of the Cholesky driver, read from the environment once per process (development sweeps).
struct PotrfTuning {
  int group = PANEL_GROUP, min_rem = PAIR_MIN_REM, tail_group = TAIL_GROUP, wide_col_rem = WIDE_COL_REM;
  int bulk_l_rank = BULK_L_RANK;  // bulk trailing updates of at least this rank use the 128 x 128 long-k shape (0: never)
  PotrfTuning() {
    if (const char* e = getenv("GPRAS_B200_BULK_L_RANK")) bulk_l_rank = atoi(e);
    if (const char* e = getenv("GPRAS_B200_PANEL_GROUP")) group = atoi(e) > 0 ? atoi(e) : 1;
    if (const char* e = getenv("GPRAS_B200_PAIR_MIN_REM")) min_rem = atoi(e);
    if (const char* e = getenv("GPRAS_B200_TAIL_GROUP")) tail_group = atoi(e) > 0 ? atoi(e) : 1;
    if (const char* e = getenv("GPRAS_B200_WIDE_COL_REM")) wide_col_rem = atoi(e);
  }
};
inline const PotrfTuning& potrf_tuning() {
  static const PotrfTuning t;
  return t;
}

#ifdef POTRF_TIMELINE
// development only (tools/microbench/chain_timing.cu): timestamps of every step's kernels
struct PotrfTimeline {
  std::vector<cudaEvent_t> ev;  // per step: [bulk start, bulk end, trsm end, diag end, leaf end, col end]
  cudaEvent_t t0;
  cudaEvent_t get(int j, int k) {
    const size_t i = (size_t)j * 6 + k;
    while (ev.size() <= i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    return ev[i];
  }
};
inline PotrfTimeline* g_timeline = nullptr;
#define TL(j, k, st) do { if (g_timeline) cudaEventRecord(g_timeline->get(j, k), st); } while (0)
#else
#define TL(j, k, st)
#endif

// Side streams + events of the Cholesky's look-ahead.
struct LookAhead {
  cudaStream_t side = nullptr;   // the chain: panel solve -> diagonal-block update -> leaf (highest priority)
  cudaStream_t side2 = nullptr;  // the rest of the next block column (needed one leaf later)
  std::vector<cudaEvent_t> ev;
  int ensure(size_t n) {
    if (!side) {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // `hi` is the greatest priority (numerically lowest)
      CU(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
      CU(cudaStreamCreateWithPriority(&side2, cudaStreamNonBlocking, hi < lo ? hi + 1 : hi));
    }
    while (ev.size() < n) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    return 0;
  }
  void destroy() {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
    if (side) cudaStreamDestroy(side);
    if (side2) cudaStreamDestroy(side2);
    side = side2 = nullptr;
  }
};

// Right-looking blocked Cholesky, panel width 128, everything in place in A.  Per step j three streams cooperate:
//   chain (high priority)  panel  L[j+1.., j] = A[j+1.., j] L_jj^-T     triangular solve, 32 rows per CTA (trsm_panel_kernel)
//                          diag   A[j+1, j+1] -= sum_pending L[j+1, k] L[j+1, k]^T           (CfgT, 8 CTAs)
//                          leaf   L_{j+1, j+1} = chol(A[j+1, j+1])                           (leaf_potrf_kernel, 1 CTA)
//   side                   col    A[j+2.., j+1] -= sum_pending L[j+2.., k] L[j+1, k]^T       (CfgT) -- needed by panel j+1 only
//   main                   rest   the trailing triangle right of block column j+1 -= (group of panels)(...)^T   (CfgS)
// The chain is what bounds the factorisation once `rest` gets short, so it carries only what the next leaf needs: the
// leaf does not invert its block (the panel is a SOLVE against L_jj, not a product with its inverse), and of block
// column j+1 only the diagonal block is updated on the chain.  The inverses W_jj of all diagonal blocks, which the
// triangular inverse starts from, come from ONE launch after the factorisation (leaf_inv_kernel, one CTA per block).
// (Round 1 had leaf + inverse 42 us, column update and panel product on the chain: ~60 us per step.)
int potrf_impl(cudaStream_t s, LookAhead& la, double* A, long lda, double* W, long ldw, int n, double* logdet_parts,
               int* info, int* launches) {
  const int nt = n / 128;
  int r;
  if (nt == 1) {  // single block: one kernel factors and inverts, no side streams
    leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, s>>>(A, lda, A, lda, W, ldw, logdet_parts, info, 0);
    if (launches) ++*launches;
    CU(cudaGetLastError());
    return 0;
  }
  if ((r = la.ensure(4 * (size_t)nt + 3))) return r;
  cudaStream_t s2 = la.side, s3 = la.side2;
  cudaEvent_t* evPanel = la.ev.data();            // [nt] panel j solved
  cudaEvent_t* evAhead = la.ev.data() + nt;       // [nt] look-ahead part of the bulk update issued at step j done
  cudaEvent_t* evCol = la.ev.data() + 2 * nt;     // [nt] rows below the diagonal of block column j up to date
  cudaEvent_t evFork = la.ev[4 * nt], evJoin = la.ev[4 * nt + 1], evJoin2 = la.ev[4 * nt + 2];
  auto leaf = [&](cudaStream_t st, int jb) -> int {
    leaf_potrf_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, st>>>(A, lda, logdet_parts, info, jb);
    if (launches) ++*launches;
    CU(cudaGetLastError());
    return 0;
  };
  if ((r = leaf(s, 0))) return r;
  CU(cudaEventRecord(evFork, s));
  CU(cudaStreamWaitEvent(s2, evFork, 0));
  CU(cudaStreamWaitEvent(s3, evFork, 0));
  // Trailing updates are applied for GROUPS of G panels (rank 128 G) while the trailing matrix is large: that divides the
  // read-modify-write traffic on the trailing matrix by G and multiplies the k extent per tile of the short-k kernel
  // (measured alone: DMMA pipe 76 % active at rank 128, 82 % at 256, 91 % at 512).  The chain brings block column j+1 fully
  // up to date itself (1 .. G pending panels).  Every bulk update is issued in two launches: first the block columns the
  // NEXT group's chain works on (look-ahead part), then the rest; the next group's chain waits for the first only, so the
  // bulk stream never idles behind the chain (with one launch it did: ~63 us per group, which cancelled the gain of
  // grouping).  Once the chain bounds the step (short trailing matrix), single panels keep the pipeline fine-grained.  The
  // switch happens at a multiple of G, where no panel is pending.
  const int G = potrf_tuning().group, Gt = potrf_tuning().tail_group;
  int j_single = nt - 1 - potrf_tuning().min_rem;  // first step whose trailing matrix is short enough for the tail grouping ...
  if (j_single < 0) j_single = 0;
  j_single = (j_single + G - 1) / G * G;  // ... rounded up to a group boundary
  int last_bulk = -1;  // step of the most recent bulk update
  for (int j = 0; j + 1 < nt; j++) {
    const int rem = nt - j - 1;  // tiles below / right of block j
    const bool grouped = j < j_single;
    const int gsz = grouped ? G : Gt;                 // size of this step's group
    const int pc = (grouped ? j % G : (j - j_single) % Gt) + 1;  // pending panels at this step
    const bool bulk = pc == gsz;                      // the group is complete: apply it to the rest of the trailing matrix
    const int kp = 128 * pc;                          // k extent of the chain's updates
    const long pcol = (long)(j - (pc - 1)) * 128;     // first column of the pending panel(s)
    double* pn = A + (long)(j + 1) * 128 * lda + pcol;  // pending panel(s), rows from tile j+1
    // ---- chain: solve panel j, bring the next diagonal block up to date, factor it ----
    if (j > 0) CU(cudaStreamWaitEvent(s2, evCol[j], 0));
    trsm_panel_kernel<<<rem * (LEAF_N / TRSM_ROWS), TRSM_THREADS, TRSM_SMEM_BYTES, s2>>>(A, lda, j);
    if (launches) ++*launches;
    CU(cudaGetLastError());
    CU(cudaEventRecord(evPanel[j], s2));
    TL(j, 2, s2);
    // (first step of a group: the previous group's bulk update must have reached this group's block columns)
    if (pc == 1 && last_bulk >= 0) CU(cudaStreamWaitEvent(s2, evAhead[last_bulk], 0));
    {
      double* djj = A + (long)(j + 1) * 128 * (lda + 1);
      GemmDesc c = make_desc(pn, lda, pn, lda, djj, lda, 2, 4, kp);
      c.alpha = -1.0, c.beta = 1.0;
      if ((r = launch_gemm(s2, false, false, c, 1, launches, SHAPE_T))) return r;
    }
    TL(j, 3, s2);
    if ((r = leaf(s2, j + 1))) return r;
    TL(j, 4, s2);
    if (rem > 1) {
      // ---- side: the rest of block column j+1 (rows from tile j+2), needed when leaf j+1 is done ----
      CU(cudaStreamWaitEvent(s3, evPanel[j], 0));
      if (pc == 1 && last_bulk >= 0) CU(cudaStreamWaitEvent(s3, evAhead[last_bulk], 0));
      {
        double* pn2 = pn + (long)128 * lda;                                     // pending panel(s), rows from tile j+2
        double* col = A + (long)(j + 2) * 128 * lda + (long)(j + 1) * 128;      // block column j+1, rows from tile j+2
        // many small tiles while this update sits next to the critical path, the throughput shape while the bulk update
        // leaves it an order of magnitude of slack
        const bool wide = rem > potrf_tuning().wide_col_rem;
        GemmDesc c = make_desc(pn2, lda, pn, lda, col, lda, wide ? rem - 1 : 2 * (rem - 1), wide ? 2 : 4, kp);
        c.alpha = -1.0, c.beta = 1.0;
        if ((r = launch_gemm(s3, false, false, c, 1, launches, wide ? SHAPE_S : SHAPE_T))) return r;
      }
      CU(cudaEventRecord(evCol[j + 1], s3));
      TL(j, 5, s3);
      // ---- main stream: the trailing triangle from block column j+2, the whole group of panels together ----
      if (bulk) {
        CU(cudaStreamWaitEvent(s, evPanel[j], 0));
        double* pn2 = pn + (long)128 * lda;  // pending panel(s): rows from tile j+2
        // look-ahead part: the block columns of the next group (rectangular launch over rows j+2.. x those columns; the few
        // tiles above the diagonal it touches are never read)
        const int next_g = (j + 1 < j_single) ? G : Gt;
        const int wa = next_g < rem - 1 ? next_g : rem - 1;  // its width in block columns
        TL(j, 0, s);
        const bool bulk_l = potrf_tuning().bulk_l_rank > 0 && kp >= potrf_tuning().bulk_l_rank;
        {
          double* ahead = A + (long)(j + 2) * 128 * (lda + 1);
          GemmDesc u = make_desc(pn2, lda, pn2, lda, ahead, lda, rem - 1, bulk_l ? wa : 2 * wa, kp);
          u.alpha = -1.0, u.beta = 1.0;
          if ((r = launch_gemm(s, false, false, u, 1, launches, bulk_l ? SHAPE_L : SHAPE_S))) return r;
        }
        CU(cudaEventRecord(evAhead[j], s));
        last_bulk = j;
        if (rem - 1 > wa) {  // the rest: lower triangle from block column j+2+wa
          double* pn3 = pn2 + (long)wa * 128 * lda;
          double* trail = A + (long)(j + 2 + wa) * 128 * (lda + 1);
          GemmDesc u = make_desc(pn3, lda, pn3, lda, trail, lda, rem - 1 - wa, (bulk_l ? 1 : 2) * (rem - 1 - wa), kp);
          u.tri = 1, u.alpha = -1.0, u.beta = 1.0;
          if ((r = launch_gemm(s, false, false, u, 1, launches, bulk_l ? SHAPE_L : SHAPE_S))) return r;
        }
        TL(j, 1, s);
      }
    }
  }
  CU(cudaEventRecord(evJoin, s2));
  CU(cudaEventRecord(evJoin2, s3));
  CU(cudaStreamWaitEvent(s, evJoin, 0));
  CU(cudaStreamWaitEvent(s, evJoin2, 0));
  // inverses of all diagonal blocks, one CTA each
  leaf_inv_kernel<<<nt, LEAF_THREADS, LEAF_SMEM_BYTES, s>>>(A, lda, W, ldw);
  if (launches) ++*launches;
  CU(cudaGetLastError());
  return 0;
}

// W = L^-1 by recursive doubling over 128-tiles: at block size b every pair of adjacent diagonal
// blocks gets W21 = -W22 (L21 W11); all pairs of a level run as one batched launch per product.
int trtri_impl(cudaStream_t s, const double* L, long ldl, double* W, long ldw, double* T, long ldt, int n,
               int* launches) {
  const int nt = n / 128;
  for (int b = 1; b < nt; b <<= 1) {
    const int full = nt / (2 * b);          // pairs with a complete second block
    const int rag = nt - full * 2 * b;      // leftover tiles
    for (int pass = 0; pass < 2; pass++) {
      int batch, m2;
      long r0;
      if (pass == 0) {
        batch = full, m2 = b, r0 = 0;
      } else {
        batch = 1, m2 = rag - b, r0 = (long)full * 2 * b * 128;  // ragged pair: first block b, second rag-b
      }
      if (batch <= 0 || m2 <= 0) continue;
      const long rb = r0 + (long)b * 128;
      const long bsL = (long)2 * b * 128 * (ldl + 1), bsW = (long)2 * b * 128 * (ldw + 1),
                 bsT = (long)2 * b * 128 * (ldt + 1);
      // T = L21 W11   (A row-major full; B = W11 k-major, lower: k >= tj)
      GemmDesc g1 = make_desc(L + rb * ldl + r0, ldl, W + r0 * (ldw + 1), ldw, T + rb * ldt + r0, ldt, m2, b, b * 128);
      g1.kb_mode = KB_TJ;
      g1.colmajor = 1;  // k range shrinks with the column tile: heaviest columns first
      g1.batchA = bsL, g1.batchB = bsW, g1.batchC = bsT;
      int r = launch_gemm(s, false, true, g1, batch, launches);
      if (r) return r;
      // W21 = -W22 T  (A = W22 row-major lower: k < (ti+1)*128; B = T k-major)
      GemmDesc g2 = make_desc(W + rb * (ldw + 1), ldw, T + rb * ldt + r0, ldt, W + rb * ldw + r0, ldw, m2, b, m2 * 128);
      g2.ke_mode = KE_TI;
      g2.alpha = -1.0;
      g2.reverse = 1;
      g2.batchA = bsW, g2.batchB = bsT, g2.batchC = bsW;
      if ((r = launch_gemm(s, false, true, g2, batch, launches))) return r;
    }
  }
  return 0;
}

// Kinv = W^T W, lower tiles; k runs from the row tile of C to n (both operands lower-triangular).
int lauum_impl(cudaStream_t s, const double* W, long ldw, double* Kinv, long ldk, int n, int* launches) {
  const int nt = n / 128;
  GemmDesc d = make_desc(W, ldw, W, ldw, Kinv, ldk, nt, nt, n);
  d.tri = 1;
  d.kb_mode = KB_TI;
  return launch_gemm(s, true, true, d, 1, launches);
}

template <int KID>
int launch_cov(cudaStream_t s, const double* Xs1, int n1, int n1_pad, const double* Xs2, int n2, int n2_pad, int D,
               const double* theta, double* out, long ldo, int square, double jitter) {
  const int t1 = n1_pad / CT, t2 = n2_pad / CT;
  const int smem = 2 * D * CT_LD * (int)sizeof(double);
  const long tiles = square ? (long)t1 * (t1 + 1) / 2 : (long)t1 * t2;
  cov_kernel<KID><<<(unsigned)tiles, PT_THREADS, smem, s>>>(Xs1, n1, Xs2, n2, D, theta, out, ldo, t2, square, square, jitter);
  CU(cudaGetLastError());
  return 0;
}

int dispatch_cov(int kid, cudaStream_t s, const double* Xs1, int n1, int n1_pad, const double* Xs2, int n2, int n2_pad,
                 int D, const double* theta, double* out, long ldo, int square, double jitter = -1.0) {
  switch (kid) {
    case K_RBF: return launch_cov<K_RBF>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square, jitter);
    case K_MATERN12: return launch_cov<K_MATERN12>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square, jitter);
    case K_MATERN32: return launch_cov<K_MATERN32>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square, jitter);
    case K_MATERN52: return launch_cov<K_MATERN52>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square, jitter);
    case K_EXPONENTIAL:
      return launch_cov<K_EXPONENTIAL>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square, jitter);
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

int dalloc(double** p, size_t count) {
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(double));
  if (e != cudaSuccess) return fail(GPRAS_E_NOMEM, "cudaMalloc", e);
  return 0;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

static __global__ void broadcast_var_kernel(const double* __restrict__ var, double* __restrict__ varm, int T, int P, long ld) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)T * ld) return;
  int c = (int)(e % ld);
  varm[e] = c < P ? var[e / ld] : 0.0;
}

}  // namespace
