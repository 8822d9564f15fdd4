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
constexpr int PANEL_GROUP = 2;    // Cholesky: trailing updates for groups of this many panels ...
constexpr int PAIR_MIN_REM = 36;  // ... while more than this many block rows remain

// ---- dense building blocks -------------------------------------------------------------------
// Side stream + events for the one-panel look-ahead of the Cholesky.
struct LookAhead {
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> ev;
  double* scratch = nullptr;  // n x 128 column block of the Cholesky chain
  size_t scratch_count = 0;
  int ensure_scratch(size_t count) {
    if (count <= scratch_count) return 0;
    if (scratch) cudaFree(scratch);
    scratch = nullptr, scratch_count = 0;
    cudaError_t e = cudaMalloc((void**)&scratch, count * sizeof(double));
    if (e != cudaSuccess) return fail(GPRAS_E_NOMEM, "cudaMalloc (Cholesky scratch)", e);
    scratch_count = count;
    return 0;
  }
  int ensure(size_t n) {
    if (!side) {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));  // `hi` is the greatest priority
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
    side = nullptr;
    if (scratch) cudaFree(scratch);
    scratch = nullptr, scratch_count = 0;
  }
};

// Right-looking blocked Cholesky, panel width 128, with one panel of look-ahead on a high-priority side stream:
//   leaf   (1 CTA)   potrf + inverse of the diagonal block  -> L_jj, W_jj
//   panel  (CfgT)    L21 = A21 W_jj^T, from the scratch column block S into A (first panel: CfgP, in place)
//   col    (CfgT)    S = block column j+1 of the trailing matrix - panel_j panel_j^T      [side stream]
//   rest   (CfgS)    the trailing triangle right of block column j+1 -= panel_j panel_j^T  [main stream]
// so leaf(j+1) and panel(j+1) run while rest(j) occupies the machine.  The chain col -> leaf -> panel is what bounds
// the factorisation once rest(j) gets short, so its two products use 64x32 tiles (16x the CTAs of a 128x128 tiling)
// and go through the scratch block S (n x 128): out of place, any tile shape is race free.
int potrf_impl(cudaStream_t s, LookAhead& la, double* A, long lda, double* W, long ldw, int n, double* logdet_parts,
               int* info, int* launches) {
  const int nt = n / 128;
  int r;
  if ((r = la.ensure(2 * (size_t)nt + 2))) return r;
  if ((r = la.ensure_scratch((size_t)n * 128))) return r;
  double* S = la.scratch;  // S[i][0..127] = updated block column for global row i
  cudaStream_t s2 = la.side;
  cudaEvent_t* evPanel = la.ev.data();        // [nt]
  cudaEvent_t* evRest = la.ev.data() + nt;    // [nt]
  cudaEvent_t evFork = la.ev[2 * nt], evJoin = la.ev[2 * nt + 1];
  auto leaf = [&](cudaStream_t st, int jb, const double* in, long ldin) -> int {
    leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, st>>>(in, ldin, A, lda, W, ldw, logdet_parts, info, jb);
    if (launches) ++*launches;
    CU(cudaGetLastError());
    return 0;
  };
  if ((r = leaf(s, 0, A, lda))) return r;
  if (nt == 1) return 0;  // single block: nothing to overlap, the side stream stays out of it
  CU(cudaEventRecord(evFork, s));
  CU(cudaStreamWaitEvent(s2, evFork, 0));
  {  // first panel, in place (rows of tiles 1.., columns of block 0)
    double* pn = A + (long)128 * lda;
    GemmDesc p = make_desc(pn, lda, W, ldw, pn, lda, 2 * (nt - 1), 1, 128);
    if ((r = launch_gemm(s, false, false, p, 1, launches, SHAPE_P))) return r;
  }
  CU(cudaEventRecord(evPanel[0], s));
  // Trailing updates are applied for GROUPS of G panels (rank 128 G) while the trailing matrix is large: that divides the
  // read-modify-write traffic on the trailing matrix by G and multiplies the k extent per tile of the short-k kernel.  The
  // chain's column update brings block column j+1 fully up to date itself (1 .. G pending panels).  Grouping pays while the
  // bulk update dominates the step; once the leaf -> panel chain does (short trailing matrix), single panels keep the pipeline
  // fine-grained.  The switch happens at a multiple of G, where no panel is pending.
  int G = PANEL_GROUP, min_rem = PAIR_MIN_REM;
  if (const char* e = getenv("GPRAS_B200_PANEL_GROUP")) G = atoi(e) > 0 ? atoi(e) : 1;
  if (const char* e = getenv("GPRAS_B200_PAIR_MIN_REM")) min_rem = atoi(e);
  int j_single = nt - 1 - min_rem;  // first step whose trailing matrix is short enough for single panels ...
  if (j_single < 0) j_single = 0;
  j_single = (j_single + G - 1) / G * G;  // ... rounded up to a group boundary
  for (int j = 0; j + 1 < nt; j++) {
    const int rem = nt - j - 1;  // tiles below / right of block j
    const bool grouped = j < j_single;
    const int pc = grouped ? (j % G) + 1 : 1;         // pending panels at this step
    const bool bulk = pc == (grouped ? G : 1);        // the group is complete: apply it to the rest of the trailing matrix
    const int kp = 128 * pc;                          // k extent of the chain's column update
    const long pcol = (long)(j - (pc - 1)) * 128;     // first column of the pending panel(s)
    double* pn = A + (long)(j + 1) * 128 * lda + pcol;  // pending panel(s), rows from tile j+1
    // ---- side stream: block column j+1 -> S, then the next diagonal block and panel ----
    CU(cudaStreamWaitEvent(s2, evPanel[j], 0));
    if (pc == 1 && j > 0) CU(cudaStreamWaitEvent(s2, evRest[j - 1], 0));  // (later steps of a group: waited at its first step)
    double* Sj = S + (long)(j + 1) * 128 * 128;  // rows from tile j+1
    {
      const double* col = A + (long)(j + 1) * 128 * (lda + 1);
      GemmDesc c = make_desc(pn, lda, pn, lda, Sj, 128, 2 * rem, 4, kp);
      c.alpha = -1.0, c.beta = 1.0;
      c.Cin = col, c.ldcin = lda;
      if ((r = launch_gemm(s2, false, false, c, 1, launches, SHAPE_T))) return r;
    }
    if ((r = leaf(s2, j + 1, Sj, 128))) return r;
    if (rem > 1) {
      {  // panel j+1: rows of tiles j+2.. of S times W_{j+1}^T -> A
        const double* wjj = W + (long)(j + 1) * 128 * (ldw + 1);
        double* out = A + (long)(j + 2) * 128 * lda + (long)(j + 1) * 128;
        GemmDesc p = make_desc(Sj + (long)128 * 128, 128, wjj, ldw, out, lda, 2 * (rem - 1), 4, 128);
        if ((r = launch_gemm(s2, false, false, p, 1, launches, SHAPE_T))) return r;
      }
      CU(cudaEventRecord(evPanel[j + 1], s2));
      // ---- main stream: the rest of the trailing triangle, the whole group of panels together ----
      if (bulk) {
        CU(cudaStreamWaitEvent(s, evPanel[j], 0));
        double* pn2 = pn + (long)128 * lda;  // pending panel(s): rows from tile j+2
        double* trail = A + (long)(j + 2) * 128 * (lda + 1);
        GemmDesc u = make_desc(pn2, lda, pn2, lda, trail, lda, rem - 1, 2 * (rem - 1), kp);
        u.tri = 1, u.alpha = -1.0, u.beta = 1.0;
        if ((r = launch_gemm(s, false, false, u, 1, launches, SHAPE_S))) return r;
        CU(cudaEventRecord(evRest[j], s));
      }
    }
  }
  CU(cudaEventRecord(evJoin, s2));
  CU(cudaStreamWaitEvent(s, evJoin, 0));
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
