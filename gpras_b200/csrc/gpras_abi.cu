// C ABI of gpras_b200 (see include/gpras_b200.h): host-side orchestration of the sm_100a kernels.
// No CPU compute path exists in this file: every entry point either launches CUDA work or fails.
#include "../../include/gpras_b200.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gemm_engine.cuh"
#include "gp_kernels.cuh"
#include "leaf.cuh"

using namespace gpras;

namespace {

thread_local std::string g_err;

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
bool g_prepared[64] = {};

template <int KID>
int prepare_kid() {
  int r;
  if ((r = opt_in_smem(cov_kernel<KID>, 200 * 1024))) return r;
  if ((r = opt_in_smem(grad_kernel<KID, 8>, 200 * 1024))) return r;
  if ((r = opt_in_smem(grad_kernel<KID, 16>, 200 * 1024))) return r;
  if ((r = opt_in_smem(grad_kernel<KID, 32>, 200 * 1024))) return r;
  if ((r = opt_in_smem(grad_kernel<KID, 64>, 200 * 1024))) return r;
  return 0;
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

enum TileShape { SHAPE_L = 0, SHAPE_S = 1, SHAPE_N = 2, SHAPE_P = 3 };

template <typename Cfg, bool AKM, bool BKM>
int launch_cfg(cudaStream_t s, const GemmDesc& d, int batch, int nz) {
  constexpr int ratio = Cfg::BM / Cfg::BN;
  long tiles = d.tri ? (long)ratio * d.m_tiles * (d.m_tiles + 1) / 2 : (long)d.m_tiles * d.n_tiles;
  dim3 grid((unsigned)tiles, (unsigned)batch, (unsigned)nz);
  gemm_tile_kernel<Cfg, AKM, BKM><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(d);
  CU(cudaGetLastError());
  return 0;
}

// m_tiles counts BM-row tiles (128; 64 for SHAPE_P); n_tiles counts BN-column tiles of the chosen shape.
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

// ---- dense building blocks -------------------------------------------------------------------
// Side stream + events for the one-panel look-ahead of the Cholesky.
struct LookAhead {
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> ev;
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
  }
};

// Right-looking blocked Cholesky, panel width 128, with one panel of look-ahead on a high-priority side stream:
//   leaf   (1 CTA)   potrf + inverse of the diagonal block  -> L_jj, W_jj
//   panel  (CfgP)    L21 = A21 W_jj^T, in place
//   col    (CfgS)    block column j+1 of the trailing matrix -= panel_j panel_j^T        [side stream]
//   rest   (CfgS)    the trailing triangle right of block column j+1 -= panel_j panel_j^T  [main stream]
// so leaf(j+1) and panel(j+1) run while rest(j) occupies the machine.
int potrf_impl(cudaStream_t s, LookAhead& la, double* A, long lda, double* W, long ldw, int n, double* logdet_parts,
               int* info, int* launches) {
  const int nt = n / 128;
  int r;
  if ((r = la.ensure(2 * (size_t)nt + 2))) return r;
  cudaStream_t s2 = la.side;
  cudaEvent_t* evPanel = la.ev.data();        // [nt]
  cudaEvent_t* evRest = la.ev.data() + nt;    // [nt]
  cudaEvent_t evFork = la.ev[2 * nt], evJoin = la.ev[2 * nt + 1];
  auto leaf = [&](cudaStream_t st, int jb) -> int {
    leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, st>>>(A, lda, W, ldw, logdet_parts, info, jb);
    if (launches) ++*launches;
    CU(cudaGetLastError());
    return 0;
  };
  auto panel = [&](cudaStream_t st, int jb) -> int {  // rows of tiles jb+1.. , columns of block jb
    const int rem = nt - jb - 1;
    double* pn = A + (long)(jb + 1) * 128 * lda + (long)jb * 128;
    const double* wjj = W + (long)jb * 128 * ldw + (long)jb * 128;
    GemmDesc p = make_desc(pn, lda, wjj, ldw, pn, lda, 2 * rem, 1, 128);
    return launch_gemm(st, false, false, p, 1, launches, SHAPE_P);
  };
  CU(cudaEventRecord(evFork, s));
  CU(cudaStreamWaitEvent(s2, evFork, 0));
  if ((r = leaf(s, 0))) return r;
  if (nt == 1) return 0;
  if ((r = panel(s, 0))) return r;
  CU(cudaEventRecord(evPanel[0], s));
  for (int j = 0; j + 1 < nt; j++) {
    const int rem = nt - j - 1;  // tiles below / right of block j
    double* pn = A + (long)(j + 1) * 128 * lda + (long)j * 128;  // panel j, rows from tile j+1
    // ---- side stream: block column j+1, then the next diagonal block and panel ----
    CU(cudaStreamWaitEvent(s2, evPanel[j], 0));
    if (j > 0) CU(cudaStreamWaitEvent(s2, evRest[j - 1], 0));
    {
      double* col = A + (long)(j + 1) * 128 * (lda + 1);
      GemmDesc c = make_desc(pn, lda, pn, lda, col, lda, rem, 2, 128);
      c.alpha = -1.0, c.beta = 1.0;
      if ((r = launch_gemm(s2, false, false, c, 1, launches, SHAPE_S))) return r;
    }
    if ((r = leaf(s2, j + 1))) return r;
    if (rem > 1) {
      if ((r = panel(s2, j + 1))) return r;
      CU(cudaEventRecord(evPanel[j + 1], s2));
      // ---- main stream: the rest of the trailing triangle ----
      if (j > 0) CU(cudaStreamWaitEvent(s, evPanel[j], 0));
      double* pn2 = pn + (long)128 * lda;  // panel j, rows from tile j+2
      double* trail = A + (long)(j + 2) * 128 * (lda + 1);
      GemmDesc u = make_desc(pn2, lda, pn2, lda, trail, lda, rem - 1, 2 * (rem - 1), 128);
      u.tri = 1, u.alpha = -1.0, u.beta = 1.0;
      if ((r = launch_gemm(s, false, false, u, 1, launches, SHAPE_S))) return r;
      CU(cudaEventRecord(evRest[j], s));
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
               const double* theta, double* out, long ldo, int square) {
  const int t1 = n1_pad / CT, t2 = n2_pad / CT;
  const int smem = 2 * D * CT_LD * (int)sizeof(double);
  const long tiles = square ? (long)t1 * (t1 + 1) / 2 : (long)t1 * t2;
  cov_kernel<KID><<<(unsigned)tiles, PT_THREADS, smem, s>>>(Xs1, n1, Xs2, n2, D, theta, out, ldo, t2, square, square);
  CU(cudaGetLastError());
  return 0;
}

int dispatch_cov(int kid, cudaStream_t s, const double* Xs1, int n1, int n1_pad, const double* Xs2, int n2, int n2_pad,
                 int D, const double* theta, double* out, long ldo, int square) {
  switch (kid) {
    case K_RBF: return launch_cov<K_RBF>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square);
    case K_MATERN12: return launch_cov<K_MATERN12>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square);
    case K_MATERN32: return launch_cov<K_MATERN32>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square);
    case K_MATERN52: return launch_cov<K_MATERN52>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square);
    case K_EXPONENTIAL:
      return launch_cov<K_EXPONENTIAL>(s, Xs1, n1, n1_pad, Xs2, n2, n2_pad, D, theta, out, ldo, square);
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

template <int KID>
int launch_grad(cudaStream_t s, const double* Xs, int n, int n_pad, int D, const double* Wt, long ldw, double* part,
                int ncols) {
  const int nt = n_pad / CT;
  const int tiles = nt * (nt + 1) / 2;
  const int smem = 2 * D * CT_LD * (int)sizeof(double);
  if (D <= 8)
    grad_kernel<KID, 8><<<tiles, PT_THREADS, smem, s>>>(Xs, n, D, Wt, ldw, part, ncols);
  else if (D <= 16)
    grad_kernel<KID, 16><<<tiles, PT_THREADS, smem, s>>>(Xs, n, D, Wt, ldw, part, ncols);
  else if (D <= 32)
    grad_kernel<KID, 32><<<tiles, PT_THREADS, smem, s>>>(Xs, n, D, Wt, ldw, part, ncols);
  else
    grad_kernel<KID, 64><<<tiles, PT_THREADS, smem, s>>>(Xs, n, D, Wt, ldw, part, ncols);
  CU(cudaGetLastError());
  return 0;
}

int dispatch_grad(int kid, cudaStream_t s, const double* Xs, int n, int n_pad, int D, const double* Wt, long ldw,
                  double* part, int ncols) {
  switch (kid) {
    case K_RBF: return launch_grad<K_RBF>(s, Xs, n, n_pad, D, Wt, ldw, part, ncols);
    case K_MATERN12: return launch_grad<K_MATERN12>(s, Xs, n, n_pad, D, Wt, ldw, part, ncols);
    case K_MATERN32: return launch_grad<K_MATERN32>(s, Xs, n, n_pad, D, Wt, ldw, part, ncols);
    case K_MATERN52: return launch_grad<K_MATERN52>(s, Xs, n, n_pad, D, Wt, ldw, part, ncols);
    case K_EXPONENTIAL: return launch_grad<K_EXPONENTIAL>(s, Xs, n, n_pad, D, Wt, ldw, part, ncols);
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

constexpr int USQ_PARTS = 296;
constexpr int PRED_TB = 2048;   // test rows per predict batch
constexpr int CELL_TB = 256;    // test rows per cell-expansion batch

}  // namespace

struct gpras_gp {
  int device = 0, kid = 0, n = 0, d = 0, p = 0, n_pad = 0, p_pad = 0, nt = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false, has_data = false, conditioned = false, pending = false, pending_grad = false;
  bool stage_timing = false;
  int launches = 0;
  // training state
  double *X = nullptr, *Xs = nullptr, *Y = nullptr, *K = nullptr, *W = nullptr, *Kinv = nullptr, *U = nullptr,
         *alpha = nullptr, *theta = nullptr, *logdet = nullptr, *gpart = nullptr, *gsum = nullptr, *usq = nullptr,
         *result = nullptr, *skinny = nullptr;
  int* info = nullptr;
  double *h_theta = nullptr, *h_result = nullptr;
  int* h_info = nullptr;
  // prediction state
  double *Xt = nullptr, *Xts = nullptr, *Ks = nullptr, *mean = nullptr, *vpart = nullptr, *var = nullptr,
         *varm = nullptr;
  // cell map
  int c = 0, c_pad = 0, p16 = 0;
  double *E1 = nullptr, *E2 = nullptr, *bias = nullptr, *zbias = nullptr, *ring_m = nullptr, *ring_v = nullptr;
  cudaEvent_t ev[8] = {};
  double stage_ms[7] = {};
  LookAhead la;
};

namespace {

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

__global__ void broadcast_var_kernel(const double* __restrict__ var, double* __restrict__ varm, int T, int P, long ld) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)T * ld) return;
  int c = (int)(e % ld);
  varm[e] = c < P ? var[e / ld] : 0.0;
}

void mark(gpras_gp* h, int i) {
  if (h->stage_timing) cudaEventRecord(h->ev[i], h->stream);
}

// cov -> potrf -> trtri -> U = W Y  (+ alpha, Kinv when want_grad / for prediction)
int factorise(gpras_gp* h, bool need_alpha, bool need_kinv) {
  cudaStream_t s = h->stream;
  const int n = h->n, n_pad = h->n_pad, D = h->d, nt = h->nt;
  const long ld = n_pad;
  int r;
  mark(h, 0);
  CU(cudaMemsetAsync(h->info, 0, sizeof(int), s));
  {
    long tot = (long)n_pad * D;
    scale_features_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->X, h->Xs, n, n_pad, D, h->theta);
    h->launches++;
    CU(cudaGetLastError());
  }
  if ((r = dispatch_cov(h->kid, s, h->Xs, n, n_pad, h->Xs, n, n_pad, D, h->theta, h->K, ld, 1))) return r;
  h->launches++;
  mark(h, 1);
  if ((r = potrf_impl(s, h->la, h->K, ld, h->W, ld, n_pad, h->logdet, h->info, &h->launches))) return r;
  mark(h, 2);
  if ((r = trtri_impl(s, h->K, ld, h->W, ld, h->Kinv, ld, n_pad, &h->launches))) return r;
  mark(h, 3);
  if (need_kinv) {
    if ((r = lauum_impl(s, h->W, ld, h->Kinv, ld, n_pad, &h->launches))) return r;
  }
  mark(h, 4);
  // U = W Y
  {
    GemmDesc g = make_desc(h->W, ld, h->Y, h->p_pad, h->U, h->p_pad, nt, h->p_pad / 32, n_pad);
    g.ke_mode = KE_TI;
    if ((r = launch_skinny(s, false, g, h->skinny, n_pad, &h->launches))) return r;
  }
  if (need_alpha) {
    GemmDesc g = make_desc(h->W, ld, h->U, h->p_pad, h->alpha, h->p_pad, nt, h->p_pad / 32, n_pad);
    g.kb_mode = KB_TI;
    if ((r = launch_skinny(s, true, g, h->skinny, n_pad, &h->launches))) return r;
  }
  mark(h, 5);
  return 0;
}

int enqueue_eval(gpras_gp* h, const double* theta, int want_grad) {
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  cudaStream_t s = h->stream;
  const int n = h->n, n_pad = h->n_pad, D = h->d, P = h->p;
  int r;
  h->launches = 0;
  h->conditioned = false;
  memcpy(h->h_theta, theta, sizeof(double) * (2 + D));
  CU(cudaMemcpyAsync(h->theta, h->h_theta, sizeof(double) * (2 + D), cudaMemcpyHostToDevice, s));
  if ((r = factorise(h, want_grad != 0, want_grad != 0))) return r;
  sumsq_partial_kernel<<<USQ_PARTS, 256, 0, s>>>(h->U, n_pad, P, h->p_pad, h->usq);
  h->launches++;
  CU(cudaGetLastError());
  if (want_grad) {
    const int ntile = h->nt * (h->nt + 1) / 2;
    // Wt = alpha alpha^T - P Kinv, in place over Kinv (lower tiles) on the DMMA engine
    GemmDesc gw = make_desc(h->alpha, h->p_pad, h->alpha, h->p_pad, h->Kinv, n_pad, h->nt, 2 * h->nt, round_up(P, 32));
    gw.tri = 1, gw.alpha = 1.0, gw.beta = -(double)P;
    if ((r = launch_gemm(s, false, false, gw, 1, &h->launches, SHAPE_S))) return r;
    if ((r = dispatch_grad(h->kid, s, h->Xs, n, n_pad, D, h->Kinv, n_pad, h->gpart, 2 + D))) return r;
    colsum_kernel<<<2 + D, 256, 0, s>>>(h->gpart, ntile, 2 + D, 2 + D, h->gsum);
    h->launches += 2;
    CU(cudaGetLastError());
  }
  finalize_kernel<<<1, 32, 0, s>>>(h->usq, USQ_PARTS, h->logdet, h->nt, h->gsum, h->theta, n, P, D, want_grad, h->result);
  h->launches++;
  CU(cudaGetLastError());
  mark(h, 6);
  CU(cudaMemcpyAsync(h->h_result, h->result, sizeof(double) * (3 + D), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h->h_info, h->info, sizeof(int), cudaMemcpyDeviceToHost, s));
  h->pending = true;
  h->pending_grad = want_grad != 0;
  return 0;
}

int fetch_eval(gpras_gp* h, double* lml, double* grad) {
  if (!h->pending) return fail(GPRAS_E_STATE, "no evaluation enqueued");
  CU(cudaStreamSynchronize(h->stream));
  h->pending = false;
  if (h->stage_timing) {
    for (int i = 0; i < 6; i++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
      h->stage_ms[i] = ms;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
    h->stage_ms[6] = ms;
  }
  if (*h->h_info != 0) {
    g_err = "covariance matrix is not positive definite";
    return *h->h_info;
  }
  if (lml) *lml = h->h_result[0];
  if (grad && h->pending_grad) memcpy(grad, h->h_result + 1, sizeof(double) * (2 + h->d));
  return 0;
}

}  // namespace

extern "C" {

int gpras_abi_version(void) { return GPRAS_B200_ABI_VERSION; }
const char* gpras_last_error(void) { return g_err.c_str(); }

int gpras_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int gpras_gp_create(gpras_gp** out, int device, int kernel_id, int n, int d, int p) {
  if (!out || n <= 0 || d <= 0 || p <= 0) return fail(GPRAS_E_ARG, "bad shape");
  if (kernel_id < 0 || kernel_id > 4) return fail(GPRAS_E_ARG, "unknown kernel id");
  if (d > 64) return fail(GPRAS_E_ARG, "d > 64 features is not supported");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  int r;
  if ((r = prepare_device())) return r;
  gpras_gp* h = new gpras_gp();
  h->device = device, h->kid = kernel_id, h->n = n, h->d = d, h->p = p;
  h->n_pad = round_up(n, 128), h->p_pad = round_up(p, 32), h->nt = h->n_pad / 128;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  const size_t nn = (size_t)h->n_pad * h->n_pad, np = (size_t)h->n_pad * h->p_pad;
  const int ntile = h->nt * (h->nt + 1) / 2;
  if ((r = dalloc(&h->X, (size_t)h->n_pad * d)) || (r = dalloc(&h->Xs, (size_t)h->n_pad * d)) || (r = dalloc(&h->Y, np)) ||
      (r = dalloc(&h->K, nn)) || (r = dalloc(&h->W, nn)) || (r = dalloc(&h->Kinv, nn)) || (r = dalloc(&h->U, np)) ||
      (r = dalloc(&h->alpha, np)) || (r = dalloc(&h->theta, 2 + d)) || (r = dalloc(&h->logdet, h->nt)) ||
      (r = dalloc(&h->gpart, (size_t)ntile * (2 + d))) || (r = dalloc(&h->gsum, 2 + d)) ||
      (r = dalloc(&h->usq, USQ_PARTS)) || (r = dalloc(&h->result, 3 + d)) ||
      (r = dalloc(&h->skinny, (size_t)SKINNY_MAX_SLABS * (h->n_pad > PRED_TB ? h->n_pad : PRED_TB) * h->p_pad))) {
    gpras_gp_destroy(h);
    return r;
  }
  CU(cudaMalloc((void**)&h->info, sizeof(int)));
  CU(cudaMallocHost((void**)&h->h_theta, sizeof(double) * (2 + d)));
  CU(cudaMallocHost((void**)&h->h_result, sizeof(double) * (3 + d)));
  CU(cudaMallocHost((void**)&h->h_info, sizeof(int)));
  // stream-ordered: the handle's stream is non-blocking, so legacy-stream memsets would race with set_data
  CU(cudaMemsetAsync(h->X, 0, sizeof(double) * h->n_pad * d, h->stream));
  CU(cudaMemsetAsync(h->Y, 0, sizeof(double) * np, h->stream));
  CU(cudaMemsetAsync(h->gsum, 0, sizeof(double) * (2 + d), h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (auto& e : h->ev) CU(cudaEventCreate(&e));
  *out = h;
  return 0;
}

int gpras_gp_destroy(gpras_gp* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  double* bufs[] = {h->X,    h->Xs,  h->Y,   h->K,  h->W,     h->Kinv, h->U,    h->alpha, h->theta, h->logdet, h->gpart,
                    h->gsum, h->usq, h->result, h->skinny, h->Xt, h->Xts, h->Ks,   h->mean, h->vpart, h->var,   h->varm,   h->E1,
                    h->E2,   h->bias, h->zbias, h->ring_m, h->ring_v};
  for (double* b : bufs)
    if (b) cudaFree(b);
  if (h->info) cudaFree(h->info);
  if (h->h_theta) cudaFreeHost(h->h_theta);
  if (h->h_result) cudaFreeHost(h->h_result);
  if (h->h_info) cudaFreeHost(h->h_info);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  h->la.destroy();
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int gpras_gp_set_stream(gpras_gp* h, void* cuda_stream) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  DeviceGuard guard(h->device);
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  if (cuda_stream) {
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
  } else {
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return 0;
}

int gpras_gp_set_data(gpras_gp* h, const double* x, const double* y, int on_device) {
  if (!h || !x || !y) return fail(GPRAS_E_ARG, "null argument");
  DeviceGuard guard(h->device);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CU(cudaMemcpyAsync(h->X, x, sizeof(double) * h->n * h->d, kind, h->stream));
  CU(cudaMemcpy2DAsync(h->Y, sizeof(double) * h->p_pad, y, sizeof(double) * h->p, sizeof(double) * h->p, h->n, kind,
                       h->stream));
  h->has_data = true;
  h->conditioned = false;
  return 0;
}

int gpras_gp_lml_grad_enqueue(gpras_gp* h, const double* theta, int want_grad) {
  if (!h || !theta) return fail(GPRAS_E_ARG, "null argument");
  DeviceGuard guard(h->device);
  return enqueue_eval(h, theta, want_grad);
}

int gpras_gp_lml_grad_fetch(gpras_gp* h, double* lml, double* grad) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  DeviceGuard guard(h->device);
  return fetch_eval(h, lml, grad);
}

int gpras_gp_lml_grad(gpras_gp* h, const double* theta, double* lml, double* grad) {
  if (!h || !theta) return fail(GPRAS_E_ARG, "null argument");
  DeviceGuard guard(h->device);
  int r = enqueue_eval(h, theta, grad != nullptr);
  if (r) return r;
  return fetch_eval(h, lml, grad);
}

int gpras_gp_lml_grad_host(gpras_gp* h, const double* x, const double* y, const double* theta, double* lml,
                           double* grad) {
  int r = gpras_gp_set_data(h, x, y, 0);
  if (r) return r;
  return gpras_gp_lml_grad(h, theta, lml, grad);
}

int gpras_gp_condition(gpras_gp* h, const double* theta) {
  if (!h || !theta) return fail(GPRAS_E_ARG, "null argument");
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  DeviceGuard guard(h->device);
  h->launches = 0;
  memcpy(h->h_theta, theta, sizeof(double) * (2 + h->d));
  CU(cudaMemcpyAsync(h->theta, h->h_theta, sizeof(double) * (2 + h->d), cudaMemcpyHostToDevice, h->stream));
  int r = factorise(h, true, false);
  if (r) return r;
  CU(cudaMemcpyAsync(h->h_info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (*h->h_info != 0) {
    g_err = "covariance matrix is not positive definite";
    return *h->h_info;
  }
  h->conditioned = true;
  return 0;
}

static int ensure_predict_buffers(gpras_gp* h) {
  if (h->Xt) return 0;
  int r;
  if ((r = dalloc(&h->Xt, (size_t)PRED_TB * h->d)) || (r = dalloc(&h->Xts, (size_t)PRED_TB * h->d)) ||
      (r = dalloc(&h->Ks, (size_t)PRED_TB * h->n_pad)) || (r = dalloc(&h->mean, (size_t)PRED_TB * h->p_pad)) ||
      (r = dalloc(&h->vpart, (size_t)h->nt * PRED_TB)) || (r = dalloc(&h->var, PRED_TB)) ||
      (r = dalloc(&h->varm, (size_t)PRED_TB * h->p_pad)))
    return r;
  return 0;
}

// One batch of <= PRED_TB test rows already staged in h->Xt (tb_pad rows, zero padded):
// leaves mean (tb_pad x p_pad) in h->mean, var (tb_pad) in h->var.
static int predict_batch(gpras_gp* h, int tb, int tb_pad) {
  cudaStream_t s = h->stream;
  const int n_pad = h->n_pad, D = h->d;
  int r;
  long tot = (long)tb_pad * D;
  scale_features_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->Xt, h->Xts, tb, tb_pad, D, h->theta);
  CU(cudaGetLastError());
  if ((r = dispatch_cov(h->kid, s, h->Xts, tb, tb_pad, h->Xs, h->n, n_pad, D, h->theta, h->Ks, n_pad, 0))) return r;
  h->launches += 2;
  // mean = Ks alpha
  GemmDesc gm = make_desc(h->Ks, n_pad, h->alpha, h->p_pad, h->mean, h->p_pad, tb_pad / 128, h->p_pad / 32, n_pad);
  if ((r = launch_skinny(s, false, gm, h->skinny, tb_pad, &h->launches))) return r;
  // |W ks|^2 : tiles of V = W Ks^T reduced on the fly to column sums of squares
  GemmDesc gv = make_desc(h->W, n_pad, h->Ks, n_pad, h->vpart, PRED_TB, h->nt, tb_pad / 128, n_pad);
  gv.ke_mode = KE_TI;
  gv.reverse = 1;
  gv.epilogue = EPI_COLSUMSQ;
  if ((r = launch_gemm(s, false, false, gv, 1, &h->launches))) return r;
  predict_var_kernel<<<(tb_pad + 127) / 128, 128, 0, s>>>(h->vpart, h->nt, tb_pad, PRED_TB, h->theta, h->var);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

static int stage_test_rows(gpras_gp* h, const double* xs, int t0, int tb, int tb_pad, int on_device) {
  cudaStream_t s = h->stream;
  CU(cudaMemsetAsync(h->Xt, 0, sizeof(double) * (size_t)tb_pad * h->d, s));
  CU(cudaMemcpyAsync(h->Xt, xs + (size_t)t0 * h->d, sizeof(double) * (size_t)tb * h->d,
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  return 0;
}

int gpras_gp_predict(gpras_gp* h, const double* xs, int t, double* mean, double* var, int on_device) {
  if (!h || !xs || t < 0) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->conditioned) return fail(GPRAS_E_STATE, "condition() has not been called");
  DeviceGuard guard(h->device);
  int r;
  if ((r = ensure_predict_buffers(h))) return r;
  cudaStream_t s = h->stream;
  h->launches = 0;
  const cudaMemcpyKind back = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  for (int t0 = 0; t0 < t; t0 += PRED_TB) {
    const int tb = t - t0 < PRED_TB ? t - t0 : PRED_TB;
    const int tb_pad = round_up(tb, 128);
    if ((r = stage_test_rows(h, xs, t0, tb, tb_pad, on_device))) return r;
    if ((r = predict_batch(h, tb, tb_pad))) return r;
    if (mean)
      CU(cudaMemcpy2DAsync(mean + (size_t)t0 * h->p, sizeof(double) * h->p, h->mean, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, back, s));
    if (var) {
      long tot = (long)tb_pad * h->p_pad;
      broadcast_var_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->var, h->varm, tb_pad, h->p, h->p_pad);
      h->launches++;
      CU(cudaGetLastError());
      CU(cudaMemcpy2DAsync(var + (size_t)t0 * h->p, sizeof(double) * h->p, h->varm, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, back, s));
    }
    if (!on_device) CU(cudaStreamSynchronize(s));  // pageable host buffers: keep batches ordered
  }
  CU(cudaStreamSynchronize(s));
  return 0;
}

int gpras_gp_set_cell_map(gpras_gp* h, const double* e_mean, const double* bias, int c) {
  if (!h || !e_mean || !bias || c <= 0) return fail(GPRAS_E_ARG, "bad argument");
  DeviceGuard guard(h->device);
  double* olds[] = {h->E1, h->E2, h->bias, h->zbias, h->ring_m, h->ring_v};
  for (double* b : olds)
    if (b) cudaFree(b);
  h->E1 = h->E2 = h->bias = h->zbias = h->ring_m = h->ring_v = nullptr;
  h->c = c, h->c_pad = round_up(c, 128), h->p16 = round_up(h->p, 32);
  const size_t ne = (size_t)h->p16 * h->c_pad;
  int r;
  if ((r = dalloc(&h->E1, ne)) || (r = dalloc(&h->E2, ne)) || (r = dalloc(&h->bias, h->c_pad)) ||
      (r = dalloc(&h->zbias, h->c_pad)))
    return r;
  std::vector<double> e1(ne, 0.0), e2(ne, 0.0), b(h->c_pad, 0.0);
  for (int pp = 0; pp < h->p; pp++)
    for (int cc = 0; cc < c; cc++) {
      double v = e_mean[(size_t)pp * c + cc];
      e1[(size_t)pp * h->c_pad + cc] = v;
      e2[(size_t)pp * h->c_pad + cc] = v * v;
    }
  memcpy(b.data(), bias, sizeof(double) * c);
  CU(cudaMemcpyAsync(h->E1, e1.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->E2, e2.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->bias, b.data(), sizeof(double) * h->c_pad, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(h->zbias, 0, sizeof(double) * h->c_pad, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

long gpras_gp_cell_pitch(gpras_gp* h) { return h ? h->c_pad : 0; }

int gpras_gp_predict_cells(gpras_gp* h, const double* xs, int t, int xs_on_device, double* mode_mean, double* mode_var,
                           double* cell_mean, double* cell_var, long ldc) {
  if (!h || !xs || t < 0) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->conditioned) return fail(GPRAS_E_STATE, "condition() has not been called");
  if (!h->E1) return fail(GPRAS_E_STATE, "set_cell_map() has not been called");
  if ((cell_mean || cell_var) && ldc < h->c_pad) return fail(GPRAS_E_ARG, "ldc smaller than gpras_gp_cell_pitch()");
  DeviceGuard guard(h->device);
  int r;
  if ((r = ensure_predict_buffers(h))) return r;
  if (!h->ring_m) {
    if ((r = dalloc(&h->ring_m, (size_t)CELL_TB * h->c_pad)) || (r = dalloc(&h->ring_v, (size_t)CELL_TB * h->c_pad)))
      return r;
  }
  cudaStream_t s = h->stream;
  h->launches = 0;
  for (int t0 = 0; t0 < t; t0 += PRED_TB) {
    const int tb = t - t0 < PRED_TB ? t - t0 : PRED_TB;
    const int tb_pad = round_up(tb, 128);
    if ((r = stage_test_rows(h, xs, t0, tb, tb_pad, xs_on_device))) return r;
    if ((r = predict_batch(h, tb, tb_pad))) return r;
    long tot = (long)tb_pad * h->p_pad;
    broadcast_var_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->var, h->varm, tb_pad, h->p, h->p_pad);
    h->launches++;
    CU(cudaGetLastError());
    if (mode_mean)
      CU(cudaMemcpy2DAsync(mode_mean + (size_t)t0 * h->p, sizeof(double) * h->p, h->mean, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, cudaMemcpyDeviceToHost, s));
    if (mode_var)
      CU(cudaMemcpy2DAsync(mode_var + (size_t)t0 * h->p, sizeof(double) * h->p, h->varm, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, cudaMemcpyDeviceToHost, s));
    // modes -> cells, CELL_TB rows at a time: cells = modes @ E (+ bias), one DMMA GEMM each for mean and variance
    for (int c0 = 0; c0 < tb_pad; c0 += CELL_TB) {
      const int cb = tb_pad - c0 < CELL_TB ? tb_pad - c0 : CELL_TB;
      double* om = cell_mean ? cell_mean + (size_t)(t0 + c0) * ldc : h->ring_m;
      double* ov = cell_var ? cell_var + (size_t)(t0 + c0) * ldc : h->ring_v;
      const long ldo_m = cell_mean ? ldc : h->c_pad, ldo_v = cell_var ? ldc : h->c_pad;
      GemmDesc gm = make_desc(h->mean + (size_t)c0 * h->p_pad, h->p_pad, h->E1, h->c_pad, om, ldo_m, cb / 128,
                              h->c_pad / 128, h->p16);
      gm.epilogue = EPI_BIAS, gm.bias = h->bias;
      if ((r = launch_gemm(s, false, true, gm, 1, &h->launches))) return r;
      GemmDesc gv = make_desc(h->varm + (size_t)c0 * h->p_pad, h->p_pad, h->E2, h->c_pad, ov, ldo_v, cb / 128,
                              h->c_pad / 128, h->p16);
      gv.epilogue = EPI_BIAS, gv.bias = h->zbias;
      if ((r = launch_gemm(s, false, true, gv, 1, &h->launches))) return r;
    }
    if (!xs_on_device) CU(cudaStreamSynchronize(s));
  }
  CU(cudaStreamSynchronize(s));
  return 0;
}

int gpras_gp_get_matrix(gpras_gp* h, int which, double* out) {
  if (!h || !out) return fail(GPRAS_E_ARG, "null argument");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  const double* src = nullptr;
  switch (which) {
    case 0:
    case 1: src = h->K; break;
    case 2: src = h->W; break;
    case 3: src = h->Kinv; break;
    case 4:
      CU(cudaMemcpy2D(out, sizeof(double) * h->p, h->alpha, sizeof(double) * h->p_pad, sizeof(double) * h->p, h->n,
                      cudaMemcpyDeviceToHost));
      return 0;
    default: return fail(GPRAS_E_ARG, "which out of range");
  }
  CU(cudaMemcpy2D(out, sizeof(double) * h->n, src, sizeof(double) * h->n_pad, sizeof(double) * h->n, h->n,
                  cudaMemcpyDeviceToHost));
  return 0;
}

int gpras_gp_last_launches(gpras_gp* h) { return h ? h->launches : 0; }

int gpras_gp_set_stage_timing(gpras_gp* h, int enabled) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  h->stage_timing = enabled != 0;
  return 0;
}

int gpras_gp_last_stage_ms(gpras_gp* h, double* ms7) {
  if (!h || !ms7) return fail(GPRAS_E_ARG, "null argument");
  memcpy(ms7, h->stage_ms, sizeof h->stage_ms);
  return 0;
}

// ---- stand-alone building blocks --------------------------------------------------------------
int gpras_dgemm_tiles(void* cuda_stream, int shape, int a_kmajor, int b_kmajor, const double* A, long lda, const double* B,
                      long ldb, double* C, long ldc, int m, int n, int k, double alpha, double beta) {
  const int bn = shape == SHAPE_L ? 128 : (shape == SHAPE_S ? 64 : 32);
  if (shape < 0 || shape > 2) return fail(GPRAS_E_ARG, "shape must be 0 (128x128), 1 (128x64) or 2 (128x32)");
  if (m % 128 || n % bn || k % 32 || m <= 0 || n <= 0 || k <= 0) return fail(GPRAS_E_ARG, "extents must be tile multiples");
  if (gpras_device_count() <= 0) return fail(GPRAS_E_CUDA, "no CUDA device (no CPU fallback)");
  int r;
  if ((r = prepare_device())) return r;
  GemmDesc d = make_desc(A, lda, B, ldb, C, ldc, m / 128, n / bn, k);
  d.alpha = alpha, d.beta = beta;
  return launch_gemm((cudaStream_t)cuda_stream, a_kmajor != 0, b_kmajor != 0, d, 1, nullptr, shape);
}

int gpras_dpotrf(void* cuda_stream, double* A, long lda, double* W, long ldw, int n, double* logdet_parts_dev,
                 int* info_dev) {
  if (n % 128 || n <= 0) return fail(GPRAS_E_ARG, "n must be a positive multiple of 128");
  if (gpras_device_count() <= 0) return fail(GPRAS_E_CUDA, "no CUDA device (no CPU fallback)");
  int r;
  if ((r = prepare_device())) return r;
  static thread_local LookAhead la;  // stand-alone entry: one side stream per calling thread
  return potrf_impl((cudaStream_t)cuda_stream, la, A, lda, W, ldw, n, logdet_parts_dev, info_dev, nullptr);
}

int gpras_dtrtri(void* cuda_stream, const double* L, long ldl, double* W, long ldw, double* scratch, long lds, int n) {
  if (n % 128 || n <= 0) return fail(GPRAS_E_ARG, "n must be a positive multiple of 128");
  if (gpras_device_count() <= 0) return fail(GPRAS_E_CUDA, "no CUDA device (no CPU fallback)");
  int r;
  if ((r = prepare_device())) return r;
  return trtri_impl((cudaStream_t)cuda_stream, L, ldl, W, ldw, scratch, lds, n, nullptr);
}

int gpras_dlauum(void* cuda_stream, const double* W, long ldw, double* Kinv, long ldk, int n) {
  if (n % 128 || n <= 0) return fail(GPRAS_E_ARG, "n must be a positive multiple of 128");
  if (gpras_device_count() <= 0) return fail(GPRAS_E_CUDA, "no CUDA device (no CPU fallback)");
  int r;
  if ((r = prepare_device())) return r;
  return lauum_impl((cudaStream_t)cuda_stream, W, ldw, Kinv, ldk, n, nullptr);
}

}  // extern "C"
