// C ABI of the PreProcessor mirror (include/gpras_b200.h, "PreProcessor" section): cells <-> modes on the device.
// Replaces the NumPy / scikit-learn arithmetic of gpras/preprocess.py:947-1094.  No CPU compute path exists here.
#include <chrono>
#include <cmath>
#include <cstdlib>

#include "host_common.cuh"
#include "eig_kernels.cuh"
#include "pre_kernels.cuh"
#include "cells_kernel.cuh"

namespace {

constexpr int PRE_MAX_MODES = 64;
constexpr int PRE_GRAM_MAX_N = 2048;  // above this many (padded) samples the PCA runs Gram-free
constexpr int PRE_REV_TB = 1024;  // rows per reverse-transform block

// wet cells: Ef[p][c] = x_std[p] E[p][c] / w[c], Ef2 = Ef^2, bias[c] = sum_p x_mean[p] E[p][c] / w[c] + mean[c];
// dry / padded cells: Ef = 0, bias = elevation (0 for "depth" and for padding)      (preprocess.py:1069-1094)
__global__ void fold_map_kernel(const double* __restrict__ E, long c_pad, int c, int p, int pk, const int* __restrict__ cls,
                                const double* __restrict__ wfull, const double* __restrict__ mean, const double* __restrict__ elev,
                                const double* __restrict__ x_mean, const double* __restrict__ x_std, int depth,
                                double* __restrict__ Ef, double* __restrict__ Ef2, double* __restrict__ bias) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c_pad) return;
  const bool wet = j < c && cls[j] != 1;
  double b = 0.0;
  for (int q = 0; q < pk; q++) {
    double v = 0.0;
    if (wet && q < p) {
      const double e = E[(long)q * c_pad + j] / wfull[j];
      v = x_std[q] * e;
      b += x_mean[q] * e;
    }
    Ef[(long)q * c_pad + j] = v;
    Ef2[(long)q * c_pad + j] = v * v;
  }
  bias[j] = wet ? b + mean[j] : ((j < c && !depth) ? elev[j] : 0.0);
}

// resid[j] = (lambda_j / lambda_0) * |Z[:, j] - U[:, j]|   (residual of Ritz pair j relative to the top eigenvalue)
__global__ void __launch_bounds__(256) ritz_residual_kernel(const double* __restrict__ Z, const double* __restrict__ U, int rows,
                                                            const double* __restrict__ lambda, double* __restrict__ out) {
  __shared__ double red[8];
  const int j = blockIdx.x, tid = threadIdx.x;
  double s = 0.0;
  for (int i = tid; i < rows; i += 256) {
    const double v = Z[(long)i * EIG_B + j] - U[(long)i * EIG_B + j];
    s += v * v;
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; i++) t += red[i];
    out[j] = lambda[0] > 0.0 ? lambda[j] / lambda[0] * sqrt(t) : 0.0;
  }
}

template <int PN>
int launch_project(cudaStream_t s, const double* X, long ldx, int n, int c, const double* elev, int clamp, const double* mean,
                   const double* wfull, const double* E, long c_pad, double* part, int* nz_out) {
  const int n_pad = round_up(n, 128);
  const int row_tiles = n_pad / 128;
  const int k_stages = (int)(c_pad / PROJ_BK);
  // Split the cells so that the grid is a whole number of waves of 2 CTAs per SM (a partial last wave costs a full CTA time:
  // measured 4.4 -> 5.0 TB/s at 8192 x 200 000, 16 modes, from fixing this alone), as many waves as keep >= 40 stages per CTA.
  static const int waves_env = getenv("GPRAS_B200_PROJ_WAVES") ? atoi(getenv("GPRAS_B200_PROJ_WAVES")) : 0;
  int nz = 1;
  for (int w = 1; w <= 8; w++) {
    const int cand = (2 * 148 * w) / row_tiles;
    if (cand < 1) continue;
    if (waves_env ? w == waves_env : (k_stages / cand >= 40 || nz == 1)) nz = cand;
  }
  if (nz > (k_stages + 31) / 32) nz = (k_stages + 31) / 32;
  if (nz < 1) nz = 1;
  const int per = (k_stages + nz - 1) / nz;
  nz = (k_stages + per - 1) / per;
  // The x stream goes through the TMA unit whenever a tensor map can describe it (16-byte aligned rows); otherwise the
  // cp.async variant of the same kernel runs.  GPRAS_B200_NO_TMA=1 forces the latter (development comparisons).
  CUtensorMap xmap;
  static const bool no_tma = getenv("GPRAS_B200_NO_TMA") != nullptr;
  if (!no_tma && tma_map_2d_f64(&xmap, X, (uint64_t)c, (uint64_t)n, (uint64_t)ldx, PROJ_BK, 128, true)) {
    project_tma_kernel<PN><<<dim3(row_tiles, nz), PROJ_THREADS, ProjTmaCfg<PN>::SMEM_BYTES, s>>>(xmap, n, c, elev, clamp, mean, wfull, E,
                                                                                               c_pad, k_stages, per, part, n_pad);
  } else {
    project_kernel<PN><<<dim3(row_tiles, nz), PROJ_THREADS, ProjCfg<PN>::SMEM_BYTES, s>>>(X, ldx, n, c, elev, clamp, mean, wfull, E,
                                                                                        c_pad, k_stages, per, part, n_pad);
  }
  CU(cudaGetLastError());
  *nz_out = nz;
  return 0;
}

int project_splits_max(int n, long c_pad) {
  const int row_tiles = round_up(n, 128) / 128;
  const int k_stages = (int)(c_pad / PROJ_BK);
  int nz = (2 * 16 * 148 + row_tiles - 1) / row_tiles;  // upper bound over every GPRAS_B200_PROJ_WAVES setting
  if (nz > (k_stages + 31) / 32) nz = (k_stages + 31) / 32;
  return nz < 1 ? 1 : nz;
}

}  // namespace

// Workspace pool: cudaMalloc / cudaFree cost hundreds of milliseconds next to multi-GB live allocations (measured: 293 ms
// for the 160 MB of the subspace iteration at config-3 sizes), so workspaces are recycled across calls and only returned
// to the driver by gpras_pre_trim / gpras_pre_destroy.
struct PoolBlock {
  void* ptr;
  size_t bytes;
  bool in_use;
};

struct gpras_pre {
  std::vector<PoolBlock> pool;
  int device = 0, c = 0, hp = 0;
  long c_pad = 0;
  double wet_threshold = 0.03;
  cudaStream_t stream = nullptr;
  bool fitted = false, map_ready = false;
  int p = 0, n_fit = 0, n_eig = 0, iters = 0, launches = 0;
  double *elev = nullptr, *weights_in = nullptr, *mean = nullptr, *wfull = nullptr, *E = nullptr, *x_mean = nullptr,
         *x_std = nullptr, *lambda = nullptr, *resid = nullptr, *Ef = nullptr, *Ef2 = nullptr, *bias = nullptr;
  int* cls = nullptr;
  double h_lambda[EIG_B] = {}, h_resid[EIG_B] = {};
  double stage_ms[7] = {};
  cudaEvent_t ev[8] = {};
};

namespace {

int pool_take(gpras_pre* h, void** out, size_t bytes) {
  if (bytes == 0) bytes = 8;
  int best = -1;
  for (int i = 0; i < (int)h->pool.size(); i++) {
    const PoolBlock& b = h->pool[i];
    if (!b.in_use && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096 && (best < 0 || b.bytes < h->pool[best].bytes)) best = i;
  }
  if (best >= 0) {
    h->pool[best].in_use = true;
    *out = h->pool[best].ptr;
    return 0;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    // give unused blocks back to the driver and retry once
    for (auto it = h->pool.begin(); it != h->pool.end();)
      if (!it->in_use) {
        cudaFree(it->ptr);
        it = h->pool.erase(it);
      } else {
        ++it;
      }
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(GPRAS_E_NOMEM, "cudaMalloc (workspace)", e);
  }
  h->pool.push_back({p, bytes, true});
  *out = p;
  return 0;
}

int palloc(gpras_pre* h, double** out, size_t count) { return pool_take(h, (void**)out, count * sizeof(double)); }

void pfree(gpras_pre* h, void* p) {
  if (!p) return;
  for (auto& b : h->pool)
    if (b.ptr == p) b.in_use = false;
}

void pool_release(gpras_pre* h) {
  for (auto& b : h->pool) cudaFree(b.ptr);
  h->pool.clear();
}

int pn_of(int p) { return p <= 8 ? 8 : (p <= 16 ? 16 : (p <= 32 ? 32 : 64)); }

// Device copy of a (n x c) matrix with a 16-byte aligned pitch of c_pad columns; returns the input itself when it already
// is one.  *owned receives the buffer to free (or NULL).
int stage_matrix(gpras_pre* h, const double* x, long ldx, int n, int on_device, const double** xd, long* ldd, double** owned) {
  *owned = nullptr;
  if (on_device && ldx >= h->c_pad && ldx % 2 == 0 && ((uintptr_t)x % 16) == 0) {
    *xd = x, *ldd = ldx;
    return 0;
  }
  int r;
  if ((r = palloc(h, owned, (size_t)n * h->c_pad))) return r;
  cudaError_t e = cudaMemcpy2DAsync(*owned, sizeof(double) * h->c_pad, x, sizeof(double) * ldx, sizeof(double) * h->c, n,
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) {
    pfree(h, *owned);
    *owned = nullptr;
    return fail(GPRAS_E_CUDA, "staging copy of the input matrix", e);
  }
  *xd = *owned, *ldd = h->c_pad;
  return 0;
}

// scores (n x pn, raw or standardised) of the staged matrix; Z is a device buffer of n x ldz
int run_project(gpras_pre* h, const double* xd, long ldd, int n, int pn, int p, int standardise, double* Z, long ldz) {
  cudaStream_t s = h->stream;
  const int n_pad = round_up(n, 128);
  double* part = nullptr;
  int r, nz = 0;
  if ((r = palloc(h, &part, (size_t)project_splits_max(n, h->c_pad) * n_pad * pn))) return r;
  const int clamp = h->hp == HP_DEPTH;
  switch (pn) {
    case 8: r = launch_project<8>(s, xd, ldd, n, h->c, h->elev, clamp, h->mean, h->wfull, h->E, h->c_pad, part, &nz); break;
    case 16: r = launch_project<16>(s, xd, ldd, n, h->c, h->elev, clamp, h->mean, h->wfull, h->E, h->c_pad, part, &nz); break;
    case 32: r = launch_project<32>(s, xd, ldd, n, h->c, h->elev, clamp, h->mean, h->wfull, h->E, h->c_pad, part, &nz); break;
    default: r = launch_project<64>(s, xd, ldd, n, h->c, h->elev, clamp, h->mean, h->wfull, h->E, h->c_pad, part, &nz); break;
  }
  if (!r) {
    const long tot = (long)n * pn;
    project_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(part, nz, n_pad, pn, n, p, h->x_mean, h->x_std, standardise, Z,
                                                                      ldz);
    if (cudaGetLastError() != cudaSuccess) r = fail(GPRAS_E_CUDA, "project_finish_kernel");
    h->launches += 2;
  }
  cudaStreamSynchronize(s);
  pfree(h, part);
  return r;
}

int build_map(gpras_pre* h) {
  const int pk = round_up(h->p, 32);
  fold_map_kernel<<<(unsigned)((h->c_pad + 255) / 256), 256, 0, h->stream>>>(h->E, h->c_pad, h->c, h->p, pk, h->cls, h->wfull, h->mean,
                                                                            h->elev, h->x_mean, h->x_std, h->hp == HP_DEPTH, h->Ef,
                                                                            h->Ef2, h->bias);
  CU(cudaGetLastError());
  h->launches++;
  h->map_ready = true;
  return 0;
}

// Leading eigenpairs of the symmetric PSD matrix G = Xw Xw^T by blocked subspace iteration with a Rayleigh-Ritz step per
// iteration.  G != NULL: explicit Gram matrix (n_pad x n_pad, full storage); G == NULL: operator form, Y = Xw (Xw^T Q) as
// two skinny products per iteration -- 4 n c 128 flop instead of the n^2 c of forming G, which wins once n exceeds a few
// thousand samples, and never squares the data in memory.  On return U (n_pad x 128) holds the Ritz vectors,
// h->h_lambda / h_resid the values.
int subspace_eig(gpras_pre* h, const double* G, const double* Xw, long c_pad, int n, int n_pad, int kconv, double tol,
                 int max_iter, double* U) {
  cudaStream_t s = h->stream;
  const bool trace = getenv("GPRAS_B200_TRACE") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(s);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[subspace_eig] %-14s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
    t_start = now;
  };
  stamp("enter");
  const size_t nb = (size_t)n_pad * EIG_B, bb = (size_t)EIG_B * EIG_B;
  double *Q = nullptr, *Y = nullptr, *Z = nullptr, *small = nullptr, *skinny = nullptr, *logdet = nullptr, *Tc = nullptr;
  int* info = nullptr;
  int r = 0;
  const int mt = n_pad / 128;
  // operator form: split of the long k = c_pad range of Y = Xw T into nz chunks (mt * nz CTAs ~ 3 waves)
  int nz_op = (3 * 148 + mt - 1) / mt;
  if (nz_op > (int)(c_pad / 512)) nz_op = (int)(c_pad / 512);
  if (nz_op < 1) nz_op = 1;
  const int ks_op = round_up((int)((c_pad + nz_op - 1) / nz_op), 128);
  nz_op = (int)((c_pad + ks_op - 1) / ks_op);
  const int n_slabs = G ? SKINNY_MAX_SLABS : (nz_op > SKINNY_MAX_SLABS ? nz_op : SKINNY_MAX_SLABS);
  auto cleanup = [&]() {
    cudaStreamSynchronize(s);
    pfree(h, Q), pfree(h, Y), pfree(h, Z), pfree(h, small), pfree(h, skinny), pfree(h, logdet), pfree(h, info), pfree(h, Tc);
  };
  if ((r = palloc(h, &Q, nb)) || (r = palloc(h, &Y, nb)) || (r = palloc(h, &Z, nb)) || (r = palloc(h, &small, 6 * bb)) ||
      (r = palloc(h, &skinny, (size_t)n_slabs * nb)) || (r = palloc(h, &logdet, 1)) ||
      (!G && (r = palloc(h, &Tc, (size_t)c_pad * EIG_B))) ||
      (r = pool_take(h, (void**)&info, sizeof(int)))) {
    cleanup();
    return r;
  }
  stamp("alloc");
  double *H = small, *B = small + bb, *V = small + 2 * bb, *Vs = small + 3 * bb, *L = small + 4 * bb, *W = small + 5 * bb;
  cudaMemsetAsync(W, 0, sizeof(double) * bb, s);  // the leaf never writes above the diagonal
  auto orthonormalise = [&](const double* Zin, double* Qout) -> int {
    // B = Zin^T Zin (split-k), guarded, factored by the Cholesky leaf; Qout = Zin W^T with W = L^-1
    GemmDesc g = make_desc(Zin, EIG_B, Zin, EIG_B, B, EIG_B, 1, EIG_B / 32, n_pad);
    int rr;
    if ((rr = launch_skinny(s, true, g, skinny, EIG_B, &h->launches))) return rr;
    gram_guard_kernel<<<1, 256, 0, s>>>(B, EIG_B);
    cudaMemsetAsync(info, 0, sizeof(int), s);
    leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, s>>>(B, EIG_B, L, EIG_B, W, EIG_B, logdet, info, 0);
    h->launches += 2;
    CU(cudaGetLastError());
    GemmDesc p = make_desc(Zin, EIG_B, W, EIG_B, Qout, EIG_B, 2 * mt, 4, EIG_B);
    return launch_gemm(s, false, false, p, 1, &h->launches, SHAPE_T);
  };
  subspace_init_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(Z, n, n_pad);
  h->launches++;
  if ((r = orthonormalise(Z, Q))) {
    cleanup();
    return r;
  }
  stamp("init+orth");
  bool converged = false;
  h->iters = 0;
  for (int it = 0; it < max_iter && !converged; it++) {
    h->iters = it + 1;
    if (G) {  // Y = G Q
      GemmDesc gy = make_desc(G, n_pad, Q, EIG_B, Y, EIG_B, mt, EIG_B / 32, n_pad);
      if ((r = launch_skinny(s, false, gy, skinny, n_pad, &h->launches))) break;
    } else {  // T = Xw^T Q (c_pad x 128), Y = Xw T (split over the cells, fixed-order reduction)
      GemmDesc gt = make_desc(Xw, c_pad, Q, EIG_B, Tc, EIG_B, (int)(c_pad / 128), 1, n_pad);
      if ((r = launch_gemm(s, true, true, gt, 1, &h->launches))) break;
      GemmDesc gy = make_desc(Xw, c_pad, Tc, EIG_B, skinny, EIG_B, mt, 1, (int)c_pad);
      gy.k_split = ks_op, gy.splitC = (long)nb;
      if ((r = launch_gemm(s, false, true, gy, 1, &h->launches, SHAPE_L, nz_op))) break;
      splitk_reduce_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(skinny, (long)nb, nz_op, (long)nb, Y);
      h->launches++;
    }
    // H = Q^T Y
    GemmDesc gh = make_desc(Q, EIG_B, Y, EIG_B, H, EIG_B, 1, EIG_B / 32, n_pad);
    if ((r = launch_skinny(s, true, gh, skinny, EIG_B, &h->launches))) break;
    jacobi_eig128_kernel<<<1, EIG_THREADS, EIG_SMEM_BYTES, s>>>(H, EIG_B, 1e-13, h->lambda, V, Vs, nullptr);
    h->launches++;
    // U = Q V (Ritz vectors), Z = Y V diag(1/lambda) (one power step applied to them, normalised)
    GemmDesc gu = make_desc(Q, EIG_B, V, EIG_B, U, EIG_B, mt, EIG_B / 32, EIG_B);
    if ((r = launch_gemm(s, false, true, gu, 1, &h->launches, SHAPE_N))) break;
    GemmDesc gz = make_desc(Y, EIG_B, Vs, EIG_B, Z, EIG_B, mt, EIG_B / 32, EIG_B);
    if ((r = launch_gemm(s, false, true, gz, 1, &h->launches, SHAPE_N))) break;
    ritz_residual_kernel<<<EIG_B, 256, 0, s>>>(Z, U, n_pad, h->lambda, h->resid);
    h->launches++;
    if (cudaMemcpyAsync(h->h_lambda, h->lambda, sizeof(double) * EIG_B, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(h->h_resid, h->resid, sizeof(double) * EIG_B, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      r = fail(GPRAS_E_CUDA, "subspace iteration", cudaGetLastError());
      break;
    }
    converged = true;
    for (int j = 0; j < kconv; j++)
      if (!(h->h_resid[j] <= tol)) converged = false;
    if (!converged && (r = orthonormalise(Z, Q))) break;
    stamp("iteration");
  }
  cleanup();
  stamp("cleanup");
  if (r) return r;
  return converged ? 0 : 1;
}

}  // namespace

extern "C" {

int gpras_pre_create(gpras_pre** out, int device, int c, int hydraulic, double wet_threshold) {
  if (!out || c <= 0 || hydraulic < 0 || hydraulic > 2) return fail(GPRAS_E_ARG, "bad argument");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  int r;
  if ((r = prepare_device())) return r;
  static std::atomic<bool> attr_done[64] = {};  // benign if two threads both set the (idempotent) attributes
  if (device < 64 && !attr_done[device]) {
    if ((r = opt_in_smem(project_kernel<8>, ProjCfg<8>::SMEM_BYTES)) || (r = opt_in_smem(project_kernel<16>, ProjCfg<16>::SMEM_BYTES)) ||
        (r = opt_in_smem(project_kernel<32>, ProjCfg<32>::SMEM_BYTES)) || (r = opt_in_smem(project_kernel<64>, ProjCfg<64>::SMEM_BYTES)) ||
        (r = opt_in_smem(project_tma_kernel<8>, ProjTmaCfg<8>::SMEM_BYTES)) || (r = opt_in_smem(project_tma_kernel<16>, ProjTmaCfg<16>::SMEM_BYTES)) ||
        (r = opt_in_smem(project_tma_kernel<32>, ProjTmaCfg<32>::SMEM_BYTES)) || (r = opt_in_smem(project_tma_kernel<64>, ProjTmaCfg<64>::SMEM_BYTES)) ||
        (r = opt_in_smem(colstats_tma_kernel, CS_SMEM_BYTES)) || (r = opt_in_smem(jacobi_eig128_kernel, EIG_SMEM_BYTES)))
      return r;
    attr_done[device] = true;
  }
  gpras_pre* h = new gpras_pre();
  h->device = device, h->c = c, h->hp = hydraulic, h->wet_threshold = wet_threshold, h->c_pad = round_up(c, 128);
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t cp = h->c_pad;
  if ((r = dalloc(&h->elev, cp)) || (r = dalloc(&h->weights_in, cp)) || (r = dalloc(&h->mean, cp)) || (r = dalloc(&h->wfull, cp)) ||
      (r = dalloc(&h->E, PRE_MAX_MODES * cp)) || (r = dalloc(&h->x_mean, PRE_MAX_MODES)) || (r = dalloc(&h->x_std, PRE_MAX_MODES)) ||
      (r = dalloc(&h->lambda, EIG_B)) || (r = dalloc(&h->resid, EIG_B)) || (r = dalloc(&h->Ef, PRE_MAX_MODES * cp)) ||
      (r = dalloc(&h->Ef2, PRE_MAX_MODES * cp)) || (r = dalloc(&h->bias, cp)) ||
      cudaMalloc((void**)&h->cls, sizeof(int) * cp) != cudaSuccess) {
    gpras_pre_destroy(h);
    return r ? r : fail(GPRAS_E_NOMEM, "cudaMalloc");
  }
  CU(cudaMemsetAsync(h->elev, 0, sizeof(double) * cp, h->stream));
  CU(cudaMemsetAsync(h->weights_in, 0, sizeof(double) * cp, h->stream));
  CU(cudaMemsetAsync(h->E, 0, sizeof(double) * PRE_MAX_MODES * cp, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (auto& e : h->ev) CU(cudaEventCreate(&e));
  *out = h;
  return 0;
}

int gpras_pre_destroy(gpras_pre* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  double* bufs[] = {h->elev, h->weights_in, h->mean, h->wfull, h->E, h->x_mean, h->x_std, h->lambda, h->resid, h->Ef, h->Ef2, h->bias};
  for (double* b : bufs)
    if (b) cudaFree(b);
  if (h->cls) cudaFree(h->cls);
  pool_release(h);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int gpras_pre_fit(gpras_pre* h, const double* x, long ldx, int n, int on_device, const double* elevations, const double* weights,
                  int modes, double tol, int max_iter) {
  if (!h || !x || !elevations || !weights || n < 2 || ldx < h->c) return fail(GPRAS_E_ARG, "bad argument");
  if (modes > PRE_MAX_MODES) return fail(GPRAS_E_ARG, "more than 64 spatial modes are not supported");
  if (n > 65535 - 127) return fail(GPRAS_E_ARG, "more than 65 408 samples are not supported (one grid row per sample)");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int c = h->c, n_pad = round_up(n, 128);
  const long c_pad = h->c_pad;
  const int clamp = h->hp == HP_DEPTH;
  h->launches = 0, h->fitted = false, h->map_ready = false;
  int r;
  CU(cudaEventRecord(h->ev[0], s));
  CU(cudaMemcpyAsync(h->elev, elevations, sizeof(double) * c, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->weights_in, weights, sizeof(double) * c, cudaMemcpyHostToDevice, s));
  const double* xd = nullptr;
  long ldd = 0;
  double *owned = nullptr, *part = nullptr, *Xw = nullptr, *G = nullptr, *U = nullptr, *Ft = nullptr, *Zs = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(s);
    pfree(h, owned), pfree(h, part), pfree(h, Xw), pfree(h, G), pfree(h, U), pfree(h, Ft), pfree(h, Zs);
  };
#define PRE_TRY(expr)          \
  do {                         \
    if ((r = (expr))) {        \
      cleanup();               \
      return r;                \
    }                          \
  } while (0)
#define PRE_CU(expr)                                   \
  do {                                                 \
    cudaError_t e__ = (expr);                          \
    if (e__ != cudaSuccess) {                          \
      cleanup();                                       \
      return fail(GPRAS_E_CUDA, #expr, e__);           \
    }                                                  \
  } while (0)
  PRE_TRY(stage_matrix(h, x, ldx, n, on_device, &xd, &ldd, &owned));
  // ---- 1. column statistics, wetness classes, input mean ----
  {
    const int gx = (c + 255) / 256;
    int splits = (8 * 148 + gx - 1) / gx;
    if (splits > (n + 31) / 32) splits = (n + 31) / 32;
    if (splits < 1) splits = 1;
    const int rows_per = (n + splits - 1) / splits;
    splits = (n + rows_per - 1) / rows_per;
    PRE_TRY(palloc(h, &part, (size_t)splits * 3 * c_pad));
    CUtensorMap xmap;
    static const bool no_tma = getenv("GPRAS_B200_NO_TMA") != nullptr;
    if (!no_tma && tma_map_2d_f64(&xmap, xd, (uint64_t)c, (uint64_t)n, (uint64_t)ldd, CS_COLS, CS_ROWS))
      colstats_tma_kernel<<<dim3(gx, splits), 128, CS_SMEM_BYTES, s>>>(xmap, n, c, h->elev, clamp, rows_per, part, c_pad);
    else
      colstats_kernel<<<dim3(gx, splits), 128, 0, s>>>(xd, ldd, n, c, h->elev, clamp, rows_per, part, c_pad);
    PRE_CU(cudaGetLastError());
    colstats_finish_kernel<<<(unsigned)((c_pad + 255) / 256), 256, 0, s>>>(part, splits, c_pad, c, n, h->hp, h->elev, h->weights_in,
                                                                          h->wet_threshold, h->mean, h->wfull, h->cls);
    PRE_CU(cudaGetLastError());
    h->launches += 2;
  }
  PRE_CU(cudaEventRecord(h->ev[1], s));
  // ---- 2. centred, weighted samples ----
  PRE_TRY(palloc(h, &Xw, (size_t)n_pad * c_pad));
  center_weight_kernel<<<dim3((unsigned)((c_pad + 255) / 256), n_pad), 256, 0, s>>>(xd, ldd, n, c, h->elev, clamp, h->mean, h->wfull, Xw,
                                                                                  c_pad);
  PRE_CU(cudaGetLastError());
  h->launches++;
  PRE_CU(cudaEventRecord(h->ev[2], s));
  // ---- 3. Gram matrix G = Xw Xw^T on the DMMA engine (lower tiles), mirrored to full storage -- only while forming it
  //         is cheaper than applying Xw and Xw^T once per subspace iteration (see subspace_eig) ----
  const bool use_gram = n_pad <= PRE_GRAM_MAX_N;
  if (use_gram) {
    PRE_TRY(palloc(h, &G, (size_t)n_pad * n_pad));
    GemmDesc g = make_desc(Xw, c_pad, Xw, c_pad, G, n_pad, n_pad / 128, n_pad / 128, (int)c_pad);
    g.tri = 1;
    PRE_TRY(launch_gemm(s, false, false, g, 1, &h->launches));
    mirror_lower_full_kernel<<<dim3((n_pad + 255) / 256, n_pad), 256, 0, s>>>(G, n_pad, n_pad);
    PRE_CU(cudaGetLastError());
    h->launches++;
  }
  PRE_CU(cudaEventRecord(h->ev[3], s));
  // ---- 4. leading eigenpairs ----
  PRE_TRY(palloc(h, &U, (size_t)n_pad * EIG_B));
  const int keep = modes > 0 ? modes : PRE_MAX_MODES;
  int kconv = keep < n - 1 ? keep : n - 1;
  if (kconv > c) kconv = c;
  r = subspace_eig(h, use_gram ? G : nullptr, Xw, c_pad, n, n_pad, kconv, tol, max_iter, U);
  if (r < 0) {
    cleanup();
    return r;
  }
  const bool converged = r == 0;
  PRE_CU(cudaEventRecord(h->ev[4], s));
  // ---- 5. EOFs: Ft = Xw^T U (c_pad x 64), scaled by 1 / singular value, sign-normalised ----
  PRE_TRY(palloc(h, &Ft, (size_t)c_pad * EIG_B));
  {
    GemmDesc g = make_desc(Xw, c_pad, U, EIG_B, Ft, EIG_B, (int)(c_pad / 128), PRE_MAX_MODES / 32, n_pad);
    PRE_TRY(launch_gemm(s, true, true, g, 1, &h->launches, SHAPE_N));
    eof_finish_kernel<<<PRE_MAX_MODES, 256, 0, s>>>(Ft, EIG_B, c, c_pad, h->lambda, h->E);
    PRE_CU(cudaGetLastError());
    h->launches++;
  }
  PRE_CU(cudaEventRecord(h->ev[5], s));
  // ---- 6. raw scores of the training samples -> mean / population std per mode ----
  PRE_TRY(palloc(h, &Zs, (size_t)n * PRE_MAX_MODES));
  PRE_TRY(run_project(h, xd, ldd, n, PRE_MAX_MODES, PRE_MAX_MODES, 0, Zs, PRE_MAX_MODES));
  score_stats_kernel<<<PRE_MAX_MODES, 256, 0, s>>>(Zs, PRE_MAX_MODES, n, h->x_mean, h->x_std);
  PRE_CU(cudaGetLastError());
  h->launches++;
  PRE_CU(cudaEventRecord(h->ev[6], s));
  PRE_CU(cudaStreamSynchronize(s));
#undef PRE_TRY
#undef PRE_CU
  cleanup();
  for (int i = 0; i < 6; i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
    h->stage_ms[i] = ms;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
  h->stage_ms[6] = ms;
  h->n_fit = n;
  h->n_eig = EIG_B < n ? EIG_B : n;
  if (h->n_eig > c) h->n_eig = c;
  h->p = keep < h->n_eig ? keep : h->n_eig;
  h->fitted = true;
  if (!converged) return fail(GPRAS_E_STATE, "the retained eigenpairs did not converge within max_iter subspace iterations");
  return 0;
}

int gpras_pre_set_modes(gpras_pre* h, int modes) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  if (!h->fitted) return fail(GPRAS_E_STATE, "fit() / set_state() has not been called");
  if (modes < 0 || modes > PRE_MAX_MODES || modes > h->n_eig) return fail(GPRAS_E_ARG, "modes out of range");
  h->p = modes;
  h->map_ready = false;
  return 0;
}

int gpras_pre_set_state(gpras_pre* h, const unsigned char* dry, const double* input_mean, const double* weights,
                        const double* eofs, const double* x_mean, const double* x_std, const double* elevations, int p) {
  if (!h || !dry || !input_mean || !weights || !eofs || !x_mean || !x_std || !elevations || p < 0 || p > PRE_MAX_MODES)
    return fail(GPRAS_E_ARG, "bad argument");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int c = h->c;
  std::vector<int> cls(h->c_pad, 1);
  std::vector<double> wf(h->c_pad, 0.0), mu(h->c_pad, 0.0);
  for (int j = 0; j < c; j++) {
    cls[j] = dry[j] ? 1 : 2;
    wf[j] = dry[j] ? 0.0 : weights[j];
    mu[j] = input_mean[j];
  }
  CU(cudaMemcpyAsync(h->cls, cls.data(), sizeof(int) * h->c_pad, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->wfull, wf.data(), sizeof(double) * h->c_pad, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->mean, mu.data(), sizeof(double) * h->c_pad, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->elev, elevations, sizeof(double) * c, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->weights_in, weights, sizeof(double) * c, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(h->E, 0, sizeof(double) * PRE_MAX_MODES * h->c_pad, s));
  if (p > 0) {
    CU(cudaMemcpy2DAsync(h->E, sizeof(double) * h->c_pad, eofs, sizeof(double) * c, sizeof(double) * c, p, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->x_mean, x_mean, sizeof(double) * p, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->x_std, x_std, sizeof(double) * p, cudaMemcpyHostToDevice, s));
  }
  CU(cudaStreamSynchronize(s));
  h->p = p, h->n_eig = p, h->fitted = true, h->map_ready = false;
  memset(h->h_lambda, 0, sizeof h->h_lambda);
  memset(h->h_resid, 0, sizeof h->h_resid);
  return 0;
}

int gpras_pre_get(gpras_pre* h, int which, double* out) {
  if (!h || !out) return fail(GPRAS_E_ARG, "null argument");
  if (!h->fitted) return fail(GPRAS_E_STATE, "fit() / set_state() has not been called");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  const int c = h->c;
  switch (which) {
    case 0: {
      std::vector<int> cls(c);
      CU(cudaMemcpy(cls.data(), h->cls, sizeof(int) * c, cudaMemcpyDeviceToHost));
      for (int j = 0; j < c; j++) out[j] = cls[j];
      return 0;
    }
    case 1: CU(cudaMemcpy(out, h->mean, sizeof(double) * c, cudaMemcpyDeviceToHost)); return 0;
    case 2: CU(cudaMemcpy(out, h->wfull, sizeof(double) * c, cudaMemcpyDeviceToHost)); return 0;
    case 3:
      if (h->p > 0)
        CU(cudaMemcpy2D(out, sizeof(double) * c, h->E, sizeof(double) * h->c_pad, sizeof(double) * c, h->p, cudaMemcpyDeviceToHost));
      return 0;
    case 4:
      for (int j = 0; j < h->n_eig; j++) out[j] = h->n_fit > 1 ? h->h_lambda[j] / (double)(h->n_fit - 1) : 0.0;
      return 0;
    case 5: CU(cudaMemcpy(out, h->x_mean, sizeof(double) * h->p, cudaMemcpyDeviceToHost)); return 0;
    case 6: CU(cudaMemcpy(out, h->x_std, sizeof(double) * h->p, cudaMemcpyDeviceToHost)); return 0;
    case 7:
      for (int j = 0; j < h->n_eig; j++) out[j] = h->h_resid[j];
      return 0;
  }
  return fail(GPRAS_E_ARG, "which out of range");
}

int gpras_pre_trim(gpras_pre* h) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  for (auto it = h->pool.begin(); it != h->pool.end();)
    if (!it->in_use) {
      cudaFree(it->ptr);
      it = h->pool.erase(it);
    } else {
      ++it;
    }
  return 0;
}

int gpras_pre_modes(gpras_pre* h) { return h ? h->p : 0; }
int gpras_pre_eigen_count(gpras_pre* h) { return h ? h->n_eig : 0; }
int gpras_pre_iterations(gpras_pre* h) { return h ? h->iters : 0; }
int gpras_pre_last_launches(gpras_pre* h) { return h ? h->launches : 0; }

int gpras_pre_last_stage_ms(gpras_pre* h, double* ms7) {
  if (!h || !ms7) return fail(GPRAS_E_ARG, "null argument");
  memcpy(ms7, h->stage_ms, sizeof h->stage_ms);
  return 0;
}

int gpras_pre_transform(gpras_pre* h, const double* x, long ldx, int n, int on_device, double* z) {
  if (!h || !x || !z || n <= 0 || ldx < h->c) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->fitted || h->p <= 0) return fail(GPRAS_E_STATE, "fit() / set_state() has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  h->launches = 0;
  const double* xd = nullptr;
  long ldd = 0;
  double *owned = nullptr, *zd = nullptr;
  int r;
  if ((r = stage_matrix(h, x, ldx, n, on_device, &xd, &ldd, &owned))) return r;
  double* zout = z;
  if (!on_device) {
    if ((r = palloc(h, &zd, (size_t)n * h->p))) {
      pfree(h, owned);
      return r;
    }
    zout = zd;
  }
  r = run_project(h, xd, ldd, n, pn_of(h->p), h->p, 1, zout, h->p);
  if (!r && !on_device && cudaMemcpy(z, zd, sizeof(double) * (size_t)n * h->p, cudaMemcpyDeviceToHost) != cudaSuccess)
    r = fail(GPRAS_E_CUDA, "copy of the scores", cudaGetLastError());
  cudaStreamSynchronize(s);
  pfree(h, owned), pfree(h, zd);
  return r;
}

int gpras_pre_reverse(gpras_pre* h, const double* mean, const double* var, int t, double* cell_mean, double* cell_var) {
  if (!h || !mean || !cell_mean || t <= 0 || (var && !cell_var)) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->fitted || h->p <= 0) return fail(GPRAS_E_STATE, "fit() / set_state() has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  h->launches = 0;
  int r;
  if (!h->map_ready && (r = build_map(h))) return r;
  const int p = h->p, pk = round_up(p, 32), c = h->c;
  const long c_pad = h->c_pad;
  double *M = nullptr, *V = nullptr, *om = nullptr, *ov = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(s);
    pfree(h, M), pfree(h, V), pfree(h, om), pfree(h, ov);
  };
  if ((r = palloc(h, &M, (size_t)PRE_REV_TB * pk)) || (r = palloc(h, &om, (size_t)PRE_REV_TB * c_pad)) ||
      (var && ((r = palloc(h, &V, (size_t)PRE_REV_TB * pk)) || (r = palloc(h, &ov, (size_t)PRE_REV_TB * c_pad))))) {
    cleanup();
    return r;
  }
  for (int t0 = 0; t0 < t && !r; t0 += PRE_REV_TB) {
    const int tb = t - t0 < PRE_REV_TB ? t - t0 : PRE_REV_TB;
    const int tb_pad = round_up(tb, 128);
    cudaMemsetAsync(M, 0, sizeof(double) * (size_t)tb_pad * pk, s);
    cudaMemcpy2DAsync(M, sizeof(double) * pk, mean + (size_t)t0 * p, sizeof(double) * p, sizeof(double) * p, tb, cudaMemcpyHostToDevice, s);
    GemmDesc gm = make_desc(M, pk, h->Ef, c_pad, om, c_pad, tb_pad / 128, (int)(c_pad / 128), pk);
    gm.epilogue = EPI_BIAS, gm.bias = h->bias;
    if ((r = launch_gemm(s, false, true, gm, 1, &h->launches))) break;
    cudaMemcpy2DAsync(cell_mean + (size_t)t0 * c, sizeof(double) * c, om, sizeof(double) * c_pad, sizeof(double) * c, tb,
                      cudaMemcpyDeviceToHost, s);
    if (var) {
      cudaMemsetAsync(V, 0, sizeof(double) * (size_t)tb_pad * pk, s);
      cudaMemcpy2DAsync(V, sizeof(double) * pk, var + (size_t)t0 * p, sizeof(double) * p, sizeof(double) * p, tb, cudaMemcpyHostToDevice, s);
      GemmDesc gv = make_desc(V, pk, h->Ef2, c_pad, ov, c_pad, tb_pad / 128, (int)(c_pad / 128), pk);
      if ((r = launch_gemm(s, false, true, gv, 1, &h->launches))) break;
      cudaMemcpy2DAsync(cell_var + (size_t)t0 * c, sizeof(double) * c, ov, sizeof(double) * c_pad, sizeof(double) * c, tb,
                        cudaMemcpyDeviceToHost, s);
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) r = fail(GPRAS_E_CUDA, "reverse transform", cudaGetLastError());
  }
  cleanup();
  return r;
}

long gpras_pre_cell_pitch(gpras_pre* h) { return h ? h->c_pad : 0; }

// reverse_transform with the (T x C) results left ON THE DEVICE (or streamed through a ring buffer when the caller keeps
// nothing): one fused kernel per block of events, general variance map (one hyperparameter set per mode).
int gpras_pre_reverse_device(gpras_pre* h, const double* mean, const double* var, int t, int on_device, double* cell_mean,
                             double* cell_var, long ldc) {
  if (!h || !mean || !var || t <= 0) return fail(GPRAS_E_ARG, "bad argument");
  if ((cell_mean != nullptr) != (cell_var != nullptr)) return fail(GPRAS_E_ARG, "pass both cell_mean and cell_var, or neither");
  if (!h->fitted || h->p <= 0) return fail(GPRAS_E_STATE, "fit() / set_state() has not been called");
  if (cell_mean && ldc < h->c_pad) return fail(GPRAS_E_ARG, "ldc smaller than gpras_pre_cell_pitch()");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  h->launches = 0;
  int r;
  if (!h->map_ready && (r = build_map(h))) return r;
  static std::atomic<bool> attr_done[64] = {};
  if (h->device < 64 && !attr_done[h->device]) {
    if ((r = opt_in_smem(cells_general_kernel<32>, CellsGenCfg<32>::SMEM_BYTES)) ||
        (r = opt_in_smem(cells_general_kernel<64>, CellsGenCfg<64>::SMEM_BYTES)))
      return r;
    attr_done[h->device] = true;
  }
  const int p = h->p, pk = round_up(p, 32);
  const long c_pad = h->c_pad;
  constexpr int RING_ROWS = 256;
  const bool keep = cell_mean != nullptr;
  double *M = nullptr, *V = nullptr, *ring_m = nullptr, *ring_v = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(s);
    pfree(h, M), pfree(h, V), pfree(h, ring_m), pfree(h, ring_v);
  };
  if ((r = palloc(h, &M, (size_t)PRE_REV_TB * pk)) || (r = palloc(h, &V, (size_t)PRE_REV_TB * pk)) ||
      (!keep && ((r = palloc(h, &ring_m, (size_t)RING_ROWS * c_pad)) || (r = palloc(h, &ring_v, (size_t)RING_ROWS * c_pad))))) {
    cleanup();
    return r;
  }
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int t0 = 0; t0 < t && !r; t0 += PRE_REV_TB) {
    const int tb = t - t0 < PRE_REV_TB ? t - t0 : PRE_REV_TB;
    const int tb_pad = round_up(tb, 128);
    cudaMemsetAsync(M, 0, sizeof(double) * (size_t)tb_pad * pk, s);
    cudaMemsetAsync(V, 0, sizeof(double) * (size_t)tb_pad * pk, s);
    cudaMemcpy2DAsync(M, sizeof(double) * pk, mean + (size_t)t0 * p, sizeof(double) * p, sizeof(double) * p, tb, kind, s);
    cudaMemcpy2DAsync(V, sizeof(double) * pk, var + (size_t)t0 * p, sizeof(double) * p, sizeof(double) * p, tb, kind, s);
    double* om = keep ? cell_mean + (size_t)t0 * ldc : ring_m;
    double* ov = keep ? cell_var + (size_t)t0 * ldc : ring_v;
    const long ldo = keep ? ldc : c_pad;
    const int t_tiles = (keep ? round_up(tb, CELLS_ROWS) : tb_pad) / CELLS_ROWS;  // kept output: only whole 64-row tiles that exist
    const int per_cta = (t_tiles + 1) / 2;
    dim3 grid((unsigned)(c_pad / 128), (unsigned)((t_tiles + per_cta - 1) / per_cta));
    if (pk == 32)
      cells_general_kernel<32><<<grid, CELLS_THREADS, CellsGenCfg<32>::SMEM_BYTES, s>>>(M, V, pk, h->Ef, c_pad, h->bias, om, ov, ldo, t_tiles,
                                                                                      per_cta, keep ? (1 << 30) : RING_ROWS);
    else
      cells_general_kernel<64><<<grid, CELLS_THREADS, CellsGenCfg<64>::SMEM_BYTES, s>>>(M, V, pk, h->Ef, c_pad, h->bias, om, ov, ldo, t_tiles,
                                                                                      per_cta, keep ? (1 << 30) : RING_ROWS);
    h->launches++;
    if (cudaGetLastError() != cudaSuccess) r = fail(GPRAS_E_CUDA, "cells_general_kernel");
    if (!on_device && cudaStreamSynchronize(s) != cudaSuccess) r = fail(GPRAS_E_CUDA, "reverse transform", cudaGetLastError());
  }
  cleanup();
  return r;
}

// reverse_transform fused with the metrics: per-mode variances, truth streamed, nothing written per cell-depth
// (gpr.predict of per-column models -> reverse_transform -> wse_2_depth -> export_metric_summary, pipeline.py:260-286).
int gpras_pre_reverse_metrics(gpras_pre* h, gpras_metrics* m, const double* mean, const double* var, int t, int on_device,
                              const double* truth, long ldx, int truth_on_device) {
  if (!h || !m || !mean || !var || t <= 0) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->fitted || h->p <= 0) return fail(GPRAS_E_STATE, "fit() / set_state() has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  int r;
  if (!h->map_ready && (r = build_map(h))) return r;
  const int p = h->p, pk = round_up(p, 32);
  const long c_pad = h->c_pad;
  double *M = nullptr, *V = nullptr, *X = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(s);
    pfree(h, M), pfree(h, V), pfree(h, X);
  };
  if ((r = palloc(h, &M, (size_t)PRE_REV_TB * pk)) || (r = palloc(h, &V, (size_t)PRE_REV_TB * pk)) ||
      (truth && !truth_on_device && (r = palloc(h, &X, (size_t)PRE_REV_TB * c_pad)))) {
    cleanup();
    return r;
  }
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int t0 = 0; t0 < t && !r; t0 += PRE_REV_TB) {
    const int tb = t - t0 < PRE_REV_TB ? t - t0 : PRE_REV_TB;
    const int tb_pad = round_up(tb, 128);
    cudaMemsetAsync(M, 0, sizeof(double) * (size_t)tb_pad * pk, s);
    cudaMemsetAsync(V, 0, sizeof(double) * (size_t)tb_pad * pk, s);
    cudaMemcpy2DAsync(M, sizeof(double) * pk, mean + (size_t)t0 * p, sizeof(double) * p, sizeof(double) * p, tb, kind, s);
    cudaMemcpy2DAsync(V, sizeof(double) * pk, var + (size_t)t0 * p, sizeof(double) * p, sizeof(double) * p, tb, kind, s);
    const double* xd = nullptr;
    long ldd = 0;
    if (truth) {
      if (truth_on_device) {
        xd = truth + (size_t)t0 * ldx, ldd = ldx;
      } else {
        cudaMemcpy2DAsync(X, sizeof(double) * c_pad, truth + (size_t)t0 * ldx, sizeof(double) * ldx, sizeof(double) * h->c, tb,
                          cudaMemcpyHostToDevice, s);
        xd = X, ldd = c_pad;
      }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      r = fail(GPRAS_E_CUDA, "staging the mode-space block", cudaGetLastError());
      break;
    }
    r = gpras_metrics_update_modes(m, M, V, pk, pk, h->Ef, c_pad, h->bias, xd, ldd, tb);
  }
  cleanup();
  return r;
}

int gpras_dsyev128(void* cuda_stream, const double* H, double* lambda, double* V) {
  if (!H || !lambda || !V) return fail(GPRAS_E_ARG, "null argument");
  if (gpras_device_count() <= 0) return fail(GPRAS_E_CUDA, "no CUDA device (no CPU fallback)");
  int r;
  if ((r = opt_in_smem(jacobi_eig128_kernel, EIG_SMEM_BYTES))) return r;
  double* Vs = nullptr;
  if ((r = dalloc(&Vs, (size_t)EIG_B * EIG_B))) return r;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  jacobi_eig128_kernel<<<1, EIG_THREADS, EIG_SMEM_BYTES, s>>>(H, EIG_B, 0.0, lambda, V, Vs, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(Vs);
  if (e != cudaSuccess) return fail(GPRAS_E_CUDA, "jacobi_eig128_kernel", e);
  return 0;
}

}  // extern "C"
