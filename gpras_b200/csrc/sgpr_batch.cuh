// The per-column sparse models of one GPRAS.fit call, trained TOGETHER and ON THE DEVICE.
//
// gpras/gpr.py:273-274 loops over one gpflow SGPR per target column; its default recipe ("two-stage", gpr.py:112-127) is
// Adam on the inducing inputs followed by Adam on the hyperparameters (gpr.py:147-173).  At reference scale (N = 5 000,
// M = 50) one evaluation is ~30 launches of microsecond kernels, so a fit is bound by launches and by the host round trip
// of every Adam step.  Here
//   * the P models live in one arena at a fixed stride, and every kernel of the evaluation (sgpr_abi.cuh's sequence, the same
//     kernels) runs ONCE for all models with the model index in blockIdx.y -- bitwise the single-model results;
//   * the Adam step itself (GPflow's softplus / log transforms, the LogNormal(0, 1) priors, the chain rule to the
//     unconstrained variables, Keras' update rule and the reference's early-stopping rule, per model) is a kernel, so
//     [unpack u -> theta, Z] -> [evaluation] -> [Adam step] is one CUDA graph replayed max_iter times with no host
//     round trip; the host reads the variables and the loss history back once per stage.
// Restricted to M <= 128 inducing points (one 128-tile: the factorisations are single leaf launches).
#pragma once
#include "sgpr_abi.cuh"
#include "sgpr_fused.cuh"

struct SgprAdamCfg {
  int n_ls, train_hypers, train_z, transform, priors;
  int rule;  // 0: Keras Adam + the reference's early-stopping rule (gpr.py:147-173); 1: Keras Adadelta, fixed number of steps (gpr.py:176-192)
  double lr, jitter, noise_floor;
  bool operator==(const SgprAdamCfg& o) const {
    return n_ls == o.n_ls && train_hypers == o.train_hypers && train_z == o.train_z && transform == o.transform &&
           priors == o.priors && rule == o.rule && lr == o.lr && jitter == o.jitter && noise_floor == o.noise_floor;
  }
};

struct gpras_sgpr_batch {
  int device = 0, kid = 0, n = 0, d = 0, m = 0, p = 0, n_pad = 0, m_pad = 128, r_pad = 32, ntn = 0;
  int nz_aat = 1, ks_aat = 0, nu = 0;
  long bs = 0;  // doubles between the buffers of consecutive models
  cudaStream_t stream = nullptr;
  bool has_data = false, warmed = false;
  int launches = 0;
  void* arena = nullptr;
  int* info = nullptr;
  // model 0's buffers (model b: + b * bs); names as in gpras_sgpr
  double *X = nullptr, *Xs = nullptr, *Z = nullptr, *Zs = nullptr, *Y = nullptr;
  double *Kuf = nullptr, *Kuu = nullptr, *WL = nullptr, *Ap = nullptr, *slabs = nullptr, *AATs = nullptr, *B = nullptr,
         *WB = nullptr, *Binv = nullptr, *Rm = nullptr, *RA = nullptr, *RW = nullptr, *T1 = nullptr, *Guu = nullptr,
         *Guf1 = nullptr;
  double *ae = nullptr, *c = nullptr, *chat = nullptr, *u = nullptr, *skinny = nullptr;
  double *theta = nullptr, *logdetL = nullptr, *logdetB = nullptr, *scal = nullptr, *partA = nullptr, *partB = nullptr,
         *zpA = nullptr, *zpB = nullptr, *result = nullptr;
  // trainer state per model: unconstrained variables [variance, noise, lengthscale(s), Z], Adam moments, bookkeeping
  double *au = nullptr, *amom = nullptr, *avel = nullptr, *ast = nullptr;
  double* losses = nullptr;  // max_iter x p, device
  int losses_cap = 0;
  double* h_pinned = nullptr;  // p x (nu + 8) staging
  std::vector<std::pair<SgprAdamCfg, cudaGraphExec_t>> graphs;
  // fused evaluation (sgpr_fused.cuh) for m <= 64, d <= 32: four kernels per evaluation, inputs shared by the models
  bool fused = false, conditioned = false;
  gpras::SfArgs fa = {};
  double *Xsh = nullptr, *yv = nullptr, *yy = nullptr;
  double *xt = nullptr, *pmean = nullptr, *pvar = nullptr;  // prediction staging: test inputs, mean / variance (t x p)
  int pred_cap = 0;
  // trainer: the models run as staggered lanes (consecutive groups of models, one stream each), so that one lane's
  // one-CTA-per-model kernels overlap another lane's tile passes; a graph holds a chunk of consecutive steps of all lanes
  static constexpr int MAX_LANES = 4;
  cudaStream_t lane_stream[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};  // [0] is `stream`
  cudaEvent_t ev_stagger[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr}, ev_join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  struct ChunkGraph {
    SgprAdamCfg cfg;
    int steps, lanes;
    cudaGraphExec_t exec;
  };
  std::vector<ChunkGraph> chunk_graphs;
};

namespace gpras {

constexpr int AST = 8;  // per-model trainer state: best, count, active, t, failed info, last loss

__device__ __forceinline__ double sb_forward(double u, int transform) {
  if (transform == 1) return exp(u);
  return (u > 0.0 ? u : 0.0) + log1p(exp(-fabs(u)));  // softplus, as numpy.logaddexp(0, u)
}
__device__ __forceinline__ double sb_dforward(double u, int transform) {
  if (transform == 1) return exp(u);
  return 0.5 * (1.0 + tanh(0.5 * u));
}

// Keras Adam on one variable, in the operation order of the NumPy restatement (gpras_b200/gpr.py:_optimize_adam) and without
// fused multiply-adds, so that the device loop and the host loop round alike.
__device__ __forceinline__ double sb_adam_update(double u, double g, double& mom, double& vel, double alpha) {
  const double b1 = 0.9, b2 = 0.999, eps = 1e-7;
  mom = __dadd_rn(__dmul_rn(b1, mom), __dmul_rn(1.0 - b1, g));
  vel = __dadd_rn(__dmul_rn(b2, vel), __dmul_rn(__dmul_rn(1.0 - b2, g), g));
  return __dsub_rn(u, __ddiv_rn(__dmul_rn(alpha, mom), __dadd_rn(sqrt(vel), eps)));
}

// Keras Adadelta (rho 0.95, eps 1e-7) on one variable, operation order of gpras_b200/gpr.py:_optimize_adadelta; acc_g / acc_d
// live in the two moment buffers.
__device__ __forceinline__ double sb_adadelta_update(double u, double g, double& acc_g, double& acc_d, double lr) {
  const double rho = 0.95, eps = 1e-7;
  acc_g = __dadd_rn(__dmul_rn(rho, acc_g), __dmul_rn(__dmul_rn(1.0 - rho, g), g));
  const double upd = __ddiv_rn(__dmul_rn(g, sqrt(__dadd_rn(acc_d, eps))), sqrt(__dadd_rn(acc_g, eps)));
  acc_d = __dadd_rn(__dmul_rn(rho, acc_d), __dmul_rn(__dmul_rn(1.0 - rho, upd), upd));
  return __dsub_rn(u, __dmul_rn(lr, upd));
}

__device__ __forceinline__ double sb_update(int rule, double u, double g, double& m1, double& m2, double alpha, double lr) {
  return rule == 0 ? sb_adam_update(u, g, m1, m2, alpha) : sb_adadelta_update(u, g, m1, m2, lr);
}

// u -> theta (constrained) and Z, info = 0 (pointers of one model).
__device__ __forceinline__ void sgpr_unpack_body(const double* __restrict__ au, double* __restrict__ theta, double* __restrict__ Z,
                                                 int* __restrict__ info, int D, int m, int n_ls, int transform, double noise_floor) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    *info = 0;
    theta[0] = sb_forward(au[0], transform);
    theta[1] = sb_forward(au[1], transform) + noise_floor;
  }
  for (int dd = tid; dd < D; dd += blockDim.x) theta[2 + dd] = sb_forward(au[2 + (n_ls == 1 ? 0 : dd)], transform);
  for (int e = tid; e < m * D; e += blockDim.x) Z[e] = au[2 + n_ls + e];
}

// One CTA per model.
static __global__ void sgpr_batch_unpack_kernel(const double* __restrict__ au, double* __restrict__ theta, double* __restrict__ Z,
                                         int* __restrict__ info, int D, int m, int n_ls, int transform, double noise_floor,
                                         long bs) {
  const long o = blockIdx.x * bs;
  sgpr_unpack_body(au + o, theta + o, Z + o, info + o * 2, D, m, n_ls, transform, noise_floor);
}

// One Adam step per still-active model (gpr.py:147-173; Keras Adam: lr, beta 0.9 / 0.999, eps 1e-7):
//   loss = -(ELBO + log prior of the trainable hyperparameters), gradient w.r.t. the trainable unconstrained variables by the
//   chain rule (result holds dELBO/dlog theta and dELBO/dZ), update, then the reference's early-stopping bookkeeping.
// Pointers of model b; called by every thread of the model's CTA.
__device__ __forceinline__ void sgpr_adam_body(const double* __restrict__ result, const int* __restrict__ info,
                                               double* __restrict__ au, double* __restrict__ amom, double* __restrict__ avel,
                                               double* __restrict__ ast, double* __restrict__ losses, int p, int b, int D, int m,
                                               const SgprAdamCfg& cfg) {
  const int tid = threadIdx.x;
  if (ast[2] == 0.0) return;  // stopped earlier
  __shared__ double s_alpha;
  __shared__ int s_fail;
  const int n_ls = cfg.n_ls, nh = 2 + n_ls;
  const double b1 = 0.9, b2 = 0.999, tol = 10e-6;
  const int patience = 50;
  const int t = (int)ast[3] + 1;
  if (tid == 0) {
    s_fail = *info;
    s_alpha = cfg.lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
  }
  __syncthreads();
  if (s_fail != 0) {  // Kuu or B lost positive definiteness: the model stops here, the host raises
    if (tid == 0) ast[2] = 0.0, ast[4] = (double)s_fail;
    return;
  }
  const double alpha = s_alpha;
  // hyperparameters: a handful of scalars, thread 0 (fixed order)
  if (tid == 0) {
    double lp = 0.0;
    if (cfg.train_hypers) {
      const double LOG_2PI = 1.8378770664093453;
      double gls = 0.0;
      if (n_ls == 1)
        for (int dd = 0; dd < D; dd++) gls += result[3 + dd];
      for (int k = 0; k < nh; k++) {
        const double uk = au[k];
        const double v = sb_forward(uk, cfg.transform) + (k == 1 ? cfg.noise_floor : 0.0);
        const double gl = k < 2 ? result[1 + k] : (n_ls == 1 ? gls : result[3 + (k - 2)]);
        double dlp = 0.0;
        if (cfg.priors) {
          const double lv = log(v);
          lp += -lv - 0.5 * LOG_2PI - 0.5 * lv * lv;
          dlp = -(1.0 + lv) / v;
        }
        const double g = -((gl / v + dlp) * sb_dforward(uk, cfg.transform));
        au[k] = sb_update(cfg.rule, uk, g, amom[k], avel[k], alpha, cfg.lr);
      }
    }
    const double loss = -(result[0] + lp);
    losses[(long)(t - 1) * p + b] = loss;
    ast[5] = loss;
    ast[3] = (double)t;
    if (cfg.rule == 0) {  // the reference's early-stopping rule belongs to its Adam loop only
      if ((ast[0] - loss) / fabs(loss) > tol) {
        ast[0] = loss, ast[1] = 0.0;
      } else {
        ast[1] += 1.0;
        if (ast[1] > (double)patience) ast[2] = 0.0;
      }
    }
  }
  if (cfg.train_z) {
    for (int e = tid; e < m * D; e += blockDim.x) {
      const int k = nh + e;
      const double g = -result[3 + D + e];
      au[k] = sb_update(cfg.rule, au[k], g, amom[k], avel[k], alpha, cfg.lr);
    }
  }
}

// One CTA per model.
static __global__ void sgpr_batch_adam_kernel(const double* __restrict__ result, const int* __restrict__ info,
                                       double* __restrict__ au, double* __restrict__ amom, double* __restrict__ avel,
                                       double* __restrict__ ast, double* __restrict__ losses, int p, int D, int m,
                                       SgprAdamCfg cfg, long bs) {
  const long o = blockIdx.x * bs;
  sgpr_adam_body(result + o, info + o * 2, au + o, amom + o, avel + o, ast + o, losses, p, blockIdx.x, D, m, cfg);
}

// Fused path, trainer: [u -> theta, Z] + sf_prep in one launch, and [bound / gradient from the pieces] + [Adam step] in another.
template <int KID>
__global__ void __launch_bounds__(SF_THREADS, 1) sf_prep_train_kernel(const SfArgs a, const double* __restrict__ au, int n_ls,
                                                                      int transform, double noise_floor) {
  const long o = (long)blockIdx.y * a.bs;
  sgpr_unpack_body(au + o, const_cast<double*>(a.theta) + o, const_cast<double*>(a.Z) + o, a.info + o * 2, a.D, a.m, n_ls, transform,
                   noise_floor);
  __syncthreads();  // theta and Z are read back from global memory by this CTA only
  sf_prep_body<KID>(a);
}

static __global__ void __launch_bounds__(SF_THREADS) sf_finish_train_kernel(const SfArgs a, double* __restrict__ result,
                                                                     double* __restrict__ au, double* __restrict__ amom,
                                                                     double* __restrict__ avel, double* __restrict__ ast,
                                                                     double* __restrict__ losses, int p, SgprAdamCfg cfg) {
  const long o = (long)blockIdx.y * a.bs;
  sf_finalize_body(a, result + o);
  __syncthreads();
  sgpr_adam_body(result + o, a.info + o * 2, au + o, amom + o, avel + o, ast + o, losses, p, blockIdx.y, a.D, a.m, cfg);
}

// start of a stage: zero moments, best = +inf, count = 0, active = 1, t = 0, failed = 0
static __global__ void sgpr_batch_reset_kernel(double* __restrict__ amom, double* __restrict__ avel, double* __restrict__ ast, int nu,
                                        long bs) {
  amom += blockIdx.x * bs, avel += blockIdx.x * bs, ast += blockIdx.x * bs;
  for (int e = threadIdx.x; e < nu; e += blockDim.x) amom[e] = avel[e] = 0.0;
  if (threadIdx.x == 0) {
    ast[0] = __longlong_as_double(0x7ff0000000000000LL);
    ast[1] = 0.0, ast[2] = 1.0, ast[3] = 0.0, ast[4] = 0.0, ast[5] = 0.0;
  }
}

static __global__ void fill_nan_kernel(double* __restrict__ v, long count) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < count) v[e] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace gpras

namespace {

// batched forms of the host launchers (blockIdx.y = model)
int sb_gemm(gpras_sgpr_batch* h, bool akm, bool bkm, GemmDesc g, int shape = SHAPE_L, int nz = 1) {
  g.batchA = g.batchB = g.batchC = h->bs;
  return launch_gemm(h->stream, akm, bkm, g, h->p, &h->launches, shape, nz);
}

int sb_skinny(gpras_sgpr_batch* h, bool akm, GemmDesc d, int rows) {
  int ks = d.K / 16;
  ks = (ks + 127) / 128 * 128;
  if (ks < 512) ks = 512;
  const int nz = (d.K + ks - 1) / ks;
  double* out = d.C;
  const long slab = (long)rows * d.ldc;
  d.k_split = ks, d.splitC = slab, d.C = h->skinny;
  int r = sb_gemm(h, akm, true, d, SHAPE_N, nz);
  if (r) return r;
  splitk_reduce_kernel<<<dim3((unsigned)((slab + 255) / 256), h->p), 256, 0, h->stream>>>(h->skinny, slab, nz, slab, out, h->bs);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int KID>
int sb_cov_t(gpras_sgpr_batch* h, const double* Xs1, int n1, int n1_pad, const double* Xs2, int n2, int n2_pad, double* out,
             long ldo, int square, double jitter) {
  const int t1 = n1_pad / CT, t2 = n2_pad / CT;
  const int smem = 2 * h->d * CT_LD * (int)sizeof(double);
  const long tiles = square ? (long)t1 * (t1 + 1) / 2 : (long)t1 * t2;
  cov_kernel<KID><<<dim3((unsigned)tiles, h->p), PT_THREADS, smem, h->stream>>>(Xs1, n1, Xs2, n2, h->d, h->theta, out, ldo, t2, square,
                                                                               square, jitter, h->bs);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int sb_cov(gpras_sgpr_batch* h, const double* Xs1, int n1, int n1_pad, const double* Xs2, int n2, int n2_pad, double* out, long ldo,
           int square, double jitter = -1.0) {
  switch (h->kid) {
    case K_RBF: return sb_cov_t<K_RBF>(h, Xs1, n1, n1_pad, Xs2, n2, n2_pad, out, ldo, square, jitter);
    case K_MATERN12: return sb_cov_t<K_MATERN12>(h, Xs1, n1, n1_pad, Xs2, n2, n2_pad, out, ldo, square, jitter);
    case K_MATERN32: return sb_cov_t<K_MATERN32>(h, Xs1, n1, n1_pad, Xs2, n2, n2_pad, out, ldo, square, jitter);
    case K_MATERN52: return sb_cov_t<K_MATERN52>(h, Xs1, n1, n1_pad, Xs2, n2, n2_pad, out, ldo, square, jitter);
    case K_EXPONENTIAL: return sb_cov_t<K_EXPONENTIAL>(h, Xs1, n1, n1_pad, Xs2, n2, n2_pad, out, ldo, square, jitter);
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

template <int KID>
int sb_chain_t(gpras_sgpr_batch* h, bool uu, const double* Xs, int n, const double* G1, long ldg, int tiles_x, double* part,
               double* zpart) {
  const int D = h->d, R = 1;
  const int smem = (2 * D * CT_LD + CT * D + 2 * R * CT_LD) * (int)sizeof(double);
  static std::atomic<bool> attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    int r;
    if ((r = opt_in_smem(sgpr_chain_kernel<KID, 16>, 220 * 1024)) || (r = opt_in_smem(sgpr_chain_kernel<KID, 64>, 220 * 1024)))
      return r;
    attr_done[dev] = true;
  }
  const dim3 grid(tiles_x, h->p);
  if (D <= 16)
    sgpr_chain_kernel<KID, 16><<<grid, PT_THREADS, smem, h->stream>>>(uu, h->Zs, h->m, Xs, n, D, G1, ldg, h->u, h->r_pad, h->Y, h->r_pad,
                                                                    R, h->theta, tiles_x, part, 1 + D, zpart, h->m_pad, h->bs);
  else
    sgpr_chain_kernel<KID, 64><<<grid, PT_THREADS, smem, h->stream>>>(uu, h->Zs, h->m, Xs, n, D, G1, ldg, h->u, h->r_pad, h->Y, h->r_pad,
                                                                    R, h->theta, tiles_x, part, 1 + D, zpart, h->m_pad, h->bs);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int sb_chain(gpras_sgpr_batch* h, bool uu, const double* Xs, int n, const double* G1, long ldg, int tiles_x, double* part,
             double* zpart) {
  switch (h->kid) {
    case K_RBF: return sb_chain_t<K_RBF>(h, uu, Xs, n, G1, ldg, tiles_x, part, zpart);
    case K_MATERN12: return sb_chain_t<K_MATERN12>(h, uu, Xs, n, G1, ldg, tiles_x, part, zpart);
    case K_MATERN32: return sb_chain_t<K_MATERN32>(h, uu, Xs, n, G1, ldg, tiles_x, part, zpart);
    case K_MATERN52: return sb_chain_t<K_MATERN52>(h, uu, Xs, n, G1, ldg, tiles_x, part, zpart);
    case K_EXPONENTIAL: return sb_chain_t<K_EXPONENTIAL>(h, uu, Xs, n, G1, ldg, tiles_x, part, zpart);
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

// One evaluation (bound + gradient) of all models from theta / Z in device memory -> result (per model).
// Same sequence and kernels as sgpr_forward + sgpr_record_eval (sgpr_abi.cuh) with m_pad = 128.
int sb_record_eval(gpras_sgpr_batch* h, double jitter) {
  cudaStream_t s = h->stream;
  const int n = h->n, D = h->d, m = h->m, n_pad = h->n_pad, m_pad = h->m_pad, ntn = h->ntn, rp = h->r_pad, P = h->p;
  const long bs = h->bs, mm = (long)m_pad * m_pad;
  int r;
  scale_features_kernel<<<dim3((unsigned)(((long)n_pad * D + 255) / 256), P), 256, 0, s>>>(h->X, h->Xs, n, n_pad, D, h->theta, bs);
  scale_features_kernel<<<dim3((unsigned)(((long)m_pad * D + 255) / 256), P), 256, 0, s>>>(h->Z, h->Zs, m, m_pad, D, h->theta, bs);
  h->launches += 2;
  CU(cudaGetLastError());
  if ((r = sb_cov(h, h->Zs, m, m_pad, h->Xs, n, n_pad, h->Kuf, n_pad, 0))) return r;
  if ((r = sb_cov(h, h->Zs, m, m_pad, h->Zs, m, m_pad, h->Kuu, m_pad, 1, jitter))) return r;
  // L = chol(Kuu) in place, WL = L^-1
  leaf_potrf_inv_kernel<<<dim3(1, P), LEAF_THREADS, LEAF_SMEM_BYTES, s>>>(h->Kuu, m_pad, h->Kuu, m_pad, h->WL, m_pad, h->logdetL, h->info,
                                                                       0, bs);
  h->launches++;
  CU(cudaGetLastError());
  {  // A' = WL Kuf
    GemmDesc g = make_desc(h->WL, m_pad, h->Kuf, n_pad, h->Ap, n_pad, 1, ntn, m_pad);
    g.ke_mode = KE_TI;
    if ((r = sb_gemm(h, false, true, g))) return r;
  }
  {  // AATs = A' A'^T / s2, B = I + AATs
    GemmDesc g = make_desc(h->Ap, n_pad, h->Ap, n_pad, h->slabs, m_pad, 1, 1, n_pad);
    g.k_split = h->ks_aat, g.splitC = mm;
    if ((r = sb_gemm(h, false, false, g, SHAPE_L, h->nz_aat))) return r;
    sgpr_finish_b_kernel<<<dim3((unsigned)((mm + 255) / 256), P), 256, 0, s>>>(h->slabs, mm, h->nz_aat, h->theta, m_pad, h->AATs, h->B,
                                                                            bs);
    h->launches++;
    CU(cudaGetLastError());
  }
  leaf_potrf_inv_kernel<<<dim3(1, P), LEAF_THREADS, LEAF_SMEM_BYTES, s>>>(h->B, m_pad, h->B, m_pad, h->WB, m_pad, h->logdetB, h->info, 0,
                                                                       bs);
  h->launches++;
  CU(cudaGetLastError());
  {  // ae = A' Y / s2 ; c = WB ae ; chat = WB^T c ; u = WL^T chat
    GemmDesc g = make_desc(h->Ap, n_pad, h->Y, rp, h->ae, rp, 1, rp / 32, n_pad);
    if ((r = sb_skinny(h, false, g, m_pad))) return r;
    sgpr_scale_noise_kernel<<<dim3((unsigned)(((long)m_pad * rp + 255) / 256), P), 256, 0, s>>>(h->ae, (long)m_pad * rp, h->theta, bs);
    h->launches++;
    GemmDesc g2 = make_desc(h->WB, m_pad, h->ae, rp, h->c, rp, 1, rp / 32, m_pad);
    g2.ke_mode = KE_TI;
    if ((r = sb_skinny(h, false, g2, m_pad))) return r;
    GemmDesc g3 = make_desc(h->WB, m_pad, h->c, rp, h->chat, rp, 1, rp / 32, m_pad);
    g3.kb_mode = KB_TI;
    if ((r = sb_skinny(h, true, g3, m_pad))) return r;
    GemmDesc g4 = make_desc(h->WL, m_pad, h->chat, rp, h->u, rp, 1, rp / 32, m_pad);
    g4.kb_mode = KB_TI;
    if ((r = sb_skinny(h, true, g4, m_pad))) return r;
  }
  {  // Binv = WB^T WB
    GemmDesc g = make_desc(h->WB, m_pad, h->WB, m_pad, h->Binv, m_pad, 1, 1, m_pad);
    g.tri = 1, g.kb_mode = KB_TI;
    if ((r = sb_gemm(h, true, true, g))) return r;
  }
  mirror_lower_kernel<<<dim3((unsigned)((mm + 255) / 256), P), 256, 0, s>>>(h->Binv, m_pad, m_pad, bs);
  sgpr_scalars_kernel<<<dim3(1, P), 256, 0, s>>>(h->AATs, h->Binv, h->c, h->chat, rp, 1, h->Y, rp, n, m, m_pad, h->scal, bs);
  sgpr_build_r_kernel<<<dim3((unsigned)((mm + 255) / 256), P), 256, 0, s>>>(h->Binv, h->AATs, h->chat, rp, 1, m_pad, h->Rm, h->RA, bs);
  h->launches += 3;
  CU(cudaGetLastError());
  // RW = WL^T Rm ;  Guf1 = RW A' ;  Guu = 1/2 WL^T (Rm - R AATs) WL
  GemmDesc g1 = make_desc(h->WL, m_pad, h->Rm, m_pad, h->RW, m_pad, 1, 1, m_pad);
  g1.kb_mode = KB_TI;
  if ((r = sb_gemm(h, true, true, g1))) return r;
  GemmDesc g2 = make_desc(h->RW, m_pad, h->Ap, n_pad, h->Guf1, n_pad, 1, ntn, m_pad);
  if ((r = sb_gemm(h, false, true, g2))) return r;
  GemmDesc g3 = make_desc(h->WL, m_pad, h->RA, m_pad, h->T1, m_pad, 1, 1, m_pad);
  g3.kb_mode = KB_TI;
  if ((r = sb_gemm(h, true, true, g3))) return r;
  GemmDesc g4 = make_desc(h->T1, m_pad, h->WL, m_pad, h->Guu, m_pad, 1, 1, m_pad);
  g4.kb_mode = KB_TJ;
  g4.alpha = 0.5;
  if ((r = sb_gemm(h, false, true, g4))) return r;
  if ((r = sb_chain(h, false, h->Xs, n, h->Guf1, n_pad, ntn, h->partA, h->zpA))) return r;
  if ((r = sb_chain(h, true, h->Zs, m, h->Guu, m_pad, 1, h->partB, h->zpB))) return r;
  sgpr_finalize_kernel<<<dim3(8, P), 256, 0, s>>>(h->scal, h->logdetB, 1, h->partA, ntn, h->partB, 1, 1 + D, h->zpA, ntn, h->zpB, 1,
                                                h->theta, n, m, m_pad, D, 1, h->result, bs);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

// |y|^2 per model, summed like sgpr_scalars_kernel does (256 strided partial sums, then a fixed tree)
__global__ void sf_yy_kernel(const double* __restrict__ yv, int n, long n_pad, double* __restrict__ yy) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const double* y = yv + blockIdx.x * n_pad;
  double s = 0.0;
  for (int e = tid; e < n; e += 256) s = fma(y[e], y[e], s);
  red[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) yy[blockIdx.x] = red[0];
}

template <int KID>
int sf_attrs_t() {
  int r;
  if ((r = opt_in_smem(sf_prep_kernel<KID>, 220 * 1024)) || (r = opt_in_smem(sf_prep_train_kernel<KID>, 220 * 1024)) ||
      (r = opt_in_smem(sf_predict_kernel<KID>, 200 * 1024)) ||  // (+ 24 KB static)
      (r = opt_in_smem(sf_forward_kernel<KID>, 220 * 1024)) || (r = opt_in_smem(sf_mid_kernel<KID>, 220 * 1024)) ||
      (r = opt_in_smem(sf_backward_kernel<KID>, 220 * 1024)))
    return r;
  return 0;
}

// shared-memory opt-in of the fused kernels of one covariance function on the current device (idempotent)
int sf_prepare(int kid) {
  switch (kid) {
    case K_RBF: return sf_attrs_t<K_RBF>();
    case K_MATERN12: return sf_attrs_t<K_MATERN12>();
    case K_MATERN32: return sf_attrs_t<K_MATERN32>();
    case K_MATERN52: return sf_attrs_t<K_MATERN52>();
    case K_EXPONENTIAL: return sf_attrs_t<K_EXPONENTIAL>();
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

// The models [model0, model0 + count) of the batch on stream s (a "lane"): every per-model pointer advanced by model0 strides.
struct SfLane {
  cudaStream_t s;
  int model0, count;
  cudaEvent_t after_forward;  // recorded behind the forward pass when not null (staggers the second lane)
};

SfArgs sf_shift(const SfArgs& a0, int model0) {
  SfArgs a = a0;
  const long o = (long)model0 * a0.bs;
  a.yv += (long)model0 * a0.n_pad, a.yy += model0;
  a.theta += o, a.Z += o, a.Zs += o, a.W += o, a.Ap += o, a.Kv += o, a.Fv += o, a.slabs += o, a.aep += o, a.aats += o, a.aes += o,
      a.RW += o, a.WBg += o, a.uvec += o, a.scal += o, a.logdetB += o, a.partA += o, a.zpA += o, a.partB += o, a.zpB += o;
  a.info += o * 2;
  return a;
}

// cfg == nullptr: theta / Z in device memory -> result.  cfg != nullptr: one whole Adam step, u -> u (six launches).
template <int KID>
int sf_launch_t(gpras_sgpr_batch* h, const SfArgs& a0, const SgprAdamCfg* cfg, const SfLane& lane) {
  cudaStream_t s = lane.s;
  const int P = lane.count;
  const SfArgs a = sf_shift(a0, lane.model0);
  const long o = (long)lane.model0 * a0.bs;
  const int tile_smem = sf_tile_doubles(a.D, a.mp) * (int)sizeof(double);
  if (cfg)
    sf_prep_train_kernel<KID><<<dim3(1, P), SF_THREADS, sf_prep_smem(a.D), s>>>(a, h->au + o, cfg->n_ls, cfg->transform, cfg->noise_floor);
  else
    sf_prep_kernel<KID><<<dim3(1, P), SF_THREADS, sf_prep_smem(a.D), s>>>(a);
  sf_forward_kernel<KID><<<dim3(a.ntn, P), SF_THREADS, tile_smem, s>>>(a);
  if (lane.after_forward) CU(cudaEventRecord(lane.after_forward, s));
  sf_reduce_kernel<<<dim3((a.mp * a.mp + a.mp + SF_THREADS - 1) / SF_THREADS, P), SF_THREADS, 0, s>>>(a);
  sf_mid_kernel<KID><<<dim3(1, P), SF_THREADS, sf_mid_smem(a.D), s>>>(a);
  sf_backward_kernel<KID><<<dim3(a.ntn, P), SF_THREADS, tile_smem, s>>>(a);
  if (cfg)
    sf_finish_train_kernel<<<dim3(1, P), SF_THREADS, 0, s>>>(a, h->result + o, h->au + o, h->amom + o, h->avel + o, h->ast + o,
                                                            h->losses + lane.model0, h->p, *cfg);
  else
    sf_finalize_kernel<<<dim3(1, P), SF_THREADS, 0, s>>>(a, h->result + o);
  h->launches += 6;
  CU(cudaGetLastError());
  return 0;
}

int sf_record_lane(gpras_sgpr_batch* h, double jitter, const SgprAdamCfg* cfg, const SfLane& lane) {
  SfArgs a = h->fa;
  a.jitter = jitter;
  switch (h->kid) {
    case K_RBF: return sf_launch_t<K_RBF>(h, a, cfg, lane);
    case K_MATERN12: return sf_launch_t<K_MATERN12>(h, a, cfg, lane);
    case K_MATERN32: return sf_launch_t<K_MATERN32>(h, a, cfg, lane);
    case K_MATERN52: return sf_launch_t<K_MATERN52>(h, a, cfg, lane);
    case K_EXPONENTIAL: return sf_launch_t<K_EXPONENTIAL>(h, a, cfg, lane);
  }
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

// conditioning for prediction: the first four kernels of the evaluation (W, WB, u per model)
template <int KID>
int sf_condition_t(gpras_sgpr_batch* h, const SfArgs& a) {
  cudaStream_t s = h->stream;
  const int P = h->p;
  const int tile_smem = sf_tile_doubles(a.D, a.mp) * (int)sizeof(double);
  sf_prep_kernel<KID><<<dim3(1, P), SF_THREADS, sf_prep_smem(a.D), s>>>(a);
  sf_forward_kernel<KID><<<dim3(a.ntn, P), SF_THREADS, tile_smem, s>>>(a);
  sf_reduce_kernel<<<dim3((a.mp * a.mp + a.mp + SF_THREADS - 1) / SF_THREADS, P), SF_THREADS, 0, s>>>(a);
  sf_mid_kernel<KID><<<dim3(1, P), SF_THREADS, sf_mid_smem(a.D), s>>>(a);
  h->launches += 4;
  CU(cudaGetLastError());
  return 0;
}

template <int KID>
int sf_predict_t(gpras_sgpr_batch* h, const SfArgs& a, int t) {
  const int tile_smem = sf_tile_doubles(a.D, a.mp) * (int)sizeof(double);
  sf_predict_kernel<KID><<<dim3((t + SF_TN - 1) / SF_TN, h->p), SF_THREADS, tile_smem, h->stream>>>(a, h->xt, t, h->pmean, h->pvar, h->p);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

#define GPRAS_SF_DISPATCH(FN, ...)                                   \
  switch (h->kid) {                                                  \
    case K_RBF: return FN<K_RBF>(__VA_ARGS__);                       \
    case K_MATERN12: return FN<K_MATERN12>(__VA_ARGS__);             \
    case K_MATERN32: return FN<K_MATERN32>(__VA_ARGS__);             \
    case K_MATERN52: return FN<K_MATERN52>(__VA_ARGS__);             \
    case K_EXPONENTIAL: return FN<K_EXPONENTIAL>(__VA_ARGS__);       \
  }                                                                  \
  return fail(GPRAS_E_ARG, "unknown kernel id")

int sf_condition(gpras_sgpr_batch* h, double jitter) {
  SfArgs a = h->fa;
  a.jitter = jitter;
  GPRAS_SF_DISPATCH(sf_condition_t, h, a);
}

int sf_predict(gpras_sgpr_batch* h, int t) {
  GPRAS_SF_DISPATCH(sf_predict_t, h, h->fa, t);
}

int sf_record_eval(gpras_sgpr_batch* h, double jitter, const SgprAdamCfg* cfg = nullptr) {
  return sf_record_lane(h, jitter, cfg, SfLane{h->stream, 0, h->p, nullptr});
}

int sb_eval(gpras_sgpr_batch* h, double jitter) { return h->fused ? sf_record_eval(h, jitter) : sb_record_eval(h, jitter); }

}  // namespace

extern "C" {

int gpras_sgpr_batch_destroy(gpras_sgpr_batch* h);

int gpras_sgpr_batch_create(gpras_sgpr_batch** out, int device, int kernel_id, int n, int d, int m, int p) {
  if (!out || n <= 0 || d <= 0 || m <= 0 || p <= 0) return fail(GPRAS_E_ARG, "bad shape");
  if (kernel_id < 0 || kernel_id > 4) return fail(GPRAS_E_ARG, "unknown kernel id");
  if (d > 64) return fail(GPRAS_E_ARG, "d > 64 features is not supported");
  if (m > 128) return fail(GPRAS_E_ARG, "the batched trainer holds at most 128 inducing points per model");
  if (p > 65535) return fail(GPRAS_E_ARG, "more than 65535 models");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  int rc;
  if ((rc = prepare_device())) return rc;
  gpras_sgpr_batch* h = new gpras_sgpr_batch();
  h->device = device, h->kid = kernel_id, h->n = n, h->d = d, h->m = m, h->p = p;
  h->n_pad = round_up(n, 128), h->ntn = h->n_pad / 128;
  h->ks_aat = round_up((h->n_pad + 31) / 32, 128);
  if (h->ks_aat < 512) h->ks_aat = 512;
  h->nz_aat = (h->n_pad + h->ks_aat - 1) / h->ks_aat;
  h->nu = 2 + d + m * d;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t mm = (size_t)h->m_pad * h->m_pad, mn = (size_t)h->m_pad * h->n_pad, mr = (size_t)h->m_pad * h->r_pad;
  struct Carve {
    double** ptr;
    size_t count;
  };
  h->fused = m <= SF_MP && d <= SF_MAX_D && !getenv("GPRAS_B200_SGPR_UNFUSED");
  std::vector<Carve> parts;
  SfArgs& fa = h->fa;
  if (h->fused) {
    // one CTA per tile of 128 training rows and model; every tile leaves a partial M x M matrix that sf_reduce_kernel adds up
    fa.n = n, fa.n_pad = h->n_pad, fa.D = d, fa.m = m, fa.mp = round_up(m, 8), fa.ntn = h->ntn;
    const size_t nt = (size_t)h->ntn;
    parts = {{&h->theta, (size_t)2 + d}, {&h->Z, (size_t)m * d}, {&fa.Zs, (size_t)SF_MP * d}, {&fa.W, (size_t)SF_MP * SF_MP},
             {&fa.Ap, (size_t)fa.mp * h->n_pad}, {&fa.Kv, (size_t)fa.mp * h->n_pad}, {&fa.Fv, (size_t)fa.mp * h->n_pad},
             {&fa.slabs, nt * SF_MP * SF_MP}, {&fa.aep, nt * SF_MP}, {&fa.aats, (size_t)SF_MP * SF_MP}, {&fa.aes, (size_t)SF_MP},
             {&fa.RW, (size_t)SF_MP * SF_MP}, {&fa.WBg, (size_t)SF_MP * SF_MP}, {&fa.uvec, (size_t)SF_MP}, {&fa.scal, 8}, {&fa.logdetB, 1},
             {&fa.partA, nt * (1 + d)}, {&fa.zpA, nt * SF_MP * d}, {&fa.partB, (size_t)1 + d},
             {&fa.zpB, (size_t)SF_MP * d}, {&h->result, 3 + d + (size_t)m * d}, {&h->au, (size_t)h->nu}, {&h->amom, (size_t)h->nu},
             {&h->avel, (size_t)h->nu}, {&h->ast, (size_t)gpras::AST}};
  } else {
    parts = {
      {&h->X, (size_t)h->n_pad * d}, {&h->Xs, (size_t)h->n_pad * d}, {&h->Z, (size_t)h->m_pad * d}, {&h->Zs, (size_t)h->m_pad * d},
      {&h->Y, (size_t)h->n_pad * h->r_pad}, {&h->Kuf, mn}, {&h->Kuu, mm}, {&h->WL, mm}, {&h->Ap, mn}, {&h->slabs, mm * h->nz_aat},
      {&h->AATs, mm}, {&h->B, mm}, {&h->WB, mm}, {&h->Binv, mm}, {&h->Rm, mm}, {&h->RA, mm}, {&h->RW, mm}, {&h->T1, mm}, {&h->Guu, mm},
      {&h->Guf1, mn}, {&h->ae, mr}, {&h->c, mr}, {&h->chat, mr}, {&h->u, mr}, {&h->skinny, (size_t)SKINNY_MAX_SLABS * mr},
      {&h->theta, (size_t)2 + d}, {&h->logdetL, 1}, {&h->logdetB, 1}, {&h->scal, 8}, {&h->partA, (size_t)h->ntn * (1 + d)},
      {&h->partB, (size_t)1 + d}, {&h->zpA, (size_t)h->ntn * h->m_pad * d}, {&h->zpB, (size_t)h->m_pad * d},
      {&h->result, 3 + d + (size_t)m * d}, {&h->au, (size_t)h->nu}, {&h->amom, (size_t)h->nu}, {&h->avel, (size_t)h->nu},
      {&h->ast, (size_t)gpras::AST}};
  }
  size_t per_model = 64;
  for (const Carve& c : parts) per_model += (c.count * sizeof(double) + 255) / 256 * 256;
  // (fused layout: the SfArgs pointers were carved through references into h->fa)
  h->bs = (long)(per_model / sizeof(double));
  {
    cudaError_t e = cudaMalloc(&h->arena, per_model * p);
    if (e != cudaSuccess) {
      gpras_sgpr_batch_destroy(h);
      return fail(GPRAS_E_NOMEM, "cudaMalloc", e);
    }
    char* cur = (char*)h->arena;
    h->info = (int*)cur;
    cur += 64;
    for (const Carve& c : parts) {
      *c.ptr = (double*)cur;
      cur += (c.count * sizeof(double) + 255) / 256 * 256;
    }
  }
  // zero everything once: padding rows of X / Z / Y and the never-written upper triangles of WL / WB must be zero
  CU(cudaMemsetAsync(h->arena, 0, per_model * p, h->stream));
  if (h->fused) {
    if ((rc = dalloc(&h->Xsh, (size_t)h->n_pad * d)) || (rc = dalloc(&h->yv, (size_t)p * h->n_pad)) || (rc = dalloc(&h->yy, p))) {
      gpras_sgpr_batch_destroy(h);
      return rc;
    }
    CU(cudaMemsetAsync(h->Xsh, 0, sizeof(double) * h->n_pad * d, h->stream));
    CU(cudaMemsetAsync(h->yv, 0, sizeof(double) * (size_t)p * h->n_pad, h->stream));
    if ((rc = sf_prepare(kernel_id))) {
      gpras_sgpr_batch_destroy(h);
      return rc;
    }
    fa.X = h->Xsh, fa.yv = h->yv, fa.yy = h->yy, fa.theta = h->theta, fa.Z = h->Z, fa.info = h->info, fa.bs = h->bs;
  }
  CU(cudaMallocHost(&h->h_pinned, sizeof(double) * (size_t)p * (h->nu + gpras::AST)));
  CU(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

int gpras_sgpr_batch_destroy(gpras_sgpr_batch* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& g : h->graphs)
    if (g.second) cudaGraphExecDestroy(g.second);
  for (auto& g : h->chunk_graphs) cudaGraphExecDestroy(g.exec);
  for (int l = 1; l < gpras_sgpr_batch::MAX_LANES; l++) {
    if (h->ev_stagger[l]) cudaEventDestroy(h->ev_stagger[l]);
    if (h->ev_join[l]) cudaEventDestroy(h->ev_join[l]);
    if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
  }
  if (h->arena) cudaFree(h->arena);
  if (h->Xsh) cudaFree(h->Xsh);
  if (h->yv) cudaFree(h->yv);
  if (h->yy) cudaFree(h->yy);
  if (h->xt) cudaFree(h->xt);
  if (h->pmean) cudaFree(h->pmean);
  if (h->pvar) cudaFree(h->pvar);
  if (h->losses) cudaFree(h->losses);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

// x: n x d (shared by the models), y: n x p row-major (column b is model b's target)
int gpras_sgpr_batch_set_data(gpras_sgpr_batch* h, const double* x, const double* y) {
  if (!h || !x || !y) return fail(GPRAS_E_ARG, "null argument");
  DeviceGuard guard(h->device);
  if (h->fused) {  // one copy of the inputs; targets transposed on the host, one vector per model
    std::vector<double> yt((size_t)h->p * h->n_pad, 0.0);
    for (int i = 0; i < h->n; i++)
      for (int b = 0; b < h->p; b++) yt[(size_t)b * h->n_pad + i] = y[(size_t)i * h->p + b];
    CU(cudaMemcpyAsync(h->Xsh, x, sizeof(double) * h->n * h->d, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->yv, yt.data(), sizeof(double) * yt.size(), cudaMemcpyHostToDevice, h->stream));
    sf_yy_kernel<<<h->p, 256, 0, h->stream>>>(h->yv, h->n, h->n_pad, h->yy);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    h->has_data = true;
    h->conditioned = false;
    return 0;
  }
  for (int b = 0; b < h->p; b++) {
    CU(cudaMemcpyAsync(h->X + b * h->bs, x, sizeof(double) * h->n * h->d, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpy2DAsync(h->Y + b * h->bs, sizeof(double) * h->r_pad, y + b, sizeof(double) * h->p, sizeof(double), h->n,
                         cudaMemcpyHostToDevice, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  h->has_data = true;
  return 0;
}

// Bound and gradient of every model at host-supplied (theta [p x (2+d)], z [p x m x d]); outputs elbo [p],
// grad_theta [p x (2+d)] (d/dlog theta), grad_z [p x m x d], info [p] (0, or the failing pivot of a model).
int gpras_sgpr_batch_elbo_grad(gpras_sgpr_batch* h, const double* theta, const double* z, double jitter, double* elbo,
                               double* grad_theta, double* grad_z, int* info) {
  if (!h || !theta || !z || !elbo || !info) return fail(GPRAS_E_ARG, "null argument");
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int D = h->d, m = h->m, P = h->p;
  const size_t pitch = sizeof(double) * h->bs;
  CU(cudaMemcpy2DAsync(h->theta, pitch, theta, sizeof(double) * (2 + D), sizeof(double) * (2 + D), P, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpy2DAsync(h->Z, pitch, z, sizeof(double) * m * D, sizeof(double) * m * D, P, cudaMemcpyHostToDevice, s));
  CU(cudaMemset2DAsync(h->info, pitch, 0, sizeof(int), P, s));
  h->launches = 0;
  h->conditioned = false;
  int r;
  if ((r = sb_eval(h, jitter))) return r;
  h->warmed = true;
  const size_t nres = 3 + D + (size_t)m * D;
  std::vector<double> res(nres * P);
  CU(cudaMemcpy2DAsync(res.data(), sizeof(double) * nres, h->result, pitch, sizeof(double) * nres, P, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpy2DAsync(info, sizeof(int), h->info, pitch, sizeof(int), P, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int b = 0; b < P; b++) {
    const double* rb = res.data() + nres * b;
    elbo[b] = rb[0];
    if (grad_theta) memcpy(grad_theta + (size_t)b * (2 + D), rb + 1, sizeof(double) * (2 + D));
    if (grad_z) memcpy(grad_z + (size_t)b * m * D, rb + 3 + D, sizeof(double) * m * D);
  }
  return 0;
}

// One Adam stage (gpr.py:147-173) for all models on the device.
//   u          p x nu in/out, nu = 2 + n_ls + m d: unconstrained [variance, noise, lengthscale(s), Z] per model
//   n_ls       1 (one lengthscale, the reference) or d (one per feature)
//   train_*    which variables this stage updates (gpflow.set_trainable choreography of the recipes, gpr.py:112-127);
//              the log prior of the hyperparameters is part of the loss only while they are trainable
//   transform  0: value = softplus(u) (+ noise_floor for the noise), GPflow; 1: value = exp(u)
//   losses     max_iter x p out: the loss every model saw at every step (NaN where it had stopped)
//   iters      p out: steps taken;  info p out: 0, or the failing pivot of a model whose Kuu / B lost positive definiteness
int gpras_sgpr_batch_train(gpras_sgpr_batch* h, double* u, int n_ls, int train_hypers, int train_z, int max_iter, double lr,
                           double jitter, int transform, int priors, double noise_floor, int rule, double* losses, int* iters,
                           int* info);

int gpras_sgpr_batch_adam(gpras_sgpr_batch* h, double* u, int n_ls, int train_hypers, int train_z, int max_iter, double lr,
                          double jitter, int transform, int priors, double noise_floor, double* losses, int* iters, int* info) {
  return gpras_sgpr_batch_train(h, u, n_ls, train_hypers, train_z, max_iter, lr, jitter, transform, priors, noise_floor, 0, losses,
                                iters, info);
}

// The same loop with the update rule as a parameter: rule 0 = Adam (above), 1 = Keras Adadelta (_optimize_adadelta,
// gpr.py:176-192: fixed max_iter steps, no early stopping).
int gpras_sgpr_batch_train(gpras_sgpr_batch* h, double* u, int n_ls, int train_hypers, int train_z, int max_iter, double lr,
                           double jitter, int transform, int priors, double noise_floor, int rule, double* losses, int* iters,
                           int* info) {
  if (!h || !u || !iters || !info) return fail(GPRAS_E_ARG, "null argument");
  if (rule != 0 && rule != 1) return fail(GPRAS_E_ARG, "unknown update rule");
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  if ((n_ls != 1 && n_ls != h->d) || max_iter < 0 || (transform != 0 && transform != 1))
    return fail(GPRAS_E_ARG, "bad trainer configuration");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int D = h->d, m = h->m, P = h->p;
  const int nu = 2 + n_ls + m * D;
  const size_t pitch = sizeof(double) * h->bs;
  h->conditioned = false;
  const SgprAdamCfg cfg{n_ls, train_hypers != 0, train_z != 0, transform, priors != 0, rule, lr, jitter, noise_floor};
  int r;
  if (max_iter > h->losses_cap) {
    for (auto& g : h->graphs)  // the loss-history pointer is a captured kernel argument
      if (g.second) cudaGraphExecDestroy(g.second);
    h->graphs.clear();
    for (auto& g : h->chunk_graphs) cudaGraphExecDestroy(g.exec);
    h->chunk_graphs.clear();
    CU(cudaStreamSynchronize(s));
    if (h->losses) cudaFree(h->losses);
    h->losses = nullptr;
    const int cap = max_iter > 256 ? max_iter : 256;
    if ((r = dalloc(&h->losses, (size_t)cap * P))) return r;
    h->losses_cap = cap;
  }
  memcpy(h->h_pinned, u, sizeof(double) * (size_t)P * nu);
  CU(cudaMemcpy2DAsync(h->au, pitch, h->h_pinned, sizeof(double) * nu, sizeof(double) * nu, P, cudaMemcpyHostToDevice, s));
  sgpr_batch_reset_kernel<<<P, 128, 0, s>>>(h->amom, h->avel, h->ast, nu, h->bs);
  if (max_iter > 0) fill_nan_kernel<<<(unsigned)(((long)max_iter * P + 255) / 256), 256, 0, s>>>(h->losses, (long)max_iter * P);
  CU(cudaGetLastError());
  auto record_step = [&]() -> int {
    if (h->fused) return sf_record_eval(h, jitter, &cfg);
    sgpr_batch_unpack_kernel<<<P, 128, 0, s>>>(h->au, h->theta, h->Z, h->info, D, m, n_ls, transform, noise_floor, h->bs);
    h->launches++;
    CU(cudaGetLastError());
    int rr = sb_eval(h, jitter);
    if (rr) return rr;
    sgpr_batch_adam_kernel<<<P, 128, 0, s>>>(h->result, h->info, h->au, h->amom, h->avel, h->ast, h->losses, P, D, m, cfg, h->bs);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
  };
  if (!h->warmed && max_iter > 0) {  // first evaluation eagerly (kernel attributes); its results are recomputed by step 1
    sgpr_batch_unpack_kernel<<<P, 128, 0, s>>>(h->au, h->theta, h->Z, h->info, D, m, n_ls, transform, noise_floor, h->bs);
    CU(cudaGetLastError());
    if ((r = sb_eval(h, jitter))) return r;
    h->warmed = true;
  }
  double* h_st = h->h_pinned + (size_t)P * h->nu;
  auto all_stopped = [&](bool* stopped) -> int {  // has every model stopped early?
    CU(cudaMemcpy2DAsync(h_st, sizeof(double) * gpras::AST, h->ast, pitch, sizeof(double) * gpras::AST, P, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    bool any = false;
    for (int b = 0; b < P; b++) any = any || h_st[(size_t)b * gpras::AST + 2] != 0.0;
    *stopped = !any;
    return 0;
  };
  const bool use_graphs = !getenv("GPRAS_B200_NO_GRAPHS");
  bool chunked_done = false;
  if (h->fused && use_graphs && max_iter > 0) {
    // Fused path: graphs of up to CHUNK consecutive steps.  With four or more models the batch runs as two lanes (halves of the
    // models on two streams, the second lane starting behind the first lane's first forward pass): one lane's one-CTA-per-model
    // kernels (Cholesky chains of prep / mid) then overlap the other lane's tile passes, and each tile pass is a single wave.
    constexpr int CHUNK = 25;
    int lanes = P >= 4 ? 2 : 1;
    if (const char* e = getenv("GPRAS_B200_SGPR_LANES")) lanes = atoi(e);
    if (lanes > gpras_sgpr_batch::MAX_LANES) lanes = gpras_sgpr_batch::MAX_LANES;
    if (lanes > P) lanes = P;
    if (lanes < 1) lanes = 1;
    h->lane_stream[0] = s;
    for (int l = 1; l < lanes; l++)
      if (!h->lane_stream[l]) {
        CU(cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_stagger[l], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_join[l], cudaEventDisableTiming));
      }
    auto chunk_graph = [&](int steps, cudaGraphExec_t* out) -> int {
      for (auto& g : h->chunk_graphs)
        if (g.cfg == cfg && g.steps == steps && g.lanes == lanes) {
          *out = g.exec;
          return 0;
        }
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      h->launches = 0;
      CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int rr = 0;
      for (int it = 0; it < steps && rr == 0; it++) {
        for (int l = 0; l < lanes && rr == 0; l++) {  // lane l: models [l P / lanes, (l + 1) P / lanes)
          const int m0 = (int)((long)l * P / lanes), m1 = (int)((long)(l + 1) * P / lanes);
          // lane l + 1 starts behind lane l's first forward pass
          cudaEvent_t after = (it == 0 && l + 1 < lanes) ? h->ev_stagger[l + 1] : nullptr;
          if (it == 0 && l > 0 && cudaStreamWaitEvent(h->lane_stream[l], h->ev_stagger[l], 0) != cudaSuccess) rr = GPRAS_E_CUDA;
          if (rr == 0) rr = sf_record_lane(h, jitter, &cfg, SfLane{h->lane_stream[l], m0, m1 - m0, after});
        }
      }
      for (int l = 1; l < lanes && rr == 0; l++)
        if (cudaEventRecord(h->ev_join[l], h->lane_stream[l]) != cudaSuccess || cudaStreamWaitEvent(s, h->ev_join[l], 0) != cudaSuccess)
          rr = GPRAS_E_CUDA;
      cudaError_t e = cudaStreamEndCapture(s, &graph);
      if (rr == 0 && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (rr != 0 || e != cudaSuccess || !exec) {
        cudaGetLastError();
        return fail(GPRAS_E_CUDA, "capturing the trainer's graph failed", e);
      }
      h->launches /= steps;  // launches of one step (all lanes)
      if (h->chunk_graphs.size() >= 8) {
        cudaGraphExecDestroy(h->chunk_graphs.front().exec);
        h->chunk_graphs.erase(h->chunk_graphs.begin());
      }
      h->chunk_graphs.push_back({cfg, steps, lanes, exec});
      *out = exec;
      return 0;
    };
    chunked_done = true;
    for (int done = 0; done < max_iter;) {
      const int steps = max_iter - done < CHUNK ? max_iter - done : CHUNK;
      cudaGraphExec_t exec = nullptr;
      if ((r = chunk_graph(steps, &exec))) {
        if (done > 0) return r;
        chunked_done = false;  // nothing has run yet: eager launches below (still the CUDA path, never a CPU one)
        break;
      }
      CU(cudaGraphLaunch(exec, s));
      done += steps;
      if (done < max_iter && rule == 0) {
        bool stopped = false;
        if ((r = all_stopped(&stopped))) return r;
        if (stopped) break;
      }
    }
  }
  if (!chunked_done) {
    cudaGraphExec_t exec = nullptr;
    for (auto& g : h->graphs)
      if (g.first == cfg) exec = g.second;
    if (!exec && max_iter > 0 && use_graphs && !h->fused) {
      cudaGraph_t graph = nullptr;
      h->launches = 0;
      CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      r = record_step();
      cudaError_t e = cudaStreamEndCapture(s, &graph);
      if (r == 0 && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (r != 0 || e != cudaSuccess || !exec) {
        cudaGetLastError();
        exec = nullptr;  // eager launches below (still the CUDA path)
      } else {
        if (h->graphs.size() >= 8) {
          cudaGraphExecDestroy(h->graphs.front().second);
          h->graphs.erase(h->graphs.begin());
        }
        h->graphs.emplace_back(cfg, exec);
      }
    }
    for (int it = 0; it < max_iter; it++) {
      if (exec)
        CU(cudaGraphLaunch(exec, s));
      else {
        h->launches = 0;
        if ((r = record_step())) return r;
      }
      if ((it + 1) % 32 == 0 && it + 1 < max_iter) {
        bool stopped = false;
        if ((r = all_stopped(&stopped))) return r;
        if (stopped) break;
      }
    }
  }
  CU(cudaMemcpy2DAsync(h->h_pinned, sizeof(double) * nu, h->au, pitch, sizeof(double) * nu, P, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpy2DAsync(h_st, sizeof(double) * gpras::AST, h->ast, pitch, sizeof(double) * gpras::AST, P, cudaMemcpyDeviceToHost, s));
  if (losses && max_iter > 0)
    CU(cudaMemcpyAsync(losses, h->losses, sizeof(double) * (size_t)max_iter * P, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  memcpy(u, h->h_pinned, sizeof(double) * (size_t)P * nu);
  for (int b = 0; b < P; b++) {
    iters[b] = (int)h_st[(size_t)b * gpras::AST + 3];
    info[b] = (int)h_st[(size_t)b * gpras::AST + 4];
  }
  return 0;
}

// predict_y (gpr.py:337) of all models: condition at host-supplied (theta [p x (2+d)], z [p x m x d]); info [p] as in elbo_grad.
// Fused layout only (m <= 64, d <= 32); otherwise GPRAS_E_ARG and the caller uses one gpras_sgpr handle per model.
int gpras_sgpr_batch_condition(gpras_sgpr_batch* h, const double* theta, const double* z, double jitter, int* info) {
  if (!h || !theta || !z || !info) return fail(GPRAS_E_ARG, "null argument");
  if (!h->fused) return fail(GPRAS_E_ARG, "batched prediction needs m <= 64 and d <= 32");
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int D = h->d, m = h->m, P = h->p;
  const size_t pitch = sizeof(double) * h->bs;
  h->conditioned = false;
  CU(cudaMemcpy2DAsync(h->theta, pitch, theta, sizeof(double) * (2 + D), sizeof(double) * (2 + D), P, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpy2DAsync(h->Z, pitch, z, sizeof(double) * m * D, sizeof(double) * m * D, P, cudaMemcpyHostToDevice, s));
  h->launches = 0;
  int r;
  if ((r = sf_condition(h, jitter))) return r;
  CU(cudaMemcpy2DAsync(info, sizeof(int), h->info, pitch, sizeof(int), P, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int b = 0; b < P; b++)
    if (info[b]) return 0;  // reported per model; the handle stays unconditioned
  h->conditioned = true;
  return 0;
}

// xs: t x d host array; mean, var: t x p host arrays (variance includes the likelihood noise).
int gpras_sgpr_batch_predict(gpras_sgpr_batch* h, const double* xs, int t, double* mean, double* var) {
  if (!h || !xs || t < 0 || !mean || !var) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->conditioned) return fail(GPRAS_E_STATE, "condition() has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int D = h->d, P = h->p;
  constexpr int CHUNK = 1 << 16;  // test rows per launch
  int r;
  if (h->pred_cap == 0) {
    if ((r = dalloc(&h->xt, (size_t)CHUNK * D)) || (r = dalloc(&h->pmean, (size_t)CHUNK * P)) || (r = dalloc(&h->pvar, (size_t)CHUNK * P)))
      return r;
    h->pred_cap = CHUNK;
  }
  h->launches = 0;
  for (int t0 = 0; t0 < t; t0 += CHUNK) {
    const int tb = t - t0 < CHUNK ? t - t0 : CHUNK;
    CU(cudaMemcpyAsync(h->xt, xs + (size_t)t0 * D, sizeof(double) * (size_t)tb * D, cudaMemcpyHostToDevice, s));
    if ((r = sf_predict(h, tb))) return r;
    CU(cudaMemcpyAsync(mean + (size_t)t0 * P, h->pmean, sizeof(double) * (size_t)tb * P, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(var + (size_t)t0 * P, h->pvar, sizeof(double) * (size_t)tb * P, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return 0;
}

int gpras_sgpr_batch_last_launches(gpras_sgpr_batch* h) { return h ? h->launches : 0; }

}  // extern "C"
