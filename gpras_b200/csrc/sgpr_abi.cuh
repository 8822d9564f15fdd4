// Sparse (inducing-point) model: host orchestration and C ABI (gpras_sgpr_*).  Compiled as sgpr.cu.
// Math and notation: sgpr_kernels.cuh / oracle/sgpr_analytic.py.  Replaces gpflow.models.SGPR as built at
// gpras/gpr.py:293-308: training_loss + gradient (hyperparameters and inducing inputs) and predict_y.
#pragma once
#include "host_common.cuh"
#include "sgpr_kernels.cuh"

struct gpras_sgpr {
  int device = 0, kid = 0, n = 0, d = 0, m = 0, r = 0, n_pad = 0, m_pad = 0, r_pad = 0, ntm = 0, ntn = 0;
  int nz_aat = 1, ks_aat = 0;
  double jitter = 1e-6;
  cudaStream_t stream = nullptr;
  bool has_data = false, conditioned = false;
  int launches = 0;
  LookAhead la;
  std::vector<double*> owned;
  double *X = nullptr, *Xs = nullptr, *Z = nullptr, *Zs = nullptr, *Y = nullptr;
  double *Kuf = nullptr, *Kuu = nullptr, *WL = nullptr, *Ap = nullptr, *slabs = nullptr, *AATs = nullptr, *B = nullptr,
         *WB = nullptr, *Binv = nullptr, *T = nullptr, *Rm = nullptr, *RA = nullptr, *RW = nullptr, *T1 = nullptr,
         *Guu = nullptr, *Guf1 = nullptr;
  double *ae = nullptr, *c = nullptr, *chat = nullptr, *u = nullptr, *skinny = nullptr;
  double *theta = nullptr, *logdetL = nullptr, *logdetB = nullptr, *scal = nullptr, *partA = nullptr, *partB = nullptr,
         *zpA = nullptr, *zpB = nullptr, *result = nullptr;
  int* info = nullptr;
  double *h_theta = nullptr, *h_z = nullptr, *h_result = nullptr;
  int* h_info = nullptr;
  void *arena = nullptr, *h_arena = nullptr;  // one device / one pinned allocation carved into the training buffers
  // asynchronous objective + CUDA-graph replay (index: want_grad); the jitter is baked into the captured kernel arguments
  bool pending = false, pending_grad = false, use_graphs = true, graph_failed = false, eager_done[2] = {false, false};
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  double graph_jitter = -1.0;
  int graph_launches[2] = {0, 0};
  // prediction
  double *Xt = nullptr, *Xts = nullptr, *Kus = nullptr, *tmp1 = nullptr, *p2 = nullptr, *q1 = nullptr, *mean = nullptr,
         *var = nullptr, *varm = nullptr;
};

namespace {

constexpr int SGPR_TB = 2048;

int sgpr_alloc(gpras_sgpr* h, double** p, size_t count) {
  int r = dalloc(p, count);
  if (r == 0) h->owned.push_back(*p);
  return r;
}

// generic split-k product on the L shape: partial slabs reduced elsewhere
int launch_splitk_L(cudaStream_t s, bool akm, bool bkm, GemmDesc d, double* slabs, long slab, int ks, int nz, int* launches) {
  d.k_split = ks;
  d.splitC = slab;
  d.C = slabs;
  return launch_gemm(s, akm, bkm, d, 1, launches, SHAPE_L, nz);
}

template <int KID>
int launch_chain(cudaStream_t s, bool uu, const double* Zs, int m, const double* Xs, int n, int D, const double* G1, long ldg,
                 const double* u, long ldu, const double* Y, long ldy, int R, const double* theta, int tiles_y, int tiles_x,
                 double* part, int ncols, double* zpart, int m_pad) {
  const int smem = (2 * D * CT_LD + CT * D + 2 * R * CT_LD) * (int)sizeof(double);
  const int grid = tiles_y * tiles_x;
  static std::atomic<bool> attr_done[64] = {};  // benign if two threads both set the (idempotent) attributes
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    int r;
    if ((r = opt_in_smem(sgpr_chain_kernel<KID, 16>, 220 * 1024)) || (r = opt_in_smem(sgpr_chain_kernel<KID, 64>, 220 * 1024)))
      return r;
    attr_done[dev] = true;
  }
  if (D <= 16)
    sgpr_chain_kernel<KID, 16><<<grid, PT_THREADS, smem, s>>>(uu, Zs, m, Xs, n, D, G1, ldg, u, ldu, Y, ldy, R, theta, tiles_x, part,
                                                              ncols, zpart, m_pad);
  else
    sgpr_chain_kernel<KID, 64><<<grid, PT_THREADS, smem, s>>>(uu, Zs, m, Xs, n, D, G1, ldg, u, ldu, Y, ldy, R, theta, tiles_x, part,
                                                              ncols, zpart, m_pad);
  CU(cudaGetLastError());
  return 0;
}

int dispatch_chain(int kid, cudaStream_t s, bool uu, const double* Zs, int m, const double* Xs, int n, int D, const double* G1,
                   long ldg, const double* u, long ldu, const double* Y, long ldy, int R, const double* theta, int tiles_y,
                   int tiles_x, double* part, int ncols, double* zpart, int m_pad) {
#define GPRAS_CH(K) return launch_chain<K>(s, uu, Zs, m, Xs, n, D, G1, ldg, u, ldu, Y, ldy, R, theta, tiles_y, tiles_x, part, ncols, zpart, m_pad)
  switch (kid) {
    case K_RBF: GPRAS_CH(K_RBF);
    case K_MATERN12: GPRAS_CH(K_MATERN12);
    case K_MATERN32: GPRAS_CH(K_MATERN32);
    case K_MATERN52: GPRAS_CH(K_MATERN52);
    case K_EXPONENTIAL: GPRAS_CH(K_EXPONENTIAL);
  }
#undef GPRAS_CH
  return fail(GPRAS_E_ARG, "unknown kernel id");
}

// Everything up to c, chat, u (shared by the objective and by prediction).
int sgpr_forward(gpras_sgpr* h, const double* theta, const double* z) {
  cudaStream_t s = h->stream;
  const int n = h->n, D = h->d, m = h->m, n_pad = h->n_pad, m_pad = h->m_pad, ntm = h->ntm, ntn = h->ntn, rp = h->r_pad;
  int r;
  if (theta) {  // NULL: the pinned staging buffers were filled by the caller (graph capture / replay)
    memcpy(h->h_theta, theta, sizeof(double) * (2 + D));
    memcpy(h->h_z, z, sizeof(double) * (size_t)m * D);
  }
  CU(cudaMemcpyAsync(h->theta, h->h_theta, sizeof(double) * (2 + D), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->Z, h->h_z, sizeof(double) * (size_t)m * D, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(h->info, 0, sizeof(int), s));
  scale_features_kernel<<<(unsigned)(((long)n_pad * D + 255) / 256), 256, 0, s>>>(h->X, h->Xs, n, n_pad, D, h->theta);
  scale_features_kernel<<<(unsigned)(((long)m_pad * D + 255) / 256), 256, 0, s>>>(h->Z, h->Zs, m, m_pad, D, h->theta);
  h->launches += 2;
  CU(cudaGetLastError());
  if ((r = dispatch_cov(h->kid, s, h->Zs, m, m_pad, h->Xs, n, n_pad, D, h->theta, h->Kuf, n_pad, 0))) return r;
  if ((r = dispatch_cov(h->kid, s, h->Zs, m, m_pad, h->Zs, m, m_pad, D, h->theta, h->Kuu, m_pad, 1, h->jitter))) return r;
  h->launches += 2;
  // L = chol(Kuu), WL = L^-1
  if ((r = potrf_impl(s, h->la, h->Kuu, m_pad, h->WL, m_pad, m_pad, h->logdetL, h->info, &h->launches))) return r;
  if ((r = trtri_impl(s, h->Kuu, m_pad, h->WL, m_pad, h->T, m_pad, m_pad, &h->launches))) return r;
  // A' = WL Kuf
  {
    GemmDesc g = make_desc(h->WL, m_pad, h->Kuf, n_pad, h->Ap, n_pad, ntm, ntn, m_pad);
    g.ke_mode = KE_TI;
    if ((r = launch_gemm(s, false, true, g, 1, &h->launches))) return r;
  }
  // AATs = A' A'^T / s2, B = I + AATs  (split-k over the N rows)
  {
    GemmDesc g = make_desc(h->Ap, n_pad, h->Ap, n_pad, h->slabs, m_pad, ntm, ntm, n_pad);
    if ((r = launch_splitk_L(s, false, false, g, h->slabs, (long)m_pad * m_pad, h->ks_aat, h->nz_aat, &h->launches))) return r;
    sgpr_finish_b_kernel<<<(unsigned)(((long)m_pad * m_pad + 255) / 256), 256, 0, s>>>(h->slabs, (long)m_pad * m_pad, h->nz_aat,
                                                                                  h->theta, m_pad, h->AATs, h->B);
    h->launches++;
    CU(cudaGetLastError());
  }
  // LB = chol(B), WB = LB^-1
  if ((r = potrf_impl(s, h->la, h->B, m_pad, h->WB, m_pad, m_pad, h->logdetB, h->info, &h->launches))) return r;
  if ((r = trtri_impl(s, h->B, m_pad, h->WB, m_pad, h->T, m_pad, m_pad, &h->launches))) return r;
  // ae = A' Y / s2 ; c = WB ae ; chat = WB^T c ; u = WL^T chat
  {
    GemmDesc g = make_desc(h->Ap, n_pad, h->Y, rp, h->ae, rp, ntm, rp / 32, n_pad);
    if ((r = launch_skinny(s, false, g, h->skinny, m_pad, &h->launches))) return r;
    sgpr_scale_noise_kernel<<<(unsigned)(((long)m_pad * rp + 255) / 256), 256, 0, s>>>(h->ae, (long)m_pad * rp, h->theta);
    h->launches++;
    GemmDesc g2 = make_desc(h->WB, m_pad, h->ae, rp, h->c, rp, ntm, rp / 32, m_pad);
    g2.ke_mode = KE_TI;
    if ((r = launch_skinny(s, false, g2, h->skinny, m_pad, &h->launches))) return r;
    GemmDesc g3 = make_desc(h->WB, m_pad, h->c, rp, h->chat, rp, ntm, rp / 32, m_pad);
    g3.kb_mode = KB_TI;
    if ((r = launch_skinny(s, true, g3, h->skinny, m_pad, &h->launches))) return r;
    GemmDesc g4 = make_desc(h->WL, m_pad, h->chat, rp, h->u, rp, ntm, rp / 32, m_pad);
    g4.kb_mode = KB_TI;
    if ((r = launch_skinny(s, true, g4, h->skinny, m_pad, &h->launches))) return r;
  }
  (void)m;
  return 0;
}

int sgpr_check_info(gpras_sgpr* h) {
  if (*h->h_info != 0) {
    g_err = "Kuu or B lost positive definiteness";
    return *h->h_info;
  }
  return 0;
}

}  // namespace

extern "C" {

int gpras_sgpr_create(gpras_sgpr** out, int device, int kernel_id, int n, int d, int m, int r) {
  if (!out || n <= 0 || d <= 0 || m <= 0 || r <= 0) return fail(GPRAS_E_ARG, "bad shape");
  if (kernel_id < 0 || kernel_id > 4) return fail(GPRAS_E_ARG, "unknown kernel id");
  if (d > 64) return fail(GPRAS_E_ARG, "d > 64 features is not supported");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  int rc;
  if ((rc = prepare_device())) return rc;
  gpras_sgpr* h = new gpras_sgpr();
  h->device = device, h->kid = kernel_id, h->n = n, h->d = d, h->m = m, h->r = r;
  h->n_pad = round_up(n, 128), h->m_pad = round_up(m, 128), h->r_pad = round_up(r, 32);
  h->ntm = h->m_pad / 128, h->ntn = h->n_pad / 128;
  h->ks_aat = round_up((h->n_pad + 31) / 32, 128);
  if (h->ks_aat < 512) h->ks_aat = 512;
  h->nz_aat = (h->n_pad + h->ks_aat - 1) / h->ks_aat;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t mm = (size_t)h->m_pad * h->m_pad, mn = (size_t)h->m_pad * h->n_pad, mr = (size_t)h->m_pad * h->r_pad;
  const size_t sk = (size_t)SKINNY_MAX_SLABS * (h->m_pad > SGPR_TB ? h->m_pad : SGPR_TB) * h->r_pad;
  // one device allocation and one pinned allocation, carved (a handle per model is created when models train in lock-step)
  struct Carve {
    double** p;
    size_t count;
  };
  const Carve parts[] = {
      {&h->X, (size_t)h->n_pad * d}, {&h->Xs, (size_t)h->n_pad * d}, {&h->Z, (size_t)h->m_pad * d}, {&h->Zs, (size_t)h->m_pad * d},
      {&h->Y, (size_t)h->n_pad * h->r_pad}, {&h->Kuf, mn}, {&h->Kuu, mm}, {&h->WL, mm}, {&h->Ap, mn}, {&h->slabs, mm * h->nz_aat},
      {&h->AATs, mm}, {&h->B, mm}, {&h->WB, mm}, {&h->Binv, mm}, {&h->T, mm}, {&h->Rm, mm}, {&h->RA, mm}, {&h->RW, mm}, {&h->T1, mm},
      {&h->Guu, mm}, {&h->Guf1, mn}, {&h->ae, mr}, {&h->c, mr}, {&h->chat, mr}, {&h->u, mr}, {&h->skinny, sk},
      {&h->theta, (size_t)2 + d}, {&h->logdetL, (size_t)h->ntm}, {&h->logdetB, (size_t)h->ntm}, {&h->scal, 8},
      {&h->partA, (size_t)h->ntm * h->ntn * (1 + d)}, {&h->partB, (size_t)h->ntm * h->ntm * (1 + d)},
      {&h->zpA, (size_t)h->ntn * h->m_pad * d}, {&h->zpB, (size_t)h->ntm * h->m_pad * d}, {&h->result, 3 + d + (size_t)m * d}};
  size_t total = 64;
  for (const Carve& c : parts) total += (c.count * sizeof(double) + 255) / 256 * 256;
  {
    cudaError_t e = cudaMalloc(&h->arena, total);
    if (e != cudaSuccess) {
      gpras_sgpr_destroy(h);
      return fail(GPRAS_E_NOMEM, "cudaMalloc", e);
    }
    char* cur = (char*)h->arena;
    h->info = (int*)cur;
    cur += 64;
    for (const Carve& c : parts) {
      *c.p = (double*)cur;
      cur += (c.count * sizeof(double) + 255) / 256 * 256;
    }
  }
  const size_t nres = 3 + d + (size_t)m * d;
  CU(cudaMallocHost(&h->h_arena, sizeof(double) * ((2 + d) + (size_t)m * d + nres) + 64));
  h->h_theta = (double*)h->h_arena;
  h->h_z = h->h_theta + (2 + d);
  h->h_result = h->h_z + (size_t)m * d;
  h->h_info = (int*)(h->h_result + nres);
  h->use_graphs = !getenv("GPRAS_B200_NO_GRAPHS");
  CU(cudaMemsetAsync(h->X, 0, sizeof(double) * h->n_pad * d, h->stream));
  CU(cudaMemsetAsync(h->Z, 0, sizeof(double) * h->m_pad * d, h->stream));
  CU(cudaMemsetAsync(h->Y, 0, sizeof(double) * h->n_pad * h->r_pad, h->stream));
  CU(cudaMemsetAsync(h->WL, 0, sizeof(double) * mm, h->stream));  // the leaves never write above the diagonal
  CU(cudaMemsetAsync(h->WB, 0, sizeof(double) * mm, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

int gpras_sgpr_destroy(gpras_sgpr* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (double* b : h->owned) cudaFree(b);
  if (h->arena) cudaFree(h->arena);
  if (h->h_arena) cudaFreeHost(h->h_arena);
  for (auto& g : h->graph)
    if (g) cudaGraphExecDestroy(g);
  h->la.destroy();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int gpras_sgpr_set_data(gpras_sgpr* h, const double* x, const double* y, int on_device) {
  if (!h || !x || !y) return fail(GPRAS_E_ARG, "null argument");
  DeviceGuard guard(h->device);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CU(cudaMemcpyAsync(h->X, x, sizeof(double) * h->n * h->d, kind, h->stream));
  CU(cudaMemcpy2DAsync(h->Y, sizeof(double) * h->r_pad, y, sizeof(double) * h->r, sizeof(double) * h->r, h->n, kind,
                       h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->has_data = true;
  h->conditioned = false;
  return 0;
}

// Enqueue every operation of one objective evaluation on h->stream (theta and z are already in the pinned staging buffers).
static int sgpr_record_eval(gpras_sgpr* h, bool want_grad) {
  cudaStream_t s = h->stream;
  const int n = h->n, D = h->d, m = h->m, n_pad = h->n_pad, m_pad = h->m_pad, ntm = h->ntm, ntn = h->ntn, rp = h->r_pad, R = h->r;
  int r;
  if ((r = sgpr_forward(h, nullptr, nullptr))) return r;
  // Binv = WB^T WB (needed for the bound's gradient and its noise derivative)
  if ((r = lauum_impl(s, h->WB, m_pad, h->Binv, m_pad, m_pad, &h->launches))) return r;
  mirror_lower_kernel<<<(unsigned)(((long)m_pad * m_pad + 255) / 256), 256, 0, s>>>(h->Binv, m_pad, m_pad);
  sgpr_scalars_kernel<<<1, 256, 0, s>>>(h->AATs, h->Binv, h->c, h->chat, rp, R, h->Y, rp, n, m, m_pad, h->scal);
  h->launches += 2;
  CU(cudaGetLastError());
  int na = 0, nb = 0;
  if (want_grad) {
    sgpr_build_r_kernel<<<(unsigned)(((long)m_pad * m_pad + 255) / 256), 256, 0, s>>>(h->Binv, h->AATs, h->chat, rp, R, m_pad,
                                                                                 h->Rm, h->RA);
    h->launches++;
    CU(cudaGetLastError());
    // RW = WL^T Rm ;  Guf1 = RW A'
    GemmDesc g1 = make_desc(h->WL, m_pad, h->Rm, m_pad, h->RW, m_pad, ntm, ntm, m_pad);
    g1.kb_mode = KB_TI;
    if ((r = launch_gemm(s, true, true, g1, 1, &h->launches))) return r;
    GemmDesc g2 = make_desc(h->RW, m_pad, h->Ap, n_pad, h->Guf1, n_pad, ntm, ntn, m_pad);
    if ((r = launch_gemm(s, false, true, g2, 1, &h->launches))) return r;
    // Guu = 1/2 WL^T (Rm - R AATs) WL
    GemmDesc g3 = make_desc(h->WL, m_pad, h->RA, m_pad, h->T1, m_pad, ntm, ntm, m_pad);
    g3.kb_mode = KB_TI;
    if ((r = launch_gemm(s, true, true, g3, 1, &h->launches))) return r;
    GemmDesc g4 = make_desc(h->T1, m_pad, h->WL, m_pad, h->Guu, m_pad, ntm, ntm, m_pad);
    g4.kb_mode = KB_TJ;
    g4.alpha = 0.5;
    if ((r = launch_gemm(s, false, true, g4, 1, &h->launches))) return r;
    // chain rule through the kernel: UF block and UU block
    if ((r = dispatch_chain(h->kid, s, false, h->Zs, m, h->Xs, n, D, h->Guf1, n_pad, h->u, rp, h->Y, rp, R, h->theta, ntm, ntn,
                                   h->partA, 1 + D, h->zpA, m_pad)))
      return r;
    if ((r = dispatch_chain(h->kid, s, true, h->Zs, m, h->Zs, m, D, h->Guu, m_pad, h->u, rp, h->Y, rp, R, h->theta, ntm, ntm,
                                  h->partB, 1 + D, h->zpB, m_pad)))
      return r;
    h->launches += 2;
    na = ntm * ntn, nb = ntm * ntm;
  }
  sgpr_finalize_kernel<<<8, 256, 0, s>>>(h->scal, h->logdetB, ntm, h->partA, na, h->partB, nb, 1 + D, h->zpA, want_grad ? ntn : 0,
                                         h->zpB, want_grad ? ntm : 0, h->theta, n, m, m_pad, D, R, h->result);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h->h_result, h->result, sizeof(double) * (3 + D + (size_t)m * D), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h->h_info, h->info, sizeof(int), cudaMemcpyDeviceToHost, s));
  return 0;
}

// One evaluation is ~40 launches of microsecond kernels at reference scale (M = 50): it is captured once per
// (handle, want_grad, jitter) into a CUDA graph and replayed (the first evaluation runs eagerly: it creates events and sets
// kernel attributes).  Works like the exact model's enqueue / fetch pair (gpras_abi.cu).
int gpras_sgpr_elbo_grad_enqueue(gpras_sgpr* h, const double* theta, const double* z, double jitter, int want_grad) {
  if (!h || !theta || !z) return fail(GPRAS_E_ARG, "null argument");
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  // the pinned staging buffers below belong to the evaluation in flight until it has been fetched
  if (h->pending) return fail(GPRAS_E_STATE, "an evaluation is already enqueued on this handle: fetch it first");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  h->conditioned = false;
  if (jitter != h->graph_jitter) {  // the jitter is a captured kernel argument: new value, new graphs
    for (auto& g : h->graph)
      if (g) cudaGraphExecDestroy(g), g = nullptr;
    h->graph_jitter = jitter;
  }
  h->jitter = jitter;
  memcpy(h->h_theta, theta, sizeof(double) * (2 + h->d));
  memcpy(h->h_z, z, sizeof(double) * (size_t)h->m * h->d);
  const int g = want_grad ? 1 : 0;
  int r;
  if (h->use_graphs && h->graph[g]) {
    h->launches = h->graph_launches[g];
    CU(cudaGraphLaunch(h->graph[g], s));
  } else if (h->use_graphs && h->eager_done[g] && !h->graph_failed) {
    h->launches = 0;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    r = sgpr_record_eval(h, want_grad != 0);
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (r == 0 && e == cudaSuccess && graph) e = cudaGraphInstantiate(&h->graph[g], graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (r != 0 || e != cudaSuccess || !h->graph[g]) {
      cudaGetLastError();
      h->graph[g] = nullptr;
      h->graph_failed = true;  // eager launches from now on (still the CUDA path, never a CPU one)
      h->launches = 0;
      if ((r = sgpr_record_eval(h, want_grad != 0))) return r;
    } else {
      h->graph_launches[g] = h->launches;
      CU(cudaGraphLaunch(h->graph[g], s));
    }
  } else {
    h->launches = 0;
    if ((r = sgpr_record_eval(h, want_grad != 0))) return r;
    h->eager_done[g] = true;
  }
  h->pending = true;
  h->pending_grad = want_grad != 0;
  return 0;
}

int gpras_sgpr_elbo_grad_fetch(gpras_sgpr* h, double* elbo, double* grad_theta, double* grad_z) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  if (!h->pending) return fail(GPRAS_E_STATE, "no evaluation enqueued");
  DeviceGuard guard(h->device);
  CU(cudaStreamSynchronize(h->stream));
  h->pending = false;
  int r;
  if ((r = sgpr_check_info(h))) return r;
  const int D = h->d;
  if (elbo) *elbo = h->h_result[0];
  if (grad_theta && h->pending_grad) memcpy(grad_theta, h->h_result + 1, sizeof(double) * (2 + D));
  if (grad_z && h->pending_grad) memcpy(grad_z, h->h_result + 3 + D, sizeof(double) * (size_t)h->m * D);
  return 0;
}

int gpras_sgpr_elbo_grad(gpras_sgpr* h, const double* theta, const double* z, double jitter, double* elbo,
                         double* grad_theta, double* grad_z) {
  int r = gpras_sgpr_elbo_grad_enqueue(h, theta, z, jitter, grad_theta != nullptr || grad_z != nullptr);
  if (r) return r;
  return gpras_sgpr_elbo_grad_fetch(h, elbo, grad_theta, grad_z);
}

int gpras_sgpr_condition(gpras_sgpr* h, const double* theta, const double* z, double jitter) {
  if (!h || !theta || !z) return fail(GPRAS_E_ARG, "null argument");
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  DeviceGuard guard(h->device);
  h->launches = 0;
  h->jitter = jitter;
  int r = sgpr_forward(h, theta, z);
  if (r) return r;
  CU(cudaMemcpyAsync(h->h_info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if ((r = sgpr_check_info(h))) return r;
  h->conditioned = true;
  return 0;
}

int gpras_sgpr_predict(gpras_sgpr* h, const double* xs, int t, double* mean, double* var) {
  if (!h || !xs || t < 0) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->conditioned) return fail(GPRAS_E_STATE, "condition() has not been called");
  DeviceGuard guard(h->device);
  cudaStream_t s = h->stream;
  const int D = h->d, m = h->m, m_pad = h->m_pad, ntm = h->ntm, rp = h->r_pad;
  int r;
  if (!h->Xt) {
    if ((r = sgpr_alloc(h, &h->Xt, (size_t)SGPR_TB * D)) || (r = sgpr_alloc(h, &h->Xts, (size_t)SGPR_TB * D)) ||
        (r = sgpr_alloc(h, &h->Kus, (size_t)m_pad * SGPR_TB)) || (r = sgpr_alloc(h, &h->tmp1, (size_t)m_pad * SGPR_TB)) ||
        (r = sgpr_alloc(h, &h->p2, (size_t)ntm * SGPR_TB)) || (r = sgpr_alloc(h, &h->q1, SGPR_TB)) ||
        (r = sgpr_alloc(h, &h->mean, (size_t)SGPR_TB * rp)) || (r = sgpr_alloc(h, &h->var, SGPR_TB)) ||
        (r = sgpr_alloc(h, &h->varm, (size_t)SGPR_TB * rp)))
      return r;
  }
  h->launches = 0;
  for (int t0 = 0; t0 < t; t0 += SGPR_TB) {
    const int tb = t - t0 < SGPR_TB ? t - t0 : SGPR_TB;
    const int tb_pad = round_up(tb, 128);
    CU(cudaMemsetAsync(h->Xt, 0, sizeof(double) * (size_t)tb_pad * D, s));
    CU(cudaMemcpyAsync(h->Xt, xs + (size_t)t0 * D, sizeof(double) * (size_t)tb * D, cudaMemcpyHostToDevice, s));
    scale_features_kernel<<<(unsigned)(((long)tb_pad * D + 255) / 256), 256, 0, s>>>(h->Xt, h->Xts, tb, tb_pad, D, h->theta);
    h->launches++;
    CU(cudaGetLastError());
    // Kus (m_pad x tb_pad), tmp1 = WL Kus, q1 = colsumsq(tmp1), p2 = colsumsq tiles of WB tmp1, mean = Kus^T u
    if ((r = dispatch_cov(h->kid, s, h->Zs, m, m_pad, h->Xts, tb, tb_pad, D, h->theta, h->Kus, SGPR_TB, 0))) return r;
    h->launches++;
    GemmDesc g1 = make_desc(h->WL, m_pad, h->Kus, SGPR_TB, h->tmp1, SGPR_TB, ntm, tb_pad / 128, m_pad);
    g1.ke_mode = KE_TI;
    if ((r = launch_gemm(s, false, true, g1, 1, &h->launches))) return r;
    colsumsq_kernel<<<(tb_pad + 127) / 128, 128, 0, s>>>(h->tmp1, m_pad, tb_pad, SGPR_TB, h->q1);
    GemmDesc g2 = make_desc(h->WB, m_pad, h->tmp1, SGPR_TB, h->p2, SGPR_TB, ntm, tb_pad / 128, m_pad);
    g2.ke_mode = KE_TI;
    g2.epilogue = EPI_COLSUMSQ;
    if ((r = launch_gemm(s, false, true, g2, 1, &h->launches))) return r;
    sgpr_predict_var_kernel<<<(tb_pad + 127) / 128, 128, 0, s>>>(h->p2, ntm, SGPR_TB, h->q1, tb_pad, h->theta, h->var);
    GemmDesc g3 = make_desc(h->Kus, SGPR_TB, h->u, rp, h->mean, rp, tb_pad / 128, rp / 32, m_pad);
    if ((r = launch_skinny(s, true, g3, h->skinny, tb_pad, &h->launches))) return r;
    h->launches += 2;
    CU(cudaGetLastError());
    if (mean)
      CU(cudaMemcpy2DAsync(mean + (size_t)t0 * h->r, sizeof(double) * h->r, h->mean, sizeof(double) * rp, sizeof(double) * h->r,
                           tb, cudaMemcpyDeviceToHost, s));
    if (var) {
      long tot = (long)tb_pad * rp;
      broadcast_var_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->var, h->varm, tb_pad, h->r, rp);
      h->launches++;
      CU(cudaMemcpy2DAsync(var + (size_t)t0 * h->r, sizeof(double) * h->r, h->varm, sizeof(double) * rp, sizeof(double) * h->r,
                           tb, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
  }
  return 0;
}

int gpras_sgpr_last_launches(gpras_sgpr* h) { return h ? h->launches : 0; }

}  // extern "C"
