// C ABI of gpras_b200, exact-GP part (see include/gpras_b200.h): host-side orchestration of the sm_100a kernels.
// No CPU compute path exists in this file: every entry point either launches CUDA work or fails.
#include <cmath>
#include <cstdlib>

#include "host_common.cuh"
#include "cells_kernel.cuh"

namespace {

template <int KID>
int launch_grad(cudaStream_t s, const double* Xs, int n, int n_pad, int D, const double* Wt, long ldw, double* part,
                int ncols) {
  const int nt = n_pad / CT;
  const int tiles = nt * (nt + 1) / 2;
  const int smem = 2 * D * CT_LD * (int)sizeof(double);
  static std::atomic<bool> attr_done[64] = {};  // benign if two threads both set the (idempotent) attributes
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    int r;
    if ((r = opt_in_smem(grad_kernel<KID, 16>, 200 * 1024)) || (r = opt_in_smem(grad_kernel<KID, 32>, 200 * 1024)) ||
        (r = opt_in_smem(grad_kernel<KID, 64>, 200 * 1024)))
      return r;
    attr_done[dev] = true;
  }
  if (D <= 16)
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
  void *arena = nullptr, *h_arena = nullptr;  // single device / pinned allocations the training buffers are carved from
  // prediction state
  double *Xt = nullptr, *Xts = nullptr, *Ks = nullptr, *mean = nullptr, *vpart = nullptr, *var = nullptr,
         *varm = nullptr;
  // cell map
  int c = 0, c_pad = 0, p16 = 0;
  double *E1 = nullptr, *E2 = nullptr, *rootS = nullptr, *bias = nullptr, *zbias = nullptr, *ring_m = nullptr, *ring_v = nullptr;
  cudaEvent_t ev[8] = {};
  // Host wait of an evaluation: cudaStreamSynchronize (spins: lowest latency, the default) or a blocking event (the thread
  // sleeps: for several host threads per GPU, e.g. restart lanes, where spinning threads can outnumber the cores; it costs
  // ~1 % at N = 8192 with one host thread, measured).  gpras_gp_set_blocking_wait selects it per handle.
  cudaEvent_t ev_done = nullptr;
  bool blocking_wait = false;
  // prediction pipeline: batch b's consumer (modes -> cells, or the fused metrics) runs on stream2 while batch b+1's
  // predictor runs on `stream`; the small mode-space buffers are double-buffered, `mean / var / varm` point at the current set
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_pred[2] = {}, ev_cons[2] = {};
  double *mean_b[2] = {nullptr, nullptr}, *var_b[2] = {nullptr, nullptr}, *varm_b[2] = {nullptr, nullptr};
  double stage_ms[7] = {};
  LookAhead la;
  // CUDA-graph replay of the evaluation (index: want_grad)
  bool use_graphs = true, graph_failed = false, eager_done[2] = {false, false};
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  int graph_launches[2] = {0, 0};
};

namespace {

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

// Enqueue every operation of one evaluation on h->stream (theta is already in the pinned staging buffer).
int record_eval(gpras_gp* h, int want_grad) {
  cudaStream_t s = h->stream;
  const int n = h->n, n_pad = h->n_pad, D = h->d, P = h->p;
  int r;
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
  return 0;
}

// One evaluation = ~20 + 5 N/128 launches over two streams.  theta lives in device memory, so the whole sequence
// is captured once per (handle, want_grad) into a CUDA graph and replayed: one launch per evaluation instead of
// hundreds (the first evaluation runs eagerly: it creates streams / events and sets kernel attributes).
int enqueue_eval(gpras_gp* h, const double* theta, int want_grad) {
  if (!h->has_data) return fail(GPRAS_E_STATE, "set_data has not been called");
  // the pinned theta staging buffer belongs to the evaluation in flight until it has been fetched
  if (h->pending) return fail(GPRAS_E_STATE, "an evaluation is already enqueued on this handle: fetch it first");
  cudaStream_t s = h->stream;
  h->conditioned = false;
  memcpy(h->h_theta, theta, sizeof(double) * (2 + h->d));
  const int g = want_grad ? 1 : 0;
  int r;
  if (h->use_graphs && !h->stage_timing && h->graph[g]) {
    h->launches = h->graph_launches[g];
    CU(cudaGraphLaunch(h->graph[g], s));
  } else if (h->use_graphs && !h->stage_timing && h->eager_done[g] && !h->graph_failed) {
    h->launches = 0;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    r = record_eval(h, want_grad);
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (r != 0 || e != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      h->graph_failed = true;  // fall back to eager launches (still the CUDA path, never a CPU one)
      h->launches = 0;
      if ((r = record_eval(h, want_grad))) return r;
    } else {
      e = cudaGraphInstantiate(&h->graph[g], graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        cudaGetLastError();
        h->graph[g] = nullptr;
        h->graph_failed = true;
        h->launches = 0;
        if ((r = record_eval(h, want_grad))) return r;
      } else {
        h->graph_launches[g] = h->launches;
        CU(cudaGraphLaunch(h->graph[g], s));
      }
    }
  } else {
    h->launches = 0;
    if ((r = record_eval(h, want_grad))) return r;
    h->eager_done[g] = true;
  }
  if (h->blocking_wait) CU(cudaEventRecord(h->ev_done, s));
  h->pending = true;
  h->pending_grad = want_grad != 0;
  return 0;
}

int fetch_eval(gpras_gp* h, double* lml, double* grad) {
  if (!h->pending) return fail(GPRAS_E_STATE, "no evaluation enqueued");
  if (h->blocking_wait)
    CU(cudaEventSynchronize(h->ev_done));
  else
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
  // graph replay pays off while the evaluation is launch/latency bound (measured: +6..10% for N <= 4096, -3% at
  // N = 8192 with two evaluations in flight, where eager launches interleave better across handles)
  h->use_graphs = (h->n_pad <= 4096 || getenv("GPRAS_B200_GRAPHS_ALL")) && !getenv("GPRAS_B200_NO_GRAPHS");
  const size_t nn = (size_t)h->n_pad * h->n_pad, np = (size_t)h->n_pad * h->p_pad;
  const int ntile = h->nt * (h->nt + 1) / 2;
  // One device allocation and one pinned allocation per handle, carved below: a handle costs ~1 ms to create instead of
  // ~4 ms (two dozen cudaMalloc / cudaMallocHost calls), which matters when a pool of them serves batched restarts.
  struct Carve {
    double** p;
    size_t count;
  };
  const Carve parts[] = {
      {&h->X, (size_t)h->n_pad * d}, {&h->Xs, (size_t)h->n_pad * d}, {&h->Y, np}, {&h->K, nn}, {&h->W, nn}, {&h->Kinv, nn},
      {&h->U, np}, {&h->alpha, np}, {&h->theta, (size_t)2 + d}, {&h->logdet, (size_t)h->nt},
      {&h->gpart, (size_t)ntile * (2 + d)}, {&h->gsum, (size_t)2 + d}, {&h->usq, (size_t)USQ_PARTS}, {&h->result, (size_t)3 + d},
      {&h->skinny, (size_t)SKINNY_MAX_SLABS * (h->n_pad > PRED_TB ? h->n_pad : PRED_TB) * h->p_pad}};
  size_t total = 64;  // the info word lives in the first 64 bytes
  for (const Carve& c : parts) total += (c.count * sizeof(double) + 255) / 256 * 256;
  {
    cudaError_t e = cudaMalloc((void**)&h->arena, total);
    if (e != cudaSuccess) {
      gpras_gp_destroy(h);
      return fail(GPRAS_E_NOMEM, "cudaMalloc", e);
    }
  }
  {
    char* cur = (char*)h->arena;
    h->info = (int*)cur;
    cur += 64;
    for (const Carve& c : parts) {
      *c.p = (double*)cur;
      cur += (c.count * sizeof(double) + 255) / 256 * 256;
    }
  }
  CU(cudaMallocHost((void**)&h->h_arena, sizeof(double) * (2 + d) + sizeof(double) * (3 + d) + 64));
  h->h_theta = (double*)h->h_arena;
  h->h_result = h->h_theta + (2 + d);
  h->h_info = (int*)(h->h_result + (3 + d));
  // stream-ordered: the handle's stream is non-blocking, so legacy-stream memsets would race with set_data
  CU(cudaMemsetAsync(h->X, 0, sizeof(double) * h->n_pad * d, h->stream));
  CU(cudaMemsetAsync(h->Y, 0, sizeof(double) * np, h->stream));
  CU(cudaMemsetAsync(h->gsum, 0, sizeof(double) * (2 + d), h->stream));
  CU(cudaMemsetAsync(h->W, 0, sizeof(double) * nn, h->stream));  // the leaves never write above the diagonal
  CU(cudaStreamSynchronize(h->stream));
  for (auto& e : h->ev) CU(cudaEventCreate(&e));
  h->blocking_wait = getenv("GPRAS_B200_BLOCKING_WAIT") != nullptr;  // see gpras_gp_set_blocking_wait
  CU(cudaEventCreateWithFlags(&h->ev_done, cudaEventBlockingSync | cudaEventDisableTiming));
  *out = h;
  return 0;
}

int gpras_gp_destroy(gpras_gp* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  double* bufs[] = {h->Xt, h->Xts, h->Ks, h->mean_b[0], h->mean_b[1], h->vpart, h->var_b[0], h->var_b[1], h->varm_b[0], h->varm_b[1],
                    h->E1, h->E2, h->bias, h->zbias, h->ring_m, h->ring_v, h->rootS};
  for (double* b : bufs)
    if (b) cudaFree(b);
  for (auto& e : h->ev_pred)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_cons)
    if (e) cudaEventDestroy(e);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->arena) cudaFree(h->arena);
  if (h->h_arena) cudaFreeHost(h->h_arena);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  for (auto& g : h->graph)
    if (g) cudaGraphExecDestroy(g);
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
  if (h->pending) return fail(GPRAS_E_STATE, "an evaluation is already enqueued on this handle: fetch it first");
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

static void use_predict_set(gpras_gp* h, int k) { h->mean = h->mean_b[k], h->var = h->var_b[k], h->varm = h->varm_b[k]; }

static int ensure_predict_buffers(gpras_gp* h) {
  if (h->Xt) return 0;
  int r;
  if ((r = dalloc(&h->Xt, (size_t)PRED_TB * h->d)) || (r = dalloc(&h->Xts, (size_t)PRED_TB * h->d)) ||
      (r = dalloc(&h->Ks, (size_t)PRED_TB * h->n_pad)) || (r = dalloc(&h->vpart, (size_t)h->nt * PRED_TB)))
    return r;
  for (int k = 0; k < 2; k++) {
    if ((r = dalloc(&h->mean_b[k], (size_t)PRED_TB * h->p_pad)) || (r = dalloc(&h->var_b[k], PRED_TB)) ||
        (r = dalloc(&h->varm_b[k], (size_t)PRED_TB * h->p_pad)))
      return r;
    CU(cudaEventCreateWithFlags(&h->ev_pred[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_cons[k], cudaEventDisableTiming));
  }
  CU(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  use_predict_set(h, 0);
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
  if (h->p > 64) return fail(GPRAS_E_ARG, "more than 64 modes are not supported by the cell expansion");
  DeviceGuard guard(h->device);
  double* olds[] = {h->E1, h->E2, h->bias, h->zbias, h->ring_m, h->ring_v, h->rootS};
  for (double* b : olds)
    if (b) cudaFree(b);
  h->E1 = h->E2 = h->bias = h->zbias = h->ring_m = h->ring_v = h->rootS = nullptr;
  h->c = c, h->c_pad = round_up(c, 128), h->p16 = h->p <= 32 ? 32 : 64;
  const size_t ne = (size_t)h->p16 * h->c_pad;
  int r;
  // E2 holds S[c] = sum_p E[p][c]^2: every mode shares theta, so cell variance = var[t] * S[c]
  if ((r = dalloc(&h->E1, ne)) || (r = dalloc(&h->E2, h->c_pad)) || (r = dalloc(&h->bias, h->c_pad)) ||
      (r = dalloc(&h->rootS, h->c_pad)))
    return r;
  std::vector<double> e1(ne, 0.0), sq(h->c_pad, 0.0), b(h->c_pad, 0.0), rt(h->c_pad, 0.0);
  for (int pp = 0; pp < h->p; pp++)
    for (int cc = 0; cc < c; cc++) {
      const double v = e_mean[(size_t)pp * c + cc];
      e1[(size_t)pp * h->c_pad + cc] = v;
      sq[cc] += v * v;
    }
  memcpy(b.data(), bias, sizeof(double) * c);
  for (int cc = 0; cc < c; cc++) rt[cc] = sqrt(sq[cc]);
  CU(cudaMemcpyAsync(h->rootS, rt.data(), sizeof(double) * h->c_pad, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->E1, e1.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->E2, sq.data(), sizeof(double) * h->c_pad, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->bias, b.data(), sizeof(double) * h->c_pad, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  static std::atomic<bool> attr_done[64] = {};  // benign if two threads both set the (idempotent) attributes
  if (h->device < 64 && !attr_done[h->device]) {
    if ((r = opt_in_smem(cells_kernel<32>, CellsCfg<32>::SMEM_BYTES)) || (r = opt_in_smem(cells_kernel<64>, CellsCfg<64>::SMEM_BYTES)))
      return r;
    attr_done[h->device] = true;
  }
  return 0;
}

long gpras_gp_cell_pitch(gpras_gp* h) { return h ? h->c_pad : 0; }

int gpras_gp_predict_cells(gpras_gp* h, const double* xs, int t, int xs_on_device, double* mode_mean, double* mode_var,
                           double* cell_mean, double* cell_var, long ldc) {
  if (!h || !xs || t < 0) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->conditioned) return fail(GPRAS_E_STATE, "condition() has not been called");
  if (!h->E1) return fail(GPRAS_E_STATE, "set_cell_map() has not been called");
  if ((cell_mean != nullptr) != (cell_var != nullptr)) return fail(GPRAS_E_ARG, "pass both cell_mean and cell_var, or neither");
  if (cell_mean && ldc < h->c_pad) return fail(GPRAS_E_ARG, "ldc smaller than gpras_gp_cell_pitch()");
  DeviceGuard guard(h->device);
  int r;
  if ((r = ensure_predict_buffers(h))) return r;
  if (!h->ring_m) {
    if ((r = dalloc(&h->ring_m, (size_t)CELL_TB * h->c_pad)) || (r = dalloc(&h->ring_v, (size_t)CELL_TB * h->c_pad)))
      return r;
  }
  cudaStream_t s = h->stream, s2 = h->stream2;
  h->launches = 0;
  // Software pipeline over batches: the predictor of batch b+1 (FP64 tensor pipe: the N^2 T variance product) overlaps the
  // modes -> cells expansion of batch b (HBM-write bound) on a second stream.
  int batch = 0;
  for (int t0 = 0; t0 < t; t0 += PRED_TB, batch++) {
    const int tb = t - t0 < PRED_TB ? t - t0 : PRED_TB;
    const int tb_pad = round_up(tb, 128);
    const int k = batch & 1;
    if (batch >= 2) CU(cudaStreamWaitEvent(s, h->ev_cons[k], 0));  // the expansion that read this buffer set is done
    use_predict_set(h, k);
    if ((r = stage_test_rows(h, xs, t0, tb, tb_pad, xs_on_device))) return r;
    if ((r = predict_batch(h, tb, tb_pad))) return r;
    long tot = (long)tb_pad * h->p_pad;
    broadcast_var_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->var, h->varm, tb_pad, h->p, h->p_pad);
    h->launches++;
    CU(cudaGetLastError());
    if (mode_mean)
      CU(cudaMemcpy2DAsync(mode_mean + (size_t)t0 * h->p, sizeof(double) * h->p, h->mean, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, cudaMemcpyDefault, s));
    if (mode_var)
      CU(cudaMemcpy2DAsync(mode_var + (size_t)t0 * h->p, sizeof(double) * h->p, h->varm, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, cudaMemcpyDefault, s));
    CU(cudaEventRecord(h->ev_pred[k], s));
    CU(cudaStreamWaitEvent(s2, h->ev_pred[k], 0));
    // modes -> cells: one streaming kernel per batch (column tile per CTA, row tiles streamed)
    {
      const bool keep = cell_mean != nullptr && cell_var != nullptr;
      double* om = keep ? cell_mean + (size_t)t0 * ldc : h->ring_m;
      double* ov = keep ? cell_var + (size_t)t0 * ldc : h->ring_v;
      const long ldo = keep ? ldc : h->c_pad;
      const int ring_rows = keep ? (1 << 30) : CELL_TB;
      const int t_tiles = tb_pad / CELLS_ROWS;
      const int per_cta = (t_tiles + 1) / 2;
      dim3 grid(h->c_pad / 128, (t_tiles + per_cta - 1) / per_cta);
      if (h->p16 == 32)
        cells_kernel<32><<<grid, CELLS_THREADS, CellsCfg<32>::SMEM_BYTES, s2>>>(h->mean, h->p_pad, h->var, h->E1, h->c_pad, h->bias,
                                                                              h->E2, om, ov, ldo, t_tiles, per_cta, ring_rows);
      else
        cells_kernel<64><<<grid, CELLS_THREADS, CellsCfg<64>::SMEM_BYTES, s2>>>(h->mean, h->p_pad, h->var, h->E1, h->c_pad, h->bias,
                                                                              h->E2, om, ov, ldo, t_tiles, per_cta, ring_rows);
      h->launches++;
      CU(cudaGetLastError());
    }
    CU(cudaEventRecord(h->ev_cons[k], s2));
    if (!xs_on_device) CU(cudaStreamSynchronize(s));
  }
  CU(cudaStreamSynchronize(s));
  CU(cudaStreamSynchronize(s2));
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

int gpras_gp_set_blocking_wait(gpras_gp* h, int enabled) {
  if (!h) return fail(GPRAS_E_ARG, "null handle");
  if (h->pending) return fail(GPRAS_E_STATE, "an evaluation is in flight on this handle");
  h->blocking_wait = enabled != 0;
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

#include "metrics_abi.cuh"
