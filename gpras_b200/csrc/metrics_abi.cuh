// C ABI of the streaming metrics accumulator (include/gpras_b200.h, "metrics" section).  Included by gpras_abi.cu:
// the fused entry point needs the exact-GP handle's prediction buffers.
#pragma once
#include "metrics_kernel.cuh"

namespace {
constexpr int MET_TB = 2048;        // timesteps per accumulation block (== PRED_TB)
constexpr int MET_MAX_SPLIT = 32;   // row-range partials per block
}  // namespace

struct gpras_metrics {
  int device = 0, c = 0, n_ctile = 0;
  long c_pad = 0, t_cap = 0, t_seen = 0;
  bool first = true, has_ex = false, has_ey = false;
  double v_tol = 0.0;
  int launches = 0;
  cudaStream_t stream = nullptr;
  double *state = nullptr, *rows = nullptr, *cell_part = nullptr, *row_part = nullptr, *elev_x = nullptr, *elev_y = nullptr,
         *scal = nullptr, *scal2 = nullptr, *cta_part = nullptr, *stage[3] = {nullptr, nullptr, nullptr};
};

namespace {

template <int P16>
int prepare_metrics_kernels() {
  int r;
  if ((r = opt_in_smem(metrics_stream_kernel<P16, true>, MetCfg<P16>::SMEM_BYTES))) return r;
  if ((r = opt_in_smem(metrics_plain_tma_kernel, MPT_SMEM_BYTES))) return r;
  if ((r = opt_in_smem(metrics_general_kernel<P16>, MetGenCfg<P16>::SMEM_BYTES))) return r;
  return 0;
}

// One block of t <= MET_TB rows; every pointer is a device pointer.  fused: (M, var, E, bias, rootS) describe the prediction.
int metrics_block(gpras_metrics* m, cudaStream_t s, MetricsArgs a, int t, int p16, bool fused, bool general = false) {
  if (m->t_seen + t > m->t_cap) return fail(GPRAS_E_ARG, "more timesteps than the accumulator's capacity");
  const int t_tiles = (t + MET_ROWS - 1) / MET_ROWS;
  int splits = (2 * 148 + m->n_ctile - 1) / m->n_ctile;
  if (splits > t_tiles) splits = t_tiles;
  if (splits > MET_MAX_SPLIT) splits = MET_MAX_SPLIT;
  if (splits < 1) splits = 1;
  const int per_cta = (t_tiles + splits - 1) / splits;
  splits = (t_tiles + per_cta - 1) / per_cta;
  a.t_rows = t, a.c = m->c, a.t_tiles = t_tiles, a.tiles_per_cta = per_cta, a.v_tol = m->v_tol;
  a.cell_part = m->cell_part, a.c_pad = m->c_pad, a.row_part = m->row_part, a.n_ctile = m->n_ctile, a.cta_part = m->cta_part;
  a.x_vec = a.X && (a.ldx % 2 == 0) && ((uintptr_t)a.X % 16 == 0);
  a.elev_x = m->has_ex ? m->elev_x : nullptr;
  a.elev_y = m->has_ey ? m->elev_y : nullptr;
  dim3 grid(m->n_ctile, splits);
  if (general) {  // prediction from mode-space means and PER-MODE variances
    if (p16 == 32)
      metrics_general_kernel<32><<<grid, MET_THREADS, MetGenCfg<32>::SMEM_BYTES, s>>>(a);
    else
      metrics_general_kernel<64><<<grid, MET_THREADS, MetGenCfg<64>::SMEM_BYTES, s>>>(a);
  } else if (fused) {
    if (p16 == 32)
      metrics_stream_kernel<32, true><<<grid, MET_THREADS, MetCfg<32>::SMEM_BYTES, s>>>(a);
    else
      metrics_stream_kernel<64, true><<<grid, MET_THREADS, MetCfg<64>::SMEM_BYTES, s>>>(a);
  } else {
    // resident arrays: streamed by the TMA unit whenever tensor maps can describe them (16-byte aligned rows)
    CUtensorMap xm, ym, cm;
    static const bool no_tma = getenv("GPRAS_B200_NO_TMA") != nullptr;
    const uint64_t rows = (uint64_t)t, cols = (uint64_t)m->c;
    bool ok = !no_tma && tma_map_2d_f64(&ym, a.Y, cols, rows, (uint64_t)a.ldy, 128, MET_ROWS);
    if (ok) ok = a.X ? tma_map_2d_f64(&xm, a.X, cols, rows, (uint64_t)a.ldx, 128, MET_ROWS) : (xm = ym, true);
    if (ok) ok = a.CONF ? tma_map_2d_f64(&cm, a.CONF, cols, rows, (uint64_t)a.ldconf, 128, MET_ROWS) : (cm = ym, true);
    if (ok)
      metrics_plain_tma_kernel<<<grid, MET_THREADS, MPT_SMEM_BYTES, s>>>(xm, ym, cm, a);
    else
      metrics_stream_kernel<32, false><<<grid, MET_THREADS, MET_PLAIN_SMEM_BYTES, s>>>(a);
  }
  CU(cudaGetLastError());
  metrics_fold_cells_kernel<<<(unsigned)((m->c_pad + 255) / 256), 256, 0, s>>>(m->cell_part, splits, m->c_pad, m->state, m->first ? 1 : 0);
  CU(cudaGetLastError());
  const long warps = (long)MET_ROWQ * t;
  metrics_fold_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(m->row_part, t, (long)t_tiles * MET_ROWS, m->n_ctile,
                                                                              m->rows, m->t_cap, m->t_seen);
  CU(cudaGetLastError());
  metrics_fold_cta_kernel<<<1, 32, 0, s>>>(m->cta_part, splits * m->n_ctile, m->scal2, m->first ? 1 : 0);
  CU(cudaGetLastError());
  m->launches += 4;
  m->first = false;
  m->t_seen += t;
  return 0;
}

// Staging buffer `which` (0 truth, 1 prediction, 2 confidence) for host inputs: MET_TB rows, allocated on first use, kept.
int metrics_stage(gpras_metrics* m, int which) {
  if (m->stage[which]) return 0;
  return dalloc(&m->stage[which], (size_t)MET_TB * m->c_pad);
}

}  // namespace

extern "C" {

int gpras_metrics_create(gpras_metrics** out, int device, int c, long t_capacity) {
  if (!out || c <= 0 || t_capacity <= 0) return fail(GPRAS_E_ARG, "bad shape");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  int r;
  if ((r = prepare_device())) return r;
  static std::atomic<bool> attr_done[64] = {};  // benign if two threads both set the (idempotent) attributes
  if (device < 64 && !attr_done[device]) {
    if ((r = prepare_metrics_kernels<32>()) || (r = prepare_metrics_kernels<64>())) return r;
    attr_done[device] = true;
  }
  gpras_metrics* m = new gpras_metrics();
  m->device = device, m->c = c, m->c_pad = round_up(c, 128), m->n_ctile = (int)(m->c_pad / 128), m->t_cap = t_capacity;
  CU(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  if ((r = dalloc(&m->state, (size_t)MET_CELLQ * m->c_pad)) || (r = dalloc(&m->rows, (size_t)MET_ROWQ * m->t_cap)) ||
      (r = dalloc(&m->cell_part, (size_t)MET_MAX_SPLIT * MET_CELLQ * m->c_pad)) ||
      (r = dalloc(&m->row_part, (size_t)MET_ROWQ * MET_TB * m->n_ctile)) || (r = dalloc(&m->elev_x, m->c_pad)) ||
      (r = dalloc(&m->elev_y, m->c_pad)) || (r = dalloc(&m->scal, MET_SCALARS)) || (r = dalloc(&m->scal2, MET_CTAQ)) ||
      (r = dalloc(&m->cta_part, (size_t)MET_MAX_SPLIT * m->n_ctile * MET_CTAQ))) {
    gpras_metrics_destroy(m);
    return r;
  }
  CU(cudaMemsetAsync(m->elev_x, 0, sizeof(double) * m->c_pad, m->stream));
  CU(cudaMemsetAsync(m->elev_y, 0, sizeof(double) * m->c_pad, m->stream));
  CU(cudaMemsetAsync(m->rows, 0, sizeof(double) * MET_ROWQ * m->t_cap, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  *out = m;
  return 0;
}

int gpras_metrics_destroy(gpras_metrics* m) {
  if (!m) return 0;
  DeviceGuard guard(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  double* bufs[] = {m->state, m->rows, m->cell_part, m->row_part, m->elev_x, m->elev_y, m->scal, m->scal2, m->cta_part, m->stage[0], m->stage[1], m->stage[2]};
  for (double* b : bufs)
    if (b) cudaFree(b);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return 0;
}

int gpras_metrics_set_elevations(gpras_metrics* m, const double* elev_truth, const double* elev_pred) {
  if (!m) return fail(GPRAS_E_ARG, "null handle");
  DeviceGuard guard(m->device);
  m->has_ex = elev_truth != nullptr, m->has_ey = elev_pred != nullptr;
  if (elev_truth) CU(cudaMemcpyAsync(m->elev_x, elev_truth, sizeof(double) * m->c, cudaMemcpyHostToDevice, m->stream));
  if (elev_pred) CU(cudaMemcpyAsync(m->elev_y, elev_pred, sizeof(double) * m->c, cudaMemcpyHostToDevice, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  return 0;
}

int gpras_metrics_reset(gpras_metrics* m, double v_tol) {
  if (!m) return fail(GPRAS_E_ARG, "null handle");
  m->first = true, m->t_seen = 0, m->v_tol = v_tol, m->launches = 0;
  return 0;
}

long gpras_metrics_timesteps(gpras_metrics* m) { return m ? m->t_seen : 0; }
int gpras_metrics_last_launches(gpras_metrics* m) { return m ? m->launches : 0; }

int gpras_metrics_update(gpras_metrics* m, const double* x, long ldx, const double* y, long ldy, const double* conf, long ldconf,
                         int t, int on_device) {
  if (!m || !y || t < 0) return fail(GPRAS_E_ARG, "bad argument");
  if ((x && ldx < m->c) || ldy < m->c || (conf && ldconf < m->c)) return fail(GPRAS_E_ARG, "row pitch smaller than the cell count");
  DeviceGuard guard(m->device);
  cudaStream_t s = m->stream;
  int r;
  for (int t0 = 0; t0 < t; t0 += MET_TB) {
    const int tb = t - t0 < MET_TB ? t - t0 : MET_TB;
    MetricsArgs a;
    memset(&a, 0, sizeof a);
    const double* src[3] = {x, y, conf};
    const long lds[3] = {ldx, ldy, ldconf};
    const double* dev[3] = {nullptr, nullptr, nullptr};
    long ldd[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
      if (!src[k]) continue;
      if (on_device) {
        dev[k] = src[k] + (size_t)t0 * lds[k], ldd[k] = lds[k];
      } else {
        if ((r = metrics_stage(m, k))) return r;
        CU(cudaMemcpy2DAsync(m->stage[k], sizeof(double) * m->c_pad, src[k] + (size_t)t0 * lds[k], sizeof(double) * lds[k],
                             sizeof(double) * m->c, tb, cudaMemcpyHostToDevice, s));
        dev[k] = m->stage[k], ldd[k] = m->c_pad;
      }
    }
    a.X = dev[0], a.ldx = ldd[0], a.Y = dev[1], a.ldy = ldd[1], a.CONF = dev[2], a.ldconf = ldd[2];
    if ((r = metrics_block(m, s, a, tb, 32, false))) return r;
    if (!on_device) CU(cudaStreamSynchronize(s));  // the staging buffers are reused by the next block
  }
  CU(cudaStreamSynchronize(s));
  return 0;
}

int gpras_gp_predict_metrics(gpras_gp* h, gpras_metrics* m, const double* xs, int t, int xs_on_device, const double* truth,
                             long ldx, int truth_on_device, double* mode_mean, double* mode_var) {
  if (!h || !m || !xs || t < 0) return fail(GPRAS_E_ARG, "bad argument");
  if (!h->conditioned) return fail(GPRAS_E_STATE, "condition() has not been called");
  if (!h->E1) return fail(GPRAS_E_STATE, "set_cell_map() has not been called");
  if (h->device != m->device || h->c != m->c) return fail(GPRAS_E_ARG, "metrics accumulator and model disagree on device / cell count");
  if (truth && ldx < m->c) return fail(GPRAS_E_ARG, "row pitch smaller than the cell count");
  DeviceGuard guard(h->device);
  int r;
  if ((r = ensure_predict_buffers(h))) return r;
  cudaStream_t s = h->stream, s2 = h->stream2;
  CU(cudaStreamSynchronize(m->stream));
  h->launches = 0;
  const int l0 = m->launches;
  // Same two-stream pipeline as gpras_gp_predict_cells: the metrics consumer of batch b runs while batch b+1 is predicted.
  const bool host_truth = truth && !truth_on_device;
  int batch = 0;
  for (int t0 = 0; t0 < t; t0 += PRED_TB, batch++) {
    const int tb = t - t0 < PRED_TB ? t - t0 : PRED_TB;
    const int tb_pad = round_up(tb, 128);
    const int k = batch & 1;
    if (batch >= 2) CU(cudaStreamWaitEvent(s, h->ev_cons[k], 0));
    use_predict_set(h, k);
    if ((r = stage_test_rows(h, xs, t0, tb, tb_pad, xs_on_device))) return r;
    if ((r = predict_batch(h, tb, tb_pad))) return r;
    if (mode_mean)
      CU(cudaMemcpy2DAsync(mode_mean + (size_t)t0 * h->p, sizeof(double) * h->p, h->mean, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, cudaMemcpyDefault, s));
    if (mode_var) {
      long tot = (long)tb_pad * h->p_pad;
      broadcast_var_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->var, h->varm, tb_pad, h->p, h->p_pad);
      h->launches++;
      CU(cudaGetLastError());
      CU(cudaMemcpy2DAsync(mode_var + (size_t)t0 * h->p, sizeof(double) * h->p, h->varm, sizeof(double) * h->p_pad,
                           sizeof(double) * h->p, tb, cudaMemcpyDefault, s));
    }
    MetricsArgs a;
    memset(&a, 0, sizeof a);
    a.M = h->mean, a.ldm = h->p_pad, a.var = h->var, a.E = h->E1, a.lde = h->c_pad, a.bias = h->bias, a.rootS = h->rootS;
    if (truth) {
      if (truth_on_device) {
        a.X = truth + (size_t)t0 * ldx, a.ldx = ldx;
      } else {
        if ((r = metrics_stage(m, 0))) return r;
        CU(cudaMemcpy2DAsync(m->stage[0], sizeof(double) * m->c_pad, truth + (size_t)t0 * ldx, sizeof(double) * ldx,
                             sizeof(double) * m->c, tb, cudaMemcpyHostToDevice, s));
        a.X = m->stage[0], a.ldx = m->c_pad;
      }
    }
    CU(cudaEventRecord(h->ev_pred[k], s));
    CU(cudaStreamWaitEvent(s2, h->ev_pred[k], 0));
    if ((r = metrics_block(m, s2, a, tb, h->p16, true))) return r;
    CU(cudaEventRecord(h->ev_cons[k], s2));
    if (!xs_on_device || host_truth) {  // pageable host buffers / the single truth staging buffer: keep batches ordered
      CU(cudaStreamSynchronize(s));
      if (host_truth) CU(cudaStreamSynchronize(s2));
    }
  }
  CU(cudaStreamSynchronize(s));
  CU(cudaStreamSynchronize(s2));
  h->launches += m->launches - l0;
  return 0;
}

// Accumulate t <= 2048 timesteps predicted in MODE space with one variance per mode (the reference's per-column models):
// y = max(M E + bias - elev, 0), conf = sqrt(V E^2), consumed tile by tile against the truth -- the (t x cells) prediction is
// never written.  Every pointer is a DEVICE pointer: M, V (round_up(t, 32) x ldm, zero padded, ldm = p16 = 32 or 64),
// E (p16 x lde folded EOF map, zero on dry cells), bias (lde), truth (t x ldx) or NULL.  Runs on the accumulator's stream
// and returns when the block has been consumed.  Called by gpras_pre_reverse_metrics, which owns the map.
int gpras_metrics_update_modes(gpras_metrics* m, const double* M, const double* V, long ldm, int p16, const double* E, long lde,
                               const double* bias, const double* truth, long ldx, int t) {
  if (!m || !M || !V || !E || !bias || t <= 0 || t > MET_TB) return fail(GPRAS_E_ARG, "bad argument");
  if ((p16 != 32 && p16 != 64) || ldm != p16 || lde < m->c_pad) return fail(GPRAS_E_ARG, "mode tiles must be padded to 32 or 64 columns");
  if (truth && ldx < m->c) return fail(GPRAS_E_ARG, "row pitch smaller than the cell count");
  DeviceGuard guard(m->device);
  MetricsArgs a;
  memset(&a, 0, sizeof a);
  a.M = M, a.Vm = V, a.ldm = ldm, a.E = E, a.lde = lde, a.bias = bias;
  a.X = truth, a.ldx = ldx;
  int r = metrics_block(m, m->stream, a, t, p16, true, true);
  if (r) return r;
  CU(cudaStreamSynchronize(m->stream));
  return 0;
}

int gpras_metrics_finalize(gpras_metrics* m, double depth_threshold, double* scalars, double* cells, double* rows) {
  if (!m || !scalars) return fail(GPRAS_E_ARG, "null argument");
  if (m->t_seen <= 0) return fail(GPRAS_E_STATE, "no timesteps accumulated");
  DeviceGuard guard(m->device);
  cudaStream_t s = m->stream;
  metrics_finalize_kernel<<<1, 1024, 0, s>>>(m->state, m->c_pad, m->c, m->scal2, depth_threshold, m->scal);
  CU(cudaGetLastError());
  m->launches++;
  CU(cudaMemcpyAsync(scalars, m->scal, sizeof(double) * MET_SCALARS, cudaMemcpyDeviceToHost, s));
  if (cells)
    CU(cudaMemcpy2DAsync(cells, sizeof(double) * m->c, m->state, sizeof(double) * m->c_pad, sizeof(double) * m->c, MET_CELLQ,
                         cudaMemcpyDeviceToHost, s));
  if (rows)
    CU(cudaMemcpy2DAsync(rows, sizeof(double) * m->t_seen, m->rows, sizeof(double) * m->t_cap, sizeof(double) * m->t_seen, MET_ROWQ,
                         cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

int gpras_metrics_fidelity(const double* x, long ldx, const double* y, long ldy, int t, int c, int t_tol, double v_tol, int device,
                           double* matching) {
  if (!x || !y || !matching || t <= 0 || c <= 0 || t_tol < 0) return fail(GPRAS_E_ARG, "bad argument");
  if (t > 65535) return fail(GPRAS_E_ARG, "more than 65 535 timesteps per call are not supported (one grid row per timestep)");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  const int bx = (c + 255) / 256;
  double *dx = nullptr, *dy = nullptr, *part = nullptr;
  int r;
  if ((r = dalloc(&dx, (size_t)t * c)) || (r = dalloc(&dy, (size_t)t * c)) || (r = dalloc(&part, (size_t)t * bx))) {
    cudaFree(dx), cudaFree(dy), cudaFree(part);
    return r;
  }
  cudaError_t e = cudaMemcpy2D(dx, sizeof(double) * c, x, sizeof(double) * ldx, sizeof(double) * c, t, cudaMemcpyDefault);
  if (e == cudaSuccess) e = cudaMemcpy2D(dy, sizeof(double) * c, y, sizeof(double) * ldy, sizeof(double) * c, t, cudaMemcpyDefault);
  std::vector<double> hp((size_t)t * bx);
  if (e == cudaSuccess) {
    metrics_fidelity_kernel<<<dim3(bx, t), 256>>>(dx, c, dy, c, t, c, t_tol, v_tol, part);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(hp.data(), part, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost);
  cudaFree(dx), cudaFree(dy), cudaFree(part);
  if (e != cudaSuccess) return fail(GPRAS_E_CUDA, "fidelity kernel", e);
  double sum = 0.0;
  for (double v : hp) sum += v;  // integer-valued partial counts: exact in any order
  *matching = sum;
  return 0;
}

}  // extern "C"
