/* gpras_b200 -- C ABI of the B200-native Gaussian-process-regression hot path.
 *
 * Drop-in boundary for the numerical work that fema-ffrd/gpras delegates to GPflow/TensorFlow from
 * gpras/gpr.py (the reference has no FFI of its own; its boundary is the Python class
 * gpras.gpr.GPRAS, gpr.py:217-384).  Each entry point below names the reference call it replaces.
 * The Python mirror gpras_b200/gpr.py binds these with ctypes.
 *
 * Conventions
 *   - FP64 everywhere (gpr.py:18 gpflow.config.set_default_float(tf.float64)); matrices row-major.
 *   - theta = [variance, noise, l_0 .. l_{D-1}] constrained values (kernel variance, Gaussian
 *     likelihood variance, ARD lengthscales; an isotropic model repeats one value D times).
 *   - Gradients are d/d log(theta_j); the softplus / prior chain rule is host logic.
 *   - Every function returns int: 0 ok; > 0 LAPACK-style info (index+1 of the first non-positive
 *     pivot); < 0 a GPRAS_E_* code, message via gpras_last_error().  Nothing throws across the ABI.
 *   - There is no CPU fallback: without a CUDA device every compute entry returns GPRAS_E_CUDA.
 *   - A handle owns one device, one stream and its workspace; it is not thread-safe.
 */
#ifndef GPRAS_B200_H
#define GPRAS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPRAS_B200_ABI_VERSION 4

/* kernel ids == KERNEL_FACTORY keys that are constructible in the reference (gpr.py:21-29, 298) */
#define GPRAS_KERNEL_RBF 0
#define GPRAS_KERNEL_MATERN12 1
#define GPRAS_KERNEL_MATERN32 2
#define GPRAS_KERNEL_MATERN52 3
#define GPRAS_KERNEL_EXPONENTIAL 4

#define GPRAS_E_ARG (-1)
#define GPRAS_E_CUDA (-2)
#define GPRAS_E_STATE (-3)
#define GPRAS_E_NOMEM (-4)

typedef struct gpras_gp gpras_gp; /* exact-GP model state on one GPU */

int gpras_abi_version(void);
const char* gpras_last_error(void);
/* number of CUDA devices visible (0 if none / no driver) */
int gpras_device_count(void);

/* ---- lifecycle --------------------------------------------------------------------------- */
/* Replaces SGPR(data=(x, y), kernel=...) construction, gpr.py:293-308, for the shared-theta exact model. */
int gpras_gp_create(gpras_gp** out, int device, int kernel_id, int n, int d, int p);
int gpras_gp_destroy(gpras_gp* h);
/* Run on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream. */
int gpras_gp_set_stream(gpras_gp* h, void* cuda_stream);
/* Upload / bind training data (gpr.py:265-266).  on_device != 0: pointers are device pointers. */
int gpras_gp_set_data(gpras_gp* h, const double* x, const double* y, int on_device);

/* ---- objective: replaces model.training_loss() + GradientTape (gpr.py:153-155,199) ------- */
/* Synchronous.  theta is a host array of 2 + d doubles.  grad (2 + d, d/dlog theta) may be NULL. */
int gpras_gp_lml_grad(gpras_gp* h, const double* theta, double* lml, double* grad);
/* Same, end to end from host buffers: uploads x (n*d) and y (n*p) first. */
int gpras_gp_lml_grad_host(gpras_gp* h, const double* x, const double* y, const double* theta, double* lml,
                           double* grad);
/* Asynchronous pair for throughput runs: enqueue on the handle's stream, then fetch (synchronises). */
int gpras_gp_lml_grad_enqueue(gpras_gp* h, const double* theta, int want_grad);
int gpras_gp_lml_grad_fetch(gpras_gp* h, double* lml, double* grad);

/* ---- prediction: replaces model.predict_y(x) (gpr.py:337) ------------------------------- */
/* Factorise at theta and form alpha; must precede predict. */
int gpras_gp_condition(gpras_gp* h, const double* theta);
/* mean (t*p) and var (t*p), noise included; xs is (t, d).  Host or device buffers (on_device). */
int gpras_gp_predict(gpras_gp* h, const double* xs, int t, double* mean, double* var, int on_device);
/* Bind the modes->cells map (PreProcessor.reverse_transform, preprocess.py:1052-1094):
 * e_mean[p][c] = x_std[p]*eofs[p][c]/weights[c] on wet cells (0 on dry), bias[c] = offset / dry fill.
 * Host arrays; c = number of cells. */
int gpras_gp_set_cell_map(gpras_gp* h, const double* e_mean, const double* bias, int c);
/* Predict and expand to cells on the device: cell_mean, cell_var are DEVICE buffers (t x ldc doubles, ldc >=
 * padded cells, see gpras_gp_cell_pitch) or NULL to run without keeping the cell-space output (the tiles are
 * written to an internal ring buffer).  mode_mean / mode_var (t*p, host OR device memory) may be NULL. */
int gpras_gp_predict_cells(gpras_gp* h, const double* xs, int t, int xs_on_device, double* mode_mean,
                           double* mode_var, double* cell_mean, double* cell_var, long ldc);
long gpras_gp_cell_pitch(gpras_gp* h);

/* How the host waits for an evaluation: 0 (default) cudaStreamSynchronize, which spins; 1 a blocking event, which lets the
 * thread sleep (several host threads per GPU, e.g. the restart lanes of gpras_b200/parallel.py). */
int gpras_gp_set_blocking_wait(gpras_gp* h, int enabled);

/* ---- introspection ---------------------------------------------------------------------- */
/* Copy internal matrices to host (n x n, row-major, lower triangle meaningful): which = 0 K~ as built,
 * 1 L, 2 W = L^-1, 3 K~^-1;  4 alpha (n x p). For tests. */
int gpras_gp_get_matrix(gpras_gp* h, int which, double* out);
/* Count of kernel launches issued by the last lml_grad / condition / predict call. */
int gpras_gp_last_launches(gpras_gp* h);
/* CUDA-event milliseconds of the stages of the last lml_grad call: [cov, potrf, trtri, lauum, alpha, grad, total]. */
int gpras_gp_last_stage_ms(gpras_gp* h, double* ms7);
int gpras_gp_set_stage_timing(gpras_gp* h, int enabled);

/* ---- sparse (inducing-point) model: replaces gpflow.models.SGPR as built at gpr.py:293-308 ---------- */
typedef struct gpras_sgpr gpras_sgpr;
/* n rows, d features, m inducing points, r output columns sharing the model (1 in the reference). */
int gpras_sgpr_create(gpras_sgpr** out, int device, int kernel_id, int n, int d, int m, int r);
int gpras_sgpr_destroy(gpras_sgpr* h);
int gpras_sgpr_set_data(gpras_sgpr* h, const double* x, const double* y, int on_device);
/* Collapsed bound (ELBO) of SGPR.training_loss (gpr.py:154,199; priors are host logic) and its gradient:
 * theta (2 + d) and z (m x d) are host arrays; grad_theta (2 + d) is d/dlog theta, grad_z (m x d) is d/dz.
 * Either gradient pointer may be NULL; both NULL skips the backward pass (differential evolution, gpr.py:62). */
int gpras_sgpr_elbo_grad(gpras_sgpr* h, const double* theta, const double* z, double jitter, double* elbo,
                         double* grad_theta, double* grad_z);
/* Asynchronous pair (the evaluation is replayed from a CUDA graph): enqueue on the handle's stream, then fetch (synchronises).
 * Several handles -- one per per-column model -- keep their evaluations in flight together. */
int gpras_sgpr_elbo_grad_enqueue(gpras_sgpr* h, const double* theta, const double* z, double jitter, int want_grad);
int gpras_sgpr_elbo_grad_fetch(gpras_sgpr* h, double* elbo, double* grad_theta, double* grad_z);
/* predict_y (gpr.py:337): condition at (theta, z), then mean (t x r) and variance (t x r, noise included). */
int gpras_sgpr_condition(gpras_sgpr* h, const double* theta, const double* z, double jitter);
int gpras_sgpr_predict(gpras_sgpr* h, const double* xs, int t, double* mean, double* var);
int gpras_sgpr_last_launches(gpras_sgpr* h);

/* ---- the per-column sparse models of ONE fit, batched and trained on the device ----------------------------------
 * Replaces the loop `for model in self.models: opt(model)` (gpr.py:273-274) for the Adam-based recipes (gpr.py:112-127,
 * 147-173): p independent SGPR models (one target column each, r = 1) over the same inputs, m <= 128 inducing points.
 * Every kernel of the evaluation runs once for all models (model index in blockIdx.y: bitwise the single-model results of
 * gpras_sgpr_elbo_grad), and one Adam step of all models -- transforms, LogNormal(0, 1) priors, chain rule, Keras' update,
 * the reference's early-stopping rule -- is one CUDA graph with no host round trip. */
typedef struct gpras_sgpr_batch gpras_sgpr_batch;
int gpras_sgpr_batch_create(gpras_sgpr_batch** out, int device, int kernel_id, int n, int d, int m, int p);
int gpras_sgpr_batch_destroy(gpras_sgpr_batch* h);
/* x: n x d (shared), y: n x p row-major (column b = model b's targets); host arrays. */
int gpras_sgpr_batch_set_data(gpras_sgpr_batch* h, const double* x, const double* y);
/* Bound + gradient of every model: theta p x (2 + d), z p x m x d -> elbo p, grad_theta p x (2 + d) (d/dlog theta),
 * grad_z p x m x d, info p (0, or the failing pivot of a model whose Kuu / B lost positive definiteness). */
int gpras_sgpr_batch_elbo_grad(gpras_sgpr_batch* h, const double* theta, const double* z, double jitter, double* elbo,
                               double* grad_theta, double* grad_z, int* info);
/* One Adam stage (_optimize_adam, gpr.py:147-173) of all models, device resident.
 * u: p x (2 + n_ls + m d) in/out, the UNCONSTRAINED variables [variance, noise, lengthscale(s), Z] of each model;
 * n_ls: 1 (one lengthscale, the reference) or d; train_hypers / train_z: the gpflow.set_trainable stage;
 * transform: 0 = GPflow softplus (noise: + noise_floor), 1 = exp; priors: LogNormal(0, 1) on the hyperparameters;
 * losses: max_iter x p out (NaN where a model had stopped; may be NULL); iters: p out, steps taken; info: p out. */
int gpras_sgpr_batch_adam(gpras_sgpr_batch* h, double* u, int n_ls, int train_hypers, int train_z, int max_iter, double lr,
                          double jitter, int transform, int priors, double noise_floor, double* losses, int* iters, int* info);
/* The same device-resident loop with the update rule as a parameter: 0 = Adam (as above), 1 = Keras Adadelta
 * (_optimize_adadelta / _optimize_tf, gpr.py:176-192: exactly max_iter steps, no early stopping). */
int gpras_sgpr_batch_train(gpras_sgpr_batch* h, double* u, int n_ls, int train_hypers, int train_z, int max_iter, double lr,
                           double jitter, int transform, int priors, double noise_floor, int rule, double* losses, int* iters,
                           int* info);
/* predict_y (gpr.py:337) of all models in one pass per tile of test inputs (m <= 64, d <= 32; GPRAS_E_ARG otherwise):
 * condition at (theta p x (2 + d), z p x m x d) -> info p; then mean, var (t x p, noise included) for xs (t x d). */
int gpras_sgpr_batch_condition(gpras_sgpr_batch* h, const double* theta, const double* z, double jitter, int* info);
int gpras_sgpr_batch_predict(gpras_sgpr_batch* h, const double* xs, int t, double* mean, double* var);
int gpras_sgpr_batch_last_launches(gpras_sgpr_batch* h);

/* ---- streaming accuracy metrics: replaces gpras/metrics.py:85-324 (called from production/analysis/pipeline.py:281-286) -- */
/* One accumulator = one event of up to t_capacity timesteps over c cells.  It keeps, per cell, sum(x-y), sum((x-y)^2),
 * sum(conf), max_t x, max_t y and, per timestep, sum_c(x-y), sum_c((x-y)^2), sum_c(conf), sum_c|x-y|,
 * count_c(|x-y| <= v_tol): every function of metrics.py (with t_tol == 0) is a closed form of these. */
typedef struct gpras_metrics gpras_metrics;
int gpras_metrics_create(gpras_metrics** out, int device, int c, long t_capacity);
int gpras_metrics_destroy(gpras_metrics* m);
/* Depth conversion on the fly (PreProcessor.wse_2_depth, preprocess.py:1041-1045, as applied at pipeline.py:265-274):
 * truth <- max(truth - elev_truth, 0), prediction <- max(prediction - elev_pred, 0).  Host arrays of c doubles; NULL = none. */
int gpras_metrics_set_elevations(gpras_metrics* m, const double* elev_truth, const double* elev_pred);
/* Start a new event; v_tol is fi_aoi_toi's value tolerance (metrics.py:203). */
int gpras_metrics_reset(gpras_metrics* m, double v_tol);
/* Accumulate t timesteps held as (t x ld) arrays: truth x (may be NULL = 0), prediction y, confidence conf (may be NULL).
 * on_device != 0: all three are device pointers. */
int gpras_metrics_update(gpras_metrics* m, const double* x, long ldx, const double* y, long ldy, const double* conf, long ldconf,
                         int t, int on_device);
/* Predict t events with a conditioned model whose cell map is bound and accumulate them against the truth (t x ldx, NULL = no
 * truth) WITHOUT materialising the t x c cell-space prediction: mean via the modes->cells map, conf = sqrt(cell variance).
 * mode_mean / mode_var (t*p, host) may be NULL. */
int gpras_gp_predict_metrics(gpras_gp* h, gpras_metrics* m, const double* xs, int t, int xs_on_device, const double* truth,
                             long ldx, int truth_on_device, double* mode_mean, double* mode_var);
/* scalars: 15 doubles = [sum e, sum e^2, sum conf, sum |e|, count(|e|<=v_tol), sum d, sum d^2, sum xm, sum (xm-mean xm)^2,
 * hits, misses, false alarms at depth_threshold, the same three at threshold 0] with d = max_t x - max_t y, xm = max_t x;
 * cells (5 x c) and rows (3 x timesteps) receive the raw per-cell / per-timestep reductions (host, may be NULL). */
/* One block (t <= 2048) of mode-space predictions with ONE VARIANCE PER MODE, consumed against the truth without writing the
 * (t x cells) prediction: y = max(M E + bias - elev, 0), conf = sqrt(V E^2).  All pointers are device pointers (M, V zero
 * padded to ldm = p16 = 32 or 64 columns and round_up(t, 32) rows; E, bias: the folded map of a gpras_pre handle).  The
 * caller-facing entry is gpras_pre_reverse_metrics. */
int gpras_metrics_update_modes(gpras_metrics* m, const double* M, const double* V, long ldm, int p16, const double* E, long lde,
                               const double* bias, const double* truth, long ldx, int t);
int gpras_metrics_finalize(gpras_metrics* m, double depth_threshold, double* scalars, double* cells, double* rows);
long gpras_metrics_timesteps(gpras_metrics* m);
int gpras_metrics_last_launches(gpras_metrics* m);
/* fi_aoi_toi (metrics.py:203-212) with a time tolerance: number of matching entries of the (t x c) arrays (host or device). */
int gpras_metrics_fidelity(const double* x, long ldx, const double* y, long ldy, int t, int c, int t_tol, double v_tol, int device,
                           double* matching);

/* ---- PreProcessor: replaces gpras/preprocess.py:947-1094 (cells <-> modes) ----------------------------------------- */
typedef struct gpras_pre gpras_pre;
#define GPRAS_HP_WSE 0
#define GPRAS_HP_DEPTH 1
#define GPRAS_HP_VELOCITY 2
/* c cells; hydraulic = GPRAS_HP_* (PreProcessor(hydraulic_parameter=...)); wet_threshold as in the constructor. */
int gpras_pre_create(gpras_pre** out, int device, int c, int hydraulic, double wet_threshold);
int gpras_pre_destroy(gpras_pre* h);
/* PreProcessor.fit (preprocess.py:947-1007): x is (n x ldx) samples x cells (host, or device when on_device), elevations
 * and weights host arrays of c doubles.  modes > 0: number of EOFs to keep; modes <= 0: keep max_modes and let the caller apply
 * North's rule to the eigenvalues.  After the call the handle holds wetness classes, input mean, EOFs, eigenvalues, score
 * statistics.  tol: residual tolerance of the retained eigenpairs relative to the largest eigenvalue; max_iter: subspace
 * iterations.  Returns 0, or GPRAS_E_STATE when the retained eigenpairs did not converge. */
int gpras_pre_fit(gpras_pre* h, const double* x, long ldx, int n, int on_device, const double* elevations, const double* weights,
                  int modes, double tol, int max_iter);
/* Truncate the fitted model to its first `modes` EOFs (North's rule is host logic on gpras_pre_get(.., eigenvalues)). */
int gpras_pre_set_modes(gpras_pre* h, int modes);
/* Restore a fitted state (PreProcessor.from_file, preprocess.py:1154-1161): full-cell-space host arrays --
 * dry (c bytes, 1 = always dry), input_mean (c), weights (c), eofs (p x c, zero on dry cells), x_mean (p), x_std (p),
 * elevations (c). */
int gpras_pre_set_state(gpras_pre* h, const unsigned char* dry, const double* input_mean, const double* weights,
                        const double* eofs, const double* x_mean, const double* x_std, const double* elevations, int p);
/* which: 0 wetness classes (c doubles: 0 unset, 1 AD, 2 TF, 3 AF), 1 input_mean (c), 2 weights incl. zero on dry cells (c),
 * 3 eofs (p x c), 4 eigenvalues = explained variance (gpras_pre_eigen_count values), 5 x_mean (p), 6 x_std (p),
 * 7 residuals of the eigenpairs relative to the largest eigenvalue (gpras_pre_eigen_count values). */
int gpras_pre_get(gpras_pre* h, int which, double* out);
int gpras_pre_modes(gpras_pre* h);
int gpras_pre_eigen_count(gpras_pre* h);
int gpras_pre_iterations(gpras_pre* h);
/* PreProcessor.transform (preprocess.py:1009-1039): z (n x p, host or device like x) = standardised EOF scores of x. */
int gpras_pre_transform(gpras_pre* h, const double* x, long ldx, int n, int on_device, double* z);
/* PreProcessor.reverse_transform (preprocess.py:1052-1084): mean (t x p) [and var (t x p), may be NULL] -> cell space
 * (t x c) host arrays; depth semantics follow the handle's hydraulic parameter. */
int gpras_pre_reverse(gpras_pre* h, const double* mean, const double* var, int t, double* cell_mean, double* cell_var);
/* The same map with the (t x cells) results left on the DEVICE: cell_mean / cell_var are device pointers with pitch
 * ldc >= gpras_pre_cell_pitch() and at least round_up(t, 64) rows, or both NULL (tiles go to an internal ring buffer: the
 * 3.2 TB of BASELINE config 5 cannot be kept).  mean / var (t x modes): host (on_device == 0) or device pointers.  One fused
 * kernel per block of events with the GENERAL variance map var @ (diag(x_std) eofs / w)^2 (preprocess.py:1081-1094), i.e.
 * the reference's default model family with one hyperparameter set per mode (gpr.py:293-308). */
int gpras_pre_reverse_device(gpras_pre* h, const double* mean, const double* var, int t, int on_device, double* cell_mean,
                             double* cell_var, long ldc);
long gpras_pre_cell_pitch(gpras_pre* h);
/* gpr.predict (per-column models: mean and variance per mode) -> reverse_transform -> wse_2_depth -> every metric of
 * gpras/metrics.py, pipeline.py:260-286, fused: the truth (t x ldx, host or device, may be NULL) is streamed, the (t x cells)
 * prediction and its confidence are never materialised.  m: an accumulator created for the same device and cell count. */
int gpras_pre_reverse_metrics(gpras_pre* h, gpras_metrics* m, const double* mean, const double* var, int t, int on_device,
                              const double* truth, long ldx, int truth_on_device);
/* Workspaces (staged input, centred samples, Gram matrix ...) are recycled across calls; trim returns them to the driver. */
int gpras_pre_trim(gpras_pre* h);
int gpras_pre_last_launches(gpras_pre* h);
/* CUDA-event milliseconds of the last fit: [column stats, centre+weight, Gram, subspace iteration, EOFs, scores, total]. */
int gpras_pre_last_stage_ms(gpras_pre* h, double* ms7);
/* Symmetric positive semi-definite 128 x 128 eigen-decomposition (one-sided Jacobi, one CTA): device pointers, pitch 128;
 * lambda descending, eigenvectors as columns of V.  Building block of the PCA fit, exported for tests. */
int gpras_dsyev128(void* cuda_stream, const double* H, double* lambda, double* V);

/* ---- inducing-input initialiser ------------------------------------------------------------ */
/* Lloyd iterations of sklearn.cluster.KMeans(n_clusters=m, random_state=0, n_init="auto") (gpr.py:313-315) from the given
 * initial centres (the k-means++ seeding stays on the host): x (n x d) and centers (m x d, in: seeds, out: final) are host
 * arrays, tol is the ABSOLUTE tolerance on the total squared centre shift (scikit-learn: 1e-4 * mean(var(x, axis=0))),
 * labels_out (n, may be NULL), inertia_out / n_iter_out (may be NULL).  Stopping rules, empty-cluster relocation and the
 * final label pass follow sklearn/cluster/_kmeans.py:_kmeans_single_lloyd; sums run in a fixed order (bitwise repeatable). */
int gpras_kmeans_lloyd(int device, const double* x, int n, int d, double* centers, int m, int max_iter, double tol,
                       int* labels_out, double* inertia_out, int* n_iter_out);

/* ---- stand-alone building blocks on device pointers (tests, composition) ------------------ */
/* C = alpha * A(.)B(.) + beta * C on the DMMA tile engine.  shape: 0 = 128x128 CTA tile (all four layouts),
 * 1 = 128x64 (row-major A, n-major B only), 2 = 128x32 (k-major B only).  m % 128 == 0, n % tile == 0, k % 32 == 0;
 * layout flags select row-major A[i][k] / k-major A[k][i] and n-major B[j][k] / k-major B[k][j]. */
int gpras_dgemm_tiles(void* cuda_stream, int shape, int a_kmajor, int b_kmajor, const double* A, long lda, const double* B,
                      long ldb, double* C, long ldc, int m, int n, int k, double alpha, double beta);
/* In-place lower Cholesky of the n x n (n % 128 == 0) device matrix A (the strictly upper part of the result is
 * unspecified); the diagonal 128-blocks of W = L^-1 come out as a by-product.  W must be zero-initialised by the caller
 * (only its lower triangle is ever written; the engine relies on the zeros above the diagonal).
 * info_dev: device int, logdet_parts_dev: n/128 device doubles. */
int gpras_dpotrf(void* cuda_stream, double* A, long lda, double* W, long ldw, int n, double* logdet_parts_dev,
                 int* info_dev);
/* Complete W = L^-1 (lower) given L and W's diagonal blocks; scratch: n x n device doubles. */
int gpras_dtrtri(void* cuda_stream, const double* L, long ldl, double* W, long ldw, double* scratch, long lds, int n);
/* Kinv = W^T W (lower tiles). */
int gpras_dlauum(void* cuda_stream, const double* W, long ldw, double* Kinv, long ldk, int n);

#ifdef __cplusplus
}
#endif
#endif /* GPRAS_B200_H */
