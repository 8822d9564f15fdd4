// Lloyd's k-means on the device: the inducing-input initialiser of the reference (gpras/gpr.py:310-315 calls
// sklearn.cluster.KMeans(n_clusters=M, random_state=0, n_init="auto"), i.e. ONE k-means++ seeding followed by Lloyd
// iterations).  The seeding is a handful of random draws and stays on the host (scikit-learn's own kmeans_plusplus with the
// same random state, so the initial centres are identical); the iterations -- O(N M D) per step -- run here.
//
//   assign   every point finds its nearest centre (first minimum on ties) by direct differences, centres staged through
//            shared memory in chunks; the squared distance to it is kept (inertia, empty-cluster relocation);
//   update   one CTA per centre sums its points in a fixed order (no floating-point atomics: bitwise repeatable) and
//            emits the new centre, its weight and its squared shift.
// The host drives the loop with scikit-learn's stopping rules (_kmeans_single_lloyd): stop when no label changed or when
// the total squared centre shift is <= tol; if the labels did change in the last step, assign once more so that labels
// and centres agree; inertia = sum of squared distances.  Empty clusters are re-seeded at the points farthest from
// their centres, as scikit-learn does.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "host_common.cuh"

namespace {

constexpr int KM_THREADS = 128, KM_CHUNK = 32;

// grid ceil(n / 128); thread = one point.  labels / dist out; changed[0] counts label changes (integer atomics only).
__global__ void __launch_bounds__(KM_THREADS) kmeans_assign_kernel(const double* __restrict__ X, int n, int d,
                                                                   const double* __restrict__ Cn, int m, int* __restrict__ labels,
                                                                   double* __restrict__ dist, int* __restrict__ changed) {
  extern __shared__ double sc[];  // [KM_CHUNK][d]
  const int i = blockIdx.x * KM_THREADS + threadIdx.x;
  const double* x = X + (long)(i < n ? i : n - 1) * d;
  double best = INFINITY;
  int arg = 0;
  for (int c0 = 0; c0 < m; c0 += KM_CHUNK) {
    const int cn = m - c0 < KM_CHUNK ? m - c0 : KM_CHUNK;
    __syncthreads();
    for (int e = threadIdx.x; e < cn * d; e += KM_THREADS) sc[e] = Cn[(long)c0 * d + e];
    __syncthreads();
    for (int c = 0; c < cn; c++) {
      double s = 0.0;
      for (int k = 0; k < d; k++) {
        const double t = x[k] - sc[c * d + k];
        s = fma(t, t, s);
      }
      if (s < best) best = s, arg = c0 + c;
    }
  }
  if (i < n) {
    if (labels[i] != arg) atomicAdd(changed, 1);
    labels[i] = arg;
    dist[i] = best;
  }
}

// grid m; CTA j sums the points labelled j: thread t takes points t, t + 256, ... (fixed order), then a fixed-shape tree.
__global__ void __launch_bounds__(256) kmeans_update_kernel(const double* __restrict__ X, int n, int d, const int* __restrict__ labels,
                                                            const double* __restrict__ Cold, double* __restrict__ Cnew,
                                                            double* __restrict__ weight, double* __restrict__ shift2) {
  extern __shared__ double red[];  // [256][d + 1]
  const int j = blockIdx.x, tid = threadIdx.x;
  double* mine = red + (long)tid * (d + 1);
  for (int k = 0; k <= d; k++) mine[k] = 0.0;
  for (int i = tid; i < n; i += 256)
    if (labels[i] == j) {
      const double* x = X + (long)i * d;
      for (int k = 0; k < d; k++) mine[k] += x[k];
      mine[d] += 1.0;
    }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      const double* other = red + (long)(tid + s) * (d + 1);
      for (int k = 0; k <= d; k++) mine[k] += other[k];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double w = red[d];
    double sh = 0.0;
    for (int k = 0; k < d; k++) {
      const double c = w > 0.0 ? red[k] / w : Cold[(long)j * d + k];  // an empty cluster keeps its centre (relocated by the host)
      const double t = c - Cold[(long)j * d + k];
      sh = fma(t, t, sh);
      Cnew[(long)j * d + k] = c;
    }
    weight[j] = w;
    shift2[j] = sh;
  }
}

}  // namespace

extern "C" int gpras_kmeans_lloyd(int device, const double* x, int n, int d, double* centers, int m, int max_iter, double tol,
                                  int* labels_out, double* inertia_out, int* n_iter_out) {
  if (!x || !centers || n <= 0 || d <= 0 || m <= 0 || m > n || max_iter <= 0) return fail(GPRAS_E_ARG, "bad argument");
  if (d > 64) return fail(GPRAS_E_ARG, "d > 64 features is not supported");
  if (gpras_device_count() <= device || device < 0) return fail(GPRAS_E_CUDA, "no such CUDA device (no CPU fallback)");
  DeviceGuard guard(device);
  double *dX = nullptr, *dC[2] = {nullptr, nullptr}, *dDist = nullptr, *dW = nullptr, *dS = nullptr;
  int *dL = nullptr, *dChanged = nullptr;
  auto cleanup = [&]() {
    cudaFree(dX), cudaFree(dC[0]), cudaFree(dC[1]), cudaFree(dDist), cudaFree(dW), cudaFree(dS), cudaFree(dL), cudaFree(dChanged);
  };
#define KM_CU(expr)                                  \
  do {                                               \
    cudaError_t e__ = (expr);                        \
    if (e__ != cudaSuccess) {                        \
      cleanup();                                     \
      return fail(GPRAS_E_CUDA, #expr, e__);         \
    }                                                \
  } while (0)
  KM_CU(cudaMalloc((void**)&dX, sizeof(double) * n * d));
  KM_CU(cudaMalloc((void**)&dC[0], sizeof(double) * m * d));
  KM_CU(cudaMalloc((void**)&dC[1], sizeof(double) * m * d));
  KM_CU(cudaMalloc((void**)&dDist, sizeof(double) * n));
  KM_CU(cudaMalloc((void**)&dW, sizeof(double) * m));
  KM_CU(cudaMalloc((void**)&dS, sizeof(double) * m));
  KM_CU(cudaMalloc((void**)&dL, sizeof(int) * n));
  KM_CU(cudaMalloc((void**)&dChanged, sizeof(int)));
  KM_CU(cudaMemcpy(dX, x, sizeof(double) * n * d, cudaMemcpyHostToDevice));
  KM_CU(cudaMemcpy(dC[0], centers, sizeof(double) * m * d, cudaMemcpyHostToDevice));
  KM_CU(cudaMemset(dL, 0xff, sizeof(int) * n));  // labels start at -1 (scikit-learn: np.full(n, -1))
  const int assign_smem = KM_CHUNK * d * (int)sizeof(double), update_smem = 256 * (d + 1) * (int)sizeof(double);
  KM_CU(cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, update_smem));
  std::vector<double> hW(m), hS(m), hC((size_t)m * d), hDist;
  std::vector<int> hL;
  int cur = 0, it = 0, changed = 1;
  bool strict = false;
  auto assign = [&](const double* C) -> cudaError_t {
    cudaMemset(dChanged, 0, sizeof(int));
    kmeans_assign_kernel<<<(n + KM_THREADS - 1) / KM_THREADS, KM_THREADS, assign_smem>>>(dX, n, d, C, m, dL, dDist, dChanged);
    return cudaMemcpy(&changed, dChanged, sizeof(int), cudaMemcpyDeviceToHost);
  };
  for (it = 0; it < max_iter; it++) {
    KM_CU(assign(dC[cur]));
    kmeans_update_kernel<<<m, 256, update_smem>>>(dX, n, d, dL, dC[cur], dC[cur ^ 1], dW, dS);
    KM_CU(cudaMemcpy(hW.data(), dW, sizeof(double) * m, cudaMemcpyDeviceToHost));
    KM_CU(cudaMemcpy(hS.data(), dS, sizeof(double) * m, cudaMemcpyDeviceToHost));
    int n_empty = 0;
    for (int j = 0; j < m; j++) n_empty += hW[j] == 0.0;
    if (n_empty > 0) {
      // scikit-learn's _relocate_empty_clusters_dense: the points farthest from their centres become the new centres of
      // the empty clusters and leave their old clusters (integer / O(N) bookkeeping on the host; rare)
      hDist.resize(n), hL.resize(n);
      KM_CU(cudaMemcpy(hDist.data(), dDist, sizeof(double) * n, cudaMemcpyDeviceToHost));
      KM_CU(cudaMemcpy(hL.data(), dL, sizeof(int) * n, cudaMemcpyDeviceToHost));
      KM_CU(cudaMemcpy(hC.data(), dC[cur ^ 1], sizeof(double) * m * d, cudaMemcpyDeviceToHost));
      std::vector<double> hCold((size_t)m * d);
      KM_CU(cudaMemcpy(hCold.data(), dC[cur], sizeof(double) * m * d, cudaMemcpyDeviceToHost));
      std::vector<int> order(n);
      std::iota(order.begin(), order.end(), 0);
      std::partial_sort(order.begin(), order.begin() + n_empty, order.end(),
                        [&](int a, int b) { return hDist[a] > hDist[b] || (hDist[a] == hDist[b] && a < b); });
      int k = 0;
      for (int j = 0; j < m; j++) {
        if (hW[j] != 0.0) continue;
        const int far = order[k++], old = hL[far];
        for (int q = 0; q < d; q++) {  // remove the point from its old cluster's mean, make it the empty cluster's centre
          const double xv = x[(size_t)far * d + q];
          if (hW[old] > 1.0) hC[(size_t)old * d + q] = (hC[(size_t)old * d + q] * hW[old] - xv) / (hW[old] - 1.0);
          hC[(size_t)j * d + q] = xv;
        }
        hW[old] -= 1.0, hW[j] = 1.0;
      }
      for (int j = 0; j < m; j++) {
        double sh = 0.0;
        for (int q = 0; q < d; q++) {
          const double t = hC[(size_t)j * d + q] - hCold[(size_t)j * d + q];
          sh += t * t;
        }
        hS[j] = sh;
      }
      KM_CU(cudaMemcpy(dC[cur ^ 1], hC.data(), sizeof(double) * m * d, cudaMemcpyHostToDevice));
    }
    cur ^= 1;
    if (changed == 0) {  // strict convergence: no label moved
      strict = true;
      it++;
      break;
    }
    double tot = 0.0;
    for (int j = 0; j < m; j++) tot += hS[j];
    if (tot <= tol) {
      it++;
      break;
    }
  }
  if (!strict) KM_CU(assign(dC[cur]));  // labels that agree with the final centres
  hDist.resize(n);
  KM_CU(cudaMemcpy(hDist.data(), dDist, sizeof(double) * n, cudaMemcpyDeviceToHost));
  double inertia = 0.0;
  for (int i = 0; i < n; i++) inertia += hDist[i];
  KM_CU(cudaMemcpy(centers, dC[cur], sizeof(double) * m * d, cudaMemcpyDeviceToHost));
  if (labels_out) KM_CU(cudaMemcpy(labels_out, dL, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (inertia_out) *inertia_out = inertia;
  if (n_iter_out) *n_iter_out = it;
#undef KM_CU
  cleanup();
  return 0;
}
