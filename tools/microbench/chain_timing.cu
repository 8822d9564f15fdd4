// Microbenchmark (not product): the kernels of the Cholesky's critical chain, each timed as 20 dependent launches on one
// stream (so launch + drain latency is included, as it is on the chain), plus the whole factorisation at a few sizes.
#include <cstdio>
#include <cmath>
#include <vector>
#define POTRF_TIMELINE 1
#include "../../gpras_b200/csrc/host_common.cuh"

template <typename F>
static float time_us(F f, int reps = 20) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / reps;
}

int main() {
  const int n = 8192, nt = n / 128;
  prepare_device();
  std::vector<double> a((size_t)n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) a[(size_t)i * n + j] = (i == j ? 3.0 : 0.0) + 1.0 / (1.0 + std::abs(i - j));
  double *dA, *dA0, *dW, *dld;
  int* dinfo;
  cudaMalloc(&dA, (size_t)n * n * 8), cudaMalloc(&dA0, (size_t)n * n * 8), cudaMalloc(&dW, (size_t)n * n * 8);
  cudaMalloc(&dld, nt * 8), cudaMalloc(&dinfo, 4);
  cudaMemset(dinfo, 0, 4), cudaMemset(dW, 0, (size_t)n * n * 8);
  cudaMemcpy(dA0, a.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dA, dA0, (size_t)n * n * 8, cudaMemcpyDeviceToDevice);
  const bool only_potrf = getenv("CHAIN_ONLY_POTRF") != nullptr;
  if (!only_potrf) {
  printf("leaf_potrf (1 CTA)            %7.2f us\n", time_us([&] { leaf_potrf_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES>>>(dA, n, dld, dinfo, 0); }));
  cudaMemcpy(dA, dA0, (size_t)n * n * 8, cudaMemcpyDeviceToDevice);
  printf("leaf_potrf_inv (1 CTA, r01)   %7.2f us\n", time_us([&] { leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES>>>(dA, n, dA, n, dW, n, dld, dinfo, 0); }));
  printf("leaf_inv x1                   %7.2f us\n", time_us([&] { leaf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES>>>(dA, n, dW, n); }));
  printf("leaf_inv x64 (one launch)     %7.2f us\n", time_us([&] { leaf_inv_kernel<<<64, LEAF_THREADS, LEAF_SMEM_BYTES>>>(dA, n, dW, n); }));
  for (int rem : {4, 16, 32, 48, 63})
    printf("trsm_panel rem=%2d (%3d CTAs)  %7.2f us\n", rem, rem * 4,
           time_us([&] { trsm_panel_kernel<<<rem * (LEAF_N / TRSM_ROWS), TRSM_THREADS, TRSM_SMEM_BYTES>>>(dA, n, 0); }));
  for (int kp : {128, 256}) {
    GemmDesc c = make_desc(dA + (size_t)128 * n, n, dA + (size_t)128 * n, n, dA + (size_t)128 * (n + 1), n, 2, 4, kp);
    c.alpha = -1.0, c.beta = 1.0;
    printf("diag update K=%3d (8 CTAs)    %7.2f us\n", kp, time_us([&] { launch_gemm(0, false, false, c, 1, nullptr, SHAPE_T); }));
    for (int rem : {16, 32, 63}) {
      GemmDesc cc = make_desc(dA + (size_t)256 * n, n, dA + (size_t)128 * n, n, dA + (size_t)256 * n + 128, n, 2 * (rem - 1), 4, kp);
      cc.alpha = -1.0, cc.beta = 1.0;
      printf("col update  K=%3d rem=%2d      %7.2f us\n", kp, rem, time_us([&] { launch_gemm(0, false, false, cc, 1, nullptr, SHAPE_T); }));
    }
  }
  // an empty kernel pair: the floor of two dependent launches
  printf("dependent-launch floor        %7.2f us\n", time_us([&] { splitk_reduce_kernel<<<1, 32>>>(dA, 0, 0, 0, dA); }));
  }
  // whole factorisation
  static LookAhead la;
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  for (int nn : {2048, 4096, 8192}) {
    if (only_potrf && nn != 8192) continue;
    float best = 1e30f;
    for (int rep = 0; rep < (only_potrf ? 2 : 4); rep++) {
      cudaMemcpy2DAsync(dA, (size_t)nn * 8, dA0, (size_t)n * 8, (size_t)nn * 8, nn, cudaMemcpyDeviceToDevice, s);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0), cudaEventCreate(&e1);
      cudaEventRecord(e0, s);
      int r = potrf_impl(s, la, dA, nn, dW, nn, nn, dld, dinfo, nullptr);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (r) printf("potrf_impl failed: %s\n", g_err.c_str());
      if (rep && ms < best) best = ms;
    }
    printf("potrf n=%5d                 %8.3f ms   (%.1f TFLOP/s)\n", nn, best, (double)nn * nn * nn / 3.0 / (best * 1e-3) * 1e-12);
  }
  if (getenv("CHAIN_TIMELINE")) {  // one more n = 8192 factorisation with per-step timestamps
    PotrfTimeline tl;
    cudaEventCreate(&tl.t0);
    cudaMemcpyAsync(dA, dA0, (size_t)n * n * 8, cudaMemcpyDeviceToDevice, s);
    cudaEventRecord(tl.t0, s);
    g_timeline = &tl;
    potrf_impl(s, la, dA, n, dW, n, n, dld, dinfo, nullptr);
    g_timeline = nullptr;
    cudaDeviceSynchronize();
    printf("step rem | bulk_start bulk_end (dur) | trsm_end diag_end leaf_end col_end   [us since start]\n");
    auto at = [&](int j, int k) {
      float ms = -1.f;
      if ((size_t)j * 6 + k < tl.ev.size() && cudaEventElapsedTime(&ms, tl.t0, tl.get(j, k)) != cudaSuccess) { cudaGetLastError(); ms = -1e-3f; }
      return ms * 1e3f;
    };
    for (int j = 0; j + 1 < nt; j++)
      printf("%3d %3d | %9.1f %9.1f (%7.1f) | %9.1f %9.1f %9.1f %9.1f\n", j, nt - j - 1, at(j, 0), at(j, 1), at(j, 1) - at(j, 0), at(j, 2), at(j, 3),
             at(j, 4), at(j, 5));
  }
  int info = 0;
  cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
  printf("info=%d  err=%s\n", info, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
