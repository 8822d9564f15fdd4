// Microbenchmark (not product): per-phase clock64() timing of the leaf kernel, built with -DLEAF_TIMING.
#include <cstdio>
#include <vector>
#include <cmath>
#define LEAF_TIMING 1
#include "../../gpras_b200/csrc/leaf.cuh"
using namespace gpras;
int main() {
  const int n = 128;
  std::vector<double> a(n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i * n + j] = (i == j ? 2.0 : 0.0) + 1.0 / (1.0 + std::abs(i - j));
  double *dA, *dW, *dld; int* dinfo; long long* dt;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dW, n * n * 8); cudaMalloc(&dld, 8); cudaMalloc(&dinfo, 4); cudaMalloc(&dt, 64 * 8);
  cudaMemset(dinfo, 0, 4);
  cudaFuncSetAttribute(leaf_potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM_BYTES);
  cudaMemcpyToSymbol(g_leaf_timing, &dt, sizeof(dt));
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpy(dA, a.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES>>>(dA, n, dA, n, dW, n, dld, dinfo, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long t[64]; cudaMemcpy(t, dt, sizeof t, cudaMemcpyDeviceToHost);
    printf("rep %d: %.2f us  err=%s\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    const char* names[] = {"load", "panel0", "phaseA(sum)", "phaseB-warp0 panel(sum)", "phaseB-warp1 update(sum)", "sync-wait warp1 (sum)", "logdet", "inv diag8", "inv8", "inv16", "inv32", "inv64", "store", "total"};
    for (int i = 0; i < 14; i++) printf("  %-28s %8lld cycles\n", names[i], t[i]);
  }
  return 0;
}
