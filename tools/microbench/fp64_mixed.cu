// Microbenchmark (not product): do DMMA.8x8x4 (tensor pipe) and DFMA (FP64 pipe) issue concurrently on B200?
// Half of the warps of every CTA run DMMA chains, the other half DFMA chains; compare with each alone.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA
template <int ILP>
__global__ void k_mix(double* out, int iters, double a, double b, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool do_dmma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = i; c1[i] = -i; }
  if (do_dmma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma884(c0[i], c1[i], a, b);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
#pragma unroll
      for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
#pragma unroll
      for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
#pragma unroll
      for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 20000, ILP = 4;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  printf("mode,warps_per_sm,ms,dmma_TFLOPs,dfma_TFLOPs,total_TFLOPs\n");
  for (int warps : {8, 16, 32}) {
    for (int mode = 0; mode < 3; mode++) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_mix<ILP><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999, mode);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k_mix<ILP><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double wd = mode == 0 ? warps : (mode == 2 ? warps / 2 : 0), wf = mode == 1 ? warps : (mode == 2 ? warps / 2 : 0);
      const double fd = 2.0 * 256 * ILP * iters * wd * sms, ff = 2.0 * 32 * 8 * ILP * iters * wf * sms;
      printf("%d,%d,%.4f,%.2f,%.2f,%.2f\n", mode, warps, ms, fd / ms * 1e-9, ff / ms * 1e-9, (fd + ff) / ms * 1e-9);
    }
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
