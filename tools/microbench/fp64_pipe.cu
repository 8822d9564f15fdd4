// Microbenchmark: FP64 pipe rates on B200 (sm_100a). Not product code.
// Measures DMMA.8x8x4 and DFMA throughput per SM as a function of resident warps
// and independent accumulator chains (ILP), plus dependent-issue latency.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock_khz %d\n", p.name, sms, p.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
  const int iters = 20000;
  int warps_list[] = {1, 2, 4, 8, 16, 32};
  printf("kind,ilp,warps_per_sm,ms,TFLOPs,instr_per_clk_per_sm(@1965MHz)\n");
#define RUN_DMMA(ILP) for (int w : warps_list) { \
    int thr = w * 32; int ctas = 1; if (thr > 1024) { ctas = thr / 1024; thr = 1024; } \
    float ms = time_ms([&] { k_dmma<ILP><<<sms * ctas, thr>>>(out, iters, 1.0000001, 0.9999999); }); \
    double flops = 2.0 * 256 * (double)ILP * iters * w * sms; \
    double ninstr = (double)ILP * iters * w; \
    printf("dmma,%d,%d,%.4f,%.3f,%.4f\n", ILP, w, ms, flops / ms * 1e-9, ninstr / (ms * 1e-3 * 1.965e9)); }
  RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8) RUN_DMMA(16) RUN_DMMA(32)
#define RUN_DFMA(ILP) for (int w : warps_list) { \
    int thr = w * 32; int ctas = 1; if (thr > 1024) { ctas = thr / 1024; thr = 1024; } \
    float ms = time_ms([&] { k_dfma<ILP><<<sms * ctas, thr>>>(out, iters, 1.0000001, 0.9999999); }); \
    double flops = 2.0 * 32 * (double)ILP * iters * w * sms; \
    double ninstr = (double)ILP * iters * w; \
    printf("dfma,%d,%d,%.4f,%.3f,%.4f\n", ILP, w, ms, flops / ms * 1e-9, ninstr / (ms * 1e-3 * 1.965e9)); }
  RUN_DFMA(1) RUN_DFMA(4) RUN_DFMA(8) RUN_DFMA(16)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
