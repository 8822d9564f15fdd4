// Shared device helpers for the gpras_b200 FP64 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpras {

// DMMA.8x8x4: D(8x8) = A(8x4, row) * B(4x8, col) + C.  Fragment ownership (lane = 4*g + q):
//   a = A[g][q]      b = B[q][g]      c0,c1 = C[g][2q], C[g][2q+1]
// tcgen05.mma has no f64 kind, so on sm_100a the FP64 tensor pipe is reached through mma.sync only;
// every wider f64 shape lowers to this instruction (checked with cuobjdump -sass).
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Kernel ids shared with the host (include/gpras_b200.h GPRAS_KERNEL_*).
enum KernelId { K_RBF = 0, K_MATERN12 = 1, K_MATERN32 = 2, K_MATERN52 = 3, K_EXPONENTIAL = 4 };

// Covariance value k(r2) / variance and the log-lengthscale derivative factor F / variance
// (d k / d log l_d = variance * F * s_d), SURVEY.md section 3.6 / gpras/gpr.py:21-29.
template <int KID>
__device__ __forceinline__ void kernel_eval(double r2, double& kval, double& fval) {
  if (KID == K_RBF) {
    kval = exp(-0.5 * r2);
    fval = kval;
  } else if (KID == K_MATERN12) {
    double r = sqrt(r2);
    kval = exp(-r);
    fval = r > 0.0 ? kval / r : 0.0;
  } else if (KID == K_EXPONENTIAL) {
    double r = sqrt(r2);
    kval = exp(-0.5 * r);
    fval = r > 0.0 ? kval / (2.0 * r) : 0.0;
  } else if (KID == K_MATERN32) {
    const double s3 = 1.7320508075688772;
    double r = sqrt(r2);
    double e = exp(-s3 * r);
    kval = (1.0 + s3 * r) * e;
    fval = 3.0 * e;
  } else {  // K_MATERN52
    const double s5 = 2.23606797749979;
    double r = sqrt(r2);
    double e = exp(-s5 * r);
    kval = (1.0 + s5 * r + (5.0 / 3.0) * r2) * e;
    fval = (5.0 / 3.0) * (1.0 + s5 * r) * e;
  }
}

template <int KID>
__device__ __forceinline__ double kernel_value(double r2) {
  double k, f;
  kernel_eval<KID>(r2, k, f);
  return k;
}

}  // namespace gpras
