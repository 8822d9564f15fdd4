// Microbenchmark (not product): achievable HBM READ bandwidth of the access patterns the streaming kernels use, on a
// rows x cols FP64 matrix (default 8192 x 200000 = 13 GB): contiguous grid-stride reads, the column-slab pattern of
// colstats_kernel (LDG), and TMA 2-D box loads into a shared-memory ring for several box shapes.
#include <cstdio>
#include <cstdlib>
#include "../../gpras_b200/csrc/tma.cuh"
#include "../../gpras_b200/csrc/pre_kernels.cuh"
using namespace gpras;

__global__ void __launch_bounds__(256) contiguous_kernel(const double2* __restrict__ x, long n2, double* out) {
  double s = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 8
  for (; i < n2; i += stride) {
    const double2 v = __ldg(x + i);
    s += v.x + v.y;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// CTA = slab of BOX_C columns x a row range; thread 0 keeps STAGES boxes of BOX_R rows in flight; all threads sum.
template <int BOX_C, int BOX_R, int STAGES, int THREADS>
__global__ void __launch_bounds__(THREADS) tma_slab_kernel(const __grid_constant__ CUtensorMap map, int rows_per_cta, int n, double* out) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  constexpr int STAGE_DOUBLES = BOX_C * BOX_R;
  constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * 8;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * BOX_C, r0 = blockIdx.y * rows_per_cta;
  int r1 = r0 + rows_per_cta;
  if (r1 > n) r1 = n;
  const int n_it = (r1 - r0 + BOX_R - 1) / BOX_R;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < STAGES && s < n_it; s++) {
      mbar_expect_tx(&full[s], STAGE_BYTES);
      tma_load_2d(smem + s * STAGE_DOUBLES, &map, c0, r0 + s * BOX_R, &full[s]);
    }
  double acc = 0.0;
  for (int it = 0; it < n_it; it++) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const double2* t = reinterpret_cast<const double2*>(smem + s * STAGE_DOUBLES);
#pragma unroll 4
    for (int e = tid; e < STAGE_DOUBLES / 2; e += THREADS) {
      const double2 v = t[e];
      acc += v.x + v.y;
    }
    __syncthreads();
    if (tid == 0 && it + STAGES < n_it) {
      mbar_expect_tx(&full[s], STAGE_BYTES);
      tma_load_2d(smem + s * STAGE_DOUBLES, &map, c0, r0 + (it + STAGES) * BOX_R, &full[s]);
    }
  }
  acc = warp_sum(acc);
  if ((tid & 31) == 0) atomicAdd(out, acc);
}

template <typename F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; i++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

template <int BOX_C, int BOX_R, int STAGES, int THREADS>
static void run_tma(const double* x, long ld, int n, int c, double* out, int splits) {
  CUtensorMap map;
  if (!tma_map_2d_f64(&map, x, c, n, ld, BOX_C, BOX_R)) {
    printf("tensor map failed\n");
    return;
  }
  const int smem = BOX_C * BOX_R * 8 * STAGES;
  cudaFuncSetAttribute(tma_slab_kernel<BOX_C, BOX_R, STAGES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int rows_per = (n + splits - 1) / splits;
  dim3 grid((c + BOX_C - 1) / BOX_C, splits);
  const float ms = time_ms([&] { tma_slab_kernel<BOX_C, BOX_R, STAGES, THREADS><<<grid, THREADS, smem>>>(map, rows_per, n, out); });
  printf("TMA box %3d cols x %3d rows, %d stages (%3d KB), %3d thr, %2d row splits: %7.3f ms  %7.1f GB/s  err=%s\n", BOX_C, BOX_R, STAGES,
         smem / 1024, THREADS, splits, ms, 8.0 * n * c / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8192, c = argc > 2 ? atoi(argv[2]) : 200000;
  const long ld = (c + 127) / 128 * 128;
  double *x, *out, *part, *elev;
  cudaMalloc(&x, (size_t)n * ld * 8), cudaMalloc(&out, 8), cudaMalloc(&part, (size_t)64 * 3 * ld * 8), cudaMalloc(&elev, ld * 8);
  cudaMemset(x, 0, (size_t)n * ld * 8), cudaMemset(elev, 0, ld * 8);
  printf("matrix %d x %d (pitch %ld): %.2f GB\n", n, c, ld, 8.0 * n * ld * 1e-9);
  {
    const long n2 = (long)n * ld / 2;
    const float ms = time_ms([&] { contiguous_kernel<<<148 * 8, 256>>>((const double2*)x, n2, out); });
    printf("contiguous grid-stride double2 reads:                      %7.3f ms  %7.1f GB/s\n", ms, 8.0 * n * ld / ms * 1e-6);
  }
  for (int splits : {4, 8, 16, 32}) {
    const int rows_per = (n + splits - 1) / splits;
    const float ms = time_ms([&] { colstats_kernel<<<dim3((c + 255) / 256, splits), 128>>>(x, ld, n, c, elev, 0, rows_per, part, ld); });
    printf("colstats_kernel (LDG, 256-col slabs), %2d row splits:        %7.3f ms  %7.1f GB/s\n", splits, ms, 8.0 * n * c / ms * 1e-6);
  }
  run_tma<256, 16, 4, 128>(x, ld, n, c, out, 4);
  run_tma<256, 16, 4, 128>(x, ld, n, c, out, 16);
  run_tma<256, 8, 6, 128>(x, ld, n, c, out, 8);
  run_tma<256, 32, 3, 256>(x, ld, n, c, out, 8);
  run_tma<128, 32, 4, 128>(x, ld, n, c, out, 4);
  run_tma<128, 32, 3, 128>(x, ld, n, c, out, 8);
  run_tma<128, 64, 3, 256>(x, ld, n, c, out, 4);
  run_tma<64, 64, 4, 128>(x, ld, n, c, out, 4);
  run_tma<32, 128, 4, 128>(x, ld, n, c, out, 2);
  run_tma<16, 128, 6, 128>(x, ld, n, c, out, 2);
  run_tma<16, 256, 4, 128>(x, ld, n, c, out, 1);
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
