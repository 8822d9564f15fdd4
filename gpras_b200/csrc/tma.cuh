// TMA (bulk asynchronous copy engine) helpers for the HBM-bound streaming kernels (sm_100a): tensor-map encoding on the
// host, mbarrier + cp.async.bulk.tensor on the device.  The tensor-map encoder lives in the driver; it is fetched through
// the runtime (cudaGetDriverEntryPoint), so the library still links cudart only.
#pragma once
#include <cuda.h>  // CUtensorMap and its enums (types only)
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpras {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (the TMA unit) before the first copy is issued
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for phase `parity` of the barrier.  A copy that never lands (a bad tensor map) must not hang the GPU: after ~2 s
// the kernel traps, which surfaces as a launch failure on the host.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// one box of a 2-D tensor (x = innermost coordinate, in elements) -> shared memory; completion is counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- host ----
typedef CUresult (*TmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmaEncodeTiledFn tma_encoder() {
  static TmaEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<TmaEncodeTiledFn>(p);
  }();
  return fn;
}

// Row-major FP64 matrix `rows x cols` with pitch `ld` elements (base 16-byte aligned, ld even), boxes of
// `box_rows x box_cols` (each <= 256, box_cols * 8 a multiple of 16).  Out-of-range elements of a box read as 0.
// Returns false when the map cannot be built (unaligned input, driver too old): the caller takes its non-TMA CUDA path.
// swizzle128: boxes of exactly 16 columns (128 B) land with their 16-byte chunks XOR-ed by (row & 7) (CU_TENSOR_MAP_SWIZZLE_128B),
// which, together with a matching k order of the DMMA fragments, makes 8-row x 4-column fragment reads bank-conflict free
// without padding (TMA cannot pad rows).  The shared-memory destination must then be 1024-byte aligned.
inline bool tma_map_2d_f64(CUtensorMap* map, const double* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                           uint32_t box_rows, bool swizzle128 = false) {
  TmaEncodeTiledFn enc = tma_encoder();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (ld & 1) || box_cols > 256 || box_rows > 256 || (box_cols & 1)) return false;
  if (swizzle128 && box_cols != 16) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(double)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gpras
