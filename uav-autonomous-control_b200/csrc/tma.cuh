// tma.cuh -- bulk asynchronous stores shared memory -> global memory through the TMA unit (sm_90+ PTX, sm_100a here):
// one instruction moves a whole staged tile, instead of a per-thread copy loop of LDS + STG + address arithmetic.
//   * 1-D bulk copy (cp.async.bulk.global.shared::cta): a contiguous run of bytes (K3's 32 x 11 row tile).
//   * 2-D tensor store (cp.async.bulk.tensor.2d.global.shared::cta) through a CUtensorMap: a box of a strided matrix
//     (K1's [missions][24 S] coefficient tile, the [samples x 13][B] state log of K2); out-of-range parts of a box are clipped.
// The tensor map is encoded on the host by the driver's cuTensorMapEncodeTiled, reached through cudaGetDriverEntryPoint
// (the library links cudart statically and does not link libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace uavb {

// ---- host
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Driver entry point, looked up once per process (nullptr when the driver does not offer it).
TensorMapEncodeTiledFn tensor_map_encoder();

// rank-2 map of a row-major matrix [dim1][dim0] of `dtype` elements whose rows are `row_stride_bytes` apart (a multiple of 16,
// base 16-byte aligned), moved in boxes of box1 x box0 elements.  Returns false when the map cannot be built (the callers
// then take their copy-loop path).
inline bool make_tensor_map_2d(CUtensorMap* out, void* base, CUtensorMapDataType dtype, unsigned long long dim0, unsigned long long dim1,
                               unsigned long long row_stride_bytes, unsigned box0, unsigned box1, CUtensorMapSwizzle swizzle) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || (row_stride_bytes & 15u) || dim0 == 0 || dim1 == 0) return false;
  const cuuint64_t dims[2] = {dim0, dim1};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box0, box1};
  const cuuint32_t estr[2] = {1, 1};
  return enc(out, dtype, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Orders this thread's earlier generic-proxy writes to shared memory before later async-proxy (TMA) reads of them.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// dst (global, 16-byte aligned) <- src (shared, 16-byte aligned), bytes a multiple of 16.
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// Box at (c0 = innermost coordinate, c1 = row) of the tensor <- src (shared, 128-byte aligned; 1024 for 128-byte swizzle).
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// ---- bulk asynchronous LOAD global -> shared, completion signalled on an mbarrier (K1's input prefetch)
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy before the first bulk copy names it
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst (shared, 16-byte aligned) <- src (global, 16-byte aligned), bytes a multiple of 16; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Wait until at most N of this thread's bulk groups are still READING their shared-memory source (the source may be reused).
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
#endif

}  // namespace uavb
