// TMA (cp.async.bulk.tensor) tile loads with mbarrier completion for sm_100a, and the host-side tensor-map encoder.
//
// The column kernels fetch a tile as N rows of W neighbouring kx (128-byte segments N*Ny*NXP*8 bytes apart).  With
// per-thread cp.async every thread issues 16-48 copies and the data is staged through L1, which caps the bytes in flight
// once shared memory takes most of the SM's 228 KB (measured: double-buffering the tile with cp.async made the fused z
// pass 30 % SLOWER, profiles/r02_d_zmid_variants.log).  One elected thread issuing box copies through the TMA unit has
// neither problem: no address arithmetic in the other threads, no L1 staging, completion on an mbarrier.
#pragma once
#include <cuda.h>  // CUtensorMap and its enums (types only; the encoder entry point is resolved at run time)
#include <cuda_runtime.h>
#include <cstdint>

namespace kw {

// 3-D tensor map over FP32 elements: dims (d0 fastest), byte strides of dims 1 and 2, box (b0, b1, b2).  Returns false when
// the driver entry point is missing or the encoder rejects the shape (callers fall back to the cp.async kernels).
bool make_tensor_map_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                        uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// box copy global -> shared, coordinates in elements (c0 fastest); completes `bytes of the box` on the mbarrier
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// the two tensor maps of a fused z pass: the spectrum [kz][ky][2 NXP floats] and the real multiplier [kz][ky][NXP]
struct ZMaps {
  CUtensorMap in, mul;
};

}  // namespace kw
