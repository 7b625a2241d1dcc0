// tma.cuh — TMA bulk global->shared copy (cp.async.bulk, SASS: UBLKCP) completing on an
// mbarrier.  Used to stage a deflate block and its 32 KiB window in shared memory.
// The emulator build replaces it with a plain cooperative copy.
#pragma once
#include "zles_dev.h"

namespace zles {

#ifndef ZLES_EMU
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// generic-proxy writes to smem must be ordered before the async proxy (TMA) touches the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// Copies `bytes` from global to shared memory with the whole CTA.  `bar`/`parity`
// track the mbarrier phase across calls (one call per staged block).  Falls back
// to vector/byte loads when source or size are not 16-byte aligned.
// Must be called by all threads; ends with a CTA barrier.
__device__ __forceinline__ void stage_g2s(u8 *dst, const u8 *src, u32 bytes, u64 *bar, u32 &parity) {
#ifndef ZLES_EMU
  const u32 bulk = (((uintptr_t)src & 15) == 0) ? (bytes & ~15u) : 0;
  if (bulk) {
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, bulk);
      for (u32 o = 0; o < bulk; o += 16384) bulk_g2s(dst + o, src + o, umin(16384u, bulk - o), bar);
    }
    for (u32 i = bulk + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    u32 spins = 0;
    while (!mbar_try_wait(bar, parity)) {
      if (++spins > (1u << 24)) __trap();  // never hang the GPU on a lost copy
    }
    parity ^= 1;
    __syncthreads();
    return;
  }
#else
  (void)bar; (void)parity;
#endif
  __syncthreads();
  for (u32 i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

}  // namespace zles
