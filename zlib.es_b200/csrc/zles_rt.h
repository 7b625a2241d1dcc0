// zles_rt.h — host-side runtime shim: CUDA runtime in the product build, libc in the
// emulator build used by the tests (tests/emu).  Keeps zles.cu free of #ifdefs.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ZLES_EMU
#include "cuda_emu.h"
typedef void *zrt_stream_t;
typedef int zrt_err_t;
#define ZRT_OK 0
static inline zrt_err_t zrt_set_device(int) { return 0; }
static inline zrt_err_t zrt_stream_create(zrt_stream_t *s) { *s = nullptr; return 0; }
static inline zrt_err_t zrt_stream_destroy(zrt_stream_t) { return 0; }
static inline zrt_err_t zrt_malloc(void **p, size_t n) {
  n = (n + 255) & ~(size_t)255;
  *p = aligned_alloc(256, n ? n : 256);
  if (*p) memset(*p, 0xA5, n ? n : 256);  // device memory starts undefined
  return *p ? 0 : 2;
}
static inline zrt_err_t zrt_free(void *p) { free(p); return 0; }
static inline zrt_err_t zrt_host_alloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline zrt_err_t zrt_host_free(void *p) { free(p); return 0; }
static inline zrt_err_t zrt_h2d(void *d, const void *s, size_t n, zrt_stream_t) { memcpy(d, s, n); return 0; }
static inline zrt_err_t zrt_d2h(void *d, const void *s, size_t n, zrt_stream_t) { memcpy(d, s, n); return 0; }
static inline zrt_err_t zrt_memset(void *d, int v, size_t n, zrt_stream_t) { memset(d, v, n); return 0; }
static inline zrt_err_t zrt_mail(void *h, const void *d, size_t n, zrt_stream_t) { memcpy(h, d, n); return 0; }
static inline zrt_err_t zrt_sync(zrt_stream_t) { return 0; }
static inline zrt_err_t zrt_last_error() { return 0; }
static inline const char *zrt_err_str(zrt_err_t) { return "emulator"; }
template <typename K>
static inline zrt_err_t zrt_set_smem(K, int) { return 0; }
template <typename K>
static inline zrt_err_t zrt_prefer_smem(K) { return 0; }
static inline int zrt_sm_count(int) { return 4; }
typedef int zrt_event_t;
static inline zrt_err_t zrt_event_create(zrt_event_t *e) { *e = 0; return 0; }
static inline zrt_err_t zrt_event_destroy(zrt_event_t) { return 0; }
static inline zrt_err_t zrt_event_record(zrt_event_t, zrt_stream_t) { return 0; }
static inline zrt_err_t zrt_event_elapsed(float *ms, zrt_event_t, zrt_event_t) { *ms = 0; return 0; }
static inline zrt_err_t zrt_event_sync(zrt_event_t) { return 0; }
static inline zrt_err_t zrt_copy(void *d, const void *s, size_t n, zrt_stream_t) { memmove(d, s, n); return 0; }
static inline zrt_err_t zrt_stream_wait_event(zrt_stream_t, zrt_event_t) { return 0; }
#define ZLES_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
typedef cudaStream_t zrt_stream_t;
typedef cudaError_t zrt_err_t;
#define ZRT_OK cudaSuccess
static inline zrt_err_t zrt_set_device(int d) { return cudaSetDevice(d); }
static inline zrt_err_t zrt_stream_create(zrt_stream_t *s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
static inline zrt_err_t zrt_stream_destroy(zrt_stream_t s) { return cudaStreamDestroy(s); }
static inline zrt_err_t zrt_malloc(void **p, size_t n) { return cudaMalloc(p, n ? n : 256); }
static inline zrt_err_t zrt_free(void *p) { return cudaFree(p); }
static inline zrt_err_t zrt_host_alloc(void **p, size_t n) { return cudaMallocHost(p, n ? n : 1); }
static inline zrt_err_t zrt_host_free(void *p) { return cudaFreeHost(p); }
static inline zrt_err_t zrt_h2d(void *d, const void *s, size_t n, zrt_stream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st); }
static inline zrt_err_t zrt_d2h(void *d, const void *s, size_t n, zrt_stream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st); }
static inline zrt_err_t zrt_memset(void *d, int v, size_t n, zrt_stream_t st) { return cudaMemsetAsync(d, v, n, st); }
// A few words from device memory to PINNED host memory, written by a kernel over PCIe (pinned allocations are device
// accessible under unified addressing) instead of a copy: a cudaMemcpyAsync would queue on the device-to-host copy
// engine behind whatever bulk copy another stream has in flight — the small read-backs between the steps of a pipelined
// call (candidate counts, status words) then wait for hundreds of megabytes.  n is a multiple of 4.
static __global__ void zrt_k_mail(unsigned *dst, const unsigned *src, unsigned nwords) {
  for (unsigned i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
static __global__ void zrt_k_mail_wide(unsigned *dst, const unsigned *src, size_t nwords) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
static inline zrt_err_t zrt_mail(void *h, const void *d, size_t n, zrt_stream_t st) {
  const size_t nw = n / 4;
  if (nw <= 64) zrt_k_mail<<<1, 32, 0, st>>>(reinterpret_cast<unsigned *>(h), reinterpret_cast<const unsigned *>(d), (unsigned)nw);
  else zrt_k_mail_wide<<<(unsigned)((nw + 1023) / 1024 < 64 ? (nw + 1023) / 1024 : 64), 256, 0, st>>>(reinterpret_cast<unsigned *>(h), reinterpret_cast<const unsigned *>(d), nw);
  return cudaGetLastError();
}
static inline zrt_err_t zrt_sync(zrt_stream_t st) { return cudaStreamSynchronize(st); }
static inline zrt_err_t zrt_last_error() { return cudaGetLastError(); }
static inline const char *zrt_err_str(zrt_err_t e) { return cudaGetErrorString(e); }
template <typename K>
static inline zrt_err_t zrt_set_smem(K kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
// all of the SM's L1/shared array as shared memory: for kernels whose residency is bound by shared memory per CTA
template <typename K>
static inline zrt_err_t zrt_prefer_smem(K kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
typedef cudaEvent_t zrt_event_t;
static inline zrt_err_t zrt_event_create(zrt_event_t *e) { return cudaEventCreate(e); }
static inline zrt_err_t zrt_event_destroy(zrt_event_t e) { return cudaEventDestroy(e); }
static inline zrt_err_t zrt_event_record(zrt_event_t e, zrt_stream_t s) { return cudaEventRecord(e, s); }
static inline zrt_err_t zrt_event_elapsed(float *ms, zrt_event_t a, zrt_event_t b) { return cudaEventElapsedTime(ms, a, b); }
static inline zrt_err_t zrt_event_sync(zrt_event_t e) { return cudaEventSynchronize(e); }
static inline zrt_err_t zrt_stream_wait_event(zrt_stream_t s, zrt_event_t e) { return cudaStreamWaitEvent(s, e, 0); }
// device, host or peer-mapped pointers on either side (unified addressing)
static inline zrt_err_t zrt_copy(void *d, const void *s, size_t n, zrt_stream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, st); }
static inline int zrt_sm_count(int dev) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}
#define ZLES_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
