// cuda_emu.h — a tiny CPU thread emulator for the CUDA sources (TESTS ONLY).
//
// The container that builds this repo has no GPU, and GPU time is rationed, so
// the kernels under zlib.es_b200/csrc are written against a small portability
// shim (zles_dev.h).  Compiled with -DZLES_EMU the same sources run here: one
// fiber per CUDA thread, __syncthreads() = a CTA barrier, warp intrinsics =
// a per-warp rendezvous keyed by the participation mask.  This is a debugging
// aid for logic errors (indexing, barrier placement, warp votes) and deadlocks.  It is never loaded by the product: zlib.es_b200/_capi.py only
// opens libzles.so (the nvcc build) and raises if that is missing.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__
#define __constant__ static const

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

namespace emu {

// One CUDA thread = one fiber (user-level context).  All fibers of a CTA run on one OS
// thread under a round-robin scheduler, so __syncthreads(), warp collectives and
// spin-waits are plain loops around yield() — no locks, deterministic, and fast enough
// to run whole kernels.  Several CTAs run in parallel on several OS threads.
struct Fiber {
  void *sp = nullptr;      // saved stack pointer
  void *stack = nullptr;
  bool done = false;
  uint3 tid{0, 0, 0};
};
struct WarpSlot {          // rendezvous state of one warp collective
  uint64_t vals[32];
  uint64_t out[32];
  uint32_t mask = 0;       // participation mask of the collective in flight (0 = idle)
  uint32_t arrived = 0;
  uint32_t pending_read = 0;
  uint32_t gen = 0;
};
struct BlockCtx {
  unsigned nthreads = 0;
  std::vector<Fiber> fibers;
  std::vector<WarpSlot> warps;
  unsigned barrier_arrived = 0;
  unsigned barrier_gen = 0;
  int vote = 0;            // __syncthreads_or accumulator
  uint8_t *smem = nullptr;
  uint3 bid{0, 0, 0};
  dim3 bdim, gdim;
  unsigned cur = 0;        // running fiber
  void *sched_sp = nullptr;
  const std::function<void()> *body = nullptr;
};

extern thread_local BlockCtx *g_blk;

inline uint8_t *dyn_smem() { return g_blk->smem; }
void yield();              // back to the scheduler; returns when this fiber is picked again

inline void block_barrier() {
  BlockCtx *b = g_blk;
  const unsigned gen = b->barrier_gen;
  if (++b->barrier_arrived == b->nthreads) {
    b->barrier_arrived = 0;
    b->barrier_gen++;
    return;
  }
  while (b->barrier_gen == gen) yield();
}

// Every participating lane publishes v and receives the values of all lanes in mask.
inline void warp_exchange(uint32_t mask, uint64_t v, uint64_t out[32]) {
  BlockCtx *b = g_blk;
  const unsigned t = b->cur, lane = t & 31;
  if (!((mask >> lane) & 1)) {
    fprintf(stderr, "emu: lane %u not in mask %08x\n", lane, mask);
    abort();
  }
  WarpSlot &s = b->warps[t >> 5];
  while (s.pending_read || (s.mask && s.mask != mask)) yield();  // previous collective still being read / another in flight
  s.mask = mask;
  s.vals[lane] = v;
  s.arrived |= 1u << lane;
  if (s.arrived == mask) {
    memcpy(s.out, s.vals, sizeof(s.out));
    s.pending_read = mask;
    s.arrived = 0;
    s.mask = 0;
    s.gen++;
  } else {
    const uint32_t gen = s.gen;
    while (s.gen == gen) yield();
  }
  memcpy(out, s.out, sizeof(s.out));
  s.pending_read &= ~(1u << lane);
}

// Runs `body` once per CUDA thread.
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);

}  // namespace emu

#define threadIdx (emu::g_blk->fibers[emu::g_blk->cur].tid)
#define blockIdx (emu::g_blk->bid)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)

static inline void __syncthreads() { emu::block_barrier(); }
static inline int __syncthreads_or(int pred) {
  emu::BlockCtx *b = emu::g_blk;
  if (pred) b->vote = 1;
  emu::block_barrier();
  const int r = b->vote;
  emu::block_barrier();
  if (b->cur == 0) b->vote = 0;  // thread 0 (fibers run one at a time; the next use is behind another barrier)
  emu::block_barrier();
  return r;
}
static inline int __syncthreads_and(int pred) { return !__syncthreads_or(!pred); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  uint64_t o[32];
  emu::warp_exchange(mask, 0, o);
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }
static inline void __nanosleep(unsigned) { emu::yield(); }

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 8, "shfl");
  uint64_t o[32], in = 0;
  memcpy(&in, &v, sizeof(T));
  emu::warp_exchange(mask, in, o);
  unsigned lane = threadIdx.x & 31;
  int s = (int)(lane & ~(unsigned)(width - 1)) + (src & (width - 1));
  T r;
  memcpy(&r, &o[s], sizeof(T));
  return r;
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  uint64_t o[32], in = 0;
  memcpy(&in, &v, sizeof(T));
  emu::warp_exchange(mask, in, o);
  int lane = (int)(threadIdx.x & 31);
  int s = lane - (int)delta;
  if (s < (lane & ~(width - 1))) s = lane;
  T r;
  memcpy(&r, &o[s], sizeof(T));
  return r;
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  uint64_t o[32], in = 0;
  memcpy(&in, &v, sizeof(T));
  emu::warp_exchange(mask, in, o);
  int lane = (int)(threadIdx.x & 31);
  int s = lane + (int)delta;
  if (s >= (lane & ~(width - 1)) + width) s = lane;
  T r;
  memcpy(&r, &o[s], sizeof(T));
  return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32) {
  uint64_t o[32], in = 0;
  memcpy(&in, &v, sizeof(T));
  emu::warp_exchange(mask, in, o);
  int lane = (int)(threadIdx.x & 31);
  (void)width;
  T r;
  memcpy(&r, &o[lane ^ lm], sizeof(T));
  return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint64_t o[32];
  emu::warp_exchange(mask, pred ? 1 : 0, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if (((mask >> i) & 1) && o[i]) r |= 1u << i;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <typename T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
  uint64_t o[32], in = 0;
  memcpy(&in, &v, sizeof(T));
  emu::warp_exchange(mask, in, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if (((mask >> i) & 1) && o[i] == in) r |= 1u << i;
  return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  emu::warp_exchange(mask, v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if ((mask >> i) & 1) r += (unsigned)o[i];
  return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  emu::warp_exchange(mask, v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if (((mask >> i) & 1) && (unsigned)o[i] > r) r = (unsigned)o[i];
  return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  emu::warp_exchange(mask, v, o);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; i++)
    if (((mask >> i) & 1) && (unsigned)o[i] < r) r = (unsigned)o[i];
  return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  emu::warp_exchange(mask, v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if ((mask >> i) & 1) r |= (unsigned)o[i];
  return r;
}

// atomics (relaxed is enough: ordering comes from the barriers)
template <typename T>
static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <typename T>
static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T>
static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <typename T>
static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <typename T>
static inline T atomicMax(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
template <typename T>
static inline T atomicMin(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
template <typename T>
static inline T atomicCAS(T *p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return cmp;
}

// integer intrinsics
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {  // selector nibbles 0..7 only (no sign replication)
  const unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(x);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (s & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  return (unsigned)((((((uint64_t)hi) << 32) | lo) << (s & 31)) >> 32);
}
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) {  // clamp variant: shift = min(s, 32)
  if (s > 32) s = 32;
  return (unsigned)((((((uint64_t)hi) << 32) | lo) << s) >> 32);
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) {
  if (s > 32) s = 32;
  return (unsigned)(((((uint64_t)hi) << 32) | lo) >> s);
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 255) * ((b >> (8 * i)) & 255);
  return c;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <typename T>
static inline T __ldg(const T *p) { return *p; }
template <typename T>
static inline T __ldcg(const T *p) { return *p; }
