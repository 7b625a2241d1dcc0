// zles_dev.h — device-side portability shim and small helpers.
//
// Product build: nvcc, sm_100a.  Test build: g++ -DZLES_EMU against
// tests/emu/cuda_emu.h (a CPU thread emulator used only by the test-suite to
// chase logic errors without a GPU; the product never loads that build).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef ZLES_EMU
#include "cuda_emu.h"
#define ZLES_SMEM_DECL(name) uint8_t *name = emu::dyn_smem()
#define ZLES_CONSTANT static const
#else
#include <cuda_runtime.h>
#define ZLES_SMEM_DECL(name) extern __shared__ __align__(1024) uint8_t name[]
#define ZLES_CONSTANT __constant__
#endif

#define ZLES_FULL 0xffffffffu

namespace zles {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// Format constants.  CHUNK is the reference's BLOCK_MAX_BUFFER_LEN
// (/root/reference/src/const.ts:7): every CHUNK of input is compressed with no
// reference to earlier chunks (/root/reference/src/lz77.ts:11-22,37).  SUB is the
// deflate block our encoder emits; the SUBs of a chunk may reference back
// inside the chunk up to the 32 KiB window (/root/reference/src/lz77.ts:49).
constexpr u32 CHUNK_LOG2 = 17;
constexpr u32 CHUNK = 1u << CHUNK_LOG2;
constexpr u32 SUB_LOG2 = 15;
constexpr u32 SUB = 1u << SUB_LOG2;
constexpr u32 SUBS_PER_CHUNK = CHUNK / SUB;
constexpr u32 WINDOW = 32768;
constexpr u32 MAX_MATCH = 258;
constexpr u32 MIN_MATCH = 3;
constexpr u32 ADLER_MOD = 65521;

// Symbol tables, /root/reference/src/const.ts:9-35 (RFC 1951 §3.2.5/§3.2.7).
ZLES_CONSTANT u16 c_len_base[32] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27, 31,
                                    35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0};
ZLES_CONSTANT u8 c_len_extra[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                    3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};
ZLES_CONSTANT u16 c_dist_base[32] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,   97,   129, 193,
                                     257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
ZLES_CONSTANT u8 c_dist_extra[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                     7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0};
ZLES_CONSTANT u8 c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__host__ __device__ __forceinline__ u32 umin(u32 a, u32 b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ u32 umax(u32 a, u32 b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ u64 umin64(u64 a, u64 b) { return a < b ? a : b; }

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ u32 lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1; }

// length (3..258) -> (symbol 0..28, extra-bit count, extra value); same mapping as the
// "last base <= value" search of /root/reference/src/lz77.ts:97-102.
__device__ __forceinline__ void len_to_sym(u32 len, u32 &sym, u32 &ebits, u32 &eval) {
  u32 l = len - 3;
  if (len == 258) { sym = 28; ebits = 0; eval = 0; return; }
  if (l < 8) { sym = l; ebits = 0; eval = 0; return; }
  u32 k = 31 - __clz((int)l);
  ebits = k - 2;
  sym = 4 * ebits + 4 + ((l >> ebits) & 3);
  eval = l & ((1u << ebits) - 1);
}
// distance (1..32768) -> (symbol 0..29, extra-bit count, extra value); /root/reference/src/lz77.ts:103-108.
__device__ __forceinline__ void dist_to_sym(u32 dist, u32 &sym, u32 &ebits, u32 &eval) {
  u32 d = dist - 1;
  if (d < 4) { sym = d; ebits = 0; eval = 0; return; }
  u32 k = 31 - __clz((int)d);
  ebits = k - 1;
  sym = 2 * k + ((d >> (k - 1)) & 1);
  eval = d & ((1u << ebits) - 1);
}

// One deflate block of a batch of independent buffers (zles_*_deflate_batch): where its own
// bytes start, how many there are, and how many bytes of window precede them.
struct BatchBlk {
  u64 in_off;
  u32 own_len;
  u32 hist_len;
};

// Block-wide exclusive scan of one u32 per thread (blockDim.x multiple of 32, <= 1024).
// scratch: 33 u32 in shared memory.  Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ u32 block_exscan(u32 v, u32 *scratch, u32 *total) {
  u32 lane = lane_id(), w = warp_id(), nw = (blockDim.x + 31) >> 5;
  u32 inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_up_sync(ZLES_FULL, inc, d);
    if (lane >= (u32)d) inc += t;
  }
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 31) scratch[w] = inc;
  __syncthreads();
  if (w == 0) {
    u32 s = lane < nw ? scratch[lane] : 0;
    u32 si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 t = __shfl_up_sync(ZLES_FULL, si, d);
      if (lane >= (u32)d) si += t;
    }
    if (lane < nw) scratch[lane] = si - s;
    if (lane == 31) scratch[32] = si;
  }
  __syncthreads();
  u32 r = scratch[w] + inc - v;
  *total = scratch[32];
  return r;
}

}  // namespace zles
