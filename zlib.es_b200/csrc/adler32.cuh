// adler32.cuh — kernel K8: parallel segmented Adler-32.
//
// Replaces calcAdler32, /root/reference/src/adler32.ts:1-10 (byte-serial s1/s2
// with `% 65521` per byte).  Closed form used here, for input d[0..n):
//     s1 = (1 + sum d[i])            mod 65521
//     s2 = (n + sum (n - i) * d[i])  mod 65521
// so every 16-byte vector contributes independently and partial sums combine by
// addition (the same identity gives the segment combine used by the sharded
// deflate: a segment at offset o of length m with local sums (A, B = sum (m-j) d)
// contributes A to s1 and B + (n - o - m) * A to s2).
//
// Roofline: pure streaming read, algorithmic bytes = n (read once), no reuse;
// HBM-bound.  4 + 4 dp4a per 16 bytes keep the ALU cost under the load cost.
#pragma once
#include "zles_dev.h"

namespace zles {

constexpr int ADLER_THREADS = 256;
constexpr int ADLER_SMEM = 2 * (ADLER_THREADS / 32) * 8;

__device__ __forceinline__ void adler_vec16(const uint4 v, u64 w /* n - offset */, u64 &a, u64 &b) {
  u32 a16 = __dp4a(v.x, 0x01010101u, 0u);
  a16 = __dp4a(v.y, 0x01010101u, a16);
  a16 = __dp4a(v.z, 0x01010101u, a16);
  a16 = __dp4a(v.w, 0x01010101u, a16);
  u32 s16 = __dp4a(v.x, 0x03020100u, 0u);
  s16 = __dp4a(v.y, 0x07060504u, s16);
  s16 = __dp4a(v.z, 0x0b0a0908u, s16);
  s16 = __dp4a(v.w, 0x0f0e0d0cu, s16);
  a += a16;
  b += w * a16 - s16;
}

// acc[0] += sum d, acc[1] += (sum (n - i) d[i]) mod p   (both as u64 atomics)
// A CTA takes tiles of 32 KiB (grid-stride: neighbouring CTAs read neighbouring tiles); a thread has eight 16-byte loads
// of a tile in flight (256 threads x 128 bytes = the whole tile) and sums them in 32-bit arithmetic with byte indices
// local to the tile — sum (n - i0 - j) d[j] = (n - i0) sum d - sum j d, the 64-bit product once per tile, not per vector.
// (The first version did the 64-bit weight per vector with four loads in flight: 62 registers, four CTAs per SM, and
// 1.5 TB/s on inputs that do not fit the L2.)
constexpr int ADLER_VPT = 8;
constexpr u32 ADLER_TILE_VECS = ADLER_THREADS * ADLER_VPT;  // 2,048 vectors = 32 KiB
__global__ void __launch_bounds__(ADLER_THREADS, 4) k_adler_partial(const u8 *__restrict__ in, u64 n, unsigned long long *acc) {
  ZLES_SMEM_DECL(smem_raw);
  u64 *red = reinterpret_cast<u64 *>(smem_raw);  // [2][ADLER_THREADS / 32]
  u64 a = 0, b = 0;
  const u64 head = umin64((u64)((16 - ((uintptr_t)in & 15)) & 15), n);
  const u64 nvec = (n - head) >> 4;
  const uint4 *vp = reinterpret_cast<const uint4 *>(in + head);
  const u64 ntiles = nvec / ADLER_TILE_VECS;
  u32 since_mod = 0;
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint4 *tp = vp + tile * ADLER_TILE_VECS + threadIdx.x;
    uint4 v[ADLER_VPT];
#pragma unroll
    for (int k = 0; k < ADLER_VPT; k++) v[k] = __ldg(tp + k * ADLER_THREADS);
    u32 at = 0, st = 0;
#pragma unroll
    for (int k = 0; k < ADLER_VPT; k++) {
      u32 a16 = __dp4a(v[k].x, 0x01010101u, 0u);
      a16 = __dp4a(v[k].y, 0x01010101u, a16);
      a16 = __dp4a(v[k].z, 0x01010101u, a16);
      a16 = __dp4a(v[k].w, 0x01010101u, a16);
      u32 s16 = __dp4a(v[k].x, 0x03020100u, 0u);
      s16 = __dp4a(v[k].y, 0x07060504u, s16);
      s16 = __dp4a(v[k].z, 0x0b0a0908u, s16);
      s16 = __dp4a(v[k].w, 0x0f0e0d0cu, s16);
      at += a16;
      st += ((u32)(k * ADLER_THREADS + threadIdx.x) << 4) * a16 + s16;  // < 8 * 32,768 * 4,080: fits 32 bits
    }
    a += at;
    b += (n - head - ((tile * ADLER_TILE_VECS) << 4)) * (u64)at - st;
    if (++since_mod == 64) { b %= ADLER_MOD; since_mod = 0; }  // a tile adds less than n * 2^15
  }
  b %= ADLER_MOD;
  // the vectors behind the last whole tile, spread over the grid
  for (u64 i = ntiles * ADLER_TILE_VECS + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (u64)gridDim.x * blockDim.x)
    adler_vec16(__ldg(vp + i), n - head - (i << 4), a, b);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // unaligned head and the < 16-byte tail
    for (u64 k = 0; k < head; k++) { a += in[k]; b += (n - k) * in[k]; }
    for (u64 k = head + (nvec << 4); k < n; k++) { a += in[k]; b += (n - k) * in[k]; }
  }
  b %= ADLER_MOD;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_down_sync(ZLES_FULL, a, d);
    b += __shfl_down_sync(ZLES_FULL, b, d);
  }
  if (lane_id() == 0) { red[warp_id()] = a; red[ADLER_THREADS / 32 + warp_id()] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 ta = 0, tb = 0;
    for (int w = 0; w < ADLER_THREADS / 32; w++) { ta += red[w]; tb += red[ADLER_THREADS / 32 + w]; }
    atomicAdd(acc + 0, (unsigned long long)ta);
    atomicAdd(acc + 1, (unsigned long long)(tb % ADLER_MOD));
  }
}

__device__ __forceinline__ u32 adler_finish(u64 sum_a, u64 sum_b_modp, u64 n) {
  u32 s1 = (u32)((1 + sum_a) % ADLER_MOD);
  u32 s2 = (u32)((n % ADLER_MOD + sum_b_modp % ADLER_MOD) % ADLER_MOD);
  return (s2 << 16) | s1;
}

__global__ void k_adler_final(const unsigned long long *acc, u64 n, u32 *out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = adler_finish(acc[0], acc[1], n);
}

}  // namespace zles
