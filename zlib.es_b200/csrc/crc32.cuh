// crc32.cuh — parallel, shard-combinable CRC-32 (IEEE 802.3, reflected, polynomial 0xEDB88320): the checksum of the gzip
// container (RFC 1952), the wire-format sibling of the zlib container the reference writes (the
// reference's own container check is Adler-32, /root/reference/src/adler32.ts — adler32.cuh).
//
// CRC is linear over GF(2).  With raw(M) = the CRC register after M starting from 0, without the final inversion:
//     raw(A || B) = raw(A) * x^(8 |B|)  +  raw(B)                     (mod P; + is XOR)
//     crc32(M)    = raw(M) + 0xFFFFFFFF * x^(8 |M|) + 0xFFFFFFFF      (the usual pre- and post-inversion)
// so every thread takes its own 128 bytes (slice-by-8 tables in shared memory), lanes, warps and CTAs combine their
// values in trees whose multipliers are constants (x^(8 * 128 * 2^k)), and a second kernel folds the per-CTA values.
// The same identity combines the CRCs of shards on the host (zles_crc32_combine), like the Adler-32 sums.
//
// Roofline: a streaming read, algorithmic bytes = n; HBM-bound in principle, in practice bound by the eight table
// look-ups per 8 bytes (shared-memory wavefronts with random bank conflicts).
#pragma once
#include "zles_dev.h"

namespace zles {

constexpr u32 CRC_POLY = 0xEDB88320u;
constexpr int CRC_THREADS = 256;
constexpr u32 CRC_PER_THREAD = 128;                              // bytes
constexpr u32 CRC_PER_BLOCK = CRC_THREADS * CRC_PER_THREAD;      // 32 KiB
constexpr int CRC_SMEM = 8 * 256 * 4 + CRC_THREADS / 32 * 4;
constexpr int CRC_FINAL_SMEM = 1024 * 12;

struct CrcTables {
  u32 t[8][256];   // slice-by-8
  u32 xp[40];      // xp[k] = x^(8 * CRC_PER_THREAD * 2^k) mod P: combine multipliers of the trees
};

// a * b mod P in the reflected representation (bit 31 is the coefficient of x^0)
__host__ __device__ __forceinline__ u32 crc_mul(u32 a, u32 b) {
  u32 p = 0;
#pragma unroll 4
  for (int i = 0; i < 32; i++) {
    if (a & (0x80000000u >> i)) p ^= b;
    b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
  }
  return p;
}
// x^(8 n) mod P
__host__ __device__ inline u32 crc_xpow8n(u64 n) {
  u32 p = 0x80000000u, sq = 0x00800000u;  // x^0, x^8
  while (n) {
    if (n & 1) p = crc_mul(sq, p);
    sq = crc_mul(sq, sq);
    n >>= 1;
  }
  return p;
}
// crc32(A || B) from crc32(A), crc32(B) and |B| (holds for the raw and for the conditioned values alike)
__host__ __device__ inline u32 crc_combine(u32 crc_a, u32 crc_b, u64 len_b) { return crc_mul(crc_a, crc_xpow8n(len_b)) ^ crc_b; }

inline void crc_tables_build(CrcTables *T) {
  for (u32 i = 0; i < 256; i++) {
    u32 c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
    T->t[0][i] = c;
  }
  for (u32 i = 0; i < 256; i++)
    for (int s = 1; s < 8; s++) T->t[s][i] = T->t[0][T->t[s - 1][i] & 255] ^ (T->t[s - 1][i] >> 8);
  u32 v = crc_xpow8n(CRC_PER_THREAD);
  for (int k = 0; k < 40; k++) { T->xp[k] = v; v = crc_mul(v, v); }
}

// raw CRC of 8 more bytes (lo = the first four, little endian)
__device__ __forceinline__ u32 crc_step8(const u32 (*t)[256], u32 crc, u32 lo, u32 hi) {
  lo ^= crc;
  return t[7][lo & 255] ^ t[6][(lo >> 8) & 255] ^ t[5][(lo >> 16) & 255] ^ t[4][lo >> 24] ^
         t[3][hi & 255] ^ t[2][(hi >> 8) & 255] ^ t[1][(hi >> 16) & 255] ^ t[0][hi >> 24];
}

// The message is bytes [skew, skew + n) of `base` (16-byte aligned; skew < 16): leading zero bytes do not change a raw
// CRC, so the bytes before the message are masked to zero and the pieces are cut in this aligned frame.
// part[b] = raw CRC of frame bytes [b * 32 KiB, (b + 1) * 32 KiB) for every piece that lies entirely below skew + n.
__global__ void __launch_bounds__(CRC_THREADS) k_crc_partial(const u8 *__restrict__ base, u32 skew, u64 nfull, const CrcTables *__restrict__ T,
                                                             u32 *part) {
  ZLES_SMEM_DECL(smem_raw);
  u32(*t)[256] = reinterpret_cast<u32(*)[256]>(smem_raw);
  u32 *wred = reinterpret_cast<u32 *>(smem_raw) + 8 * 256;
  for (u32 i = threadIdx.x; i < 8 * 256; i += CRC_THREADS) (&t[0][0])[i] = (&T->t[0][0])[i];
  __syncthreads();
  const u32 lane = lane_id(), w = warp_id();
  for (u64 b = blockIdx.x; b < nfull; b += gridDim.x) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base + b * CRC_PER_BLOCK + (u64)threadIdx.x * CRC_PER_THREAD);
    uint4 v[CRC_PER_THREAD / 16];
#pragma unroll
    for (u32 k = 0; k < CRC_PER_THREAD / 16; k++) v[k] = __ldg(p + k);
    if (b == 0 && threadIdx.x == 0 && skew) {  // the bytes in front of the message read as zero
      u32 wv[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
      for (u32 q = 0; q < 16; q++)
        if (q < skew) wv[q >> 2] &= ~(0xffu << ((q & 3) * 8));
      v[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    u32 c = 0;
#pragma unroll
    for (u32 k = 0; k < CRC_PER_THREAD / 16; k++) {
      c = crc_step8(t, c, v[k].x, v[k].y);
      c = crc_step8(t, c, v[k].z, v[k].w);
    }
    // lanes: the lower neighbour's bytes come first, so its value is shifted over the upper one's length
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const u32 o = __shfl_xor_sync(ZLES_FULL, c, 1 << k);
      const bool upper = (lane >> k) & 1;
      const u32 left = upper ? o : c, right = upper ? c : o;
      c = crc_mul(left, T->xp[k]) ^ right;  // both partners compute the pair's value
    }
    if (lane == 0) wred[w] = c;
    __syncthreads();
    if (w == 0) {
      u32 x = lane < CRC_THREADS / 32 ? wred[lane] : 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {  // 8 warps
        const u32 o = __shfl_xor_sync(ZLES_FULL, x, 1 << k);
        const bool upper = (lane >> k) & 1;
        const u32 left = upper ? o : x, right = upper ? x : o;
        x = crc_mul(left, T->xp[5 + k]) ^ right;
      }
      if (lane == 0) part[b] = x;
    }
    __syncthreads();
  }
}

// tree over 1024 (value, byte length) pairs in thread order; the result ends up in red[0] / len[0]
__device__ __forceinline__ void crc_tree(u32 *red, u64 *len) {
  const u32 tid = threadIdx.x;
  for (u32 s = 1; s < 1024; s <<= 1) {
    if ((tid & (2 * s - 1)) == 0) {
      const u64 rn = len[tid + s];
      if (rn) red[tid] = crc_mul(red[tid], crc_xpow8n(rn)) ^ red[tid + s];
      len[tid] += rn;
    }
    __syncthreads();
  }
}

// One CTA: folds the per-piece values in order, adds the tail (the frame bytes after the last full piece, < 32 KiB) and
// applies the conditioning.  out[0] = crc32 of the message.
__global__ void __launch_bounds__(1024) k_crc_final(const u8 *__restrict__ base, u32 skew, u64 n, u64 nfull, const u32 *__restrict__ part,
                                                    const CrcTables *__restrict__ T, u32 *out) {
  ZLES_SMEM_DECL(smem_raw);
  u64 *len = reinterpret_cast<u64 *>(smem_raw);            // [1024]
  u32 *red = reinterpret_cast<u32 *>(smem_raw + 8 * 1024);  // [1024]
  const u32 tid = threadIdx.x;
  // thread t folds pieces [a, b) sequentially: value * x^(8 * 32 KiB) + next
  const u64 per = (nfull + 1023) / 1024;
  const u64 a = umin64((u64)tid * per, nfull), b = umin64(a + per, nfull);
  const u32 xblk = T->xp[8];  // x^(8 * 128 * 256)
  u32 c = 0;
  for (u64 i = a; i < b; i++) c = crc_mul(c, xblk) ^ part[i];
  red[tid] = c;
  len[tid] = (b - a) * CRC_PER_BLOCK;
  __syncthreads();
  crc_tree(red, len);
  const u32 head = red[0];
  __syncthreads();
  // the tail: 32 frame bytes per thread, byte-wise
  const u64 t0 = nfull * CRC_PER_BLOCK, end = (u64)skew + n;
  const u64 ta = umin64(t0 + (u64)tid * 32, end), tb = umin64(ta + 32, end);
  c = 0;
  for (u64 i = ta; i < tb; i++) {
    const u32 v = i < skew ? 0u : (u32)base[i];
    c = T->t[0][(c ^ v) & 255] ^ (c >> 8);
  }
  red[tid] = c;
  len[tid] = tb - ta;
  __syncthreads();
  crc_tree(red, len);
  if (tid == 0) {
    const u32 r = crc_mul(head, crc_xpow8n(len[0])) ^ red[0];
    *out = r ^ crc_mul(0xFFFFFFFFu, crc_xpow8n(n)) ^ 0xFFFFFFFFu;
  }
}

}  // namespace zles
