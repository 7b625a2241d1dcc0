// lz77.cuh — kernels K1/K2: per-block parallel match finder, greedy/lazy parse and
// token histogram.
//
// Replaces generateLZ77IndexMap + generateLZ77Codes (/root/reference/src/lz77.ts:11-119)
// and the histogram loop of deflateDynamicBlock (/root/reference/src/deflate.ts:58-77).
//
// The reference builds, per 128 KiB block, an exact map "3-byte key -> ascending
// list of positions" and, walking the block serially, checks up to 128 of the most
// recent earlier positions of the current key (most recent first, window 32768,
// longest wins, ties keep the nearest).  Here the same index is built in parallel:
//
//   one CTA (1024 threads) per 32 KiB deflate block ("SUB"), staged in shared memory
//   together with the preceding 32 KiB of the same chunk (the window), by TMA;
//   S2  the positions are radix-sorted (stable, 2 x 8 bit) by a 16-bit hash of their
//       3-byte key: one run of the sorted array == one position list of the reference;
//   S3  every position of the block looks up its own matches in parallel by walking
//       backwards through its run (most recent first, same stop rules as
//       src/lz77.ts:66-69,86-92, depth configurable);
//   S4  the greedy parse (src/lz77.ts:39-115, nowIndex += repeatLengthMax) is a serial
//       chain; 64 walkers parse 512-position ranges speculatively and a short serial
//       pass stitches them (a walk re-synchronises with the speculative one within a
//       few tokens), giving a bitmap of token starts;
//   S5  tokens are emitted and counted in parallel from the bitmap.
//
// Differences from the reference that only shrink the output (allowed by the
// "within 3 %" rule): matches may run to the end of the block (no Q2 tail discard,
// src/lz77.ts:95), length-3 matches further than 4096 back are dropped, and the
// parse is lazy (a match is deferred by one literal when the next position has a
// longer one).
//
// Shared memory: 64 KiB+pad data | 128 KiB sorted positions (u16) reused for the
// per-position match results (u32) | 16 KiB sort histograms reused for the token
// bitmap and the symbol histograms | 4 KiB warp queues | 1 KiB misc  = 213.4 KiB.
#pragma once
#include "tma.cuh"
#include "zles_dev.h"

namespace zles {

constexpr int LZ_THREADS = 1024;
constexpr int LZ_WARPS = LZ_THREADS / 32;
constexpr u32 LZ_PAD = 320;
constexpr u32 LZ_NWALK = 64;
constexpr u32 LZ_RANGE = SUB / LZ_NWALK;  // 512 positions per speculative walker
constexpr u32 LZ_HCOPIES = 8;
constexpr u32 LZ_NSYM = 320;  // [0,288) literal/length symbols, [288,320) distance symbols

constexpr u32 LZ_OFF_DATA = 0;
constexpr u32 LZ_OFF_X = 65536 + 384;                       // 65920
constexpr u32 LZ_OFF_WH = LZ_OFF_X + 131072;                // u16[LZ_WARPS*256] | bitmap u32[1024] + hist u32[8*320]
constexpr u32 LZ_OFF_WQ = LZ_OFF_WH + 16384;                // u16[LZ_WARPS][64]
constexpr u32 LZ_OFF_MISC = LZ_OFF_WQ + 4096;               // scratch u32[40] | specexit u32[64] | mbarrier
constexpr u32 LZ_SMEM = LZ_OFF_MISC + 1024;                 // 218496 B

struct LzParams {
  const u8 *in;       // this shard's input
  u64 n;              // its length
  u32 nblocks;        // ceil(n / SUB)
  u32 *tokens;        // [nblocks][SUB]
  u32 *ntok;          // [nblocks]
  u32 *hist;          // [nblocks][LZ_NSYM]
  u32 *scratch;       // [gridDim.x][SUB] u32: sort pass buffer (as u16[65536]), then match results
  u64 *adler_part;    // [nblocks][2]: sum d, sum (len - j) d[j] over the block's own bytes
  u32 max_checks;     // FAST_INDEX_CHECK_MAX   (reference: 128, src/lz77.ts:7)
  u32 min_checks;     // FAST_INDEX_CHECK_MIN   (reference: 16,  src/lz77.ts:8)
  u32 good_len;       // FAST_REPEAT_LENGTH     (reference: 8,   src/lz77.ts:9)
  u32 lazy;           // 1: defer a match by one literal when the next position has a longer one
  const BatchBlk *table = nullptr;  // batch mode: block b is table[b] (in/n describe one stream otherwise)
};

// token encoding shared with pack.cuh: literal = byte value; match = bit31 | (len-3)<<16 | (dist-1)
__device__ __forceinline__ u32 tok_match(u32 len, u32 dist) { return 0x80000000u | ((len - 3) << 16) | (dist - 1); }

__device__ __forceinline__ u32 lz_key3(const u8 *d, u32 p) { return (u32)d[p] | ((u32)d[p + 1] << 8) | ((u32)d[p + 2] << 16); }
__device__ __forceinline__ u32 lz_hash16(u32 key) { return (key * 0x9E3779B1u) >> 16; }

// unaligned 32-bit read from shared memory (two aligned loads + funnel shift)
__device__ __forceinline__ u32 lz_ld32(const u8 *d, u32 p) {
  const u32 *w = reinterpret_cast<const u32 *>(d + (p & ~3u));
  return __funnelshift_r(w[0], w[1], (p & 3) * 8);
}

// length of the common prefix of d[c..] and d[p..], c < p, capped at maxlen.
__device__ __forceinline__ u32 lz_match_len(const u8 *d, u32 c, u32 p, u32 maxlen) {
  u32 o = 0;
  while (o < maxlen) {
    u32 x = lz_ld32(d, c + o) ^ lz_ld32(d, p + o);
    if (x) { o += (u32)(__ffs((int)x) - 1) >> 3; break; }
    o += 4;
  }
  return umin(o, maxlen);
}

// One stable counting-sort pass over `N` items on the digit `shift` of their hash.
// Items of warp w are [w*per, (w+1)*per); pass 1 takes the item index as position,
// pass 2 reads positions from `src`.  Output order is stable (ascending source index).
template <typename OutT>
__device__ __forceinline__ void lz_sort_pass(const u8 *data, const u16 *src, OutT *dst, u16 *wh, u32 *scratch, u32 N, u32 per,
                                             u32 shift) {
  const u32 lane = lane_id(), w = warp_id();
  const u32 wbeg = umin(w * per, N), wend = umin(wbeg + per, N);
  for (u32 i = threadIdx.x; i < LZ_WARPS * 256 / 2; i += LZ_THREADS) reinterpret_cast<u32 *>(wh)[i] = 0;
  __syncthreads();
  for (u32 base = wbeg; base < wend; base += 32) {
    u32 idx = base + lane;
    bool valid = idx < wend;
    u32 digit = 256 + lane;
    if (valid) {
      u32 p = src ? (u32)src[idx] : idx;
      digit = (lz_hash16(lz_key3(data, p)) >> shift) & 255;
    }
    u32 m = __match_any_sync(ZLES_FULL, digit);
    if (valid && lane == (u32)(__ffs((int)m) - 1)) wh[w * 256 + digit] = (u16)(wh[w * 256 + digit] + __popc(m));
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive scan in digit-major, warp-minor order
    u32 v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      u32 e = threadIdx.x * 8 + k;
      v[k] = wh[(e % LZ_WARPS) * 256 + e / LZ_WARPS];
      s += v[k];
    }
    u32 total;
    u32 ex = block_exscan(s, scratch, &total);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      u32 e = threadIdx.x * 8 + k;
      wh[(e % LZ_WARPS) * 256 + e / LZ_WARPS] = (u16)ex;
      ex += v[k];
    }
  }
  __syncthreads();
  for (u32 base = wbeg; base < wend; base += 32) {
    u32 idx = base + lane;
    bool valid = idx < wend;
    u32 digit = 256 + lane, p = 0;
    if (valid) {
      p = src ? (u32)src[idx] : idx;
      digit = (lz_hash16(lz_key3(data, p)) >> shift) & 255;
    }
    u32 m = __match_any_sync(ZLES_FULL, digit);
    u32 basepos = valid ? wh[w * 256 + digit] : 0;
    __syncwarp();
    if (valid) {
      dst[basepos + __popc(m & lanemask_lt())] = (OutT)p;
      if (lane == (u32)(__ffs((int)m) - 1)) wh[w * 256 + digit] = (u16)(basepos + __popc(m));
    }
    __syncwarp();
  }
  __syncthreads();
}

// S3: best earlier match of position p = X[k] (one thread).  Returns (len << 16) | dist, or 0.
__device__ __forceinline__ u32 lz_find(const u8 *data, const u16 *X, u32 k, u32 L, const LzParams &P) {
  const u32 p = X[k];
  const u32 maxlen = umin(MAX_MATCH, L - p);
  if (maxlen < MIN_MATCH) return 0;
  const u32 kp = lz_key3(data, p);
  const u32 hp = lz_hash16(kp);
  u32 best = 2, bdist = 0, checks = 0, skips = 0;
  for (u32 j = k; j-- > 0;) {
    const u32 c = X[j];
    const u32 kc = lz_key3(data, c);
    if (kc != kp) {
      if (lz_hash16(kc) != hp || ++skips > 64) break;  // left the run of this hash
      continue;                                        // hash collision inside the run
    }
    if (p - c > WINDOW) break;  // runs are ascending: everything further left is older (src/lz77.ts:49)
    checks++;
    // src/lz77.ts:72-76: a candidate that cannot beat the best is rejected from its far end first
    if (best < maxlen && data[c + best] == data[p + best]) {
      u32 len = lz_match_len(data, c, p, maxlen);
      if (len > best) {  // strictly longer wins, ties keep the nearest (src/lz77.ts:86-92)
        best = len;
        bdist = p - c;
        if (len >= maxlen) break;
      }
    }
    if (checks >= P.max_checks || (best >= P.good_len && checks >= P.min_checks)) break;  // src/lz77.ts:66-69
  }
  if (best < MIN_MATCH) return 0;
  if (best == MIN_MATCH && bdist > 4096) return 0;  // costs more than three literals
  return (best << 16) | bdist;
}

// match length the parse uses at own-relative position pos (0 = literal)
__device__ __forceinline__ u32 lz_parse_len(const u32 *XR, u32 pos, u32 own_len, u32 lazy) {
  u32 len = XR[pos] >> 16;
  if (len && lazy && pos + 1 < own_len && (XR[pos + 1] >> 16) > len) return 0;
  return len;
}

__device__ __forceinline__ void lz_clear_bits(u32 *bm, u32 a, u32 b) {  // clears [a, b)
  while (a < b) {
    u32 w = a >> 5, lo = a & 31;
    u32 hi = umin(32u, lo + (b - a));
    u32 mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1)) & ~((1u << lo) - 1);
    bm[w] &= ~mask;
    a += hi - lo;
  }
}

__global__ void __launch_bounds__(LZ_THREADS, 1) k_lz(const LzParams P) {
  ZLES_SMEM_DECL(smem);
  u8 *data = smem + LZ_OFF_DATA;
  u16 *X = reinterpret_cast<u16 *>(smem + LZ_OFF_X);
  u32 *XR = reinterpret_cast<u32 *>(smem + LZ_OFF_X);
  u16 *wh = reinterpret_cast<u16 *>(smem + LZ_OFF_WH);
  u32 *bm = reinterpret_cast<u32 *>(smem + LZ_OFF_WH);          // [1024]
  u32 *hcopies = reinterpret_cast<u32 *>(smem + LZ_OFF_WH) + 1024;  // [LZ_HCOPIES][LZ_NSYM]
  u16 *wq = reinterpret_cast<u16 *>(smem + LZ_OFF_WQ);
  u32 *scratch = reinterpret_cast<u32 *>(smem + LZ_OFF_MISC);       // [40]
  u32 *specexit = scratch + 40;                                      // [64]
  u64 *mbar = reinterpret_cast<u64 *>(scratch + 40 + 64);           // 8-byte aligned: (40+64)*4 = 416
  u64 *red = reinterpret_cast<u64 *>(scratch + 40 + 64 + 2);        // [2][LZ_WARPS] u64

  const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
  u32 parity = 0;
#ifndef ZLES_EMU
  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();
#endif
  u16 *Y = reinterpret_cast<u16 *>(P.scratch + (size_t)blockIdx.x * SUB);
  u32 *R = P.scratch + (size_t)blockIdx.x * SUB;

  for (u32 b = blockIdx.x; b < P.nblocks; b += gridDim.x) {
    u64 own_off = (u64)b * SUB;
    u32 own_len, hist_len;
    if (P.table) {
      const BatchBlk t = P.table[b];
      own_off = t.in_off; own_len = t.own_len; hist_len = t.hist_len;
    } else {
      own_len = (u32)umin64((u64)SUB, P.n - own_off);
      hist_len = (b % SUBS_PER_CHUNK) ? SUB : 0;  // window = previous SUB of the same chunk
    }
    const u32 L = hist_len + own_len;

    // S0: stage window + block, zero the pad so word reads past the end are defined
    stage_g2s(data, P.in + own_off - hist_len, L, mbar, parity);
    for (u32 i = tid; i < LZ_PAD; i += LZ_THREADS) data[L + i] = 0;
    __syncthreads();

    // S1: Adler-32 partial sums of the block's own bytes (K8 fused into the load)
    {
      u64 a = 0, bsum = 0;
      const u32 j0 = tid * 32;
      if (j0 < own_len) {
        const u32 cnt = umin(32u, own_len - j0);
        u32 sa = 0, sb = 0;
        for (u32 j = 0; j < cnt; j++) {
          u32 d = data[hist_len + j0 + j];
          sa += d;
          sb += (own_len - j0 - j) * d;
        }
        a = sa; bsum = sb;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(ZLES_FULL, a, d);
        bsum += __shfl_down_sync(ZLES_FULL, bsum, d);
      }
      if (lane == 0) { red[w] = a; red[LZ_WARPS + w] = bsum; }
      __syncthreads();
      if (tid == 0) {
        u64 ta = 0, tb = 0;
        for (int i = 0; i < LZ_WARPS; i++) { ta += red[i]; tb += red[LZ_WARPS + i]; }
        P.adler_part[2 * (size_t)b] = ta;
        P.adler_part[2 * (size_t)b + 1] = tb;
      }
    }

    // S2: stable radix sort of positions by hash16(key3)
    const u32 N = L >= 3 ? L - 2 : 0;
    const u32 per = ((N + LZ_THREADS - 1) / LZ_THREADS) * 32;
    lz_sort_pass<u16>(data, nullptr, Y, wh, scratch, N, per, 0);
    lz_sort_pass<u16>(data, Y, X, wh, scratch, N, per, 8);

    // S3: per-position match search, own positions only, compacted through a warp queue
    {
      u16 *q = wq + w * 64;
      u32 qn = 0;
      for (u32 base = w * 32; base < N; base += LZ_THREADS) {
        const u32 k = base + lane;
        const bool own = k < N && X[k] >= hist_len;
        const u32 bal = __ballot_sync(ZLES_FULL, own);
        if (own) q[qn + __popc(bal & lanemask_lt())] = (u16)k;
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 32) {
          const u32 kk = q[lane];
          R[(u32)X[kk] - hist_len] = lz_find(data, X, kk, L, P);
          __syncwarp();
          u16 t = 0;
          if (lane < qn - 32) t = q[32 + lane];
          __syncwarp();
          if (lane < qn - 32) q[lane] = t;
          qn -= 32;
          __syncwarp();
        }
      }
      if (lane < qn) {
        const u32 kk = q[lane];
        R[(u32)X[kk] - hist_len] = lz_find(data, X, kk, L, P);
      }
      // positions without a full 3-byte key (the last two of the window+block) have no match
      if (tid < 2 && own_len > tid) R[own_len - 1 - tid] = 0;
    }
    __syncthreads();

    // S4: match results into shared memory (over the sorted array), then the parse
    for (u32 i = tid; i < own_len; i += LZ_THREADS) XR[i] = R[i];
    bm[tid] = 0;
    for (u32 i = tid; i < LZ_HCOPIES * LZ_NSYM; i += LZ_THREADS) hcopies[i] = 0;
    __syncthreads();
    if (tid < LZ_NWALK) {
      const u32 s = tid * LZ_RANGE;
      if (s < own_len) {
        const u32 e = umin(s + LZ_RANGE, own_len);
        u32 pos = s;
        while (pos < e) {
          bm[pos >> 5] |= 1u << (pos & 31);
          u32 len = lz_parse_len(XR, pos, own_len, P.lazy);
          pos += len ? len : 1;
        }
        specexit[tid] = pos;
      }
    }
    __syncthreads();
    if (tid == 0) {  // stitch: the true parse enters range t where range t-1 really left off
      u32 entry = 0;
      for (u32 t = 0; t < LZ_NWALK; t++) {
        const u32 s = t * LZ_RANGE;
        if (s >= own_len) break;
        const u32 e = umin(s + LZ_RANGE, own_len);
        if (entry == s) { entry = specexit[t]; continue; }
        lz_clear_bits(bm, s, umin(entry, e));
        if (entry >= e) continue;
        u32 pos = entry;
        for (;;) {
          if (pos >= e) { entry = pos; break; }
          if ((bm[pos >> 5] >> (pos & 31)) & 1) { entry = specexit[t]; break; }  // re-synchronised
          bm[pos >> 5] |= 1u << (pos & 31);
          u32 len = lz_parse_len(XR, pos, own_len, P.lazy);
          u32 nxt = pos + (len ? len : 1);
          lz_clear_bits(bm, pos + 1, umin(nxt, e));
          pos = nxt;
        }
      }
    }
    __syncthreads();

    // S5: emit tokens and count symbols
    {
      u32 word = bm[tid];
      u32 total;
      u32 o = block_exscan((u32)__popc(word), scratch, &total);
      u32 *tok = P.tokens + (size_t)b * SUB;
      u32 *hc = hcopies + (w % LZ_HCOPIES) * LZ_NSYM;
      while (word) {
        const u32 bit = (u32)(__ffs((int)word) - 1);
        word &= word - 1;
        const u32 pos = tid * 32 + bit;
        const u32 len = lz_parse_len(XR, pos, own_len, P.lazy);
        if (len) {
          const u32 dist = XR[pos] & 0xffff;
          u32 ls, le, lv, ds, de, dv;
          len_to_sym(len, ls, le, lv);
          dist_to_sym(dist, ds, de, dv);
          atomicAdd(hc + 257 + ls, 1u);
          atomicAdd(hc + 288 + ds, 1u);
          tok[o] = tok_match(len, dist);
        } else {
          const u32 d = data[hist_len + pos];
          atomicAdd(hc + d, 1u);
          tok[o] = d;
        }
        o++;
      }
      __syncthreads();
      if (tid < LZ_NSYM) {
        u32 s = 0;
        for (u32 c = 0; c < LZ_HCOPIES; c++) s += hcopies[c * LZ_NSYM + tid];
        P.hist[(size_t)b * LZ_NSYM + tid] = s;
      }
      if (tid == 0) P.ntok[b] = total;
    }
    __syncthreads();
  }
}

}  // namespace zles
