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
//   S3  every position of the block finds its matches in parallel.  A warp sweeps its slice of the
//       sorted array 32 entries at a time; each entry's next 7 bytes go into a per-warp ring, and
//       every lane compares its own 7 bytes with those of the <= 32 entries before it in its run
//       (nearest first, inside the window — the candidate order of src/lz77.ts:64-93).  That
//       gives the exact match length up to 6 for all 32 candidates at ~13 instructions each with
//       no divergence; only candidates that match on all 7 bytes are extended byte-wise
//       (up to `deep` of them).  Longest wins, ties keep the nearest (src/lz77.ts:86-92);
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
// per-position match results (u32) | 16 KiB sort histograms reused for the candidate
// rings, then the token bitmap and the symbol histograms | 1 KiB misc  = 209.4 KiB.
#pragma once
#include "tma.cuh"
#include "zles_dev.h"

namespace zles {

constexpr int LZ_THREADS = 1024;
constexpr int LZ_WARPS = LZ_THREADS / 32;
constexpr u32 LZ_PAD = 320;
constexpr u32 LZ_NWALK = 256;
constexpr u32 LZ_RANGE = SUB / LZ_NWALK;  // 128 positions per speculative walker
constexpr u32 LZ_HCOPIES = 8;
constexpr u32 LZ_NSYM = 320;  // [0,288) literal/length symbols, [288,320) distance symbols

constexpr u32 LZ_OFF_DATA = 0;
constexpr u32 LZ_OFF_X = 65536 + 384;                       // 65920
constexpr u32 LZ_OFF_WH = LZ_OFF_X + 131072;                // u16[LZ_WARPS*256] | bitmap u32[1024] + hist u32[8*320]
constexpr u32 LZ_OFF_MISC = LZ_OFF_WH + 32768;              // scratch u32[40] | specexit u32[64] | mbarrier
constexpr u32 LZ_SMEM = LZ_OFF_MISC + 2048;                 // 231808 B (of 232448 available)
constexpr u32 LZ_SCAN = 32;                                 // candidates compared per position (at most)
constexpr u32 LZ_BGROUP = 16;                               // batch mode: blocks per group (one sort when all are small)
constexpr u32 LZ_BSLOT = (2 * SUB) / LZ_BGROUP;              // 4096: the most a block of a packed group holds
constexpr u32 LZ_REDO_SLACK = 64;                            // pair mode: block 2 is matched again with its window when it took more than
                                                            // 4/3 of block 3's tokens plus this many (see the end of a unit)
constexpr u32 LZ_SLICE = 256;                               // sorted entries per dynamically scheduled slice (multiple of 32)
// per-warp candidate ring: 64 entries of 4 bytes (lz_tag), stored twice (slot i and i + 64) so that "entry k - r"
// is a constant offset from a per-lane base and needs no wrap-around arithmetic
static_assert(LZ_WARPS * 128 * 4 <= 32768, "candidate rings overlay the sort histograms");

struct LzParams {
  const u8 *in;       // this shard's input
  u64 n;              // its length
  u32 first_block = 0; // this launch covers blocks [first_block, nblocks)
  u32 nblocks;        // ceil(n / SUB), or the end of the slab
  u32 *tokens;        // [nblocks][SUB]
  u32 *ntok;          // [nblocks]
  u32 *hist;          // [nblocks][LZ_NSYM]
  u32 *scratch;       // [gridDim.x][2 * SUB] u32: sort pass buffer, then match results
  u64 *adler_part;    // [nblocks][2]: sum d, sum (len - j) d[j] over the block's own bytes
  u32 max_checks;     // candidates compared per position, <= LZ_SCAN   (reference: FAST_INDEX_CHECK_MAX = 128, src/lz77.ts:7)
  u32 min_checks;     // reserved (reference: FAST_INDEX_CHECK_MIN = 16, src/lz77.ts:8)
  u32 good_len;       // reserved (reference: FAST_REPEAT_LENGTH = 8, src/lz77.ts:9)
  u32 lazy;           // 1: defer a match by one literal when the next position has a longer one
  u32 *unit_ctr = nullptr;  // zeroed before the launch: units are handed out from it, the two-block ones first
  u32 pair_mode = 1;  // 1: two sorts per chunk — blocks {0,1} and {2,3}: block 2 has no window (the default);
                      // 0: three — {0,1}, then {2} and {3} each with the block before as window (smaller output, slower)
  const BatchBlk *table = nullptr;  // batch mode: block b is table[b] (in/n describe one stream otherwise)
  u32 tok_stride = SUB;  // token slots per block: SUB, or less when no block of a batch is that long (a batch of 262,144
                         // buffers of 4 KiB then takes 4 GiB of token scratch instead of 32)
};

// Per-stage cycle counters (tools/lz_stages.py builds with -DZLES_STAGE_CLOCKS; off in the product build):
// thread 0 adds the cycles since the previous mark to g_lz_clk[i].  The previous mark lives in the last
// 8 bytes of the misc area of shared memory.
#ifdef ZLES_STAGE_CLOCKS
__device__ unsigned long long g_lz_clk[16];
#define LZ_CLK(scr, i)                                                                   \
  do {                                                                                   \
    if (threadIdx.x == 0) {                                                              \
      long long *cp_ = reinterpret_cast<long long *>((scr) + 510);                       \
      const long long t_ = clock64();                                                    \
      if ((i) >= 0) atomicAdd(&g_lz_clk[(i) < 0 ? 0 : (i)], (unsigned long long)(t_ - *cp_)); \
      *cp_ = t_;                                                                         \
    }                                                                                    \
  } while (0)
// inside S3: warp 0 accumulates its own cycles per sub-stage (fill + run lengths, scan, finalize)
#define LZ_WCLK(i)                                                        \
  do {                                                                    \
    if (threadIdx.x == 0) {                                               \
      const long long t_ = clock64();                                     \
      atomicAdd(&g_lz_clk[i], (unsigned long long)(t_ - wclk));           \
      wclk = t_;                                                          \
    }                                                                     \
  } while (0)
#define LZ_WCLK_DECL long long wclk = clock64()
#else
#define LZ_CLK(scr, i) do { } while (0)
#define LZ_WCLK(i) do { } while (0)
#define LZ_WCLK_DECL do { } while (0)
#endif

// token encoding shared with pack.cuh: literal = byte value; match = bit31 | (len-3)<<16 | (dist-1)
__device__ __forceinline__ u32 tok_match(u32 len, u32 dist) { return 0x80000000u | ((len - 3) << 16) | (dist - 1); }

__device__ __forceinline__ u32 lz_key3(const u8 *d, u32 p) { return (u32)d[p] | ((u32)d[p + 1] << 8) | ((u32)d[p + 2] << 16); }
// lz_mix(w), w = the four bytes at a position: (w * M) with M = C << 8, C odd, depends on the 3-byte key only (the
// fourth byte is multiplied out of the 32 bits) and is a bijection of the key onto bits 31..8.  Bits 31..16 are the
// sort hash, bits 15..8 the "key fold": two positions with equal hash AND equal fold have the same 3-byte key.
constexpr u32 LZ_HMUL = 0x3779B100u;
__device__ __forceinline__ u32 lz_mix(u32 w) { return w * LZ_HMUL; }
__device__ __forceinline__ u32 lz_hash16(u32 w) { return lz_mix(w) >> 16; }

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

// ---- S2: stable LSD radix sort of the positions by hash16(key3), 2 passes of 8 bits ----------
// Every warp owns a contiguous range of items in each pass; per-warp digit histograms (u16, in `wh`)
// make the scatter stable.  Counting uses fire-and-forget shared-memory atomics on packed u16
// pairs (no dependency chain); the scatter ranks equal digits inside a warp (lz_peers).
// Pass 1 reads the data, pass 2 reads pass 1's output Y (u32 = hash << 16 | position, in global
// memory / L2, coalesced, four loads in flight per lane).

__device__ __forceinline__ u32 lz_key3_fast(const u8 *d, u32 p) {  // the key's bytes + one more (lz_mix ignores it)
  return lz_ld32(d, p);
}

__device__ __forceinline__ void lz_hist_scan(u16 *wh, u32 *scratch) {  // exclusive scan, digit-major / warp-minor
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

__device__ __forceinline__ void lz_count(u16 *wh, u32 w, u32 digit) {
  atomicAdd(reinterpret_cast<u32 *>(wh) + ((w * 256 + digit) >> 1), 1u << ((digit & 1) * 16));
}

// lanes of the warp that hold the same 8-bit digit (valid lanes only).  BALLOT: eight ballots, a fixed ~40
// instructions; otherwise __match_any_sync, whose cost grows with the number of distinct digits in the warp
// (measured: ~7 cycles of the scheduler per distinct value).  Pass 1 sees ~28 distinct digits per warp on text
// (consecutive positions), pass 2 only a few (its input is grouped by the first digit).
template <bool BALLOT>
__device__ __forceinline__ u32 lz_peers(u32 digit, bool valid) {
  if (BALLOT) {
    u32 peers = __ballot_sync(ZLES_FULL, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const bool bit = (digit >> b) & 1;
      const u32 bal = __ballot_sync(ZLES_FULL, bit);
      peers &= bit ? bal : ~bal;
    }
    return peers;
  }
  return __match_any_sync(ZLES_FULL, valid ? digit : 256 + lane_id());
}

// rank of this lane among the lanes of the warp with the same digit, and the slot it scatters to
template <bool BALLOT>
__device__ __forceinline__ u32 lz_slot(u16 *wh, u32 w, u32 digit, bool valid) {
  const u32 lane = lane_id();
  const u32 m = lz_peers<BALLOT>(digit, valid);
  const u32 basepos = valid ? wh[w * 256 + digit] : 0;
  __syncwarp();
  if (valid && lane == (u32)(__ffs((int)m) - 1)) wh[w * 256 + digit] = (u16)(basepos + __popc(m));
  __syncwarp();
  return basepos + __popc(m & lanemask_lt());
}

// Warp w owns items [w << lper, (w + 1) << lper) of each pass.  `wh` = per-warp histograms of pass 1,
// `wh + LZ_WARPS * 256` = those of pass 2, which are counted while pass 1 scatters (the item that lands in
// slot s belongs to warp s >> lper in pass 2), so pass 2 reads Y only once.
__device__ __forceinline__ void lz_sort(const u8 *data, u32 *Y, u16 *X, u16 *wh, u32 *scratch, u32 N, u32 lper) {
  const u32 lane = lane_id(), w = warp_id();
  const u32 wbeg = umin(w << lper, N), wend = umin(wbeg + (1u << lper), N);
  u16 *wh2 = wh + LZ_WARPS * 256;
  u32 *wh32 = reinterpret_cast<u32 *>(wh);
  // ---- pass 1: low digit ----
  for (u32 i = threadIdx.x; i < LZ_WARPS * 256; i += LZ_THREADS) wh32[i] = 0;  // both histograms
  __syncthreads();
  for (u32 idx = wbeg + lane; idx < wend; idx += 32) lz_count(wh, w, lz_hash16(lz_key3_fast(data, idx)) & 255);
  __syncthreads();
  LZ_CLK(scratch, 2);
  lz_hist_scan(wh, scratch);
  __syncthreads();
  LZ_CLK(scratch, 3);
  for (u32 base = wbeg; base < wend; base += 32) {
    const u32 idx = base + lane;
    const bool valid = idx < wend;
    const u32 h = valid ? lz_hash16(lz_key3_fast(data, idx)) : 0;
    const u32 slot = lz_slot<true>(wh, w, h & 255, valid);
    if (valid) {
      Y[slot] = (h << 16) | idx;
      lz_count(wh2, slot >> lper, h >> 8);
    }
  }
  __syncthreads();  // also orders the global writes of Y before the reads below (same CTA)
  LZ_CLK(scratch, 4);
  // ---- pass 2: high digit ----
  lz_hist_scan(wh2, scratch);
  __syncthreads();
  LZ_CLK(scratch, 6);
  for (u32 base = wbeg; base < wend; base += 128) {
    u32 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const u32 idx = base + 32 * k + lane; v[k] = idx < wend ? __ldcg(Y + idx) : 0; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (base + 32 * k >= wend) break;  // warp-uniform
      const bool valid = base + 32 * k + lane < wend;
      const u32 slot = lz_slot<false>(wh2, w, v[k] >> 24, valid);
      if (valid) X[slot] = (u16)v[k];
    }
  }
  __syncthreads();
  LZ_CLK(scratch, 7);
}

// bytes d[p .. p+6] as (lo = bytes 0-3, hi = bytes 4-6); unaligned shared-memory read
__device__ __forceinline__ void lz_ld56(const u8 *d, u32 p, u32 &lo, u32 &hi) {
  const u32 *w = reinterpret_cast<const u32 *>(d + (p & ~3u));
  const u32 sh = (p & 3) * 8;
  const u32 w0 = w[0], w1 = w[1], w2 = w[2];
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh) & 0x00ffffffu;
}

// bytes d[p .. p+7] as one u64; unaligned shared-memory read (3 aligned loads)
__device__ __forceinline__ u64 lz_ld64(const u8 *d, u32 p) {
  const u32 *w = reinterpret_cast<const u32 *>(d + (p & ~3u));
  const u32 sh = (p & 3) * 8;
  const u32 w0 = w[0], w1 = w[1], w2 = w[2];
  return ((u64)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}

// ring entry of a position: [key fold, byte 3, byte 4, byte 5] (q = lz_mix(lo); lo, hi from lz_ld56)
__device__ __forceinline__ u32 lz_tag(u32 q, u32 lo, u32 hi) {
  return __byte_perm(__byte_perm(q, lo, 0x0071), hi, 0x5410);
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

// One candidate of the S3 scan (see there): uses es, rp, rrun, one; updates best.  `one` is the value 1 in a
// register the assembler cannot see through, which keeps the two multiply-adds on the FMA pipe.
#ifdef ZLES_EMU
#define LZ_CAND(r)                                                 \
  do {                                                             \
    const u32 v_ = es ^ rp[-(int)(r)];                             \
    const u32 z_ = ~v_ & (v_ * one - 1u) & 0x80808080u;            \
    if ((r) <= rrun) best = umax(best, z_ | (64u - (r)));          \
  } while (0)
#else
#define LZ_CAND(r)                                                 \
  asm("{\n\t"                                                      \
      ".reg .pred p;\n\t"                                          \
      ".reg .b32 v, t, z;\n\t"                                     \
      "xor.b32 v, %1, %2;\n\t"                                     \
      "mad.lo.u32 t, v, %3, 0xffffffff;\n\t"                       \
      "lop3.b32 z, v, t, 0x80808080, 0x08;\n\t"                    \
      "mad.lo.u32 z, z, %3, %5;\n\t"                               \
      "setp.ge.u32 p, %4, %6;\n\t"                                 \
      "@p max.u32 %0, %0, z;\n\t"                                  \
      "}"                                                          \
      : "+r"(best)                                                 \
      : "r"(es), "r"(rp[-(int)(r)]), "r"(one), "r"(rrun), "n"(64 - (r)), "n"(r))
#endif

// TABLE = false: one stream (P.in, P.n).  TABLE = true: a batch of independent buffers (P.table): groups of LZ_BGROUP
// consecutive blocks; a group whose blocks are all small first blocks (<= LZ_BSLOT bytes, no window) shares ONE sort —
// its buffers sit in 4 KiB slots of the staged 64 KiB and a position's candidates stop at its slot's start — which is
// what makes 4 KiB buffers (BASELINE configs[2]) cost 1/16 of a sort each instead of a whole one; other groups are
// processed a block at a time.
template <bool TABLE>
__device__ __forceinline__ void lz_body(const LzParams &P) {
  ZLES_SMEM_DECL(smem);
  u8 *data = smem + LZ_OFF_DATA;
  u16 *X = reinterpret_cast<u16 *>(smem + LZ_OFF_X);
  u32 *XR = reinterpret_cast<u32 *>(smem + LZ_OFF_X);
  u16 *wh = reinterpret_cast<u16 *>(smem + LZ_OFF_WH);
  u32 *bm = reinterpret_cast<u32 *>(smem + LZ_OFF_WH);          // [1024]
  u32 *hcopies = reinterpret_cast<u32 *>(smem + LZ_OFF_WH) + 1024;  // [LZ_HCOPIES][LZ_NSYM]
  u32 *scratch = reinterpret_cast<u32 *>(smem + LZ_OFF_MISC);       // [40]
  u32 *specexit = scratch + 40;                                      // [LZ_NWALK]
  u64 *mbar = reinterpret_cast<u64 *>(scratch + 40 + LZ_NWALK);     // 8-byte aligned: (40+128)*4 = 672
  u64 *red = reinterpret_cast<u64 *>(scratch + 40 + LZ_NWALK + 2);  // [2][LZ_WARPS] u64
  u32 *slice_ctr = scratch + 40 + LZ_NWALK + 2 + 4 * LZ_WARPS;       // S3 work counter
  u32 *seg_len = slice_ctr + 2;                                      // [LZ_BGROUP] bytes of every block of a packed group

  const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
  u32 parity = 0;
#ifndef ZLES_EMU
  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();
#endif
  if (tid == 0) seg_len[0] = seg_len[1] = 0;  // stream mode: [0] 1 + the block 2 this unit is judged by, [1] a pending single-block unit
  u32 *Y = P.scratch + (size_t)blockIdx.x * 2 * SUB;  // sort pass buffer, 2 * SUB entries
  u32 *R = P.scratch + (size_t)blockIdx.x * 2 * SUB;  // then the match results, SUB entries

  // A unit of work is one sort: up to 64 KiB in shared memory = [window | own blocks].  The first two blocks of a
  // chunk share one sort (block 0 has no window, block 1's window is block 0: one sorted array serves both); blocks
  // 2 and 3 each take a sort with the block before as window (pair_mode 0), or share one without window for block 2
  // (pair_mode 1).  Batch mode: one block per unit, as its table entry says.
  const u32 upc = P.pair_mode ? 2u : 3u;  // units per chunk (stream mode)
  const u32 c_begin = P.first_block / SUBS_PER_CHUNK, c_end = (P.nblocks + SUBS_PER_CHUNK - 1) / SUBS_PER_CHUNK;
  const u32 nunits = TABLE ? (P.nblocks - P.first_block + LZ_BGROUP - 1) / LZ_BGROUP : (c_end - c_begin) * upc;
  for (;;) {
    // units come from a counter (their cost differs: two blocks or one); all two-block units are handed out first.
    // Bit 31 set: not a unit of the counter but block (ui & 0x7fffffff) on its own, with the block before it as window — a
    // block 2 that this CTA matches a second time (see the end of a unit), waiting in seg_len[1].
    __syncthreads();
    if (tid == 0) {
      if (!TABLE && seg_len[1]) { *slice_ctr = seg_len[1]; seg_len[1] = 0; }
      else *slice_ctr = atomicAdd(P.unit_ctr, 1u);
    }
    __syncthreads();
    const u32 ui = *slice_ctr;
    if (!(ui & 0x80000000u) && ui >= nunits) break;
    // batch mode: is the group one packed unit, or `gcnt` units of one block each?
    u32 gfirst = 0, gcnt = 1;
    bool packed = false;
    if (TABLE) {
      gfirst = P.first_block + ui * LZ_BGROUP;
      gcnt = umin(LZ_BGROUP, P.nblocks - gfirst);
      const u32 gi = tid & (LZ_BGROUP - 1);
      int small = 1;
      if (gi < gcnt) {
        const BatchBlk t = P.table[gfirst + gi];
        small = t.hist_len == 0 && t.own_len <= LZ_BSLOT;
        if (tid < LZ_BGROUP) seg_len[tid] = t.own_len;
      } else if (tid < LZ_BGROUP) {
        seg_len[tid] = 0;
      }
      packed = __syncthreads_and(small) != 0 && gcnt > 1;
    }
    const u32 nrep = (TABLE && !packed) ? gcnt : 1u;
    for (u32 rep = 0; rep < nrep; rep++) {
    u64 own_off = 0;
    u32 own_len, hist_len, bfirst;
    if (TABLE) {
      if (packed) {
        own_len = gcnt * LZ_BSLOT; hist_len = 0; bfirst = gfirst;
      } else {
        const BatchBlk t = P.table[gfirst + rep];
        own_off = t.in_off; own_len = t.own_len; hist_len = t.hist_len; bfirst = gfirst + rep;
      }
    } else {
      u32 v;  // chunk * upc + unit within the chunk
      if (ui & 0x80000000u) v = ((ui & 0x7fffffffu) / SUBS_PER_CHUNK) * upc + 1;
      else if (ui < c_end - c_begin) v = (c_begin + ui) * upc;
      else { const u32 r = ui - (c_end - c_begin); v = (c_begin + r / (upc - 1)) * upc + 1 + r % (upc - 1); }
      const u32 chunk = v / upc, k = v % upc;
      u32 sb0, nsb;
      if (ui & 0x80000000u) { sb0 = (ui & 0x7fffffffu) % SUBS_PER_CHUNK; nsb = 1; hist_len = SUB; }
      else if (k == 0) { sb0 = 0; nsb = 2; hist_len = 0; }
      else if (P.pair_mode) {
        // blocks {2,3} share a sort, block 2 without a window — or, when the chunk ends with block 2, that block alone
        // WITH its window (nothing to share the sort with, and no block 3 to judge the missing window by: see below)
        const bool has3 = chunk * SUBS_PER_CHUNK + 3 < P.nblocks;
        sb0 = 2; nsb = has3 ? 2u : 1u; hist_len = has3 ? 0u : SUB;
      }
      else { sb0 = k + 1; nsb = 1; hist_len = SUB; }  // window = previous SUB of the same chunk
      // (remembered in shared memory, not in registers — the kernel sits at its 64-register cap: 1 + the block 2 that is
      // judged at the end of this unit, 0 for every other kind of unit)
      if (tid == 0) seg_len[0] = (!(ui & 0x80000000u) && k == 1 && P.pair_mode && nsb == 2) ? chunk * SUBS_PER_CHUNK + 3 : 0u;
      bfirst = chunk * SUBS_PER_CHUNK + sb0;
      if (bfirst >= P.nblocks) continue;  // the stream's last chunk is short (uniform over the CTA)
      own_off = (u64)bfirst * SUB;
      own_len = (u32)umin64((u64)nsb * SUB, P.n - own_off);
    }
    const u32 SEG = (TABLE && packed) ? LZ_BSLOT : SUB;  // bytes per deflate block slot of the unit
    const u32 nsub = (TABLE && packed) ? gcnt : (own_len > SUB ? 2u : 1u);  // deflate blocks in this unit
    const u32 L = hist_len + own_len;

    LZ_CLK(scratch, -1);
    // S0: stage window + block, zero the pad so word reads past the end are defined
    if (TABLE && packed) {
      // every buffer into its slot, the rest of the slot zeroed; 16 bytes at a time where the source allows it
      for (u32 i = tid; i < (L >> 4); i += LZ_THREADS) {
        const u32 sl = i / (LZ_BSLOT >> 4), o = (i % (LZ_BSLOT >> 4)) << 4;
        const u32 sl_len = seg_len[sl];
        uint4 v4 = make_uint4(0, 0, 0, 0);
        if (o < sl_len) {
          const u8 *src = P.in + P.table[gfirst + sl].in_off + o;
          if (o + 16 <= sl_len && ((uintptr_t)src & 15) == 0) {
            v4 = *reinterpret_cast<const uint4 *>(src);
          } else {
            u32 wv[4] = {0, 0, 0, 0};
#pragma unroll
            for (u32 q = 0; q < 16; q++)
              if (o + q < sl_len) wv[q >> 2] |= (u32)src[q] << ((q & 3) * 8);
            v4 = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          }
        }
        *reinterpret_cast<uint4 *>(data + (i << 4)) = v4;
      }
      __syncthreads();
    } else {
      stage_g2s(data, P.in + own_off - hist_len, L, mbar, parity);
    }
    for (u32 i = tid; i < LZ_PAD; i += LZ_THREADS) data[L + i] = 0;
    __syncthreads();
    LZ_CLK(scratch, 0);

    // S1: Adler-32 partial sums of every own block (K8 fused into the load)
    for (u32 sbi = 0; sbi < nsub; sbi++) {
      const u32 sbase = sbi * SEG, slen = (TABLE && packed) ? seg_len[sbi] : umin(SUB, own_len - sbase);
      u64 a = 0, bsum = 0;
      const u32 j0 = tid * 32;
      if (j0 < slen) {
        // 32 bytes as two 16-byte vectors; bytes past the unit's end are zero (the pad) and add nothing (a block that
        // is not the unit's last is full).  sum (W - i) d[i] = W * sum d[i] - sum i d[i], the two sums by dp4a
        const uint4 *v4 = reinterpret_cast<const uint4 *>(data + hist_len + sbase + j0);
        const u32 W = slen - j0;
        u32 sa = 0, si = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint4 x = v4[h];
          sa = __dp4a(x.x, 0x01010101u, sa); sa = __dp4a(x.y, 0x01010101u, sa);
          sa = __dp4a(x.z, 0x01010101u, sa); sa = __dp4a(x.w, 0x01010101u, sa);
          const u32 o4 = 0x10101010u * (u32)h;  // byte offsets 16 h + i
          si = __dp4a(x.x, 0x03020100u + o4, si); si = __dp4a(x.y, 0x07060504u + o4, si);
          si = __dp4a(x.z, 0x0b0a0908u + o4, si); si = __dp4a(x.w, 0x0f0e0d0cu + o4, si);
        }
        a = sa; bsum = (u64)W * sa - si;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(ZLES_FULL, a, d);
        bsum += __shfl_down_sync(ZLES_FULL, bsum, d);
      }
      if (lane == 0) { red[w] = a; red[LZ_WARPS + w] = bsum; }
      __syncthreads();
      if (tid == 0) {
        *slice_ctr = 0;  // consumed in S3, several barriers from here
        u64 ta = 0, tb = 0;
        for (int i = 0; i < LZ_WARPS; i++) { ta += red[i]; tb += red[LZ_WARPS + i]; }
        P.adler_part[2 * (size_t)(bfirst + sbi)] = ta;
        P.adler_part[2 * (size_t)(bfirst + sbi) + 1] = tb;
      }
      __syncthreads();  // red is reused by the next block
    }
    LZ_CLK(scratch, 1);

    // S2: stable radix sort of positions by hash16(key3)
    const u32 N = L >= 3 ? L - 2 : 0;
    u32 lper = 5;  // items per warp and pass: the power of two >= ceil(N / warps), at least one batch
    while (((u32)LZ_WARPS << lper) < N) lper++;
    lz_sort(data, Y, X, wh, scratch, N, lper);

    // S3: per-position match search.  Warps take slices of the sorted array.
    {
      u32 *ring = reinterpret_cast<u32 *>(smem + LZ_OFF_WH) + w * 128;
      const u32 scan = umin(P.max_checks, LZ_SCAN);
      const u32 one = (P.max_checks >> 31) + 1;  // 1 (max_checks <= LZ_SCAN is enforced by the host)
      // slices of LZ_SLICE sorted entries are handed out dynamically (their cost depends on the run structure)
      for (;;) {
      u32 slice = 0;
      if (lane == 0) slice = atomicAdd(slice_ctr, 1u);
      slice = __shfl_sync(ZLES_FULL, slice, 0);
      const u32 kbeg = slice * LZ_SLICE;
      if (kbeg >= N) break;
      const u32 kend = umin(kbeg + LZ_SLICE, N);
      u32 Bprev = 0, hprev = 0xffffffffu;
      if (kbeg > 0) {  // the 32 entries before the slice are candidates of its first ones (kbeg is a multiple of 32)
        const u32 j = kbeg - 32 + lane;
        u32 lo, hi;
        lz_ld56(data, X[j], lo, hi);
        const u32 q = lz_mix(lo);
        const u32 e = lz_tag(q, lo, hi);
        ring[j & 63] = e;
        ring[(j & 63) + 64] = e;
        const u32 hj = q >> 16;
        u32 hl = __shfl_up_sync(ZLES_FULL, hj, 1);
        if (lane == 0) hl = 0xffffffffu;  // whether entry kbeg-32 starts a run never matters (see rrun below)
        Bprev = __ballot_sync(ZLES_FULL, hj != hl);
        hprev = __shfl_sync(ZLES_FULL, hj, 31);
        __syncwarp();
      }
      for (u32 kb = kbeg; kb < kend; kb += 32) {
        LZ_WCLK_DECL;
        const u32 k = kb + lane;
        const bool valid = k < kend;
        u32 p = 0, es = 0, h = 0x10000u + lane;
        if (valid) {
          u32 lo, hi;
          p = X[k];
          lz_ld56(data, p, lo, hi);
          const u32 q = lz_mix(lo);
          es = lz_tag(q, lo, hi);
          h = q >> 16;
        }
        ring[k & 63] = es;
        ring[(k & 63) + 64] = es;
        u32 hl = __shfl_up_sync(ZLES_FULL, h, 1);
        if (lane == 0) hl = hprev;
        const u32 B = __ballot_sync(ZLES_FULL, h != hl);  // bit l: entry kb+l starts a run of equal hashes
        hprev = __shfl_sync(ZLES_FULL, h, 31);
        __syncwarp();
        // rrun = how many entries before k belong to k's run (capped at the scan width): the only ones that
        // may be looked at — what precedes them in the ring is another run, an older slice, or stale
        // the block slot of p: where its candidates start (the window, or a packed buffer's own start) and where it ends
        u32 lim = p > WINDOW ? p - WINDOW : 0, bend = (nsub == 2 && p < SUB) ? SUB : L;
        bool own = valid && p >= hist_len;
        if (TABLE && packed) {
          lim = p & ~(LZ_BSLOT - 1);
          bend = lim + seg_len[p / LZ_BSLOT];
          own = valid && p < bend;
        }
        u32 rrun = 0;
        if (own) {
          // bit 31 = entry k's run-start flag, bit 30 = entry k-1's, ...: the leading zeros count the run's earlier entries
          rrun = umin((u32)__clz((int)__funnelshift_rc(Bprev, B, lane + 1)), scan);
          if (rrun && (u32)X[k - rrun] < lim) {
            // candidates older than the window (src/lz77.ts:49) or from another buffer of a packed group: positions
            // ascend inside a run, so binary-search the largest r with X[k-r] >= lim
            u32 a = 0, b = rrun;  // X[k-a] in window (a = 0: k itself), X[k-b] not
            while (b - a > 1) {
              const u32 m = (a + b) >> 1;
              if ((u32)X[k - m] >= lim) a = m; else b = m;
            }
            rrun = a;
          }
        }
        const u32 rmax = __reduce_max_sync(ZLES_FULL, rrun);
        // Branch-free, fully unrolled body: candidate r is entry k-r of the ring, at a constant offset from `rp`.
        // v = own tag ^ candidate's tag: its zero bytes from the bottom are [fold, byte 3, byte 4, byte 5];
        // ~v & (v - 1) has bit 7 of exactly those bytes set, so z orders candidates by match length 2 + popc(z)
        // (0: another key; 3, 4, 5; 6 = at least six, to be extended), and z | (64 - r) prefers the nearer of equals
        // (src/lz77.ts:86-92).  The loop is bound by the ALU pipe (one warp instruction per 2 cycles and scheduler),
        // so LZ_CAND spells it out: 4 ALU instructions, 2 integer multiply-adds (FMA pipe) and the load.
        u32 best = 0;
        const u32 *rp = ring + (k & 63) + 64;
        LZ_WCLK(12);
        if (rmax >= 1) {
          LZ_CAND(1); LZ_CAND(2); LZ_CAND(3); LZ_CAND(4); LZ_CAND(5); LZ_CAND(6); LZ_CAND(7); LZ_CAND(8);
          if (rmax >= 9) {
            LZ_CAND(9); LZ_CAND(10); LZ_CAND(11); LZ_CAND(12); LZ_CAND(13); LZ_CAND(14); LZ_CAND(15); LZ_CAND(16);
            if (rmax >= 17) {
              LZ_CAND(17); LZ_CAND(18); LZ_CAND(19); LZ_CAND(20); LZ_CAND(21); LZ_CAND(22); LZ_CAND(23); LZ_CAND(24);
              if (rmax >= 25) {
                LZ_CAND(25); LZ_CAND(26); LZ_CAND(27); LZ_CAND(28); LZ_CAND(29); LZ_CAND(30); LZ_CAND(31); LZ_CAND(32);
              }
            }
          }
        }
        // Within five bytes of its block's end a position's tag holds bytes from behind the end (the next block, another
        // buffer of a packed group): they must not rank its candidates — whatever follows the block, the same nearest
        // candidate wins (a buffer compresses to the same bytes alone, in a batch or in a packed group).  Rare: redone
        // here with those bytes masked out.
        if (rrun && bend - p < 6) {  // (rrun != 0 implies own; no warp collective inside: plain divergence)
          const u32 m = bend - p;
          const u32 vmask = 0xffu | (m > 3 ? 0xff00u : 0u) | (m > 4 ? 0xff0000u : 0u);
          best = 0;
          for (u32 r = 1; r <= rrun; r++) {
            const u32 v = (es ^ rp[-(int)r]) & vmask;
            best = umax(best, (~v & (v - 1) & 0x80808080u) | (64u - r));
          }
        }
        LZ_WCLK(13);
        // The winner: the nearest candidate of the longest class.  Classes 3..5 are exact lengths; the class "six or
        // more" is extended — its first 8 bytes in straight-line code for the whole warp, the rest (rare on text) in a loop.
        const bool hit = own && best >= 0x80u;
        if (__any_sync(ZLES_FULL, hit)) {
          const u32 c = hit ? (u32)X[k - (64 - (best & 0x7fu))] : p;
          // a match ends with its deflate block (bend)
          const u32 maxlen = own ? umin(MAX_MATCH, bend - p) : 0;
          u32 len = 2u + (u32)__popc(best & 0x80808080u);
          if (__any_sync(ZLES_FULL, hit && len == 6)) {
            const u64 x = lz_ld64(data, p + 6) ^ lz_ld64(data, c + 6);
            if (hit && len == 6) {
              len = x ? 6 + ((u32)(__ffsll((long long)x) - 1) >> 3) : 14;
              if (!x && maxlen > 14) len = 14 + lz_match_len(data, c + 14, p + 14, maxlen - 14);
            }
          }
          len = umin(len, maxlen);
          const u32 dist = p - c;
          if (len == MIN_MATCH && dist > 4096) len = 0;  // costs more than three literals
          if (own) R[p - hist_len] = (hit && len >= MIN_MATCH) ? (len << 16) | dist : 0;
        } else if (own) {
          R[p - hist_len] = 0;
        }
        Bprev = B;
        __syncwarp();  // the next batch overwrites the older half of the ring
        LZ_WCLK(14);
      }
      }  // slices
      // positions without a full 3-byte key (the last two of the window+block) have no match
      if (tid < 2 && own_len > tid) R[own_len - 1 - tid] = 0;
    }
    __syncthreads();
    LZ_CLK(scratch, 8);

    // S4/S5 of a packed group: eight buffers (32 KiB of slots) at a time — their parses run side by side (the first
    // walker of every buffer starts at its first byte, nothing crosses a slot), tokens, counts and histograms go to
    // every buffer's own block.  Done one buffer after the other this stage cost more than sort and match together.
    if (TABLE && packed) {
      u32 *pref = reinterpret_cast<u32 *>(smem + LZ_OFF_WH + 16384);  // [1024] exclusive token counts per bitmap word
      constexpr u32 WPS = LZ_BSLOT / LZ_RANGE;                         // walkers per slot
      constexpr u32 SPH = SUB / LZ_BSLOT;                              // slots per half (8 = LZ_HCOPIES)
      static_assert(SPH == LZ_HCOPIES, "one histogram copy per slot of a half");
      for (u32 h = 0; h * SPH < gcnt; h++) {
        const u32 nsl = umin(SPH, gcnt - h * SPH), hbase = h * SUB, hlen = nsl * LZ_BSLOT;
        {
          const uint4 *src4 = reinterpret_cast<const uint4 *>(R + hbase);
          uint4 *dst4 = reinterpret_cast<uint4 *>(XR);
          const u32 n4 = hlen >> 2;
          for (u32 i0 = 0; i0 < n4; i0 += 8 * LZ_THREADS) {
            uint4 v8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const u32 i = i0 + u * LZ_THREADS + tid; if (i < n4) v8[u] = __ldcg(src4 + i); }
#pragma unroll
            for (int u = 0; u < 8; u++) { const u32 i = i0 + u * LZ_THREADS + tid; if (i < n4) dst4[i] = v8[u]; }
          }
        }
        bm[tid] = 0;
        for (u32 i = tid; i < LZ_HCOPIES * LZ_NSYM; i += LZ_THREADS) hcopies[i] = 0;
        __syncthreads();
        // the walkers: range [ws, we) of slot wsl; a slot's first walker has a fixed entry
        const u32 wsl = tid / WPS;
        const u32 sbeg = wsl * LZ_BSLOT, send = sbeg + (tid < LZ_NWALK && wsl < nsl ? seg_len[h * SPH + wsl] : 0u);
        const u32 ws = tid * LZ_RANGE, we = umin(ws + LZ_RANGE, send);
        const bool walker = tid < LZ_NWALK && ws < send;
        u32 wentry = ws;
        if (walker) {
          u32 pos = ws;
          while (pos < we) {
            bm[pos >> 5] |= 1u << (pos & 31);
            const u32 len = lz_parse_len(XR, pos, send, P.lazy);
            pos += len ? len : 1;
          }
          specexit[tid] = pos;
        }
        for (;;) {
          __syncthreads();
          const u32 entry = (walker && (tid % WPS) != 0) ? specexit[tid - 1] : ws;
          __syncthreads();
          int changed = 0;
          if (walker && entry != wentry) {
            wentry = entry;
            u32 exitpos = specexit[tid];
            lz_clear_bits(bm, ws, umin(entry, we));
            if (entry >= we) {
              exitpos = entry;
            } else {
              u32 pos = entry;
              for (;;) {
                if (pos >= we) { exitpos = pos; break; }
                if ((bm[pos >> 5] >> (pos & 31)) & 1) break;
                bm[pos >> 5] |= 1u << (pos & 31);
                const u32 len = lz_parse_len(XR, pos, send, P.lazy);
                const u32 nxt = pos + (len ? len : 1);
                lz_clear_bits(bm, pos + 1, umin(nxt, we));
                pos = nxt;
              }
            }
            if (exitpos != specexit[tid]) { specexit[tid] = exitpos; changed = 1; }
          }
          if (!__syncthreads_or(changed)) break;
        }
        // emit: bitmap word tid belongs to slot tid / 128
        {
          u32 word = bm[tid];
          u32 total;
          const u32 oex = block_exscan((u32)__popc(word), scratch, &total);
          pref[tid] = oex;
          __syncthreads();
          const u32 sl = tid >> 7;  // 128 words of 32 positions per slot
          const u32 bcur = bfirst + h * SPH + sl;
          const u32 send2 = sl * LZ_BSLOT + (sl < nsl ? seg_len[h * SPH + sl] : 0u);
          u32 o = oex - pref[tid & ~127u];
          u32 *tok = P.tokens + (size_t)bcur * P.tok_stride;
          u32 *hc = hcopies + sl * LZ_NSYM;
          while (word) {
            const u32 bit = (u32)(__ffs((int)word) - 1);
            word &= word - 1;
            const u32 pos = tid * 32 + bit;
            const u32 len = lz_parse_len(XR, pos, send2, P.lazy);
            if (len) {
              const u32 dist = XR[pos] & 0xffff;
              u32 ls, le, lv, ds, de, dv;
              len_to_sym(len, ls, le, lv);
              dist_to_sym(dist, ds, de, dv);
              atomicAdd(hc + 257 + ls, 1u);
              atomicAdd(hc + 288 + ds, 1u);
              tok[o] = tok_match(len, dist);
            } else {
              const u32 d = data[hbase + pos];
              atomicAdd(hc + d, 1u);
              tok[o] = d;
            }
            o++;
          }
          __syncthreads();
          for (u32 i = tid; i < nsl * LZ_NSYM; i += LZ_THREADS)
            P.hist[(size_t)(bfirst + h * SPH + i / LZ_NSYM) * LZ_NSYM + i % LZ_NSYM] = hcopies[i];
          if ((tid & 127) == 0 && sl < nsl) P.ntok[bcur] = (sl + 1 < SPH ? pref[tid + 128] : total) - oex;
        }
        __syncthreads();
      }
    }
    // S4/S5 once per deflate block of the unit (the match results stay in the scratch; XR takes one block's at a time)
    const u32 unit_own = own_len;
    for (u32 sbi = 0; sbi < ((TABLE && packed) ? 0u : nsub); sbi++) {
    const u32 sbase = sbi * SEG, bcur = bfirst + sbi, dbase = hist_len + sbase;
    const u32 own_len = (TABLE && packed) ? seg_len[sbi] : umin(SUB, unit_own - sbase);  // this block's bytes (shadows the unit's)
    // S4: match results into shared memory (over the sorted array), then the parse
    {  // 16 bytes per load: the copy is bound by the round trips to L2, not by their width
      const uint4 *src4 = reinterpret_cast<const uint4 *>(R + sbase);
      uint4 *dst4 = reinterpret_cast<uint4 *>(XR);
      const u32 n4 = (own_len + 3) >> 2;  // R has room for 2 * SUB entries: reading up to 3 entries past own_len is harmless
      for (u32 i0 = 0; i0 < n4; i0 += 8 * LZ_THREADS) {  // eight loads in flight per thread, then the stores
        uint4 v8[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const u32 i = i0 + u * LZ_THREADS + tid; if (i < n4) v8[u] = __ldcg(src4 + i); }
#pragma unroll
        for (int u = 0; u < 8; u++) { const u32 i = i0 + u * LZ_THREADS + tid; if (i < n4) dst4[i] = v8[u]; }
      }
    }
    bm[tid] = 0;
    for (u32 i = tid; i < LZ_HCOPIES * LZ_NSYM; i += LZ_THREADS) hcopies[i] = 0;
    __syncthreads();
    LZ_CLK(scratch, 9);
    // Every walker parses its range from the range start (speculation), recording token starts in the
    // bitmap and where it left the range.  Then, in rounds, every walker whose true entry (= where the
    // previous range was really left) differs from the entry it last used re-parses from there until
    // it meets its old path again — a greedy parse re-synchronises within a few tokens — so almost
    // all ranges are final after two rounds; the loop ends when no exit moved.
    const bool walker = tid < LZ_NWALK && tid * LZ_RANGE < own_len;
    const u32 ws = tid * LZ_RANGE, we = umin(ws + LZ_RANGE, own_len);
    u32 wentry = ws;
    if (walker) {
      u32 pos = ws;
      while (pos < we) {
        bm[pos >> 5] |= 1u << (pos & 31);
        u32 len = lz_parse_len(XR, pos, own_len, P.lazy);
        pos += len ? len : 1;
      }
      specexit[tid] = pos;
    }
    for (;;) {
      __syncthreads();
      const u32 entry = (walker && tid > 0) ? specexit[tid - 1] : ws;
      __syncthreads();  // every exit has been read before any is rewritten
      int changed = 0;
      if (walker && entry != wentry) {
        wentry = entry;
        u32 exitpos = specexit[tid];
        lz_clear_bits(bm, ws, umin(entry, we));
        if (entry >= we) {
          exitpos = entry;  // a token of the previous range covers this one entirely
        } else {
          u32 pos = entry;
          for (;;) {
            if (pos >= we) { exitpos = pos; break; }
            if ((bm[pos >> 5] >> (pos & 31)) & 1) break;  // re-synchronised: the rest of the old path, and its exit, stand
            bm[pos >> 5] |= 1u << (pos & 31);
            u32 len = lz_parse_len(XR, pos, own_len, P.lazy);
            u32 nxt = pos + (len ? len : 1);
            lz_clear_bits(bm, pos + 1, umin(nxt, we));
            pos = nxt;
          }
        }
        if (exitpos != specexit[tid]) { specexit[tid] = exitpos; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
    LZ_CLK(scratch, 10);

    // S5: emit tokens and count symbols
    {
      u32 word = bm[tid];
      u32 total;
      u32 o = block_exscan((u32)__popc(word), scratch, &total);
      u32 *tok = P.tokens + (size_t)bcur * P.tok_stride;
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
          const u32 d = data[dbase + pos];
          atomicAdd(hc + d, 1u);
          tok[o] = d;
        }
        o++;
      }
      __syncthreads();
      if (tid < LZ_NSYM) {
        u32 s = 0;
        for (u32 c = 0; c < LZ_HCOPIES; c++) s += hcopies[c * LZ_NSYM + tid];
        P.hist[(size_t)bcur * LZ_NSYM + tid] = s;
      }
      if (tid == 0) P.ntok[bcur] = total;
    }
    __syncthreads();
    }  // blocks of the unit
    __syncthreads();
    // Pair mode gave block 2 no window.  What that cost is judged by block 3, which had one (block 2): when block 2 came out
    // in a third more tokens than block 3 — a period longer than a few hundred bytes costs its literals again, a patchwork
    // repeats what block 1 held — block 2 is matched again, alone, with block 1 as window (this CTA's next unit).  Text
    // and binary data stay at 1.05–1.15 (tools/gpu_size_sweep.py), zeros at 1.0: no second pass.
    if (!TABLE && tid == 0 && seg_len[0]) {
      const u32 b2 = seg_len[0] - 1;
      const u32 nt2 = P.ntok[b2], nt3 = P.ntok[b2 + 1];  // (thread 0's own stores)
      if (3 * nt2 > 4 * nt3 + LZ_REDO_SLACK) seg_len[1] = 0x80000000u | b2;
    }
    LZ_CLK(scratch, 11);
    }  // blocks of an unpacked group
  }
}

__global__ void __launch_bounds__(LZ_THREADS, 1) k_lz(const LzParams P) { lz_body<false>(P); }
__global__ void __launch_bounds__(LZ_THREADS, 1) k_lz_batch(const LzParams P) { lz_body<true>(P); }

}  // namespace zles
