// inflate.cuh — kernels K6/K7: Huffman decode and LZ77 back-reference copy.
//
// Replaces /root/reference/src/inflate.ts:16-292 (bit-serial canonical decode, one
// `read()` per code bit, byte-serial copy), src/huffman.ts:8-53 (decode tables),
// src/utils/BitReadStream.ts and src/utils/Uint8WriteStream.ts.
//
// Three tiers, tried in this order by zles.cu:
//  1. OUR streams (this file, "fast path"): every 32 KiB block ends with an empty stored block
//     (the 00 00 FF FF sync marker), so block starts are found by a byte scan (k_mark_*), every
//     block ("segment") is Huffman-decoded into tokens on its own warp (k_inf_tokens), the result
//     is verified (k_inf_check) and the copies are resolved one warp per 128 KiB chunk
//     (k_inf_resolve).
//  2. streams of other encoders (inflate_foreign.cuh): block headers are searched at every bit.
//  3. everything else, and every error: k_inflate, one warp, sequential, exactly the reference's
//     order of events including BitReadStream's end-of-buffer behaviour.  k_inflate_batch runs
//     tier 3 on a batch of independent streams, one warp per stream.
//
// Roofline: HBM-bound in principle (algorithmic bytes = C read + U written), in practice bound
// by the serial decode chain of each warp; see DESIGN.md.
#pragma once
#include "zles_dev.h"

namespace zles {

constexpr int INF_WARPS = 4;
constexpr int INF_THREADS = INF_WARPS * 32;
constexpr int LL_ROOT = 10;
constexpr int D_ROOT = 8;
constexpr int CL_ROOT = 7;

// status codes of a segment (the error values equal the ZLES_E_* codes of include/zles.h)
constexpr u32 SEG_SYNC = 100;      // ended on an empty stored block
constexpr u32 SEG_FINAL = 101;     // ended with BFINAL
constexpr u32 SEG_E_BTYPE3 = 2;    // 'Not supported BTYPE : 3'        src/inflate.ts:32
constexpr u32 SEG_E_INSUFF = 3;    // 'Data length is insufficient'    src/inflate.ts:35
constexpr u32 SEG_E_CORRUPT = 4;   // 'Data is corrupted'              src/inflate.ts:50,88,166,247,276
constexpr u32 SEG_E_LACK = 5;      // 'Lack of data length'            src/utils/BitReadStream.ts:15
constexpr u32 SEG_E_RUNAWAY = 20;  // the reference never returns on this input (inf_coded); ZLES_E_RUNAWAY
constexpr u32 SEGF_HISTORY = 1;    // a distance reached before the segment start
constexpr u32 SEGF_OVERFLOW = 2;   // output did not fit (out_len is still the true length)

struct InfRes {
  u64 end_pos;  // byte position (in the stream) just after the segment
  u64 out_len;  // bytes this segment decodes to
  u32 status;
  u32 flags;
};

struct InfTab {  // canonical description of one code, in shared memory
  u32 first[16];
  u16 cnt[16];
  u16 off[16];
};

struct InfWarpSmem {
  u16 lut_ll[1 << LL_ROOT];
  u16 lut_d[1 << D_ROOT];  // doubles as the code-length-code LUT while a header is parsed
  u16 sorted_ll[288];
  u16 sorted_d[64];
  InfTab tab_ll, tab_d;
  u16 cur[16];
  u8 lens[352];  // [0,288) literal/length code lengths, [288,352) distance code lengths (a run of the code-length
                 // code may spill up to 5 symbols past HDIST <= 32 — src/inflate.ts:187-200 keeps them, so do we)
  u8 cl_lens[32];
};
constexpr int INF_SMEM = (int)(sizeof(InfWarpSmem) + sizeof(InfRes)) * INF_WARPS;  // per-warp tables | per-warp result (batch kernel)

// ---- bit reader: warp-uniform state, input fetched 128 B per warp at a time ----
struct InfReader {
  const u8 *in;   // stream base
  u64 n;          // stream length in bytes
  const u32 *words;  // `in` rounded down to 4 bytes
  u32 skew;          // in - words, in bytes
  u64 widx;          // next word to consume
  u64 win_base;      // word index held by lane 0 of `win`
  u32 win, win_next;
  u64 bb;
  u32 bc;

  __device__ __forceinline__ u32 load_word(u64 wi) const {
    // word wi covers stream bytes [4*wi - skew, 4*wi - skew + 4); bytes outside [0, n) read as 0
    // (/root/reference/src/utils/BitReadStream.ts:33-35 reads `undefined << k === 0` past the end)
    long long lo = (long long)(wi << 2) - (long long)skew;
    if (lo >= 0 && (u64)lo + 4 <= n) return __ldg(words + wi);
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      long long b = lo + k;
      if (b >= 0 && (u64)b < n) v |= (u32)in[b] << (8 * k);
    }
    return v;
  }
  __device__ __forceinline__ u32 next_word() {
    if (widx - win_base >= 32) {
      win_base += 32;
      win = win_next;
      win_next = load_word(win_base + 32 + lane_id());
    }
    u32 w = __shfl_sync(ZLES_FULL, win, (int)(widx & 31));
    widx++;
    return w;
  }
  __device__ __forceinline__ void init(const u8 *in_, u64 n_, u64 byte_pos) {
    in = in_; n = n_;
    skew = (u32)((uintptr_t)in_ & 3);
    words = reinterpret_cast<const u32 *>(in_ - skew);
    u64 a = byte_pos + skew;
    widx = a >> 2;
    win_base = widx & ~(u64)31;
    win = load_word(win_base + lane_id());
    win_next = load_word(win_base + 32 + lane_id());
    u32 sh = (u32)(a & 3) * 8;
    bb = (u64)(next_word() >> sh);
    bc = 32 - sh;
    refill();
  }
  __device__ __forceinline__ void refill() {
    if (bc <= 32) {
      bb |= (u64)next_word() << bc;
      bc += 32;
    }
  }
  __device__ __forceinline__ u32 peek(u32 k) const { return (u32)bb & ((1u << k) - 1); }
  __device__ __forceinline__ void skip(u32 k) { bb >>= k; bc -= k; }
  __device__ __forceinline__ u32 take(u32 k) { u32 v = peek(k); skip(k); return v; }
  // bit position (from stream byte 0) of the next unread bit
  __device__ __forceinline__ u64 bitpos() const { return (widx << 5) - bc - ((u64)skew << 3); }
};

// Builds the canonical arrays and the root LUT of one code from its lengths
// (same assignment as generateHuffmanTable, /root/reference/src/huffman.ts:8-39:
// by length, then ascending symbol; code <<= 1 per length).  n is a multiple of 32.
// canonical arrays of one code: per-length counts, first codes, offsets, and the symbols sorted by (length, symbol)
__device__ __forceinline__ void inf_canon(const u8 *lens, int n, u16 *sorted, InfTab *t, u16 *cur) {
  const u32 lane = lane_id();
  if (lane < 16) t->cnt[lane] = 0;
  __syncwarp();
  for (int g = 0; g < n; g += 32) {
    u32 l = lens[g + lane];
    u32 m = __match_any_sync(ZLES_FULL, l);
    if (l && lane == (u32)(__ffs((int)m) - 1)) t->cnt[l] = (u16)(t->cnt[l] + __popc(m));
    __syncwarp();
  }
  if (lane == 0) {
    u32 code = 0, o = 0;
    t->first[0] = 0; t->off[0] = 0; cur[0] = 0;
    for (int l = 1; l < 16; l++) {
      t->first[l] = code;
      t->off[l] = (u16)o;
      cur[l] = (u16)o;
      o += t->cnt[l];
      code = (code + t->cnt[l]) << 1;
    }
  }
  __syncwarp();
  for (int g = 0; g < n; g += 32) {
    u32 l = lens[g + lane];
    u32 m = __match_any_sync(ZLES_FULL, l);
    u32 base = cur[l];
    __syncwarp();
    if (l) {
      sorted[base + __popc(m & lanemask_lt())] = (u16)(g + lane);
      if (lane == (u32)(__ffs((int)m) - 1)) cur[l] = (u16)(base + __popc(m));
    }
    __syncwarp();
  }
}
// root-LUT entry of index i (the next `root` bits of the stream, LSB first): symbol << 4 | code length, 0 = no code of
// at most `root` bits starts like this
__device__ __forceinline__ u32 inf_lut_entry(u32 i, int root, const InfTab *t, const u16 *sorted) {
  u32 code = 0;
  for (int l = 1; l <= root; l++) {
    code = (code << 1) | ((i >> (l - 1)) & 1);
    const u32 d = code - t->first[l];
    if (code >= t->first[l] && d < t->cnt[l]) return ((u32)sorted[t->off[l] + d] << 4) | (u32)l;
  }
  return 0;
}
__device__ __forceinline__ void inf_build(const u8 *lens, int n, int root, u16 *lut, u16 *sorted, InfTab *t, u16 *cur) {
  inf_canon(lens, n, sorted, t, cur);
  for (u32 i = lane_id(); i < (1u << root); i += 32) lut[i] = (u16)inf_lut_entry(i, root, t, sorted);
  __syncwarp();
}

// Canonical decode without the LUT (codes longer than the root, or invalid): tries
// lengths in increasing order like the lookup loop at /root/reference/src/inflate.ts:238-252.
__device__ __forceinline__ bool inf_slow(u64 bb, const InfTab *t, const u16 *sorted, u32 &sym, u32 &len) {
  u32 code = 0;
  for (int l = 1; l < 16; l++) {
    code = (code << 1) | (u32)((bb >> (l - 1)) & 1);
    u32 d = code - t->first[l];
    if (code >= t->first[l] && d < t->cnt[l]) {
      sym = sorted[t->off[l] + d];
      len = (u32)l;
      return true;
    }
  }
  return false;
}

// One Huffman-coded symbol in the sequential decoder, with the reference's end-of-buffer bookkeeping.
// The reference reads code bits one at a time through BitReadStream.read()
// (/root/reference/src/utils/BitReadStream.ts:14-31): a read() that consumes the last bit of the last
// byte — or of a "phantom" zero byte that an earlier readRange() pulled in past the end — sets isEnd,
// and any read() after that throws 'Lack of data length'.  readRange() (header fields, extra bits,
// stored bytes) never sets isEnd and silently yields zeros past the end.  The symbol loops run
// `while (!stream.isEnd)` and the block loop throws 'Data length is insufficient' when a non-final
// block ends with isEnd set (/root/reference/src/inflate.ts:34-36).  Returns 0 or a SEG_E_* code.
// When isEnd is set, read() leaves the bit it just consumed in nowBits with nowBitsLength = 0, so the NEXT
// readRange() returns that bit again as its first bit, followed by zeros: `stale` carries it (inf_extra).
__device__ __forceinline__ u32 inf_coded(InfReader &r, const u16 *lut, int root, const InfTab *t, const u16 *sorted, u64 nbits,
                                         u32 &is_end, u32 &stale, u32 &sym) {
  const u32 e = lut[r.peek((u32)root)];
  u32 l = e & 15;
  sym = e >> 4;
  bool found = true;
  if (l == 0) found = inf_slow(r.bb, t, sorted, sym, l);
  const u64 P = r.bitpos();
  u32 L = l;
  if (!found) {  // the reference reads up to the longest code of the table before giving up
    L = 0;
    for (int k = 15; k >= 1; k--)
      if (t->cnt[k]) { L = (u32)k; break; }
    // empty table: codelenMin stays Number.MAX_SAFE_INTEGER (src/inflate.ts:139-147, 206-224) and readRangeCoded reads
    // bit after bit until read() throws at the end of the buffer, wherever in the stream this happens
    if (L == 0) return SEG_E_LACK;
  }
  if (P + 64 >= nbits) {  // only near the end of the buffer can the bookkeeping matter
    if (is_end) return SEG_E_LACK;
    // Past the end every bit reads as zero and isEnd is only set by a read() that takes the last bit of a byte: a symbol
    // loop whose all-zero token is a multiple of 8 bits long and avoids that bit never ends in the reference (it writes
    // until the JS heap is gone, or nothing at all).  The decoder's state is then its bit offset inside a byte, so a
    // coded symbol that STARTS >= 512 bits past the end (8 tokens of <= 48 bits without an exit) proves the cycle.
    // The oracle reports it at exactly the same point (oracle/zlibes_oracle.c, decode_symbol).
    if (P >= nbits + 512) return SEG_E_RUNAWAY;
    const u64 x0 = P > nbits - 1 ? P : nbits - 1;
    const u64 E = x0 + ((7 - (x0 & 7)) & 7);  // first bit position >= x0 that is the last bit of a byte
    if (E < P + L - 1) return SEG_E_LACK;     // a read() after the one that reached the end
    if (E == P + L - 1) {
      is_end = 1;
      stale = (u32)(r.bb >> (L - 1)) & 1;
    }
  }
  if (!found) return SEG_E_CORRUPT;
  r.skip(l);
  return 0;
}

// readRange(k) of the sequential decoder (extra bits): after isEnd the first bit is the stale one, the rest are
// phantom zeros (everything past the end of the buffer reads as zero)
__device__ __forceinline__ u32 inf_extra(InfReader &r, u32 k, u32 &stale) {
  u32 v = r.take(k);
  if (k && stale) { v |= 1; stale = 0; }
  return v;
}

// Dynamic block header (/root/reference/src/inflate.ts:121-202): HLIT/HDIST/HCLEN, the code-length
// code, then the run-length coded literal/length and distance code lengths into S->lens.
// Returns false with `status` set on error.  Warp-uniform.
__device__ __forceinline__ bool inf_read_dynamic_header(InfReader &r, InfWarpSmem *S, u64 nbits, u32 &is_end, u32 &stale, u32 &status) {
  const u32 lane = lane_id();
  const u32 HLIT = r.take(5) + 257;
  const u32 HDIST = r.take(5) + 1;
  const u32 HCLEN = r.take(4) + 4;
  S->cl_lens[lane] = 0;
  for (u32 i = lane; i < 352; i += 32) S->lens[i] = 0;
  __syncwarp();
  for (u32 i = 0; i < HCLEN; i++) {
    r.refill();
    u32 v = r.take(3);
    if (lane == 0) S->cl_lens[c_cl_order[i]] = (u8)v;
  }
  __syncwarp();
  inf_build(S->cl_lens, 32, CL_ROOT, S->lut_d, S->sorted_d, &S->tab_d, S->cur);
  const u32 total = HLIT + HDIST;
  u32 prev = 0;
  for (u32 i = 0; i < total;) {
    r.refill();
    u32 sym;
    const u32 rc = inf_coded(r, S->lut_d, CL_ROOT, &S->tab_d, S->sorted_d, nbits, is_end, stale, sym);
    if (rc) { status = rc; return false; }
    u32 rep = 1, val = sym;
    if (sym == 16) { rep = 3 + inf_extra(r, 2, stale); val = prev; }
    else if (sym == 17) { rep = 3 + inf_extra(r, 3, stale); val = 0; prev = 0; }
    else if (sym == 18) { rep = 11 + inf_extra(r, 7, stale); val = 0; prev = 0; }
    else prev = sym;
    if (val) {
      for (u32 k = lane; k < rep; k += 32) {
        u32 j = i + k;
        if (j < HLIT) { if (j < 288) S->lens[j] = (u8)val; }
        else if (j - HLIT < 64) S->lens[288 + j - HLIT] = (u8)val;
      }
    }
    i += rep;
  }
  __syncwarp();
  return true;
}

// Decodes one segment on one warp.  Returns through *res (lane 0 writes).
__device__ __forceinline__ void inf_segment(InfWarpSmem *S, const u8 *in, u64 n, u64 in_pos, u8 *out, u64 out_off, u64 cap,
                                            bool stop_at_sync, InfRes *res) {
  const u32 lane = lane_id();
  InfReader r;
  r.init(in, n, in_pos);
  const u64 obase = out_off;
  u64 opos = out_off;
  u32 status = 0, flags = 0;
  u64 end_pos = 0;

  const u64 nbits = n << 3;
  u32 is_end = 0, stale = 0;  // BitReadStream.isEnd and the bit it leaves behind, see inf_coded
  for (;;) {  // blocks, /root/reference/src/inflate.ts:22-37
    r.refill();
    if (r.bitpos() > nbits + (1u << 20)) { status = SEG_E_LACK; break; }  // cannot happen (see inf_coded); bounds the loop
    const u32 bfinal = r.take(1);
    const u32 btype = r.take(2);
    if (btype == 3) { status = SEG_E_BTYPE3; break; }
    if (btype == 0) {  // stored, src/inflate.ts:42-55: readRange only — bytes past the end are zeros, never an error
      r.skip((u32)((0 - r.bitpos()) & 7));
      r.refill();
      const u32 LEN = r.take(16);
      r.refill();
      const u32 NLEN = r.take(16);
      if (LEN + NLEN != 65535) { status = SEG_E_CORRUPT; break; }
      const u64 q = r.bitpos() >> 3;
      for (u32 i = lane; i < LEN; i += 32) {
        u8 v = (q + i < n) ? in[q + i] : (u8)0;
        if (opos + i < cap) out[opos + i] = v;
      }
      opos += LEN;
      if (bfinal) { status = SEG_FINAL; end_pos = q + LEN; break; }
      if (is_end) { status = SEG_E_INSUFF; break; }  // src/inflate.ts:34-36
      if (LEN == 0 && stop_at_sync) { status = SEG_SYNC; end_pos = q; break; }
      r.init(in, n, q + LEN);
      continue;
    }
    if (btype == 1) {  // fixed, src/huffman.ts:41-53, src/inflate.ts:57-118
      // (the distance is readRangeCoded(5), src/inflate.ts:107: 32 codes of 5 bits)
      for (u32 i = lane; i < 352; i += 32) S->lens[i] = (u8)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : i < 320 ? 5 : 0);
      __syncwarp();
    } else {  // dynamic header, src/inflate.ts:121-202
      if (!inf_read_dynamic_header(r, S, nbits, is_end, stale, status)) break;
    }
    inf_build(S->lens, 288, LL_ROOT, S->lut_ll, S->sorted_ll, &S->tab_ll, S->cur);
    inf_build(S->lens + 288, 64, D_ROOT, S->lut_d, S->sorted_d, &S->tab_d, S->cur);

    // symbol loop, src/inflate.ts:76-117 / 237-291: `while (!stream.isEnd)`
    while (!is_end) {
      r.refill();
      u32 sym;
      u32 rc = inf_coded(r, S->lut_ll, LL_ROOT, &S->tab_ll, S->sorted_ll, nbits, is_end, stale, sym);
      if (rc) { status = rc; break; }
      if (sym < 256) {
        if (lane == 0 && opos < cap) out[opos] = (u8)sym;
        opos++;
        continue;
      }
      if (sym == 256) break;
      // Length codes 29/30 (symbols 286/287) and distance codes >= 30 index past the ends of the reference's tables
      // (src/const.ts:9-31): base and extra-bit count are `undefined` there, `0 < undefined` is false (no extra bits
      // are read), an undefined length makes the copy loop `i < undefined` do nothing — the distance is still decoded —
      // and an undefined distance makes the source index NaN, so `len` zeros are written (src/inflate.ts:101-116,
      // 260-290).  Neither is an error.
      const u32 ls = sym - 257;
      const bool len_def = ls < 29;
      const u32 len = len_def ? c_len_base[ls] + inf_extra(r, c_len_extra[ls], stale) : 0;
      r.refill();
      u32 ds;
      rc = inf_coded(r, S->lut_d, D_ROOT, &S->tab_d, S->sorted_d, nbits, is_end, stale, ds);
      if (rc) { status = rc; break; }
      r.refill();
      const bool dist_def = ds < 30;
      const u32 dist = dist_def ? c_dist_base[ds] + inf_extra(r, c_dist_extra[ds], stale) : 0;
      if (!len_def) continue;
      if (!dist_def) {
        for (u32 i = lane; i < len; i += 32)
          if (opos + i < cap) out[opos + i] = 0;
        opos += len;
        __syncwarp();
        continue;
      }
      // back-reference copy, src/inflate.ts:287-290; bytes before the start read as 0
      const long long src = (long long)opos - (long long)dist;
      if (src < (long long)obase) flags |= SEGF_HISTORY;
      __syncwarp();
      if (dist >= len) {
        for (u32 i = lane; i < len; i += 32) {
          long long s = src + i;
          u8 v = (s >= (long long)obase && (u64)s < cap) ? out[s] : (u8)0;
          if (opos + i < cap) out[opos + i] = v;
        }
      } else {  // overlapping: the output is periodic with period dist
        for (u32 i = lane; i < len; i += 32) {
          long long s = src + (i % dist);
          u8 v = (s >= (long long)obase && (u64)s < cap) ? out[s] : (u8)0;
          if (opos + i < cap) out[opos + i] = v;
        }
      }
      opos += len;
      __syncwarp();
    }
    if (status) break;
    if (bfinal) { status = SEG_FINAL; end_pos = umin64((r.bitpos() + 7) >> 3, n); break; }
    if (is_end) { status = SEG_E_INSUFF; break; }  // src/inflate.ts:34-36
  }
  if (opos > cap) flags |= SEGF_OVERFLOW;
  if (lane == 0) {
    res->end_pos = end_pos;
    res->out_len = opos - obase;
    res->status = status;
    res->flags = flags;
  }
}

// Grid of persistent warps; each takes segment indices from a global counter.
// in_pos[j] = start byte of segment j; out_off == nullptr means "segment j writes
// at j * CHUNK" (the optimistic placement that is right for our own streams).
__global__ void __launch_bounds__(INF_THREADS)
k_inflate(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ in_pos, const u64 *__restrict__ out_off,
          const u32 *__restrict__ nseg_ptr, u32 nseg_cap, u8 *out, u64 cap, int stop_at_sync, InfRes *res, u32 *counter) {
  ZLES_SMEM_DECL(smem_raw);
  InfWarpSmem *S = reinterpret_cast<InfWarpSmem *>(smem_raw) + warp_id();
  const u32 nseg = umin(*nseg_ptr, nseg_cap);
  for (;;) {
    u32 j = 0;
    if (lane_id() == 0) j = atomicAdd(counter, 1u);
    j = __shfl_sync(ZLES_FULL, j, 0);
    if (j >= nseg) break;
    u64 oo = out_off ? out_off[j] : (u64)j * CHUNK;
    inf_segment(S, in, n, in_pos[j], out, oo, cap, stop_at_sync != 0, res + j);
    __syncwarp();
  }
}

// ---- sync-marker scan -------------------------------------------------------
// A candidate segment start is a byte position c with in[c-4..c) == 00 00 FF FF.
// Pass 1 counts candidates per 4 KiB tile; pass 2 (after a scan of the counts)
// writes their positions in ascending order.  cand[0] is always `first`.
constexpr int MARK_THREADS = 256;
constexpr u32 MARK_TILE = MARK_THREADS * 16;

__device__ __forceinline__ u32 mark_mask16(const u8 *__restrict__ in, u64 n, u64 first, u64 v) {
  // bit k set <=> position c = 16 v + k is a candidate
  u64 c0 = v << 4;
  if (c0 >= n) return 0;
  u32 m = 0;
  if (c0 >= 4 && c0 + 16 <= n && (((uintptr_t)in & 3) == 0)) {
    const u32 *w = reinterpret_cast<const u32 *>(in + c0 - 4);
    u32 W[5];
#pragma unroll
    for (int k = 0; k < 5; k++) W[k] = __ldg(w + k);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      u32 x = (k & 3) ? __funnelshift_r(W[k >> 2], W[(k >> 2) + 1], (k & 3) * 8) : W[k >> 2];
      if (x == 0xFFFF0000u) m |= 1u << k;
    }
  } else {
    for (int k = 0; k < 16; k++) {
      u64 c = c0 + k;
      if (c >= 4 && c < n && in[c - 4] == 0 && in[c - 3] == 0 && in[c - 2] == 0xFF && in[c - 1] == 0xFF) m |= 1u << k;
    }
  }
  // a candidate must leave at least one byte to decode and lie after the first segment start
  for (int k = 0; k < 16; k++)
    if ((m >> k) & 1) {
      u64 c = c0 + k;
      if (c <= first || c >= n) m &= ~(1u << k);
    }
  return m;
}

__global__ void __launch_bounds__(MARK_THREADS) k_mark_count(const u8 *__restrict__ in, u64 n, u64 first, u32 *tile_cnt) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u64 v = (u64)blockIdx.x * MARK_THREADS + threadIdx.x;
  u32 c = (u32)__popc(mark_mask16(in, n, first, v));
  u32 total;
  block_exscan(c, scratch, &total);
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// single CTA: exclusive scan of tile counts in place; writes *ncand = 1 + total (cand[0] = first)
__global__ void __launch_bounds__(1024) k_mark_scan(u32 *tile_cnt, u32 ntiles, u32 *ncand) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u32 carry = 1;
  for (u32 base = 0; base < ntiles; base += 1024) {
    u32 i = base + threadIdx.x;
    u32 c = i < ntiles ? tile_cnt[i] : 0;
    u32 total;
    u32 ex = block_exscan(c, scratch, &total);
    if (i < ntiles) tile_cnt[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *ncand = carry;
}

__global__ void __launch_bounds__(MARK_THREADS)
k_mark_emit(const u8 *__restrict__ in, u64 n, u64 first, const u32 *tile_base, u64 *cand, u32 cand_cap) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u64 v = (u64)blockIdx.x * MARK_THREADS + threadIdx.x;
  u32 m = mark_mask16(in, n, first, v);
  u32 total;
  u32 ex = block_exscan((u32)__popc(m), scratch, &total);
  u32 o = tile_base[blockIdx.x] + ex;
  while (m) {
    int k = __ffs((int)m) - 1;
    m &= m - 1;
    if (o < cand_cap) cand[o] = (v << 4) + k;
    o++;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) cand[0] = first;
}

// ---- K7 fast path for OUR streams: two phases ------------------------------------------------
// Our encoder ends every 32 KiB block with a sync marker and a block only references data of
// its own 128 KiB chunk.  Huffman decoding is a serial chain per block but independent between
// blocks; the LZ77 copies depend on earlier output but are cheap.  So:
//   phase A  k_inf_tokens   one warp per candidate segment: decode the symbols into a token
//                           list (literal byte | match length + distance), no output yet.
//                           Where a segment ends and how many bytes it stands for do not
//                           depend on anything outside it, so false candidates are harmless.
//   (check)  k_inf_check    every candidate must end where the next one starts and stand for
//                           exactly SUB bytes; otherwise the host filters the list (zles.cu).
//   phase B  k_inf_resolve  one warp per 128 KiB chunk (4 segments, in order): a warp-wide scan
//                           of the token lengths gives every token its output offset, literals
//                           are stored at once, and matches are copied lane-parallel, a batch
//                           of 32 tokens at a time, in rounds that respect their dependencies.
// Tokens use the encoder's format (lz77.cuh): literal = byte value; match = bit31 | (len-3)<<16 | (dist-1).
constexpr u32 SEGF_BADREF = 4;            // a distance reached before the start of the chunk
constexpr u32 SEGF_STORED = 8;            // the segment's data block is a stored one: no tokens, phase B copies the bytes
constexpr u32 SEGF_TAIL_SHIFT = 8;        // flags bits 8..15 of a stored segment: bytes between its payload's end and end_pos
// where a stored segment's payload starts in the input, or ~0 when the result is inconsistent
__host__ __device__ __forceinline__ u64 seg_stored_src(u64 end_pos, u64 out_len, u32 flags) {
  const u64 back = out_len + ((flags >> SEGF_TAIL_SHIFT) & 0xffu);
  return end_pos >= back ? end_pos - back : ~0ull;
}

// Phase A uses 32-bit LUT entries so that one shared-memory load yields everything about a symbol:
//   bits 0-3 code length (0 = longer than the root: canonical slow path) | bits 4-7 extra-bit count |
//   bits 8-23 literal value / length base / distance base | bit 24 length | bit 25 end of block | bit 26 invalid
constexpr u32 TK_LEN = 1u << 24, TK_EOB = 1u << 25, TK_INV = 1u << 26;

// What a token decoder keeps per warp besides its two 32-bit LUTs: the canonical arrays (the slow path for codes longer
// than the root tables).  What only lives while a header is parsed and the tables are built shares the LUTs' room —
// the code-length code's LUT and lengths sit where lut_ll is filled afterwards, the literal/length and distance code
// lengths where lut_d is filled last (tk_build_tables reads them before that): 6,112 bytes per warp instead of 6,752,
// which is what lets NINE four-warp CTAs of k_inf_tokens share an SM's 228 KiB instead of eight.
struct TokCore {
  u16 sorted_ll[288];
  u16 sorted_d[64];
  InfTab tab_ll, tab_d;
  u16 cur[16];
};
struct TokWarpSmem {
  union {
    u32 lut_ll[1 << LL_ROOT];
    struct {
      u16 cl_lut[1 << CL_ROOT];  // the code-length code's LUT while a header is parsed
      u8 cl_lens[32];
    };
  };
  union {
    u32 lut_d[1 << D_ROOT];
    u8 lens[352];  // [0,288) literal/length code lengths, [288,352) distance code lengths
  };
  TokCore w;
};
static_assert(sizeof(TokWarpSmem) == 4096 + 1024 + sizeof(TokCore), "header-time members must fit inside the LUTs");
constexpr int TOK_SMEM = (int)sizeof(TokWarpSmem) * INF_WARPS;

__device__ __forceinline__ u32 tk_entry_ll(u32 sym, u32 l) {
  if (sym < 256) return (sym << 8) | l;
  if (sym == 256) return TK_EOB | l;
  if (sym >= 286) return TK_INV | l;
  return TK_LEN | ((u32)c_len_base[sym - 257] << 8) | ((u32)c_len_extra[sym - 257] << 4) | l;
}
__device__ __forceinline__ u32 tk_entry_d(u32 sym, u32 l) {
  if (sym >= 30) return TK_INV | l;
  return ((u32)c_dist_base[sym] << 8) | ((u32)c_dist_extra[sym] << 4) | l;
}
// both codes of a block from T->w.lens: canonical arrays and the 32-bit root LUTs.  Symbols the reference's tables do not
// define (286, 287, distance codes 30..) are left out of the fast tables: they take the slow path, which hands the block
// to the sequential decoder.
__device__ __forceinline__ void tk_build_tables(TokWarpSmem *T) {
  TokCore *S = &T->w;
  inf_canon(T->lens, 288, S->sorted_ll, &S->tab_ll, S->cur);
  for (u32 i = lane_id(); i < (1u << LL_ROOT); i += 32) {
    const u32 e = inf_lut_entry(i, LL_ROOT, &S->tab_ll, S->sorted_ll);
    T->lut_ll[i] = ((e & 15) && (e >> 4) < 286) ? tk_entry_ll(e >> 4, e & 15) : 0;
  }
  __syncwarp();
  inf_canon(T->lens + 288, 32, S->sorted_d, &S->tab_d, S->cur);
  __syncwarp();  // the lengths have been read: lut_d takes their room
  for (u32 i = lane_id(); i < (1u << D_ROOT); i += 32) {
    const u32 e = inf_lut_entry(i, D_ROOT, &S->tab_d, S->sorted_d);
    T->lut_d[i] = ((e & 15) && (e >> 4) < 30) ? tk_entry_d(e >> 4, e & 15) : 0;
  }
  __syncwarp();
}

// warp-uniform bit reader for phase A: the bit buffer is kept as two 32-bit halves so that every
// shift is one funnel-shift instruction (no shift here exceeds 31 bits); words are handed out from
// a 128-byte window per warp.
struct TokReader {
  const u8 *in;
  u64 n;
  const u32 *words;   // `in` rounded down to 4 bytes
  u32 skew;
  u64 wabs;           // absolute word index of window position 0
  u32 wi;             // next word to consume, relative to wabs (the live window is [wi & ~31, +32))
  u32 win, win_next;
  u32 lo, hi;         // bit buffer: bits 0..31 and 32..63 of the unread stream
  u32 bc;             // valid bits in (hi:lo)

  __device__ __forceinline__ u32 load_word(u64 w) const {  // bytes outside [0, n) read as 0 (BitReadStream.ts:33-35)
    const long long l0 = (long long)(w << 2) - (long long)skew;
    if (l0 >= 0 && (u64)l0 + 4 <= n) return __ldg(words + w);
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long b = l0 + k;
      if (b >= 0 && (u64)b < n) v |= (u32)in[b] << (8 * k);
    }
    return v;
  }
  __device__ __forceinline__ u32 next_word() {
    const u32 w = __shfl_sync(ZLES_FULL, win, (int)(wi & 31));
    wi++;
    if ((wi & 31) == 0) {
      win = win_next;
      win_next = load_word(wabs + wi + 32 + lane_id());
    }
    return w;
  }
  __device__ __forceinline__ void init(const u8 *in_, u64 n_, u64 byte_pos) {
    in = in_; n = n_;
    skew = (u32)((uintptr_t)in_ & 3);
    words = reinterpret_cast<const u32 *>(in_ - skew);
    const u64 a = byte_pos + skew;
    const u64 w0 = a >> 2;
    wabs = w0 & ~(u64)31;
    wi = (u32)(w0 - wabs);
    win = load_word(wabs + lane_id());
    win_next = load_word(wabs + 32 + lane_id());
    const u32 sh = (u32)(a & 3) * 8;
    lo = next_word() >> sh;
    hi = 0;
    bc = 32 - sh;
    refill();
  }
  __device__ __forceinline__ void refill() {  // afterwards bc >= 33
    if (bc <= 32) {
      const u32 w = next_word();
      lo |= __funnelshift_lc(0u, w, bc);   // w << bc (0 when bc == 32)
      hi = __funnelshift_lc(w, 0u, bc);    // w >> (32 - bc) (w when bc == 32, 0 when bc == 0)
      bc += 32;
    }
  }
  __device__ __forceinline__ u32 peek(u32 k) const { return lo & ~(0xffffffffu << k); }  // k <= 31
  __device__ __forceinline__ void skip(u32 k) {                                           // k <= 31
    lo = __funnelshift_r(lo, hi, k);
    hi >>= k;
    bc -= k;
  }
  __device__ __forceinline__ u32 take(u32 k) { const u32 v = peek(k); skip(k); return v; }
  __device__ __forceinline__ u64 bits64() const { return ((u64)hi << 32) | lo; }
  __device__ __forceinline__ u64 bitpos() const { return ((wabs + wi) << 5) - bc - ((u64)skew << 3); }
  __device__ __forceinline__ bool past_end() const { return bitpos() > (n << 3); }  // consumed bits the buffer does not have
};

// Reader of the symbol loop of phase A.  The warp holds 256 bytes of the stream, one word (two) per lane; lane i looks
// at the stream at bit offset P + i, so the 32 lanes decode the tokens that would start at the next 32 bit positions
// all at once (inf_segment_tokens, inflate_spec.cuh).
struct SpecReader : TokReader {
  u32 P;  // bit offset of the current position from word `wabs` (lane l holds word wabs + l in win, wabs + 32 + l in win_next)

  __device__ __forceinline__ void begin(const u8 *in_, u64 n_, u64 bit_pos) {  // bit_pos relative to in_
    in = in_; n = n_;
    skew = (u32)((uintptr_t)in_ & 3);
    words = reinterpret_cast<const u32 *>(in_ - skew);
    const u64 a = bit_pos + ((u64)skew << 3);
    wabs = (a >> 5) & ~(u64)31;
    P = (u32)(a - (wabs << 5));
    win = load_word(wabs + lane_id());
    win_next = load_word(wabs + 32 + lane_id());
  }
  __device__ __forceinline__ u64 pos() const { return (wabs << 5) + P - ((u64)skew << 3); }
  __device__ __forceinline__ void advance(u32 c) {  // warp-uniform, c < 1024
    P += c;
    if (P >= 1024) {  // the first window is used up: slide (once per 128 bytes of stream)
      win = win_next;
      wabs += 32;
      win_next = load_word(wabs + 32 + lane_id());
      P -= 1024;
    }
  }
  __device__ __forceinline__ u32 word(u32 idx) const {  // idx warp-uniform, < 64
    return __shfl_sync(ZLES_FULL, idx < 32 ? win : win_next, (int)(idx & 31));
  }
  // the 64 bits at offset P + s, s <= 31: they lie in the four words from P's word on, which every lane gets by
  // shuffle (no per-round shifting of a register window, no loop)
  __device__ __forceinline__ void at(u32 s, u32 &lo, u32 &hi) const {
    const u32 j0 = P >> 5;
    const u32 X0 = word(j0), X1 = word(j0 + 1), X2 = word(j0 + 2), X3 = word(j0 + 3);
    const u32 t = (P & 31) + s;
    const bool up = t >= 32;
    const u32 a = up ? X1 : X0, b = up ? X2 : X1, c = up ? X3 : X2;
    lo = __funnelshift_r(a, b, t);  // shift taken modulo 32
    hi = __funnelshift_r(b, c, t);
  }
};

// dynamic block header for phase A (same as inf_read_dynamic_header, on the TokReader)
__device__ __forceinline__ bool tk_read_dynamic_header(TokReader &r, TokWarpSmem *T, u32 &status) {
  TokCore *S = &T->w;
  const u32 lane = lane_id();
  const u32 HLIT = r.take(5) + 257;
  const u32 HDIST = r.take(5) + 1;
  const u32 HCLEN = r.take(4) + 4;
  T->cl_lens[lane] = 0;
  for (u32 i = lane; i < 352; i += 32) T->lens[i] = 0;
  __syncwarp();
  for (u32 i = 0; i < HCLEN; i++) {
    r.refill();
    const u32 v = r.take(3);
    if (lane == 0) T->cl_lens[c_cl_order[i]] = (u8)v;
  }
  __syncwarp();
  inf_build(T->cl_lens, 32, CL_ROOT, T->cl_lut, S->sorted_d, &S->tab_d, S->cur);
  const u32 total = HLIT + HDIST;
  u32 prev = 0;
  for (u32 i = 0; i < total;) {
    r.refill();
    if (r.past_end()) { status = SEG_E_LACK; return false; }
    u32 e = T->cl_lut[r.peek(CL_ROOT)], l = e & 15, sym = e >> 4;
    if (l == 0 && !inf_slow(r.bits64(), &S->tab_d, S->sorted_d, sym, l)) { status = SEG_E_CORRUPT; return false; }
    r.skip(l);
    u32 rep = 1, val = sym;
    if (sym == 16) { rep = 3 + r.take(2); val = prev; }
    else if (sym == 17) { rep = 3 + r.take(3); val = 0; prev = 0; }
    else if (sym == 18) { rep = 11 + r.take(7); val = 0; prev = 0; }
    else prev = sym;
    if (val) {
      for (u32 k = lane; k < rep; k += 32) {
        const u32 j = i + k;
        if (j < HLIT) { if (j < 288) T->lens[j] = (u8)val; }
        else if (j - HLIT < 32) T->lens[288 + j - HLIT] = (u8)val;
      }
      // a run that gives distance symbols 32.. a length (it spills past HDIST = 32): the reference keeps those codes
      // (src/inflate.ts:187-200); the parallel tiers do not model them — such a block is left to the sequential decoder
      if (i + rep > HLIT + 32) { status = SEG_E_CORRUPT; return false; }
    }
    i += rep;
  }
  __syncwarp();
  return true;
}

// profiling build (tools/lz_stages.py): warp totals of phase A — cycles in block headers + table builds, cycles in the
// symbol loop, rounds, tokens, tokens decoded on their own, cycles re-seating the readers
#ifdef ZLES_STAGE_CLOCKS
__device__ unsigned long long g_inf_clk[8];
#define INF_CLK(i)                                                         \
  do {                                                                     \
    if (lane_id() == 0) {                                                  \
      const long long t_ = clock64();                                      \
      atomicAdd(&g_inf_clk[i], (unsigned long long)(t_ - iclk));           \
      iclk = t_;                                                           \
    }                                                                      \
  } while (0)
#define INF_CNT(i, v) do { if (lane_id() == 0) atomicAdd(&g_inf_clk[i], (unsigned long long)(v)); } while (0)
#define INF_CLK_DECL long long iclk = clock64()
#else
#define INF_CLK(i) do { } while (0)
#define INF_CNT(i, v) do { } while (0)
#define INF_CLK_DECL do { } while (0)
#endif

// pieces (may be null): the block's tokens cut into four runs of about 8 KiB of output each — [count, bytes] x 4, the layout
// the four-warp phase A leaves (inflate_spec.cuh) — so that phase B can give every piece its own warp (k_piece_sym).
__device__ __forceinline__ void inf_segment_tokens(TokWarpSmem *T, const u8 *in, u64 n, u64 in_pos, u32 *tok, InfRes *res, u32 *ntok_out,
                                                   u32 *pieces = nullptr) {
  const u32 lane = lane_id();
  u32 cut_nt0 = 0, cut_nt1 = 0, cut_nt2 = 0, cut_o0 = 0, cut_o1 = 0, cut_o2 = 0, ncut = 0, thresh = SUB / 4;
  // a cut after the tokens decoded so far, whenever the output has passed the next quarter of a block
#define ZLES_CUT()                                                                   \
  do {                                                                               \
    while (o >= thresh && ncut < 3) {                                                \
      if (ncut == 0) { cut_nt0 = nt; cut_o0 = o; }                                   \
      else if (ncut == 1) { cut_nt1 = nt; cut_o1 = o; }                              \
      else { cut_nt2 = nt; cut_o2 = o; }                                             \
      ncut++;                                                                        \
      thresh += SUB / 4;                                                             \
    }                                                                                \
  } while (0)
  TokCore *S = &T->w;
  TokReader r;
  r.init(in, n, in_pos);
  u32 o = 0, nt = 0;
  u32 status = 0, flags = 0;
  u64 end_pos = 0, stored_end = 0;
  INF_CLK_DECL;
  // a single token, outside the parallel rounds of the symbol loop
#define ZLES_EMIT(t)                 \
  do {                               \
    if (lane == 0) tok[nt] = (t);    \
    nt++;                            \
  } while (0)
  for (;;) {
    r.refill();
    if (r.past_end()) { status = SEG_E_LACK; break; }
    const u32 bfinal = r.take(1);
    const u32 btype = r.take(2);
    if (btype == 3) { status = SEG_E_BTYPE3; break; }
    if (btype == 0) {
      r.skip((u32)((0 - r.bitpos()) & 7));
      r.refill();
      const u32 LEN = r.take(16);
      r.refill();
      const u32 NLEN = r.take(16);
      if (LEN + NLEN != 65535) { status = SEG_E_CORRUPT; break; }
      const u64 q = r.bitpos() >> 3;
      if (LEN != 0) {
        // a stored data block: ours only as the one data block of its segment (incompressible input, see k_huff);
        // phase B copies the bytes [q, q + LEN) from the input.  Anything else takes the sequential path.
        if (o != 0 || nt != 0 || (flags & SEGF_STORED) || LEN > SUB || q + LEN > n) { status = SEG_E_CORRUPT; break; }
        flags |= SEGF_STORED;
        o = LEN;
        stored_end = q + LEN;
        if (bfinal) { status = SEG_FINAL; end_pos = q + LEN; break; }
        r.init(in, n, q + LEN);
        continue;
      }
      end_pos = q;
      status = bfinal ? SEG_FINAL : SEG_SYNC;
      break;
    }
    if (flags & SEGF_STORED) { status = SEG_E_CORRUPT; break; }  // a coded block after stored data: not one of ours
    if (btype == 1) {
      for (u32 i = lane; i < 320; i += 32) T->lens[i] = (u8)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
      __syncwarp();
    } else {
      if (!tk_read_dynamic_header(r, T, status)) break;
    }
    tk_build_tables(T);
    // symbol loop (/root/reference/src/inflate.ts:237-291 without the copy).  It cannot run away: every
    // token accounts for at least one output byte and a token is only emitted while o < SUB.
    // Every round, lane i decodes the token that would start i bits from the current position — both table
    // look-ups, extra bits and all — and the warp then follows the chain from lane 0: a token of b bits at lane c
    // is followed by the one at lane c + b.  Two or three tokens per round on text, for the instructions of one.
    {
      INF_CLK(0);
      SpecReader sr;
      sr.begin(in, n, r.bitpos());
      INF_CLK(5);
      bool eob = false;
      while (!eob && !status) {
        INF_CNT(2, 1);
        u32 lo, hi;
        sr.at(lane, lo, hi);
        const u32 e = T->lut_ll[lo & ((1u << LL_ROOT) - 1)];
        const u32 l1 = e & 15, eb = (e >> 4) & 15, sh2 = l1 + eb;
        const u32 len = ((e >> 8) & 0xffff) + ((lo >> l1) & ~(0xffffffffu << eb));  // a literal's value when eb == 0
        const u32 y = __funnelshift_r(lo, hi, sh2);                                 // the stream after the length's extra bits
        const u32 d = T->lut_d[y & ((1u << D_ROOT) - 1)];
        const u32 l2 = d & 15, db = (d >> 4) & 15;
        const u32 dist = (d >> 8) + ((y >> l2) & ~(0xffffffffu << db));
        const bool is_len = (e & TK_LEN) != 0;
        // pack: bits the token takes (0 = a code the fast tables do not hold: decoded on its own below) | bit 8 end of block
        u32 pack = is_len ? (l2 ? sh2 + l2 + db : 0) : l1;
        if (l1 == 0) pack = 0;
        if (e & TK_EOB) pack |= 0x100;
        const u32 tokv = is_len ? (0x80000000u | ((len - 3) << 16) | (dist - 1)) : (len & 0xff);
        // The chain: lane 0 is a token start; a token of nb bits at lane c makes lane c + nb one.  A token that
        // ends the round (end of block, or a code the fast tables do not hold) leads nowhere.
        // The chain: lane 0 is a token start; a token of nb bits at lane c makes lane c + nb one — one shuffle per
        // token.  A token that ends the round (end of block, or a code the fast tables do not hold) leads nowhere.
        u32 R = 1, last = 0, plast;
        bool last_stops;
        for (;;) {
          plast = __shfl_sync(ZLES_FULL, pack, (int)last);
          const u32 nb = plast & 0xff;
          last_stops = nb == 0 || (plast & 0x100);
          if (last_stops || last + nb >= 32) break;
          last += nb;
          R |= 1u << last;
        }
        const u32 plain = last_stops ? R & ~(1u << last) : R;              // the members that are ordinary tokens
        const bool mine = (plain >> lane) & 1;
        const u32 sum = __reduce_add_sync(ZLES_FULL, mine ? (is_len ? len : 1u) : 0u);
        u32 cur = 0;
        bool slow = false;
        if (o + sum <= SUB) {
          // every token of the round starts below SUB (what the reference checks token by token)
          if (mine) tok[nt + (u32)__popc(plain & lanemask_lt())] = tokv;
          nt += (u32)__popc(plain);
          o += sum;
          ZLES_CUT();
          if (!last_stops) cur = last + (plast & 0xff);
          else if (plast & 0x100) {
            cur = last + (plast & 0xff);
            eob = true;
          } else {
            cur = last;
            slow = true;
          }
        } else {
          // would pass SUB (not one of our blocks): token by token, to stop exactly where the reference's check does
          for (;;) {
            const u32 pk = __shfl_sync(ZLES_FULL, pack, (int)cur);
            const u32 kb = pk & 0xff;
            if (kb == 0) { slow = true; break; }
            if (pk & 0x100) { cur += kb; eob = true; break; }
            if (o >= SUB) { status = SEG_E_CORRUPT; flags |= SEGF_OVERFLOW; break; }
            const u32 t = __shfl_sync(ZLES_FULL, tokv, (int)cur);
            ZLES_EMIT(t);
            o += (t >> 31) ? ((t >> 16) & 0x1ff) + 3 : 1;
            cur += kb;
            if (cur >= 32) break;
          }
        }
        sr.advance(cur);
        if (slow) {  // one token with a code longer than the root table (or an invalid one), the canonical way
          INF_CNT(4, 1);
          sr.at(0, lo, hi);
          u32 e2 = T->lut_ll[lo & ((1u << LL_ROOT) - 1)];
          if ((e2 & 15) == 0) {
            u32 sym, l;
            if (!inf_slow(((u64)hi << 32) | lo, &S->tab_ll, S->sorted_ll, sym, l)) { status = SEG_E_CORRUPT; break; }
            e2 = tk_entry_ll(sym, l);
            if (e2 & TK_INV) { status = SEG_E_CORRUPT; break; }
          }
          if (e2 & TK_EOB) { sr.advance(e2 & 15); eob = true; break; }
          if (o >= SUB) { status = SEG_E_CORRUPT; flags |= SEGF_OVERFLOW; break; }
          if (e2 < TK_LEN) {
            sr.advance(e2 & 15);
            ZLES_EMIT(e2 >> 8);
            o++;
            continue;
          }
          const u32 eb2 = (e2 >> 4) & 15;
          const u32 len2 = ((e2 >> 8) & 0xffff) + ((lo >> (e2 & 15)) & ~(0xffffffffu << eb2));  // code <= 15 bits, extra <= 5
          sr.advance((e2 & 15) + eb2);
          sr.at(0, lo, hi);
          u32 d2 = T->lut_d[lo & ((1u << D_ROOT) - 1)];
          if ((d2 & 15) == 0) {
            u32 sym, l;
            if (!inf_slow(((u64)hi << 32) | lo, &S->tab_d, S->sorted_d, sym, l)) { status = SEG_E_CORRUPT; break; }
            d2 = tk_entry_d(sym, l);
            if (d2 & TK_INV) { status = SEG_E_CORRUPT; break; }
          }
          const u32 db2 = (d2 >> 4) & 15;
          const u32 dist2 = (d2 >> 8) + ((lo >> (d2 & 15)) & ~(0xffffffffu << db2));           // code <= 15 bits, extra <= 13
          sr.advance((d2 & 15) + db2);
          ZLES_EMIT(0x80000000u | ((len2 - 3) << 16) | (dist2 - 1));
          o += len2;
        }
      }
      // back to the block-level reader at the position the symbol loop stopped
      INF_CLK(1);
      const u64 pos = sr.pos();
      r.init(in, n, pos >> 3);
      r.skip((u32)(pos & 7));
      INF_CLK(5);
    }
    if (o > SUB) { status = SEG_E_CORRUPT; flags |= SEGF_OVERFLOW; }
    if (status) break;
    if (r.past_end()) { status = SEG_E_LACK; break; }
    if (bfinal) { status = SEG_FINAL; end_pos = (r.bitpos() + 7) >> 3; break; }
  }
#undef ZLES_EMIT
  ZLES_CUT();
#undef ZLES_CUT
  if (pieces && lane == 0) {
    // cuts that were never reached leave empty pieces; the last piece takes the rest
    const u32 n0 = ncut > 0 ? cut_nt0 : nt, o0 = ncut > 0 ? cut_o0 : o;
    const u32 n1 = ncut > 1 ? cut_nt1 : nt, o1 = ncut > 1 ? cut_o1 : o;
    const u32 n2 = ncut > 2 ? cut_nt2 : nt, o2 = ncut > 2 ? cut_o2 : o;
    pieces[0] = n0; pieces[1] = o0;
    pieces[2] = n1 - n0; pieces[3] = o1 - o0;
    pieces[4] = n2 - n1; pieces[5] = o2 - o1;
    pieces[6] = nt - n2; pieces[7] = o - o2;
  }
  INF_CLK(0);
  INF_CNT(3, nt);
  // a stored segment: how far before end_pos its payload ends (0: the data block was the final one; 5: an empty stored
  // block — the marker, or a final one as system zlib writes after a full-size stored block — follows it)
  if ((flags & SEGF_STORED) && end_pos >= stored_end) flags |= ((u32)umin64(end_pos - stored_end, 255) & 0xffu) << SEGF_TAIL_SHIFT;
  if (lane == 0) {
    res->end_pos = end_pos;
    res->out_len = o;
    res->status = status;
    res->flags = flags;
    *ntok_out = nt;
  }
}

// phase A: persistent warps take segment indices from a counter.  Segment j's tokens go to
// tokens[j * SUB ...]; candidates beyond tok_segs (more than the output could hold) are not decoded.
__global__ void __launch_bounds__(INF_THREADS)
k_inf_tokens(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ seg_pos, u32 nseg, u32 *tokens, u32 *ntok, InfRes *res,
             u32 *pinfo, u32 *counter) {
  ZLES_SMEM_DECL(smem_raw);
  TokWarpSmem *T = reinterpret_cast<TokWarpSmem *>(smem_raw) + warp_id();
  for (;;) {
    u32 j = 0;
    if (lane_id() == 0) j = atomicAdd(counter, 1u);
    j = __shfl_sync(ZLES_FULL, j, 0);
    if (j >= nseg) break;
    inf_segment_tokens(T, in, n, seg_pos[j], tokens + (size_t)j * SUB, res + j, ntok + j, pinfo ? pinfo + (size_t)j * 8 : nullptr);
    __syncwarp();
  }
}

// phase B: one warp per chunk.  seg_list (or identity when null) names the real segments in
// stream order; chunk c is made of entries [4c, 4c+4) and is written at out + c * CHUNK.
// Per batch of 32 tokens: a warp scan of the lengths gives the output offsets; literals are
// stored at once; matches whose source lies entirely before the batch ("independent", the
// common case on text) are copied lane-parallel with all loads issued before the stores; the
// others — and long ones — are copied in token order by the whole warp, reading from a 16 KiB
// shared-memory ring that mirrors the warp's most recent output, so that chains of near
// references (structured data) run at shared-memory latency instead of one HBM/L2 round trip each.
constexpr int RES_WARPS = 4;
constexpr int RES_THREADS = RES_WARPS * 32;
constexpr u32 RES_LONG = 17;         // matches at least this long are copied by the whole warp
constexpr u32 RES_RING = 16384;      // >= 2 x the most a batch can produce (32 x 258 = 8256)
constexpr int RES_SMEM = (int)(RES_WARPS * RES_RING);

// State of one warp of phase B: it writes a run of segments/blocks one after the other at `base`.
struct ResState {
  u8 *base;     // where the run's first byte goes
  u8 *ring;     // RES_RING bytes of shared memory mirroring the latest output
  u32 room;     // bytes of the run that fit in the output buffer
  u32 limit;    // the run may not decode to more than this
  u32 o;        // bytes written so far (run-relative)
  u32 bad;      // problem bits: 1 = not a stream this path handles, 2 = output buffer too small
};

// Appends the output of `nt` tokens to the run.  Warp-uniform; returns false when st.bad was set.
__device__ __forceinline__ bool res_tokens(ResState &st, const u32 *__restrict__ tok, u32 nt) {
  constexpr u32 RM = RES_RING - 1;
  const u32 lane = lane_id();
  u8 *base = st.base, *ring = st.ring;
  u32 o = st.o;
  u32 tnext = lane < nt ? __ldg(tok + lane) : 0;
  for (u32 b0 = 0; b0 < nt; b0 += 32) {
    const bool valid = b0 + lane < nt;
    const u32 t = tnext;
    tnext = b0 + 32 + lane < nt ? __ldg(tok + b0 + 32 + lane) : 0;  // next batch's tokens are in flight during this one
    const bool isM = valid && (t >> 31);
    const u32 len = !valid ? 0 : (isM ? ((t >> 16) & 255) + 3 : 1);
    const u32 dist = (t & 0x7fff) + 1;
    u32 inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 v = __shfl_up_sync(ZLES_FULL, inc, d);
      if (lane >= (u32)d) inc += v;
    }
    const u32 pos = o + inc - len;            // run-relative offset of this token's first byte
    const u32 total = __shfl_sync(ZLES_FULL, inc, 31);
    if (total > st.limit - o) { st.bad |= 1; break; }
    if (o + total > st.room) { st.bad |= 2; break; }  // output buffer too small: nothing of this batch is written
    if (__any_sync(ZLES_FULL, isM && dist > pos)) { st.bad |= 1; break; }  // a reference before the start of the run
    if (valid && !isM) { base[pos] = (u8)t; ring[pos & RM] = (u8)t; }
    const u32 src = pos - dist;
    const bool indep = isM && len < RES_LONG && src + len <= o;  // source entirely before this batch
    const u32 later = __ballot_sync(ZLES_FULL, isM && !indep);
    if (indep) {  // all loads first, then the stores: one memory round trip for the whole batch
      u8 v[RES_LONG - 1];
#pragma unroll
      for (u32 q = 0; q < RES_LONG - 1; q++) if (q < len) v[q] = base[src + q];
#pragma unroll
      for (u32 q = 0; q < RES_LONG - 1; q++) if (q < len) { base[pos + q] = v[q]; ring[(pos + q) & RM] = v[q]; }
    }
    __syncwarp();
    u32 m = later;
    while (m) {  // dependent or long matches: in token order, the whole warp copies one at a time
      const int j = __ffs((int)m) - 1;
      m &= m - 1;
      const u32 pj = __shfl_sync(ZLES_FULL, pos, j), lj = __shfl_sync(ZLES_FULL, len, j), dj = __shfl_sync(ZLES_FULL, dist, j);
      const u32 sj = pj - dj;
      // the ring holds this warp's latest output up to the end of this batch; everything from sj on is in it
      // when o + total - sj <= RES_RING (and then nothing of it has been overwritten)
      const bool in_ring = o + total - sj <= RES_RING;
      if (dj >= lj) {
        for (u32 q = lane; q < lj; q += 32) {
          const u8 b = in_ring ? ring[(sj + q) & RM] : base[sj + q];
          base[pj + q] = b;
          ring[(pj + q) & RM] = b;
        }
      } else {  // overlapping: the output is periodic with period dj
        for (u32 q = lane; q < lj; q += 32) {
          const u8 b = in_ring ? ring[(sj + q % dj) & RM] : base[sj + q % dj];
          base[pj + q] = b;
          ring[(pj + q) & RM] = b;
        }
      }
      __syncwarp();
    }
    o += total;
  }
  st.o = o;
  return st.bad == 0;
}

// Appends `len` raw bytes (a stored block's payload) to the run.
__device__ __forceinline__ bool res_bytes(ResState &st, const u8 *__restrict__ srcp, u32 len) {
  constexpr u32 RM = RES_RING - 1;
  if (len > st.limit - st.o) { st.bad |= 1; return false; }
  if (st.o + len > st.room) { st.bad |= 2; return false; }
  for (u32 q = lane_id(); q < len; q += 32) {
    const u8 v = srcp[q];
    st.base[st.o + q] = v;
    st.ring[(st.o + q) & RM] = v;
  }
  __syncwarp();
  st.o += len;
  return true;
}

__global__ void __launch_bounds__(RES_THREADS)
k_inf_resolve(const u32 *__restrict__ tokens, const u32 *__restrict__ ntok, const u32 *__restrict__ seg_list, u32 nseg,
              const u8 *__restrict__ in, const InfRes *__restrict__ res, u8 *out, u64 cap, u32 *problems) {
  ZLES_SMEM_DECL(smem_raw);
  const u32 c = blockIdx.x * RES_WARPS + warp_id();
  if (c * SUBS_PER_CHUNK >= nseg) return;
  ResState st;
  st.base = out + (u64)c * CHUNK;
  st.ring = smem_raw + warp_id() * RES_RING;
  const u64 room64 = (u64)c * CHUNK >= cap ? 0 : cap - (u64)c * CHUNK;
  st.room = (u32)umin64(room64, (u64)CHUNK);  // bytes of this chunk that fit in the output
  st.limit = CHUNK;
  st.o = 0;
  st.bad = 0;
  for (u32 k = 0; k < SUBS_PER_CHUNK; k++) {
    const u32 e = c * SUBS_PER_CHUNK + k;
    if (e >= nseg) break;
    const u32 sidx = seg_list ? seg_list[e] : e;
    const InfRes r = res[sidx];
    // This kernel also runs optimistically on candidates that k_inf_check will turn down: a segment that did not
    // decode cleanly has no usable end position (a stored block followed by garbage leaves end_pos = 0).
    if (r.status != SEG_SYNC && r.status != SEG_FINAL) { st.bad |= 1; break; }
    if (r.flags & SEGF_STORED) {  // stored block: the bytes sit in the input right before the marker / the end
      const u32 len = (u32)r.out_len;
      const u64 src = seg_stored_src(r.end_pos, r.out_len, r.flags);
      if (src == ~0ull) { st.bad |= 1; break; }
      if (!res_bytes(st, in + src, len)) break;
      continue;
    }
    if (!res_tokens(st, tokens + (size_t)sidx * SUB, umin(ntok[sidx], SUB))) break;
  }
  if (st.bad && lane_id() == 0) atomicOr(problems, st.bad);
}

// stream too short to hold a marker: the only candidate is `first`
__global__ void k_mark_none(u64 first, u64 *cand, u32 *ncand) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { cand[0] = first; *ncand = 1; }
}

// ---- acceptance of the optimistic parallel decode -----------------------------
// problems[0] stays 0 only if every segment j ended on the marker that starts
// segment j+1 and stands for exactly SUB bytes (the last one: <= SUB and BFINAL);
// otherwise the host walks the chain (zles.cu).  total[0] = sum of out_len.
__global__ void __launch_bounds__(256)
k_inf_check(const InfRes *res, const u64 *cand, u32 nseg, u64 n, u32 has_final, u32 *problems, unsigned long long *total) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  u32 bad = 0;
  u64 len = 0;
  if (j < nseg) {
    InfRes r = res[j];
    len = r.out_len;
    bool good;
    if (j + 1 < nseg) good = r.status == SEG_SYNC && r.end_pos == cand[j + 1] && r.out_len == SUB;
    else if (has_final) good = r.status == SEG_FINAL && r.out_len <= SUB;
    else good = r.status == SEG_SYNC && r.end_pos == n && r.out_len == SUB;  // a shard that is not the stream's last
    if (!good) bad |= 1;
  }
  if (bad) atomicOr(problems, bad);
  if (len) atomicAdd(total, (unsigned long long)len);
}

// ---- batches of independent zlib streams: one warp per stream, sequential inside ----------
__global__ void __launch_bounds__(INF_THREADS)
k_inflate_batch(const u8 *__restrict__ in, const u64 *__restrict__ in_off, u32 count, u8 *out, const u64 *__restrict__ out_off,
                u64 *out_len, int32_t *status, u32 *counter, u32 *first_err) {
  ZLES_SMEM_DECL(smem_raw);
  InfWarpSmem *S = reinterpret_cast<InfWarpSmem *>(smem_raw) + warp_id();
  InfRes *s_res = reinterpret_cast<InfRes *>(smem_raw + sizeof(InfWarpSmem) * INF_WARPS);
  for (;;) {
    u32 j = 0;
    if (lane_id() == 0) j = atomicAdd(counter, 1u);
    j = __shfl_sync(ZLES_FULL, j, 0);
    if (j >= count) break;
    const u8 *sin = in + in_off[j];
    const u64 sn = in_off[j + 1] - in_off[j];
    u32 code = 0;
    u64 olen = 0;
    if (((sn ? sin[0] : 0) & 15) != 8) {
      code = 1;  // 'Not compressed by deflate', /root/reference/src/zlib.ts:13-16
    } else {
      InfRes *r = &s_res[warp_id()];
      inf_segment(S, sin, sn, 2, out, out_off[j], out_off[j + 1], false, r);
      __syncwarp();
      olen = r->out_len;
      if (r->status != SEG_FINAL) code = r->status;
      else if (r->flags & SEGF_OVERFLOW) code = 16;  // ZLES_E_OUTPUT_FULL
      __syncwarp();
    }
    if (lane_id() == 0) {
      out_len[j] = olen;
      status[j] = (int32_t)code;
      if (code) atomicMax(first_err, code);
    }
    __syncwarp();
  }
}

}  // namespace zles
