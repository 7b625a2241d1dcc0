// inflate_fblk.cuh — the block-parallel tier for streams of OTHER encoders (zlib.es itself, system zlib), second
// version: every block, dynamic or fixed, is cut into pieces that are Huffman-decoded side by side — without guessing.
//
// A piece of a block's coded bits can only be decoded from a position where a token starts, and that is only known once
// everything before it has been decoded.  Speculation (start anywhere, wait for the parse to fall into step with the
// true one: inflate_spec.cuh, pugz, rapidgzip) works on text, but a block whose codes all have about the same length
// (incompressible data: 8-bit literals) never falls into step.  Here nothing is guessed:
//   1. for every piece, the token length at EVERY bit position of the piece is computed (32 positions per round,
//      lane i decodes the token that would start at bit i: both table look-ups, a tile of 4,096 positions at a time
//      kept in shared memory) and the 64 positions a piece can be entered at (a token is at most 48 bits long) are
//      walked through it side by side, one or two chains per lane: entry -> exit, or "end-of-block code at ...", or
//      "no code";
//   2. one thread chains the pieces: the block's first token starts after its header, piece k + 1 is entered where
//      piece k was left, until the end-of-block code — wherever it is: the block's end is not known in advance, the
//      next candidate of the header scan (hint_end) only steers how the work is cut;
//   3. every piece on the chain is decoded from its true entry: token lengths again, one lane walks the chain and
//      marks the token starts in a bitmap, all lanes decode the marked tokens and store them.
// Steps 1 and 3 run a warp per piece over ALL blocks of the stream at once (k_fblk_map, k_fblk_prefix), step 2 a thread
// per block (k_fblk_chain); k_fblk_head parses the headers and builds the tables once per block.
// A block that does not end within the room its hint gave it is reported as FB_LONG and decoded again with a longer
// hint; one that does not decode at all is reported (status 0) and the stream goes to the sequential decoder, which is
// exact about the reference's behaviour on damaged input (/root/reference/src/inflate.ts:57-118 fixed blocks,
// :120-292 dynamic blocks).
//
// Phase B works on the pieces, not on blocks: k_fpiece_sym resolves every piece on its own warp into 16-bit symbols
// (a byte that comes from before the piece stays a reference into the 32 KiB before it), k_frun_merge makes the pieces
// of a run of blocks (>= 128 KiB) concrete in order — what reaches before the run stays a reference into the run's
// window — and k_win_propagate / k_sym_finalize (inflate_foreign.cuh) finish as before.  Streams whose blocks never
// reach before their own start (zlib.es) leave no window references and skip the propagation.
#pragma once
#include "inflate_foreign.cuh"
#include "inflate_spec.cuh"

namespace zles {

constexpr u32 FB_PLAN = 32;             // a short block is cut into about this many pieces
constexpr u32 FB_PIECE_TILES = 4;       // a piece is at most this many tiles: long blocks are cut into more pieces, not longer ones
constexpr int FB_WARPS = 16;
constexpr int FB_THREADS = FB_WARPS * 32;
constexpr u32 TA_TILE = 4096;           // bit positions per tile
constexpr u32 TA_ENT = 64;              // entry positions per piece (a token is at most 15 + 5 + 15 + 13 = 48 bits)
constexpr u32 TA_WORDS = TA_TILE / 32 + 4;
constexpr u32 TA_EOB = 64;              // nx[]: bits 0-5 the token's length, bit 6 "a chain stops here": the end-of-block code (with its
                                        // length), or no code at all (length 0)
constexpr u32 TA_X_EOB = 0x80000000u;   // exit of a chain: it read the end-of-block code, which ends at (x & 0x7fffffff)
constexpr u32 TA_X_BAD = 0xffffffffu;   // ... it ran into something that is no code
constexpr u32 FB_JOB_COMPACT = 1;       // leave the block's tokens contiguous at the start of its room (blocks decoded on demand)
constexpr u32 FB_JOB_CONT = 2;          // `bit` is a token start inside the block whose header and tables are at `aux`: decode on from there
constexpr u32 FB_LONG = 2;              // FbRes.status: the block did not end within the room its hint gave it; the result stands for its
                                        // tokens up to end_bit (a token start): decode on from there (FB_JOB_CONT)

struct FbJob {
  u64 bit;        // position of the block's BFINAL bit
  u64 hint_end;   // where the block probably ends (bit position); only steers how the work is cut
  u32 *tok;       // room for the block's tokens
  u32 tok_cap;    // ... fb_tok_room(hint_end - bit) of them
  u32 flags;
  u32 piece0;     // its first piece in the per-piece arrays (maps, infos; its token runs are pieces[2 * piece0 ...])
  u32 np;         // pieces planned: fb_planned_pieces(hint_end - bit)
  u32 aux;        // where its header / tables are kept between the kernels (FbMeta, TokWarpSmem)
  u32 pad;
};
struct FbItem {   // a work item of k_fblk_map / k_fblk_prefix: pieces [p0, p0 + FB_WARPS) of job `ji` of the launch
  u32 ji, p0;
};

struct FbPiece {  // piece p of a decoded block, in stream order (unused slots are zero)
  u32 tok_off;    // first token, relative to the job's `tok`
  u32 cnt;        // tokens
  u32 bytes;      // bytes they stand for
  u32 pad;
};

// how a block whose hint spans `span` bits is cut: bits per piece (whole tiles), and the token room it needs — one
// token per two bits of a piece (fewer bits per token only in degenerate blocks: those go to the sequential decoder)
__host__ __device__ __forceinline__ u32 fb_piece_bits(u64 span) {
  const u64 per = (span + FB_PLAN - 1) / FB_PLAN;
  u64 tiles = per ? (per + TA_TILE - 1) / TA_TILE : 1;
  if (tiles > FB_PIECE_TILES) tiles = FB_PIECE_TILES;
  return (u32)(tiles * TA_TILE);
}
__host__ __device__ __forceinline__ u32 fb_planned_pieces(u64 span) {
  const u32 pb = fb_piece_bits(span);
  const u64 np = (span + pb - 1) / pb;  // (the pieces start behind the header: they reach at least as far as the hint)
  return (u32)(np < 1 ? 1 : np);
}
__host__ __device__ __forceinline__ u64 fb_tok_room(u64 span) { return (u64)fb_planned_pieces(span) * (fb_piece_bits(span) / 2); }

constexpr u32 TA_MARK = 128;            // nx[]: a token of the chain being decoded starts here (ta_walk_marking)
constexpr u32 TA_CHUNK = 512;           // positions whose marked tokens are decoded together
struct TaWarp {
  u8 nx[TA_TILE];
  u32 words[TA_WORDS];
  u16 list[TA_CHUNK];
};
struct TaPiece {      // what step 1 leaves about a piece
  u32 mpos;           // where all its live chains had become one; 0 = they never did
  u32 soff;           // where the merged chain's tokens are stored in the piece's token room (what comes before them: fewer)
  u32 sexit;          // the merged chain's exit: a position, TA_X_EOB | position, or TA_X_BAD
  u32 snt, sob, sfin; // tokens / bytes of the merged chain from mpos on (stored at tokP + soff), the bit after its end-of-block code
  u32 entry;          // k_fblk_chain: where the block's chain enters the piece,
  u32 merged;         // ... whether that leads into the merged chain,
  u32 on_chain, pad;  // ... whether the chain gets here at all (1; 2: it ends here)
};
constexpr u32 TA_X_MERGED = 0xfffffffeu;  // map[]: this entry's chain is the merged one

struct FbMeta {       // a block's header, parsed (k_fblk_head)
  u32 mode;           // 1: tables built, the fields below are valid
  u32 bfinal, sym_start, piece_bits, pcap, hard_end, fixed, pad1;
};
struct FbShared {
  TokWarpSmem T;
  TaWarp tw[FB_WARPS];
  u32 item;
};
constexpr int FB_SMEM = (int)sizeof(FbShared);

// the stream as 32-bit words, bytes outside [0, n) reading as zero
struct TaSrc {
  const u8 *in;
  u64 n;
  const u32 *words;
  u32 skew;
  u32 fixed_run;  // the block uses the fixed code: an end-of-block code followed by the header of another non-final fixed block
                  // (0, 01) is walked over like a token — a run of fixed blocks decodes as one (ta_tile)
  __device__ __forceinline__ void init(const u8 *in_, u64 n_) {
    fixed_run = 0;
    in = in_; n = n_;
    skew = (u32)((uintptr_t)in_ & 3);
    words = reinterpret_cast<const u32 *>(in_ - skew);
  }
  __device__ __forceinline__ u32 load_word(u64 w) const {
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
};

// The token that starts at the 64 bits (hi:lo): its length in bits (| TA_EOB for the end-of-block code), 0 when no code
// starts like this (or one of the symbols the reference's tables do not define: the sequential decoder knows what
// the reference does with them); tokv / olen as in the other decoders.
__device__ __forceinline__ u32 ta_decode(const TokWarpSmem *T, u32 lo, u32 hi, u32 &tokv, u32 &olen) {
  const TokCore *S = &T->w;
  tokv = 0; olen = 0;
  u32 e = T->lut_ll[lo & ((1u << LL_ROOT) - 1)];
  if ((e & 15) == 0) {
    u32 sym, l;
    if (!inf_slow(((u64)hi << 32) | lo, &S->tab_ll, S->sorted_ll, sym, l)) return 0;
    e = tk_entry_ll(sym, l);
    if (e & TK_INV) return 0;
  }
  const u32 l1 = e & 15;
  if (e & TK_EOB) return l1 | TA_EOB;
  if (!(e & TK_LEN)) { tokv = (e >> 8) & 0xff; olen = 1; return l1; }
  const u32 eb = (e >> 4) & 15, sh2 = l1 + eb;
  const u32 len = ((e >> 8) & 0xffff) + ((lo >> l1) & ~(0xffffffffu << eb));
  const u32 y = __funnelshift_r(lo, hi, sh2);  // the stream after the length's extra bits (sh2 <= 20)
  u32 d = T->lut_d[y & ((1u << D_ROOT) - 1)];
  if ((d & 15) == 0) {
    u32 sym, l;
    if (!inf_slow((((u64)hi << 32) | lo) >> sh2, &S->tab_d, S->sorted_d, sym, l)) return 0;
    d = tk_entry_d(sym, l);
    if (d & TK_INV) return 0;
  }
  const u32 l2 = d & 15, db = (d >> 4) & 15;
  const u32 dist = (d >> 8) + ((y >> l2) & ~(0xffffffffu << db));
  tokv = 0x80000000u | ((len - 3) << 16) | (dist - 1);
  olen = len;
  return sh2 + l2 + db;
}

// ta_decode's length for the common case, without a branch: TA_MARK when a code is longer than the root tables
__device__ __forceinline__ u32 ta_fast(const TokWarpSmem *T, u32 lo, u32 hi) {
  const u32 e = T->lut_ll[lo & ((1u << LL_ROOT) - 1)];
  const u32 l1 = e & 15, eb = (e >> 4) & 15, sh2 = l1 + eb;
  const u32 y = __funnelshift_r(lo, hi, sh2);
  const u32 d = T->lut_d[y & ((1u << D_ROOT) - 1)];  // (looked up for literals too: no branch)
  const u32 l2 = d & 15, db = (d >> 4) & 15;
  const bool is_len = (e & TK_LEN) != 0;
  u32 v = is_len ? sh2 + l2 + db : l1;
  if (e & TK_EOB) v |= TA_EOB;
  if (l1 == 0 || (is_len && l2 == 0)) v = TA_MARK;
  return v;
}

// the 64 bits at position p of the staged tile
__device__ __forceinline__ void ta_bits_at(const TaWarp *W, u32 sh, u32 p, u32 &lo, u32 &hi) {
  const u32 t = sh + p, j = t >> 5;
  const u32 w0 = W->words[j], w1 = W->words[j + 1], w2 = W->words[j + 2];
  lo = __funnelshift_r(w0, w1, t);  // shift taken modulo 32
  hi = __funnelshift_r(w1, w2, t);
}

// Stages the tile that starts at block-relative bit t0 and fills nx[] for its positions; positions at or behind
// hard_end hold no code.  Returns sh, the bit offset of position 0 in words[0].
__device__ __forceinline__ u32 ta_tile(TaWarp *W, const TokWarpSmem *T, const TaSrc &src, u64 bit0, u32 t0, u32 hard_end, u32 upto = TA_TILE) {
  const u32 lane = lane_id();
  const u64 a = bit0 + t0 + ((u64)src.skew << 3);
  const u64 w0 = a >> 5;
  const u32 sh = (u32)(a & 31);
  const u32 nr = umin(TA_TILE / 32, ((upto + 127) / 128) * 4);  // rounds: only the first `upto` positions are needed
  __syncwarp();
  for (u32 i = lane; i < nr + 4; i += 32) W->words[i] = src.load_word(w0 + i);
  __syncwarp();
  for (u32 r = 0; r < nr; r += 4) {  // four independent positions per lane in flight
    u32 v[4];
#pragma unroll
    for (u32 k = 0; k < 4; k++) {
      u32 lo, hi;
      ta_bits_at(W, sh, 32 * (r + k) + lane, lo, hi);
      v[k] = ta_fast(T, lo, hi);
    }
#pragma unroll
    for (u32 k = 0; k < 4; k++) {
      const u32 p = 32 * (r + k) + lane;
      if (v[k] == TA_MARK) {  // a code the root tables do not hold
        u32 lo, hi, tokv, olen;
        ta_bits_at(W, sh, p, lo, hi);
        v[k] = ta_decode(T, lo, hi, tokv, olen);
      }
      if (src.fixed_run && (v[k] & TA_EOB) && (v[k] & 63)) {  // an end-of-block code of the fixed code: what follows it?
        u32 lo, hi;
        ta_bits_at(W, sh, p, lo, hi);
        if (((lo >> (v[k] & 63)) & 7) == 2) v[k] = (v[k] & 63) + 3;  // BFINAL = 0, BTYPE = 01: the chain goes on behind the header
      }
      if (v[k] == 0 || t0 + p >= hard_end) v[k] = TA_EOB;  // no code starts here
      W->nx[p] = (u8)v[k];
    }
  }
  // (a partial tile: no stale marks up to the end of the last chunk ta_emit_marked looks at)
  for (u32 p = nr * 32 + lane; p < umin(TA_TILE, ((upto + TA_CHUNK - 1) / TA_CHUNK) * TA_CHUNK); p += 32) W->nx[p] = (u8)TA_EOB;
  __syncwarp();
  return sh;
}

// One lane walks the chain that is at `c` through the tile [t0, t0 + TA_TILE) — one dependent shared-memory load per
// token — up to `stop`, marking the token starts in nx[] (TA_MARK).  state: 0 = still going, 1 = read the
// end-of-block code (fin = the bit after it), 2 = ran into something that is no code.  Warp-uniform on return.
__device__ __forceinline__ void ta_walk_marking(TaWarp *W, u32 t0, u32 stop, u32 &c, u32 &state, u32 &fin) {
  const u32 tend = umin(t0 + TA_TILE, stop);
  if (lane_id() == 0) {
    while (c < tend) {
      const u32 v = W->nx[c - t0];
      if (v & TA_EOB) {
        if (v & 63) { state = 1; fin = c + (v & 63); } else state = 2;
        break;
      }
      W->nx[c - t0] = (u8)(v | TA_MARK);
      c += v;
    }
  }
  c = __shfl_sync(ZLES_FULL, c, 0);
  state = __shfl_sync(ZLES_FULL, state, 0);
  fin = __shfl_sync(ZLES_FULL, fin, 0);
  __syncwarp();
}

// The marked tokens, decoded a lane per token and appended to tokP[nt ...] (room for cap): 512 positions at a time
// (16 per lane, their marks gathered from nx[]), their positions compacted into a list first.  acc collects, per
// lane, the bytes they stand for.  Returns false when the room is used up.
__device__ __forceinline__ bool ta_emit_marked(TaWarp *W, const TokWarpSmem *T, u32 sh, u32 *tokP, u32 cap, u32 &nt, u32 &acc, u32 upto = TA_TILE) {
  const u32 lane = lane_id();
  const u32 nch = umin(TA_TILE / TA_CHUNK, (upto + TA_CHUNK - 1) / TA_CHUNK);  // (marks only in the first `upto` positions)
  for (u32 ch = 0; ch < nch; ch++) {
    const uint4 q = *reinterpret_cast<const uint4 *>(W->nx + ch * TA_CHUNK + lane * 16);
    // bit 7 of the 16 bytes -> a 16-bit mask (byte k of a word to bit k: the multiply gathers the four bits)
    const u32 n0 = ((((q.x >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 15, n1 = ((((q.y >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 15;
    const u32 n2 = ((((q.z >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 15, n3 = ((((q.w >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 15;
    const u32 word = n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
    const u32 cnt = (u32)__popc(word);
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 t = __shfl_up_sync(ZLES_FULL, inc, d);
      if (lane >= (u32)d) inc += t;
    }
    const u32 total = __shfl_sync(ZLES_FULL, inc, 31);
    if (total == 0) continue;
    if (nt + total > cap) return false;
    u32 o = inc - cnt, wv = word;
    while (wv) {
      const u32 b = (u32)(__ffs((int)wv) - 1);
      wv &= wv - 1;
      W->list[o++] = (u16)(ch * TA_CHUNK + lane * 16 + b);
    }
    __syncwarp();
    for (u32 k0 = 0; k0 < total; k0 += 32) {
      const u32 k = k0 + lane;
      u32 tokv = 0, olen = 0;
      bool keep = false;
      if (k < total) {
        u32 lo, hi;
        ta_bits_at(W, sh, W->list[k], lo, hi);
        keep = (ta_decode(T, lo, hi, tokv, olen) & TA_EOB) == 0;  // (an end-of-block code inside a run of fixed blocks is no token)
      }
      const u32 bal = __ballot_sync(ZLES_FULL, keep);
      if (keep) {
        tokP[nt + (u32)__popc(bal & lanemask_lt())] = tokv;
        acc += olen;
      }
      nt += (u32)__popc(bal);
    }
    __syncwarp();
  }
  return true;
}

// Step 1 for the piece [q0, q0 + piece_bits): where each of the 64 entries q0 + i leaves it.  Chains entered at
// different positions fall into step after a few tokens on most data; once all chains that are still alive are at the
// same position, what follows does not depend on the entry any more: from there on one lane walks and the tokens are
// decoded and stored right away (step 3 then only has to redo the tiles before that point).
__device__ __forceinline__ void ta_map_piece(TaWarp *W, const TokWarpSmem *T, const TaSrc &src, u64 bit0, u32 q0, u32 piece_bits, u32 hard_end,
                                             u32 *map, TaPiece *info, u32 *tokP, u32 pcap) {
  const u32 lane = lane_id();
  u32 cA = q0 + lane, cB = q0 + 32 + lane;
  u32 mpos = 0, soff = 0, c = 0, state = 0, fin = 0, snt = 0, acc = 0;
  const u32 qend = q0 + piece_bits;
  for (u32 t0 = q0; t0 - q0 < piece_bits && t0 < hard_end; t0 += TA_TILE) {
    const u32 tend = t0 + TA_TILE;
    if (mpos != 0 && (state != 0 || c >= tend)) continue;
    const u32 sh = ta_tile(W, T, src, bit0, t0, hard_end);
    if (mpos == 0) {
      // The chains, position by position: everybody steps up to the foremost live chain; when all live chains stand on
      // that position they have become one (most data: after a few tokens).  Data whose codes all have about the same
      // length never gets there: after a while the chains just run to the end of the tile.
      bool merged = false;
      for (u32 round = 0; round < 48; round++) {
        const bool liveA = cA < TA_X_EOB, liveB = cB < TA_X_EOB;
        const u32 front = __reduce_max_sync(ZLES_FULL, umax(liveA ? cA : 0u, liveB ? cB : 0u));
        if (front == 0) break;  // no chain left
        const bool same = (!liveA || cA == front) && (!liveB || cB == front);
        if (__all_sync(ZLES_FULL, same)) { merged = front < umin(qend, hard_end); break; }
        if (front >= tend) break;
#pragma unroll
        for (int u = 0; u < 3; u++) {
          if (cA < front) {
            const u32 v = W->nx[cA - t0];
            cA = (v & TA_EOB) ? ((v & 63) ? ((cA + (v & 63)) | TA_X_EOB) : TA_X_BAD) : cA + v;
          }
          if (cB < front) {
            const u32 v = W->nx[cB - t0];
            cB = (v & TA_EOB) ? ((v & 63) ? ((cB + (v & 63)) | TA_X_EOB) : TA_X_BAD) : cB + v;
          }
        }
      }
      if (!merged) {
        bool more;
        do {  // (the vote every fourth step: a chain that has left the tile just idles)
          more = false;
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (cA < tend) {
              const u32 v = W->nx[cA - t0];
              cA = (v & TA_EOB) ? ((v & 63) ? ((cA + (v & 63)) | TA_X_EOB) : TA_X_BAD) : cA + v;
              more = true;
            }
            if (cB < tend) {
              const u32 v = W->nx[cB - t0];
              cB = (v & TA_EOB) ? ((v & 63) ? ((cB + (v & 63)) | TA_X_EOB) : TA_X_BAD) : cB + v;
              more = true;
            }
          }
        } while (__any_sync(ZLES_FULL, more));
        // at the end of a tile every live chain stands on its first token start behind it: have they become one?
        const bool liveA = cA < TA_X_EOB, liveB = cB < TA_X_EOB;
        const u32 front = __reduce_max_sync(ZLES_FULL, umax(liveA ? cA : 0u, liveB ? cB : 0u));
        const bool same = (!liveA || cA == front) && (!liveB || cB == front);
        merged = front != 0 && __all_sync(ZLES_FULL, same) && front < umin(qend, hard_end);
      }
      if (merged) {
        const bool liveA = cA < TA_X_EOB, liveB = cB < TA_X_EOB;
        mpos = liveA ? cA : cB;
        mpos = __reduce_max_sync(ZLES_FULL, (liveA || liveB) ? mpos : 0u);
        c = mpos;
        soff = umin(((mpos - q0) / 2 + 64) & ~31u, pcap);
        if (liveA) cA = TA_X_MERGED;
        if (liveB) cB = TA_X_MERGED;
      }
    }
    if (mpos != 0 && state == 0 && c < tend) {  // the merged chain: its tokens right away
      ta_walk_marking(W, t0, 0xffffffffu, c, state, fin);
      if (!ta_emit_marked(W, T, sh, tokP + soff, pcap - soff, snt, acc)) state = 2;
    }
  }
  if (mpos == 0) {
    // a chain that is still inside the piece ran into the end of the stream
    if (cA < qend) cA = TA_X_BAD;
    if (cB < qend) cB = TA_X_BAD;
  }
  map[lane] = cA;
  map[32 + lane] = cB;
  const u32 sob = __reduce_add_sync(ZLES_FULL, acc);
  if (lane == 0) {
    info->mpos = mpos;
    info->soff = soff;
    info->snt = snt;
    info->sob = sob;
    info->sfin = fin;
    info->sexit = state == 1 ? (TA_X_EOB | (fin & 0x7fffffffu)) : (state == 2 || c < qend) ? TA_X_BAD : c;
  }
}

// Step 3 for the piece [q0, q0 + piece_bits), entered at `entry`: the tokens that start before `stop` to tokP[0 .. nt)
// (room for cap).  Returns 0 = reached stop / left the piece (fin = where), 1 = read the end-of-block code (fin = the bit
// after it), 2 = failed.
__device__ __forceinline__ u32 ta_emit_piece(TaWarp *W, const TokWarpSmem *T, const TaSrc &src, u64 bit0, u32 q0, u32 piece_bits, u32 hard_end,
                                             u32 entry, u32 stop, u32 *tokP, u32 cap, u32 &nt, u32 &ob, u32 &fin) {
  u32 c = entry, state = 0, acc = 0;
  nt = 0;
  for (u32 t0 = q0; t0 - q0 < piece_bits && t0 < hard_end && state == 0 && c < stop; t0 += TA_TILE) {
    if (c >= t0 + TA_TILE) continue;  // (a token longer than what is left of a tile)
    const u32 upto = stop - t0 < TA_TILE ? stop - t0 : TA_TILE;  // nothing behind `stop` is looked at
    const u32 sh = ta_tile(W, T, src, bit0, t0, hard_end, upto);
    ta_walk_marking(W, t0, stop, c, state, fin);
    if (!ta_emit_marked(W, T, sh, tokP, cap, nt, acc, upto)) state = 2;
  }
  ob = __reduce_add_sync(ZLES_FULL, acc);
  if (state == 0) {
    if (c < umin(stop, q0 + piece_bits)) state = 2;  // the stream ended inside the piece
    fin = c;
  }
  return state;
}

// The kernels.  Job j of a launch is jobs[job0 + j]; its header and tables live at index jobs[].aux between the kernels
// (blocks decoded on demand share one such slot).
// k_fblk_head: a warp per block — the header, the tables (stored for the other kernels), how the coded bits are cut.
__global__ void __launch_bounds__(INF_THREADS)
k_fblk_head(const u8 *__restrict__ in, u64 n, const FbJob *__restrict__ jobs, u32 njobs, u32 job0, FbMeta *meta, TokWarpSmem *tabs) {
  ZLES_SMEM_DECL(smem_raw);
  TokWarpSmem *T = reinterpret_cast<TokWarpSmem *>(smem_raw) + warp_id();
  TokCore *S = &T->w;
  const u32 lane = lane_id();
  const u64 nbits = n << 3;
  const u32 ji = blockIdx.x * INF_WARPS + warp_id();
  if (ji >= njobs) return;
  const FbJob J = jobs[job0 + ji];
  const u64 bit0 = J.bit & ~7ull;
  const u64 span = J.hint_end > J.bit ? J.hint_end - J.bit : 0;
  FbMeta m;
  if (J.flags & FB_JOB_CONT) {  // the tables are there already: only the cut changes
    m = meta[J.aux];
    if (m.mode) {
      m.mode = 0;
      m.hard_end = (u32)umin64(nbits - bit0, 0x7fff0000ull);
      m.sym_start = (u32)(J.bit - bit0);
      m.piece_bits = fb_piece_bits(span);
      m.pcap = m.piece_bits / 2;
      if (m.sym_start < m.hard_end && J.np >= 1 && (u64)J.np * m.pcap <= J.tok_cap) m.mode = 1;
    }
    __syncwarp();
    if (lane == 0) meta[J.aux] = m;
    return;
  }
  m.mode = 0; m.bfinal = 0; m.sym_start = 0; m.piece_bits = TA_TILE; m.pcap = 0; m.hard_end = 0; m.fixed = 0; m.pad1 = 0;
  if (J.bit + 3 <= nbits) {
    TokReader r;
    r.init(in, n, J.bit >> 3);
    r.skip((u32)(J.bit & 7));
    r.refill();
    m.bfinal = r.take(1);
    const u32 btype = r.take(2);
    u32 status = 0;
    bool ok = false;
    if (btype == 2) {
      ok = tk_read_dynamic_header(r, T, status);
    } else if (btype == 1) {  // the fixed code, /root/reference/src/huffman.ts:41-53
      for (u32 i = lane; i < 352; i += 32) T->lens[i] = (u8)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : i < 320 ? 5 : 0);
      __syncwarp();
      ok = true;
      m.fixed = m.bfinal ? 0 : 1;  // (a final block is followed by nothing)
    }
    if (ok && !r.past_end()) {
      tk_build_tables(T);
      const u64 bp = r.bitpos();
      m.hard_end = (u32)umin64(nbits - bit0, 0x7fff0000ull);
      if (bp - bit0 < m.hard_end) {
        m.sym_start = (u32)(bp - bit0);
        m.piece_bits = fb_piece_bits(span);
        m.pcap = m.piece_bits / 2;
        if (J.np >= 1 && (u64)J.np * m.pcap <= J.tok_cap) m.mode = 1;
      }
    }
  }
  __syncwarp();
  if (m.mode) {  // the tables, for k_fblk_map and k_fblk_prefix
    const u32 *s4 = reinterpret_cast<const u32 *>(T);
    u32 *d4 = reinterpret_cast<u32 *>(tabs + J.aux);
    for (u32 i = lane; i < sizeof(TokWarpSmem) / 4; i += 32) d4[i] = s4[i];
  }
  if (lane == 0) meta[J.aux] = m;
}

// shared by k_fblk_map and k_fblk_prefix: loads the tables of the item's job.  False = nothing to do for this item.
__device__ __forceinline__ bool fb_item_begin(FbShared *Sh, const FbItem it, const FbMeta *__restrict__ meta, const TokWarpSmem *__restrict__ tabs,
                                              u32 aux, FbMeta &m) {
  m = meta[aux];
  if (!m.mode) return false;
  // pieces that start behind the end of the stream do not exist
  if ((u64)m.sym_start + (u64)it.p0 * m.piece_bits >= m.hard_end) return false;
  const u32 *s4 = reinterpret_cast<const u32 *>(tabs + aux);
  u32 *d4 = reinterpret_cast<u32 *>(&Sh->T);
  for (u32 i = threadIdx.x; i < sizeof(TokWarpSmem) / 4; i += FB_THREADS) d4[i] = s4[i];
  __syncthreads();
  return true;
}

// k_fblk_map: step 1.  A work item is sixteen pieces of one block, a warp each.
__global__ void __launch_bounds__(FB_THREADS)
k_fblk_map(const u8 *__restrict__ in, u64 n, const FbJob *__restrict__ jobs, const FbItem *__restrict__ items, u32 nitems, u32 job0,
           const FbMeta *__restrict__ meta, const TokWarpSmem *__restrict__ tabs, u32 *maps, TaPiece *infos, u32 *counter) {
  ZLES_SMEM_DECL(smem_raw);
  FbShared *Sh = reinterpret_cast<FbShared *>(smem_raw);
  TaWarp *W = &Sh->tw[warp_id()];
  const u32 w = warp_id();
  TaSrc src;
  src.init(in, n);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) Sh->item = atomicAdd(counter, 1u);
    __syncthreads();
    if (Sh->item >= nitems) break;
    const FbItem it = items[Sh->item];
    const FbJob J = jobs[job0 + it.ji];
    FbMeta m;
    if (!fb_item_begin(Sh, it, meta, tabs, J.aux, m)) continue;
    src.fixed_run = m.fixed;
    const u32 p = it.p0 + w;
    if (p >= J.np) continue;
    const u64 q0 = (u64)m.sym_start + (u64)p * m.piece_bits;
    u32 *map = maps + (size_t)(J.piece0 + p) * TA_ENT;
    TaPiece *info = infos + (size_t)J.piece0 + p;
    if (q0 >= m.hard_end) continue;  // behind the end of the stream (k_fblk_chain does not look at it)
    ta_map_piece(W, &Sh->T, src, J.bit & ~7ull, (u32)q0, m.piece_bits, m.hard_end, map, info, J.tok + (size_t)p * m.pcap, m.pcap);
  }
}

// k_fblk_chain: step 2, a warp per block.  The block's first token starts after its header; piece p + 1 is entered where
// piece p was left.  Where a piece is left does not depend on where it was entered once its chains have become one
// (TA_X_MERGED -> sexit), so 32 pieces are chained per round: lane i assumes piece i - 1 was entered through its merged
// chain — true as long as every lane before it found just that — and the round ends at the first piece for which it is
// not, whose exit then is exact.  Leaves the result (status, end, what the merged chains stand for) and the token runs of
// the merged chains; k_fblk_prefix adds the runs before them.
__global__ void __launch_bounds__(128)
k_fblk_chain(const FbJob *__restrict__ jobs, u32 njobs, u32 job0, const FbMeta *__restrict__ meta, const u32 *__restrict__ maps, TaPiece *infos,
             FbRes *res, FbPiece *pieces, u64 n) {
  const u32 ji = blockIdx.x * 4 + warp_id(), lane = lane_id();
  if (ji >= njobs) return;
  const u32 j = job0 + ji;
  const FbJob J = jobs[j];
  const FbMeta m = meta[J.aux];
  const u64 bit0 = J.bit & ~7ull;
  FbPiece *pc = pieces + 2 * (size_t)J.piece0;
  TaPiece *inf = infos + J.piece0;
  u32 status = 0, last = 0, fin = 0, total_tok = 0, total_out = 0;
  if (m.mode) {
    u32 e = m.sym_start;  // where the next piece is entered
    status = FB_LONG;
    for (u32 p0 = 0; p0 < J.np && status == FB_LONG;) {
      const u32 p = p0 + lane;
      const bool have = p < J.np;
      const u64 q0 = (u64)m.sym_start + (u64)p * m.piece_bits;
      // the exit of the piece before (lane 0: exact; the others: if it was entered through its merged chain)
      u32 prev = e;
      if (lane > 0 && have) { prev = inf[p - 1].mpos ? inf[p - 1].sexit : TA_X_BAD; }
      u32 x = TA_X_BAD;
      const bool enter = have && q0 < m.hard_end && prev < TA_X_EOB && prev >= q0 && prev - q0 < TA_ENT;
      if (enter) x = maps[(size_t)(J.piece0 + p) * TA_ENT + (prev - (u32)q0)];
      const bool mg = x == TA_X_MERGED;
      const u32 exit = mg ? inf[p].sexit : x;
      const bool stop = have && (!mg || exit >= TA_X_EOB);  // the chain of assumptions ends here: this piece's exit is exact
      const u32 nhave = (u32)__popc(__ballot_sync(ZLES_FULL, have));  // lanes 0 .. nhave - 1 have a piece (nhave >= 1)
      const u32 sb = __ballot_sync(ZLES_FULL, stop);
      const u32 first = sb ? (u32)__ffs((int)sb) - 1 : 32u;
      const u32 upto = umin(first, nhave - 1);
      if (lane <= upto) {
        inf[p].entry = prev;
        inf[p].merged = mg;
        inf[p].on_chain = 1;
        if (mg) {
          total_tok += inf[p].snt;
          total_out += inf[p].sob;
        }
      }
      const u32 xe = __shfl_sync(ZLES_FULL, exit, (int)upto);
      last = p0 + upto;
      if (xe == TA_X_BAD) status = 0;
      else if (xe & TA_X_EOB) { status = FB_OK; fin = xe & 0x7fffffffu; }
      else { e = xe; fin = xe; }
      p0 += upto + 1;
    }
    if (status == FB_OK && lane == 0) inf[last].on_chain = 2;
    if (bit0 + fin > (n << 3)) status = 0;  // the end-of-block code took bits the stream does not have
  }
  total_tok = __reduce_add_sync(ZLES_FULL, total_tok);
  total_out = __reduce_add_sync(ZLES_FULL, total_out);
  __syncwarp();
  if (status) {  // the token runs of the merged chains (those before them: k_fblk_prefix)
    for (u32 p = lane; p <= last; p += 32) {
      FbPiece e0, e1;
      e0.tok_off = 0; e0.cnt = 0; e0.bytes = 0; e0.pad = 0;
      e1 = e0;
      if (inf[p].merged) {
        e1.tok_off = p * m.pcap + inf[p].soff;
        e1.cnt = inf[p].snt;
        e1.bytes = inf[p].sob;
      }
      pc[2 * p] = e0;
      pc[2 * p + 1] = e1;
    }
    for (u32 p = last + 1 + lane; p < J.np; p += 32) inf[p].on_chain = 0;
  } else {
    for (u32 p = lane; p < J.np; p += 32) inf[p].on_chain = 0;  // nothing for k_fblk_prefix to do
  }
  if (lane == 0) {
    FbRes r0;
    r0.end_bit = bit0 + fin;
    r0.out_len = status ? total_out : 0;
    r0.ntok = status ? total_tok : 0;
    r0.status = status;
    r0.bfinal = m.bfinal;
    r0.npieces = status ? 2 * (last + 1) : 0;
    r0.pad = 0;
    res[j] = r0;
  }
}

// k_fblk_prefix: step 3.  The pieces of the chain from their true entries, up to where their chains had become one (a
// piece whose chains never did: all of it).  Adds its token runs to the block's result; a piece that does not come out
// as step 1 said fails the block.
__global__ void __launch_bounds__(FB_THREADS)
k_fblk_prefix(const u8 *__restrict__ in, u64 n, const FbJob *__restrict__ jobs, const FbItem *__restrict__ items, u32 nitems, u32 job0,
              const FbMeta *__restrict__ meta, const TokWarpSmem *__restrict__ tabs, const TaPiece *__restrict__ infos, FbRes *res, FbPiece *pieces,
              u32 *counter) {
  ZLES_SMEM_DECL(smem_raw);
  FbShared *Sh = reinterpret_cast<FbShared *>(smem_raw);
  TaWarp *W = &Sh->tw[warp_id()];
  const u32 w = warp_id(), lane = lane_id();
  TaSrc src;
  src.init(in, n);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) Sh->item = atomicAdd(counter, 1u);
    __syncthreads();
    if (Sh->item >= nitems) break;
    const FbItem it = items[Sh->item];
    const u32 j = job0 + it.ji;
    const FbJob J = jobs[j];
    // nothing to do for a block that did not chain up, or for pieces behind its last (the chain fills on_chain from piece 0 on)
    if (res[j].status == 0 || infos[(size_t)J.piece0 + it.p0].on_chain == 0) continue;
    FbMeta m;
    if (!fb_item_begin(Sh, it, meta, tabs, J.aux, m)) continue;
    src.fixed_run = m.fixed;
    const u32 p = it.p0 + w;
    if (p >= J.np) continue;
    const TaPiece I = infos[(size_t)J.piece0 + p];
    if (I.on_chain == 0) continue;
    const bool mg = I.merged != 0, is_last = I.on_chain == 2;
    u32 nt = 0, ob = 0, fin = 0;
    const u32 st = ta_emit_piece(W, &Sh->T, src, J.bit & ~7ull, m.sym_start + p * m.piece_bits, m.piece_bits, m.hard_end, I.entry,
                                 mg ? I.mpos : 0xffffffffu, J.tok + (size_t)p * m.pcap, mg ? I.soff : m.pcap, nt, ob, fin);
    if (lane == 0) {
      bool ok;
      if (mg) ok = st == 0 && fin == I.mpos;  // the run must end exactly where the merged chain starts
      else ok = st != 2 && (st == 1) == is_last && (!is_last || (J.bit & ~7ull) + fin == res[j].end_bit);  // (on_chain 2: the end-of-block code is in it)
      if (!ok) {
        atomicExch(&res[j].status, 0u);
      } else {
        FbPiece e;
        e.tok_off = p * m.pcap; e.cnt = nt; e.bytes = ob; e.pad = 0;
        pieces[2 * ((size_t)J.piece0 + p)] = e;
        atomicAdd(&res[j].ntok, nt);
        atomicAdd(&res[j].out_len, ob);
      }
    }
  }
}

// k_fblk_compact: a block decoded on demand leaves its `nslots` token runs contiguous at the start of its room (one CTA)
__global__ void __launch_bounds__(FB_THREADS) k_fblk_compact(const FbJob *__restrict__ jobs, u32 j, const FbRes *__restrict__ res, FbPiece *pieces) {
  const FbJob J = jobs[j];
  u32 *tok = J.tok;
  FbPiece *pc = pieces + 2 * (size_t)J.piece0;
  const u32 tid = threadIdx.x;
  const u32 nslots = res[j].status ? res[j].npieces : 0;
  u32 total = 0;
  for (u32 q = 0; q < nslots; q++) {  // runs move down, never up, in order
    const u32 cnt = pc[q].cnt, src0 = pc[q].tok_off, dst0 = total;
    total += cnt;
    __syncthreads();  // everybody has read pc[q]
    if (tid == 0) pc[q].tok_off = dst0;
    if (cnt == 0 || src0 == dst0) continue;
    for (u32 i0 = 0; i0 < cnt; i0 += FB_THREADS) {
      const u32 i = i0 + tid;
      u32 v = 0;
      if (i < cnt) v = tok[src0 + i];
      __syncthreads();  // a store below may land on what another thread of this round has just read
      if (i < cnt) tok[dst0 + i] = v;
    }
  }
}

// ---- phase B on pieces ------------------------------------------------------------------------------------
struct FbEnt {      // one block of the accepted chain
  u64 out_off;      // where its bytes go
  u64 src;          // stored: byte offset of the payload in the input
  u32 job;          // coded: index of its job / result
  u32 len;          // bytes it stands for
  u32 stored;
  u32 warp0;        // k_fpiece_sym: its first warp (a warp per token run; FB_STORED_WARPS for a stored block)
  u32 slot0;        // its first token run in the stream-wide numbering of token runs (a stored block counts as one)
  u32 nslots;
};
constexpr u32 FB_STORED_WARPS = 16;

// the last entry e with key(e) <= g (ents[0] has key 0)
template <typename F>
__device__ __forceinline__ u32 fb_find_ent(u32 nent, u32 g, F key) {
  u32 lo = 0, hi = nent;
  while (hi - lo > 1) {
    const u32 mid = (lo + hi) >> 1;
    if (key(mid) <= g) lo = mid; else hi = mid;
  }
  return lo;
}

// where every token run of a coded entry starts inside the entry (FbPiece.pad): a warp per entry
__global__ void __launch_bounds__(128)
k_fslot_scan(const FbJob *__restrict__ jobs, FbPiece *pieces, const FbEnt *__restrict__ ents, u32 nent) {
  const u32 e = blockIdx.x * 4 + warp_id(), lane = lane_id();
  if (e >= nent) return;
  const FbEnt en = ents[e];
  if (en.stored) return;
  FbPiece *pc = pieces + 2 * (size_t)jobs[en.job].piece0;
  u32 base = 0;
  for (u32 q0 = 0; q0 < en.nslots; q0 += 32) {
    const u32 q = q0 + lane;
    const u32 b = q < en.nslots ? pc[q].bytes : 0;
    u32 inc = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 t = __shfl_up_sync(ZLES_FULL, inc, d);
      if (lane >= (u32)d) inc += t;
    }
    if (q < en.nslots) pc[q].pad = base + inc - b;
    base += __shfl_sync(ZLES_FULL, inc, 31);
  }
}

// run r starts with the stream's token run run_first[r]: where its bytes start (run_off[nruns] = total)
__global__ void __launch_bounds__(128)
k_frun_offsets(const FbJob *__restrict__ jobs, const FbPiece *__restrict__ pieces, const FbEnt *__restrict__ ents, u32 nent,
               const u32 *__restrict__ run_first, u32 nruns, u64 total, u64 *run_off) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nruns) return;
  if (r == nruns) { run_off[r] = total; return; }
  const u32 f = run_first[r];
  const u32 e = fb_find_ent(nent, f, [&](u32 i) { return ents[i].slot0; });
  const FbEnt en = ents[e];
  run_off[r] = en.out_off + (en.stored ? 0u : pieces[2 * (size_t)jobs[en.job].piece0 + (f - en.slot0)].pad);
}

// one warp per (chain entry, token run): the run's tokens into 16-bit symbols at sym + where its bytes go
__global__ void __launch_bounds__(RES_THREADS)
k_fpiece_sym(const FbJob *__restrict__ jobs, const FbPiece *__restrict__ pieces, const FbEnt *__restrict__ ents, u32 nent, u32 nwarps,
             const u8 *__restrict__ in, u16 *sym) {
  ZLES_SMEM_DECL(smem_raw);
  const u32 g = blockIdx.x * RES_WARPS + warp_id();
  if (g >= nwarps) return;
  const FbEnt en = ents[fb_find_ent(nent, g, [&](u32 i) { return ents[i].warp0; })];
  const u32 p = g - en.warp0;
  if (en.stored) {  // the entry's warps share the copy
    const u32 a = (u32)((u64)en.len * p / FB_STORED_WARPS), b = (u32)((u64)en.len * (p + 1) / FB_STORED_WARPS);
    const u8 *src = in + en.src;
    u16 *dst = sym + en.out_off;
    for (u32 i = a + lane_id(); i < b; i += 32) dst[i] = src[i];
    return;
  }
  const FbJob J = jobs[en.job];
  const FbPiece me = pieces[2 * (size_t)J.piece0 + p];
  if (me.cnt == 0) return;
  SymState st;
  st.ring = reinterpret_cast<u16 *>(smem_raw) + warp_id() * SEG_RING;
  st.base = sym + en.out_off + me.pad;
  st.o = 0;
  st.refs = 0;
  st.vfrom = 0;
  sym_tokens<SEG_RING>(st, J.tok + me.tok_off, me.cnt);
}

// run r = the stream's token runs [run_first[r], run_first[r + 1]), made concrete in order.  A reference that lands
// inside the run takes what is there (a byte, or a reference into the run's window already); one that reaches before the
// run becomes a reference into the 32 KiB before the run.  any_refs is set when a run other than the first keeps one.
constexpr int MRG_THREADS = 512;
__global__ void __launch_bounds__(MRG_THREADS)
k_frun_merge(const FbJob *__restrict__ jobs, const FbPiece *__restrict__ pieces, const FbEnt *__restrict__ ents, u32 nent,
             const u32 *__restrict__ run_first, const u64 *__restrict__ run_off, u32 nruns, u16 *sym, u32 *any_refs) {
  for (u32 r = blockIdx.x; r < nruns; r += gridDim.x) {
    const u32 f0 = run_first[r], f1 = run_first[r + 1];
    const u64 run_base = run_off[r];
    u32 kept = 0;
    u32 e = fb_find_ent(nent, f0, [&](u32 i) { return ents[i].slot0; });
    for (u32 f = f0; f < f1; f++) {
      while (e + 1 < nent && ents[e + 1].slot0 <= f) e++;
      const FbEnt en = ents[e];
      if (en.stored) continue;  // concrete already
      const FbPiece me = pieces[2 * (size_t)jobs[en.job].piece0 + (f - en.slot0)];
      const u32 bytes = me.bytes;
      if (bytes == 0) continue;
      const u64 pstart = en.out_off + me.pad;
      u16 *s = sym + pstart;
      const long long rel = (long long)(pstart - run_base) - (long long)SYM_WIN;
      for (u32 i0 = threadIdx.x; i0 < bytes; i0 += 4 * MRG_THREADS) {  // four independent elements per thread and round
        u32 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const u32 i = i0 + u * MRG_THREADS; v[u] = i < bytes ? s[i] : 0u; }
        bool ch[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          ch[u] = v[u] >= SYM_REF;
          if (ch[u]) {
            const long long q = rel + (long long)(v[u] & 0x7fff);
            if (q >= 0) v[u] = sym[run_base + (u64)q];
            else v[u] = SYM_REF | (u32)(q + (long long)SYM_WIN);
            if (v[u] >= SYM_REF) kept = 1;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { const u32 i = i0 + u * MRG_THREADS; if (i < bytes && ch[u]) s[i] = (u16)v[u]; }
      }
      __syncthreads();  // the next token run reads what this one wrote
    }
    if (kept && r > 0) atomicOr(any_refs, 1u);
  }
}

}  // namespace zles
