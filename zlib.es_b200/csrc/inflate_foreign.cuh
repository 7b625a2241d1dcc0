// inflate_foreign.cuh — parallel inflate of zlib streams made by OTHER encoders, first of all by
// zlib.es itself: its deflate bit-concatenates independent 128 KiB dynamic blocks
// (/root/reference/src/deflate.ts:20-34, src/lz77.ts:11-22,37) without any marker, so block
// starts have to be found.
//
//   k_hdr_scan      every bit position of the stream is tested for being the start of a dynamic
//                   block (BTYPE=2): field ranges, a complete code-length code, code lengths that
//                   decode without overrun into complete literal/length and distance codes with an
//                   end-of-block code.  What passes is a candidate (false positives are harmless).
//   k_blk_tokens    one warp per candidate decodes that ONE block into tokens (same decoder as
//                   phase A of inflate.cuh) and records where it ends, how many bytes it stands
//                   for and how far before its own start it reaches.
//   (host)          walks the chain from the first block: the block after one that ends at bit e is
//                   the candidate at bit e, until a BFINAL block.  Any gap (a stored or fixed block,
//                   an error) hands the stream to the sequential decoder, which is exact about the
//                   reference's behaviour on every input.
//   k_blk_resolve   phase B when no block reaches before its own start (zlib.es streams): every
//                   block is copied on its own warp.
//   k_run_resolve / k_win_propagate / k_sym_finalize
//                   phase B when blocks use the 32 KiB before them (system zlib): the chain is cut into
//                   runs of >= 256 KiB; every run is resolved in parallel into 16-bit symbols, a byte
//                   that comes from the unknown 32 KiB window before the run staying symbolic
//                   (0x8000 | window offset, copied around like any other value); then the windows are
//                   made concrete run after run (32 KiB each, cheap), and a last parallel pass
//                   substitutes them (the two-pass scheme of pugz / rapidgzip).
#pragma once
#include "inflate.cuh"

namespace zles {

constexpr u32 FB_TOK = 131072 + 32;  // token capacity per block (a zlib.es block of 128 KiB of literals fits)
constexpr u32 FB_OK = 1;       // a decoded dynamic block

struct FbRes {
  u64 end_bit;     // bit position just after the block's end-of-block code
  u32 out_len;     // bytes the block stands for
  u32 ntok;
  u32 status;      // FB_OK or 0
  u32 bfinal;
  u32 hist_need;   // how many bytes before the block's own start its matches reach
  u32 pad;
};

// k <= 24 bits at an arbitrary bit position; bytes outside [0, n) read as 0
__device__ __forceinline__ u32 fb_bits(const u8 *__restrict__ in, u64 n, u64 bit, u32 k) {
  const u64 B = bit >> 3;
  u32 w = 0;
  if (B + 4 <= n) {
    w = (u32)in[B] | ((u32)in[B + 1] << 8) | ((u32)in[B + 2] << 16) | ((u32)in[B + 3] << 24);
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (B + q < n) w |= (u32)in[B + q] << (8 * q);
  }
  return (w >> (bit & 7)) & ((1u << k) - 1);
}

// Is `bit` plausibly the first bit (BFINAL) of a dynamic block?  Thread-level.
__device__ __forceinline__ bool fb_header_ok(const u8 *__restrict__ in, u64 n, u64 bit) {
  const u32 h = fb_bits(in, n, bit, 17);
  if (((h >> 1) & 3) != 2) return false;                      // BTYPE
  const u32 hlit = (h >> 3) & 31, hdist = (h >> 8) & 31, hclen = ((h >> 13) & 15) + 4;
  if (hlit > 29 || hdist > 29) return false;                  // more than 286 / 30 codes
  // code-length code: (hclen) 3-bit lengths in the order of src/const.ts:33-35
  u32 cl[19];
#pragma unroll
  for (int i = 0; i < 19; i++) cl[i] = 0;
  u32 kraft = 0, used = 0;
  u64 p = bit + 17;
  for (u32 i = 0; i < hclen; i += 8) {
    const u32 v = fb_bits(in, n, p, 24);
    p += 24;
#pragma unroll
    for (u32 q = 0; q < 8; q++) {
      if (i + q < hclen) {
        const u32 l = (v >> (3 * q)) & 7;
        cl[c_cl_order[i + q]] = l;
        if (l) { kraft += 128u >> l; used++; }
      }
    }
  }
  p = bit + 17 + 3 * hclen;
  if (!(kraft == 128 || used == 1)) return false;             // encoders emit complete codes (or a single one)
  // canonical code-length code (codes by length, then symbol: src/huffman.ts:8-39)
  u32 cnt[8], first[8], offs[8];
  u8 sorted[19];
#pragma unroll
  for (int l = 0; l < 8; l++) cnt[l] = 0;
  for (int i = 0; i < 19; i++) cnt[cl[i]]++;
  {
    u32 code = 0, o = 0;
    first[0] = 0; offs[0] = 0;
    for (int l = 1; l < 8; l++) {
      first[l] = code;
      offs[l] = o;
      o += cnt[l];
      code = (code + cnt[l]) << 1;
    }
    u32 cur[8];
    for (int l = 0; l < 8; l++) cur[l] = offs[l];
    for (int i = 0; i < 19; i++)
      if (cl[i]) sorted[cur[cl[i]]++] = (u8)i;
  }
  // the HLIT + HDIST code lengths: only their Kraft sums are kept
  const u32 nll = hlit + 257, total = nll + hdist + 1;
  u32 kll = 0, kd = 0, nd = 0, nl = 0, eob = 0, prev = 0;
  u32 i = 0;
  while (i < total) {
    const u32 v = fb_bits(in, n, p, 14);  // <= 7 code bits + <= 7 extra bits
    u32 code = 0, sym = 99, l = 0;
    for (l = 1; l < 8; l++) {
      code = (code << 1) | ((v >> (l - 1)) & 1);
      const u32 d = code - first[l];
      if (code >= first[l] && d < cnt[l]) { sym = sorted[offs[l] + d]; break; }
    }
    if (sym == 99) return false;
    u32 rep = 1, val = sym;
    if (sym == 16) { if (i == 0) return false; rep = 3 + ((v >> l) & 3); val = prev; l += 2; }
    else if (sym == 17) { rep = 3 + ((v >> l) & 7); val = 0; l += 3; }
    else if (sym == 18) { rep = 11 + ((v >> l) & 127); val = 0; l += 7; }
    if (sym <= 15) prev = sym; else if (sym != 16) prev = 0;
    p += l;
    if (i + rep > total) return false;                        // a run past the last code
    if (val) {
      const u32 w = 32768u >> val;
      for (u32 q = 0; q < rep; q++) {
        const u32 j = i + q;
        if (j < nll) { kll += w; nl++; if (j == 256) eob = 1; } else { kd += w; nd++; }
      }
    }
    i += rep;
  }
  if (p > (n << 3)) return false;                             // header runs past the buffer
  if (!eob) return false;
  if (!(kll == 32768 || nl == 1)) return false;
  if (!(kd == 32768 || nd <= 1)) return false;
  return true;
}

struct FbStored {   // a stored-block candidate found by the scan
  u64 bit;          // position of its BFINAL bit
  u32 len;          // LEN
  u32 bfinal;
};

struct FbChainEnt { // one block of the accepted chain, as phase B needs it
  u64 a;            // dynamic: offset of its tokens in the token buffer; stored: byte offset of its payload in the input
  u32 b;            // dynamic: number of tokens; stored: LEN
  u32 stored;
};

// thread per byte of the stream, 8 bit offsets each; candidates are appended in no particular order:
// dynamic-block headers to cand[], stored-block headers to st[]; cnt[0] / cnt[1] count them
__global__ void __launch_bounds__(256) k_hdr_scan(const u8 *__restrict__ in, u64 n, u64 first_bit, u64 *cand, u32 cap, FbStored *st, u32 st_cap,
                                                  u32 *cnt) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 B = (u64)blockIdx.x * blockDim.x + threadIdx.x; B < n; B += stride) {
    const u32 w = fb_bits(in, n, B << 3, 24);
    for (u32 s = 0; s < 8; s++) {
      const u32 h = w >> s;
      const u64 bit = (B << 3) + s;
      if (bit < first_bit) continue;
      if (((h >> 1) & 3) == 0) {
        // a stored block (RFC 1951 3.2.4): after the 3 header bits and the pad, LEN and its complement
        const u64 q = (bit + 3 + 7) >> 3;
        if (q + 4 <= n) {
          const u32 len = (u32)in[q] | ((u32)in[q + 1] << 8), nlen = (u32)in[q + 2] | ((u32)in[q + 3] << 8);
          if ((len ^ nlen) == 0xffffu && q + 4 + len <= n) {
            const u32 o = atomicAdd(cnt + 1, 1u);
            if (o < st_cap) { st[o].bit = bit; st[o].len = len; st[o].bfinal = h & 1; }
          }
        }
      } else if (((h >> 1) & 3) == 2 && ((h >> 3) & 31) <= 29 && ((h >> 8) & 31) <= 29) {  // quick reject, then the full test
        if (fb_header_ok(in, n, bit)) {
          const u32 o = atomicAdd(cnt, 1u);
          if (o < cap) cand[o] = bit;
        }
      }
    }
  }
}

// One block, from its BFINAL bit to its end-of-block code, into tokens.
__device__ __forceinline__ void fb_block_tokens(TokWarpSmem *T, const u8 *in, u64 n, u64 bit, u32 *tok, u32 tok_cap, FbRes *res) {
  const u32 lane = lane_id();
  TokCore *S = &T->w;
  TokReader r;
  r.init(in, n, bit >> 3);
  r.skip((u32)(bit & 7));
  r.refill();
  u32 o = 0, nt = 0, mytok = 0, turn = lane, left = 32, ok = 0, hist = 0;
  const u32 bfinal = r.take(1);
  const u32 btype = r.take(2);
  u32 status = 0;
  if (btype == 2 && tk_read_dynamic_header(r, S, status)) {
    tk_build_tables(T);
    __syncwarp();
    for (;;) {
      r.refill();
      u32 e = T->lut_ll[r.lo & ((1u << LL_ROOT) - 1)];
      if ((e & 15) == 0) {
        u32 sym, l;
        if (!inf_slow(r.bits64(), &S->tab_ll, S->sorted_ll, sym, l)) break;
        e = tk_entry_ll(sym, l);
        if (e & TK_INV) break;
      }
      r.skip(e & 15);
      if (e & TK_EOB) { ok = 1; break; }
      if (nt >= tok_cap || o >= 0x7fff0000u) break;
      u32 t;
      if (e < TK_LEN) {
        t = e >> 8;
        o++;
      } else {
        const u32 len = ((e >> 8) & 0xffff) + r.take((e >> 4) & 15);
        r.refill();
        u32 d = T->lut_d[r.lo & ((1u << D_ROOT) - 1)];
        if ((d & 15) == 0) {
          u32 sym, l;
          if (!inf_slow(r.bits64(), &S->tab_d, S->sorted_d, sym, l)) break;
          d = tk_entry_d(sym, l);
          if (d & TK_INV) break;
        }
        r.skip(d & 15);
        const u32 dist = (d >> 8) + r.take((d >> 4) & 15);
        if (dist > o) hist = umax(hist, dist - o);
        t = 0x80000000u | ((len - 3) << 16) | (dist - 1);
        o += len;
      }
      if (turn == 0) mytok = t;
      turn = (turn - 1) & 31;
      nt++;
      if (--left == 0) { tok[nt - 32 + lane] = mytok; left = 32; }
    }
    if (r.past_end()) ok = 0;  // consumed bits the buffer does not have: the sequential decoder knows what the reference does
  }
  if (lane < (nt & 31)) tok[(nt & ~31u) + lane] = mytok;
  if (lane == 0) {
    res->end_bit = r.bitpos();
    res->out_len = o;
    res->ntok = nt;
    res->status = ok ? FB_OK : 0;
    res->bfinal = bfinal;
    res->hist_need = hist;
    res->pad = 0;
  }
}

// candidate j's tokens go to tokens + tok_off[j] (room for tok_cap[j], a multiple of 32)
__global__ void __launch_bounds__(INF_THREADS)
k_blk_tokens(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ cand, u32 ncand, u32 *tokens, const u64 *__restrict__ tok_off,
             const u32 *__restrict__ tok_cap, FbRes *res, u32 *counter) {
  ZLES_SMEM_DECL(smem_raw);
  TokWarpSmem *T = reinterpret_cast<TokWarpSmem *>(smem_raw) + warp_id();
  for (;;) {
    u32 j = 0;
    if (lane_id() == 0) j = atomicAdd(counter, 1u);
    j = __shfl_sync(ZLES_FULL, j, 0);
    if (j >= ncand) break;
    fb_block_tokens(T, in, n, cand[j], tokens + tok_off[j], tok_cap[j], res + j);
    __syncwarp();
  }
}

// run r = chain entries [run_first[r], run_first[r + 1]) written at out + run_off[r]
__global__ void __launch_bounds__(RES_THREADS)
k_blk_resolve(const u32 *__restrict__ tokens, const FbChainEnt *__restrict__ chain, const u32 *__restrict__ run_first,
              const u64 *__restrict__ run_off, u32 nruns, const u8 *__restrict__ in, u8 *out, u64 cap, u32 *problems) {
  ZLES_SMEM_DECL(smem_raw);
  const u32 r = blockIdx.x * RES_WARPS + warp_id();
  if (r >= nruns) return;
  ResState st;
  const u64 off = run_off[r];
  st.base = out + off;
  st.ring = smem_raw + warp_id() * RES_RING;
  st.room = (u32)umin64(off >= cap ? 0 : cap - off, 0xffffffffull);
  st.limit = 0xffffffffu;
  st.o = 0;
  st.bad = 0;
  for (u32 i = run_first[r]; i < run_first[r + 1]; i++) {
    const FbChainEnt e = chain[i];
    if (e.stored) {
      if (!res_bytes(st, in + e.a, e.b)) break;
    } else if (!res_tokens(st, tokens + e.a, e.b)) {
      break;
    }
  }
  if (st.bad && lane_id() == 0) atomicOr(problems, st.bad);
}

// ---- chains with history across blocks: symbolic 32 KiB windows -------------------------------------
constexpr u32 SYM_WIN = 32768;          // deflate's window
constexpr u32 SYM_REF = 0x8000;         // symbol >= SYM_REF: byte (sym & 0x7fff) of the window before the run
constexpr u32 SYM_RING = 16384;         // u16 entries mirrored in shared memory per warp (>= 32 x 258)
constexpr int SYM_SMEM = (int)(RES_WARPS * SYM_RING * 2);
constexpr u32 SYM_RUN = 262144;         // target bytes per run

struct SymState {
  u16 *base;    // the run's symbols
  u16 *ring;
  u32 o;
  u32 refs;     // set once a source before the run was seen (a window reference was written)
  u32 vfrom;    // the ring mirrors run positions >= vfrom only (a batch wider than the ring leaves it undefined)
};

// Same as res_tokens, on 16-bit symbols; a source position before the run (negative) yields a window reference.
// RING = entries of the warp's shared-memory mirror (a power of two; a source is read from it when it is still there).
template <u32 RING = SYM_RING>
__device__ __forceinline__ void sym_tokens(SymState &st, const u32 *__restrict__ tok, u32 nt) {
  constexpr u32 RM = RING - 1;
  const u32 lane = lane_id();
  u16 *base = st.base, *ring = st.ring;
  u32 o = st.o;
  u32 tnext = lane < nt ? __ldg(tok + lane) : 0;
  for (u32 b0 = 0; b0 < nt; b0 += 32) {
    const bool valid = b0 + lane < nt;
    const u32 t = tnext;
    tnext = b0 + 32 + lane < nt ? __ldg(tok + b0 + 32 + lane) : 0;
    const bool isM = valid && (t >> 31);
    const u32 len = !valid ? 0 : (isM ? ((t >> 16) & 255) + 3 : 1);
    const u32 dist = (t & 0x7fff) + 1;
    u32 inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 v = __shfl_up_sync(ZLES_FULL, inc, d);
      if (lane >= (u32)d) inc += v;
    }
    const u32 pos = o + inc - len;
    const u32 total = __shfl_sync(ZLES_FULL, inc, 31);
    if (valid && !isM) { base[pos] = (u16)t; ring[pos & RM] = (u16)t; }
    const bool indep = isM && len < RES_LONG && dist <= pos && pos - dist + len <= o;  // source inside the run, before this batch
    const bool before = isM && len < RES_LONG && pos + len <= dist;                     // source entirely before the run: no load at all
    const u32 later = __ballot_sync(ZLES_FULL, isM && !indep && !before);
    if (__any_sync(ZLES_FULL, before)) {
      st.refs = 1;
      if (before) {
        const u32 w0 = SYM_WIN + pos - dist;  // window offset of the first source byte (dist <= 32768)
#pragma unroll
        for (u32 q = 0; q < RES_LONG - 1; q++)
          if (q < len) { const u16 v = (u16)(SYM_REF | (w0 + q)); base[pos + q] = v; ring[(pos + q) & RM] = v; }
      }
    }
    if (indep) {
      const u32 src = pos - dist;
      u16 v[RES_LONG - 1];
#pragma unroll
      for (u32 q = 0; q < RES_LONG - 1; q++) if (q < len) v[q] = base[src + q];
#pragma unroll
      for (u32 q = 0; q < RES_LONG - 1; q++) if (q < len) { base[pos + q] = v[q]; ring[(pos + q) & RM] = v[q]; }
    }
    __syncwarp();
    u32 m = later;
    while (m) {
      const int j = __ffs((int)m) - 1;
      m &= m - 1;
      const u32 pj = __shfl_sync(ZLES_FULL, pos, j), lj = __shfl_sync(ZLES_FULL, len, j), dj = __shfl_sync(ZLES_FULL, dist, j);
      const int sj = (int)pj - (int)dj;  // may be negative: before the run
      if (sj < 0) st.refs = 1;
      // a batch wider than the ring aliases its own positions (their stores are not ordered): nothing of it is read
      // from the ring, now or later (vfrom)
      const bool in_ring = sj >= 0 && (u32)sj >= st.vfrom && total <= RING && o + total - (u32)sj <= RING;
      for (u32 q = lane; q < lj; q += 32) {
        const int s = sj + (int)(dj >= lj ? q : q % dj);
        u16 b;
        if (s < 0) b = (u16)(SYM_REF | (u32)((int)SYM_WIN + s));  // dist <= 32768, so SYM_WIN + s >= 0
        else b = in_ring ? ring[(u32)s & RM] : base[s];
        base[pj + q] = b;
        ring[(pj + q) & RM] = b;
      }
      __syncwarp();
    }
    o += total;
    if (total > RING) st.vfrom = o;
  }
  st.o = o;
}

template <u32 RING = SYM_RING>
__device__ __forceinline__ void sym_bytes(SymState &st, const u8 *__restrict__ srcp, u32 len) {
  constexpr u32 RM = RING - 1;
  for (u32 q = lane_id(); q < len; q += 32) {
    const u16 v = srcp[q];
    st.base[st.o + q] = v;
    st.ring[(st.o + q) & RM] = v;
  }
  __syncwarp();
  st.o += len;
}

// pass 1: run r = chain entries [run_first[r], run_first[r + 1]), symbols at sym + run_off[r]
__global__ void __launch_bounds__(RES_THREADS)
k_run_resolve(const u32 *__restrict__ tokens, const FbChainEnt *__restrict__ chain, const u32 *__restrict__ run_first,
              const u64 *__restrict__ run_off, u32 nruns, const u8 *__restrict__ in, u16 *sym) {
  ZLES_SMEM_DECL(smem_raw);
  const u32 r = blockIdx.x * RES_WARPS + warp_id();
  if (r >= nruns) return;
  SymState st;
  st.base = sym + run_off[r];
  st.ring = reinterpret_cast<u16 *>(smem_raw) + warp_id() * SYM_RING;
  st.o = 0;
  st.refs = 0;
  st.vfrom = 0;
  for (u32 i = run_first[r]; i < run_first[r + 1]; i++) {
    const FbChainEnt e = chain[i];
    if (e.stored) sym_bytes(st, in + e.a, e.b);
    else sym_tokens(st, tokens + e.a, e.b);
  }
}

// pass 2, one CTA: win[r] = the last 32 KiB of run r, concrete; win[-1] (before the stream) is all zeros, which is
// what the reference's inflate reads there (/root/reference/src/inflate.ts:287-290).  Runs are >= 32 KiB except the last.
__global__ void __launch_bounds__(1024) k_win_propagate(const u16 *__restrict__ sym, const u64 *__restrict__ run_off, u32 nruns, u8 *win) {
  for (u32 r = 0; r + 1 < nruns; r++) {
    const u64 end = run_off[r + 1];
    const u8 *prev = r ? win + (size_t)(r - 1) * SYM_WIN : nullptr;
    for (u32 k = threadIdx.x; k < SYM_WIN; k += 1024) {
      const u16 v = sym[end - SYM_WIN + k];
      u8 b;
      if (v < SYM_REF) b = (u8)v;
      else b = prev ? prev[v & 0x7fff] : (u8)0;
      win[(size_t)r * SYM_WIN + k] = b;
    }
    __syncthreads();
  }
}

// pass 3: symbols -> bytes
__global__ void __launch_bounds__(256) k_sym_finalize(const u16 *__restrict__ sym, const u64 *__restrict__ run_off, u32 nruns,
                                                      const u8 *__restrict__ win, u8 *out) {
  for (u32 r = blockIdx.y; r < nruns; r += gridDim.y) {
    const u64 a = run_off[r], b = run_off[r + 1];
    const u8 *prev = r ? win + (size_t)(r - 1) * SYM_WIN : nullptr;
    for (u64 i = a + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (u64)gridDim.x * blockDim.x) {
      const u16 v = sym[i];
      out[i] = v < SYM_REF ? (u8)v : (prev ? prev[v & 0x7fff] : (u8)0);
    }
  }
}

// ---- OUR streams when there are too few 128 KiB chunks to fill the GPU with one warp each (k_inf_resolve): the same
// two-pass idea at a finer grain.  Phase A (k_inf_tokens4) leaves every 32 KiB block as up to four pieces — token
// count and bytes each, pinfo — and k_piece_sym resolves every piece on its own warp into 16-bit symbols, a byte
// that comes from before the piece (<= 32 KiB back) staying symbolic; k_chunk_final then makes the pieces of a chunk
// concrete in order, each a fully parallel pass.  Sixteen times the warps of k_inf_resolve.
constexpr u32 SEG_RING = 2048;  // a small mirror keeps many warps per SM; text batches produce ~200 bytes
constexpr int SEG_SMEM = (int)(RES_WARPS * SEG_RING * 2);
constexpr u32 SEG_PIECES = 4;

// piece p of the segment with result r: its tokens [tok_off, +cnt) and bytes [out_off, +bytes); false = inconsistent
__device__ __forceinline__ bool seg_piece(const u32 *__restrict__ pinfo, u32 sidx, const InfRes &r, u32 nt, u32 p, u32 &tok_off, u32 &cnt,
                                          u32 &out_off, u32 &bytes) {
  tok_off = 0; out_off = 0; cnt = 0; bytes = 0;
  if (!pinfo) {  // phase A ran one warp per block: the block is one piece
    if (p == 0) { cnt = umin(nt, SUB); bytes = (u32)r.out_len; }
    return true;
  }
  const u32 *pi = pinfo + (size_t)sidx * 2 * SEG_PIECES;
  for (u32 k = 0; k < p; k++) { tok_off += pi[2 * k]; out_off += pi[2 * k + 1]; }
  cnt = pi[2 * p];
  bytes = pi[2 * p + 1];
  return tok_off <= SUB && cnt <= SUB - tok_off && out_off <= SUB && bytes <= SUB - out_off && out_off + bytes <= r.out_len;
}

__global__ void __launch_bounds__(RES_THREADS)
k_piece_sym(const u32 *__restrict__ tokens, const u32 *__restrict__ ntok, const u32 *__restrict__ pinfo, const u32 *__restrict__ seg_list, u32 nseg,
            u32 seg0, const u8 *__restrict__ in, const InfRes *__restrict__ res, u16 *sym, u8 *out, u64 cap, u32 *problems) {
  // this launch covers entries seg0 .. of the list (a multiple of four: whole chunks); `sym` holds the launch's symbols only
  ZLES_SMEM_DECL(smem_raw);
  const u32 g = blockIdx.x * RES_WARPS + warp_id();
  const u32 el = g / SEG_PIECES, p = g % SEG_PIECES;
  const u32 e = seg0 + el;
  if (e >= nseg) return;
  const u32 sidx = seg_list ? seg_list[e] : e;
  const InfRes r = res[sidx];
  // this kernel also runs optimistically on candidates that k_inf_check will turn down
  if ((r.status != SEG_SYNC && r.status != SEG_FINAL) || r.out_len > SUB) {
    if (p == 0 && lane_id() == 0) atomicOr(problems, 1u);
    return;
  }
  SymState st;
  st.ring = reinterpret_cast<u16 *>(smem_raw) + warp_id() * SEG_RING;
  st.o = 0;
  st.refs = 0;
  st.vfrom = 0;
  if (r.flags & SEGF_STORED) {  // a stored block is concrete already: its payload goes straight to the output, four warps a block
    const u64 so = seg_stored_src(r.end_pos, r.out_len, r.flags);
    if (so == ~0ull) { if (p == 0 && lane_id() == 0) atomicOr(problems, 1u); return; }
    const u8 *src = in + so;
    const u64 off = (u64)e * SUB;
    u32 len = (u32)r.out_len;
    if (off + len > cap) {
      if (p == 0 && lane_id() == 0) atomicOr(problems, 2u);
      len = off >= cap ? 0 : (u32)(cap - off);
    }
    for (u32 i = p * 32 + lane_id(); i < len; i += SEG_PIECES * 32) out[off + i] = src[i];
    return;
  }
  u32 tok_off, cnt, out_off, bytes;
  if (!seg_piece(pinfo, sidx, r, ntok[sidx], p, tok_off, cnt, out_off, bytes)) {
    if (lane_id() == 0) atomicOr(problems, 1u);
    return;
  }
  if (cnt == 0) return;
  st.base = sym + (size_t)el * SUB + out_off;
  sym_tokens<SEG_RING>(st, tokens + (size_t)sidx * SUB + tok_off, cnt);  // phase A made sure they stand for `bytes` bytes
}

constexpr int FIN_THREADS = 512;
__global__ void __launch_bounds__(FIN_THREADS)
k_chunk_final(const u16 *__restrict__ sym, const u32 *__restrict__ ntok, const u32 *__restrict__ pinfo, const u32 *__restrict__ seg_list, u32 nseg,
              u32 seg0, const InfRes *__restrict__ res, u8 *out, u64 cap, u32 *problems) {
  const u32 c = seg0 / SUBS_PER_CHUNK + blockIdx.x;
  u8 *cbase = out + (u64)c * CHUNK;
  for (u32 k = 0; k < SUBS_PER_CHUNK; k++) {
    const u32 e = c * SUBS_PER_CHUNK + k;
    if (e >= nseg) break;
    const u32 sidx = seg_list ? seg_list[e] : e;
    const InfRes r = res[sidx];
    if ((r.status != SEG_SYNC && r.status != SEG_FINAL) || r.out_len > SUB) break;  // flagged by k_piece_sym
    if (r.flags & SEGF_STORED) continue;  // written by k_piece_sym (an earlier launch)
    for (u32 p = 0; p < SEG_PIECES; p++) {
      u32 tok_off, cnt, out_off, bytes;
      if (!seg_piece(pinfo, sidx, r, ntok[sidx], p, tok_off, cnt, out_off, bytes)) break;  // flagged by k_piece_sym
      if (bytes == 0) continue;
      const u32 pstart = k * SUB + out_off;            // where the piece starts in the chunk
      const u64 off = (u64)c * CHUNK + pstart;
      u32 len = bytes;
      if (off + len > cap) {  // output buffer too small
        if (threadIdx.x == 0) atomicOr(problems, 2u);
        len = off >= cap ? 0 : (u32)(cap - off);
      }
      const u16 *s = sym + (size_t)(e - seg0) * SUB + out_off;
      u32 bad = 0;
      // four independent elements per thread and round: the loads of a round are all in flight together
      for (u32 i0 = threadIdx.x; i0 < len; i0 += 4 * FIN_THREADS) {
        u32 v[4];
        u8 b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const u32 i = i0 + u * FIN_THREADS; v[u] = i < len ? s[i] : 0u; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          b[u] = (u8)v[u];
          if (v[u] >= SYM_REF) {  // byte (v & 0x7fff) of the 32 KiB before the piece
            const int q = (int)pstart + (int)(v[u] & 0x7fff) - (int)SYM_WIN;
            if (q < 0) { bad = 1; b[u] = 0; }  // before the chunk: not something our encoder writes
            else b[u] = cbase[q];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { const u32 i = i0 + u * FIN_THREADS; if (i < len) cbase[pstart + i] = b[u]; }
      }
      if (bad) atomicOr(problems, 1u);
      __syncthreads();  // the next piece reads what this one wrote
    }
  }
}

}  // namespace zles
