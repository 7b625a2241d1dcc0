// inflate_foreign.cuh — parallel inflate of zlib streams made by OTHER encoders, first of all by
// zlib.es itself: its deflate bit-concatenates independent 128 KiB dynamic blocks
// (/root/reference/src/deflate.ts:20-34, src/lz77.ts:11-22,37) without any marker, so block
// starts have to be found.
//
//   k_hdr_filter / k_hdr_verify
//                   every bit position of the stream is tested for being the start of a dynamic
//                   block (BTYPE=2): field ranges, a complete code-length code, code lengths that
//                   decode without overrun into complete literal/length and distance codes with an
//                   end-of-block code.  What passes is a candidate (false positives are harmless).
//   k_fblk_tokens   (inflate_fblk.cuh) one CTA per candidate decodes that ONE block, dynamic or fixed, into tokens — its
//                   coded bits cut into up to 16 pieces decoded side by side — and records where it ends and how
//                   many bytes it stands for.
//   (host)          walks the chain from the first block: the block after one that ends at bit e is the candidate at
//                   bit e, until a BFINAL block; a block the scan cannot recognise (a fixed block: its header is three
//                   bits) is decoded on demand when the chain reaches it.  An error hands the stream to the sequential
//                   decoder, which is exact about the reference's behaviour on every input.
//   k_fpiece_sym / k_frun_merge / k_win_propagate / k_sym_finalize
//                   phase B: every piece is resolved on its own warp into 16-bit symbols, a byte that comes from
//                   before the piece staying symbolic (0x8000 | offset into the 32 KiB before it, copied around like
//                   any other value); the pieces of a run (>= 128 KiB, mostly) are made concrete in order; what
//                   reaches before a run refers to the run's window; the windows are made concrete run after run
//                   (32 KiB each, cheap; skipped when no run kept a reference — zlib.es streams), and a last parallel
//                   pass substitutes them (the two-pass scheme of pugz / rapidgzip).
#pragma once
#include "inflate.cuh"

namespace zles {

constexpr u32 FB_TOK = 1u << 22;      // token capacity per block at most
constexpr u32 FB_OK = 1;       // a decoded dynamic block

struct FbRes {
  u64 end_bit;     // bit position just after the block's end-of-block code
  u32 out_len;     // bytes the block stands for
  u32 ntok;
  u32 status;      // FB_OK or 0
  u32 bfinal;
  u32 npieces;     // pieces it was decoded in (inflate_fblk.cuh)
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

// The scan, in two steps.  k_hdr_filter: a thread per byte of the stream, 8 bit offsets each, everything from registers
// (the CTA's 256 bytes + 16 are staged in shared memory once): stored-block headers (LEN = ~NLEN) go to st[], and a
// position that could start a dynamic block — BTYPE, HLIT / HDIST in range, and a code-length code that is complete
// (or a single code), its Kraft sum taken three 3-bit lengths at a time from a 512-entry table — to surv[]: about one
// bit position in a thousand.  k_hdr_verify: a thread per survivor runs the full test (fb_header_ok).  Candidates are
// appended in no particular order.  cnt[0] = dynamic candidates, cnt[1] = stored candidates, cnt[2] = survivors.
constexpr int HS_THREADS = 256;
constexpr int HS_HALO_WORDS = 4;
constexpr int HS_SMEM = 1024 + (HS_THREADS / 4 + HS_HALO_WORDS) * 4;
__global__ void __launch_bounds__(HS_THREADS) k_hdr_filter(const u8 *__restrict__ in, u64 n, u64 first_bit, u64 *surv, u32 surv_cap, FbStored *st,
                                                           u32 st_cap, u32 *cnt) {
  ZLES_SMEM_DECL(smem_raw);
  u16 *lut = reinterpret_cast<u16 *>(smem_raw);                 // [512]
  u32 *tile = reinterpret_cast<u32 *>(smem_raw + 1024);         // [HS_THREADS / 4 + HS_HALO_WORDS]
  const u32 tid = threadIdx.x;
  for (u32 i = tid; i < 512; i += HS_THREADS) {
    u32 k = 0, used = 0;
    for (u32 q = 0; q < 3; q++) {
      const u32 l = (i >> (3 * q)) & 7;
      if (l) { k += 128u >> l; used++; }
    }
    lut[i] = (u16)(k | (used << 12));
  }
  const bool aligned = ((uintptr_t)in & 3) == 0;
  const u64 ntiles = (n + HS_THREADS - 1) / HS_THREADS;
  for (u64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    __syncthreads();
    const u64 base = t * HS_THREADS;
    if (tid < HS_THREADS / 4 + HS_HALO_WORDS) {
      const u64 b = base + 4 * tid;
      u32 w = 0;
      if (aligned && b + 4 <= n) {
        w = __ldg(reinterpret_cast<const u32 *>(in + b));
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (b + q < n) w |= (u32)in[b + q] << (8 * q);
      }
      tile[tid] = w;
    }
    __syncthreads();
    const u64 B = base + tid;
    if (B >= n) continue;
    const u32 k = tid >> 2, sh = (tid & 3) * 8;
    const u32 w0 = tile[k], w1 = tile[k + 1], w2 = tile[k + 2], w3 = tile[k + 3];
    const u32 x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh), x2 = __funnelshift_r(w2, w3, sh);  // the 96 bits from byte B on
#pragma unroll
    for (u32 s = 0; s < 8; s++) {
      const u64 bit = (B << 3) + s;
      if (bit < first_bit) continue;
      const u32 h = __funnelshift_r(x0, x1, s);
      const u32 bt = (h >> 1) & 3;
      if (bt == 0) {
        // a stored block (RFC 1951 3.2.4): after the 3 header bits and the pad, LEN and its complement
        const u32 qo = (s + 10) >> 3;  // 1 or 2 bytes after B
        const u32 v = __funnelshift_r(x0, x1, 8 * qo);
        const u32 len = v & 0xffffu, nlen = v >> 16;
        if ((len ^ nlen) == 0xffffu && B + qo + 4 + len <= n) {
          const u32 o = atomicAdd(cnt + 1, 1u);
          if (o < st_cap) { st[o].bit = bit; st[o].len = len; st[o].bfinal = h & 1; }
        }
      } else if (bt == 2 && ((h >> 3) & 31) <= 29 && ((h >> 8) & 31) <= 29) {
        const u32 hclen = ((h >> 13) & 15) + 4;
        const u32 lo = __funnelshift_r(x0, x1, s + 17), hi = __funnelshift_r(x1, x2, s + 17);  // the 3-bit lengths of the code-length code
        const u64 v = (((u64)hi << 32) | lo) & ((1ull << (3 * hclen)) - 1);
        u32 acc = 0;
#pragma unroll
        for (u32 q = 0; q < 7; q++) acc += lut[(u32)(v >> (9 * q)) & 511];
        if ((acc & 0xfff) == 128 || (acc >> 12) == 1) {  // encoders emit complete codes (or a single one)
          const u32 o = atomicAdd(cnt + 2, 1u);
          if (o < surv_cap) surv[o] = bit;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_hdr_verify(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ surv, u32 surv_cap, u64 *cand, u32 cap,
                                                    u32 *cnt) {
  const u32 ns = umin(cnt[2], surv_cap);
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
    const u64 bit = surv[i];
    if (fb_header_ok(in, n, bit)) {
      const u32 o = atomicAdd(cnt, 1u);
      if (o < cap) cand[o] = bit;
    }
  }
}

// ---- chains with history across blocks: symbolic 32 KiB windows -------------------------------------
constexpr u32 SYM_WIN = 32768;          // deflate's window
constexpr u32 SYM_REF = 0x8000;         // symbol >= SYM_REF: byte (sym & 0x7fff) of the window before the run
constexpr u32 SYM_RING = 16384;         // u16 entries mirrored in shared memory per warp (>= 32 x 258)
constexpr int SYM_SMEM = (int)(RES_WARPS * SYM_RING * 2);
constexpr u32 SYM_RUN = 131072;         // bytes per run at least (the last run of a stream may be shorter)

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

// pass 2, one CTA: win[r] = the 32 KiB of output that end where run r ends, concrete; win[-1] (before the stream) is all
// zeros, which is what the reference's inflate reads there (/root/reference/src/inflate.ts:287-290).  A run shorter
// than 32 KiB takes the rest from the window before it.  Every thread keeps 32 symbols in flight.
__global__ void __launch_bounds__(1024) k_win_propagate(const u16 *__restrict__ sym, const u64 *__restrict__ run_off, u32 nruns, u8 *win,
                                                        const u32 *__restrict__ any_refs) {
  if (*any_refs == 0) return;  // no run refers to its window: nothing to make concrete (k_frun_merge)
  for (u32 r = 0; r + 1 < nruns; r++) {
    const u64 beg = run_off[r], end = run_off[r + 1];
    const u32 own = (u32)umin64(end - beg, (u64)SYM_WIN);  // the window's last `own` bytes are this run's
    const u8 *prev = r ? win + (size_t)(r - 1) * SYM_WIN : nullptr;
    u8 *cur = win + (size_t)r * SYM_WIN;
    u32 v[SYM_WIN / 1024];
#pragma unroll
    for (u32 i = 0; i < SYM_WIN / 1024; i++) {
      const u32 k = threadIdx.x + i * 1024;
      v[i] = k >= SYM_WIN - own ? (u32)sym[end - SYM_WIN + k] : SYM_REF | (k + own);  // (k + own < SYM_WIN: the window before, shifted)
    }
#pragma unroll
    for (u32 i = 0; i < SYM_WIN / 1024; i++) {
      const u32 k = threadIdx.x + i * 1024;
      cur[k] = v[i] < SYM_REF ? (u8)v[i] : (prev ? prev[v[i] & 0x7fff] : (u8)0);
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
