// huffman.cuh — kernel K3: per-block length-limited Huffman codes, code-length RLE,
// dynamic-block header bits and the block's exact bit length.  One warp per block.
//
// Replaces generateDeflateHuffmanTable (/root/reference/src/huffman.ts:55-153:
// histogram + 15 (or 7) rounds of package-merge + canonical code assignment) and the
// header half of deflateDynamicBlock (/root/reference/src/deflate.ts:78-181).
//
// The reference's package-merge yields an optimal length-limited code.  Here the
// lengths come from an in-place Huffman construction over the sorted counts
// (Moffat/Katajainen) followed by a Kraft-sum repair when a length exceeds the
// limit; the two differ only when the unrestricted Huffman tree is deeper than 15
// (7) levels, which costs a few bits per block at most (sizes are gated at <= 1.03x
// the oracle's in tests/).  Canonical code assignment is the reference's
// (by length, then ascending symbol; src/huffman.ts:135-151).
#pragma once
#include "zles_dev.h"

namespace zles {

constexpr int HUF_WARPS = 4;
constexpr int HUF_THREADS = HUF_WARPS * 32;
constexpr u32 HDR_BYTES = 576;  // >= ceil((14 + 19*3 + 316*14) / 8) = 562
constexpr u32 HUF_STORED = 0xffffffffu;  // BlockCodes::hdr_nbits of a block that is smaller stored (BTYPE=0) than coded
constexpr u32 HUF_FIXED = 0xfffffffeu;   // ... of a block that is smallest with the fixed code (BTYPE=1): no header, ll/d hold the fixed codes
// code length of literal/length symbol s in the fixed code (RFC 1951 3.2.6; /root/reference/src/huffman.ts:41-53)
__host__ __device__ __forceinline__ u32 fixed_ll_len(u32 s) { return s < 144 ? 8u : s < 256 ? 9u : s < 280 ? 7u : 8u; }
// bytes a block of `bits` bits takes in the stream: padded to a byte when it ends the stream, else followed by the empty
// stored block that makes the next block start on a byte (3 bits + pad, LEN, NLEN); bit 31 set: `bits` holds the bytes already
__host__ __device__ __forceinline__ u32 seg_bytes_of(u32 bits, bool final_block) {
  if (bits & 0x80000000u) return bits & 0x7fffffffu;
  return final_block ? (bits + 7) >> 3 : ((bits + 3 + 7) >> 3) + 4;
}

struct BlockCodes {       // written by k_huff, read by k_pack
  u32 ll[288];            // (bit-reversed code << 8) | length, 0 for unused symbols
  u32 d[32];
  u32 hdr_nbits;          // HLIT..code lengths, starts right after BFINAL/BTYPE; HUF_STORED: emit the block stored
  u32 pad_[3];
  u8 hdr[HDR_BYTES];      // header bits, LSB first
};

struct HufWarpSmem {
  u32 freq[320];
  u32 skey[288];   // counts in ascending order, then the construction's work array
  u16 ssym[288];   // symbols in the same order
  u8 len[320];     // [0,288) literal/length, [288,320) distance
  u32 code[320];
  u32 clfreq[32];
  u8 cllen[32];
  u32 clcode[32];
  u16 rle[328];    // (symbol << 8) | extra value, one per code-length symbol
  u32 nrle;
  u32 cnt[16];     // codes per length / next code
  u32 hdr[HDR_BYTES / 4];
};
constexpr int HUF_SMEM = (int)sizeof(HufWarpSmem) * HUF_WARPS;

// Computes code lengths (<= maxlen) for n symbols (n <= 288, multiple of 32 not
// required) with counts freq[0..n).  Unused symbols get 0; a single used symbol gets 1
// (like src/huffman.ts:71-75).  Warp-cooperative; all lanes must call.
__device__ __forceinline__ void huf_lengths(HufWarpSmem *S, const u32 *freq, u32 n, u32 maxlen, u8 *len_out) {
  const u32 lane = lane_id();
  // compact the used symbols, ascending symbol order
  u32 nused = 0;
  for (u32 g = 0; g < n; g += 32) {
    u32 s = g + lane;
    u32 f = s < n ? freq[s] : 0;
    if (s < n) len_out[s] = 0;
    u32 bal = __ballot_sync(ZLES_FULL, f != 0);
    if (f) {
      u32 o = nused + __popc(bal & lanemask_lt());
      S->skey[o] = f;
      S->ssym[o] = (u16)s;
    }
    nused += __popc(bal);
  }
  __syncwarp();
  if (nused == 0) return;
  if (nused == 1) {
    if (lane == 0) len_out[S->ssym[0]] = 1;
    __syncwarp();
    return;
  }
  // rank sort by (count, symbol); up to 9 entries per lane
  u32 f[9], s[9], rk[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    u32 i = lane + 32 * k;
    f[k] = i < nused ? S->skey[i] : 0;
    s[k] = i < nused ? S->ssym[i] : 0;
    rk[k] = 0;
  }
  for (u32 j = 0; j < nused; j++) {
    u32 fj = S->skey[j];
#pragma unroll
    for (int k = 0; k < 9; k++) {
      u32 i = lane + 32 * k;
      rk[k] += (fj < f[k] || (fj == f[k] && j < i)) ? 1u : 0u;
    }
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 9; k++) {
    u32 i = lane + 32 * k;
    if (i < nused) { S->skey[rk[k]] = f[k]; S->ssym[rk[k]] = (u16)s[k]; }
  }
  __syncwarp();
  if (lane == 0) {
    // in-place minimum-redundancy code lengths over ascending counts (Moffat & Katajainen)
    u32 *A = S->skey;
    const int nn = (int)nused;
    A[0] += A[1];
    int root = 0, leaf = 2;
    for (int next = 1; next < nn - 1; next++) {
      if (leaf >= nn || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = (u32)next; }
      else A[next] = A[leaf++];
      if (leaf >= nn || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = (u32)next; }
      else A[next] += A[leaf++];
    }
    A[nn - 2] = 0;
    for (int next = nn - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
    int avbl = 1, used = 0, dpth = 0;
    root = nn - 2;
    int next = nn - 1;
    while (avbl > 0) {
      while (root >= 0 && (int)A[root] == dpth) { used++; root--; }
      while (avbl > used) { A[next--] = (u32)dpth; avbl--; }
      avbl = 2 * used;
      dpth++;
      used = 0;
    }
    // A[i] = length of the i-th least frequent symbol (non-increasing in i)
    u32 *cnt = S->cnt;
    for (u32 l = 0; l < 16; l++) cnt[l] = 0;
    for (int i = 0; i < nn; i++) cnt[A[i] > maxlen ? maxlen : A[i]]++;
    u32 total = 0;
    for (u32 l = maxlen; l > 0; l--) total += cnt[l] << (maxlen - l);
    while (total != (1u << maxlen)) {  // over-subscribed after clamping: lengthen the cheapest codes
      cnt[maxlen]--;
      for (u32 l = maxlen - 1; l > 0; l--)
        if (cnt[l]) { cnt[l]--; cnt[l + 1] += 2; break; }
      total--;
    }
    int j = nn;
    for (u32 l = 1; l <= maxlen; l++)
      for (u32 c = cnt[l]; c > 0; c--) len_out[S->ssym[--j]] = (u8)l;
  }
  __syncwarp();
}

// canonical codes from lengths, bit-reversed for LSB-first emission; n multiple of 32
__device__ __forceinline__ void huf_codes(HufWarpSmem *S, const u8 *len, u32 n, u32 *code_out) {
  const u32 lane = lane_id();
  u32 *cnt = S->cnt;
  if (lane < 16) cnt[lane] = 0;
  __syncwarp();
  for (u32 g = 0; g < n; g += 32) {
    u32 l = len[g + lane];
    u32 m = __match_any_sync(ZLES_FULL, l);
    if (l && lane == (u32)(__ffs((int)m) - 1)) cnt[l] += __popc(m);
    __syncwarp();
  }
  if (lane == 0) {  // cnt[l] := first code of length l
    u32 code = 0;
    for (u32 l = 1; l < 16; l++) {
      u32 c = cnt[l];
      cnt[l] = code;
      code = (code + c) << 1;
    }
  }
  __syncwarp();
  for (u32 g = 0; g < n; g += 32) {
    u32 l = len[g + lane];
    u32 m = __match_any_sync(ZLES_FULL, l);
    u32 base = cnt[l];
    __syncwarp();
    u32 c = 0;
    if (l) {
      u32 code = base + __popc(m & lanemask_lt());
      c = ((__brev(code) >> (32 - l)) << 8) | l;
      if (lane == (u32)(__ffs((int)m) - 1)) cnt[l] = base + __popc(m);
    }
    code_out[g + lane] = c;
    __syncwarp();
  }
}

__device__ __forceinline__ void huf_putbits(u32 *buf, u32 &pos, u32 v, u32 nb) {  // single thread
  if (!nb) return;
  u32 w = pos >> 5, sh = pos & 31;
  buf[w] |= v << sh;
  if (sh + nb > 32) buf[w + 1] |= v >> (32 - sh);
  pos += nb;
}

// Codes, header bits and coded size of one block from its symbol counts in S->freq (the end-of-block symbol is added
// here): leaves the codes in S->code, the header in S->hdr; hbits = header bits after BFINAL / BTYPE, bits = what the
// symbols and their extra bits take.  One warp.
__device__ __forceinline__ void huf_block(HufWarpSmem *S, u32 &hbits_out, u32 &bits_out) {
  const u32 lane = lane_id();
  __syncwarp();
  if (lane == 0) S->freq[256] = 1;  // EOB, src/deflate.ts:58
  __syncwarp();
  huf_lengths(S, S->freq, 288, 15, S->len);          // src/deflate.ts:78
  huf_lengths(S, S->freq + 288, 32, 15, S->len + 288);  // src/deflate.ts:79
  huf_codes(S, S->len, 288, S->code);
  huf_codes(S, S->len + 288, 32, S->code + 288);

  // HLIT / HDIST (src/deflate.ts:81-97) and the run-length coding of the lengths
  // (src/deflate.ts:99-139; here the standard 16/17/18 greedy, runs may span both alphabets)
  if (lane < 32) S->clfreq[lane] = 0;
  for (u32 i = lane; i < HDR_BYTES / 4; i += 32) S->hdr[i] = 0;
  __syncwarp();
  u32 hlit = 257, hdist = 1;
  if (lane == 0) {
    for (u32 i = 287; i >= 257; i--)
      if (S->len[i]) { hlit = i + 1; break; }
    for (u32 i = 31; i >= 1; i--)
      if (S->len[288 + i]) { hdist = i + 1; break; }
    const u32 total = hlit + hdist;
    u32 nr = 0;
    for (u32 i = 0; i < total;) {
      const u32 v = i < hlit ? S->len[i] : S->len[288 + i - hlit];
      u32 run = 1;
      while (i + run < total) {
        u32 j = i + run;
        u32 vv = j < hlit ? S->len[j] : S->len[288 + j - hlit];
        if (vv != v) break;
        run++;
      }
      i += run;
      if (v == 0) {
        while (run >= 11) { u32 r = run > 138 ? 138 : run; S->rle[nr++] = (u16)((18 << 8) | (r - 11)); S->clfreq[18]++; run -= r; }
        if (run >= 3) { S->rle[nr++] = (u16)((17 << 8) | (run - 3)); S->clfreq[17]++; run = 0; }
        while (run--) { S->rle[nr++] = 0; S->clfreq[0]++; }
      } else {
        S->rle[nr++] = (u16)(v << 8); S->clfreq[v]++; run--;
        while (run >= 3) { u32 r = run > 6 ? 6 : run; S->rle[nr++] = (u16)((16 << 8) | (r - 3)); S->clfreq[16]++; run -= r; }
        while (run--) { S->rle[nr++] = (u16)(v << 8); S->clfreq[v]++; }
      }
    }
    S->nrle = nr;
  }
  __syncwarp();
  hlit = __shfl_sync(ZLES_FULL, hlit, 0);
  hdist = __shfl_sync(ZLES_FULL, hdist, 0);
  huf_lengths(S, S->clfreq, 19, 7, S->cllen);  // src/deflate.ts:141
  if (lane >= 19) S->cllen[lane] = 0;
  __syncwarp();
  huf_codes(S, S->cllen, 32, S->clcode);

  u32 hbits = 0;
  if (lane == 0) {  // header bits, src/deflate.ts:150-181
    u32 hclen = 4;
    for (u32 i = 0; i < 19; i++)
      if (S->cllen[c_cl_order[i]]) hclen = i + 1 > 4 ? i + 1 : 4;
    u32 pos = 0;
    huf_putbits(S->hdr, pos, hlit - 257, 5);
    huf_putbits(S->hdr, pos, hdist - 1, 5);
    huf_putbits(S->hdr, pos, hclen - 4, 4);
    for (u32 i = 0; i < hclen; i++) huf_putbits(S->hdr, pos, S->cllen[c_cl_order[i]], 3);
    for (u32 i = 0; i < S->nrle; i++) {
      u32 sym = S->rle[i] >> 8, ev = S->rle[i] & 255;
      u32 c = S->clcode[sym];
      huf_putbits(S->hdr, pos, c >> 8, c & 255);
      if (sym == 16) huf_putbits(S->hdr, pos, ev, 2);
      else if (sym == 17) huf_putbits(S->hdr, pos, ev, 3);
      else if (sym == 18) huf_putbits(S->hdr, pos, ev, 7);
    }
    hbits = pos;
  }
  __syncwarp();
  hbits = __shfl_sync(ZLES_FULL, hbits, 0);

  // exact size of the block: BFINAL/BTYPE + header + sum count * (code length + extra bits)
  u32 bits = 0;
  for (u32 i = lane; i < 320; i += 32) {
    u32 e = 0;
    if (i >= 257 && i < 286) e = c_len_extra[i - 257];
    else if (i >= 288 && i < 318) e = c_dist_extra[i - 288];
    bits += S->freq[i] * (S->len[i] + e);
  }
  bits = __reduce_add_sync(ZLES_FULL, bits);
  hbits_out = hbits;
  bits_out = bits;
}

constexpr u32 BLK_BYTES = 0x80000000u;   // blk_bits[]: bit 31 set = the low bits are the block's BYTES in the stream, marker or final pad
                                         // included (a chunk written as one block: its first block carries all of it, the others 0)
constexpr u32 HUF_MERGE_MAX = 8192;      // chunks that compress to less than this many bytes are tried as one block
constexpr u32 HUF_MERGED = 0x4d524730u;  // BlockCodes::pad_[0] of the first block of such a chunk; pad_[1] = its blocks

// own_len[b] = input bytes of block b: n and table describe them as in LzParams (one stream, or a batch table).
__global__ void __launch_bounds__(HUF_THREADS) k_huff(const u32 *__restrict__ hist, u32 first_block, u32 nblocks, BlockCodes *codes, u32 *blk_bits,
                                                      u64 n, const BatchBlk *__restrict__ table) {
  ZLES_SMEM_DECL(smem_raw);
  HufWarpSmem *S = reinterpret_cast<HufWarpSmem *>(smem_raw) + warp_id();
  const u32 lane = lane_id();
  const u32 b = first_block + blockIdx.x * HUF_WARPS + warp_id();
  if (b >= nblocks) return;  // whole warp leaves; no CTA barrier below
  for (u32 i = lane; i < 320; i += 32) S->freq[i] = hist[(size_t)b * 320 + i];
  u32 hbits, bits;
  huf_block(S, hbits, bits);
  // The same symbols under the fixed code (BTYPE=1: no header at all).  The reference never writes it (src/deflate.ts:28);
  // it wins on inputs of a few hundred bytes, where the dynamic header is most of the block.
  u32 fbits = 0;
  for (u32 i = lane; i < 320; i += 32) {
    u32 e = 0, l = 5;
    if (i < 288) l = fixed_ll_len(i);
    if (i >= 257 && i < 286) e = c_len_extra[i - 257];
    else if (i >= 288 && i < 318) e = c_dist_extra[i - 288];
    fbits += S->freq[i] * (l + e);
  }
  fbits = __reduce_add_sync(ZLES_FULL, fbits);
  const u32 own_len = table ? table[b].own_len : (u32)umin64((u64)SUB, n - (u64)b * SUB);
  // The reference always emits BTYPE=2 (src/deflate.ts:28) and so expands incompressible data; a stored block
  // (3 header bits + pad, LEN, NLEN, the bytes: RFC 1951 3.2.4) is taken when it is smaller.  Blocks start on a
  // byte boundary, so a stored block costs exactly 5 + own_len bytes.
  const u32 coded = 3 + hbits + bits, fixed = 3 + fbits, stored = 8 * (5 + own_len);
  const u32 best = umin(coded, fixed);
  const bool use_stored = stored < ((best + 7) & ~7u), use_fixed = !use_stored && fixed < coded;
  BlockCodes *C = codes + b;
  if (use_fixed) {
    // canonical fixed codes, bit-reversed like huf_codes': 8-bit 0x30.. for 0-143, 9-bit 0x190.. for 144-255,
    // 7-bit 0 .. for 256-279, 8-bit 0xC0.. for 280-287; distances are their own 5-bit numbers
    for (u32 i = lane; i < 288; i += 32) {
      const u32 l = fixed_ll_len(i);
      const u32 code = i < 144 ? 0x30 + i : i < 256 ? 0x190 + (i - 144) : i < 280 ? i - 256 : 0xC0 + (i - 280);
      C->ll[i] = ((__brev(code) >> (32 - l)) << 8) | l;
    }
    C->d[lane] = ((__brev(lane) >> 27) << 8) | 5;
  } else {
    for (u32 i = lane; i < 288; i += 32) C->ll[i] = S->code[i];
    C->d[lane] = S->code[288 + lane];
    for (u32 i = lane; i < HDR_BYTES / 4; i += 32) reinterpret_cast<u32 *>(C->hdr)[i] = S->hdr[i];
  }
  if (lane == 0) {
    C->hdr_nbits = use_stored ? HUF_STORED : use_fixed ? HUF_FIXED : hbits;
    blk_bits[b] = use_stored ? stored : use_fixed ? fixed : coded;
  }
}

// One block per chunk where four headers cost more than they save: the reference writes one dynamic block per 128 KiB
// chunk (/root/reference/src/deflate.ts:20-34), we write four — on input that compresses to a few hundred bytes per
// chunk (zeros, short periods) the three extra headers and markers are a third of the output.  A warp per chunk whose
// four blocks are all coded and small: the code of the summed counts, and if header + symbols + one marker come out
// smaller than the four blocks, the chunk's first block takes the code and the whole chunk's bytes, the others nothing
// (k_pack then writes one header, the tokens of all four, one end-of-block code, one marker).  Such a stream is still
// plain zlib, but no longer a run of 32 KiB blocks: our inflate decodes it with the block-parallel tier for other
// encoders' streams (inflate_fblk.cuh).
__global__ void __launch_bounds__(HUF_THREADS) k_huff_merge(const u32 *__restrict__ hist, u32 first_block, u32 end_block, u32 nblocks_total,
                                                            u32 last_is_final, BlockCodes *codes, u32 *blk_bits) {
  ZLES_SMEM_DECL(smem_raw);
  HufWarpSmem *S = reinterpret_cast<HufWarpSmem *>(smem_raw) + warp_id();
  const u32 lane = lane_id();
  const u32 b0 = (first_block / SUBS_PER_CHUNK + blockIdx.x * HUF_WARPS + warp_id()) * SUBS_PER_CHUNK;
  if (b0 < first_block || b0 >= end_block) return;
  const u32 cnt = umin(SUBS_PER_CHUNK, nblocks_total - b0);
  if (cnt < 2 || b0 + cnt > end_block) return;
  // what the blocks take on their own
  u32 sep = 0;
  bool ok = true;
  for (u32 k = 0; k < cnt; k++) {
    if (codes[b0 + k].hdr_nbits == HUF_STORED) ok = false;
    sep += seg_bytes_of(blk_bits[b0 + k], last_is_final && b0 + k + 1 == nblocks_total);
  }
  if (!ok || sep >= HUF_MERGE_MAX) return;
  for (u32 i = lane; i < 320; i += 32) {
    u32 f = 0;
    for (u32 k = 0; k < cnt; k++) f += hist[(size_t)(b0 + k) * 320 + i];
    S->freq[i] = f;
  }
  __syncwarp();
  u32 hbits, bits;
  huf_block(S, hbits, bits);  // (sets the end-of-block count to 1)
  const u32 merged = seg_bytes_of(3 + hbits + bits, last_is_final && b0 + cnt == nblocks_total);
  if (merged >= sep) return;
  BlockCodes *C = codes + b0;
  for (u32 i = lane; i < 288; i += 32) C->ll[i] = S->code[i];
  C->d[lane] = S->code[288 + lane];
  for (u32 i = lane; i < HDR_BYTES / 4; i += 32) reinterpret_cast<u32 *>(C->hdr)[i] = S->hdr[i];
  if (lane == 0) {
    C->hdr_nbits = hbits;
    C->pad_[0] = HUF_MERGED;
    C->pad_[1] = cnt;
    blk_bits[b0] = BLK_BYTES | merged;
    for (u32 k = 1; k < cnt; k++) blk_bits[b0 + k] = BLK_BYTES;
  }
}

}  // namespace zles
