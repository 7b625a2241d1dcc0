// pack.cuh — kernels K4/K5: output layout (exclusive scan of compressed sizes) and the
// bit-packing encoder.
//
// Replaces the sequential BitWriteStream of the reference
// (/root/reference/src/utils/BitWriteStream.ts:14-46, one call per output bit) and the
// emit loops of deflateDynamicBlock (/root/reference/src/deflate.ts:150-226) plus the block
// loop of deflate (/root/reference/src/deflate.ts:20-38).
//
// Layout of one deflate block ("segment") in the stream.  Every block is byte aligned, so
// blocks are written independently — to local memory or straight into a peer GPU's buffer —
// and can be located and decoded in parallel by inflate.cuh:
//   [dynamic block, BFINAL=0]  000 + pad to byte + 00 00 FF FF   (empty stored block = sync marker)
// The last block of the stream sets BFINAL, is padded to a byte (src/deflate.ts:35-37) and has
// no marker.  A block of chunk-relative index 1..3 may reference the 32 KiB before it (same
// chunk); index 0 references nothing outside itself.
//
// K5 computes every token's code + extra bits (<= 48 bits), a block-wide exclusive
// scan of the bit lengths gives each token its bit offset, tokens are OR-ed into a
// shared-memory staging window and the window is written out with aligned,
// coalesced 32-bit stores (head/tail bytes of a block with byte stores).
#pragma once
#include "huffman.cuh"
#include "zles_dev.h"

namespace zles {

constexpr int PACK_THREADS = 512;
constexpr int PACK_ITEMS = 4;                                   // tokens per thread per tile
constexpr u32 PACK_TILE = PACK_THREADS * PACK_ITEMS;            // 2048 tokens
constexpr u32 PACK_STAGE_WORDS = PACK_TILE * 48 / 32 + 8;       // 3080 words
constexpr u32 PACK_SMEM = PACK_STAGE_WORDS * 4 + 320 * 4 + 40 * 4 + 48 * 8;  // stage | code table | scan scratch | reduction

struct LayoutParams {
  const u32 *blk_bits;   // [nblocks]
  u32 nblocks;
  u32 last_is_final;     // this shard ends the stream
  u64 n;                 // shard length in bytes
  const u64 *adler_part; // [nblocks][2]
  u64 *blk_off;          // [nblocks + 1] byte offsets relative to the shard's first byte; [nblocks] = total
  u64 *summary;          // [0]=total bytes, [1]=sum d mod p, [2]=sum (n - i) d[i] mod p (i local to the shard)
};

__host__ __device__ __forceinline__ u32 seg_bytes(u32 bits, bool final_block) { return seg_bytes_of(bits, final_block); }

// single CTA
__global__ void __launch_bounds__(1024) k_layout(const LayoutParams P) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u64 *red = reinterpret_cast<u64 *>(smem_raw + 256);
  u64 carry = 0;
  u64 sa = 0, sb = 0;
  for (u32 base = 0; base < P.nblocks; base += 1024) {
    const u32 b = base + threadIdx.x;
    u32 bytes = 0;
    if (b < P.nblocks) {
      bytes = seg_bytes(P.blk_bits[b], P.last_is_final && b + 1 == P.nblocks);
      const u64 off = (u64)b * SUB;
      const u64 len = umin64((u64)SUB, P.n - off);
      const u64 A = P.adler_part[2 * (size_t)b], B = P.adler_part[2 * (size_t)b + 1];
      sa += A;
      sb += (B + ((P.n - off - len) % ADLER_MOD) * (A % ADLER_MOD)) % ADLER_MOD;
    }
    u32 total;
    u32 ex = block_exscan(bytes, scratch, &total);
    if (b < P.nblocks) P.blk_off[b] = carry + ex;
    carry += total;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sa += __shfl_down_sync(ZLES_FULL, sa, d);
    sb += __shfl_down_sync(ZLES_FULL, sb, d);
  }
  __syncthreads();
  if (lane_id() == 0) { red[warp_id()] = sa; red[32 + warp_id()] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 ta = 0, tb = 0;
    for (int w = 0; w < 32; w++) { ta += red[w]; tb += red[32 + w]; }
    P.blk_off[P.nblocks] = carry;
    P.summary[0] = carry;
    P.summary[1] = ta % ADLER_MOD;
    P.summary[2] = tb % ADLER_MOD;
  }
}
constexpr u32 LAYOUT_SMEM = 256 + 64 * 8;

// Byte offsets of the blocks [b0, b1) continuing from *carry (the end of the previous slab): lets the host-buffer
// deflate pack and copy out a slab while the matcher works on the next one (zles_deflate).  Single CTA.
__global__ void __launch_bounds__(1024) k_layout_slab(const u32 *__restrict__ blk_bits, u32 b0, u32 b1, u32 nblocks, u32 last_is_final,
                                                      u64 *carry_io, u64 *blk_off, u64 *slab_end) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u64 carry = *carry_io;
  __syncthreads();  // everybody has read the carry before thread 0 rewrites it
  for (u32 base = b0; base < b1; base += 1024) {
    const u32 b = base + threadIdx.x;
    const u32 bytes = b < b1 ? seg_bytes(blk_bits[b], last_is_final && b + 1 == nblocks) : 0;
    u32 total;
    const u32 ex = block_exscan(bytes, scratch, &total);
    if (b < b1) blk_off[b] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    blk_off[b1] = carry;
    *carry_io = carry;
    *slab_end = carry;
  }
}

struct PackParams {
  const u32 *tokens;       // [nblocks][SUB]
  const u32 *ntok;         // [nblocks]
  const BlockCodes *codes; // [nblocks]
  const u32 *blk_bits;     // [nblocks]
  const u64 *blk_off;      // [nblocks + 1]
  u32 nblocks;
  u32 first_block = 0;     // this launch packs blocks first_block + blockIdx.x
  u32 last_is_final;
  u8 *out;                 // destination of this shard's first byte (local or peer memory)
  const u8 *in;            // the shard's input (stored blocks copy from it)
  u64 n;
};

struct PackState {
  u32 *stage;     // staging words; word 0 holds the carry
  u32 *scratch;   // block_exscan scratch
  u32 *gw;        // aligned word pointer at/below the first byte
  u32 head;       // bytes of gw[0] that belong to somebody else
  u32 wcur;       // words already written
  u32 carry_bits; // valid bits in stage[0]
};

__device__ __forceinline__ void pack_begin(PackState &st, u32 *stage, u32 *scratch, u8 *dst) {
  st.stage = stage;
  st.scratch = scratch;
  st.head = (u32)((uintptr_t)dst & 3);
  st.gw = reinterpret_cast<u32 *>(dst - st.head);
  st.wcur = 0;
  st.carry_bits = st.head * 8;  // the foreign bytes of gw[0] are never stored (see pack_emit)
  for (u32 i = threadIdx.x; i < PACK_STAGE_WORDS; i += PACK_THREADS) stage[i] = 0;
  __syncthreads();
}

// All threads contribute up to PACK_ITEMS consecutive bit strings (nb may be 0);
// the strings are concatenated in thread order and appended to the output.
__device__ __forceinline__ void pack_emit(PackState &st, const u64 *bits, const u32 *nb) {
  u32 mine = 0;
#pragma unroll
  for (int k = 0; k < PACK_ITEMS; k++) mine += nb[k];
  u32 total;
  u32 off = st.carry_bits + block_exscan(mine, st.scratch, &total);
#pragma unroll
  for (int k = 0; k < PACK_ITEMS; k++) {
    if (nb[k]) {
      const u32 w = off >> 5, sh = off & 31;
      const u64 v = bits[k] << sh;
      atomicOr(st.stage + w, (u32)v);
      if (sh + nb[k] > 32) atomicOr(st.stage + w + 1, (u32)(v >> 32));
      if (sh + nb[k] > 64) atomicOr(st.stage + w + 2, (u32)(bits[k] >> (64 - sh)));
      off += nb[k];
    }
  }
  __syncthreads();
  const u32 avail = st.carry_bits + total;
  const u32 nfull = avail >> 5;
  // full words go out with 128-bit stores (coalesced: a warp writes 512 contiguous bytes per instruction — to local HBM
  // or, through a peer-mapped pointer, over NVLink); the words before the first 16-byte boundary of the destination and
  // the up to three after the last one with 32-bit stores, the block's first partial word byte by byte
  {
    u32 *g = st.gw + st.wcur;
    u32 pre = (u32)((16 - ((uintptr_t)g & 15)) & 15) >> 2;  // words up to the 16-byte boundary
    if (st.wcur == 0 && st.head && pre == 0) pre = 4;       // the block's first word holds somebody else's bytes: never in a vector
    pre = umin(pre, nfull);
    const u32 nvec = (nfull - pre) >> 2;
    for (u32 i = threadIdx.x; i < pre; i += PACK_THREADS) {
      const u32 v = st.stage[i];
      if (st.wcur + i == 0 && st.head) {
        u8 *p = reinterpret_cast<u8 *>(st.gw);
        for (u32 k = st.head; k < 4; k++) p[k] = (u8)(v >> (8 * k));
      } else {
        g[i] = v;
      }
    }
    uint4 *g4 = reinterpret_cast<uint4 *>(g + pre);
    for (u32 i = threadIdx.x; i < nvec; i += PACK_THREADS) {
      const u32 *sp = st.stage + pre + 4 * i;
      g4[i] = make_uint4(sp[0], sp[1], sp[2], sp[3]);
    }
    for (u32 i = pre + 4 * nvec + threadIdx.x; i < nfull; i += PACK_THREADS) g[i] = st.stage[i];
  }
  const u32 carry = st.stage[nfull];
  __syncthreads();
  for (u32 i = threadIdx.x; i <= nfull + 2 && i < PACK_STAGE_WORDS; i += PACK_THREADS) st.stage[i] = (i == 0) ? carry : 0u;
  st.wcur += nfull;
  st.carry_bits = avail & 31;
  __syncthreads();
}

// the output ends on a byte boundary; write the bytes left in the carry word
__device__ __forceinline__ void pack_finish(PackState &st) {
  if (threadIdx.x == 0) {
    const u32 nbytes = st.carry_bits >> 3;
    const u32 v = st.stage[0];
    u8 *p = reinterpret_cast<u8 *>(st.gw + st.wcur);
    for (u32 k = (st.wcur == 0 ? st.head : 0); k < nbytes; k++) p[k] = (u8)(v >> (8 * k));
  }
}

// Appends deflate blocks [b0, b1) bit-concatenated and then either the final pad
// (src/deflate.ts:35-37) or the empty stored block that makes the next block byte aligned.
// raw[b] = the block's own input bytes (for stored blocks), raw_len = their number.
// merged: the blocks are written as ONE block — the header and code of block b0, the tokens of all, one end-of-block code
// (k_huff_merge).
__device__ __forceinline__ void pack_chunk(PackState &st, u32 *ctab, const u32 *tokens, const u32 *ntok, const BlockCodes *codes,
                                           u32 b0, u32 b1, bool final_chunk, const u8 *raw, u32 raw_len, bool merged = false,
                                           u32 tok_stride = SUB) {
  const u32 tid = threadIdx.x;
  u64 bits[PACK_ITEMS];
  u32 nb[PACK_ITEMS];
  for (u32 b = b0; b < b1; b++) {
    const BlockCodes *C = codes + (merged ? b0 : b);
    if (!merged && C->hdr_nbits == HUF_STORED) {  // BTYPE=0: header byte, LEN, NLEN, the bytes; then the marker unless final
      const bool fin = final_chunk && b + 1 == b1;
      const u32 nitems = 2 + (raw_len + 3) / 4 + (fin ? 0 : 2);  // 8 + 32 header bits | payload words | 8 + 32 marker bits
      for (u32 base = 0; base < nitems; base += PACK_TILE) {
#pragma unroll
        for (int k = 0; k < PACK_ITEMS; k++) {
          const u32 i = base + tid * PACK_ITEMS + k;
          bits[k] = 0; nb[k] = 0;
          if (i == 0) { bits[k] = fin ? 1u : 0u; nb[k] = 8; }                          // BFINAL, BTYPE=00, 5 pad bits
          else if (i == 1) { bits[k] = raw_len | ((raw_len ^ 0xffffu) << 16); nb[k] = 32; }  // LEN, NLEN
          else if (i - 2 < (raw_len + 3) / 4) {
            const u32 o = (i - 2) * 4, cnt = umin(4u, raw_len - o);
            u32 v = 0;
            for (u32 q = 0; q < cnt; q++) v |= (u32)raw[o + q] << (8 * q);
            bits[k] = v; nb[k] = 8 * cnt;
          } else if (i < nitems) {
            const u32 m = i - 2 - (raw_len + 3) / 4;                                     // the sync marker: 00 | 00 00 FF FF
            if (m == 0) { bits[k] = 0; nb[k] = 8; } else { bits[k] = 0xFFFF0000u; nb[k] = 32; }
          }
        }
        pack_emit(st, bits, nb);
      }
      continue;
    }
    if (!merged || b == b0) {
      for (u32 i = tid; i < 320; i += PACK_THREADS) ctab[i] = i < 288 ? C->ll[i] : C->d[i - 288];
      __syncthreads();
    }
    // block header: BFINAL, BTYPE=2 (src/deflate.ts:21-28), then the code-length header in 32-bit pieces
    const bool fixed = C->hdr_nbits == HUF_FIXED;  // BTYPE=1: the three bits are the whole header (k_huff)
    const u32 hbits = fixed ? 0u : C->hdr_nbits;
    const u32 hwords = (merged && b != b0) ? 0xffffffffu : (hbits + 31) >> 5;  // (no header inside a merged chunk)
    const u32 bfinal = (final_chunk && (merged || b + 1 == b1)) ? 1u : 0u;
    for (u32 base = 0; hwords != 0xffffffffu && base < hwords + 1; base += PACK_TILE) {
#pragma unroll
      for (int k = 0; k < PACK_ITEMS; k++) {
        const u32 i = base + tid * PACK_ITEMS + k;
        bits[k] = 0; nb[k] = 0;
        if (i == 0) { bits[k] = bfinal | ((fixed ? 1u : 2u) << 1); nb[k] = 3; }
        else if (i <= hwords) {
          bits[k] = reinterpret_cast<const u32 *>(C->hdr)[i - 1];
          nb[k] = umin(32u, hbits - (i - 1) * 32);
        }
      }
      pack_emit(st, bits, nb);
    }
    // tokens, src/deflate.ts:183-220
    const u32 nt = ntok[b];
    const u32 *tok = tokens + (size_t)b * tok_stride;
    for (u32 base = 0; base < nt; base += PACK_TILE) {
#pragma unroll
      for (int k = 0; k < PACK_ITEMS; k++) {
        const u32 i = base + tid * PACK_ITEMS + k;
        bits[k] = 0; nb[k] = 0;
        if (i < nt) {
          const u32 t = tok[i];
          if (t & 0x80000000u) {
            const u32 len = ((t >> 16) & 255) + 3, dist = (t & 0x7fff) + 1;
            u32 ls, le, lv, ds, de, dv;
            len_to_sym(len, ls, le, lv);
            dist_to_sym(dist, ds, de, dv);
            const u32 cl = ctab[257 + ls], cd = ctab[288 + ds];
            u64 v = cl >> 8;
            u32 n = cl & 255;
            v |= (u64)lv << n; n += le;
            v |= (u64)(cd >> 8) << n; n += cd & 255;
            v |= (u64)dv << n; n += de;
            bits[k] = v; nb[k] = n;
          } else {
            const u32 cl = ctab[t];
            bits[k] = cl >> 8; nb[k] = cl & 255;
          }
        }
      }
      pack_emit(st, bits, nb);
    }
    // end of block (src/deflate.ts:222-226); after the chunk's last block the sync marker or the final pad
    if (!merged || b + 1 == b1) {
      const u32 cur = (st.wcur << 5) + st.carry_bits;  // bit position inside the aligned word stream
#pragma unroll
      for (int k = 0; k < PACK_ITEMS; k++) { bits[k] = 0; nb[k] = 0; }
      if (tid == 0) {
        const u32 ce = ctab[256];
        bits[0] = ce >> 8; nb[0] = ce & 255;
        if (b + 1 == b1) {
          const u32 after = cur + nb[0];
          if (final_chunk) {
            bits[1] = 0; nb[1] = (0u - after) & 7;  // src/deflate.ts:35-37
          } else {
            bits[1] = 0; nb[1] = 3 + ((0u - (after + 3)) & 7);  // BFINAL=0, BTYPE=00, pad to byte
            bits[2] = 0xFFFF0000u; nb[2] = 32;                   // LEN=0000, NLEN=FFFF
          }
        }
      }
      pack_emit(st, bits, nb);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PACK_THREADS) k_pack(const PackParams P) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *stage = reinterpret_cast<u32 *>(smem_raw);
  u32 *ctab = stage + PACK_STAGE_WORDS;  // [320]
  u32 *scratch = ctab + 320;
  const u32 b = P.first_block + blockIdx.x;
  if (b >= P.nblocks) return;
  if (P.blk_bits[b] == BLK_BYTES) return;  // inside a chunk written as one block: its first block's CTA writes it
  PackState st;
  pack_begin(st, stage, scratch, P.out + P.blk_off[b]);
  if (P.blk_bits[b] & BLK_BYTES) {
    const u32 cnt = P.codes[b].pad_[1];
    pack_chunk(st, ctab, P.tokens, P.ntok, P.codes, b, b + cnt, P.last_is_final && b + cnt == P.nblocks, P.in, 0, true);
  } else {
    pack_chunk(st, ctab, P.tokens, P.ntok, P.codes, b, b + 1, P.last_is_final && b + 1 == P.nblocks, P.in + (u64)b * SUB,
               (u32)umin64((u64)SUB, P.n - (u64)b * SUB));
  }
  pack_finish(st);
}

// ---- batches of independent buffers: every buffer is a complete zlib stream ---------------
// (header 78 9C, its chunks, Adler-32 trailer: /root/reference/src/zlib.ts:25-49 per buffer)

// blk_first[i] = index of buffer i's first deflate block; blk_first[count] = total; blk_first[count + 1] = the longest
// block of the batch in bytes (= the token slots a block can need).  Single CTA.
__global__ void __launch_bounds__(1024) k_batch_count(const u64 *__restrict__ in_off, u32 count, u64 *blk_first) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u64 carry = 0;
  u32 longest = 0;
  for (u32 base = 0; base < count; base += 1024) {
    const u32 i = base + threadIdx.x;
    u32 nb = 0;
    if (i < count) {
      const u64 len = in_off[i + 1] - in_off[i];
      nb = (u32)((len + SUB - 1) / SUB);
      if (nb == 0) nb = 1;
      longest = umax(longest, (u32)umin64(len, (u64)SUB));
    }
    u32 total;
    const u32 ex = block_exscan(nb, scratch, &total);
    if (i < count) blk_first[i] = carry + ex;
    carry += total;
  }
  __syncthreads();
  if (threadIdx.x == 0) { blk_first[count] = carry; scratch[0] = 0; }
  __syncthreads();
  atomicMax(scratch, longest);
  __syncthreads();
  if (threadIdx.x == 0) blk_first[count + 1] = scratch[0];
}

// A buffer of more than one block: every block but a chunk's first gets the 32 KiB before it as window.  (The stream form
// leaves a chunk's third block without one so that blocks {2,3} share a sort — LzParams::pair_mode — and repairs the cases
// where that costs size with a second pass; a batch matches the blocks of a long buffer one at a time anyway, so the
// window is free.  A buffer of at most 64 KiB compresses to the same bytes alone and in a batch; a longer one may come out
// smaller in a batch.)  pair_mode is accepted and ignored.
__global__ void __launch_bounds__(256) k_batch_table(const u64 *__restrict__ in_off, u32 count, const u64 *__restrict__ blk_first,
                                                     BatchBlk *table, u32 pair_mode) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const u64 beg = in_off[i], len = in_off[i + 1] - beg;
  const u64 b0 = blk_first[i], nb = blk_first[i + 1] - b0;
  for (u64 k = 0; k < nb; k++) {
    BatchBlk t;
    t.in_off = beg + k * SUB;
    t.own_len = (u32)umin64((u64)SUB, len - k * SUB);
    const u32 kc = (u32)(k % SUBS_PER_CHUNK);
    t.hist_len = kc == 0 ? 0 : SUB;
    (void)pair_mode;
    table[b0 + k] = t;
  }
}

// ---- host forms of the batch calls: only the bytes that were produced travel back ----------------------------
// coff[i] = where buffer i's result goes in the compacted output (exclusive scan of the produced lengths; a buffer whose
// status is not 0 produced nothing); coff[count] = total.  Single CTA.
__global__ void __launch_bounds__(1024) k_batch_prefix(const u64 *__restrict__ out_len, const int32_t *__restrict__ status, u32 count, u64 *coff) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *scratch = reinterpret_cast<u32 *>(smem_raw);
  u64 carry = 0;
  for (u32 base = 0; base < count; base += 1024) {
    const u32 i = base + threadIdx.x;
    const u32 v = (i < count && status[i] == 0) ? (u32)out_len[i] : 0;  // one result is < 4 GiB (a slab holds less than that)
    u32 total;
    const u32 ex = block_exscan(v, scratch, &total);
    if (i < count) coff[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) coff[count] = carry;
}

// one CTA per buffer: its result from out + out_off[i] to packed + coff[i]
__global__ void __launch_bounds__(128) k_batch_gather(const u8 *__restrict__ out, const u64 *__restrict__ out_off, const u64 *__restrict__ coff,
                                                      u32 count, u8 *packed) {
  for (u32 i = blockIdx.x; i < count; i += gridDim.x) {
    const u64 n = coff[i + 1] - coff[i];
    const u8 *src = out + out_off[i];
    u8 *dst = packed + coff[i];
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      const u64 nv = n >> 4;
      for (u64 k = threadIdx.x; k < nv; k += blockDim.x) reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(src)[k];
      for (u64 k = (nv << 4) + threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
    } else {
      for (u64 k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
    }
  }
}

struct BatchPackParams {
  const u32 *tokens;
  const u32 *ntok;
  const BlockCodes *codes;
  const u32 *blk_bits;
  const u64 *adler_part;   // [nblocks][2]
  const BatchBlk *table;
  const u64 *blk_first;    // [count + 1]
  const u64 *in_off;       // [count + 1]
  const u64 *out_off;      // [count + 1]
  u32 count;
  const u8 *in;            // the batch's input (stored blocks copy from it)
  u8 *out;
  u64 *out_len;            // [count]
  int32_t *status;         // [count]
  u32 *first_err;          // lowest non-zero status of the batch (0 if none)
  u32 tok_stride = SUB;    // token slots per block (LzParams::tok_stride)
};

// one CTA per buffer
__global__ void __launch_bounds__(PACK_THREADS) k_pack_batch(const BatchPackParams P) {
  ZLES_SMEM_DECL(smem_raw);
  u32 *stage = reinterpret_cast<u32 *>(smem_raw);
  u32 *ctab = stage + PACK_STAGE_WORDS;
  u32 *scratch = ctab + 320;
  u64 *red = reinterpret_cast<u64 *>(scratch + 40);  // [3][16]
  const u32 i = blockIdx.x, tid = threadIdx.x;
  if (i >= P.count) return;
  const u64 len = P.in_off[i + 1] - P.in_off[i];
  const u32 b0 = (u32)P.blk_first[i], b1 = (u32)P.blk_first[i + 1];

  // size of the stream and Adler-32 sums of the buffer
  u64 bytes = 0, sa = 0, sb = 0;
  for (u32 b = b0 + tid; b < b1; b += PACK_THREADS) {
    const u64 off = (u64)(b - b0) * SUB;
    const u64 A = P.adler_part[2 * (size_t)b], B = P.adler_part[2 * (size_t)b + 1];
    sa += A;
    sb += (B + ((len - off - P.table[b].own_len) % ADLER_MOD) * (A % ADLER_MOD)) % ADLER_MOD;
    bytes += seg_bytes(P.blk_bits[b], b + 1 == b1);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    bytes += __shfl_down_sync(ZLES_FULL, bytes, d);
    sa += __shfl_down_sync(ZLES_FULL, sa, d);
    sb += __shfl_down_sync(ZLES_FULL, sb, d);
  }
  if (lane_id() == 0) { red[warp_id()] = bytes; red[16 + warp_id()] = sa; red[32 + warp_id()] = sb; }
  __syncthreads();
  bytes = 0; sa = 0; sb = 0;
  for (int w = 0; w < PACK_THREADS / 32; w++) { bytes += red[w]; sa += red[16 + w]; sb += red[32 + w]; }
  const u64 need = bytes + 6;
  const u64 cap = P.out_off[i + 1] - P.out_off[i];
  if (tid == 0) {
    P.out_len[i] = need;
    P.status[i] = need > cap ? 16 /* ZLES_E_OUTPUT_FULL */ : 0;
    if (need > cap) atomicMax(P.first_err, 16u);
  }
  if (need > cap) return;
  const u32 s1 = (u32)((1 + sa) % ADLER_MOD);
  const u32 s2 = (u32)((len % ADLER_MOD + sb % ADLER_MOD) % ADLER_MOD);
  const u32 adler = (s2 << 16) | s1;

  PackState st;
  pack_begin(st, stage, scratch, P.out + P.out_off[i]);
  u64 bits[PACK_ITEMS];
  u32 nb[PACK_ITEMS];
#pragma unroll
  for (int k = 0; k < PACK_ITEMS; k++) { bits[k] = 0; nb[k] = 0; }
  if (tid == 0) { bits[0] = 0x9C78u; nb[0] = 16; }  // CMF = 78, FLG = 9C (src/zlib.ts:28-34)
  pack_emit(st, bits, nb);
  for (u32 b = b0; b < b1; b++)
    pack_chunk(st, ctab, P.tokens, P.ntok, P.codes, b, b + 1, b + 1 == b1, P.in + P.table[b].in_off, P.table[b].own_len, false, P.tok_stride);
  nb[0] = 0;
  if (tid == 0) {  // big-endian Adler-32 (src/zlib.ts:36-40)
    bits[0] = ((adler >> 24) & 0xff) | ((adler >> 8) & 0xff00) | ((adler << 8) & 0xff0000) | (adler << 24);
    nb[0] = 32;
  }
  pack_emit(st, bits, nb);
  pack_finish(st);
}

}  // namespace zles
