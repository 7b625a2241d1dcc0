// inflate_spec.cuh — phase A of inflate with four warps per 32 KiB block.
//
// k_inf_tokens (inflate.cuh) gives every block one warp; a 64 MiB stream has 2,048 blocks, a B200 room for 9,472
// warps, and a warp's decode is one dependent chain.  Here the block's Huffman-coded bits are cut into four
// quarters.  Warp 0 starts where the symbols start; warps 1-3 start at an arbitrary bit and decode speculatively:
// a Huffman/LZ77 token stream decoded from a wrong position falls into step with the true one after a few tokens
// and stays there.  Every warp records the bit position of its first SPEC_K tokens; after the quarters are done,
// warp k decodes on from where it stopped until it lands on a position in warp k+1's record — from that token on
// warp k+1's output is the true continuation, what it produced before is dropped.  Nothing is assumed: a block
// whose quarters do not fall into step within the record, or that is not one dynamic block followed by the
// encoder's marker, is decoded again by warp 0 alone with the one-warp decoder (inf_segment_tokens), which is
// also what reports every error.  The tokens end up contiguous in tokens[j * SUB ...] exactly as k_inf_tokens
// leaves them, so phase B and k_inf_check do not know which kernel ran.
// Replaces the symbol loop of /root/reference/src/inflate.ts:237-291 (see inflate.cuh for the rest).
#pragma once
#include "inflate.cuh"

namespace zles {

constexpr int SPEC_WARPS = 4;
constexpr int SPEC_THREADS = SPEC_WARPS * 32;
constexpr u32 SPEC_K = 512;             // token positions recorded per quarter
constexpr u32 SPEC_CAP = SUB / 4;       // token room per quarter (a quarter that needs more: one-warp decoder)
constexpr u32 SPEC_MIN_BITS = 4 * 4096; // shorter symbol streams are not worth cutting
constexpr u32 SP_RUN = 0, SP_EOB = 1, SP_FAIL = 2, SP_SYNC = 3, SP_PAST = 4;

struct SpecShared {
  TokWarpSmem T;                 // one set of tables for the four warps
  u32 bnd[SPEC_WARPS][SPEC_K];   // bit position (relative to the block's first byte) of a quarter's first tokens
  u32 nt0[SPEC_WARPS];           // tokens a quarter had when its own range was done (the valid part of bnd)
  u32 nt[SPEC_WARPS], ob[SPEC_WARPS], fin[SPEC_WARPS], flag[SPEC_WARPS];
  u32 next[SPEC_WARPS], skipn[SPEC_WARPS];  // the quarter this one fell in step with (SPEC_WARPS: none, it read the end of block itself) and from which token
  u32 skipb[SPEC_WARPS];
  u32 seg, mode, bfinal, sym_start, span;
};
constexpr int SPEC_SMEM = (int)sizeof(SpecShared);

// What lane `lane` sees at its bit offset: pack = bits the token takes (0: a code the fast tables do not hold) |
// bit 8 end of block; tokv = the token; olen = bytes it stands for.  Same arithmetic as inf_segment_tokens.
__device__ __forceinline__ void spec_decode_at(const SpecReader &sr, const TokWarpSmem *T, u32 lane, u32 &pack, u32 &tokv, u32 &olen) {
  u32 lo, hi;
  sr.at(lane, lo, hi);
  const u32 e = T->lut_ll[lo & ((1u << LL_ROOT) - 1)];
  const u32 l1 = e & 15, eb = (e >> 4) & 15, sh2 = l1 + eb;
  const u32 len = ((e >> 8) & 0xffff) + ((lo >> l1) & ~(0xffffffffu << eb));
  const u32 y = __funnelshift_r(lo, hi, sh2);
  const u32 d = T->lut_d[y & ((1u << D_ROOT) - 1)];
  const u32 l2 = d & 15, db = (d >> 4) & 15;
  const u32 dist = (d >> 8) + ((y >> l2) & ~(0xffffffffu << db));
  const bool is_len = (e & TK_LEN) != 0;
  pack = is_len ? (l2 ? sh2 + l2 + db : 0) : l1;
  if (l1 == 0) pack = 0;
  if (e & TK_EOB) pack |= 0x100;
  tokv = is_len ? (0x80000000u | ((len - 3) << 16) | (dist - 1)) : (len & 0xff);
  olen = is_len ? len : 1u;
}

// One token the canonical way (codes longer than the root tables), at the reader's position; the reader moves past it.
// Returns 0 = a token (tokv, olen), 1 = end of block, 2 = not a valid code.
__device__ __forceinline__ u32 spec_slow_token(SpecReader &sr, const TokWarpSmem *T, u32 &tokv, u32 &olen, u32 &bits) {
  const TokCore *S = &T->w;
  u32 lo, hi;
  sr.at(0, lo, hi);
  u32 e = T->lut_ll[lo & ((1u << LL_ROOT) - 1)];
  if ((e & 15) == 0) {
    u32 sym, l;
    if (!inf_slow(((u64)hi << 32) | lo, &S->tab_ll, S->sorted_ll, sym, l)) return 2;
    e = tk_entry_ll(sym, l);
    if (e & TK_INV) return 2;
  }
  if (e & TK_EOB) { bits = e & 15; sr.advance(bits); return 1; }
  if (e < TK_LEN) { bits = e & 15; sr.advance(bits); tokv = e >> 8; olen = 1; return 0; }
  const u32 eb = (e >> 4) & 15;
  const u32 len = ((e >> 8) & 0xffff) + ((lo >> (e & 15)) & ~(0xffffffffu << eb));
  bits = (e & 15) + eb;
  sr.advance(bits);
  sr.at(0, lo, hi);
  u32 d = T->lut_d[lo & ((1u << D_ROOT) - 1)];
  if ((d & 15) == 0) {
    u32 sym, l;
    if (!inf_slow(((u64)hi << 32) | lo, &S->tab_d, S->sorted_d, sym, l)) return 2;
    d = tk_entry_d(sym, l);
    if (d & TK_INV) return 2;
  }
  const u32 db = (d >> 4) & 15;
  const u32 dist = (d >> 8) + ((lo >> (d & 15)) & ~(0xffffffffu << db));
  sr.advance((d & 15) + db);
  bits += (d & 15) + db;
  tokv = 0x80000000u | ((len - 3) << 16) | (dist - 1);
  olen = len;
  return 0;
}

// index of the first entry >= v in an ascending shared-memory list (warp-uniform)
__device__ __forceinline__ u32 spec_lower_bound(const u32 *list, u32 m, u32 v) {
  u32 lo = 0, hi = m;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (list[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Decodes tokens from bit `start` (relative to bit0) and appends them to tokR[nt...] until a token starts at or after
// `stop` (fin = that position) or — is_last — the end-of-block code has been read (flag = SP_EOB, fin = the bit
// after it).  The first SPEC_K token positions go to bnd.  A speculative quarter (may_restart) that meets an end-of-
// block code or an invalid code cannot be in step yet: it drops what it has and starts again one bit further.
// With a `sync` list (the positions another quarter recorded) the run ends as soon as a token starts on one of them
// (SP_SYNC, sync_idx = which) or once it is past all of them (SP_PAST); `stop` is then taken from the list.
__device__ __forceinline__ void spec_run(const TokWarpSmem *T, const u8 *in, u64 n, u64 bit0, u32 start, u32 stop, u32 hard_end, bool is_last,
                                         bool may_restart, u32 *tokR, u32 *bnd, u32 &nt, u32 &ob, u32 &fin, u32 &flag,
                                         const u32 *sync = nullptr, u32 sync_m = 0, u32 *sync_idx = nullptr) {
  const u32 lane = lane_id();
  SpecReader sr;
  u32 pos = start;
  sr.begin(in, n, bit0 + pos);
  flag = SP_RUN;
  for (;;) {
    if (sync) {
      const u32 idx = spec_lower_bound(sync, sync_m, pos);
      if (idx >= sync_m) { flag = SP_PAST; break; }
      if (sync[idx] == pos) { flag = SP_SYNC; *sync_idx = idx; break; }
      stop = sync[idx];
    }
    if (!is_last && pos >= stop) break;
    if (nt + 32 > SPEC_CAP || pos >= hard_end) { flag = SP_FAIL; break; }
    u32 pack, tokv, olen;
    spec_decode_at(sr, T, lane, pack, tokv, olen);
    const u32 nb = pack & 0xff;
    const bool stops = nb == 0 || (pack & 0x100);
    const u32 step = (!stops && lane + nb < 32) ? 1u << (lane + nb) : 0;
    u32 R = 1;
    for (;;) {
      const u32 add = __reduce_or_sync(ZLES_FULL, ((R >> lane) & 1) ? step : 0u);
      if ((add & ~R) == 0) break;
      R |= add;
    }
    if (!is_last && stop - pos < 32) R &= (1u << (stop - pos)) - 1;   // tokens from `stop` on belong to the next quarter (bit 0 stays: pos < stop)
    const u32 last = 31u - (u32)__clz((int)R);
    const u32 plast = __shfl_sync(ZLES_FULL, pack, (int)last);
    const bool last_stops = (plast & 0xff) == 0 || (plast & 0x100);
    const u32 plain = last_stops ? R & ~(1u << last) : R;
    const bool mine = (plain >> lane) & 1;
    const u32 idx = nt + (u32)__popc(plain & lanemask_lt());
    if (mine) {
      tokR[idx] = tokv;
      if (idx < SPEC_K) bnd[idx] = pos + lane;
    }
    ob += __reduce_add_sync(ZLES_FULL, mine ? olen : 0u);
    nt += (u32)__popc(plain);
    if (!last_stops) {
      const u32 cur = last + (plast & 0xff);
      sr.advance(cur);
      pos += cur;
      continue;
    }
    // the chain ended on an end-of-block code or on a code that has to be decoded on its own
    u32 outcome;  // 0 token, 1 end of block, 2 invalid
    u32 t = 0, ol = 0, bits = 0;
    if (plast & 0x100) {
      outcome = 1;
      bits = plast & 0xff;
      pos += last;
    } else {
      sr.advance(last);
      pos += last;
      outcome = spec_slow_token(sr, T, t, ol, bits);
    }
    if (outcome == 0) {
      if (lane == 0) {
        tokR[nt] = t;
        if (nt < SPEC_K) bnd[nt] = pos;
      }
      nt++;
      ob += ol;
      pos += bits;
      continue;
    }
    if (outcome == 1 && is_last) {
      pos += bits;
      flag = SP_EOB;
      break;
    }
    if (!may_restart) { flag = SP_FAIL; break; }
    pos += 1;  // not in step with the real token stream yet
    nt = 0;
    ob = 0;
    sr.begin(in, n, bit0 + pos);
  }
  fin = pos;
}

// a block decoded by the one-warp decoder is one piece (lane 0 of the warp that decoded it; res/ntok are its own writes)
__device__ __forceinline__ void spec_qinfo_whole(u32 *qinfo, u32 j, const InfRes *res, const u32 *ntok) {
  __syncwarp();
  if (lane_id() == 0) {
    u32 *qi = qinfo + (size_t)j * 2 * SPEC_WARPS;
    qi[0] = ntok[j];
    qi[1] = (u32)umin64(res[j].out_len, 0xffffffffull);
    for (u32 k = 1; k < SPEC_WARPS; k++) { qi[2 * k] = 0; qi[2 * k + 1] = 0; }
  }
}

// 9 CTAs per SM = 56 registers: what the kernel was tuned at (fewer registers spill, more CTAs per SM measured slower)
__global__ void __launch_bounds__(SPEC_THREADS, 9)
k_inf_tokens4(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ seg_pos, u32 nseg, u32 *tokens, u32 *ntok, InfRes *res, u32 *qinfo,
              u32 *counter) {
  ZLES_SMEM_DECL(smem_raw);
  SpecShared *Sh = reinterpret_cast<SpecShared *>(smem_raw);
  TokWarpSmem *T = &Sh->T;
  TokCore *S = &T->w;
  const u32 lane = lane_id(), w = warp_id(), tid = threadIdx.x;
  for (;;) {
    __syncthreads();  // everybody is done with the previous block's shared state
    if (tid == 0) Sh->seg = atomicAdd(counter, 1u);
    __syncthreads();
    const u32 j = Sh->seg;
    if (j >= nseg) break;
    u32 *tok = tokens + (size_t)j * SUB;
    const u64 in_pos = seg_pos[j];
    const u64 seg_end = j + 1 < nseg ? seg_pos[j + 1] : n;
    const u64 bit0 = in_pos << 3;

    // 1. warp 0: the block header and the tables
    if (w == 0) {
      u32 mode = 0, bfinal = 0, sym_start = 0, span = 0;
      TokReader r;
      r.init(in, n, in_pos);
      if (!r.past_end()) {
        bfinal = r.take(1);
        const u32 btype = r.take(2);
        u32 status = 0;
        if (btype == 2 && tk_read_dynamic_header(r, T, status)) {
          tk_build_tables(T);
          const u64 bp = r.bitpos();
          if (seg_end > in_pos && (seg_end << 3) > bp && (seg_end - in_pos) < (1u << 19)) {
            sym_start = (u32)(bp - bit0);
            span = (u32)((seg_end << 3) - bp);
            if (span >= SPEC_MIN_BITS) mode = 1;
          }
        }
      }
      if (lane == 0) { Sh->mode = mode; Sh->bfinal = bfinal; Sh->sym_start = sym_start; Sh->span = span; }
    }
    __syncthreads();
    if (Sh->mode == 0) {  // not one dynamic block of a useful size: the one-warp decoder
      if (w == 0) {
        inf_segment_tokens(T, in, n, in_pos, tok, res + j, ntok + j);
        spec_qinfo_whole(qinfo, j, res, ntok);
      }
      continue;
    }

    // 2. every warp its quarter
    const u32 sym_start = Sh->sym_start, span = Sh->span;
    const u32 hard_end = sym_start + span;
    u32 nt = 0, ob = 0, fin = 0, flag = SP_RUN;
    {
      const u32 q0 = sym_start + w * (span / SPEC_WARPS);
      const u32 q1 = w + 1 < SPEC_WARPS ? sym_start + (w + 1) * (span / SPEC_WARPS) : hard_end;
      spec_run(T, in, n, bit0, q0, q1, hard_end, w + 1 == SPEC_WARPS, w > 0, tok + w * SPEC_CAP, Sh->bnd[w], nt, ob, fin, flag);
      if (lane == 0) { Sh->nt0[w] = nt; Sh->flag[w] = flag; }
    }
    __syncthreads();

    // 3. warp k decodes on until it is in step with a later quarter.  One that it runs past without meeting any of
    //    its recorded positions never fell in step in time: its output is dropped and its range decoded here.
    u32 next = SPEC_WARPS, skipn = 0;
    if (flag != SP_FAIL && flag != SP_EOB) {
      for (u32 nk = w + 1;; nk++) {
        if (nk >= SPEC_WARPS) {  // nobody left to meet: read on to the end of the block
          spec_run(T, in, n, bit0, fin, hard_end, hard_end, true, false, tok + w * SPEC_CAP, Sh->bnd[w], nt, ob, fin, flag);
          break;
        }
        if (Sh->flag[nk] == SP_FAIL) continue;
        u32 idx = 0;
        INF_CNT(4, 1);
        spec_run(T, in, n, bit0, fin, 0, hard_end, false, false, tok + w * SPEC_CAP, Sh->bnd[w], nt, ob, fin, flag, Sh->bnd[nk],
                 umin(Sh->nt0[nk], SPEC_K), &idx);
        if (flag == SP_SYNC) { next = nk; skipn = idx; break; }
        if (flag != SP_PAST) break;  // SP_FAIL
      }
    }
    __syncthreads();  // every warp has read the others' flag and nt0 of step 2
    if (lane == 0) { Sh->nt[w] = nt; Sh->ob[w] = ob; Sh->fin[w] = fin; Sh->flag[w] = flag; Sh->next[w] = next; Sh->skipn[w] = skipn; }
    __syncthreads();

    // 4. the chain of quarters that make up the block, what each keeps, and what the dropped tokens stood for
    bool good = true;
    u32 skip[SPEC_WARPS], off[SPEC_WARPS], cnt[SPEC_WARPS];
    u32 last_piece = 0;
#pragma unroll
    for (int k = 0; k < SPEC_WARPS; k++) { skip[k] = 0xffffffffu; off[k] = 0; cnt[k] = 0; }
    skip[0] = 0;
    {
      u32 k = 0;
      for (int it = 0; it < SPEC_WARPS; it++) {
        last_piece = k;
        const u32 f = Sh->flag[k], nx = Sh->next[k];
        if (f == SP_FAIL || f == SP_RUN || f == SP_PAST) { good = false; break; }
        if (f == SP_EOB) break;
        if (nx >= SPEC_WARPS || nx <= k) { good = false; break; }  // SP_SYNC
        skip[nx] = Sh->skipn[k];
        k = nx;
      }
      if (Sh->flag[last_piece] != SP_EOB) good = false;
    }
    {
      const u32 sk = skip[w] == 0xffffffffu ? 0u : umin(skip[w], Sh->nt[w]);
      const u32 *tr = tok + w * SPEC_CAP;
      u32 sb = 0;
      for (u32 i = lane; i < sk; i += 32) {
        const u32 t = tr[i];
        sb += (t >> 31) ? ((t >> 16) & 0x1ff) + 3 : 1;
      }
      sb = __reduce_add_sync(ZLES_FULL, sb);
      if (lane == 0) Sh->skipb[w] = sb;
    }
    __syncthreads();
    u32 total_tok = 0, total_out = 0;
#pragma unroll
    for (int k = 0; k < SPEC_WARPS; k++) {
      if (skip[k] == 0xffffffffu) continue;  // not part of the block
      if (skip[k] > Sh->nt[k]) { good = false; continue; }
      off[k] = total_tok;
      cnt[k] = Sh->nt[k] - skip[k];
      total_tok += cnt[k];
      total_out += Sh->ob[k] - Sh->skipb[k];
    }
    if (total_out > SUB || total_tok > SUB) good = false;

    // 5. the marker (or the end of the stream) after the end-of-block code — warp 0 decides, everybody follows
    if (w == 0) {
      u32 status = 0;
      u64 end_pos = 0;
      if (good) {
        const u64 pos = bit0 + Sh->fin[last_piece];
        TokReader r;
        r.init(in, n, pos >> 3);
        r.skip((u32)(pos & 7));
        if (r.past_end()) good = false;
        else if (Sh->bfinal) { status = SEG_FINAL; end_pos = (r.bitpos() + 7) >> 3; }
        else {
          r.refill();
          const u32 bf2 = r.take(1), bt2 = r.take(2);
          if (r.past_end() || bt2 != 0) good = false;
          else {
            r.skip((u32)((0 - r.bitpos()) & 7));
            r.refill();
            const u32 LEN = r.take(16);
            r.refill();
            const u32 NLEN = r.take(16);
            if (LEN != 0 || NLEN != 0xffff || r.past_end()) good = false;   // anything but the empty stored block: one-warp decoder
            else { end_pos = r.bitpos() >> 3; status = bf2 ? SEG_FINAL : SEG_SYNC; }
          }
        }
      }
      if (lane == 0) {
        Sh->mode = good ? 2 : 0;
        if (good) {
          res[j].end_pos = end_pos;
          res[j].out_len = total_out;
          res[j].status = status;
          res[j].flags = 0;
          ntok[j] = total_tok;
          // the pieces in stream order, for a phase B that gives every piece its own warp (k_piece_sym)
          u32 *qi = qinfo + (size_t)j * 2 * SPEC_WARPS;
          u32 slot = 0;
#pragma unroll
          for (int k = 0; k < SPEC_WARPS; k++) {
            if (skip[k] == 0xffffffffu) continue;
            qi[2 * slot] = cnt[k];
            qi[2 * slot + 1] = Sh->ob[k] - Sh->skipb[k];
            slot++;
          }
          for (; slot < SPEC_WARPS; slot++) { qi[2 * slot] = 0; qi[2 * slot + 1] = 0; }
        }
      }
    }
    __syncthreads();
    if (Sh->mode == 0) {
      if (w == 0) {
        INF_CNT(7, 1);
        inf_segment_tokens(T, in, n, in_pos, tok, res + j, ntok + j);
        spec_qinfo_whole(qinfo, j, res, ntok);
      }
      continue;
    }
    if (w == 0) INF_CNT(6, 1);

    // 6. make the token list contiguous: quarter k's kept tokens move down to off[k]
#pragma unroll
    for (int k = 1; k < SPEC_WARPS; k++) {
      if (cnt[k] == 0) continue;
      const u32 src0 = (u32)k * SPEC_CAP + skip[k];
      if (src0 == off[k]) continue;
      for (u32 i0 = 0; i0 < cnt[k]; i0 += SPEC_THREADS) {
        const u32 i = i0 + tid;
        u32 v = 0;
        if (i < cnt[k]) v = tok[src0 + i];
        __syncthreads();  // a store below may land on what another thread of this round has just read
        if (i < cnt[k]) tok[off[k] + i] = v;
      }
      __syncthreads();
    }
  }
}

}  // namespace zles
