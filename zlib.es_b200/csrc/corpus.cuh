// corpus.cuh — the synthetic corpora BASELINE.json's configs name,
// generated directly in HBM so that benchmark inputs never cross PCIe, with a
// bit-identical host version for tests and the CPU-baseline sample.
//
// Counter based: u(seed, i) = mix64(seed + GOLDEN * (i + 1)).  Data is produced in
// independent 64 KiB pages keyed by the absolute page index, so any shard of a
// corpus can be generated on any GPU.
//   kind 0  T  English-like text: order-2 character Markov chain trained on
//              corpus_text.h, every page starts in state "e "
//   kind 1  B  structured binary: 32-byte little-endian records
//   kind 2  R  random bytes
//   kind 3  mixed: per 128 KiB segment s, u(MIX, s) % 20: 0-1 R, 2-10 T, 11-19 B
#pragma once
#include "zles_dev.h"

namespace zles {

constexpr u32 CORPUS_PAGE = 65536;
constexpr u32 CORPUS_NSYM = 59;
constexpr u64 CORPUS_SEED_T = 0xB2000001ull, CORPUS_SEED_B = 0xB2000002ull, CORPUS_SEED_R = 0xB2000003ull,
              CORPUS_SEED_MIX = 0xB20000FFull;

struct CorpusTable {
  u16 cdf[CORPUS_NSYM * CORPUS_NSYM][64];  // cdf[state][k] = P(next <= k) * 65536, saturated; [59..63] = 65535
  u8 alphabet[64];
};

__host__ __device__ __forceinline__ u64 corpus_u(u64 seed, u64 i) {
  u64 z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ int corpus_kind_of_page(int kind, u64 page) {
  if (kind != 3) return kind;
  u64 c = corpus_u(CORPUS_SEED_MIX, page >> 1) % 20;
  return c < 2 ? 2 : (c < 11 ? 0 : 1);
}

// Writes the bytes of `page` that fall inside [lo, hi) (absolute offsets) to dst - lo.
__host__ __device__ inline void corpus_page(const CorpusTable *T, int kind, u64 page, u64 lo, u64 hi, u8 *dst) {
  const u64 base = page * CORPUS_PAGE;
  const int k = corpus_kind_of_page(kind, page);
#define ZLES_PUT(j, v)                                  \
  do {                                                  \
    u64 a_ = base + (j);                                \
    if (a_ >= lo && a_ < hi) dst[a_ - lo] = (u8)(v);    \
  } while (0)
  if (k == 2) {
    for (u32 j = 0; j < CORPUS_PAGE; j += 8) {
      u64 r = corpus_u(CORPUS_SEED_R ^ page, j >> 3);
      for (u32 b = 0; b < 8; b++) ZLES_PUT(j + b, r >> (8 * b));
    }
  } else if (k == 0) {
    u32 s1 = 4 /* 'e' */, s2 = 52 /* ' ' */;
    for (u32 j = 0; j < CORPUS_PAGE; j++) {
      u32 r = (u32)(corpus_u(CORPUS_SEED_T ^ page, j) >> 24) & 0xFFFF;
      const u16 *cdf = T->cdf[s1 * CORPUS_NSYM + s2];
      u32 lo_i = 0, hi_i = CORPUS_NSYM - 1;  // first index with cdf > r
      while (lo_i < hi_i) {
        u32 mid = (lo_i + hi_i) >> 1;
        if (cdf[mid] > r) hi_i = mid; else lo_i = mid + 1;
      }
      ZLES_PUT(j, T->alphabet[lo_i]);
      s1 = s2;
      s2 = lo_i;
    }
  } else {
    const float fbase[6] = {0.f, 0.5f, 1.f, 1.5f, 2.25f, 100.f};
    u32 ts = (u32)(page * 51200u);
    for (u32 rec = 0; rec < CORPUS_PAGE / 32; rec++) {
      u64 r0 = corpus_u(CORPUS_SEED_B ^ page, 2 * (u64)rec), r1 = corpus_u(CORPUS_SEED_B ^ page, 2 * (u64)rec + 1);
      u32 id = (u32)(page * 2048u + rec);
      ts += (u32)(r0 % 50);
      float f = fbase[(r0 >> 8) % 6] + (float)((r0 >> 16) % 4) * 0.25f;
      u32 fbits;
      memcpy(&fbits, &f, 4);
      u32 h = (u32)((r0 >> 24) % 300);
      u32 tag = (u32)(r1 % 64);
      u32 j = rec * 32;
      for (u32 b = 0; b < 4; b++) ZLES_PUT(j + b, id >> (8 * b));
      for (u32 b = 0; b < 4; b++) ZLES_PUT(j + 4 + b, ts >> (8 * b));
      for (u32 b = 0; b < 4; b++) ZLES_PUT(j + 8 + b, fbits >> (8 * b));
      ZLES_PUT(j + 12, h);
      ZLES_PUT(j + 13, h >> 8);
      // 8-byte ASCII tag from a 64-entry table: "TAG" + 2 letters + 3 digits
      ZLES_PUT(j + 14, 'T'); ZLES_PUT(j + 15, 'A'); ZLES_PUT(j + 16, 'G');
      ZLES_PUT(j + 17, 'A' + (tag >> 3)); ZLES_PUT(j + 18, 'a' + (tag & 7));
      ZLES_PUT(j + 19, '0' + (tag % 10)); ZLES_PUT(j + 20, '0' + ((tag * 7) % 10)); ZLES_PUT(j + 21, '0' + ((tag * 3) % 10));
      for (u32 b = 22; b < 32; b++) ZLES_PUT(j + b, 0);
    }
  }
#undef ZLES_PUT
}

__global__ void __launch_bounds__(128) k_corpus(const CorpusTable *T, int kind, u64 offset, u8 *out, u64 n) {
  const u64 first_page = offset / CORPUS_PAGE;
  const u64 last_page = (offset + n + CORPUS_PAGE - 1) / CORPUS_PAGE;
  const u64 page = first_page + (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (page < last_page) corpus_page(T, kind, page, offset, offset + n, out);
}

}  // namespace zles
