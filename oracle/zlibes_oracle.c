/*
 * zlibes_oracle.c — CPU restatement of zprodev/zlib.es v0.6.0.
 * TEST INFRASTRUCTURE ONLY (see zlibes_oracle.h for the rules and the pinning
 * status: inflate + Adler-32 pinned by the reference's vectors; deflate's exact
 * bits are PARITY UNPINNED by the reference itself).
 *
 * Every function cites the reference lines it follows (paths are relative to
 * /root/reference/src).  JavaScript semantics that matter are restated
 * explicitly: ascending integer-key order of Object.keys, stable
 * Array.prototype.sort, `undefined` for out-of-range typed-array reads.
 */
#include "zlibes_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- src/const.ts:7-35 -------------------------------------------------- */
#define BLOCK_MAX_BUFFER_LEN 131072u
static const int LENGTH_EXTRA_BIT_LEN[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                             2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int LENGTH_EXTRA_BIT_BASE[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                              31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const int DISTANCE_EXTRA_BIT_BASE[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                                33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                                1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const int DISTANCE_EXTRA_BIT_LEN[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                               6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const int CODELEN_VALUES[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

const char *zo_strerror(int code) {
  switch (code) {
    case ZO_OK: return "";
    case ZO_E_NOT_DEFLATE: return "Not compressed by deflate";
    case ZO_E_BTYPE3: return "Not supported BTYPE : 3";
    case ZO_E_INSUFFICIENT: return "Data length is insufficient";
    case ZO_E_CORRUPTED: return "Data is corrupted";
    case ZO_E_LACK: return "Lack of data length";
    case ZO_E_NOMEM: return "out of memory";
    case ZO_E_RUNAWAY: return "stream never ends";
  }
  return "unknown";
}
void zo_free(void *p) { free(p); }

/* ---- src/adler32.ts:1-10 ------------------------------------------------- */
uint32_t zo_adler32(const uint8_t *in, size_t n) {
  uint32_t s1 = 1, s2 = 0;
  for (size_t i = 0; i < n; i++) {
    s1 = (s1 + in[i]) % 65521u; /* :5 */
    s2 = (s1 + s2) % 65521u;    /* :6 */
  }
  return (s2 << 16) + s1; /* :9, compared as >>> 0 */
}

/* ---- src/utils/BitWriteStream.ts:1-47 ------------------------------------ */
typedef struct {
  uint8_t *buffer;
  size_t length;
  size_t bufferIndex;
  uint32_t nowBits;
  int nowBitsIndex;
  int isEnd;
  int err;
} BitWriteStream;

static void bws_write(BitWriteStream *s, int bit) { /* :14-28 */
  if (s->isEnd) { s->err = ZO_E_LACK; return; }
  s->nowBits += (uint32_t)bit << s->nowBitsIndex;
  s->nowBitsIndex++;
  if (s->nowBitsIndex >= 8) {
    s->buffer[s->bufferIndex] = (uint8_t)s->nowBits;
    s->bufferIndex++;
    s->nowBits = 0;
    s->nowBitsIndex = 0;
    if (s->length <= s->bufferIndex) s->isEnd = 1;
  }
}
static void bws_writeRange(BitWriteStream *s, uint32_t value, int length) { /* :29-37, LSB first */
  uint32_t mask = 1;
  for (int i = 0; i < length; i++) {
    bws_write(s, (value & mask) ? 1 : 0);
    mask <<= 1;
  }
}
static void bws_writeRangeCoded(BitWriteStream *s, uint32_t value, int length) { /* :38-46, MSB first */
  uint32_t mask = 1u << (length - 1);
  for (int i = 0; i < length; i++) {
    bws_write(s, (value & mask) ? 1 : 0);
    mask >>= 1;
  }
}

/* ---- src/lz77.ts --------------------------------------------------------- */
#define REPEAT_LEN_MIN 3
#define FAST_INDEX_CHECK_MAX 128
#define FAST_INDEX_CHECK_MIN 16
#define FAST_REPEAT_LENGTH 8

/* one LZ77 code: literal [v] or [lenSym, distSym, len, dist]; `undef` marks a
 * literal that read `undefined` (block length 0 or 1, quirk Q1). */
typedef struct {
  int32_t v0, v1, v2, v3;
  uint8_t is_match, undef;
} Lz77Code;

/* generateLZ77IndexMap, src/lz77.ts:11-22: exact 24-bit key -> ascending
 * position list, positions start .. start+len-3 of THIS block only.  Restated
 * as a stable LSD radix sort of the positions by key; list of a key = one run. */
typedef struct {
  uint32_t *sorted;  /* positions grouped by key, ascending inside a group */
  uint32_t *grp;     /* per position (relative): start of its group in sorted[] */
  uint32_t *glen;    /* per group start: group length */
  uint32_t *cur_s;   /* startIndexMap, src/lz77.ts:35 (per key == per group) */
  uint32_t *cur_e;   /* endIndexMap,   src/lz77.ts:36 */
  size_t cnt;
} IndexMap;

static int indexmap_build(IndexMap *m, const uint8_t *in, size_t start, size_t len) {
  memset(m, 0, sizeof(*m));
  if (len < REPEAT_LEN_MIN) return 0;
  size_t cnt = len - REPEAT_LEN_MIN + 1; /* i = start .. start+len-3 inclusive */
  m->cnt = cnt;
  uint32_t *a = (uint32_t *)malloc(cnt * sizeof(uint32_t));
  uint32_t *b = (uint32_t *)malloc(cnt * sizeof(uint32_t));
  uint32_t *key = (uint32_t *)malloc(cnt * sizeof(uint32_t));
  m->grp = (uint32_t *)malloc(cnt * sizeof(uint32_t));
  m->glen = (uint32_t *)calloc(cnt, sizeof(uint32_t));
  m->cur_s = (uint32_t *)calloc(cnt, sizeof(uint32_t));
  m->cur_e = (uint32_t *)calloc(cnt, sizeof(uint32_t));
  if (!a || !b || !key || !m->grp || !m->glen || !m->cur_s || !m->cur_e) return ZO_E_NOMEM;
  for (size_t i = 0; i < cnt; i++) {
    const uint8_t *p = in + start + i;
    key[i] = (uint32_t)p[0] << 16 | (uint32_t)p[1] << 8 | p[2]; /* :15 */
    a[i] = (uint32_t)i;
  }
  for (int pass = 0; pass < 3; pass++) {
    size_t hist[257] = {0};
    int sh = pass * 8;
    for (size_t i = 0; i < cnt; i++) hist[((key[a[i]] >> sh) & 255) + 1]++;
    for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
    for (size_t i = 0; i < cnt; i++) b[hist[(key[a[i]] >> sh) & 255]++] = a[i];
    uint32_t *t = a; a = b; b = t;
  }
  /* a[] = relative positions sorted by (key, position) */
  size_t g = 0;
  for (size_t i = 0; i < cnt; i++) {
    if (i > 0 && key[a[i]] != key[a[i - 1]]) g = i;
    m->grp[a[i]] = (uint32_t)g;
    m->glen[g]++;
  }
  for (size_t i = 0; i < cnt; i++) a[i] += (uint32_t)start; /* absolute indices like the reference */
  m->sorted = a;
  free(b);
  free(key);
  return 0;
}
static void indexmap_free(IndexMap *m) {
  free(m->sorted); free(m->grp); free(m->glen); free(m->cur_s); free(m->cur_e);
}

/* `input[k]` with JS out-of-range semantics: 256 stands for `undefined`
 * (!== every byte, === itself). */
static inline int js_at(const uint8_t *in, size_t n, size_t k) { return k < n ? in[k] : 256; }

/* generateLZ77Codes, src/lz77.ts:24-119.  codes must hold len+2 entries. */
static int lz77_codes(const uint8_t *in, size_t n, size_t start, size_t len, Lz77Code *codes, size_t *n_codes) {
  size_t nowIndex = start;
  /* endIndex = start + len - 3 may be "negative" (len < 3): keep it signed. */
  long long endIndex = (long long)start + (long long)len - REPEAT_LEN_MIN; /* :26 */
  size_t nc = 0;
  IndexMap m;
  int rc = indexmap_build(&m, in, start, len); /* :37 */
  if (rc) { indexmap_free(&m); return rc; }

  while ((long long)nowIndex <= endIndex) { /* :39 */
    size_t rel = nowIndex - start;
    uint32_t g = m.grp[rel];
    const uint32_t *indexes = m.sorted + g;
    uint32_t ilen = m.glen[g];
    if (ilen <= 1) { /* :44-48 */
      codes[nc].is_match = 0; codes[nc].undef = 0; codes[nc].v0 = in[nowIndex]; nc++;
      nowIndex++;
      continue;
    }
    size_t slideIndexBase = (nowIndex > 0x8000) ? nowIndex - 0x8000 : 0; /* :49 */
    int repeatLengthMax = 0;
    size_t repeatLengthMaxIndex = 0;

    uint32_t skip = m.cur_s[g]; /* :53-57 */
    while (indexes[skip] < slideIndexBase) skip++;
    m.cur_s[g] = skip;
    skip = m.cur_e[g]; /* :58-62 */
    while (indexes[skip] < nowIndex) skip++;
    m.cur_e[g] = skip;

    int checkCount = 0;
    for (long long i = (long long)m.cur_e[g] - 1, iMin = m.cur_s[g]; iMin <= i; i--) { /* :65 */
      if (checkCount >= FAST_INDEX_CHECK_MAX ||
          (repeatLengthMax >= FAST_REPEAT_LENGTH && checkCount >= FAST_INDEX_CHECK_MIN)) { /* :66-69 */
        break;
      }
      checkCount++;
      size_t index = indexes[i];
      int rejected = 0;
      for (int j = repeatLengthMax - 1; 0 < j; j--) { /* :72-76 */
        if (js_at(in, n, index + j) != js_at(in, n, nowIndex + j)) { rejected = 1; break; }
      }
      if (rejected) continue;
      int repeatLength = 258; /* :78 */
      for (int j = repeatLengthMax; j <= 258; j++) { /* :80-85 */
        if (js_at(in, n, index + j) != js_at(in, n, nowIndex + j)) { repeatLength = j; break; }
      }
      if (repeatLengthMax < repeatLength) { /* :86-92 */
        repeatLengthMax = repeatLength;
        repeatLengthMaxIndex = index;
        if (258 <= repeatLength) break;
      }
    }

    if (repeatLengthMax >= 3 && (long long)nowIndex + repeatLengthMax <= endIndex) { /* :95 */
      int distance = (int)(nowIndex - repeatLengthMaxIndex);
      int lc = 0, dc = 0;
      for (int i = 0; i < 29; i++) { /* :97-102 */
        if (LENGTH_EXTRA_BIT_BASE[i] > repeatLengthMax) break;
        lc = i;
      }
      for (int i = 0; i < 30; i++) { /* :103-108 */
        if (DISTANCE_EXTRA_BIT_BASE[i] > distance) break;
        dc = i;
      }
      codes[nc].is_match = 1; codes[nc].undef = 0;
      codes[nc].v0 = lc; codes[nc].v1 = dc; codes[nc].v2 = repeatLengthMax; codes[nc].v3 = distance;
      nc++;
      nowIndex += (size_t)repeatLengthMax; /* :110 */
    } else {
      codes[nc].is_match = 0; codes[nc].undef = 0; codes[nc].v0 = in[nowIndex]; nc++; /* :112 */
      nowIndex++;
    }
  }
  /* :116-117 — the last two bytes are always literals; may read `undefined`. */
  for (int k = 0; k < 2; k++) {
    int v = js_at(in, n, nowIndex + k);
    codes[nc].is_match = 0; codes[nc].undef = (v == 256); codes[nc].v0 = v; nc++;
  }
  indexmap_free(&m);
  *n_codes = nc;
  return 0;
}

int zo_lz77_count(const uint8_t *in, size_t n, size_t start, size_t len, uint32_t *n_tokens, uint32_t *n_matches) {
  Lz77Code *codes = (Lz77Code *)malloc((len + 2) * sizeof(Lz77Code));
  if (!codes) return ZO_E_NOMEM;
  size_t nc = 0;
  int rc = lz77_codes(in, n, start, len, codes, &nc);
  uint32_t nm = 0;
  for (size_t i = 0; i < nc; i++) nm += codes[i].is_match;
  *n_tokens = (uint32_t)nc;
  *n_matches = nm;
  free(codes);
  return rc;
}

/* ---- src/huffman.ts:55-153 generateDeflateHuffmanTable -------------------- */
#define HT_MAXSYM 320 /* symbols are < 288 on every call site */
typedef struct {
  int has[HT_MAXSYM];
  uint32_t code[HT_MAXSYM];
  int bitlen[HT_MAXSYM];
} DeflateHuffmanTable;

typedef struct {
  uint64_t count;
  uint8_t mult[HT_MAXSYM]; /* multiset of `simbles`, indexed by rank of the symbol among present symbols */
} Package;

/* stable merge sort of package indices by count (Array.prototype.sort with the
 * comparator at src/huffman.ts:95-99 is stable on ES2019+/Node >= 11). */
static void stable_sort_idx(int *idx, int *tmp, int n, const Package *p) {
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int i = lo, j = mid, k = lo;
      while (i < mid && j < hi) tmp[k++] = (p[idx[j]].count < p[idx[i]].count) ? idx[j++] : idx[i++];
      while (i < mid) tmp[k++] = idx[i++];
      while (j < hi) tmp[k++] = idx[j++];
    }
    memcpy(idx, tmp, (size_t)n * sizeof(int));
  }
}

/* values: the symbol list; the histogram is what matters (:59-66). */
static int deflate_huffman_table(const uint32_t *hist_in, int nsym_space, int maxLength, DeflateHuffmanTable *t) {
  memset(t, 0, sizeof(*t));
  int keys[HT_MAXSYM]; /* Object.keys(valuesCount): ascending integer order (:67) */
  int m = 0;
  for (int s = 0; s < nsym_space; s++)
    if (hist_in[s]) keys[m++] = s;
  int codelen_of[HT_MAXSYM];
  memset(codelen_of, 0, sizeof(codelen_of));
  if (m == 0) return 0; /* empty table (:77-104 run on empty lists) */
  if (m == 1) {
    codelen_of[0] = 1; /* :71-75 then :107-115 */
  } else {
    Package *packages = (Package *)malloc(sizeof(Package) * (size_t)(2 * m + 2));
    Package *tmpPackages = (Package *)malloc(sizeof(Package) * (size_t)(2 * m + 2));
    int *idx = (int *)malloc(sizeof(int) * (size_t)(2 * m + 2) * 2);
    Package *sorted = (Package *)malloc(sizeof(Package) * (size_t)(2 * m + 2));
    if (!packages || !tmpPackages || !idx || !sorted) return ZO_E_NOMEM;
    int ntmp = 0, np = 0;
    for (int round = 0; round < maxLength; round++) { /* :77 */
      np = 0;
      for (int k = 0; k < m; k++) { /* :79-85 fresh leaves, ascending symbol order */
        packages[np].count = hist_in[keys[k]];
        memset(packages[np].mult, 0, (size_t)m);
        packages[np].mult[k] = 1;
        np++;
      }
      for (int ti = 0; ti + 2 <= ntmp; ti += 2) { /* :86-94 pair up the previous round */
        packages[np].count = tmpPackages[ti].count + tmpPackages[ti + 1].count;
        for (int k = 0; k < m; k++) packages[np].mult[k] = (uint8_t)(tmpPackages[ti].mult[k] + tmpPackages[ti + 1].mult[k]);
        np++;
      }
      for (int i = 0; i < np; i++) idx[i] = i; /* :95-99 stable sort by count */
      stable_sort_idx(idx, idx + np, np, packages);
      for (int i = 0; i < np; i++) {
        sorted[i].count = packages[idx[i]].count;
        memcpy(sorted[i].mult, packages[idx[i]].mult, (size_t)m);
      }
      if (np % 2 != 0) np--; /* :100-102 pop the largest */
      Package *sw = tmpPackages; tmpPackages = sorted; sorted = sw;
      ntmp = np;
    }
    for (int i = 0; i < ntmp; i++) /* :106-115 code length = multiplicity over the final list */
      for (int k = 0; k < m; k++) codelen_of[k] += tmpPackages[i].mult[k];
    free(packages); free(tmpPackages); free(idx); free(sorted);
  }
  /* :117-151 canonical codes: by length, then ascending symbol; code <<= 1 per length */
  int lmin = 1 << 30, lmax = 0;
  for (int k = 0; k < m; k++) {
    if (codelen_of[k] < lmin) lmin = codelen_of[k];
    if (codelen_of[k] > lmax) lmax = codelen_of[k];
  }
  uint32_t code = 0;
  for (int l = lmin; l <= lmax; l++) {
    for (int k = 0; k < m; k++) {
      if (codelen_of[k] == l) {
        t->has[keys[k]] = 1; t->code[keys[k]] = code; t->bitlen[keys[k]] = l;
        code++;
      }
    }
    code <<= 1;
  }
  return 0;
}

/* ---- src/deflate.ts:56-227 deflateDynamicBlock ---------------------------- */
static int deflate_dynamic_block(BitWriteStream *stream, const uint8_t *in, size_t n, size_t startIndex, size_t targetLength) {
  Lz77Code *lz = (Lz77Code *)malloc((targetLength + 2) * sizeof(Lz77Code));
  if (!lz) return ZO_E_NOMEM;
  size_t nlz = 0;
  int rc = lz77_codes(in, n, startIndex, targetLength, lz, &nlz); /* :57 */
  if (rc) { free(lz); return rc; }

  uint32_t clHist[HT_MAXSYM] = {0}, distHist[HT_MAXSYM] = {0};
  int sawUndefined = 0;
  clHist[256] = 1; /* :58 */
  int clCodeValueMax = 256, distanceCodeValueMax = 0;
  for (size_t i = 0; i < nlz; i++) { /* :62-77 */
    if (lz[i].is_match) {
      int cl = lz[i].v0 + 257;
      distHist[lz[i].v1]++;
      if (distanceCodeValueMax < lz[i].v1) distanceCodeValueMax = lz[i].v1;
      clHist[cl]++;
      if (clCodeValueMax < cl) clCodeValueMax = cl;
    } else if (lz[i].undef) {
      /* Q1: `undefined` becomes histogram key 'undefined' -> NaN; the code is
       * built, but Map.get(undefined) later misses -> 'Data is corrupted'
       * (src/deflate.ts:214-217).  Nothing of the output survives the throw. */
      sawUndefined = 1;
    } else {
      clHist[lz[i].v0]++;
    }
  }
  if (sawUndefined) { free(lz); return ZO_E_CORRUPTED; }

  DeflateHuffmanTable *dataT = (DeflateHuffmanTable *)malloc(sizeof(DeflateHuffmanTable) * 3);
  if (!dataT) { free(lz); return ZO_E_NOMEM; }
  DeflateHuffmanTable *distT = dataT + 1, *clT = dataT + 2;
  rc = deflate_huffman_table(clHist, 288, 15, dataT); /* :78 */
  if (!rc) rc = deflate_huffman_table(distHist, 30, 15, distT); /* :79 */
  if (rc) { free(lz); free(dataT); return rc; }

  int codelens[320];
  int ncl = 0;
  for (int i = 0; i <= clCodeValueMax; i++) codelens[ncl++] = dataT->has[i] ? dataT->bitlen[i] : 0; /* :82-88 */
  int HLIT = ncl;
  for (int i = 0; i <= distanceCodeValueMax; i++) codelens[ncl++] = distT->has[i] ? distT->bitlen[i] : 0; /* :90-96 */
  int HDIST = ncl - HLIT;

  int runLengthCodes[320], runLengthRepeatCount[320];
  int nrl = 0, nrc = 0;
  uint32_t rlHist[HT_MAXSYM] = {0};
  for (int i = 0; i < ncl; i++) { /* :103-139 */
    int codelen = codelens[i];
    int repeatLength = 1;
    while (i + 1 < ncl && codelen == codelens[i + 1]) {
      repeatLength++;
      i++;
      if (codelen == 0) {
        if (138 <= repeatLength) break;
      } else {
        if (6 <= repeatLength) break;
      }
    }
    if (4 <= repeatLength) {
      if (codelen == 0) {
        runLengthCodes[nrl++] = (11 <= repeatLength) ? 18 : 17;
      } else {
        runLengthCodes[nrl++] = codelen;
        runLengthRepeatCount[nrc++] = 1;
        repeatLength--;
        runLengthCodes[nrl++] = 16;
      }
      runLengthRepeatCount[nrc++] = repeatLength;
    } else {
      for (int j = 0; j < repeatLength; j++) {
        runLengthCodes[nrl++] = codelen;
        runLengthRepeatCount[nrc++] = 1;
      }
    }
  }
  for (int i = 0; i < nrl; i++) rlHist[runLengthCodes[i]]++;
  rc = deflate_huffman_table(rlHist, 19, 7, clT); /* :141 */
  if (rc) { free(lz); free(dataT); return rc; }

  int HCLEN = 0;
  for (int i = 0; i < 19; i++) /* :143-148 */
    if (clT->has[CODELEN_VALUES[i]]) HCLEN = i + 1;

  bws_writeRange(stream, (uint32_t)(HLIT - 257), 5); /* :151 */
  bws_writeRange(stream, (uint32_t)(HDIST - 1), 5);  /* :153 */
  bws_writeRange(stream, (uint32_t)(HCLEN - 4), 4);  /* :155 */
  for (int i = 0; i < HCLEN; i++) {                  /* :158-165 */
    int v = CODELEN_VALUES[i];
    bws_writeRange(stream, clT->has[v] ? (uint32_t)clT->bitlen[v] : 0, 3);
  }
  for (int i = 0; i < nrl; i++) { /* :167-181 */
    int v = runLengthCodes[i];
    if (!clT->has[v]) { rc = ZO_E_CORRUPTED; goto done; }
    bws_writeRangeCoded(stream, clT->code[v], clT->bitlen[v]);
    if (v == 18) bws_writeRange(stream, (uint32_t)(runLengthRepeatCount[i] - 11), 7);
    else if (v == 17) bws_writeRange(stream, (uint32_t)(runLengthRepeatCount[i] - 3), 3);
    else if (v == 16) bws_writeRange(stream, (uint32_t)(runLengthRepeatCount[i] - 3), 2);
  }
  for (size_t i = 0; i < nlz; i++) { /* :183-220 */
    if (lz[i].is_match) {
      int clv = lz[i].v0, dv = lz[i].v1;
      if (!dataT->has[clv + 257]) { rc = ZO_E_CORRUPTED; goto done; }
      bws_writeRangeCoded(stream, dataT->code[clv + 257], dataT->bitlen[clv + 257]);
      if (0 < LENGTH_EXTRA_BIT_LEN[clv])
        bws_writeRange(stream, (uint32_t)(lz[i].v2 - LENGTH_EXTRA_BIT_BASE[clv]), LENGTH_EXTRA_BIT_LEN[clv]);
      if (!distT->has[dv]) { rc = ZO_E_CORRUPTED; goto done; }
      bws_writeRangeCoded(stream, distT->code[dv], distT->bitlen[dv]);
      if (0 < DISTANCE_EXTRA_BIT_LEN[dv])
        bws_writeRange(stream, (uint32_t)(lz[i].v3 - DISTANCE_EXTRA_BIT_BASE[dv]), DISTANCE_EXTRA_BIT_LEN[dv]);
    } else {
      int v = lz[i].v0;
      if (!dataT->has[v]) { rc = ZO_E_CORRUPTED; goto done; }
      bws_writeRangeCoded(stream, dataT->code[v], dataT->bitlen[v]);
    }
  }
  if (!dataT->has[256]) { rc = ZO_E_CORRUPTED; goto done; } /* :222-226 */
  bws_writeRangeCoded(stream, dataT->code[256], dataT->bitlen[256]);
  if (stream->err) rc = stream->err;
done:
  free(lz);
  free(dataT);
  return rc;
}

/* ---- src/deflate.ts:14-39 deflate ----------------------------------------- */
int zo_deflate_raw_blk(const uint8_t *in, size_t n, size_t block_len, uint8_t **out, size_t *out_len) {
  size_t streamHeap = (n < BLOCK_MAX_BUFFER_LEN / 2) ? BLOCK_MAX_BUFFER_LEN : n * 2; /* :16 */
  BitWriteStream s;
  memset(&s, 0, sizeof(s));
  s.buffer = (uint8_t *)calloc(streamHeap, 1); /* :17 (Uint8Array is zero-filled) */
  if (!s.buffer) return ZO_E_NOMEM;
  s.length = streamHeap;
  size_t processedLength = 0, targetLength = 0;
  int rc = 0;
  for (;;) { /* :20-34 */
    if (processedLength + block_len >= n) {
      targetLength = n - processedLength;
      bws_writeRange(&s, 1, 1);
    } else {
      targetLength = block_len;
      bws_writeRange(&s, 0, 1);
    }
    bws_writeRange(&s, 2 /* BTYPE.DYNAMIC */, 2); /* :28 */
    rc = deflate_dynamic_block(&s, in, n, processedLength, targetLength);
    if (rc) break;
    processedLength += block_len;
    if (processedLength >= n) break;
  }
  if (!rc && s.nowBitsIndex != 0) bws_writeRange(&s, 0, 8 - s.nowBitsIndex); /* :35-37 */
  if (!rc && s.err) rc = s.err;
  if (rc) { free(s.buffer); *out = NULL; *out_len = 0; return rc; }
  *out = s.buffer;
  *out_len = s.bufferIndex; /* :38 */
  return 0;
}
int zo_deflate_raw(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  return zo_deflate_raw_blk(in, n, BLOCK_MAX_BUFFER_LEN, out, out_len);
}

/* ---- src/zlib.ts:25-49 deflate (framing) ---------------------------------- */
int zo_deflate(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  uint8_t *data = NULL;
  size_t dlen = 0;
  int rc = zo_deflate_raw(in, n, &data, &dlen);
  if (rc) { *out = NULL; *out_len = 0; return rc; }
  uint8_t *o = (uint8_t *)malloc(dlen + 6);
  if (!o) { free(data); return ZO_E_NOMEM; }
  o[0] = 8 | (7 << 4);                  /* CMF: CM=8, CINFO=7            :28-30 */
  o[1] = (uint8_t)(28 | (0 << 5) | (2 << 6)); /* FLG: FCHECK=28, FDICT=0, FLEVEL=2 :31-34 */
  memcpy(o + 2, data, dlen);
  uint32_t a = zo_adler32(in, n); /* :36 */
  o[dlen + 2] = (uint8_t)(a >> 24); /* :37-40 big endian */
  o[dlen + 3] = (uint8_t)(a >> 16);
  o[dlen + 4] = (uint8_t)(a >> 8);
  o[dlen + 5] = (uint8_t)a;
  free(data);
  *out = o;
  *out_len = dlen + 6;
  return 0;
}

/* ---- src/utils/BitReadStream.ts:1-50 -------------------------------------- */
typedef struct {
  const uint8_t *buffer;
  size_t length;
  size_t bufferIndex;
  uint32_t nowBits;
  int nowBitsLength;
  int isEnd;
  int err;
} BitReadStream;

static inline uint32_t brs_byte(const BitReadStream *s, size_t i) { return i < s->length ? s->buffer[i] : 0; /* undefined << k === 0 */ }

static void brs_init(BitReadStream *s, const uint8_t *buf, size_t n, size_t offset) { /* :7-12 */
  s->buffer = buf; s->length = n; s->bufferIndex = offset;
  s->nowBits = offset < n ? buf[offset] : 0;
  s->nowBitsLength = 8; s->isEnd = 0; s->err = 0;
}
static int brs_read(BitReadStream *s) { /* :14-31 */
  if (s->isEnd) { s->err = ZO_E_LACK; return 0; }
  int bit = (int)(s->nowBits & 1);
  if (s->nowBitsLength > 1) {
    s->nowBitsLength--;
    s->nowBits >>= 1;
  } else {
    s->bufferIndex++;
    if (s->bufferIndex < s->length) {
      s->nowBits = s->buffer[s->bufferIndex];
      s->nowBitsLength = 8;
    } else {
      s->nowBitsLength = 0;
      s->isEnd = 1;
    }
  }
  return bit;
}
static uint32_t brs_readRange(BitReadStream *s, int length) { /* :32-41 */
  while (s->nowBitsLength <= length) {
    s->nowBits |= brs_byte(s, ++s->bufferIndex) << s->nowBitsLength;
    s->nowBitsLength += 8;
  }
  uint32_t bits = s->nowBits & ((1u << length) - 1);
  s->nowBits >>= length;
  s->nowBitsLength -= length;
  return bits;
}
static uint32_t brs_readRangeCoded(BitReadStream *s, int length) { /* :42-49 */
  uint32_t bits = 0;
  for (int i = 0; i < length; i++) {
    bits <<= 1;
    bits |= (uint32_t)brs_read(s);
    if (s->err) return 0;
  }
  return bits;
}

/* ---- src/utils/Uint8WriteStream.ts:1-25 (growth policy is not observable) -- */
typedef struct {
  uint8_t *buffer;
  size_t index, length;
  int err;
} Uint8WriteStream;
static void u8ws_write(Uint8WriteStream *b, int value) {
  if (b->length <= b->index) {
    size_t nl = b->length ? b->length * 2 : 4096;
    uint8_t *nb = (uint8_t *)realloc(b->buffer, nl);
    if (!nb) { b->err = ZO_E_NOMEM; return; }
    b->buffer = nb; b->length = nl;
  }
  b->buffer[b->index++] = (uint8_t)value;
}
/* buffer.buffer[k] with k possibly negative: `undefined`, stored as 0. */
static inline int u8ws_at(const Uint8WriteStream *b, long long k) { return (k >= 0 && (size_t)k < b->index) ? b->buffer[k] : 0; }

/* ---- src/huffman.ts:8-39 generateHuffmanTable ------------------------------ */
/* table[bitlen][code] = value; restated as dense arrays per bit length. */
#define DT_MAXLEN 16
typedef struct {
  int lmin, lmax;       /* over lengths that occur; lmin > lmax when empty */
  uint32_t first[DT_MAXLEN]; /* first code of this length */
  int count[DT_MAXLEN];
  int offset[DT_MAXLEN];
  int syms[HT_MAXSYM];  /* values sorted by (length, value) */
} DecodeTable;

static void decode_table_build(DecodeTable *t, const int *lens, int n) {
  memset(t, 0, sizeof(*t));
  t->lmin = 1 << 30; t->lmax = 0;
  for (int i = 0; i < n; i++)
    if (lens[i] > 0) {
      t->count[lens[i]]++;
      if (lens[i] < t->lmin) t->lmin = lens[i];
      if (lens[i] > t->lmax) t->lmax = lens[i];
    }
  int off = 0;
  uint32_t code = 0;
  for (int l = t->lmin; l <= t->lmax && l < DT_MAXLEN; l++) { /* :22-37 */
    t->first[l] = code;
    t->offset[l] = off;
    for (int i = 0; i < n; i++)
      if (lens[i] == l) t->syms[off++] = i; /* values.sort ascending (:25-29) */
    code += (uint32_t)t->count[l];
    code <<= 1;
  }
}
/* the lookup loop at src/inflate.ts:238-252 / 84-96 / 155-168 */
static int decode_symbol(BitReadStream *s, const DecodeTable *t, int *err) {
  if (t->lmin > t->lmax) {
    /* empty table: Math.min over no keys leaves codelenMin = Number.MAX_SAFE_INTEGER (src/inflate.ts:139-147,
     * 206-224), so readRangeCoded(codelenMin) (src/utils/BitReadStream.ts:42-49) keeps calling read() until the
     * buffer is exhausted and read() throws */
    *err = ZO_E_LACK;
    return -1;
  }
  /* NOT in the reference: a guard for inputs on which the reference never returns.  Past the end of the buffer every
   * bit reads as zero (readRange, src/utils/BitReadStream.ts:33-35) and isEnd is only set by a read() that takes the
   * last bit of a byte (:21-28).  A symbol loop whose all-zero token is a multiple of 8 bits long and whose read()s avoid
   * that bit goes on for ever (writing output until the JS heap is exhausted, or nothing at all).  Its state is the bit
   * offset inside a byte, so 8 tokens (<= 48 bits each) past the end without an exit prove the cycle: a coded symbol
   * that STARTS 512 or more bits past the end is reported as ZO_E_RUNAWAY.  The GPU decoder applies the same rule
   * (zlib.es_b200/csrc/inflate.cuh, inf_coded). */
  if (!s->isEnd) {
    const unsigned long long P = ((unsigned long long)s->bufferIndex + 1) * 8 - (unsigned long long)s->nowBitsLength;
    if (P >= (unsigned long long)s->length * 8 + 512) { *err = ZO_E_RUNAWAY; return -1; }
  }
  int codelen = t->lmin;
  uint32_t code = brs_readRangeCoded(s, t->lmin);
  if (s->err) { *err = s->err; return -1; }
  for (;;) {
    if (code >= t->first[codelen] && code - t->first[codelen] < (uint32_t)t->count[codelen])
      return t->syms[t->offset[codelen] + (int)(code - t->first[codelen])];
    if (t->lmax <= codelen) { *err = ZO_E_CORRUPTED; return -1; }
    codelen++;
    code <<= 1;
    code |= (uint32_t)brs_read(s);
    if (s->err) { *err = s->err; return -1; }
  }
}

/* The match part of the symbol loop, src/inflate.ts:101-116 (fixed) / :260-290 (dynamic).  lenCode 29/30 (symbols
 * 286/287) and distance codes 30/31 index past the ends of the const.ts tables: JS yields `undefined` there, and
 *   - `0 < undefined` is false, so no extra bits are read for such a code;
 *   - an undefined length makes `i < repeatLengthValue` false at once: the distance is still decoded (and its extra
 *     bits consumed), nothing is written, no error;
 *   - an undefined distance makes repeatStartIndex NaN, buffer.buffer[NaN + i] is undefined, and
 *     Uint8WriteStream.write(undefined) stores 0: `len` zero bytes are written, no error. */
static int copy_match(BitReadStream *s, Uint8WriteStream *b, int lenCode, const DecodeTable *distT, int fixedDist) {
  int err = 0;
  const int lenDefined = lenCode < 29;
  int len = lenDefined ? LENGTH_EXTRA_BIT_BASE[lenCode] : 0;
  if (lenDefined && 0 < LENGTH_EXTRA_BIT_LEN[lenCode]) len += (int)brs_readRange(s, LENGTH_EXTRA_BIT_LEN[lenCode]);
  int dc;
  if (fixedDist) {
    dc = (int)brs_readRangeCoded(s, 5); /* src/inflate.ts:107 */
    if (s->err) return s->err;
  } else {
    dc = decode_symbol(s, distT, &err); /* :267-281 */
    if (err) return err;
  }
  const int distDefined = dc < 30;
  int dist = distDefined ? DISTANCE_EXTRA_BIT_BASE[dc] : 0;
  if (distDefined && 0 < DISTANCE_EXTRA_BIT_LEN[dc]) dist += (int)brs_readRange(s, DISTANCE_EXTRA_BIT_LEN[dc]);
  if (!lenDefined) return 0;                      /* `i < undefined` never holds */
  long long startIdx = (long long)b->index - dist; /* :287 */
  for (int i = 0; i < len; i++) {                 /* :288-290 */
    u8ws_write(b, distDefined ? u8ws_at(b, startIdx + i) : 0);
    if (b->err) return b->err;
  }
  return 0;
}

/* src/inflate.ts:42-55 */
static int inflate_uncompressed_block(BitReadStream *s, Uint8WriteStream *b) {
  if (s->nowBitsLength < 8) brs_readRange(s, s->nowBitsLength);
  uint32_t LEN = brs_readRange(s, 8); LEN |= brs_readRange(s, 8) << 8;
  uint32_t NLEN = brs_readRange(s, 8); NLEN |= brs_readRange(s, 8) << 8;
  if (LEN + NLEN != 65535) return ZO_E_CORRUPTED;
  for (uint32_t i = 0; i < LEN; i++) {
    u8ws_write(b, (int)brs_readRange(s, 8));
    if (b->err) return b->err;
  }
  return 0;
}

/* shared symbol loop of src/inflate.ts:76-117 (fixed) and :237-291 (dynamic) */
static int inflate_symbols(BitReadStream *s, Uint8WriteStream *b, const DecodeTable *dataT, const DecodeTable *distT, int fixedDist) {
  while (!s->isEnd) {
    int err = 0;
    int v = decode_symbol(s, dataT, &err);
    if (err) return err;
    if (v < 256) {
      u8ws_write(b, v);
      if (b->err) return b->err;
      continue;
    }
    if (v == 256) break;
    err = copy_match(s, b, v - 257, distT, fixedDist);
    if (err) return err;
  }
  return 0;
}

/* src/inflate.ts:14,57-118 */
static int inflate_fixed_block(BitReadStream *s, Uint8WriteStream *b) {
  int lens[288];
  for (int i = 0; i <= 287; i++) lens[i] = (i <= 143) ? 8 : (i <= 255) ? 9 : (i <= 279) ? 7 : 8; /* src/huffman.ts:41-53 */
  DecodeTable t;
  decode_table_build(&t, lens, 288);
  return inflate_symbols(s, b, &t, NULL, 1);
}

/* src/inflate.ts:120-292 */
static int inflate_dynamic_block(BitReadStream *s, Uint8WriteStream *b) {
  int HLIT = (int)brs_readRange(s, 5) + 257;
  int HDIST = (int)brs_readRange(s, 5) + 1;
  int HCLEN = (int)brs_readRange(s, 4) + 4;
  int cll[19] = {0};
  for (int i = 0; i < HCLEN; i++) cll[CODELEN_VALUES[i]] = (int)brs_readRange(s, 3); /* :125-136 */
  DecodeTable clT;
  decode_table_build(&clT, cll, 19);

  int dataLens[320] = {0}, distLens[64] = {0};
  int codesNumber = HLIT + HDIST;
  int codelen = 0;
  for (int i = 0; i < codesNumber;) { /* :156-202 */
    int err = 0;
    int rl = decode_symbol(s, &clT, &err);
    if (err) return err;
    int repeat;
    if (rl == 16) {
      repeat = 3 + (int)brs_readRange(s, 2);
    } else if (rl == 17) {
      repeat = 3 + (int)brs_readRange(s, 3);
      codelen = 0;
    } else if (rl == 18) {
      repeat = 11 + (int)brs_readRange(s, 7);
      codelen = 0;
    } else {
      repeat = 1;
      codelen = rl;
    }
    if (codelen <= 0) {
      i += repeat;
    } else {
      while (repeat) {
        if (i < HLIT) { if (i < 320) dataLens[i] = codelen; i++; }
        else { if (i - HLIT < 64) distLens[i - HLIT] = codelen; i++; }
        repeat--;
      }
    }
  }
  DecodeTable *dataT = (DecodeTable *)malloc(sizeof(DecodeTable) * 2);
  if (!dataT) return ZO_E_NOMEM;
  DecodeTable *distT = dataT + 1;
  decode_table_build(dataT, dataLens, 320); /* :203 */
  decode_table_build(distT, distLens, 64);  /* :204 */
  int rc = inflate_symbols(s, b, dataT, distT, 0);
  free(dataT);
  return rc;
}

/* src/inflate.ts:16-40 */
int zo_inflate_raw(const uint8_t *in, size_t n, size_t offset, uint8_t **out, size_t *out_len) {
  Uint8WriteStream b = {0};
  BitReadStream s;
  brs_init(&s, in, n, offset);
  int bFinal = 0, rc = 0;
  while (bFinal != 1) {
    bFinal = (int)brs_readRange(&s, 1);
    int bType = (int)brs_readRange(&s, 2);
    if (bType == 0) rc = inflate_uncompressed_block(&s, &b);
    else if (bType == 1) rc = inflate_fixed_block(&s, &b);
    else if (bType == 2) rc = inflate_dynamic_block(&s, &b);
    else rc = ZO_E_BTYPE3;
    if (rc) break;
    if (bFinal == 0 && s.isEnd) { rc = ZO_E_INSUFFICIENT; break; }
  }
  if (rc) { free(b.buffer); *out = NULL; *out_len = 0; return rc; }
  if (!b.buffer) b.buffer = (uint8_t *)malloc(1);
  *out = b.buffer;
  *out_len = b.index;
  return 0;
}

/* src/zlib.ts:11-23 */
int zo_inflate(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  BitReadStream s;
  brs_init(&s, in, n, 0);
  uint32_t CM = brs_readRange(&s, 4);
  if (CM != 8) { *out = NULL; *out_len = 0; return ZO_E_NOT_DEFLATE; }
  /* CINFO, FCHECK, FDICT, FLEVEL are read and ignored (:17-20); the Adler-32
   * trailer is never read (:22). */
  return zo_inflate_raw(in, n, 2, out, out_len);
}
