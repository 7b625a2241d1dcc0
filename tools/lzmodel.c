/* lzmodel.c — CPU model of the GPU matcher/parse for ratio studies (design tool, not product,
 * not oracle).  usage: lzmodel <file> <mode> <D> <min_checks> <good_len> <lazy> [exact_key]
 *   mode 0: sequential candidate walk with early stop (max D / min_checks once best >= good_len)
 *   mode 1: fixed depth: the D nearest entries of the hash run are all evaluated
 * Prints the estimated deflate size (dynamic Huffman per 32 KiB block, exact header cost). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define SUB 32768u
#define CHUNK 131072u
static int KB = 3, T3 = 4096, T4 = 1 << 20, SCORE = 0;
static uint32_t key3(const uint8_t *d, uint32_t p) { uint32_t k = d[p] | (d[p + 1] << 8) | (d[p + 2] << 16); if (KB == 4) k |= (uint32_t)d[p + 3] << 24; return k; }
static uint32_t hash16(uint32_t k) { return (k * 0x9E3779B1u) >> 16; }
static int dbits(uint32_t dist) { int b = 0; uint32_t x = dist - 1; while (x >= 4) { x >>= 1; b++; } return b; }
static int benefit(uint32_t len, uint32_t dist) { return (int)len * 9 - 2 * dbits(dist); }
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }
static uint32_t mlen(const uint8_t *d, uint32_t c, uint32_t p, uint32_t maxlen) { uint32_t o = 0; while (o < maxlen && d[c + o] == d[p + o]) o++; return o; }

/* Huffman code lengths (unlimited depth is fine for a model, clamp 15) */
static void huff(const uint32_t *f, int n, uint8_t *len) {
  int idx[320], m = 0; uint64_t w[640]; int par[640];
  for (int i = 0; i < n; i++) { len[i] = 0; if (f[i]) idx[m++] = i; }
  if (m == 0) return; if (m == 1) { len[idx[0]] = 1; return; }
  int tot = m; int alive[640]; for (int i = 0; i < m; i++) { w[i] = f[idx[i]]; alive[i] = 1; par[i] = -1; }
  for (int r = 0; r < m - 1; r++) { int a = -1, b = -1; for (int i = 0; i < tot; i++) if (alive[i]) { if (a < 0 || w[i] < w[a]) { b = a; a = i; } else if (b < 0 || w[i] < w[b]) b = i; }
    w[tot] = w[a] + w[b]; alive[a] = alive[b] = 0; alive[tot] = 1; par[a] = par[b] = tot; par[tot] = -1; tot++; }
  for (int i = 0; i < m; i++) { int l = 0, x = i; while (par[x] >= 0) { x = par[x]; l++; } len[idx[i]] = l > 15 ? 15 : l; }
}
static const int LBASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const int LEXT[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const int DBASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const int DEXT[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static int lsym(int l) { int s = 28; while (LBASE[s] > l) s--; return s; }
static int dsym(int d) { int s = 29; while (DBASE[s] > d) s--; return s; }

static uint64_t block_cost(const uint32_t *fl, const uint32_t *fd) {
  uint8_t ll[288], dl[32]; huff(fl, 288, ll); huff(fd, 32, dl);
  uint64_t bits = 3;
  for (int i = 0; i < 286; i++) bits += (uint64_t)fl[i] * (ll[i] + (i >= 257 ? LEXT[i - 257] : 0));
  for (int i = 0; i < 30; i++) bits += (uint64_t)fd[i] * (dl[i] + DEXT[i]);
  int hlit = 257, hdist = 1; for (int i = 285; i >= 257; i--) if (ll[i]) { hlit = i + 1; break; } for (int i = 29; i >= 1; i--) if (dl[i]) { hdist = i + 1; break; }
  uint8_t v[320]; int n = 0; for (int i = 0; i < hlit; i++) v[n++] = ll[i]; for (int i = 0; i < hdist; i++) v[n++] = dl[i];
  uint32_t cf[19] = {0}; int extra = 0;
  for (int i = 0; i < n;) { int run = 1; while (i + run < n && v[i + run] == v[i]) run++; int x = v[i]; i += run;
    if (x == 0) { while (run >= 11) { int r = run > 138 ? 138 : run; cf[18]++; extra += 7; run -= r; } if (run >= 3) { cf[17]++; extra += 3; run = 0; } cf[0] += run; }
    else { cf[x]++; run--; while (run >= 3) { int r = run > 6 ? 6 : run; cf[16]++; extra += 2; run -= r; } cf[x] += run; } }
  uint8_t cl[19]; huff(cf, 19, cl); for (int i = 0; i < 19; i++) if (cl[i] > 7) cl[i] = 7;
  bits += 14 + 19 * 3 + extra; for (int i = 0; i < 19; i++) bits += cf[i] * cl[i];
  return bits;
}

int main(int argc, char **argv) {
  if (argc < 7) return 2;
  FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  uint8_t *d = malloc(n + 512); memset(d, 0, n + 512); if (fread(d, 1, n, f) != (size_t)n) return 3; fclose(f);
  int mode = atoi(argv[2]), D = atoi(argv[3]), minc = atoi(argv[4]), good = atoi(argv[5]), lazy = atoi(argv[6]), exact = argc > 7 ? atoi(argv[7]) : 0;
  int merge = argc > 8 ? atoi(argv[8]) : 0; if (argc > 9) KB = atoi(argv[9]); if (argc > 10) T3 = atoi(argv[10]); if (argc > 11) T4 = atoi(argv[11]); if (argc > 12) SCORE = atoi(argv[12]);
  uint64_t total_bits = 0, pairs = 0, npos = 0, ntok = 0;
  uint64_t *srt = malloc(65536 * 8); uint32_t *X = malloc(65536 * 4); uint32_t *R = malloc(SUB * 4);
  uint32_t cfl[288], cfd[32]; uint64_t chunk_sep_bits = 0;
  for (long cs = 0; cs < n; cs += CHUNK) {
    long ce = cs + CHUNK < n ? cs + CHUNK : n; memset(cfl, 0, sizeof cfl); memset(cfd, 0, sizeof cfd); chunk_sep_bits = 0;
    for (long bs = cs; bs < ce; bs += SUB) {
      uint32_t own = (uint32_t)((bs + SUB < ce ? bs + SUB : ce) - bs), hist = (getenv("PAIRS") ? (((bs - cs) / SUB) & 1) : (bs > cs)) ? SUB : 0, L = hist + own;
      const uint8_t *b = d + bs - hist; uint32_t N = L >= 3 ? L - 2 : 0;
      for (uint32_t p = 0; p < N; p++) { uint32_t k = key3(b, p); srt[p] = ((uint64_t)(exact ? k : hash16(k)) << 32) | p; }
      qsort(srt, N, 8, cmp_u64); for (uint32_t i = 0; i < N; i++) X[i] = (uint32_t)srt[i];
      memset(R, 0, own * 4);
      for (uint32_t k = 0; k < N; k++) { uint32_t p = X[k]; if (p < hist) continue; npos++;
        uint32_t maxlen = L - p < 258 ? L - p : 258; if (maxlen < 3) continue; uint32_t hp = (uint32_t)(srt[k] >> 32), kp = key3(b, p);
        uint32_t best = 2, bd = 0, checks = 0;
        for (uint32_t j = k; j-- > 0;) { if ((uint32_t)(srt[j] >> 32) != hp) break; uint32_t c = X[j];
          if (mode == 0) { if (key3(b, c) != kp) continue; if (p - c > 32768) break; checks++; pairs++; uint32_t l = mlen(b, c, p, maxlen); if (l > best) { best = l; bd = p - c; if (l >= maxlen) break; }
            if (checks >= (uint32_t)D || (best >= (uint32_t)good && checks >= (uint32_t)minc)) break; }
          else if (mode == 3) { /* tags: scan minc nearest in the hash run; evaluate up to D whose bytes 4 and 5 match, else the nearest whose byte 4 matches, else the nearest */
            static uint32_t ev5, n4, n3; if (checks == 0) { ev5 = 0; n4 = n3 = 0xffffffff; }
            if (checks >= (uint32_t)minc) break; checks++; if (p - c > 32768) continue;
            if (n3 == 0xffffffff) n3 = c;
            if (b[c + 3] == b[p + 3]) { if (b[c + 4] == b[p + 4]) { if (ev5 < (uint32_t)D) { ev5++; pairs++; uint32_t l = mlen(b, c, p, maxlen); if (l > best) { best = l; bd = p - c; } } } else if (n4 == 0xffffffff) n4 = c; }
            if (j == 0 || (uint32_t)(srt[j - 1] >> 32) != hp || checks >= (uint32_t)minc) { /* end of scan: fall-backs */
              if (best < 5 && n4 != 0xffffffff) { pairs++; uint32_t l = mlen(b, n4, p, maxlen); if (l > best) { best = l; bd = p - n4; } }
              if (best < 4 && n3 != 0xffffffff) { pairs++; uint32_t l = mlen(b, n3, p, maxlen); if (l > best) { best = l; bd = p - n3; } } } }
          else if (mode == 2) { /* scan minc nearest; evaluate the first `good` regardless and up to D whose 4th byte matches */
            if (checks >= (uint32_t)minc) break; checks++; if (p - c > 32768) continue; static uint32_t ev4; if (checks == 1) ev4 = 0; int take = checks <= (uint32_t)good; if (!take && ev4 < (uint32_t)D && b[c + 3] == b[p + 3]) { take = 1; } if (take) { if (b[c+3]==b[p+3]) ev4++; pairs++; uint32_t l = mlen(b, c, p, maxlen); if (l > best) { best = l; bd = p - c; } } }
          else { if (checks >= (uint32_t)D) break; checks++; pairs++; if (p - c > 32768) continue; uint32_t l = mlen(b, c, p, maxlen); if (l >= 3 && (SCORE ? (bd == 0 || benefit(l, p - c) > benefit(best, bd)) : l > best)) { best = l; bd = p - c; } } }
        if (best >= 3 && bd && !(best == 3 && bd > (uint32_t)T3) && !(best == 4 && bd > (uint32_t)T4)) R[p - hist] = (best << 16) | bd; }
      uint32_t fl[288] = {0}, fd[32] = {0}; fl[256] = 1;
      for (uint32_t pos = 0; pos < own;) { uint32_t len = R[pos] >> 16; if (len && lazy && pos + 1 < own && (R[pos + 1] >> 16) > len) len = 0; ntok++;
        if (len) { fl[257 + lsym(len)]++; fd[dsym(R[pos] & 0xffff)]++; pos += len; } else { fl[b[hist + pos]]++; pos++; } }
      chunk_sep_bits += block_cost(fl, fd);
      for (int i = 0; i < 288; i++) cfl[i] += fl[i]; for (int i = 0; i < 32; i++) cfd[i] += fd[i];
    }
    cfl[256] = 1; uint64_t merged = block_cost(cfl, cfd);
    uint64_t bits = (merge && merged < chunk_sep_bits) ? merged : chunk_sep_bits;
    total_bits += ((bits + 3 + 7) / 8 + 4) * 8;
  }
  printf("size %llu  ratio %.4f  pairs/pos %.2f  tok %llu\n", (unsigned long long)(total_bits / 8 + 6), (double)n / (total_bits / 8.0), (double)pairs / npos, (unsigned long long)ntok);
  return 0;
}
