/*
 * zlibes_oracle.h — CPU restatement of zprodev/zlib.es v0.6.0 (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for the B200 hot path.  It restates, function by
 * function, the reference's TypeScript (citations are file:line under
 * /root/reference/src).  Nothing in the product (zlib.es_b200/, include/) may
 * link, import or call it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg do.
 *
 * Pinning status:
 *   inflate  — pinned by the reference's own 4 known-answer vectors
 *              (test/index.js:7-10,37-42 + test/data/compressed.bin).
 *   adler32  — pinned by the trailers of those vectors (2B23056C, 140FA15B).
 *   deflate  — exact bits are PARITY UNPINNED by the reference (its tests only
 *              round-trip, test/index.js:46-109, and no JS engine exists in the
 *              image to run it).  The restatement is cross-checked against the
 *              survey's independent model table and against
 *              system zlib as a second decoder.
 */
#ifndef ZLIBES_ORACLE_H
#define ZLIBES_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes; zo_strerror() returns the reference's exact Error messages. */
enum {
  ZO_OK = 0,
  ZO_E_NOT_DEFLATE = 1,   /* 'Not compressed by deflate'      src/zlib.ts:15 */
  ZO_E_BTYPE3 = 2,        /* 'Not supported BTYPE : 3'        src/inflate.ts:32 */
  ZO_E_INSUFFICIENT = 3,  /* 'Data length is insufficient'    src/inflate.ts:35 */
  ZO_E_CORRUPTED = 4,     /* 'Data is corrupted'              src/inflate.ts:50,88,166,247,276; src/deflate.ts:172-224 */
  ZO_E_LACK = 5,          /* 'Lack of data length'            src/utils/BitReadStream.ts:15, BitWriteStream.ts:15 */
  ZO_E_NOMEM = 6,
  ZO_E_RUNAWAY = 20       /* the reference never returns on this input (see decode_symbol in zlibes_oracle.c) */
};

const char *zo_strerror(int code);
void zo_free(void *p);

/* src/adler32.ts:1-10 (returned unsigned, i.e. the JS value >>> 0). */
uint32_t zo_adler32(const uint8_t *in, size_t n);

/* src/zlib.ts:25-49 — zlib framing around src/deflate.ts:14-39.  *out is malloc'd. */
int zo_deflate(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
/* src/deflate.ts:14-39 — raw deflate only. */
int zo_deflate_raw(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);

/* src/zlib.ts:11-23 + src/inflate.ts:16-40. *out is malloc'd. */
int zo_inflate(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
/* src/inflate.ts:16 with explicit offset (raw deflate when offset = 0). */
int zo_inflate_raw(const uint8_t *in, size_t n, size_t offset, uint8_t **out, size_t *out_len);

/* src/lz77.ts:24-119 — token count of one block (the trailing two literals
 * included, EOB excluded), and the number of match tokens. */
int zo_lz77_count(const uint8_t *in, size_t n, size_t start, size_t len,
                  uint32_t *n_tokens, uint32_t *n_matches);

/* Variant used only for design studies: same algorithm with a
 * different block length than BLOCK_MAX_BUFFER_LEN.  Not part of parity. */
int zo_deflate_raw_blk(const uint8_t *in, size_t n, size_t block_len, uint8_t **out, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
