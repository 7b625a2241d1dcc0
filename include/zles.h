/*
 * zles.h — C ABI of the B200-native zlib codec (libzles.so).
 *
 * This is the drop-in boundary for zprodev/zlib.es.  The reference has no FFI of
 * its own; its public surface is two functions,
 *     export declare function inflate(input: Uint8Array): Uint8Array;
 *     export declare function deflate(input: Uint8Array): Uint8Array;
 * (/root/reference/dist/tsc/zlib.d.ts:4-5, implemented at /root/reference/src/zlib.ts:11-49),
 * plus synchronous `throw new Error(msg)` with five fixed messages.  The entry
 * points below are what an N-API / ctypes binding for that surface binds; the
 * binding a maintainer would add is shown in INTEGRATION.md and implemented in
 * zlib.es_b200/node/addon.c (N-API) and zlib.es_b200/_capi.py (ctypes).
 *
 * Conventions: plain pointers and sizes, no exceptions, int status (0 = ok),
 * caller owns every buffer.  "host" functions take host pointers and do the
 * host<->device copies themselves; "dev" functions take device pointers (the
 * data is already in HBM) and run on the handle's stream.  There is no CPU
 * fallback: every call fails with ZLES_E_CUDA when no usable GPU is present.
 */
#ifndef ZLES_H
#define ZLES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  1..5 carry the reference's exact Error messages (zles_strerror). */
enum {
  ZLES_OK = 0,
  ZLES_E_NOT_DEFLATE = 1,  /* 'Not compressed by deflate'    /root/reference/src/zlib.ts:15 */
  ZLES_E_BTYPE3 = 2,       /* 'Not supported BTYPE : 3'      /root/reference/src/inflate.ts:32 */
  ZLES_E_INSUFFICIENT = 3, /* 'Data length is insufficient'  /root/reference/src/inflate.ts:35 */
  ZLES_E_CORRUPTED = 4,    /* 'Data is corrupted'            /root/reference/src/inflate.ts:50,88,166,247,276 */
  ZLES_E_LACK = 5,         /* 'Lack of data length'          /root/reference/src/utils/BitReadStream.ts:15 */
  ZLES_E_OUTPUT_FULL = 16, /* output buffer too small; *out_len holds the size needed */
  ZLES_E_CUDA = 17,        /* CUDA runtime error (zles_last_cuda_error) */
  ZLES_E_ARG = 18,
  ZLES_E_NOMEM = 19,
  ZLES_E_CHECKSUM = 21,    /* zles_gzip_inflate only: CRC-32 or ISIZE of the trailer do not match the data */
  ZLES_E_RUNAWAY = 20      /* 'stream never ends': the reference does not return on this input — past the end of the
                              buffer it reads zero bits for ever (/root/reference/src/utils/BitReadStream.ts:33-35 never
                              sets isEnd) in a symbol loop that only stops on isEnd (/root/reference/src/inflate.ts:78,237) */
};

typedef struct zles_ctx zles_ctx;

/* Library / error helpers. */
const char *zles_version(void);
const char *zles_strerror(int code);
const char *zles_last_cuda_error(void);

/* One context per (device, stream); calls on one context are serialised by the caller. */
int zles_ctx_create(int device, zles_ctx **ctx);
void zles_ctx_destroy(zles_ctx *ctx);
/* Use an existing CUDA stream (cudaStream_t passed as void*) instead of the context's own. */
int zles_ctx_set_stream(zles_ctx *ctx, void *cuda_stream);
/* Encoder search depth (the reference's FAST_INDEX_CHECK_MAX, /root/reference/src/lz77.ts:7, scaled for an
 * all-positions search): max_checks = candidates compared per position (1..32, default 32; larger values are clamped
 * to 32); lazy = defer a match by one literal when the next position has a longer one (default 1).  min_checks and
 * good_len are accepted for signature compatibility with the reference's constants and have NO effect: the nearest
 * candidate of the longest class is the one that is extended. */
int zles_ctx_set_level(zles_ctx *ctx, uint32_t max_checks, uint32_t min_checks, uint32_t good_len, uint32_t lazy);
/* Which window the third 32 KiB block of a 128 KiB chunk sees (/root/reference/src/lz77.ts:49 gives every position the
 * 32 KiB before it inside its chunk).  1 (default): none, like block 0 — blocks {0,1} and {2,3} each share one match-finder
 * pass; a third block that took a third more tokens than the fourth (periodic or repeated material: what the window is
 * for) is matched again with its window.  0: the block before it, like blocks 1 and 3 — three passes per chunk, ~25 %
 * slower, output ~1.4 % smaller on text.
 * Either way the stream stays within 3 % of the reference's size on the benchmark corpora ("Size"). */
int zles_ctx_set_window_mode(zles_ctx *ctx, uint32_t mode);
/* Host-buffer inflate of our own streams runs slab by slab (finished slabs are copied to the host while the next one is
 * decoded): blocks of 32 KiB per slab, a multiple of 4; 0 (default) = automatic (a quarter of the stream, 64..512 MiB; streams below 128 MiB are decoded in one go). */
int zles_ctx_set_slab_blocks(zles_ctx *ctx, uint32_t blocks);
/* Host-buffer inflate: a stream of at least `bytes` compressed bytes (default 96 MiB) is copied to the device in pieces
 * (a sixth of the threshold, doubling up to eight times that) that are scanned and decoded as they land, so that the copy
 * in, the decode and the copy out all overlap.  Streams that turn out not to be ours take the general path afterwards. */
int zles_ctx_set_stream_min(zles_ctx *ctx, size_t bytes);
/* Host batch calls go through the device a slab of buffers at a time (the workspace is sized by the slab; the results of
 * slab k travel to the host and into the caller's buffers while slab k + 1 is computed): at most `buffers` per slab
 * (default and maximum 16,384; also bounded by 64 MiB of input and 256 MiB of output room). */
int zles_ctx_set_batch_slab(zles_ctx *ctx, uint32_t buffers);
/* Number of kernels launched through this context since creation (bench.py's gpu_launches). */
uint64_t zles_ctx_launches(const zles_ctx *ctx);
/* Per-kernel device timing: when on, every launch is bracketed by CUDA events on the context's
 * stream; zles_ctx_kernel_time returns the summed duration and launch count of one kernel
 * (e.g. "k_lz") since timing was switched on.  Both calls synchronise the stream. */
int zles_ctx_set_timing(zles_ctx *ctx, int on);
int zles_ctx_kernel_time(zles_ctx *ctx, const char *kernel, double *ms_total, uint64_t *launches);
/* Plain device memory (cudaMalloc/cudaFree), e.g. for a buffer that is exported with zles_ipc_export,
 * and a synchronous copy between any two device / host / peer-mapped pointers. */
int zles_dev_alloc(zles_ctx *ctx, size_t n, void **d_ptr);
int zles_dev_free(zles_ctx *ctx, void *d_ptr);
int zles_dev_copy(zles_ctx *ctx, void *dst, const void *src, size_t n);
/* The same without waiting: the copy is ordered on the context's stream; a host-side source must be pinned and stay
 * untouched until the stream has been synchronised (zles_ctx_sync or any synchronising call). */
int zles_dev_copy_async(zles_ctx *ctx, void *dst, const void *src, size_t n);
int zles_ctx_sync(zles_ctx *ctx);

/* ---- the drop-in pair: host buffers in, host buffers out ----------------------
 * zles_deflate   replaces zlib.deflate  (/root/reference/src/zlib.ts:25-49)
 * zles_inflate   replaces zlib.inflate  (/root/reference/src/zlib.ts:11-23)
 * Both use a process-wide default context on device 0 when ctx is NULL. */
size_t zles_deflate_bound(size_t n);
int zles_deflate(zles_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int zles_inflate(zles_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
/* inflate with a library-allocated result (the reference's return-a-new-array shape); free with zles_free. */
int zles_inflate_alloc(zles_ctx *ctx, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
void zles_free(void *p);
/* calcAdler32 >>> 0 (/root/reference/src/adler32.ts:1-10). */
int zles_adler32(zles_ctx *ctx, const uint8_t *in, size_t n, uint32_t *adler);

/* ---- wire-format siblings behind the same kernels --------------------------------
 * raw:  the deflate data alone (RFC 1951) — what the reference's deflate core returns (/root/reference/src/deflate.ts:14)
 *       and what its inflate(input, offset = 0) reads from byte `offset` on (/root/reference/src/inflate.ts:16).
 * gzip: RFC 1952 — 10-byte header (no name, no time), the same deflate data, CRC-32 and ISIZE (little endian).
 *       zles_gzip_inflate reads one member (any header flags), takes the trailer from the last eight bytes of the buffer and
 *       verifies both fields (ZLES_E_CHECKSUM); other errors are those of zles_inflate.  Use zles_deflate_bound for `cap`.
 * zles_crc32 is the container's checksum (IEEE 802.3), computed on the GPU in shard-combinable form:
 * crc32(A || B) = zles_crc32_combine(crc32(A), crc32(B), len(B)). */
int zles_deflate_raw(zles_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int zles_inflate_raw(zles_ctx *ctx, const uint8_t *in, size_t n, size_t offset, uint8_t *out, size_t cap, size_t *out_len);
int zles_gzip_deflate(zles_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int zles_gzip_inflate(zles_ctx *ctx, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int zles_crc32(zles_ctx *ctx, const uint8_t *in, size_t n, uint32_t *crc);
int zles_dev_crc32(zles_ctx *ctx, const uint8_t *d_in, size_t n, uint32_t *crc);
uint32_t zles_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* ---- batches of independent buffers (BASELINE config 3: 262,144 x 4 KiB) --------
 * Buffer i is in[in_off[i] .. in_off[i+1]); its result is written at out + out_off[i]
 * (capacity out_off[i+1] - out_off[i]) and its length to out_len[i]; status[i] gets
 * the per-buffer code.  Host pointers.  Returns 0 when every buffer succeeded, ZLES_E_CUDA / ZLES_E_ARG / ZLES_E_NOMEM when
 * the call as a whole failed (status[] is then untouched), otherwise the LARGEST per-buffer status — e.g. OUTPUT_FULL (16)
 * hides CORRUPTED (4): inspect status[], do not branch on the return value alone.  For inflate, out_len[i] of a buffer
 * with status OUTPUT_FULL holds the size it needs; only the bytes actually produced are copied back to the host.
 * deflate: a buffer of at most 64 KiB compresses to exactly the bytes zles_deflate gives it on its own; a longer one may
 * come out a little smaller (every block but a chunk's first sees its 32 KiB window in a batch). */
int zles_deflate_batch(zles_ctx *ctx, const uint8_t *in, const uint64_t *in_off, uint32_t count, uint8_t *out,
                       const uint64_t *out_off, uint64_t *out_len, int32_t *status);
int zles_inflate_batch(zles_ctx *ctx, const uint8_t *in, const uint64_t *in_off, uint32_t count, uint8_t *out,
                       const uint64_t *out_off, uint64_t *out_len, int32_t *status);

/* ---- device-resident forms (inputs and outputs already in HBM) -------------------
 * Same results as the host forms, no host<->device payload copies. */
int zles_dev_deflate(zles_ctx *ctx, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t cap, size_t *out_len);
int zles_dev_inflate(zles_ctx *ctx, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t cap, size_t *out_len);
int zles_dev_adler32(zles_ctx *ctx, const uint8_t *d_in, size_t n, uint32_t *adler);
int zles_dev_deflate_batch(zles_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint32_t count, uint8_t *d_out,
                           const uint64_t *d_out_off, uint64_t *d_out_len, int32_t *d_status);
int zles_dev_inflate_batch(zles_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint32_t count, uint8_t *d_out,
                           const uint64_t *d_out_off, uint64_t *d_out_len, int32_t *d_status);

/* ---- sharded deflate: one stream, chunks split over the GPUs of a node -----------
 * Rank r compresses a contiguous shard (a multiple of 128 KiB except on the last
 * rank).  Phase 1 runs the match finder, code construction and layout and reports
 * the shard's compressed size and Adler-32 partial sums; the ranks exchange those
 * (NCCL all-gather in zlib.es_b200/dist.py); phase 2 writes the shard's bytes at
 * its global offset — d_dst may be a peer-mapped pointer (zles_ipc_*), so the
 * blocks are stored straight into the final stream over NVLink.  Every block is
 * byte aligned (it ends with an empty stored block), which is what lets shards
 * be written independently and lets inflate decode blocks in parallel. */
typedef struct {
  uint64_t comp_bytes;   /* compressed size of the shard (no zlib header/trailer) */
  uint64_t raw_bytes;    /* shard length */
  uint64_t adler_a;      /* sum of bytes mod 65521 */
  uint64_t adler_b;      /* sum of (raw_bytes - i) * byte[i] mod 65521 */
  uint64_t n_blocks;     /* 32 KiB deflate blocks in the shard */
} zles_shard_info;
int zles_dev_deflate_phase1(zles_ctx *ctx, const uint8_t *d_in, size_t n, int is_last_shard, zles_shard_info *info);
/* byte offsets of the shard's blocks after the last phase 1 (device pointer to n_blocks+1 uint64) */
int zles_dev_deflate_block_offsets(zles_ctx *ctx, const uint64_t **d_offsets);
int zles_dev_deflate_phase2(zles_ctx *ctx, uint8_t *d_dst);
/* Combine the per-shard sums of all ranks (in rank order) into the stream's Adler-32. */
uint32_t zles_adler32_combine_shards(const zles_shard_info *infos, uint32_t count);
/* Inflate `n` bytes of marker-delimited blocks that start at a 128 KiB chunk boundary of the
 * original data (no zlib header); used by the sharded inflate.  has_final = 0: the bytes are a
 * shard that ends on a block boundary before the stream's last block. */
int zles_dev_inflate_segment(zles_ctx *ctx, const uint8_t *d_in, size_t n, int has_final, uint8_t *d_out, size_t cap,
                             size_t *out_len);

/* ---- one process, several GPUs (sharding behind the drop-in boundary) --------------
 * The reference is a synchronous library inside ONE process (/root/reference/src/zlib.ts:11-49); a Node process that
 * loads the addon can only use more than one GPU if the library shards below this ABI.  A zles_mgpu owns a context and
 * a worker thread per device.  Host buffers in, host buffers out, same results byte for byte as the single-device calls
 * (chunks are independent; shards are contiguous runs of whole 128 KiB chunks):
 *   deflate  every device compresses its shard; one host-side exchange of the shards' compressed sizes (what the
 *            multi-process form all-gathers over NCCL) places every shard in the caller's buffer;
 *   inflate  the shards are found from the stream itself (marker scan of equal byte slices on every device), decoded
 *            slab by slab with the copies back overlapped.  Streams that are not ours go to device 0's general path.
 * zles_init(device_mask) installs a process-wide zles_mgpu over the devices whose bit is set (bit d = CUDA device d) as
 * the default that zles_deflate / zles_inflate / zles_inflate_alloc use when ctx is NULL — what the N-API addon calls at
 * load time with ZLES_DEVICES; mask 0 or 1 keeps the single-device default.  zles_shutdown releases the defaults. */
typedef struct zles_mgpu zles_mgpu;
int zles_init(uint32_t device_mask);
void zles_shutdown(void);
int zles_mgpu_create(const int *devices, int count, zles_mgpu **out);
void zles_mgpu_destroy(zles_mgpu *m);
int zles_mgpu_device_count(const zles_mgpu *m);
/* Inputs with fewer than `bytes` per device use fewer devices (default 4 MiB). */
int zles_mgpu_set_min_shard(zles_mgpu *m, size_t bytes);
zles_ctx *zles_mgpu_ctx(zles_mgpu *m, int index);   /* the context of device `index` (encoder settings, timing) */
uint64_t zles_mgpu_launches(const zles_mgpu *m);
int zles_mgpu_deflate(zles_mgpu *m, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int zles_mgpu_inflate(zles_mgpu *m, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int zles_mgpu_inflate_alloc(zles_mgpu *m, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);

/* Block starts of one of our streams, found from the bytes themselves: `first` and every position in d_in[0 .. n) that
 * follows a 00 00 FF FF marker (ascending, relative to d_in, starts[0] = first).  The sharded inflate runs it on every
 * rank's slice of the compressed bytes and all-gathers the results (zlib.es_b200/dist.py, inflate_from_stream).
 * ZLES_E_OUTPUT_FULL: *count holds the number found; ZLES_E_CORRUPTED: more candidates than a stream of ours can have. */
int zles_dev_scan_blocks(zles_ctx *ctx, const uint8_t *d_in, size_t n, uint64_t first, uint64_t *starts, size_t cap, size_t *count);

/* CUDA IPC helpers so that another process can map a device buffer (64-byte handles). */
int zles_ipc_export(const void *d_ptr, uint8_t handle[64]);
int zles_ipc_open(const uint8_t handle[64], void **d_ptr);
int zles_ipc_close(void *d_ptr);

/* Synthetic corpora of BASELINE.json / SURVEY.md §8d, generated in HBM (kind: 0 text,
 * 1 structured binary, 2 random, 3 mixed); `offset` is the absolute byte offset of the
 * first byte so that shards of one corpus can be generated independently. */
int zles_dev_corpus(zles_ctx *ctx, int kind, uint64_t offset, uint8_t *d_out, size_t n);
/* Same bytes on the host (single thread; for tests and the CPU baseline sample). */
int zles_host_corpus(int kind, uint64_t offset, uint8_t *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
