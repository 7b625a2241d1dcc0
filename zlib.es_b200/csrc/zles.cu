// zles.cu — host side of libzles.so: the C ABI of include/zles.h over the sm_100a
// kernels in this directory.
//
// Replaces, for callers, zlib.deflate / zlib.inflate of the reference
// (/root/reference/src/zlib.ts:11-49).  The zlib framing (CMF/FLG header, Adler-32
// trailer, /root/reference/src/zlib.ts:28-46) is done here on the host; everything
// else runs on the GPU.  There is no CPU fallback: without a usable device every
// entry point returns ZLES_E_CUDA.
//
// The same file builds with -DZLES_EMU against tests/emu (CPU thread emulator,
// tests only).
#include "../../include/zles.h"

#include <stdio.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "zles_rt.h"
// kernels
#include "adler32.cuh"
#include "corpus.cuh"
#include "corpus_text.h"
#include "crc32.cuh"
#include "huffman.cuh"
#include "inflate.cuh"
#include "inflate_foreign.cuh"
#include "inflate_spec.cuh"
#include "inflate_fblk.cuh"
#include "lz77.cuh"
#include "pack.cuh"

using namespace zles;

#define ZLES_VERSION_STR "zles-b200 0.1.0 (sm_100a)"

static thread_local std::string g_cuda_err;

static int cuda_fail(zrt_err_t e, const char *what) {
  g_cuda_err = std::string(what) + ": " + zrt_err_str(e);
  return ZLES_E_CUDA;
}
#define CK(expr)                                          \
  do {                                                    \
    zrt_err_t e_ = (expr);                                \
    if (e_ != ZRT_OK) return cuda_fail(e_, #expr);        \
  } while (0)
#define RET(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_) return rc_;     \
  } while (0)

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) zrt_free(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    zrt_err_t e = zrt_malloc(&p, want);
    if (e != ZRT_OK) {
      p = nullptr;
      e = zrt_malloc(&p, bytes);
      if (e != ZRT_OK) { p = nullptr; zrt_last_error(); return cuda_fail(e, "device allocation"); }
      want = bytes;
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) zrt_free(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

// small pinned block for results read back from the device
struct HostMail {
  u64 summary[4];       // deflate: total bytes, adler a, adler b
  u32 ok;               // batch calls: largest per-buffer status
  u32 ncand;
  unsigned long long total;
  u32 adler;
  u8 head[8];           // zlib header / trailer staging
  InfRes res0;
  u64 fres0[4];         // an FbRes (a block decoded on demand)
};

}  // namespace

constexpr u32 BATCH_SLAB_STREAMS = 16384;  // host batch calls: buffers per slab at most

#include "stager.inl"

struct zles_ctx {
  int device = 0;
  zrt_stream_t stream{};
  zrt_stream_t copy_stream{};  // host<->device copies that overlap kernels on `stream`
  zrt_stream_t out_stream{};   // device-to-host copies of finished output slabs (inflate), so that they do not queue behind
                               // host-to-device copies on copy_stream
  bool own_stream = false;
  int sm_count = 148;
  uint64_t launches = 0;
  // encoder search depth (see zles_ctx_set_level)
  u32 max_checks = 32, min_checks = 1, good_len = 8, lazy = 1;
  u32 pair_mode = 1;  // zles_ctx_set_window_mode
  bool no_dev_slabs = false;  // ZLES_NO_DEV_SLABS=1: zles_dev_deflate never cuts a long input into slabs (a measuring aid)
  bool merge_chunks = true;  // k_huff_merge (ZLES_NO_MERGE=1 in the environment turns it off: a debugging aid)
  size_t inf_stream_min = (size_t)96 << 20;  // host-buffer inflate: streams at least this long are copied in pieces, scanned and decoded as they land
  u32 inf_slab_blocks = 0;  // host-buffer inflate of our own streams: blocks per slab (inflate_slabs_to_host); 0 = automatic
  DevBuf unit_ctr;    // k_lz hands its units out from this counter
  // deflate workspace
  DevBuf tokens, ntok, hist, scratch, adler_part, codes, blk_bits, blk_off, summary;
  // inflate workspace
  DevBuf tile_cnt, cand, res, ctl, seg_pos, seg_off, fres, run_first, fstored, fchain, fsym, fwin, pinfo, fjobs, fpieces, fsurv, fmeta, ftab, fmaps, finfo, fitems;
  std::vector<DevBuf> dem_slabs;  // token room of blocks decoded on demand (inflate_foreign)
  bool pinfo_valid = false;  // phase A left the blocks as pieces (k_inf_tokens4) for a piece-parallel phase B
  // adler / misc
  DevBuf acc;
  // staging for the host forms
  DevBuf d_in, d_out, d_off_in, d_off_out, d_len, d_status, d_coff, d_pack;
  u32 batch_slab_streams = 16384;  // host batch calls: buffers per slab at most (BATCH_SLAB_STREAMS)
  u8 *batch_stage = nullptr;  // pinned: compacted results of the host batch calls, double buffered
  size_t batch_stage_cap = 0;
  HostMail *mail = nullptr;  // pinned
  u64 *slab_mail = nullptr;  // pinned: where each slab of a pipelined host-buffer deflate ends (bytes)
  size_t slab_mail_cap = 0;
  StageRing ring_in, ring_out;  // pinned staging of large pageable host buffers (stager.inl)
  u64 *cand_mail = nullptr;  // pinned: block starts read back by scan_block_starts
  size_t cand_mail_cap = 0;
  u64 *slab_cand = nullptr;  // pinned: the block starts of one slab on their way back to the device (inflate_known_starts)
  size_t slab_cand_cap = 0;
  CorpusTable *d_corpus = nullptr;
  CrcTables *d_crc = nullptr;  // CRC-32 tables (gzip), uploaded on first use
  DevBuf crc_part;
  // per-kernel timing
  bool timing = false;
  struct TimedLaunch { const char *name; zrt_event_t e0, e1; };
  std::vector<TimedLaunch> timed;
  std::vector<zrt_event_t> event_pool;
  struct KernelTotal { std::string name; double ms; uint64_t n; };
  std::vector<KernelTotal> totals;
  // state between phase 1 and phase 2 of a deflate
  bool p1_valid = false;
  const u8 *p1_in = nullptr;
  u64 p1_n = 0;
  u32 p1_nblocks = 0, p1_nchunks = 0;
  int p1_final = 0;
  u64 p1_comp = 0;
};

// Per-kernel device timing (zles_ctx_set_timing): an event pair around every launch, read back by
// zles_ctx_kernel_time.  Off by default; bench.py turns it on to get the dominant kernel's duration.
static void timing_begin(zles_ctx *c, const char *name);
static void timing_end(zles_ctx *c);

#define LAUNCH(ctx, kern, grid, block, smem, ...)                        \
  do {                                                                   \
    if ((ctx)->timing) timing_begin((ctx), #kern);                       \
    ZLES_LAUNCH(kern, grid, block, smem, (ctx)->stream, __VA_ARGS__);    \
    if ((ctx)->timing) timing_end((ctx));                                \
    (ctx)->launches++;                                                   \
    if (debug_sync()) debug_check((ctx), #kern);                         \
  } while (0)

// ZLES_DEBUG_SYNC=1 in the environment: synchronise after every launch and name the kernel that faulted
// on stderr (compute-sanitizer is not available on the GPU pool).
static bool debug_sync() {
  static const bool on = [] { const char *e = getenv("ZLES_DEBUG_SYNC"); return e && *e && *e != '0'; }();
  return on;
}
static void debug_check(zles_ctx *c, const char *kern);

// ZLES_TRACE=1 in the environment: host-side timestamps of the host-buffer calls on stderr (where does the wall time go)
#include <chrono>
static bool trace_on() {
  static const bool on = [] { const char *e = getenv("ZLES_TRACE"); return e && *e && *e != '0'; }();
  return on;
}
static double trace_now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TRACE(...) do { if (trace_on()) { fprintf(stderr, "[zles %.3f] ", trace_now()); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static int resolve_ctx(zles_ctx *&c);
struct zles_mgpu;
static zles_mgpu *default_mgpu();  // mgpu.inl: the multi-GPU default installed by zles_init(device_mask), or null

static zrt_event_t timing_event(zles_ctx *c) {
  if (!c->event_pool.empty()) { zrt_event_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
  zrt_event_t e{};
  zrt_event_create(&e);
  return e;
}
static void debug_check(zles_ctx *c, const char *kern) {
  zrt_err_t e = zrt_sync(c->stream);
  if (e == ZRT_OK) e = zrt_last_error();
  if (e != ZRT_OK) fprintf(stderr, "[zles debug] %s: %s\n", kern, zrt_err_str(e));
}
static void timing_begin(zles_ctx *c, const char *name) {
  zles_ctx::TimedLaunch t{name, timing_event(c), timing_event(c)};
  zrt_event_record(t.e0, c->stream);
  c->timed.push_back(t);
}
static void timing_end(zles_ctx *c) { zrt_event_record(c->timed.back().e1, c->stream); }
static void timing_collect(zles_ctx *c) {
  zrt_sync(c->stream);
  for (auto &t : c->timed) {
    float ms = 0;
    zrt_event_elapsed(&ms, t.e0, t.e1);
    bool found = false;
    for (auto &k : c->totals)
      if (k.name == t.name) { k.ms += ms; k.n++; found = true; break; }
    if (!found) c->totals.push_back({t.name, (double)ms, 1});
    c->event_pool.push_back(t.e0);
    c->event_pool.push_back(t.e1);
  }
  c->timed.clear();
}

// ---- library / context ----------------------------------------------------------------

extern "C" const char *zles_version(void) { return ZLES_VERSION_STR; }

extern "C" const char *zles_strerror(int code) {
  switch (code) {
    case ZLES_OK: return "";
    case ZLES_E_NOT_DEFLATE: return "Not compressed by deflate";      // src/zlib.ts:15
    case ZLES_E_BTYPE3: return "Not supported BTYPE : 3";             // src/inflate.ts:32
    case ZLES_E_INSUFFICIENT: return "Data length is insufficient";   // src/inflate.ts:35
    case ZLES_E_CORRUPTED: return "Data is corrupted";                // src/inflate.ts:50,88,166,247,276
    case ZLES_E_LACK: return "Lack of data length";                   // src/utils/BitReadStream.ts:15
    case ZLES_E_OUTPUT_FULL: return "output buffer too small";
    case ZLES_E_CUDA: return "CUDA error";
    case ZLES_E_ARG: return "invalid argument";
    case ZLES_E_NOMEM: return "out of memory";
    case ZLES_E_RUNAWAY: return "stream never ends";
    case ZLES_E_CHECKSUM: return "gzip checksum mismatch";
  }
  return "unknown error";
}

extern "C" const char *zles_last_cuda_error(void) { return g_cuda_err.c_str(); }


static int set_kernel_attrs(int device) {
  // opt in to > 48 KiB dynamic shared memory where needed
  zrt_err_t e = zrt_set_smem(k_lz, (int)LZ_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_lz)");
  e = zrt_set_smem(k_lz_batch, (int)LZ_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_lz_batch)");
  e = zrt_set_smem(k_huff, HUF_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_huff)");
  e = zrt_set_smem(k_huff_merge, HUF_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_huff_merge)");
  e = zrt_set_smem(k_inflate, INF_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_inflate)");
  e = zrt_set_smem(k_inf_resolve, RES_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_inf_resolve)");
  e = zrt_set_smem(k_fblk_map, FB_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_fblk_map)");
  e = zrt_set_smem(k_fblk_prefix, FB_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_fblk_prefix)");
  e = zrt_set_smem(k_fblk_head, TOK_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_fblk_head)");
  e = zrt_set_smem(k_fpiece_sym, SEG_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_fpiece_sym)");
  e = zrt_set_smem(k_piece_sym, SEG_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_piece_sym)");
  e = zrt_set_smem(k_inf_tokens4, SPEC_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_inf_tokens4)");
  e = zrt_prefer_smem(k_inf_tokens);  // nine CTAs of 24,448 + 1,024 bytes per SM need all 228 KiB as shared memory
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_inf_tokens, carveout)");
  e = zrt_set_smem(k_inf_tokens, TOK_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_inf_tokens)");
  e = zrt_set_smem(k_inflate_batch, INF_SMEM);
  if (e != ZRT_OK) return cuda_fail(e, "cudaFuncSetAttribute(k_inflate_batch)");
  (void)device;
  return 0;
}

extern "C" int zles_ctx_create(int device, zles_ctx **out) {
  if (!out || device < 0) return ZLES_E_ARG;
  *out = nullptr;
  CK(zrt_set_device(device));
  zles_ctx *c = new (std::nothrow) zles_ctx();
  if (!c) return ZLES_E_NOMEM;
  c->device = device;
  zrt_err_t e = zrt_stream_create(&c->stream);
  if (e != ZRT_OK) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
  c->own_stream = true;
  e = zrt_stream_create(&c->copy_stream);
  if (e != ZRT_OK) { zrt_stream_destroy(c->stream); delete c; return cuda_fail(e, "cudaStreamCreate"); }
  e = zrt_stream_create(&c->out_stream);
  if (e != ZRT_OK) { zrt_stream_destroy(c->stream); zrt_stream_destroy(c->copy_stream); delete c; return cuda_fail(e, "cudaStreamCreate"); }
  c->sm_count = zrt_sm_count(device);
  { const char *e = getenv("ZLES_NO_MERGE"); if (e && *e && *e != '0') c->merge_chunks = false; }
  { const char *e = getenv("ZLES_NO_DEV_SLABS"); if (e && *e && *e != '0') c->no_dev_slabs = true; }
  void *m = nullptr;
  e = zrt_host_alloc(&m, sizeof(HostMail));
  if (e != ZRT_OK) { zrt_stream_destroy(c->stream); zrt_stream_destroy(c->copy_stream); zrt_stream_destroy(c->out_stream); delete c; return cuda_fail(e, "cudaMallocHost"); }
  c->mail = reinterpret_cast<HostMail *>(m);
  memset(c->mail, 0, sizeof(HostMail));
  int rc = set_kernel_attrs(device);
  if (rc) { zles_ctx_destroy(c); return rc; }
  *out = c;
  return 0;
}

extern "C" void zles_ctx_destroy(zles_ctx *c) {
  if (!c) return;
  zrt_set_device(c->device);
  zrt_sync(c->stream);
  DevBuf *bufs[] = {&c->tokens, &c->ntok,     &c->hist,      &c->scratch, &c->adler_part, &c->codes,  &c->blk_bits, &c->blk_off,
                    &c->summary, &c->tile_cnt, &c->cand,    &c->res,        &c->ctl,    &c->seg_pos,  &c->seg_off, &c->fres, &c->run_first, &c->fstored, &c->fchain, &c->fsym, &c->fwin, &c->pinfo, &c->fjobs, &c->fpieces, &c->fsurv, &c->fmeta, &c->ftab, &c->fmaps, &c->finfo, &c->fitems, &c->unit_ctr,
                    &c->acc,    &c->d_in,     &c->d_out,     &c->d_off_in, &c->d_off_out, &c->d_len,  &c->d_status, &c->d_coff, &c->d_pack};
  for (DevBuf *b : bufs) b->release();
  for (DevBuf &b : c->dem_slabs) b.release();
  if (c->slab_mail) zrt_host_free(c->slab_mail);
  c->slab_mail = nullptr;
  if (c->cand_mail) zrt_host_free(c->cand_mail);
  c->cand_mail = nullptr;
  if (c->slab_cand) zrt_host_free(c->slab_cand);
  c->slab_cand = nullptr;
  if (c->batch_stage) zrt_host_free(c->batch_stage);
  c->batch_stage = nullptr;
  c->ring_in.release();
  c->ring_out.release();
  timing_collect(c);
  for (zrt_event_t e : c->event_pool) zrt_event_destroy(e);
  if (c->d_corpus) zrt_free(c->d_corpus);
  if (c->d_crc) zrt_free(c->d_crc);
  c->crc_part.release();
  if (c->mail) zrt_host_free(c->mail);
  if (c->own_stream) zrt_stream_destroy(c->stream);
  zrt_stream_destroy(c->copy_stream);
  zrt_stream_destroy(c->out_stream);
  delete c;
}

extern "C" int zles_ctx_set_stream(zles_ctx *c, void *cuda_stream) {
  if (!c) return ZLES_E_ARG;
  CK(zrt_sync(c->stream));
  if (c->own_stream) zrt_stream_destroy(c->stream);
  c->stream = (zrt_stream_t)cuda_stream;
  c->own_stream = false;
  return 0;
}

extern "C" int zles_ctx_set_level(zles_ctx *c, uint32_t max_checks, uint32_t min_checks, uint32_t good_len, uint32_t lazy) {
  if (!c || max_checks == 0) return ZLES_E_ARG;
  if (max_checks > LZ_SCAN) max_checks = LZ_SCAN;  // the matcher compares at most LZ_SCAN candidates per position (and relies on it)
  c->max_checks = max_checks;
  c->min_checks = min_checks ? min_checks : 1;
  c->good_len = good_len;
  c->lazy = lazy ? 1 : 0;
  return 0;
}

extern "C" int zles_ctx_set_window_mode(zles_ctx *c, uint32_t mode) {
  if (!c || mode > 1) return ZLES_E_ARG;
  c->pair_mode = mode;
  return 0;
}

extern "C" int zles_ctx_set_slab_blocks(zles_ctx *c, uint32_t blocks) {
  if (!c || (blocks && blocks < SUBS_PER_CHUNK) || blocks > (1u << 20)) return ZLES_E_ARG;
  c->inf_slab_blocks = (blocks / SUBS_PER_CHUNK) * SUBS_PER_CHUNK;
  return 0;
}

extern "C" int zles_ctx_set_batch_slab(zles_ctx *c, uint32_t buffers) {
  if (!c || buffers == 0) return ZLES_E_ARG;
  c->batch_slab_streams = buffers < BATCH_SLAB_STREAMS ? buffers : BATCH_SLAB_STREAMS;
  return 0;
}

extern "C" int zles_ctx_set_stream_min(zles_ctx *c, size_t bytes) {
  if (!c) return ZLES_E_ARG;
  c->inf_stream_min = bytes;
  return 0;
}

extern "C" uint64_t zles_ctx_launches(const zles_ctx *c) { return c ? c->launches : 0; }

extern "C" int zles_ctx_set_timing(zles_ctx *c, int on) {
  if (!c) return ZLES_E_ARG;
  CK(zrt_set_device(c->device));
  timing_collect(c);
  c->totals.clear();
  c->timing = on != 0;
  return 0;
}

extern "C" int zles_ctx_kernel_time(zles_ctx *c, const char *kernel, double *ms_total, uint64_t *launches) {
  if (!c || !kernel || !ms_total || !launches) return ZLES_E_ARG;
  CK(zrt_set_device(c->device));
  timing_collect(c);
  *ms_total = 0;
  *launches = 0;
  for (auto &k : c->totals)
    if (k.name == kernel) { *ms_total = k.ms; *launches = k.n; }
  return 0;
}

extern "C" int zles_dev_alloc(zles_ctx *c, size_t n, void **d_ptr) {
  if (!d_ptr) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  CK(zrt_malloc(d_ptr, n));
  return 0;
}
extern "C" int zles_dev_free(zles_ctx *c, void *d_ptr) {
  RET(resolve_ctx(c));
  if (d_ptr) CK(zrt_free(d_ptr));
  return 0;
}
extern "C" int zles_dev_copy(zles_ctx *c, void *dst, const void *src, size_t n) {
  if ((!dst || !src) && n) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  if (n) CK(zrt_copy(dst, src, n, c->stream));
  CK(zrt_sync(c->stream));
  return 0;
}

extern "C" int zles_dev_copy_async(zles_ctx *c, void *dst, const void *src, size_t n) {
  if ((!dst || !src) && n) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  if (n) CK(zrt_copy(dst, src, n, c->stream));
  return 0;
}

extern "C" int zles_ctx_sync(zles_ctx *c) {
  RET(resolve_ctx(c));
  CK(zrt_sync(c->stream));
  return 0;
}

static std::mutex g_default_mu;
static zles_ctx *g_default_ctx = nullptr;

static int resolve_ctx(zles_ctx *&c) {
  if (c) {
    CK(zrt_set_device(c->device));
    return 0;
  }
  std::lock_guard<std::mutex> lk(g_default_mu);
  if (!g_default_ctx) RET(zles_ctx_create(0, &g_default_ctx));
  c = g_default_ctx;
  CK(zrt_set_device(c->device));
  return 0;
}

extern "C" void zles_free(void *p) { free(p); }

// ---- Adler-32 (K8) -----------------------------------------------------------------------

static int dev_adler32(zles_ctx *c, const u8 *d_in, size_t n, uint32_t *adler) {
  RET(c->acc.reserve(64));
  unsigned long long *acc = c->acc.as<unsigned long long>();
  CK(zrt_memset(acc, 0, 16, c->stream));
  if (n) {
    u64 nvec = (n >> 4) + 1;
    u64 want = (nvec + ADLER_TILE_VECS - 1) / ADLER_TILE_VECS;  // tiles of 32 KiB, handed out grid-stride
    u32 grid = (u32)umin64(want, (u64)c->sm_count * 4);          // four resident CTAs per SM (64 registers x 256 threads)
    if (grid == 0) grid = 1;
    LAUNCH(c, k_adler_partial, grid, ADLER_THREADS, ADLER_SMEM, d_in, (u64)n, acc);
  }
  LAUNCH(c, k_adler_final, 1, 32, 0, acc, (u64)n, reinterpret_cast<u32 *>(acc + 2));
  CK(zrt_last_error());
  CK(zrt_mail(&c->mail->adler, acc + 2, 4, c->stream));
  CK(zrt_sync(c->stream));
  *adler = c->mail->adler;
  return 0;
}

extern "C" int zles_dev_adler32(zles_ctx *c, const uint8_t *d_in, size_t n, uint32_t *adler) {
  if (!adler || (!d_in && n)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  return dev_adler32(c, d_in, n, adler);
}

extern "C" int zles_adler32(zles_ctx *c, const uint8_t *in, size_t n, uint32_t *adler) {
  if (!adler || (!in && n)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  RET(c->d_in.reserve(n + 16));
  if (n) CK(zrt_h2d(c->d_in.p, in, n, c->stream));
  return dev_adler32(c, c->d_in.as<u8>(), n, adler);
}

// ---- deflate (K1..K5) --------------------------------------------------------------------

extern "C" size_t zles_deflate_bound(size_t n) {
  // every 32 KiB block costs at most what it costs stored (5 bytes + its input) plus the 5-byte marker;
  // 6 bytes of zlib framing.  (Rounded up generously: callers size buffers with this.)
  size_t nblocks = (n + SUB - 1) / SUB;
  if (nblocks == 0) nblocks = 1;
  return 6 + n + nblocks * 16 + 64;
}

// h_src != nullptr: d_in is a staging buffer that is filled from host memory slab by slab on the copy
// stream while the matcher already works on earlier slabs (blocks are independent per 128 KiB chunk).
// Host-buffer deflate only: where the packed stream goes (device staging and the caller's buffer).  When the input is
// long enough to be cut into slabs, every slab is laid out, packed and copied to the host while the matcher works on
// the next one; `done` says so, `overflow` that the caller's buffer was too small.
struct DeflatePipe {
  u8 *d_out;      // device staging for the raw deflate bytes (room for zles_deflate_bound)
  u8 *h_out;      // caller's buffer for them
  size_t h_cap;   // its capacity in bytes
  Drainer *drain = nullptr;  // the caller's buffer is pageable: copies to it go through the pinned staging ring
  bool defer = false;  // pack slab by slab into d_out but leave the copy to the host to the caller (multi-GPU: a shard's
                       // place in the stream is only known once every shard before it has been laid out)
  bool dev = false;    // the input is already on the device and d_out is where the stream goes (zles_dev_deflate on a long
                       // input): slabs of DEV_SLAB_BLOCKS blocks, so that the token scratch is a slab's and not the input's
  bool done = false, overflow = false;
};

#ifdef ZLES_EMU
constexpr u32 DEV_SLAB_BLOCKS = 8;      // (the emulator tests run the slab logic on small inputs)
#else
constexpr u32 DEV_SLAB_BLOCKS = 32768;  // 1 GiB of input per slab of a device-resident deflate (4 GiB of token slots)
#endif

static int deflate_phase1(zles_ctx *c, const u8 *d_in, size_t n, int is_last, zles_shard_info *info, const u8 *h_src = nullptr,
                          DeflatePipe *pipe = nullptr) {
  c->p1_valid = false;
  if (!is_last && (n % CHUNK) != 0) return ZLES_E_ARG;
  if (!is_last && n == 0) {  // an empty shard in front of the last one contributes nothing
    c->p1_valid = true;
    c->p1_n = 0;
    c->p1_nblocks = 0;
    c->p1_nchunks = 0;
    c->p1_final = 0;
    c->p1_comp = 0;
    if (info) memset(info, 0, sizeof(*info));
    return 0;
  }
  u64 nb64 = ((u64)n + SUB - 1) / SUB;
  if (nb64 == 0) nb64 = 1;
  if (nb64 > 0x7fffffffull / LZ_NSYM) return ZLES_E_ARG;
  const u32 nblocks = (u32)nb64;
  const u32 nchunks = (nblocks + SUBS_PER_CHUNK - 1) / SUBS_PER_CHUNK;
  const u32 grid_lz = (u32)umin64((u64)nblocks, (u64)c->sm_count);

  RET(c->ntok.reserve((size_t)nblocks * 4));
  RET(c->hist.reserve((size_t)nblocks * LZ_NSYM * 4));
  RET(c->scratch.reserve((size_t)grid_lz * 2 * SUB * 4));
  RET(c->adler_part.reserve((size_t)nblocks * 16));
  RET(c->codes.reserve((size_t)nblocks * sizeof(BlockCodes)));
  RET(c->blk_bits.reserve((size_t)nblocks * 4));
  RET(c->blk_off.reserve(((size_t)nblocks + 1) * 8));
  RET(c->summary.reserve(64));

  LzParams lp;
  lp.in = d_in;
  lp.n = n;
  lp.nblocks = nblocks;
  lp.ntok = c->ntok.as<u32>();
  lp.hist = c->hist.as<u32>();
  lp.scratch = c->scratch.as<u32>();
  lp.adler_part = c->adler_part.as<u64>();
  lp.max_checks = c->max_checks;
  lp.min_checks = c->min_checks;
  lp.good_len = c->good_len;
  lp.lazy = c->lazy;
  lp.pair_mode = c->pair_mode;
  RET(c->unit_ctr.reserve(4));
  lp.unit_ctr = c->unit_ctr.as<u32>();
  // Host input arrives slab by slab: one wave of CTAs first (the matcher starts as soon as 4.6 MiB are on the device),
  // then 2, then 4 waves per slab (a multiple of the SM count keeps the tail of every launch short); whole chunks.
  std::vector<u32> slab_begin;  // first block of every slab, and nblocks at the end
  {
    const u32 wave = ((u32)c->sm_count / SUBS_PER_CHUNK) * SUBS_PER_CHUNK;
    if (!h_src && pipe && pipe->dev && nblocks >= 2 * DEV_SLAB_BLOCKS) {
      for (u32 b = 0; b < nblocks; b += DEV_SLAB_BLOCKS) slab_begin.push_back(b);
      if (nblocks - slab_begin.back() < DEV_SLAB_BLOCKS / 2) slab_begin.pop_back();
      slab_begin.push_back(nblocks);
    } else if (!h_src || wave == 0 || nblocks < 4 * wave) {
      slab_begin = {0, nblocks};
    } else {
      // slabs double from one wave up to an eighth of the input (at least 4, at most 64 waves): a long input gets slabs
      // with dozens of units per CTA — every launch ends with a tail in which the last CTAs finish alone
      u32 cap_sz = (nblocks / 8 / wave) * wave;
      if (cap_sz < 4 * wave) cap_sz = 4 * wave;
      if (cap_sz > 64 * wave) cap_sz = 64 * wave;
      u32 b = 0, sz = wave;
      while (b < nblocks) {
        slab_begin.push_back(b);
        b += sz;
        if (sz < cap_sz) sz = sz * 2 < cap_sz ? sz * 2 : cap_sz;
      }
      if (nblocks - slab_begin.back() < wave && slab_begin.size() > 1) slab_begin.pop_back();  // no sliver at the end
      slab_begin.push_back(nblocks);
    }
  }
  const u32 nslabs = (u32)slab_begin.size() - 1;
  const bool piped = pipe && (h_src || pipe->dev) && nslabs > 1;
  // Token slots: 4 bytes per input byte of whatever is between the matcher and the packer at one time — the whole input
  // when the packer runs afterwards (sharded form: a shard's place in the stream is known only after the exchange), ONE
  // SLAB when every slab is packed before the next is matched (same stream: k_pack of slab s precedes k_lz of slab s + 1).
  // Block b's row is then at (b - slab's first block): the kernels index by absolute block, so they get a shifted base.
  u32 tok_blocks = nblocks;
  if (piped) {
    tok_blocks = 0;
    for (u32 si = 0; si < nslabs; si++) tok_blocks = std::max(tok_blocks, slab_begin[si + 1] - slab_begin[si]);
  }
  RET(c->tokens.reserve((size_t)tok_blocks * SUB * 4));
  // a large pageable source is staged through pinned memory by helper threads (stager.inl), slab by slab
  Feeder feeder;
  bool staged = false;
  if (h_src && n >= STAGE_MIN && host_is_pageable(h_src)) {
    std::vector<StagePiece> pieces;
    for (u32 si = 0; si < nslabs; si++) {
      const size_t off = (size_t)slab_begin[si] * SUB, len = (size_t)umin64((u64)(slab_begin[si + 1] - slab_begin[si]) * SUB, (u64)n - off);
      for (size_t o = 0; o < len; o += STAGE_SLOT)
        pieces.push_back(StagePiece{h_src + off + o, const_cast<u8 *>(d_in) + off + o, std::min<size_t>(STAGE_SLOT, len - o), si});
    }
    staged = feeder.start(c->device, &c->ring_in, c->copy_stream, std::move(pieces), nslabs) == 0;
  }
  std::vector<zrt_event_t> slab_ev;
  if (piped && !pipe->defer) {
    if (c->slab_mail_cap < nslabs) {
      if (c->slab_mail) zrt_host_free(c->slab_mail);
      c->slab_mail = nullptr;
      c->slab_mail_cap = 0;
      CK(zrt_host_alloc(reinterpret_cast<void **>(&c->slab_mail), (size_t)nslabs * 8));
      c->slab_mail_cap = nslabs;
    }
  }
  if (piped) CK(zrt_memset(c->summary.p, 0, 64, c->stream));  // summary[4] carries the running offset between slabs
  for (u32 si = 0; si < nslabs; si++) {
    const u32 b0 = slab_begin[si], b1 = slab_begin[si + 1];
    if (h_src) {
      const size_t off = (size_t)b0 * SUB, len = (size_t)umin64((u64)(b1 - b0) * SUB, (u64)n - off);
      if (staged) {
        if (!feeder.wait_group(si)) return cuda_fail(zrt_last_error(), "staged copy to device");
      } else if (len) {
        CK(zrt_h2d(const_cast<u8 *>(d_in) + off, h_src + off, len, c->copy_stream));
      }
      zrt_event_t ev = timing_event(c);
      CK(zrt_event_record(ev, c->copy_stream));
      CK(zrt_stream_wait_event(c->stream, ev));
      c->event_pool.push_back(ev);  // recorded and waited on in stream order; reusable once this call has synchronised
    }
    lp.first_block = b0;
    lp.nblocks = b1;
    lp.tokens = c->tokens.as<u32>() - (piped ? (size_t)b0 * SUB : 0);
    CK(zrt_memset(lp.unit_ctr, 0, 4, c->stream));
    LAUNCH(c, k_lz, (u32)umin64((u64)(b1 - b0), (u64)c->sm_count), LZ_THREADS, LZ_SMEM, lp);
    LAUNCH(c, k_huff, (b1 - b0 + HUF_WARPS - 1) / HUF_WARPS, HUF_THREADS, HUF_SMEM, (const u32 *)c->hist.as<u32>(), b0, b1,
           c->codes.as<BlockCodes>(), c->blk_bits.as<u32>(), (u64)n, (const BatchBlk *)nullptr);
    if (c->merge_chunks) {  // one block per chunk where four headers cost more than they save (k_huff_merge)
      const u32 nch = (b1 - b0 + SUBS_PER_CHUNK - 1) / SUBS_PER_CHUNK;
      LAUNCH(c, k_huff_merge, (nch + HUF_WARPS - 1) / HUF_WARPS, HUF_THREADS, HUF_SMEM, (const u32 *)c->hist.as<u32>(), b0, b1, nblocks,
             is_last ? 1u : 0u, c->codes.as<BlockCodes>(), c->blk_bits.as<u32>());
    }
    if (piped) {  // this slab's offsets, its bits, and where it ends (for the host, which copies it out below)
      LAUNCH(c, k_layout_slab, 1, 1024, LAYOUT_SMEM, (const u32 *)c->blk_bits.as<u32>(), b0, b1, nblocks, is_last ? 1u : 0u,
             c->summary.as<u64>() + 4, c->blk_off.as<u64>(), c->summary.as<u64>() + 5);
      PackParams pp;
      pp.tokens = c->tokens.as<u32>() - (size_t)b0 * SUB;
      pp.ntok = c->ntok.as<u32>();
      pp.codes = c->codes.as<BlockCodes>();
      pp.blk_bits = c->blk_bits.as<u32>();
      pp.blk_off = c->blk_off.as<u64>();
      pp.nblocks = nblocks;
      pp.first_block = b0;
      pp.last_is_final = is_last ? 1u : 0u;
      pp.out = pipe->d_out;
      pp.in = d_in;
      pp.n = n;
      LAUNCH(c, k_pack, b1 - b0, PACK_THREADS, PACK_SMEM, pp);
      if (!pipe->defer) {
        CK(zrt_mail(c->slab_mail + si, c->summary.as<u64>() + 5, 8, c->stream));
        zrt_event_t ev = timing_event(c);
        CK(zrt_event_record(ev, c->stream));
        slab_ev.push_back(ev);
      }
    }
  }
  if (piped && pipe->defer) pipe->done = true;
  if (piped && !pipe->defer) {
    // the GPU works through the slabs in order; as each one's end offset arrives, its bytes go to the caller's buffer
    u64 prev = 0;
    for (u32 si = 0; si < nslabs; si++) {
      CK(zrt_event_sync(slab_ev[si]));
      const u64 end = c->slab_mail[si];
      if (end > pipe->h_cap) pipe->overflow = true;
      if (!pipe->overflow && end > prev) {
        if (pipe->drain) { if (!pipe->drain->push(pipe->h_out + prev, pipe->d_out + prev, (size_t)(end - prev))) return cuda_fail(zrt_last_error(), "staged copy to host"); }
        else CK(zrt_d2h(pipe->h_out + prev, pipe->d_out + prev, (size_t)(end - prev), c->copy_stream));
      }
      prev = end;
      c->event_pool.push_back(slab_ev[si]);
    }
    if (pipe->drain && !pipe->drain->finish()) return cuda_fail(zrt_last_error(), "staged copy to host");
    zrt_event_t ev = timing_event(c);
    CK(zrt_event_record(ev, c->copy_stream));
    CK(zrt_stream_wait_event(c->stream, ev));
    c->event_pool.push_back(ev);
    pipe->done = true;
  }
  if (staged && !feeder.finish()) return cuda_fail(zrt_last_error(), "staged copy to device");

  LayoutParams yp;
  yp.blk_bits = c->blk_bits.as<u32>();
  yp.nblocks = nblocks;
  yp.last_is_final = is_last ? 1u : 0u;
  yp.n = n;
  yp.adler_part = c->adler_part.as<u64>();
  yp.blk_off = c->blk_off.as<u64>();
  yp.summary = c->summary.as<u64>();
  LAUNCH(c, k_layout, 1, 1024, LAYOUT_SMEM, yp);
  CK(zrt_last_error());

  CK(zrt_mail(c->mail->summary, c->summary.p, 24, c->stream));
  CK(zrt_sync(c->stream));
  c->p1_valid = !piped;  // (slab by slab: the stream has been written, the tokens are gone — there is no phase 2)
  c->p1_in = d_in;
  c->p1_n = n;
  c->p1_nblocks = nblocks;
  c->p1_nchunks = nchunks;
  c->p1_final = is_last;
  c->p1_comp = c->mail->summary[0];
  if (info) {
    info->comp_bytes = c->mail->summary[0];
    info->raw_bytes = n;
    info->adler_a = c->mail->summary[1];
    info->adler_b = c->mail->summary[2];
    info->n_blocks = nblocks;
  }
  return 0;
}

static int deflate_phase2(zles_ctx *c, u8 *d_dst) {
  if (!c->p1_valid) return ZLES_E_ARG;
  if (c->p1_nblocks == 0) return 0;
  PackParams pp;
  pp.tokens = c->tokens.as<u32>();
  pp.ntok = c->ntok.as<u32>();
  pp.codes = c->codes.as<BlockCodes>();
  pp.blk_bits = c->blk_bits.as<u32>();
  pp.blk_off = c->blk_off.as<u64>();
  pp.nblocks = c->p1_nblocks;
  pp.last_is_final = c->p1_final ? 1u : 0u;
  pp.out = d_dst;
  pp.in = c->p1_in;
  pp.n = c->p1_n;
  LAUNCH(c, k_pack, c->p1_nblocks, PACK_THREADS, PACK_SMEM, pp);
  CK(zrt_last_error());
  return 0;
}

extern "C" uint32_t zles_adler32_combine_shards(const zles_shard_info *infos, uint32_t count) {
  // stream d[0..N) = shard 0 | shard 1 | ...; shard s at offset o_s of length m_s with
  // A_s = sum d, B_s = sum (m_s - j) d[j]:  s1 = 1 + sum A_s,  s2 = N + sum (B_s + (N - o_s - m_s) A_s)
  u64 N = 0;
  for (uint32_t i = 0; i < count; i++) N += infos[i].raw_bytes;
  u64 s1 = 1, s2 = N % ADLER_MOD, o = 0;
  for (uint32_t i = 0; i < count; i++) {
    const u64 A = infos[i].adler_a % ADLER_MOD, B = infos[i].adler_b % ADLER_MOD;
    const u64 after = (N - o - infos[i].raw_bytes) % ADLER_MOD;
    s1 = (s1 + A) % ADLER_MOD;
    s2 = (s2 + B + after * A) % ADLER_MOD;
    o += infos[i].raw_bytes;
  }
  return (uint32_t)((s2 << 16) | s1);
}

static void put_zlib_header(u8 *p) {
  // CMF: CM = 8, CINFO = 7; FLG: FCHECK = 28, FDICT = 0, FLEVEL = 2  (src/zlib.ts:28-34) -> 78 9C
  p[0] = 0x78;
  p[1] = 0x9C;
}
static void put_be32(u8 *p, u32 v) {  // src/zlib.ts:37-40
  p[0] = (u8)(v >> 24);
  p[1] = (u8)(v >> 16);
  p[2] = (u8)(v >> 8);
  p[3] = (u8)v;
}

extern "C" int zles_dev_deflate_phase1(zles_ctx *c, const uint8_t *d_in, size_t n, int is_last_shard, zles_shard_info *info) {
  if (!d_in && n) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  return deflate_phase1(c, d_in, n, is_last_shard, info);
}
extern "C" int zles_dev_deflate_block_offsets(zles_ctx *c, const uint64_t **d_offsets) {
  if (!c || !d_offsets || !c->p1_valid) return ZLES_E_ARG;
  *d_offsets = c->blk_off.as<u64>();
  return 0;
}
extern "C" int zles_dev_deflate_phase2(zles_ctx *c, uint8_t *d_dst) {
  if (!c || !d_dst) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  RET(deflate_phase2(c, d_dst));
  CK(zrt_sync(c->stream));
  return 0;
}

extern "C" int zles_dev_deflate(zles_ctx *c, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t cap, size_t *out_len) {
  if ((!d_in && n) || !out_len) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  zles_shard_info info;
  DeflatePipe pipe;  // a long input with room for the worst case: packed slab by slab straight into d_out
  pipe.dev = pipe.defer = true;
  pipe.d_out = d_out + 2;
  const bool slabs = d_out && cap >= zles_deflate_bound(n) && !c->no_dev_slabs;
  RET(deflate_phase1(c, d_in, n, 1, &info, nullptr, slabs ? &pipe : nullptr));
  const size_t need = (size_t)info.comp_bytes + 6;
  *out_len = need;
  if (!d_out || cap < need) return ZLES_E_OUTPUT_FULL;
  if (!pipe.done) RET(deflate_phase2(c, d_out + 2));
  put_zlib_header(c->mail->head);
  put_be32(c->mail->head + 2, zles_adler32_combine_shards(&info, 1));
  CK(zrt_h2d(d_out, c->mail->head, 2, c->stream));
  CK(zrt_h2d(d_out + need - 4, c->mail->head + 2, 4, c->stream));
  CK(zrt_sync(c->stream));
  return 0;
}

// ---- CRC-32 (gzip's checksum; crc32.cuh) --------------------------------------------------------
static int dev_crc32(zles_ctx *c, const u8 *d_in, size_t n, uint32_t *crc) {
  if (!c->d_crc) {
    CrcTables *T = new (std::nothrow) CrcTables();
    if (!T) return ZLES_E_NOMEM;
    crc_tables_build(T);
    void *p = nullptr;
    zrt_err_t e = zrt_malloc(&p, sizeof(CrcTables));
    if (e == ZRT_OK) e = zrt_h2d(p, T, sizeof(CrcTables), c->stream);
    if (e == ZRT_OK) e = zrt_sync(c->stream);
    delete T;
    if (e != ZRT_OK) { if (p) zrt_free(p); return cuda_fail(e, "CRC tables"); }
    c->d_crc = reinterpret_cast<CrcTables *>(p);
  }
  const u32 skew = (u32)((uintptr_t)d_in & 15);
  const u8 *base = d_in - skew;
  const u64 nfull = ((u64)skew + n) / CRC_PER_BLOCK;
  RET(c->crc_part.reserve((size_t)(nfull + 1) * 4 + 16));
  u32 *part = c->crc_part.as<u32>();
  if (nfull) {
    const u32 grid = (u32)umin64(nfull, (u64)c->sm_count * 8);
    LAUNCH(c, k_crc_partial, grid, CRC_THREADS, CRC_SMEM, base, skew, nfull, (const CrcTables *)c->d_crc, part);
  }
  LAUNCH(c, k_crc_final, 1, 1024, CRC_FINAL_SMEM, base, skew, (u64)n, nfull, (const u32 *)part, (const CrcTables *)c->d_crc, part + nfull);
  CK(zrt_last_error());
  CK(zrt_mail(&c->mail->adler, part + nfull, 4, c->stream));
  CK(zrt_sync(c->stream));
  *crc = c->mail->adler;
  return 0;
}

extern "C" int zles_dev_crc32(zles_ctx *c, const uint8_t *d_in, size_t n, uint32_t *crc) {
  if (!crc || (!d_in && n)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  return dev_crc32(c, d_in, n, crc);
}

extern "C" int zles_crc32(zles_ctx *c, const uint8_t *in, size_t n, uint32_t *crc) {
  if (!crc || (!in && n)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  RET(c->d_in.reserve(n + 16));
  if (n) CK(zrt_h2d(c->d_in.p, in, n, c->stream));
  return dev_crc32(c, c->d_in.as<u8>(), n, crc);
}

extern "C" uint32_t zles_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) { return crc_combine(crc_a, crc_b, len_b); }

// ---- containers around the raw deflate data ---------------------------------------------------------
// zlib (RFC 1950): 78 9C | data | Adler-32 big endian — what the reference writes (/root/reference/src/zlib.ts:25-49).
// raw  (RFC 1951): data only — the reference's deflate core (/root/reference/src/deflate.ts:14) and its
//                  inflate(input, offset) (/root/reference/src/inflate.ts:16).
// gzip (RFC 1952): 10-byte header | data | CRC-32 | ISIZE, both little endian — the sibling format.
enum { FMT_ZLIB = 0, FMT_RAW = 1, FMT_GZIP = 2 };
static size_t fmt_head(int fmt) { return fmt == FMT_ZLIB ? 2 : fmt == FMT_GZIP ? 10 : 0; }
static size_t fmt_tail(int fmt) { return fmt == FMT_ZLIB ? 4 : fmt == FMT_GZIP ? 8 : 0; }
static void put_le32(u8 *p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }
static u32 get_le32(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
static void put_gzip_header(u8 *p) {
  // ID1 ID2 CM=8 FLG=0 MTIME=0 XFL=0 OS=255 (unknown)
  static const u8 h[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 255};
  memcpy(p, h, 10);
}

static int deflate_host(zles_ctx *c, int fmt, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  const size_t H = fmt_head(fmt), T = fmt_tail(fmt);
  RET(resolve_ctx(c));
  RET(c->d_in.reserve(n + 16));
  zles_shard_info info;
  DeflatePipe pipe;
  DeflatePipe *pp = nullptr;
  Drainer drain;
  if (out && cap >= H + T && c->d_out.reserve(zles_deflate_bound(n) + 16) == 0) {  // raw deflate bytes go to out + H .. cap - T
    pipe.d_out = c->d_out.as<u8>() + 2;
    pipe.h_out = out + H;
    pipe.h_cap = cap - H - T;
    if (n >= STAGE_MIN && host_is_pageable(out) && drain.start(c->device, &c->ring_out, c->out_stream) == 0) pipe.drain = &drain;
    pp = &pipe;
  }
  // the host-to-device copy is pipelined with the matcher, and (long inputs) packing and the copy back with it too
  RET(deflate_phase1(c, c->d_in.as<u8>(), n, 1, &info, in, pp));
  const size_t need = (size_t)info.comp_bytes + H + T;
  *out_len = need;
  if (!out || cap < need) return ZLES_E_OUTPUT_FULL;
  if (!pipe.done || pipe.overflow) {
    RET(c->d_out.reserve(need + 16));
    RET(deflate_phase2(c, c->d_out.as<u8>() + 2));
    CK(zrt_d2h(out + H, c->d_out.as<u8>() + 2, info.comp_bytes, c->stream));
  }
  if (fmt == FMT_ZLIB) {
    put_zlib_header(out);
    put_be32(out + need - 4, zles_adler32_combine_shards(&info, 1));
  } else if (fmt == FMT_GZIP) {
    u32 crc = 0;
    RET(dev_crc32(c, c->d_in.as<u8>(), n, &crc));  // the input is in device memory: its CRC-32 costs no extra copy
    put_gzip_header(out);
    put_le32(out + need - 8, crc);
    put_le32(out + need - 4, (u32)n);  // ISIZE = length mod 2^32
  }
  CK(zrt_sync(c->stream));
  return 0;
}

extern "C" int zles_deflate(zles_ctx *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len) return ZLES_E_ARG;
  if (!c)
    if (zles_mgpu *m = default_mgpu()) return zles_mgpu_deflate(m, in, n, out, cap, out_len);
  return deflate_host(c, FMT_ZLIB, in, n, out, cap, out_len);
}
extern "C" int zles_deflate_raw(zles_ctx *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len) return ZLES_E_ARG;
  return deflate_host(c, FMT_RAW, in, n, out, cap, out_len);
}
extern "C" int zles_gzip_deflate(zles_ctx *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len) return ZLES_E_ARG;
  return deflate_host(c, FMT_GZIP, in, n, out, cap, out_len);
}

// ---- inflate (K6/K7) ---------------------------------------------------------------------

static int seg_status_to_code(u32 st) {
  switch (st) {
    case SEG_E_BTYPE3: return ZLES_E_BTYPE3;
    case SEG_E_INSUFF: return ZLES_E_INSUFFICIENT;
    case SEG_E_CORRUPT: return ZLES_E_CORRUPTED;
    case SEG_E_LACK: return ZLES_E_LACK;
    case SEG_E_RUNAWAY: return ZLES_E_RUNAWAY;
  }
  return ZLES_E_CORRUPTED;
}

// control block in device memory
struct InfCtl {
  u32 ncand;
  u32 counter;
  u32 ok;        // problem bits of k_inf_check
  u32 ok_res;    // problem bits of k_inf_resolve
  unsigned long long total;
};

static u32 inflate_grid(const zles_ctx *c, u64 nseg) {
  u64 want = (nseg + INF_WARPS - 1) / INF_WARPS;
  u64 cap = (u64)c->sm_count * 16;
  if (want < 1) want = 1;
  return (u32)(want < cap ? want : cap);
}

// Step 1 of inflate: candidate segment starts = `first` and every position that follows a
// 00 00 FF FF marker.  Leaves them in c->cand (device) and returns their number; *cand_cap is the
// list capacity (more candidates than that: not one of our streams).
static int inflate_scan(zles_ctx *c, const u8 *d_in, size_t n, u64 first, u32 *ncand_all, u32 *cand_cap_out) {
  const u64 nvec = ((u64)n + 15) >> 4;
  const u32 ntiles = (u32)((nvec + MARK_THREADS - 1) / MARK_THREADS);
  RET(c->ctl.reserve(sizeof(InfCtl)));
  InfCtl *ctl = c->ctl.as<InfCtl>();
  RET(c->tile_cnt.reserve(((size_t)ntiles + 1) * 4));
  // Our own streams have one marker per 32 KiB of input (>= ~40 B of stream even for zeros).  Streams
  // with more candidates than n / 32 are somebody else's and take the sequential path.
  const u64 cand_cap64 = (u64)n / 32 + 64;
  if (cand_cap64 > 0x7fffffffull) return ZLES_E_ARG;
  const u32 cand_cap = (u32)cand_cap64;
  RET(c->cand.reserve((size_t)cand_cap * 8));
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  if (ntiles) {
    LAUNCH(c, k_mark_count, ntiles, MARK_THREADS, 64 * 4, d_in, (u64)n, first, c->tile_cnt.as<u32>());
    LAUNCH(c, k_mark_scan, 1, 1024, 64 * 4, c->tile_cnt.as<u32>(), ntiles, &ctl->ncand);
    LAUNCH(c, k_mark_emit, ntiles, MARK_THREADS, 64 * 4, d_in, (u64)n, first, (const u32 *)c->tile_cnt.as<u32>(), c->cand.as<u64>(),
           cand_cap);
  } else {
    LAUNCH(c, k_mark_none, 1, 32, 0, first, c->cand.as<u64>(), &ctl->ncand);
  }
  CK(zrt_last_error());
  CK(zrt_mail(&c->mail->ncand, &ctl->ncand, 4, c->stream));
  CK(zrt_sync(c->stream));
  *ncand_all = c->mail->ncand;
  *cand_cap_out = cand_cap;
  return 0;
}

static int read_ctl(zles_ctx *c, InfCtl *h) {
  CK(zrt_mail(&c->mail->summary[0], c->ctl.p, sizeof(InfCtl), c->stream));
  CK(zrt_sync(c->stream));
  memcpy(h, &c->mail->summary[0], sizeof(InfCtl));
  return 0;
}

// The four kernels that decode jobs [job0, job0 + njobs) of c->fjobs (host copy: hjobs[0 .. njobs)) into tokens
// (inflate_fblk.cuh).  Results in c->fres / c->fpieces.
static int fblk_decode(zles_ctx *c, const u8 *d_in, size_t n, const FbJob *hjobs, u32 njobs, u32 job0) {
  InfCtl *ctl = c->ctl.as<InfCtl>();
  const FbJob *jobs = c->fjobs.as<FbJob>();
  std::vector<FbItem> items;
  for (u32 i = 0; i < njobs; i++)
    for (u32 p0 = 0; p0 < hjobs[i].np; p0 += FB_WARPS) items.push_back(FbItem{i, p0});
  const u32 nitems = (u32)items.size();
  if (nitems == 0) return 0;
  RET(c->fitems.reserve((size_t)nitems * sizeof(FbItem)));
  const u32 grid = nitems < 2 * (u32)c->sm_count ? nitems : 2 * (u32)c->sm_count;  // two CTAs per SM (shared memory, registers)
  CK(zrt_h2d(c->fitems.p, items.data(), (size_t)nitems * sizeof(FbItem), c->stream));
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  LAUNCH(c, k_fblk_head, (njobs + INF_WARPS - 1) / INF_WARPS, INF_THREADS, TOK_SMEM, d_in, (u64)n, jobs, njobs, job0, c->fmeta.as<FbMeta>(),
         c->ftab.as<TokWarpSmem>());
  LAUNCH(c, k_fblk_map, grid, FB_THREADS, FB_SMEM, d_in, (u64)n, jobs, (const FbItem *)c->fitems.as<FbItem>(), nitems, job0,
         (const FbMeta *)c->fmeta.as<FbMeta>(), (const TokWarpSmem *)c->ftab.as<TokWarpSmem>(), c->fmaps.as<u32>(), c->finfo.as<TaPiece>(), &ctl->counter);
  LAUNCH(c, k_fblk_chain, (njobs + 3) / 4, 128, 0, jobs, njobs, job0, (const FbMeta *)c->fmeta.as<FbMeta>(), (const u32 *)c->fmaps.as<u32>(),
         c->finfo.as<TaPiece>(), c->fres.as<FbRes>(), c->fpieces.as<FbPiece>(), (u64)n);
  LAUNCH(c, k_fblk_prefix, grid, FB_THREADS, FB_SMEM, d_in, (u64)n, jobs, (const FbItem *)c->fitems.as<FbItem>(), nitems, job0,
         (const FbMeta *)c->fmeta.as<FbMeta>(), (const TokWarpSmem *)c->ftab.as<TokWarpSmem>(), (const TaPiece *)c->finfo.as<TaPiece>(),
         c->fres.as<FbRes>(), c->fpieces.as<FbPiece>(), &ctl->ok);
  CK(zrt_last_error());
  CK(zrt_sync(c->stream));  // `items` is host heap memory (and the caller reads the results next)
  return 0;
}

// Returns 0 (decoded), a positive status, or -1 when the stream is not something this path handles.
constexpr u32 FB_DEMAND_MAX = 16384;               // blocks decoded on demand per stream at most (a few launches and a read-back each)
constexpr u64 FB_DEMAND_SPAN = 8ull << 20;         // hint for a block decoded on demand: 1 MiB of stream at most (all its pieces at once)
constexpr size_t FB_DEM_SLAB = (size_t)16 << 20;   // tokens per slab of on-demand token room (64 MiB)
static int inflate_foreign(zles_ctx *c, const u8 *d_in, size_t n, u64 first, u8 *d_out, size_t cap, size_t *out_len) {
  if (n < first + 1 || (u64)n >= (1ull << 40)) return -1;
  InfCtl *ctl = c->ctl.as<InfCtl>();
  const u64 nbits = (u64)n * 8;
  const u64 dcap64 = (u64)n / 64 + 1024;   // a dynamic block is rarely shorter than 64 bytes
  const u64 scap64 = (u64)n / 16 + 4096;   // stored-block candidates: any LEN / ~LEN pair in the data looks like one
  const u64 vcap64 = (u64)n / 8 + 4096;    // survivors of the filter: about a bit position in a thousand
  const u32 dcap = (u32)dcap64, scap = (u32)scap64, vcap = (u32)vcap64;
  if (c->cand.reserve((size_t)dcap * 8) || c->fstored.reserve((size_t)scap * sizeof(FbStored)) || c->fsurv.reserve((size_t)vcap * 8)) return -1;
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  {
    const u64 want = ((u64)n + HS_THREADS - 1) / HS_THREADS;
    const u32 grid = (u32)umin64(want, (u64)c->sm_count * 32);
    LAUNCH(c, k_hdr_filter, grid, HS_THREADS, HS_SMEM, d_in, (u64)n, first * 8, c->fsurv.as<u64>(), vcap, c->fstored.as<FbStored>(), scap, &ctl->ncand);
    LAUNCH(c, k_hdr_verify, (u32)c->sm_count * 4, 128, 0, d_in, (u64)n, (const u64 *)c->fsurv.as<u64>(), vcap, c->cand.as<u64>(), dcap, &ctl->ncand);
  }
  CK(zrt_last_error());
  InfCtl h;
  RET(read_ctl(c, &h));
  const u32 ncand = h.ncand, nst = h.counter;  // the scan counts into (ncand, counter, ok)
  if (ncand > dcap || nst > scap || h.ok > vcap) return -1;
  std::vector<u64> cand(ncand);
  std::vector<FbStored> stv(nst);
  if (ncand) CK(zrt_d2h(cand.data(), c->cand.p, (size_t)ncand * 8, c->stream));
  if (nst) CK(zrt_d2h(stv.data(), c->fstored.p, (size_t)nst * sizeof(FbStored), c->stream));
  CK(zrt_sync(c->stream));
  std::sort(cand.begin(), cand.end());
  std::sort(stv.begin(), stv.end(), [](const FbStored &a, const FbStored &b) { return a.bit < b.bit; });
  // jobs: one per dynamic candidate — the next candidate's start is the hint of where the block ends — and room for the
  // blocks decoded on demand; pieces: theirs, and as many again plus a piece per tile of the stream for those
  std::vector<FbJob> jobs(ncand);
  u64 tok_total = 0, npieces = 0;
  for (u32 i = 0; i < ncand; i++) {
    const u64 hint = i + 1 < ncand ? cand[i + 1] : nbits;
    FbJob &J = jobs[i];
    J.bit = cand[i];
    J.hint_end = hint;
    J.tok_cap = (u32)fb_tok_room(hint - cand[i]);
    J.flags = 0;
    J.piece0 = (u32)npieces;
    J.np = fb_planned_pieces(hint - cand[i]);
    J.aux = i;
    J.pad = 0;
    J.tok = reinterpret_cast<u32 *>(tok_total);  // offset for now
    tok_total += J.tok_cap;
    npieces += J.np;
  }
  const u64 pieces_cap = npieces * 2 + nbits / TA_TILE + 4096;
  const size_t njob_cap = (size_t)ncand + FB_DEMAND_MAX, naux = (size_t)ncand + 1;
  if (pieces_cap > 0x7fffffffull) return -1;
  if (c->fjobs.reserve(njob_cap * sizeof(FbJob)) || c->fres.reserve(njob_cap * sizeof(FbRes)) || c->fpieces.reserve((size_t)pieces_cap * 2 * sizeof(FbPiece)) ||
      c->fmaps.reserve((size_t)pieces_cap * TA_ENT * 4) || c->finfo.reserve((size_t)pieces_cap * sizeof(TaPiece)) ||
      c->fmeta.reserve(naux * sizeof(FbMeta)) || c->ftab.reserve(naux * sizeof(TokWarpSmem)) || c->tokens.reserve((size_t)tok_total * 4 + 256))
    return -1;
  std::vector<FbRes> res(ncand);
  if (ncand) {
    for (u32 i = 0; i < ncand; i++) jobs[i].tok = c->tokens.as<u32>() + reinterpret_cast<u64>(jobs[i].tok);
    CK(zrt_h2d(c->fjobs.p, jobs.data(), (size_t)ncand * sizeof(FbJob), c->stream));
    RET(fblk_decode(c, d_in, n, jobs.data(), ncand, 0u));
    CK(zrt_d2h(res.data(), c->fres.p, (size_t)ncand * sizeof(FbRes), c->stream));
    CK(zrt_sync(c->stream));
  }
  // Blocks the scan did not find (fixed blocks, dynamic blocks with an incomplete code) are decoded when the chain
  // reaches them; a block that turns out longer than the room its hint gave it (FB_LONG: a false candidate inside it, or
  // a hint that was a guess) is decoded on from where that room ended, with four times the room.
  u32 ndem = 0;
  size_t dem_slab = 0, dem_used = 0;
  auto demand = [&](u64 pos, u64 hint, u32 flags, u32 aux, FbRes &r) -> int {
    if (ndem >= FB_DEMAND_MAX) return -1;
    const u64 capt = fb_tok_room(hint - pos);
    const u32 np = fb_planned_pieces(hint - pos);
    if (capt > 0xffffffe0ull || npieces + np > pieces_cap) return -1;
    // room: the current slab, or the next one
    for (;; dem_slab++, dem_used = 0) {
      if (dem_slab >= c->dem_slabs.size()) c->dem_slabs.emplace_back();
      DevBuf &sl = c->dem_slabs[dem_slab];
      if (sl.cap == 0 && sl.reserve((FB_DEM_SLAB > capt ? FB_DEM_SLAB : (size_t)capt) * 4)) return -1;
      if ((dem_used + capt) * 4 <= sl.cap) break;
    }
    FbJob J;
    J.bit = pos; J.hint_end = hint; J.tok = c->dem_slabs[dem_slab].as<u32>() + dem_used; J.tok_cap = (u32)capt; J.flags = FB_JOB_COMPACT | flags;
    J.piece0 = (u32)npieces; J.np = np; J.aux = aux; J.pad = 0;
    const u32 ji = ncand + ndem;
    CK(zrt_h2d(c->fjobs.as<FbJob>() + ji, &J, sizeof(FbJob), c->stream));
    RET(fblk_decode(c, d_in, n, &J, 1u, ji));
    LAUNCH(c, k_fblk_compact, 1, FB_THREADS, 0, (const FbJob *)c->fjobs.as<FbJob>(), ji, (const FbRes *)c->fres.as<FbRes>(), c->fpieces.as<FbPiece>());
    CK(zrt_last_error());
    CK(zrt_mail(&c->mail->fres0[0], c->fres.as<FbRes>() + ji, sizeof(FbRes), c->stream));
    CK(zrt_sync(c->stream));  // also: J is on the stack
    memcpy(&r, &c->mail->fres0[0], sizeof(FbRes));
    ndem++;
    if (r.status != FB_OK && r.status != FB_LONG) return -1;
    dem_used += ((size_t)r.ntok + 31) & ~(size_t)31;
    npieces += (r.npieces + 1) / 2;  // the pieces the block really has keep their slots
    return 0;
  };
  // chain walk: the block after one that ends at bit e starts at bit e
  std::vector<FbEnt> chain;
  u64 pos = first * 8, total = 0, nwarps = 0, nflat = 0;
  auto push = [&](FbEnt en) {
    en.out_off = total;
    en.warp0 = (u32)nwarps;
    en.slot0 = (u32)nflat;
    nwarps += en.stored ? FB_STORED_WARPS : en.nslots;
    nflat += en.nslots;
    total += en.len;
    chain.push_back(en);
  };
  for (;;) {
    if (chain.size() > (size_t)ncand + nst + 2 * (size_t)FB_DEMAND_MAX) return -1;
    if (pos + 3 > nbits) return -1;
    const auto it = std::lower_bound(cand.begin(), cand.end(), pos);
    const auto is = std::lower_bound(stv.begin(), stv.end(), pos, [](const FbStored &a, u64 p) { return a.bit < p; });
    const bool is_cand = it != cand.end() && *it == pos && res[(size_t)(it - cand.begin())].status != 0;
    u32 bfinal;
    u64 next;
    FbEnt en;
    memset(&en, 0, sizeof(en));
    if (!is_cand && is != stv.end() && is->bit == pos) {
      const u64 q = (pos + 3 + 7) >> 3;
      en.src = q + 4;
      en.len = is->len;
      en.stored = 1;
      en.nslots = 1;
      bfinal = is->bfinal;
      next = (q + 4 + is->len) * 8;
      push(en);
    } else {
      FbRes r;
      u32 aux;
      u64 span;
      if (is_cand) {
        const size_t ci = (size_t)(it - cand.begin());
        r = res[ci];
        en.job = (u32)ci;
        aux = (u32)ci;
        span = jobs[ci].hint_end - pos;
      } else {
        // not a block this path decodes (BTYPE 3, a damaged stored block, garbage)?  Then the sequential path.
        const auto nx = std::upper_bound(cand.begin(), cand.end(), pos);
        u64 hint = nx != cand.end() ? *nx : nbits;
        if (hint > pos + FB_DEMAND_SPAN) hint = pos + FB_DEMAND_SPAN;
        span = hint - pos;
        aux = ncand;
        const int rc = demand(pos, hint, 0, aux, r);
        if (rc) return rc;
        en.job = ncand + ndem - 1;
      }
      for (;;) {
        en.len = r.out_len;
        en.nslots = r.npieces;
        push(en);
        if (r.status == FB_OK) break;
        // FB_LONG: on from where the room ended
        const u64 from = r.end_bit;
        if (from <= pos || from >= nbits) return -1;
        span *= 4;
        u64 hint = from + span;
        if (hint > nbits) hint = nbits;
        const int rc = demand(from, hint, FB_JOB_CONT, aux, r);
        if (rc) return rc;
        en.job = ncand + ndem - 1;
        pos = from;
      }
      bfinal = r.bfinal;
      next = r.end_bit;
    }
    if (total >= (1ull << 46) || nwarps >= 0x7fff0000ull || nflat >= 0x7fff0000ull) return -1;
    if (bfinal) break;
    if (next <= pos) return -1;
    pos = next;
  }
  *out_len = (size_t)total;
  if (total > cap) return ZLES_E_OUTPUT_FULL;
  const u32 nblk = (u32)chain.size();
  // Runs: about `target` bytes of output each, cut at token runs — a long block (a run of fixed blocks decoded as one can
  // be the whole stream) is cut in proportion to its token runs; where exactly a run starts is worked out on the device
  // (k_frun_offsets).  Long streams take longer runs so that the serial window propagation stays short.
  std::vector<u32> run_first;
  {
    u64 target = total / 1024;
    if (target < SYM_RUN) target = SYM_RUN;
    u64 acc = 0;
    for (u32 i = 0; i < nblk; i++) {
      const FbEnt &en = chain[i];
      if (i == 0 || acc >= target) { run_first.push_back(en.slot0); acc = 0; }
      u64 k = en.len / target;  // cut this entry into k parts?
      if (k > en.nslots) k = en.nslots;
      if (k >= 2) {
        for (u64 q = 1; q < k; q++) {
          const u32 f = en.slot0 + (u32)(q * en.nslots / k);
          if (f > run_first.back()) run_first.push_back(f);
        }
        acc = en.len / k;
      } else {
        acc += en.len;
      }
    }
  }
  const u32 nruns = (u32)run_first.size();
  run_first.push_back((u32)nflat);
  if (c->fsym.reserve((size_t)total * 2 + 64) || c->fwin.reserve((size_t)nruns * SYM_WIN)) return -1;
  RET(c->fchain.reserve((size_t)nblk * sizeof(FbEnt)));
  RET(c->run_first.reserve((size_t)(nruns + 1) * 4));
  RET(c->seg_off.reserve((size_t)(nruns + 1) * 8));
  CK(zrt_h2d(c->fchain.p, chain.data(), (size_t)nblk * sizeof(FbEnt), c->stream));
  CK(zrt_h2d(c->run_first.p, run_first.data(), (size_t)(nruns + 1) * 4, c->stream));
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  const FbEnt *d_ents = c->fchain.as<FbEnt>();
  LAUNCH(c, k_fslot_scan, (nblk + 3) / 4, 128, 0, (const FbJob *)c->fjobs.as<FbJob>(), c->fpieces.as<FbPiece>(), d_ents, nblk);
  LAUNCH(c, k_frun_offsets, (nruns + 1 + 127) / 128, 128, 0, (const FbJob *)c->fjobs.as<FbJob>(), (const FbPiece *)c->fpieces.as<FbPiece>(), d_ents, nblk,
         (const u32 *)c->run_first.as<u32>(), nruns, (u64)total, c->seg_off.as<u64>());
  LAUNCH(c, k_fpiece_sym, (u32)((nwarps + RES_WARPS - 1) / RES_WARPS), RES_THREADS, SEG_SMEM, (const FbJob *)c->fjobs.as<FbJob>(),
         (const FbPiece *)c->fpieces.as<FbPiece>(), d_ents, nblk, (u32)nwarps, d_in, c->fsym.as<u16>());
  LAUNCH(c, k_frun_merge, nruns < (u32)c->sm_count * 8 ? nruns : (u32)c->sm_count * 8, MRG_THREADS, 0, (const FbJob *)c->fjobs.as<FbJob>(),
         (const FbPiece *)c->fpieces.as<FbPiece>(), d_ents, nblk, (const u32 *)c->run_first.as<u32>(), (const u64 *)c->seg_off.as<u64>(), nruns,
         c->fsym.as<u16>(), &ctl->ok_res);
  LAUNCH(c, k_win_propagate, 1, 1024, 0, (const u16 *)c->fsym.as<u16>(), (const u64 *)c->seg_off.as<u64>(), nruns, c->fwin.as<u8>(),
         (const u32 *)&ctl->ok_res);
  {
    dim3 grid(16, nruns < 4096 ? nruns : 4096);
    LAUNCH(c, k_sym_finalize, grid, 256, 0, (const u16 *)c->fsym.as<u16>(), (const u64 *)c->seg_off.as<u64>(), nruns,
           (const u8 *)c->fwin.as<u8>(), d_out);
  }
  CK(zrt_last_error());
  CK(zrt_sync(c->stream));  // the vectors are host heap memory
  return 0;
}

// Phase B of our own streams: the copies of nseg segments (seg_list, or candidates 0..nseg-1 when null) into d_out;
// problem bits go to ctl->ok_res.  One warp per piece of a block (phase A leaves up to four; one when it ran one warp per
// block) into 16-bit symbols, then the pieces of each chunk made concrete in order (k_piece_sym + k_chunk_final) —
// sixteen times as many warps as one warp per 128 KiB chunk (k_inf_resolve, kept as the fallback when the symbol buffer
// cannot be allocated), for 2 bytes of scratch per output byte of a group.
constexpr u32 SYM_GROUP_SEGS = 8192;  // blocks per launch pair of the piece-parallel phase B (512 MiB of 16-bit symbols)
constexpr u32 SPEC_MAX_SEGS = 4096;   // phase A with four warps per block up to 128 MiB of output; beyond, one warp per block fills the GPU
static int launch_phase_b(zles_ctx *c, const u32 *d_seg_list, u32 nseg, const u8 *d_in, u8 *d_out, size_t cap) {
  InfCtl *ctl = c->ctl.as<InfCtl>();
  const u32 nchunks = (nseg + SUBS_PER_CHUNK - 1) / SUBS_PER_CHUNK;
  static const bool no_sym = [] { const char *e = getenv("ZLES_NO_SYM"); return e && *e && *e != '0'; }();  // debugging aid
  const u32 group = nseg < SYM_GROUP_SEGS ? nseg : SYM_GROUP_SEGS;
  if (!no_sym && c->fsym.reserve((size_t)group * SUB * 2) == 0) {
    // one warp per piece into 16-bit symbols, then every chunk's pieces made concrete in order; long runs go through the
    // symbol buffer a group of whole chunks at a time (measured: 59 ms for 8 GiB, against 82 ms with one warp per chunk)
    const u32 *pinfo = c->pinfo_valid ? (const u32 *)c->pinfo.as<u32>() : nullptr;
    for (u32 seg0 = 0; seg0 < nseg; seg0 += group) {
      const u32 cnt = nseg - seg0 < group ? nseg - seg0 : group;
      const u32 nwarps = cnt * SEG_PIECES;
      LAUNCH(c, k_piece_sym, (nwarps + RES_WARPS - 1) / RES_WARPS, RES_THREADS, SEG_SMEM, (const u32 *)c->tokens.as<u32>(),
             (const u32 *)c->ntok.as<u32>(), pinfo, d_seg_list, seg0 + cnt, seg0, d_in, (const InfRes *)c->res.as<InfRes>(), c->fsym.as<u16>(), d_out,
             (u64)cap, &ctl->ok_res);
      LAUNCH(c, k_chunk_final, (cnt + SUBS_PER_CHUNK - 1) / SUBS_PER_CHUNK, FIN_THREADS, 0, (const u16 *)c->fsym.as<u16>(),
             (const u32 *)c->ntok.as<u32>(), pinfo, d_seg_list, seg0 + cnt, seg0, (const InfRes *)c->res.as<InfRes>(), d_out, (u64)cap, &ctl->ok_res);
    }
    return 0;
  }
  LAUNCH(c, k_inf_resolve, (nchunks + RES_WARPS - 1) / RES_WARPS, RES_THREADS, RES_SMEM, (const u32 *)c->tokens.as<u32>(),
         (const u32 *)c->ntok.as<u32>(), d_seg_list, nseg, d_in, (const InfRes *)c->res.as<InfRes>(), d_out, (u64)cap, &ctl->ok_res);
  return 0;
}

// Steps 2.. of inflate.  On success *out_len = decoded size.  ZLES_E_OUTPUT_FULL: *out_len = size needed.
// ours_only: give up (ZLES_E_CORRUPTED) instead of trying the other tiers when the bytes are not a run of our own blocks.
static int inflate_decode(zles_ctx *c, const u8 *d_in, size_t n, u64 first, u32 ncand_all, u32 cand_cap, u8 *d_out, size_t cap,
                          size_t *out_len, bool has_final = true, bool ours_only = false) {
  InfCtl *ctl = c->ctl.as<InfCtl>();
  bool fast = ncand_all >= 1 && ncand_all <= cand_cap;
  // one of our blocks takes at most SUB + 10 bytes of stream (stored): a stream with fewer markers than that is somebody
  // else's, and decoding its first block on one warp only to find that out costs more than the whole block-parallel tier
  if ((u64)n > (u64)ncand_all * (SUB + 4096)) fast = false;
  const u32 ncand = ncand_all;
  if (fast) {  // workspace for phase A; failing to get it only costs the fast path
    if (c->tokens.reserve((size_t)ncand * SUB * 4) || c->ntok.reserve((size_t)ncand * 4) || c->res.reserve((size_t)ncand * sizeof(InfRes)) ||
        c->pinfo.reserve((size_t)ncand * 2 * SEG_PIECES * 4))
      fast = false;
  }
  if (fast) {
    // 2. phase A on every candidate, acceptance check, and — optimistically — phase B with candidate j
    //    taken as block j of the stream (true unless a marker pattern occurs inside compressed data)
    // few blocks: four warps per block, the quarters decoded speculatively (inflate_spec.cuh); many: one warp per block
    static const bool no_spec = [] { const char *e = getenv("ZLES_NO_SPEC"); return e && *e && *e != '0'; }();  // debugging aid
    if (ncand <= SPEC_MAX_SEGS && !no_spec) {
      const u32 grid = ncand < (u32)c->sm_count * 16 ? ncand : (u32)c->sm_count * 16;
      LAUNCH(c, k_inf_tokens4, grid, SPEC_THREADS, SPEC_SMEM, d_in, (u64)n, (const u64 *)c->cand.as<u64>(), ncand, c->tokens.as<u32>(),
             c->ntok.as<u32>(), c->res.as<InfRes>(), c->pinfo.as<u32>(), &ctl->counter);
      c->pinfo_valid = true;
    } else {
      c->pinfo_valid = true;  // the one-warp decoder cuts every block into four pieces too
      LAUNCH(c, k_inf_tokens, inflate_grid(c, ncand), INF_THREADS, TOK_SMEM, d_in, (u64)n, (const u64 *)c->cand.as<u64>(), ncand,
             c->tokens.as<u32>(), c->ntok.as<u32>(), c->res.as<InfRes>(), c->pinfo.as<u32>(), &ctl->counter);
    }
    LAUNCH(c, k_inf_check, (ncand + 255) / 256, 256, 0, (const InfRes *)c->res.as<InfRes>(), (const u64 *)c->cand.as<u64>(), ncand,
           (u64)n, has_final ? 1u : 0u, &ctl->ok, &ctl->total);
    const bool room = (u64)(ncand - 1) * SUB < (u64)cap + 1;  // otherwise the result cannot fit: size query only
    if (room) RET(launch_phase_b(c, nullptr, ncand, d_in, d_out, cap));
    CK(zrt_last_error());
    InfCtl h;
    RET(read_ctl(c, &h));
    if (h.ok == 0) {
      *out_len = (size_t)h.total;
      if (!room || h.total > cap || h.ok_res == 2) return ZLES_E_OUTPUT_FULL;
      if (h.ok_res == 0) return 0;
      // a reference before the start of a chunk: not ours after all -> sequential path
    } else {
      // 3. some candidate is not a block start.  Where a segment ends and how much it stands for do not
      //    depend on the others, so walk the chain on the host: candidate 0 is real, the next real one
      //    starts where it ended.
      std::vector<InfRes> res(ncand);
      std::vector<u64> cand(ncand);
      CK(zrt_d2h(res.data(), c->res.p, (size_t)ncand * sizeof(InfRes), c->stream));
      CK(zrt_d2h(cand.data(), c->cand.p, (size_t)ncand * 8, c->stream));
      CK(zrt_sync(c->stream));
      std::vector<u32> list;
      u64 total = 0;
      bool chain_ok = true;
      for (size_t j = 0;;) {
        const InfRes &r = res[j];
        if (r.status != SEG_SYNC && r.status != SEG_FINAL) { chain_ok = false; break; }
        if (r.status == SEG_SYNC && r.out_len != SUB) { chain_ok = false; break; }
        list.push_back((u32)j);
        total += r.out_len;
        if (r.status == SEG_FINAL) { if (!has_final) chain_ok = false; break; }
        if (!has_final && r.end_pos == n) break;
        size_t lo = j + 1, hi = ncand;
        while (lo < hi) {
          size_t mid = (lo + hi) >> 1;
          if (cand[mid] < r.end_pos) lo = mid + 1; else hi = mid;
        }
        if (lo >= ncand || cand[lo] != r.end_pos) { chain_ok = false; break; }
        j = lo;
      }
      if (chain_ok) {
        *out_len = (size_t)total;
        if (total > cap) return ZLES_E_OUTPUT_FULL;
        const u32 nseg = (u32)list.size();
        RET(c->seg_pos.reserve((size_t)nseg * 4));
        CK(zrt_h2d(c->seg_pos.p, list.data(), (size_t)nseg * 4, c->stream));
        CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
        CK(zrt_sync(c->stream));  // list is host heap memory
        RET(launch_phase_b(c, (const u32 *)c->seg_pos.as<u32>(), nseg, d_in, d_out, cap));
        CK(zrt_last_error());
        RET(read_ctl(c, &h));
        if (h.ok_res == 0) return 0;
      }
    }
  }

  if (!has_final || ours_only) return ZLES_E_CORRUPTED;  // a shard / slab of one of our streams must have decoded above

  // 3b. a stream from another encoder (zlib.es itself, system zlib): find the dynamic blocks, decode them in
  //     parallel, chain them up.  Only a complete, consistent chain that ends in a BFINAL block is accepted;
  //     everything else is left to the sequential decoder below.
  {
    int rc = inflate_foreign(c, d_in, n, first, d_out, cap, out_len);
    if (rc >= 0) return rc;  // decoded (0) or ZLES_E_OUTPUT_FULL / CUDA error; -1 = not handled
  }

  // 4. sequential decode of the whole stream on one warp: exactly the reference's order of
  //    events (src/inflate.ts:22-37), used for everything else and for error reporting.
  {
    const u32 one = 1;
    const u64 zero = 0;
    RET(c->seg_pos.reserve(8));
    RET(c->seg_off.reserve(8));
    RET(c->res.reserve(sizeof(InfRes)));
    CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
    CK(zrt_h2d(&ctl->ncand, &one, 4, c->stream));
    CK(zrt_h2d(c->seg_pos.p, &first, 8, c->stream));
    CK(zrt_h2d(c->seg_off.p, &zero, 8, c->stream));
    CK(zrt_sync(c->stream));
    LAUNCH(c, k_inflate, 1, INF_THREADS, INF_SMEM, d_in, (u64)n, (const u64 *)c->seg_pos.as<u64>(), (const u64 *)c->seg_off.as<u64>(),
           (const u32 *)&ctl->ncand, 1u, d_out, (u64)cap, 0, c->res.as<InfRes>(), &ctl->counter);
    CK(zrt_last_error());
    CK(zrt_mail(&c->mail->res0, c->res.p, sizeof(InfRes), c->stream));
    CK(zrt_sync(c->stream));
    const InfRes r = c->mail->res0;
    if (r.status != SEG_FINAL) return seg_status_to_code(r.status);
    *out_len = (size_t)r.out_len;
    if (r.flags & SEGF_OVERFLOW) return ZLES_E_OUTPUT_FULL;
    return 0;
  }
}

static int inflate_body(zles_ctx *c, const u8 *d_in, size_t n, u64 first, u8 *d_out, size_t cap, size_t *out_len, bool has_final = true,
                        bool ours_only = false) {
  u32 ncand = 0, cand_cap = 0;
  RET(inflate_scan(c, d_in, n, first, &ncand, &cand_cap));
  return inflate_decode(c, d_in, n, first, ncand, cand_cap, d_out, cap, out_len, has_final, ours_only);
}


// ---- host-buffer inflate of OUR streams, slab by slab ----------------------------------------------
// The blocks of our own streams are byte aligned (every one ends with the 00 00 FF FF marker), 32 KiB of output each,
// and chunks of four never refer to one another.  So a long run of them can be decoded a slab of whole chunks at a
// time: while slab k is decoded on the context's stream, slab k - 1 (final once decoded) travels to the host on the
// copy stream — the device-to-host copy of the output no longer waits for the whole decode.
// d_in[0 .. n): marker-delimited blocks whose starts (relative to d_in, ascending, starts[0] = first block) are given;
// the run stands for whole chunks except for its last block when has_final.  Output goes to h_out (capacity h_cap).
// Returns 0, ZLES_E_OUTPUT_FULL (*out_len = size needed), ZLES_E_CUDA, or -1: not decodable this way (the caller then
// runs the general path on the same device bytes).
constexpr size_t INF_SLAB_MIN_STREAM = 1u << 16;  // shorter streams are not worth the extra scan
constexpr u32 INF_SLAB_BLOCKS = 16384;  // 512 MiB of output per slab: enough blocks for one warp per block to fill the GPU
static int scan_block_starts(zles_ctx *c, const u8 *d_in, size_t n, u64 first, std::vector<u64> &starts);

// inflate_body for a run of our blocks whose starts the host already knows (h_starts[0 .. ncand), `rebase` is subtracted:
// they become offsets into d_in): the marker scan — two kernels and a wait per slab — is skipped, the list goes to the
// device with the launches.  Our blocks only (anything else: ZLES_E_CORRUPTED, the caller takes the general path).
static int inflate_known_starts(zles_ctx *c, const u8 *d_in, size_t n, const u64 *h_starts, u32 ncand, u64 rebase, u8 *d_out, size_t cap,
                                size_t *out_len, bool has_final) {
  RET(c->ctl.reserve(sizeof(InfCtl)));
  InfCtl *ctl = c->ctl.as<InfCtl>();
  const u64 cand_cap64 = (u64)n / 32 + 64;
  if (cand_cap64 > 0x7fffffffull) return ZLES_E_ARG;
  if (ncand == 0 || ncand > cand_cap64) return ZLES_E_CORRUPTED;
  RET(c->cand.reserve((size_t)cand_cap64 * 8));
  if (c->slab_cand_cap < (size_t)ncand + 1) {
    if (c->slab_cand) zrt_host_free(c->slab_cand);
    c->slab_cand = nullptr;
    c->slab_cand_cap = 0;
    const size_t want = (size_t)ncand + ((size_t)ncand >> 2) + 64;
    CK(zrt_host_alloc(reinterpret_cast<void **>(&c->slab_cand), want * 8));
    c->slab_cand_cap = want;
  }
  for (u32 j = 0; j < ncand; j++) c->slab_cand[j] = h_starts[j] - rebase;
  c->slab_cand[ncand] = ncand;  // (its low word is what goes into ctl->ncand)
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  // read from the pinned list by a kernel: a copy would queue behind the bulk host-to-device copies of the stream's pieces
  // (measured: 55 ms per 8 GiB lost that way)
  CK(zrt_mail(c->cand.p, c->slab_cand, (size_t)ncand * 8, c->stream));
  CK(zrt_mail(&ctl->ncand, c->slab_cand + ncand, 4, c->stream));
  // (the pinned list is rewritten for the next slab only after inflate_decode has waited for this one's result)
  return inflate_decode(c, d_in, n, c->slab_cand[0], ncand, (u32)cand_cap64, d_out, cap, out_len, has_final, /*ours_only=*/true);
}

// One slab after the other: decode(b0, b1) decodes blocks [b0, b1) of the run into one of two device buffers and queues
// its copy to the host on the output stream; finish() waits for the copies.
struct SlabDecoder {
  zles_ctx *c;
  const u8 *d_in;
  bool has_final;
  u8 *h_out;
  size_t h_cap;
  u32 slab_blocks;
  Drainer *drain = nullptr;  // the caller's buffer is pageable: finished slabs go through the pinned staging ring
  size_t slab_out = 0;
  zrt_event_t done[2], copied[2];
  bool copied_valid[2] = {false, false};
  bool have_events = false;
  u32 k = 0;
  size_t total = 0;

  int begin() {
    slab_blocks = (slab_blocks / SUBS_PER_CHUNK) * SUBS_PER_CHUNK;
    if (slab_blocks == 0) slab_blocks = SUBS_PER_CHUNK;
    slab_out = (size_t)slab_blocks * SUB;
    RET(c->d_out.reserve(2 * slab_out + 64));
    for (int i = 0; i < 2; i++) { done[i] = timing_event(c); copied[i] = timing_event(c); }
    have_events = true;
    return 0;
  }
  // blocks [b0, b1) start at starts[b0 ..]; `end` = where the slab's bytes end (the next block's start, or the run's end);
  // last: the run's final slab.  Returns 0, ZLES_E_OUTPUT_FULL (total = size needed so far), ZLES_E_CUDA or -1.
  int decode(const std::vector<u64> &starts, size_t b0, size_t b1, u64 end, bool last) {
    const u64 al = (starts[b0] + (u64)((uintptr_t)d_in & 15)) & 15;  // the slab's bytes from a 16-byte aligned address on
    const u64 in0 = starts[b0] - al;                                  // (al <= starts[b0]: device buffers are 256-byte aligned)
    u8 *d_slab = c->d_out.as<u8>() + (size_t)(k & 1) * (slab_out + 32);
    const size_t off = b0 * (size_t)SUB;
    const size_t cap = off >= h_cap ? 0 : (h_cap - off < slab_out ? h_cap - off : slab_out);
    if (copied_valid[k & 1]) CK(zrt_stream_wait_event(c->stream, copied[k & 1]));  // the copy out of this buffer (slab k - 2) is done
    size_t olen = 0;
    TRACE("slab %u: blocks [%zu, %zu) begin", k, b0, b1);
    int rc = inflate_known_starts(c, d_in + in0, (size_t)(end - in0), starts.data() + b0, (u32)(b1 - b0), in0, d_slab, cap, &olen, has_final && last);
    TRACE("slab %u: decoded rc=%d olen=%zu", k, rc, olen);
    if (rc == ZLES_E_OUTPUT_FULL) { total = off + olen; return rc; }
    if (rc == ZLES_E_CUDA) return rc;
    if (rc) return -1;
    if (!(last && has_final) && olen != (b1 - b0) * (size_t)SUB) return -1;
    CK(zrt_event_record(done[k & 1], c->stream));
    CK(zrt_stream_wait_event(c->out_stream, done[k & 1]));
    if (olen) {
      if (drain) { if (!drain->push(h_out + off, d_slab, olen)) return cuda_fail(zrt_last_error(), "staged copy to host"); }
      else CK(zrt_d2h(h_out + off, d_slab, olen, c->out_stream));
    }
    CK(zrt_event_record(copied[k & 1], c->out_stream));
    copied_valid[k & 1] = true;
    total = off + olen;
    k++;
    return 0;
  }
  int finish() {
    bool drained = true;
    if (drain) drained = drain->finish();
    zrt_err_t e = zrt_sync(c->out_stream);
    if (e == ZRT_OK && !drained) e = zrt_last_error();
    TRACE("slabs: copies done");
    if (have_events)
      for (int i = 0; i < 2; i++) { c->event_pool.push_back(done[i]); c->event_pool.push_back(copied[i]); }
    have_events = false;
    if (e != ZRT_OK) return cuda_fail(e, "copy to host");
    return 0;
  }
};

static int inflate_slabs_to_host(zles_ctx *c, const u8 *d_in, size_t n, const std::vector<u64> &starts, bool has_final, u8 *h_out, size_t h_cap,
                                 size_t *out_len, u32 slab_blocks = INF_SLAB_BLOCKS) {
  const size_t B = starts.size();
  if (B == 0) return -1;
  // every block but the last stands for exactly 32 KiB: when those cannot fit, the general path reports the exact size
  if (B > 1 && (u64)(B - 1) * SUB > (u64)h_cap) return -1;
  Drainer drain;
  SlabDecoder sd{c, d_in, has_final, h_out, h_cap, slab_blocks};
  if (B * (size_t)SUB >= STAGE_MIN && host_is_pageable(h_out) && drain.start(c->device, &c->ring_out, c->out_stream) == 0) sd.drain = &drain;
  RET(sd.begin());
  int rc = 0;
  for (size_t b0 = 0; b0 < B && rc == 0; b0 += sd.slab_blocks) {
    const size_t b1 = b0 + sd.slab_blocks < B ? b0 + sd.slab_blocks : B;
    rc = sd.decode(starts, b0, b1, b1 == B ? (u64)n : starts[b1], b1 == B);
  }
  const int rf = sd.finish();
  *out_len = sd.total;
  if (rc == 0 && rf) return rf;
  return rc;
}

// The same for a stream that is still on its way to the device: the compressed bytes are copied in pieces (16 MiB, then
// doubling up to 128 MiB) on the copy stream; as each piece lands it is scanned for block starts and every slab whose
// bytes are complete is decoded, its output leaving on the output stream — host-to-device copy, decode and device-to-host
// copy all overlap.  h_in[0 .. n) is the whole stream, `first` its first block.  Returns like inflate_slabs_to_host; on -1
// the whole stream is in device memory (the caller runs the general path on it).
static int inflate_streaming_to_host(zles_ctx *c, const u8 *h_in, size_t n, u64 first, u8 *h_out, size_t h_cap, size_t *out_len) {
  u8 *d_in = c->d_in.as<u8>();
  std::vector<size_t> pb{0};
  size_t piece0 = (c->inf_stream_min / 6) & ~(size_t)4095;  // 16 MiB with the default threshold
  if (piece0 < 65536) piece0 = 65536;
  if (piece0 > ((size_t)16 << 20)) piece0 = (size_t)16 << 20;
  for (size_t sz = piece0; pb.back() < n; sz = sz < 8 * piece0 ? sz * 2 : sz) pb.push_back(pb.back() + sz < n ? pb.back() + sz : n);
  const size_t np = pb.size() - 1;
  std::vector<zrt_event_t> ev(np);
  Feeder feeder;
  bool staged = false;
  if (n >= STAGE_MIN && host_is_pageable(h_in)) {  // a pageable source goes through the pinned staging ring (stager.inl)
    std::vector<StagePiece> pieces;
    for (size_t i = 0; i < np; i++)
      for (size_t o = pb[i]; o < pb[i + 1]; o += STAGE_SLOT) pieces.push_back(StagePiece{h_in + o, d_in + o, std::min<size_t>(STAGE_SLOT, pb[i + 1] - o), (u32)i});
    staged = feeder.start(c->device, &c->ring_in, c->copy_stream, std::move(pieces), (u32)np) == 0;
  }
  for (size_t i = 0; i < np; i++) {
    ev[i] = timing_event(c);
    if (staged) continue;  // recorded below, once the piece's copies have been enqueued
    CK(zrt_h2d(d_in + pb[i], h_in + pb[i], pb[i + 1] - pb[i], c->copy_stream));
    CK(zrt_event_record(ev[i], c->copy_stream));
  }
  Drainer drain;
  SlabDecoder sd{c, d_in, true, h_out, h_cap, c->inf_slab_blocks ? c->inf_slab_blocks : INF_SLAB_BLOCKS};
  if (h_cap >= STAGE_MIN && host_is_pageable(h_out) && drain.start(c->device, &c->ring_out, c->out_stream) == 0) sd.drain = &drain;
  int rc = sd.begin();
  std::vector<u64> starts, found;
  size_t next = 0;  // first block not decoded yet
  for (size_t i = 0; i < np && rc == 0; i++) {
    if (staged) {
      if (!feeder.wait_group((u32)i)) { rc = cuda_fail(zrt_last_error(), "staged copy to device"); break; }
      CK(zrt_event_record(ev[i], c->copy_stream));
    }
    CK(zrt_stream_wait_event(c->stream, ev[i]));
    const size_t lo = i == 0 ? 0 : pb[i] - 16, hi = pb[i + 1];
    const int rs = scan_block_starts(c, d_in + lo, hi - lo, i == 0 ? first : 15, found);
    if (rs) { rc = rs > 0 ? rs : -1; break; }
    for (size_t j = i == 0 ? 0 : 1; j < found.size(); j++) starts.push_back(found[j] + lo);
    TRACE("piece %zu: %zu blocks so far", i, starts.size());
    const bool all = i + 1 == np;
    // every block but the last stands for exactly 32 KiB: when those cannot fit, the general path reports the exact size
    if (starts.size() > 1 && (u64)(starts.size() - 1) * SUB > (u64)h_cap) { rc = -1; break; }
    while (rc == 0 && (all ? next < starts.size() : next + sd.slab_blocks < starts.size())) {
      const size_t b1 = next + sd.slab_blocks < starts.size() ? next + sd.slab_blocks : starts.size();
      const bool last = all && b1 == starts.size();
      rc = sd.decode(starts, next, b1, last ? (u64)n : starts[b1], last);
      next = b1;
    }
  }
  for (size_t i = 0; i < np; i++) c->event_pool.push_back(ev[i]);
  if (staged && !feeder.finish() && rc == 0) rc = cuda_fail(zrt_last_error(), "staged copy to device");
  zrt_err_t e = zrt_sync(c->copy_stream);  // whatever happens next needs the whole stream on the device
  const int rf = sd.finish();
  if (e != ZRT_OK) return cuda_fail(e, "copy to device");
  *out_len = sd.total;
  if (rc == 0 && rf) return rf;
  return rc;
}

// blocks per slab for a run of B blocks: the context's setting, or (0 = automatic) a quarter of the run, between 64 MiB
// and 512 MiB of output — enough work per slab to fill the GPU, enough slabs for the copies to overlap
static u32 inflate_slab_size(const zles_ctx *c, size_t B) {
  if (c->inf_slab_blocks) return c->inf_slab_blocks;
  size_t q = (B / 4 / SUBS_PER_CHUNK) * SUBS_PER_CHUNK;
  if (q < 2048) q = 2048;  // 64 MiB: a smaller slab does not fill the GPU (measured: 64 MiB in four slabs decodes in 4.7 ms, in one go in 2.1)
  if (q > INF_SLAB_BLOCKS) q = INF_SLAB_BLOCKS;
  return (u32)q;
}

// The block starts of one of our streams, on the host: candidates found by the marker scan of d_in[0 .. n) (first = where
// the first block starts).  Returns 0 and fills `starts`, or -1 when the scan says "not ours" (too many candidates).
static int scan_block_starts(zles_ctx *c, const u8 *d_in, size_t n, u64 first, std::vector<u64> &starts) {
  u32 ncand = 0, cand_cap = 0;
  RET(inflate_scan(c, d_in, n, first, &ncand, &cand_cap));
  if (ncand == 0 || ncand > cand_cap) return -1;
  // through pinned memory, written by a kernel: a copy would queue behind whatever bulk device-to-host copy is in flight
  if (c->cand_mail_cap < ncand) {
    if (c->cand_mail) zrt_host_free(c->cand_mail);
    c->cand_mail = nullptr;
    c->cand_mail_cap = 0;
    const size_t want = (size_t)ncand + ((size_t)ncand >> 1) + 1024;
    CK(zrt_host_alloc(reinterpret_cast<void **>(&c->cand_mail), want * 8));
    c->cand_mail_cap = want;
  }
  CK(zrt_mail(c->cand_mail, c->cand.p, (size_t)ncand * 8, c->stream));
  CK(zrt_sync(c->stream));
  starts.assign(c->cand_mail, c->cand_mail + ncand);
  return 0;
}

// header check of zlib.inflate (src/zlib.ts:12-16): only CM is looked at.
static int check_zlib_header(const u8 *h, size_t n) {
  const u32 b0 = n ? h[0] : 0;  // reading past the end yields 0 bits
  if ((b0 & 15) != 8) return ZLES_E_NOT_DEFLATE;
  return 0;
}

extern "C" int zles_dev_inflate(zles_ctx *c, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t cap, size_t *out_len) {
  if ((!d_in && n) || !out_len || (!d_out && cap)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  c->mail->head[0] = 0;
  if (n) {
    CK(zrt_d2h(c->mail->head, d_in, 1, c->stream));
    CK(zrt_sync(c->stream));
  }
  RET(check_zlib_header(c->mail->head, n));
  return inflate_body(c, d_in, n, 2, d_out, cap, out_len);
}

extern "C" int zles_dev_inflate_segment(zles_ctx *c, const uint8_t *d_in, size_t n, int has_final, uint8_t *d_out, size_t cap,
                                        size_t *out_len) {
  if ((!d_in && n) || !out_len) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  if (n == 0 && !has_final) { *out_len = 0; return 0; }
  return inflate_body(c, d_in, n, 0, d_out, cap, out_len, has_final != 0);
}

extern "C" int zles_dev_scan_blocks(zles_ctx *c, const uint8_t *d_in, size_t n, uint64_t first, uint64_t *starts, size_t cap, size_t *count) {
  if ((!d_in && n) || !count || (!starts && cap)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  std::vector<u64> v;
  const int rs = scan_block_starts(c, d_in, n, first, v);
  if (rs > 0) return rs;
  if (rs < 0) { *count = 0; return ZLES_E_CORRUPTED; }  // more candidates than one of our streams can hold
  *count = v.size();
  if (v.size() > cap) return ZLES_E_OUTPUT_FULL;
  memcpy(starts, v.data(), v.size() * 8);
  return 0;
}

// Host-buffer inflate of the raw deflate data that starts at byte `first` of in[0 .. n).  crc: when not null, the
// CRC-32 of the output is computed on the device (gzip) — the whole output then stays in device memory until it is checked.
static int inflate_host(zles_ctx *c, const uint8_t *in, size_t n, u64 first, uint8_t *out, size_t cap, size_t *out_len, uint32_t *crc = nullptr) {
  RET(resolve_ctx(c));
  TRACE("inflate: n=%zu cap=%zu", n, cap);
  RET(c->d_in.reserve(n + 16));
  if (!crc && n >= c->inf_stream_min && n >= 262144) {
    // long stream: pieces of it are scanned and decoded while the rest is still being copied in
    const int rc = inflate_streaming_to_host(c, in, n, first, out, cap, out_len);
    TRACE("inflate: streaming rc=%d", rc);
    if (rc >= 0) return rc;
    RET(c->d_out.reserve(cap + 16));  // not a long run of our blocks: the general path, on the bytes already in device memory
    const int rg = inflate_body(c, c->d_in.as<u8>(), n, first, c->d_out.as<u8>(), cap, out_len);
    if (rg) return rg;
    if (*out_len) CK(zrt_d2h(out, c->d_out.p, *out_len, c->stream));
    CK(zrt_sync(c->stream));
    return 0;
  }
  if (n) CK(zrt_h2d(c->d_in.p, in, n, c->stream));
  TRACE("inflate: h2d enqueued");
  if (!crc) {
    // one of our own streams, long enough to be worth it: decode slab by slab, copying finished slabs out meanwhile
    std::vector<u64> starts;
    int rs = n >= INF_SLAB_MIN_STREAM ? scan_block_starts(c, c->d_in.as<u8>(), n, first, starts) : -1;
    if (rs > 0) return rs;
    const u32 slab = inflate_slab_size(c, starts.size());
    TRACE("inflate: scan done rs=%d blocks=%zu slab=%u", rs, starts.size(), slab);
    if (rs == 0 && starts.size() >= 2 * (size_t)slab) {
      int rc = inflate_slabs_to_host(c, c->d_in.as<u8>(), n, starts, true, out, cap, out_len, slab);
      if (rc >= 0) return rc;
    }
  }
  RET(c->d_out.reserve(cap + 16));
  int rc = inflate_body(c, c->d_in.as<u8>(), n, first, c->d_out.as<u8>(), cap, out_len);
  if (rc) return rc;
  if (crc) RET(dev_crc32(c, c->d_out.as<u8>(), *out_len, crc));
  if (*out_len) CK(zrt_d2h(out, c->d_out.p, *out_len, c->stream));
  CK(zrt_sync(c->stream));
  return 0;
}

extern "C" int zles_inflate(zles_ctx *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len || (!out && cap)) return ZLES_E_ARG;
  if (!c)
    if (zles_mgpu *m = default_mgpu()) return zles_mgpu_inflate(m, in, n, out, cap, out_len);
  RET(check_zlib_header(in, n));
  return inflate_host(c, in, n, 2, out, cap, out_len);
}

// inflate(input, offset = 0) of the reference's core (/root/reference/src/inflate.ts:16): raw deflate data from byte `offset` on
extern "C" int zles_inflate_raw(zles_ctx *c, const uint8_t *in, size_t n, size_t offset, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len || (!out && cap)) return ZLES_E_ARG;
  return inflate_host(c, in, n, offset, out, cap, out_len);
}

// gzip member header (RFC 1952 2.3): returns its length, or 0 with *err set
static size_t gzip_header_len(const u8 *in, size_t n, int *err) {
  *err = 0;
  if (n < 10) { *err = ZLES_E_LACK; return 0; }
  if (in[0] != 0x1f || in[1] != 0x8b || in[2] != 8) { *err = ZLES_E_NOT_DEFLATE; return 0; }
  const u32 flg = in[3];
  size_t p = 10;
  if (flg & 4) {  // FEXTRA
    if (p + 2 > n) { *err = ZLES_E_LACK; return 0; }
    p += 2 + ((size_t)in[p] | ((size_t)in[p + 1] << 8));
  }
  for (u32 bit = 8; bit <= 16; bit <<= 1)  // FNAME, FCOMMENT: zero-terminated
    if (flg & bit) {
      while (p < n && in[p]) p++;
      p++;
    }
  if (flg & 2) p += 2;  // FHCRC
  if (p > n) { *err = ZLES_E_LACK; return 0; }
  return p;
}

// One gzip member: header, deflate data, CRC-32 and ISIZE in the last eight bytes of the buffer; both are verified.
extern "C" int zles_gzip_inflate(zles_ctx *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len || (!out && cap)) return ZLES_E_ARG;
  int err = 0;
  const size_t first = gzip_header_len(in, n, &err);
  if (err) return err;
  uint32_t crc = 0;
  RET(inflate_host(c, in, n, first, out, cap, out_len, &crc));
  if (n < first + 8) return ZLES_E_LACK;
  if (get_le32(in + n - 8) != crc || get_le32(in + n - 4) != (u32)*out_len) return ZLES_E_CHECKSUM;
  return 0;
}

extern "C" int zles_inflate_alloc(zles_ctx *c, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  if ((!in && n) || !out || !out_len) return ZLES_E_ARG;
  if (!c)
    if (zles_mgpu *m = default_mgpu()) return zles_mgpu_inflate_alloc(m, in, n, out, out_len);
  *out = nullptr;
  *out_len = 0;
  RET(check_zlib_header(in, n));
  RET(resolve_ctx(c));
  RET(c->d_in.reserve(n + 16));
  if (n) CK(zrt_h2d(c->d_in.p, in, n, c->stream));
  // capacity: the reference's own initial guess, 10 x input (src/inflate.ts:17), or what our own format
  // implies (32 KiB per marker-delimited block), whichever is larger; retry once with the exact size
  u32 ncand = 0, cand_cap = 0;
  RET(inflate_scan(c, c->d_in.as<u8>(), n, 2, &ncand, &cand_cap));
  size_t cap = n * 10 + CHUNK;
  if (ncand <= cand_cap && (size_t)ncand * SUB > cap) cap = (size_t)ncand * SUB;
  size_t need = 0;
  RET(c->d_out.reserve(cap + 16));
  int rc = inflate_decode(c, c->d_in.as<u8>(), n, 2, ncand, cand_cap, c->d_out.as<u8>(), cap, &need);
  if (rc == ZLES_E_OUTPUT_FULL) {
    cap = need;
    RET(c->d_out.reserve(cap + 16));
    rc = inflate_body(c, c->d_in.as<u8>(), n, 2, c->d_out.as<u8>(), cap, &need);
  }
  if (rc) return rc;
  u8 *buf = (u8 *)malloc(need ? need : 1);
  if (!buf) return ZLES_E_NOMEM;
  if (need) {
    zrt_err_t e = zrt_d2h(buf, c->d_out.p, need, c->stream);
    if (e == ZRT_OK) e = zrt_sync(c->stream);
    if (e != ZRT_OK) { free(buf); return cuda_fail(e, "copy to host"); }
  }
  *out = buf;
  *out_len = need;
  return 0;
}

// ---- batches of independent buffers --------------------------------------------------------

static int dev_deflate_batch(zles_ctx *c, const u8 *d_in, const u64 *d_in_off, const u64 *h_in_off, u32 count, u8 *d_out,
                             const u64 *d_out_off, u64 *d_out_len, int32_t *d_status);

extern "C" int zles_dev_inflate_batch(zles_ctx *c, const uint8_t *d_in, const uint64_t *d_in_off, uint32_t count, uint8_t *d_out,
                                      const uint64_t *d_out_off, uint64_t *d_out_len, int32_t *d_status) {
  if (!count) return 0;
  if (!d_in || !d_in_off || !d_out || !d_out_off || !d_out_len || !d_status) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  RET(c->ctl.reserve(sizeof(InfCtl)));
  InfCtl *ctl = c->ctl.as<InfCtl>();
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  LAUNCH(c, k_inflate_batch, inflate_grid(c, count), INF_THREADS, INF_SMEM, d_in, d_in_off, count, d_out, d_out_off, d_out_len, d_status,
         &ctl->counter, &ctl->ok);
  CK(zrt_last_error());
  CK(zrt_mail(&c->mail->ok, &ctl->ok, 4, c->stream));
  CK(zrt_sync(c->stream));
  return (int)c->mail->ok;  // the LARGEST per-stream status (atomicMax), 0 if none: callers inspect status[]
}

// ---- host forms of the batch calls ---------------------------------------------------------------------------
// The batch goes through the device a slab of buffers at a time (at most BATCH_SLAB_STREAMS buffers, about 64 MiB of
// input and 256 MiB of output room): the workspace is sized by the slab, not by the batch; on the device the results are
// compacted (k_batch_prefix + k_batch_gather) so that only the bytes that were produced cross PCIe, into a pinned staging
// ring; a helper thread moves every result from the ring to its place in the caller's buffer while the GPU works on the
// next slab.
constexpr u64 BATCH_SLAB_IN = (u64)64 << 20, BATCH_SLAB_OUT = (u64)256 << 20;

struct BatchJob {
  u32 s0, s1;           // buffers [s0, s1)
  int buf;              // staging buffer
  zrt_event_t ready;    // recorded after the slab's copies to the staging buffer
};

static int batch_host(zles_ctx *c, bool inflate, const uint8_t *in, const uint64_t *in_off, uint32_t count, uint8_t *out,
                      const uint64_t *out_off, uint64_t *out_len, int32_t *status) {
  RET(resolve_ctx(c));
  // slabs
  std::vector<u32> sb{0};
  for (u32 i = 0; i < count;) {
    u32 j = i + 1;
    while (j < count && j - i < c->batch_slab_streams && in_off[j + 1] - in_off[i] <= BATCH_SLAB_IN && out_off[j + 1] - out_off[i] <= BATCH_SLAB_OUT) j++;
    sb.push_back(j);
    i = j;
  }
  const u32 nslab = (u32)sb.size() - 1;
  u64 max_in = 0, max_out = 0;
  u32 max_cnt = 0;
  for (u32 k = 0; k < nslab; k++) {
    max_in = std::max<u64>(max_in, in_off[sb[k + 1]] - in_off[sb[k]]);
    max_out = std::max<u64>(max_out, out_off[sb[k + 1]] - out_off[sb[k]]);
    max_cnt = std::max(max_cnt, sb[k + 1] - sb[k]);
  }
  if (max_out >= 0xffffffffull) return ZLES_E_ARG;  // one buffer's room must be below 4 GiB
  RET(c->d_in.reserve((size_t)max_in + 16));
  RET(c->d_out.reserve((size_t)max_out + 16));
  RET(c->d_off_in.reserve(((size_t)max_cnt + 1) * 8));
  RET(c->d_off_out.reserve(((size_t)max_cnt + 1) * 8));
  RET(c->d_len.reserve((size_t)max_cnt * 8));
  RET(c->d_status.reserve((size_t)max_cnt * 4));
  RET(c->d_coff.reserve(((size_t)max_cnt + 1) * 8));
  // device-side compacted results, and their pinned landing place, double buffered: [bytes | out_len | status]
  const size_t meta = (size_t)max_cnt * 12 + 64;
  const size_t stage_bytes = (((size_t)max_out + 15) & ~(size_t)15) + meta;
  RET(c->d_pack.reserve(2 * stage_bytes));
  if (c->batch_stage_cap < 2 * stage_bytes) {
    if (c->batch_stage) zrt_host_free(c->batch_stage);
    c->batch_stage = nullptr;
    c->batch_stage_cap = 0;
    CK(zrt_host_alloc(reinterpret_cast<void **>(&c->batch_stage), 2 * stage_bytes));
    c->batch_stage_cap = 2 * stage_bytes;
  }
  // helper thread: results from the staging ring to the caller's buffers
  std::mutex mu;
  std::condition_variable cv;
  std::deque<BatchJob> jobs;
  bool closing = false;
  int helper_err = 0;
  u32 done_slabs = 0;
  auto helper = [&]() {
    zrt_set_device(c->device);
    for (;;) {
      BatchJob j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return closing || !jobs.empty(); });
        if (jobs.empty()) return;
        j = jobs.front();
        jobs.pop_front();
      }
      if (zrt_event_sync(j.ready) != ZRT_OK) helper_err = 1;
      const u8 *st = c->batch_stage + (size_t)j.buf * stage_bytes;
      const u32 cnt = j.s1 - j.s0;
      const u64 *lens = reinterpret_cast<const u64 *>(st + (((size_t)max_out + 15) & ~(size_t)15));
      const int32_t *stat = reinterpret_cast<const int32_t *>(lens + max_cnt);
      u64 o = 0;
      for (u32 i = 0; i < cnt; i++) {
        out_len[j.s0 + i] = lens[i];
        status[j.s0 + i] = stat[i];
        if (stat[i] == 0 && lens[i]) {
          memcpy(out + out_off[j.s0 + i], st + o, (size_t)lens[i]);
          o += lens[i];
        }
      }
      {
        std::unique_lock<std::mutex> lk(mu);
        done_slabs++;
      }
      cv.notify_all();
    }
  };
#ifndef ZLES_EMU
  std::thread th(helper);
#endif
  int rc = 0, worst = 0;
  std::vector<zrt_event_t> evs;
  std::vector<u64> oi, oo;
  for (u32 k = 0; k < nslab && rc == 0; k++) {
    const u32 s0 = sb[k], s1 = sb[k + 1], cnt = s1 - s0;
    const int buf = (int)(k & 1);
    oi.resize((size_t)cnt + 1);
    oo.resize((size_t)cnt + 1);
    for (u32 i = 0; i <= cnt; i++) { oi[i] = in_off[s0 + i] - in_off[s0]; oo[i] = out_off[s0 + i] - out_off[s0]; }
    const size_t nin = (size_t)oi[cnt];
    zrt_err_t e = ZRT_OK;
    if (nin) e = zrt_h2d(c->d_in.p, in + in_off[s0], nin, c->stream);
    if (e == ZRT_OK) e = zrt_h2d(c->d_off_in.p, oi.data(), ((size_t)cnt + 1) * 8, c->stream);
    if (e == ZRT_OK) e = zrt_h2d(c->d_off_out.p, oo.data(), ((size_t)cnt + 1) * 8, c->stream);
    if (e == ZRT_OK) e = zrt_sync(c->stream);  // the offset vectors are reused by the next slab
    if (e != ZRT_OK) { rc = cuda_fail(e, "copy to device"); break; }
    int r = inflate ? zles_dev_inflate_batch(c, c->d_in.as<u8>(), c->d_off_in.as<u64>(), cnt, c->d_out.as<u8>(), c->d_off_out.as<u64>(),
                                             c->d_len.as<u64>(), c->d_status.as<int32_t>())
                    : dev_deflate_batch(c, c->d_in.as<u8>(), c->d_off_in.as<u64>(), oi.data(), cnt, c->d_out.as<u8>(), c->d_off_out.as<u64>(),
                                        c->d_len.as<u64>(), c->d_status.as<int32_t>());
    if (r == ZLES_E_CUDA || r == ZLES_E_ARG || r == ZLES_E_NOMEM) { rc = r; break; }
    if (r > worst) worst = r;
    // the staging buffer of slab k - 2 must have been emptied by the helper before slab k's results land in it
    if (k >= 2) {
      std::unique_lock<std::mutex> lk(mu);
#ifdef ZLES_EMU
      (void)lk;
#else
      cv.wait(lk, [&] { return done_slabs + 2 > k; });
#endif
    }
    u8 *d_pk = c->d_pack.as<u8>() + (size_t)buf * stage_bytes;
    u8 *h_st = c->batch_stage + (size_t)buf * stage_bytes;
    const size_t meta_off = ((size_t)max_out + 15) & ~(size_t)15;
    LAUNCH(c, k_batch_prefix, 1, 1024, 64 * 4, (const u64 *)c->d_len.as<u64>(), (const int32_t *)c->d_status.as<int32_t>(), cnt, c->d_coff.as<u64>());
    LAUNCH(c, k_batch_gather, cnt < 65535u ? cnt : 65535u, 128, 0, (const u8 *)c->d_out.as<u8>(), (const u64 *)c->d_off_out.as<u64>(),
           (const u64 *)c->d_coff.as<u64>(), cnt, d_pk);
    e = zrt_last_error();
    // lengths and status words ride along in the packed buffer (the next slab's kernels rewrite d_len / d_status)
    if (e == ZRT_OK) e = zrt_copy(d_pk + meta_off, c->d_len.p, (size_t)cnt * 8, c->stream);
    if (e == ZRT_OK) e = zrt_copy(d_pk + meta_off + (size_t)max_cnt * 8, c->d_status.p, (size_t)cnt * 4, c->stream);
    if (e == ZRT_OK) e = zrt_mail(&c->mail->total, c->d_coff.as<u64>() + cnt, 8, c->stream);
    if (e == ZRT_OK) e = zrt_sync(c->stream);  // everything of this slab is in d_pk now
    if (e != ZRT_OK) { rc = cuda_fail(e, "batch compaction"); break; }
    const size_t produced = (size_t)c->mail->total;
    zrt_event_t ev = timing_event(c);
    evs.push_back(ev);
    if (produced) e = zrt_d2h(h_st, d_pk, produced, c->out_stream);
    if (e == ZRT_OK) e = zrt_d2h(h_st + meta_off, d_pk + meta_off, (size_t)max_cnt * 12, c->out_stream);
    if (e == ZRT_OK) e = zrt_event_record(ev, c->out_stream);
    if (e != ZRT_OK) { rc = cuda_fail(e, "copy to host"); break; }
    {
      std::unique_lock<std::mutex> lk(mu);
      jobs.push_back(BatchJob{s0, s1, buf, ev});
    }
    cv.notify_all();
#ifdef ZLES_EMU
    closing = true;
    helper();  // the emulator build has no threads: the helper's loop runs inline and returns when the queue is empty
    closing = false;
#endif
  }
  {
    std::unique_lock<std::mutex> lk(mu);
    closing = true;
  }
  cv.notify_all();
#ifndef ZLES_EMU
  th.join();
#endif
  for (zrt_event_t ev : evs) c->event_pool.push_back(ev);
  if (rc) return rc;
  if (helper_err) return cuda_fail(zrt_last_error(), "copy to host");
  return worst;
}

extern "C" int zles_inflate_batch(zles_ctx *c, const uint8_t *in, const uint64_t *in_off, uint32_t count, uint8_t *out,
                                  const uint64_t *out_off, uint64_t *out_len, int32_t *status) {
  if (!count) return 0;
  if (!in || !in_off || !out || !out_off || !out_len || !status) return ZLES_E_ARG;
  return batch_host(c, true, in, in_off, count, out, out_off, out_len, status);
}

extern "C" int zles_dev_deflate_batch(zles_ctx *c, const uint8_t *d_in, const uint64_t *d_in_off, uint32_t count, uint8_t *d_out,
                                      const uint64_t *d_out_off, uint64_t *d_out_len, int32_t *d_status) {
  if (!count) return 0;
  if (!d_in || !d_in_off || !d_out || !d_out_off || !d_out_len || !d_status) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  return dev_deflate_batch(c, d_in, d_in_off, nullptr, count, d_out, d_out_off, d_out_len, d_status);
}

extern "C" int zles_deflate_batch(zles_ctx *c, const uint8_t *in, const uint64_t *in_off, uint32_t count, uint8_t *out,
                                  const uint64_t *out_off, uint64_t *out_len, int32_t *status) {
  if (!count) return 0;
  if (!in || !in_off || !out || !out_off || !out_len || !status) return ZLES_E_ARG;
  return batch_host(c, false, in, in_off, count, out, out_off, out_len, status);
}

// Batch deflate: every buffer is its own zlib stream (header, blocks, Adler-32 trailer all
// written on the device so that one launch sequence serves the whole batch).
static int dev_deflate_batch(zles_ctx *c, const u8 *d_in, const u64 *d_in_off, const u64 *h_in_off, u32 count, u8 *d_out,
                             const u64 *d_out_off, u64 *d_out_len, int32_t *d_status) {
  (void)h_in_off;
  RET(c->ctl.reserve(sizeof(InfCtl)));
  InfCtl *ctl = c->ctl.as<InfCtl>();
  CK(zrt_memset(ctl, 0, sizeof(InfCtl), c->stream));
  // per-buffer block counts -> block table
  RET(c->seg_pos.reserve(((size_t)count + 2) * 8));
  u64 *d_blk_first = c->seg_pos.as<u64>();  // [count + 1] first block index of each buffer, then the longest block's bytes
  LAUNCH(c, k_batch_count, 1, 1024, 64 * 4, d_in_off, count, d_blk_first);
  CK(zrt_last_error());
  CK(zrt_mail(c->mail->summary, d_blk_first + count, 16, c->stream));
  CK(zrt_sync(c->stream));
  const u64 nb64 = c->mail->summary[0];
  // token slots per block: a block of b bytes has at most b tokens (16-slot granularity keeps rows 64-byte aligned)
  const u32 tok_stride = (u32)std::max<u64>(16, (umin64(c->mail->summary[1], (u64)SUB) + 15) & ~15ull);
  if (nb64 > 0x7fffffffull / LZ_NSYM) return ZLES_E_ARG;
  const u32 nblocks = (u32)nb64;
  const u32 grid_lz = (u32)umin64((u64)nblocks, (u64)c->sm_count);
  RET(c->seg_off.reserve((size_t)nblocks * sizeof(BatchBlk)));
  BatchBlk *d_tab = c->seg_off.as<BatchBlk>();
  RET(c->tokens.reserve((size_t)nblocks * tok_stride * 4));
  RET(c->ntok.reserve((size_t)nblocks * 4));
  RET(c->hist.reserve((size_t)nblocks * LZ_NSYM * 4));
  RET(c->scratch.reserve((size_t)grid_lz * 2 * SUB * 4));
  RET(c->adler_part.reserve((size_t)nblocks * 16));
  RET(c->codes.reserve((size_t)nblocks * sizeof(BlockCodes)));
  RET(c->blk_bits.reserve((size_t)nblocks * 4));
  c->p1_valid = false;
  LAUNCH(c, k_batch_table, (count + 255) / 256, 256, 0, d_in_off, count, (const u64 *)d_blk_first, d_tab, c->pair_mode);

  LzParams lp;
  lp.in = d_in;
  lp.n = 0;
  lp.nblocks = nblocks;
  lp.tokens = c->tokens.as<u32>();
  lp.ntok = c->ntok.as<u32>();
  lp.hist = c->hist.as<u32>();
  lp.scratch = c->scratch.as<u32>();
  lp.adler_part = c->adler_part.as<u64>();
  lp.max_checks = c->max_checks;
  lp.min_checks = c->min_checks;
  lp.good_len = c->good_len;
  lp.lazy = c->lazy;
  lp.table = d_tab;
  lp.tok_stride = tok_stride;
  RET(c->unit_ctr.reserve(4));
  lp.unit_ctr = c->unit_ctr.as<u32>();
  CK(zrt_memset(lp.unit_ctr, 0, 4, c->stream));
  LAUNCH(c, k_lz_batch, grid_lz, LZ_THREADS, LZ_SMEM, lp);
  LAUNCH(c, k_huff, (nblocks + HUF_WARPS - 1) / HUF_WARPS, HUF_THREADS, HUF_SMEM, (const u32 *)c->hist.as<u32>(), 0u, nblocks,
         c->codes.as<BlockCodes>(), c->blk_bits.as<u32>(), (u64)0, (const BatchBlk *)d_tab);
  BatchPackParams bp;
  bp.tokens = c->tokens.as<u32>();
  bp.ntok = c->ntok.as<u32>();
  bp.codes = c->codes.as<BlockCodes>();
  bp.blk_bits = c->blk_bits.as<u32>();
  bp.adler_part = c->adler_part.as<u64>();
  bp.table = d_tab;
  bp.blk_first = d_blk_first;
  bp.in_off = d_in_off;
  bp.out_off = d_out_off;
  bp.count = count;
  bp.in = d_in;
  bp.out = d_out;
  bp.out_len = d_out_len;
  bp.status = d_status;
  bp.first_err = &ctl->ok;
  bp.tok_stride = tok_stride;
  LAUNCH(c, k_pack_batch, count, PACK_THREADS, PACK_SMEM, bp);
  CK(zrt_last_error());
  CK(zrt_mail(&c->mail->ok, &ctl->ok, 4, c->stream));
  CK(zrt_sync(c->stream));
  return (int)c->mail->ok;
}

// ---- CUDA IPC (peer-mapped destination of the sharded deflate) ----------------------------------

extern "C" int zles_ipc_export(const void *d_ptr, uint8_t handle[64]) {
#ifdef ZLES_EMU
  (void)d_ptr; (void)handle;
  return ZLES_E_CUDA;
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!d_ptr || !handle) return ZLES_E_ARG;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle, &h, 64);
  return 0;
#endif
}
extern "C" int zles_ipc_open(const uint8_t handle[64], void **d_ptr) {
#ifdef ZLES_EMU
  (void)handle; (void)d_ptr;
  return ZLES_E_CUDA;
#else
  if (!d_ptr || !handle) return ZLES_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
#endif
}
extern "C" int zles_ipc_close(void *d_ptr) {
#ifdef ZLES_EMU
  (void)d_ptr;
  return ZLES_E_CUDA;
#else
  if (!d_ptr) return ZLES_E_ARG;
  CK(cudaIpcCloseMemHandle(d_ptr));
  return 0;
#endif
}

// ---- synthetic corpora ------------------------------------------------------------------------

static std::once_flag g_corpus_once;
static CorpusTable *g_corpus_host = nullptr;

static int corpus_sym_of(int ch) {
  if (ch >= 'a' && ch <= 'z') return ch - 'a';
  if (ch >= 'A' && ch <= 'Z') return 26 + ch - 'A';
  switch (ch) {
    case ' ': return 52;
    case ',': return 53;
    case '.': return 54;
    case ';': return 55;
    case '\'': return 56;
    case '-': return 57;
    case '\n': return 58;
  }
  return -1;
}

static void corpus_build_host() {
  CorpusTable *T = new CorpusTable();
  static const char alpha[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ ,.;'-\n";
  memset(T->alphabet, ' ', sizeof(T->alphabet));
  for (u32 i = 0; i < CORPUS_NSYM; i++) T->alphabet[i] = (u8)alpha[i];
  std::vector<u32> c2((size_t)CORPUS_NSYM * CORPUS_NSYM * CORPUS_NSYM, 0), c1((size_t)CORPUS_NSYM * CORPUS_NSYM, 0), c0(CORPUS_NSYM, 0);
  int s1 = 4, s2 = 52;
  for (const char *p = ZLES_CORPUS_PROSE; *p; p++) {
    int s = corpus_sym_of((unsigned char)*p);
    if (s < 0) continue;
    c2[((size_t)s1 * CORPUS_NSYM + s2) * CORPUS_NSYM + s]++;
    c1[(size_t)s2 * CORPUS_NSYM + s]++;
    c0[s]++;
    s1 = s2;
    s2 = s;
  }
  for (u32 st = 0; st < CORPUS_NSYM * CORPUS_NSYM; st++) {
    const u32 *cnt = &c2[(size_t)st * CORPUS_NSYM];
    u64 tot = 0;
    for (u32 k = 0; k < CORPUS_NSYM; k++) tot += cnt[k];
    if (tot == 0) {  // unseen pair: back off to the order-1, then the order-0 statistics
      cnt = &c1[(size_t)(st % CORPUS_NSYM) * CORPUS_NSYM];
      for (u32 k = 0; k < CORPUS_NSYM; k++) tot += cnt[k];
      if (tot == 0) {
        cnt = c0.data();
        for (u32 k = 0; k < CORPUS_NSYM; k++) tot += cnt[k];
      }
    }
    u64 cum = 0;
    for (u32 k = 0; k < 64; k++) {
      if (k < CORPUS_NSYM) cum += cnt[k];
      u64 v = tot ? (cum * 65536ull) / tot : 65535;
      T->cdf[st][k] = (u16)(v > 65535 ? 65535 : v);
    }
  }
  g_corpus_host = T;
}

static const CorpusTable *corpus_host_table() {
  std::call_once(g_corpus_once, corpus_build_host);
  return g_corpus_host;
}

extern "C" int zles_host_corpus(int kind, uint64_t offset, uint8_t *out, size_t n) {
  if (kind < 0 || kind > 3 || (!out && n)) return ZLES_E_ARG;
  const CorpusTable *T = corpus_host_table();
  if (!n) return 0;
  const u64 p0 = offset / CORPUS_PAGE, p1 = (offset + n + CORPUS_PAGE - 1) / CORPUS_PAGE;
  for (u64 p = p0; p < p1; p++) corpus_page(T, kind, p, offset, offset + n, out);
  return 0;
}

extern "C" int zles_dev_corpus(zles_ctx *c, int kind, uint64_t offset, uint8_t *d_out, size_t n) {
  if (kind < 0 || kind > 3 || (!d_out && n)) return ZLES_E_ARG;
  RET(resolve_ctx(c));
  if (!c->d_corpus) {
    void *p = nullptr;
    CK(zrt_malloc(&p, sizeof(CorpusTable)));
    c->d_corpus = reinterpret_cast<CorpusTable *>(p);
    CK(zrt_h2d(c->d_corpus, corpus_host_table(), sizeof(CorpusTable), c->stream));
    CK(zrt_sync(c->stream));
  }
  if (!n) return 0;
  const u64 p0 = offset / CORPUS_PAGE, p1 = (offset + n + CORPUS_PAGE - 1) / CORPUS_PAGE;
  const u32 grid = (u32)((p1 - p0 + 31) / 32);
  LAUNCH(c, k_corpus, grid, 32, 0, (const CorpusTable *)c->d_corpus, kind, (u64)offset, d_out, (u64)n);
  CK(zrt_last_error());
  CK(zrt_sync(c->stream));
  return 0;
}

#include "mgpu.inl"

#ifdef ZLES_STAGE_CLOCKS
// profiling build only (tools/lz_stages.py): per-stage cycle totals of k_lz, summed over CTAs
extern "C" int zles_debug_lz_clocks(unsigned long long *out16, int reset) {
  if (out16) CK(cudaMemcpyFromSymbol(out16, g_lz_clk, sizeof(unsigned long long) * 16));
  if (reset) {
    unsigned long long z[16] = {0};
    CK(cudaMemcpyToSymbol(g_lz_clk, z, sizeof z));
  }
  return ZLES_OK;
}
extern "C" int zles_debug_inf_clocks(unsigned long long *out8, int reset) {
  if (out8) CK(cudaMemcpyFromSymbol(out8, g_inf_clk, sizeof(unsigned long long) * 8));
  if (reset) {
    unsigned long long z[8] = {0};
    CK(cudaMemcpyToSymbol(g_inf_clk, z, sizeof z));
  }
  return ZLES_OK;
}
#endif
