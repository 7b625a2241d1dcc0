// mgpu.inl — one process, several GPUs, behind the drop-in boundary (included at the end of zles.cu).
//
// The reference is a synchronous two-function library (/root/reference/src/zlib.ts:11-49); a Node process that loads
// the N-API addon is ONE process, so sharding over the GPUs of a box has to live below the C ABI (SURVEY.md §8b,
// §8e).  A zles_mgpu owns one zles_ctx and one worker thread per device.
//
//   deflate   the input's 128 KiB chunks are cut into contiguous shards, one per device.  Every worker copies its
//             shard in and runs phase 1 (match finder, codes, layout; the packer runs slab by slab into device
//             staging meanwhile).  ONE exchange — on the host, a prefix sum of the shards' compressed sizes, the same
//             numbers the multi-process form all-gathers over NCCL — then every worker copies its blocks to their
//             place in the caller's buffer.  Header and Adler-32 trailer are written by the calling thread
//             (/root/reference/src/zlib.ts:28-46 stays on the host).
//   inflate   the shards are found from the stream itself: every worker copies in an equal slice of the compressed
//             bytes and scans it for block markers; the host concatenates the block starts (block j of our streams
//             stands for bytes [32 KiB j, 32 KiB (j + 1)) of the output, chunks of four are independent), assigns whole
//             chunks to devices, and every worker fetches the bytes of its share it does not have yet and decodes slab
//             by slab, copying finished slabs out meanwhile.  Anything that is not a well-formed stream of ours is
//             handed to the single-device path, which knows every other kind of stream and every error.
//
// Host buffers only cross PCIe once per direction; no peer traffic is needed (the stream is assembled in host memory).
#include <condition_variable>
#include <functional>
#include <thread>

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wsubobject-linkage"  // DevBuf lives in zles.cu's anonymous namespace; this file is part of that TU
#endif
struct zles_mgpu {
  std::vector<int> devices;
  std::vector<zles_ctx *> ctx;
  size_t min_shard = (size_t)4 << 20;  // inputs shorter than this per device are not worth sharding
  // worker pool: run(fn) executes fn(r) on worker r for every device and waits for all of them
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<int(int)> task;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> err;
  // inflate: per-device staging of the slice that was scanned
  std::vector<DevBuf> slice, range;

  int run(const std::function<int(int)> &fn) {
    const int R = (int)devices.size();
#ifdef ZLES_EMU
    // the CPU thread emulator is not re-entrant: the "devices" take their turns on the calling thread
    for (int r = 0; r < R; r++) { rc[r] = fn(r); err[r] = g_cuda_err; }
#else
    {
      std::unique_lock<std::mutex> lk(mu);
      task = fn;
      pending = R;
      generation++;
    }
    cv_go.notify_all();
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_done.wait(lk, [&] { return pending == 0; });
    }
#endif
    for (int r = 0; r < R; r++)
      if (rc[r]) {
        if (rc[r] == ZLES_E_CUDA) g_cuda_err = "device " + std::to_string(devices[r]) + ": " + err[r];
        return rc[r];
      }
    return 0;
  }

  void worker_main(int r) {
    uint64_t seen = 0;
    zrt_set_device(devices[r]);
    for (;;) {
      std::function<int(int)> fn;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_go.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        fn = task;
      }
      int v = fn(r);
      {
        std::unique_lock<std::mutex> lk(mu);
        rc[r] = v;
        err[r] = g_cuda_err;
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
};

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic pop
#endif

extern "C" int zles_mgpu_create(const int *devices, int count, zles_mgpu **out) {
  if (!devices || count < 1 || count > 64 || !out) return ZLES_E_ARG;
  *out = nullptr;
  zles_mgpu *m = new (std::nothrow) zles_mgpu();
  if (!m) return ZLES_E_NOMEM;
  for (int i = 0; i < count; i++) {
    zles_ctx *c = nullptr;
    int rc = zles_ctx_create(devices[i], &c);
    if (rc) {
      for (zles_ctx *x : m->ctx) zles_ctx_destroy(x);
      delete m;
      return rc;
    }
    m->devices.push_back(devices[i]);
    m->ctx.push_back(c);
  }
  m->rc.assign(count, 0);
  m->err.assign(count, std::string());
  m->slice.resize(count);
  m->range.resize(count);
#ifndef ZLES_EMU
  for (int r = 0; r < count; r++) m->workers.emplace_back([m, r] { m->worker_main(r); });
#endif
  *out = m;
  return 0;
}

extern "C" void zles_mgpu_destroy(zles_mgpu *m) {
  if (!m) return;
  {
    std::unique_lock<std::mutex> lk(m->mu);
    m->stop = true;
  }
  m->cv_go.notify_all();
  for (std::thread &t : m->workers) t.join();
  for (size_t r = 0; r < m->ctx.size(); r++) {
    zrt_set_device(m->devices[r]);
    m->slice[r].release();
    m->range[r].release();
    zles_ctx_destroy(m->ctx[r]);
  }
  delete m;
}

extern "C" int zles_mgpu_device_count(const zles_mgpu *m) { return m ? (int)m->devices.size() : 0; }

extern "C" int zles_mgpu_set_min_shard(zles_mgpu *m, size_t bytes) {
  if (!m) return ZLES_E_ARG;
  m->min_shard = bytes;
  return 0;
}

extern "C" zles_ctx *zles_mgpu_ctx(zles_mgpu *m, int index) {
  if (!m || index < 0 || index >= (int)m->ctx.size()) return nullptr;
  return m->ctx[index];
}

extern "C" uint64_t zles_mgpu_launches(const zles_mgpu *m) {
  uint64_t s = 0;
  if (m) for (const zles_ctx *c : m->ctx) s += c->launches;
  return s;
}

// contiguous shards of whole 128 KiB chunks, the last one takes the ragged tail (zlib.es_b200/dist.py: shard_bounds)
static void mgpu_shards(size_t n, int R, std::vector<size_t> &begin) {
  const u64 nchunks = ((u64)n + CHUNK - 1) / CHUNK;
  begin.resize(R + 1);
  for (int r = 0; r < R; r++) begin[r] = (size_t)umin64((u64)n, (nchunks * (u64)r / (u64)R) * CHUNK);
  begin[R] = n;
}

extern "C" int zles_mgpu_deflate(zles_mgpu *m, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if (!m || (!in && n) || !out_len) return ZLES_E_ARG;
  const int R = (int)m->devices.size();
  if (R == 1 || n < m->min_shard * 2) return zles_deflate(m->ctx[0], in, n, out, cap, out_len);
  int use = R;
  while (use > 1 && n / (size_t)use < m->min_shard) use--;
  std::vector<size_t> begin;
  mgpu_shards(n, use, begin);
  std::vector<zles_shard_info> info(use);
  std::vector<DeflatePipe> pipes(use);
  // phase 1 on every device: copy the shard in (pipelined with the matcher), codes, layout, and the packer into staging
  int rc = m->run([&](int r) -> int {
    if (r >= use) return 0;
    zles_ctx *c = m->ctx[r];
    RET(resolve_ctx(c));
    const size_t a = begin[r], len = begin[r + 1] - a;
    RET(c->d_in.reserve(len + 16));
    RET(c->d_out.reserve(zles_deflate_bound(len) + 16));
    pipes[r].d_out = c->d_out.as<u8>();
    pipes[r].h_out = nullptr;
    pipes[r].h_cap = 0;
    pipes[r].defer = true;
    return deflate_phase1(c, c->d_in.as<u8>(), len, r == use - 1, &info[r], in + a, &pipes[r]);
  });
  if (rc) return rc;
  // the exchange: every shard's place in the stream
  std::vector<size_t> off(use + 1, 0);
  for (int r = 0; r < use; r++) off[r + 1] = off[r] + (size_t)info[r].comp_bytes;
  const size_t need = off[use] + 6;
  *out_len = need;
  if (!out || cap < need) return ZLES_E_OUTPUT_FULL;
  rc = m->run([&](int r) -> int {
    if (r >= use) return 0;
    zles_ctx *c = m->ctx[r];
    RET(resolve_ctx(c));
    if (!pipes[r].done) RET(deflate_phase2(c, c->d_out.as<u8>()));
    if (info[r].comp_bytes) CK(zrt_d2h(out + 2 + off[r], c->d_out.p, (size_t)info[r].comp_bytes, c->stream));
    CK(zrt_sync(c->stream));
    return 0;
  });
  if (rc) return rc;
  put_zlib_header(out);
  put_be32(out + need - 4, zles_adler32_combine_shards(info.data(), (uint32_t)use));
  return 0;
}

extern "C" int zles_mgpu_inflate(zles_mgpu *m, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
  if (!m || (!in && n) || !out_len || (!out && cap)) return ZLES_E_ARG;
  RET(check_zlib_header(in, n));
  const int R = (int)m->devices.size();
  if (R == 1 || n < m->min_shard) return zles_inflate(m->ctx[0], in, n, out, cap, out_len);
  int use = R;
  while (use > 1 && cap / (size_t)use < m->min_shard) use--;
  if (use == 1) return zles_inflate(m->ctx[0], in, n, out, cap, out_len);
  // 1. equal slices of the compressed bytes, each scanned for block markers on its device.  A block start c needs the
  //    four bytes before it, so slice r is staged from 16 bytes before its first position (keeps 16-byte alignment).
  std::vector<size_t> sb(use + 1);
  for (int r = 0; r <= use; r++) sb[r] = r == use ? n : (size_t)(2 + ((u64)(n - 2) * (u64)r / (u64)use)) & ~(size_t)15;
  sb[0] = 0;
  std::vector<std::vector<u64>> found(use);
  std::vector<int> ours(use, 1);
  int rc = m->run([&](int r) -> int {
    if (r >= use) return 0;
    zles_ctx *c = m->ctx[r];
    RET(resolve_ctx(c));
    const size_t lo = r == 0 ? 0 : sb[r] - 16, hi = sb[r + 1];
    RET(m->slice[r].reserve(hi - lo + 16));
    CK(zrt_h2d(m->slice[r].p, in + lo, hi - lo, c->stream));
    // positions are relative to the staged bytes; candidates lie after `first`: 2 for the stream's first block (which the
    // scan always reports as cand[0]), 15 elsewhere (block starts at local 16.. = global sb[r]..; cand[0] is then not a block)
    const int rs = scan_block_starts(c, m->slice[r].as<u8>(), hi - lo, r == 0 ? 2 : 15, found[r]);
    if (rs > 0) return rs;
    if (rs < 0) { ours[r] = 0; return 0; }
    for (u64 &p : found[r]) p += lo;
    if (r > 0) found[r].erase(found[r].begin());
    return 0;
  });
  if (rc) return rc;
  bool ok = true;
  for (int r = 0; r < use; r++) ok = ok && ours[r];
  std::vector<u64> starts;
  if (ok) {
    for (int r = 0; r < use; r++) starts.insert(starts.end(), found[r].begin(), found[r].end());
    for (size_t i = 1; i < starts.size() && ok; i++) ok = starts[i] > starts[i - 1];
  }
  const size_t B = starts.size();
  if (!ok || B < (size_t)use * SUBS_PER_CHUNK * 2 || (u64)(B - 1) * SUB > (u64)cap) return zles_inflate(m->ctx[0], in, n, out, cap, out_len);
  // 2. whole chunks to devices; device r decodes blocks [bb[r], bb[r + 1]) = stream bytes [starts[bb[r]], starts[bb[r + 1]])
  const u64 nchunks = ((u64)B + SUBS_PER_CHUNK - 1) / SUBS_PER_CHUNK;
  std::vector<size_t> bb(use + 1);
  for (int r = 0; r <= use; r++) bb[r] = (size_t)umin64((u64)B, (nchunks * (u64)r / (u64)use) * SUBS_PER_CHUNK);
  std::vector<size_t> got(use, 0);
  std::vector<int> handled(use, 1);
  rc = m->run([&](int r) -> int {
    if (r >= use) return 0;
    zles_ctx *c = m->ctx[r];
    RET(resolve_ctx(c));
    const size_t a = (size_t)starts[bb[r]] & ~(size_t)15, b = r == use - 1 ? n : (size_t)starts[bb[r + 1]];
    const size_t lo = r == 0 ? 0 : sb[r] - 16, hi = sb[r + 1];  // what the device already holds
    RET(m->range[r].reserve(b - a + 32));
    u8 *d = m->range[r].as<u8>();
    // the part it holds moves device to device; the rest comes from the host
    const size_t ia = a > lo ? a : lo, ib = b < hi ? b : hi;
    if (ia < ib) {
      CK(zrt_copy(d + (ia - a), m->slice[r].as<u8>() + (ia - lo), ib - ia, c->stream));
      if (a < ia) CK(zrt_h2d(d, in + a, ia - a, c->stream));
      if (ib < b) CK(zrt_h2d(d + (ib - a), in + ib, b - ib, c->stream));
    } else {
      CK(zrt_h2d(d, in + a, b - a, c->stream));
    }
    std::vector<u64> local(starts.begin() + (long)bb[r], starts.begin() + (long)bb[r + 1]);
    for (u64 &p : local) p -= a;
    const size_t off = bb[r] * (size_t)SUB;
    const size_t room = off >= cap ? 0 : cap - off;
    const int rs = inflate_slabs_to_host(c, d, b - a, local, r == use - 1, out + off, room, &got[r], inflate_slab_size(c, local.size()));
    if (rs < 0) { handled[r] = 0; return 0; }
    if (rs == ZLES_E_OUTPUT_FULL && r == use - 1) return 0;  // reported below with the total size
    return rs;
  });
  if (rc) return rc;
  for (int r = 0; r < use; r++)
    if (!handled[r]) return zles_inflate(m->ctx[0], in, n, out, cap, out_len);
  const size_t total = bb[use - 1] * (size_t)SUB + got[use - 1];
  *out_len = total;
  if (total > cap) return ZLES_E_OUTPUT_FULL;
  for (int r = 0; r + 1 < use; r++)
    if (got[r] != (bb[r + 1] - bb[r]) * (size_t)SUB) return zles_inflate(m->ctx[0], in, n, out, cap, out_len);
  return 0;
}

extern "C" int zles_mgpu_inflate_alloc(zles_mgpu *m, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  if (!m || (!in && n) || !out || !out_len) return ZLES_E_ARG;
  *out = nullptr;
  *out_len = 0;
  // capacity: the reference's own initial guess, 10 x input + slack (src/inflate.ts:17); retried once with the exact size
  size_t cap = n * 10 + CHUNK;
  for (int attempt = 0; attempt < 2; attempt++) {
    u8 *buf = (u8 *)malloc(cap ? cap : 1);
    if (!buf) return ZLES_E_NOMEM;
    size_t need = 0;
    const int rc = zles_mgpu_inflate(m, in, n, buf, cap, &need);
    if (rc == 0) {
      *out = buf;
      *out_len = need;
      return 0;
    }
    free(buf);
    if (rc != ZLES_E_OUTPUT_FULL || attempt == 1) return rc;
    cap = need;
  }
  return ZLES_E_CORRUPTED;
}

// ---- zles_init: the process-wide default used by zles_deflate / zles_inflate(NULL, ...) -------------------------
static zles_mgpu *g_default_mgpu = nullptr;

extern "C" int zles_init(uint32_t device_mask) {
  std::lock_guard<std::mutex> lk(g_default_mu);
  if (g_default_mgpu) { zles_mgpu_destroy(g_default_mgpu); g_default_mgpu = nullptr; }
  if (device_mask == 0) device_mask = 1;
  std::vector<int> devs;
  for (int d = 0; d < 32; d++)
    if (device_mask & (1u << d)) devs.push_back(d);
  if (devs.size() == 1 && devs[0] == 0) return 0;  // the single-device default context serves
  return zles_mgpu_create(devs.data(), (int)devs.size(), &g_default_mgpu);
}

extern "C" void zles_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_default_mu);
  if (g_default_mgpu) { zles_mgpu_destroy(g_default_mgpu); g_default_mgpu = nullptr; }
  if (g_default_ctx) { zles_ctx_destroy(g_default_ctx); g_default_ctx = nullptr; }
}

static zles_mgpu *default_mgpu() {
  std::lock_guard<std::mutex> lk(g_default_mu);
  return g_default_mgpu;
}
