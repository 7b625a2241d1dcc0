// stager.inl — pageable host buffers (included by zles.cu).
//
// What the N-API addon hands the drop-in calls are plain ArrayBuffers: pageable memory.  A cudaMemcpyAsync from / to
// pageable memory goes through the driver's own bounce buffer — one thread, synchronous, measured on the B200 boxes
// at ~11 GB/s host to device and ~19 GB/s device to host against 55 GB/s for pinned memory — and does not overlap
// with anything.  Large pageable transfers are therefore staged here: a ring of pinned slots per direction, a few
// helper threads that memcpy between the caller's buffer and the slots (several cores copy faster than one), and
// ordinary asynchronous copies between the slots and the device.
//
//   Feeder   host -> device.  Given the whole list of pieces up front (every piece belongs to a group: a slab of the
//            pipelined deflate, a piece of the streaming inflate), the helper threads work through it in order:
//            wait for a free slot, memcpy, enqueue the slot's copy on the copy stream.  wait_group(g) blocks until
//            every piece of groups <= g has been ENQUEUED, so that the caller can record an event behind them.
//   Drainer  device -> host.  push() enqueues device -> slot copies on the output stream (waiting for a free slot)
//            and hands the slot to the helper threads, which wait for the copy and memcpy into the caller's buffer.
//
// Pinned memory (cudaHostAlloc / cudaHostRegister) is detected and never staged.
namespace {

constexpr int STAGE_SLOTS = 8;                     // per direction: 128 MiB of pinned memory per direction and context
constexpr int STAGE_THREADS = 4;                   // helper threads per direction
#ifdef ZLES_EMU
// the emulator tests run the staging logic on small inputs: ZLES_EMU_STAGE=1 shrinks slot and threshold
inline bool stage_small() { static const bool v = [] { const char *e = getenv("ZLES_EMU_STAGE"); return e && *e && *e != '0'; }(); return v; }
#define STAGE_SLOT (stage_small() ? (size_t)40000 : (size_t)16 << 20)
#define STAGE_MIN (stage_small() ? (size_t)100000 : (size_t)32 << 20)
#else
constexpr size_t STAGE_SLOT = (size_t)16 << 20;   // bytes per ring slot
constexpr size_t STAGE_MIN = (size_t)32 << 20;     // smaller pageable transfers take the driver's path
#endif

inline bool host_is_pageable(const void *p) {
#ifdef ZLES_EMU
  static const bool force = [] { const char *e = getenv("ZLES_EMU_STAGE"); return e && *e && *e != '0'; }();
  (void)p;
  return force;
#else
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
#endif
}

struct StageRing {
  u8 *base = nullptr;
  zrt_event_t ev[STAGE_SLOTS];
  bool have = false;
  int ensure() {
    if (have) return 0;
    void *p = nullptr;
    if (zrt_host_alloc(&p, ((size_t)16 << 20) * STAGE_SLOTS) != ZRT_OK) { zrt_last_error(); return 1; }
    base = reinterpret_cast<u8 *>(p);
    for (int i = 0; i < STAGE_SLOTS; i++)
      if (zrt_event_create(&ev[i]) != ZRT_OK) return 1;
    have = true;
    return 0;
  }
  void release() {
    if (!have) return;
    for (int i = 0; i < STAGE_SLOTS; i++) zrt_event_destroy(ev[i]);
    zrt_host_free(base);
    base = nullptr;
    have = false;
  }
  u8 *slot(int i) const { return base + (size_t)i * ((size_t)16 << 20); }
};

struct StagePiece {
  const u8 *h;   // caller's memory
  u8 *d;         // device memory
  size_t len;
  u32 group;
};

class Feeder {
 public:
  // pieces in the order they are needed; groups ascend
  int start(int device, StageRing *ring, zrt_stream_t copy_stream, std::vector<StagePiece> pieces, u32 ngroups) {
    if (ring->ensure()) return 1;
    ring_ = ring;
    device_ = device;
    stream_ = copy_stream;
    pieces_ = std::move(pieces);
    left_.assign(ngroups, 0);
    for (const StagePiece &p : pieces_) left_[p.group]++;
    next_ = 0;
    enq_.assign(pieces_.size(), 0);
    err_ = false;
    const int nt = (int)std::min<size_t>(STAGE_THREADS, pieces_.size());
#ifdef ZLES_EMU
    (void)nt;
    run();  // the emulator build stages inline
#else
    for (int t = 0; t < nt; t++) threads_.emplace_back([this] { run(); });
#endif
    return 0;
  }
  // every piece of groups <= g has been enqueued on the copy stream (or an error occurred: returns false)
  bool wait_group(u32 g) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] {
      if (err_) return true;
      for (u32 k = 0; k <= g && k < left_.size(); k++)
        if (left_[k]) return false;
      return true;
    });
    return !err_;
  }
  bool finish() {
    for (std::thread &t : threads_) t.join();
    threads_.clear();
    return !err_;
  }
  ~Feeder() { finish(); }

 private:
  void run() {
    zrt_set_device(device_);
    for (;;) {
      size_t i;
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (err_ || next_ >= pieces_.size()) return;
        i = next_++;
      }
      const int s = (int)(i % STAGE_SLOTS);
      // the slot's previous copy (piece i - STAGE_SLOTS) must have left it; pieces are taken in order, so that piece has
      // been claimed — wait until its thread has recorded the event, then for the event
      if (i >= (size_t)STAGE_SLOTS) {
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return err_ || enq_[i - STAGE_SLOTS]; });
          if (err_) return;
        }
        if (zrt_event_sync(ring_->ev[s]) != ZRT_OK) { fail(); return; }
      }
      const StagePiece &p = pieces_[i];
      memcpy(ring_->slot(s), p.h, p.len);
      zrt_err_t e = zrt_h2d(p.d, ring_->slot(s), p.len, stream_);
      if (e == ZRT_OK) e = zrt_event_record(ring_->ev[s], stream_);
      if (e != ZRT_OK) { fail(); return; }
      {
        std::unique_lock<std::mutex> lk(mu_);
        enq_[i] = true;
        left_[p.group]--;
      }
      cv_.notify_all();
    }
  }
  void fail() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      err_ = true;
    }
    cv_.notify_all();
  }
  StageRing *ring_ = nullptr;
  int device_ = 0;
  zrt_stream_t stream_{};
  std::vector<StagePiece> pieces_;
  std::vector<u32> left_;
  std::vector<char> enq_;  // piece i's copy has been enqueued
  std::mutex mu_;
  std::condition_variable cv_;
  size_t next_ = 0;
  bool err_ = false;
  std::vector<std::thread> threads_;
};

class Drainer {
 public:
  int start(int device, StageRing *ring, zrt_stream_t out_stream) {
    if (ring->ensure()) return 1;
    ring_ = ring;
    device_ = device;
    stream_ = out_stream;
    free_.assign(STAGE_SLOTS, true);
    closing_ = false;
    err_ = false;
    pending_ = 0;
#ifndef ZLES_EMU
    for (int t = 0; t < STAGE_THREADS; t++) threads_.emplace_back([this] { run(); });
#endif
    return 0;
  }
  // device bytes [d, d + len) -> caller's memory h, in pieces; returns false on a CUDA error
  bool push(u8 *h, const u8 *d, size_t len) {
    for (size_t o = 0; o < len; o += STAGE_SLOT) {
      const size_t m = std::min<size_t>(STAGE_SLOT, len - o);
      int s = -1;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
          if (err_) return true;
          for (int k = 0; k < STAGE_SLOTS; k++)
            if (free_[k]) { s = k; return true; }
          return false;
        });
        if (err_) return false;
        free_[s] = false;
      }
      zrt_err_t e = zrt_d2h(ring_->slot(s), d + o, m, stream_);
      if (e == ZRT_OK) e = zrt_event_record(ring_->ev[s], stream_);
      if (e != ZRT_OK) { fail(); return false; }
      {
        std::unique_lock<std::mutex> lk(mu_);
        jobs_.push_back(Job{s, h + o, m});
        pending_++;
      }
      cv_.notify_all();
#ifdef ZLES_EMU
      drain_one();
#endif
    }
    return true;
  }
  // all pushed bytes are in the caller's memory
  bool finish() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return err_ || pending_ == 0; });
      closing_ = true;
    }
    cv_.notify_all();
    for (std::thread &t : threads_) t.join();
    threads_.clear();
    return !err_;
  }
  bool active() const { return ring_ != nullptr; }
  ~Drainer() { if (ring_) finish(); }

 private:
  struct Job { int slot; u8 *h; size_t len; };
  bool drain_one() {
    Job j;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return closing_ || err_ || !jobs_.empty(); });
      if (jobs_.empty()) return false;
      j = jobs_.front();
      jobs_.pop_front();
    }
    if (zrt_event_sync(ring_->ev[j.slot]) != ZRT_OK) { fail(); return false; }
    memcpy(j.h, ring_->slot(j.slot), j.len);
    {
      std::unique_lock<std::mutex> lk(mu_);
      free_[j.slot] = true;
      pending_--;
    }
    cv_.notify_all();
    return true;
  }
  void run() {
    zrt_set_device(device_);
    while (drain_one()) {}
  }
  void fail() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      err_ = true;
      closing_ = true;
    }
    cv_.notify_all();
  }
  StageRing *ring_ = nullptr;
  int device_ = 0;
  zrt_stream_t stream_{};
  std::vector<bool> free_;
  std::deque<Job> jobs_;
  std::mutex mu_;
  std::condition_variable cv_;
  size_t pending_ = 0;
  bool closing_ = false, err_ = false;
  std::vector<std::thread> threads_;
};

}  // namespace
