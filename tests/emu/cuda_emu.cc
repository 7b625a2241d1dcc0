// cuda_emu.cc — fiber scheduler of the CPU thread emulator (TESTS ONLY; see cuda_emu.h).
#include "cuda_emu.h"

#include <sys/mman.h>

#include <thread>

namespace emu {

thread_local BlockCtx *g_blk = nullptr;

// Minimal x86-64 System V context switch: saves the callee-saved registers on the
// current stack, stores the stack pointer through `from`, loads `to`.
extern "C" void emu_switch(void **from, void *to);
__asm__(
    ".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n.size emu_switch,.-emu_switch\n");

static const size_t kStack = 256 * 1024;

static void fiber_main() {
  BlockCtx *b = g_blk;
  (*b->body)();
  b->fibers[b->cur].done = true;
  void *dummy;
  emu_switch(&dummy, b->sched_sp);  // never returns
  abort();
}

void yield() {
  BlockCtx *b = g_blk;
  emu_switch(&b->fibers[b->cur].sp, b->sched_sp);
}

static void prepare(Fiber &f) {
  // initial frame: six callee-saved registers, then the return address = fiber_main.
  // At fiber_main's entry rsp must be 8 modulo 16 (as after a call).
  uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
  uint64_t *sp = (uint64_t *)(top - 8);   // slot that a caller's return address would occupy
  *--sp = (uint64_t)(uintptr_t)&fiber_main;
  for (int i = 0; i < 6; i++) *--sp = 0;
  f.sp = sp;
  f.done = false;
}

static void run_block(BlockCtx *b, size_t smem_bytes) {
  g_blk = b;
  memset(b->smem, 0xCD, smem_bytes);  // shared memory starts undefined
  b->barrier_arrived = 0;
  for (auto &w : b->warps) w = WarpSlot();
  for (unsigned t = 0; t < b->nthreads; t++) prepare(b->fibers[t]);
  unsigned live = b->nthreads;
  unsigned long idle_rounds = 0;
  while (live) {
    unsigned progressed = 0;
    for (unsigned t = 0; t < b->nthreads; t++) {
      Fiber &f = b->fibers[t];
      if (f.done) continue;
      b->cur = t;
      emu_switch(&b->sched_sp, f.sp);
      if (f.done) { live--; progressed++; }
    }
    (void)progressed;
    if (++idle_rounds > (1ul << 34)) { fprintf(stderr, "emu: scheduler livelock\n"); abort(); }
  }
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  const unsigned T = block.x * block.y * block.z;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  if (T == 0 || nblocks == 0) return;
  unsigned workers = std::thread::hardware_concurrency();
  if (workers < 1) workers = 1;
  if (workers > 8) workers = 8;
  if (workers > nblocks) workers = nblocks;
  std::atomic<unsigned> next_block{0};
  auto worker = [&]() {
    BlockCtx ctx;
    ctx.nthreads = T;
    ctx.fibers.resize(T);
    ctx.warps.resize((T + 31) / 32);
    ctx.bdim = block;
    ctx.gdim = grid;
    ctx.body = &body;
    const size_t smem_alloc = ((smem_bytes + 1023) / 1024 + 1) * 1024;
    ctx.smem = (uint8_t *)aligned_alloc(1024, smem_alloc);
    char *stacks = (char *)mmap(nullptr, kStack * T, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == (char *)MAP_FAILED) { perror("emu: mmap"); abort(); }
    for (unsigned t = 0; t < T; t++) {
      ctx.fibers[t].stack = stacks + (size_t)t * kStack;
      ctx.fibers[t].tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
    }
    for (;;) {
      unsigned bidx = next_block.fetch_add(1);
      if (bidx >= nblocks) break;
      ctx.bid = {bidx % grid.x, (bidx / grid.x) % grid.y, bidx / (grid.x * grid.y)};
      run_block(&ctx, smem_bytes);
    }
    munmap(stacks, kStack * T);
    free(ctx.smem);
    g_blk = nullptr;
  };
  if (workers == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (unsigned i = 0; i < workers; i++) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
}

}  // namespace emu
