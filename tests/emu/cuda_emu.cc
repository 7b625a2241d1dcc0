// cuda_emu.cc — launch loop of the CPU thread emulator (TESTS ONLY; see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

thread_local BlockCtx *g_blk = nullptr;
thread_local uint3 g_tid = {0, 0, 0}, g_bid = {0, 0, 0};
thread_local dim3 g_bdim, g_gdim;

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  const unsigned T = block.x * block.y * block.z;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  if (T == 0 || nblocks == 0) return;
  unsigned groups = 2048 / T;
  if (groups < 1) groups = 1;
  if (groups > 16) groups = 16;
  if (groups > nblocks) groups = nblocks;
  std::atomic<unsigned> next_block{0};
  std::vector<std::thread> all;
  std::vector<BlockCtx *> ctxs;
  for (unsigned g = 0; g < groups; g++) {
    BlockCtx *ctx = new BlockCtx();
    ctx->nthreads = T;
    ctx->warps = std::vector<WarpRv>((T + 31) / 32);
    ctx->smem = (uint8_t *)aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
    ctxs.push_back(ctx);
    // per-group shared "current block" variable, written by thread 0 between barriers
    unsigned *cur = new unsigned(0);
    for (unsigned t = 0; t < T; t++) {
      all.emplace_back([=, &next_block, &body] {
        g_blk = ctx;
        g_bdim = block;
        g_gdim = grid;
        g_tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        for (;;) {
          if (t == 0) {
            *cur = next_block.fetch_add(1);
            memset(ctx->smem, 0xCD, smem_bytes);  // shared memory starts undefined
          }
          block_barrier();
          unsigned b = *cur;
          if (b >= nblocks) break;
          g_bid = {b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
          body();
          block_barrier();
        }
      });
    }
  }
  for (auto &th : all) th.join();
  for (auto *c : ctxs) {
    free(c->smem);
    delete c;
  }
}

}  // namespace emu
