"""Sharded deflate / inflate over the GPUs of one node (one process per GPU).

The reference has no notion of more than one thread; the shape below is SURVEY.md §8(e):

deflate   rank r compresses a contiguous shard of the input (a multiple of 128 KiB except on
          the last rank) with no communication (phase 1: match finder, code construction,
          layout).  ONE exchange: an all-gather of 5 x u64 per rank (compressed size, raw size,
          Adler-32 partial sums, block count) over NCCL.  Every rank then knows every shard's
          offset in the final stream and phase 2 (the bit packer) stores its blocks straight
          into the destination buffer on the owning rank through a peer-mapped pointer
          (CUDA IPC; NVLink P2P stores issued by k_pack itself).  The owner adds the zlib
          header and the combined Adler-32 trailer (/root/reference/src/zlib.ts:28-46).
inflate   rank r pulls its shard's compressed bytes from the owner (peer read), decodes them
          (blocks are byte aligned and chunk-independent) into its local output shard.

The transport that turns "a buffer on rank 0" into a pointer usable by every rank is pluggable:
``IpcTransport`` (CUDA IPC) is the product path; the CPU tests plug in a shared-memory one.
"""
from __future__ import annotations

import contextlib
import ctypes
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from .codec import Codec, combine_adler

CHUNK = 131072


def shard_bounds(total: int, world: int) -> list[tuple[int, int]]:
    """Contiguous shards, whole 128 KiB chunks each (the last one takes the ragged tail)."""
    nchunks = (total + CHUNK - 1) // CHUNK
    out = []
    for r in range(world):
        c0 = nchunks * r // world
        c1 = nchunks * (r + 1) // world
        a = min(total, c0 * CHUNK)
        b = total if r == world - 1 else min(total, c1 * CHUNK)
        out.append((a, b))
    return out


class IpcTransport:
    """Destination buffer in rank 0's HBM, mapped into every rank with CUDA IPC."""

    def __init__(self, codec: Codec):
        self.codec = codec
        self.base = 0
        self.owner = False

    def create(self, nbytes: int) -> bytes:
        self.base = self.codec.dev_alloc(nbytes)
        self.owner = True
        return self.codec.ipc_export(self.base)

    def open(self, handle: bytes):
        self.base = self.codec.ipc_open(handle)

    def close(self):
        if not self.base:
            return
        if self.owner:
            self.codec.dev_free(self.base)
        else:
            self.codec.ipc_close(self.base)
        self.base = 0


@dataclass
class ShardLayout:
    offsets: list[int]      # byte offset of each rank's blocks inside the raw deflate data
    comp: list[int]         # compressed bytes per rank
    raw: list[int]          # raw bytes per rank
    total_comp: int         # whole zlib stream length (header + data + trailer)
    adler: int


class ShardedCodec:
    def __init__(self, codec: Codec, transport, rank: int | None = None, world: int | None = None, group=None,
                 comm_device: str | torch.device = "cuda"):
        self.c = codec
        self.t = transport
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.comm_device = torch.device(comm_device)
        self.capacity = 0
        self.layout: ShardLayout | None = None
        cuda = self.comm_device.type == "cuda"
        # persistent exchange buffers: pinned on the host side, so that nothing but the one read-back waits
        self._h_mine = torch.empty(5, dtype=torch.int64, pin_memory=cuda)
        self._h_all = torch.empty(self.world * 5, dtype=torch.int64, pin_memory=cuda)
        self._d_mine = torch.empty(5, dtype=torch.int64, device=self.comm_device)
        self._d_all = torch.empty(self.world * 5, dtype=torch.int64, device=self.comm_device)
        self._d_flag = torch.zeros(1, dtype=torch.int32, device=self.comm_device)
        self._frame = torch.empty(8, dtype=torch.uint8, pin_memory=cuda)
        # numpy views of the pinned buffers: one assignment instead of a tensor operation per element
        self._np_mine = self._h_mine.numpy()
        self._np_all = self._h_all.numpy()
        self._np_frame = self._frame.numpy()

    def _ordered(self):
        """Collectives issued inside this context are ordered on the codec's stream (when it runs on a caller's CUDA
        stream): no host synchronisation is needed between the codec's kernels and the exchange."""
        if self.comm_device.type == "cuda" and getattr(self.c, "stream_ptr", None):
            return torch.cuda.stream(torch.cuda.ExternalStream(self.c.stream_ptr, device=self.comm_device))
        return contextlib.nullcontext()

    def _rendezvous(self):
        """Every rank's work issued so far on the codec's stream is complete before anything issued after this, on any
        rank (what the packer's peer stores need before the stream is read)."""
        if self.comm_device.type == "cuda" and getattr(self.c, "stream_ptr", None):
            with self._ordered():
                dist.all_reduce(self._d_flag, group=self.group)  # stream-ordered: nobody's stream passes before everybody's arrives
        else:
            self.c.sync()
            dist.barrier(group=self.group)

    # -- destination buffer -----------------------------------------------------------------
    def setup(self, capacity: int):
        """Allocate the final-stream buffer on rank 0 and map it everywhere (once, outside any timing)."""
        self.capacity = capacity
        box = [None]
        if self.rank == 0:
            box[0] = self.t.create(capacity)
        dist.broadcast_object_list(box, src=0, group=self.group)
        if self.rank != 0:
            self.t.open(box[0])
        dist.barrier(group=self.group)

    def teardown(self):
        dist.barrier(group=self.group)
        if self.rank != 0:
            self.t.close()
        dist.barrier(group=self.group)
        if self.rank == 0:
            self.t.close()

    # -- deflate ------------------------------------------------------------------------------
    def deflate(self, d_in: int, n_local: int) -> ShardLayout:
        """All ranks call this with their shard (device pointer).  Afterwards the zlib stream lies at
        ``self.t.base`` on rank 0 (``layout.total_comp`` bytes)."""
        info = self.c.dev_deflate_phase1(d_in, n_local, self.rank == self.world - 1)
        self._np_mine[:] = (info.comp_bytes, info.raw_bytes, info.adler_a, info.adler_b, info.n_blocks)
        with self._ordered():
            self._d_mine.copy_(self._h_mine, non_blocking=True)
            dist.all_gather_into_tensor(self._d_all, self._d_mine, group=self.group)  # the one exchange step of the path
            self._h_all.copy_(self._d_all, non_blocking=True)
            if self.comm_device.type == "cuda":
                torch.cuda.current_stream(self.comm_device).synchronize()
        rows = self._np_all.reshape(self.world, 5).tolist()
        infos = []
        offsets, comp, raw = [], [], []
        off = 0
        for cb, rb, a, b, nb in rows:
            si = _capi.ShardInfo()
            si.comp_bytes, si.raw_bytes, si.adler_a, si.adler_b, si.n_blocks = int(cb), int(rb), int(a), int(b), int(nb)
            infos.append(si)
            offsets.append(off)
            comp.append(int(cb))
            raw.append(int(rb))
            off += int(cb)
        total = off + 6
        if total > self.capacity:
            raise RuntimeError("destination buffer too small: need %d, have %d" % (total, self.capacity))
        adler = combine_adler(infos, self.c.L)
        # phase 2: the packer stores this shard's blocks at their global offset — local memory on rank 0,
        # peer-mapped memory (NVLink) elsewhere
        self.c.dev_deflate_phase2(self.t.base + 2 + offsets[self.rank])
        if self.rank == 0:  # framing stays on the host (/root/reference/src/zlib.ts:28-46)
            self._np_frame[:6] = (0x78, 0x9C, (adler >> 24) & 255, (adler >> 16) & 255, (adler >> 8) & 255, adler & 255)
            self.c.dev_copy_async(self.t.base, self._frame.data_ptr(), 2)
            self.c.dev_copy_async(self.t.base + total - 4, self._frame.data_ptr() + 2, 4)
        self._rendezvous()
        self.layout = ShardLayout(offsets, comp, raw, total, adler)
        return self.layout

    # -- inflate ------------------------------------------------------------------------------
    def inflate(self, d_stage: int, d_out: int, cap: int, layout: ShardLayout | None = None) -> int:
        """Rank r decodes shard r of the stream at ``self.t.base`` into d_out; d_stage is local scratch of at
        least layout.comp[rank] bytes.  Returns the decoded length."""
        lay = layout or self.layout
        n = lay.comp[self.rank]
        self.c.dev_copy_async(d_stage, self.t.base + 2 + lay.offsets[self.rank], n)  # peer read of this shard's bytes
        return self.c.dev_inflate_segment(d_stage, n, d_out, cap, has_final=self.rank == self.world - 1)

    def inflate_from_stream(self, total_comp: int, d_slice: int, d_range: int, stage_cap: int, d_out: int, cap: int) -> int:
        """The same without any knowledge of how the stream was made (a consumer that only holds the zlib bytes,
        /root/reference/src/zlib.ts:11-23): the stream of ``total_comp`` bytes lies at ``self.t.base`` on rank 0.  Every rank
        pulls an equal slice of its bytes and scans it for block markers; ONE all-gather of the block starts; block j stands
        for output bytes [32 KiB j, 32 KiB (j + 1)) and chunks of four blocks are independent, so rank r takes the chunks
        ``shard_bounds`` gives it, completes the bytes of its range it does not hold yet (peer read) and decodes them into
        d_out.  d_slice / d_range: local scratch of stage_cap bytes each.  Returns the decoded length of this rank's shard.
        Works for a stream made by any number of ranks (or by the single-device call)."""
        n, R, r = total_comp, self.world, self.rank
        sb = [0] + [((2 + (n - 2) * k // R) & ~15) for k in range(1, R)] + [n]
        lo, hi = (0 if r == 0 else sb[r] - 16), sb[r + 1]
        if hi - lo > stage_cap:
            raise RuntimeError("scratch too small for the slice: need %d, have %d" % (hi - lo, stage_cap))
        self.c.dev_copy_async(d_slice, self.t.base + lo, hi - lo)
        mine = self.c.dev_scan_blocks(d_slice, hi - lo, 2 if r == 0 else 15)
        mine = (mine[1:] if r else mine) + np.uint64(lo)  # cand[0] is the scan's `first`: a block only for the stream's first slice
        # the one exchange: counts, then the (padded) lists
        dev = self.comm_device
        with self._ordered():
            cnt = torch.tensor([len(mine)], dtype=torch.int64, device=dev)
            cnts = torch.empty(R, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(cnts, cnt, group=self.group)
            counts = cnts.cpu().tolist()
            width = max(1, max(counts))
            pad = np.zeros(width, dtype=np.int64)
            pad[:len(mine)] = mine.astype(np.int64)
            allpos = torch.empty(R * width, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allpos, torch.from_numpy(pad).to(dev), group=self.group)
            allpos = allpos.cpu().numpy().reshape(R, width)
        starts = np.concatenate([allpos[k, :counts[k]] for k in range(R)]).astype(np.uint64)
        B = len(starts)
        if B == 0 or (B > 1 and not bool((starts[1:] > starts[:-1]).all())):
            raise RuntimeError("not one of our streams")
        nchunks = (B + 3) // 4
        bb = [min(B, (nchunks * k // R) * 4) for k in range(R)] + [B]
        if bb[r] == bb[r + 1]:
            return 0
        a = int(starts[bb[r]]) & ~15
        b = n if r == R - 1 else int(starts[bb[r + 1]])
        if b - a > stage_cap:
            raise RuntimeError("scratch too small for the range: need %d, have %d" % (b - a, stage_cap))
        ia, ib = max(a, lo), min(b, hi)
        if ia < ib:  # what the slice already holds moves locally; the rest is read from the owner
            self.c.dev_copy_async(d_range + (ia - a), d_slice + (ia - lo), ib - ia)
            if a < ia:
                self.c.dev_copy_async(d_range, self.t.base + a, ia - a)
            if ib < b:
                self.c.dev_copy_async(d_range + (ib - a), self.t.base + ib, b - ib)
        else:
            self.c.dev_copy_async(d_range, self.t.base + a, b - a)
        first = int(starts[bb[r]]) - a
        return self.c.dev_inflate_segment(d_range + first, b - a - first, d_out, cap, has_final=r == R - 1)
