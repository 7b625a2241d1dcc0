"""The N > 1 path on the CPU: world_size 2 and 3 over gloo, the emulator library standing in
for the GPU and shared memory for CUDA IPC.  Checks the one exchange step (all-gather of shard
sizes), the offsets every rank derives from it, the peer writes into the final stream and the
sharded inflate."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import emu_lib
    import vectors as T
    from shm_transport import ShmTransport
    import zles
    from zles import dist as zdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = emu_lib.codec()
        data = (T.gen("G5", total // 2) + T.fixture_raw())[:total]
        a, b = zdist.shard_bounds(total, world)[rank]
        shard = np.frombuffer(data[a:b], dtype=np.uint8).copy()
        sc = zdist.ShardedCodec(c, ShmTransport(), rank, world, comm_device="cpu")
        sc.setup(c.deflate_bound(total) + 64 * world)
        lay = sc.deflate(shard.ctypes.data if shard.size else 0, shard.size)
        dist.barrier()
        if rank == 0:
            stream = bytes((np.ctypeslib.as_array((__import__("ctypes").c_uint8 * lay.total_comp).from_address(sc.t.base))))
            assert zlib.decompress(stream) == data
            assert stream == c.deflate(data)  # sharding does not change a single bit of the stream
            assert lay.adler == zlib.adler32(data)
        stage = np.zeros(lay.comp[rank] + 16, dtype=np.uint8)
        out = np.zeros(b - a + 16, dtype=np.uint8)
        n = sc.inflate(stage.ctypes.data, out.ctypes.data, b - a)
        assert n == b - a and out[:n].tobytes() == data[a:b]
        dist.barrier()
        # the same stream decoded WITHOUT the encoder's layout: shards found from the bytes (marker scan + all-gather).
        # The stream in the shared buffer is replaced by the single-call one first (made by "a different world size").
        if rank == 0:
            whole = np.frombuffer(c.deflate(data), dtype=np.uint8)
            assert whole.size == lay.total_comp
            __import__("ctypes").memmove(sc.t.base, whole.ctypes.data, whole.size)
        dist.barrier()
        cap = c.deflate_bound(total) + 64
        sl = np.zeros(cap, dtype=np.uint8)
        rg = np.zeros(cap, dtype=np.uint8)
        out2 = np.zeros(b - a + 16, dtype=np.uint8)
        n2 = sc.inflate_from_stream(lay.total_comp, sl.ctypes.data, rg.ctypes.data, cap, out2.ctypes.data, b - a)
        if total >= 131072 * world:
            assert n2 == b - a and out2[:n2].tobytes() == data[a:b], (rank, n2, b - a)
        else:  # fewer chunks than ranks: the chunks go where shard_bounds over the BLOCK count puts them
            got = __import__("torch").tensor([n2])
            dist.all_reduce(got)
            assert int(got) == total
        dist.barrier()
        sc.teardown()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.emu
@pytest.mark.parametrize("world,total", [(2, 300000), (3, 131072 * 3 + 5), (2, 100)])
def test_sharded_stream_over_gloo(world, total):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_lib
    emu_lib.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_shard_bounds():
    sys.path.insert(0, ROOT)
    import zles
    from zles import dist as zdist
    for total in (0, 1, 131072, 131073, 10 * 131072 + 7, 8 << 30):
        for world in (1, 2, 3, 4, 8):
            b = zdist.shard_bounds(total, world)
            assert b[0][0] == 0 and b[-1][1] == total
            for (a0, b0), (a1, b1) in zip(b[:-1], b[1:]):
                assert b0 == a1 and b0 % 131072 == 0 or b0 == total
