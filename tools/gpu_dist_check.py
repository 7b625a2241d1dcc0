"""torchrun check of the sharded path on real GPUs: the stream assembled on rank 0 over NCCL + CUDA IPC
must be a valid zlib stream that system zlib inflates to the concatenation of all shards, and be
bit-identical to the single-GPU deflate of the same bytes."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist, zles
from zles import dist as zdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
c = zles.Codec(local)
total = int(sys.argv[1]) if len(sys.argv) > 1 else (48 << 20) + 12345
kind = 3
a, b = zdist.shard_bounds(total, world)[rank]
src = torch.empty(max(1, b - a), dtype=torch.uint8, device="cuda"); c.dev_corpus(kind, a, src.data_ptr(), b - a)
sc = zdist.ShardedCodec(c, zdist.IpcTransport(c), rank, world)
sc.setup(c.deflate_bound(total) + 64 * world)
lay = sc.deflate(src.data_ptr(), b - a)
dist.barrier()
ok = True
if rank == 0:
    host = np.empty(lay.total_comp, dtype=np.uint8)
    c.dev_copy(host.ctypes.data, sc.t.base, lay.total_comp)
    stream = host.tobytes()
    whole = c.host_corpus(kind, 0, total).tobytes()
    ok = zlib.decompress(stream) == whole
    single = c.deflate(whole)
    same = single == stream
    print("rank0: stream %d bytes, zlib ok=%s, identical to single-GPU deflate=%s, adler ok=%s" % (len(stream), ok, same, lay.adler == zlib.adler32(whole)), flush=True)
    ok = ok and same
stage = torch.empty(lay.comp[rank] + 16, dtype=torch.uint8, device="cuda"); back = torch.zeros(max(1, b - a), dtype=torch.uint8, device="cuda")
n = sc.inflate(stage.data_ptr(), back.data_ptr(), b - a)
ok2 = n == b - a and torch.equal(src[: b - a], back[: b - a])
print("rank %d: shard [%d,%d) inflate ok=%s" % (rank, a, b, ok2), flush=True)
flag = torch.tensor([0 if (ok and ok2) else 1], device="cuda"); dist.all_reduce(flag)
sc.teardown(); dist.destroy_process_group()
sys.exit(int(flag.item()) != 0)
