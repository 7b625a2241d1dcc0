"""BASELINE.json configs[4]: the mixed corpus sharded over the GPUs of one box into ONE zlib stream (torchrun).

usage: python -m torch.distributed.run --nproc-per-node N tools/gpu_config5.py [GiB total = 8] [reps = 3]
Every rank generates its contiguous shard of the corpus on its GPU, then: sharded deflate (phase 1, NCCL all-gather of
the shard sizes, packer storing into rank 0's buffer over NVLink), sharded inflate (every rank peer-reads and decodes
its shard).  Timed on the device with CUDA events, barrier on both sides, max over ranks, best of `reps`.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist, zles
from zles import dist as zdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
c = zles.Codec(local)
st = torch.cuda.Stream(); c.set_stream(st.cuda_stream)
total = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 8 << 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a, b = zdist.shard_bounds(total, world)[rank]
with torch.cuda.stream(st):
    src = torch.empty(b - a, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, a, src.data_ptr(), b - a)
    back = torch.empty(b - a, dtype=torch.uint8, device="cuda")
    sc = zdist.ShardedCodec(c, zdist.IpcTransport(c), rank, world)
    sc.setup(c.deflate_bound(total) + 64 * world)
    stage = None
    best_d = best_i = 1e9
    for _ in range(reps + 1):  # first pass = warm-up (allocations)
        st.synchronize(); dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        lay = sc.deflate(src.data_ptr(), b - a)
        e1.record(st); st.synchronize(); dist.barrier()
        if stage is None: stage = torch.empty(max(lay.comp) + 16, dtype=torch.uint8, device="cuda")
        st.synchronize(); dist.barrier()
        e1b = torch.cuda.Event(enable_timing=True); e1b.record(st)
        n = sc.inflate(stage.data_ptr(), back.data_ptr(), b - a)
        e2.record(st); st.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), e1b.elapsed_time(e2)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if _ > 0:
            best_d = min(best_d, float(t[0])); best_i = min(best_i, float(t[1]))
    ok = torch.tensor([1 if (n == b - a and torch.equal(src, back)) else 0], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"config": "mixed corpus %.2f GiB over %d GPU(s), one zlib stream" % (total / (1 << 30), world), "deflate_GBps": round(total / best_d / 1e6, 2),
                      "inflate_GBps": round(total / best_i / 1e6, 2), "ratio": round(total / lay.total_comp, 4), "roundtrip_ok": bool(ok.item()),
                      "deflate_ms": round(best_d, 2), "inflate_ms": round(best_i, 2)}), flush=True)
sc.teardown(); dist.destroy_process_group()
