"""Throughput of every BASELINE.json config on one GPU (device-resident, CUDA-event timed, best of 3)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
c = zles.Codec(0)
st = torch.cuda.Stream(); c.set_stream(st.cuda_stream)
def timed(fn, reps=3):
    best = 1e9; r = None
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); r = fn(); e1.record(st); st.synchronize(); best = min(best, e0.elapsed_time(e1))
    return r, best
with torch.cuda.stream(st):
    for label, kind, n in [("config2 text 64MiB", 0, 64 << 20), ("binary 64MiB", 1, 64 << 20), ("random 64MiB", 2, 64 << 20), ("config4 mixed 1GiB", 3, 1 << 30), ("text 1GiB", 0, 1 << 30)]:
        src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(kind, 0, src.data_ptr(), n)
        cap = c.deflate_bound(n); comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); back = torch.empty(n, dtype=torch.uint8, device="cuda")
        clen, td = timed(lambda: c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap))
        olen, ti = timed(lambda: c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n))
        ok = olen == n and torch.equal(src, back)
        print(json.dumps({"config": label, "deflate_GBps": round(n / td / 1e6, 2), "inflate_GBps": round(n / ti / 1e6, 2), "ratio": round(n / clen, 4), "ok": bool(ok)}), flush=True)
        del src, comp, back
    # config 3: 262,144 x 4 KiB buffers of the mixed corpus, each its own zlib stream
    count = 262144; n = count * 4096
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
    in_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * 4096
    bound = c.deflate_bound(4096)
    out_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * bound
    out = torch.empty(count * bound, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(count, dtype=torch.int64, device="cuda"); status = torch.zeros(count, dtype=torch.int32, device="cuda")
    st.synchronize()
    rc, td = timed(lambda: c.dev_deflate_batch(src.data_ptr(), in_off.data_ptr(), count, out.data_ptr(), out_off.data_ptr(), out_len.data_ptr(), status.data_ptr()))
    back = torch.zeros(n, dtype=torch.uint8, device="cuda"); blen = torch.zeros(count, dtype=torch.int64, device="cuda"); st2 = torch.zeros(count, dtype=torch.int32, device="cuda")
    st.synchronize()
    rc2, ti = timed(lambda: c.dev_inflate_batch(out.data_ptr(), out_off.data_ptr(), count, back.data_ptr(), in_off.data_ptr(), blen.data_ptr(), st2.data_ptr()))
    st.synchronize()
    print(json.dumps({"config": "config3 262144 x 4KiB mixed", "deflate_GBps": round(n / td / 1e6, 2), "inflate_GBps": round(n / ti / 1e6, 2),
                      "ratio": round(n / int(out_len.sum()), 4), "ok": bool(rc == 0 and rc2 == 0 and torch.equal(src, back))}), flush=True)
