"""BASELINE config 5 on ONE GPU: 8 GiB mixed corpus as a single zlib stream (correctness + throughput)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
c = zles.Codec(0)
st = torch.cuda.Stream(); c.set_stream(st.cuda_stream)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) << 30
with torch.cuda.stream(st):
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
    cap = c.deflate_bound(n); comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); back = torch.empty(n, dtype=torch.uint8, device="cuda")
    def timed(fn):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); r = fn(); e1.record(st); st.synchronize(); return r, e0.elapsed_time(e1)
    clen, td = timed(lambda: c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap))
    olen, ti = timed(lambda: c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n))
    ok = olen == n and torch.equal(src, back)
    adler_ok = c.dev_adler32(src.data_ptr(), n) == int.from_bytes(comp[clen - 4:clen].cpu().numpy().tobytes(), "big")
    print(json.dumps({"config": "mixed %d GiB single stream, 1 GPU" % (n >> 30), "deflate_GBps": round(n / td / 1e6, 2), "inflate_GBps": round(n / ti / 1e6, 2),
                      "ratio": round(n / clen, 4), "roundtrip": bool(ok), "adler_trailer_ok": bool(adler_ok), "peak_mem_GiB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}))
