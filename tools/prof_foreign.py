"""One inflate of a zlib.es-made stream and of a system-zlib stream (for ncu: the foreign tier's kernels)."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, zles, oracle as O
c = zles.Codec(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 16) << 20
raw = c.host_corpus(3, 0, n).tobytes()
for z in (O.deflate(raw), zlib.compress(raw, 6)):
    d_in = torch.frombuffer(bytearray(z), dtype=torch.uint8).cuda(); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        m = c.dev_inflate(d_in.data_ptr(), len(z), d_out.data_ptr(), n)
    torch.cuda.synchronize()
    assert m == n and d_out.cpu().numpy().tobytes() == raw
print("ok")
