"""One deflate + one inflate of a synthetic corpus (for ncu): python tools/prof_step.py [kind] [MiB] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zles
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
c = zles.Codec(0)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
c.dev_corpus(kind, 0, src.data_ptr(), n)
cap = c.deflate_bound(n)
comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
    olen = c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n)
assert olen == n and torch.equal(src, back)
print("ok", n, clen, c.launches)
