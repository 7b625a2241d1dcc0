"""Size and speed of deflate at different search depths (zles_ctx_set_level) on the synthetic corpora."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, zles
c = zles.Codec(0)
oracle8 = {"text": 3405877, "binary": 2629584, "mixed": 3342227}
n = 64 << 20
for kind, label in [(0, "text"), (1, "binary"), (3, "mixed")]:
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(kind, 0, src.data_ptr(), n)
    cap = c.deflate_bound(n); comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    for scan, deep, lazy in [(32, 4, 1), (24, 4, 1), (16, 4, 1), (32, 2, 1), (32, 1, 1), (24, 2, 1), (8, 2, 1), (32, 4, 0)]:
        c.set_level(scan, deep, 8, bool(lazy))
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t = time.perf_counter(); clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        s8 = c.dev_deflate(src.data_ptr(), 8 << 20, comp.data_ptr(), cap)
        print(label, "scan", scan, "deep", deep, "lazy", lazy, "GB/s %.2f" % (n / best / 1e9), "ratio %.4f" % (n / clen), "vs_oracle %.4f" % (s8 / oracle8[label]), flush=True)
