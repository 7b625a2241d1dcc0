"""Compressed size at the default search level against the oracle (restatement of zlib.es) on 8 MiB samples of the
synthetic corpora and on the reference's fixture; oracle sizes computed once on the CPU (oracle.deflate) and kept here."""
import os, sys, json, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, zles
import vectors as T
c = zles.Codec(0)
if len(sys.argv) > 1:
    c.set_window_mode(int(sys.argv[1]))
oracle8 = {"text": 3405877, "binary": 2629584, "random": 8397052, "mixed": 3342227}
n = 8 << 20
for kind, label in [(0, "text"), (1, "binary"), (2, "random"), (3, "mixed")]:
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(kind, 0, src.data_ptr(), n)
    cap = c.deflate_bound(n); comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
    host = bytes(src.cpu().numpy())
    print(json.dumps({"corpus": label, "ours": clen, "oracle": oracle8[label], "ours/oracle": round(clen / oracle8[label], 4), "zlib-6": len(zlib.compress(host, 6))}), flush=True)
raw = T.fixture_raw()
z = c.deflate(raw)
print(json.dumps({"corpus": "reference fixture raw.bin", "ours": len(z), "oracle": 191734, "ours/oracle": round(len(z) / 191734, 4), "zlib-6": len(zlib.compress(raw, 6))}), flush=True)
