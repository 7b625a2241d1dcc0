import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch, zles
c = zles.Codec(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n = count * 4096
src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
in_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * 4096
bound = c.deflate_bound(4096)
out_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * bound
out = torch.empty(count * bound, dtype=torch.uint8, device="cuda")
out_len = torch.zeros(count, dtype=torch.int64, device="cuda")
status = torch.ones(count, dtype=torch.int32, device="cuda")
rc = c.dev_deflate_batch(src.data_ptr(), in_off.data_ptr(), count, out.data_ptr(), out_off.data_ptr(), out_len.data_ptr(), status.data_ptr())
print("deflate rc", rc, int(status.abs().sum()))
back = torch.zeros(n, dtype=torch.uint8, device="cuda")
blen = torch.zeros(count, dtype=torch.int64, device="cuda")
st2 = torch.ones(count, dtype=torch.int32, device="cuda")
# inflate straight from the padded layout: stream i = [out_off[i], out_off[i+1]) (trailing garbage is ignored like the reference does)
rc = c.dev_inflate_batch(out.data_ptr(), out_off.data_ptr(), count, back.data_ptr(), in_off.data_ptr(), blen.data_ptr(), st2.data_ptr())
bad = torch.nonzero(st2).flatten()
print("inflate rc", rc, "bad streams", bad.numel(), bad[:10].tolist(), st2[bad[:10]].tolist())
print("equal", torch.equal(src, back))
hs = src.cpu().numpy(); ho = out.cpu().numpy(); ol = out_len.cpu().numpy()
for i in bad[:5].tolist() + list(range(0, count, max(1, count // 50))):
    z = ho[i * bound: i * bound + ol[i]].tobytes(); raw = hs[i * 4096:(i + 1) * 4096].tobytes()
    try:
        ok = zlib.decompress(z) == raw
    except Exception as e:
        ok = repr(e)
    if ok is not True or i in bad[:5].tolist():
        print(i, "len", ol[i], "zlib:", ok, "kind-seg", i * 4096 // 131072)
        try:
            r = c.inflate(z); print("   single inflate ok", r == raw)
        except Exception as e:
            print("   single inflate", e)
