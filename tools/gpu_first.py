"""First-contact GPU script: parity on the reference vectors, then rough timings."""
import os, sys, time, zlib, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import zles
import vectors as T
import oracle as O

c = zles.Codec(0)
print(c.L.zles_version(), torch.cuda.get_device_name(0), flush=True)
bad = 0
def check(name, cond):
    global bad
    print(("ok   " if cond else "FAIL ") + name, flush=True)
    bad += 0 if cond else 1

check("adler RAW", c.adler32(T.RAW) == 0x2B23056C)
check("adler fixture", c.adler32(T.fixture_raw()) == 0x140FA15B)
for name in ("UNCOMPRESSED", "FIXED", "DYNAMIC"):
    check("inflate " + name, c.inflate(getattr(T, name)) == T.RAW)
check("inflate fixture", c.inflate(T.fixture_compressed()) == T.fixture_raw())
for name, n in [("RAW", 0), ("REPEAT", 0), ("FIXTURE", 0), ("G1", 2), ("G1", 1), ("G1", 0), ("G1", 4096), ("G1", 131072), ("G1", 131073), ("G1", 300000),
                ("G2", 4096), ("G2", 131072), ("G3", 4096), ("G3", 65536), ("G4", 4096), ("G5", 4096), ("G5", 200000)]:
    d = T.gen(name, n)
    z = c.deflate(d)
    try:
        osz = len(O.deflate(d))
    except Exception:
        osz = None
    ok = zlib.decompress(z) == d and O.inflate(z) == d and c.inflate(z) == d
    check("deflate %s-%d: %d -> %d (oracle %s, zlib6 %d)" % (name, len(d), len(d), len(z), osz, len(zlib.compress(d, 6))), ok and (osz is None or len(z) <= 1.03 * osz))

# corpora: host and device generators agree
for kind in range(4):
    n = 3 * 65536 + 777
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    c.dev_corpus(kind, 12345, dev.data_ptr(), n)
    host = c.host_corpus(kind, 12345, n)
    check("corpus kind %d host == device" % kind, bool((dev.cpu().numpy() == host).all()))

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return r, best

for kind, label, n in [(0, "text", 64 << 20), (1, "binary", 64 << 20), (2, "random", 64 << 20), (3, "mixed", 256 << 20)]:
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    c.dev_corpus(kind, 0, src.data_ptr(), n)
    cap = c.deflate_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    clen, td = timed(lambda: c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap))
    olen, ti = timed(lambda: c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n))
    _, ta = timed(lambda: c.dev_adler32(src.data_ptr(), n))
    same = olen == n and bool(torch.equal(src, back))
    sample = src[: 8 << 20].cpu().numpy().tobytes()
    zc = comp[:clen].cpu().numpy().tobytes() if n <= (64 << 20) else None
    zok = (zlib.decompress(zc) == src.cpu().numpy().tobytes()) if zc is not None else None
    t0 = time.perf_counter(); osz = len(O.deflate(sample)); to = time.perf_counter() - t0
    ours_sample = c.deflate(sample)
    print(json.dumps({"corpus": label, "n": n, "comp": clen, "ratio": round(n / clen, 4), "deflate_GBps": round(n / td / 1e9, 3),
                      "inflate_GBps": round(n / ti / 1e9, 3), "adler_GBps": round(n / ta / 1e9, 1), "roundtrip": same, "syszlib_ok": zok,
                      "sample8M_ours": len(ours_sample), "sample8M_oracle": osz, "size_vs_oracle": round(len(ours_sample) / osz, 4),
                      "sample8M_zlib6": len(zlib.compress(sample, 6)), "oracle_deflate_MBps": round(len(sample) / to / 1e6, 2)}), flush=True)
    bad += 0 if same else 1
print("launches", c.launches)
print("FAILED %d" % bad if bad else "ALL OK")
sys.exit(1 if bad else 0)
