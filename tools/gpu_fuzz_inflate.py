"""Damaged streams on the GPU against the oracle: random multi-bit flips and truncations of streams of every kind
(ours, zlib.es-made, system zlib dynamic / fixed / stored, the hand-built undefined-code vectors).  Same bytes or the same
error text, and never a hang (the whole run sits under `timeout`).  usage: python tools/gpu_fuzz_inflate.py [seed] [cases]"""
import os, random, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as O, vectors as T, zles, parity_cases as P
c = zles.Codec(0)
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 77)
d = T.gen("G5", 700)
big = T.gen("G5", 90000) + T.fixture_raw()[:120000]
co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
streams = [zlib.compress(d, 6), O.deflate(d), zlib.compress(T.gen("G3", 100), 0), T.FIXED, T.DYNAMIC, co.compress(d) + co.flush(), c.deflate(d),
           c.deflate(big), zlib.compress(big, 6), O.deflate(big)]
streams += [s for _, s, _ in T.undefined_code_vectors()]
bad = n = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3000):
    z = rnd.choice(streams)
    zz = bytearray(z)
    for _k in range(rnd.choice([1, 1, 2, 3, 5, 8])):
        zz[rnd.randrange(2, len(zz))] ^= 1 << rnd.randrange(8)
    if rnd.random() < 0.2:
        zz = zz[:rnd.randrange(2, len(zz))]
    a, b = P._outcome(O.inflate, bytes(zz)), P._outcome(c.inflate, bytes(zz))
    n += 1
    if a != b:
        bad += 1
        print("MISMATCH", a[0], a[1][-20:] if a[0] == "err" else len(a[1]), b[0], b[1][-20:] if b[0] == "err" else len(b[1]), bytes(zz).hex()[:300], flush=True)
print("done", n, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
