"""Randomised round trips through the C ABI on a GPU (not a benchmark): our deflate against system zlib's inflate and
ours, system zlib's deflate (random level / strategy) against our inflate, single calls and batches.
usage: python tools/gpu_stress.py [cases=200] [seed=1] [log2 of the largest size=22] [window mode] [scan width]"""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import zles
import vectors as T
c = zles.Codec(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
maxlog = int(sys.argv[3]) if len(sys.argv) > 3 else 22
if len(sys.argv) > 4:  # window mode, then optionally the scan width
    c.set_window_mode(int(sys.argv[4]))
if len(sys.argv) > 5:
    c.set_level(int(sys.argv[5]), 1, 8, True)
raw = T.fixture_raw()

import stress_cases as S


def make(n):
    return S.make(rng, n, raw)


bad = 0
for i in range(cases):
    n = S.size(rng, maxlog)
    d = make(n)
    z = c.deflate(d)
    ok = zlib.decompress(z) == d and c.inflate(z) == d
    lvl = int(rng.integers(0, 10)); strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
    co = zlib.compressobj(lvl, zlib.DEFLATED, 15, 8, strat)
    f = co.compress(d) + co.flush()
    ok = ok and c.inflate(f) == d
    if not ok:
        bad += 1
        print("FAIL case", i, "n", n, "level", lvl, "strategy", strat, flush=True)
for r in range(max(1, cases // 40)):
    bufs = [make(int(rng.integers(0, 9000 if rng.integers(0, 2) else 4097))) for _ in range(int(rng.integers(1, 200)))]
    zs = c.deflate_batch(bufs)
    ok = all(zlib.decompress(z) == b for b, z in zip(bufs, zs)) and c.inflate_batch(zs) == bufs and all(z == c.deflate(b) for b, z in zip(bufs[:30], zs[:30]))
    if not ok:
        bad += 1
        print("FAIL batch", r, flush=True)
print("stress: %d cases, %d failures" % (cases, bad))
sys.exit(1 if bad else 0)
