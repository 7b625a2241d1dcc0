"""Size against the oracle (the restatement of zlib.es) on random inputs: the north_star's "within 3 %" on more than the
named corpora.  usage: python tools/gpu_size_sweep.py [cases=200] [seed=1] [log2 of the largest size=21]
Prints every case above 1.03 and the worst ratios per kind; exit code 1 if any case is above the bound."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import oracle as O
import zles
import vectors as T
import stress_cases as S

c = zles.Codec(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
maxlog = int(sys.argv[3]) if len(sys.argv) > 3 else 21
rng = np.random.default_rng(seed)
raw = T.fixture_raw()
worst, bad, tot_o, tot_r = [], 0, 0, 0
for i in range(cases):
    n = S.size(rng, maxlog)
    state = rng.bit_generator.state
    d = S.make(rng, n, raw)
    z = c.deflate(d)
    assert zlib.decompress(z) == d, ("round trip", i, n)
    try:
        r = len(O.deflate(d))
    except O.OracleError:
        continue  # a length on which the reference throws
    ratio = len(z) / r
    tot_o += len(z); tot_r += r
    worst.append((ratio, i, n, len(z), r))
    if len(z) > 1.03 * r:
        bad += 1
        print("ABOVE case %d n=%d ours=%d oracle=%d ratio=%.4f first bytes %r" % (i, n, len(z), r, ratio, d[:24]), flush=True)
worst.sort(reverse=True)
print("cases %d  above the bound %d  total ours/oracle %.4f" % (len(worst), bad, tot_o / max(1, tot_r)))
for w in worst[:8]:
    print("  ratio %.4f  case %d  n=%d  ours=%d  oracle=%d" % w)
sys.exit(1 if bad else 0)
