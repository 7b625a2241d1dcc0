"""The pageable-buffer staging (zlib.es_b200/csrc/stager.inl: pinned ring + helper threads) on the CPU emulator build.
ZLES_EMU_STAGE=1 makes the emulator library treat every host buffer as pageable and shrinks slot size and threshold, so the
multi-piece logic runs on small inputs; the variable is read when the library loads, hence the subprocess."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, zlib
sys.path[:0] = [%(root)r, %(root)r + "/tests", %(root)r + "/oracle"]
import numpy as np
import emu_lib, zles, vectors as T, oracle as O, parity_cases as P
c = emu_lib.codec()
data = (T.fixture_raw() + T.gen("G5", 400000) + bytes(200000) + T.gen("G3", 100000))[:1100000]
z = c.deflate(data)                      # staged source (slab by slab) and staged result
assert zlib.decompress(z) == data
for smin, slab in ((1 << 40, 0), (262144, 8), (300000, 16)):   # plain path, streaming path with two slab sizes
    c.set_stream_min(smin); c.set_slab_blocks(slab)
    assert c.inflate(z) == data, (smin, slab)
    assert c.inflate(zlib.compress(data, 6)) == data
c.set_stream_min(96 << 20); c.set_slab_blocks(0)
mc = zles.MultiCodec([0, 0, 0], lib=c.L)
mc.set_min_shard(65536)
assert mc.deflate(data) == z and mc.inflate(z) == data
out = np.zeros(len(data) - 1, dtype=np.uint8)
try:
    c.inflate_into(z, out)
except zles.ZlesError as e:
    assert e.code == 16
else:
    raise AssertionError("no error")
P.slabbed_host_inflate(c, 900000)
print("staging ok")
"""


@pytest.mark.emu
def test_pageable_buffers_are_staged_through_pinned_memory():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_lib
    emu_lib.build()
    env = dict(os.environ, ZLES_EMU_STAGE="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "staging ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
