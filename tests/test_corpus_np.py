"""bench/corpus_np.py (the numpy port of the corpus generator that the `--impl reference` arm of bench.py uses so that the CPU
arm never loads the product library) gives exactly the bytes of zles_host_corpus / zles_dev_corpus."""
import os
import sys

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bench"))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_numpy_corpus_equals_the_library_generator(kind):
    import corpus_np
    G.build()
    from zles import _capi
    lib = _capi.lib()
    for off, n in ((0, 140000), (65536 * 3 + 17, 131072 * 2 + 5), (8 << 30, 70000)):
        want = np.empty(n, dtype=np.uint8)
        assert lib.zles_host_corpus(kind, off, want.ctypes.data, n) == 0
        got = corpus_np.corpus(kind, off, n)
        assert got.shape == want.shape and bool((got == want).all()), (kind, off, n)
