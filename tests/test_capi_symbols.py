"""CPU-side checks of the drop-in boundary: libzles.so loads, exports every symbol that
include/zles.h declares, and the host-only entry points behave (no GPU needed)."""
import ctypes
import os
import re
import zlib

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    G.build()
    import zles
    from zles import _capi
    return _capi.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zles.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zles_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from zles import _capi
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libzles.so does not export %s" % n
        assert n in _capi.PROTOTYPES, "%s has no ctypes prototype" % n
    assert sorted(_capi.PROTOTYPES) == names


def test_error_strings_are_the_references(lib):
    # /root/reference/src/zlib.ts:15, src/inflate.ts:32,35,50, src/utils/BitReadStream.ts:15
    want = {1: "Not compressed by deflate", 2: "Not supported BTYPE : 3", 3: "Data length is insufficient",
            4: "Data is corrupted", 5: "Lack of data length"}
    for code, msg in want.items():
        assert lib.zles_strerror(code).decode() == msg
    assert lib.zles_strerror(0) == b""
    assert b"sm_100a" in lib.zles_version()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert lib.zles_ctx_create(0, ctypes.byref(h)) == 17  # ZLES_E_CUDA
    assert lib.zles_last_cuda_error() != b""
    out = ctypes.c_size_t()
    buf = (ctypes.c_uint8 * 64)()
    assert lib.zles_deflate(None, b"abc", 3, buf, 64, ctypes.byref(out)) == 17
    import zles
    with pytest.raises(zles.ZlesError, match="CUDA error"):
        zles.deflate(b"abc")


def test_deflate_bound_monotone(lib):
    prev = 0
    for n in [0, 1, 2, 100, 4096, 32768, 32769, 131072, 131073, 1 << 20, 1 << 30]:
        b = lib.zles_deflate_bound(n)
        assert b >= n + 6 and b >= prev
        prev = b


def test_adler_combine_shards(lib):
    from zles import _capi
    rng = np.random.default_rng(7)
    data = rng.integers(0, 256, size=500001, dtype=np.uint8).tobytes()
    cuts = [0, 131072, 131072 * 3, len(data)]
    infos = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        seg = np.frombuffer(data[a:b], dtype=np.uint8).astype(np.uint64)
        m = len(seg)
        info = _capi.ShardInfo()
        info.raw_bytes = m
        info.adler_a = int(seg.sum() % 65521)
        info.adler_b = int(((np.arange(m, 0, -1, dtype=np.uint64) % 65521) * seg % 65521).sum() % 65521)
        infos.append(info)
    arr = (_capi.ShardInfo * len(infos))(*infos)
    assert lib.zles_adler32_combine_shards(arr, len(infos)) == zlib.adler32(data)
    one = (_capi.ShardInfo * 1)(_capi.ShardInfo())
    assert lib.zles_adler32_combine_shards(one, 1) == 1  # empty input


def test_host_corpus_deterministic_and_windowed(lib):
    n = 200000
    full = np.empty(n, dtype=np.uint8)
    for kind in range(4):
        assert lib.zles_host_corpus(kind, 0, full.ctypes.data, n) == 0
        part = np.empty(70000, dtype=np.uint8)
        assert lib.zles_host_corpus(kind, 65000, part.ctypes.data, 70000) == 0
        assert (part == full[65000:135000]).all()
    text = np.empty(1 << 20, dtype=np.uint8)
    lib.zles_host_corpus(0, 0, text.ctypes.data, text.size)
    ratio = text.size / len(zlib.compress(text.tobytes(), 6))
    assert 1.8 < ratio < 3.5, ratio  # English-like order-2 Markov text
    rnd = np.empty(1 << 18, dtype=np.uint8)
    lib.zles_host_corpus(2, 0, rnd.ctypes.data, rnd.size)
    assert len(zlib.compress(rnd.tobytes(), 6)) > rnd.size
