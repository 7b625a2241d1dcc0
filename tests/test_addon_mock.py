"""The N-API addon (zlib.es_b200/node/addon.c) EXECUTED, not just compiled: linked against tests/napi_mock (a stand-in for
the handful of N-API calls it makes; the image has no Node.js) and driven through ctypes.

* CPU (this container): against the emulator build of the library — the addon's argument checks, result shapes, error
  mapping and batch paths run for real; and against libzles.so without a GPU — every call must surface ZLES_E_CUDA as a
  thrown Error, the batch forms included (they used to return empty arrays).
* GPU (`-m gpu`): the same cases against libzles.so.

The TypeScript layer above it (node/index.ts) only re-exports these functions.
"""
import ctypes
import os
import subprocess
import zlib

import pytest

import oracle as O
import vectors as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_addon(libpath: str, tag: str) -> ctypes.CDLL:
    out = os.path.join(ROOT, "build", "addon_mock_%s.so" % tag)
    srcs = [os.path.join(ROOT, "zlib.es_b200", "node", "addon.c"), os.path.join(ROOT, "tests", "napi_mock", "napi_mock.c")]
    deps = srcs + [os.path.join(ROOT, "include", "zles.h"), libpath]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-g", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", out] + srcs +
                              [libpath, "-Wl,-rpath," + os.path.dirname(libpath)])
    L = ctypes.CDLL(out)
    L.mock_export_name.restype = ctypes.c_char_p
    L.mock_call1.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                             ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
    L.mock_call_batch.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p),
                                  ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
    L.mock_free.argtypes = [ctypes.c_void_p]
    assert L.mock_load() >= 4
    return L


class JsError(Exception):
    pass


class Addon:
    """What `require('./zles.node')` gives index.ts."""

    def __init__(self, L):
        self.L = L
        self.exports = [L.mock_export_name(i).decode() for i in range(L.mock_load())]

    def call(self, name: str, data, kind: int = 0) -> bytes:
        out, n, err = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.create_string_buffer(512)
        b = bytes(data)
        rc = self.L.mock_call1(name.encode(), kind, b, len(b), ctypes.byref(out), ctypes.byref(n), err, 512)
        if rc:
            e = JsError(err.value.decode())
            e.type_error = rc == 2
            raise e
        try:
            return ctypes.string_at(out.value, n.value)
        finally:
            self.L.mock_free(out)

    def batch(self, name: str, bufs) -> list:
        lens = (ctypes.c_size_t * max(1, len(bufs)))(*[len(b) for b in bufs])
        out, err = ctypes.c_void_p(), ctypes.create_string_buffer(512)
        olens = (ctypes.c_size_t * max(1, len(bufs)))()
        rc = self.L.mock_call_batch(name.encode(), b"".join(bytes(b) for b in bufs) or b"\0", lens, len(bufs), ctypes.byref(out), olens, err, 512)
        if rc:
            raise JsError(err.value.decode())
        try:
            blob = ctypes.string_at(out.value, sum(olens[i] for i in range(len(bufs))))
        finally:
            self.L.mock_free(out)
        res, o = [], 0
        for i in range(len(bufs)):
            res.append(blob[o:o + olens[i]])
            o += olens[i]
        return res


def drop_in_cases(a: Addon, big: int):
    """test/index.js of the reference, through the addon: the inflate known-answer vectors, deflate round trips through
    the addon, the oracle and system zlib, the reference's error texts as thrown Errors; then the extra exports."""
    assert set(a.exports) >= {"deflate", "inflate", "deflateBatch", "inflateBatch", "deflateRaw", "inflateRaw", "gzip", "gunzip"}
    for name in ("UNCOMPRESSED", "FIXED", "DYNAMIC"):                    # test/index.js:15-35
        assert a.call("inflate", getattr(T, name)) == T.RAW
    assert a.call("inflate", T.fixture_compressed()) == T.fixture_raw()    # test/index.js:37-42
    for data in (T.RAW, T.repeat_input(), T.fixture_raw()[:big], b""):     # test/index.js:56-109
        z = a.call("deflate", data)
        assert a.call("inflate", z) == data and O.inflate(z) == data and zlib.decompress(z) == data
    for stream, msg in ((b"\x77\x9c\x03\x00", "Not compressed by deflate"), (b"\x78\x9c\x07\x00\x00\x00\x00\x00", "Not supported BTYPE : 3"),
                        (b"\x78\x9c\x01\x05\x00\x00\x00hello", "Data is corrupted")):
        with pytest.raises(JsError) as e:
            a.call("inflate", stream)
        assert str(e.value) == msg and not e.value.type_error
    z = zlib.compress(T.gen("G5", 3000))
    for cut in (40, 41, 300, len(z) - 5):   # truncated streams: the reference's outcome, an Error text or partial bytes
        try:
            want = ("ok", O.inflate(z[:cut]))
        except O.OracleError as oe:
            want = ("err", str(oe))
        try:
            got = ("ok", a.call("inflate", z[:cut]))
        except JsError as je:
            got = ("err", str(je))
        assert got == want, cut
    for kind in (1, 2):                                                   # not a Uint8Array / no argument: TypeError
        for fn in ("deflate", "inflate"):
            with pytest.raises(JsError) as e:
                a.call(fn, b"", kind)
            assert e.value.type_error
    # a result far larger than 10 x the input (the first capacity guess, src/inflate.ts:17)
    z = zlib.compress(bytes(big * 4), 9)
    assert a.call("inflate", z) == bytes(big * 4)
    bufs = [T.gen("G5", 4096), b"", T.gen("G1", 100), T.fixture_raw()[:big // 2], b"x", bytes(50000)]
    zs = a.batch("deflateBatch", bufs)
    assert [zlib.decompress(z) for z in zs] == bufs and zs[0] == a.call("deflate", bufs[0])
    assert a.batch("inflateBatch", zs) == bufs
    assert a.batch("inflateBatch", [zlib.compress(bytes(300000), 9), zs[0]]) == [bytes(300000), bufs[0]]  # one needs a retry with more room
    with pytest.raises(JsError, match="Not compressed by deflate"):
        a.batch("inflateBatch", [zs[0], b"\x77\x00"])
    assert a.batch("deflateBatch", []) == []
    data = T.fixture_raw()[:big]
    assert zlib.decompress(a.call("deflateRaw", data), -15) == data and a.call("inflateRaw", a.call("deflateRaw", data)) == data
    assert zlib.decompress(a.call("gzip", data), 31) == data and a.call("gunzip", a.call("gzip", data)) == data
    g = bytearray(a.call("gzip", data))
    g[-7] ^= 1
    with pytest.raises(JsError, match="gzip checksum mismatch"):
        a.call("gunzip", bytes(g))


@pytest.mark.emu
def test_addon_runs_against_the_emulator_library():
    import emu_lib
    a = Addon(build_addon(emu_lib.build(), "emu"))
    drop_in_cases(a, 60000)
    os.environ["NAPI_MOCK_NO_EXTERNAL"] = "1"   # engines that forbid external array buffers: the addon copies instead
    try:
        assert a.call("inflate", T.FIXED) == T.RAW
    finally:
        del os.environ["NAPI_MOCK_NO_EXTERNAL"]


def test_addon_throws_when_the_library_has_no_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import __graft_entry__ as G
    G.build()
    a = Addon(build_addon(os.path.join(ROOT, "zlib.es_b200", "libzles.so"), "nogpu"))
    for fn in ("deflate", "inflate", "gzip", "deflateRaw"):
        with pytest.raises(JsError, match="(?i)cuda|device|driver"):
            a.call(fn, zlib.compress(b"abc") if fn == "inflate" else b"abc")
    for fn in ("deflateBatch", "inflateBatch"):   # used to return empty arrays: the whole-call failure left status[] untouched
        with pytest.raises(JsError, match="(?i)cuda|device|driver"):
            a.batch(fn, [zlib.compress(b"abc"), zlib.compress(b"def")])


@pytest.mark.gpu
def test_addon_runs_against_libzles_on_the_gpu():
    a = Addon(build_addon(os.path.join(ROOT, "zlib.es_b200", "libzles.so"), "gpu"))
    drop_in_cases(a, 400000)
