"""ctypes front end of the CPU oracle (oracle/zlibes_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product package
(zlib.es_b200/) never imports this module.

The functions mirror the reference's export surface
(/root/reference/dist/tsc/zlib.d.ts:4-5): ``deflate(bytes) -> bytes`` and
``inflate(bytes) -> bytes`` raising ``OracleError`` with the reference's exact
message strings (/root/reference/src/zlib.ts:15, src/inflate.ts:32,35,50).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libzlibes_oracle.so")
_lib = None


class OracleError(Exception):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "zlibes_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "zlibes_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.c_char_p
        pp = ctypes.POINTER(ctypes.c_void_p)
        szp = ctypes.POINTER(ctypes.c_size_t)
        L.zo_strerror.restype = ctypes.c_char_p
        L.zo_strerror.argtypes = [ctypes.c_int]
        L.zo_free.argtypes = [ctypes.c_void_p]
        L.zo_adler32.restype = ctypes.c_uint32
        L.zo_adler32.argtypes = [u8p, ctypes.c_size_t]
        for name in ("zo_deflate", "zo_deflate_raw", "zo_inflate"):
            f = getattr(L, name)
            f.restype = ctypes.c_int
            f.argtypes = [u8p, ctypes.c_size_t, pp, szp]
        L.zo_inflate_raw.restype = ctypes.c_int
        L.zo_inflate_raw.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, pp, szp]
        L.zo_deflate_raw_blk.restype = ctypes.c_int
        L.zo_deflate_raw_blk.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, pp, szp]
        L.zo_lz77_count.restype = ctypes.c_int
        L.zo_lz77_count.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                    ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        _lib = L
    return _lib


def _call(fn, data: bytes, *extra) -> bytes:
    L = lib()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    rc = fn(bytes(data), len(data), *extra, ctypes.byref(out), ctypes.byref(n))
    if rc != 0:
        raise OracleError(rc, L.zo_strerror(rc).decode())
    try:
        return ctypes.string_at(out.value, n.value)
    finally:
        L.zo_free(out)


def deflate(data: bytes) -> bytes:
    """zlib.deflate, /root/reference/src/zlib.ts:25-49."""
    return _call(lib().zo_deflate, data)


def deflate_raw(data: bytes, block_len: int | None = None) -> bytes:
    if block_len is None:
        return _call(lib().zo_deflate_raw, data)
    L = lib()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    rc = L.zo_deflate_raw_blk(bytes(data), len(data), block_len, ctypes.byref(out), ctypes.byref(n))
    if rc != 0:
        raise OracleError(rc, L.zo_strerror(rc).decode())
    try:
        return ctypes.string_at(out.value, n.value)
    finally:
        L.zo_free(out)


def inflate(data: bytes) -> bytes:
    """zlib.inflate, /root/reference/src/zlib.ts:11-23."""
    return _call(lib().zo_inflate, data)


def inflate_raw(data: bytes, offset: int = 0) -> bytes:
    return _call(lib().zo_inflate_raw, data, ctypes.c_size_t(offset))


def adler32(data: bytes) -> int:
    """calcAdler32 >>> 0, /root/reference/src/adler32.ts:1-10."""
    return int(lib().zo_adler32(bytes(data), len(data)))


def lz77_count(data: bytes, start: int, length: int) -> tuple[int, int]:
    nt = ctypes.c_uint32()
    nm = ctypes.c_uint32()
    rc = lib().zo_lz77_count(bytes(data), len(data), start, length, ctypes.byref(nt), ctypes.byref(nm))
    if rc:
        raise OracleError(rc, lib().zo_strerror(rc).decode())
    return nt.value, nm.value
