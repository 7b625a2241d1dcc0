"""Host-side mirror of the reference's operator surface over the C ABI.

``deflate`` / ``inflate`` keep the reference's argument meaning and error
behaviour (/root/reference/src/zlib.ts:11-49): one buffer in, a new buffer
out, synchronous, errors raised with the reference's message strings.

Deliberate differences (documented in DESIGN.md):
  * ``deflate`` succeeds for the lengths on which the reference throws
    (0, 1 and n = 1 mod 131072) — it returns a valid stream.
  * The compressed bytes differ from the reference's (its exact bits are not
    pinned by its own tests); they inflate to the same input under the
    reference's inflate, ours and system zlib, and are not larger than 1.03 x.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence

import numpy as np

from . import _capi


class ZlesError(Exception):
    """Mirrors ``throw new Error(msg)`` of the reference; ``code`` is the ZLES_E_* status."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def _addr(buf) -> tuple[int, int, object]:
    """(address, length, keep-alive) of a bytes-like object without copying when possible."""
    if isinstance(buf, np.ndarray):
        a = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        return a.ctypes.data, a.size, a
    if isinstance(buf, (bytes, bytearray, memoryview)):
        a = np.frombuffer(buf, dtype=np.uint8)
        return (a.ctypes.data if a.size else 0), a.size, a
    a = np.frombuffer(bytes(buf), dtype=np.uint8)
    return (a.ctypes.data if a.size else 0), a.size, a


class Codec:
    """One zles context (device + stream).  Not thread-safe; use one per thread."""

    def __init__(self, device: int = 0, lib: ctypes.CDLL | None = None):
        self.L = lib if lib is not None else _capi.lib()
        h = ctypes.c_void_p()
        self._check(self.L.zles_ctx_create(device, ctypes.byref(h)))
        self.h = h
        self.device = device
        self.stream_ptr = None

    def close(self):
        if getattr(self, "h", None):
            self.L.zles_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc == 0:
            return
        msg = self.L.zles_strerror(rc).decode()
        if rc == _capi.E_CUDA:
            msg += ": " + self.L.zles_last_cuda_error().decode()
        raise ZlesError(rc, msg)

    def set_window_mode(self, mode: int):
        """0: every block but a chunk's first sees the 32 KiB before it; 1: a chunk's third block sees none (faster)."""
        self._check(self.L.zles_ctx_set_window_mode(self.h, mode))

    def set_level(self, max_checks: int, min_checks: int, good_len: int, lazy: bool = True):
        self._check(self.L.zles_ctx_set_level(self.h, max_checks, min_checks, good_len, 1 if lazy else 0))

    def set_slab_blocks(self, blocks: int):
        """Blocks of 32 KiB per slab of the host-buffer inflate of our own streams (0 = automatic)."""
        self._check(self.L.zles_ctx_set_slab_blocks(self.h, blocks))

    def set_batch_slab(self, buffers: int):
        """Buffers per slab of the host batch calls (at most 16,384)."""
        self._check(self.L.zles_ctx_set_batch_slab(self.h, buffers))

    def set_stream_min(self, nbytes: int):
        """Host-buffer inflate copies streams of at least ``nbytes`` in pieces that are decoded as they land."""
        self._check(self.L.zles_ctx_set_stream_min(self.h, nbytes))

    def set_stream(self, cuda_stream: int):
        self._check(self.L.zles_ctx_set_stream(self.h, ctypes.c_void_p(cuda_stream)))
        self.stream_ptr = cuda_stream  # the caller's stream the codec's kernels are ordered on (None: its own)

    def sync(self):
        self._check(self.L.zles_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.L.zles_ctx_launches(self.h))

    def set_timing(self, on: bool):
        self._check(self.L.zles_ctx_set_timing(self.h, 1 if on else 0))

    def kernel_time(self, kernel: str) -> tuple[float, int]:
        """(summed device milliseconds, launches) of one kernel since set_timing(True)."""
        ms = ctypes.c_double()
        n = ctypes.c_uint64()
        self._check(self.L.zles_ctx_kernel_time(self.h, kernel.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)

    def dev_alloc(self, n: int) -> int:
        p = ctypes.c_void_p()
        self._check(self.L.zles_dev_alloc(self.h, n, ctypes.byref(p)))
        return int(p.value)

    def dev_free(self, ptr: int):
        self._check(self.L.zles_dev_free(self.h, ctypes.c_void_p(ptr)))

    def dev_copy(self, dst: int, src: int, n: int):
        self._check(self.L.zles_dev_copy(self.h, ctypes.c_void_p(dst), ctypes.c_void_p(src), n))

    def dev_copy_async(self, dst: int, src: int, n: int):
        """Ordered on the codec's stream, not waited for (a host source must be pinned and stay alive)."""
        self._check(self.L.zles_dev_copy_async(self.h, ctypes.c_void_p(dst), ctypes.c_void_p(src), n))

    def ipc_export(self, ptr: int) -> bytes:
        buf = (ctypes.c_uint8 * 64)()
        self._check(self.L.zles_ipc_export(ctypes.c_void_p(ptr), buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (ctypes.c_uint8 * 64)(*handle)
        p = ctypes.c_void_p()
        self._check(self.L.zles_ipc_open(buf, ctypes.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._check(self.L.zles_ipc_close(ctypes.c_void_p(ptr)))

    def deflate_bound(self, n: int) -> int:
        return int(self.L.zles_deflate_bound(n))

    # -- the drop-in pair (host buffers) -------------------------------------------------
    def deflate(self, data) -> bytes:
        """zlib.deflate, /root/reference/src/zlib.ts:25-49."""
        p, n, keep = _addr(data)
        cap = self.deflate_bound(n)
        out = np.empty(cap, dtype=np.uint8)
        olen = ctypes.c_size_t()
        self._check(self.L.zles_deflate(self.h, p, n, out.ctypes.data, cap, ctypes.byref(olen)))
        del keep
        return out[:olen.value].tobytes()

    def inflate(self, data) -> bytes:
        """zlib.inflate, /root/reference/src/zlib.ts:11-23."""
        p, n, keep = _addr(data)
        optr = ctypes.c_void_p()
        olen = ctypes.c_size_t()
        self._check(self.L.zles_inflate_alloc(self.h, p, n, ctypes.byref(optr), ctypes.byref(olen)))
        del keep
        try:
            return ctypes.string_at(optr.value, olen.value)
        finally:
            self.L.zles_free(optr)

    def inflate_into(self, data, out: np.ndarray) -> int:
        p, n, keep = _addr(data)
        olen = ctypes.c_size_t()
        self._check(self.L.zles_inflate(self.h, p, n, out.ctypes.data, out.size, ctypes.byref(olen)))
        del keep
        return olen.value

    def deflate_into(self, data, out: np.ndarray) -> int:
        p, n, keep = _addr(data)
        olen = ctypes.c_size_t()
        self._check(self.L.zles_deflate(self.h, p, n, out.ctypes.data, out.size, ctypes.byref(olen)))
        del keep
        return olen.value

    def adler32(self, data) -> int:
        """calcAdler32(...) >>> 0, /root/reference/src/adler32.ts:1-10."""
        p, n, keep = _addr(data)
        v = ctypes.c_uint32()
        self._check(self.L.zles_adler32(self.h, p, n, ctypes.byref(v)))
        del keep
        return int(v.value)

    # -- wire-format siblings: raw deflate data and gzip ---------------------------------------
    def _deflate_fmt(self, fn, data) -> bytes:
        p, n, keep = _addr(data)
        cap = self.deflate_bound(n) + 16
        out = np.empty(cap, dtype=np.uint8)
        olen = ctypes.c_size_t()
        self._check(fn(self.h, p, n, out.ctypes.data, cap, ctypes.byref(olen)))
        del keep
        return out[:olen.value].tobytes()

    def _inflate_fmt(self, call, n: int) -> bytes:
        cap = n * 10 + 131072  # the reference's initial guess (src/inflate.ts:17); retried once with the exact size
        for _ in range(2):
            out = np.empty(cap, dtype=np.uint8)
            olen = ctypes.c_size_t()
            rc = call(out.ctypes.data, cap, ctypes.byref(olen))
            if rc == _capi.E_OUTPUT_FULL:
                cap = olen.value
                continue
            self._check(rc)
            return out[:olen.value].tobytes()
        self._check(rc)

    def deflate_raw(self, data) -> bytes:
        """The reference's deflate core (/root/reference/src/deflate.ts:14): raw RFC 1951 data, no container."""
        return self._deflate_fmt(self.L.zles_deflate_raw, data)

    def inflate_raw(self, data, offset: int = 0) -> bytes:
        """The reference's inflate core, ``inflate(input, offset = 0)`` (/root/reference/src/inflate.ts:16)."""
        p, n, keep = _addr(data)
        return self._inflate_fmt(lambda o, cap, olen: self.L.zles_inflate_raw(self.h, p, n, offset, o, cap, olen), n)

    def gzip_deflate(self, data) -> bytes:
        return self._deflate_fmt(self.L.zles_gzip_deflate, data)

    def gzip_inflate(self, data) -> bytes:
        p, n, keep = _addr(data)
        return self._inflate_fmt(lambda o, cap, olen: self.L.zles_gzip_inflate(self.h, p, n, o, cap, olen), n)

    def crc32(self, data) -> int:
        p, n, keep = _addr(data)
        v = ctypes.c_uint32()
        self._check(self.L.zles_crc32(self.h, p, n, ctypes.byref(v)))
        del keep
        return int(v.value)

    # -- batches of independent buffers ----------------------------------------------------
    def deflate_batch(self, bufs: Sequence) -> list[bytes]:
        """``[deflate(b) for b in bufs]`` in one launch sequence."""
        count = len(bufs)
        if count == 0:
            return []
        lens = np.array([len(b) for b in bufs], dtype=np.uint64)
        in_off = np.zeros(count + 1, dtype=np.uint64)
        np.cumsum(lens, out=in_off[1:])
        blob = np.frombuffer(b"".join(bytes(b) for b in bufs), dtype=np.uint8) if int(in_off[-1]) else np.zeros(1, np.uint8)
        caps = np.array([self.deflate_bound(int(n)) for n in lens], dtype=np.uint64)
        out_off = np.zeros(count + 1, dtype=np.uint64)
        np.cumsum(caps, out=out_off[1:])
        out = np.empty(int(out_off[-1]), dtype=np.uint8)
        out_len = np.zeros(count, dtype=np.uint64)
        status = np.zeros(count, dtype=np.int32)
        rc = self.L.zles_deflate_batch(self.h, blob.ctypes.data, in_off.ctypes.data, count, out.ctypes.data, out_off.ctypes.data,
                                       out_len.ctypes.data, status.ctypes.data)
        self._check(rc)
        return [out[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() for i in range(count)]

    def inflate_batch(self, bufs: Sequence, out_sizes: Iterable[int] | None = None, raise_on_error: bool = True):
        """``[inflate(b) for b in bufs]``; ``out_sizes`` are capacity hints (default 10 x input + 64 each, the reference's own
        first guess, /root/reference/src/inflate.ts:17; a stream that needs more is retried with the size it reports)."""
        count = len(bufs)
        if count == 0:
            return []
        lens = np.array([len(b) for b in bufs], dtype=np.uint64)
        in_off = np.zeros(count + 1, dtype=np.uint64)
        np.cumsum(lens, out=in_off[1:])
        blob = np.frombuffer(b"".join(bytes(b) for b in bufs), dtype=np.uint8) if int(in_off[-1]) else np.zeros(1, np.uint8)
        if out_sizes is None:
            caps = lens * np.uint64(10) + np.uint64(64)
        else:
            caps = np.array(list(out_sizes), dtype=np.uint64)
        for _attempt in range(2):
            out_off = np.zeros(count + 1, dtype=np.uint64)
            np.cumsum(caps, out=out_off[1:])
            out = np.empty(max(1, int(out_off[-1])), dtype=np.uint8)
            out_len = np.zeros(count, dtype=np.uint64)
            status = np.zeros(count, dtype=np.int32)
            rc = self.L.zles_inflate_batch(self.h, blob.ctypes.data, in_off.ctypes.data, count, out.ctypes.data, out_off.ctypes.data,
                                           out_len.ctypes.data, status.ctypes.data)
            if rc in (_capi.E_CUDA, _capi.E_ARG):
                self._check(rc)
            full = status == _capi.E_OUTPUT_FULL
            if not full.any():
                break
            caps = np.where(full, out_len, caps).astype(np.uint64)  # out_len holds the size needed
        res = []
        for i in range(count):
            if status[i] != 0:
                if raise_on_error:
                    self._check(int(status[i]))
                res.append(ZlesError(int(status[i]), self.L.zles_strerror(int(status[i])).decode()))
            else:
                res.append(out[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes())
        return res

    # -- device-resident forms (raw device addresses as ints) ---------------------------------
    def dev_deflate(self, d_in: int, n: int, d_out: int, cap: int) -> int:
        olen = ctypes.c_size_t()
        self._check(self.L.zles_dev_deflate(self.h, d_in, n, d_out, cap, ctypes.byref(olen)))
        return olen.value

    def dev_inflate(self, d_in: int, n: int, d_out: int, cap: int) -> int:
        olen = ctypes.c_size_t()
        self._check(self.L.zles_dev_inflate(self.h, d_in, n, d_out, cap, ctypes.byref(olen)))
        return olen.value

    def dev_adler32(self, d_in: int, n: int) -> int:
        v = ctypes.c_uint32()
        self._check(self.L.zles_dev_adler32(self.h, d_in, n, ctypes.byref(v)))
        return int(v.value)

    def dev_deflate_batch(self, d_in: int, d_in_off: int, count: int, d_out: int, d_out_off: int, d_out_len: int, d_status: int) -> int:
        rc = self.L.zles_dev_deflate_batch(self.h, d_in, d_in_off, count, d_out, d_out_off, d_out_len, d_status)
        if rc in (_capi.E_CUDA, _capi.E_ARG):
            self._check(rc)
        return rc

    def dev_inflate_batch(self, d_in: int, d_in_off: int, count: int, d_out: int, d_out_off: int, d_out_len: int, d_status: int) -> int:
        rc = self.L.zles_dev_inflate_batch(self.h, d_in, d_in_off, count, d_out, d_out_off, d_out_len, d_status)
        if rc in (_capi.E_CUDA, _capi.E_ARG):
            self._check(rc)
        return rc

    def dev_deflate_phase1(self, d_in: int, n: int, is_last: bool) -> _capi.ShardInfo:
        info = _capi.ShardInfo()
        self._check(self.L.zles_dev_deflate_phase1(self.h, d_in, n, 1 if is_last else 0, ctypes.byref(info)))
        return info

    def dev_deflate_phase2(self, d_dst: int):
        self._check(self.L.zles_dev_deflate_phase2(self.h, d_dst))

    def dev_inflate_segment(self, d_in: int, n: int, d_out: int, cap: int, has_final: bool = True) -> int:
        olen = ctypes.c_size_t()
        self._check(self.L.zles_dev_inflate_segment(self.h, d_in, n, 1 if has_final else 0, d_out, cap, ctypes.byref(olen)))
        return olen.value

    def dev_scan_blocks(self, d_in: int, n: int, first: int) -> np.ndarray:
        """Block starts of one of our streams in d_in[0 .. n) (uint64, relative to d_in; [0] = first)."""
        cap = n // 32 + 64
        out = np.empty(cap, dtype=np.uint64)
        cnt = ctypes.c_size_t()
        self._check(self.L.zles_dev_scan_blocks(self.h, d_in, n, first, out.ctypes.data, cap, ctypes.byref(cnt)))
        return out[:cnt.value]

    def dev_corpus(self, kind: int, offset: int, d_out: int, n: int):
        self._check(self.L.zles_dev_corpus(self.h, kind, offset, d_out, n))

    def host_corpus(self, kind: int, offset: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint8)
        self._check(self.L.zles_host_corpus(kind, offset, out.ctypes.data, n))
        return out


class MultiCodec:
    """One process, several GPUs (zles_mgpu_*, include/zles.h): the same ``deflate`` / ``inflate`` on host buffers, the input's
    128 KiB chunks sharded over the devices inside the library.  Results are byte for byte those of ``Codec``."""

    def __init__(self, devices: Sequence[int], lib: ctypes.CDLL | None = None):
        self.L = lib if lib is not None else _capi.lib()
        arr = (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        self._check(self.L.zles_mgpu_create(arr, len(devices), ctypes.byref(h)))
        self.h = h
        self.devices = list(devices)

    _check = Codec._check

    def close(self):
        if getattr(self, "h", None):
            self.L.zles_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.L.zles_mgpu_launches(self.h))

    def set_min_shard(self, nbytes: int):
        self._check(self.L.zles_mgpu_set_min_shard(self.h, nbytes))

    def set_slab_blocks(self, blocks: int):
        for i in range(len(self.devices)):
            self._check(self.L.zles_ctx_set_slab_blocks(ctypes.c_void_p(self.L.zles_mgpu_ctx(self.h, i)), blocks))

    def deflate_bound(self, n: int) -> int:
        return int(self.L.zles_deflate_bound(n))

    def deflate(self, data) -> bytes:
        p, n, keep = _addr(data)
        cap = self.deflate_bound(n)
        out = np.empty(cap, dtype=np.uint8)
        olen = ctypes.c_size_t()
        self._check(self.L.zles_mgpu_deflate(self.h, p, n, out.ctypes.data, cap, ctypes.byref(olen)))
        del keep
        return out[:olen.value].tobytes()

    def inflate(self, data) -> bytes:
        p, n, keep = _addr(data)
        optr = ctypes.c_void_p()
        olen = ctypes.c_size_t()
        self._check(self.L.zles_mgpu_inflate_alloc(self.h, p, n, ctypes.byref(optr), ctypes.byref(olen)))
        del keep
        try:
            return ctypes.string_at(optr.value, olen.value)
        finally:
            self.L.zles_free(optr)

    def deflate_into(self, data, out: np.ndarray) -> int:
        p, n, keep = _addr(data)
        olen = ctypes.c_size_t()
        self._check(self.L.zles_mgpu_deflate(self.h, p, n, out.ctypes.data, out.size, ctypes.byref(olen)))
        del keep
        return olen.value

    def inflate_into(self, data, out: np.ndarray) -> int:
        p, n, keep = _addr(data)
        olen = ctypes.c_size_t()
        self._check(self.L.zles_mgpu_inflate(self.h, p, n, out.ctypes.data, out.size, ctypes.byref(olen)))
        del keep
        return olen.value


def combine_adler(infos: Sequence[_capi.ShardInfo], lib: ctypes.CDLL | None = None) -> int:
    L = lib if lib is not None else _capi.lib()
    arr = (_capi.ShardInfo * len(infos))(*infos)
    return int(L.zles_adler32_combine_shards(arr, len(infos)))


_default: Codec | None = None


def default_codec() -> Codec:
    global _default
    if _default is None:
        _default = Codec(0)
    return _default


def deflate(data) -> bytes:
    """Drop-in for ``zlib.deflate(input: Uint8Array): Uint8Array`` (/root/reference/dist/tsc/zlib.d.ts:5)."""
    return default_codec().deflate(data)


def inflate(data) -> bytes:
    """Drop-in for ``zlib.inflate(input: Uint8Array): Uint8Array`` (/root/reference/dist/tsc/zlib.d.ts:4)."""
    return default_codec().inflate(data)


def adler32(data) -> int:
    return default_codec().adler32(data)


def deflate_batch(bufs: Sequence) -> list[bytes]:
    return default_codec().deflate_batch(bufs)


def inflate_batch(bufs: Sequence, out_sizes=None) -> list[bytes]:
    return default_codec().inflate_batch(bufs, out_sizes)
