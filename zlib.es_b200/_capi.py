"""ctypes binding of libzles.so (the C ABI declared in include/zles.h).

This is the Python stand-in for the N-API addon (zlib.es_b200/node/addon.c): the
image has no Node.js, so tests/, bench.py and __graft_entry__ drive the same C
entry points from here.  There is no CPU fallback: if libzles.so (the nvcc
sm_100a build) is missing, importing the codec raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzles.so")

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_szp = ctypes.POINTER(ctypes.c_size_t)
c_vp = ctypes.c_void_p


class ShardInfo(ctypes.Structure):
    """zles_shard_info"""
    _fields_ = [("comp_bytes", ctypes.c_uint64), ("raw_bytes", ctypes.c_uint64), ("adler_a", ctypes.c_uint64),
                ("adler_b", ctypes.c_uint64), ("n_blocks", ctypes.c_uint64)]


# name -> (restype, argtypes); every symbol include/zles.h declares
PROTOTYPES = {
    "zles_version": (ctypes.c_char_p, []),
    "zles_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "zles_last_cuda_error": (ctypes.c_char_p, []),
    "zles_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "zles_ctx_destroy": (None, [c_vp]),
    "zles_ctx_set_stream": (ctypes.c_int, [c_vp, c_vp]),
    "zles_ctx_set_level": (ctypes.c_int, [c_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]),
    "zles_ctx_set_window_mode": (ctypes.c_int, [c_vp, ctypes.c_uint32]),
    "zles_ctx_set_slab_blocks": (ctypes.c_int, [c_vp, ctypes.c_uint32]),
    "zles_ctx_set_stream_min": (ctypes.c_int, [c_vp, ctypes.c_size_t]),
    "zles_ctx_set_batch_slab": (ctypes.c_int, [c_vp, ctypes.c_uint32]),
    "zles_ctx_launches": (ctypes.c_uint64, [c_vp]),
    "zles_init": (ctypes.c_int, [ctypes.c_uint32]),
    "zles_shutdown": (None, []),
    "zles_mgpu_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(c_vp)]),
    "zles_mgpu_destroy": (None, [c_vp]),
    "zles_mgpu_device_count": (ctypes.c_int, [c_vp]),
    "zles_mgpu_set_min_shard": (ctypes.c_int, [c_vp, ctypes.c_size_t]),
    "zles_mgpu_ctx": (c_vp, [c_vp, ctypes.c_int]),
    "zles_mgpu_launches": (ctypes.c_uint64, [c_vp]),
    "zles_mgpu_deflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_mgpu_inflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_mgpu_inflate_alloc": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp), c_szp]),
    "zles_ctx_set_timing": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "zles_ctx_kernel_time": (ctypes.c_int, [c_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "zles_dev_alloc": (ctypes.c_int, [c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp)]),
    "zles_dev_free": (ctypes.c_int, [c_vp, c_vp]),
    "zles_dev_copy": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
    "zles_dev_copy_async": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
    "zles_ctx_sync": (ctypes.c_int, [c_vp]),
    "zles_deflate_bound": (ctypes.c_size_t, [ctypes.c_size_t]),
    "zles_deflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_inflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_inflate_alloc": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp), c_szp]),
    "zles_free": (None, [c_vp]),
    "zles_adler32": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]),
    "zles_deflate_raw": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_inflate_raw": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_gzip_deflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_gzip_inflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_crc32": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]),
    "zles_dev_crc32": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]),
    "zles_crc32_combine": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]),
    "zles_deflate_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "zles_inflate_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "zles_dev_deflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_dev_inflate": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, c_szp]),
    "zles_dev_adler32": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]),
    "zles_dev_deflate_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "zles_dev_inflate_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "zles_dev_deflate_phase1": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ShardInfo)]),
    "zles_dev_deflate_block_offsets": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "zles_dev_deflate_phase2": (ctypes.c_int, [c_vp, c_vp]),
    "zles_adler32_combine_shards": (ctypes.c_uint32, [ctypes.POINTER(ShardInfo), ctypes.c_uint32]),
    "zles_dev_inflate_segment": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, c_vp, ctypes.c_size_t, c_szp]),
    "zles_dev_scan_blocks": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_uint64, c_vp, ctypes.c_size_t, c_szp]),
    "zles_ipc_export": (ctypes.c_int, [c_vp, c_vp]),
    "zles_ipc_open": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "zles_ipc_close": (ctypes.c_int, [c_vp]),
    "zles_dev_corpus": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_uint64, c_vp, ctypes.c_size_t]),
    "zles_host_corpus": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, c_vp, ctypes.c_size_t]),
}

# status codes of include/zles.h
OK, E_NOT_DEFLATE, E_BTYPE3, E_INSUFFICIENT, E_CORRUPTED, E_LACK = 0, 1, 2, 3, 4, 5
E_OUTPUT_FULL, E_CUDA, E_ARG, E_NOMEM, E_RUNAWAY, E_CHECKSUM = 16, 17, 18, 19, 20, 21


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    """Attach the prototypes; raises AttributeError if a declared symbol is missing."""
    for name, (res, args) in PROTOTYPES.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    return lib


_lib = None


def lib() -> ctypes.CDLL:
    """libzles.so, loaded once.  Fails loudly when the CUDA build is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "zles: %s not found — build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib
