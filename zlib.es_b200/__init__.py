"""zles — B200-native zlib codec with the zlib.es API.

Host-side mirror of the reference's export surface
(/root/reference/dist/tsc/zlib.d.ts:4-5, /root/reference/src/zlib.ts:11-49):

    deflate(input: bytes-like) -> bytes      # zlib.deflate
    inflate(input: bytes-like) -> bytes      # zlib.inflate

Errors are raised as ``ZlesError`` whose message is the reference's exact
``Error`` text (``'Not compressed by deflate'``, ``'Data is corrupted'`` ...).
Everything runs on the GPU through libzles.so (include/zles.h); there is no CPU
path.  The directory name contains a dot, so import it through the ``zles``
shim at the repository root (``import zles``).
"""
from __future__ import annotations

from .codec import (Codec, MultiCodec, ZlesError, adler32, default_codec, deflate, deflate_batch, inflate, inflate_batch)  # noqa: F401
from . import _capi  # noqa: F401

__all__ = ["deflate", "inflate", "adler32", "deflate_batch", "inflate_batch", "Codec", "MultiCodec", "ZlesError", "default_codec"]
