"""Builds and loads the CUDA sources through the CPU thread emulator (tests/emu).

TESTS ONLY.  The emulator library is a debugging aid for the kernels' logic
(barriers, warp votes, indexing) in a container without a GPU; the product
never loads it (zlib.es_b200/_capi.py opens libzles.so only).  The parity
tests proper are the `-m gpu` tests, which run the nvcc build on a B200.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "zlib.es_b200", "csrc")
OUT = os.path.join(ROOT, "build", "libzles_emu.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "tests", "emu", f) for f in ("cuda_emu.h", "cuda_emu.cc")]
    srcs.append(os.path.join(ROOT, "include", "zles.h"))
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-g", "-DZLES_EMU", "-x", "c++", "-I" + os.path.join(ROOT, "tests", "emu"), "-I" + SRC,
               "-fPIC", "-shared", "-o", OUT, os.path.join(SRC, "zles.cu"), os.path.join(ROOT, "tests", "emu", "cuda_emu.cc"), "-lpthread"]
        subprocess.check_call(cmd)
    return OUT


_codec = None


def codec():
    """A Codec bound to the emulator library."""
    global _codec
    if _codec is None:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import zles
        from zles import _capi
        lib = _capi.bind(ctypes.CDLL(build()))
        _codec = zles.Codec(0, lib=lib)
    return _codec
