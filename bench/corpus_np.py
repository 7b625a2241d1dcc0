"""numpy port of the synthetic corpora (zlib.es_b200/csrc/corpus.cuh).

BENCH / TEST INFRASTRUCTURE: the `--impl reference` arm of bench.py generates its input with
this module so that the CPU arm never loads the product library (libzles.so).
tests/test_corpus_np.py checks it byte for byte against zles_host_corpus.

Counter based: u(seed, i) = mix64(seed + GOLDEN * (i + 1)); data comes in independent 64 KiB
pages keyed by the absolute page index.
  kind 0  text      order-2 character Markov chain trained on corpus_text.h, every page starts in state "e "
  kind 1  binary    32-byte little-endian records
  kind 2  random
  kind 3  mixed     per 128 KiB segment s: u(MIX, s) % 20 — 0-1 random, 2-10 text, 11-19 binary
"""
from __future__ import annotations

import os
import re

import numpy as np

PAGE = 65536
NSYM = 59
SEED_T, SEED_B, SEED_R, SEED_MIX = 0xB2000001, 0xB2000002, 0xB2000003, 0xB20000FF
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_ALPHA = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ ,.;'-\n"
_HERE = os.path.dirname(os.path.abspath(__file__))
_PROSE_H = os.path.join(os.path.dirname(_HERE), "zlib.es_b200", "csrc", "corpus_text.h")


def u64(seed, i):
    """corpus_u: splitmix64 finaliser of seed + GOLDEN * (i + 1); seed and i broadcast (uint64 arrays)."""
    with np.errstate(over="ignore"):
        z = np.asarray(seed, dtype=np.uint64) + _GOLDEN * (np.asarray(i, dtype=np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _prose() -> bytes:
    """The C string literal ZLES_CORPUS_PROSE of corpus_text.h (adjacent literals concatenated, escapes resolved)."""
    src = open(_PROSE_H, encoding="utf-8").read()
    body = src[src.index("ZLES_CORPUS_PROSE[]"):]
    out = bytearray()
    for lit in re.findall(r'"((?:[^"\\]|\\.)*)"', body):
        out += lit.encode("latin-1").decode("unicode_escape").encode("latin-1")
    return bytes(out)


_table = None


def text_table():
    """(cdf uint16[NSYM*NSYM, 64], alphabet uint8[64]) — corpus_build_host of zles.cu."""
    global _table
    if _table is not None:
        return _table
    sym_of = {ord(ch): k for k, ch in enumerate(_ALPHA)}
    c2 = np.zeros((NSYM * NSYM, NSYM), dtype=np.uint64)
    c1 = np.zeros((NSYM, NSYM), dtype=np.uint64)
    c0 = np.zeros(NSYM, dtype=np.uint64)
    s1, s2 = 4, 52
    for b in _prose():
        s = sym_of.get(b)
        if s is None:
            continue
        c2[s1 * NSYM + s2, s] += 1
        c1[s2, s] += 1
        c0[s] += 1
        s1, s2 = s2, s
    cdf = np.zeros((NSYM * NSYM, 64), dtype=np.uint16)
    for st in range(NSYM * NSYM):
        cnt = c2[st]
        if cnt.sum() == 0:  # unseen pair: back off to the order-1, then the order-0 statistics
            cnt = c1[st % NSYM]
            if cnt.sum() == 0:
                cnt = c0
        tot = int(cnt.sum())
        cum = np.cumsum(cnt.astype(np.uint64))
        full = np.concatenate([cum, np.full(64 - NSYM, cum[-1], dtype=np.uint64)])
        v = (full * np.uint64(65536)) // np.uint64(tot) if tot else np.full(64, 65535, dtype=np.uint64)
        cdf[st] = np.minimum(v, 65535).astype(np.uint16)
    alphabet = np.full(64, ord(" "), dtype=np.uint8)
    alphabet[:NSYM] = np.frombuffer(_ALPHA.encode("latin-1"), dtype=np.uint8)
    _table = (cdf, alphabet)
    return _table


def kind_of_page(kind: int, pages: np.ndarray) -> np.ndarray:
    if kind != 3:
        return np.full(pages.shape, kind, dtype=np.int64)
    c = u64(np.uint64(SEED_MIX), pages >> np.uint64(1)) % np.uint64(20)
    return np.where(c < 2, 2, np.where(c < 11, 0, 1)).astype(np.int64)


def _random_pages(pages: np.ndarray) -> np.ndarray:
    j = np.arange(PAGE // 8, dtype=np.uint64)
    r = u64((np.uint64(SEED_R) ^ pages)[:, None], j[None, :])       # byte b of word j = r >> 8b: little endian
    return r.astype("<u8").view(np.uint8).reshape(len(pages), PAGE)


def _text_pages(pages: np.ndarray) -> np.ndarray:
    cdf, alphabet = text_table()
    P = len(pages)
    out = np.empty((P, PAGE), dtype=np.uint8)
    seeds = np.uint64(SEED_T) ^ pages
    s1 = np.full(P, 4, dtype=np.int64)
    s2 = np.full(P, 52, dtype=np.int64)
    cdf32 = cdf[:, :NSYM].astype(np.int32)
    BLK = 256
    for j0 in range(0, PAGE, BLK):
        # the random draws of BLK steps at once; the chain itself is sequential in j
        r = ((u64(seeds[:, None], np.arange(j0, j0 + BLK, dtype=np.uint64)[None, :]) >> np.uint64(24)) & np.uint64(0xFFFF)).astype(np.int32)
        for k in range(BLK):
            rows = cdf32[s1 * NSYM + s2]                              # [P, NSYM]
            # first index with cdf > r among 0..NSYM-2, else NSYM-1 (the binary search of corpus_page)
            nxt = (rows[:, :NSYM - 1] <= r[:, k:k + 1]).sum(axis=1)
            # the binary search assumes a monotone row; cumulative sums are, so counting "<= r" gives the same index
            out[:, j0 + k] = alphabet[nxt]
            s1, s2 = s2, nxt
    return out


def _binary_pages(pages: np.ndarray) -> np.ndarray:
    P = len(pages)
    NREC = PAGE // 32
    rec = np.arange(NREC, dtype=np.uint64)
    seeds = (np.uint64(SEED_B) ^ pages)[:, None]
    r0 = u64(seeds, 2 * rec[None, :])
    r1 = u64(seeds, 2 * rec[None, :] + np.uint64(1))
    out = np.zeros((P, NREC, 32), dtype=np.uint8)
    with np.errstate(over="ignore"):
        ident = ((pages[:, None] * np.uint64(2048) + rec[None, :]) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        ts0 = ((pages * np.uint64(51200)) & np.uint64(0xFFFFFFFF)).astype(np.uint64)
        ts = ((ts0[:, None] + np.cumsum(r0 % np.uint64(50), axis=1)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    fbase = np.array([0.0, 0.5, 1.0, 1.5, 2.25, 100.0], dtype=np.float32)
    f = fbase[((r0 >> np.uint64(8)) % np.uint64(6)).astype(np.int64)] + ((r0 >> np.uint64(16)) % np.uint64(4)).astype(np.float32) * np.float32(0.25)
    h = ((r0 >> np.uint64(24)) % np.uint64(300)).astype(np.uint16)
    tag = (r1 % np.uint64(64)).astype(np.uint32)
    out[:, :, 0:4] = ident.astype("<u4")[..., None].view(np.uint8).reshape(P, NREC, 4)
    out[:, :, 4:8] = ts.astype("<u4")[..., None].view(np.uint8).reshape(P, NREC, 4)
    out[:, :, 8:12] = f.astype("<f4")[..., None].view(np.uint8).reshape(P, NREC, 4)
    out[:, :, 12:14] = h.astype("<u2")[..., None].view(np.uint8).reshape(P, NREC, 2)
    out[:, :, 14] = ord("T")
    out[:, :, 15] = ord("A")
    out[:, :, 16] = ord("G")
    out[:, :, 17] = (ord("A") + (tag >> 3)).astype(np.uint8)
    out[:, :, 18] = (ord("a") + (tag & 7)).astype(np.uint8)
    out[:, :, 19] = (ord("0") + tag % 10).astype(np.uint8)
    out[:, :, 20] = (ord("0") + (tag * 7) % 10).astype(np.uint8)
    out[:, :, 21] = (ord("0") + (tag * 3) % 10).astype(np.uint8)
    return out.reshape(P, PAGE)


def corpus(kind: int, offset: int, n: int) -> np.ndarray:
    """n bytes of corpus `kind` starting at absolute byte `offset` (uint8 array) — same bytes as zles_host_corpus."""
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    p0, p1 = offset // PAGE, (offset + n + PAGE - 1) // PAGE
    pages = np.arange(p0, p1, dtype=np.uint64)
    kinds = kind_of_page(kind, pages)
    buf = np.empty((len(pages), PAGE), dtype=np.uint8)
    for k, gen in ((0, _text_pages), (1, _binary_pages), (2, _random_pages)):
        idx = np.nonzero(kinds == k)[0]
        for a in range(0, len(idx), 1024):  # bounded working set
            sel = idx[a:a + 1024]
            buf[sel] = gen(pages[sel])
    flat = buf.reshape(-1)
    lo = offset - p0 * PAGE
    return flat[lo:lo + n]
