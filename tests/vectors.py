"""Shared test inputs: the reference's known-answer vectors (tests/golden/vectors.json,
made by tests/golden/make_golden.py from /root/reference/test/index.js) and the
SURVEY.md §8c generators."""
import json
import os
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
V = json.load(open(os.path.join(GOLD, "vectors.json")))

RAW = bytes(V["RAW"])
UNCOMPRESSED = bytes(V["UNCOMPRESSED"])
FIXED = bytes(V["FIXED"])
DYNAMIC = bytes(V["DYNAMIC"])


def undefined_code_vectors():
    """Hand-built streams that use literal/length symbols 286/287 and distance codes 30.. (tests/golden/make_undefined_vectors.py):
    [(name, stream, expected output)]."""
    doc = json.load(open(os.path.join(GOLD, "undefined_codes.json")))
    return [(v["name"], bytes.fromhex(v["stream"]), bytes.fromhex(v["expect"])) for v in doc["vectors"]]


def repeat_input() -> bytes:
    rep = ""
    while len(rep) < 1000:
        rep += V["REPEAT_ALPHABET"]
    assert len(rep) == V["REPEAT_LEN"]
    return rep.encode()


def fixture_compressed() -> bytes:
    return open(os.path.join(GOLD, "ref_fixture_compressed.zlib"), "rb").read()


_raw_cache = None


def fixture_raw() -> bytes:
    """test/data/raw.bin, reproduced from the compressed fixture with system zlib
    and checked against the recorded sha256."""
    global _raw_cache
    if _raw_cache is None:
        import hashlib
        raw = zlib.decompress(fixture_compressed())
        assert hashlib.sha256(raw).hexdigest() == V["fixture"]["raw_sha256"]
        _raw_cache = raw
    return _raw_cache


def _lcg(n):
    x = 1
    out = []
    for _ in range(n):
        x = (1103515245 * x + 12345) & 0x7FFFFFFF
        out.append(x >> 16)
    return out


def gen(name: str, n: int) -> bytes:
    if name == "RAW":
        return RAW
    if name == "REPEAT":
        return repeat_input()
    if name == "FIXTURE":
        return fixture_raw()
    if name == "G1":
        return bytes(((7 * i + 3) & 255) for i in range(n))
    if name == "G2":
        return bytes(n)
    if name == "G3":
        return bytes(v & 255 for v in _lcg(n))
    if name == "G4":
        return (b"abc" * (n // 3 + 1))[:n]
    if name == "G5":
        return bytes(97 + (v & 15) for v in _lcg(n))
    raise KeyError(name)
