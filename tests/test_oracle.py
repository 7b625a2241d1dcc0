"""Pins the CPU oracle (oracle/zlibes_oracle.c) before anything is graded against it.

* inflate: the reference's own 4 known-answer vectors (/root/reference/test/index.js:15-42).
* Adler-32: the trailers of those vectors.
* deflate: round trip through the oracle's inflate and system zlib (what
  test/index.js:56-109 asserts), plus the SURVEY.md §8c table produced by an
  independent restatement (exact bits of deflate are parity-unpinned by the
  reference itself).
"""
import hashlib
import zlib

import pytest

import oracle as O
import vectors as T


@pytest.mark.parametrize("name", ["UNCOMPRESSED", "FIXED", "DYNAMIC"])
def test_inflate_kat(name):
    assert O.inflate(getattr(T, name)) == T.RAW


def test_inflate_fixture():
    assert O.inflate(T.fixture_compressed()) == T.fixture_raw()


def test_adler_kat():
    assert "%08x" % O.adler32(T.RAW) == T.V["adler_kat"]["RAW"]
    assert "%08x" % O.adler32(T.fixture_raw()) == T.V["adler_kat"]["FIXTURE"]
    assert O.adler32(T.fixture_raw()) == zlib.adler32(T.fixture_raw())
    assert O.adler32(b"") == 1


@pytest.mark.parametrize("row", T.V["model_table"], ids=lambda r: "%s-%d" % (r[0], r[1]))
def test_deflate_model_table(row):
    name, n, adler, size, sha16, toks = row
    data = T.gen(name, n)
    assert len(data) == n
    assert "%08x" % O.adler32(data) == adler
    out = O.deflate(data)
    assert len(out) == size
    assert hashlib.sha256(out).hexdigest()[:16] == sha16
    assert [O.lz77_count(data, s, min(131072, n - s))[0] for s in range(0, n, 131072)] == toks
    # test/index.js:56-109: both decoders give the input back
    assert O.inflate(out) == data
    assert zlib.decompress(out) == data
    assert out[:2] == b"\x78\x9c"


@pytest.mark.parametrize("n", T.V["throwing_lengths"])
def test_deflate_throwing_lengths(n):
    with pytest.raises(O.OracleError, match="Data is corrupted"):
        O.deflate(T.gen("G1", n))


def test_dynamic_vector_prefix():
    # the DYNAMIC vector shares its first 21 bytes with our deflate(RAW)
    assert O.deflate(T.RAW)[:21] == T.DYNAMIC[:21]


def test_inflate_errors():
    with pytest.raises(O.OracleError, match="Not compressed by deflate"):
        O.inflate(b"\x77\x9c\x03\x00")
    with pytest.raises(O.OracleError, match="Not supported BTYPE : 3"):
        O.inflate(b"\x78\x9c\x07\x00\x00\x00\x00\x00")
    with pytest.raises(O.OracleError, match="Data is corrupted"):
        O.inflate(b"\x78\x9c\x01\x05\x00\x00\x00hello")  # LEN + NLEN != 65535


def test_inflate_system_zlib_levels():
    data = T.gen("G5", 50000) + T.fixture_raw()[:70000]
    for level in (0, 1, 6, 9):
        assert O.inflate(zlib.compress(data, level)) == data


def test_inflate_lenient_like_reference():
    # src/zlib.ts:22 — trailer never read: a corrupt Adler still "succeeds"
    z = bytearray(zlib.compress(b"hello hello hello hello"))
    z[-1] ^= 0xFF
    assert O.inflate(bytes(z)) == b"hello hello hello hello"
    # src/inflate.ts:287-290 — a distance before the start of the output yields zeros
    co = zlib.compressobj(wbits=-15, zdict=b"abcdefgh")
    body = co.compress(b"abcdefghabcdefgh") + co.flush()
    out = O.inflate(b"\x78\x9c" + body)
    assert out == (b"a" + bytes(7)) * 2  # literal 'a', then length-15 match at distance 8 into nothing
