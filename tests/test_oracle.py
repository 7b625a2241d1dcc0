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


# ---- codes the reference's tables do not define, and a second reading of the source ----------------------------
# tests/golden/js_model.py is a line-by-line Python model of src/inflate.ts with JavaScript's `undefined` semantics;
# tests/golden/undefined_codes.json holds hand-built streams whose expected output was derived from the source by hand
# (make_undefined_vectors.py).  The C oracle must agree with both.

@pytest.mark.parametrize("vec", T.undefined_code_vectors(), ids=lambda v: v[0])
def test_undefined_length_and_distance_codes(vec):
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import js_model as JS
    _, stream, expect = vec
    assert JS.inflate(stream) == expect
    assert O.inflate(stream) == expect  # no 'Data is corrupted': src/inflate.ts:98-117, 260-290


def test_oracle_agrees_with_the_js_model_on_damaged_streams():
    import os, random, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import js_model as JS

    def outcome(f, z):
        try:
            return ("ok", f(z))
        except (O.OracleError, JS.JsError) as e:
            return ("err", str(e))

    rnd = random.Random(1)
    d = T.gen("G5", 400)
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    streams = [zlib.compress(d, 6), O.deflate(d), zlib.compress(T.gen("G3", 100), 0), T.FIXED, T.DYNAMIC, co.compress(d) + co.flush()]
    streams += [s for _, s, _ in T.undefined_code_vectors()]
    for z in streams:
        for cut in range(0, len(z) + 1, 3):
            assert outcome(O.inflate, z[:cut]) == outcome(JS.inflate, z[:cut]), ("cut", cut, z.hex())
        for _ in range(60):
            zz = bytearray(z)
            for _k in range(rnd.choice([1, 1, 1, 2, 3])):
                zz[rnd.randrange(len(zz))] ^= 1 << rnd.randrange(8)
            assert outcome(O.inflate, bytes(zz)) == outcome(JS.inflate, bytes(zz)), bytes(zz).hex()


def test_stream_on_which_the_reference_never_returns():
    # past the end of the buffer the reference reads zero bits for ever unless a read() lands on the last bit of a byte
    # (src/utils/BitReadStream.ts:21-28,33-35); this damaged stream makes its symbol loop cycle without writing anything.
    # The oracle (and the GPU decoder) report it instead of hanging: 'stream never ends'.
    z = bytes.fromhex("789cfdde010900000080a0adfd3f5047c22592f50800000000")
    with pytest.raises(O.OracleError, match="stream never ends"):
        O.inflate(z)
