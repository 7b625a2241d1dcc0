"""Parity cases shared by the GPU tests (tests/test_gpu_*.py, through libzles.so on a B200)
and the emulator tests (tests/test_emu_*.py, same sources on the CPU thread emulator,
small sizes only).  Every function takes a ``zles.Codec``.

They restate /root/reference/test/index.js in pytest form: the 4 inflate known-answer
vectors, and round trips of deflate output through (a) our inflate, (b) the oracle's
restatement of the reference's inflate and (c) system zlib (the stand-in for Node's
zlib.inflateSync, test/index.js:68,83,105), plus the size bound against the oracle.
"""
import zlib

import oracle as O
import vectors as T

SIZE_SLACK = 1.03  # BASELINE.json north_star: compressed size within 3 % of the reference's


def inflate_kats(c):
    for name in ("UNCOMPRESSED", "FIXED", "DYNAMIC"):  # test/index.js:15-35
        assert c.inflate(getattr(T, name)) == T.RAW, name


def inflate_fixture(c):  # test/index.js:37-42
    assert c.inflate(T.fixture_compressed()) == T.fixture_raw()


def adler_kats(c):
    assert c.adler32(T.RAW) == 0x2B23056C
    assert c.adler32(b"") == 1
    assert c.adler32(b"\x00") == O.adler32(b"\x00")


def roundtrip(c, data: bytes, check_size: bool = True, oracle_decode: bool = True):
    """deflate(data) decodes to data under all three decoders; size <= 1.03 x oracle's."""
    z = c.deflate(data)
    assert z[:2] == b"\x78\x9c"  # src/zlib.ts:28-34
    assert int.from_bytes(z[-4:], "big") == zlib.adler32(data)  # src/zlib.ts:36-40
    assert zlib.decompress(z) == data  # Node's inflateSync stand-in (verifies the trailer too)
    if oracle_decode:
        assert O.inflate(z) == data
    assert c.inflate(z) == data
    if check_size:
        try:
            ref = O.deflate(data)
        except O.OracleError:
            ref = None  # the reference throws on this length; nothing to compare
        if ref is not None:
            assert len(z) <= SIZE_SLACK * len(ref), (len(z), len(ref))
    return z


def tiny_inputs_take_the_cheapest_block_type(c):
    """the reference always writes a dynamic block (src/deflate.ts:28; its stored-block writer is dead
    code, src/deflate.ts:41-54).  We write whichever of stored / fixed / dynamic is smallest: never larger than the
    reference's stream, and never larger than system zlib -6 by more than a byte, on inputs from 1 byte to 4 KiB."""
    seen = set()
    for name in ("G1", "G3", "G5", "RAWx"):
        for n in (1, 2, 3, 5, 10, 30, 64, 100, 200, 400, 800, 1500, 3000, 4096):
            d = (T.RAW * 200)[:n] if name == "RAWx" else T.gen(name, n)
            z = roundtrip(c, d)
            seen.add((z[2] >> 1) & 3)  # BTYPE of the first block
            assert len(z) <= len(zlib.compress(d, 6)) + 1, (name, n, len(z))
            if n > 1:
                assert len(z) <= len(O.deflate(d)), (name, n)
    assert seen == {0, 1, 2}, seen  # all three block types were chosen somewhere


def batch_token_rows(c):
    """A batch's token scratch has one row per block, as long as the batch's longest block (not 32 Ki slots: 262,144
    buffers of 4 KiB would take 32 GiB): batches whose longest buffer is 1, 17, 1,003 and 5,000 bytes, and one with a
    buffer of several blocks, give every buffer the bytes it gets on its own."""
    import numpy as np
    rng = np.random.default_rng(21)
    src = T.fixture_raw() + T.gen("G5", 100000)
    for longest, count in ((1, 5), (17, 40), (1003, 70), (5000, 40), (70000, 6)):
        lens = [longest] + [int(x) for x in rng.integers(0, longest + 1, size=count - 1)]
        bufs = []
        for n in lens:
            o = int(rng.integers(0, len(src) - n))
            bufs.append(src[o:o + n])
        zs = c.deflate_batch(bufs)
        for b, z in zip(bufs, zs):
            assert zlib.decompress(z) == b
        for i in (0, 1, count // 2, count - 1):
            assert zs[i] == c.deflate(bufs[i]), (longest, i)
        assert c.inflate_batch(zs) == bufs


def inflate_matches_oracle(c, stream: bytes):
    """Same bytes or the same error as the reference's inflate."""
    try:
        want = O.inflate(stream)
    except O.OracleError as e:
        want = e
    try:
        got = c.inflate(stream)
    except Exception as e:  # ZlesError
        got = e
    if isinstance(want, Exception):
        assert isinstance(got, Exception), "reference throws %r, we returned %d bytes" % (str(want), len(got))
        assert str(got) == str(want)
    else:
        assert not isinstance(got, Exception), "we raise %r, reference returns %d bytes" % (str(got), len(want))
        assert got == want


def lenient_like_reference(c):
    # src/zlib.ts:22 — the Adler-32 trailer is never read
    z = bytearray(zlib.compress(b"hello hello hello hello"))
    z[-1] ^= 0xFF
    assert c.inflate(bytes(z)) == b"hello hello hello hello"
    # trailer absent / trailing garbage
    z = zlib.compress(b"hello hello hello hello")
    assert c.inflate(z[:-4]) == b"hello hello hello hello"
    assert c.inflate(z + b"garbage") == b"hello hello hello hello"
    # src/inflate.ts:287-290 — a distance before the start of the output yields zeros
    co = zlib.compressobj(wbits=-15, zdict=b"abcdefgh")
    body = co.compress(b"abcdefghabcdefgh") + co.flush()
    inflate_matches_oracle(c, b"\x78\x9c" + body)
    # FDICT / FCHECK ignored (src/zlib.ts:17-20)
    inflate_matches_oracle(c, b"\x78\xff" + zlib.compress(b"abc" * 50)[2:])


def undefined_codes(c):
    """Literal/length symbols 286/287 copy nothing, distance codes 30.. write zeros — neither is an error in the reference
    (/root/reference/src/inflate.ts:98-117, 260-290).  Hand-built vectors (tests/golden/make_undefined_vectors.py), then
    every prefix and random bit flips of them against the oracle, and a stream on which the reference never returns."""
    vecs = T.undefined_code_vectors()
    for name, stream, expect in vecs:
        assert c.inflate(stream) == expect, name
        assert O.inflate(stream) == expect, name
    truncation_sweep(c, [s for _, s, _ in vecs])
    bitflip_sweep(c, [s for _, s, _ in vecs], trials=25, seed=9)
    runaway = bytes.fromhex("789cfdde010900000080a0adfd3f5047c22592f50800000000")
    inflate_matches_oracle(c, runaway)
    try:
        c.inflate(runaway)
    except Exception as e:
        assert str(e) == "stream never ends" and getattr(e, "code", None) == 20
    else:
        raise AssertionError("no error for the runaway stream")
    assert c.inflate_batch([vecs[0][1], runaway, vecs[2][1]], raise_on_error=False)[::2] == [vecs[0][2], vecs[2][2]]


def error_strings(c):
    for stream, msg in [(b"\x77\x9c\x03\x00", "Not compressed by deflate"),   # src/zlib.ts:15
                        (b"", "Not compressed by deflate"),
                        (b"\x78\x9c\x07\x00\x00\x00\x00\x00", "Not supported BTYPE : 3"),  # src/inflate.ts:32
                        (b"\x78\x9c\x01\x05\x00\x00\x00hello", "Data is corrupted")]:    # src/inflate.ts:50
        try:
            c.inflate(stream)
        except Exception as e:
            assert str(e) == msg, (stream, str(e))
        else:
            raise AssertionError("no error for %r" % stream)


def output_full_protocol(c):
    """ZLES_E_OUTPUT_FULL reports the size needed; a second call with that size succeeds."""
    import numpy as np
    data = T.gen("G5", 100000)
    small = np.zeros(10, dtype=np.uint8)
    try:
        c.deflate_into(data, small)
    except Exception as e:
        assert getattr(e, "code", None) == 16, e
    else:
        raise AssertionError("deflate into 10 bytes succeeded")
    z = c.deflate(data)
    for cap in (0, 1, 50000, 99999):
        out = np.zeros(max(cap, 1), dtype=np.uint8)[:cap]
        try:
            c.inflate_into(z, out)
        except Exception as e:
            assert getattr(e, "code", None) == 16, (cap, e)
        else:
            raise AssertionError("inflate into %d bytes succeeded" % cap)
    out = np.zeros(100000, dtype=np.uint8)
    assert c.inflate_into(z, out) == 100000 and out.tobytes() == data
    # a foreign stream (sequential path) follows the same protocol
    zf = zlib.compress(data, 6)
    try:
        c.inflate_into(zf, np.zeros(5000, dtype=np.uint8))
    except Exception as e:
        assert getattr(e, "code", None) == 16, e
    else:
        raise AssertionError("inflate of a foreign stream into 5000 bytes succeeded")
    assert c.inflate_into(zf, out) == 100000 and out.tobytes() == data


def _outcome(f, z):
    try:
        return ("ok", f(z))
    except Exception as e:
        return ("err", str(e))


def truncation_sweep(c, streams, step=1):
    """Every prefix of a stream gives the reference's outcome: the same error text, or — the reference returns
    partial data without an error when the last bit of the buffer is consumed by a Huffman-code read inside a
    final block (BitReadStream.isEnd, /root/reference/src/utils/BitReadStream.ts:14-31, src/inflate.ts:34-36,76,237)
    — the same partial bytes, including the effect of the stale bit read() leaves behind."""
    for z in streams:
        for cut in range(0, len(z), step):
            a, b = _outcome(O.inflate, z[:cut]), _outcome(c.inflate, z[:cut])
            assert a == b, (cut, a[0], a[1][-16:] if a[0] == "ok" else a[1], b[0], b[1][-16:] if b[0] == "ok" else b[1])


def bitflip_sweep(c, streams, trials, seed=5):
    """Single-bit corruptions: same garbage or same error as the reference."""
    import random
    rnd = random.Random(seed)
    for z in streams:
        for _ in range(trials):
            pos = rnd.randrange(2, len(z))
            zz = bytearray(z)
            zz[pos] ^= 1 << rnd.randrange(8)
            a, b = _outcome(O.inflate, bytes(zz)), _outcome(c.inflate, bytes(zz))
            assert a == b, (pos, a[0], a[1][-16:] if a[0] == "ok" else a[1], b[0], b[1][-16:] if b[0] == "ok" else b[1])


def damaged_streams(c):
    d = T.gen("G5", 3000)
    return [c.deflate(d), zlib.compress(d, 6), O.deflate(d), zlib.compress(T.gen("G3", 300), 0), T.FIXED, c.deflate(T.gen("G5", 40000))]


def foreign_tier(c, n):
    """Streams of other encoders — zlib.es's own bit-concatenated blocks (no markers), system zlib at several
    levels (32 KiB history across blocks, stored and empty blocks) — are decoded by the block-parallel path
    (k_hdr_filter / k_fblk_map / k_fpiece_sym and the symbolic-window passes), not by the sequential warp, and
    give the reference's bytes."""
    import numpy as np
    rng = np.random.default_rng(11)
    data = (T.gen("G5", n // 2) + T.fixture_raw() + rng.integers(0, 256, 20000, dtype=np.uint8).tobytes())[:n]
    co = zlib.compressobj(6)
    flushed = co.compress(data[:50000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[50000:]) + co.flush()
    cf = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)   # fixed-Huffman blocks only (src/inflate.ts:57-118)
    fixed = cf.compress(data) + cf.flush()
    # fixed, stored and dynamic blocks in one stream: raw-deflate pieces, each flushed to a byte boundary, under one zlib frame
    def raw_piece(b, level, strategy, last):
        cr = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        return cr.compress(b) + (cr.flush() if last else cr.flush(zlib.Z_FULL_FLUSH))
    mixed = (b"\x78\x9c" + raw_piece(data[:70000], 6, zlib.Z_FIXED, False) + raw_piece(data[70000:90000], 0, zlib.Z_DEFAULT_STRATEGY, False)
             + raw_piece(data[90000:], 6, zlib.Z_DEFAULT_STRATEGY, True) + zlib.adler32(data).to_bytes(4, "big"))
    assert zlib.decompress(mixed) == data
    for name, z in [("zlib.es", O.deflate(data)), ("zlib1", zlib.compress(data, 1)), ("zlib6", zlib.compress(data, 6)),
                    ("zlib9", zlib.compress(data, 9)), ("zlib0", zlib.compress(data[:200000], 0)), ("sync-flush", flushed),
                    ("fixture", T.fixture_compressed()), ("fixed", fixed), ("fixed+stored+dynamic", mixed)]:
        c.set_timing(True)
        out = c.inflate(z)
        used = c.kernel_time("k_fpiece_sym")[1]
        seq = c.kernel_time("k_inflate")[1]
        c.set_timing(False)
        assert out == O.inflate(z), name
        assert used >= 1 and seq == 0, (name, used, seq)


def multi_device(c, mc, n):
    """One process, several devices (zles_mgpu_*): the stream is bit for bit the single-device one, inflate finds its
    shards from the stream itself, other encoders' streams and damaged ones give the reference's outcome through the
    fallback, and the output-size protocol holds.  `mc` may name the same device several times (one-GPU boxes)."""
    import numpy as np
    rng = np.random.default_rng(5)
    data = (T.fixture_raw() + T.gen("G5", n // 2) + bytes(n // 8) + rng.integers(0, 256, n // 8, dtype=np.uint8).tobytes() + T.gen("G5", n))[:n]
    mc.set_min_shard(65536)
    z = mc.deflate(data)
    assert z == c.deflate(data)
    assert zlib.decompress(z) == data
    for slab in (0, 8, 64):
        mc.set_slab_blocks(slab)
        assert mc.inflate(z) == data, slab
    mc.set_slab_blocks(0)
    for cut in (n // 3, 131072 * 3, 131072 * 3 + 1, 100, 0):  # ragged sizes, fewer chunks than devices, empty
        d = data[:cut]
        zz = mc.deflate(d)
        assert zz == c.deflate(d) and mc.inflate(zz) == d, cut
    assert mc.inflate(zlib.compress(data, 6)) == data
    assert mc.inflate(O.deflate(data[:300000])) == data[:300000]
    small = np.zeros(n - 1, dtype=np.uint8)
    try:
        mc.inflate_into(z, small)
    except Exception as e:
        assert getattr(e, "code", None) == 16, e
    else:
        raise AssertionError("inflate into n - 1 bytes succeeded")
    out = np.zeros(n, dtype=np.uint8)
    assert mc.inflate_into(z, out) == n and out.tobytes() == data
    for cut in (len(z) // 2, len(z) - 3, 100):
        assert _outcome(O.inflate, z[:cut]) == _outcome(mc.inflate, z[:cut]), cut
    zz = bytearray(z)
    zz[len(z) // 3] ^= 4
    assert _outcome(O.inflate, bytes(zz)) == _outcome(mc.inflate, bytes(zz))


def slabbed_host_inflate(c, n):
    """zles_inflate of one of our streams decodes slab by slab (copies back overlapped): same bytes for every slab size."""
    data = (T.gen("G5", n // 2) + T.fixture_raw() * 3 + bytes(n))[:n]
    z = c.deflate(data)
    for slab in (4, 8, 12, 0):
        c.set_slab_blocks(slab)
        try:
            assert c.inflate(z) == data, slab
        finally:
            c.set_slab_blocks(0)
    # the same with the stream copied in pieces that are scanned and decoded as they land
    import numpy as np
    zf = zlib.compress(data, 1)
    for smin, slab in ((262144, 8), (400000, 16), (len(z) // 2, 4)):
        c.set_stream_min(smin)
        c.set_slab_blocks(slab)
        try:
            assert c.inflate(z) == data, (smin, slab)
            assert c.inflate(zf) == data                     # not ours: falls through to the general path
            assert _outcome(O.inflate, z[:len(z) * 2 // 3]) == _outcome(c.inflate, z[:len(z) * 2 // 3])
            try:
                c.inflate_into(z, np.zeros(n - 1, dtype=np.uint8))
            except Exception as e:
                assert getattr(e, "code", None) == 16, e
            else:
                raise AssertionError("inflate into n - 1 bytes succeeded")
        finally:
            c.set_stream_min(96 << 20)
            c.set_slab_blocks(0)


def spurious_markers_in_slabs(c):
    """Incompressible data that contains the marker bytes 00 00 FF FF: its blocks are stored, so the pattern appears
    verbatim in the stream and the marker scan reports block starts that are none.  The slab-by-slab host inflate (which
    decodes from the list of starts the host holds) must notice and hand the stream to the general path: same bytes for
    every slab size, with and without the piece-wise copy in."""
    import numpy as np
    rng = np.random.default_rng(5)
    a = bytearray(rng.integers(0, 256, size=600000, dtype=np.uint8).tobytes())
    for o in range(1000, len(a) - 10, 7001):
        a[o:o + 4] = b"\x00\x00\xff\xff"
    d = bytes(a) + T.gen("G5", 200000)
    z = c.deflate(d)
    assert zlib.decompress(z) == d
    try:
        for smin, slab in ((1 << 40, 0), (1 << 40, 4), (1 << 40, 8), (200000, 4), (300000, 8)):
            c.set_stream_min(smin)
            c.set_slab_blocks(slab)
            assert c.inflate(z) == d, (smin, slab)
    finally:
        c.set_stream_min(96 << 20)
        c.set_slab_blocks(0)


def wire_format_siblings(c, n):
    """raw deflate data (the reference's cores, /root/reference/src/deflate.ts:14 and src/inflate.ts:16 with its
    `offset` parameter) and gzip (RFC 1952, CRC-32) behind the same kernels — against system zlib / Python gzip / the oracle."""
    import gzip
    import io
    import os
    rnd = os.urandom
    for m in (0, 1, 5, 15, 16, 17, 127, 128, 129, 32767, 32768, 32769, 100000, n):
        d = rnd(m)
        assert c.crc32(d) == zlib.crc32(d), m
        assert c.crc32(memoryview(b"xyz" + d)[3:]) == zlib.crc32(d), m  # a source that is not 16-byte aligned
    a, b = rnd(1000), rnd(77777)
    assert c.L.zles_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
    data = (T.fixture_raw() + T.gen("G5", n))[:n]
    r = c.deflate_raw(data)
    assert r == c.deflate(data)[2:-4]                     # the same deflate data as inside the zlib container
    assert zlib.decompress(r, -15) == data and O.inflate_raw(r) == data
    assert c.inflate_raw(r) == data and c.inflate_raw(b"junk!" + r, 5) == data   # inflate(input, offset), src/inflate.ts:16
    assert c.inflate_raw(zlib.compress(data, 6)[2:-4]) == data
    zo = O.deflate(data[:200000])
    assert c.inflate_raw(zo, 2) == data[:200000]
    g = c.gzip_deflate(data)
    assert gzip.decompress(g) == data and zlib.decompress(g, 31) == data
    assert g[10:-8] == r
    assert c.gzip_inflate(g) == data
    assert c.gzip_inflate(gzip.compress(data, 6)) == data
    buf = io.BytesIO()
    with gzip.GzipFile(filename="name.txt", mode="wb", fileobj=buf, mtime=12345) as f:
        f.write(data)
    assert c.gzip_inflate(buf.getvalue()) == data         # FNAME set
    assert c.gzip_inflate(c.gzip_deflate(b"")) == b""
    for damage, code in ((lambda z: z[:-6] + bytes([z[-6] ^ 1]) + z[-5:], 21), (lambda z: z[:-1] + bytes([z[-1] ^ 1]), 21),
                         (lambda z: b"\x1f\x8b\x07" + z[3:], 1), (lambda z: z[:5], 5)):
        try:
            c.gzip_inflate(damage(g))
        except Exception as e:
            assert getattr(e, "code", None) == code, (code, e)
        else:
            raise AssertionError("no error")


def batch_in_slabs(c, count):
    """The host batch calls go through the device a slab of buffers at a time (results compacted on the device, moved into
    the caller's buffers by a helper thread): same results whatever the slab size, errors and size retries included."""
    import numpy as np
    rng = np.random.default_rng(21)
    src = T.fixture_raw() + T.gen("G5", 300000)
    bufs = []
    for i in range(count):
        n = int(rng.integers(0, 6000)) if i % 7 else int(rng.integers(30000, 70000))
        o = int(rng.integers(0, len(src) - n))
        bufs.append(src[o:o + n])
    want = None
    for slab in (16384, 5, 64):
        c.set_batch_slab(slab)
        try:
            zs = c.deflate_batch(bufs)
            if want is None:
                want = zs
                for b, z in zip(bufs[:12], zs[:12]):
                    assert zlib.decompress(z) == b and z == c.deflate(b)
            assert zs == want, slab
            assert c.inflate_batch(zs) == bufs, slab
            mixed = list(zs)
            mixed[3] = b"\x77\x00"
            mixed[count // 2] = zlib.compress(bytes(100000), 9)      # needs far more room than 10 x its length: retried
            res = c.inflate_batch(mixed, raise_on_error=False)
            assert str(res[3]) == "Not compressed by deflate" and res[count // 2] == bytes(100000)
            assert all(r == b for i, (r, b) in enumerate(zip(res, bufs)) if i not in (3, count // 2)), slab
        finally:
            c.set_batch_slab(16384)
