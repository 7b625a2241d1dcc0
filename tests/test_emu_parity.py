"""The CUDA sources run through the CPU thread emulator (tests/emu) at small sizes.

This is NOT the parity gate (that is tests/test_gpu_*.py on a B200); it keeps the kernels'
logic — barriers, warp votes, indexing, the host-side chain walk — checked in a container
without a GPU.  The emulator library is loaded only here (tests/emu_lib.py).
"""
import zlib

import numpy as np
import pytest

import emu_lib
import oracle as O
import parity_cases as P
import vectors as T

pytestmark = pytest.mark.emu


@pytest.fixture(scope="module")
def c():
    return emu_lib.codec()


def test_inflate_kats(c):
    P.inflate_kats(c)


def test_inflate_fixture(c):
    P.inflate_fixture(c)


def test_adler(c):
    P.adler_kats(c)
    for n in (1, 15, 16, 17, 4095, 70001):
        d = T.gen("G3", n)
        assert c.adler32(d) == O.adler32(d)


@pytest.mark.parametrize("name,n", [("RAW", 0), ("REPEAT", 0), ("G1", 0), ("G1", 1), ("G1", 2), ("G1", 3), ("G1", 4096), ("G2", 4096),
                                    ("G3", 4096), ("G4", 4096), ("G5", 4096), ("G5", 32768), ("G5", 32769), ("G5", 70000), ("G1", 131073),
                                    ("G5", 140000), ("G2", 200000)])
def test_deflate_roundtrip(c, name, n):
    data = T.gen(name, n)
    P.roundtrip(c, data)


def test_batch_token_rows_follow_the_longest_block(c):
    P.batch_token_rows(c)


def test_tiny_inputs_take_the_cheapest_block_type(c):
    P.tiny_inputs_take_the_cheapest_block_type(c)


def test_batch_of_small_buffers_shares_sorts(c):
    # groups of 16 consecutive buffers of at most 4 KiB each share one pass of the match finder (BASELINE configs[2]):
    # every buffer must still compress to exactly the bytes it gets on its own, ragged and empty ones included
    rng = np.random.default_rng(7)
    lens = [4096] * 16 + [0, 1, 2, 3, 15, 16, 17, 100, 1000, 4095, 4096, 33, 2048, 7, 4000, 64] + [int(x) for x in rng.integers(0, 4097, size=21)]
    src = T.fixture_raw() + T.gen("G5", 200000)
    bufs, o = [], 0
    for n in lens:
        bufs.append(src[o:o + n]); o = (o + n + 37) % (len(src) - 5000)
    zs = c.deflate_batch(bufs)
    for b, z in zip(bufs, zs):
        assert zlib.decompress(z) == b
        assert z == c.deflate(b)
    assert c.inflate_batch(zs) == bufs


def test_long_matches_between_text(c):
    # runs of zeros (258-byte matches at distance 1) between stretches of text: a batch of 32 such tokens is wider than
    # the small shared-memory mirror of the piece-parallel copy pass (regression: its slots aliased inside one batch)
    rng = np.random.default_rng(11)
    parts = []
    for _ in range(14):
        parts.append(T.gen("G5", int(rng.integers(1, 3000))))
        parts.append(bytes(int(rng.integers(1, 40000))) if rng.integers(0, 3) else b"ab" * int(rng.integers(1, 9000)))
    data = b"".join(parts)
    P.roundtrip(c, data, check_size=False)
    P.inflate_matches_oracle(c, zlib.compress(data, 6))


def test_stored_block_followed_by_a_final_empty_one(c):
    # what system zlib writes at level 0 for exactly 32 KiB: a NON-final stored block, then a final empty stored block
    # (regression: the payload was looked for 5 bytes too late, as if the data block itself had been the final one)
    import struct
    for data in (T.fixture_raw()[5000:5000 + 32768], T.gen("G5", 1000), T.fixture_raw()[:32768], b"\x00\x00\xff\xff" * 300):
        n = len(data)
        z = b"\x78\x01" + b"\x00" + struct.pack("<HH", n, n ^ 0xffff) + data + b"\x01\x00\x00\xff\xff" + struct.pack(">I", zlib.adler32(data))
        assert zlib.decompress(z) == data
        P.inflate_matches_oracle(c, z)
        assert c.inflate(z) == data


def test_random_round_trips_small(c):
    # the randomised checks of tests/test_gpu_stress.py at sizes the emulator finishes in seconds
    import stress_cases as S
    rng = np.random.default_rng(3)
    raw = T.fixture_raw()
    for i in range(14):
        d = S.make(rng, S.size(rng, 17) % 70000, raw)
        z = c.deflate(d)
        assert zlib.decompress(z) == d and c.inflate(z) == d, (i, len(d))
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, 15, 8, int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_RLE, zlib.Z_FIXED])))
        f = co.compress(d) + co.flush()
        assert c.inflate(f) == d, (i, len(d))
        if rng.integers(0, 2):
            s = bytearray(z if rng.integers(0, 2) else f)
            if len(s) > 8:
                s[int(rng.integers(2, len(s)))] ^= 1 << int(rng.integers(0, 8))
            P.inflate_matches_oracle(c, bytes(s))


def test_window_modes(c):
    data = T.fixture_raw()[:150000] + T.gen("G5", 120000)
    z1 = P.roundtrip(c, data, check_size=False, oracle_decode=False)
    c.set_window_mode(0)
    try:
        z0 = P.roundtrip(c, data, check_size=False, oracle_decode=False)
    finally:
        c.set_window_mode(1)
    assert len(z0) <= len(z1)


def test_host_deflate_in_slabs(c):
    # long enough for the host-buffer deflate to run slab by slab (the emulator has 4 SMs: slabs of 8 blocks): every
    # slab is laid out, packed and copied out on its own; the stream must equal the device-resident form's
    data = T.gen("G5", 300000) + T.fixture_raw()[:200000] + bytes(70000) + T.gen("G3", 40000)
    z = P.roundtrip(c, data, check_size=False, oracle_decode=False)
    import numpy as np
    out = np.zeros(len(z) - 1, dtype=np.uint8)  # one byte short: the size query protocol still holds
    with pytest.raises(Exception) as e:  # ZLES_E_OUTPUT_FULL (the caller's buffer is one byte short)
        c.deflate_into(data, out)
    assert getattr(e.value, "code", None) == 16


def test_fixture_roundtrip_and_size(c):
    z = P.roundtrip(c, T.fixture_raw())
    assert len(z) <= 1.03 * T.V["fixture"]["oracle_deflate_size"] if "oracle_deflate_size" in T.V["fixture"] else True


def test_inflate_reference_streams(c):
    # zlib.es's own output: bit-concatenated 128 KiB blocks, no markers -> sequential path
    for name, n in [("G5", 4096), ("G5", 140000), ("G1", 131074)]:
        P.inflate_matches_oracle(c, O.deflate(T.gen(name, n)))


def test_inflate_system_zlib_streams(c):
    data = T.gen("G5", 30000) + T.fixture_raw()[:40000]
    for level in (0, 1, 6, 9):
        P.inflate_matches_oracle(c, zlib.compress(data, level))
    co = zlib.compressobj(6)
    z = co.compress(data[:20000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[20000:]) + co.flush(zlib.Z_FULL_FLUSH) + co.flush()
    P.inflate_matches_oracle(c, z)  # markers with history across them: must not be taken for our format


def test_foreign_streams_decode_in_parallel(c):
    P.foreign_tier(c, 300000)


def test_lenient_and_errors(c):
    P.lenient_like_reference(c)
    P.error_strings(c)


def test_codes_the_reference_leaves_undefined(c):
    P.undefined_codes(c)


def test_truncated_and_corrupted_streams_match_the_reference(c):
    streams = P.damaged_streams(c)
    P.truncation_sweep(c, streams[:5], step=41)
    P.truncation_sweep(c, streams[5:], step=997)
    P.bitflip_sweep(c, streams, trials=4)


def test_output_full_protocol(c):
    P.output_full_protocol(c)


def test_false_marker_in_payload(c):
    # the marker bytes 00 00 FF FF inside a stored block are not block boundaries
    payload = (b"\x00\x00\xff\xff" * 50 + b"abc") * 20
    P.inflate_matches_oracle(c, zlib.compress(payload, 0))


def test_batch(c):
    bufs = [T.gen("G5", 4096), b"", T.gen("G1", 100), T.gen("G3", 5000), T.gen("G5", 40000), b"x"]
    zs = c.deflate_batch(bufs)
    for b, z in zip(bufs, zs):
        assert zlib.decompress(z) == b
        assert z == c.deflate(b)  # a batch entry is exactly the single-call stream
    assert c.inflate_batch(zs) == bufs
    mixed = [zlib.compress(bufs[0], 6), O.deflate(bufs[3]), zs[4]]
    assert c.inflate_batch(mixed) == [bufs[0], bufs[3], bufs[4]]
    res = c.inflate_batch([zs[0], b"\x77\x00"], raise_on_error=False)
    assert res[0] == bufs[0] and str(res[1]) == "Not compressed by deflate"


def test_sharded_phases_match_single_call(c):
    # independent chunks: compressing shards separately and concatenating gives the single-call bytes
    import zles
    from zles import _capi
    data = T.gen("G5", 300000)
    whole = c.deflate(data)
    cuts = [0, 131072, 262144, len(data)]
    arr = np.frombuffer(data, dtype=np.uint8)
    infos, parts = [], []
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        seg = np.ascontiguousarray(arr[a:b])
        info = c.dev_deflate_phase1(seg.ctypes.data, b - a, k == len(cuts) - 2)
        out = np.zeros(info.comp_bytes + 8, dtype=np.uint8)
        c.dev_deflate_phase2(out.ctypes.data)
        parts.append(out[:info.comp_bytes].tobytes())
        infos.append(info)
    adler = zles.codec.combine_adler(infos, c.L)
    stream = b"\x78\x9c" + b"".join(parts) + adler.to_bytes(4, "big")
    assert stream == whole
    assert zlib.decompress(stream) == data


def test_multi_device_context(c):
    import zles
    mc = zles.MultiCodec([0, 0, 0], lib=c.L)
    try:
        P.multi_device(c, mc, 1200000)
    finally:
        mc.close()


def test_host_inflate_in_slabs(c):
    P.slabbed_host_inflate(c, 1500000)


def test_raw_deflate_and_gzip(c):
    P.wire_format_siblings(c, 300001)


def test_spurious_markers_in_slabs(c):
    P.spurious_markers_in_slabs(c)


def test_host_batch_in_slabs(c):
    P.batch_in_slabs(c, 40)
