"""Parity of the CUDA path (libzles.so on a B200, through the C ABI) against the oracle.

Restates /root/reference/test/index.js: the 4 inflate known-answer vectors, and deflate
round trips through our inflate, the oracle's restatement of the reference's inflate and
system zlib (Node's inflateSync stand-in), plus SURVEY.md §8c inputs, edge lengths, the
lenient behaviours of the reference's inflate and the size bound (<= 1.03 x oracle).
Bar: bit-exact (byte / integer work, no tolerance).
"""
import zlib

import numpy as np
import pytest

import oracle as O
import parity_cases as P
import vectors as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c():
    import torch
    import zles
    codec = zles.Codec(0)
    # run on torch's current stream: the tests fill device tensors with torch right before calling the codec
    codec.set_stream(torch.cuda.current_stream().cuda_stream)
    return codec


def test_native_library_is_loaded(c):
    import zles
    assert zles._capi.LIB_PATH.endswith("zlib.es_b200/libzles.so")
    assert c.launches == 0 or c.launches > 0
    before = c.launches
    c.adler32(b"abc")
    assert c.launches > before  # kernels really launched


def test_inflate_kats(c):
    P.inflate_kats(c)


def test_inflate_fixture(c):
    P.inflate_fixture(c)


def test_adler_kats_and_sizes(c):
    P.adler_kats(c)
    assert c.adler32(T.fixture_raw()) == 0x140FA15B
    rng = np.random.default_rng(1)
    for n in (1, 2, 15, 16, 17, 31, 4095, 4096, 65521, 131072, 1 << 20, (1 << 24) + 13):
        d = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert c.adler32(d) == zlib.adler32(d)
    d = b"\xff" * (1 << 24)  # largest per-byte value: exercises the deferred modulo
    assert c.adler32(d) == zlib.adler32(d)


@pytest.mark.parametrize("row", T.V["model_table"], ids=lambda r: "%s-%d" % (r[0], r[1]))
def test_deflate_model_table(c, row):
    name, n = row[0], row[1]
    data = T.gen(name, n)
    # the 3 % size bound holds on every row: a chunk whose four blocks' headers cost more than they save is written as one
    # block (k_huff_merge), tiny inputs with the fixed code (k_huff)
    P.roundtrip(c, data)


@pytest.mark.parametrize("n", [0, 1, 131073, 262145])
def test_lengths_on_which_the_reference_throws(c, n):
    # zlib.es throws for these; we return a valid stream (documented difference)
    P.roundtrip(c, T.gen("G5", n), check_size=False)


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 255, 256, 257, 258, 259, 32767, 32768, 32769, 65535, 65536, 131071, 131072, 131074, 400000])
def test_ragged_lengths(c, n):
    P.roundtrip(c, T.gen("G5", n))


def test_batch_of_small_buffers_shares_sorts(c):
    # groups of 16 consecutive buffers of at most 4 KiB each share one pass of the match finder (BASELINE configs[2]):
    # every buffer must still compress to exactly the bytes it gets on its own, ragged and empty ones included
    rng = np.random.default_rng(7)
    lens = [4096] * 16 + [0, 1, 2, 3, 15, 16, 17, 100, 1000, 4095, 4096, 33, 2048, 7, 4000, 64] + [int(x) for x in rng.integers(0, 4097, size=85)]
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
    for _ in range(60):
        parts.append(T.gen("G5", int(rng.integers(1, 3000))))
        parts.append(bytes(int(rng.integers(1, 40000))) if rng.integers(0, 3) else b"ab" * int(rng.integers(1, 9000)))
    data = b"".join(parts)
    P.roundtrip(c, data)
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


def test_window_modes(c):
    # zles_ctx_set_window_mode: 1 (default) sorts blocks {0,1} and {2,3} of a chunk together; 0 gives block 2 the block
    # before it as window (three sorts per chunk): smaller or equal output, same decoders, same size bound
    data = T.fixture_raw() + T.gen("G5", 300000)
    z1 = P.roundtrip(c, data)
    c.set_window_mode(0)
    try:
        z0 = P.roundtrip(c, data)
    finally:
        c.set_window_mode(1)
    assert len(z0) <= len(z1)


def test_long_matches_and_overlaps(c):
    for d in (bytes(300000), b"a" * 70000, b"ab" * 50000, b"abc" * 40000, T.repeat_input() * 40, bytes(range(256)) * 600):
        P.roundtrip(c, d)


def test_tiny_inputs_take_the_cheapest_block_type(c):
    P.tiny_inputs_take_the_cheapest_block_type(c)


def test_incompressible(c):
    rng = np.random.default_rng(2)
    d = rng.integers(0, 256, size=300001, dtype=np.uint8).tobytes()
    z = P.roundtrip(c, d)
    assert len(z) < len(d) * 1.01


def test_skewed_histograms_deep_trees(c):
    # Fibonacci-like byte frequencies force Huffman trees deeper than 15 (7 for the code-length code):
    # the length-limited construction must still give complete codes that zlib accepts
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    rng = np.random.default_rng(3)
    syms = np.repeat(np.arange(24, dtype=np.uint8) * 7 + 3, fib)
    for size in (32768, 100000):
        d = rng.permutation(np.resize(syms, size)).tobytes()
        P.roundtrip(c, d)


def test_inflate_reference_streams(c):
    # streams made by (the restatement of) zlib.es itself: bit-concatenated blocks, no markers
    for name, n in [("RAW", 0), ("REPEAT", 0), ("G5", 4096), ("G5", 200000), ("G1", 131074), ("G3", 65536), ("FIXTURE", 0)]:
        P.inflate_matches_oracle(c, O.deflate(T.gen(name, n)))


def test_inflate_system_zlib_streams(c):
    data = T.gen("G5", 50000) + T.fixture_raw()[:150000]
    for level in range(0, 10):
        P.inflate_matches_oracle(c, zlib.compress(data, level))
    co = zlib.compressobj(6)
    z = co.compress(data[:20000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[20000:90000]) + co.flush(zlib.Z_FULL_FLUSH)
    z += co.compress(data[90000:]) + co.flush()
    P.inflate_matches_oracle(c, z)
    co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)
    P.inflate_matches_oracle(c, co.compress(data) + co.flush())  # fixed Huffman blocks only


def test_foreign_streams_decode_in_parallel(c):
    P.foreign_tier(c, 600000)


def test_lenient_and_errors(c):
    P.lenient_like_reference(c)
    P.error_strings(c)


def test_codes_the_reference_leaves_undefined(c):
    P.undefined_codes(c)


def test_output_full_protocol(c):
    P.output_full_protocol(c)


def test_truncated_and_corrupted_streams_match_the_reference(c):
    streams = P.damaged_streams(c)
    P.truncation_sweep(c, streams[:5], step=5)
    P.truncation_sweep(c, streams[5:], step=211)
    P.bitflip_sweep(c, streams, trials=60)


def test_marker_bytes_inside_data(c):
    payload = (b"\x00\x00\xff\xff" * 50 + b"abc") * 2000
    P.inflate_matches_oracle(c, zlib.compress(payload, 0))  # stored blocks: the pattern appears verbatim
    P.roundtrip(c, payload, check_size=False)
    # a marker pattern followed by something that looks like a stored block and then by garbage: that candidate
    # fails half-way (regression: the optimistic copy pass used its unset end position)
    bait = b"\x00\x00\xff\xff" + b"\x00\x05\x00\xfa\xffabcde" + b"\x07" + b"\x00\x00\xff\xff" + b"\x00\xff\x7f\x00\x80" + b"\x06"
    payload = (T.gen("G5", 3000) + bait) * 60
    P.inflate_matches_oracle(c, zlib.compress(payload, 0))
    P.inflate_matches_oracle(c, zlib.compress(payload, 1))
    # our own stream with a marker pattern spliced between two blocks' worth of compressed bytes cannot
    # be built by hand; instead check that removing candidates is exercised: a stream = ours + foreign tail
    z = c.deflate(T.gen("G5", 100000))
    assert c.inflate(z + b"\x00\x00\xff\xff" * 8) == T.gen("G5", 100000)


def test_batch(c):
    rng = np.random.default_rng(4)
    bufs = [T.gen("G5", 4096), b"", T.gen("G1", 100), T.gen("G3", 5000), T.gen("G5", 140000), b"x", T.fixture_raw()[:4096 * 3]]
    bufs += [rng.integers(0, 64, size=int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 9000, size=200)]
    zs = c.deflate_batch(bufs)
    assert len(zs) == len(bufs)
    for b, z in zip(bufs[:20], zs[:20]):
        assert zlib.decompress(z) == b and O.inflate(z) == b
        if len(b) <= 65536:
            assert z == c.deflate(b)          # the same bytes alone and in a batch
        else:
            assert len(z) <= len(c.deflate(b))  # a long buffer's third blocks get their windows in a batch: never larger
    for b, z in zip(bufs, zs):
        assert zlib.decompress(z) == b
    assert c.inflate_batch(zs) == bufs
    foreign = [zlib.compress(b, 6) for b in bufs[:10]] + [O.deflate(bufs[3])]
    assert c.inflate_batch(foreign) == bufs[:10] + [bufs[3]]
    res = c.inflate_batch([zs[0], b"\x77\x00", b"\x78\x9c\x07"], raise_on_error=False)
    assert res[0] == bufs[0] and str(res[1]) == "Not compressed by deflate" and str(res[2]) == "Not supported BTYPE : 3"


def test_batch_token_rows_follow_the_longest_block(c):
    P.batch_token_rows(c)


def test_sharded_phases_match_single_call(c):
    import torch
    import zles
    data = T.fixture_raw() + T.gen("G5", 300000)
    whole = c.deflate(data)
    cuts = [0, 131072, 131072 * 3, len(data)]
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    infos, parts = [], []
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        info = c.dev_deflate_phase1(src.data_ptr() + a, b - a, k == len(cuts) - 2)
        out = torch.zeros(info.comp_bytes + 8, dtype=torch.uint8, device="cuda")
        c.dev_deflate_phase2(out.data_ptr())
        parts.append(out[:info.comp_bytes].cpu().numpy().tobytes())
        infos.append(info)
    adler = zles.codec.combine_adler(infos)
    stream = b"\x78\x9c" + b"".join(parts) + adler.to_bytes(4, "big")
    assert stream == whole
    assert zlib.decompress(stream) == data
    # sharded inflate: every shard decodes on its own
    for k, part in enumerate(parts):
        d_in = torch.frombuffer(bytearray(part), dtype=torch.uint8).cuda()
        d_out = torch.zeros(cuts[k + 1] - cuts[k], dtype=torch.uint8, device="cuda")
        n = c.dev_inflate_segment(d_in.data_ptr(), len(part), d_out.data_ptr(), d_out.numel(), has_final=k == len(parts) - 1)
        assert n == cuts[k + 1] - cuts[k]
        assert d_out.cpu().numpy().tobytes() == data[cuts[k]:cuts[k + 1]]


def test_multi_device_context(c):
    # zles_mgpu_*: one process driving several devices; on a one-GPU box the same device is named three times, which
    # exercises the same sharding, exchange and stream-discovered inflate
    import torch
    import zles
    ndev = torch.cuda.device_count()
    mc = zles.MultiCodec(list(range(ndev)) if ndev > 1 else [0, 0, 0])
    try:
        P.multi_device(c, mc, 24 << 20)
    finally:
        mc.close()


def test_host_inflate_in_slabs(c):
    P.slabbed_host_inflate(c, 40 << 20)


def test_spurious_markers_in_slabs(c):
    P.spurious_markers_in_slabs(c)


def test_raw_deflate_and_gzip(c):
    P.wire_format_siblings(c, 20971527)


def test_host_batch_in_slabs(c):
    P.batch_in_slabs(c, 3000)


def test_large_pageable_buffers_are_staged(c):
    # what the N-API addon passes are pageable ArrayBuffers: at 32 MiB and more they go through the pinned staging ring and
    # its helper threads (stager.inl) — same bytes as with pinned buffers, in both directions, one and several devices
    import torch
    import zles
    n = 200 << 20
    with torch.cuda.stream(torch.cuda.current_stream()):
        src = torch.empty(n, dtype=torch.uint8, device="cuda")
        c.dev_corpus(3, 0, src.data_ptr(), n)
    pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
    pinned.copy_(src)
    torch.cuda.synchronize()
    pageable = np.array(pinned.numpy(), copy=True)
    cap = c.deflate_bound(n)
    z_pin = torch.empty(cap, dtype=torch.uint8).pin_memory()
    z_pag = np.empty(cap, dtype=np.uint8)
    m1 = c.deflate_into(pinned.numpy(), z_pin.numpy())
    m2 = c.deflate_into(pageable, z_pag)
    assert m1 == m2 and bool((z_pin.numpy()[:m1] == z_pag[:m2]).all())
    back = np.empty(n, dtype=np.uint8)
    for smin in (96 << 20, 1 << 40):       # streaming path and plain path
        c.set_stream_min(smin)
        try:
            back[:] = 0
            assert c.inflate_into(z_pag[:m2], back) == n and bool((back == pageable).all())
        finally:
            c.set_stream_min(96 << 20)
    mc = zles.MultiCodec([0, 0] if torch.cuda.device_count() < 2 else [0, 1])
    try:
        z2 = np.empty(cap, dtype=np.uint8)
        assert mc.deflate_into(pageable, z2) == m1 and bool((z2[:m1] == z_pag[:m1]).all())
        back[:] = 0
        assert mc.inflate_into(z2[:m1], back) == n and bool((back == pageable).all())
    finally:
        mc.close()
