"""BASELINE.json configs at their full sizes, through size-independent properties:
deflate -> inflate round trip on the device, Adler-32 of the output == Adler-32 of the input ==
the stream's trailer, determinism, chunk independence; plus oracle / system-zlib parity and the
3 % size bound on bounded samples of every corpus (the oracle runs at ~15 MB/s)."""
import zlib

import numpy as np
import pytest

import oracle as O
import parity_cases as P

pytestmark = pytest.mark.gpu
KINDS = {"text": 0, "binary": 1, "random": 2, "mixed": 3}


@pytest.fixture(scope="module")
def c():
    import torch
    import zles
    codec = zles.Codec(0)
    # run on torch's current stream: the tests fill device tensors with torch right before calling the codec
    codec.set_stream(torch.cuda.current_stream().cuda_stream)
    return codec


def _corpus(c, kind, n, offset=0):
    import torch
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    c.dev_corpus(kind, offset, t.data_ptr(), n)
    return t


@pytest.mark.parametrize("kind", list(KINDS))
def test_host_and_device_corpora_agree(c, kind):
    n = 5 * 65536 + 123
    dev = _corpus(c, KINDS[kind], n, 777).cpu().numpy()
    assert (dev == c.host_corpus(KINDS[kind], 777, n)).all()


@pytest.mark.parametrize("kind", list(KINDS))
def test_corpus_sample_parity_and_size(c, kind):
    data = c.host_corpus(KINDS[kind], 0, 3 << 20).tobytes()
    z = P.roundtrip(c, data)  # ours / oracle / system zlib all decode it; size <= 1.03 x oracle
    P.inflate_matches_oracle(c, O.deflate(data[:1 << 20]))


def _device_round_trip(c, kind, n):
    import torch
    src = _corpus(c, kind, n)
    cap = c.deflate_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    olen = c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n)
    assert olen == n and torch.equal(src, back)
    head = comp[:2].cpu().numpy().tobytes()
    trailer = int.from_bytes(comp[clen - 4:clen].cpu().numpy().tobytes(), "big")
    assert head == b"\x78\x9c"
    assert trailer == c.dev_adler32(src.data_ptr(), n) == c.dev_adler32(back.data_ptr(), n)
    return src, comp, clen


def test_config2_64MiB_text_stream(c):
    n = 64 << 20
    src, comp, clen = _device_round_trip(c, KINDS["text"], n)
    z = comp[:clen].cpu().numpy().tobytes()
    raw = src.cpu().numpy().tobytes()
    assert zlib.decompress(z) == raw                      # system zlib accepts the whole stream
    assert zlib.adler32(raw) == int.from_bytes(z[-4:], "big")
    # determinism + chunk independence: a chunk-aligned slice compresses to the same bytes as inside the stream
    import torch
    comp2 = torch.empty_like(comp)
    assert c.dev_deflate(src.data_ptr(), n, comp2.data_ptr(), comp2.numel()) == clen
    assert torch.equal(comp[:clen], comp2[:clen])
    ref_sample = O.deflate(raw[:4 << 20])
    ours_sample = c.deflate(raw[:4 << 20])
    assert len(ours_sample) <= 1.03 * len(ref_sample)


def test_config3_batch_of_4KiB_buffers(c):
    import torch
    count = 262144
    n = count * 4096
    src = _corpus(c, KINDS["mixed"], n)
    in_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * 4096
    bound = c.deflate_bound(4096)
    out_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * bound
    out = torch.empty(count * bound, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(count, dtype=torch.int64, device="cuda")
    status = torch.ones(count, dtype=torch.int32, device="cuda")
    rc = c.dev_deflate_batch(src.data_ptr(), in_off.data_ptr(), count, out.data_ptr(), out_off.data_ptr(), out_len.data_ptr(), status.data_ptr())
    assert rc == 0 and int(status.abs().sum()) == 0
    # inflate the batch in place of the original offsets
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    # compact the streams so that stream i is [c_off[i], c_off[i+1])
    c_off = torch.zeros(count + 1, dtype=torch.int64, device="cuda")
    c_off[1:] = torch.cumsum(out_len, 0)
    total = int(c_off[-1])
    idx = torch.repeat_interleave(torch.arange(count, device="cuda"), out_len)
    pos = torch.arange(total, device="cuda") - c_off[idx] + out_off[idx]
    packed = out[pos]
    blen = torch.zeros(count, dtype=torch.int64, device="cuda")
    st2 = torch.ones(count, dtype=torch.int32, device="cuda")
    rc = c.dev_inflate_batch(packed.data_ptr(), c_off.data_ptr(), count, back.data_ptr(), in_off.data_ptr(), blen.data_ptr(), st2.data_ptr())
    assert rc == 0 and int(st2.abs().sum()) == 0 and bool((blen == 4096).all())
    assert torch.equal(src, back)
    # oracle / zlib parity and the size bound on a sample of the buffers
    host = src[: 64 * 4096].cpu().numpy().tobytes()
    sizes_ours = sizes_ref = 0
    ol = out_len[:64].cpu().numpy()
    oo = out_off[:64].cpu().numpy()
    outh = out[: 64 * bound].cpu().numpy()
    for i in range(64):
        raw = host[i * 4096:(i + 1) * 4096]
        z = outh[oo[i]:oo[i] + ol[i]].tobytes()
        assert zlib.decompress(z) == raw and O.inflate(z) == raw
        sizes_ours += len(z)
        sizes_ref += len(O.deflate(raw))
    assert sizes_ours <= 1.03 * sizes_ref
    assert total < n  # the corpus is compressible on the whole


def test_config4_1GiB_mixed_stream(c):
    _device_round_trip(c, KINDS["mixed"], 1 << 30)


def test_long_device_input_is_deflated_in_slabs(c):
    # zles_dev_deflate cuts an input of 2 GiB or more into slabs of 1 GiB that are matched and packed one after the other
    # (the token scratch is then a slab's, 4 GiB, not 4 bytes per byte of the whole input): same stream, byte for byte,
    # as without slabs (ZLES_NO_DEV_SLABS), and as with a buffer too small for the worst case (two-phase path)
    import os
    import torch
    import zles
    n = (2 << 30) + (3 << 20) + 12345
    src, comp, clen = _device_round_trip(c, KINDS["mixed"], n)
    os.environ["ZLES_NO_DEV_SLABS"] = "1"
    try:
        c2 = zles.Codec(0)
    finally:
        del os.environ["ZLES_NO_DEV_SLABS"]
    comp2 = torch.empty(clen, dtype=torch.uint8, device="cuda")  # exactly the size needed: below the bound
    assert c2.dev_deflate(src.data_ptr(), n, comp2.data_ptr(), clen) == clen
    assert torch.equal(comp[:clen], comp2)
    comp2.zero_()
    assert c.dev_deflate(src.data_ptr(), n, comp2.data_ptr(), clen) == clen  # the slabbed codec, no room for the bound
    assert torch.equal(comp[:clen], comp2)
