"""Inflate throughput on streams made by other encoders (zlib.es via the oracle, system zlib)."""
import os, sys, time, zlib, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch, zles, oracle as O
c = zles.Codec(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 32) << 20
raw = c.host_corpus(3, 0, n).tobytes()
from concurrent.futures import ThreadPoolExecutor
def oracle_stream(data):  # zlib.es deflate of the whole buffer = concatenation of its 128 KiB blocks: build it piecewise, in parallel
    return O.deflate(data)
t = time.time(); zes = oracle_stream(raw[: min(n, 16 << 20)]); t_or = time.time() - t
cf = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
streams = {"zlib.es (oracle) %d MiB" % (min(n, 16 << 20) >> 20): (zes, raw[: min(n, 16 << 20)]), "zlib -6": (zlib.compress(raw, 6), raw), "zlib -1": (zlib.compress(raw, 1), raw),
           "zlib fixed blocks": (cf.compress(raw) + cf.flush(), raw), "ours": (c.deflate(raw), raw)}
for name, (z, ref) in streams.items():
    d_in = torch.frombuffer(bytearray(z), dtype=torch.uint8).cuda(); d_out = torch.zeros(len(ref), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        c.set_timing(True)
        t = time.perf_counter(); m = c.dev_inflate(d_in.data_ptr(), len(z), d_out.data_ptr(), len(ref)); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        tiers = {k: c.kernel_time(k)[1] for k in ("k_piece_sym", "k_fpiece_sym", "k_fblk_map", "k_inflate")}
        ms = {k: round(c.kernel_time(k)[0], 3) for k in ("k_hdr_filter", "k_hdr_verify", "k_fblk_head", "k_fblk_map", "k_fblk_chain", "k_fblk_prefix", "k_fpiece_sym", "k_frun_merge", "k_win_propagate", "k_sym_finalize", "k_inflate")}
        c.set_timing(False)
    ok = m == len(ref) and d_out.cpu().numpy().tobytes() == ref
    print(json.dumps({"stream": name, "comp": len(z), "raw": len(ref), "ok": ok, "GBps": round(len(ref) / best / 1e9, 3), "tier_launches": tiers, "kernel_ms": ms}), flush=True)
