"""Standalone Adler-32 (K8) against the HBM roofline: device time of k_adler_partial for inputs from 64 MiB (L2-warm) to
8 GiB, checked against system zlib on a sample.  usage: python tools/gpu_adler_bw.py"""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
c = zles.Codec(0)
peak = 6539.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
src = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
c.dev_corpus(3, 0, src.data_ptr(), src.numel())
h = bytes(src[:(64 << 20) + 3].cpu().numpy())
assert c.dev_adler32(src.data_ptr(), 64 << 20) == zlib.adler32(h[:64 << 20])
assert c.dev_adler32(src.data_ptr() + 3, 64 << 20) == zlib.adler32(h[3:])
for mib in (64, 256, 1024, 4096, 8192):
    n = mib << 20
    for _ in range(2):
        c.dev_adler32(src.data_ptr(), n)
    c.set_timing(True)
    for _ in range(5):
        c.dev_adler32(src.data_ptr(), n)
    ms, cnt = c.kernel_time("k_adler_partial")
    c.set_timing(False)
    gbs = n / (ms / cnt * 1e-3) / 1e9
    print(json.dumps({"MiB": mib, "k_adler_partial_ms": round(ms / cnt, 4), "GB/s": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / peak, 3)}), flush=True)
