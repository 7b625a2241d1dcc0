"""zles_deflate / zles_inflate on PAGEABLE host buffers (what the N-API addon passes) against pinned ones.
usage: python tools/gpu_pageable_probe.py [MiB = 1024]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, zles
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
c = zles.Codec(0)
st = torch.cuda.Stream(); c.set_stream(st.cuda_stream)
with torch.cuda.stream(st):
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(src); st.synchronize(); del src
cap = c.deflate_bound(n)
bufs = {"pinned": (h_in.numpy(), torch.empty(cap, dtype=torch.uint8).pin_memory().numpy(), torch.empty(n, dtype=torch.uint8).pin_memory().numpy()),
        "pageable": (np.array(h_in.numpy(), copy=True), np.empty(cap, dtype=np.uint8), np.empty(n, dtype=np.uint8))}
for name, (a, z, b) in bufs.items():
    for it in range(3):
        t0 = time.perf_counter(); m = c.deflate_into(a, z); t1 = time.perf_counter(); o = c.inflate_into(z[:m], b); t2 = time.perf_counter()
    assert o == n and bool((a == b).all())
    print(json.dumps({"buffers": name, "MiB": n >> 20, "deflate_GBps": round(n / (t1 - t0) / 1e9, 2), "inflate_GBps": round(n / (t2 - t1) / 1e9, 2),
                      "round_trip_GBps": round(n / (t2 - t0) / 1e9, 2)}), flush=True)
