"""Where the time of the host-buffer calls goes: zles_deflate / zles_inflate on pinned buffers against the device-resident
forms, with per-kernel device time.  usage: python tools/gpu_e2e_probe.py [MiB = 1024] [slab_blocks ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
slabs = [int(x) for x in sys.argv[2:]] or [0]
c = zles.Codec(0)
st = torch.cuda.Stream(); c.set_stream(st.cuda_stream)
KS = ("k_lz", "k_huff", "k_pack", "k_layout_slab", "k_inf_tokens", "k_inf_tokens4", "k_inf_resolve", "k_piece_sym", "k_chunk_final", "k_mark_count", "k_mark_emit", "k_inf_check")
with torch.cuda.stream(st):
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(src); st.synchronize()
    cap = c.deflate_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); back = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(st); clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap); e1.record(st)
        olen = c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n); e2.record(st); st.synchronize()
    print(json.dumps({"device_resident_ms": {"deflate": round(e0.elapsed_time(e1), 2), "inflate": round(e1.elapsed_time(e2), 2)}}), flush=True)
    del src, comp, back
torch.cuda.empty_cache()
h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory(); h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
for slab in slabs:
    c.set_slab_blocks(slab)
    for it in range(2):
        c.set_timing(it == 1)
        t0 = time.perf_counter(); hc = c.deflate_into(h_in.numpy(), h_comp.numpy()); t1 = time.perf_counter()
        kd = {k: round(c.kernel_time(k)[0], 2) for k in KS if c.kernel_time(k)[1]} if it == 1 else None
        c.set_timing(it == 1)
        t2 = time.perf_counter(); ho = c.inflate_into(h_comp.numpy()[:hc], h_back.numpy()); t3 = time.perf_counter()
        ki = {k: (round(c.kernel_time(k)[0], 2), c.kernel_time(k)[1]) for k in KS if c.kernel_time(k)[1]} if it == 1 else None
    c.set_timing(False)
    assert ho == n and torch.equal(h_back, h_in)
    print(json.dumps({"slab_blocks": slab, "host_ms": {"deflate": round((t1 - t0) * 1e3, 2), "inflate": round((t3 - t2) * 1e3, 2)},
                      "GBps": {"deflate": round(n / (t1 - t0) / 1e9, 2), "inflate": round(n / (t3 - t2) / 1e9, 2)}, "deflate_kernels_ms": kd, "inflate_kernels_ms(n)": ki}), flush=True)
