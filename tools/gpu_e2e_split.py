"""Host-buffer (pinned) deflate and inflate timed separately, against the device-resident calls (one GPU, 64 MiB text)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
c = zles.Codec(0)
n = 64 << 20
src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(0, 0, src.data_ptr(), n)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(src.cpu())
cap = c.deflate_bound(n)
h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory(); h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
best = {"deflate_host": 1e9, "inflate_host": 1e9, "deflate_dev": 1e9, "inflate_dev": 1e9}
for _ in range(6):
    torch.cuda.synchronize(); t = time.perf_counter(); clen = c.deflate_into(h_in.numpy(), h_comp.numpy()); best["deflate_host"] = min(best["deflate_host"], time.perf_counter() - t)
    t = time.perf_counter(); m = c.inflate_into(h_comp.numpy()[:clen], h_back.numpy()); best["inflate_host"] = min(best["inflate_host"], time.perf_counter() - t)
    torch.cuda.synchronize(); t = time.perf_counter(); cl2 = c.dev_deflate(src.data_ptr(), n, d_comp.data_ptr(), cap); torch.cuda.synchronize(); best["deflate_dev"] = min(best["deflate_dev"], time.perf_counter() - t)
    t = time.perf_counter(); c.dev_inflate(d_comp.data_ptr(), cl2, d_back.data_ptr(), n); torch.cuda.synchronize(); best["inflate_dev"] = min(best["inflate_dev"], time.perf_counter() - t)
assert m == n and torch.equal(h_back, h_in)
print(json.dumps({k: round(v * 1e3, 3) for k, v in best.items()} | {"comp_bytes": clen, "unit": "ms, best of 6, wall clock around the call"}))
