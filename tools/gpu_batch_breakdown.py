"""Kernel-time breakdown of the batch calls on BASELINE configs[2] (262,144 x 4 KiB buffers of the mixed corpus)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
c = zles.Codec(0)
count = 262144; n = count * 4096
src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
in_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * 4096
bound = c.deflate_bound(4096)
out_off = torch.arange(0, count + 1, dtype=torch.int64, device="cuda") * bound
out = torch.empty(count * bound, dtype=torch.uint8, device="cuda")
out_len = torch.zeros(count, dtype=torch.int64, device="cuda"); status = torch.zeros(count, dtype=torch.int32, device="cuda")
back = torch.zeros(n, dtype=torch.uint8, device="cuda"); blen = torch.zeros(count, dtype=torch.int64, device="cuda"); st2 = torch.zeros(count, dtype=torch.int32, device="cuda")
for it in range(2):
    if it == 1: c.set_timing(True)
    c.dev_deflate_batch(src.data_ptr(), in_off.data_ptr(), count, out.data_ptr(), out_off.data_ptr(), out_len.data_ptr(), status.data_ptr())
    c.dev_inflate_batch(out.data_ptr(), out_off.data_ptr(), count, back.data_ptr(), in_off.data_ptr(), blen.data_ptr(), st2.data_ptr())
torch.cuda.synchronize()
res = {}
for k in ["k_batch_count", "k_batch_table", "k_lz_batch", "k_lz", "k_huff", "k_pack_batch", "k_inflate_batch", "k_layout"]:
    ms, cnt = c.kernel_time(k)
    if cnt: res[k] = round(ms, 3)
print(json.dumps(res), "ok" if torch.equal(src, back) else "MISMATCH")
