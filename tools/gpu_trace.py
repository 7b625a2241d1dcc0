import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zles
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
c = zles.Codec(0)
st = torch.cuda.Stream(); c.set_stream(st.cuda_stream)
with torch.cuda.stream(st):
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(3, 0, src.data_ptr(), n)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(src); st.synchronize(); del src
cap = c.deflate_bound(n)
h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory(); h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
for it in range(3):
    print("---- iteration", it, file=sys.stderr, flush=True)
    t0 = time.perf_counter(); hc = c.deflate_into(h_in.numpy(), h_comp.numpy()); t1 = time.perf_counter()
    ho = c.inflate_into(h_comp.numpy()[:hc], h_back.numpy()); t2 = time.perf_counter()
    print("deflate %.2f ms  inflate %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr, flush=True)
