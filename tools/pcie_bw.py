"""Host<->device copy bandwidth of this box (pinned and pageable), per GPU and all GPUs at once: the bound of every
host-buffer (`e2e`) figure in bench.py."""
import json, sys, time, threading
import torch
n = 1 << 30
res = {}
ng = torch.cuda.device_count()
def one(dev, pinned, out):
    torch.cuda.set_device(dev)
    h = torch.empty(n, dtype=torch.uint8)
    if pinned: h = h.pin_memory()
    h.fill_(3)
    d = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % dev)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3): fn()
        torch.cuda.synchronize(dev)
        out[name] = round(3 * n / (time.perf_counter() - t0) / 1e9, 2)
for pinned in (True, False):
    o = {}; one(0, pinned, o); res["gpu0_%s" % ("pinned" if pinned else "pageable")] = o
if ng > 1:
    outs = [dict() for _ in range(ng)]
    th = [threading.Thread(target=one, args=(g, True, outs[g])) for g in range(ng)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    res["all_gpus_concurrently_pinned"] = outs
print(json.dumps(res))
