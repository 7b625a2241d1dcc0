#!/usr/bin/env python
"""bench.py — deflate + inflate throughput of the zlib.es hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path over one batch of synthetic input: `deflate` of a
64 MiB English-like Markov text stream per GPU (BASELINE.json configs[1]) followed by
`inflate` of the result, both through the C ABI (include/zles.h).  Inputs are generated in
HBM before the timed region; `value` is uncompressed bytes / (deflate time + inflate time),
summed over all GPUs (weak scaling: every rank adds one 64 MiB shard; for N > 1 the shards
form ONE zlib stream — an NCCL all-gather of the shard sizes, then NVLink peer stores of the
compressed blocks into the stream on rank 0).  L2 is flushed before every timed call.

`--impl reference` times the CPU oracle (oracle/zlibes_oracle.c: a C restatement of zlib.es —
the reference itself is TypeScript and no JS engine exists in the image) on all host cores
over the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "deflate+inflate round-trip throughput (uncompressed bytes)"
UNIT = "GB/s"
SHARD = 64 << 20            # BASELINE.json configs[1]: single 64 MiB synthetic English-like Markov text stream
KIND_TEXT = 0
CHUNK = 131072


def measured_peaks():
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary (or None)."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        return j["k_lz"]["dram_bytes_per_launch"]
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host cores, one worker per core over independent 128 KiB-aligned pieces
# ---------------------------------------------------------------------------------------------
def cpu_round_trip(data: bytes, cores: int) -> dict:
    import oracle as O
    O.lib()
    n = len(data)
    nchunks = (n + CHUNK - 1) // CHUNK
    pieces = []
    for w in range(cores):
        a = min(n, (nchunks * w // cores) * CHUNK)
        b = n if w == cores - 1 else min(n, (nchunks * (w + 1) // cores) * CHUNK)
        if b > a:
            pieces.append(data[a:b])
    with ThreadPoolExecutor(max_workers=cores) as ex:  # ctypes releases the GIL inside the C calls
        t0 = time.perf_counter()
        comp = list(ex.map(O.deflate, pieces))
        t1 = time.perf_counter()
        back = list(ex.map(O.inflate, comp))
        t2 = time.perf_counter()
    assert b"".join(back) == data
    return {"deflate_s": t1 - t0, "inflate_s": t2 - t1, "comp_bytes": sum(len(c) for c in comp), "pieces": len(pieces)}


def cpu_single_thread(data: bytes) -> dict:
    """The same oracle on ONE core, and system zlib -6 on one core for context, on the first 4 MiB of the stream (the
    north_star asks for the CPU path both single-threaded and with one worker per core)."""
    import zlib as syszlib
    import oracle as O
    O.lib()
    sample = data[:4 << 20]
    t0 = time.perf_counter(); z = O.deflate(sample); t1 = time.perf_counter(); back = O.inflate(z); t2 = time.perf_counter()
    assert back == sample
    t3 = time.perf_counter(); z6 = syszlib.compress(sample, 6); t4 = time.perf_counter(); syszlib.decompress(z6); t5 = time.perf_counter()
    m = len(sample)
    return {"value": round(m / (t2 - t0) / 1e9, 5), "unit": UNIT, "cores": 1, "sample": "the first 4 MiB of the stream, once",
            "deflate_gbs": round(m / (t1 - t0) / 1e9, 5), "inflate_gbs": round(m / (t2 - t1) / 1e9, 5),
            "system_zlib_level6": {"deflate_gbs": round(m / (t4 - t3) / 1e9, 5), "inflate_gbs": round(m / (t5 - t4) / 1e9, 5),
                                   "ratio": round(m / len(z6), 4)}}


def host_text(n: int) -> bytes:
    import ctypes
    import numpy as np
    from zles import _capi
    out = np.empty(n, dtype=np.uint8)
    rc = _capi.lib().zles_host_corpus(KIND_TEXT, 0, out.ctypes.data, n)
    assert rc == 0
    return out.tobytes()


def node_reference(data: bytes):
    """The unmodified zlib.es under Node (bench/node_ref.mjs), when this machine has `node` and ZLIBES_DIST names the
    reference's dist/cjs/zlib.js.  Neither the build image nor the GPU boxes do: returns None there."""
    import shutil, subprocess, tempfile
    node, dist_js = shutil.which("node"), os.environ.get("ZLIBES_DIST")
    if not node or not dist_js or not os.path.exists(dist_js):
        return None
    try:
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(data); f.flush()
            out = subprocess.run([node, os.path.join(ROOT, "bench", "node_ref.mjs"), dist_js, f.name], capture_output=True, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    data = host_text(SHARD)
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_round_trip(data, cores)
        if i >= args.warmup:
            times.append(r)
    td = sum(t["deflate_s"] for t in times) / len(times)
    ti = sum(t["inflate_s"] for t in times) / len(times)
    value = SHARD / (td + ti) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round((td + ti) * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "64 MiB synthetic English-like order-2 Markov text, deflate then inflate; CPU oracle "
                               "(C restatement of zlib.es) on all host cores, one worker per core over independent 128 KiB-aligned pieces"},
        "deflate_gbs": round(SHARD / td / 1e9, 5), "inflate_gbs": round(SHARD / ti / 1e9, 5),
        "ratio": round(SHARD / times[-1]["comp_bytes"], 4),
        "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "the full 64 MiB stream per step, %d pieces" % times[-1]["pieces"],
                         "single_thread": cpu_single_thread(data)},
        "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    node = node_reference(data)
    if node is not None:
        line["node_zlib_es"] = node  # the reference itself, when a Node runtime exists on the box
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import zles
    from zles import dist as zdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the codec has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = zles.Codec(local)
    stream = torch.cuda.Stream()
    c.set_stream(stream.cuda_stream)
    n = SHARD
    dev = torch.device("cuda", local)
    with torch.cuda.stream(stream):
        src = torch.empty(n, dtype=torch.uint8, device=dev)
        c.dev_corpus(KIND_TEXT, rank * n, src.data_ptr(), n)  # rank r holds bytes [r*64Mi, (r+1)*64Mi) of the corpus
        cap = c.deflate_bound(n)
        comp = torch.empty(cap, dtype=torch.uint8, device=dev)
        back = torch.empty(n, dtype=torch.uint8, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    sc = None
    if multi:
        sc = zdist.ShardedCodec(c, zdist.IpcTransport(c), rank, world)
        sc.setup(c.deflate_bound(n) * world + 64)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def l2_flush():
        with torch.cuda.stream(stream):
            flush.fill_(rank + 1)

    def step(timed: bool):
        """Returns (deflate_ms, inflate_ms, compressed bytes of this rank's shard)."""
        l2_flush()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        if multi:
            dist.barrier()
        e0.record(stream)
        if multi:
            lay = sc.deflate(src.data_ptr(), n)
            clen = lay.comp[rank]
        else:
            clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
        e1.record(stream)
        l2_flush()
        if multi:
            dist.barrier()
        e2.record(stream)
        if multi:
            olen = sc.inflate(comp.data_ptr(), back.data_ptr(), n)
        else:
            olen = c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n)
        e3.record(stream)
        stream.synchronize()
        assert olen == n
        return e0.elapsed_time(e1), e2.elapsed_time(e3), clen

    for _ in range(args.warmup):
        step(False)
    with torch.cuda.stream(stream):
        assert torch.equal(src, back), "round trip mismatch"

    # ---- timed region: exactly K steps --------------------------------------------------------
    c.set_timing(True)
    launches0 = c.launches
    sampler = ClockSampler(local) if rank == 0 else None
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    td = ti = 0.0
    clen = 0
    for _ in range(args.steps):
        a, b, clen = step(True)
        td += a
        ti += b
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = c.launches - launches0
    lz_ms, lz_n = c.kernel_time("k_lz")
    tok_ms, tok_n = c.kernel_time("k_inf_tokens")  # phase A: one of the two forms runs (zles.cu inflate_decode)
    tok4_ms, tok4_n = c.kernel_time("k_inf_tokens4")
    tok_ms += tok4_ms
    tok_n += tok4_n
    res_ms, res_n = c.kernel_time("k_inf_resolve")  # phase B: one of the two forms runs (zles.cu launch_phase_b)
    sym_ms, _ = c.kernel_time("k_piece_sym")
    fin_ms, _ = c.kernel_time("k_chunk_final")
    res_ms += sym_ms + fin_ms
    pack_ms, pack_n = c.kernel_time("k_pack")
    c.set_timing(False)

    # ---- end to end through the drop-in C ABI with HOST buffers (pinned), copies inside the timing ------------
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.copy_(src)
    h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    e2e_t = 0.0
    e2e_steps = max(1, min(args.steps, 5))
    hc = 0
    for i in range(1 + e2e_steps):
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        t0 = time.perf_counter()
        hc = c.deflate_into(h_in.numpy(), h_comp.numpy())            # H2D 64 MiB, kernels, D2H compressed
        ho = c.inflate_into(h_comp.numpy()[:hc], h_back.numpy())     # H2D compressed, kernels, D2H 64 MiB
        t1 = time.perf_counter()
        assert ho == n
        if i > 0:
            e2e_t += t1 - t0
    assert bool((h_back.numpy() == h_in.numpy()).all())
    e2e_ms = e2e_t / e2e_steps * 1e3

    # ---- max over ranks --------------------------------------------------------------------------
    vals = torch.tensor([td / args.steps, ti / args.steps, e2e_ms, float(clen), float(launches), lz_ms / max(1, lz_n)], dtype=torch.float64,
                        device=dev)
    if multi:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    d_ms, i_ms, e_ms = float(mx[0]), float(mx[1]), float(mx[2])
    comp_total = float(sm[3])
    launches_total = int(float(sm[4]))
    lz_avg_ms = float(mx[5])

    cpu = None
    if rank == 0 and not multi:
        cores = os.cpu_count() or 1
        r = cpu_round_trip(src.cpu().numpy().tobytes(), cores)
        cpu = {"value": round(n / (r["deflate_s"] + r["inflate_s"]) / 1e9, 5), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "the full 64 MiB stream once, %d independent pieces (one per core)" % r["pieces"],
               "deflate_gbs": round(n / r["deflate_s"] / 1e9, 5), "inflate_gbs": round(n / r["inflate_s"] / 1e9, 5),
               "ratio": round(n / r["comp_bytes"], 4), "single_thread": cpu_single_thread(src[: 4 << 20].cpu().numpy().tobytes())}

    if rank == 0:
        total = n * world
        peak, peak_src = measured_peaks()
        algo = n + comp_total / world  # deflate = U read + C written, per launch of the matcher over one shard
        achieved = algo / (lz_avg_ms * 1e-3) / 1e9 if lz_avg_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(total / ((d_ms + i_ms) * 1e-3) / 1e9, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(d_ms + i_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "64 MiB synthetic English-like order-2 Markov text per GPU (BASELINE.json configs[1]); "
                                   "deflate then inflate through the C ABI, device-resident; for N > 1 the shards form one zlib stream "
                                   "(NCCL all-gather of shard sizes + NVLink peer stores into rank 0)",
                       "bytes_per_gpu": n, "l2": "flushed (256 MiB fill) before every timed call", "timing": "CUDA events on the codec's stream, max over ranks"},
            "deflate_gbs": round(total / (d_ms * 1e-3) / 1e9, 4), "inflate_gbs": round(total / (i_ms * 1e-3) / 1e9, 4),
            "ratio": round(total / comp_total, 4), "wall_s_timed_region": round(wall, 3),
            "e2e": {"value": round(total / (e_ms * 1e-3) / 1e9, 4), "unit": UNIT, "h2d_bytes_per_step": int(total + comp_total),
                    "d2h_bytes_per_step": int(total + comp_total),
                    "how": "zles_deflate + zles_inflate on pinned host buffers, one independent 64 MiB stream per GPU"},
            "gpu_launches": launches_total,
            "kernels_ms_per_step": {"k_lz": round(lz_ms / args.steps, 4), "k_pack": round(pack_ms / args.steps, 4),
                                    "phase_a(k_inf_tokens|k_inf_tokens4)": round(tok_ms / args.steps, 4), "phase_b(k_inf_resolve|k_piece_sym+k_chunk_final)": round(res_ms / args.steps, 4),
                                    "k_piece_sym": round(sym_ms / args.steps, 4), "k_chunk_final": round(fin_ms / args.steps, 4)},
            "roofline": {"bound": "hbm", "kernel": "k_lz", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo)},
            "roofline_inflate": {"bound": "hbm", "kernels": "k_inf_tokens + phase B",
                                 "achieved": round(algo / ((tok_ms + res_ms) / max(1, tok_n) * 1e-3) / 1e9, 2) if tok_ms + res_ms > 0 else 0.0,
                                 "peak": peak, "unit": "GB/s",
                                 "frac": round(algo / ((tok_ms + res_ms) / max(1, tok_n) * 1e-3) / 1e9 / peak, 5) if tok_ms + res_ms > 0 else 0.0,
                                 "algorithmic_bytes_per_launch": int(algo)},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if multi:
        sc.teardown()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
