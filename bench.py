#!/usr/bin/env python
"""bench.py — deflate + inflate throughput of the zlib.es hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[4], the configuration the north_star metric is quoted on): the 8 GiB synthetic
mixed corpus (per 128 KiB segment 45 % English-like text, 45 % structured binary, 10 % random) compressed into ONE
zlib stream and inflated again.  The stream's 65,536 independent 128 KiB chunks are sharded contiguously over the
N GPUs — strong scaling: the total stays 8 GiB for N = 1, 2, 4, 8.  A step is one deflate of the whole corpus plus
one inflate of the resulting stream, both through the C ABI (include/zles.h), inputs generated in HBM before the
timed region.  For N > 1 (one process per GPU under torchrun) the ranks exchange their shards' compressed sizes with
one NCCL all-gather and the bit packer stores every shard's blocks straight into the stream on rank 0 over NVLink
(peer-mapped memory); inflate finds the shards from the stream itself (marker scan + all-gather), every rank
peer-reads and decodes its share.  `value` = 8 GiB / (deflate time + inflate time), CUDA events, max over ranks.
Every shard (>= 1 GiB) is far larger than the 126 MB L2, so no L2 flush is needed between steps.

`e2e` is the same round trip through the drop-in host-buffer calls (zles_deflate / zles_inflate) on pinned host
buffers, host<->device copies inside the timed region; for N > 1 it runs in ONE process (rank 0) that drives all N
GPUs through the library's own multi-GPU context (zles_mgpu_*, what the N-API addon binds after zles_init(mask)).

`--impl reference` times the CPU oracle (oracle/zlibes_oracle.c: a C restatement of zlib.es — the reference itself is
TypeScript and no JS engine exists in the image) on all host cores over a bounded sample of the same workload; rank 0
only.  Its input comes from bench/corpus_np.py (a numpy port of the corpus generator): the CPU arm never loads the
product library.
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
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "bench")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "deflate+inflate round-trip throughput (uncompressed bytes)"
UNIT = "GB/s"
CHUNK = 131072
KIND_TEXT, KIND_MIXED = 0, 3
TOTAL = int(os.environ.get("ZLES_BENCH_TOTAL_MIB", str(8 << 10))) << 20     # 8 GiB unless overridden (smoke runs)
CPU_SAMPLE = min(TOTAL, int(os.environ.get("ZLES_BENCH_CPU_SAMPLE_MIB", "128")) << 20)


def workload_config(n_gpus: int) -> dict:
    """Identical in both arms (the driver compares them)."""
    return {
        "workload": "BASELINE.json configs[4]: %g GiB synthetic mixed corpus (per 128 KiB segment 45%% order-2 Markov text / 45%% "
                    "structured binary / 10%% random) as ONE zlib stream, deflate then inflate, 128 KiB chunks sharded "
                    "contiguously over the GPUs" % (TOTAL / (1 << 30)),
        "total_bytes": TOTAL, "chunk_bytes": CHUNK, "shards": n_gpus, "corpus_kind": "mixed",
        "l2": "every shard (>= 1 GiB) exceeds the 126 MB L2; no flush needed",
    }


def measured_peaks():
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str, launch_input_bytes: float):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, scaled to this launch's input size
    (the capture's own size and git hash are reported next to it).  None when there is no capture."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        k = j[kernel]
        per_byte = k["dram_bytes_per_launch"] / k["input_bytes_per_launch"]
        return {"traffic": int(per_byte * launch_input_bytes), "dram_bytes_per_input_byte": round(per_byte, 4),
                "profile_git": j.get("_meta", {}).get("git"), "profile_input_bytes": k["input_bytes_per_launch"]}
    except Exception:
        return None


def issue_roofline(kernel: str, launch_input_bytes: float, launch_ms: float, sm_mhz: float, sms: int):
    """The bound that actually holds for the byte-granular integer kernels (DRAM is 1 % busy): warp instructions issued per
    second against the SMs' issue rate (4 schedulers per SM, one warp instruction per cycle each).  The instruction count
    per input byte comes from the committed ncu capture (smsp__inst_executed.sum), the time is this run's."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        k = j[kernel]
        inst = k["warp_inst"] / k["input_bytes_per_launch"] * launch_input_bytes
        peak = sms * 4 * sm_mhz * 1e6
        ach = inst / (launch_ms * 1e-3)
        return {"bound": "issue", "kernel": kernel, "achieved": round(ach / 1e9, 1), "peak": round(peak / 1e9, 1), "unit": "G warp-instructions/s",
                "frac": round(ach / peak, 4), "warp_instructions_per_input_byte": round(k["warp_inst"] / k["input_bytes_per_launch"], 2),
                "peak_source": "%d SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the timed region)" % (sms, sm_mhz),
                "profile_git": j.get("_meta", {}).get("git")}
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
    """The same oracle on ONE core, and system zlib -6 on one core for context, on the first 4 MiB of the sample (the
    north_star asks for the CPU path both single-threaded and with one worker per core)."""
    import zlib as syszlib
    import oracle as O
    O.lib()
    sample = data[:4 << 20]
    t0 = time.perf_counter(); z = O.deflate(sample); t1 = time.perf_counter(); back = O.inflate(z); t2 = time.perf_counter()
    assert back == sample
    t3 = time.perf_counter(); z6 = syszlib.compress(sample, 6); t4 = time.perf_counter(); syszlib.decompress(z6); t5 = time.perf_counter()
    m = len(sample)
    return {"value": round(m / (t2 - t0) / 1e9, 5), "unit": UNIT, "cores": 1, "sample": "the first 4 MiB of the corpus, once",
            "deflate_gbs": round(m / (t1 - t0) / 1e9, 5), "inflate_gbs": round(m / (t2 - t1) / 1e9, 5),
            "system_zlib_level6": {"deflate_gbs": round(m / (t4 - t3) / 1e9, 5), "inflate_gbs": round(m / (t5 - t4) / 1e9, 5),
                                   "ratio": round(m / len(z6), 4)}}


CPU_SAMPLE_DESC = "the first %d MiB of the corpus (1,024 of its 65,536 chunks) per step, one worker per core over independent 128 KiB-aligned pieces"


def node_reference(data: bytes):
    """The unmodified zlib.es under Node (bench/node_ref.mjs), when this machine has `node` and ZLIBES_DIST names the
    reference's dist/cjs/zlib.js.  Neither the build image nor the GPU boxes do: returns None there."""
    import shutil
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
    import corpus_np  # numpy port of the generator: this arm never loads libzles.so
    cores = os.cpu_count() or 1
    data = corpus_np.corpus(KIND_MIXED, 0, CPU_SAMPLE).tobytes()
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_round_trip(data, cores)
        if i >= args.warmup:
            times.append(r)
    td = sum(t["deflate_s"] for t in times) / len(times)
    ti = sum(t["inflate_s"] for t in times) / len(times)
    value = CPU_SAMPLE / (td + ti) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round((td + ti) * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args.gpus),
        "deflate_gbs": round(CPU_SAMPLE / td / 1e9, 5), "inflate_gbs": round(CPU_SAMPLE / ti / 1e9, 5),
        "ratio": round(CPU_SAMPLE / times[-1]["comp_bytes"], 4),
        "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "port",
                         "what": "oracle/zlibes_oracle.c, a C restatement of zlib.es (an upper bound on what V8 achieves)",
                         "sample": CPU_SAMPLE_DESC % (CPU_SAMPLE >> 20), "single_thread": cpu_single_thread(data)},
        "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    node = node_reference(data[:16 << 20])
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
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        # "under load": samples taken while the GPU drew more than half of the run's peak power
        pmax = max(power) if power else 0.0
        busy = [v for v, p in zip(sm, power) if p >= 0.5 * pmax and v > 0] or [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def numa_nodes() -> int:
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        return 1


def pinned_host(torch, nbytes: int, interleave: bool):
    """A page-locked host buffer of nbytes (uint8 tensor).  interleave: its pages are spread over all NUMA nodes of the box
    (set_mempolicy(MPOL_INTERLEAVE) around the first touch, then cudaHostRegister) — one process feeding GPUs on both
    sockets from a buffer that sits on one node is limited by the socket interconnect, not by PCIe."""
    import ctypes
    import numpy as np
    if interleave and numa_nodes() > 1:
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            nn = numa_nodes()
            mask = ctypes.c_ulong((1 << nn) - 1)
            if libc.syscall(238, 3, ctypes.byref(mask), ctypes.c_ulong(64)) == 0:   # set_mempolicy(MPOL_INTERLEAVE, all nodes)
                try:
                    arr = np.empty(nbytes, dtype=np.uint8)
                    arr[::4096] = 0                                                  # first touch under the policy
                finally:
                    libc.syscall(238, 0, None, ctypes.c_ulong(0))                    # MPOL_DEFAULT
                t = torch.from_numpy(arr)
                if int(torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)) == 0:
                    return t, "interleaved over %d NUMA nodes, cudaHostRegister" % nn
        except Exception:
            pass
    return torch.empty(nbytes, dtype=torch.uint8).pin_memory(), "cudaHostAlloc (torch pin_memory)"


def other_configs(c, torch, stream, dev):
    """BASELINE.json configs[1..3] on one GPU (parity-test cases; reported for context, best of 3, device-resident) and the
    parallel inflate of streams made by other encoders."""
    import numpy as np
    out = {}

    def timed(fn, reps=3):
        best, r = 1e30, None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); r = fn(); e1.record(stream); stream.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return r, best

    with torch.cuda.stream(stream):
        for key, kind, n in (("configs[1] 64 MiB text stream", KIND_TEXT, 64 << 20), ("configs[3] 1 GiB mixed stream", KIND_MIXED, 1 << 30)):
            src = torch.empty(n, dtype=torch.uint8, device=dev)
            c.dev_corpus(kind, 0, src.data_ptr(), n)
            cap = c.deflate_bound(n)
            comp = torch.empty(cap, dtype=torch.uint8, device=dev)
            back = torch.empty(n, dtype=torch.uint8, device=dev)
            clen, td = timed(lambda: c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap))
            olen, ti = timed(lambda: c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n))
            ok = bool(olen == n and torch.equal(src, back))
            out[key] = {"deflate_gbs": round(n / td / 1e6, 2), "inflate_gbs": round(n / ti / 1e6, 2), "ratio": round(n / clen, 4), "roundtrip_ok": ok}
            del src, comp, back
        # configs[2]: 262,144 x 4 KiB buffers of the mixed corpus, each its own zlib stream
        count = 262144
        n = count * 4096
        src = torch.empty(n, dtype=torch.uint8, device=dev)
        c.dev_corpus(KIND_MIXED, 0, src.data_ptr(), n)
        in_off = torch.arange(0, count + 1, dtype=torch.int64, device=dev) * 4096
        bound = c.deflate_bound(4096)
        out_off = torch.arange(0, count + 1, dtype=torch.int64, device=dev) * bound
        outb = torch.empty(count * bound, dtype=torch.uint8, device=dev)
        out_len = torch.zeros(count, dtype=torch.int64, device=dev)
        status = torch.zeros(count, dtype=torch.int32, device=dev)
        stream.synchronize()
        rc, td = timed(lambda: c.dev_deflate_batch(src.data_ptr(), in_off.data_ptr(), count, outb.data_ptr(), out_off.data_ptr(), out_len.data_ptr(), status.data_ptr()))
        back = torch.zeros(n, dtype=torch.uint8, device=dev)
        blen = torch.zeros(count, dtype=torch.int64, device=dev)
        st2 = torch.zeros(count, dtype=torch.int32, device=dev)
        stream.synchronize()
        rc2, ti = timed(lambda: c.dev_inflate_batch(outb.data_ptr(), out_off.data_ptr(), count, back.data_ptr(), in_off.data_ptr(), blen.data_ptr(), st2.data_ptr()))
        stream.synchronize()
        out["configs[2] 262,144 x 4 KiB buffers"] = {"deflate_gbs": round(n / td / 1e6, 2), "inflate_gbs": round(n / ti / 1e6, 2),
                                                     "ratio": round(n / int(out_len.sum()), 4), "roundtrip_ok": bool(rc == 0 and rc2 == 0 and torch.equal(src, back))}
        del src, outb, back
    # streams of other encoders (host-made, decoded from HBM): zlib.es's own bit-concatenated 128 KiB blocks, system zlib
    try:
        import zlib as syszlib
        import oracle as O
        foreign = {}
        text = c.host_corpus(KIND_MIXED, 0, 32 << 20).tobytes()
        for key, make, m in (("zlib.es-made 16 MiB", lambda d: O.deflate(d), 16 << 20), ("system zlib -6 32 MiB", lambda d: syszlib.compress(d, 6), 32 << 20),
                             ("system zlib -1 32 MiB", lambda d: syszlib.compress(d, 1), 32 << 20)):
            z = make(text[:m])
            with torch.cuda.stream(stream):
                d_z = torch.frombuffer(bytearray(z), dtype=torch.uint8).to(dev)
                d_o = torch.empty(m, dtype=torch.uint8, device=dev)
                olen, ti = timed(lambda: c.dev_inflate(d_z.data_ptr(), len(z), d_o.data_ptr(), m))
                ok = olen == m and d_o.cpu().numpy().tobytes() == text[:m]
            foreign[key] = {"inflate_gbs": round(m / ti / 1e6, 2), "ok": bool(ok)}
        out["foreign-stream inflate (parallel tier)"] = foreign
    except Exception as e:  # pragma: no cover
        out["foreign-stream inflate (parallel tier)"] = {"error": str(e)}
    return out


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
    host_group = None
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")  # waits that must not keep a GPU busy (an NCCL barrier spins on the device)
    c = zles.Codec(local)
    stream = torch.cuda.Stream()
    c.set_stream(stream.cuda_stream)
    dev = torch.device("cuda", local)
    a, b = zdist.shard_bounds(TOTAL, world)[rank]
    n = b - a                                   # this rank's shard: chunks [a / CHUNK, b / CHUNK) of the corpus
    with torch.cuda.stream(stream):
        src = torch.empty(n, dtype=torch.uint8, device=dev)
        c.dev_corpus(KIND_MIXED, a, src.data_ptr(), n)
        back = torch.empty(n, dtype=torch.uint8, device=dev)
        cap = c.deflate_bound(TOTAL if not multi else n) + 64
        comp = torch.empty(cap, dtype=torch.uint8, device=dev)  # N = 1: the stream; N > 1: this rank's share of it (the stream is on rank 0)
        slice_ = torch.empty(cap, dtype=torch.uint8, device=dev) if multi else None  # N > 1: the slice of the stream this rank scans
    sc = None
    if multi:
        sc = zdist.ShardedCodec(c, zdist.IpcTransport(c), rank, world)
        sc.setup(c.deflate_bound(TOTAL) + 64 * world)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step():
        """Returns (deflate_ms, inflate_ms, compressed bytes of the whole stream)."""
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        if multi:
            dist.barrier()
        e0.record(stream)
        if multi:
            lay = sc.deflate(src.data_ptr(), n)
            clen = lay.total_comp
        else:
            clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
        e1.record(stream)
        if multi:
            dist.barrier()
        e2.record(stream)
        if multi:
            olen = sc.inflate_from_stream(clen, slice_.data_ptr(), comp.data_ptr(), cap, back.data_ptr(), n)  # shards found from the stream itself
        else:
            olen = c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n)
        e3.record(stream)
        stream.synchronize()
        assert olen == n, (olen, n)
        return e0.elapsed_time(e1), e2.elapsed_time(e3), clen

    for _ in range(args.warmup):
        step()
    with torch.cuda.stream(stream):
        assert torch.equal(src, back), "round trip mismatch"
        back.zero_()

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
        d_ms, i_ms, clen = step()
        td += d_ms
        ti += i_ms
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = c.launches - launches0
    free_b, total_b = torch.cuda.mem_get_info(local)  # the codec's workspaces only grow: what is in use now is the step's peak
    dev_mem_gib = round((total_b - free_b) / (1 << 30), 2)
    with torch.cuda.stream(stream):
        assert torch.equal(src, back), "round trip mismatch after the timed steps"
    kt = {k: c.kernel_time(k) for k in ("k_lz", "k_huff", "k_pack", "k_inf_tokens", "k_inf_tokens4", "k_inf_resolve", "k_piece_sym", "k_chunk_final",
                                        "k_mark_count", "k_mark_emit")}
    c.set_timing(False)
    # the one pure streaming kernel of the path on its own: Adler-32 of this rank's shard (standalone K8; in deflate the sums are
    # fused into k_lz's load), against the same HBM peak
    adler = None
    try:
        c.dev_adler32(src.data_ptr(), n)
        c.set_timing(True)
        for _ in range(3):
            c.dev_adler32(src.data_ptr(), n)
        a_ms, a_n = c.kernel_time("k_adler_partial")
        c.set_timing(False)
        if a_n:
            adler = {"kernel": "k_adler_partial", "bytes": n, "launch_ms": round(a_ms / a_n, 4), "gbs": round(n / (a_ms / a_n * 1e-3) / 1e9, 1)}
    except Exception as e:  # never let the side measurement break the line
        adler = {"error": str(e)}
    lz_ms, lz_n = kt["k_lz"]
    tok_ms = kt["k_inf_tokens"][0] + kt["k_inf_tokens4"][0]
    tok_n = kt["k_inf_tokens"][1] + kt["k_inf_tokens4"][1]
    res_ms = kt["k_inf_resolve"][0] + kt["k_piece_sym"][0] + kt["k_chunk_final"][0]

    # ---- the other BASELINE configs and foreign streams (context, N = 1 only) ---------------------
    extras = None
    if not multi and not args.no_extras:
        del back
        torch.cuda.empty_cache()
        extras = other_configs(c, torch, stream, dev)

    # ---- end to end through the drop-in C ABI with HOST buffers, copies inside the timing ------------
    # N = 1: zles_deflate + zles_inflate of the 8 GiB corpus on this GPU.  N > 1: one process (rank 0) drives all N GPUs
    # through the library's multi-GPU context; the other ranks wait.  Device tensors of the resident run are released first.
    src_host = None
    host_how = None
    if rank == 0:
        src_host, host_how = pinned_host(torch, TOTAL, multi)
    del src, comp, slice_
    try:
        del back
    except NameError:
        pass
    if multi:
        sc.teardown()
        sc = None
    torch.cuda.empty_cache()
    if multi:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
    e2e = None
    pageable = None
    if rank == 0:
        slab = 1 << 30
        with torch.cuda.stream(stream):
            tmp = torch.empty(min(slab, TOTAL), dtype=torch.uint8, device=dev)
            for o in range(0, TOTAL, slab):
                m = min(slab, TOTAL - o)
                c.dev_corpus(KIND_MIXED, o, tmp.data_ptr(), m)
                src_host[o:o + m].copy_(tmp[:m], non_blocking=True)
            stream.synchronize()
            del tmp
        torch.cuda.empty_cache()
        hcap = c.deflate_bound(TOTAL)
        h_comp, _ = pinned_host(torch, hcap, multi)
        h_back, _ = pinned_host(torch, TOTAL, multi)
        codec = zles.MultiCodec(list(range(world))) if multi else c
        e2e_steps = max(1, min(args.steps, 3))
        t_d = t_i = 0.0
        hc = 0
        for i in range(1 + e2e_steps):
            t0 = time.perf_counter()
            hc = codec.deflate_into(src_host.numpy(), h_comp.numpy())            # H2D 8 GiB, kernels, D2H compressed
            t1 = time.perf_counter()
            ho = codec.inflate_into(h_comp.numpy()[:hc], h_back.numpy())         # H2D compressed, kernels, D2H 8 GiB
            t2 = time.perf_counter()
            assert ho == TOTAL
            if i > 0:
                t_d += t1 - t0
                t_i += t2 - t1
        assert torch.equal(h_back, src_host), "e2e round trip mismatch"
        e2e = {"value": round(TOTAL / ((t_d + t_i) / e2e_steps) / 1e9, 4), "unit": UNIT, "h2d_bytes_per_step": int(TOTAL + hc), "d2h_bytes_per_step": int(TOTAL + hc),
               "deflate_gbs": round(TOTAL / (t_d / e2e_steps) / 1e9, 4), "inflate_gbs": round(TOTAL / (t_i / e2e_steps) / 1e9, 4), "steps": e2e_steps,
               "host_buffers": host_how, "numa_nodes": numa_nodes(),
               "how": ("zles_deflate + zles_inflate on pinned host buffers, wall clock around the calls" if not multi else
                       "zles_mgpu_deflate + zles_mgpu_inflate (one process, %d GPUs, pinned host buffers), wall clock around the calls" % world)}
        # the same calls with PAGEABLE host buffers (what a Node ArrayBuffer is): 1 GiB of the corpus, one warm-up + one timed pass
        pn = min(TOTAL, 1 << 30)
        p_in = np.array(src_host.numpy()[:pn], copy=True)
        p_comp = np.empty(c.deflate_bound(pn), dtype=np.uint8)
        p_back = np.empty(pn, dtype=np.uint8)
        for i in range(2):
            t0 = time.perf_counter()
            pc = codec.deflate_into(p_in, p_comp)
            t1 = time.perf_counter()
            po = codec.inflate_into(p_comp[:pc], p_back)
            t2 = time.perf_counter()
        assert po == pn and bool((p_back == p_in).all())
        pageable = {"value": round(pn / (t2 - t0) / 1e9, 4), "unit": UNIT, "deflate_gbs": round(pn / (t1 - t0) / 1e9, 4), "inflate_gbs": round(pn / (t2 - t1) / 1e9, 4),
                    "bytes": pn, "how": "the same calls on pageable (malloc'ed) host buffers, first 1 GiB of the corpus"}
        if multi:
            codec.close()
    if multi:
        dist.barrier(group=host_group)  # the other ranks wait on the host: their GPUs belong to rank 0's multi-GPU context meanwhile

    # ---- max over ranks --------------------------------------------------------------------------
    vals = torch.tensor([td / args.steps, ti / args.steps, float(launches), lz_ms / max(1, args.steps), tok_ms / max(1, args.steps), res_ms / max(1, args.steps), lz_n / max(1, args.steps)],
                        dtype=torch.float64, device=dev)
    if multi:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    d_ms, i_ms = float(mx[0]), float(mx[1])
    launches_total = int(float(sm[2]))
    lz_step_ms = float(mx[3])                    # k_lz per step on the slowest rank: all its launches (one per slab of the shard)
    lz_launches = max(1, round(float(mx[6])))
    lz_avg_ms = lz_step_ms / lz_launches

    cpu = None
    if rank == 0 and not multi and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = src_host.numpy()[:CPU_SAMPLE].tobytes()
        r = cpu_round_trip(sample, cores)
        cpu = {"value": round(CPU_SAMPLE / (r["deflate_s"] + r["inflate_s"]) / 1e9, 5), "unit": UNIT, "cores": cores, "kind": "port",
               "what": "oracle/zlibes_oracle.c, a C restatement of zlib.es (an upper bound on what V8 achieves)",
               "sample": CPU_SAMPLE_DESC % (CPU_SAMPLE >> 20) + ", once",
               "deflate_gbs": round(CPU_SAMPLE / r["deflate_s"] / 1e9, 5), "inflate_gbs": round(CPU_SAMPLE / r["inflate_s"] / 1e9, 5),
               "ratio": round(CPU_SAMPLE / r["comp_bytes"], 4), "single_thread": cpu_single_thread(sample)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        shard = TOTAL / world
        algo_shard = shard + (clen - 6) / world  # deflate = U read + C written by the matcher's pass over one shard
        algo = algo_shard / lz_launches         # ... per launch: a long resident shard is matched in slabs of 1 GiB, one launch each
        achieved = algo / (lz_avg_ms * 1e-3) / 1e9 if lz_avg_ms > 0 else 0.0
        tr = ncu_traffic("k_lz", shard / lz_launches)
        inf_ms = (float(mx[4]) + float(mx[5]))
        line = {
            "metric": METRIC, "value": round(TOTAL / ((d_ms + i_ms) * 1e-3) / 1e9, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(d_ms + i_ms, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(world),
            "timing": "CUDA events on the codec's stream around each call, max over ranks; inputs resident in HBM",
            "deflate_gbs": round(TOTAL / (d_ms * 1e-3) / 1e9, 4), "inflate_gbs": round(TOTAL / (i_ms * 1e-3) / 1e9, 4),
            "targets_8gpu": {"deflate_gbs": 100, "inflate_gbs": 200},
            "ratio": round(TOTAL / clen, 4), "compressed_bytes": int(clen), "wall_s_timed_region": round(wall, 3),
            "e2e": e2e, "e2e_pageable": pageable,
            "gpu_launches": launches_total,
            "device_mem_gib": {"rank0_in_use_after_the_timed_steps": dev_mem_gib,
                               "of_which_bench_tensors": round((2 * n + cap + (cap if multi else 0)) / (1 << 30), 2),
                               "note": "the rest is the codec's workspace: inflate keeps token rows for every block of the stream it decodes (4 B per output "
                                       "byte: 32 GiB for 8 GiB) plus 16-bit symbols for one group of 8,192 blocks; deflate keeps token rows for one "
                                       "1 GiB slab of a long resident input (4 GiB; the whole shard in the two-phase sharded form), histograms and "
                                       "codes per block; buffers are over-allocated by 1/8 and never shrink"},
            "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in kt.items() if v[1]},
            "roofline": {"bound": "hbm", "kernel": "k_lz", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": tr["traffic"] if tr else None, "traffic_source": tr, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo), "launch_ms": round(lz_avg_ms, 4), "launches_per_step": lz_launches},
            "roofline_issue": issue_roofline("k_lz", shard / lz_launches, lz_avg_ms, float((clocks or {}).get("sm_mhz") or 1965.0),
                                             torch.cuda.get_device_properties(local).multi_processor_count),
            "roofline_adler": dict(adler, bound="hbm", peak=peak, unit="GB/s", frac=round(adler["gbs"] / peak, 4),
                                   note="a read-only stream; the peak is a measured COPY bandwidth, so a pure read can exceed it") if adler and "gbs" in adler else adler,
            "roofline_inflate": {"bound": "hbm", "kernels": "phase A (k_inf_tokens | k_inf_tokens4) + phase B (k_inf_resolve | k_piece_sym + k_chunk_final)",
                                 "achieved": round(algo_shard / (inf_ms * 1e-3) / 1e9, 2) if inf_ms > 0 else 0.0, "peak": peak, "unit": "GB/s",
                                 "frac": round(algo_shard / (inf_ms * 1e-3) / 1e9 / peak, 5) if inf_ms > 0 else 0.0, "algorithmic_bytes_per_launch": int(algo_shard)},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if extras:
            line["other_configs"] = extras
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs / foreign streams (context only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
