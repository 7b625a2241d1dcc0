"""Summarise .ncu-rep files into profiles/ (tracked): per-kernel key metrics as JSON + markdown.
usage: python tools/ncu_summary.py <report.ncu-rep> <out-prefix>"""
import csv, io, json, subprocess, sys
rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_inst",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
        "launch__shared_mem_per_block_dynamic": "dyn_smem"}
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return x
def to_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(u, 1)
out = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("zles::", "")
    d = {}
    for k, short in want.items():
        if k in hdr:
            i = hdr.index(k); v = num(r[i])
            if short in ("dram_read", "dram_write") and isinstance(v, float): v = to_bytes(v, units[i])
            if short == "duration" and isinstance(v, float): v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(units[i], 1)  # -> ms
            d[short] = v
    st = [(num(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
          if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i]]
    d["top_stalls"] = [[h, v] for v, h in sorted(st, reverse=True)[:6]]
    if "dram_read" in d and "dram_write" in d: d["dram_bytes_per_launch"] = d["dram_read"] + d["dram_write"]
    out.setdefault(name, d)
json.dump(out, open(prefix + ".json", "w"), indent=1)
with open(prefix + ".md", "w") as f:
    f.write("| kernel | ms | DRAM read MB | DRAM write MB | DRAM %% | occupancy %% | issue active %% | threads/inst | regs | smem bank conflicts | top stalls |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for k, d in out.items():
        f.write("| %s | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %s | %s |\n" % (
            k, d.get("duration", 0), d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6, d.get("dram_pct", 0), d.get("occupancy_pct", 0),
            d.get("issue_active_pct", 0), d.get("threads_per_inst", 0), int(d.get("regs", 0)), int(d.get("smem_bank_conflicts", 0)),
            ", ".join("%s %.2f" % (h, v) for h, v in d["top_stalls"][:4])))
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "top_stalls"} for k, v in out.items()}, indent=0)[:1500])
