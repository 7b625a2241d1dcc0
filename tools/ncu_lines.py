"""Per-source-line instruction and stall-sample totals for one kernel of an .ncu-rep.

ncu's CSV source page has no CUDA-line view with metrics, so this joins its SASS page
(by instruction order) with `nvdisasm -g` line info of the same cubin.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled-name substring> [top]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
inst = []
for r in rows[2:]:
    if len(r) <= ismp or r[0] == "Kernel Name" or r[0] == "Address":
        if r and r[0] == "Kernel Name" and inst:
            break  # first launch only
        continue
    inst.append((r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "zlib.es_b200", "libzles.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
lines = []
cur = None
active = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        active = mangled in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(1).strip()))
if len(lines) != len(inst):
    print("warning: %d SASS instructions in the report, %d in the cubin" % (len(inst), len(lines)))
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for (src, ex, smp), (loc, _) in zip(inst, lines):
    agg[loc][0] += ex
    agg[loc][1] += smp
    tot_i += ex
    tot_s += smp
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
srcs = {}
def text(loc):
    if not loc:
        return ""
    f = os.path.join(ROOT, "zlib.es_b200", "csrc", loc[0])
    if f not in srcs:
        srcs[f] = open(f).read().splitlines() if os.path.exists(f) else []
    return srcs[f][loc[1] - 1].strip()[:90] if loc[1] - 1 < len(srcs[f]) else ""
for loc, (ex, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%s  %s" % (100.0 * smp / max(1, tot_s), 100.0 * ex / max(1, tot_i), loc[0] if loc else "?", loc[1] if loc else "?", text(loc)))
