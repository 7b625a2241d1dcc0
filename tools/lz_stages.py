"""Per-stage cycle breakdown of k_lz (profiling build, -DZLES_STAGE_CLOCKS; not the product library).

usage: python tools/lz_stages.py [kind=0 text|1 binary|2 random|3 mixed] [MiB=64] [scan deep [window_mode]]
Builds build/libzles_prof.so on first use (here, before gpurun), then runs one warm deflate and prints the
share of each stage in thread 0's cycles summed over all CTAs.
"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "build", "libzles_prof.so")
SRC = os.path.join(ROOT, "zlib.es_b200", "csrc", "zles.cu")
NAMES = ["S0 stage(TMA)", "S1 adler", "S2 p1 count", "S2 p1 scan", "S2 p1 scatter", "S2 p2 count", "S2 p2 scan", "S2 p2 scatter",
         "S3 match", "S4 copy", "S4 parse", "S5 emit"]


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-shared", "-DZLES_STAGE_CLOCKS", "-o", LIB, SRC])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build(); sys.exit(0)
    if not os.path.exists(LIB):
        build()
    import torch, zles
    from zles import _capi
    L = ctypes.CDLL(LIB); _capi.bind(L)
    c = zles.Codec(0, lib=L)
    kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 20
    if len(sys.argv) > 4:  # search level: scan width, extensions
        c.set_level(int(sys.argv[3]), int(sys.argv[4]), 8, True)
    if len(sys.argv) > 5:  # window mode (zles_ctx_set_window_mode)
        c.set_window_mode(int(sys.argv[5]))
    src = torch.empty(n, dtype=torch.uint8, device="cuda"); c.dev_corpus(kind, 0, src.data_ptr(), n)
    cap = c.deflate_bound(n); comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    L.zles_debug_lz_clocks(None, 1)
    clen = c.dev_deflate(src.data_ptr(), n, comp.data_ptr(), cap)
    torch.cuda.synchronize()
    L.zles_debug_lz_clocks(out, 0)
    tot = sum(out[:12])
    print("kind", kind, "MiB", n >> 20, "level", sys.argv[3:6], "comp", clen, "ratio %.4f" % (n / clen), "cycles per 32 KiB block: %.0f" % (tot / (n / 32768)))
    for i, nm in enumerate(NAMES):
        print("  %-16s %6.2f %%   %8.0f cycles/block" % (nm, 100.0 * out[i] / tot, out[i] / (n / 32768)))
    w = sum(out[12:15]) or 1
    print("  inside S3 (warp 0's own cycles): fill + run lengths %.1f %%, scan %.1f %%, finalize %.1f %%" % tuple(100.0 * out[i] / w for i in (12, 13, 14)))
    # phase A of inflate on the stream just made
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n)
    torch.cuda.synchronize()
    io = (ctypes.c_ulonglong * 8)()
    L.zles_debug_inf_clocks(None, 1)
    assert c.dev_inflate(comp.data_ptr(), clen, back.data_ptr(), n) == n
    torch.cuda.synchronize()
    L.zles_debug_inf_clocks(io, 0)
    nseg = (n + 32767) // 32768
    print("  k_inf_tokens4: %d blocks decoded four-way, %d fell back to one warp after trying, %.2f stitch continuations per block" % (io[6], io[7], io[4] / nseg))
    print("  k_inf_tokens per segment: header+tables %.0f cycles, symbol loop %.0f cycles, reader re-seating %.0f cycles; %.0f tokens, %.0f rounds (%.2f tokens/round, %.0f cycles/round), %.1f tokens decoded alone"
          % (io[0] / nseg, io[1] / nseg, io[5] / nseg, io[3] / nseg, io[2] / nseg, io[3] / max(io[2], 1), io[1] / max(io[2], 1), io[4] / nseg))
