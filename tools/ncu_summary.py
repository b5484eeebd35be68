"""Summarise an .ncu-rep (read on the CPU box): per-kernel headline metrics + stall mix + opcode mix.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if sub not in name:
        continue
    print("=== %s" % name[:70])
    for w in WANT:
        if w in hdr:
            print("  %-66s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = {h: float(r[i]) for i, h in enumerate(hdr)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]}
    tot = sum(st.values()) or 1.0
    print("  stall mix: " + ", ".join("%s %.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot)
                                        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name" and sub in rows[k][1]:
        hdr = rows[k + 1]
        ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        g = defaultdict(lambda: [0, 0])
        j = k + 2
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            r = rows[j]
            op = [o for o in r[ia].split() if not o.startswith("@")][0].split(".")[0]
            g[op][0] += int(r[isamp]); g[op][1] += int(r[iex])
            j += 1
        ts = sum(v[0] for v in g.values()) or 1; te = sum(v[1] for v in g.values()) or 1
        print("=== opcode mix of %s (executed %d warp-instr)" % (rows[k][1][:50], te))
        print("  " + ", ".join("%s %.1f%%/%.1f%%s" % (op, 100 * e / te, 100 * s / ts)
                                 for op, (s, e) in sorted(g.items(), key=lambda kv: -kv[1][1])[:18]))
        break
    k += 1
