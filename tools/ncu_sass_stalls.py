"""Top SASS instructions by warp-stall samples (first kernel of an .ncu-rep), with the dominant stall reason.
usage: python tools/ncu_sass_stalls.py rep.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1: break
        continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data); totx = sum(int(r[iex]) for r in data)
print("instructions (static) %d, executed %d, samples %d" % (len(data), totx, tot))
agg = {}
for r in data:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print("stall totals: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:top]
for k in sorted(idx):
    r = data[k]
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print("%5d %5.2f%% smp %5.2f%% ex  %-60s %s" % (k, 100 * int(r[isamp]) / tot, 100 * int(r[iex]) / totx, r[isrc].strip()[:60],
                                              " ".join("%s=%d" % (h[6:], n) for n, h in st if n)))
