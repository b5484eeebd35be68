"""Dump SASS of the first kernel of an .ncu-rep with executed counts (per warp-quad) and samples.
usage: python tools/ncu_sass_dump.py rep.ncu-rep [min_exec_fraction]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; minf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1: break
        continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
ex = [int(r[iex]) for r in data]
ref = sorted(ex)[int(len(ex) * 0.9)]  # a typical hot-path count
tot = sum(ex)
print("hot-path reference count %d; total executed %d = %.1f x ref" % (ref, tot, tot / ref))
for k, r in enumerate(data):
    if ex[k] >= minf * ref:
        print("%5d %6.2f %5d  %s" % (k, ex[k] / ref, int(r[isamp]), r[isrc].strip()[:90]))
