"""Executed-instruction / stall-sample attribution per CUDA source line from an .ncu-rep.
usage: python tools/ncu_lines.py rep.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = []; tot = tots = 0; first = None; skip = False
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name":
        fn = r[1]
        if first is None: first = fn
        skip = fn != first
        continue
    if skip: continue
    if len(r) > 7 and r[2] == "-" and r[0].isdigit():
        agg.append((int(r[7]), int(r[6]), cur, r[0], r[1].strip()[:105]))
        tot += int(r[7]); tots += int(r[6])
agg.sort(reverse=True)
print("total executed %d, samples %d" % (tot, tots))
for e, s, f, l, src in agg[:top]:
    print("%5.1f%% ex %5.1f%% smp  %s:%s  %s" % (100 * e / tot, 100 * s / max(tots, 1), f, l, src))
