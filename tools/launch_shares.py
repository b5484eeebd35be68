"""Kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
t = defaultdict(float); n = defaultdict(int)
for r in rows[1:]:
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    k = r[ik].split("(")[0]; t[k] += v; n[k] += 1
tot = sum(t.values())
for k in sorted(t, key=lambda k: -t[k]):
    print("%-50s n=%4d total %10.1f us  share %5.1f%%  avg %8.1f us" % (k[:50], n[k], t[k] / 1e3, 100 * t[k] / tot, t[k] / n[k] / 1e3))
