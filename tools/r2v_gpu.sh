#!/bin/bash
set -u
timeout 300 python tools/quick_fan.py 400 8 3 2>&1 | grep -v Voxel
timeout 300 python tools/quick_fan.py 400 8 3 u16 2>&1 | grep -v Voxel
timeout 300 python -X importtime -c "pass" 2>/dev/null
timeout 300 python - <<'PY'
import cProfile, pstats, sys, io
sys.argv = ["quick_fan.py", "300", "8", "3"]
pr = cProfile.Profile()
pr.enable()
exec(open("tools/quick_fan.py").read())
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[-3500:])
PY
