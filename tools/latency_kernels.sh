#!/bin/bash
# Per-kernel execution times of ONE single call (ncu, serialised) next to the event-bracketed stage times of the same call.
python tools/latency_breakdown.py 2>&1 | tail -3
python tools/one_call.py 3 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/one_call.csv python tools/one_call.py 3 > gpurun_out/one_call.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/one_call.csv") if l.startswith('"'))]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
r = rows[1:]
n = len(r)
# the last call = the last third of the launches after the detector's
per = [(x[k].split("(")[0][-40:], float(x[v].replace(",", "")), x[u]) for x in r]
for name, val, unit in per[-((n - 0) // 3 + 2):]:
    print("%-42s %9.2f %s" % (name, val, unit))
PY
