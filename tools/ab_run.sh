#!/bin/bash
# usage (under gpurun): tools/ab_run.sh tagA tagB ...   -- checks every variant against the oracle, then alternates them twice
for t in "$@"; do
DR3LK_LIB=$PWD/3dr_b200/lib/libdr3lk_$t.so python - <<PY
import importlib, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import oracle
from _common import load_gray, golden_case
m = importlib.import_module("3dr_b200")
ok = True
with m.Context(0) as ctx:
    for case in ("c1_default_21x21", "c1_31x31_L4", "c1_reference_30x30_initflow"):
        g = golden_case(case)
        a, b = load_gray(g["prev"]), load_gray(g["next"])
        got = ctx.calc_optical_flow_pyr_lk(a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"], g["flags"])
        exp = oracle.calc_optical_flow_pyr_lk(a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"], g["flags"])
        ok &= all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(got, exp))
print("$t", "bit-exact vs oracle" if ok else "*** WRONG RESULTS ***")
PY
done
for rep in 1 2; do
for t in "$@"; do
  DR3LK_LIB=$PWD/3dr_b200/lib/libdr3lk_$t.so python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$t', 'lk_ms %.1f pyr_ms %.1f value %.4g' % (r['lk_ms_per_launch'], r['pyramid_ms_per_step'], d['value']))"
done
done
