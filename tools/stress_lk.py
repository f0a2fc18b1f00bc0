#!/usr/bin/env python
"""Stress: repeat the parameter-sweep cases of tests/test_gpu_lk.py many times against oracle results computed once;
report any run whose GPU result is not bit-identical.  usage (on a GPU box): python tools/stress_lk.py [reps [case,case,...]]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from _common import load_gray, random_points  # noqa: E402

dr3 = importlib.import_module("3dr_b200")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None  # case indices
CASES = [((21, 21), 3, (3, 30, 0.01), 0), ((3, 3), 5, (3, 10, 0.01), 0), ((5, 7), 2, (1, 4, 0.0), 0), ((21, 21), 0, (3, 30, 0.01), 0),
         ((31, 31), 4, (2, 0, 0.03), 0), ((30, 30), 4, (3, 1000, 1e-3), 4), ((21, 21), 3, (3, 0, 0.01), 0), ((45, 33), 3, (3, 30, 0.01), 8),
         ((21, 21), 3, (3, 30, 0.01), 12)]
a, b = load_gray("kitti3.png"), load_gray("kitti4.png")
h, w = a.shape
data = []
for win, ml, crit, flags in CASES:
    rng = np.random.default_rng(win[0] * 100 + ml)
    pts = random_points(rng, w, h, 1500)
    pts[::7] = np.round(pts[::7]); pts[1::11] = np.round(pts[1::11] * 2) / 2
    init = (pts + rng.normal(0, 2.0, pts.shape)).astype(np.float32) if flags & 4 else None
    data.append((pts, init, oracle.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags)))
bad = 0
with dr3.Context(0) as ctx:
    for r in range(reps):
        for ci, ((win, ml, crit, flags), (pts, init, exp)) in enumerate(zip(CASES, data)):
            if only is not None and ci not in only:
                continue
            p, s, e = ctx.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags)
            ds = np.where(s != exp[1])[0]
            dp = np.where((p.view(np.uint32) != exp[0].view(np.uint32)).any(axis=1))[0]
            de = np.where(e.view(np.uint32) != exp[2].view(np.uint32))[0]
            if ds.size or dp.size or de.size:
                bad += 1
                print("rep %d case %d %s: status %d pos %d err %d differ; first idx %s; pts %s got %s exp %s err got %s exp %s" % (
                    r, ci, (win, ml, crit, flags), ds.size, dp.size, de.size, (list(ds[:3]), list(dp[:3]), list(de[:3])),
                    pts[de[:2]] if de.size else pts[dp[:2]], p[dp[:2]], exp[0][dp[:2]], e[de[:3]], exp[2][de[:3]]), flush=True)
print("stress: %d runs, %d mismatching" % (reps * (len(only) if only else len(CASES)), bad))
