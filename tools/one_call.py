#!/usr/bin/env python
"""A few single calcOpticalFlowPyrLK calls (kitti0 -> kitti1, the reference detector's corners) for `ncu`:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/one_call.py
The per-kernel durations of the last call, next to the event-bracketed stage times of tools/latency_breakdown.py, show how much
of the call's GPU time is kernel execution and how much is the gap between dependent launches."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import load_gray
dr3 = importlib.import_module("3dr_b200")
a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
with dr3.Context(0) as ctx:
    xy, _, _ = ctx.fast_detect(a)
    pts = xy.astype(np.float32)
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        ctx.calc_optical_flow_pyr_lk(a, b, pts)
    print("points", len(pts))
