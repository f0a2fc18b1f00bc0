#!/usr/bin/env python
"""Single-call latency of dr3lk_calc_optical_flow_pyr_lk (C1) next to the GPU time of its pyramid and LK kernels
(library-side CUDA events), for the full FAST corner set and for the reference detector's operating point (<= 546 points).
usage (on a GPU box): python tools/latency_breakdown.py"""
import importlib, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import golden_case, load_gray
dr3 = importlib.import_module("3dr_b200")
a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
pts = golden_case("c1_default_21x21")["prev_pts"][:4607]
with dr3.Context(0) as ctx:
    for n in (4607, 546):
        p = pts[:n]
        for _ in range(5): ctx.calc_optical_flow_pyr_lk(a, b, p)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter(); ctx.calc_optical_flow_pyr_lk(a, b, p); ts.append(time.perf_counter() - t0)
        ctx.profile_read(); ctx.set_profiling(True)
        for _ in range(20): ctx.calc_optical_flow_pyr_lk(a, b, p)
        lk, nl, py, _ = ctx.profile_read(); ctx.set_profiling(False)
        print("n=%d wall %.1f us  (min %.1f)  gpu: pyramids %.1f us, LK %.1f us" % (n, 1e6 * np.median(ts), 1e6 * min(ts), 1e3 * py / nl, 1e3 * lk / nl))
