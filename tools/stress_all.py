#!/usr/bin/env python
"""Determinism stress over every entry point: each path runs once for a reference result, then `reps` more times; any
repetition that is not bit-identical is reported.  Rare races (a copy that had not landed, a stale scratch buffer) show
up here long before they show up in a single test run.  usage (on a GPU box): python tools/stress_all.py [reps]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import golden_case, load_gray, random_points  # noqa: E402

dr3 = importlib.import_module("3dr_b200")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
frames = [load_gray("kitti%d.png" % i) for i in range(10)]
odd = load_gray("kitti_000000.png")  # 1241 wide
rng = np.random.default_rng(7)
pts = golden_case("c1_default_21x21")["prev_pts"][:3000]


def flat(x):
    if x is None:
        return b""
    if isinstance(x, (tuple, list)):
        return b"".join(flat(y) for y in x)
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x).tobytes()
    return repr(x).encode()


def paths(ctx):
    prev = np.ascontiguousarray(np.stack(frames[:8])); nxt = np.ascontiguousarray(np.stack(frames[1:9]))
    plist = [random_points(np.random.default_rng(40 + i), 1240, 376, n) for i, n in enumerate((900, 0, 1, 333, 1500, 64, 2048, 700))]
    offs = np.concatenate([[0], np.cumsum([len(p) for p in plist])]).astype(np.int32)
    allp = np.concatenate(plist).astype(np.float32)
    yield "single 21x21", lambda: ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], pts)
    yield "single 21x21 zero iterations", lambda: ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], pts, None, (21, 21), 3, (3, 0, 0.01))
    yield "single 30x30 warm start", lambda: ctx.calc_optical_flow_pyr_lk(frames[0], frames[2], pts, pts + np.float32(0.5), (30, 30), 4, (3, 1000, 1e-3), 4)
    yield "single 31x31 odd width", lambda: ctx.calc_optical_flow_pyr_lk(odd, odd[::-1].copy()[::-1], pts[:1500], None, (31, 31), 4)
    yield "single generic 9x13", lambda: ctx.calc_optical_flow_pyr_lk(frames[3], frames[4], pts[:800], None, (9, 13), 2)
    yield "host batch ragged (chunk ramp)", lambda: ctx.track_batch_host(prev, nxt, allp, offs, want_stats=True)
    yield "host batch ragged chunk 3", lambda: ctx.track_batch_host(prev, nxt, allp, offs, want_stats=True, chunk_pairs=3)

    def chain():
        out, cur = [], pts
        pyr = dr3.Pyramid(ctx, frames[0], (21, 21), 3)
        for i in range(5):
            p, s, e, nx = ctx.track_frame(pyr, frames[i + 1], cur, keep_next=2)
            out += [p, s, e]
            pyr.close(); pyr, cur = nx, p[s == 1]
        pyr.close()
        return out
    yield "track_frame chain", chain
    yield "lk pyramid + derivatives", lambda: ctx.build_lk_pyramid(odd, (21, 21), 3, True)
    yield "box pyramid", lambda: ctx.box_pyramid(frames[5], 3)
    yield "fast detector", lambda: ctx.fast_detect(frames[6])
    p1, s1, _ = ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], pts)
    yield "filter tracks", lambda: ctx.filter_tracks(pts, p1, s1, 718.856, 718.856, 607.19, 185.22)
    yield "filter tracks (distorted pinhole)", lambda: ctx.filter_tracks(pts, p1, s1, 458.654, 457.296, 367.215, 248.375, (-0.2834, 0.0740, 1.9e-4, 1.8e-5, 0.0))
    F = np.random.default_rng(5).normal(0, 1e-3, (200, 3, 3)).astype(np.float32)
    yield "score fundamental", lambda: ctx.score_fundamental(F, pts[s1 == 1], p1[s1 == 1], 1.0)

    def init_chain():
        xy, lv, sc, pyr = ctx.init_first_frame(frames[0])
        k = xy.astype(np.float32)
        out = ctx.init_second_frame(pyr, frames[1], k, k, fx=718.856, fy=718.856, cx=607.19, cy=185.22)
        res = ctx.init_score_fundamental(F, len(out["ref"]))
        pyr.close()
        return [xy, sc, out["ref"], out["cur"], out["disparity"], out["bearing"], out["status"], res[0], res[1]]
    yield "init chain (first frame, second frame, scoring)", init_chain
    mc = dr3.MultiContext([0, 0, 0])
    yield "multi-context host batch (3 ranks on one device)", lambda: mc.track_batch_host(prev, nxt, allp, offs, want_stats=True)


bad = 0
with dr3.Context(0) as ctx:
    lst = list(paths(ctx))
    refs = [flat(fn()) for _, fn in lst]
    for r in range(reps):
        for (name, fn), ref in zip(lst, refs):   # interleaved on purpose: scratch buffers change hands between paths
            if flat(fn()) != ref:
                bad += 1
                print("rep %d: %s is not reproducible" % (r, name), flush=True)
print("stress_all: %d paths x %d repetitions, %d mismatching" % (len(lst), reps, bad))
sys.exit(1 if bad else 0)
