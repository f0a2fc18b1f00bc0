#!/usr/bin/env python
"""Single-call latency of the LK path on one KITTI pair (kitti0 -> kitti1, 21x21, 4 levels), for the full FAST corner set and
for the reference detector's operating point (<= 546 points):
  * through the Python binding, frames in pageable memory and in page-locked memory at the device pitch,
  * the GPU time of the pyramid and LK stages of the same call (library-side CUDA events),
  * from compiled C++ (3dr_b200/host/call_latency: the header shim over the C ABI, no Python in the way), pageable / pinned / the
    frame-to-frame form with the previous frame's pyramid kept on the device.
usage (on a GPU box): python tools/latency_breakdown.py"""
import importlib, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import golden_case, load_gray
dr3 = importlib.import_module("3dr_b200")
a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
h, w = a.shape
pts = golden_case("c1_default_21x21")["prev_pts"][:4607]


def wall(fn, n=100, warm=10):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts)), 1e6 * min(ts)


with dr3.Context(0) as ctx:
    pins = [dr3.PinnedArray((h, (w + 15) // 16 * 16), np.uint8) for _ in range(2)]
    for pa, f in zip(pins, (a, b)):
        pa.array[:, :w] = f
    pa_, pb_ = pins[0].array[:, :w], pins[1].array[:, :w]
    for n in (4607, 546):
        p = pts[:n]
        med, mn = wall(lambda: ctx.calc_optical_flow_pyr_lk(a, b, p))
        medp, mnp = wall(lambda: ctx.calc_optical_flow_pyr_lk(pa_, pb_, p))
        ctx.profile_read(); ctx.set_profiling(True)
        for _ in range(20): ctx.calc_optical_flow_pyr_lk(a, b, p)
        lk, nl, py, _ = ctx.profile_read(); ctx.set_profiling(False)
        print("n=%d  python binding: pageable frames %.1f us (min %.1f), pinned frames at the device pitch %.1f us (min %.1f);  "
              "gpu stages: pyramids %.1f us, LK %.1f us" % (n, med, mn, medp, mnp, 1e3 * py / nl, 1e3 * lk / nl))
    for pa in pins:
        pa.free()

exe = os.path.join(ROOT, "3dr_b200", "host", "call_latency")
if os.path.exists(exe):
    with tempfile.TemporaryDirectory() as td:
        for name, f in (("a.pgm", a), ("b.pgm", b)):
            with open(os.path.join(td, name), "wb") as fh:
                fh.write(b"P5\n%d %d\n255\n" % (w, h)); fh.write(np.ascontiguousarray(f).tobytes())
        for n in (4607, 546):
            np.savetxt(os.path.join(td, "pts.txt"), pts[:n], fmt="%.9g")
            r = subprocess.run([exe, os.path.join(td, "a.pgm"), os.path.join(td, "b.pgm"), os.path.join(td, "pts.txt"), "300"], capture_output=True, text=True)
            d = json.loads(r.stdout) if r.returncode == 0 else {"failed": r.stderr[-200:]}
            print("n=%d  compiled C++ caller: %s" % (n, json.dumps(d)))
