#!/usr/bin/env python
"""Informational numbers for the BASELINE.json configs that are not the bench.py headline (C1, C2, C4, C5):
single-call latency through the host-buffer C ABI and batched device-resident throughput, next to cv2 on the host.
Run on the GPU box:  python tools/bench_configs.py > gpurun_out/configs.json"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from _common import golden_case, load_gray  # noqa: E402
from oracle import cv2_ref  # noqa: E402
from tools import synth  # noqa: E402

dr3 = importlib.import_module("3dr_b200")
ctx = dr3.Context(0)
out = {}


def time_call(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def cpu_time(a, b, pts, win, ml, crit, flags=0, init=None, n=5):
    return time_call(lambda: cv2_ref.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags), n=n, warm=1)


# ---- C1: kitti0 -> kitti1, FAST corners; default and the reference's literal parameters
a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
pts = golden_case("c1_default_21x21")["prev_pts"][:4607]
for name, (win, ml, crit, flags) in {"c1_21x21": ((21, 21), 3, (3, 30, 0.01), 0), "c1_reference_30x30": ((30, 30), 4, (3, 1000, 1e-3), 4)}.items():
    init = pts.copy() if flags & 4 else None
    t = time_call(lambda: ctx.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags))
    p, s, e = ctx.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags)
    tc = cpu_time(a, b, pts, win, ml, crit, flags, init)
    out[name] = {"points": len(pts), "tracked": int(s.sum()), "gpu_call_ms": 1e3 * t, "gpu_features_per_s": len(pts) / t,
                 "cv2_call_ms": 1e3 * tc, "cv2_features_per_s": len(pts) / tc, "cv2_threads": cv2_ref.cv2.getNumThreads()}

# ---- C2: kitti0..9 frame-to-frame chain (9 dependent calls)
frames = [load_gray("kitti%d.png" % i) for i in range(10)]


def chain(track):
    cur, surv = pts, []
    for i in range(9):
        p, s, _ = track(frames[i], frames[i + 1], cur)
        cur = p[s == 1]; surv.append(len(cur))
    return surv


t = time_call(lambda: chain(lambda x, y, c: ctx.calc_optical_flow_pyr_lk(x, y, c)), n=10)
tc = time_call(lambda: chain(lambda x, y, c: cv2_ref.calc_optical_flow_pyr_lk(x, y, c)), n=3, warm=1)

def chain_cached():
    # streaming use: one pyramid per incoming frame, the previous frame's pyramid is reused as the LK template side
    prev_pyr = dr3.Pyramid(ctx, frames[0], (21, 21), 3)
    cur, surv = pts, []
    for i in range(9):
        nxt_pyr = dr3.Pyramid(ctx, frames[i + 1], (21, 21), 3)
        p, s, _ = ctx.calc_optical_flow_pyr_lk_cached(prev_pyr, nxt_pyr, cur)
        cur = p[s == 1]; surv.append(len(cur))
        prev_pyr.close()
        prev_pyr = nxt_pyr
    prev_pyr.close()
    return surv


t_cached = time_call(chain_cached, n=10)
out["c2_chain_cached_pyramids_ms"] = 1e3 * t_cached


def chain_streaming():
    # one call per incoming frame: upload + pyramid of the new frame + LK against the previous frame's cached pyramid
    prev_pyr = dr3.Pyramid(ctx, frames[0], (21, 21), 3)
    cur, surv = pts, []
    for i in range(9):
        p, s, _, nxt_pyr = ctx.track_frame(prev_pyr, frames[i + 1], cur, keep_next=2)
        cur = p[s == 1]; surv.append(len(cur))
        prev_pyr.close()
        prev_pyr = nxt_pyr
    prev_pyr.close()
    return surv


out["c2_chain_streaming_ms"] = 1e3 * time_call(chain_streaming, n=10)
assert chain_streaming() == chain_cached()
out["c2_chain"] = {"survivors": chain(lambda x, y, c: ctx.calc_optical_flow_pyr_lk(x, y, c)), "gpu_chain_ms": 1e3 * t, "cv2_chain_ms": 1e3 * tc}


def device_batch(prev, nxt, pts_list, win, ml, crit, reps=3):
    B, h, w = prev.shape
    offs = np.concatenate([[0], np.cumsum([len(p) for p in pts_list])]).astype(np.int32)
    allp = np.concatenate(pts_list).astype(np.float32)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dp, dn = torch.from_numpy(prev).cuda(), torch.from_numpy(nxt).cuda()
        dpts = torch.from_numpy(allp).cuda(); dnext = torch.zeros_like(dpts)
        dst = torch.zeros(len(allp), dtype=torch.uint8, device="cuda"); derr = torch.zeros(len(allp), device="cuda")
        ctx.set_stream(st.cuda_stream)
        run = lambda: ctx.track_batch(dp.data_ptr(), dn.data_ptr(), w, h, w, h * w, B, dpts.data_ptr(), dnext.data_ptr(), dst.data_ptr(),
                                      derr.data_ptr(), offs, None, win, ml, crit, 0)
        run(); run(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            run()
        e1.record(st); st.synchronize()
        ctx.set_stream(None)
        return e0.elapsed_time(e1) / reps * 1e-3, int(dst.sum().item()), len(allp)


# ---- C4: 3840x2160, 50k corners, 31x31, 5 levels; batch of 8 pairs
a4, b4, _ = synth.make_pair(2000, 3840, 2160)
p4 = synth.corners(a4, 50000, 5)
t, tracked, n = device_batch(np.stack([a4] * 8), np.stack([b4] * 8), [p4] * 8, (31, 31), 4, (3, 30, 0.01))
tc = cpu_time(a4, b4, p4, (31, 31), 4, (3, 30, 0.01), n=3)
out["c4_4k_31x31"] = {"pairs": 8, "points_per_pair": len(p4), "gpu_features_per_s": n / t, "gpu_ms_per_pair": 1e3 * t / 8,
                      "cv2_features_per_s": len(p4) / tc, "cv2_ms_per_pair": 1e3 * tc}

# ---- C5: semi-dense lattice on 1241x376, batch of 256 pairs
a5, b5, _ = synth.make_pair(1000)
p5 = synth.lattice(1241, 376)
t, tracked, n = device_batch(np.stack([a5] * 256), np.stack([b5] * 256), [p5] * 256, (21, 21), 3, (3, 30, 0.01))
tc = cpu_time(a5, b5, p5, (21, 21), 3, (3, 30, 0.01), n=3)
out["c5_semidense"] = {"pairs": 256, "points_per_pair": len(p5), "tracked_fraction": tracked / n, "gpu_features_per_s": n / t,
                       "cv2_features_per_s": len(p5) / tc}
# ---- KITTI batch: the 9 bundled consecutive pairs kitti0..9 tiled to 2304 device-resident pairs, FAST corners of each prev frame
kp = [cv2_ref.fast_corners(frames[i])[0] for i in range(9)]
reps = 256
prev_k = np.stack([frames[i] for i in range(9)] * reps); next_k = np.stack([frames[i + 1] for i in range(9)] * reps)
t, tracked, n = device_batch(prev_k, next_k, kp * reps, (21, 21), 3, (3, 30, 0.01))
tc = sum(cpu_time(frames[i], frames[i + 1], kp[i], (21, 21), 3, (3, 30, 0.01), n=3) for i in range(9))
out["kitti_batch_21x21"] = {"pairs": 9 * reps, "points": n, "tracked_fraction": tracked / n, "gpu_features_per_s": n / t,
                            "gpu_tracked_per_s": tracked / t, "cv2_features_per_s": sum(len(k) for k in kp) / tc}
# ---- a-1..a-3: the Frame box pyramid (utils::create_img_pyramid, 3 levels) on device-resident batches
def box_batch(img, n_img=1024, reps=5):
    h, w = img.shape
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d0 = torch.from_numpy(np.ascontiguousarray(img)).cuda().unsqueeze(0).repeat(n_img, 1, 1).contiguous()
        d1 = torch.empty((n_img, h // 2, w // 2), dtype=torch.uint8, device="cuda")
        d2 = torch.empty((n_img, h // 4, w // 4), dtype=torch.uint8, device="cuda")
        ctx.set_stream(st.cuda_stream)
        run = lambda: ctx.box_pyramid_device(d0.data_ptr(), w, h, w, w * h, n_img, [d1.data_ptr(), d2.data_ptr()])
        run(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            run()
        e1.record(st); st.synchronize()
        ctx.set_stream(None)
    t = e0.elapsed_time(e1) / reps * 1e-3
    nbytes = n_img * (w * h + 2 * (w // 2) * (h // 2) + (w // 4) * (h // 4))  # read L0, write+read L1, write L2
    tc = time_call(lambda: oracle.box_pyramid(img, 3), n=20)
    return {"images": n_img, "gpu_us_per_image": 1e6 * t / n_img, "gpu_gb_per_s": nbytes / t / 1e9, "cpu_restatement_us_per_image": 1e6 * tc,
            "cpu": "oracle/lk_oracle.c (1 thread)"}


import oracle  # noqa: E402
out["a1_box_pyramid_1240x376"] = box_batch(frames[0])
out["a1_box_pyramid_1241x376_sheared"] = box_batch(load_gray("kitti_000000.png"))

# ---- "next" rows of SURVEY.md 8f through their host-buffer C-ABI calls, next to the CPU restatements (oracle/) on the host
import oracle  # noqa: E402
from oracle import postfilter  # noqa: E402
t = time_call(lambda: ctx.fast_detect(frames[0]), n=20)
tc = time_call(lambda: oracle.fast_detector(frames[0]), n=3, warm=1)
xy, lv, sc = ctx.fast_detect(frames[0])
out["f1_fast_detect_kitti0"] = {"corners": int(len(xy)), "gpu_call_ms": 1e3 * t, "cpu_restatement_ms": 1e3 * tc, "cpu": "oracle/fast_oracle.c, 1 thread"}
p1, s1, _ = ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], pts)
t = time_call(lambda: ctx.filter_tracks(pts, p1, s1, 718.856, 718.856, 607.19, 185.22), n=20)
tc = time_call(lambda: postfilter.filter_tracks(pts, p1, s1, 718.856, 718.856, 607.19, 185.22), n=5, warm=1)
out["f3_filter_tracks"] = {"points": int(len(pts)), "gpu_call_ms": 1e3 * t, "cpu_restatement_ms": 1e3 * tc, "cpu": "oracle/postfilter.py (numpy)"}
rng = np.random.default_rng(5)
keep = s1 == 1
F = rng.normal(0, 1e-3, (200, 3, 3)).astype(np.float32)
t = time_call(lambda: ctx.score_fundamental(F, pts[keep], p1[keep], 1.0), n=20)
tc = time_call(lambda: postfilter.check_fundamental(F, pts[keep], p1[keep], 1.0), n=1, warm=0)
out["f4_score_fundamental"] = {"hypotheses": 200, "matches": int(keep.sum()), "gpu_call_ms": 1e3 * t, "cpu_restatement_ms": 1e3 * tc,
                               "cpu": "oracle/postfilter.py (numpy across hypotheses, python loop over matches in the reference's fp32 order)"}
out["host"] = {"cpus": os.cpu_count(), "cv2": cv2_ref.CV2_VERSION, "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out, indent=1))
