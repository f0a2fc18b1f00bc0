"""Generates tests/golden/*.npz and pyramid_hashes.json from the real OpenCV code path (cv2 wheel).

Run in the build container (cv2 4.13.0 present):   python tests/golden/make_golden.py
The fixtures pin oracle/lk_oracle.c (and through it the CUDA path) to what cv::calcOpticalFlowPyrLK /
cv::buildOpticalFlowPyramid -- the functions the reference calls at src/initialization.cpp:608-613 --
actually compute.  Inputs are the reference's bundled KITTI frames (copied to data/).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from oracle import cv2_ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    return cv2.imread(os.path.join(ROOT, "data", name), cv2.IMREAD_GRAYSCALE)


# name -> (prev, next, win, maxLevel, criteria, flags, init)   init: None | "prev" | "noisy"
LK_CASES = {
    "c1_default_21x21": ("kitti0.png", "kitti1.png", (21, 21), 3, (3, 30, 0.01), 0, None),
    "c1_reference_30x30_initflow": ("kitti0.png", "kitti1.png", (30, 30), 4, (3, 1000, 1e-3), 4, "prev"),
    "c1_31x31_L4": ("kitti0.png", "kitti1.png", (31, 31), 4, (3, 30, 0.01), 0, None),
    "c1_mineig_21x21": ("kitti0.png", "kitti1.png", (21, 21), 3, (3, 30, 0.01), 8, None),
    "oddwidth_21x21": ("kitti_000000.png", "kitti_000001.png", (21, 21), 3, (3, 30, 0.01), 0, None),
    "c1_noisy_init_15x9": ("kitti0.png", "kitti2.png", (15, 9), 2, (3, 20, 0.03), 4, "noisy"),
}


def main():
    hashes = {"cv2_version": cv2.__version__, "images": {}}
    for name in ["kitti0.png", "kitti1.png", "kitti_000000.png", "sample_gray_500x375.png"]:
        im = load(name)
        lv, dv = cv2_ref.build_lk_pyramid(im, (21, 21), 3, True)
        hashes["images"][name] = {"shape": list(im.shape), "gauss": [sha(a) for a in lv], "scharr": [sha(a) for a in dv]}
    with open(os.path.join(OUT, "pyramid_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)

    rng = np.random.default_rng(20261018)
    for case, (pa, pb, win, ml, crit, flags, init) in LK_CASES.items():
        a, b = load(pa), load(pb)
        pts, _ = cv2_ref.fast_corners(a)
        h, w = a.shape
        extra = np.stack([rng.uniform(-40, w + 40, 300), rng.uniform(-40, h + 40, 300)], 1).astype(np.float32)
        allp = np.concatenate([pts, extra]).astype(np.float32)
        ini = None
        if init == "prev":
            ini = allp.copy()
        elif init == "noisy":
            ini = (allp + rng.normal(0, 1.5, allp.shape)).astype(np.float32)
        p1, st, err = cv2_ref.calc_optical_flow_pyr_lk(a, b, allp, ini, win, ml, crit, flags)
        err = np.where(st == 1, err, 0).astype(np.float32) if not (flags & 8) else err  # lost-point err is uninitialised in OpenCV
        np.savez_compressed(os.path.join(OUT, "lk_%s.npz" % case), prev=pa, next=pb, win=np.array(win), max_level=ml,
                            crit=np.array(crit, np.float64), flags=flags, prev_pts=allp,
                            init=ini if ini is not None else np.zeros((0, 2), np.float32), next_pts=p1, status=st, err=err)
        print(case, "n", len(allp), "tracked", int(st.sum()))

    # cv::undistortPoints as Pinhole::cam2world calls it for a distorted camera (src/camera.cpp:32-40): float K / D, 32FC2 points
    cams = [(458.654, 457.296, 367.215, 248.375, (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0), 752, 480),
            (718.856, 718.856, 607.1928, 185.2157, (-0.3, 0.1, 0.001, -0.002, 0.05), 1241, 376),
            (300.0, 300.0, 320.0, 240.0, (0.9, 2.5, 0.01, 0.01, 1.0), 640, 480)]  # the last one reaches the icdist < 0 exit
    und = {}
    for ci, (fx, fy, cx, cy, dist, w, h) in enumerate(cams):
        uv = np.stack([rng.uniform(-100, w + 100, 2000), rng.uniform(-100, h + 100, 2000)], 1).astype(np.float32)
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        D = np.array([dist], np.float32)
        und["cam%d" % ci] = np.array([fx, fy, cx, cy] + list(dist), np.float64)
        und["uv%d" % ci] = uv
        und["xy%d" % ci] = cv2.undistortPoints(uv.reshape(-1, 1, 2), K, D).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "undistort.npz"), **und)
    print("undistort", {k: v.shape for k, v in und.items()})

    # frame-to-frame chain kitti0..9 (config C2): survivors per step with default parameters
    frames = [load("kitti%d.png" % i) for i in range(10)]
    pts, _ = cv2_ref.fast_corners(frames[0])
    surv = [len(pts)]
    cur = pts
    for i in range(9):
        p1, st, _ = cv2_ref.calc_optical_flow_pyr_lk(frames[i], frames[i + 1], cur, None, (21, 21), 3, (3, 30, 0.01), 0)
        cur = p1[st == 1]
        surv.append(len(cur))
    # reference-style: anchored on kitti0 with warm start, 30x30
    ref, curp = pts.copy(), pts.copy()
    anch = []
    for i in range(1, 10):
        p1, st, _ = cv2_ref.calc_optical_flow_pyr_lk(frames[0], frames[i], ref, curp, (30, 30), 4, (3, 1000, 1e-3), 4)
        ref, curp = ref[st == 1], p1[st == 1]
        anch.append(len(ref))
    with open(os.path.join(OUT, "chain_counts.json"), "w") as f:
        json.dump({"chain_21x21": surv, "anchored_30x30": anch}, f)
    print("chain", surv, "anchored", anch)


if __name__ == "__main__":
    main()
