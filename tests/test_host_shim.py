"""GPU test of the C++14 host shim: the compiled mirror of the reference call site (3dr_b200/host/init_frontend.cpp,
built against include/dr3lk.hpp) must reproduce the oracle run with the reference's literal parameters."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from _common import ROOT, golden_case, load_gray

pytestmark = pytest.mark.gpu


def _write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img).tobytes())


def test_init_frontend_matches_oracle(tmp_path):
    exe = os.path.join(ROOT, "3dr_b200", "host", "init_frontend")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe)], check=True)
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = golden_case("c1_reference_30x30_initflow")["prev_pts"][:1500]
    _write_pgm(tmp_path / "a.pgm", a)
    _write_pgm(tmp_path / "b.pgm", b)
    np.savetxt(tmp_path / "pts.txt", pts, fmt="%.9g")
    r = subprocess.run([exe, str(tmp_path / "a.pgm"), str(tmp_path / "b.pgm"), str(tmp_path / "pts.txt"), str(tmp_path / "out.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Outlier count from optical flow" in r.stdout and "Average disparity" in r.stdout
    with open(tmp_path / "out.txt") as f:
        lines = f.read().split("\n")
    outliers, matches, mean_disp = lines[0].split()
    # the reference's literal call: 30x30, maxLevel 4, (COUNT+EPS, 1000, 1e-3), OPTFLOW_USE_INITIAL_FLOW with nextPts = prevPts
    p, s, e = oracle.calc_optical_flow_pyr_lk(a, b, pts, pts.copy(), (30, 30), 4, (3, 1000, 1e-3), 4)
    assert int(outliers) == int((s == 0).sum()) and int(matches) == int(s.sum())
    disp = np.linalg.norm(pts[s == 1].astype(np.float64) - p[s == 1].astype(np.float64), axis=1)
    assert abs(float(mean_disp) - disp.mean()) < 1e-5
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[3:] if ln.strip()], np.float64)
    assert rows.shape == (int(s.sum()), 4)
    assert np.abs(rows[:, 0:2] - pts[s == 1]).max() < 1e-4 and np.abs(rows[:, 2:4] - p[s == 1]).max() < 1e-4
    box = oracle.box_pyramid(a, 3)
    assert [int(v) for v in lines[1].split()] == [620, 188, 310, 94]
    assert [int(v) for v in lines[2].split()] == [int(box[1].sum()), int(box[2].sum())]

    # The same call site with the reference's OWN types (cv::Mat, std::vector<cv::Point2f>, cv::TermCriteria ...) through
    # include/dr3lk_opencv.hpp: the swap is a namespace change.  Built against tests/mock_opencv (no OpenCV headers in this
    # image); its results must be the raw-pointer shim's, line for line.
    exe2 = os.path.join(ROOT, "3dr_b200", "host", "opencv_callsite")
    if not os.path.exists(exe2):
        subprocess.run(["make", "-C", os.path.dirname(exe2), "opencv_callsite"], check=True)
    r2 = subprocess.run([exe2, str(tmp_path / "a.pgm"), str(tmp_path / "b.pgm"), str(tmp_path / "pts.txt"), str(tmp_path / "out2.txt")],
                        capture_output=True, text=True)
    assert r2.returncode == 0, (r2.returncode, r2.stderr)
    with open(tmp_path / "out2.txt") as f:
        lines2 = f.read().split("\n")
    assert lines2 == lines


def test_default_context_device_selection(tmp_path):
    """Context::thread_default() honours DR3LK_DEVICE; a device that does not exist is an error, not a silent fallback"""
    exe = os.path.join(ROOT, "3dr_b200", "host", "init_frontend")
    a = load_gray("kitti0.png")
    _write_pgm(tmp_path / "a.pgm", a)
    np.savetxt(tmp_path / "pts.txt", np.array([[100.0, 100.0]]), fmt="%.9g")
    args = [exe, str(tmp_path / "a.pgm"), str(tmp_path / "a.pgm"), str(tmp_path / "pts.txt"), str(tmp_path / "o.txt")]
    assert subprocess.run(args, env=dict(os.environ, DR3LK_DEVICE="0"), capture_output=True).returncode == 0
    r = subprocess.run(args, env=dict(os.environ, DR3LK_DEVICE="4242"), capture_output=True, text=True)
    assert r.returncode != 0


def test_call_latency_program_gets_identical_results_from_all_three_forms(tmp_path):
    """3dr_b200/host/call_latency.cpp (the compiled-C++ leg of bench.py's latency block): the plain call with pageable frames, the
    same call with page-locked frames at the device pitch (direct upload) and the frame-to-frame form must return identical
    points and status (the program compares them itself; the oracle parity of each form is in test_gpu_lk / test_gpu_next_rows)."""
    exe = os.path.join(ROOT, "3dr_b200", "host", "call_latency")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe)], check=True)
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = golden_case("c1_default_21x21")["prev_pts"][:400]
    _write_pgm(tmp_path / "a.pgm", a)
    _write_pgm(tmp_path / "b.pgm", b)
    np.savetxt(tmp_path / "pts.txt", pts, fmt="%.9g")
    more = []
    for i in (2, 3):
        _write_pgm(tmp_path / ("f%d.pgm" % i), load_gray("kitti%d.png" % i))
        more.append(str(tmp_path / ("f%d.pgm" % i)))
    r = subprocess.run([exe, str(tmp_path / "a.pgm"), str(tmp_path / "b.pgm"), str(tmp_path / "pts.txt"), "20"] + more, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    import json
    d = json.loads(r.stdout)
    assert d["identical_results"] is True and d["points"] == 400
    assert 0 < d["c_abi_call_us_pinned"] and 0 < d["c_abi_call_us_pageable"] and 0 < d["c_abi_track_frame_us_pinned"]
    # the chain kitti0 -> 1 -> 2 -> 3, survivors carried forward: as many as the oracle keeps
    cur = pts
    for i in range(3):
        p, s, _ = oracle.calc_optical_flow_pyr_lk(load_gray("kitti%d.png" % i), load_gray("kitti%d.png" % (i + 1)), cur)
        cur = p[s == 1]
    assert d["chain_frames"] == 4 and d["chain_survivors"] == len(cur) and d["chain_ms_pinned"] > 0
