"""CPU tests: the oracle restatement against the golden vectors produced by the real OpenCV code path, against cv2
live when importable, and the box-pyramid known answers of SURVEY.md 8c."""
import numpy as np
import pytest

import oracle
from oracle import cv2_ref
from _common import ALL_TRACKED_BOUNDS, GOLDEN_CASES, IMAGES, compare_lk, golden_case, golden_json, load_gray, sha


def test_pyramids_match_cv2_golden_hashes():
    g = golden_json("pyramid_hashes.json")["images"]
    for name in IMAGES:
        im = load_gray(name)
        lv, dv = oracle.build_lk_pyramid(im, (21, 21), 3, True)
        assert [sha(a) for a in lv] == g[name]["gauss"], name
        assert [sha(a) for a in dv] == g[name]["scharr"], name


def test_survey_known_answers():
    # sha256 prefixes frozen at survey time (kitti0 + odd-width KITTI/000000)
    im = load_gray("kitti0.png")
    lv, dv = oracle.build_lk_pyramid(im, (21, 21), 3, True)
    assert [sha(a)[:16] for a in lv] == ["74d8a6e48432626b", "272aab6ee0ca620a", "efe3073117dfcdef", "6f401f3f92345363"]
    assert [sha(a)[:16] for a in dv] == ["37a8b631234b6e1a", "85dbf32ff48495cd", "9f694d439613368d", "3303b37025493205"]
    box = oracle.box_pyramid(im, 3)
    assert [b.shape for b in box] == [(376, 1240), (188, 620), (94, 310)]
    assert [sha(b)[:16] for b in box[1:]] == ["8befb7026fb4f64b", "6a236274936f5ccc"]
    assert [int(b.sum()) for b in box[1:]] == [10339525, 2574585]
    k1 = oracle.box_pyramid(load_gray("kitti1.png"), 3)
    assert [sha(b)[:16] for b in k1[1:]] == ["2cfaba3deb13b3be", "7864e59236789993"]
    odd = oracle.box_pyramid(load_gray("kitti_000000.png"), 3)  # 1241 wide: the sheared pointer walk
    assert [sha(b)[:16] for b in odd[1:]] == ["31c9d41ecba6551a", "0d90d380417b6584"]


def test_box_modes_and_unsupported_shapes():
    rng = np.random.default_rng(1)
    im = rng.integers(0, 256, (64, 96), dtype=np.uint8)  # 96 % 16 == 0 -> SSE2 rounding on x86
    a = im.astype(np.int32)
    tr = ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]) // 4).astype(np.uint8)
    v0, v1 = (a[0::2, 0::2] + a[1::2, 0::2] + 1) >> 1, (a[0::2, 1::2] + a[1::2, 1::2] + 1) >> 1
    ss = ((v0 + v1 + 1) >> 1).astype(np.uint8)
    assert np.array_equal(oracle.box_half(im, oracle.BOX_TRUNC), tr)
    assert np.array_equal(oracle.box_half(im, oracle.BOX_SSE2), ss)
    assert np.array_equal(oracle.box_half(im, oracle.BOX_AUTO_X86), ss)
    # scalar pointer walk (src/utils.cpp:401-418) emulated literally, incl. odd widths and an ROI whose step != cols
    for (h, w, stride) in [(64, 90, 90), (64, 91, 91), (376, 1241, 1241), (64, 90, 92), (10, 7, 7)]:
        buf = rng.integers(0, 256, (h, stride), dtype=np.uint8)
        flat, out, top, bottom, end = buf.reshape(-1).astype(np.int32), [], 0, stride, stride * h
        while bottom < end:
            for _ in range(w // 2):
                out.append((flat[top] + flat[top + 1] + flat[bottom] + flat[bottom + 1]) // 4)
                top += 2; bottom += 2
            top += stride; bottom += stride
        exp = np.array(out, np.uint8).reshape(h // 2, w // 2)
        assert np.array_equal(oracle.box_half(buf[:, :w], oracle.BOX_AUTO_X86), exp), (h, w, stride)
    with pytest.raises(ValueError):
        oracle.box_half(im[:, :90], oracle.BOX_AUTO_X86)  # step 96 != cols 90: the reference's walk overruns the output
    with pytest.raises(ValueError):
        oracle.box_half(np.zeros((375, 501), np.uint8))  # odd x odd: the reference writes one row too many
    with pytest.raises(ValueError):
        oracle.box_half(np.zeros((64, 90), np.uint8), oracle.BOX_SSE2)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_lk_oracle_vs_cv2_golden(case):
    g = golden_case(case)
    a, b = load_gray(g["prev"]), load_gray(g["next"])
    p, s, e, tr = oracle.calc_optical_flow_pyr_lk(a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"],
                                                  g["flags"], trace=True)
    max_count = min(max(g["crit"][1], 0), 100) if g["crit"][0] & 1 else 30
    converged = tr["iters"][:, 0] < max_count  # terminated by eps / oscillation at level 0, not by the iteration cap
    m = compare_lk(p, s, e, g["next_pts"], g["status"], g["err"], converged)
    # north_star gates: status agreement >= 99.9 %, |dpos| <= 0.01 px on jointly converged points
    assert m["status_agree"] >= 0.999, m
    assert m["max_dpos_converged"] <= 0.01, m
    # ... and over ALL jointly tracked points, iteration-capped ones included: explicit count and size of the outliers
    n_over, max_d = ALL_TRACKED_BOUNDS[case]
    print("%s: %d jointly tracked, %d over 0.01 px (max %.4f px); converged subset: %d over, max %.4f px" % (
        case, m["n_both"], m["n_over_0p01_tracked"], m["max_dpos_tracked"], m["n_over_0p01_converged"], m["max_dpos_converged"]))
    assert m["n_over_0p01_converged"] == 0, m
    assert m["n_over_0p01_tracked"] <= n_over and m["max_dpos_tracked"] <= max_d, m
    assert m["frac_within_0p01"] >= 0.997, m
    if g["flags"] & 8:
        assert m["max_derr"] <= 1e-5, m
    else:
        assert m["max_derr"] <= 0.05, m


@pytest.mark.skipif(not cv2_ref.HAVE_CV2, reason="cv2 not importable")
def test_oracle_vs_cv2_live_pyramids_and_edge_semantics():
    rng = np.random.default_rng(7)
    for (h, w) in [(47, 155), (64, 64), (33, 70), (240, 135)]:
        im = rng.integers(0, 256, (h, w), dtype=np.uint8)
        lv, dv = oracle.build_lk_pyramid(im, (5, 5), 3, True)
        lc, dc = cv2_ref.build_lk_pyramid(im, (5, 5), 3, True)
        assert len(lv) == len(lc)
        assert all(np.array_equal(x, y) for x, y in zip(lv, lc)) and all(np.array_equal(x, y) for x, y in zip(dv, dc))
    a, b = load_gray("kitti0.png"), load_gray("kitti2.png")
    pts = np.concatenate([cv2_ref.fast_corners(a)[0][::12], np.stack([rng.uniform(-30, 1270, 60), rng.uniform(-30, 400, 60)], 1)]).astype(np.float32)
    for crit in [(2, 0, 0.03), (1, 5, 0.0), (3, 0, 0.01), (3, 200, 50.0)]:
        p1, s1, e1 = cv2_ref.calc_optical_flow_pyr_lk(a, b, pts, None, (21, 21), 3, crit, 0)
        p2, s2, e2 = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, (21, 21), 3, crit, 0)
        m = compare_lk(p1, s1, np.where(s1 == 1, e1, 0), p2, s2, e2)
        assert m["status_agree"] == 1.0 and m["frac_within_0p01"] >= 0.99, (crit, m)


def test_undistort_points_vs_cv2_golden():
    """oracle/postfilter.undistort_points (the distorted branch of Pinhole::cam2world, src/camera.cpp:32-40) against
    cv2.undistortPoints outputs frozen by tests/golden/make_golden.py -- bit-exact -- and live when cv2 is importable"""
    import os
    from oracle import postfilter
    from _common import GOLDEN
    z = np.load(os.path.join(GOLDEN, "undistort.npz"))
    for ci in range(3):
        cam, uv, xy = z["cam%d" % ci], z["uv%d" % ci], z["xy%d" % ci]
        o = postfilter.undistort_points(uv, cam[0], cam[1], cam[2], cam[3], tuple(cam[4:9]))
        assert np.array_equal(o.view(np.uint32), xy.view(np.uint32)), ci
        if cv2_ref.HAVE_CV2:
            import cv2
            K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
            live = cv2.undistortPoints(uv.reshape(-1, 1, 2), K, cam[4:9].astype(np.float32).reshape(1, 5)).reshape(-1, 2)
            assert np.array_equal(live.view(np.uint32), xy.view(np.uint32)), ci


def test_lk_oracle_argument_errors_and_empty():
    a = load_gray("kitti0.png")
    with pytest.raises(ValueError):
        oracle.calc_optical_flow_pyr_lk(a, a, np.zeros((1, 2), np.float32), win=(2, 2))
    with pytest.raises(ValueError):
        oracle.calc_optical_flow_pyr_lk(a, a, np.zeros((1, 2), np.float32), max_level=-1)
    p, s, e = oracle.calc_optical_flow_pyr_lk(a, a, np.zeros((0, 2), np.float32))
    assert p.shape == (0, 2) and s.shape == (0,)
    # identical frames: zero flow, everything with texture is tracked, err == 0
    pts = np.array([[100.5, 100.25], [620, 188], [-100, -100]], np.float32)
    p, s, e = oracle.calc_optical_flow_pyr_lk(a, a, pts)
    assert np.allclose(p[:2], pts[:2], atol=1e-3) and s[2] == 0 and e[2] == 0
