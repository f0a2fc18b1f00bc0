"""GPU against oracle/_ref = the reference's OWN functions compiled unmodified from /root/reference (oracle/build_ref.sh):
box pyramid (src/utils.cpp:324-430), Shi-Tomasi scores of the detector's features (src/utils.cpp:282-321),
CheckFundamental (src/initialization.cpp:171-249), cam2world bearings incl. the distorted pinhole (src/camera.cpp:25-41).
Bit-exact everywhere.  The library travels with the snapshot; /root/reference is not read here."""
import numpy as np
import pytest

from oracle import ref
from _common import golden_case, load_gray, random_points

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref3dr.so did not travel with the snapshot")]


@pytest.mark.parametrize("name", ["kitti0.png", "kitti1.png", "kitti_000000.png"])
def test_box_pyramid_fixtures(ctx, name):
    img = load_gray(name)  # 1240 wide: scalar walk; 1241 wide: the sheared odd-width walk
    got, exp = ctx.box_pyramid(img, 3), ref.box_pyramid(img, 3)
    assert all(np.array_equal(a, b) for a, b in zip(got[1:], exp[1:]))


@pytest.mark.parametrize("shape,levels", [((376, 1241), 3), ((64, 96), 3), ((270, 960), 3), ((2160, 3840), 3), ((100, 90), 3),
                                          ((52, 48), 2), ((50, 35), 2), ((375, 500), 3), ((33, 64), 4), ((128, 256), 5)])
def test_box_pyramid_shapes(ctx, dr3, shape, levels):
    """SSE2 rounding (cols % 16 == 0, incl. 3840 = C4 and levels that leave the SSE2 path again), truncating walk, odd sizes"""
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    try:
        exp = ref.box_pyramid(img, levels)
    except ValueError:
        with pytest.raises(dr3.Dr3lkError):  # the reference would leave its buffers: the product refuses the shape
            ctx.box_pyramid(img, levels)
        return
    got = ctx.box_pyramid(img, levels)
    assert all(np.array_equal(a, b) for a, b in zip(got[1:], exp[1:]))


@pytest.mark.parametrize("name", ["kitti0.png", "kitti5.png", "kitti_000000.png", "sample_gray_500x375.png"])
def test_detector_scores_are_the_reference_shi_tomasi(ctx, name):
    """every feature FastDetector::detect keeps carries utils::shi_tomasi_score of its pyramid level at its corner"""
    img = load_gray(name)
    xy, lv, sc = ctx.fast_detect(img, 3, 30, 20, 20.0)
    pyr = ref.box_pyramid(img, 3)
    assert len(xy) > 100
    for level in range(3):
        m = lv == level
        if not m.any():
            continue
        uv = (xy[m] >> level).astype(np.int32)  # Feature.px = corner * 2^level (src/features.cpp:83)
        assert np.array_equal(uv << level, xy[m])
        r = ref.shi_tomasi(pyr[level], uv)
        assert np.array_equal(r.view(np.uint32), sc[m].view(np.uint32)), (name, level)


def test_score_fundamental(ctx):
    rng = np.random.default_rng(33)
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = golden_case("c1_default_21x21")["prev_pts"][:4607]
    p, s, _ = ctx.calc_optical_flow_pyr_lk(a, b, pts)
    p1, p2 = pts[s == 1], p[s == 1]
    F0 = np.array([[0, -1e-6, 2e-4], [1e-6, 0, -3e-3], [-2e-4, 3e-3, 0]], np.float32)
    F = (F0[None] * (1 + 0.3 * rng.standard_normal((200, 3, 3)))).astype(np.float32)
    F[11] = np.eye(3, dtype=np.float32)
    for sigma in (1.0, 1.5):
        sc, inl, best = ctx.score_fundamental(F, p1, p2, sigma)
        rs, ri = ref.check_fundamental(F, p1, p2, sigma)
        assert np.array_equal(sc.view(np.uint32), rs.view(np.uint32)) and np.array_equal(inl, ri)
    assert inl.sum() > 1000 and best == int(np.argmax(rs))


CAMS = [(718.856, 718.856, 607.1928, 185.2157, None, 1241, 376),
        (458.654, 457.296, 367.215, 248.375, (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0), 752, 480),
        (718.856, 718.856, 607.1928, 185.2157, (-0.3, 0.1, 0.001, -0.002, 0.05), 1241, 376),
        (300.0, 300.0, 320.0, 240.0, (0.9, 2.5, 0.01, 0.01, 1.0), 640, 480),       # reaches undistortPoints' icdist < 0 exit
        (500.0, 500.0, 320.0, 240.0, (1e-8, 0.5, 0.0, 0.0, 0.0), 640, 480)]        # |d0| <= 1e-7: Pinhole::_distortion stays false


@pytest.mark.parametrize("cam", CAMS)
def test_bearings_are_the_reference_cam2world(ctx, cam):
    fx, fy, cx, cy, dist, w, h = cam
    rng = np.random.default_rng(int(fx))
    n = 5000
    refp = random_points(rng, w, h, n, margin=60)
    cur = (refp + rng.normal(0, 3, refp.shape)).astype(np.float32)
    st = (rng.random(n) < 0.9).astype(np.uint8)
    r, c, d, bear = ctx.filter_tracks(refp, cur, st, fx, fy, cx, cy, dist)
    assert np.array_equal(c, cur[st == 1]) and np.array_equal(r, refp[st == 1])
    exp = ref.cam2world(c, fx, fy, cx, cy, dist if dist is not None else (0, 0, 0, 0, 0))
    assert np.array_equal(bear.view(np.uint64), exp.view(np.uint64))
