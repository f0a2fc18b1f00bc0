"""The CPU restatements (oracle/lk_oracle.c, oracle/fast_oracle.c, oracle/postfilter.py) against oracle/_ref = the
reference's own functions compiled unmodified from /root/reference (oracle/build_ref.sh):
src/utils.cpp:282-430, src/initialization.cpp:171-249, src/camera.cpp:25-41.  Bit-exact everywhere."""
import numpy as np
import pytest

import oracle
from oracle import postfilter, ref
from _common import load_gray

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref3dr.so not built (needs /root/reference)")


def _rand_img(rng, h, w):
    # smooth + noise so that 2x2 means hit every rounding case
    return rng.integers(0, 256, (h, w), dtype=np.uint8)


@pytest.mark.parametrize("name", ["kitti0.png", "kitti1.png", "kitti_000000.png"])
def test_box_pyramid_fixtures(name):
    """1240 wide (scalar walk), 1241 wide (the sheared odd-width walk, src/utils.cpp:410-417)"""
    img = load_gray(name)
    a, b = oracle.box_pyramid(img, 3), ref.box_pyramid(img, 3)
    assert [x.shape for x in a] == [x.shape for x in b]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("shape,levels", [((376, 1241), 3), ((376, 1240), 3), ((64, 96), 3), ((270, 960), 3), ((2160, 3840), 3),
                                          ((100, 90), 3), ((52, 48), 2), ((50, 35), 2), ((375, 500), 3), ((33, 64), 4), ((128, 256), 5)])
def test_box_pyramid_random(shape, levels):
    """cols % 16 == 0 -> halfSampleSSE2 rounding (src/utils.cpp:337-341) incl. levels that leave it again (960 -> 480 -> 240),
    everything else -> truncating scalar walk; odd heights; odd widths at level 0 and at deeper levels"""
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    img = _rand_img(rng, *shape)
    try:
        b = ref.box_pyramid(img, levels)
    except ValueError:
        # the reference's walk would leave its buffers: the restatement must refuse the same shape
        with pytest.raises(ValueError):
            oracle.box_pyramid(img, levels)
        return
    a = oracle.box_pyramid(img, levels)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_box_pyramid_unsupported_shapes_agree():
    """odd x odd: the reference writes one row too many -> both sides refuse"""
    img = np.zeros((35, 51), np.uint8)
    with pytest.raises(ValueError):
        ref.box_pyramid(img, 2)
    with pytest.raises(ValueError):
        oracle.box_pyramid(img, 2)


def test_shi_tomasi_scores():
    for name in ["kitti0.png", "sample_gray_500x375.png"]:
        img = load_gray(name)
        h, w = img.shape
        rng = np.random.default_rng(5)
        uv = np.stack([rng.integers(-2, w + 2, 20000), rng.integers(-2, h + 2, 20000)], 1).astype(np.int32)
        uv = np.concatenate([uv, [[4, 4], [5, 5], [w - 6, h - 6], [w - 5, h - 5], [5, h - 6], [w - 6, 5]]]).astype(np.int32)
        r = ref.shi_tomasi(img, uv)
        o = np.array([oracle.shi_tomasi(img, u, v) for u, v in uv], np.float32)
        assert np.array_equal(r.view(np.uint32), o.view(np.uint32))
        assert (r > 0).sum() > 1000


def _hypotheses(rng, n_hyp):
    F = rng.normal(0, 1, (n_hyp, 9)).astype(np.float32)
    F[:, [0, 1, 3, 4]] *= 1e-5
    F[:, [2, 5, 6, 7]] *= 1e-2
    return F


def test_check_fundamental():
    rng = np.random.default_rng(11)
    n = 1500
    p1 = np.stack([rng.uniform(0, 1240, n), rng.uniform(0, 376, n)], 1).astype(np.float32)
    p2 = (p1 + rng.normal(0, 2, p1.shape)).astype(np.float32)
    F = _hypotheses(rng, 24)
    # one near-true hypothesis (pure x translation: F = [t]_x) so that the inlier branch is exercised heavily
    F[0] = np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32)
    p2[:, 1] = p1[:, 1] + rng.normal(0, 0.7, n).astype(np.float32)
    for sigma in (1.0, 2.0):
        sr, ir = ref.check_fundamental(F, p1, p2, sigma)
        so, io = postfilter.check_fundamental(F, p1, p2, sigma)
        assert np.array_equal(sr.view(np.uint32), so.view(np.uint32))
        assert np.array_equal(ir, io)
    assert ir[0].sum() > 100


def test_check_fundamental_degenerate():
    """zero / non-finite lines: 0/0 -> NaN chi2, `chi2 > th` false -> the reference ADDS thScore - NaN; restated as is"""
    p1 = np.array([[0, 0], [10, 20], [3, 4]], np.float32)
    p2 = np.array([[0, 0], [11, 19], [3, 4]], np.float32)
    F = np.zeros((2, 9), np.float32)
    F[1, 8] = 1.0
    sr, ir = ref.check_fundamental(F, p1, p2)
    so, io = postfilter.check_fundamental(F, p1, p2)
    assert np.array_equal(sr.view(np.uint32), so.view(np.uint32)) or (np.isnan(sr) == np.isnan(so)).all()
    assert np.array_equal(ir, io)


def test_cam2world_undistorted():
    rng = np.random.default_rng(3)
    uv = np.stack([rng.uniform(-50, 1300, 5000), rng.uniform(-50, 420, 5000)], 1).astype(np.float32)
    fx, fy, cx, cy = 718.856, 718.856, 607.1928, 185.2157
    r = ref.cam2world(uv, fx, fy, cx, cy)
    st = np.ones(len(uv), np.uint8)
    _, _, _, b = postfilter.filter_tracks(uv, uv, st, fx, fy, cx, cy)
    assert np.array_equal(r.view(np.uint64), b.view(np.uint64))


def test_cam2world_distorted():
    rng = np.random.default_rng(4)
    uv = np.stack([rng.uniform(0, 752, 3000), rng.uniform(0, 480, 3000)], 1).astype(np.float32)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    dist = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
    r = ref.cam2world(uv, fx, fy, cx, cy, dist)
    st = np.ones(len(uv), np.uint8)
    _, _, _, b = postfilter.filter_tracks(uv, uv, st, fx, fy, cx, cy, dist)
    assert np.array_equal(r.view(np.uint64), b.view(np.uint64))
    # and the distortion does something
    u = ref.cam2world(uv, fx, fy, cx, cy)
    assert np.abs(u - r).max() > 1e-3
