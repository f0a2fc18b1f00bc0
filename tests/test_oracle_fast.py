"""CPU tests of the FAST-10 / grid-selection restatement (oracle/fast_oracle.c).
The `fast` library the reference links is absent; what can be pinned is checked here: corner test, score and non-max
against cv2's FAST-9 by running the same code with arc length 9 (identical key points and responses), the Shi-Tomasi half
against the reference's own function (tests/test_oracle_vs_ref.py), and internal consistency of score / non-max / grid."""
import numpy as np
import pytest

import oracle
from oracle import cv2_ref
from _common import load_gray


@pytest.mark.skipif(not cv2_ref.HAVE_CV2, reason="cv2 not importable")
@pytest.mark.parametrize("name", ["kitti0.png", "kitti_000000.png", "sample_gray_500x375.png"])
def test_corner_test_matches_cv2_fast9(name):
    im = load_gray(name)
    det = cv2_ref.cv2.FastFeatureDetector_create(20, False, cv2_ref.cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    cvp = sorted((int(k.pt[1]), int(k.pt[0])) for k in det.detect(im, None))
    mine = sorted((int(y), int(x)) for x, y in oracle.fast_detect(im, 20, 9))
    assert mine == cvp and len(mine) > 1000


@pytest.mark.skipif(not cv2_ref.HAVE_CV2, reason="cv2 not importable")
@pytest.mark.parametrize("name", ["kitti0.png", "kitti_000000.png", "sample_gray_500x375.png"])
def test_score_and_nonmax_match_cv2_fast9(name):
    """The whole FAST machinery of the restatement -- arc test, score (largest threshold that keeps the corner) and the 3x3
    non-maximum suppression -- against OpenCV's FAST with suppression on, at arc length 9: identical key-point sets and
    identical responses.  The reference's detector (uzh-rpg `fast`, absent here) is the same algorithm at arc length 10:
    the one constant by which the pinned code path and the used one differ."""
    im = load_gray(name)
    det = cv2_ref.cv2.FastFeatureDetector_create(20, True, cv2_ref.cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    cvd = {(int(k.pt[0]), int(k.pt[1])): int(k.response) for k in det.detect(im, None)}
    xy = oracle.fast_detect(im, 20, 9)
    sc = oracle.fast_score(im, xy, 20, 9)
    keep = oracle.fast_nonmax(xy, sc)
    mine = {(int(xy[i, 0]), int(xy[i, 1])): int(sc[i]) for i in keep}
    assert len(mine) > 4000 and mine == cvd


def test_score_is_largest_threshold_and_nonmax_is_8_neighbour():
    im = load_gray("kitti1.png")
    xy = oracle.fast_detect(im, 20, 10)
    sc = oracle.fast_score(im, xy, 20, 10)
    assert len(xy) > 5000 and sc.min() >= 20 and sc.max() <= 254
    # raising the threshold to score keeps the corner, score + 1 drops it
    for i in np.random.default_rng(0).choice(len(xy), 40, replace=False):
        x, y, s = int(xy[i, 0]), int(xy[i, 1]), int(sc[i])
        patch = np.ascontiguousarray(im[y - 3:y + 4, x - 3:x + 4])
        assert len(oracle.fast_detect(patch, s, 10)) == 1 and len(oracle.fast_detect(patch, s + 1, 10)) == 0
    keep = oracle.fast_nonmax(xy, sc)
    smap = np.zeros(im.shape, np.int32)
    smap[xy[:, 1], xy[:, 0]] = sc
    pad = np.pad(smap, 1)
    nb = np.max([pad[1 + dy:1 + dy + im.shape[0], 1 + dx:1 + dx + im.shape[1]] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (0, 0)], axis=0)
    exp = [i for i in range(len(xy)) if nb[xy[i, 1], xy[i, 0]] < sc[i]]
    assert list(keep) == exp


def test_detector_grid_properties():
    im = load_gray("kitti0.png")
    xy, lv, sc = oracle.fast_detector(im, 3, 30, 20, 20.0)
    assert 200 < len(xy) <= 42 * 13 and (sc > 20.0).all() and set(np.unique(lv)) <= {0, 1, 2}
    cells = (xy[:, 1] // 30) * 42 + xy[:, 0] // 30
    assert (np.diff(cells) > 0).all()                         # one feature per cell, in cell order
    assert ((xy % (1 << lv)[:, None]) == 0).all()             # level-l corners sit on the 2^l lattice
    # occupied cells are skipped
    occ = np.zeros(42 * 13, np.uint8)
    occ[cells[::2]] = 1
    xy2, _, _ = oracle.fast_detector(im, 3, 30, 20, 20.0, occupancy=occ)
    cells2 = (xy2[:, 1] // 30) * 42 + xy2[:, 0] // 30
    assert set(cells2) == set(cells[1::2])


@pytest.mark.skipif(not cv2_ref.HAVE_CV2, reason="cv2 not importable")
@pytest.mark.parametrize("name", ["kitti0.png", "sample_gray_500x375.png"])
def test_whole_detector_at_arc9_equals_cv2_plus_reference_functions(name):
    """FastDetector::detect end to end with nothing of the restatement on the expected side: OpenCV FAST-9 key points on the
    reference's own reduce_to_half levels, ranked by the reference's own shi_tomasi_score, through the grid rule (numpy)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from _common import expected_from_cv2_and_reference
    im = load_gray(name)
    im = np.ascontiguousarray(im[: im.shape[0] // 4 * 4, : im.shape[1] // 4 * 4])
    for nl, cell, thr, det in [(3, 30, 20, 20.0), (4, 25, 10, 5.0)]:
        if nl == 4:
            im = np.ascontiguousarray(im[: im.shape[0] // 8 * 8, : im.shape[1] // 8 * 8])
        got = oracle.fast_detector(im, nl, cell, thr, det, arc=9)
        exp = expected_from_cv2_and_reference(im, nl, cell, thr, det)
        assert len(exp[0]) > 100
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        assert np.array_equal(got[2].view(np.uint32), exp[2].view(np.uint32))
