"""The GPU corner detector against code that is not the oracle.

The FAST library the reference links (uzh-rpg `fast`) is absent, so the detector kernels are pinned the other way round:
switched to arc length 9 (dr3lk_debug_set_fast_arc) they must reproduce, feature for feature,
    OpenCV's FAST-9 with non-maximum suppression   (corner test + score + 3x3 non-max), run on
    the reference's own reduce_to_half pyramid       (oracle/_ref, src/utils.cpp:323-430), ranked by
    the reference's own shi_tomasi_score             (oracle/_ref, src/utils.cpp:282-321), through
    the grid rule of FastDetector::detect            (src/features.cpp:75-95, ten lines of numpy in tests/_common.py).
Arc length 10 -- what the reference runs -- differs from this path by one template constant in fast.cu and is checked
against the restatement in tests/test_gpu_next_rows.py."""
import numpy as np
import pytest

from oracle import cv2_ref, ref
from _common import load_gray, expected_from_cv2_and_reference

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not cv2_ref.HAVE_CV2, reason="cv2 not importable"),
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


@pytest.mark.parametrize("name", ["kitti0.png", "kitti1.png", "kitti_000000.png", "sample_gray_500x375.png"])
def test_detector_at_arc9_equals_cv2_fast9_on_reference_pyramid(dr3, name):
    im = load_gray(name)
    im = np.ascontiguousarray(im[: im.shape[0] // 4 * 4, : im.shape[1] // 4 * 4])
    with dr3.Context(0) as c:
        c.debug_set_fast_arc(9)
        xy, lv, sc = c.fast_detect(im, 3, 30, 20, 20.0)
        exy, elv, esc = expected_from_cv2_and_reference(im, 3, 30, 20, 20.0)
        assert len(exy) > 150 and len(set(elv.tolist())) >= 2
        assert np.array_equal(xy, exy) and np.array_equal(lv, elv) and np.array_equal(sc.view(np.uint32), esc.view(np.uint32))
        # and the hook really switches the segment test: arc 10 finds a different (sparser-cornered) set
        c.debug_set_fast_arc(10)
        xy10, _, _ = c.fast_detect(im, 3, 30, 20, 20.0)
        assert not (len(xy10) == len(xy) and np.array_equal(xy10, xy))


def test_detector_at_arc9_other_parameters_and_occupancy(dr3):
    im = load_gray("kitti0.png")
    im = np.ascontiguousarray(im[:368, :1232])
    rng = np.random.default_rng(5)
    with dr3.Context(0) as c:
        c.debug_set_fast_arc(9)
        for nl, cell, thr, det in [(1, 30, 20, 20.0), (4, 25, 10, 5.0), (2, 40, 35, 50.0)]:
            got = c.fast_detect(im, nl, cell, thr, det)
            exp = expected_from_cv2_and_reference(im, nl, cell, thr, det)
            assert len(exp[0]) > 50
            for g, e in zip(got, exp):
                assert np.array_equal(g, e)
        occ = (rng.random((-(-368 // 30)) * (-(-1232 // 30))) < 0.5).astype(np.uint8)
        got = c.fast_detect(im, 3, 30, 20, 20.0, occ)
        exp = expected_from_cv2_and_reference(im, 3, 30, 20, 20.0, occ)
        assert 50 < len(exp[0]) and all(np.array_equal(g, e) for g, e in zip(got, exp))
        with pytest.raises(dr3.Dr3lkError):
            c.debug_set_fast_arc(12)
