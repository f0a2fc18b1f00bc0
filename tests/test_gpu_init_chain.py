"""The two-frame initialiser front end in three device-resident calls (dr3lk_init_first_frame / _second_frame /
_score_fundamental; reference src/initialization.cpp:546-661): bit-identical to the four separate host-buffer calls (FAST
detect, pyramid + LK, status filter / disparity / bearing, hypothesis scoring) and to the CPU restatements, with one upload
and one download per frame instead of four."""
import numpy as np
import pytest

import oracle
from oracle import postfilter
from _common import load_gray

pytestmark = pytest.mark.gpu

CAM = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157)
REF_LK = dict(max_level=4, criteria=(3, 1000, 1e-3), flags=4)   # src/initialization.cpp:593-613


def _eq(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("dist", [None, (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)])
def test_init_chain_matches_separate_calls_and_oracle(ctx, dr3, dist):
    frames = [load_gray("kitti%d.png" % i) for i in range(4)]
    # first frame: corners + LK pyramid from one upload
    xy, lv, sc, ref_pyr = ctx.init_first_frame(frames[0])
    exy, elv, esc = ctx.fast_detect(frames[0])
    assert len(xy) > 100 and _eq(xy, exy) and _eq(lv, elv) and _eq(sc, esc)
    oxy, olv, osc = oracle.fast_detector(frames[0])
    assert _eq(xy, oxy) and _eq(sc, osc)
    kps_ref = xy.astype(np.float32)      # Feature.px of every corner (src/initialization.cpp:567-576)
    kps_cur = kps_ref.copy()             # line 578
    launches0 = ctx.launch_count
    for k in (1, 2, 3):                  # the handler stays in SECOND_FRAME on failure: same reference, warm start (src/handler.cpp:67-72)
        out = ctx.init_second_frame(ref_pyr, frames[k], kps_ref, kps_cur, dist=dist, **CAM, **REF_LK)
        # the separate calls
        p, s, e = ctx.calc_optical_flow_pyr_lk(frames[0], frames[k], kps_ref, kps_cur, (30, 30), 4, (3, 1000, 1e-3), dr3.USE_INITIAL_FLOW)
        r2, c2, d2, b2 = ctx.filter_tracks(kps_ref, p, s, dist=dist, **CAM)
        assert _eq(out["status"], s) and _eq(out["err"][s == 1], e[s == 1])
        assert _eq(out["ref"], r2) and _eq(out["cur"], c2) and _eq(out["disparity"], d2) and _eq(out["bearing"], b2)
        # the CPU restatements
        po, so, eo = oracle.calc_optical_flow_pyr_lk(frames[0], frames[k], kps_ref, kps_cur, (30, 30), 4, (3, 1000, 1e-3), 4)
        ro, co, do, bo = postfilter.filter_tracks(kps_ref, po, so, dist=dist, **CAM)
        assert _eq(out["ref"], ro) and _eq(out["cur"], co) and _eq(out["disparity"], do) and _eq(out["bearing"], bo)
        # hypothesis scoring on the resident tracks == scoring of the downloaded ones
        n = len(out["ref"])
        rng = np.random.default_rng(k)
        F0 = np.array([[0, -1e-6, 2e-4], [1e-6, 0, -3e-3], [-2e-4, 3e-3, 0]], np.float32)
        F = (F0[None] * (1 + 0.3 * rng.standard_normal((200, 3, 3)))).astype(np.float32)
        s1, i1, b1 = ctx.init_score_fundamental(F, n)
        s0, i0, b0 = ctx.score_fundamental(F, out["ref"], out["cur"])
        assert _eq(s1, s0) and _eq(i1, i0) and b1 == b0
        # erase loop semantics: the survivors are the next call's inputs
        kps_ref, kps_cur = out["ref"].copy(), out["cur"].copy()
        assert n >= 100
    assert ctx.launch_count > launches0
    ref_pyr.close()


def test_init_chain_errors(ctx, dr3):
    a = load_gray("kitti0.png")
    xy, lv, sc, pyr = ctx.init_first_frame(a, win=(21, 21), max_level=3)
    with pytest.raises(dr3.Dr3lkError):
        ctx.init_second_frame(pyr, a[:, :-8], xy.astype(np.float32), xy.astype(np.float32))
    with pytest.raises(dr3.Dr3lkError):   # USE_INITIAL_FLOW without an initial guess
        ctx.init_second_frame(pyr, a, xy.astype(np.float32), None)
    out = ctx.init_second_frame(pyr, a, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert len(out["ref"]) == 0
    with pytest.raises(dr3.Dr3lkError):   # nothing resident after an empty call
        ctx.init_score_fundamental(np.eye(3, dtype=np.float32)[None], 0)
    with pytest.raises(dr3.Dr3lkError):
        ctx.init_first_frame(a, win=(2, 2))
    pyr.close()
