"""GPU parity of the SURVEY.md 8(f) rows built so far: f-2 pyramid caching and f-3 status filter / disparity / bearings."""
import numpy as np
import pytest

import oracle
from oracle import postfilter
from _common import golden_case, load_gray, random_points

pytestmark = pytest.mark.gpu


def test_cached_pyramids_give_identical_results(ctx, dr3):
    frames = [load_gray("kitti%d.png" % i) for i in range(4)]
    pts = golden_case("c1_default_21x21")["prev_pts"][:2000]
    pyrs = [dr3.Pyramid(ctx, f, (21, 21), 3) for f in frames]
    assert all(p.levels == 4 for p in pyrs)
    cur = pts
    for i in range(3):  # frame-to-frame chain: every pyramid is built once and used as next, then as prev
        got = ctx.calc_optical_flow_pyr_lk_cached(pyrs[i], pyrs[i + 1], cur)
        exp = ctx.calc_optical_flow_pyr_lk(frames[i], frames[i + 1], cur)
        orc = oracle.calc_optical_flow_pyr_lk(frames[i], frames[i + 1], cur)
        for a, b, c in zip(got, exp, orc):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)) and np.array_equal(a.view(np.uint8), c.view(np.uint8))
        cur = got[0][got[1] == 1]
    # anchored retries against the same reference frame with warm start (the reference's actual loop)
    init = pts + np.float32(0.75)
    got = ctx.calc_optical_flow_pyr_lk_cached(pyrs[0], pyrs[2], pts, init, 3, (3, 30, 0.01), dr3.USE_INITIAL_FLOW)
    orc = oracle.calc_optical_flow_pyr_lk(frames[0], frames[2], pts, init, (21, 21), 3, (3, 30, 0.01), dr3.USE_INITIAL_FLOW)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, orc))
    # fewer levels than the pyramids hold, and mismatched windows / sizes are rejected like OpenCV does
    got = ctx.calc_optical_flow_pyr_lk_cached(pyrs[0], pyrs[1], pts, max_level=1)
    orc = oracle.calc_optical_flow_pyr_lk(frames[0], frames[1], pts, None, (21, 21), 1)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, orc))
    other = dr3.Pyramid(ctx, frames[0][:, :-8], (21, 21), 3)
    with pytest.raises(dr3.Dr3lkError):
        ctx.calc_optical_flow_pyr_lk_cached(pyrs[0], other, pts)
    for p in pyrs + [other]:
        p.close()


def test_track_frame_streaming_matches_plain_calls(ctx, dr3):
    """dr3lk_track_frame: one call per new frame against a frame that is already on the device -- a frame-to-frame chain
    (keep_next=2), the reference's anchored warm-start loop (keep_next=0 / 1) and the 30x30 window; bit-identical to the
    plain two-image call and to the oracle."""
    frames = [load_gray("kitti%d.png" % i) for i in range(5)]
    pts = golden_case("c1_default_21x21")["prev_pts"][:1500]
    prev = dr3.Pyramid(ctx, frames[0], (21, 21), 3)
    cur = pts
    for i in range(4):
        p, s, e, nxt = ctx.track_frame(prev, frames[i + 1], cur, keep_next=2)
        orc = oracle.calc_optical_flow_pyr_lk(frames[i], frames[i + 1], cur)
        for a, c in zip((p, s, e), orc):
            assert np.array_equal(a.view(np.uint8), c.view(np.uint8)), i
        assert nxt.levels == 4
        prev.close()
        prev, cur = nxt, p[s == 1]
    prev.close()
    # anchored: the reference frame stays, every new frame is tracked with a warm start (src/initialization.cpp:608-613)
    anchor = dr3.Pyramid(ctx, frames[0], (30, 30), 4)
    ref, curp = pts.copy(), pts.copy()
    for i in (1, 2, 3):
        p, s, e, keep = ctx.track_frame(anchor, frames[i], ref, curp, 4, (3, 1000, 1e-3), dr3.USE_INITIAL_FLOW, keep_next=i % 2)
        orc = oracle.calc_optical_flow_pyr_lk(frames[0], frames[i], ref, curp, (30, 30), 4, (3, 1000, 1e-3), dr3.USE_INITIAL_FLOW)
        for a, c in zip((p, s, e), orc):
            assert np.array_equal(a.view(np.uint8), c.view(np.uint8)), i
        if keep is not None:
            # a Gaussian-only pyramid can be the next side again, but not the previous side
            q = ctx.calc_optical_flow_pyr_lk_cached(anchor, keep, ref, curp, 4, (3, 1000, 1e-3), dr3.USE_INITIAL_FLOW)
            assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(q, orc))
            with pytest.raises(dr3.Dr3lkError):
                ctx.calc_optical_flow_pyr_lk_cached(keep, anchor, ref)
            keep.close()
        ref, curp = ref[s == 1], p[s == 1]
    # no points: the new frame's pyramid is still built when asked for
    p, s, e, nxt = ctx.track_frame(anchor, frames[4], np.zeros((0, 2), np.float32), keep_next=2)
    assert p.shape == (0, 2) and nxt is not None and nxt.levels == 4
    q = ctx.calc_optical_flow_pyr_lk_cached(nxt, anchor, pts)
    orc = oracle.calc_optical_flow_pyr_lk(frames[4], frames[0], pts, None, (30, 30), 4)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(q, orc))
    with pytest.raises(dr3.Dr3lkError):
        ctx.track_frame(anchor, frames[1][:, :-8], pts)
    nxt.close(); anchor.close()


def test_filter_tracks_matches_restatement(ctx):
    rng = np.random.default_rng(12)
    for n in (1, 7, 1023, 1024, 1025, 4607, 20000):
        ref = random_points(rng, 1240, 376, n)
        cur = (ref + rng.normal(0, 3, ref.shape)).astype(np.float32)
        st = (rng.random(n) < 0.9).astype(np.uint8)
        got = ctx.filter_tracks(ref, cur, st, 718.856, 718.856, 607.1928, 185.2157)  # KITTI sequence-00 intrinsics
        exp = postfilter.filter_tracks(ref, cur, st, 718.856, 718.856, 607.1928, 185.2157)
        assert len(got[0]) == int(st.sum())
        for a, b in zip(got, exp):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), n
        assert np.allclose(np.linalg.norm(got[3], axis=1), 1.0, atol=1e-15)
        dist = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)  # Pinhole with distortion, src/camera.cpp:32-40
        got = ctx.filter_tracks(ref, cur, st, 458.654, 457.296, 367.215, 248.375, dist)
        exp = postfilter.filter_tracks(ref, cur, st, 458.654, 457.296, 367.215, 248.375, dist)
        for a, b in zip(got, exp):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), ("distorted", n)
    r, c, d, b = ctx.filter_tracks(ref, cur, np.zeros(len(ref), np.uint8))
    assert len(r) == 0 and len(d) == 0 and b is None
    r, c, d, b = ctx.filter_tracks(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.uint8))
    assert len(r) == 0


@pytest.mark.parametrize("name", ["kitti0.png", "kitti5.png", "kitti_000000.png", "sample_gray_500x375.png"])
def test_fast_detector_matches_restatement(ctx, name):
    """f-1: FAST-10 + score + 3x3 non-max + per-cell Shi-Tomasi selection on the Frame's box pyramid."""
    im = load_gray(name)
    got = ctx.fast_detect(im, 3, 30, 20, 20.0)
    exp = oracle.fast_detector(im, 3, 30, 20, 20.0)
    assert len(got[0]) == len(exp[0]) > 100
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
    assert np.array_equal(got[2].view(np.uint32), exp[2].view(np.uint32))


def test_fast_detector_parameters_and_occupancy(ctx, dr3):
    from tools import synth
    rng = np.random.default_rng(21)
    a, _, _ = synth.make_pair(3001, 1280, 720)  # 1280 % 16 == 0: the box pyramid takes the SSE2 rounding on x86
    for (nl, cell, thr, det, mode) in [(3, 30, 20, 20.0, 0), (1, 16, 10, 5.0, 0), (4, 40, 35, 50.0, 1), (2, 25, 20, 0.0, 2)]:
        got = ctx.fast_detect(a, nl, cell, thr, det, None, mode)
        exp = oracle.fast_detector(a, nl, cell, thr, det, None, mode)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2]), (nl, cell, thr)
    ncell = (-(-1280 // 30)) * (-(-720 // 30))
    occ = (rng.random(ncell) < 0.5).astype(np.uint8)
    got = ctx.fast_detect(a, 3, 30, 20, 20.0, occ)
    exp = oracle.fast_detector(a, 3, 30, 20, 20.0, occ)
    assert len(got[0]) > 50 and np.array_equal(got[0], exp[0]) and np.array_equal(got[2], exp[2])
    cells = (got[0][:, 1] // 30) * (-(-1280 // 30)) + got[0][:, 0] // 30
    assert not occ[cells].any()
    flat = np.full((100, 120), 9, np.uint8)
    assert len(ctx.fast_detect(flat)[0]) == 0
    with pytest.raises(dr3.Dr3lkError):
        ctx.fast_detect(np.zeros((375, 501), np.uint8))  # odd x odd level: the reference overruns its pyramid buffers
    # the detector's corners feed the LK call, as in Init::process_first_frame -> process_second_frame
    b = load_gray("kitti1.png")
    k0 = load_gray("kitti0.png")
    xy, lv, sc = ctx.fast_detect(k0)
    p, s, e = ctx.calc_optical_flow_pyr_lk(k0, b, xy.astype(np.float32), xy.astype(np.float32), (30, 30), 4, (3, 1000, 1e-3), dr3.USE_INITIAL_FLOW)
    po, so, eo = oracle.calc_optical_flow_pyr_lk(k0, b, xy.astype(np.float32), xy.astype(np.float32), (30, 30), 4, (3, 1000, 1e-3), 4)
    assert s.sum() >= 100 and np.array_equal(s, so) and np.array_equal(p.view(np.uint32), po.view(np.uint32))


def test_score_fundamental_matches_restatement(ctx):
    """f-4: 200 RANSAC hypotheses scored against the tracked matches, bit-identical to the scalar fp32 evaluation."""
    rng = np.random.default_rng(33)
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = golden_case("c1_default_21x21")["prev_pts"][:4607]
    p, s, _ = ctx.calc_optical_flow_pyr_lk(a, b, pts)
    p1, p2 = pts[s == 1][:546], p[s == 1][:546]
    # hypotheses: perturbations of a plausible forward-motion F plus a few degenerate ones
    F0 = np.array([[0, -1e-6, 2e-4], [1e-6, 0, -3e-3], [-2e-4, 3e-3, 0]], np.float32)
    F = (F0[None] * (1 + 0.3 * rng.standard_normal((200, 3, 3)))).astype(np.float32)
    F[7] = 0
    F[11] = np.eye(3, dtype=np.float32)
    sc, inl, best = ctx.score_fundamental(F, p1, p2, 1.0)
    esc, einl = postfilter.check_fundamental(F, p1, p2, 1.0)
    # bit-identical, except that a NaN score (the all-zero hypothesis: 0/0) carries a different payload on the GPU
    nan = np.isnan(esc)
    assert nan[7] and np.array_equal(np.isnan(sc), nan)
    assert np.array_equal(sc[~nan].view(np.uint32), esc[~nan].view(np.uint32)) and np.array_equal(inl, einl)
    exp_best, bs = -1, np.float32(0)
    for i in range(200):
        if esc[i] > bs:
            bs, exp_best = esc[i], i
    assert best == exp_best and np.nanmax(sc) > 0
    sc2, inl2, _ = ctx.score_fundamental(F[:3], p1[:5], p2[:5], 2.0, want_inliers=False)
    assert inl2 is None and np.array_equal(sc2, postfilter.check_fundamental(F[:3], p1[:5], p2[:5], 2.0)[0])


def test_pyramid_outliving_its_context_is_safe(dr3):
    """dr3lk.h lifetime rule: dr3lk_destroy orphans the pyramids that are still alive; destroying them afterwards is legal
    (Python's garbage collector and C++ static destructors do exactly that), using them is an error, not a crash."""
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = random_points(np.random.default_rng(3), 1240, 376, 200)
    c1 = dr3.Context(0)
    p1, p2 = dr3.Pyramid(c1, a), dr3.Pyramid(c1, b)
    exp = c1.calc_optical_flow_pyr_lk_cached(p1, p2, pts)
    p2.close()               # normal order for one of them ...
    c1.close()               # ... the other one outlives its context
    c2 = dr3.Context(0)
    q1, q2 = dr3.Pyramid(c2, a), dr3.Pyramid(c2, b)
    with pytest.raises(dr3.Dr3lkError) as e:
        c2.calc_optical_flow_pyr_lk_cached(p1, q2, pts)   # orphaned pyramid: refused
    assert e.value.code == dr3.E_ARG
    got = c2.calc_optical_flow_pyr_lk_cached(q1, q2, pts)
    assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    p1.close()               # destroy after the context is gone: no use-after-free
    p1.close()               # idempotent in the wrapper
    q1.close(); q2.close(); c2.close()
    for _ in range(20):      # contexts and pyramids created and dropped in every order
        c = dr3.Context(0)
        ps = [dr3.Pyramid(c, a) for _ in range(3)]
        ps[0].close(); c.close(); ps[1].close()
        del ps


def test_set_stream_switch_keeps_results(ctx, dr3):
    """dr3lk_set_stream joins the previous stream before switching: back-to-back calls on alternating streams stay bit-exact"""
    import torch
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = random_points(np.random.default_rng(4), 1240, 376, 1500)
    exp = ctx.calc_optical_flow_pyr_lk(a, b, pts)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    try:
        for i in range(12):
            ctx.set_stream((s1 if i % 2 else s2).cuda_stream)
            got = ctx.calc_optical_flow_pyr_lk(a, b, pts) if i % 3 else ctx.calc_optical_flow_pyr_lk(b, a, pts)
            if i % 3:
                assert all(np.array_equal(x, y) for x, y in zip(got, exp)), i
    finally:
        ctx.set_stream(None)
    assert all(np.array_equal(x, y) for x, y in zip(ctx.calc_optical_flow_pyr_lk(a, b, pts), exp))


def test_pooled_pyramid_buffers_changing_roles_stay_correct(ctx, dr3):
    """Destroyed pyramids hand their device buffers to a pool; a derivative buffer whose zero aprons were cleared once is not
    cleared again when it is reused for the same layout -- but it must be when it served as an IMAGE buffer in between (the
    image of a 2x larger frame is about as large as the derivatives of the small one).  Windows hanging over every border read
    the aprons, so stale apron contents show up as differences against the oracle."""
    import cv2
    small_a, small_b = load_gray("kitti0.png"), load_gray("kitti1.png")
    big_a = cv2.resize(small_a, None, fx=2, fy=2, interpolation=cv2.INTER_LINEAR)
    big_b = cv2.resize(small_b, None, fx=2, fy=2, interpolation=cv2.INTER_LINEAR)
    rng = np.random.default_rng(17)

    def border_points(img, n=700):
        h, w = img.shape
        p = random_points(rng, w, h, n, margin=25)
        edge = np.array([[x, y] for x in (-9.5, 0.0, 3.3, w - 4.2, w - 1.0, w + 8.0) for y in (-9.0, 0.0, 2.7, h - 3.1, h - 1.0, h + 7.5)], np.float32)
        return np.concatenate([p, edge]).astype(np.float32)

    ps, pb = border_points(small_a), border_points(big_a)
    exp_s = oracle.calc_optical_flow_pyr_lk(small_a, small_b, ps)
    exp_b = oracle.calc_optical_flow_pyr_lk(big_a, big_b, pb)
    for rep in range(4):
        # small frame with derivatives, then destroyed: image + derivative buffers go to the pool
        p1, p2 = dr3.Pyramid(ctx, small_a), dr3.Pyramid(ctx, small_b)
        got = ctx.calc_optical_flow_pyr_lk_cached(p1, p2, ps)
        assert all(np.array_equal(x, y) for x, y in zip(got, exp_s)), ("small", rep)
        p1.close(); p2.close()
        # big frames: their Gaussian-only pyramids (keep_next = 1) are about the size of the small frame's derivatives
        q1 = dr3.Pyramid(ctx, big_a)
        np_, st, er, q2 = ctx.track_frame(q1, big_b, pb, keep_next=1)
        assert all(np.array_equal(x, y) for x, y in zip((np_, st, er), exp_b)), ("big", rep)
        q1.close(); q2.close()


def test_page_locked_images_skip_the_staging_copy_and_change_nothing(ctx, dr3):
    """Images in page-locked memory (dr3lk_host_alloc) whose rows are continuous or sit at the aligned device pitch (width
    rounded up to 16) are handed to the copy engine directly instead of being packed into the context's mirror; pinned images
    at any other step are packed like pageable ones.  Whole frames and a strided ROI inside a pinned frame, through the two-image call,
    dr3lk_pyramid_create and dr3lk_track_frame -- results identical to pageable inputs and to the oracle."""
    frames = [load_gray("kitti%d.png" % i) for i in range(3)]
    h, w = frames[0].shape
    pts = golden_case("c1_default_21x21")["prev_pts"][:1200]
    pins = [dr3.PinnedArray((h, (w + 15) // 16 * 16), np.uint8) for _ in frames]
    for pa, f in zip(pins, frames):
        pa.array[...] = 0xA5                                   # the pad columns hold garbage that must never be read as pixels
        pa.array[:, :w] = f
    pf = [pa.array[:, :w] for pa in pins]                      # row step == device pitch: the direct path
    assert pf[0].strides[0] == (w + 15) // 16 * 16 and w % 16 != 0
    orc = oracle.calc_optical_flow_pyr_lk(frames[0], frames[1], pts)
    for a_img, b_img in [(pf[0], pf[1]), (pf[0], frames[1]), (frames[0], pf[1])]:   # either side pinned
        got = ctx.calc_optical_flow_pyr_lk(a_img, b_img, pts)
        assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, orc))
    # a strided ROI of pinned frames: odd width, row step not the device pitch of that width -> packed
    r0, r1 = pf[0][11:311, 37:900], pf[1][11:311, 37:900]
    assert not r0.flags["C_CONTIGUOUS"]
    q = (pts[(pts[:, 0] < 850) & (pts[:, 1] < 290)] + np.float32(3.25))[:600]
    got = ctx.calc_optical_flow_pyr_lk(r0, r1, q, None, (21, 21), 3, (3, 30, 0.01), 0)
    exp = oracle.calc_optical_flow_pyr_lk(np.ascontiguousarray(r0), np.ascontiguousarray(r1), q)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, exp))
    # a band of full-width rows: same step, fewer rows -> direct, and the copy must stop at the last pixel of the last row
    r0, r1 = pf[0][5:205], pf[1][5:205]
    q = pts[pts[:, 1] < 195][:600]
    got = ctx.calc_optical_flow_pyr_lk(r0, r1, q)
    exp = oracle.calc_optical_flow_pyr_lk(np.ascontiguousarray(r0), np.ascontiguousarray(r1), q)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, exp))
    # CONTINUOUS pinned frames (row step == width, the usual cv::Mat): also direct, at an unaligned device pitch -- the odd
    # 1241-wide pair and the even 1240-wide one; one side continuous and the other at the aligned pitch -> both are packed
    odd = [load_gray("kitti_000000.png"), load_gray("kitti_000001.png")]
    assert odd[0].shape[1] == 1241
    for fa, fb in [(frames[0], frames[1]), (odd[0], odd[1])]:
        ca, cb = dr3.PinnedArray(fa.shape, np.uint8), dr3.PinnedArray(fb.shape, np.uint8)
        ca.array[...] = fa; cb.array[...] = fb
        exp = oracle.calc_optical_flow_pyr_lk(fa, fb, pts)
        for a_img, b_img in [(ca.array, cb.array), (ca.array, fb), (fa, cb.array)]:
            got = ctx.calc_optical_flow_pyr_lk(a_img, b_img, pts)
            assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, exp))
        # the reference's literal parameters (30x30 window) and a window without a specialised kernel (generic path reads level 0 raw)
        for win, ml in [((30, 30), 4), ((13, 17), 2)]:
            got = ctx.calc_optical_flow_pyr_lk(ca.array, cb.array, pts[:500], None, win, ml, (3, 30, 0.01), 0)
            exp2 = oracle.calc_optical_flow_pyr_lk(fa, fb, pts[:500], None, win, ml, (3, 30, 0.01), 0)
            assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, exp2)), (win, fa.shape)
        if fa.shape == frames[0].shape:
            got = ctx.calc_optical_flow_pyr_lk(ca.array, pf[1], pts)        # continuous + aligned pitch: mixed, packed
            assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, exp))
        pc = dr3.Pyramid(ctx, ca.array, (21, 21), 3)
        p_, s_, e_, _ = ctx.track_frame(pc, cb.array, pts, keep_next=0)
        assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip((p_, s_, e_), exp))
        pc.close(); ca.free(); cb.free()
    # memory the caller already owns, page-locked in place (dr3lk_host_register): same path, same results; and back to pageable
    ra, rb = frames[0].copy(), frames[1].copy()
    dr3.host_register(ra); dr3.host_register(rb)
    got = ctx.calc_optical_flow_pyr_lk(ra, rb, pts)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, orc))
    dr3.host_unregister(ra); dr3.host_unregister(rb)
    got = ctx.calc_optical_flow_pyr_lk(ra, rb, pts)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, orc))
    with pytest.raises(dr3.Dr3lkError):
        dr3.host_unregister(ra)                                  # not registered any more
    # cached pyramids and the streaming call
    prev = dr3.Pyramid(ctx, pf[0], (21, 21), 3)
    p, s, e, nxt = ctx.track_frame(prev, pf[1], pts, keep_next=2)
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip((p, s, e), orc))
    p2, s2, e2, _ = ctx.track_frame(nxt, pf[2], p[s == 1], keep_next=0)
    orc2 = oracle.calc_optical_flow_pyr_lk(frames[1], frames[2], p[s == 1])
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip((p2, s2, e2), orc2))
    # no points, pinned image: only the pyramid is built
    _, _, _, only = ctx.track_frame(prev, pf[2], np.zeros((0, 2), np.float32), keep_next=2)
    got = ctx.calc_optical_flow_pyr_lk_cached(nxt, only, p[s == 1])
    assert all(np.array_equal(a.view(np.uint8), c.view(np.uint8)) for a, c in zip(got, orc2))
    for x in (prev, nxt, only):
        x.close()
    for pa in pins:
        pa.free()
