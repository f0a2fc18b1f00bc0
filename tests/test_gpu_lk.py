"""GPU parity of the LK tracker: bit-exact against the CPU oracle, within the north_star tolerances against the
golden vectors of the real OpenCV path, plus edge cases and size-independent properties."""
import numpy as np
import pytest

import oracle
from _common import ALL_TRACKED_BOUNDS, GOLDEN_CASES, compare_lk, golden_case, golden_json, load_gray, random_points

pytestmark = pytest.mark.gpu


def _assert_bit_exact(got, exp, what):
    p, s, e = got
    po, so, eo = exp
    assert np.array_equal(s, so), (what, "status", int((s != so).sum()))
    bad = np.where((p.view(np.uint32) != po.view(np.uint32)).any(axis=1))[0]
    assert bad.size == 0, (what, "positions differ at", bad[:10], p[bad[:5]], po[bad[:5]])
    if e is not None:
        bad = np.where(e.view(np.uint32) != eo.view(np.uint32))[0]
        assert bad.size == 0, (what, "err differs at", bad[:10], e[bad[:5]], eo[bad[:5]])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_lk_matches_oracle_bit_exact_and_cv2_golden(ctx, case):
    g = golden_case(case)
    a, b = load_gray(g["prev"]), load_gray(g["next"])
    got = ctx.calc_optical_flow_pyr_lk(a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"], g["flags"])
    po, so, eo, tr = oracle.calc_optical_flow_pyr_lk(a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"],
                                                      g["flags"], trace=True)
    _assert_bit_exact(got, (po, so, eo), case)
    max_count = min(max(g["crit"][1], 0), 100) if g["crit"][0] & 1 else 30
    m = compare_lk(got[0], got[1], got[2], g["next_pts"], g["status"], g["err"], tr["iters"][:, 0] < max_count)
    assert m["status_agree"] >= 0.999 and m["max_dpos_converged"] <= 0.01 and m["n_over_0p01_converged"] == 0, m
    # all jointly tracked points, iteration-capped ones included: bounded count and size of the outliers (see _common.py)
    n_over, max_d = ALL_TRACKED_BOUNDS[case]
    print("%s: %d jointly tracked, %d over 0.01 px (max %.4f px)" % (case, m["n_both"], m["n_over_0p01_tracked"], m["max_dpos_tracked"]))
    assert m["n_over_0p01_tracked"] <= n_over and m["max_dpos_tracked"] <= max_d and m["frac_within_0p01"] >= 0.997, m


@pytest.mark.parametrize("win,ml,crit,flags", [
    ((21, 21), 3, (3, 30, 0.01), 0), ((3, 3), 5, (3, 10, 0.01), 0), ((5, 7), 2, (1, 4, 0.0), 0), ((21, 21), 0, (3, 30, 0.01), 0),
    ((31, 31), 4, (2, 0, 0.03), 0), ((30, 30), 4, (3, 1000, 1e-3), 4), ((21, 21), 3, (3, 0, 0.01), 0), ((45, 33), 3, (3, 30, 0.01), 8),
    ((21, 21), 3, (3, 30, 0.01), 12)])
def test_lk_parameter_sweep_bit_exact(ctx, win, ml, crit, flags):
    a, b = load_gray("kitti3.png"), load_gray("kitti4.png")
    rng = np.random.default_rng(win[0] * 100 + ml)
    h, w = a.shape
    pts = random_points(rng, w, h, 1500)
    pts[::7] = np.round(pts[::7])          # integer coordinates (weights exactly 1, 0, 0, 0)
    pts[1::11] = np.round(pts[1::11] * 2) / 2  # half-pixel ties for cvRound
    init = (pts + rng.normal(0, 2.0, pts.shape)).astype(np.float32) if flags & 4 else None
    got = ctx.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags)
    exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, init, win, ml, crit, flags)
    _assert_bit_exact(got, exp, (win, ml, crit, flags))


def test_lk_zero_iterations_repeatable(ctx):
    """maxCount = 0 runs the levels of a feature back to back (template, no iterations, next template ...), the tightest
    schedule the asynchronous staging of the kernel sees: every repetition must reproduce the oracle bit for bit.
    (A kernel that waited on two copy groups in flight with wait_group 1 returned a wrong err for ~1 feature in 10^5 here.)"""
    a, b = load_gray("kitti3.png"), load_gray("kitti4.png")
    pts = random_points(np.random.default_rng(2103), a.shape[1], a.shape[0], 1500)
    pts[::7] = np.round(pts[::7])
    exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, (21, 21), 3, (3, 0, 0.01), 0)
    for rep in range(400):
        _assert_bit_exact(ctx.calc_optical_flow_pyr_lk(a, b, pts, None, (21, 21), 3, (3, 0, 0.01), 0), exp, ("rep", rep))
    exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, (31, 31), 4, (3, 0, 0.01), 0)
    for rep in range(100):
        _assert_bit_exact(ctx.calc_optical_flow_pyr_lk(a, b, pts, None, (31, 31), 4, (3, 0, 0.01), 0), exp, ("rep31", rep))


def test_lk_without_err_output(ctx):
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = random_points(np.random.default_rng(2), 1240, 376, 800)
    p, s, e = ctx.calc_optical_flow_pyr_lk(a, b, pts, want_err=False)
    po, so, _ = oracle.calc_optical_flow_pyr_lk(a, b, pts, want_err=False)
    assert e is None and np.array_equal(s, so) and np.array_equal(p.view(np.uint32), po.view(np.uint32))


def test_lk_small_images_and_borders(ctx):
    rng = np.random.default_rng(11)
    for (h, w, win) in [(24, 24, (21, 21)), (40, 30, (9, 9)), (12, 200, (5, 5)), (23, 23, (21, 21)), (16, 16, (21, 21))]:
        base = rng.integers(0, 256, (h + 8, w + 8)).astype(np.float32)
        base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) / 4  # some smoothness
        a = base[4:4 + h, 4:4 + w].astype(np.uint8)
        b = base[3:3 + h, 5:5 + w].astype(np.uint8)
        pts = random_points(rng, w, h, 300, margin=win[0] + 3)
        got = ctx.calc_optical_flow_pyr_lk(a, b, pts, None, win, 3, (3, 30, 0.01), 0)
        exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, win, 3, (3, 30, 0.01), 0)
        _assert_bit_exact(got, exp, (h, w, win))


def test_lk_apron_layout_borders_and_layout_changes(ctx):
    """The specialised kernels read every level through its apron (reflected pixels, zero derivatives): windows hanging
    over all four borders, levels just above the smallest accepted size, and back-to-back calls whose layouts differ
    (the zero derivative aprons are cleared per layout) must stay bit-exact."""
    rng = np.random.default_rng(23)
    shapes = [(23, 34, (21, 21), 0), (47, 156, (21, 21), 1), (120, 97, (31, 31), 1), (64, 33, (30, 30), 0), (200, 333, (21, 21), 3),
              (90, 70, (45, 33), 1), (200, 333, (21, 21), 3), (33, 131, (31, 31), 2)]
    for (h, w, win, ml) in shapes:
        base = rng.integers(0, 256, (h + 8, w + 8)).astype(np.float32)
        base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) / 4
        a = base[4:4 + h, 4:4 + w].astype(np.uint8)
        b = base[3:3 + h, 5:5 + w].astype(np.uint8)
        pts = random_points(rng, w, h, 600, margin=win[0] + 3)
        # a ring of points hugging the four borders and corners
        edge = np.array([[x, y] for x in (-0.4, 0.0, 0.6, w / 2, w - 1.3, w - 1.0, w - 0.2) for y in (-0.3, 0.0, 0.7, h / 2, h - 1.4, h - 1.0, h - 0.1)],
                        np.float32)
        pts = np.concatenate([pts, edge])
        init = (pts + rng.normal(0, 3.0, pts.shape)).astype(np.float32)
        for flags, ini in ((0, None), (4, init)):
            got = ctx.calc_optical_flow_pyr_lk(a, b, pts, ini, win, ml, (3, 30, 0.01), flags)
            exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, ini, win, ml, (3, 30, 0.01), flags)
            _assert_bit_exact(got, exp, (h, w, win, ml, flags))


def test_lk_errors_and_empty(ctx, dr3):
    a = load_gray("kitti0.png")
    one = np.zeros((1, 2), np.float32)
    for kw in [dict(win=(2, 2)), dict(max_level=-1), dict(win=(21, 2))]:
        with pytest.raises(dr3.Dr3lkError) as e:
            ctx.calc_optical_flow_pyr_lk(a, a, one, **kw)
        assert e.value.code == dr3.E_ARG and "maxLevel >= 0 && winSize.width > 2" in str(e.value)
    with pytest.raises(dr3.Dr3lkError):
        ctx.calc_optical_flow_pyr_lk(a, a[:, :-1], one)
    with pytest.raises(dr3.Dr3lkError):
        ctx.calc_optical_flow_pyr_lk(a, a, one, None, flags=dr3.USE_INITIAL_FLOW)
    p, s, e = ctx.calc_optical_flow_pyr_lk(a, a, np.zeros((0, 2), np.float32))
    assert p.shape == (0, 2) and s.shape == (0,) and e.shape == (0,)
    # a context stays usable after an error
    p, s, e = ctx.calc_optical_flow_pyr_lk(a, a, np.array([[300.25, 200.5]], np.float32))
    assert s[0] == 1 and np.allclose(p[0], [300.25, 200.5], atol=1e-3) and e[0] == 0


def test_chain_tracking_counts_match_reference_path(ctx):
    """C2: kitti0..9 frame-to-frame chain and the reference-style anchored warm-start loop."""
    g = golden_json("chain_counts.json")
    frames = [load_gray("kitti%d.png" % i) for i in range(10)]
    pts = golden_case("c1_default_21x21")["prev_pts"][:g["chain_21x21"][0]]
    cur, surv = pts, [len(pts)]
    for i in range(9):
        p, s, _ = ctx.calc_optical_flow_pyr_lk(frames[i], frames[i + 1], cur)
        cur = p[s == 1]
        surv.append(len(cur))
    # survivor counts can differ from cv2 by a borderline point or two per step (fp32 accumulation order in OpenCV)
    assert all(abs(x - y) <= max(3, 0.002 * y) for x, y in zip(surv, g["chain_21x21"])), (surv, g["chain_21x21"])
    ref, curp, anch = pts.copy(), pts.copy(), []
    for i in range(1, 10):
        p, s, _ = ctx.calc_optical_flow_pyr_lk(frames[0], frames[i], ref, curp, (30, 30), 4, (3, 1000, 1e-3), 4)
        ref, curp = ref[s == 1], p[s == 1]
        anch.append(len(ref))
    assert all(abs(x - y) <= max(3, 0.002 * y) for x, y in zip(anch, g["anchored_30x30"])), (anch, g["anchored_30x30"])


def _torch_batch(ctx, dr3, prev, nxt, pts_list, **kw):
    import torch
    B, h, w = prev.shape
    offs = np.zeros(B + 1, np.int32)
    offs[1:] = np.cumsum([len(p) for p in pts_list])
    allp = np.concatenate(pts_list).astype(np.float32)
    stream = torch.cuda.Stream()  # torch and the library share one non-default stream
    with torch.cuda.stream(stream):
        dp, dn = torch.from_numpy(prev).cuda(), torch.from_numpy(nxt).cuda()
        dpts = torch.from_numpy(allp).cuda()
        dnext = torch.zeros_like(dpts)
        dst = torch.zeros(len(allp), dtype=torch.uint8, device="cuda")
        derr = torch.zeros(len(allp), dtype=torch.float32, device="cuda")
        dstats = torch.zeros(len(allp), dtype=torch.int32, device="cuda")
        ctx.set_stream(stream.cuda_stream)
        ctx.track_batch(dp.data_ptr(), dn.data_ptr(), w, h, w, h * w, B, dpts.data_ptr(), dnext.data_ptr(), dst.data_ptr(),
                        derr.data_ptr(), offs, dstats.data_ptr(), **kw)
        ctx.synchronize()
        ctx.set_stream(None)
    return dnext.cpu().numpy(), dst.cpu().numpy(), derr.cpu().numpy(), dstats.cpu().numpy().view(np.uint32), offs


def test_batch_device_and_host_paths_match_single_calls(ctx, dr3):
    frames = np.stack([load_gray("kitti%d.png" % i) for i in range(6)])
    prev, nxt = np.ascontiguousarray(frames[:5]), np.ascontiguousarray(frames[1:6])
    rng = np.random.default_rng(4)
    pts_list = [random_points(rng, 1240, 376, n) for n in (700, 0, 1, 333, 1024)]  # ragged, incl. an empty pair
    p, s, e, stats, offs = _torch_batch(ctx, dr3, prev, nxt, pts_list)
    hp, hs, he, hstats = ctx.track_batch_host(prev, nxt, np.concatenate(pts_list), offs, want_stats=True, chunk_pairs=2)
    assert np.array_equal(hp.view(np.uint32), p.view(np.uint32)) and np.array_equal(hs, s) and np.array_equal(he, e)
    assert np.array_equal(hstats, stats)
    tot_it = 0
    for b in range(5):
        sl = slice(offs[b], offs[b + 1])
        po, so, eo, tr = oracle.calc_optical_flow_pyr_lk(prev[b], nxt[b], pts_list[b], trace=True)
        _assert_bit_exact((p[sl], s[sl], e[sl]), (po, so, eo), ("pair", b))
        it, tl, ep = dr3.decode_stats(stats[sl])
        assert np.array_equal(it, tr["iters"].sum(1)) and np.array_equal(tl, (tr["code"] >= 2).sum(1))
        tot_it += it.sum()
    assert tot_it > 0 and dr3.algorithmic_bytes(stats, (21, 21)) > 0


def test_properties_full_size_synthetic(ctx, dr3):
    """Size-independent properties on a full-size (1241x376) synthetic pair: identity gives zero flow and zero error;
    a pure integer translation is recovered for textured points; results do not depend on batch position."""
    rng = np.random.default_rng(99)
    h, w = 376, 1241
    big = rng.normal(0, 1, (h + 64, w + 64)).astype(np.float32)
    for _ in range(3):
        big = (big + np.roll(big, 1, 0) + np.roll(big, -1, 0) + np.roll(big, 1, 1) + np.roll(big, -1, 1)) / 5
    big = np.clip(128 + 48 * big / big.std(), 0, 255).astype(np.uint8)
    a = big[32:32 + h, 32:32 + w]
    b = big[32 - 3:32 - 3 + h, 32 + 5:32 + 5 + w]  # content moves by (-5, +3)
    xs, ys = np.meshgrid(np.arange(40, w - 40, 16, dtype=np.float32), np.arange(40, h - 40, 16, dtype=np.float32))
    pts = np.stack([xs.ravel(), ys.ravel()], 1)
    p, s, e = ctx.calc_optical_flow_pyr_lk(a, a, pts)
    assert s.all() and np.abs(p - pts).max() < 1e-3 and np.abs(e).max() == 0
    p, s, e = ctx.calc_optical_flow_pyr_lk(a, b, pts)
    assert s.mean() > 0.99
    d = p[s == 1] - pts[s == 1] - np.array([-5, 3], np.float32)
    assert np.median(np.abs(d)) < 0.02 and np.percentile(np.abs(d), 99) < 0.2
    prev = np.ascontiguousarray(np.stack([a, a, a]))
    nxt = np.ascontiguousarray(np.stack([b, a, b]))
    bp, bs, be, _, offs = _torch_batch(ctx, dr3, prev, nxt, [pts, pts, pts])
    n = len(pts)
    assert np.array_equal(bp[:n].view(np.uint32), p.view(np.uint32)) and np.array_equal(bp[2 * n:].view(np.uint32), p.view(np.uint32))
    assert np.array_equal(bs[:n], s) and np.array_equal(be[2 * n:], e) and np.abs(bp[n:2 * n] - pts).max() < 1e-3


def test_c5_semidense_lattice_bit_exact(ctx):
    """C5: 4-px lattice (310 x 94 = 29140 points) on a 1241x376 synthetic pair -- many texture-less points, exercises the
    minEig rejection path and status parity at full size."""
    from tools import synth
    a, b, _ = synth.make_pair(1007)
    pts = synth.lattice(1241, 376)
    assert len(pts) == 29140
    got = ctx.calc_optical_flow_pyr_lk(a, b, pts)
    exp = oracle.calc_optical_flow_pyr_lk(a, b, pts)
    _assert_bit_exact(got, exp, "c5")
    # flat images: every lattice point is rejected by the min-eigenvalue test
    flat = np.full((376, 1241), 77, np.uint8)
    p, s, e = ctx.calc_optical_flow_pyr_lk(flat, flat, pts)
    assert not s.any() and not e.any()


def test_c4_4k_31x31_five_levels_bit_exact(ctx):
    """C4: 3840x2160, 31x31 window, maxLevel 4 (5 levels, none dropped), corners from goodFeaturesToTrack."""
    from tools import synth
    a, b, M = synth.make_pair(2000, 3840, 2160)
    pts = synth.corners(a, 3000, 5)
    got = ctx.calc_optical_flow_pyr_lk(a, b, pts, None, (31, 31), 4, (3, 30, 0.01), 0)
    exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, (31, 31), 4, (3, 30, 0.01), 0)
    _assert_bit_exact(got, exp, "c4")
    gt = pts @ M[:, :2].T.astype(np.float32) + M[:, 2].astype(np.float32)
    d = np.linalg.norm(got[0] - gt, axis=1)[got[1] == 1]
    assert got[1].mean() > 0.98 and np.median(d) < 0.1
    # the box pyramid of a 3840-wide frame takes the SSE2 rounding path on x86 (3840 % 16 == 0), SURVEY.md 8d
    box = ctx.box_pyramid(a, 3)
    exp_box = oracle.box_pyramid(a, 3)
    assert all(np.array_equal(x, y) for x, y in zip(box[1:], exp_box[1:]))


def test_lk_tma_staging_variant_is_bit_exact():
    """DR3LK_TMA=1 selects the kernels that stage their three regions with cp.async.bulk.tensor boxes + per-warp mbarriers
    instead of cp.async rounds (kept as a measured alternative, DESIGN.md: 201.2 vs 190.9 ms per C3 launch).  The switch is
    read once per process, so the check runs in a child process: golden cases of all three specialised kernels, a ragged
    batch, zero iterations and windows hanging over every border -- bit-exact against the oracle."""
    import os
    import subprocess
    import sys
    from _common import ROOT
    code = r'''
import importlib, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle
from _common import golden_case, load_gray, random_points
m = importlib.import_module("3dr_b200")
def same(got, exp):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(got, exp))
with m.Context(0) as ctx:
    for case in ("c1_default_21x21", "c1_31x31_L4", "c1_reference_30x30_initflow", "oddwidth_21x21"):
        g = golden_case(case)
        a, b = load_gray(g["prev"]), load_gray(g["next"])
        args = (a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"], g["flags"])
        assert same(ctx.calc_optical_flow_pyr_lk(*args), oracle.calc_optical_flow_pyr_lk(*args)), case
    a, b = load_gray("kitti3.png"), load_gray("kitti4.png")
    rng = np.random.default_rng(5)
    pts = random_points(rng, a.shape[1], a.shape[0], 3000, margin=45)
    for crit in ((3, 0, 0.01), (3, 30, 0.01)):
        exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, (21, 21), 3, crit, 0)
        for rep in range(20):
            assert same(ctx.calc_optical_flow_pyr_lk(a, b, pts, None, (21, 21), 3, crit, 0), exp)
    prev = np.stack([a, b, a]); nxt = np.stack([b, a, b])
    offs = np.array([0, 700, 700, 3000], np.int32)
    got = ctx.track_batch_host(prev, nxt, pts, offs)
    for i in range(3):
        sl = slice(offs[i], offs[i + 1])
        po, so, eo = oracle.calc_optical_flow_pyr_lk(prev[i], nxt[i], pts[sl])
        assert np.array_equal(got[1][sl], so) and np.array_equal(got[0][sl].view(np.uint32), po.view(np.uint32)) and np.array_equal(got[2][sl], eo)
print("tma ok")
''' % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DR3LK_TMA="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "tma ok" in r.stdout, r.stderr[-3000:]


@pytest.mark.parametrize("win,ml", [((21, 21), 3), ((31, 31), 4), ((9, 13), 2)])
def test_lk_non_finite_and_huge_points_are_lost_not_fatal(ctx, dr3, win, ml):
    """NaN / inf / huge coordinates in prevPts (and in the initial flow): x86 OpenCV's cvFloor saturates them to INT_MIN, every
    bounds test fails and the point is reported lost with err 0 -- same here, in the specialised and in the generic kernel,
    and the finite points of the same call are unaffected."""
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    rng = np.random.default_rng(77)
    good = random_points(rng, a.shape[1], a.shape[0], 400, margin=0)
    nan, inf, big = np.float32(np.nan), np.float32(np.inf), np.float32(1e30)
    weird = np.array([[nan, 10], [10, nan], [nan, nan], [inf, 5], [5, -inf], [big, 3], [3, -big], [2.2e9, 1], [1, -2.2e9], [4294967296.0, 7],
                      [1e-42, 1e-42], [-0.0, -0.0], [a.shape[1] - 1, a.shape[0] - 1]], np.float32)
    pts = np.concatenate([good[:200], weird, good[200:]]).astype(np.float32)
    lost = slice(200, 200 + 10)  # the first ten weird points can never be inside a frame
    with np.errstate(all="ignore"):
        exp = oracle.calc_optical_flow_pyr_lk(a, b, pts, None, win, ml)
        got = ctx.calc_optical_flow_pyr_lk(a, b, pts, None, win, ml)
        assert np.array_equal(got[1], exp[1]) and not got[1][lost].any() and not got[2][lost].any()
        ok = got[1] == 1
        assert ok.sum() > 300 and np.array_equal(got[0][ok].view(np.uint32), exp[0][ok].view(np.uint32)) and np.array_equal(got[2], exp[2])
        # the same garbage as the INITIAL FLOW of otherwise fine points
        init = pts.copy()
        init[200:213] = weird[::-1]
        prev2 = pts.copy()
        prev2[200:213] = good[:13]
        exp = oracle.calc_optical_flow_pyr_lk(a, b, prev2, init, win, ml, flags=4)
        got = ctx.calc_optical_flow_pyr_lk(a, b, prev2, init, win, ml, flags=dr3.USE_INITIAL_FLOW)
        assert np.array_equal(got[1], exp[1])
        ok = got[1] == 1
        assert np.array_equal(got[0][ok].view(np.uint32), exp[0][ok].view(np.uint32)) and np.array_equal(got[2][ok], exp[2][ok])


def test_single_call_with_more_points_than_the_mapped_output_limit(ctx, dr3):
    """Above 16384 points the single-pair calls copy the results back instead of letting the LK kernel write them into the
    mapped pinned mirror: 20000 points through the two-image call and the streaming call, bit-exact against the oracle."""
    a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
    pts = random_points(np.random.default_rng(20000), a.shape[1], a.shape[0], 20000, margin=10)
    exp = oracle.calc_optical_flow_pyr_lk(a, b, pts)
    _assert_bit_exact(ctx.calc_optical_flow_pyr_lk(a, b, pts), exp, "20000 points")
    init = pts + np.float32(0.5)
    exp2 = oracle.calc_optical_flow_pyr_lk(a, b, pts, init, (21, 21), 3, (3, 30, 0.01), dr3.USE_INITIAL_FLOW)
    prev = dr3.Pyramid(ctx, a, (21, 21), 3)
    p, s, e, _ = ctx.track_frame(prev, b, pts, init, 3, (3, 30, 0.01), dr3.USE_INITIAL_FLOW, keep_next=0)
    _assert_bit_exact((p, s, e), exp2, "20000 points, streaming, initial flow")
    prev.close()


def test_latency_measures_switched_off_give_the_same_results():
    """The single-pair latency measures (programmatic dependent launches, results written into the mapped pinned mirror, points
    read from it, direct upload of continuous frames) each have a knob that restores the plain path (tools/latency_ab.sh).
    The knobs are read once per process, so the plain paths run in a child process: golden cases, pinned and pageable frames,
    the streaming call -- bit-exact against the oracle."""
    import os
    import subprocess
    import sys
    from _common import ROOT
    code = r'''
import importlib, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle
from _common import golden_case, load_gray, random_points
m = importlib.import_module("3dr_b200")
def same(got, exp):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(got, exp))
with m.Context(0) as ctx:
    for case in ("c1_default_21x21", "c1_reference_30x30_initflow", "oddwidth_21x21"):
        g = golden_case(case)
        a, b = load_gray(g["prev"]), load_gray(g["next"])
        args = (a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"], g["flags"])
        exp = oracle.calc_optical_flow_pyr_lk(*args)
        for rep in range(5):
            assert same(ctx.calc_optical_flow_pyr_lk(*args), exp), case
        pa, pb = m.PinnedArray(a.shape, np.uint8), m.PinnedArray(b.shape, np.uint8)
        pa.array[...] = a; pb.array[...] = b
        assert same(ctx.calc_optical_flow_pyr_lk(pa.array, pb.array, *args[2:]), exp), case
        if tuple(g["win"]) == (21, 21):
            prev = m.Pyramid(ctx, pa.array, (21, 21), g["max_level"])
            p, s, e, _ = ctx.track_frame(prev, pb.array, g["prev_pts"], g["init"], g["max_level"], g["crit"], g["flags"], keep_next=0)
            assert same((p, s, e), exp), case
            prev.close()
        pa.free(); pb.free()
print("plain ok")
''' % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, DR3LK_NO_PDL="1", DR3LK_NO_DIRECT_OUT="1", DR3LK_NO_MAPPED_PTS="1", DR3LK_PACK_PAGEABLE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "plain ok" in r.stdout, r.stderr[-3000:]
