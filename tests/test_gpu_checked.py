"""The checked build of the library (`make -C 3dr_b200/csrc checked`, -DDR3LK_CHECKED): compute-sanitizer is closed on this
GPU pool, so the specialised LK kernels carry their own bounds checks -- every staged rectangle against the apron-carrying
allocation it is copied from, every shared-memory load against its region and the staged rows, every output index against
the batch.  The cases that stress those bounds (windows hanging over every border, levels just above the smallest accepted
size, odd widths, ragged batches, cached pyramids, both staging flavours) run under it in a child process; they must raise
no violation, execute a large number of checks, and stay bit-exact against the oracle."""
import os
import subprocess
import sys

import pytest

from _common import ROOT

pytestmark = pytest.mark.gpu

CHECKED = os.path.join(ROOT, "3dr_b200", "lib", "libdr3lk_checked.so")

CODE = r'''
import importlib, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle
from _common import golden_case, load_gray, random_points
m = importlib.import_module("3dr_b200")
def same(got, exp):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(got, exp))
rng = np.random.default_rng(23)
with m.Context(0) as ctx:
    assert ctx.debug_check_read()[0] == 0
    # windows hanging over all four borders, levels just above the smallest accepted size, all three specialised windows
    for (h, w, win, ml) in [(23, 34, (21, 21), 0), (47, 156, (21, 21), 1), (120, 97, (31, 31), 1), (64, 33, (30, 30), 0), (200, 333, (21, 21), 3),
                            (33, 131, (31, 31), 2), (376, 1241, (21, 21), 3), (375, 500, (30, 30), 4)]:
        base = rng.integers(0, 256, (h + 8, w + 8)).astype(np.float32)
        base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) / 4
        a = base[4:4 + h, 4:4 + w].astype(np.uint8); b = base[3:3 + h, 5:5 + w].astype(np.uint8)
        pts = random_points(rng, w, h, 800, margin=win[0] + 3)
        edge = np.array([[x, y] for x in (-win[0], -0.4, 0.0, w / 2, w - 1.0, w - 0.2, w + win[0] - 1) for y in (-win[1], -0.3, 0.0, h / 2, h - 1.0, h - 0.1, h + win[1] - 1)], np.float32)
        pts = np.concatenate([pts, edge]).astype(np.float32)
        init = (pts + rng.normal(0, 4.0, pts.shape)).astype(np.float32)
        for flags, ini in ((0, None), (4, init)):
            assert same(ctx.calc_optical_flow_pyr_lk(a, b, pts, ini, win, ml, (3, 30, 0.01), flags), oracle.calc_optical_flow_pyr_lk(a, b, pts, ini, win, ml, (3, 30, 0.01), flags)), (h, w, win)
    # golden cases, a ragged host batch, cached pyramids + streaming
    for case in ("c1_default_21x21", "c1_31x31_L4", "c1_reference_30x30_initflow", "oddwidth_21x21"):
        g = golden_case(case)
        a, b = load_gray(g["prev"]), load_gray(g["next"])
        args = (a, b, g["prev_pts"], g["init"], g["win"], g["max_level"], g["crit"], g["flags"])
        assert same(ctx.calc_optical_flow_pyr_lk(*args), oracle.calc_optical_flow_pyr_lk(*args)), case
    fr = [load_gray("kitti%%d.png" %% i) for i in range(4)]
    prev = np.stack(fr[:3]); nxt = np.stack(fr[1:])
    pts = random_points(rng, 1240, 376, 2500, margin=30)
    offs = np.array([0, 900, 900, 2500], np.int32)
    got = ctx.track_batch_host(prev, nxt, pts, offs)
    for i in range(3):
        sl = slice(offs[i], offs[i + 1])
        assert same((got[0][sl], got[1][sl], got[2][sl]), oracle.calc_optical_flow_pyr_lk(prev[i], nxt[i], pts[sl]))
    pyr = m.Pyramid(ctx, fr[0])
    p, s, e, nx = ctx.track_frame(pyr, fr[1], pts, keep_next=2)
    assert same((p, s, e), oracle.calc_optical_flow_pyr_lk(fr[0], fr[1], pts))
    pyr.close(); nx.close()
    viol, kind, info, n_checks = ctx.debug_check_read()
    print("checked: %%d violations (first kind %%d, detail %%d), %%d checks" %% (viol, kind, info, n_checks))
    assert viol == 0 and n_checks > 100000
    # negative control: flags bit 30 makes the checked kernels pretend every previous-frame allocation ends one window early;
    # template windows at the lower border must then be reported (kind 1), nothing else changes
    a, b = fr[0], fr[1]
    low = np.stack([rng.uniform(0, 1240, 300), rng.uniform(376 - 12, 376 + 9, 300)], 1).astype(np.float32)
    exp = oracle.calc_optical_flow_pyr_lk(a, b, low)
    assert same(ctx.calc_optical_flow_pyr_lk(a, b, low, flags=0x40000000), exp)
    viol, kind, info, n_checks = ctx.debug_check_read()
    print("negative control: %%d violations, first kind %%d" %% (viol, kind))
    assert viol > 0 and kind == 1
    assert same(ctx.calc_optical_flow_pyr_lk(a, b, low), exp) and ctx.debug_check_read()[0] == 0
print("checked ok")
''' % (ROOT, os.path.join(ROOT, "tests"))


@pytest.mark.parametrize("tma", ["0", "1"])
def test_checked_build_reports_no_violation(tma):
    if not os.path.exists(CHECKED):
        subprocess.run(["make", "-C", os.path.join(ROOT, "3dr_b200", "csrc"), "checked"], check=True, capture_output=True)
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, DR3LK_LIB=CHECKED, DR3LK_TMA=tma), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "checked ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_default_build_has_no_checks(ctx, dr3):
    with pytest.raises(dr3.Dr3lkError) as e:
        ctx.debug_check_read()
    assert e.value.code == dr3.E_UNSUPPORTED
