"""The headline configuration (C3: synthetic 1241x376 pairs, 8192 goodFeaturesToTrack corners, 4-level 21x21, (30, 0.01);
SURVEY.md 8d) on 64 host-generated pairs: GPU (through the C ABI's host-batch entry point) against the real OpenCV code path
(cv2.calcOpticalFlowPyrLK, what the reference calls at src/initialization.cpp:608-613) and, bit for bit, against the oracle.

north_star gates: status agreement >= 99.9 %, |dpos| <= 0.01 px for points both implementations mark converged.  The position
gate is evaluated and BOUNDED on two populations: all jointly tracked points, and the subset whose level-0 iterations ended
before the iteration cap ("converged")."""
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import cv2_ref
from _common import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not cv2_ref.HAVE_CV2, reason="cv2 not importable")]

N_PAIRS = 64


def test_c3_64_pairs_vs_cv2_and_oracle(ctx):
    import cv2
    import synth
    prev, nxt, pts, offs = synth.make_batch(N_PAIRS)  # seeds 1000 .. 1063, the recipe bench.py times
    p, s, e, stats = ctx.track_batch_host(prev, nxt, pts, offs, want_stats=True)
    tot = dict(n=0, agree=0, both=0, over=0, conv=0, over_conv=0, maxd=0.0, maxd_conv=0.0)
    for i in range(N_PAIRS):
        sl = slice(offs[i], offs[i + 1])
        # bit-exact against the oracle (positions, status, err) on the configuration the headline is quoted on
        po, so, eo, tr = oracle.calc_optical_flow_pyr_lk(prev[i], nxt[i], pts[sl], trace=True)
        assert np.array_equal(s[sl], so), ("status vs oracle", i)
        assert np.array_equal(p[sl].view(np.uint32), po.view(np.uint32)), ("positions vs oracle", i)
        assert np.array_equal(e[sl].view(np.uint32), eo.view(np.uint32)), ("err vs oracle", i)
        # the kernel's own iteration counts (the input of the algorithmic-bytes model) equal the oracle's
        assert np.array_equal((stats[sl] & 0xFFFF).astype(np.int64), tr["iters"].sum(1)), ("iterations vs oracle", i)
        pc, sc, _ = cv2.calcOpticalFlowPyrLK(prev[i], nxt[i], pts[sl].reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                             criteria=(3, 30, 0.01))
        pc, sc = pc.reshape(-1, 2), sc.ravel()
        both = (sc == 1) & (s[sl] == 1)
        conv = both & (tr["iters"][:, 0] < 30)
        d = np.linalg.norm(pc.astype(np.float64) - p[sl], axis=1)
        tot["n"] += len(sc); tot["agree"] += int((sc == s[sl]).sum()); tot["both"] += int(both.sum()); tot["conv"] += int(conv.sum())
        tot["over"] += int((d[both] > 0.01).sum()); tot["over_conv"] += int((d[conv] > 0.01).sum())
        tot["maxd"] = max(tot["maxd"], float(d[both].max())); tot["maxd_conv"] = max(tot["maxd_conv"], float(d[conv].max()))
    print("C3 parity vs cv2 over %d pairs: %s" % (N_PAIRS, tot))
    assert tot["n"] == N_PAIRS * 8192
    assert tot["agree"] / tot["n"] >= 0.999, tot           # measured: 524288 / 524288
    assert tot["over_conv"] == 0 and tot["maxd_conv"] <= 0.01, tot
    assert tot["over"] == 0 and tot["maxd"] <= 0.01, tot   # ALL jointly tracked points, capped ones included (measured max 0.0024 px)
