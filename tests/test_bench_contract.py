"""CPU tests of bench.py: the reference arm (the reference's own CPU path, runnable without a GPU) prints ONE JSON line with
the contract's keys, and the named workloads produce the shapes BASELINE.json quotes."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from _common import ROOT

cv2 = pytest.importorskip("cv2")


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly one JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run("--impl", "reference", "--cpu-sample-pairs", "2", "--base-pairs", "2", "--steps", "1", "--warmup", "1")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "features/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "features/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C3") and d["config"]["corners_per_pair"] == 8192 and d["config"]["pairs_per_gpu"] == 4096
    # both arms print the SAME config dict (the driver compares them): the keys our arm prints are exactly these
    assert sorted(d["config"]) == ["corners_per_pair", "distinct_pairs_per_gpu", "l2", "max_level", "pairs_per_gpu", "sharding", "win", "workload"]
    assert "larger than L2" in d["config"]["l2"] and d["cpu_sample_pairs_per_step"] == 2


def test_our_arm_builds_the_same_config_dict():
    """static check (no GPU here): our arm emits `cfg` itself as "config" -- nothing is merged into it after the reference
    arm's copy was built"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": cfg') >= 3 and "dict(cfg," not in src and 'cfg["cpu_affinity"]' not in src


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_named_workloads():
    sys.path.insert(0, ROOT)
    import bench
    prev, nxt, pts, offs = bench.make_workload("kitti", 3, 1000)
    assert prev.shape == (3, 376, 1240) and nxt.shape == prev.shape and prev.dtype == np.uint8
    assert offs[0] == 0 and offs[-1] == len(pts) and pts.dtype == np.float32 and len(pts) > 3 * 2000
    bench.W_IMG, bench.H_IMG, bench.CORNERS = 1241, 376, 8192
    prev, nxt, pts, offs = bench.make_workload("c5", 1, 1000)
    assert prev.shape == (1, 376, 1241) and len(pts) == 29140 and np.array_equal(pts[0], [2.0, 2.0])
    prev, nxt, pts, offs = bench.make_workload("c3", 1, 1000)
    assert len(pts) == 8192 and (pts[:, 0] < 1241).all() and (pts[:, 1] < 376).all()
