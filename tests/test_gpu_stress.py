"""Reduced-repetition runs of the two stress tools under `pytest -m gpu` (compute-sanitizer is closed on this GPU pool, so
rare races -- a copy that had not landed, a stale scratch buffer changing hands between entry points -- have to show up as
non-reproducible results).  tools/stress_all.py: every entry point against its own first result, interleaved;
tools/stress_lk.py: the LK parameter sweep against oracle results.  The full-length runs are `python tools/stress_all.py 500` /
`python tools/stress_lk.py 2000`."""
import os
import subprocess
import sys

import pytest

from _common import ROOT

pytestmark = pytest.mark.gpu


def _run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *args], capture_output=True, text=True, timeout=900)
    return r.returncode, r.stdout[-3000:] + r.stderr[-2000:]


def test_stress_all_entry_points_reproducible():
    rc, out = _run("stress_all.py", "12")
    assert rc == 0 and "0 mismatching" in out, out


def test_stress_lk_sweep_against_oracle():
    rc, out = _run("stress_lk.py", "25")
    assert rc == 0 and ", 0 mismatching" in out, out
