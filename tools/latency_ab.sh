#!/bin/bash
# Same-box A/B of the single-pair latency measures (compiled C++ caller, 3dr_b200/host/call_latency): each knob switches ONE of them
# off; configurations alternate three times.  usage (GPU box): bash tools/latency_ab.sh [tag ...]
set -e
python - <<'PY'
import os, sys, numpy as np, importlib
sys.path.insert(0, "tests")
from _common import load_gray
dr3 = importlib.import_module("3dr_b200")
a, b = load_gray("kitti0.png"), load_gray("kitti1.png")
os.makedirs("gpurun_out/lat", exist_ok=True)
for n, f in (("a", a), ("b", b)):
    with open("gpurun_out/lat/%s.pgm" % n, "wb") as fh:
        fh.write(b"P5\n%d %d\n255\n" % (f.shape[1], f.shape[0])); fh.write(np.ascontiguousarray(f).tobytes())
with dr3.Context(0) as c:
    xy, _, _ = c.fast_detect(a)
np.savetxt("gpurun_out/lat/pts.txt", xy.astype(np.float32), fmt="%.9g")
PY
run() { env "$@" 3dr_b200/host/call_latency gpurun_out/lat/a.pgm gpurun_out/lat/b.pgm gpurun_out/lat/pts.txt 400; }
# other builds of the library (tools/ab_build.sh tags) given as arguments run through LD_LIBRARY_PATH
for t in "$@"; do mkdir -p /tmp/dr3lk_ab/$t; cp 3dr_b200/lib/libdr3lk_$t.so /tmp/dr3lk_ab/$t/libdr3lk.so; done
for rep in 1 2 3; do
  echo "all_on        $(run X=1)"
  for t in "$@"; do echo "lib_$t   $(run LD_LIBRARY_PATH=/tmp/dr3lk_ab/$t)"; done
  echo "no_pdl        $(run DR3LK_NO_PDL=1)"
  echo "no_direct_out $(run DR3LK_NO_DIRECT_OUT=1)"
  echo "no_mapped_pts $(run DR3LK_NO_MAPPED_PTS=1)"
  echo "all_off       $(run DR3LK_NO_PDL=1 DR3LK_NO_DIRECT_OUT=1 DR3LK_NO_MAPPED_PTS=1)"
done
