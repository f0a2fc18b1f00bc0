#!/bin/bash
# quick LK-only numbers: tests + short device-resident bench
python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/benchq.json 2> gpurun_out/benchq.err
python - <<PY
import json
d=json.load(open("gpurun_out/benchq.json"))
r=d["roofline"]
print("value %.4g feat/s  ms/step %.1f  lk_ms %.1f  pyr_ms %.1f  frac %.3f  launches %d" % (d["value"], d["ms_per_step"], r["lk_ms_per_launch"], r["pyramid_ms_per_step"], r["frac"], d["gpu_launches"]))
PY
tail -2 gpurun_out/benchq.err
