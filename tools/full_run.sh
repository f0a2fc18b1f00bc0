#!/bin/bash
# Round-end style run on the GPU box: smoke, GPU tests, both bench arms, ncu launch list + full capture of the LK kernel,
# ncu capture of the pyramid kernels, the other BASELINE configs and the instruction microbenchmarks.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -5 | tee gpurun_out/tests.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.err
bash tools/ncu_lk.sh > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
bash tools/ncu_pyr.sh
python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
[ -x ./tools/microbench ] && ./tools/microbench > gpurun_out/microbench.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ours.json")); r=json.load(open("gpurun_out/bench_reference.json"))
print("ours value %.4g e2e %.4g | reference %.4g | value/ref %.1f e2e/ref %.1f | lk_ms %.1f pyr_ms %.1f frac %.3f" % (
    d["value"], d["e2e"]["value"], r["value"], d["value"]/r["value"], d["e2e"]["value"]/r["value"],
    d["roofline"]["lk_ms_per_launch"], d["roofline"]["pyramid_ms_per_step"], d["roofline"]["frac"]))
PY
