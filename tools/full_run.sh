#!/bin/bash
# Round-end style run on the GPU box: smoke, GPU tests, both bench arms, ncu launch list + full capture of the LK kernel.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -5 | tee gpurun_out/tests.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.err
bash tools/ncu_lk.sh > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
./tools/microbench > gpurun_out/microbench.log 2>&1
