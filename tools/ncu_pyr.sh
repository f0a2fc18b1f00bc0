#!/bin/bash
# ncu capture of the pyramid kernels (level-0 pyr_level + pad_level0) on a small C3 batch (run under gpurun)
CMD="python bench.py --pairs 296 --base-pairs 8 --steps 1 --warmup 1 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k "regex:pyr_level|pad_level0" -s 5 -c 2 -f -o gpurun_out/pyr_prof $CMD > gpurun_out/ncu_pyr.log 2>&1
tail -2 gpurun_out/ncu_pyr.log
