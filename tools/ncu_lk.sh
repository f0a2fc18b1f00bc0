#!/bin/bash
# ncu capture of the LK kernel on a small C3 batch (run under gpurun; writes into gpurun_out/)
set -e
CMD="python bench.py --pairs 296 --base-pairs 8 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lk_fast -s 1 -c 1 -f -o gpurun_out/lk_prof $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_full.log
