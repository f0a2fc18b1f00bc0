#!/bin/bash
# The TMA-staging LK kernel variant (DR3LK_TMA=1): parity tests + A/B against the default (tools/tma_ab.sh), then an ncu capture
# on a small C3 batch (run under gpurun; writes into gpurun_out/)
bash tools/tma_ab.sh 2>&1 | tee gpurun_out/tma2.log
export DR3LK_TMA=1
CMD="python bench.py --pairs 296 --base-pairs 8 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_tma_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lk_fast -s 1 -c 1 -f -o gpurun_out/lk_prof_tma $CMD > gpurun_out/ncu_tma_full.log 2>&1
tail -3 gpurun_out/ncu_tma_full.log
