#!/bin/bash
# usage (under gpurun): tools/ab_large.sh tagA tagB ...  -- A/B of library variants (tools/ab_build.sh) on the configs that
# use the 31x31 / 30x30 kernels (C4, the reference's literal parameters) plus the single-call latencies
for rep in 1 2; do for t in "$@"; do
DR3LK_LIB=$PWD/3dr_b200/lib/libdr3lk_$t.so python tools/bench_configs.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print('$t', 'c4 %.4g feat/s  c1_30x30 %.3f ms  c1_21 %.3f ms  chain %.2f ms' % (d['c4_4k_31x31']['gpu_features_per_s'], d['c1_reference_30x30']['gpu_call_ms'], d['c1_21x21']['gpu_call_ms'], d['c2_chain']['gpu_chain_ms']))"
done; done
