#!/bin/bash
for t in "$@"; do
  DR3LK_LIB=$PWD/3dr_b200/lib/libdr3lk_$t.so python tools/bench_configs.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print('$t', 'c1 %.3f ms  c1ref %.3f ms  chain %.2f ms  c4 %.3g  c5 %.3g' % (d['c1_21x21']['gpu_call_ms'], d['c1_reference_30x30']['gpu_call_ms'], d['c2_chain']['gpu_chain_ms'], d['c4_4k_31x31']['gpu_features_per_s'], d['c5_semidense']['gpu_features_per_s']))"
done
