#!/bin/bash
# usage (under gpurun): tools/ab_run.sh tagA tagB ...   -- alternates the variants twice
for rep in 1 2; do
for t in "$@"; do
  DR3LK_LIB=$PWD/3dr_b200/lib/libdr3lk_$t.so python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$t', 'lk_ms %.1f pyr_ms %.1f value %.4g' % (r['lk_ms_per_launch'], r['pyramid_ms_per_step'], d['value']))"
done
done
