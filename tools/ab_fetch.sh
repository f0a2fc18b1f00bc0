bash tools/ab_run.sh cur jn
for fm in 4 16 32; do
DR3LK_FETCH_MAX=$fm DR3LK_LIB=$PWD/3dr_b200/lib/libdr3lk_cur.so python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('fetch_max $fm', 'lk_ms %.1f value %.4g' % (r['lk_ms_per_launch'], d['value']))"
done
