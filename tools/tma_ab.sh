export DR3LK_TMA=1
timeout 600 python -m pytest tests/test_gpu_lk.py tests/test_gpu_next_rows.py tests/test_gpu_c3_parity.py tests/test_gpu_multi.py -m gpu -q --no-header -x 2>&1 | tail -6
for rep in 1 2; do for t in 0 1; do
DR3LK_TMA=$t timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('TMA=$t', 'lk_ms %.1f pyr_ms %.1f value %.4g' % (r['lk_ms_per_launch'], r['pyramid_ms_per_step'], d['value']))"
done; done
