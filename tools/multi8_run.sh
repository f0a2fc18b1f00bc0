#!/bin/bash
# usage (under gpurun --gpus N): tools/multi8_run.sh N -- scaling evidence of one N-GPU box: concurrent H2D ceiling, the contract
# arm under torchrun (default and 4 pipeline slots), the single-process arm, the 2+-device split test
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_probe.py 1024 12 2>/dev/null | tail -1 | tee gpurun_out/h2d_probe_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 tools/h2d_probe.py 1024 12 2>/dev/null | tail -1 | tee gpurun_out/h2d_probe_1gpu.json
python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header 2>&1 | tail -2
run() {
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu \
      > gpurun_out/multi_${label}_${N}gpu.json 2> gpurun_out/multi_${label}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/multi_${label}_${N}gpu.json"))
    print("${label} N=$N value %.4g e2e %.4g  h2d/gpu %.1f GB/s  ms/step %.1f e2e ms %.1f" % (d["value"], d["e2e"]["value"], d["e2e"]["h2d_GBps_per_gpu"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("${label} failed", e)
PY
}
run default DR3LK_SLOTS=3
run slots4 DR3LK_SLOTS=4
python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu > gpurun_out/multi_n1_on_${N}gpu_box.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/multi_n1_on_${N}gpu_box.json')); print('N=1 on this box: value %.4g e2e %.4g h2d/gpu %.1f GB/s' % (d['value'], d['e2e']['value'], d['e2e']['h2d_GBps_per_gpu']))"
python bench.py --gpus $N --single-process --steps 3 --warmup 2 > gpurun_out/multi_single_process_${N}gpu.json 2> gpurun_out/multi_single_process_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/multi_single_process_${N}gpu.json')); print('single-process N=$N e2e %.4g  h2d/gpu %.1f GB/s  replicas_consistent %s' % (d['value'], d['e2e']['h2d_GBps_per_gpu'], d['replicas_consistent']))"
