#!/bin/bash
# usage (under gpurun --gpus N): tools/multi_run.sh N  -- the multi-GPU checks of one box:
#   the dr3lk_multi tests (one batch split over all devices == one device, bit for bit), the contract arm under torchrun,
#   the single-process arm (dr3lk_multi from one process) and the host-feed knobs of the end-to-end path
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header 2>&1 | tail -3
run() { # label, env..., then bench args
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu \
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
run chunk256 DR3LK_SLOTS=3 DR3LK_CHUNK_MB=256
run slots4_chunk64 DR3LK_SLOTS=4 DR3LK_CHUNK_MB=64
python bench.py --gpus $N --single-process --steps 3 --warmup 2 > gpurun_out/multi_single_process_${N}gpu.json 2> gpurun_out/multi_single_process_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/multi_single_process_${N}gpu.json"))
    print("single-process N=$N e2e %.4g  h2d/gpu %.1f GB/s  replicas_consistent %s launches %d" % (d["value"], d["e2e"]["h2d_GBps_per_gpu"], d["replicas_consistent"], d["gpu_launches"]))
except Exception as e:
    print("single-process failed", e)
PY
