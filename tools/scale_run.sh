#!/bin/bash
# usage (under gpurun --gpus N): tools/scale_run.sh N ["c3 kitti ..."]  -- the headline workload and the other named shapes at N GPUs
N=$1
WLS=${2:-"c3 kitti c5 c4"}
mkdir -p gpurun_out
for wl in $WLS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}_${N}gpu.json 2> gpurun_out/bench_${wl}_${N}gpu.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_${N}gpu.json"))
print("$wl", "n_gpus", d["n_gpus"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms %.1f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "tracked %.4f" % d["tracked_fraction"])
PY
done
