CMD="python bench.py --workload c4 --pairs 8 --base-pairs 1 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/c4_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:lk_fast" -s 1 -c 1 -f -o gpurun_out/lk_c4_prof $CMD > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
