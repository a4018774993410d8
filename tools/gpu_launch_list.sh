#!/bin/bash
# ncu launch list (gpu__time_duration.sum) of bench.py, eager launches, first ~4500 kernels = DAgger iteration + one PPO iteration
tag=$1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-graphs"
$B > gpurun_out/bench_for_ncu_$tag.json 2> gpurun_out/bench_for_ncu_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
wc -l gpurun_out/launches_$tag.csv
