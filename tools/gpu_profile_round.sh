#!/bin/bash
# Round evidence: (1) plain bench (must exit 0 first), (2) ncu launch list of the same command (~15 min of GPU time: 14 k
# launches), (3) ncu --set full of post_physics_kernel in situ.  GEMM captures: tools/gpu_profile_gemm.sh / gpu_profile_one.sh.
# Run under gpurun; outputs land in gpurun_out/.  Summarise the launch list with the snippet in profiles/README.md.
tag=$1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$B > gpurun_out/bench_for_ncu_$tag.json 2> gpurun_out/bench_for_ncu_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:post_physics_kernel -s 60 -c 1 -o gpurun_out/prof_post_physics_$tag -f $B > gpurun_out/ncu_full_env_$tag.log 2>&1
ls -la gpurun_out/*$tag*
