#!/bin/bash
# Round evidence: (1) plain bench (exit 0 first), (2) ncu launch list of the same command, (3) ncu --set full of the dominant GEMM
# and of post_physics_kernel, in situ.  Run under gpurun; outputs land in gpurun_out/.
tag=$1
set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_for_ncu_$tag.json 2> gpurun_out/bench_for_ncu_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel<0, 256, true>' -s 60 -c 1 -o gpurun_out/prof_tc_fwd_pair_$tag -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_fwd_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel<1, 256, false>' -s 40 -c 1 -o gpurun_out/prof_tc_dgrad_$tag -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_dgrad_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'post_physics_kernel' -s 60 -c 1 -o gpurun_out/prof_post_physics_$tag -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full_env_$tag.log 2>&1
ls -la gpurun_out/*$tag*
