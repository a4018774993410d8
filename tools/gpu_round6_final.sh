#!/bin/bash
# Final 1-GPU record of round 2 (tag r6a): GPU test suite, smoke, the driver's bench line, the reference arm, the ncu launch list
# of the same bench command, and ncu --set full of the dominant GEMM and the env kernel (DRAM traffic per launch).
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r6a_gpu_suite.txt; cat gpurun_out/r6a_gpu_suite.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r6a_bench_1gpu.json 2> gpurun_out/r6a_bench_1gpu.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r6a_bench_reference_arm.json 2> gpurun_out/r6a_ref.err; echo "ref rc $?"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-graphs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches_r6a.csv $B > gpurun_out/ncu_launches_r6a.log 2>&1
wc -l gpurun_out/launches_r6a.csv
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_gemm_kernel<\(int\)1, \(int\)256, \(bool\)0, \(int\)1, \(bool\)1>" -s 4 -c 1 -o gpurun_out/prof_r6a_dgrad_epi256 -f $B > gpurun_out/ncu_r6a_dgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post_physics_tile_kernel -s 30 -c 1 -o gpurun_out/prof_r6a_env_tile -f $B > gpurun_out/ncu_r6a_env.log 2>&1
ls -la gpurun_out/prof_r6a_* 2>&1 | tail -3
