#!/bin/bash
# ncu --set full of the dominant GEMM kernels in situ (inside bench.py's update), a few launches each
tag=$1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-graphs"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_gemm_kernel<\(int\)0, \(int\)256, \(bool\)1>' -s 8 -c 2 -o gpurun_out/prof_tc_fwd_pair256_$tag -f $B > gpurun_out/ncu_g1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_gemm_kernel<\(int\)1, \(int\)128, \(bool\)0>' -s 12 -c 4 -o gpurun_out/prof_tc_dgrad128_$tag -f $B > gpurun_out/ncu_g2_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tc_gemm_kernel<\(int\)2, \(int\)256, \(bool\)0>' -s 12 -c 3 -o gpurun_out/prof_tc_wgrad256_$tag -f $B > gpurun_out/ncu_g3_$tag.log 2>&1
ls -la gpurun_out/prof_*_$tag.ncu-rep; tail -2 gpurun_out/ncu_g1_$tag.log
