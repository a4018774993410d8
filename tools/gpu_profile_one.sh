#!/bin/bash
# ncu --set full of one tcgen05 GEMM instantiation in situ: gpu_profile_one.sh <tag> <mode> <bn> <pair 0|1> <skip> <count>
tag=$1; mode=$2; bn=$3; pair=$4; skip=$5; cnt=$6
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-graphs"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_gemm_kernel<\(int\)$mode, \(int\)$bn, \(bool\)$pair>" -s $skip -c $cnt -o gpurun_out/prof_tc_${mode}_${bn}_${pair}_$tag -f $B > gpurun_out/ncu_one_$tag.log 2>&1
ls -la gpurun_out/prof_tc_${mode}_${bn}_${pair}_$tag.ncu-rep
