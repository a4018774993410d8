#!/bin/bash
# Multi-GPU record of one round: bash tools/gpu_multi_round.sh <N> <tag> [sweep sizes...]
#   default task at K = 20 / W = 5 (the driver's window), BASELINE config 4 (go2_parkour_finetune), and the env-count sweep
#   points of config 5 given on the command line -- all launched the way the driver launches bench.py.
N=$1; tag=$2; shift 2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
run --steps 20 --warmup 5 > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err
run --steps 20 --warmup 5 --e2e-steps 0 --task go2_parkour_finetune > gpurun_out/bench_${tag}_${N}gpu_finetune.json 2> gpurun_out/bench_${tag}_${N}gpu_finetune.err
for n in "$@"; do
  run --steps 3 --warmup 3 --e2e-steps 0 --num-envs $n > gpurun_out/bench_${tag}_${N}gpu_envs$n.json 2> gpurun_out/bench_${tag}_${N}gpu_envs$n.err
done
grep -ho "\"value\": [0-9.]*, \"unit\": \"env-steps/s\", \"n_gpus\": [0-9]*, \"steps\": [0-9]*, \"warmup\": [0-9]*, \"ms_per_step\": [0-9.]*" gpurun_out/bench_${tag}_${N}gpu*.json
