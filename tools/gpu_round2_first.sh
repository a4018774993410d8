#!/bin/bash
# First GPU visit of the next round (nothing here could be run at the end of round 1: the GPU budget was spent).
#   1. compute-sanitizer memcheck over smoke() and the golden env replays, racecheck over the env kernel's shared-memory
#      phases (SURVEY.md §5 "race detection": absent in the reference, wanted here);
#   2. the minibatch timeline with the side-chain SM cap on (tools/trace_update.py) -- which kernels still starve;
#   3. default bench line (with cpu_baseline) for profiles/.
# Usage under gpurun (1 GPU, ~4 GPU-minutes):  bash tools/gpu_round2_first.sh r4a
tag=${1:-r4a}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/sanitize_memcheck_smoke_$tag.log \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_smoke_$tag.out 2>&1
echo "memcheck smoke rc=$?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/sanitize_memcheck_env_$tag.log \
  python -m pytest tests/test_env_gpu.py -q -x -k "golden" > gpurun_out/sanitize_env_$tag.out 2>&1
echo "memcheck env golden rc=$?"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/sanitize_racecheck_env_$tag.log \
  python -m pytest tests/test_env_gpu.py -q -x -k "golden and go2_parkour-go2-layout-baked-in" > gpurun_out/sanitize_race_$tag.out 2>&1
echo "racecheck env rc=$?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/sanitize_memcheck_learner_$tag.log \
  python -m pytest tests/test_learner_gpu.py -q -x -k "update_matches or dagger_matches or adaptive" > gpurun_out/sanitize_learner_$tag.out 2>&1
echo "memcheck learner rc=$?"
timeout 300 python tools/trace_update.py gpurun_out/minibatch_timeline_$tag.json > gpurun_out/minibatch_timeline_$tag.txt 2> gpurun_out/minibatch_timeline_$tag.err
timeout 600 python bench.py > gpurun_out/bench_${tag}_1gpu.json 2> gpurun_out/bench_${tag}.err
tail -c 600 gpurun_out/bench_${tag}_1gpu.json
