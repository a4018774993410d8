#!/bin/bash
# last GPU visit of round 1 (6 GPU-minutes left): new tests first, then the prefetch A/B, then as much of the suite as fits
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time timeout 120 python -u -m pytest tests/test_export.py tests/test_env_gpu.py -m gpu -q -k "deploy_policy or golden" ) > gpurun_out/r3a_new_tests.log 2>&1
tail -4 gpurun_out/r3a_new_tests.log
( time timeout 70 python -u tools/prefetch_ab.py --steps 30 ) > gpurun_out/r3a_prefetch_ab.log 2>&1
cat gpurun_out/r3a_prefetch_ab.log | grep -v Warning | tail -10
( time timeout 150 python -u -m pytest tests -m gpu -x -v --deselect tests/test_export.py --durations=10 ) > gpurun_out/r3a_gpu_suite.log 2>&1
tail -25 gpurun_out/r3a_gpu_suite.log
