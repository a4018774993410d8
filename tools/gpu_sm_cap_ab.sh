#!/bin/bash
# A/B of PPO.side_sm_cap (b200_tc_set_stream_sm_cap on the low-priority chains of a minibatch): one bench line per setting
cd "${GRAFT_REPO_ROOT:-.}"
# arguments: <cap> ...
for spec in "$@"; do
  cap=$spec
  tag=$spec
  timeout 90 python bench.py --no-cpu-baseline --e2e-steps 0 --side-sm-cap $cap > gpurun_out/bench_smcap_$tag.json 2> gpurun_out/bench_smcap_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_smcap_$tag.json").read().strip().splitlines()[-1])
    print("cap $spec:", round(d["value"]), "env-steps/s", round(d["ms_per_step"], 3), "ms", d["split"])
except Exception as e:
    print("cap $spec: failed", e, open("gpurun_out/bench_smcap_$tag.err").read()[-600:])
PY
done
