#!/bin/bash
# A/B of environment-variable switches on one box: ab_env.sh "VAR=1 VAR2=0" "..."  (one bench.py run per argument)
i=0
for envs in "$@"; do
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/abe_$i.log 2>&1
  python - "$envs" gpurun_out/abe_$i.log <<'P'
import json, sys
line = [l for l in open(sys.argv[2]) if l.startswith("{")]
if not line:
    print(sys.argv[1], "FAILED"); print(open(sys.argv[2]).read()[-1500:])
else:
    d = json.loads(line[-1])
    print(f"[{sys.argv[1]}] ms_per_step {d['ms_per_step']:.3f} value {d['value']:.0f} split {d.get('split')}")
P
  i=$((i+1))
done
