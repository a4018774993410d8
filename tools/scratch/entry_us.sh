#!/bin/bash
# eager per-call time of one ABI entry from a short bench run: entry_us.sh <entry-substring>
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/entry_us.log 2>&1
python - "$1" <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/entry_us.log") if l.startswith("{")][-1])
for n, k in d["kernels"].items():
    if sys.argv[1] in n: print(n, k, "->", round(1e3 * k["ms"] / k["calls"], 2), "us per call")
print("ms_per_step", d["ms_per_step"])
P
