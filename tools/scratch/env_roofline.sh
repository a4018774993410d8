#!/bin/bash
# roofline_env regimes of the current build (short bench run)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/env_roofline.log 2>&1
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/env_roofline.log") if l.startswith("{")][-1])
r = d["roofline_env"]
print("rotating", round(r["avg_us"], 2), round(r["frac"], 3), {k: (round(v.get("avg_us", v.get("us", 0)), 2), round(v["frac"], 3)) for k, v in r.items() if isinstance(v, dict) and "frac" in v})
print("ms_per_step", d["ms_per_step"], d["split"])
P
