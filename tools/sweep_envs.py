"""BASELINE config 5: env-count sweep (go2_parkour, envs per GPU 1024 .. 65536) on one GPU; one bench.py line per size."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384, 32768, 65536]
rows = []
for n in sizes:
    steps = 3 if n <= 16384 else 2
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--num-envs", str(n), "--steps", str(steps), "--warmup", "3",
                          "--e2e-steps", "0", "--no-cpu-baseline"], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        rows.append({"num_envs": n, "error": out.stderr[-400:]})
        continue
    d = json.loads(line[-1])
    k = d["kernels"].get("b200_post_physics_step_dev") or d["kernels"].get("b200_post_physics_step") or {}
    re_ = d.get("roofline_env") or {}
    rows.append({"num_envs": n, "env_steps_per_s": d["value"], "ms_per_iteration": d["ms_per_step"], "split": d["split"],
                 "post_physics": {"kernel": re_.get("kernel"), "inputs_larger_than_l2_us": re_.get("avg_us"), "frac_of_hbm_peak": re_.get("frac"),
                                  "l2_resident_graph": re_.get("l2_resident_graph"), "l2_flushed_single_launch": re_.get("l2_flushed_single_launch"),
                                  "device_span": re_.get("device_span")},
                 "roofline": d["roofline"]})
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_envs.json"), "w"), indent=1)
