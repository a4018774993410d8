for n in 16384 65536; do
timeout 400 python bench.py --num-envs $n --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/envs_$n.log 2>&1
python - gpurun_out/envs_$n.log <<'P'
import json, sys
line = [l for l in open(sys.argv[1]) if l.startswith("{")]
if not line: print("FAILED", open(sys.argv[1]).read()[-1500:])
else:
    d = json.loads(line[-1]); print(d["config"]["num_envs_per_gpu"], d["value"], json.dumps(d.get("roofline_env")))
P
done
