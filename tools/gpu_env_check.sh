#!/bin/bash
# dev helper: env parity tests + ncu timing of the env kernels (run under gpurun)
tag=$1
python -m pytest tests/test_env_gpu.py -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"post_physics|pd_torques|extras" -c 60 --csv --log-file gpurun_out/launches_env_$tag.csv python tools/bench_env.py --steps 10 --warmup 2 --no-flush > /dev/null 2>&1
python - <<PY
import csv,collections,re
lines=[l for l in open('gpurun_out/launches_env_$tag.csv') if l.startswith('"')]
agg=collections.defaultdict(list)
for d in csv.DictReader(lines):
    agg[(re.sub(r'\(.*','',d['Kernel Name'])[:40], d['Metric Name'])].append(float(d['Metric Value']))
for k,v in agg.items(): print(k, len(v), 'med %.1f min %.1f'%(sorted(v)[len(v)//2], min(v)))
PY
