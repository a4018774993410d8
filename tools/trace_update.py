"""Timeline of the tcgen05 GEMM launches of ONE PPO minibatch as the GPU actually ran them inside its CUDA graph (streams
included), from the kernels' own %globaltimer stamps (b200_tc_set_trace).  Development / evidence tool.
usage: python tools/trace_update.py [out.json]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib, configs  # noqa: E402
from legged_gym_custom_b200.env import Go2Env  # noqa: E402
from legged_gym_custom_b200.runner import OnPolicyRunner, class_to_dict  # noqa: E402

DEV = torch.device("cuda:0")
env_cfg, train_cfg = configs.TASKS["go2_parkour"]
env = Go2Env(env_cfg, sim_device="cuda:0", seed=1234)
tc = class_to_dict(train_cfg)
tc["runner"]["resume"] = False
runner = OnPolicyRunner(env, tc, log_dir=None, device=DEV)
lib = _lib.lib()
CAP = 512
dev = torch.zeros(CAP, 2, dtype=torch.int64, device=DEV)
meta = np.zeros((CAP, 4), dtype=np.int64)
alg = runner.alg
runner.enable_graphs()
runner.iteration(0)                    # DAgger iteration (eager + allocations)
# the minibatch graphs are captured during the first PPO update (epoch 0 eager, epoch 1 captures): bake the trace slots in,
# slot numbering restarting for every minibatch
orig = alg._minibatch


def traced(r0, M):
    lib.b200_tc_set_trace(C.c_void_p(dev.data_ptr()), meta.ctypes.data_as(C.c_void_p), CAP)      # restart at slot 0
    return orig(r0, M)


alg._minibatch = traced
runner.iteration(1)                    # captures (trace pointers baked into the nodes) and replays
lib.b200_tc_set_trace(None, None, 0)   # the rollout's GEMMs are not traced
runner.iteration(2)
runner.iteration(3)
torch.cuda.synchronize()
# one clean replay of minibatch slot 0
dev[:, 0] = torch.iinfo(torch.int64).max
dev[:, 1] = 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
alg._graphs[("ppo", 0)].replay()
e1.record()
torch.cuda.synchronize()
t = dev.cpu().numpy()
rows = []
for i in range(CAP):
    if t[i, 1] == 0:
        continue
    m = meta[i]
    mode = ["fwd", "dgrad", "wgrad"][int(m[3]) // 1000]
    bn = int(m[3]) % 1000
    rows.append(dict(slot=i, kind=mode, M=int(m[0]), N=int(m[1]), K=int(m[2]), tile=("pair" if bn >= 500 else "") + str(bn % 500),
                     start=int(t[i, 0]), end=int(t[i, 1])))
t0 = min(r["start"] for r in rows)
for r in rows:
    r["start_us"], r["end_us"] = (r["start"] - t0) / 1e3, (r["end"] - t0) / 1e3
    r["us"] = r["end_us"] - r["start_us"]
rows.sort(key=lambda r: r["start_us"])
span = max(r["end_us"] for r in rows)
# union of busy intervals and sum of durations
ev = sorted([(r["start_us"], 1) for r in rows] + [(r["end_us"], -1) for r in rows])
busy, depth, last, conc = 0.0, 0, 0.0, 0.0
for x, d in ev:
    if depth > 0:
        busy += x - last
        conc += (x - last) * depth
    depth += d
    last = x
print(f"minibatch graph: {e0.elapsed_time(e1) * 1e3:.0f} us by CUDA events; {len(rows)} tcgen05 GEMM launches spanning {span:.0f} us; "
      f"at least one GEMM running {busy:.0f} us; sum of GEMM durations {sum(r['us'] for r in rows):.0f} us (mean concurrency {conc / busy:.2f})")
print(f"{'start':>8} {'end':>8} {'us':>7}  kind   out-rows x out-cols x reduction  tile")
for r in rows:
    print(f"{r['start_us']:8.1f} {r['end_us']:8.1f} {r['us']:7.1f}  {r['kind']:5s}  {r['M']:6d} x {r['N']:4d} x {r['K']:6d}  {r['tile']}")
if len(sys.argv) > 1:
    json.dump(dict(events_us=e0.elapsed_time(e1) * 1e3, span_us=span, busy_us=busy, rows=[{k: v for k, v in r.items() if k not in ("start", "end")} for r in rows]),
              open(sys.argv[1], "w"), indent=1)
