"""Timeline of the tcgen05 GEMM launches of the rollout graph (24 env steps), from the kernels' own %globaltimer stamps
(b200_tc_set_trace).  Prints env steps 10 and 11.  Development / evidence tool."""
import ctypes as C
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
CAP = 1024
dev = torch.zeros(CAP, 2, dtype=torch.int64, device=DEV)
meta = np.zeros((CAP, 4), dtype=np.int64)
runner.enable_graphs()
runner.iteration(0)                    # DAgger iteration: adaptation-mode rollout (eager)
runner.iteration(1)                    # first normal rollout: eager
lib.b200_tc_set_trace(C.c_void_p(dev.data_ptr()), meta.ctypes.data_as(C.c_void_p), CAP)
runner.rollout(False)                  # captures (trace slots baked in) and replays
lib.b200_tc_set_trace(None, None, 0)
runner.alg.storage.clear()
torch.cuda.synchronize()
dev[:, 0] = torch.iinfo(torch.int64).max
dev[:, 1] = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
runner.rollout(False)
e1.record()
torch.cuda.synchronize()
t = dev.cpu().numpy()
rows = [dict(slot=i, M=int(meta[i, 0]), N=int(meta[i, 1]), K=int(meta[i, 2]), code=int(meta[i, 3]), start=int(t[i, 0]), end=int(t[i, 1]))
        for i in range(CAP) if t[i, 1] != 0]
t0 = min(r["start"] for r in rows)
per_step = len(rows) // 24
print(f"rollout graph: {e0.elapsed_time(e1) * 1e3:.0f} us by CUDA events for 24 env steps ({e0.elapsed_time(e1) * 1e3 / 24:.0f} us per step); "
      f"{len(rows)} tcgen05 GEMM launches ({per_step} per step), GEMM time {sum(r['end'] - r['start'] for r in rows) / 1e3:.0f} us in total")
for r in rows[10 * per_step:12 * per_step + 1]:
    bn = r["code"] % 1000
    print(f"{(r['start'] - t0) / 1e3:9.1f} {(r['end'] - t0) / 1e3:9.1f} {(r['end'] - r['start']) / 1e3:6.1f}  {r['M']:5d} x {r['N']:4d} x {r['K']:4d}  {'pair' if bn >= 500 else ''}{bn % 500}")
