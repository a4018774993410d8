"""Kernel-level timeline of ONE PPO minibatch graph replay (every kernel) through torch.profiler (CUPTI).  Development tool:
CUPTI adds overhead per kernel, read the ORDER and the relative sizes."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import configs  # noqa: E402
from legged_gym_custom_b200.env import Go2Env  # noqa: E402
from legged_gym_custom_b200.runner import OnPolicyRunner, class_to_dict  # noqa: E402

DEV = torch.device("cuda:0")
env_cfg, train_cfg = configs.TASKS["go2_parkour"]
env = Go2Env(env_cfg, sim_device="cuda:0", seed=1234)
tc = class_to_dict(train_cfg)
tc["runner"]["resume"] = False
runner = OnPolicyRunner(env, tc, log_dir=None, device=DEV)
runner.enable_graphs()
runner.capture_graphs()
for it in range(3):
    runner.iteration(it)
runner.rollout(False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner.alg.update()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Event" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
adv = [i for i, e in enumerate(evs) if "adam_advance" in e.name]
print(f"{len(evs)} device activities in one update, span {(evs[-1].time_range.end - evs[0].time_range.start):.0f} us; {len(adv)} adam_advance launches")
# minibatch 3 of epoch 0: between the 6th and the 8th adam_advance (two optimisers per minibatch)
lo, hi = adv[5] + 1, adv[7] + 1
t0 = evs[lo].time_range.start
for e in evs[lo:hi]:
    name = e.name.replace("void (anonymous namespace)::", "")
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - t0:8.1f} {e.time_range.end - e.time_range.start:6.1f}  {name[:70]}")
print("before the first minibatch:")
t0 = evs[0].time_range.start
for e in evs[:24]:
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - t0:8.1f} {e.time_range.end - e.time_range.start:6.1f}  {e.name[:70]}")
