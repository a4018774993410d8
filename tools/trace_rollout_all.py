"""Kernel-level timeline of ONE replay of the rollout graph (every kernel, not only the GEMMs) through torch.profiler (CUPTI).
Prints env steps 10 and 11.  Development / evidence tool: CUPTI adds overhead per kernel, read the ORDER and relative sizes."""
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
if os.environ.get("B200_HOST_PHYSX") == "1":       # the end-to-end arm of bench.py: PhysX frames in pinned host memory
    from legged_gym_custom_b200.env import HostPhysX
    env.physx = HostPhysX(env.num_envs, env.bufs["env_origins"], DEV, seed=1234, decimation=env.params.decimation)
tc = class_to_dict(train_cfg)
tc["runner"]["resume"] = False
runner = OnPolicyRunner(env, tc, log_dir=None, device=DEV)
runner.enable_graphs()
runner.capture_graphs()
for it in range(3):
    runner.iteration(it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner.rollout(False)
    torch.cuda.synchronize()
runner.alg.storage.clear()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Event" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
posts = [i for i, e in enumerate(evs) if "post_physics" in e.name]
print(f"{len(evs)} device activities, {len(posts)} post_physics launches, span {(evs[-1].time_range.end - t0):.0f} us")
lo, hi = posts[9] + 1, posts[11] + 4
for e in evs[lo:hi]:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:9.1f} {e.time_range.end - e.time_range.start:6.1f}  {e.name[:90]}")
