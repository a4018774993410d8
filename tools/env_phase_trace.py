"""Phase timeline of post_physics_kernel from its built-in %globaltimer trace (b200_env_set_phase_trace)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib, configs  # noqa: E402
from legged_gym_custom_b200.env import Go2Env  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


class Cfg(configs.Go2ParkourCfg):
    class env(configs.Go2ParkourCfg.env):
        num_envs = N


env = Go2Env(Cfg, sim_device="cuda:0", terrain_tiles=bool(int(os.environ.get("B200_TERRAIN_TILES", "0"))))
env.reset()
env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=1000)
actions = torch.randn(N, 12, device="cuda:0")
for _ in range(5):
    env.step(actions)
ctas = (N + 7) // 8
trace = torch.zeros(ctas, 8, dtype=torch.int64, device="cuda:0")
_lib.check(env.lib.b200_env_set_phase_trace(env._handle, C.c_void_p(trace.data_ptr())))
for rep in range(3):
    env.step(actions)
    torch.cuda.synchronize()
    t = trace.cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t[:, :6] - t0) / 1e3
    names = ["start", "A rows+scan", "E items", "B1 terms+hist ld", "B2 sum/reset+hist st", "C obs+write-back"]
    print(f"rep {rep}: kernel span {rel.max():.1f} us")
    for i, n in enumerate(names):
        print(f"   {n:22s} reached at: min {rel[:, i].min():6.1f}  median {np.median(rel[:, i]):6.1f}  max {rel[:, i].max():6.1f} us")
    d = np.median(rel[:, 1:6] - rel[:, 0:5], axis=0)
    print("   median phase durations (us, incl. the wait at the barrier before): " + "  ".join(f"{n.split()[0]} {x:.1f}" for n, x in zip(names[1:], d)))
_lib.check(env.lib.b200_env_set_phase_trace(env._handle, None))
