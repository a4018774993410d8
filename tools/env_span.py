"""post_physics_kernel device span (its own %globaltimer trace) at several env counts -- development tool."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib, configs  # noqa: E402
from legged_gym_custom_b200.env import Go2Env  # noqa: E402

for N in [int(x) for x in sys.argv[1:]] or [4096, 65536]:
    class Cfg(configs.Go2ParkourCfg):
        class env(configs.Go2ParkourCfg.env):
            num_envs = N
    env = Go2Env(Cfg, sim_device="cuda:0")
    env.reset()
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=1000)
    actions = torch.randn(N, 12, device="cuda:0")
    for _ in range(5):
        env.step(actions)
    trace = torch.zeros((N + 7) // 8, 8, dtype=torch.int64, device="cuda:0")
    _lib.check(env.lib.b200_env_set_phase_trace(env._handle, C.c_void_p(trace.data_ptr())))
    spans = []
    for _ in range(8):
        env.step(actions)
        torch.cuda.synchronize()
        t = trace.cpu().numpy()
        spans.append((t[:, 5].max() - t[:, 0].min()) / 1e3)
    _lib.check(env.lib.b200_env_set_phase_trace(env._handle, None))
    us = float(np.median(spans))
    print(f"N={N}: span {us:.1f} us  {12618 * N / us / 1e3:.0f} GB/s  {12618 * N / us / 1e3 / 6555.2:.3f} of HBM peak", flush=True)
    del env
