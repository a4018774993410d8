"""A/B of b200_env_set_prefetch on post_physics_kernel alone (development tool): one process, settings interleaved
(0, 1, 0, 1) so that clock / box drift cancels; CUDA events, L2 flushed before every launch and back to back."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib, configs  # noqa: E402
from legged_gym_custom_b200.env import Go2Env  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--num-envs", type=int, nargs="+", default=[4096, 65536])
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--warmup", type=int, default=5)
args = ap.parse_args()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")

for N in args.num_envs:
    class Cfg(configs.Go2ParkourCfg):
        class env(configs.Go2ParkourCfg.env):
            num_envs = N

    env = Go2Env(Cfg, sim_device="cuda:0")
    env.reset()
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=1000)
    env.step(torch.randn(N, 12, device="cuda:0"))
    lib, h, st, b = env.lib, env._handle, _lib.stream_ptr(), env.bufs
    step = [env.common_step_counter]

    def post():
        step[0] += 1
        lib.b200_post_physics_step_parts(h, C.byref(b.struct), step[0], 1, st)      # part 1 = post_physics_kernel only

    for rep in range(2):
        for pf in (0, 1):
            _lib.check(lib.b200_env_set_prefetch(h, pf))
            cold = []
            for i in range(args.warmup + args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                post()
                e1.record()
                torch.cuda.synchronize()
                if i >= args.warmup:
                    cold.append(e0.elapsed_time(e1) * 1e3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(5):
                post()
            e0.record()
            for _ in range(30):
                post()
            e1.record()
            torch.cuda.synchronize()
            warm = e0.elapsed_time(e1) * 1e3 / 30
            print(json.dumps({"num_envs": N, "prefetch": pf, "rep": rep, "cold_us_median": round(float(np.median(cold)), 2),
                              "cold_us_min": round(float(np.min(cold)), 2), "back_to_back_us": round(warm, 2),
                              "cold_frac_of_6555GBs": round(12618.0 * N / (float(np.median(cold)) * 1e-6) / 6555.2e9, 3)}), flush=True)
    del env
