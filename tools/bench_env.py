"""Micro-benchmark of the env kernels alone (K1 x4 + K2 + extras) -- development tool; the judged
numbers come from bench.py.  Times with CUDA events on the launching stream, flushes L2 between
timed steps, reports achieved algorithmic GB/s (13.8 KB per env-step)."""
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
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--warmup", type=int, default=10)
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--prefetch", type=int, default=0, help="b200_env_set_prefetch")
args = ap.parse_args()


class Cfg(configs.Go2ParkourCfg):
    class env(configs.Go2ParkourCfg.env):
        num_envs = args.num_envs


env = Go2Env(Cfg, sim_device="cuda:0")
env.reset()
env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=1000)
N = args.num_envs
actions = torch.randn(N, 12, device="cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
lib, h, st = env.lib, env._handle, _lib.stream_ptr()
_lib.check(lib.b200_env_set_prefetch(h, args.prefetch))


def timed(fn, n, warm):
    ts = []
    for i in range(warm + n):
        if not args.no_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


res = {}
res["env_step_us(med,min)"] = timed(lambda: env.step(actions), args.steps, args.warmup)
b = env.bufs
res["pd_torques_us"] = timed(lambda: lib.b200_pd_torques(h, C.byref(b.struct), C.c_void_p(actions.data_ptr()), 1, st), args.steps, args.warmup)
step = [env.common_step_counter]


def post():
    step[0] += 1
    lib.b200_post_physics_step(h, C.byref(b.struct), step[0], st)


res["post_physics+extras_us"] = timed(post, args.steps, args.warmup)
alg = 13.8e3 * N
res["algorithmic_GBs_env_step(min)"] = alg / (res["env_step_us(med,min)"][1] * 1e-6) / 1e9
res["algorithmic_GBs_post_physics(min)"] = 12.618e3 * N / (res["post_physics+extras_us"][1] * 1e-6) / 1e9
res["num_envs"] = N
res["prefetch"] = args.prefetch
res["resets_last_step"] = int(b["reset_count"].item())
print(json.dumps(res))
