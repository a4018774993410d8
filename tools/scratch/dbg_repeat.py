"""run-to-run spread of two learning iterations from the same seeds (split-K wgrads add with fp32 atomics): our runner, eager"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from legged_gym_custom_b200 import configs
from legged_gym_custom_b200.env import Go2Env
from legged_gym_custom_b200.runner import OnPolicyRunner, class_to_dict
DEV = "cuda:0"
env_cfg, train_cfg = configs.TASKS["go2_parkour"]
class Cfg(env_cfg):
    class env(env_cfg.env):
        pass
Cfg.env.num_envs = 256
tc = class_to_dict(train_cfg); tc["seed"] = 0
sds = []
for rep in range(6):
    env = Go2Env(Cfg, sim_device=DEV, seed=5)
    r = OnPolicyRunner(env, tc, log_dir=None, device=DEV)
    if rep >= 4:
        r.alg.defer_store = False; env.extras_stream = None
    for it in range(2):
        r.iteration(it)
    torch.cuda.synchronize()
    sds.append({k: v.clone() for k, v in r.alg.actor_critic.state_dict().items()})
for i in range(1, 6):
    worst = max((float((sds[0][k] - sds[i][k]).abs().max()), k) for k in sds[0])
    rms = max((float((sds[0][k] - sds[i][k]).pow(2).mean().sqrt()), k) for k in sds[0])
    print("run 0 vs run", i, "(no deferral)" if i >= 4 else "", "max |diff|", worst, "worst per-tensor rms", rms)
