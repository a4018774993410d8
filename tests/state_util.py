"""Random-but-plausible persistent env state shared by the oracle and the kernels in large-N tests."""
import numpy as np
import torch

from legged_gym_custom_b200.params import NUM_DOF, REWARD_TERMS
from oracle.go2_oracle import Go2Oracle


def random_statics(p, rng, height_samples=None, terrain_origins=None):
    N = p.num_envs
    return dict(kp_kd_multipliers=rng.uniform(0.8, 1.2, (2, N, NUM_DOF)).astype(np.float32),
                privileged_mass_params=np.concatenate([rng.uniform(0, 3, (N, 1)), rng.uniform(-0.15, 0.15, (N, 3))], 1).astype(np.float32),
                privileged_friction_coeffs=rng.uniform(0.1, 1.0, (N, 1)).astype(np.float32),
                height_samples=height_samples, terrain_origins=terrain_origins)


def random_state(p, rng, terrain_origins=None, step0=395):
    N = p.num_envs
    if terrain_origins is not None:
        levels = torch.from_numpy(rng.integers(0, p.max_terrain_level, N))
        types = torch.div(torch.arange(N), (N / p.terrain_cols), rounding_mode="floor").to(torch.long)
        origins = torch.from_numpy(terrain_origins)[levels, types]
    else:
        levels = types = None
        origins = torch.from_numpy(np.stack([rng.integers(0, 64, N) * 3.0, rng.integers(0, 64, N) * 3.0, np.zeros(N)], 1).astype(np.float32))
    st = Go2Oracle.fresh_state(p, origins, levels, types)
    f = lambda *s, scale=1.0: torch.from_numpy((rng.normal(0, scale, s)).astype(np.float32))
    ep = rng.integers(0, 1001, N)
    ep[: min(N, 8)] = [498, 499, 997, 999, 1000, 0, 1, 500][: min(N, 8)]
    st["episode_length_buf"] = torch.from_numpy(ep.astype(np.int64))
    cmd = np.zeros((N, 4), dtype=np.float32)
    cmd[:, 0] = rng.uniform(0.75, 1.5, N) * (rng.random(N) > 0.1)
    cmd[:, 2] = rng.uniform(-1, 1, N) * (cmd[:, 0] > 0)
    cmd[:, 3] = rng.uniform(-0.2, 0.2, N)
    st["commands"] = torch.from_numpy(cmd)
    st["actions"], st["torques"] = f(N, NUM_DOF), f(N, NUM_DOF, scale=5.0)
    st["last_actions"], st["last_dof_vel"], st["last_torques"] = f(N, NUM_DOF), f(N, NUM_DOF, scale=1.5), f(N, NUM_DOF, scale=5.0)
    st["last_root_vel"], st["last_base_lin_vel"] = f(N, 6, scale=0.5), f(N, 3, scale=0.5)
    st["obs_history_buf"] = f(N, p.history_len, p.num_proprio)
    st["last_contacts"] = torch.from_numpy(rng.random((N, 4)) < 0.5)
    st["last_contact_heights"] = torch.from_numpy(rng.uniform(0.0, 0.1, (N, 4)).astype(np.float32))
    st["feet_air_time"] = torch.from_numpy(rng.uniform(0.0, 0.4, (N, 4)).astype(np.float32))
    st["jump_flags"] = torch.from_numpy((rng.random((N, 1)) < 0.3).astype(np.float32))
    # inactive terms have no episode sum; an active term's sum carries the sign of its scale (no artificial cancellation)
    sign = torch.tensor([float(np.sign(p.reward_scales[i])) for i in range(len(REWARD_TERMS))])
    st["episode_sums"] = f(len(REWARD_TERMS), N, scale=0.3).abs() * sign[:, None]
    st["reset_buf"] = torch.zeros(N, dtype=torch.bool)
    st["common_step_counter"] = torch.tensor(step0, dtype=torch.int64)
    return st
