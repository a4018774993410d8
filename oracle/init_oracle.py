"""Env-creation-time domain randomisation, restated.

TEST INFRASTRUCTURE (oracle/): imported only by tests/.

Follows LeggedRobot._process_rigid_shape_props (legged_robot.py:306-329: 64 friction buckets ~ U(range), every env picks
one), _process_rigid_body_props (:358-380: added base mass, centre-of-mass shift), the kp/kd multipliers of _init_buffers
(:696-701: U(range) of shape [2, N, 12]) and _get_env_origins (:897-930: initial terrain level ~ randint(0, max_init + 1),
type = floor(i / (N / num_cols)), origin = terrain_origins[level, type]; or the env_spacing grid on a plane).  The
reference draws from torch's global stream; like every other draw of this repo the kernels use keyed Philox
(oracle/philox.py), sites 8-12 with env = index and step = 0, so the restatement is bit-exact against them.
"""
import numpy as np

from . import philox as px

SITE_INIT_FRICTION_BUCKET, SITE_INIT_FRICTION_PICK, SITE_INIT_MASS, SITE_INIT_KPKD, SITE_INIT_LEVEL = 8, 9, 10, 11, 12
f32 = np.float32


def init_randomisation(seed, num_envs, ip, terrain_origins=None):
    """ip: an object with the B200InitParams fields.  Returns the six tensors the kernel writes."""
    N = num_envs
    env = np.arange(N)
    out = {}
    if ip.randomize_friction:
        buckets = (f32(ip.friction_hi) - f32(ip.friction_lo)) * px.keyed_uniform(seed, SITE_INIT_FRICTION_BUCKET, 0, np.arange(64), [0])[:, 0] \
            + f32(ip.friction_lo)
        pick = px.keyed_u32(seed, SITE_INIT_FRICTION_PICK, 0, env, [0])[:, 0] % np.uint32(64)
        out["priv_friction"] = buckets[pick].astype(f32).reshape(N, 1)
    else:
        out["priv_friction"] = np.full((N, 1), ip.dynamic_friction, f32)
    u = px.keyed_uniform(seed, SITE_INIT_MASS, 0, env, [0, 1, 2, 3])
    mass = np.zeros((N, 4), f32)
    if ip.randomize_base_mass:
        mass[:, 0] = (f32(ip.mass_hi) - f32(ip.mass_lo)) * u[:, 0] + f32(ip.mass_lo)
    if ip.randomize_com:
        mass[:, 1:] = (f32(ip.com_hi) - f32(ip.com_lo)) * u[:, 1:] + f32(ip.com_lo)
    out["priv_mass_params"] = mass
    u = px.keyed_uniform(seed, SITE_INIT_KPKD, 0, env, np.arange(24))
    kpkd = (f32(ip.kp_kd_hi) - f32(ip.kp_kd_lo)) * u + f32(ip.kp_kd_lo)
    out["kp_kd_multipliers"] = np.stack([kpkd[:, :12], kpkd[:, 12:]]).astype(f32)
    if ip.num_init_levels > 0:
        levels = (px.keyed_u32(seed, SITE_INIT_LEVEL, 0, env, [0])[:, 0] % np.uint32(ip.num_init_levels)).astype(np.int64)
        # torch.div(arange(N), N / num_cols, rounding_mode='floor') (legged_robot.py:909) is fmod-based floor division in fp32
        # (ATen div_floor_floating): (a - fmod(a, b)) / b, floored with a half-ulp guard -- not floor(a / b)
        a, b = env.astype(f32), f32(N / ip.terrain_cols)
        mod = np.fmod(a, b).astype(f32)
        div = ((a - mod).astype(f32) / b).astype(f32)
        div = np.where((mod != 0) & ((b < 0) != (mod < 0)), div - f32(1), div).astype(f32)
        fl = np.floor(div).astype(f32)
        fl = np.where(div - fl > f32(0.5), fl + f32(1), fl)
        types = fl.astype(np.int64)
        out["terrain_levels"], out["terrain_types"] = levels, types
        out["env_origins"] = np.asarray(terrain_origins, f32)[levels, types]
    else:
        o = np.zeros((N, 3), f32)
        o[:, 0] = f32(ip.env_spacing) * (env // ip.grid_cols).astype(f32)
        o[:, 1] = f32(ip.env_spacing) * (env % ip.grid_cols).astype(f32)
        out["env_origins"] = o
        out["terrain_levels"], out["terrain_types"] = np.zeros(N, np.int64), np.zeros(N, np.int64)
    return out
