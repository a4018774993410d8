"""ctypes mirror of include/b200gym.h (B200EnvParams / B200EnvBuffers) and the packing of a
reference-style cfg into it.

What the reference bakes at init (legged_robot.py:933-955 `_parse_cfg`, :344-357 soft dof
limits, :625-727 `_init_buffers`, :730-754 `_prepare_reward_function`, go2.py:110-129
`_get_noise_scale_vec`) becomes one POD struct that the kernels read from constant memory.
Python scalars are evaluated in double exactly as the reference does and stored as the fp32
value torch broadcasts them to.
"""
import ctypes as C

import numpy as np

NUM_DOF = 12
NUM_BODIES = 19
NUM_FEET = 4
MAX_SCAN_AXIS = 24
MAX_PROPRIO = 64
ABI_VERSION = 3

# go2.urdf with collapse_fixed_joints (Head_*/foot joints are dont_collapse) -- SURVEY.md §8(c)
BODY_NAMES = ["base", "Head_upper", "Head_lower"] + [
    f"{leg}_{part}" for leg in ("FL", "FR", "RL", "RR") for part in ("hip", "thigh", "calf", "foot")]
DOF_NAMES = [f"{leg}_{j}_joint" for leg in ("FL", "FR", "RL", "RR") for j in ("hip", "thigh", "calf")]
DOF_LOWER = [-1.0472, -1.5708, -2.7227] * 2 + [-1.0472, -0.5236, -2.7227] * 2
DOF_UPPER = [1.0472, 3.4907, -0.83776] * 2 + [1.0472, 4.5379, -0.83776] * 2
DOF_VELOCITY = [30.1, 30.1, 20.07] * 4
DOF_EFFORT = [23.7, 23.7, 35.55] * 4

# every `_reward_*` of the reference; sorted() == the order class_to_dict/dir() yields (helpers.py:45)
REWARD_TERMS = sorted([
    "lin_vel_z", "ang_vel_xy", "orientation", "base_height", "torques", "dof_vel", "dof_acc", "action_rate",
    "collision", "dof_pos_limits", "dof_vel_limits", "torque_limits", "tracking_lin_vel", "tracking_ang_vel",
    "stumble_feet", "stand_still", "feet_contact_forces", "delta_torques", "dof_error", "zero_cmd_dof_error",
    "hip_pos", "thigh_pos", "calf_pos", "phase_contact_match", "phase_foot_lifting", "stumble_calves",
    "calf_collision", "tracking_pitch", "tracking_roll", "thigh_symmetry", "calf_symmetry", "heading_alignment",
    "reverse_penalty", "jump_zone_forward_vel", "jump_zone_upward_vel", "min_height", "feet_air_time",
]) + ["termination"]
NUM_REWARD_TERMS = len(REWARD_TERMS)
REWARD_INDEX = {n: i for i, n in enumerate(REWARD_TERMS)}

f32, i32 = C.c_float, C.c_int32


class EnvParams(C.Structure):
    _fields_ = [
        ("abi_version", i32), ("num_envs", i32), ("num_proprio", i32), ("history_len", i32), ("num_priv", i32),
        ("num_est", i32), ("num_scan", i32), ("control_type", i32), ("randomize_kp_kd", i32), ("decimation", i32),
        ("sim_dt", f32), ("dt", f32), ("action_scale", f32), ("clip_actions", f32), ("clip_obs", f32),
        ("p_gains", f32 * NUM_DOF), ("d_gains", f32 * NUM_DOF), ("default_dof_pos", f32 * NUM_DOF),
        ("torque_limits", f32 * NUM_DOF), ("dof_pos_lo", f32 * NUM_DOF), ("dof_pos_hi", f32 * NUM_DOF),
        ("dof_vel_limits", f32 * NUM_DOF),
        ("max_episode_length", i32), ("max_episode_length_s", f32), ("resample_interval", i32),
        ("push_robots", i32), ("push_interval", i32), ("max_push_vel", f32),
        ("period", f32), ("fr_offset", f32), ("bl_offset", f32), ("fl_offset", f32), ("br_offset", f32),
        ("cmd_lo", f32 * 4), ("cmd_span", f32 * 4), ("heading_command", i32), ("heading_error_gain", f32),
        ("zero_command", i32), ("zero_command_prob", f32),
        ("has_height_samples", i32), ("hs_rows", i32), ("hs_cols", i32), ("border_size", f32),
        ("horizontal_scale", f32), ("vertical_scale", f32), ("index_div_mode", i32), ("parkour", i32),
        ("curriculum", i32), ("custom_origins", i32), ("promote_dist", f32), ("demote_threshold", f32),
        ("max_terrain_level", i32), ("terrain_cols", i32), ("scan_nx", i32), ("scan_ny", i32),
        ("scan_x", f32 * MAX_SCAN_AXIS), ("scan_y", f32 * MAX_SCAN_AXIS),
        ("base_init_state", f32 * 13), ("dof_reset_lo", f32), ("dof_reset_span", f32),
        ("obs_lin_vel", f32), ("obs_ang_vel", f32), ("obs_dof_pos", f32), ("obs_dof_vel", f32),
        ("add_noise", i32), ("noise_vec", f32 * MAX_PROPRIO),
        ("reward_scales", f32 * NUM_REWARD_TERMS), ("only_positive_rewards", i32),
        ("tracking_sigma", f32), ("base_height_target", f32), ("max_contact_force", f32), ("max_foot_height", f32),
        ("stance_threshold", f32), ("soft_dof_vel_limit", f32), ("soft_torque_limit", f32),
        ("pitch_deg_target", f32), ("roll_deg_target", f32),
        ("feet", i32 * NUM_FEET), ("calves", i32 * NUM_FEET), ("n_penalised", i32), ("penalised", i32 * NUM_BODIES),
        ("n_termination", i32), ("termination", i32 * NUM_BODIES),
        ("hip_joints", i32 * 4), ("thigh_joints", i32 * 4), ("calf_joints", i32 * 4),
        ("contact_thr2_term", f32), ("contact_thr2_collision", f32), ("alias_outputs", i32),
        ("seed", C.c_uint64),
        ("cc_vel_increment", C.c_double), ("cc_max_forward_vel", C.c_double), ("cc_max_reverse_vel", C.c_double),
        ("cc_range0", C.c_double * 2),
        ("cc_threshold", f32), ("command_curriculum", i32), ("hs_pitch", i32), ("terrain_tiles", i32),
    ]

    def reward_names(self):
        """active terms in summation order, `termination` excluded (legged_robot.py:745-750)."""
        return [n for n in REWARD_TERMS[:-1] if self.reward_scales[REWARD_INDEX[n]] != 0.0]


# field order must match B200EnvBuffers in include/b200gym.h
BUFFER_FIELDS = [
    "root_states", "dof_state", "contact_forces", "rigid_body_states", "kp_kd_multipliers", "priv_mass_params",
    "priv_friction", "height_samples", "terrain_origins", "actions", "torques", "commands", "episode_length_buf",
    "last_actions", "last_dof_vel", "last_root_vel", "last_base_lin_vel", "last_torques", "obs_history_buf",
    "last_contacts", "last_contact_heights", "feet_air_time", "jump_flags", "episode_sums", "terrain_levels",
    "terrain_types", "env_origins", "base_lin_vel", "base_ang_vel", "projected_gravity", "rpy", "measured_heights",
    "height_index", "phases", "foot_contacts", "obs_buf", "privileged_obs_buf", "critic_obs_buf",
    "estimated_obs_buf", "scan_obs_buf", "rew_buf", "reset_buf", "time_out_buf", "extras_time_outs",
    "extras_episode", "reset_count", "reset_episode_sums", "command_ranges", "cc_value", "cc_reset",
]


class EnvBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


class InitParams(C.Structure):
    """B200InitParams (include/b200gym.h): env-creation-time randomisation constants."""
    _fields_ = [("randomize_friction", C.c_int32), ("friction_lo", C.c_float), ("friction_hi", C.c_float), ("dynamic_friction", C.c_float),
                ("randomize_base_mass", C.c_int32), ("mass_lo", C.c_float), ("mass_hi", C.c_float),
                ("randomize_com", C.c_int32), ("com_lo", C.c_float), ("com_hi", C.c_float),
                ("kp_kd_lo", C.c_float), ("kp_kd_hi", C.c_float), ("num_init_levels", C.c_int32), ("terrain_cols", C.c_int32),
                ("env_spacing", C.c_float), ("grid_cols", C.c_int32)]


def init_params_from_cfg(cfg, num_envs, has_height_field):
    """legged_robot.py:306-380 (friction / mass / com ranges), :696-701 (kp_kd_range), :897-930 (origins)."""
    dr, ter = cfg.domain_rand, cfg.terrain
    ip = InitParams()
    ip.randomize_friction = int(bool(getattr(dr, "randomize_friction", False)))
    ip.friction_lo, ip.friction_hi = getattr(dr, "friction_range", (1.0, 1.0))
    ip.dynamic_friction = float(getattr(ter, "dynamic_friction", 1.0))
    ip.randomize_base_mass = int(bool(getattr(dr, "randomize_base_mass", False)))
    ip.mass_lo, ip.mass_hi = getattr(dr, "added_mass_range", (0.0, 0.0))
    ip.randomize_com = int(bool(getattr(dr, "randomize_center_of_mass", False)))
    ip.com_lo, ip.com_hi = getattr(dr, "added_com_range", (0.0, 0.0))
    ip.kp_kd_lo, ip.kp_kd_hi = getattr(dr, "kp_kd_range", (1.0, 1.0))
    if has_height_field:
        max_init = ter.max_init_terrain_level if ter.curriculum else ter.num_rows - 1
        ip.num_init_levels, ip.terrain_cols = int(max_init) + 1, int(ter.num_cols)
    else:
        ip.num_init_levels, ip.terrain_cols = 0, 0
    ip.env_spacing = float(getattr(cfg.env, "env_spacing", 3.0))
    ip.grid_cols = int(np.floor(np.sqrt(num_envs)))
    return ip


def sqrt_threshold_squared(t):
    """largest float32 s with sqrt_rn(s) <= t, so that  sqrt(s) > t  <=>  s > result  (exactly, for every float32 s)"""
    t = np.float32(t)
    s = np.float32(t * t)
    while np.sqrt(s, dtype=np.float32) <= t:
        s = np.nextafter(s, np.float32(np.inf), dtype=np.float32)
    while np.sqrt(s, dtype=np.float32) > t:
        s = np.nextafter(s, np.float32(-np.inf), dtype=np.float32)
    return float(s)


def _get(ns, name, default=None):
    return getattr(ns, name, default)


def _names_containing(keys, names):
    out = []
    for k in keys:
        out.extend(i for i, n in enumerate(names) if k in n)
    return out


def env_params_from_cfg(cfg, num_envs=None, seed=1234, index_div_mode=0, hs_shape=None, alias_outputs=False, terrain_tiles=False):
    """Pack a reference-style env cfg (class namespace or instance) into EnvParams.

    `hs_shape` = (rows, cols) of height_samples for heightfield/trimesh terrains.  `alias_outputs`: the observation outputs
    exist once, as rows of critic_obs_buf (include/b200gym.h).
    """
    p = EnvParams()
    p.alias_outputs = int(bool(alias_outputs))
    p.terrain_tiles = int(bool(terrain_tiles))
    env, ter, cmd, ctl, dr = cfg.env, cfg.terrain, cfg.commands, cfg.control, cfg.domain_rand
    rew, norm, noise = cfg.rewards, cfg.normalization, cfg.noise
    p.abi_version = ABI_VERSION
    p.num_envs = int(num_envs if num_envs is not None else env.num_envs)
    p.num_proprio, p.history_len = int(env.num_proprio), int(env.history_buffer_length)
    p.num_priv, p.num_est, p.num_scan = int(env.num_privileged_obs), int(env.num_estimated_obs), int(env.num_scan_obs)
    assert env.num_actions == NUM_DOF and p.num_proprio <= MAX_PROPRIO
    assert env.num_observations == p.num_proprio * (p.history_len + 1)
    p.control_type = {"P": 0, "V": 1, "T": 2}[ctl.control_type]
    p.randomize_kp_kd = int(bool(_get(dr, "randomize_kp_kd", False)))
    p.decimation = int(ctl.decimation)
    sim_dt = float(cfg.sim.dt)
    dt = ctl.decimation * sim_dt                         # legged_robot.py:946
    p.sim_dt, p.dt = sim_dt, dt
    p.action_scale, p.clip_actions, p.clip_obs = ctl.action_scale, norm.clip_actions, norm.clip_observations

    for i, name in enumerate(DOF_NAMES):                 # legged_robot.py:706-724, :344-357
        p.default_dof_pos[i] = cfg.init_state.default_joint_angles[name]
        for key in ctl.stiffness:
            if key in name:
                p.p_gains[i], p.d_gains[i] = ctl.stiffness[key], ctl.damping[key]
        p.torque_limits[i], p.dof_vel_limits[i] = DOF_EFFORT[i], DOF_VELOCITY[i]
        lo, hi = np.float32(DOF_LOWER[i]), np.float32(DOF_UPPER[i])
        m = np.float32(np.float32(lo + hi) / np.float32(2))
        r = np.float32(hi - lo)
        half = np.float32(np.float32(np.float32(0.5) * r) * np.float32(rew.soft_dof_pos_limit))
        p.dof_pos_lo[i], p.dof_pos_hi[i] = np.float32(m - half), np.float32(m + half)

    p.max_episode_length_s = env.episode_length_s
    p.max_episode_length = int(np.ceil(env.episode_length_s / dt))
    p.resample_interval = int(cmd.resampling_time / dt)
    p.push_robots = int(bool(_get(dr, "push_robots", False)))
    p.push_interval = int(np.ceil(_get(dr, "push_interval_s", 15) / dt))
    p.max_push_vel = _get(dr, "max_push_vel_xy", 1.0)
    p.period = _get(env, "period", 0.4)
    p.fr_offset, p.bl_offset = _get(env, "fr_offset", 0.0), _get(env, "bl_offset", 0.5)
    p.fl_offset, p.br_offset = _get(env, "fl_offset", 0.0), _get(env, "br_offset", 0.5)

    rng = cmd.ranges
    for i, key in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading")):
        lo, hi = getattr(rng, key)
        p.cmd_lo[i], p.cmd_span[i] = lo, hi - lo           # torch_rand_float: (upper - lower) * u + lower
    p.heading_command = int(bool(cmd.heading_command))
    p.heading_error_gain = _get(cmd, "heading_error_gain", 0.5)
    p.zero_command, p.zero_command_prob = int(bool(_get(cmd, "zero_command", False))), _get(cmd, "zero_command_prob", 0.1)
    assert len(_get(cmd, "user_command", [])) == 0, "user_command override is a play-time feature, not on the hot path"
    # command curriculum (go2.py:80-107): the thresholds are Python floats compared with / clipped against fp32 tensors
    p.command_curriculum = int(bool(_get(cmd, "curriculum", False)))
    p.cc_vel_increment = float(_get(cmd, "vel_increment", 0.0))
    p.cc_max_forward_vel, p.cc_max_reverse_vel = float(_get(cmd, "max_forward_vel", 0.0)), float(_get(cmd, "max_reverse_vel", 0.0))
    p.cc_range0[0], p.cc_range0[1] = float(rng.lin_vel_x[0]), float(rng.lin_vel_x[1])

    mesh = ter.mesh_type
    p.has_height_samples = int(mesh in ("heightfield", "trimesh"))
    p.custom_origins = p.has_height_samples
    if p.has_height_samples:
        assert hs_shape is not None
        p.hs_rows, p.hs_cols = int(hs_shape[0]), int(hs_shape[1])
        p.hs_pitch = (p.hs_cols + 7) // 8 * 8          # 16-byte rows: an env's scan neighbourhood is one 2-D TMA box
    p.border_size, p.horizontal_scale, p.vertical_scale = ter.border_size, ter.horizontal_scale, ter.vertical_scale
    p.index_div_mode = int(index_div_mode)
    p.parkour = int(bool(_get(ter, "parkour", False)))
    p.curriculum = int(bool(ter.curriculum) and p.has_height_samples)      # legged_robot.py:950-951
    p.promote_dist = ter.terrain_length * ter.promote_threshold
    p.demote_threshold = ter.demote_threshold
    p.max_terrain_level, p.terrain_cols = int(ter.num_rows), int(ter.num_cols)
    xs, ys = list(ter.measured_points_x), list(ter.measured_points_y)
    p.scan_nx, p.scan_ny = len(xs), len(ys)
    assert p.scan_nx * p.scan_ny == p.num_scan and max(p.scan_nx, p.scan_ny) <= MAX_SCAN_AXIS
    for i, v in enumerate(xs):
        p.scan_x[i] = v
    for i, v in enumerate(ys):
        p.scan_y[i] = v

    init = cfg.init_state
    for i, v in enumerate(list(init.pos) + list(init.rot) + list(init.lin_vel) + list(init.ang_vel)):
        p.base_init_state[i] = v
    p.dof_reset_lo, p.dof_reset_span = 0.0, 0.9 - 0.0

    sc = norm.obs_scales
    p.obs_lin_vel, p.obs_ang_vel, p.obs_dof_pos, p.obs_dof_vel = sc.lin_vel, sc.ang_vel, sc.dof_pos, sc.dof_vel
    p.add_noise = int(bool(noise.add_noise))
    nv = np.zeros(MAX_PROPRIO, dtype=np.float32)          # go2.py:115-127 (53-wide layout on a 52-wide vector)
    ns, lvl = noise.noise_scales, noise.noise_level
    nv[0:3] = ns.ang_vel * lvl * sc.ang_vel
    nv[3:5] = ns.imu * lvl
    nv[9:21] = ns.dof_pos * lvl * sc.dof_pos
    nv[21:33] = ns.dof_vel * lvl * sc.dof_vel
    nv[p.num_proprio:] = 0.0
    for i in range(MAX_PROPRIO):
        p.noise_vec[i] = nv[i]

    scales = {k: getattr(rew.scales, k) for k in dir(rew.scales) if not k.startswith("_")}
    for name, val in scales.items():
        if not isinstance(val, (int, float)) or val == 0:
            continue
        if name not in REWARD_INDEX:
            raise KeyError(f"reward scale '{name}' has no _reward_{name} in the reference")
        p.reward_scales[REWARD_INDEX[name]] = val * dt     # legged_robot.py:740
    p.only_positive_rewards = int(bool(rew.only_positive_rewards))
    p.cc_threshold = 0.8 * (_get(rew.scales, "tracking_lin_vel", 0.0) * dt)     # go2.py:91 (0.8 * the dt-scaled scale)
    p.tracking_sigma, p.base_height_target = rew.tracking_sigma, rew.base_height_target
    p.max_contact_force, p.max_foot_height = rew.max_contact_force, _get(rew, "max_foot_height", 0.08)
    p.stance_threshold = 2.0 * _get(rew, "percent_time_on_ground", 0.5) - 1.0
    p.soft_dof_vel_limit, p.soft_torque_limit = _get(rew, "soft_dof_vel_limit", 1.0), _get(rew, "soft_torque_limit", 1.0)
    p.pitch_deg_target, p.roll_deg_target = _get(rew, "pitch_deg_target", 0.0), _get(rew, "roll_deg_target", 0.0)

    asset = cfg.asset
    feet = [i for i, n in enumerate(BODY_NAMES) if asset.foot_name in n]
    calves = [i for i, n in enumerate(BODY_NAMES) if "calf" in n]
    pen = _names_containing(asset.penalize_contacts_on, BODY_NAMES)
    term = _names_containing(asset.terminate_after_contacts_on, BODY_NAMES)
    assert len(feet) == NUM_FEET and len(calves) == NUM_FEET
    for i in range(NUM_FEET):
        p.feet[i], p.calves[i] = feet[i], calves[i]
    p.n_penalised, p.n_termination = len(pen), len(term)
    for i, b in enumerate(pen):
        p.penalised[i] = b
    for i, b in enumerate(term):
        p.termination[i] = b
    for i, leg in enumerate(("FL", "FR", "RL", "RR")):
        p.hip_joints[i] = DOF_NAMES.index(f"{leg}_hip_joint")
        p.thigh_joints[i] = DOF_NAMES.index(f"{leg}_thigh_joint")
        p.calf_joints[i] = DOF_NAMES.index(f"{leg}_calf_joint")
    p.contact_thr2_term = sqrt_threshold_squared(1.0)
    p.contact_thr2_collision = sqrt_threshold_squared(0.1)
    p.seed = int(seed)
    return p
