"""Shared replay helpers for the golden fixtures in tests/golden (made by oracle/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import torch

from legged_gym_custom_b200 import configs, terrain
from legged_gym_custom_b200.params import env_params_from_cfg, REWARD_TERMS, REWARD_INDEX, NUM_DOF, NUM_BODIES

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TASKS = ("go2_parkour", "go2_parkour_finetune", "go2")
# replays with cfg.commands.curriculum = True (oracle/make_golden.py CC_SCENARIOS); every helper below takes either a task
# name or one of these scenario names
CC_SCENARIOS = ("go2_parkour+cc_move", "go2+cc_move", "go2_parkour+cc_hold")
_terrain_cache = {}


def task_of(name):
    return name.split("+")[0]


def cfg_for(name):
    cfg = configs.TASKS[task_of(name)][0]
    if "+cc" in name:
        commands = type("commands", (cfg.commands,), {"curriculum": True})
        cfg = type(cfg.__name__ + "CommandCurriculum", (cfg,), {"commands": commands})
    return cfg


def load(task):
    return np.load(os.path.join(GOLDEN_DIR, f"env_{task}.npz"), allow_pickle=False)


def terrain_for(task):
    """(height_samples int16 | None, terrain_origins | None), regenerated and sha-pinned."""
    task = task_of(task)
    cfg = configs.TASKS[task][0]
    if cfg.terrain.mesh_type == "plane":
        return None, None
    if task not in _terrain_cache:
        hs, origins = terrain.make_parkour_terrain(cfg.terrain)
        with open(os.path.join(GOLDEN_DIR, "terrain_sha.json")) as f:
            sha = json.load(f)[task]
        assert hashlib.sha256(hs.tobytes()).hexdigest() == sha["sha256"]
        assert hashlib.sha256(origins.tobytes()).hexdigest() == sha["origins_sha256"]
        _terrain_cache[task] = (hs, origins)
    return _terrain_cache[task]


def params_for(task, g, index_div_mode=0):
    cfg = cfg_for(task)
    hs, _ = terrain_for(task)
    return env_params_from_cfg(cfg, num_envs=int(g["num_envs"]), seed=int(g["seed"]), index_div_mode=index_div_mode,
                               hs_shape=None if hs is None else hs.shape)


def statics_for(task, g):
    hs, origins = terrain_for(task)
    return dict(kp_kd_multipliers=g["static/kp_kd_multipliers"], privileged_mass_params=g["static/privileged_mass_params"],
                privileged_friction_coeffs=g["static/privileged_friction_coeffs"], height_samples=hs, terrain_origins=origins)


def init_state(g, p):
    """golden 'init/*' -> persistent-state dict with the oracle / B200EnvBuffers names."""
    N = p.num_envs
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g["init/" + k]))
    sums = torch.zeros(len(REWARD_TERMS), N)
    ep = torch.zeros(len(REWARD_TERMS) + 1)
    for i, name in enumerate(REWARD_TERMS):
        if "init/episode_sums/" + name in g.files:
            sums[i] = t("episode_sums/" + name)
        if "init/extras/episode/rew_" + name in g.files:
            ep[i] = float(g["init/extras/episode/rew_" + name])
    if "init/extras/episode/terrain_level" in g.files:
        ep[len(REWARD_TERMS)] = float(g["init/extras/episode/terrain_level"])
    st = dict(
        root_states=t("root_states"), dof_state=t("dof_state"), contact_forces=torch.zeros(N * NUM_BODIES, 3),
        rigid_body_states=torch.zeros(N * NUM_BODIES, 13), actions=t("actions"), torques=t("torques"),
        commands=t("commands"), episode_length_buf=t("episode_length_buf"), last_actions=t("last_actions"),
        last_dof_vel=t("last_dof_vel"), last_root_vel=t("last_root_vel"), last_base_lin_vel=t("last_base_lin_vel"),
        last_torques=t("last_torques"), obs_history_buf=t("obs_history_buf"), last_contacts=t("last_contacts").bool(),
        last_contact_heights=t("last_contact_heights"), feet_air_time=t("feet_air_time"), jump_flags=t("jump_flags"),
        episode_sums=sums, terrain_levels=t("terrain_levels") if "init/terrain_levels" in g.files else torch.zeros(N, dtype=torch.int64),
        terrain_types=t("terrain_types") if "init/terrain_types" in g.files else torch.zeros(N, dtype=torch.int64),
        env_origins=t("env_origins"), reset_buf=t("reset_buf").bool(), time_out_buf=t("time_out_buf").bool(),
        extras_time_outs=t("extras/time_outs").bool() if "init/extras/time_outs" in g.files else torch.zeros(N, dtype=torch.bool),
        extras_episode=ep, common_step_counter=torch.tensor(int(g["init/common_step_counter"]), dtype=torch.int64),
    )
    if "init/command_ranges" in g.files:
        st["command_ranges"] = t("command_ranges")
    return st


def frames_of(g, t):
    return {k: g[f"step{t}/in/{k}"] for k in ("dof", "root", "contact", "rigid")}


def expected(g, t):
    """dict name -> numpy of the reference's tensors after step t, mapped to our names."""
    pre = f"step{t}/out/"
    e = {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
    N = int(g["num_envs"])
    sums = np.zeros((len(REWARD_TERMS), N), dtype=np.float32)
    ep = {}
    for k, v in list(e.items()):
        if k.startswith("episode_sums/"):
            sums[REWARD_INDEX[k.split("/", 1)[1]]] = v
        if k.startswith("extras/episode/rew_"):
            ep[REWARD_INDEX[k[len("extras/episode/rew_"):]]] = float(v)
        if k == "extras/episode/terrain_level":
            ep[len(REWARD_TERMS)] = float(v)
    e["episode_sums"] = sums
    e["extras_episode"] = ep
    return e


def critic_sha(obs, priv, est, scan):
    c = np.concatenate([obs, priv, est, scan], axis=-1).astype(np.float32)
    return hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()


# ---- comparison of a BufferSet (host emulator or CUDA) with the reference's tensors ---------
RTOL = 1e-5          # BASELINE.json north_star: "within 1e-5 relative in fp32"
EXACT = ("episode_length_buf", "last_contacts", "terrain_levels", "reset_buf", "time_out_buf", "height_index",
         "foot_contacts", "jump_flags")
# golden name -> (buffer name, transform)
FLOAT_MAP = {
    "base_lin_vel": "base_lin_vel", "base_ang_vel": "base_ang_vel", "projected_gravity": "projected_gravity",
    "measured_heights": "measured_heights", "rew_buf": "rew_buf", "obs_buf": "obs_buf",
    "privileged_obs_buf": "privileged_obs_buf", "estimated_obs_buf": "estimated_obs_buf", "scan_obs_buf": "scan_obs_buf",
    "actions": "actions", "torques": "torques", "commands": "commands", "last_actions": "last_actions",
    "last_dof_vel": "last_dof_vel", "last_root_vel": "last_root_vel", "last_base_lin_vel": "last_base_lin_vel",
    "last_torques": "last_torques", "last_contact_heights": "last_contact_heights", "env_origins": "env_origins",
    "feet_air_time": "feet_air_time", "root_states": "root_states", "dof_state": "dof_state",
}


def rel_err(a, b, scale=None):
    """max |a-b| / max(|b|, s): relative to the element, floored at a scale s so that exact zeros and
    cancellations do not make the ratio meaningless.  s = the tensor's own rms (>= 1e-6) by default;
    for a sum of signed terms (rew_buf) pass `scale` = sum of |terms| per element, the magnitude its
    rounding errors are proportional to."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    s = max(float(np.sqrt(np.mean(b * b))), 1e-6) if scale is None else np.maximum(np.asarray(scale, dtype=np.float64), 1e-6)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), s)))


def check_step(bufs, exp, t, rtol=RTOL, report=None):
    """bufs: BufferSet after step t; exp: expected(g, t). Raises AssertionError on mismatch."""
    get = lambda name: bufs[name].detach().cpu().numpy()
    worst = {}
    for gname, bname in FLOAT_MAP.items():
        if gname not in exp:
            continue
        a, b = get(bname), exp[gname]
        assert a.shape == b.shape, (gname, a.shape, b.shape)
        assert np.isfinite(a).all(), f"{gname} step {t}: non-finite values"
        worst[gname] = rel_err(a, b, scale=exp.get("rew_terms_abs") if gname == "rew_buf" else None)
    rpy = get("rpy")
    for i, k in enumerate(("roll", "pitch", "yaw")):
        worst[k] = rel_err(rpy[:, i], exp[k])
    ph = get("phases")
    for i, k in enumerate(("phase", "phase_fr", "phase_fl", "phase_bl", "phase_br")):
        worst[k] = rel_err(ph[:, i], exp[k])
    worst["episode_sums"] = rel_err(get("episode_sums").T, exp["episode_sums"])
    crit = get("critic_obs_buf")
    worst["critic_obs_buf"] = rel_err(crit, np.concatenate([exp["obs_buf"], exp["privileged_obs_buf"], exp["estimated_obs_buf"],
                                                            exp["scan_obs_buf"]], axis=-1))
    ep = get("extras_episode")
    for i, v in exp["extras_episode"].items():
        worst[f"extras_episode[{i}]"] = rel_err(ep[i:i + 1], np.array([v]))
    if report is not None:
        for k, v in worst.items():
            report[k] = max(report.get(k, 0.0), v)
    bad = {k: v for k, v in worst.items() if not v <= rtol}
    assert not bad, f"step {t}: relative error above {rtol}: {bad}"
    # bit-exact: indices, masks, levels
    fc = get("foot_contacts")
    for i, k in enumerate(("fl_contact", "fr_contact", "bl_contact", "br_contact")):
        assert (fc[:, i].astype(bool) == exp[k].astype(bool)).all(), (k, t)
    for k in ("episode_length_buf", "terrain_levels"):
        if k in exp:
            assert (get(k) == exp[k]).all(), (k, t)
    for k in ("reset_buf", "time_out_buf", "last_contacts"):
        assert (get(k).astype(bool) == exp[k].astype(bool)).all(), (k, t)
    assert (get("jump_flags") == exp["jump_flags"]).all(), ("jump_flags", t)
    if bufs["height_index"] is not None:
        assert (get("height_index") == exp["height_index"]).all(), ("height_index", t)
    if "extras/time_outs" in exp:
        assert (get("extras_time_outs").astype(bool) == exp["extras/time_outs"].astype(bool)).all(), ("extras_time_outs", t)
    assert int(get("reset_count")[0]) == int(exp["n_reset"]), ("reset_count", t)
    if "command_ranges" in exp:        # Python floats in the reference, doubles here: identical
        assert (get("command_ranges")[2:4] == exp["command_ranges"]).all(), ("command_ranges", t, get("command_ranges"), exp["command_ranges"])


def oracle_expected(orc, out):
    """oracle tensors -> the dict shape golden_util.check_step expects."""
    st = orc.st
    e = {k: out[k].numpy() for k in ("base_lin_vel", "base_ang_vel", "projected_gravity", "roll", "pitch", "yaw",
                                     "measured_heights", "phase", "phase_fr", "phase_fl", "phase_bl", "phase_br",
                                     "fl_contact", "fr_contact", "bl_contact", "br_contact", "rew_buf", "obs_buf",
                                     "privileged_obs_buf", "estimated_obs_buf", "scan_obs_buf", "height_index")}
    for k in ("actions", "torques", "commands", "episode_length_buf", "last_actions", "last_dof_vel", "last_root_vel",
              "last_base_lin_vel", "last_torques", "last_contacts", "last_contact_heights", "jump_flags", "terrain_levels",
              "env_origins", "reset_buf", "time_out_buf", "feet_air_time", "root_states", "dof_state"):
        e[k] = st[k].numpy()
    e["episode_sums"] = st["episode_sums"].numpy()
    e["extras_episode"] = {i: float(v) for i, v in enumerate(st["extras_episode"].numpy())
                           if i == len(st["extras_episode"]) - 1 or orc.p.reward_scales[i] != 0.0}
    if not orc.p.curriculum:
        e["extras_episode"].pop(len(st["extras_episode"]) - 1, None)
    e["extras/time_outs"] = st["extras_time_outs"].numpy()
    e["n_reset"] = out["reset_count"]
    e["rew_terms_abs"] = out["rew_terms_abs"].numpy()
    if orc.p.command_curriculum:
        e["command_ranges"] = st["command_ranges"].numpy()
    return e


