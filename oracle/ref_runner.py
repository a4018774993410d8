"""Run the UNMODIFIED reference (/root/reference) on CPU behind the fake isaacgym.

TEST INFRASTRUCTURE, authoring container only (/root/reference does not exist on the GPU
box).  Used by oracle/make_golden.py to produce tests/golden/*.npz and by
tests/test_oracle_vs_reference.py (skipped when /root/reference is absent).

What it does (recipe = SURVEY.md Appendix A):
  * puts oracle/refshim (fake isaacgym + empty matplotlib) and the reference on sys.path,
  * builds `Go2Robot` through the reference's own task_registry with sim_device='cpu'
    (= the reference's --sim_device=cpu --rl_device=cpu path, base_task.py:50-53),
  * replays synthetic PhysX frames through the stub's simulate / refresh hooks,
  * reroutes the reference's RNG draw sites to the keyed Philox function (oracle/keyed_rng.py)
    by wrapping -- not editing -- the methods that draw.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_INSTALLED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")      # baseline/install_ref.sh (travels to the GPU box)
REFERENCE_ROOT = os.environ.get("B200GYM_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/legged_gym") else _INSTALLED)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "legged_gym"))


def _setup_path():
    shim = os.path.join(_HERE, "refshim")
    for p in (os.path.join(REFERENCE_ROOT, "rsl_rl"), REFERENCE_ROOT, shim, os.path.dirname(_HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)


def make_args(task, num_envs, seed=1, max_iterations=1):
    from isaacgym import gymapi
    return argparse.Namespace(task=task, resume=False, experiment_name=None, run_name=None, load_run=None,
                              checkpoint=None, headless=True, horovod=False, rl_device='cpu', num_envs=num_envs,
                              seed=seed, max_iterations=max_iterations, physics_engine=gymapi.SIM_PHYSX,
                              sim_device='cpu', use_gpu=False, subscenes=0, use_gpu_pipeline=False, num_threads=0)


def build_env(task="go2_parkour", num_envs=32, seed=1, cfg_patch=None, keyed_rng=True):
    """-> (env, env_cfg, gym). The env is the reference's own Go2Robot instance.  `keyed_rng=False` leaves the reference's
    own torch RNG draws in place (timing runs: nothing is wrapped)."""
    _setup_path()
    import isaacgym  # noqa: F401  (the stub)
    from isaacgym import gymapi
    from legged_gym.envs import task_registry  # registers go2 / go2_parkour / go2_parkour_finetune
    args = make_args(task, num_envs, seed)
    env_cfg, _ = task_registry.get_cfgs(task)
    if cfg_patch is not None:
        cfg_patch(env_cfg)
    env, env_cfg = task_registry.make_env(task, args, env_cfg)
    if keyed_rng:
        install_keyed_rng(env)
    return env, env_cfg, gymapi._GYM


def build_training_run(task="go2_parkour", num_envs=4096, seed=1, ring=4, resume=None):
    """The reference's own training set-up on CPU (scripts/train.py:33-44 with --sim_device=cpu --rl_device=cpu): its
    Go2Robot from task_registry.make_env and its OnPolicyRunner from task_registry.make_alg_runner, nothing wrapped or
    edited; PhysX is the stub replaying a ring of `ring` synthetic frame sets (the same generator bench.py's GPU arm
    replays).  -> (runner, env); drive it with runner.learn(n)."""
    env, _, gym = build_env(task, num_envs, seed, keyed_rng=False)
    from legged_gym.envs import task_registry
    from legged_gym_custom_b200 import synth
    args = make_args(task, num_envs, seed)
    if resume is not None:
        _, train_cfg = task_registry.get_cfgs(task)
        train_cfg.runner.resume = False            # no checkpoint ships with the reference: random-init weights either way
    runner, _ = task_registry.make_alg_runner(env=env, name=task, args=args, log_root=temp_log_dir())
    rng = np.random.default_rng(seed)
    origins = env.env_origins.detach().cpu().numpy()
    frames = [{k: torch.from_numpy(v) for k, v in synth.make_frames(num_envs, origins, rng).items()} for _ in range(ring)]
    cur = {"frame": -1, "sub": 0}

    def on_sim(g):                                  # gym.simulate, once per decimation substep
        if cur["sub"] == 0:
            cur["frame"] = (cur["frame"] + 1) % ring
        g.dof.copy_(frames[cur["frame"]]["dof"][cur["sub"]])
        cur["sub"] = (cur["sub"] + 1) % frames[0]["dof"].shape[0]

    def on_root(g):                                 # gym.refresh_actor_root_state_tensor, top of post_physics_step
        f = frames[max(cur["frame"], 0)]
        g.root.copy_(f["root"]); g.contact.copy_(f["contact"]); g.rigid.copy_(f["rigid"])

    gym.on_simulate, gym.on_refresh_root = on_sim, on_root
    return runner, env


def install_keyed_rng(env):
    """Wrap the reference methods that draw random numbers so each draw is keyed by
    (site, common_step_counter, env id, lane).  Sites: SURVEY.md §8(c)."""
    from oracle import keyed_rng, philox
    state = {"in_reset": False}

    def ids(t):
        return t.detach().cpu().numpy().astype(np.int64).reshape(-1)

    all_ids = np.arange(env.num_envs)
    orig = {name: getattr(env, name) for name in
            ("reset_idx", "_resample_commands", "_push_robots", "_update_terrain_curriculum", "_reset_dofs",
             "_reset_root_states", "compute_observations")}

    def reset_idx(env_ids):
        state["in_reset"] = True
        try:
            return orig["reset_idx"](env_ids)
        finally:
            state["in_reset"] = False

    def resample(env_ids):
        s = philox.SITE_CMD_RESET if state["in_reset"] else philox.SITE_CMD_PERIODIC
        with keyed_rng.site(s, env.common_step_counter, ids(env_ids)):
            return orig["_resample_commands"](env_ids)

    def push():
        with keyed_rng.site(philox.SITE_PUSH, env.common_step_counter, all_ids):
            return orig["_push_robots"]()

    def curriculum(env_ids):
        with keyed_rng.site(philox.SITE_CURRICULUM, env.common_step_counter, ids(env_ids)):
            return orig["_update_terrain_curriculum"](env_ids)

    def reset_dofs(env_ids):
        with keyed_rng.site(philox.SITE_RESET_DOFS, env.common_step_counter, ids(env_ids)):
            return orig["_reset_dofs"](env_ids)

    def reset_root(env_ids):
        with keyed_rng.site(philox.SITE_RESET_ROOT, env.common_step_counter, ids(env_ids)):
            return orig["_reset_root_states"](env_ids)

    def observe():
        with keyed_rng.site(philox.SITE_OBS_NOISE, env.common_step_counter, all_ids):
            return orig["compute_observations"]()

    env.reset_idx = reset_idx
    env._resample_commands = resample
    env._push_robots = push
    env._update_terrain_curriculum = curriculum
    env._reset_dofs = reset_dofs
    env._reset_root_states = reset_root
    env.compute_observations = observe


def attach_frames(gym, frames):
    """Install replay hooks so the next env.step() consumes `frames` (synth.make_frames)."""
    k = {"i": 0}
    dof = torch.from_numpy(frames["dof"])

    def on_sim(g):
        g.dof.copy_(dof[k["i"]])
        k["i"] += 1

    def on_root(g):
        g.root.copy_(torch.from_numpy(frames["root"]))
        g.contact.copy_(torch.from_numpy(frames["contact"]))
        g.rigid.copy_(torch.from_numpy(frames["rigid"]))

    gym.on_simulate = on_sim
    gym.on_refresh_root = on_root


def np_(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().copy()
    return np.asarray(t)


PERSISTENT = ("commands", "episode_length_buf", "last_actions", "last_dof_vel", "last_root_vel",
              "last_base_lin_vel", "last_torques", "torques", "obs_history_buf", "last_contacts",
              "last_contact_heights", "jump_flags", "terrain_levels", "terrain_types", "env_origins",
              "reset_buf", "time_out_buf", "feet_air_time")
STATICS = ("kp_kd_multipliers", "privileged_mass_params", "privileged_friction_coeffs", "terrain_origins",
           "noise_scale_vec", "default_dof_pos", "torque_limits", "p_gains", "d_gains", "dof_pos_limits",
           "feet_indices", "penalised_contact_indices", "termination_contact_indices", "height_points")
DERIVED = ("base_lin_vel", "base_ang_vel", "projected_gravity", "roll", "pitch", "yaw", "measured_heights",
           "phase", "phase_fr", "phase_fl", "phase_bl", "phase_br", "fl_contact", "fr_contact", "bl_contact",
           "br_contact", "rew_buf", "obs_buf", "privileged_obs_buf", "critic_obs_buf", "estimated_obs_buf",
           "scan_obs_buf", "actions")


def snapshot(env, gym, names=PERSISTENT):
    """Persistent env state as numpy (what must survive between env steps)."""
    s = {k: np_(getattr(env, k)) for k in names if hasattr(env, k)}
    s["root_states"] = np_(gym.root)
    s["dof_state"] = np_(gym.dof)
    s["common_step_counter"] = np.int64(env.common_step_counter)
    for k, v in env.episode_sums.items():
        s["episode_sums/" + k] = np_(v)
    if "time_outs" in env.extras:
        s["extras/time_outs"] = np_(env.extras["time_outs"])
    for k, v in env.extras.get("episode", {}).items():
        s["extras/episode/" + k] = np_(v)
    return s


def statics(env):
    s = {k: np_(getattr(env, k)) for k in STATICS if hasattr(env, k)}
    if env.height_samples is not None:
        s["height_samples"] = np_(env.height_samples)
    s["reward_names"] = np.array(env.reward_names)
    s["reward_scales"] = np.array([env.reward_scales[n] for n in env.reward_names], dtype=np.float64)
    return s


def step_outputs(env, gym):
    o = {k: np_(getattr(env, k)) for k in DERIVED if hasattr(env, k)}
    o.update(snapshot(env, gym))
    return o


def temp_log_dir():
    return tempfile.mkdtemp(prefix="b200gym_ref_")
