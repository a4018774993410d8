"""The CUDA kernel SOURCE (csrc/env_core.cuh), compiled for the host and run lane by lane, against
the reference's golden tensors: bit-exact indices / masks / levels, <= 1e-5 relative floats.
The same checks run against the real kernels on the GPU in test_env_gpu.py."""
import ctypes as C

import numpy as np
import pytest
import torch

import golden_util as gu
from host_emul import emul
from legged_gym_custom_b200.buffers import BufferSet


@pytest.fixture(scope="module")
def lib():
    return emul.load()


@pytest.mark.parametrize("force_generic,alias", [(0, 0), (1, 0), (0, 1)], ids=["go2-layout-baked-in", "layout-generic", "aliased-output-rows"])
@pytest.mark.parametrize("task", gu.TASKS + gu.CC_SCENARIOS)
def test_kernel_source_matches_reference(lib, task, force_generic, alias):
    g = gu.load(task)
    p = gu.params_for(task, g)
    p.alias_outputs = alias            # obs / priv / est / scan are column slices of the critic rows, written once
    bufs = BufferSet(p, "cpu", record_height_index=True)
    bufs.load_statics(gu.statics_for(task, g))
    st = gu.init_state(g, p)
    step = int(st.pop("common_step_counter"))
    bufs.load_state(st)
    report = {}
    for t in range(int(g["steps"])):
        fr = gu.frames_of(g, t)
        actions = torch.from_numpy(g[f"step{t}/in/actions"]).contiguous()
        for k in range(p.decimation):
            lib.emul_pd_torques(C.byref(p), C.byref(bufs.struct), C.c_void_p(actions.data_ptr()), int(k == 0))
            bufs["dof_state"].copy_(torch.from_numpy(fr["dof"][k]))
        bufs["root_states"].copy_(torch.from_numpy(fr["root"]))
        bufs["contact_forces"].copy_(torch.from_numpy(fr["contact"]))
        bufs["rigid_body_states"].copy_(torch.from_numpy(fr["rigid"]))
        step += 1
        lib.emul_post_physics_step_variant(C.byref(p), C.byref(bufs.struct), step, force_generic)
        gu.check_step(bufs, gu.expected(g, t), t, report=report)
    hist = bufs["obs_history_buf"].numpy()
    assert gu.rel_err(hist, g["final/obs_history_buf"]) <= gu.RTOL
    print(task, "worst relative errors:", {k: f"{v:.1e}" for k, v in sorted(report.items(), key=lambda kv: -kv[1])[:6]})


def test_kernel_source_every_reward_term_active(lib):
    """all 38 reward terms switched on at once (including the stateful feet_air_time and the termination reward), on a
    rough height field with curriculum: the oracle and the kernel source must still agree."""
    import state_util as su
    from legged_gym_custom_b200 import configs, synth
    from legged_gym_custom_b200.params import NUM_DOF, REWARD_TERMS, env_params_from_cfg
    from oracle.go2_oracle import Go2Oracle

    class AllCfg(configs.Go2ParkourCfg):
        class rewards(configs.Go2ParkourCfg.rewards):
            only_positive_rewards = False
            soft_dof_vel_limit, soft_torque_limit = 0.05, 0.3

            class scales:
                pass
    for i, name in enumerate(REWARD_TERMS):
        setattr(AllCfg.rewards.scales, name, (-1.0) ** i * (0.3 + 0.1 * i))
    hs, origins = gu.terrain_for("go2_parkour")
    N = 400
    p = env_params_from_cfg(AllCfg, num_envs=N, seed=5, hs_shape=hs.shape)
    assert len(p.reward_names()) == len(REWARD_TERMS) - 1 and p.reward_scales[len(REWARD_TERMS) - 1] != 0
    rng = np.random.default_rng(3)
    statics, st = su.random_statics(p, rng, hs, origins), su.random_state(p, rng, origins)
    orc = Go2Oracle(p, statics, st)
    bufs = BufferSet(p, "cpu", record_height_index=True)
    bufs.load_statics(statics)
    st2 = dict(st)
    step = int(st2.pop("common_step_counter"))
    bufs.load_state(st2)
    for t in range(3):
        frames = synth.make_frames(N, st["env_origins"].numpy(), rng, hole_prob=0.02, flip_prob=0.02, body_hit_prob=0.05)
        actions = torch.from_numpy(rng.normal(0, 1.5, (N, NUM_DOF)).astype(np.float32))
        out = orc.step(actions, frames)
        for k in range(p.decimation):
            lib.emul_pd_torques(C.byref(p), C.byref(bufs.struct), C.c_void_p(actions.data_ptr()), int(k == 0))
            bufs["dof_state"].copy_(torch.from_numpy(frames["dof"][k]))
        for name, key in (("root_states", "root"), ("contact_forces", "contact"), ("rigid_body_states", "rigid")):
            bufs[name].copy_(torch.from_numpy(frames[key]))
        step += 1
        lib.emul_post_physics_step(C.byref(p), C.byref(bufs.struct), step)
        gu.check_step(bufs, gu.oracle_expected(orc, out), t)


@pytest.mark.parametrize("task,num_envs,steps", [("go2_parkour", 1500, 3), ("go2", 333, 2), ("go2_parkour", 1, 2), ("go2_parkour", 9, 2)])
def test_kernel_source_matches_oracle_at_scale(lib, task, num_envs, steps):
    """same harness as tests/test_env_gpu.py::test_cuda_env_matches_oracle_at_scale, on the host emulator."""
    import state_util as su
    from legged_gym_custom_b200 import configs, synth
    from legged_gym_custom_b200.params import NUM_DOF, env_params_from_cfg
    from oracle.go2_oracle import Go2Oracle
    cfg = configs.TASKS[task][0]
    hs, origins = gu.terrain_for(task)
    p = env_params_from_cfg(cfg, num_envs=num_envs, seed=99, hs_shape=None if hs is None else hs.shape)
    rng = np.random.default_rng(7)
    statics = su.random_statics(p, rng, hs, origins)
    st = su.random_state(p, rng, origins)
    orc = Go2Oracle(p, statics, st)
    bufs = BufferSet(p, "cpu", record_height_index=True)
    bufs.load_statics(statics)
    st2 = dict(st)
    step = int(st2.pop("common_step_counter"))
    bufs.load_state(st2)
    origins0 = st["env_origins"].numpy()
    total = 0
    for t in range(steps):
        frames = synth.make_frames(num_envs, origins0, rng, hole_prob=0.01, flip_prob=0.01)
        actions = torch.from_numpy(rng.normal(0, 1.5, (num_envs, NUM_DOF)).astype(np.float32))
        out = orc.step(actions, frames)
        for k in range(p.decimation):
            lib.emul_pd_torques(C.byref(p), C.byref(bufs.struct), C.c_void_p(actions.data_ptr()), int(k == 0))
            bufs["dof_state"].copy_(torch.from_numpy(frames["dof"][k]))
        bufs["root_states"].copy_(torch.from_numpy(frames["root"]))
        bufs["contact_forces"].copy_(torch.from_numpy(frames["contact"]))
        bufs["rigid_body_states"].copy_(torch.from_numpy(frames["rigid"]))
        step += 1
        lib.emul_post_physics_step(C.byref(p), C.byref(bufs.struct), step)
        gu.check_step(bufs, gu.oracle_expected(orc, out), t)
        total += out["reset_count"]
    assert total > 0 or num_envs < 64


@pytest.mark.parametrize("flag", ["noise.add_noise", "rewards.only_positive_rewards", "domain_rand.push_robots", "commands.zero_command"])
def test_kernel_source_with_cfg_switches_flipped(lib, flag):
    """the cfg switches every shipped go2 cfg leaves at one value, flipped: observation noise off, negative total rewards
    kept (legged_robot.py:230-231), no pushes, no zero-command draws (go2.py:450-456).  5 steps from counter 397, so the
    push step (400) is inside the window.  Oracle (pinned to the reference for the shipped values) vs kernel source."""
    import state_util as su
    from legged_gym_custom_b200 import configs, synth
    from legged_gym_custom_b200.params import NUM_DOF, env_params_from_cfg
    from oracle.go2_oracle import Go2Oracle
    base = configs.TASKS["go2_parkour"][0]
    group, name = flag.split(".")
    sub = getattr(base, group)
    cfg = type("Cfg", (base,), {group: type(group, (sub,), {name: not getattr(sub, name)})})
    hs, origins = gu.terrain_for("go2_parkour")
    N = 300
    p = env_params_from_cfg(cfg, num_envs=N, seed=17, hs_shape=hs.shape)
    assert getattr(p, {"add_noise": "add_noise", "only_positive_rewards": "only_positive_rewards", "push_robots": "push_robots",
                       "zero_command": "zero_command"}[name]) == int(not getattr(sub, name))
    rng = np.random.default_rng(23)
    statics, st = su.random_statics(p, rng, hs, origins), su.random_state(p, rng, origins, step0=397)
    orc = Go2Oracle(p, statics, st)
    bufs = BufferSet(p, "cpu", record_height_index=True)
    bufs.load_statics(statics)
    st2 = dict(st)
    step = int(st2.pop("common_step_counter"))
    bufs.load_state(st2)
    origins0 = st["env_origins"].numpy()
    negative = 0
    for t in range(5):
        frames = synth.make_frames(N, origins0, rng, hole_prob=0.01, flip_prob=0.01)
        actions = torch.from_numpy(rng.normal(0, 1.5, (N, NUM_DOF)).astype(np.float32))
        out = orc.step(actions, frames)
        for k in range(p.decimation):
            lib.emul_pd_torques(C.byref(p), C.byref(bufs.struct), C.c_void_p(actions.data_ptr()), int(k == 0))
            bufs["dof_state"].copy_(torch.from_numpy(frames["dof"][k]))
        bufs["root_states"].copy_(torch.from_numpy(frames["root"]))
        bufs["contact_forces"].copy_(torch.from_numpy(frames["contact"]))
        bufs["rigid_body_states"].copy_(torch.from_numpy(frames["rigid"]))
        step += 1
        lib.emul_post_physics_step(C.byref(p), C.byref(bufs.struct), step)
        gu.check_step(bufs, gu.oracle_expected(orc, out), t)
        negative += int((out["rew_buf"] < 0).sum())
    if name == "only_positive_rewards":
        assert negative > 0                    # the un-clipped branch was really exercised


@pytest.mark.parametrize("control_type,randomize", [("V", True), ("T", True), ("P", False)])
def test_kernel_source_torque_control_variants(lib, control_type, randomize):
    """_compute_torques' other branches (legged_robot.py:456-471): velocity control, direct torque control, and position
    control without kp / kd randomisation -- no go2 cfg selects them, the oracle and the kernel source must still agree
    bit for bit (same fp32 operations in the same order)."""
    import state_util as su
    from legged_gym_custom_b200 import configs
    from legged_gym_custom_b200.params import NUM_DOF, env_params_from_cfg
    from oracle.go2_oracle import Go2Oracle
    base = configs.TASKS["go2"][0]
    cfg = type("Cfg", (base,), {"control": type("control", (base.control,), {"control_type": control_type}),
                                "domain_rand": type("domain_rand", (base.domain_rand,), {"randomize_kp_kd": randomize})})
    N = 257
    p = env_params_from_cfg(cfg, num_envs=N, seed=5)
    assert p.control_type == {"P": 0, "V": 1, "T": 2}[control_type] and p.randomize_kp_kd == int(randomize)
    rng = np.random.default_rng(11)
    statics, st = su.random_statics(p, rng), su.random_state(p, rng)
    st["dof_state"] = torch.from_numpy(np.stack([rng.normal(0, 0.5, N * NUM_DOF), rng.normal(0, 2.0, N * NUM_DOF)], 1).astype(np.float32))
    orc = Go2Oracle(p, statics, st)
    bufs = BufferSet(p, "cpu")
    bufs.load_statics(statics)
    st2 = dict(st)
    st2.pop("common_step_counter")
    bufs.load_state(st2)
    actions = torch.from_numpy(rng.normal(0, 2.5, (N, NUM_DOF)).astype(np.float32))     # beyond clip_actions and the torque limits
    orc.clip_actions(actions)
    want = orc.compute_torques()
    lib.emul_pd_torques(C.byref(p), C.byref(bufs.struct), C.c_void_p(actions.data_ptr()), 1)
    assert torch.equal(bufs["actions"], orc.st["actions"])
    assert torch.equal(bufs["torques"], want)
    if control_type != "T":                      # direct torques (action x 0.25) never reach the limits; the PD modes must
        assert (want.abs() == torch.tensor(list(p.torque_limits))).any() and (want.abs() < torch.tensor(list(p.torque_limits))).any()


def test_command_curriculum_rule_matches_reference_arithmetic(lib):
    """go2.py:87-107 restated with the reference's own types (fp32 0-dim tensor mean, Python-float thresholds, np.clip on
    Python floats) against command_curriculum_rule of the kernel source, over random ranges / limits / sums -- including the
    positive-max_reverse_vel branch whose np.clip upper bound is the value itself and sums right at the 80 % mark."""
    from legged_gym_custom_b200.params import EnvParams
    rng = np.random.default_rng(5)
    moved = 0
    for trial in range(400):
        p = EnvParams()
        scale_dt = float(rng.uniform(0.01, 0.1))
        p.max_episode_length = int(rng.integers(2, 2000))
        p.cc_threshold = 0.8 * scale_dt
        p.cc_vel_increment = float(rng.choice([0.05, 0.1, 0.25]))
        p.cc_max_forward_vel = float(rng.uniform(0.5, 3.0))
        p.cc_max_reverse_vel = float(rng.uniform(-2.0, 1.0))
        lo, hi = sorted(float(v) for v in rng.uniform(-1.5, 2.5, 2))
        n = int(rng.integers(0, 6))
        frac = float(rng.choice([0.5, 0.79999, 0.8, 0.80001, 1.2]))
        sums = torch.full((max(n, 1),), frac * scale_dt * p.max_episode_length) * torch.from_numpy(rng.uniform(0.999, 1.001, max(n, 1))).float()
        want_lo, want_hi = lo, hi
        if n > 0:                                                        # reset_idx returns before the curriculum when nothing resets
            mean = torch.mean(sums[:n]) / p.max_episode_length
            if mean > 0.8 * scale_dt:
                d = p.cc_vel_increment
                if p.cc_max_reverse_vel < 0.0:
                    want_lo = float(np.clip(lo - d, p.cc_max_reverse_vel, 0.))
                else:
                    want_lo = float(np.clip(lo - d, p.cc_max_reverse_vel, lo - d))
                want_hi = float(np.clip(hi + d, 0., p.cc_max_forward_vel))
                moved += 1
        src, dst = (C.c_double * 2)(lo, hi), (C.c_double * 2)()
        total = float(sums[:n].double().sum()) if n else 0.0
        lib.emul_command_curriculum_rule(C.byref(p), n, total, src, dst)
        # the kernel forms the mean from an fp64 sum: equal to torch's fp32 mean up to its last bit, so only trials whose mean
        # sits within one ulp of the mark may legitimately differ -- there are none by construction (0.79999 / 0.80001)
        assert (dst[0], dst[1]) == (want_lo, want_hi), (trial, n, frac, (lo, hi), (dst[0], dst[1]), (want_lo, want_hi))
    assert 50 < moved < 350
