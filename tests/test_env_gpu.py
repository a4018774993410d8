"""Parity of the CUDA env kernels, called through the C ABI (libb200gym.so), with the reference
(golden fixtures) and with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): bit-exact for height-sample indices, termination / reset /
time-out masks, contact flags, jump flags, episode lengths and curriculum levels; <= 1e-5 relative
(golden_util.rel_err) for every fp32 tensor."""
import ctypes as C

import numpy as np
import pytest
import torch

import golden_util as gu
import state_util as su
from legged_gym_custom_b200 import _lib, configs, synth
from legged_gym_custom_b200.buffers import BufferSet
from legged_gym_custom_b200.params import NUM_DOF, env_params_from_cfg
from oracle.go2_oracle import Go2Oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class CudaEnv:
    """thin driver of the C ABI over a BufferSet (what Go2Env does, minus the PhysX provider)."""

    def __init__(self, p, record_height_index=True, force_generic=0, prefetch=0, alias=0, terrain_tiles=0):
        """alias=1: observation outputs aliased into the critic rows (b200gym.h `alias_outputs`); with the go2 layout and no
        height-index recording that is post_physics_tile_kernel; terrain_tiles=1: its height scan reads a TMA-staged
        shared-memory tile of the field"""
        p.alias_outputs = int(alias)
        p.terrain_tiles = int(bool(terrain_tiles) and bool(p.has_height_samples))
        self.lib, self.p = _lib.lib(), p
        self.bufs = BufferSet(p, DEV, record_height_index=record_height_index)
        self.h = C.c_void_p()
        _lib.check(self.lib.b200_env_create(C.byref(p), 0, C.byref(self.h)))
        _lib.check(self.lib.b200_env_force_generic_layout(self.h, force_generic))
        _lib.check(self.lib.b200_env_set_prefetch(self.h, prefetch))

    def step(self, actions, frames, step, counter=None):
        """`counter`: int64 device tensor holding the step count BEFORE this step -> the graph-replay flavour of the call"""
        b, st = self.bufs, _lib.stream_ptr()
        a = torch.as_tensor(actions).to(DEV).contiguous()
        for k in range(self.p.decimation):
            _lib.check(self.lib.b200_pd_torques(self.h, C.byref(b.struct), C.c_void_p(a.data_ptr()), int(k == 0), st))
            b["dof_state"].copy_(torch.as_tensor(frames["dof"][k]))
        for name, key in (("root_states", "root"), ("contact_forces", "contact"), ("rigid_body_states", "rigid")):
            b[name].copy_(torch.as_tensor(frames[key]))
        if counter is not None:
            _lib.check(self.lib.b200_post_physics_step_dev(self.h, C.byref(b.struct), C.c_void_p(counter.data_ptr()), st))
        else:
            _lib.check(self.lib.b200_post_physics_step(self.h, C.byref(b.struct), step, st))
        torch.cuda.synchronize()

    def close(self):
        self.lib.b200_env_destroy(self.h)


@pytest.mark.parametrize("force_generic,prefetch,alias,record,tiles",
                         [(0, 0, 0, 1, 0), (1, 0, 0, 1, 0), (0, 1, 0, 1, 0), (1, 1, 0, 1, 0), (0, 0, 1, 0, 0), (1, 0, 1, 1, 0), (0, 0, 1, 1, 0), (0, 0, 1, 0, 1)],
                         ids=["go2-layout-baked-in", "layout-generic", "go2-layout-baked-in+prefetch", "layout-generic+prefetch",
                              "tile-kernel(aliased-rows)", "layout-generic+aliased-rows", "go2-layout-baked-in+aliased-rows",
                              "tile-kernel+smem-terrain-tiles"])
@pytest.mark.parametrize("task", gu.TASKS + gu.CC_SCENARIOS)
def test_cuda_env_matches_reference_golden(task, force_generic, prefetch, alias, record, tiles):
    g = gu.load(task)
    p = gu.params_for(task, g)
    env = CudaEnv(p, force_generic=force_generic, prefetch=prefetch, alias=alias, record_height_index=bool(record), terrain_tiles=tiles)
    env.bufs.load_statics(gu.statics_for(task, g))
    st = gu.init_state(g, p)
    step = int(st.pop("common_step_counter"))
    env.bufs.load_state(st)
    report = {}
    for t in range(int(g["steps"])):
        step += 1
        env.step(g[f"step{t}/in/actions"], gu.frames_of(g, t), step)
        gu.check_step(env.bufs, gu.expected(g, t), t, report=report)
    assert gu.rel_err(env.bufs["obs_history_buf"].cpu().numpy(), g["final/obs_history_buf"]) <= gu.RTOL
    print(task, {k: f"{v:.1e}" for k, v in sorted(report.items(), key=lambda kv: -kv[1])[:5]})
    env.close()


@pytest.mark.parametrize("task", gu.CC_SCENARIOS)
def test_command_curriculum_with_device_step_counter(task):
    """b200_post_physics_step_dev (common_step_counter in device memory, as under CUDA-graph replay): the probe pass and the
    range update of the command curriculum (go2.py:80-107, :222-223) take their decision from the device counter"""
    g = gu.load(task)
    p = gu.params_for(task, g)
    env = CudaEnv(p)
    env.bufs.load_statics(gu.statics_for(task, g))
    st = gu.init_state(g, p)
    counter = torch.tensor([int(st.pop("common_step_counter"))], dtype=torch.int64, device=DEV)
    env.bufs.load_state(st)
    for t in range(int(g["steps"])):
        env.step(g[f"step{t}/in/actions"], gu.frames_of(g, t), None, counter=counter)
        gu.check_step(env.bufs, gu.expected(g, t), t)
    assert int(counter.item()) == int(g["init/common_step_counter"]) + int(g["steps"])
    env.close()


@pytest.mark.parametrize("tile", [0, 1, 2], ids=["separate-outputs", "tile-kernel(aliased-rows)", "tile-kernel+smem-terrain-tiles"])
@pytest.mark.parametrize("task,num_envs,steps", [("go2_parkour", 4096, 3), ("go2_parkour_finetune", 1000, 2), ("go2", 777, 2),
                                                 ("go2_parkour", 1, 2), ("go2_parkour", 7, 2), ("go2_parkour", 9, 2),
                                                 ("go2_parkour", 65536, 1)])
def test_cuda_env_matches_oracle_at_scale(task, num_envs, steps, tile):
    """BASELINE config sizes (4096 envs; 65536 = the top of the env-count sweep), ragged sizes (not a multiple of the CTA's
    8 envs), fewer envs than one CTA holds, all three tasks."""
    cfg = configs.TASKS[task][0]
    hs, origins = gu.terrain_for(task)
    p = env_params_from_cfg(cfg, num_envs=num_envs, seed=99, hs_shape=None if hs is None else hs.shape)
    rng = np.random.default_rng(7)
    statics = su.random_statics(p, rng, hs, origins)
    st = su.random_state(p, rng, origins)
    orc = Go2Oracle(p, statics, st)
    env = CudaEnv(p, alias=int(tile > 0), record_height_index=not tile, terrain_tiles=int(tile == 2))
    env.bufs.load_statics(statics)
    st2 = dict(st)
    step = int(st2.pop("common_step_counter"))
    env.bufs.load_state(st2)
    origins0 = st["env_origins"].numpy()
    total_resets = 0
    for t in range(steps):
        frames = synth.make_frames(num_envs, origins0, rng, hole_prob=0.01, flip_prob=0.01)
        actions = rng.normal(0, 1.5, (num_envs, NUM_DOF)).astype(np.float32)
        out = orc.step(torch.from_numpy(actions), frames)
        step += 1
        env.step(actions, frames, step)
        gu.check_step(env.bufs, gu.oracle_expected(orc, out), t)
        total_resets += out["reset_count"]
    assert total_resets > 0 or num_envs < 64
    env.close()


def test_height_index_torch_cuda_division_mode():
    """index_div_mode=1 reproduces torch-CUDA's `points / horizontal_scale` (= points * fp32(1/scale)):
    compared bit-for-bit with the same torch ops executed on this GPU (SURVEY.md §7 hard part 1)."""
    task = "go2_parkour"
    cfg = configs.TASKS[task][0]
    hs, origins = gu.terrain_for(task)
    N = 4096
    p = env_params_from_cfg(cfg, num_envs=N, seed=5, hs_shape=hs.shape, index_div_mode=1)
    rng = np.random.default_rng(3)
    env = CudaEnv(p)
    env.bufs.load_statics(su.random_statics(p, rng, hs, origins))
    st = su.random_state(p, rng, origins)
    frames = synth.make_frames(N, st["env_origins"].numpy(), rng)
    env.bufs["root_states"].copy_(torch.from_numpy(frames["root"]))
    _lib.check(env.lib.b200_get_heights(env.h, C.byref(env.bufs.struct), _lib.stream_ptr()))
    torch.cuda.synchronize()
    # the reference's ops (legged_robot.py:1018-1032, math.py:38-42) on the GPU
    root = env.bufs["root_states"]
    xs = torch.tensor(list(p.scan_x)[:p.scan_nx], device=DEV)
    ys = torch.tensor(list(p.scan_y)[:p.scan_ny], device=DEV)
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")
    pts = torch.zeros(N, p.num_scan, 3, device=DEV)
    pts[:, :, 0], pts[:, :, 1] = gx.flatten(), gy.flatten()
    q = root[:, 3:7].repeat(1, p.num_scan).view(-1, 4).clone()
    q[:, :2] = 0.
    q = q / q.norm(p=2, dim=-1).clamp(min=1e-9).unsqueeze(-1)
    v = pts.view(-1, 3)
    xyz = q[:, :3]
    tt = xyz.cross(v, dim=-1) * 2
    rot = (v + q[:, 3:] * tt + xyz.cross(tt, dim=-1)).view(N, p.num_scan, 3) + root[:, :3].unsqueeze(1)
    rot += cfg.terrain.border_size
    idx = (rot / cfg.terrain.horizontal_scale).long()
    px = torch.clip(idx[:, :, 0], 0, hs.shape[0] - 2)
    py = torch.clip(idx[:, :, 1], 0, hs.shape[1] - 2)
    mine = env.bufs["height_index"]
    mism = int(((mine[..., 0] != px) | (mine[..., 1] != py)).sum())
    hsd = env.bufs["height_samples"]
    h = torch.min(torch.min(hsd[px, py], hsd[px + 1, py]), hsd[px, py + 1]).float() * cfg.terrain.vertical_scale
    print("torch-CUDA division-mode mismatches:", mism, "of", px.numel())
    assert mism == 0
    assert torch.equal(h, env.bufs["measured_heights"])
    env.close()


def test_go2env_api_smoke():
    """the drop-in surface: 8-tuple step, reset(), assignable episode_length_buf, extras on device."""
    from legged_gym_custom_b200.env import Go2Env

    class Cfg(configs.Go2ParkourCfg):
        class env(configs.Go2ParkourCfg.env):
            num_envs = 256
    env = Go2Env(Cfg, sim_device=DEV)
    obs, priv, crit, est, scan = env.reset()
    assert obs.shape == (256, 572) and priv.shape == (256, 29) and crit.shape == (256, 736) and est.shape == (256, 3) and scan.shape == (256, 132)
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
    for _ in range(5):
        out = env.step(torch.randn(256, 12, device=DEV))
    assert len(out) == 8 and out[5].shape == (256,) and out[6].dtype == torch.bool
    assert "time_outs" in out[7] and "rew_tracking_lin_vel" in out[7]["episode"] and "terrain_level" in out[7]["episode"]
    assert torch.isfinite(out[0]).all() and torch.isfinite(out[2]).all()
    assert torch.equal(out[2][:, :572], out[0])
    assert len(env.step5(torch.zeros(256, 12, device=DEV))) == 5


def test_go2_rough_terrain_matches_oracle():
    """BASELINE config 2: the `go2` task on the reference's DEFAULT terrain curriculum (legged_robot_config.py:20-53:
    trimesh, 10 x 20 tiles of 8 m, proportions [0.1, 0.1, 0.35, 0.25, 0.2]: smooth / rough slopes, stairs down / up,
    discrete obstacles) at 4096 envs, built by legged_gym_custom_b200.terrain.make_curriculum_terrain (deterministic tiles
    pinned to the reference's Terrain, tests/test_terrain_curriculum.py) -> height_samples [1300, 2100].  base_height, scan
    observations, curriculum levels and the bit-exact height indices are all exercised."""
    from legged_gym_custom_b200 import terrain as terrain_mod

    class RoughCfg(configs.Go2Cfg):
        class terrain(configs.Go2Cfg.terrain):
            mesh_type, curriculum, measure_heights, parkour = "trimesh", True, True, False
            num_rows, num_cols, terrain_length, terrain_width = 10, 20, 8., 8.
            max_init_terrain_level = 5
            terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2, 0.0, 0.0]

        class rewards(configs.Go2Cfg.rewards):
            class scales(configs.Go2Cfg.rewards.scales):
                base_height = -20.0
                stumble_feet = -1.0
                feet_air_time = 1.0
    rng = np.random.default_rng(11)
    hs, origins = terrain_mod.make_terrain(RoughCfg.terrain, seed=21)
    assert hs.shape == (10 * 80 + 500, 20 * 80 + 500) and hs.min() < -20 and hs.max() > 60
    N = 4096
    p = env_params_from_cfg(RoughCfg, num_envs=N, seed=21, hs_shape=hs.shape)
    assert p.has_height_samples and p.curriculum and not p.parkour and p.reward_scales[gu.REWARD_INDEX["base_height"]] != 0
    statics = su.random_statics(p, rng, hs, origins)
    st = su.random_state(p, rng, origins)
    orc = Go2Oracle(p, statics, st)
    env = CudaEnv(p)
    env.bufs.load_statics(statics)
    st2 = dict(st)
    step = int(st2.pop("common_step_counter"))
    env.bufs.load_state(st2)
    origins0 = st["env_origins"].numpy()
    for t_ in range(3):
        frames = synth.make_frames(N, origins0, rng, hole_prob=0.0, flip_prob=0.01)
        frames["root"][:, 0] = origins0[:, 0] + rng.uniform(-3.5, 6.0, N).astype(np.float32)   # around the tile centre: promote + demote
        actions = rng.normal(0, 1.5, (N, NUM_DOF)).astype(np.float32)
        out = orc.step(torch.from_numpy(actions), frames)
        step += 1
        env.step(actions, frames, step)
        gu.check_step(env.bufs, gu.oracle_expected(orc, out), t_)
    env.close()


def test_every_reward_term_active_matches_oracle():
    """all 38 `_reward_*` terms on at once (stateful feet_air_time and the post-clip termination reward included)."""
    from legged_gym_custom_b200.params import REWARD_TERMS

    class AllCfg(configs.Go2ParkourCfg):
        class rewards(configs.Go2ParkourCfg.rewards):
            only_positive_rewards = False
            soft_dof_vel_limit, soft_torque_limit = 0.05, 0.3

            class scales:
                pass
    for i, name in enumerate(REWARD_TERMS):
        setattr(AllCfg.rewards.scales, name, (-1.0) ** i * (0.3 + 0.1 * i))
    hs, origins = gu.terrain_for("go2_parkour")
    N = 3000
    p = env_params_from_cfg(AllCfg, num_envs=N, seed=5, hs_shape=hs.shape)
    rng = np.random.default_rng(3)
    statics, st = su.random_statics(p, rng, hs, origins), su.random_state(p, rng, origins)
    orc = Go2Oracle(p, statics, st)
    env = CudaEnv(p)
    env.bufs.load_statics(statics)
    st2 = dict(st)
    step = int(st2.pop("common_step_counter"))
    env.bufs.load_state(st2)
    for t_ in range(3):
        frames = synth.make_frames(N, st["env_origins"].numpy(), rng, hole_prob=0.02, flip_prob=0.02, body_hit_prob=0.05)
        actions = rng.normal(0, 1.5, (N, NUM_DOF)).astype(np.float32)
        out = orc.step(torch.from_numpy(actions), frames)
        step += 1
        env.step(actions, frames, step)
        gu.check_step(env.bufs, gu.oracle_expected(orc, out), t_)
    env.close()


def test_host_physx_zero_copy_matches_copy():
    """HostPhysX (bench.py's end-to-end arm): reading rigid_body_states / dof_state in place from pinned host memory and writing
    the torques of substeps 0..2 straight into the pinned mirror give bit-identical steps -- and identical host-side mirrors --
    to copying whole tensors both ways."""
    from legged_gym_custom_b200.env import Go2Env, HostPhysX

    class Cfg(configs.Go2ParkourCfg):
        class env(configs.Go2ParkourCfg.env):
            num_envs = 512
    outs = []
    for zero_copy in (True, False):
        env = Go2Env(Cfg, sim_device=DEV, seed=3)
        env.physx = HostPhysX(512, env.bufs["env_origins"], torch.device(DEV), seed=3, decimation=env.params.decimation,
                              zero_copy_rigid=zero_copy, zero_copy_dof=zero_copy, zero_copy_torques=zero_copy)
        env.reset()
        g = torch.Generator(device=DEV).manual_seed(0)
        for _ in range(5):
            out = env.step(torch.randn(512, NUM_DOF, device=DEV, generator=g))
        env.wait_extras()                          # joins the copy-back stream of the last step's state push
        torch.cuda.synchronize()
        px = env.physx
        assert float(px.host_torques.abs().max()) > 0.0 and torch.equal(px.host_torques[-1], env.bufs["torques"].cpu())
        outs.append([t.clone() for t in out[:7]] + [env.bufs["last_contact_heights"].clone()] +
                    [px.host_torques.clone(), px.host_root.clone(), px.host_dof.clone(), px.host_reset.clone(), px.host_rew.clone()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)        # incl. what went back to the host: the torques of all 4 substeps, pushed state, rewards
