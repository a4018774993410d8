"""`Go2Env`: the drop-in for the reference's `Go2Robot` / `LeggedRobot` env object, with the
torch op-by-op path replaced by the sm_100a kernels of libb200gym.so.

Mirrors the Python surface OnPolicyRunner / play.py use: constructor
`(cfg, sim_params, physics_engine, sim_device, headless)`, `step(actions)` returning the
reference's 8-tuple (legged_robot.py:100), `reset()`, `reset_idx()`, `get_*observations()`
(base_task.py:112-135) and the buffer attributes, all backed by device tensors that the kernels
update in place.  PhysX is out of scope (BASELINE.json north_star): it is an opaque producer of
the four state tensors, represented here by a small provider object (`SyntheticPhysX` replays
pre-generated frames; an Isaac Gym adapter would wrap gym.simulate / refresh_*).

There is no CPU path: constructing a Go2Env without CUDA or without the built library raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, synth, terrain as terrain_mod
from .buffers import BufferSet
from .params import NUM_BODIES, NUM_DOF, REWARD_INDEX, REWARD_TERMS, env_params_from_cfg, init_params_from_cfg


class SyntheticPhysX:
    """Replays a ring of pre-generated PhysX frame sets (synth.make_frames) with zero copies: the
    env's PhysX-owned buffer slots are re-pointed at the next frame, as if the simulator had just
    written them."""

    def __init__(self, num_envs, env_origins, device, ring=4, seed=1234, decimation=4, **frame_kw):
        rng = np.random.default_rng(seed)
        origins = env_origins.detach().cpu().numpy() if isinstance(env_origins, torch.Tensor) else np.asarray(env_origins)
        self.frames = []
        for _ in range(ring):
            f = synth.make_frames(num_envs, origins, rng, decimation=decimation, **frame_kw)
            self.frames.append({k: torch.from_numpy(v).to(device).contiguous() for k, v in f.items()})
        self.cursor = -1

    def begin_step(self, env):
        self.cursor = (self.cursor + 1) % len(self.frames)

    def simulate(self, env, substep):
        """gym.set_dof_actuation_force_tensor + gym.simulate + refresh_dof_state_tensor (legged_robot.py:81-85)."""
        env.bufs.rebind("dof_state", self.frames[self.cursor]["dof"][substep])

    def refresh(self, env):
        """refresh_actor_root_state / net_contact_force / rigid_body_state tensors (go2.py:352-353, :272)."""
        f = self.frames[self.cursor]
        env.bufs.rebind("root_states", f["root"])
        env.bufs.rebind("contact_forces", f["contact"])
        env.bufs.rebind("rigid_body_states", f["rigid"])

    def push_state(self, env):
        """set_dof_state_tensor_indexed / set_actor_root_state_tensor_indexed: nothing to do for a replay."""


class ExternalPhysX:
    """The caller writes the PhysX tensors itself (tests replaying recorded frames)."""

    def begin_step(self, env):
        pass

    def simulate(self, env, substep):
        if self.on_simulate is not None:
            self.on_simulate(env, substep)

    def refresh(self, env):
        if self.on_refresh is not None:
            self.on_refresh(env)

    def push_state(self, env):
        pass

    def __init__(self, on_simulate=None, on_refresh=None):
        self.on_simulate, self.on_refresh = on_simulate, on_refresh


def _buffer_property(name):
    def get(self):
        return self.bufs[name]

    def set_(self, value):                 # the runner ASSIGNS episode_length_buf (on_policy_runner.py:122)
        self.bufs[name].copy_(torch.as_tensor(value, device=self.device))

    return property(get, set_)


class Go2Env:
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True, *, physx=None,
                 seed=None, index_div_mode=0, height_samples=None, terrain_origins=None, record_height_index=False,
                 alias_outputs=True, terrain_tiles=False):
        if not torch.cuda.is_available():
            raise RuntimeError("Go2Env needs a CUDA device: the hot path has no CPU fallback")
        self.lib = _lib.lib()
        self.cfg = cfg
        self.sim_params = sim_params
        self.device = torch.device(sim_device)
        self.headless = headless
        seed = int(seed if seed is not None else getattr(cfg, "seed", 1))
        N = int(cfg.env.num_envs)

        hs = origins = None
        if cfg.terrain.mesh_type in ("heightfield", "trimesh"):
            if height_samples is None:      # the layout the reference's Terrain would build: parkour, or the default curriculum
                if getattr(cfg.terrain, "parkour", False):          # built on the device (csrc/terrain_kernels.cu)
                    height_samples, terrain_origins = terrain_mod.make_parkour_terrain_gpu(cfg.terrain, self.device)
                else:
                    height_samples, terrain_origins = terrain_mod.make_terrain(cfg.terrain, seed)
            hs = height_samples if isinstance(height_samples, torch.Tensor) else np.asarray(height_samples, dtype=np.int16)
            origins = np.asarray(terrain_origins, dtype=np.float32)
        # alias_outputs (default): obs / privileged / estimated / scan observations are column slices of the critic rows
        # (identical values, go2.py:538-563) -- strided views instead of four more buffers; `bind_output_rows` lets a
        # runner point the rows at its rollout-storage slot.  Pass False for separate contiguous buffers.
        self.params = p = env_params_from_cfg(cfg, num_envs=N, seed=seed, index_div_mode=index_div_mode,
                                              hs_shape=None if hs is None else hs.shape, alias_outputs=alias_outputs,
                                              terrain_tiles=terrain_tiles and hs is not None)
        self.bufs = BufferSet(p, self.device, record_height_index=record_height_index)
        self._handle = C.c_void_p()
        _lib.check(self.lib.b200_env_create(C.byref(p), self.device.index or 0, C.byref(self._handle)))

        # sizes (base_task.py:57-66)
        self.num_envs, self.num_actions = N, NUM_DOF
        self.num_proprio, self.history_buffer_length = p.num_proprio, p.history_len
        self.num_obs = p.num_proprio * (p.history_len + 1)
        self.num_privileged_obs, self.num_estimated_obs, self.num_scan_obs = p.num_priv, p.num_est, p.num_scan
        self.num_critic_obs = self.num_obs + p.num_priv + p.num_est + p.num_scan
        self.dt = cfg.control.decimation * float(cfg.sim.dt)
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = float(np.ceil(self.max_episode_length_s / self.dt))
        self.common_step_counter = 0
        self.step_counter_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        # set by a caller that consumes `extras` (time_outs, episode sums) on that stream or after wait_extras() only
        # (runner.OnPolicyRunner: the critic / bookkeeping stream of its PPO); None: everything on the current stream
        self.extras_stream = None
        self._extras_done = None
        self.use_device_counter = False        # True: common_step_counter lives on the device (CUDA-graph replay of the rollout)
        self.init_done = False
        self.reward_names = p.reward_names()
        self.feet_indices = torch.tensor(list(p.feet), device=self.device)
        self.penalised_contact_indices = torch.tensor(list(p.penalised)[:p.n_penalised], device=self.device)
        self.termination_contact_indices = torch.tensor(list(p.termination)[:p.n_termination], device=self.device)
        self.default_dof_pos = torch.tensor(list(p.default_dof_pos), device=self.device).unsqueeze(0)
        self._extras = {}

        self._init_domain_randomisation(cfg, hs, origins, seed)
        self.physx = physx if physx is not None else SyntheticPhysX(N, self.bufs["env_origins"], self.device, seed=seed,
                                                                    decimation=p.decimation)
        if hasattr(self.physx, "bind"):                 # a simulator-backed provider wraps its own state tensors (integration.py)
            self.physx.bind(self)
        self.init_done = True

    # ---- env-creation-time randomisation (legged_robot.py:306-380, :696-701, :897-930): one kernel over the envs
    #      (b200_env_init_randomisation, keyed Philox draws) instead of the reference's per-env Python loop
    def _init_domain_randomisation(self, cfg, hs, origins, seed):
        b = self.bufs
        if hs is not None:
            b["height_samples"].copy_(hs if isinstance(hs, torch.Tensor) else torch.from_numpy(hs))
            b["terrain_origins"].copy_(torch.from_numpy(origins))
        self.init_params = init_params_from_cfg(cfg, self.num_envs, hs is not None)
        _lib.check(self.lib.b200_env_init_randomisation(self._handle, C.byref(b.struct), C.byref(self.init_params), _lib.stream_ptr()))

    def __del__(self):
        try:
            if self._handle:
                self.lib.b200_env_destroy(self._handle)
        except Exception:
            pass

    # ---- the VecEnv contract -----------------------------------------------------------------------
    def step(self, actions):
        """legged_robot.py:67-100. -> (obs, privileged_obs, critic_obs, estimated_obs, scan_obs, rew, reset, extras)"""
        actions = actions.to(self.device, torch.float32).contiguous()
        st = _lib.stream_ptr()
        b = self.bufs
        self.physx.begin_step(self)
        for k in range(self.params.decimation):
            _lib.check(self.lib.b200_pd_torques(self._handle, C.byref(b.struct), C.c_void_p(actions.data_ptr()), int(k == 0), st))
            self.physx.simulate(self, k)
        self.physx.refresh(self)
        self.common_step_counter += 1
        if self.use_device_counter and self.extras_stream is not None:
            # the episode statistics / time-out copy (extras_kernel, which also commits the device step counter) leave the
            # critical path: launched on `extras_stream` behind the per-env kernel; the NEXT step's per-env kernel waits for it
            ctr = C.c_void_p(self.step_counter_dev.data_ptr())
            if self._extras_done is not None:
                torch.cuda.current_stream().wait_event(self._extras_done)
            _lib.check(self.lib.b200_post_physics_step_dev_parts(self._handle, C.byref(b.struct), ctr, 1, st))
            ev = torch.cuda.Event()
            ev.record()
            self.extras_stream.wait_event(ev)
            with torch.cuda.stream(self.extras_stream):
                _lib.check(self.lib.b200_post_physics_step_dev_parts(self._handle, C.byref(b.struct), ctr, 2, _lib.stream_ptr()))
                self._extras_done = torch.cuda.Event()
                self._extras_done.record()
        elif self.use_device_counter:
            _lib.check(self.lib.b200_post_physics_step_dev(self._handle, C.byref(b.struct), C.c_void_p(self.step_counter_dev.data_ptr()), st))
        else:
            _lib.check(self.lib.b200_post_physics_step(self._handle, C.byref(b.struct), self.common_step_counter, st))
        self.physx.push_state(self)
        return (b["obs_buf"], b["privileged_obs_buf"], b["critic_obs_buf"], b["estimated_obs_buf"], b["scan_obs_buf"],
                b["rew_buf"], b["reset_buf"], self.extras)

    def set_device_counter(self, enabled=True):
        self.wait_extras()
        self.step_counter_dev.fill_(self.common_step_counter)
        self.use_device_counter = bool(enabled)

    def wait_extras(self):
        """order the current stream after the last step's extras launch (only needed with `extras_stream`)"""
        if self._extras_done is not None:
            torch.cuda.current_stream().wait_event(self._extras_done)
            self._extras_done = None
        if hasattr(self.physx, "finish"):          # HostPhysX: the last step's state push rides on a copy stream
            self.physx.finish()

    @property
    def supports_output_binding(self):
        return bool(self.params.alias_outputs)

    def bind_output_rows(self, rows):
        """the next step() writes its observation rows [N, num_critic_obs] into `rows` (alias_outputs only)"""
        self.bufs.bind_output_rows(rows)

    def step5(self, actions):
        """upstream rsl_rl VecEnv 5-tuple (rsl_rl/env/vec_env.py:28): obs, privileged_obs, rew, done, info."""
        obs, priv, _, _, _, rew, done, info = self.step(actions)
        return obs, priv, rew, done, info

    def reset_idx(self, env_ids=None):
        """go2.py:207-263 outside a step.  Only the all-envs form is used by the reference (BaseTask.reset)."""
        if env_ids is not None and len(env_ids) != self.num_envs:
            raise NotImplementedError("partial reset_idx outside step() is not on the hot path; step() resets flagged envs itself")
        _lib.check(self.lib.b200_reset_all(self._handle, C.byref(self.bufs.struct), self.common_step_counter,
                                           int(self.init_done), _lib.stream_ptr()))
        self.physx.push_state(self)

    def reset(self):
        """base_task.py:131-135."""
        self.reset_idx()
        out = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        return out[:5]

    def get_observations(self):
        return self.bufs["obs_buf"]

    def get_privileged_observations(self):
        return self.bufs["privileged_obs_buf"]

    def get_critic_observations(self):
        return self.bufs["critic_obs_buf"]

    def get_estimated_observations(self):
        return self.bufs["estimated_obs_buf"]

    def get_scan_observations(self):
        return self.bufs["scan_obs_buf"]

    def get_heights(self):
        """LeggedRobot._get_heights as a standalone launch (legged_robot.py:997-1032)."""
        _lib.check(self.lib.b200_get_heights(self._handle, C.byref(self.bufs.struct), _lib.stream_ptr()))
        return self.bufs["measured_heights"]

    # ---- extras: device-resident, refreshed by the kernels only on steps with >= 1 reset (go2.py:246-263)
    @property
    def extras(self):
        if not self._extras:
            ep = self.bufs["extras_episode"]
            episode = {"rew_" + n: ep[REWARD_INDEX[n]] for n in REWARD_TERMS if self.params.reward_scales[REWARD_INDEX[n]] != 0.0}
            if self.params.curriculum:
                episode["terrain_level"] = ep[len(REWARD_TERMS)]
            if self.params.command_curriculum:             # go2.py:255-259; the lin_vel_x range lives on the device
                cr, rng = self.bufs["command_ranges"], self.cfg.commands.ranges
                episode.update(max_command_x=cr[3], min_command_x=cr[2], max_command_y=rng.lin_vel_y[1], max_command_yaw=rng.ang_vel_yaw[1])
            self._extras = {"episode": episode}
            if getattr(self.cfg.env, "send_timeouts", True):
                self._extras["time_outs"] = self.bufs["extras_time_outs"]
        return self._extras

    @property
    def command_ranges(self):
        """legged_robot.py:949 `command_ranges` (lists of Python floats); with a command curriculum the lin_vel_x entry is
        read back from the device buffer the kernels move (one sync)"""
        rng = self.cfg.commands.ranges
        out = {k: list(getattr(rng, k)) for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading")}
        if self.params.command_curriculum:
            out["lin_vel_x"] = self.bufs["command_ranges"][2:4].tolist()
        return out

    @property
    def episode_sums(self):
        s = self.bufs["episode_sums"]
        return {n: s[:, REWARD_INDEX[n]] for n in REWARD_TERMS if self.params.reward_scales[REWARD_INDEX[n]] != 0.0}

    @property
    def dof_pos(self):
        return self.bufs["dof_state"].view(self.num_envs, NUM_DOF, 2)[..., 0]

    @property
    def dof_vel(self):
        return self.bufs["dof_state"].view(self.num_envs, NUM_DOF, 2)[..., 1]

    @property
    def contact_forces(self):
        return self.bufs["contact_forces"].view(self.num_envs, NUM_BODIES, 3)

    @property
    def base_quat(self):
        return self.bufs["root_states"][:, 3:7]

    @property
    def roll(self):
        return self.bufs["rpy"][:, 0]

    @property
    def pitch(self):
        return self.bufs["rpy"][:, 1]

    @property
    def yaw(self):
        return self.bufs["rpy"][:, 2]

    @property
    def privileged_mass_params(self):
        return self.bufs["priv_mass_params"]

    @property
    def privileged_friction_coeffs(self):
        return self.bufs["priv_friction"]


for _name in ("obs_buf", "privileged_obs_buf", "critic_obs_buf", "estimated_obs_buf", "scan_obs_buf", "rew_buf", "reset_buf",
              "time_out_buf", "obs_history_buf", "commands", "torques", "actions", "base_lin_vel", "base_ang_vel",
              "projected_gravity", "root_states", "dof_state", "rigid_body_states", "episode_length_buf", "last_actions",
              "last_dof_vel", "last_root_vel", "last_base_lin_vel", "last_torques", "last_contacts", "last_contact_heights",
              "feet_air_time", "jump_flags", "terrain_levels", "terrain_types", "env_origins", "measured_heights",
              "kp_kd_multipliers", "height_samples", "terrain_origins"):
    setattr(Go2Env, _name, _buffer_property(_name))


class HostPhysX:
    """PhysX living on the HOST (the reference's --sim_device=cpu pipeline: the simulator reads and writes host tensors).
    Every tensor that crosses the simulator boundary crosses the bus inside `Go2Env.step`, on the current stream:

    host -> device (pinned frames): dof_state after every substep (legged_robot.py:84-85), root_states, contact_forces and
        rigid_body_states after the last one (go2.py:352-353, :272).  dof_state, root_states and contact_forces are read
        densely and are copied -- except the dof_state of substeps 0..2, whose only reader is the next PD-torque kernel:
        that kernel streams the pinned frame in place (`zero_copy_dof`); rigid_body_states [N*19,13] is read at 4 floats
        per env (the feet heights), so the kernel reads it in place too (`zero_copy_rigid`) instead of copying 4 MB.
    device -> host (pinned mirrors): the torques of EVERY substep (gym.set_dof_actuation_force_tensor, legged_robot.py:81-83;
        substeps 0..2 are written by the PD kernel straight into the mirror -- `zero_copy_torques` -- the last one is copied),
        and after the step the state the reference pushes back with set_dof_state_tensor_indexed /
        set_actor_root_state_tensor_indexed (legged_robot.py:504-506, :530-532, :539): root_states, dof_state and the reset
        flags that select the rows -- copied whole (a fixed-size, graph-capturable transfer; an upper bound of the rows the
        reference sends), plus the step's rewards and dones (`read_back_results`).
    `bytes_per_step` / `d2h_bytes_per_step` count what crosses the bus per env step.  Used by bench.py's end-to-end leg."""

    def __init__(self, num_envs, env_origins, device, ring=4, seed=1234, decimation=4, zero_copy_rigid=True, zero_copy_dof=True,
                 read_back_results=True, zero_copy_torques=True, **frame_kw):
        rng = np.random.default_rng(seed)
        origins = env_origins.detach().cpu().numpy() if isinstance(env_origins, torch.Tensor) else np.asarray(env_origins)
        self.frames = []
        for _ in range(ring):
            f = synth.make_frames(num_envs, origins, rng, decimation=decimation, **frame_kw)
            self.frames.append({k: torch.from_numpy(v).pin_memory() for k, v in f.items()})
        self.cursor, self.h2d_bytes, self.d2h_bytes = -1, 0, 0
        self._pushed = None
        self.zero_copy_torques = bool(zero_copy_torques)
        self.zero_copy_rigid = bool(zero_copy_rigid)
        self.zero_copy_dof, self.decimation = bool(zero_copy_dof), int(decimation)
        f0 = self.frames[0]
        # zero-copy reads: 4 feet per env, one 32-byte sector each
        self.rigid_bytes = num_envs * 4 * 32 if self.zero_copy_rigid else 4 * f0["rigid"].numel()
        self.bytes_per_step = 4 * (f0["dof"].numel() + f0["root"].numel() + f0["contact"].numel()) + self.rigid_bytes
        pin = lambda *shape, dt=torch.float32: torch.zeros(*shape, dtype=dt).pin_memory()
        self.host_torques = pin(decimation, num_envs, NUM_DOF)       # what the host simulator is actuated with
        self.host_root, self.host_dof = pin(num_envs, 13), pin(num_envs * NUM_DOF, 2)
        self.host_reset = pin(num_envs, dt=torch.bool)
        self.read_back_results = bool(read_back_results)
        self.host_rew, self.host_done = pin(num_envs), pin(num_envs, dt=torch.bool)
        self.d2h_bytes_per_step = 4 * (self.host_torques.numel() + self.host_root.numel() + self.host_dof.numel()) + num_envs \
            + (5 * num_envs if self.read_back_results else 0)

    def begin_step(self, env):
        self.cursor = (self.cursor + 1) % len(self.frames)
        self.finish()                  # the previous step's state push reads buffers this step's frames overwrite
        if self.zero_copy_torques and self.decimation > 1:
            env.bufs.rebind_host_mapped("torques", self.host_torques[0])

    def finish(self):
        """order the current stream after the last push_state() (issued on the copy-back stream)"""
        if self._pushed is not None:
            torch.cuda.current_stream().wait_event(self._pushed)
            self._pushed = None

    def simulate(self, env, substep):
        """torques of this substep -> host; dof_state after substep k -> device.  Between substeps the only reader of
        dof_state is the next PD-torque kernel (legged_robot.py:81-85), which streams it once: that kernel reads the pinned
        frame in place (`zero_copy_dof`); the frame of the LAST substep is what post_physics_step and the env's `dof_state`
        attribute see, so it is copied."""
        last = substep == self.decimation - 1
        if self.zero_copy_torques:
            # substeps 0..D-2: the PD kernel has written its torques straight into the pinned mirror (posted writes over the
            # bus, no copy launch); only the last substep's torques have a device-side reader (the reward terms) and are copied
            if last:
                self.host_torques[substep].copy_(env.bufs["torques"], non_blocking=True)
            elif substep + 1 < self.decimation - 1:
                env.bufs.rebind_host_mapped("torques", self.host_torques[substep + 1])
            else:
                env.bufs.unbind_host_mapped("torques")
        else:
            self.host_torques[substep].copy_(env.bufs["torques"], non_blocking=True)
        self.d2h_bytes += 4 * env.bufs["torques"].numel()
        src = self.frames[self.cursor]["dof"][substep]
        if self.zero_copy_dof and not last:
            env.bufs.rebind_host_mapped("dof_state", src)
        else:
            if self.zero_copy_dof:
                env.bufs.unbind_host_mapped("dof_state")
            env.bufs["dof_state"].copy_(src, non_blocking=True)
        self.h2d_bytes += src.numel() * 4

    def refresh(self, env):
        """root_states and contact_forces become available together (after the last substep): their two copies go out on
        two streams, so they share the bus / copy engines instead of queueing behind each other."""
        f = self.frames[self.cursor]
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=env.device)
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        self._copy_stream.wait_event(fork)
        with torch.cuda.stream(self._copy_stream):
            env.bufs["root_states"].copy_(f["root"], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        env.bufs["contact_forces"].copy_(f["contact"], non_blocking=True)
        main.wait_event(done)
        self.h2d_bytes += 4 * (f["root"].numel() + f["contact"].numel())
        if self.zero_copy_rigid:
            env.bufs.rebind_host_mapped("rigid_body_states", f["rigid"])
        else:
            env.bufs["rigid_body_states"].copy_(f["rigid"], non_blocking=True)
        self.h2d_bytes += self.rigid_bytes

    def push_state(self, env):
        """the reset / pushed rows go back to the host simulator (whole tensors + the flags that select the rows), and the
        step's rewards / dones to the host caller"""
        b = env.bufs
        # The host simulator needs these before its NEXT simulate(), i.e. after the next policy inference: the copies go out
        # on their own stream behind the step's kernels and are joined at the next begin_step() (finish()), so the policy
        # inference of the next step overlaps them instead of queueing behind five bus transfers.
        if not hasattr(self, "_back_stream"):
            self._back_stream = torch.cuda.Stream(device=env.device)
        fork = torch.cuda.Event()
        fork.record()
        self._back_stream.wait_event(fork)
        with torch.cuda.stream(self._back_stream):
            self.host_root.copy_(b["root_states"], non_blocking=True)
            self.host_dof.copy_(b["dof_state"], non_blocking=True)
            self.host_reset.copy_(b["reset_buf"], non_blocking=True)
            if self.read_back_results:
                self.host_rew.copy_(b["rew_buf"], non_blocking=True)
                self.host_done.copy_(b["reset_buf"], non_blocking=True)
            self._pushed = torch.cuda.Event()
            self._pushed.record()
        self.d2h_bytes += self.d2h_bytes_per_step - 4 * self.host_torques.numel()
