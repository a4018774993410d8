"""CPU oracle for the env half of the hot path: Go2Robot.step / post_physics_step.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.  The product path (legged_gym_custom_b200)
never does; it fails loudly when libb200gym.so is missing.

A restatement, in plain fp32 torch-CPU ops, of what the reference computes between two
PhysX steps.  Every block cites the reference lines it follows.  Parity status: PINNED --
oracle/make_golden.py runs the unmodified reference (behind oracle/refshim) on the same
frames and keyed random numbers, and tests/test_oracle_golden.py requires this file to
reproduce every tensor of those runs bit-for-bit (tests/golden/*.npz).

Third-party arithmetic: isaacgym.torch_utils (Isaac Gym Preview 4, unpinned binary, absent
from /root/reference) -- quat_rotate_inverse, quat_apply, normalize, torch_rand_float are
restated from their published definitions.

State layout: dict of torch tensors with the reference's attribute names.
"""
import math

import numpy as np
import torch

from . import philox
from legged_gym_custom_b200.params import NUM_DOF, NUM_BODIES, REWARD_TERMS, REWARD_INDEX

TWO_PI = 2 * math.pi


# ---- isaacgym.torch_utils restated ------------------------------------------------------
def quat_rotate_inverse(q, v):
    w = q[:, 3]
    xyz = q[:, :3]
    a = v * (2.0 * w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(xyz, v, dim=-1) * w.unsqueeze(-1) * 2.0
    c = xyz * torch.bmm(xyz.view(-1, 1, 3), v.view(-1, 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def quat_apply(q, v):
    xyz = q[:, :3]
    t = xyz.cross(v, dim=-1) * 2
    return v + q[:, 3:] * t + xyz.cross(t, dim=-1)


def wrap_to_pi(angles):
    """legged_gym/utils/math.py:45-48 (in place on its argument, like the reference)."""
    angles %= TWO_PI
    angles -= TWO_PI * (angles > math.pi)
    return angles


def euler_from_quat(q):
    """go2.py:11-31."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    pitch = torch.asin(torch.clip(2.0 * (w * y - z * x), -1, 1))
    yaw = torch.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return roll, pitch, yaw


class Go2Oracle:
    def __init__(self, params, statics, state):
        """params: EnvParams; statics: kp_kd_multipliers [2,N,12], privileged_mass_params [N,4],
        privileged_friction_coeffs [N,1], height_samples int16 [R,C] | None, terrain_origins
        [L,T,3] | None; state: persistent buffers (see `fresh_state`)."""
        self.p = p = params
        self.N = p.num_envs
        t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        self.st = {k: t(v).clone() for k, v in state.items()}
        self.kp_kd = t(statics["kp_kd_multipliers"]).float()
        self.mass = t(statics["privileged_mass_params"]).float()
        self.fric = t(statics["privileged_friction_coeffs"]).float()
        hs = statics.get("height_samples")
        self.height_samples = None if hs is None or not p.has_height_samples else t(hs)
        to = statics.get("terrain_origins")
        self.terrain_origins = None if to is None else t(to).float()
        f = lambda arr: torch.tensor(list(arr), dtype=torch.float32)
        self.p_gains, self.d_gains = f(p.p_gains), f(p.d_gains)
        self.default_dof_pos = f(p.default_dof_pos).unsqueeze(0)
        self.torque_limits = f(p.torque_limits)
        self.dof_pos_limits = torch.stack([f(p.dof_pos_lo), f(p.dof_pos_hi)], dim=1)
        self.dof_vel_limits = f(p.dof_vel_limits)
        self.noise_vec = f(p.noise_vec)[:p.num_proprio]
        self.base_init_state = f(p.base_init_state)
        self.commands_scale = torch.tensor([p.obs_lin_vel, p.obs_lin_vel, p.obs_ang_vel])
        gx, gy = torch.meshgrid(f(p.scan_x)[:p.scan_nx], f(p.scan_y)[:p.scan_ny], indexing="ij")
        self.height_points = torch.zeros(self.N, p.num_scan, 3)          # legged_robot.py:980-994
        self.height_points[:, :, 0] = gx.flatten()
        self.height_points[:, :, 1] = gy.flatten()
        self.feet = list(p.feet)
        self.calves = list(p.calves)
        self.pen = list(p.penalised)[:p.n_penalised]
        self.term = list(p.termination)[:p.n_termination]
        self.hip_j, self.thigh_j, self.calf_j = list(p.hip_joints), list(p.thigh_joints), list(p.calf_joints)
        self.gravity_vec = torch.tensor([0., 0., -1.]).repeat(self.N, 1)
        self.forward_vec = torch.tensor([1., 0., 0.]).repeat(self.N, 1)
        self.active = [n for n in REWARD_TERMS[:-1] if p.reward_scales[REWARD_INDEX[n]] != 0.0]
        self.out = {}
        # command_ranges["lin_vel_x"] (legged_robot.py:949): Python floats, moved by update_command_curriculum
        if "command_ranges" not in self.st:
            self.st["command_ranges"] = torch.tensor([p.cc_range0[0], p.cc_range0[1]], dtype=torch.float64)

    # ---- construction helpers ------------------------------------------------------------
    @staticmethod
    def fresh_state(p, env_origins, terrain_levels=None, terrain_types=None):
        """All-zero persistent buffers (legged_robot.py:648-681, go2.py:132-172, base_task.py:76-86)."""
        N = p.num_envs
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
        root = z(N, 13)
        root[:, 6] = 1.0
        return dict(
            root_states=root, dof_state=z(N * NUM_DOF, 2), contact_forces=z(N * NUM_BODIES, 3),
            rigid_body_states=z(N * NUM_BODIES, 13), actions=z(N, NUM_DOF), torques=z(N, NUM_DOF),
            commands=z(N, 4), episode_length_buf=z(N, dt=torch.int64), last_actions=z(N, NUM_DOF),
            last_dof_vel=z(N, NUM_DOF), last_root_vel=z(N, 6), last_base_lin_vel=z(N, 3), last_torques=z(N, NUM_DOF),
            obs_history_buf=z(N, p.history_len, p.num_proprio), last_contacts=z(N, 4, dt=torch.bool),
            last_contact_heights=z(N, 4), feet_air_time=z(N, 4), jump_flags=z(N, 1),
            episode_sums=z(len(REWARD_TERMS), N),
            terrain_levels=(terrain_levels if terrain_levels is not None else z(N, dt=torch.int64)),
            terrain_types=(terrain_types if terrain_types is not None else z(N, dt=torch.int64)),
            env_origins=torch.as_tensor(env_origins, dtype=torch.float32).clone(),
            reset_buf=torch.ones(N, dtype=torch.bool), time_out_buf=z(N, dt=torch.bool),
            extras_time_outs=z(N, dt=torch.bool), extras_episode=z(len(REWARD_TERMS) + 1),
            common_step_counter=torch.zeros((), dtype=torch.int64),
        )

    # ---- keyed uniforms --------------------------------------------------------------------
    def _u(self, site, env_ids, lanes):
        ids = env_ids.numpy() if isinstance(env_ids, torch.Tensor) else np.asarray(env_ids)
        step = int(self.st["common_step_counter"])
        return torch.from_numpy(philox.keyed_uniform(self.p.seed, site, step, ids, lanes))

    def _u32(self, site, env_ids, lanes):
        ids = env_ids.numpy() if isinstance(env_ids, torch.Tensor) else np.asarray(env_ids)
        step = int(self.st["common_step_counter"])
        return philox.keyed_u32(self.p.seed, site, step, ids, lanes)

    # ---- views -----------------------------------------------------------------------------
    @property
    def dof_pos(self):
        return self.st["dof_state"].view(self.N, NUM_DOF, 2)[..., 0]

    @property
    def dof_vel(self):
        return self.st["dof_state"].view(self.N, NUM_DOF, 2)[..., 1]

    @property
    def contact(self):
        return self.st["contact_forces"].view(self.N, NUM_BODIES, 3)

    # ---- K1: legged_robot.py:74-75 + :440-478 -------------------------------------------
    def clip_actions(self, actions):
        c = self.p.clip_actions
        self.st["actions"] = torch.clip(actions.float(), -c, c)

    def compute_torques(self):
        p, a = self.p, self.st["actions"]
        scaled = a * p.action_scale
        if p.control_type == 0:
            if p.randomize_kp_kd:
                tq = self.kp_kd[0] * self.p_gains * (scaled + self.default_dof_pos - self.dof_pos) \
                    - self.kp_kd[1] * self.d_gains * self.dof_vel
            else:
                tq = self.p_gains * (scaled + self.default_dof_pos - self.dof_pos) - self.d_gains * self.dof_vel
        elif p.control_type == 1:
            tq = self.p_gains * (scaled - self.dof_vel) - self.d_gains * (self.dof_vel - self.st["last_dof_vel"]) / p.sim_dt
        else:
            tq = scaled
        self.st["torques"] = torch.clip(tq, -self.torque_limits, self.torque_limits)
        return self.st["torques"]

    def load_physx(self, frames, substep=None):
        """stand-in for gym.simulate / refresh_*: copy a synthetic frame into the state tensors."""
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a
        if substep is not None:
            self.st["dof_state"].copy_(t(frames["dof"][substep]))
        else:
            self.st["root_states"].copy_(t(frames["root"]))
            self.st["contact_forces"].copy_(t(frames["contact"]))
            self.st["rigid_body_states"].copy_(t(frames["rigid"]))

    def step(self, actions, frames):
        """LeggedRobot.step (legged_robot.py:67-100) with PhysX replaced by `frames`."""
        self.clip_actions(actions)
        for k in range(self.p.decimation):
            self.compute_torques()
            self.load_physx(frames, substep=k)
        self.load_physx(frames)
        self.post_physics_step()
        return self.out

    # ---- command resampling: go2.py:413-464 ------------------------------------------------
    def _resample_commands(self, env_ids, site):
        p, cmd = self.p, self.st["commands"]
        if len(env_ids) == 0:
            return
        u = self._u(site, env_ids, [0, 1, 2, 3])
        lo_x, span_x = p.cmd_lo[0], p.cmd_span[0]
        if p.command_curriculum:      # torch_rand_float(lower, upper): (upper - lower) in double, then fp32 arithmetic
            lo, hi = (float(v) for v in self.st["command_ranges"])
            lo_x, span_x = float(np.float32(lo)), float(np.float32(hi - lo))
        cmd[env_ids, 0] = span_x * u[:, 0] + lo_x
        cmd[env_ids, 1] = p.cmd_span[1] * u[:, 1] + p.cmd_lo[1]
        if p.heading_command:
            cmd[env_ids, 3] = p.cmd_span[3] * u[:, 2] + p.cmd_lo[3]
        else:
            cmd[env_ids, 2] = p.cmd_span[2] * u[:, 2] + p.cmd_lo[2]
        cmd[env_ids, :2] *= (torch.norm(cmd[env_ids, :2], dim=1) > 0.2).unsqueeze(1)
        if p.zero_command:
            idx = env_ids[u[:, 3] < p.zero_command_prob]
            cmd[idx, 0:3] *= 0.0
            if p.heading_command:
                fwd = quat_apply(self.st["root_states"][idx, 3:7], self.forward_vec[idx])
                cmd[idx, 3] = torch.atan2(fwd[:, 1], fwd[:, 0])

    # ---- height scan: legged_robot.py:997-1032, math.py:38-42 -----------------------------
    def get_heights(self):
        p, N = self.p, self.N
        if not p.has_height_samples:
            self.out["height_index"] = torch.zeros(N, p.num_scan, 2, dtype=torch.int64)
            return torch.zeros(N, p.num_scan)
        q = self.st["root_states"][:, 3:7].repeat(1, p.num_scan).view(-1, 4).clone()
        q[:, :2] = 0.
        q = q / q.norm(p=2, dim=-1).clamp(min=1e-9).unsqueeze(-1)
        pts = quat_apply(q, self.height_points.view(-1, 3)).view(N, p.num_scan, 3) + self.st["root_states"][:, :3].unsqueeze(1)
        pts += p.border_size
        if p.index_div_mode == 0:
            pts = (pts / p.horizontal_scale).long()
        else:   # torch-CUDA evaluates tensor / python_scalar as tensor * fp32(1 / scalar)
            pts = (pts * float(np.float32(1.0) / np.float32(p.horizontal_scale))).long()
        px = torch.clip(pts[:, :, 0].reshape(-1), 0, self.height_samples.shape[0] - 2)
        py = torch.clip(pts[:, :, 1].reshape(-1), 0, self.height_samples.shape[1] - 2)
        self.out["height_index"] = torch.stack([px, py], dim=1).view(N, p.num_scan, 2)
        h = torch.min(torch.min(self.height_samples[px, py], self.height_samples[px + 1, py]), self.height_samples[px, py + 1])
        return h.view(N, -1) * p.vertical_scale

    # ---- rewards: legged_robot.py:1036-1148, go2.py:578-831 -------------------------------
    def _reward(self, name):
        p, st, o = self.p, self.st, self.out
        root, cmd, F = st["root_states"], st["commands"], self.contact
        dq = self.dof_pos - self.default_dof_pos
        sq = torch.square
        cmd_norm3 = torch.norm(cmd[:, :3], dim=1)

        def stance(ph):
            return torch.sin(TWO_PI * ph) <= p.stance_threshold

        if name == "action_rate":
            return torch.sum(sq(st["last_actions"] - st["actions"]), dim=1)
        if name == "ang_vel_xy":
            return torch.sum(sq(o["base_ang_vel"][:, :2]), dim=1)
        if name == "base_height":
            bh = torch.mean(root[:, 2].unsqueeze(1) - o["measured_heights"], dim=1)
            return sq(bh - p.base_height_target)
        if name == "calf_collision":
            return torch.sum(1.0 * (torch.norm(F[:, self.calves, :], dim=-1) > 0.1), dim=1)
        if name == "calf_pos":
            return torch.sum(sq(dq[:, self.calf_j]), dim=1)
        if name == "calf_symmetry":
            c = self.calf_j
            return torch.sum(torch.abs(self.dof_pos[:, [c[0], c[2]]] - self.dof_pos[:, [c[1], c[3]]]), dim=1)
        if name == "collision":
            return torch.sum(1. * (torch.norm(F[:, self.pen, :], dim=-1) > 0.1), dim=1)
        if name == "delta_torques":
            return torch.sum(sq(st["torques"] - st["last_torques"]), dim=1)
        if name == "dof_acc":
            return torch.sum(sq((st["last_dof_vel"] - self.dof_vel) / p.dt), dim=1)
        if name == "dof_error":
            return torch.sum(sq(dq), dim=1)
        if name == "dof_pos_limits":
            out = -(self.dof_pos - self.dof_pos_limits[:, 0]).clip(max=0.)
            out += (self.dof_pos - self.dof_pos_limits[:, 1]).clip(min=0.)
            return torch.sum(out, dim=1)
        if name == "dof_vel":
            return torch.sum(sq(self.dof_vel), dim=1)
        if name == "dof_vel_limits":
            return torch.sum((torch.abs(self.dof_vel) - self.dof_vel_limits * p.soft_dof_vel_limit).clip(min=0., max=1.), dim=1)
        if name == "feet_air_time":                                        # go2.py:819-831 (stateful)
            contact = F[:, self.feet, 2] > 1.
            filt = torch.logical_or(contact, st["last_contacts"])
            first = (st["feet_air_time"] > 0.) * filt
            st["feet_air_time"] += p.dt
            r = torch.sum((st["feet_air_time"] - 0.5) * first, dim=1)
            r *= torch.norm(cmd[:, :2], dim=1) > 0.1
            st["feet_air_time"] *= ~filt
            return r
        if name == "feet_contact_forces":
            return torch.sum((torch.norm(F[:, self.feet, :], dim=-1) - p.max_contact_force).clip(min=0.), dim=1)
        if name == "heading_alignment":                                    # go2.py:734-756
            fwd = quat_apply(root[:, 3:7], self.forward_vec)
            heading = torch.atan2(fwd[:, 1], fwd[:, 0])
            desired = wrap_to_pi(cmd[:, 3]) if p.heading_command else torch.zeros_like(heading)  # in place on cmd[:,3]
            err = wrap_to_pi(desired - heading)
            return sq(err) * (cmd_norm3 >= 0.2).float()
        if name == "hip_pos":
            return torch.sum(sq(dq[:, self.hip_j]), dim=1)
        if name == "jump_zone_forward_vel":
            return torch.clamp(root[:, 7], min=0.0) * (st["jump_flags"][:, 0] > 0.0).float() * (cmd_norm3 >= 0.2).float()
        if name == "jump_zone_upward_vel":
            return torch.clamp(root[:, 9], min=0.0) * (st["jump_flags"][:, 0] > 0.0).float() * (cmd_norm3 >= 0.2).float()
        if name == "lin_vel_z":
            return sq(o["base_lin_vel"][:, 2])
        if name == "min_height":
            z_err = torch.clip(p.base_height_target - root[:, 2], min=0.0, max=p.base_height_target)
            return z_err * (st["jump_flags"][:, 0] > 0.0).float()
        if name == "orientation":
            return torch.sum(sq(o["projected_gravity"][:, :2]), dim=1)
        if name == "phase_contact_match":                                  # go2.py:621-644
            r = torch.zeros(self.N)
            for leg in ("fl", "fr", "bl", "br"):
                r += torch.where(~(o[leg + "_contact"] ^ stance(o["phase_" + leg])), 0.25, -0.25)
            return r
        if name == "phase_foot_lifting":                                   # go2.py:647-678
            fz = st["rigid_body_states"].view(self.N, NUM_BODIES, 13)[:, self.feet, 2]
            h = torch.clamp(fz - st["last_contact_heights"], min=0.0, max=p.max_foot_height)
            swing = torch.stack([~stance(o["phase_" + leg]) for leg in ("fl", "fr", "bl", "br")], dim=1)
            nh = h / p.max_foot_height
            return torch.sum(torch.where(swing, nh, -nh), dim=1) / 2.0
        if name == "reverse_penalty":
            return -torch.clamp(root[:, 7], max=0.0)
        if name == "stand_still":
            return torch.sum(torch.abs(dq), dim=1) * (torch.norm(cmd[:, :2], dim=1) < 0.1)
        if name == "stumble_calves":
            return torch.any(torch.norm(F[:, self.calves, :2], dim=2) > 5 * torch.abs(F[:, self.calves, 2]), dim=1)
        if name == "stumble_feet":
            return torch.any(torch.norm(F[:, self.feet, :2], dim=2) > 5 * torch.abs(F[:, self.feet, 2]), dim=1)
        if name == "thigh_pos":
            return torch.sum(sq(dq[:, self.thigh_j]), dim=1)
        if name == "thigh_symmetry":
            c = self.thigh_j
            return torch.sum(torch.abs(self.dof_pos[:, [c[0], c[2]]] - self.dof_pos[:, [c[1], c[3]]]), dim=1)
        if name == "torque_limits":
            return torch.sum((torch.abs(st["torques"]) - self.torque_limits * p.soft_torque_limit).clip(min=0.), dim=1)
        if name == "torques":
            return torch.sum(sq(st["torques"]), dim=1)
        if name == "tracking_ang_vel":
            return torch.exp(-sq(cmd[:, 2] - o["base_ang_vel"][:, 2]) / p.tracking_sigma)
        if name == "tracking_lin_vel":
            return torch.exp(-torch.sum(sq(cmd[:, :2] - o["base_lin_vel"][:, :2]), dim=1) / p.tracking_sigma)
        if name == "tracking_pitch":
            return torch.exp(-sq(o["pitch"] * (180.0 / math.pi) - p.pitch_deg_target) / p.tracking_sigma)
        if name == "tracking_roll":
            return torch.exp(-sq(o["roll"] * (180.0 / math.pi) - p.roll_deg_target) / p.tracking_sigma)
        if name == "zero_cmd_dof_error":
            return torch.sum(sq(dq), dim=1) * (cmd_norm3 < 0.2).float()
        raise KeyError(name)

    # ---- command curriculum: go2.py:80-107 ------------------------------------------------------
    def update_command_curriculum(self, env_ids):
        p, st = self.p, self.st
        mean = torch.mean(st["episode_sums"][REWARD_INDEX["tracking_lin_vel"]][env_ids]) / p.max_episode_length
        lo, hi = (float(v) for v in st["command_ranges"])
        delta = p.cc_vel_increment
        if mean > torch.tensor(p.cc_threshold):                            # fp32 tensor vs Python float: compared in fp32
            if p.cc_max_reverse_vel < 0.0:
                lo = np.clip(lo - delta, p.cc_max_reverse_vel, 0.)
            else:                                                          # go2.py:100-103: a_max is the value itself
                lo = np.clip(lo - delta, p.cc_max_reverse_vel, lo - delta)
            hi = np.clip(hi + delta, 0., p.cc_max_forward_vel)
        st["command_ranges"] = torch.tensor([float(lo), float(hi)], dtype=torch.float64)

    # ---- reset: go2.py:207-263, legged_robot.py:481-574 -------------------------------------
    def reset_idx(self, env_ids, init_done=True):
        p, st = self.p, self.st
        n = len(env_ids)
        if n == 0:
            return
        root = st["root_states"]
        if p.curriculum and init_done:                                     # legged_robot.py:543-574
            dist = torch.norm(root[env_ids, :2] - st["env_origins"][env_ids, :2], dim=1)
            up = dist > p.promote_dist
            expected = torch.norm(st["commands"][env_ids, :2], dim=1) * p.max_episode_length_s
            down = dist < expected * p.demote_threshold
            lv = st["terrain_levels"]
            lv[env_ids[up]] += 1
            lv[env_ids[down]] -= 1
            rnd = torch.from_numpy((self._u32(philox.SITE_CURRICULUM, env_ids, [0])[:, 0] % np.uint32(p.max_terrain_level)).astype(np.int64))
            lv[env_ids] = torch.where(lv[env_ids] >= p.max_terrain_level, rnd, torch.clip(lv[env_ids], 0))
            st["env_origins"][env_ids] = self.terrain_origins[lv[env_ids], st["terrain_types"][env_ids]]
        # go2.py:222-223 (reset() at step 0 sees all-zero episode sums: no move, so only in-step resets are restated)
        if p.command_curriculum and init_done and int(st["common_step_counter"]) % p.max_episode_length == 0:
            self.update_command_curriculum(env_ids)
        # _reset_dofs (legged_robot.py:481-506)
        u = self._u(philox.SITE_RESET_DOFS, env_ids, list(range(NUM_DOF)))
        dof = st["dof_state"].view(self.N, NUM_DOF, 2)
        dof[env_ids, :, 0] = self.default_dof_pos + (p.dof_reset_span * u + p.dof_reset_lo)
        dof[env_ids, :, 1] = 0.
        # _reset_root_states (legged_robot.py:509-532)
        # lanes: xy offset (only drawn with custom origins) first, then the six velocities
        v0 = 2 if p.custom_origins else 0
        u = self._u(philox.SITE_RESET_ROOT, env_ids, list(range(v0 + 6)))
        root[env_ids] = self.base_init_state
        root[env_ids, :3] += st["env_origins"][env_ids]
        if p.custom_origins:
            root[env_ids, :2] += 2.0 * u[:, 0:2] + -1.0
        root[env_ids, 7:13] = 1.0 * u[:, v0:v0 + 6] + -0.5
        self._resample_commands(env_ids, philox.SITE_CMD_RESET)
        for k in ("last_actions", "last_dof_vel", "last_root_vel", "last_base_lin_vel", "last_torques",
                  "feet_air_time", "last_contact_heights"):
            st[k][env_ids] = 0.
        st["obs_history_buf"][env_ids, :, :] = 0.
        st["episode_length_buf"][env_ids] = 0
        st["reset_buf"][env_ids] = True
        st["last_contacts"][env_ids] = False
        # extras (go2.py:246-263)
        ep = st["extras_episode"]
        for i, name in enumerate(REWARD_TERMS):
            if p.reward_scales[i] != 0.0:
                ep[i] = torch.mean(st["episode_sums"][i][env_ids]) / p.max_episode_length_s
                st["episode_sums"][i][env_ids] = 0.
        if p.curriculum:
            ep[len(REWARD_TERMS)] = torch.mean(st["terrain_levels"].float())
        st["extras_time_outs"] = st["time_out_buf"].clone()
        self.out["reset_count"] = n

    def reset_all(self, init_done=False):
        """BaseTask.reset's first half (base_task.py:131-133)."""
        self.reset_idx(torch.arange(self.N), init_done=init_done)

    # ---- go2.py:345-387 -----------------------------------------------------------------------
    def post_physics_step(self):
        p, st, o, N = self.p, self.st, self.out, self.N
        root = st["root_states"]
        st["episode_length_buf"] += 1
        st["common_step_counter"] += 1
        q = root[:, 3:7]
        o["base_lin_vel"] = quat_rotate_inverse(q, root[:, 7:10])
        o["base_ang_vel"] = quat_rotate_inverse(q, root[:, 10:13])
        o["projected_gravity"] = quat_rotate_inverse(q, self.gravity_vec)

        # update_feet_states (go2.py:266-328)
        ph = (st["episode_length_buf"] * p.dt) % p.period / p.period
        o["phase"] = ph
        small = torch.norm(st["commands"][:, :3], dim=1) < 0.2
        keep = torch.where(small, 0.0, 1.0)
        for leg, off in (("fr", p.fr_offset), ("bl", p.bl_offset), ("fl", p.fl_offset), ("br", p.br_offset)):
            o["phase_" + leg] = ((ph + off) % 1) * keep
        feet_z = st["rigid_body_states"].view(N, NUM_BODIES, 13)[:, self.feet, 2]
        cur = self.contact[:, self.feet, 2] > 1.0
        filt = torch.logical_or(cur, st["last_contacts"])
        for i, leg in enumerate(("fl", "fr", "bl", "br")):
            o[leg + "_contact"] = filt[:, i]
        st["last_contacts"] = cur.clone()
        st["last_contact_heights"] = torch.where(filt, feet_z, st["last_contact_heights"])

        o["roll"], o["pitch"], o["yaw"] = euler_from_quat(q)

        # _post_physics_step_callback (go2.py:390-410)
        ids = (st["episode_length_buf"] % p.resample_interval == 0).nonzero(as_tuple=False).flatten()
        self._resample_commands(ids, philox.SITE_CMD_PERIODIC)
        if p.heading_command:
            fwd = quat_apply(q, self.forward_vec)
            heading = torch.atan2(fwd[:, 1], fwd[:, 0])
            st["commands"][:, 2] = torch.clip(wrap_to_pi(st["commands"][:, 3] - heading) * p.heading_error_gain, -1., 1.)
        o["measured_heights"] = self.get_heights()
        if p.push_robots and int(st["common_step_counter"]) % p.push_interval == 0:      # legged_robot.py:535-540
            u = self._u(philox.SITE_PUSH, torch.arange(N), [0, 1])
            mv = p.max_push_vel
            root[:, 7:9] = (mv - -mv) * u + -mv

        # check_termination (go2.py:186-204)
        reset = torch.any(torch.norm(self.contact[:, self.term, :], dim=-1) > 1., dim=1)
        st["time_out_buf"] = st["episode_length_buf"] > p.max_episode_length
        reset |= st["time_out_buf"]
        reset |= o["projected_gravity"][:, 2] > 0.
        if p.parkour:
            reset |= root[:, 2] < -1.0
        st["reset_buf"] = reset

        # compute_reward (legged_robot.py:216-237)
        rew = torch.zeros(N)
        mag = torch.zeros(N)          # sum of |terms|: the scale rounding errors of the (cancelling) sum live on
        for name in self.active:
            i = REWARD_INDEX[name]
            r = self._reward(name) * p.reward_scales[i]
            rew += r
            mag += r.abs()
            st["episode_sums"][i] += r
        o["rew_terms_abs"] = mag
        if p.only_positive_rewards:
            rew = torch.clip(rew, min=0.)
        ti = REWARD_INDEX["termination"]
        if p.reward_scales[ti] != 0.0:
            r = (st["reset_buf"] * ~st["time_out_buf"]) * p.reward_scales[ti]
            rew += r
            st["episode_sums"][ti] += r
        o["rew_buf"] = rew

        o["reset_count"] = 0
        self.reset_idx(st["reset_buf"].nonzero(as_tuple=False).flatten())
        self.compute_observations()

        st["last_actions"] = st["actions"].clone()
        st["last_dof_vel"] = self.dof_vel.clone()
        st["last_root_vel"] = root[:, 7:13].clone()
        st["last_base_lin_vel"] = o["base_lin_vel"].clone()
        st["last_torques"] = st["torques"].clone()

        c = p.clip_obs                                                     # legged_robot.py:91-95
        for k in ("obs_buf", "privileged_obs_buf", "critic_obs_buf", "estimated_obs_buf"):
            o[k] = torch.clip(o[k], -c, c)
        o["reset_buf"], o["time_out_buf"] = st["reset_buf"], st["time_out_buf"]

    # ---- go2.py:467-574 -------------------------------------------------------------------------
    def compute_observations(self):
        p, st, o, N = self.p, self.st, self.out, self.N
        feat = []
        for leg in ("fr", "fl", "bl", "br"):
            feat += [torch.sin(TWO_PI * o["phase_" + leg]), torch.cos(TWO_PI * o["phase_" + leg])]
        if p.parkour:
            outliers = torch.sum(torch.abs(o["measured_heights"]) > 0.1, dim=1)
            st["jump_flags"] = (outliers >= 8).unsqueeze(1).float()
        cur = torch.cat((o["base_ang_vel"] * p.obs_ang_vel,
                         torch.stack((o["roll"], o["pitch"]), dim=1),
                         st["commands"][:, :3] * self.commands_scale,
                         (self.dof_pos - self.default_dof_pos) * p.obs_dof_pos,
                         self.dof_vel * p.obs_dof_vel,
                         st["actions"],
                         torch.stack(feat, dim=1)), dim=-1)
        if p.add_noise:
            u = self._u(philox.SITE_OBS_NOISE, torch.arange(N), philox.noise_lane(np.arange(p.num_proprio)))
            cur += (2 * u - 1) * self.noise_vec
        o["obs_buf"] = torch.cat([st["obs_history_buf"].view(N, -1), cur], dim=-1)
        o["privileged_obs_buf"] = torch.cat((self.mass, self.fric, self.kp_kd[0] - 1, self.kp_kd[1] - 1), dim=-1)
        o["estimated_obs_buf"] = o["base_lin_vel"] * p.obs_lin_vel
        o["scan_obs_buf"] = torch.clip(st["root_states"][:, 2].unsqueeze(1) - 0.3 - o["measured_heights"], -1, 1.)
        o["critic_obs_buf"] = torch.cat((o["obs_buf"], o["privileged_obs_buf"], o["estimated_obs_buf"], o["scan_obs_buf"]), dim=-1)
        st["obs_history_buf"] = torch.where((st["episode_length_buf"] <= 1)[:, None, None],
                                            torch.stack([cur] * p.history_len, dim=1),
                                            torch.cat([st["obs_history_buf"][:, 1:], cur.unsqueeze(1)], dim=1))
