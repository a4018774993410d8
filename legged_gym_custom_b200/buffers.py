"""Allocation of the B200EnvBuffers block (include/b200gym.h) as torch tensors.

All memory is owned by PyTorch on the Python side and outlives every library call
(SURVEY.md §8(b) ownership rule); the struct handed to the C ABI only carries data_ptr()s.
`device` is a CUDA device for the product path; tests/host_emul passes 'cpu' to run the
kernel source under the host emulator.
"""
import ctypes as C

import torch

from .params import BUFFER_FIELDS, EnvBuffers, NUM_BODIES, NUM_DOF, NUM_REWARD_TERMS


def buffer_specs(p):
    """name -> (shape, dtype) for every pointer of B200EnvBuffers."""
    N, NP, H, NS = p.num_envs, p.num_proprio, p.history_len, p.num_scan
    obs = NP * (H + 1)
    crit = obs + p.num_priv + p.num_est + NS
    f, i64, b = torch.float32, torch.int64, torch.bool
    return {
        "root_states": ((N, 13), f), "dof_state": ((N * NUM_DOF, 2), f), "contact_forces": ((N * NUM_BODIES, 3), f),
        "rigid_body_states": ((N * NUM_BODIES, 13), f), "kp_kd_multipliers": ((2, N, NUM_DOF), f),
        "priv_mass_params": ((N, 4), f), "priv_friction": ((N, 1), f),
        "height_samples": ((p.hs_rows, p.hs_cols), torch.int16) if p.has_height_samples else None,
        "terrain_origins": ((p.max_terrain_level, p.terrain_cols, 3), f) if p.has_height_samples else None,
        "actions": ((N, NUM_DOF), f), "torques": ((N, NUM_DOF), f), "commands": ((N, 4), f),
        "episode_length_buf": ((N,), i64), "last_actions": ((N, NUM_DOF), f), "last_dof_vel": ((N, NUM_DOF), f),
        "last_root_vel": ((N, 6), f), "last_base_lin_vel": ((N, 3), f), "last_torques": ((N, NUM_DOF), f),
        "obs_history_buf": ((N, H, NP), f), "last_contacts": ((N, 4), b), "last_contact_heights": ((N, 4), f),
        "feet_air_time": ((N, 4), f), "jump_flags": ((N, 1), f), "episode_sums": ((N, NUM_REWARD_TERMS), f),
        "terrain_levels": ((N,), i64), "terrain_types": ((N,), i64), "env_origins": ((N, 3), f),
        "base_lin_vel": ((N, 3), f), "base_ang_vel": ((N, 3), f), "projected_gravity": ((N, 3), f), "rpy": ((N, 3), f),
        "measured_heights": ((N, NS), f), "height_index": None, "phases": ((N, 5), f), "foot_contacts": ((N, 4), b),
        "obs_buf": ((N, obs), f), "privileged_obs_buf": ((N, p.num_priv), f), "critic_obs_buf": ((N, crit), f),
        "estimated_obs_buf": ((N, p.num_est), f), "scan_obs_buf": ((N, NS), f), "rew_buf": ((N,), f),
        "reset_buf": ((N,), b), "time_out_buf": ((N,), b), "extras_time_outs": ((N,), b),
        "extras_episode": ((NUM_REWARD_TERMS + 1,), f), "reset_count": ((1,), torch.int32),
        "reset_episode_sums": ((N, NUM_REWARD_TERMS), f),
        "command_ranges": ((4,), torch.float64), "cc_value": ((N,), f), "cc_reset": ((N,), b),
    }


class BufferSet:
    """Named torch tensors + the ctypes struct that points at them."""

    def __init__(self, p, device, record_height_index=False):
        self.p, self.device = p, torch.device(device)
        self.t = {}
        for name, spec in buffer_specs(p).items():
            if name == "height_index" and record_height_index:
                spec = ((p.num_envs, p.num_scan, 2), torch.int64)
            self.t[name] = None if spec is None else torch.zeros(spec[0], dtype=spec[1], device=self.device)
        if self.t["height_samples"] is not None and p.hs_pitch > p.hs_cols:      # rows padded to params.hs_pitch elements
            self._hs_storage = torch.zeros(p.hs_rows, p.hs_pitch, dtype=torch.int16, device=self.device)
            self.t["height_samples"] = self._hs_storage[:, :p.hs_cols]
        self.t["root_states"][:, 6] = 1.0
        self.t["reset_buf"].fill_(True)                      # base_task.py:83
        self.set_command_range(p.cc_range0[0], p.cc_range0[1])
        self.struct = EnvBuffers()
        if p.alias_outputs:
            self._alias_views(self.t["critic_obs_buf"])
        self.refresh_pointers()

    ALIASED = ("obs_buf", "privileged_obs_buf", "estimated_obs_buf", "scan_obs_buf")

    def _alias_views(self, rows):
        """alias_outputs: obs / priv / est / scan ARE column slices of the critic rows (go2.py:538-563 concatenates them)"""
        p = self.p
        obs = p.num_proprio * (p.history_len + 1)
        c0, c1 = obs + p.num_priv, obs + p.num_priv + p.num_est
        self.t["obs_buf"], self.t["privileged_obs_buf"] = rows[:, :obs], rows[:, obs:c0]
        self.t["estimated_obs_buf"], self.t["scan_obs_buf"] = rows[:, c0:c1], rows[:, c1:]

    def bind_output_rows(self, rows):
        """alias_outputs only: the step's observation rows land in `rows` [N, critic width] from now on -- e.g. a slot of the
        rollout storage, so that the env writes the learner's transition in place"""
        assert self.p.alias_outputs, "bind_output_rows needs alias_outputs"
        old = self.t["critic_obs_buf"]
        assert rows.shape == old.shape and rows.dtype == old.dtype and rows.device == old.device and rows.is_contiguous()
        self.t["critic_obs_buf"] = rows
        self._alias_views(rows)
        for name in ("critic_obs_buf",) + self.ALIASED:
            setattr(self.struct, name, C.c_void_p(self.t[name].data_ptr()))

    def refresh_pointers(self):
        for name in BUFFER_FIELDS:
            t = self.t[name]
            if t is not None and not (self.p.alias_outputs and name in self.ALIASED) and name != "height_samples":
                assert t.is_contiguous(), name
            setattr(self.struct, name, None if t is None else C.c_void_p(t.data_ptr()))

    def rebind(self, name, tensor):
        """Point a PhysX-owned slot at another tensor (zero-copy replay of pre-generated frames)."""
        old = self.t[name]
        assert tensor.shape == old.shape and tensor.dtype == old.dtype and tensor.device == old.device and tensor.is_contiguous()
        self.t[name] = tensor
        setattr(self.struct, name, C.c_void_p(tensor.data_ptr()))

    def rebind_host_mapped(self, name, tensor):
        """Point a READ-ONLY PhysX-owned slot at a pinned host tensor: under unified addressing the kernels read it in
        place over PCIe / NVLink-C2C (zero-copy).  Worth it only for sparsely read tensors (rigid_body_states: 4 of 247
        floats per env are read)."""
        old = self.t[name]
        assert tensor.shape == old.shape and tensor.dtype == old.dtype and tensor.is_pinned() and tensor.is_contiguous()
        setattr(self.struct, name, C.c_void_p(tensor.data_ptr()))

    def unbind_host_mapped(self, name):
        """back to the device tensor of the slot"""
        setattr(self.struct, name, C.c_void_p(self.t[name].data_ptr()))

    def set_command_range(self, lo, hi):
        """lin_vel_x range of the command curriculum (go2.py:80-107): {lo, hi} in force = {lo, hi} of this step's resets"""
        self.t["command_ranges"].copy_(torch.tensor([lo, hi, lo, hi], dtype=torch.float64))

    def __getitem__(self, name):
        return self.t[name]

    def load_state(self, st):
        """Copy a persistent-state dict (oracle / golden naming) into the buffers."""
        for k, v in st.items():
            if k == "episode_sums":
                self.t[k].copy_(torch.as_tensor(v).t())
            elif k == "command_ranges":
                lo, hi = (float(x) for x in torch.as_tensor(v).flatten()[:2])
                self.set_command_range(lo, hi)
            elif k in self.t and self.t[k] is not None:
                self.t[k].copy_(torch.as_tensor(v).reshape(self.t[k].shape).to(self.t[k].dtype))

    def load_statics(self, statics):
        self.t["kp_kd_multipliers"].copy_(torch.as_tensor(statics["kp_kd_multipliers"]))
        self.t["priv_mass_params"].copy_(torch.as_tensor(statics["privileged_mass_params"]))
        self.t["priv_friction"].copy_(torch.as_tensor(statics["privileged_friction_coeffs"]).reshape(-1, 1))
        if self.t["height_samples"] is not None:
            self.t["height_samples"].copy_(torch.as_tensor(statics["height_samples"]))
            self.t["terrain_origins"].copy_(torch.as_tensor(statics["terrain_origins"]))
