"""The oracle must reproduce the reference run stored in tests/golden BIT-FOR-BIT (CPU fp32
torch ops in the same order).  This is what pins oracle/go2_oracle.py to the reference."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle.go2_oracle import Go2Oracle

OUT_KEYS = ("base_lin_vel", "base_ang_vel", "projected_gravity", "roll", "pitch", "yaw", "measured_heights",
            "phase", "phase_fr", "phase_fl", "phase_bl", "phase_br", "fl_contact", "fr_contact", "bl_contact",
            "br_contact", "rew_buf", "obs_buf", "privileged_obs_buf", "estimated_obs_buf", "scan_obs_buf",
            "height_index")
STATE_KEYS = ("actions", "torques", "commands", "episode_length_buf", "last_actions", "last_dof_vel", "last_root_vel",
              "last_base_lin_vel", "last_torques", "last_contacts", "last_contact_heights", "jump_flags",
              "terrain_levels", "env_origins", "reset_buf", "time_out_buf", "feet_air_time", "root_states",
              "dof_state", "episode_sums")


def _same(a, b, name, t):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (name, t, a.shape, b.shape)
    if a.dtype.kind == "f":
        ok = (a.view(np.int32) == b.astype(np.float32).view(np.int32)) | ((a == 0) & (b == 0))
    else:
        ok = a.astype(np.int64) == b.astype(np.int64)
    assert ok.all(), f"{name} step {t}: {int((~ok).sum())} of {ok.size} differ; max abs {np.abs(a.astype(np.float64) - b).max()}"


@pytest.mark.parametrize("task", gu.TASKS + gu.CC_SCENARIOS)
def test_oracle_reproduces_reference(task):
    g = gu.load(task)
    p = gu.params_for(task, g)
    replay_oracle(g, p, gu.statics_for(task, g))


def replay_oracle(g, p, statics):
    """step the oracle through the replay `g` (npz or dict-like with .files) and require every tensor bit for bit"""
    orc = Go2Oracle(p, statics, gu.init_state(g, p))
    for t in range(int(g["steps"])):
        out = orc.step(torch.from_numpy(g[f"step{t}/in/actions"]), gu.frames_of(g, t))
        exp = gu.expected(g, t)
        for k in OUT_KEYS:
            if k in exp:
                _same(out[k], exp[k], k, t)
        for k in STATE_KEYS:
            if k in exp:
                _same(orc.st[k], exp[k], k, t)
        assert out["reset_count"] == int(exp["n_reset"])
        if "command_ranges" in exp:                       # command curriculum: Python floats moved by np.clip in double
            assert orc.st["command_ranges"].tolist() == exp["command_ranges"].tolist(), (t, orc.st["command_ranges"], exp["command_ranges"])
        assert gu.critic_sha(out["obs_buf"].numpy(), out["privileged_obs_buf"].numpy(), out["estimated_obs_buf"].numpy(),
                             out["scan_obs_buf"].numpy()) == str(exp["critic_sha"])
        _same(out["critic_obs_buf"], np.concatenate([exp["obs_buf"], exp["privileged_obs_buf"], exp["estimated_obs_buf"],
                                                      exp["scan_obs_buf"]], axis=-1), "critic_obs_buf", t)
        if "extras/time_outs" in exp:
            _same(orc.st["extras_time_outs"], exp["extras/time_outs"], "extras_time_outs", t)
        for i, v in exp["extras_episode"].items():
            assert np.float32(orc.st["extras_episode"][i].item()) == np.float32(v), ("extras_episode", i, t)
    _same(orc.st["obs_history_buf"], g["final/obs_history_buf"], "obs_history_buf", "final")


@pytest.mark.parametrize("task", gu.TASKS)
def test_params_match_reference_constants(task):
    """cfg packing (params.py + configs.py) against what the reference resolved at runtime."""
    g = gu.load(task)
    p = gu.params_for(task, g)
    assert p.reward_names() == [str(n) for n in g["static/reward_names"]]
    ref_scales = g["static/reward_scales"]
    for n, s in zip(p.reward_names(), ref_scales):
        assert np.float32(p.reward_scales[gu.REWARD_INDEX[n]]) == np.float32(s), n
    np.testing.assert_array_equal(np.array(p.noise_vec[:p.num_proprio], dtype=np.float32), g["static/noise_scale_vec"])
    np.testing.assert_array_equal(np.array(p.default_dof_pos, dtype=np.float32), g["static/default_dof_pos"].reshape(-1))
    np.testing.assert_array_equal(np.array(p.torque_limits, dtype=np.float32), g["static/torque_limits"])
    np.testing.assert_array_equal(np.array(p.p_gains, dtype=np.float32), g["static/p_gains"])
    np.testing.assert_array_equal(np.array(p.d_gains, dtype=np.float32), g["static/d_gains"])
    np.testing.assert_array_equal(np.array(p.dof_pos_lo, dtype=np.float32), g["static/dof_pos_limits"][:, 0])
    np.testing.assert_array_equal(np.array(p.dof_pos_hi, dtype=np.float32), g["static/dof_pos_limits"][:, 1])
    assert list(p.feet) == list(g["static/feet_indices"])
    assert list(p.penalised)[:p.n_penalised] == list(g["static/penalised_contact_indices"])
    assert list(p.termination)[:p.n_termination] == list(g["static/termination_contact_indices"])
