"""Pins the oracle's NON-default branches to the reference itself, run live (authoring container only: skipped where
/root/reference is absent).  The committed golden replays cover the cfg values the go2 tasks ship; here the reference's
`Go2Robot` is built with one cfg switch flipped (or another control type), stepped on seeded synthetic frames exactly like
oracle/make_golden.py does, and the oracle must reproduce every tensor bit for bit."""
import os

import pytest

import golden_util as gu
from legged_gym_custom_b200 import configs
from legged_gym_custom_b200.params import env_params_from_cfg
from oracle import ref_runner
from test_oracle_golden import replay_oracle

pytestmark = pytest.mark.skipif(not ref_runner.reference_available(), reason="reference not present (GPU box)")


class _Replay(dict):
    @property
    def files(self):
        return list(self.keys())


def _flip(group, name, value, undo):
    """the registry hands out ONE cfg object per task: remember the old value so the test can put it back"""
    def patch(cfg):
        sub = getattr(cfg, group)
        undo.append((sub, name, getattr(sub, name)))
        setattr(sub, name, value)
    return patch


CASES = {
    "noise-off": ("go2_parkour", "noise", "add_noise", False),
    "negative-rewards-kept": ("go2_parkour", "rewards", "only_positive_rewards", False),
    "no-pushes": ("go2_parkour", "domain_rand", "push_robots", False),
    "no-zero-commands": ("go2_parkour", "commands", "zero_command", False),
    "velocity-control": ("go2", "control", "control_type", "V"),
    "torque-control": ("go2", "control", "control_type", "T"),
    "kp-kd-not-randomised": ("go2", "domain_rand", "randomize_kp_kd", False),
}


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_live_reference_with_switch_flipped(case):
    from oracle import make_golden
    task, group, name, value = CASES[case]
    undo = []
    try:
        data, _ = make_golden.run(task, 12, 5, cfg_patch=_flip(group, name, value, undo))
    finally:
        for sub, attr, old in undo:
            setattr(sub, attr, old)
    g = _Replay(data)
    base = configs.TASKS[task][0]
    sub = getattr(base, group)
    assert getattr(sub, name) != value, "the switch must differ from the shipped value"
    cfg = type("Cfg", (base,), {group: type(group, (sub,), {name: value})})
    hs, _ = gu.terrain_for(task)
    p = env_params_from_cfg(cfg, num_envs=int(g["num_envs"]), seed=int(g["seed"]), hs_shape=None if hs is None else hs.shape)
    replay_oracle(g, p, gu.statics_for(task, g))
