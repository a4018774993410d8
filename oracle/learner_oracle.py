"""CPU oracle for the learner half of the hot path (rsl_rl): GAE, the time-out bootstrap, the
networks and the PPO / DAgger losses.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.

Plain fp32 torch-CPU restatement; each function cites the reference lines it follows.  Pinned by
tests/test_learner_oracle_vs_reference.py, which (in the authoring container, where
/root/reference exists) runs the reference's own RolloutStorage / ActorCritic / PPO on the same
tensors and requires identical results, and by tests/golden/learner_*.npz on the GPU box.
"""
import torch
import torch.nn.functional as F


def compute_returns(rewards, dones, values, last_values, gamma, lam):
    """rollout_storage.py:110-124.  rewards/values [T,N,1] fp32, dones [T,N,1] uint8, last_values [N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + not_terminal * gamma * next_values - values[step]
        advantage = delta + not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    adv = returns - values
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    return returns, adv


def bootstrap_rewards(rew, values, time_outs, gamma):
    """ppo.py:160-166: rewards += gamma * squeeze(values * time_outs.unsqueeze(1), 1)."""
    out = rew.clone()
    if time_outs is not None:
        out += gamma * torch.squeeze(values * time_outs.unsqueeze(1), 1)
    return out
