"""Pins oracle/learner_oracle.py to the reference's own rsl_rl code (runs only where /root/reference
exists, i.e. the authoring container): same weights, same storage, same minibatch indices ->
identical GAE, act statistics, losses and post-update parameters for PPO.update and update_dagger."""
import os
import sys

import pytest
import torch

import learner_util as lu
from oracle import learner_oracle as lo

REF = "/root/reference/rsl_rl"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present (GPU box)")


def _reference_ppo(T, N, seed=0, schedule='fixed'):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from rsl_rl.algorithms import PPO
    from rsl_rl.modules import ActorCritic
    from rsl_rl.modules.support_networks import MlpEstimator
    torch.manual_seed(seed)
    ac = ActorCritic(52, 29, 736, 3, 132, 12, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128],
                     priv_encoder_hidden_dims=[64, 20], scan_encoder_hidden_dims=[128, 64], latent_encoder_output_dim=20,
                     scan_encoder_output_dim=32, activation='elu', init_noise_std=1.0)
    est = MlpEstimator(52, 10, 3, hidden_dims=[256, 128], activation='elu', use_history=True)
    ppo = PPO(ac, est, num_learning_epochs=2, num_mini_batches=2, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=2e-4, estimator_learning_rate=1e-4, max_grad_norm=1.0, use_clipped_value_loss=True,
              schedule=schedule, desired_kl=0.01, resume=True, device='cpu')       # resume=True -> ROA coef 0.1 from the second update
    ppo.init_storage(N, T, [572], [29], [736], [3], [132], [12])
    return ppo


def _fill(ppo, st):
    s = ppo.storage
    s.observations.copy_(st["obs"]); s.privileged_observations.copy_(st["priv"]); s.critic_observations.copy_(st["critic_obs"])
    s.true_estimated_observations.copy_(st["true_est"]); s.scan_observations.copy_(st["scan"]); s.actions.copy_(st["actions"])
    s.values.copy_(st["values"]); s.returns.copy_(st["returns"]); s.advantages.copy_(st["adv"])
    s.actions_log_prob.copy_(st["old_logp"]); s.mu.copy_(st["mu"]); s.sigma.copy_(st["sigma"])


def test_gae_matches_reference_storage():
    T, N = 24, 64
    ppo = _reference_ppo(T, N)
    g = torch.Generator().manual_seed(3)
    s = ppo.storage
    s.rewards.copy_(torch.rand(T, N, 1, generator=g)); s.values.copy_(torch.randn(T, N, 1, generator=g))
    s.dones.copy_((torch.rand(T, N, 1, generator=g) < 0.05).byte())
    last = torch.randn(N, 1, generator=g)
    s.compute_returns(last, 0.99, 0.95)
    ret, adv = lo.compute_returns(s.rewards, s.dones, s.values, last, 0.99, 0.95)
    assert torch.equal(ret, s.returns) and torch.equal(adv, s.advantages)


@pytest.mark.parametrize("total_updates", [0.0, 3.0])
def test_update_matches_reference(total_updates, monkeypatch):
    T, N = 6, 32
    ppo = _reference_ppo(T, N)
    ppo.total_updates = total_updates                      # 0 -> reg coef 0.0 ; 3 -> 0.1 (resume schedule)
    st = lu.random_storage(T, N, seed=1)
    _fill(ppo, st)
    sd = {k: v.detach().clone() for k, v in ppo.actor_critic.state_dict().items()}
    sd_est = {k: v.detach().clone() for k, v in ppo.estimator.state_dict().items()}
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(5))
    monkeypatch.setattr(torch, "randperm", lambda *a, **k: perm.clone())
    v, sur, reg, coef, el = ppo.update()
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4)
    stage = min(max((total_updates - 0) / 1, 0.0), 1.0)
    assert coef == 0.1 * stage
    logs = []
    mb = T * N // 2
    for _ in range(2):
        for i in range(2):
            logs.append(orc.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=coef))
    mean = lambda k: sum(l[k] for l in logs) / len(logs)
    assert abs(mean("value") - v) <= 1e-6 * abs(v) and abs(mean("surrogate") - sur) <= 1e-6 * abs(sur)
    assert abs(mean("reg") - reg) <= 1e-6 * abs(reg) and abs(mean("estimator") - el) <= 1e-6 * abs(el)
    new = ppo.actor_critic.state_dict()
    for k in orc.main_keys:
        ref = new[k]
        mine = orc.sd[k].detach() if k != "std" else torch.min(orc.sd[k].detach(), torch.tensor(1.0))   # enforce_max_std
        assert torch.allclose(mine, ref, rtol=1e-6, atol=5e-8), k
    for k, ref in ppo.estimator.state_dict().items():
        assert torch.allclose(orc.sd_est[k].detach(), ref, rtol=1e-6, atol=5e-8), k


@pytest.mark.parametrize("near", [False, True], ids=["old-policy-far:lr-shrinks", "old-policy-near:lr-grows"])
def test_adaptive_schedule_matches_reference(near, monkeypatch):
    """schedule='adaptive' (ppo.py:233-246): same learning-rate trajectory and post-update parameters.  `near`: the stored
    mu / sigma are the current policy's own (KL ~ 1e-4 < desired/2 -> lr x 1.5 until the policy has moved), else random (KL >> 2 desired)."""
    T, N = 6, 32
    ppo = _reference_ppo(T, N, schedule='adaptive')
    st = lu.random_storage(T, N, seed=2)
    sd = {k: v.detach().clone() for k, v in ppo.actor_critic.state_dict().items()}
    sd_est = {k: v.detach().clone() for k, v in ppo.estimator.state_dict().items()}
    if near:
        f = lambda t: t.flatten(0, 1)
        with torch.no_grad():
            mu = lo.actor_mean(sd, f(st["obs"]), f(st["priv"]), f(st["true_est"]), f(st["scan"]))
        st["mu"], st["sigma"] = mu.view(T, N, -1).clone(), (mu * 0. + sd["std"]).view(T, N, -1).clone()
    _fill(ppo, st)
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(6))
    monkeypatch.setattr(torch, "randperm", lambda *a, **k: perm.clone())
    ppo.update()
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4, desired_kl=0.01)
    mb = T * N // 2
    for _ in range(2):
        for i in range(2):
            orc.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=0.0)
    assert orc.lr == ppo.learning_rate, orc.kl_log
    if near:      # grows while the policy is still close, then holds (desired/2 <= KL <= 2 desired) once it has moved
        assert ppo.learning_rate > 2e-4
    else:
        assert ppo.learning_rate == pytest.approx(2e-4 / 1.5 ** 4, rel=1e-12)
    assert ppo.optimizer.param_groups[0]['lr'] == orc.lr
    new = ppo.actor_critic.state_dict()
    for k in orc.main_keys:
        mine = orc.sd[k].detach() if k != "std" else torch.min(orc.sd[k].detach(), torch.tensor(1.0))
        # the 1-ulp-per-Adam-step difference of the other update tests, at up to 2.25 x their learning rate
        assert torch.allclose(mine, new[k], rtol=1e-6, atol=2e-7), (k, float((mine - new[k]).abs().max()))


def test_update_dagger_matches_reference(monkeypatch):
    T, N = 6, 32
    ppo = _reference_ppo(T, N)
    st = lu.random_storage(T, N, seed=2)
    _fill(ppo, st)
    sd = {k: v.detach().clone() for k, v in ppo.actor_critic.state_dict().items()}
    sd_est = {k: v.detach().clone() for k, v in ppo.estimator.state_dict().items()}
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(6))
    monkeypatch.setattr(torch, "randperm", lambda *a, **k: perm.clone())
    loss = ppo.update_dagger()
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4)
    mb, losses = T * N // 2, []
    for _ in range(2):
        for i in range(2):
            losses.append(orc.dagger_minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb])))
    assert abs(sum(losses) / 4 - loss) <= 1e-6 * abs(loss)
    new = ppo.actor_critic.state_dict()
    for k in orc.adapt_keys:
        assert torch.allclose(orc.sd[k].detach(), new[k], rtol=1e-6, atol=5e-8), k


def test_act_statistics_match_reference():
    """estimator -> latent -> actor -> Normal: mu, values, log-prob of given actions (sampling itself is keyed on our side)."""
    T, N = 2, 48
    ppo = _reference_ppo(T, N)
    st = lu.random_storage(T, N, seed=4)
    b = lu.minibatch(st, torch.arange(N))
    sd = dict(ppo.actor_critic.state_dict())
    sd_est = dict(ppo.estimator.state_dict())
    for mode in (False, True):
        with torch.no_grad():
            est_hat = ppo.estimator(b["obs"])
            ppo.actor_critic.update_distribution(b["obs"], b["priv"], est_hat, b["scan"], adaptation_mode=mode)
            mu_ref, val_ref = ppo.actor_critic.action_mean, ppo.actor_critic.evaluate(b["critic_obs"])
            lp_ref = ppo.actor_critic.get_actions_log_prob(b["actions"])
            mu = lo.actor_mean(sd, b["obs"], b["priv"], lo.estimator(sd_est, b["obs"]), b["scan"], mode)
            assert torch.equal(mu, mu_ref) and torch.equal(lo.critic_value(sd, b["critic_obs"]), val_ref)
            assert torch.equal(lo.normal_log_prob(b["actions"], mu, mu * 0. + sd["std"]), lp_ref)
