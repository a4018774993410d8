"""Generate tests/golden/learner_small.npz by running the reference's own rsl_rl classes (CPU).

TEST INFRASTRUCTURE; authoring container only.  `python -m oracle.make_golden_learner`

Small hidden sizes keep the fixture ~1 MB; the layer structure, observation widths and every code path
of PPO.update / update_dagger / RolloutStorage.compute_returns are the reference's.  Stored: initial state
dicts, the storage contents, the permutation handed to mini_batch_generator (torch.randperm is replaced by
a fixed tensor), the returned loss means and the post-update state dicts.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/rsl_rl")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import learner_util as lu  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "learner_small.npz")
HID = dict(actor=[64, 32, 16], critic=[64, 32, 16], priv=[16, 12], scan=[32, 16], est=[32, 16])
T, N, EPOCHS, MBS = 6, 32, 2, 2


def build(resume):
    from rsl_rl.algorithms import PPO
    from rsl_rl.modules import ActorCritic
    from rsl_rl.modules.support_networks import MlpEstimator
    torch.manual_seed(7)
    ac = ActorCritic(52, 29, 736, 3, 132, 12, 10, actor_hidden_dims=HID["actor"], critic_hidden_dims=HID["critic"],
                     priv_encoder_hidden_dims=HID["priv"], scan_encoder_hidden_dims=HID["scan"], latent_encoder_output_dim=20,
                     scan_encoder_output_dim=32, activation='elu', init_noise_std=0.8)
    est = MlpEstimator(52, 10, 3, hidden_dims=HID["est"], activation='elu', use_history=True)
    ppo = PPO(ac, est, num_learning_epochs=EPOCHS, num_mini_batches=MBS, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=2e-4, estimator_learning_rate=1e-4, max_grad_norm=1.0, use_clipped_value_loss=True,
              schedule='fixed', desired_kl=0.01, resume=resume, device='cpu')
    ppo.init_storage(N, T, [572], [29], [736], [3], [132], [12])
    return ppo


def fill(ppo, st):
    s = ppo.storage
    s.observations.copy_(st["obs"]); s.privileged_observations.copy_(st["priv"]); s.critic_observations.copy_(st["critic_obs"])
    s.true_estimated_observations.copy_(st["true_est"]); s.scan_observations.copy_(st["scan"]); s.actions.copy_(st["actions"])
    s.values.copy_(st["values"]); s.returns.copy_(st["returns"]); s.advantages.copy_(st["adv"])
    s.actions_log_prob.copy_(st["old_logp"]); s.mu.copy_(st["mu"]); s.sigma.copy_(st["sigma"])


def main():
    data = {}
    st = lu.random_storage(T, N, seed=11)
    for k, v in st.items():
        data["storage/" + k] = v.numpy()
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(13))
    data["perm"] = perm.numpy()
    real_randperm = torch.randperm
    torch.randperm = lambda *a, **k: perm.clone()
    try:
        # PPO.update with the ROA coefficient active (resume schedule, second update)
        ppo = build(resume=True)
        ppo.total_updates = 2.0
        fill(ppo, st)
        for k, v in ppo.actor_critic.state_dict().items():
            data["init/ac/" + k] = v.numpy().copy()
        for k, v in ppo.estimator.state_dict().items():
            data["init/est/" + k] = v.numpy().copy()
        out = ppo.update()
        data["update/returned"] = np.array(out, dtype=np.float64)
        for k, v in ppo.actor_critic.state_dict().items():
            data["update/ac/" + k] = v.numpy().copy()
        for k, v in ppo.estimator.state_dict().items():
            data["update/est/" + k] = v.numpy().copy()
        # PPO.update_dagger from the same initial weights
        ppo = build(resume=True)
        fill(ppo, st)
        data["dagger/returned"] = np.array([ppo.update_dagger()], dtype=np.float64)
        for k, v in ppo.actor_critic.state_dict().items():
            data["dagger/ac/" + k] = v.numpy().copy()
        # update_dagger THEN update on the same PPO object (what OnPolicyRunner.learn does: iteration 0 is a DAgger
        # iteration): the adaptation encoder's stale post-clip .grad enters every later clip_grad_norm_ (ppo.py:274).
        # Its own storage ("seq_storage/"): on-policy old log-probs / values and small advantages, so that the main
        # gradient norm is of the order of max_grad_norm -- with the 1000x larger norms of the storage above the stale
        # term would vanish in the total norm and the fixture would not pin it.
        ppo = build(resume=True)
        st2 = lu.random_storage(T, N, seed=12)
        with torch.no_grad():
            f = lambda t: t.flatten(0, 1)
            ppo.actor_critic.update_distribution(f(st2["obs"]), f(st2["priv"]), f(st2["true_est"]), f(st2["scan"]), adaptation_mode=False)
            g2 = torch.Generator().manual_seed(19)
            mu = ppo.actor_critic.action_mean
            st2["actions"] = (mu + 0.8 * torch.randn(mu.shape, generator=g2)).view(T, N, -1)
            st2["old_logp"] = ppo.actor_critic.get_actions_log_prob(f(st2["actions"])).view(T, N, 1).clone()
            st2["mu"] = mu.view(T, N, -1).clone()
            st2["sigma"] = ppo.actor_critic.action_std.reshape(T, N, -1).clone()
            st2["values"] = ppo.actor_critic.evaluate(f(st2["critic_obs"])).view(T, N, 1).clone()
            st2["returns"] = st2["values"] + 0.3 * torch.randn(T, N, 1, generator=g2)
            st2["adv"] = 1.0 * torch.randn(T, N, 1, generator=g2)
        for k, v in st2.items():
            data["seq_storage/" + k] = v.numpy()
        fill(ppo, st2)
        data["seq/dagger_returned"] = np.array([ppo.update_dagger()], dtype=np.float64)
        norms = []
        real_clip = torch.nn.utils.clip_grad_norm_
        def spy(params, max_norm, *a, **k):
            out = real_clip(params, max_norm, *a, **k)
            norms.append(float(out))
            return out
        torch.nn.utils.clip_grad_norm_ = spy
        import rsl_rl.algorithms.ppo as ppo_mod
        ppo_mod.nn.utils.clip_grad_norm_ = spy
        try:
            data["seq/update_returned"] = np.array(ppo.update(), dtype=np.float64)
        finally:
            torch.nn.utils.clip_grad_norm_ = real_clip
            ppo_mod.nn.utils.clip_grad_norm_ = real_clip
        data["seq/clip_total_norms"] = np.array(norms, dtype=np.float64)      # estimator, main, estimator, main, ...
        print("seq: total norms seen by clip_grad_norm_ (estimator / main alternating):", [round(n, 4) for n in norms])
        for k, v in ppo.actor_critic.state_dict().items():
            data["seq/ac/" + k] = v.numpy().copy()
        for k, v in ppo.estimator.state_dict().items():
            data["seq/est/" + k] = v.numpy().copy()
        # act statistics + GAE
        ppo = build(resume=True)
        b = lu.minibatch(st, torch.arange(N))
        with torch.no_grad():
            for mode in (False, True):
                ppo.actor_critic.update_distribution(b["obs"], b["priv"], ppo.estimator(b["obs"]), b["scan"], adaptation_mode=mode)
                data[f"act/mu_{int(mode)}"] = ppo.actor_critic.action_mean.numpy().copy()
                data[f"act/logp_{int(mode)}"] = ppo.actor_critic.get_actions_log_prob(b["actions"]).numpy().copy()
            data["act/value"] = ppo.actor_critic.evaluate(b["critic_obs"]).numpy().copy()
            data["act/est"] = ppo.estimator(b["obs"]).numpy().copy()
        g = torch.Generator().manual_seed(17)
        s = ppo.storage
        s.rewards.copy_(torch.rand(T, N, 1, generator=g)); s.values.copy_(torch.randn(T, N, 1, generator=g))
        s.dones.copy_((torch.rand(T, N, 1, generator=g) < 0.1).byte())
        last = torch.randn(N, 1, generator=g)
        s.compute_returns(last, 0.99, 0.95)
        for k, v in (("rewards", s.rewards), ("values", s.values), ("dones", s.dones), ("last_values", last), ("returns", s.returns),
                     ("advantages", s.advantages)):
            data["gae/" + k] = v.numpy().copy()
    finally:
        torch.randperm = real_randperm
    data["hid"] = np.array(str(HID))
    np.savez_compressed(OUT, **data)
    print(OUT, f"{os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
