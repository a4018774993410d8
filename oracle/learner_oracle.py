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


# ---- networks, restated functionally over reference-layout state dicts --------------------------
import math

import numpy as np

from . import philox


def mlp(sd, prefix, x, n_layers, final_act=False):
    """nn.Sequential(Linear, ELU, ..., Linear) with keys `<prefix>.<2i>.{weight,bias}`
    (actor_critic.py:84-108, support_networks.py:24-35, :70-80, :108-116)."""
    for i in range(n_layers):
        x = F.linear(x, sd[f"{prefix}.{2 * i}.weight"], sd[f"{prefix}.{2 * i}.bias"])
        if i < n_layers - 1 or final_act:
            x = F.elu(x)
    return x


def n_layers(sd, prefix):
    return len([k for k in sd if k.startswith(prefix + ".") and k.endswith(".weight")])


def adaptation_encoder(sd, obs, num_proprio=52, hist=10):
    """AdaptationEncoder.forward on obs[:, :-num_proprio] (actor_critic.py:174-180, support_networks.py:128-175)."""
    h = obs[:, :-num_proprio].reshape(-1, hist, num_proprio)
    p = "adaptation_encoder_."
    x = F.elu(F.linear(h, sd[p + "fc_encoder.0.weight"], sd[p + "fc_encoder.0.bias"])).permute(0, 2, 1)
    x = F.elu(F.conv1d(x, sd[p + "conv_layers.0.weight"], sd[p + "conv_layers.0.bias"], stride=2))
    x = F.elu(F.conv1d(x, sd[p + "conv_layers.2.weight"], sd[p + "conv_layers.2.bias"], stride=1))
    return F.elu(F.linear(x.flatten(1), sd[p + "fc_final.0.weight"], sd[p + "fc_final.0.bias"]))


def privileged_encoder(sd, priv):
    pre = "privileged_encoder_.priv_encoder"
    return mlp(sd, pre, priv, n_layers(sd, pre))


def scan_encoder(sd, scan):
    pre = "scan_encoder.scan_encoder"
    return mlp(sd, pre, scan, n_layers(sd, pre))


def actor_mean(sd, obs, priv, est, scan, adaptation_mode=False):
    """ActorCritic.update_distribution (actor_critic.py:190-197)."""
    latent = adaptation_encoder(sd, obs) if adaptation_mode else privileged_encoder(sd, priv)
    x = torch.cat((obs, latent, scan_encoder(sd, scan), est), dim=-1)
    return mlp(sd, "actor", x, n_layers(sd, "actor"))


def critic_value(sd, critic_obs):
    return mlp(sd, "critic", critic_obs, n_layers(sd, "critic"))


def estimator(sd_est, obs):
    return mlp(sd_est, "estimator", obs, n_layers(sd_est, "estimator"))


def normal_log_prob(actions, mu, std):
    return (-((actions - mu) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)


def keyed_standard_normal(seed, step, num_envs, num_actions):
    """Box-Muller on keyed uniforms -- twin of sample_actions_kernel (csrc/learner_kernels.cu)."""
    lanes = np.arange(2 * num_actions)
    r = philox.keyed_u32(seed, philox.SITE_ACTION_NOISE, step, np.arange(num_envs), lanes)
    u1 = (((r[:, 0::2] >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)).astype(np.float32)
    u2 = ((r[:, 1::2] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    z = torch.sqrt(-2.0 * torch.log(torch.from_numpy(u1))) * torch.cos(torch.from_numpy(u2) * 6.283185307179586)
    return z


def ppo_act(sd, sd_est, obs, priv, critic_obs, scan, seed, step, adaptation_mode=False):
    """PPO.act (ppo.py:129-153): acts on the ESTIMATED obs; -> actions, values, log-prob, mu, sigma."""
    with torch.no_grad():
        est_hat = estimator(sd_est, obs)
        mu = actor_mean(sd, obs, priv, est_hat, scan, adaptation_mode)
        std = sd["std"]
        actions = mu + std * keyed_standard_normal(seed, step, obs.shape[0], mu.shape[1])
        values = critic_value(sd, critic_obs)
        return actions, values, normal_log_prob(actions, mu, mu * 0. + std), mu, (mu * 0. + std)


def ppo_losses(sd, sd_est, b, clip=0.2, value_coef=1.0, entropy_coef=0.01, reg_coef=0.0, use_clipped_value_loss=True):
    """The loss block of PPO.update for one minibatch `b` (dict of tensors), ppo.py:199-270.
    -> dict(loss, surrogate, value, reg, entropy, estimator)"""
    mu = actor_mean(sd, b["obs"], b["priv"], b["true_est"], b["scan"], False)
    sigma = mu * 0. + sd["std"]
    logp = normal_log_prob(b["actions"], mu, sigma)
    value = critic_value(sd, b["critic_obs"])
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(dim=-1)
    lat_p = privileged_encoder(sd, b["priv"])
    with torch.no_grad():
        lat_a = adaptation_encoder(sd, b["obs"])
    reg = (lat_p - lat_a.detach()).norm(p=2, dim=1).mean()
    est_loss = (estimator(sd_est, b["obs"]) - b["true_est"]).norm(p=2, dim=1).pow(2).mean()
    ratio = torch.exp(logp - torch.squeeze(b["old_logp"]))
    adv = torch.squeeze(b["adv"])
    surrogate = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
    if use_clipped_value_loss:
        vc = b["values"] + (value - b["values"]).clamp(-clip, clip)
        vloss = torch.max((value - b["returns"]).pow(2), (vc - b["returns"]).pow(2)).mean()
    else:
        vloss = (b["returns"] - value).pow(2).mean()
    loss = surrogate + value_coef * vloss - entropy_coef * entropy.mean() + reg_coef * reg
    return dict(loss=loss, surrogate=surrogate, value=vloss, reg=reg, entropy=entropy.mean(), estimator=est_loss,
                mu=mu, value_out=value, lat_p=lat_p, lat_a=lat_a)


def adaptive_lr(lr, mu, sigma, old_mu, old_sigma, desired_kl):
    """The schedule == 'adaptive' block of PPO.update (ppo.py:233-246) -> (new learning rate, kl_mean)."""
    with torch.no_grad():
        kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) + (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma))
                       - 0.5, axis=-1)
        kl_mean = torch.mean(kl)
        if kl_mean > desired_kl * 2.0:
            lr = max(1e-5, lr / 1.5)
        elif kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
            lr = min(1e-2, lr * 1.5)
    return lr, float(kl_mean)


def dagger_loss(sd, b):
    """PPO.update_dagger's loss (ppo.py:322-333)."""
    with torch.no_grad():
        lat_p = privileged_encoder(sd, b["priv"])
    lat_a = adaptation_encoder(sd, b["obs"])
    return (lat_p.detach() - lat_a).norm(p=2, dim=1).mean()


def clip_and_adam(params, grads, state, lr, max_norm, betas=(0.9, 0.999), eps=1e-8):
    """nn.utils.clip_grad_norm_ + torch.optim.Adam.step on lists of tensors; `state` = dict(step, m, v)."""
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    state["step"] += 1
    t = state["step"]
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g.detach() * coef
        state["m"][i].lerp_(g, 1 - betas[0])
        state["v"][i].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = (state["v"][i].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(state["m"][i], denom, value=-lr / bc1)


MAIN_PREFIXES = ("actor.", "critic.", "privileged_encoder_.", "std", "scan_encoder.")


class LearnerOracle:
    """PPO.update / update_dagger over explicit minibatch index lists (the reference draws them with randperm)."""

    def __init__(self, sd, sd_est, lr=2e-4, est_lr=1e-4, max_grad_norm=1.0, desired_kl=None, **loss_kw):
        self.sd = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()}
        self.sd_est = {k: v.clone().float().requires_grad_(True) for k, v in sd_est.items()}
        self.lr, self.est_lr, self.max_grad_norm, self.loss_kw = lr, est_lr, max_grad_norm, loss_kw
        self.desired_kl, self.kl_log = desired_kl, []      # desired_kl set = schedule 'adaptive'
        self.main_keys = [k for k in self.sd if k.startswith(MAIN_PREFIXES)]
        self.adapt_keys = [k for k in self.sd if k.startswith("adaptation_encoder_.")]
        mk = lambda keys, d: dict(step=0, m=[torch.zeros_like(d[k]) for k in keys], v=[torch.zeros_like(d[k]) for k in keys])
        self.opt_main, self.opt_adapt = mk(self.main_keys, self.sd), mk(self.adapt_keys, self.sd)
        self.est_keys = list(self.sd_est)
        self.opt_est = mk(self.est_keys, self.sd_est)

    def minibatch(self, b, reg_coef=0.0):
        out = ppo_losses(self.sd, self.sd_est, b, reg_coef=reg_coef, **self.loss_kw)
        g_est = torch.autograd.grad(out["estimator"], [self.sd_est[k] for k in self.est_keys])
        clip_and_adam([self.sd_est[k] for k in self.est_keys], g_est, self.opt_est, self.est_lr, self.max_grad_norm)
        g = torch.autograd.grad(out["loss"], [self.sd[k] for k in self.main_keys], allow_unused=True)
        g = [torch.zeros_like(self.sd[k]) if gi is None else gi for k, gi in zip(self.main_keys, g)]
        self.last_grads = dict(zip(self.main_keys, g), **dict(zip(self.est_keys, g_est)))
        if self.desired_kl is not None:
            mu = out["mu"].detach()
            self.lr, kl = adaptive_lr(self.lr, mu, mu * 0. + self.sd["std"].detach(), b["mu"], b["sigma"], self.desired_kl)
            self.kl_log.append((kl, self.lr))
        clip_and_adam([self.sd[k] for k in self.main_keys], g, self.opt_main, self.lr, self.max_grad_norm)
        return {k: float(v.detach()) for k, v in out.items() if v.dim() == 0}

    def dagger_minibatch(self, b):
        loss = dagger_loss(self.sd, b)
        g = torch.autograd.grad(loss, [self.sd[k] for k in self.adapt_keys])
        self.last_grads = dict(zip(self.adapt_keys, g))
        clip_and_adam([self.sd[k] for k in self.adapt_keys], g, self.opt_adapt, self.lr, self.max_grad_norm)
        return float(loss)
